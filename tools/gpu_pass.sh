#!/bin/bash
# persistent pass kernel: parity tests, phase timeline, ms per token with / without it
mkdir -p gpurun_out/pass
timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -m gpu -k "persistent_pass" > gpurun_out/pass/pytest_pass.log 2>&1
tail -5 gpurun_out/pass/pytest_pass.log
DD_PASS=1 timeout 300 python tools/pass_timeline.py 2>&1 | tail -1 | tee gpurun_out/pass/timeline.json
: > gpurun_out/pass/step.jsonl
for v in "DD_PASS=0" "DD_PASS=1"; do
  echo "$v" | tee -a gpurun_out/pass/step.jsonl
  env $v timeout 300 python tools/stepbench.py --batch 1 2>&1 | tail -1 | tee -a gpurun_out/pass/step.jsonl
done
