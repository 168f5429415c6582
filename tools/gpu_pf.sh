#!/bin/bash
# L2-prefetch sweep (round 2): ms per token of a one-sequence pass for several DD_PF_MB settings.
mkdir -p gpurun_out/pf
python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "l2_prefetch or rmsnorm or decode_attention" > gpurun_out/pf/pytest.log 2>&1
tail -3 gpurun_out/pf/pytest.log
: > gpurun_out/pf/sweep.jsonl
for pf in ${PF_LIST:-0 0,48,0 0,24,0 0,0.01,0 0,80,0 24,48,24}; do
  echo "pf=$pf" | tee -a gpurun_out/pf/sweep.jsonl
  DD_PF_MB=$pf python tools/stepbench.py --batch 1 2>&1 | tail -1 | tee -a gpurun_out/pf/sweep.jsonl
done
