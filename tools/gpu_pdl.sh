#!/bin/bash
# dependent-launch modes of the <= 16-row passes: ms per token
mkdir -p gpurun_out/pdl
: > gpurun_out/pdl/step.jsonl
for v in "DD_PDL_SKINNY=0" "DD_PDL_SKINNY=2" "DD_PDL_SKINNY=1" "DD_PDL_SKINNY=2 DD_LIB=dropoutdecoding_b200/_build/libdropdec_trig1.so" "DD_PDL_SKINNY=1 DD_LIB=dropoutdecoding_b200/_build/libdropdec_trig1.so"; do
  echo "$v" | tee -a gpurun_out/pdl/step.jsonl
  env $v timeout 300 python tools/stepbench.py --batch 1 2>&1 | tail -1 | tee -a gpurun_out/pdl/step.jsonl
done
env DD_PDL_SKINNY=2 timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -m gpu -k "bit_exact or bit_for_bit or two_pass" 2>&1 | tail -3
