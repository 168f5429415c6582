#!/bin/bash
OUT=gpurun_out/${1:-spec2}; mkdir -p $OUT
timeout 600 python tools/sg_stress.py > $OUT/stress.log 2>&1; tail -20 $OUT/stress.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "eager_tcgen05_gemm or bit_for_bit" 2>&1 | tail -4; done
