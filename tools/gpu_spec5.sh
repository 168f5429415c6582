#!/bin/bash
OUT=gpurun_out/${1:-spec5}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log
for b in 16 64; do for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
DD_GEMM=stream timeout 300 python tools/stepbench.py --batch 64 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl; tail -3 $OUT/stepbench.err
