#!/bin/bash
OUT=gpurun_out/${1:-sg8}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
for b in 1 8 64; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
DD_GEMM=stream timeout 300 python tools/stepbench.py --batch 64 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
DD_GEMM=skinny timeout 300 python tools/stepbench.py --batch 8 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
DD_GEMM=library timeout 300 python tools/stepbench.py --batch 8 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl; tail -3 $OUT/stepbench.err
