#!/bin/bash
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "attention" > $OUT/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -5 $OUT/pytest_attn.log
timeout 300 python tools/microbench.py attention > $OUT/attn.jsonl 2> $OUT/attn.err; cat $OUT/attn.jsonl; tail -3 $OUT/attn.err
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
for b in 1 8 64; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; cat $OUT/stepbench.jsonl
timeout 300 python tools/microbench.py uncertainty_first > $OUT/unc.jsonl 2> $OUT/unc.err; cat $OUT/unc.jsonl
