#!/bin/bash
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "uncertainty" > $OUT/pytest_unc.log 2>&1; echo "pytest unc rc=$?"; tail -5 $OUT/pytest_unc.log
timeout 300 python tools/microbench.py uncertainty > $OUT/unc.jsonl 2> $OUT/unc.err; cat $OUT/unc.jsonl; tail -3 $OUT/unc.err
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
for b in 1 8; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; cat $OUT/stepbench.jsonl
