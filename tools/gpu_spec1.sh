#!/bin/bash
OUT=gpurun_out/${1:-spec1}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
for b in 1 8; do for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --batch $b --steps 48 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
cat $OUT/stepbench.jsonl; tail -3 $OUT/stepbench.err
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
