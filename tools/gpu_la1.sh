#!/bin/bash
OUT=gpurun_out/${1:-la1}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|Error" $OUT/pytest_gpu.log | head -30
for la in 1 0; do DD_LOOKAHEAD=$la timeout 300 python tools/stepbench.py --batch 1 --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
cat $OUT/stepbench.jsonl; tail -5 $OUT/stepbench.err
for la in 1 0; do DD_LOOKAHEAD=$la timeout 900 python bench.py > $OUT/bench_la$la.json 2> $OUT/bench_la$la.err; echo "bench rc=$?"; cat $OUT/bench_la$la.json; tail -3 $OUT/bench_la$la.err; done
