#!/bin/bash
OUT=gpurun_out/${1:-la2}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|Error" $OUT/pytest_gpu.log | head -30
for b in 2 8 16; do for la in 1 0; do DD_BATCH_LA=$la timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
cat $OUT/stepbench.jsonl; tail -5 $OUT/stepbench.err
