#!/bin/bash
OUT=gpurun_out/${1:-spec10}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $OUT/pytest_gpu.log | head -20
for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --model next --batch 16 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --model next --batch 2 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --model iblip --batch 32 --mode quantile --vote-on hidden --topk 10 --seq 52 --n-vis 32 --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
cat $OUT/stepbench.jsonl; tail -5 $OUT/stepbench.err
