#!/bin/bash
OUT=gpurun_out/${1:-spec8}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed" $OUT/pytest_gpu.log | head -20
for b in 2 4 8; do timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
timeout 300 python tools/stepbench.py --batch 64 --steps 127 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
timeout 300 python tools/stepbench.py --model next --batch 16 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl; tail -3 $OUT/stepbench.err
timeout 900 python bench.py --batch 8 --no-cpu-baseline > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/bench_b8.json"))
print({k:d[k] for k in ("value","e2e","ms_per_step","decode_schedule","gpu_launches")})
PY
