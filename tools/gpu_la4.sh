#!/bin/bash
OUT=gpurun_out/${1:-la4}; mkdir -p $OUT
for b in 3 4 32; do for la in 1 0; do DD_BATCH_LA=$la timeout 600 python bench.py --batch $b --new-tokens 128 --steps 2 --no-cpu-baseline > $OUT/bench_b${b}_la$la.json 2> $OUT/bench_b${b}_la$la.err; python - <<PY
import json
d=json.load(open("$OUT/bench_b${b}_la$la.json"))
print("B=$b la=$la", d["value"], d["e2e"]["value"], d["ms_per_token"], d["decode_schedule"]["steps"], d["decode_schedule"]["fallback_rate"])
PY
done; done
for b in 2 16; do for la in 1 0; do DD_BATCH_LA=$la timeout 300 python tools/stepbench.py --model next --batch $b --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
cat $OUT/stepbench.jsonl; tail -5 $OUT/stepbench.err
