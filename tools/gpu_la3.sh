#!/bin/bash
OUT=gpurun_out/${1:-la3}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED\|Error" $OUT/pytest_gpu.log | head -30
for b in 2 8 16; do for la in 1 0; do DD_BATCH_LA=$la timeout 600 python bench.py --batch $b --new-tokens 128 --steps 2 --no-cpu-baseline > $OUT/bench_b${b}_la$la.json 2> $OUT/bench_b${b}_la$la.err; python - <<PY
import json
d=json.load(open("$OUT/bench_b${b}_la$la.json"))
print("B=$b la=$la", d["value"], d["e2e"]["value"], d["ms_per_token"], d["decode_schedule"]["steps"], d["decode_schedule"]["fallback_rate"])
PY
done; done
