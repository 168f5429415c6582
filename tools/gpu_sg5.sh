#!/bin/bash
OUT=gpurun_out/${1:-sg5}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stream_gemm" > $OUT/pytest_sg.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_sg.log
timeout 300 python tools/sg_ts.py > $OUT/ts.log 2>&1; cat $OUT/ts.log | tail -40
timeout 600 python tools/microbench.py stream_gemm > $OUT/sgemm.jsonl 2> $OUT/sgemm.err; python - <<'PY'
import json,sys
for l in open('$OUT/sgemm.jsonl'):
    d=json.loads(l); print(d['M'],d['N'],d['K'],d['epi'],'us',d['us'],'cublas',d['cublas_us'],'skinny',d.get('skinny_us'),'frac',d['frac_hbm'])
PY; tail -5 $OUT/sgemm.err
