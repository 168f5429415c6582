#!/bin/bash
# Last visit of the round: full parity suite, smoke, the driver's bench line, one ncu --set full capture of the
# uncertainty kernel after the late changes (grid split, epilogue).
OUT=gpurun_out/${1:-r01x}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-900 $OUT/bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uncertainty_kernel -c 2 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
ncu -i $OUT/unc.ncu-rep --page raw --csv > $OUT/unc_raw.csv 2>/dev/null
rm -f $OUT/unc.ncu-rep
ls -la $OUT
