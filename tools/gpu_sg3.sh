#!/bin/bash
TAG=${1:-sgncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_gemm -s 2 -c 2 -o $OUT/sg -f python tools/microbench.py sgemm1 > $OUT/ncu_sg.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_sg.log
ncu -i $OUT/sg.ncu-rep --page raw --csv > $OUT/sg_raw.csv 2>/dev/null
ncu -i $OUT/sg.ncu-rep --page details --csv > $OUT/sg_details.csv 2>/dev/null
ls -la $OUT
