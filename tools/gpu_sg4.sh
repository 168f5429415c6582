#!/bin/bash
OUT=gpurun_out/${1:-sgts}; mkdir -p $OUT
timeout 300 python tools/sg_ts.py > $OUT/ts.log 2>&1; cat $OUT/ts.log | tail -40
for d in 0 1 2 4 8 15; do DD_SG_QUICK=1 DD_SG_DBG=$d timeout 300 python tools/microbench.py stream_gemm >> $OUT/sgemm.jsonl 2>> $OUT/sgemm.err; done
cut -c1-140 $OUT/sgemm.jsonl; tail -5 $OUT/sgemm.err
