#!/bin/bash
TAG=${1:-sgdbg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for d in 0 1 2 4 6 7 8 15; do DD_SG_QUICK=1 DD_SG_DBG=$d timeout 300 python tools/microbench.py stream_gemm >> $OUT/sgemm.jsonl 2>> $OUT/sgemm.err; done
cat $OUT/sgemm.jsonl; tail -5 $OUT/sgemm.err
