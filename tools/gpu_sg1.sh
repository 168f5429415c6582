#!/bin/bash
# stream GEMM bring-up: parity tests (bounded), micro-benchmark
TAG=${1:-sg1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stream_gemm" > $OUT/pytest_sg.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_sg.log
timeout 600 python tools/microbench.py stream_gemm > $OUT/sgemm.jsonl 2> $OUT/sgemm.err; echo "micro rc=$?"; cat $OUT/sgemm.jsonl; tail -5 $OUT/sgemm.err
