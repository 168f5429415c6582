#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ./dropoutdecoding_b200/_build/umma_probe > $OUT/umma_probe.jsonl 2> $OUT/umma_probe.err; echo "probe rc=$?"; cat $OUT/umma_probe.jsonl; tail -3 $OUT/umma_probe.err
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for b in 1 8; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; cat $OUT/stepbench.jsonl
