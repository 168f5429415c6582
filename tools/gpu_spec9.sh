#!/bin/bash
OUT=gpurun_out/${1:-spec9}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "^E  \|passed\|failed\|FAILED" $OUT/pytest_gpu.log | head -20
