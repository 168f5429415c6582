#!/bin/bash
mkdir -p gpurun_out/agg
timeout 2400 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/agg/pytest_gpu.log
