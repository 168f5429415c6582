#!/bin/bash
mkdir -p gpurun_out/vis
timeout 900 python -m pytest tests/test_wrappers_gpu.py -q -x -m gpu 2>&1 | tail -5
DD_VISION_GRAPH=0 timeout 600 python tools/e2e_profile.py 2>&1 | tail -2 | tee gpurun_out/vis/e2e_eager.jsonl
DD_VISION_GRAPH=1 timeout 600 python tools/e2e_profile.py 2>&1 | tail -2 | tee gpurun_out/vis/e2e_graph.jsonl
