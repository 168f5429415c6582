#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), micro-benchmarks, ncu launch list.
# Outputs under gpurun_out/.  Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
timeout 600 python tools/microbench.py all > $OUT/microbench.jsonl 2> $OUT/microbench.err; echo "micro rc=$?"; cat $OUT/microbench.jsonl
for b in 1 8 64; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; cat $OUT/stepbench.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dd -c 1400 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 2 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/pytest_gpu.log $OUT/smoke.log
