#!/bin/bash
# Round-2 end-of-round visit (one GPU): parity tests, smoke, bench (both arms), kernel micro-benchmarks, step timings.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_r02_final.sh [tag]'; the ncu visit is tools/gpu_r02_ncu.sh
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench rc=$?"; cut -c1-700 $OUT/bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-500 $OUT/bench_ref.json
timeout 300 python tools/microbench.py uncertainty > $OUT/uncertainty_onesweep.jsonl 2> $OUT/microbench.err; echo "microbench unc rc=$?"; cut -c1-200 $OUT/uncertainty_onesweep.jsonl
timeout 300 python tools/microbench.py attention > $OUT/attention_microbench.jsonl 2>> $OUT/microbench.err; echo "microbench attn rc=$?"
timeout 300 ./tools/_bin/sgemm_bench 4 16 24 32 64 128 192 256 > $OUT/stream_gemm_v5.jsonl 2> $OUT/sgemm.err; echo "sgemm rc=$?"
DD_PDL=1 timeout 300 ./tools/_bin/sgemm_bench 32 64 128 192 256 > $OUT/stream_gemm_v5_chain_launch.jsonl 2>> $OUT/sgemm.err; echo "sgemm (dependent launch) rc=$?"
for b in 1 2 4 8 16 32 64; do timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
for b in 8 64; do DD_PDL=0 timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; DD_PAIR_GEMM=0 timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; DD_GEMM=library timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
timeout 300 python tools/stepbench.py --model next --batch 16 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
timeout 300 python tools/stepbench.py --model iblip --batch 32 --mode quantile --vote-on hidden --topk 10 --seq 52 --n-vis 32 --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cut -c1-230 $OUT/stepbench.jsonl
ls -la $OUT | head -40
