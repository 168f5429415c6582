#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), micro-benchmarks, step timings, ncu launch list and
# full captures of the three north-star kernels.  Outputs under gpurun_out/<tag>/.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KREG='regex:attn_|skinny_gemm|aggregate_kernel|commit_winner|embed_gather|rmsnorm|rope_stage|sampler_kernel|step_advance|uncertainty_kernel|finalize_kernel|logpavg'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
timeout 600 python tools/microbench.py all > $OUT/microbench.jsonl 2> $OUT/microbench.err; echo "micro rc=$?"; cat $OUT/microbench.jsonl
for b in 1 8 64; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
DD_PDL=1 timeout 300 python tools/stepbench.py --batch 1 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 1100 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 2 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uncertainty_kernel -c 2 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_partial -s 1 -c 1 -o $OUT/attn -f python tools/microbench.py attention64 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 1 -c 7 -o $OUT/gemm -f python tools/microbench.py gemm1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la $OUT | head -30
