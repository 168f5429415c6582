#!/bin/bash
# End-of-round visit without the ncu --set full captures (kernels unchanged since tools/gpu_round.sh ran):
# parity tests, smoke, bench (both arms, two-pass, lookahead off), step timings, ncu launch list.
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KREG='regex:attn_|skinny_gemm|stream_gemm|aggregate_kernel|commit_winner|embed_gather|rmsnorm|rope_stage|sampler_kernel|step_advance|token_stats|uncertainty_kernel|finalize_kernel|logpavg'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
DD_SPEC=off timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_two_pass.json 2> $OUT/bench_two_pass.err; echo "bench two-pass rc=$?"
DD_LOOKAHEAD=0 timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_no_lookahead.json 2> $OUT/bench_no_lookahead.err; echo "bench no-lookahead rc=$?"
for b in 1 8 16 64; do for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
timeout 300 python tools/stepbench.py --model next --batch 16 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
timeout 300 python tools/stepbench.py --model next --batch 2 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --topk 10 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
timeout 300 python tools/stepbench.py --model iblip --batch 32 --mode quantile --vote-on hidden --topk 10 --seq 52 --n-vis 32 --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 1600 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 4 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ls -la $OUT | head -40
