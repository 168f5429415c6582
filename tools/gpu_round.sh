#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), micro-benchmarks, step timings, ncu launch list and
# full captures of the north-star kernels.  Outputs under gpurun_out/<tag>/.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
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
timeout 900 python tools/microbench.py all > $OUT/microbench.jsonl 2> $OUT/microbench.err; echo "micro rc=$?"; cat $OUT/microbench.jsonl
for b in 1 8 16 64; do for sp in auto off; do DD_SPEC=$sp timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
DD_GEMM=stream timeout 300 python tools/stepbench.py --batch 64 --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
timeout 300 python tools/stepbench.py --model next --batch 16 --probs 0.1,0.2,0.3,0.4,0.5,0.6,0.7,0.8 --mode independent --seq 2900 --n-vis 2880 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 1400 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 4 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_gemm -s 2 -c 6 -o $OUT/sgemm -f python tools/microbench.py sgemm1 > $OUT/ncu_sgemm.log 2>&1; echo "ncu sgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uncertainty_kernel -c 2 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_partial -s 1 -c 1 -o $OUT/attn -f python tools/microbench.py attention64 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 1 -c 7 -o $OUT/gemm -f python tools/microbench.py gemm1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
for r in sgemm unc attn gemm; do ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/${r}_raw.csv 2>/dev/null; done
ls -la $OUT | head -40
