#!/bin/bash
# Round-2 late visit (one GPU): parity tests, smoke, bench (both arms), step timings, the persistent pass kernel's
# timeline, ncu launch list of the bench command and full captures of the kernels changed late in the round.
# Usage (from the repo root): gpurun --timeout 2700 -- 'bash tools/gpu_r02_late.sh [tag]'
TAG=${1:-r02late}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench rc=$?"; cut -c1-900 $OUT/bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $OUT/bench_ref.json
for b in 1 2 4 8 16 32 64; do timeout 300 python tools/stepbench.py --batch $b --steps 63 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
DD_PASS=1 timeout 300 python tools/stepbench.py --batch 1 --steps 63 >> $OUT/stepbench_pass_kernel.jsonl 2>> $OUT/stepbench.err
DD_PASS=1 timeout 300 python tools/pass_timeline.py > $OUT/pass_timeline.json 2>> $OUT/stepbench.err
cut -c1-200 $OUT/stepbench.jsonl $OUT/stepbench_pass_kernel.jsonl $OUT/pass_timeline.json
timeout 300 python tools/e2e_profile.py 2>/dev/null | tail -2 > $OUT/e2e_profile.jsonl; cat $OUT/e2e_profile.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-target --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches (bench) rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv --from-first dd:: > $OUT/launches_bench_b1.md 2>/dev/null; head -14 $OUT/launches_bench_b1.md
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:aggregate_kernel|token_stats_kernel|rmsnorm_kernel" -s 20 -c 6 -o $OUT/agg -f python tools/stepbench.py --batch 1 --steps 3 --reps 1 --no-graph > $OUT/ncu_agg.log 2>&1; echo "ncu agg rc=$?"
DD_PASS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:decoder_pass_kernel" -s 2 -c 1 -o $OUT/pass -f python tools/stepbench.py --batch 1 --steps 3 --reps 1 --no-graph > $OUT/ncu_pass.log 2>&1; echo "ncu pass rc=$?"
for r in agg pass; do ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/${r}_raw.csv 2>/dev/null; done
rm -f $OUT/*.ncu-rep
ls -la $OUT | head -40
