#!/bin/bash
# Round-2 ncu visit: launch list of the bench command + full captures of the three north-star kernels.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_r02_ncu.sh [tag]'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
# launch list of the SAME command the driver benches (one timed step + one warm-up; the graph's kernel nodes are listed one by one)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-target --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches (bench) rc=$?"
# a batch pass (8 sequences, lookahead: 56 rows -> dd_stream_gemm)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_b8.csv python tools/stepbench.py --batch 8 --steps 2 --reps 1 --no-graph > $OUT/ncu_launch_b8.log 2>&1; echo "ncu launches (b8) rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:stream_gemm|pair_gemm" -s 2 -c 6 -o $OUT/sgemm -f python tools/microbench.py sgemm1 > $OUT/ncu_sgemm.log 2>&1; echo "ncu sgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:logits_stats|entropy_kernel|finalize" -s 4 -c 4 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_partial -s 1 -c 1 -o $OUT/attn -f python tools/microbench.py attention64 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 300 python tools/microbench.py uncertainty > $OUT/uncertainty_microbench.jsonl 2>&1; echo "microbench unc rc=$?"
for r in sgemm unc attn; do ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/${r}_raw.csv 2>/dev/null; done
ls -la $OUT | head -30
