#!/bin/bash
# ncu captures of the hot kernels (one GPU).  Usage: gpurun --timeout 1800 -- 'bash tools/gpu_prof.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KREG='regex:attn_|skinny_gemm|aggregate_kernel|commit_winner|embed_gather|rmsnorm|rope_stage|sampler_kernel|step_advance|uncertainty_kernel|finalize_kernel|logpavg'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 1400 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 2 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uncertainty_kernel -c 2 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_partial -s 1 -c 1 -o $OUT/attn -f python tools/microbench.py attention64 > $OUT/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 1 -c 7 -o $OUT/gemm -f python tools/microbench.py gemm1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la $OUT
