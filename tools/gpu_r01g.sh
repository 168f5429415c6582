#!/bin/bash
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "uncertainty" > $OUT/pytest_unc.log 2>&1; echo "pytest unc rc=$?"; tail -5 $OUT/pytest_unc.log
timeout 300 python tools/microbench.py uncertainty_first > $OUT/unc.jsonl 2> $OUT/unc.err
DD_LIB=$PWD/dropoutdecoding_b200/_build/libdropdec_noepi.so timeout 300 python tools/microbench.py uncertainty_first >> $OUT/unc.jsonl 2>> $OUT/unc.err
cat $OUT/unc.jsonl; tail -3 $OUT/unc.err
for b in 1 8; do timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done
DD_FUSE_NORM=1 timeout 300 python tools/stepbench.py --batch 1 --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err
cat $OUT/stepbench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uncertainty_kernel -c 2 -o $OUT/unc -f python tools/microbench.py uncertainty1 > $OUT/ncu_unc.log 2>&1; echo "ncu unc rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_|skinny_gemm|aggregate_kernel|commit_winner|embed_gather|rmsnorm|rope_stage|sampler_kernel|step_advance" -c 1400 --csv --log-file $OUT/launches.csv python tools/stepbench.py --batch 1 --steps 2 --reps 1 --no-graph > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
