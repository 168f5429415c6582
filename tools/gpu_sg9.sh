#!/bin/bash
OUT=gpurun_out/${1:-sg9}; mkdir -p $OUT
timeout 600 python tools/microbench.py attention_sweep > $OUT/attn_sweep.jsonl 2> $OUT/attn_sweep.err; cat $OUT/attn_sweep.jsonl; tail -3 $OUT/attn_sweep.err
for b in 16 21; do for g in auto stream; do DD_GEMM=$g timeout 300 python tools/stepbench.py --batch $b --steps 32 >> $OUT/stepbench.jsonl 2>> $OUT/stepbench.err; done; done
cat $OUT/stepbench.jsonl; tail -3 $OUT/stepbench.err
