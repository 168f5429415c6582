#!/bin/bash
# Uncertainty-kernel experiment variants (build.py --variant=...), one timing line each.
TAG=${1:-exp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "" wpf noepi wpf_noepi wpf_krot wpf_krot_noepi; do
  if [ -z "$v" ]; then unset DD_LIB; else export DD_LIB=$PWD/dropoutdecoding_b200/_build/libdropdec_$v.so; fi
  timeout 300 python tools/microbench.py uncertainty_first >> $OUT/unc_variants.jsonl 2>> $OUT/unc_variants.err
done
cat $OUT/unc_variants.jsonl
export DD_LIB=$PWD/dropoutdecoding_b200/_build/libdropdec_wpf_krot.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k uncertainty > $OUT/pytest_unc_wpf_krot.log 2>&1; tail -3 $OUT/pytest_unc_wpf_krot.log
unset DD_LIB
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -15 $OUT/pytest_gpu.log
