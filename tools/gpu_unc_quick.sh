#!/bin/bash
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "uncertainty" > $OUT/pytest_unc.log 2>&1; echo "pytest unc rc=$?"; tail -5 $OUT/pytest_unc.log
timeout 300 python tools/microbench.py uncertainty_first > $OUT/unc.jsonl 2> $OUT/unc.err
DD_LIB=$PWD/dropoutdecoding_b200/_build/libdropdec_noepi.so timeout 300 python tools/microbench.py uncertainty_first >> $OUT/unc.jsonl 2>> $OUT/unc.err
cat $OUT/unc.jsonl; tail -3 $OUT/unc.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:uncertainty_kernel -c 4 --csv --log-file $OUT/unc_launches.csv python tools/microbench.py uncertainty1 > $OUT/ncu.log 2>&1; grep uncertainty_kernel $OUT/unc_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-150
