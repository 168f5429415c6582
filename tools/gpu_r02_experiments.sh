#!/bin/bash
# Round-2 late experiments on the one-sequence pass (one GPU); each section is what produced a file under profiles/r02/.
# Usage (from the repo root): gpurun --timeout 1800 -- 'bash tools/gpu_r02_experiments.sh [section ...]'
OUT=gpurun_out/exp; mkdir -p $OUT
want() { [ $# -eq 0 ] && return 0; for s in "${SECTIONS[@]}"; do [ "$s" = "$1" ] && return 0; done; [ ${#SECTIONS[@]} -eq 0 ]; }
SECTIONS=("$@")
step() { python tools/stepbench.py --batch 1 2>&1 | tail -1; }
if want prefetch; then   # l2_prefetch_sweep.jsonl: dd_hint_l2_prefetch, MB hinted to {RMSNorm before QKV, attention, RMSNorm before gate/up}
  for pf in 0 0,48,0 0,24,0 0,0.01,0 0,80,0 24,48,24; do echo "pf=$pf"; DD_PF_MB=$pf step; done | tee $OUT/l2_prefetch_sweep.jsonl
  python tools/pf_probe.py | tee $OUT/pf_probe.jsonl
fi
if want pdl; then        # pdl_modes.jsonl: dependent-launch modes of the <= 16-row passes
  for m in 0 1 2 3 0 3; do echo "DD_PDL_SKINNY=$m"; DD_PDL_SKINNY=$m step; DD_PDL_SKINNY=$m python tools/pass_forms.py | tail -1; done | tee $OUT/pdl_modes.jsonl
fi
if want splits; then     # attention_splits_sweep.txt
  for sp in 2 3 4 5 6 8; do echo "splits=$sp"; DD_ATTN_SPLITS=$sp step; DD_ATTN_SPLITS=$sp python tools/pass_forms.py | tail -1; done | tee $OUT/attention_splits.jsonl
  for v in 1 0; do echo "DD_TWO_WINDOW_ATTN=$v"; DD_TWO_WINDOW_ATTN=$v python tools/pass_forms.py | tail -1; done | tee -a $OUT/attention_splits.jsonl
fi
if want pass; then       # pass_timeline.json, stepbench_pass_kernel.jsonl: the persistent pass kernel
  python -m pytest tests/test_engine_gpu.py -q -x -m gpu -k persistent_pass | tail -2
  DD_PASS=1 python tools/pass_timeline.py | tee $OUT/pass_timeline.json
  DD_PASS=1 step | tee $OUT/stepbench_pass_kernel.jsonl
fi
