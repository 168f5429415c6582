#!/usr/bin/env python
"""Experiment: where the uncertainty kernel's cycles go (needs the experiment build:
python -m dropoutdecoding_b200.build --variant=exp -DDD_SG_EXPERIMENT -DDD_UNC_EXPERIMENT;
DD_LIB=dropoutdecoding_b200/_build/libdropdec_exp.so DD_UNC_DBG=1 python tools/unc_cycles.py [n_images])."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dropoutdecoding_b200 import _lib, ops  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(n_img, 576, 4096, generator=g, device=dev).bfloat16()
w = (torch.randn(32064, 4096, generator=g, device=dev) * 0.02).bfloat16()
L = _lib.load()
L.dd_uncertainty_debug_read.restype = C.c_int
L.dd_uncertainty_debug_read.argtypes = [C.c_void_p, C.c_int]
for _ in range(3):
    ops.uncertainty_project(h, w, topk=5, round_logits=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.uncertainty_project(h, w, topk=5, round_logits=True, check=False); e1.record()
torch.cuda.synchronize()
n = 148
buf = np.zeros((n, 8), dtype=np.int64)
rc = L.dd_uncertainty_debug_read(buf.ctypes.data, n)
names = ["mma_loop", "mma_wait_w", "mma_wait_act", "mma_wait_tmem_empty", "epi_loop", "epi_wait_tmem_full"]
print(f"n_img={n_img}: {e0.elapsed_time(e1) * 1e3:.1f} us per call; cycle counters of the LAST sweep (FUSED), median over CTAs (rc={rc})")
live = buf[buf[:, 0] > 0]
for i, nm in enumerate(names):
    print(f"  {nm:22s} {int(np.median(live[:, i])):>9d}   (min {int(live[:, i].min())}, max {int(live[:, i].max())})")
print(f"  ctas with data: {len(live)}")
