#!/usr/bin/env python
"""Experiment: per-CTA phase timestamps of dd_stream_gemm (DD_SG_DBG=16)."""
import math, os, sys
os.environ["DD_SG_DBG"] = str(16 | int(os.environ.get("DD_SG_DBG", "0")))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200 import ops
dev = torch.device("cuda:0")
for (M, N, K) in [(192, 4096, 4096), (192, 12288, 4096), (24, 4096, 4096)]:
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    ws = ops.stream_gemm_workspace(M, [(N, K, 0)], dev)
    for _ in range(3):
        ops.stream_gemm(x, w, workspace=ws)
    torch.cuda.synchronize()
    G = 148
    ts = ws[-G * 64:].view(torch.int64).view(G, 8).cpu()
    t0 = ts[:, 0].min()
    rel = (ts - t0).float() / 1e3
    rel[ts == 0] = float("nan")
    names = ["entry", "prologue", "last_tfull", "segs_done", "wait0_done", "wait1_done", "epi_done", "all_done"]
    print(f"M={M} N={N} K={K}  (us since first CTA entry: min / median / max)")
    for i, n in enumerate(names):
        col = rel[:, i]; col = col[~col.isnan()]
        if len(col):
            print(f"  {n:12s} {col.min():8.1f} {col.median():8.1f} {col.max():8.1f}   n={len(col)}")
    worst = rel[:, 7].argmax().item()
    print("  slowest CTA", worst, [round(v, 1) for v in rel[worst].tolist()])
