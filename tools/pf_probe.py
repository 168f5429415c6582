#!/usr/bin/env python
"""Does L2 residency speed up the skinny GEMM, and does dd_hint_l2_prefetch achieve it?  (round-2 probe)

For each decoder shape at M=4: the GEMM timed (CUDA events around the one launch) with the weights cold (L2 flushed),
warm (the same launch repeated; only shapes that fit L2 stay resident) and after a hinted RMSNorm launch that
prefetches the first `mb` MB.  Also the hinted RMSNorm / attention launches themselves.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
M = 4
xn = torch.randn(M, 4096, device=dev).bfloat16()
nw = torch.ones(4096, device=dev).bfloat16()


def timed(fn, pre=None, reps=7):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        if pre:
            pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 2)


for name, N, K, epi in (("o_proj", 4096, 4096, ops.EPI_RESIDUAL), ("qkv", 12288, 4096, ops.EPI_NONE),
                        ("gate_up", 11008, 4096, ops.EPI_SWIGLU), ("down", 4096, 11008, ops.EPI_RESIDUAL)):
    w = torch.randn((2 * N if epi == ops.EPI_SWIGLU else N), K, device=dev).bfloat16()
    x = torch.randn(M, K, device=dev).bfloat16()
    res = torch.randn(M, N, device=dev).bfloat16() if epi == ops.EPI_RESIDUAL else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    gemm = lambda: ops.skinny_gemm(x, w, epi, residual=res, out=out)
    mbytes = w.numel() * 2 / 2**20
    row = dict(shape=name, weight_mb=round(mbytes, 1), cold_us=timed(gemm), warm_us=timed(gemm, pre=gemm))
    for mb in (8, 24, 48, 96):
        if mb > mbytes + 24:
            continue
        def pre(mb=mb):
            ops.hint_l2_prefetch(w, 0, mb << 20)
            ops.rmsnorm(xn, nw, 1e-5)
            torch.cuda._sleep(40000)     # ~20 us: let the prefetch land
        row[f"after_pf{mb}_us"] = timed(gemm, pre=pre)
        def pre0(mb=mb):
            ops.hint_l2_prefetch(w, 0, mb << 20)
            ops.rmsnorm(xn, nw, 1e-5)
        row[f"after_pf{mb}_nosleep_us"] = timed(gemm, pre=pre0)
    print(json.dumps(row), flush=True)

big = torch.randn(96 << 20, device=dev).bfloat16()
row = dict(kernel="rmsnorm", plain_us=timed(lambda: ops.rmsnorm(xn, nw, 1e-5)))
for mb in (8, 24, 48, 96):
    def f(mb=mb):
        ops.hint_l2_prefetch(big, 0, mb << 20)
        ops.rmsnorm(xn, nw, 1e-5)
    row[f"pf{mb}_us"] = timed(f)
print(json.dumps(row), flush=True)
