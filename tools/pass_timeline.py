#!/usr/bin/env python
"""Phase timeline of dd_decoder_pass (CTA 0's %globaltimer stamps): us of work and of barrier wait per phase kind.

    python tools/pass_timeline.py [--rows 4|7] [--model 7b]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="7b")
ap.add_argument("--seq", type=int, default=593)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = {"7b": DecoderConfig.llava15_7b, "tiny": DecoderConfig.tiny}[a.model]()
with torch.no_grad():
    w = DecoderWeights.random(cfg, dev)
    eng = DropoutDecodeEngine(w, 1, 576, a.seq + 72, max_steps=72, est_seq=a.seq + 32)
    eng.k_cache.normal_(); eng.v_cache.normal_()
    eng.seq_lens.fill_(a.seq); eng.vis_start.fill_(5); eng.cur_token.fill_(1234)
    L = cfg.n_layers
    tl = torch.zeros(2 * (5 * L + 2) + 64, dtype=torch.int64, device=dev)
    eng.pass_timeline = tl
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._forward(eng.pc, eng.spec_mask)
        e1.record()
        torch.cuda.synchronize()
    tr = tl.cpu()[2 * (5 * L + 2):].view(4, 16).double() / 1e3
    t = tl.cpu()[:2 * (5 * L + 2)].view(-1, 2).double() / 1e3          # [phase, {start, end}] us
    names = ["qkv", "attn", "o_proj", "gate_up", "down"]
    work = {n: 0.0 for n in names}
    wait = {n: 0.0 for n in names}
    for l in range(L):
        for i, n in enumerate(names):
            ph = 1 + 5 * l + i
            work[n] += float(t[ph, 1] - t[ph, 0]) / L
            wait[n] += float(t[ph + 1, 0] - t[ph, 1]) / L
    print(json.dumps(dict(rows=eng.pc.M, pass_us=round(float(t[-1, 1] - t[0, 0]), 1), event_ms=round(e0.elapsed_time(e1), 3),
                          embed_us=round(float(t[1, 0] - t[0, 0]), 2), head_us=round(float(t[-1, 1] - t[-1, 0]), 2),
                          work_us={k: round(v, 2) for k, v in work.items()},
                          barrier_wait_after_us={k: round(v, 2) for k, v in wait.items()})))
    if float(tr.abs().sum()) > 0:     # library built with -DDD_PASS_TRACE: CTA 0, layer 1
        for k, n in enumerate(["qkv", "o_proj", "gate_up", "down"]):
            ph = 1 + 5 * 1 + (k if k == 0 else k + 1)
            row = tr[k]
            rel = [round(float(v - t[ph, 0]), 2) for v in row if v > 0]
            print(n, "phase start -> {entry, tile red start, tile red end, ...} us:", rel, "phase end", round(float(t[ph, 1] - t[ph, 0]), 2))
