#!/usr/bin/env python
"""ms per pass of the two pass forms of the one-sequence schedule, replayed from CUDA graphs on fixed random state:
the speculative head pass (K+1 rows, one attention window) and the lookahead chain pass (2K+1 rows, two windows).
    [DD_ATTN_SPLITS=n] python tools/pass_forms.py [--seq 625]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, default=625)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = DecoderConfig.llava15_7b()
with torch.no_grad():
    w = DecoderWeights.random(cfg, dev)
    eng = DropoutDecodeEngine(w, 1, 576, a.seq + 80, max_steps=80, est_seq=a.seq)
    eng.k_cache.normal_(); eng.v_cache.normal_()
    eng.epi.uniform_(0.75, 0.95); eng.topk_ids.random_(0, cfg.vocab)
    eng.seq_lens.fill_(a.seq); eng.vis_start.fill_(5); eng.cur_token.fill_(1234)
    eng._g_win.fill_(1); eng._g_tok.fill_(777)
    out = {"splits": eng.pc.splits, "seq": a.seq}
    for name, fn in (("head_pass_ms", lambda: eng._forward(eng.pc, eng.spec_mask)), ("chain_pass_ms", eng._spec_chain)):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / 20, 4)
    print(json.dumps(out))
