#!/usr/bin/env python
"""Decode-step timing on random state (no HF model, no prefill): ms per ensemble step.

    [DD_LIB=... DD_PDL=0|1] python tools/stepbench.py [--batch B] [--steps N] [--seq S] [--model 7b|next|tiny]
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
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=63)
ap.add_argument("--seq", type=int, default=593)
ap.add_argument("--n-vis", type=int, default=576)
ap.add_argument("--model", default="7b")
ap.add_argument("--probs", default="0.3,0.5,0.7")
ap.add_argument("--mode", default="cumulative")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--vote-on", default="logits")
ap.add_argument("--topk", type=int, default=5)
ap.add_argument("--stop-token", type=int, default=None, help="decode with an EOS id the model never emits (host-side EOS checks on)")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = {"7b": DecoderConfig.llava15_7b, "next": DecoderConfig.llava_next_mistral_7b, "tiny": DecoderConfig.tiny,
       "iblip": lambda: DecoderConfig(32, 4096, 32, 32, 128, 11008, 32001, 1e-5, 10000.0)}[a.model]()
probs = [float(p) for p in a.probs.split(",")]
torch.manual_seed(0)
with torch.no_grad():
    w = DecoderWeights.random(cfg, dev)
    eng = DropoutDecodeEngine(w, a.batch, a.n_vis, a.seq + a.steps + 8, probs=probs, mode=a.mode, topk=a.topk,
                              max_steps=a.steps + 8, est_seq=a.seq + a.steps // 2, vote_on=a.vote_on)
    eng.k_cache.normal_(); eng.v_cache.normal_()
    eng.epi.uniform_(0.75, 0.95)
    eng.topk_ids.random_(0, cfg.vocab)

    def reset():
        eng.seq_lens.fill_(a.seq); eng.vis_start.fill_(5); eng.step_ctr.zero_(); eng.cur_token.fill_(1234)

    reset()
    eng.decode(2, use_graph=not a.no_graph)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode(a.steps, use_graph=not a.no_graph, stop_token=a.stop_token)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    wbytes = 2 * w.stream_bytes_per_pass()
    kv = 2 * a.batch * (a.seq + a.steps / 2) * cfg.n_layers * cfg.n_kv_heads * cfg.head_dim * 2 * 2
    print(json.dumps(dict(lib=os.environ.get("DD_LIB", "default"), pdl=os.environ.get("DD_PDL", "auto"), pair_gemm=os.environ.get("DD_PAIR_GEMM", "1"), model=a.model,
                          batch=a.batch, K=len(probs), ms_per_step=round(best, 4),
                          tokens_per_s=round(a.batch * 1e3 / best, 1),
                          step_roofline_frac=round((wbytes + kv) / (best * 1e-3) / 6551.4e9, 3),
                          launches=eng.launches_per_step, spec=("host" if eng.spec_active else "batch" if eng.spec_batch else "off"),
                          gemm=eng.gemm_backend, lookahead=([eng.n_chain, eng.n_chain_ok] if getattr(eng, "lookahead", False) else ("batch" if getattr(eng, "batch_la", False) else None)), passes=eng.n_spec_steps, fallback_rate=round(eng.fallback_rate(), 3))))
