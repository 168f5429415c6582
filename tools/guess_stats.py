#!/usr/bin/env python
"""How often is the unmasked argmax y_init of a step also the step's voted token, with ensemble member 0 the winner?
(the guess a pipelined-preliminary schedule would build the NEXT step's unmasked context on).  bench.py's workload
(LLaVA-1.5-7B random init, synthetic images), reference two-pass order, eager steps.
Usage: python tools/guess_stats.py [--batch 16] [--tokens 64]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--tokens", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
with torch.no_grad():
    model = bench.build_llava15_7b(dev)
    ids_h, px_h = bench.synthetic_inputs(a.batch, seed0=1001)
    merged, start, n_vis = model._merge(ids_h.to(dev), px_h.to(dev, torch.bfloat16))
    S0 = merged.shape[1]
    probe = model._engine(1, n_vis, S0 + 8, est_seq=S0 + 4)
    eng = bench._fresh_engine(model, probe, a.batch, n_vis, S0 + a.tokens + 8, S0 + a.tokens // 2, speculative="off")
    hidden = eng.prefill(merged.to(torch.bfloat16), [start] * a.batch)
    eng.seq_lens.fill_(S0)
    eng.uncertainty_stage(hidden[:, start:start + n_vis].contiguous())
    eng.first_token(hidden[:, -1].contiguous())
    n = tok_eq = both = y_in_members = 0
    relevant = 0
    for _ in range(a.tokens - 1):
        eng.step()
        y, t, w = eng.y_init.cpu(), eng.next_token.cpu(), eng.winner.cpu()
        ids = eng.member_ids.cpu()
        n += a.batch
        tok_eq += int((y == t).sum())
        both += int(((y == t) & (w == 0)).sum())
        y_in_members += int((ids[:, 0] == y).sum())
    print(json.dumps({"batch": a.batch, "steps": a.tokens - 1, "p_vote_is_y_init": round(tok_eq / n, 4),
                      "p_vote_is_y_init_and_member0_wins": round(both / n, 4),
                      "p_member0_argmax_is_y_init": round(y_in_members / n, 4)}))
