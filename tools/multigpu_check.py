#!/usr/bin/env python
"""Multi-GPU checks, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

1. member sharding: a MemberShardedEngine (K=4 members over the ranks, per-step all_gather of the member ids and
   of the staged K/V rows) must generate exactly the tokens of the single-GPU engine with all 4 members;
2. sequence sharding: B sequences split over the ranks with seq_base-keyed Philox masks must equal the
   single-GPU batch, after gather_tokens.
Prints one JSON line on rank 0 and exits non-zero on any mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200 import distributed as dd_dist  # noqa: E402
from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine  # noqa: E402


def main():
    rank, local, world = dd_dist.init_from_env()
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    cfg = DecoderConfig.tiny()
    w = DecoderWeights.random(cfg, dev, seed=0, head_std=0.08)          # same seed -> same weights on every rank
    probs, n_vis, S0, T = [0.1, 0.3, 0.5, 0.7], 576, 593, 12
    ok = True
    report = {"world": world}

    def prime(eng, B, seq0=0):
        g = torch.Generator(device=dev).manual_seed(5)
        kc = torch.randn(cfg.n_layers, 8, cfg.n_kv_heads, S0, cfg.head_dim, generator=g, device=dev)
        vc = torch.randn(cfg.n_layers, 8, cfg.n_kv_heads, S0, cfg.head_dim, generator=g, device=dev)
        eng.load_prefill(kc[:, seq0:seq0 + B], vc[:, seq0:seq0 + B], [5] * B)
        epi = torch.rand(8, n_vis, generator=g, device=dev) * 0.2 + 0.75
        ids = torch.randint(0, cfg.vocab, (8, n_vis, 5), generator=g, device=dev, dtype=torch.int32)
        eng.epi.copy_(epi[seq0:seq0 + B]); eng.topk_ids.copy_(ids[seq0:seq0 + B])
        eng.cur_token.fill_(1234); eng.tokens_out[:, 0] = 1234

    with torch.no_grad():
        # ---- 1. members over ranks (B = 2 sequences on every rank) ----------------------------------------
        ref = DropoutDecodeEngine(w, 2, n_vis, S0 + T + 4, probs=probs, max_steps=32, seed=24, speculative="off")
        prime(ref, 2)
        ref.decode(T, use_graph=False)
        for graph in (True, False):       # the whole step incl. the NCCL all-gather as one CUDA graph, and eagerly
            shard = dd_dist.make_member_sharded_engine(w, 2, n_vis, S0 + T + 4, probs, max_steps=32, seed=24)
            prime(shard, 2)
            shard.decode(T, use_graph=graph)
            torch.cuda.synchronize()
            same = (torch.equal(ref.generated(), shard.generated()) and torch.equal(ref.k_cache, shard.k_cache)
                    and torch.equal(ref.token_stats[:T + 1], shard.token_stats[:T + 1]))
            report["member_sharding_tokens_equal" + ("_graph" if graph else "_eager")] = bool(same)
            ok &= same
            del shard
        # ---- 2. sequences over ranks (B = 8 in total) ------------------------------------------------------
        full = DropoutDecodeEngine(w, 8, n_vis, S0 + T + 4, probs=probs[:3], max_steps=32, seed=24)
        prime(full, 8)
        full.decode(T, use_graph=True)
        # on one GPU emulate rank 1 of 2: sequences 4..7 decoded on their own
        lo, hi = dd_dist.shard_range(8, world, rank) if world > 1 else (4, 8)
        # (the split-KV factor of the attention is chosen from the number of (sequence, kv head) pairs on the GPU and is
        #  part of the numerics -- another split, another fp32 summation order; a shard that has to reproduce the unsharded
        #  tokens bit for bit pins it.  Unsplit ranges -- every batch that fills the SMs -- need nothing.)
        part = DropoutDecodeEngine(w, hi - lo, n_vis, S0 + T + 4, probs=probs[:3], max_steps=32, seed=24, seq_base=lo,
                                   attn_splits=full.pb.splits)
        prime(part, hi - lo, seq0=lo)
        part.decode(T, use_graph=True)
        torch.cuda.synchronize()
        if world > 1:
            same = torch.equal(dd_dist.gather_tokens(part.generated().contiguous(), 8), full.generated())
        else:
            same = torch.equal(part.generated(), full.generated()[lo:hi])
        report["sequence_sharding_tokens_equal"] = bool(same)
        ok &= same
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["ok"] = bool(flag.item())
        print(json.dumps(report), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
