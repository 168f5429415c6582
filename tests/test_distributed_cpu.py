"""Host-side multi-GPU logic (dropoutdecoding_b200/distributed.py) on CPU: world_size-2 gloo process group.
The combine rules are checked against the oracle's select_by_vote / select_by_average
(models/llava.py:22-52), the sequence sharding against a single-process run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from dropoutdecoding_b200 import distributed as dd_dist


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, K: int, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    r, _, w = dd_dist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        g = torch.Generator().manual_seed(7)
        B, V, T = 5, 97, 6
        # ---- axis 2: members.  Logits with deliberate ties between ids so that the tie rules matter.
        logits = torch.randn(B, K, V, generator=g)
        logits[0, :, 11] = 50.0                       # all members agree
        logits[1, 0, 3] = 50.0; logits[1, 1, 4] = 50.0; logits[1, 2 % K, 4] = 60.0     # majority != first member
        for m in range(K):
            logits[2, m, m] = 50.0                    # all different: first member wins
        lo, hi = dd_dist.member_slice(K, world, rank)
        ids_local = logits[:, lo:hi].argmax(-1).to(torch.int32)
        ids = dd_dist.gather_member_ids(ids_local)
        winner, token = dd_dist.vote_from_ids(ids)
        for b in range(B):
            win, _ = oracle.select_by_vote([logits[b, m] for m in range(K)])
            assert int(winner[b]) == win and int(token[b]) == int(logits[b, win].argmax())
        mean = dd_dist.mean_logits(logits[:, lo:hi].sum(1).clone(), K)
        for b in range(B):
            want = oracle.select_by_average([logits[b, m] for m in range(K)])
            assert torch.allclose(mean[b], want, atol=1e-5)
        rows = torch.arange(2 * B * K * 3 * 4, dtype=torch.float32).view(2, B, K, 3, 4).to(torch.bfloat16)
        got = dd_dist.gather_member_rows(rows[:, :, lo:hi].contiguous())
        assert torch.equal(got, rows)
        # ---- axis 1: sequences
        n_total = 7
        all_tokens = torch.randint(0, 1000, (n_total, T), generator=g, dtype=torch.int32)
        s_lo, s_hi = dd_dist.shard_range(n_total, world, rank)
        out = dd_dist.gather_tokens(all_tokens[s_lo:s_hi].clone(), n_total)
        assert torch.equal(out, all_tokens)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [2, 4])
def test_two_rank_gloo_combine_rules(K):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "worker failed (see stderr)"
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [dd_dist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dd_dist.member_slice(3, 2, 0)
    assert dd_dist.member_slice(8, 4, 3) == (6, 8)


def test_vote_from_ids_single_process_matches_reference_rule():
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 4, (200, 5), generator=g, dtype=torch.int32)      # few distinct ids -> many ties
    winner, token = dd_dist.vote_from_ids(ids)
    for b in range(ids.shape[0]):
        fake = [torch.nn.functional.one_hot(ids[b, m].long(), 8).float() for m in range(5)]
        win, _ = oracle.select_by_vote(fake)
        assert int(winner[b]) == win and int(token[b]) == int(ids[b, win])
