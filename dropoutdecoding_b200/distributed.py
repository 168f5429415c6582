"""Multi-GPU layout of the Dropout-Decoding hot path on one NVLink/NVSwitch box.

The reference has no distributed code at all (it decodes one image at a time, chair_test/chair_test.py:274);
the path nevertheless shards along two independent axes, one process per GPU, `torch.distributed`
(NCCL over NVLink on the GPUs, gloo in the CPU tests) as plumbing:

1. **sequences** (default, `shard_range` / `gather_tokens`): every rank owns B/G sequences with all K
   members co-located, so the shared prefix K/V is still read once per pass; weights are replicated; there
   is NO per-step collective.  The sampler's Philox stream is keyed by the global sequence index
   (`seq_base`), so the tokens do not depend on G.  One `all_gather` of the [B/G, T] token block at the end.
2. **ensemble members** (`make_member_sharded_engine`, for B < G -- BASELINE.json configs 2 and 5): every rank runs
   the preliminary pass (replicated) and K/G of the masked members; per step ONE collective: an all-gather of
   each rank's exchange record {member argmax ids, token statistics, staged new-token K/V rows (and the logits in
   mean mode)}, which the rank's kernels write directly into the send buffer and `dd_member_finish` consumes
   directly (vote + commit of the winner's rows) -- the caches stay bit-identical on all ranks; the step is one
   CUDA graph.

The combine rules restate `select_by_vote` / `select_by_average` (models/llava.py:22-52) on gathered
tensors; they run on whatever device the tensors live on (CUDA in production, CPU under gloo in tests).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------
# process group plumbing
# ------------------------------------------------------------------------------------------------
def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default group when world > 1.
    backend defaults to nccl when CUDA is available, else gloo."""
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend, **kw)
    return rank, local, world


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def rank_of(group=None) -> int:
    return dist.get_rank(group) if dist.is_initialized() else 0


# ------------------------------------------------------------------------------------------------
# axis 1: sequences
# ------------------------------------------------------------------------------------------------
def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for `rank` (the first n % world ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_tokens(local_tokens: torch.Tensor, n_total: int, group=None, pad_value: int = 0) -> torch.Tensor:
    """local_tokens int [b_local, T] of this rank's shard (shard_range order) -> [n_total, T] on every rank."""
    world = world_size(group)
    if world == 1:
        return local_tokens
    T = local_tokens.shape[1]
    b_max = -(-n_total // world)
    buf = torch.full((b_max, T), pad_value, dtype=local_tokens.dtype, device=local_tokens.device)
    buf[:local_tokens.shape[0]] = local_tokens
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        parts.append(out[r][:hi - lo])
    return torch.cat(parts, dim=0)


# ------------------------------------------------------------------------------------------------
# axis 2: ensemble members -- the combine rules
# ------------------------------------------------------------------------------------------------
def member_slice(K: int, world: int, rank: int) -> Tuple[int, int]:
    """Members [lo, hi) run by `rank`.  K must be a multiple of the world size so that every rank stages the
    same number of rows (static shapes: one all_gather, no padding)."""
    if K % world:
        raise ValueError(f"K={K} members do not divide over {world} ranks")
    per = K // world
    return rank * per, (rank + 1) * per


def vote_from_ids(member_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """select_by_vote (models/llava.py:22-36) on ids: member_ids int [B, K] -> (winner [B], token [B]).
    The winning id is the most frequent one, ties between ids go to the id seen first (Counter.most_common),
    and the winner is the first member holding that id."""
    B, K = member_ids.shape
    same = member_ids[:, :, None] == member_ids[:, None, :]
    counts = same.sum(-1)                                                # [B, K] votes for each member's id
    idx = torch.arange(K, device=member_ids.device)
    score = counts * (K + 1) - idx                                       # max count, then the earliest member
    winner = torch.argmax(score, dim=1)
    token = member_ids.gather(1, winner[:, None])[:, 0]
    return winner.to(torch.int32), token.to(member_ids.dtype)


def gather_member_ids(local_ids: torch.Tensor, group=None) -> torch.Tensor:
    """local_ids int32 [B, K/G] (this rank's members, member_slice order) -> [B, K] on every rank."""
    world = world_size(group)
    if world == 1:
        return local_ids
    out = [torch.empty_like(local_ids) for _ in range(world)]
    dist.all_gather(out, local_ids.contiguous(), group=group)
    return torch.cat(out, dim=1)


def mean_logits(local_sum: torch.Tensor, K: int, group=None) -> torch.Tensor:
    """select_by_average (models/llava.py:37-52): local_sum fp32 [B, V] = sum of this rank's member logits ->
    the member mean on every rank (one all_reduce)."""
    if world_size(group) > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group)
    return local_sum / float(K)


def gather_member_rows(rows_local: torch.Tensor, group=None) -> torch.Tensor:
    """rows_local bf16 [L, B, K/G, H_kv, D] (staged new-token K or V rows) -> [L, B, K, H_kv, D]."""
    world = world_size(group)
    if world == 1:
        return rows_local
    out = [torch.empty_like(rows_local) for _ in range(world)]
    dist.all_gather(out, rows_local.contiguous(), group=group)
    return torch.cat(out, dim=2)


# ------------------------------------------------------------------------------------------------
# member-sharded engine (CUDA only; imported lazily so that the helpers above work without the library)
# ------------------------------------------------------------------------------------------------
def exchange_layout(B: int, mpr: int, n_layers: int, h_kv: int, d: int):
    """Byte layout of one rank's per-step exchange record (dd_member_finish, include/dropdec.h):
    ids int32 [B, mpr] | token statistics fp32 [mpr, B, 3] | staged K rows bf16 [L, B, mpr, H_kv, D] | staged V rows.
    Returns (ids_off, stats_off, k_off, v_off, record_bytes), every offset a multiple of 256."""
    al = lambda x: -(-x // 256) * 256
    ids_off = 0
    stats_off = al(ids_off + B * mpr * 4)
    k_off = al(stats_off + mpr * B * 3 * 4)
    rows = n_layers * B * mpr * h_kv * d * 2
    v_off = al(k_off + rows)
    return ids_off, stats_off, k_off, v_off, al(v_off + rows)


def make_member_sharded_engine(weights, batch: int, n_vis: int, max_seq: int, probs: Sequence[float], group=None,
                               **kw):
    """A DropoutDecodeEngine whose masked pass runs only this rank's K/G members (B < G: BASELINE configs 2 and 5).

    Per step and rank: the preliminary pass (replicated -- every rank needs y_init for the masks), all K masks
    (replicated, cheap; cumulative masks depend on the earlier members), the rank's K/G masked contexts, then ONE
    collective: an all-gather of the rank's exchange record.  The record IS the output buffer of the rank's kernels --
    the QKV projection's RoPE epilogue stages the members' new K/V rows in it, dd_aggregate writes the member argmax
    ids into it, dd_token_stats the per-member statistics -- so nothing is packed or copied before the collective, and
    dd_member_finish consumes the gathered records directly: vote, winner, token, statistics and the commit of the
    winner's K/V rows into the cache, in one launch.  The whole step (both passes, the NCCL all-gather, the finish) is
    one CUDA graph; the caches stay bit-identical on all ranks, and the tokens equal the single-GPU engine's.
    Logit-mean mode (select_by_average) gathers the members' logits as well and sums them in member order, i.e. exactly
    the sequential fp32 sum of the single-GPU kernel and of the reference's numpy mean."""
    from . import ops
    from .engine import DropoutDecodeEngine

    class MemberShardedEngine(DropoutDecodeEngine):
        def __init__(self):
            self.group = group
            self.world, self.rank = world_size(group), rank_of(group)
            self.K_total = len(probs)
            self.m_lo, self.m_hi = member_slice(self.K_total, self.world, self.rank)
            self.mpr = self.m_hi - self.m_lo
            kw["speculative"] = "off"       # members live on different ranks: the reference's two-pass order
            super().__init__(weights, batch, n_vis, max_seq, probs=probs, **kw)
            if self.vote_on != "logits":
                raise ops.DropDecError("member sharding votes on logits (the InstructBLIP hidden-state vote is not sharded)")
            cfg, dev, B, mpr = self.cfg, self.dev, batch, self.mpr
            L, Hkv, D, V = cfg.n_layers, cfg.n_kv_heads, cfg.head_dim, cfg.vocab
            self._ids_off, self._stats_off, self._k_off, self._v_off, self._rec = exchange_layout(B, mpr, L, Hkv, D)
            mean = self.aggregate == "mean"
            self._logit_off = self._rec
            if mean:                        # the members' logits ride along: fp32 [B, mpr, V]
                self._rec += -(-B * mpr * V * 4 // 256) * 256
            self._send = torch.zeros(self._rec, dtype=torch.uint8, device=dev)
            self._recv = torch.zeros(self.world * self._rec, dtype=torch.uint8, device=dev)
            view = lambda off, n, dt, shape: self._send[off:off + n].view(dt).view(*shape)
            self._ids_local = view(self._ids_off, B * mpr * 4, torch.int32, (B, mpr))
            self._stats_local = view(self._stats_off, mpr * B * 3 * 4, torch.float32, (mpr, B, 3))
            rows = L * B * mpr * Hkv * D * 2
            # pass B of the local members: its K/V staging buffers live inside the send buffer
            self.pb = self._alloc_pass(mpr, kw.get("est_seq") or max_seq)
            self.pb.k_new = view(self._k_off, rows, torch.bfloat16, (L, B, mpr, Hkv, D))
            self.pb.v_new = view(self._v_off, rows, torch.bfloat16, (L, B, mpr, Hkv, D))
            if mean:
                self._logits_local = view(self._logit_off, B * mpr * V * 4, torch.float32, (B, mpr, V))
                self._logits_all = torch.zeros(B, self.K_total, V, dtype=torch.float32, device=dev)
            self._win_local = torch.zeros(B, dtype=torch.int32, device=dev)
            self._tok_local = torch.zeros(B, dtype=torch.int32, device=dev)
            self._members = [torch.full((B,), j, dtype=torch.int32, device=dev) for j in range(mpr)]
            self._local_mask = torch.zeros(B, mpr, self.words, dtype=torch.int32, device=dev)
            self.use_graph = os.environ.get("DD_MEMBER_GRAPH", "1") == "1"

        def _exchange(self) -> None:
            if self.world > 1:
                dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
            else:
                self._recv.copy_(self._send)

        def step(self) -> None:
            self.launches_per_step = 0
            V, mean = self.cfg.vocab, self.aggregate == "mean"
            logits_a = self._forward(self.pa, None)                       # replicated preliminary pass
            ops.aggregate(logits_a, "vote", member_ids=self.member_ids_a, winner=self.winner_a, token=self.y_init,
                          vocab=V)
            self._sample(self._step_uniforms(), self.keep_relevant, self.bitmask, self.n_masked)   # all K masks
            self._local_mask.copy_(self.bitmask[:, self.m_lo:self.m_hi])
            logits_b = self._forward(self.pb, self._local_mask)           # this rank's members
            # the rank's record: ids, per-member statistics (and logits in mean mode) go straight into the send buffer
            ops.aggregate(logits_b, "vote", member_ids=self._ids_local, winner=self._win_local, token=self._tok_local,
                          vocab=V)
            for j in range(self.mpr):
                ops.token_stats(logits_b, V, self._stats_local, member=self._members[j], slot_offset=j)
            if mean:
                self._logits_local.copy_(logits_b[:, :, :V])
            self._exchange()                                              # the per-step collective
            if mean:
                rec = self._recv.view(self.world, self._rec)
                for r in range(self.world):     # member order = rank order x local order
                    self._logits_all[:, r * self.mpr:(r + 1) * self.mpr] = \
                        rec[r, self._logit_off:self._logit_off + self._logits_local.numel() * 4].view(torch.float32).view(
                            self.B, self.mpr, V)
                ops.aggregate(self._logits_all, "mean", member_ids=self.member_ids, winner=self.winner,
                              token=self.next_token, out_logits=self.out_logits, vocab=V)
                ops.token_stats(self.out_logits, V, self.token_stats, slot=self.step_ctr, slot_offset=1)
            ops.member_finish(self._recv, self._rec, self._ids_off, self._stats_off, self._k_off, self._v_off, self.world,
                              self.mpr, self.seq_lens, self.step_ctr, self.member_ids, self.winner, self.next_token,
                              None if mean else self.token_stats, self.k_cache, self.v_cache,
                              forced_winner=0 if mean else -1, forced_token=self.next_token if mean else None)
            ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
            self.cur_token.copy_(self.next_token)
            self.launches_per_step += 6 + self.mpr

        def decode(self, n_steps: int, use_graph: Optional[bool] = None, stop_token: Optional[int] = None) -> int:
            """n_steps ensemble steps on every rank (the ranks must call this together); returns the steps done.
            stop_token is accepted for interface parity and ignored: the step count is fixed so that the ranks stay in
            lock step (callers trim at EOS afterwards, as generate() does)."""
            return super().decode(n_steps, use_graph=self.use_graph if use_graph is None else use_graph)

    return MemberShardedEngine()
