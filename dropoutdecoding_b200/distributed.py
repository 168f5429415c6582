"""Multi-GPU layout of the Dropout-Decoding hot path on one NVLink/NVSwitch box.

The reference has no distributed code at all (it decodes one image at a time, chair_test/chair_test.py:274);
the path nevertheless shards along two independent axes, one process per GPU, `torch.distributed`
(NCCL over NVLink on the GPUs, gloo in the CPU tests) as plumbing:

1. **sequences** (default, `shard_range` / `gather_tokens`): every rank owns B/G sequences with all K
   members co-located, so the shared prefix K/V is still read once per pass; weights are replicated; there
   is NO per-step collective.  The sampler's Philox stream is keyed by the global sequence index
   (`seq_base`), so the tokens do not depend on G.  One `all_gather` of the [B/G, T] token block at the end.
2. **ensemble members** (`MemberShardedEngine`, for B < G -- BASELINE.json configs 2 and 5): every rank runs
   the preliminary pass (replicated) and K/G of the masked members; per step ONE exchange:
   `all_gather` of the member argmax ids (vote) or `all_reduce(sum)` of the [B, V] fp32 logits (mean),
   then an `all_gather` of the members' staged new-token K/V rows from which every rank commits the
   winner's -- the caches stay bit-identical on all ranks.

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
def make_member_sharded_engine(weights, batch: int, n_vis: int, max_seq: int, probs: Sequence[float], group=None,
                               **kw):
    """A DropoutDecodeEngine whose masked pass runs only this rank's K/G members; see the module docstring."""
    from . import ops
    from .engine import DropoutDecodeEngine

    class MemberShardedEngine(DropoutDecodeEngine):
        def __init__(self):
            self.group = group
            self.world, self.rank = world_size(group), rank_of(group)
            self.K_total = len(probs)
            self.m_lo, self.m_hi = member_slice(self.K_total, self.world, self.rank)
            kw.setdefault("speculative", "off")     # members live on different ranks: the reference's two-pass order
            super().__init__(weights, batch, n_vis, max_seq, probs=probs, **kw)
            if self.vote_on != "logits":
                raise ops.DropDecError("member sharding votes on logits (the InstructBLIP hidden-state vote is not sharded)")
            # pass B buffers for the local members only
            self.pb = self._alloc_pass(self.m_hi - self.m_lo, kw.get("est_seq") or max_seq)
            self._ids_local = torch.zeros(batch, self.m_hi - self.m_lo, dtype=torch.int32, device=self.dev)
            self._win_local = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            self._tok_local = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            self._sum_local = torch.zeros(batch, self.cfg.vocab, dtype=torch.float32, device=self.dev)

        def step(self) -> None:
            self.launches_per_step = 0
            logits_a = self._forward(self.pa, None)                       # replicated preliminary pass
            ops.aggregate(logits_a, "vote", member_ids=self.member_ids_a, winner=self.winner_a, token=self.y_init,
                          vocab=self.cfg.vocab)
            u = None
            if self.uniforms is not None:
                self._u_step.copy_(self.uniforms.index_select(
                    0, self.step_ctr.long().clamp(max=self.uniforms.shape[0] - 1))[0])
                u = self._u_step
            # all K masks are sampled everywhere (cumulative masks depend on the earlier members); cheap
            ops.sample_masks(self.epi, self.topk_ids, self.y_init, self.probs, mode=self.mode,
                             keep_relevant=self.keep_relevant, uniforms=u, seed=self.seed, offset=0,
                             step=self.step_ctr, out=self.bitmask, n_masked=self.n_masked, seq_base=self.seq_base)
            local_mask = self.bitmask[:, self.m_lo:self.m_hi].contiguous()
            logits_b = self._forward(self.pb, local_mask)                 # this rank's members
            V = self.cfg.vocab
            if self.aggregate == "vote":
                ops.aggregate(logits_b, "vote", member_ids=self._ids_local, winner=self._win_local,
                              token=self._tok_local, vocab=V)
                self.member_ids.copy_(gather_member_ids(self._ids_local, self.group))      # the per-step exchange
                winner, token = vote_from_ids(self.member_ids)
                self.winner.copy_(winner)
                self.next_token.copy_(token)
            else:
                self._sum_local.copy_(logits_b[:, :, :V].sum(1))
                self.out_logits.copy_(mean_logits(self._sum_local, self.K_total, self.group))
                self.next_token.copy_(torch.argmax(self.out_logits, dim=-1).to(torch.int32))
                self.winner.zero_()
            k_all = gather_member_rows(self.pb.k_new, self.group)         # [L, B, K, H_kv, D]
            v_all = gather_member_rows(self.pb.v_new, self.group)
            ops.commit_winner(k_all, v_all, self.winner, self.seq_lens, self.k_cache, self.v_cache)
            ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
            self.cur_token.copy_(self.next_token)

        def decode(self, n_steps: int, use_graph: bool = False) -> None:
            # eager: the per-step NCCL exchange is enqueued on the compute stream between the kernels
            super().decode(n_steps, use_graph=False)

    return MemberShardedEngine()
