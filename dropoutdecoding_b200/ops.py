"""Torch-tensor front end of the C ABI (include/dropdec.h).

Every function checks dtype / device / contiguity, passes raw device pointers and torch's current
CUDA stream to libdropdec.so, and raises DropDecError on a non-zero return code.  Nothing here
computes on the CPU and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DropDecError, check

check_rc = check

MASK_CUMULATIVE, MASK_INDEPENDENT, MASK_QUANTILE = 0, 1, 2
AGG_VOTE, AGG_MEAN = 0, 1
EPI_NONE, EPI_RESIDUAL, EPI_F32, EPI_SWIGLU, EPI_ROPE = 0, 1, 2, 3, 4
MASK_MODES = {"cumulative": MASK_CUMULATIVE, "independent": MASK_INDEPENDENT, "quantile": MASK_QUANTILE}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise DropDecError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise DropDecError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise DropDecError(f"{name} must be contiguous")
    return t


def set_chain_launch(mode) -> int:
    """dd_set_chain_launch: programmatic dependent launch for the dd_* kernels launched from now on -- False / 0 off,
    True / 1 every kernel, 2 only the skinny weight GEMMs, 3 every kernel but those (returns the previous mode)."""
    return int(_lib.load().dd_set_chain_launch(int(mode)))


def hint_l2_prefetch(t0: Optional[torch.Tensor], off0: int = 0, bytes0: int = 0, t1: Optional[torch.Tensor] = None,
                     off1: int = 0, bytes1: int = 0) -> None:
    """dd_hint_l2_prefetch: the next rmsnorm / decode_attention call also prefetches bytes [off, off + bytes) of up to
    two (contiguous) tensors into L2; ranges are clipped to the tensors."""
    def rng(t, off, n):
        if t is None or n <= 0:
            return None, 0
        total = t.numel() * t.element_size()
        off = min(off, total) & ~15
        return t.data_ptr() + off, (min(n, total - off) & ~15)
    p0, b0 = rng(t0, off0, bytes0)
    p1, b1 = rng(t1, off1, bytes1)
    if b0 == 0 and b1 == 0:
        return
    check(_lib.load().dd_hint_l2_prefetch(p0 if b0 else None, b0, p1 if b1 else None, b1), "dd_hint_l2_prefetch")


def mask_words(n_vis: int) -> int:
    return (n_vis + 31) // 32


# --------------------------------------------------------------------------------------------
def uncertainty_workspace(n_img: int, n_vis: int, vocab: int, topk: int, device) -> torch.Tensor:
    """Scratch of dd_uncertainty_project (dd_uncertainty_workspace_bytes): the logits stash of one image group plus the
    per-row partials; reusable across calls of the same shape."""
    need = _lib.load().dd_uncertainty_workspace_bytes(n_img, n_vis, vocab, topk)
    return torch.empty(max(need, 1), dtype=torch.uint8, device=device)


def uncertainty_project(h_vis: torch.Tensor, w_lm: torch.Tensor, topk: int = 5, round_logits: bool = False,
                        workspace: Optional[torch.Tensor] = None, check: bool = True, return_overflow: bool = False,
                        out: Optional[Sequence[torch.Tensor]] = None):
    """h_vis bf16 [n_img, n_vis, d], w_lm bf16 [V, d] -> (epi, alea, var fp32 [n_img, n_vis], topk_ids int32).
    out = (epi, alea, var, topk_ids): contiguous result tensors to write into (the engine's own state).

    A row whose top-k candidate list overflowed on the device (more logits tied at its threshold than the list holds)
    is an error, never a silently shorter list: check=True reads the device counter (one stream synchronisation) and
    raises; return_overflow=True appends the int32 [1] device counter to the result instead, for callers that check
    it at their next synchronisation point (the decode engine)."""
    L = _lib.load()
    _need(h_vis, torch.bfloat16, "h_vis"); _need(w_lm, torch.bfloat16, "w_lm")
    if h_vis.dim() == 2:
        h_vis = h_vis[None]
    n_img, n_vis, d = h_vis.shape
    V = w_lm.shape[0]
    if w_lm.shape[1] != d:
        raise DropDecError("w_lm and h_vis disagree on d")
    dev = h_vis.device
    if out is not None:
        epi, alea, var, ids = out
        for t, name in ((epi, "epi"), (alea, "alea"), (var, "var")):
            _need(t, torch.float32, name)
            if tuple(t.shape) != (n_img, n_vis):
                raise DropDecError(f"{name} must be [{n_img}, {n_vis}]")
        _need(ids, torch.int32, "topk_ids")
        if tuple(ids.shape) != (n_img, n_vis, topk):
            raise DropDecError(f"topk_ids must be [{n_img}, {n_vis}, {topk}]")
    else:
        epi = torch.empty(n_img, n_vis, dtype=torch.float32, device=dev)
        alea, var = torch.empty_like(epi), torch.empty_like(epi)
        ids = torch.empty(n_img, n_vis, topk, dtype=torch.int32, device=dev)
    need = L.dd_uncertainty_workspace_bytes(n_img, n_vis, V, topk)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    check_rc(L.dd_uncertainty_project(_ptr(h_vis), _ptr(w_lm), n_img, n_vis, d, V, topk, int(round_logits),
                                      _ptr(epi), _ptr(alea), _ptr(var), _ptr(ids), _ptr(workspace),
                                      workspace.numel() * workspace.element_size(), _stream()),
             "dd_uncertainty_project")
    off = L.dd_uncertainty_overflow_offset(n_img, n_vis, V, topk)
    overflow = workspace[off:off + 4].view(torch.int32)
    if return_overflow:
        return epi, alea, var, ids, overflow
    if check:
        raise_on_overflow(overflow)
    return epi, alea, var, ids


def raise_on_overflow(overflow: torch.Tensor) -> None:
    n = int(overflow.item())
    if n:
        raise DropDecError(f"dd_uncertainty_project: the top-k candidate list of {n} visual row(s) overflowed "
                           f"(DD_ERR_OVERFLOW): too many logits tied at the k-th largest value")


def sample_masks(epi: torch.Tensor, topk_ids: Optional[torch.Tensor], y_init: Optional[torch.Tensor],
                 probs: Sequence[float], mode: str = "cumulative", keep_relevant: bool = True,
                 uniforms: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
                 step: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 n_masked: Optional[torch.Tensor] = None, seq_base: int = 0,
                 step_per_seq: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """epi fp32 [B, n_vis] -> (bitmask uint32-as-int32 [B, K, words], n_masked int32 [B, K]).
    seq_base: global index of row 0 (keys the Philox stream, so sharded batches draw the same numbers)."""
    L = _lib.load()
    _need(epi, torch.float32, "epi")
    B, n_vis = epi.shape
    K = len(probs)
    words = mask_words(n_vis)
    topk = 0
    if keep_relevant:
        if topk_ids is None or y_init is None:
            raise DropDecError("keep_relevant needs topk_ids and y_init")
        _need(topk_ids, torch.int32, "topk_ids"); _need(y_init, torch.int32, "y_init")
        topk = topk_ids.shape[-1]
    if uniforms is not None:
        _need(uniforms, torch.float32, "uniforms")
        if tuple(uniforms.shape) != (B, K, n_vis):
            raise DropDecError(f"uniforms must be [{B},{K},{n_vis}]")
    if out is None:
        out = torch.empty(B, K, words, dtype=torch.int32, device=epi.device)
    if n_masked is None:
        n_masked = torch.empty(B, K, dtype=torch.int32, device=epi.device)
    if step is not None:
        _need(step, torch.int32, "step")
    probs_c = (C.c_double * K)(*[float(p) for p in probs])
    fn = L.dd_sample_masks_seq if step_per_seq else L.dd_sample_masks      # step: [B] counters vs one scalar
    check(fn(_ptr(epi), _ptr(topk_ids) if keep_relevant else None, topk,
                            _ptr(y_init) if keep_relevant else None, probs_c, K, _ptr(uniforms), seed, offset,
                            _ptr(step), MASK_MODES[mode], int(keep_relevant), B, n_vis, int(seq_base), _ptr(out), out.shape[-1],
             _ptr(n_masked), _stream()), "dd_sample_masks")
    return out, n_masked


def schedule_flags(flags: torch.Tensor, mask_a: Optional[torch.Tensor] = None, mask_b: Optional[torch.Tensor] = None,
                   tok_a: Optional[torch.Tensor] = None, tok_b: Optional[torch.Tensor] = None,
                   win_a: Optional[torch.Tensor] = None, win_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dd_schedule_flags: flags int32 [2] <- {all(tok_a == tok_b) & all(win_a == win_b), any(mask_a != mask_b)}."""
    _need(flags, torch.int32, "flags")
    for t, n in ((mask_a, "mask_a"), (mask_b, "mask_b"), (tok_a, "tok_a"), (tok_b, "tok_b"), (win_a, "win_a"),
                 (win_b, "win_b")):
        if t is not None:
            _need(t, torch.int32, n)
            if not t.is_contiguous():
                raise DropDecError(f"schedule_flags: {n} must be contiguous")
    if mask_a is not None and mask_a.numel() != mask_b.numel():
        raise DropDecError("schedule_flags: masks of different sizes")
    check(_lib.load().dd_schedule_flags(_ptr(tok_a), _ptr(tok_b), _ptr(win_a), _ptr(win_b),
                                        tok_a.numel() if tok_a is not None else 0, _ptr(mask_a), _ptr(mask_b),
                                        mask_a.numel() if mask_a is not None else 0, _ptr(flags), _stream()),
          "dd_schedule_flags")
    return flags


def philox_uniform(seed: int, offset: int, n: int, device="cuda") -> torch.Tensor:
    L = _lib.load()
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(L.dd_philox_uniform(seed, offset, n, _ptr(out), _stream()), "dd_philox_uniform")
    return out


def attention_workspace(B: int, M: int, H: int, D: int, num_splits: int, device) -> torch.Tensor:
    L = _lib.load()
    # zeroed once: the tail holds the ticket counters of the fused split merge (each call restores them to zero)
    return torch.zeros(max(L.dd_attention_workspace_bytes(B, M, H, D, num_splits), 1), dtype=torch.uint8,
                       device=device)


def decode_attention(q: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor,
                     v_cache: torch.Tensor, seq_lens: torch.Tensor, vis_start: Optional[torch.Tensor],
                     bitmask: Optional[torch.Tensor], n_vis: int, scale: float, num_splits: int = 1,
                     out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                     seq_active: Optional[torch.Tensor] = None, member_offset: int = 0,
                     members: Optional[int] = None, seq_active2: Optional[torch.Tensor] = None,
                     len_extra: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B,M,H,D], k_new/v_new [B,M,H_kv,D], caches [B,H_kv,S_max,D] (bf16) -> out [B,M,H,D] bf16.
    seq_active int32 [B]: sequences with 0 are skipped (their rows of `out` stay untouched) unless seq_active2 is set.
    member_offset / members: serve only members [member_offset, member_offset + members) of the M stored.
    len_extra int32 [B, M] in {0, 1}: member m attends keys [0, seq_lens[b] + len_extra[b, m]) -- two member windows
    (verify + lookahead contexts) in one launch, the K/V read once; needs num_splits == 1."""
    L = _lib.load()
    for t, n in ((q, "q"), (k_new, "k_new"), (v_new, "v_new"), (k_cache, "k_cache"), (v_cache, "v_cache")):
        _need(t, torch.bfloat16, n)
    _need(seq_lens, torch.int32, "seq_lens")
    B, M, H, D = q.shape
    H_kv, S_max = k_cache.shape[1], k_cache.shape[2]
    words = 0
    if bitmask is not None:
        _need(bitmask, torch.int32, "bitmask"); _need(vis_start, torch.int32, "vis_start")
        words = bitmask.shape[-1]
    if out is None:
        out = torch.empty_like(q)
    if workspace is None:
        workspace = attention_workspace(B, M, H, D, num_splits, q.device)
    if seq_active is not None:
        _need(seq_active, torch.int32, "seq_active")
    if seq_active2 is not None:
        _need(seq_active2, torch.int32, "seq_active2")
    if len_extra is not None:
        _need(len_extra, torch.int32, "len_extra")
        if tuple(len_extra.shape) != (B, M):
            raise DropDecError(f"len_extra must be [{B},{M}]")
    Mc = M - member_offset if members is None else members
    check(L.dd_decode_attention_windows(_ptr(q), _ptr(k_new), _ptr(v_new), _ptr(k_cache), _ptr(v_cache), _ptr(seq_lens),
                                        _ptr(vis_start), _ptr(bitmask), _ptr(seq_active), _ptr(seq_active2),
                                        _ptr(len_extra), B, Mc, member_offset, M, H, H_kv, D, S_max, n_vis, words,
                                        float(scale), num_splits, _ptr(out), _ptr(workspace),
                                        workspace.numel() * workspace.element_size(), _stream()), "dd_decode_attention")
    return out


def decode_attention_two_windows(q: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor,
                                 v_cache: torch.Tensor, seq_lens_a: torch.Tensor, seq_lens_b: torch.Tensor,
                                 members_a: int, vis_start: Optional[torch.Tensor], bitmask: Optional[torch.Tensor],
                                 n_vis: int, scale: float, num_splits: int, out: torch.Tensor,
                                 workspace: torch.Tensor) -> torch.Tensor:
    """dd_decode_attention_two_windows: members [0, members_a) attend keys [0, seq_lens_a), the rest [0, seq_lens_b), split
    key ranges, one launch.  workspace: two_window_attention_workspace."""
    L = _lib.load()
    for t, n in ((q, "q"), (k_new, "k_new"), (v_new, "v_new"), (k_cache, "k_cache"), (v_cache, "v_cache"), (out, "out")):
        _need(t, torch.bfloat16, n)
    _need(seq_lens_a, torch.int32, "seq_lens_a"); _need(seq_lens_b, torch.int32, "seq_lens_b")
    B, M, H, D = q.shape
    H_kv, S_max = k_cache.shape[1], k_cache.shape[2]
    words = 0
    if bitmask is not None:
        _need(bitmask, torch.int32, "bitmask"); _need(vis_start, torch.int32, "vis_start")
        words = bitmask.shape[-1]
    check(L.dd_decode_attention_two_windows(_ptr(q), _ptr(k_new), _ptr(v_new), _ptr(k_cache), _ptr(v_cache),
                                            _ptr(seq_lens_a), _ptr(seq_lens_b), _ptr(vis_start), _ptr(bitmask), B,
                                            members_a, M - members_a, H, H_kv, D, S_max, n_vis, words, float(scale),
                                            num_splits, _ptr(out), _ptr(workspace),
                                            workspace.numel() * workspace.element_size(), _stream()),
          "dd_decode_attention_two_windows")
    return out


def two_window_attention_workspace(B: int, M: int, H: int, D: int, num_splits: int, device) -> torch.Tensor:
    one = (_lib.load().dd_attention_workspace_bytes(B, M, H, D, num_splits) + 255) // 256 * 256
    return torch.zeros(2 * one, dtype=torch.uint8, device=device)


def commit_winner(k_new_all: torch.Tensor, v_new_all: torch.Tensor, winner: torch.Tensor, seq_lens: torch.Tensor,
                  k_cache_all: torch.Tensor, v_cache_all: torch.Tensor) -> None:
    """k_new_all [L,B,M,H_kv,D] -> caches [L,B,H_kv,S_max,D] at row seq_lens[b] of the winner member."""
    L = _lib.load()
    _need(k_new_all, torch.bfloat16, "k_new_all"); _need(k_cache_all, torch.bfloat16, "k_cache_all")
    _need(v_new_all, torch.bfloat16, "v_new_all"); _need(v_cache_all, torch.bfloat16, "v_cache_all")
    _need(winner, torch.int32, "winner"); _need(seq_lens, torch.int32, "seq_lens")
    Ln, B, M, H_kv, D = k_new_all.shape
    S_max = k_cache_all.shape[3]
    check(L.dd_commit_winner(_ptr(k_new_all), _ptr(v_new_all), _ptr(winner), _ptr(seq_lens), _ptr(k_cache_all),
                             _ptr(v_cache_all), Ln, B, M, H_kv, D, S_max, _stream()), "dd_commit_winner")


def aggregate(logits: torch.Tensor, mode: str = "vote", want_logits: bool = False,
              member_ids: Optional[torch.Tensor] = None, winner: Optional[torch.Tensor] = None,
              token: Optional[torch.Tensor] = None, out_logits: Optional[torch.Tensor] = None,
              vocab: Optional[int] = None, first_member: int = 0, members: Optional[int] = None):
    """logits fp32 [B, K, ld] (first `vocab` entries of each row valid; default all) ->
    (member_ids [B,K], winner [B], token [B], out_logits [B,vocab] | None).
    first_member / members select rows [first_member, first_member + members) of every sequence."""
    L = _lib.load()
    _need(logits, torch.float32, "logits")
    B, K_all, ld = logits.shape
    K = K_all - first_member if members is None else members
    if first_member < 0 or K < 1 or first_member + K > K_all:
        raise DropDecError("aggregate: member slice out of range")
    V = ld if vocab is None else vocab
    dev = logits.device
    member_ids = torch.empty(B, K, dtype=torch.int32, device=dev) if member_ids is None else member_ids
    winner = torch.empty(B, dtype=torch.int32, device=dev) if winner is None else winner
    token = torch.empty(B, dtype=torch.int32, device=dev) if token is None else token
    if out_logits is None and (want_logits or mode == "mean"):
        out_logits = torch.empty(B, V, dtype=torch.float32, device=dev)
    check(L.dd_aggregate_strided(logits.data_ptr() + first_member * ld * 4, B, K, V, ld, K_all * ld,
                                 AGG_VOTE if mode == "vote" else AGG_MEAN, _ptr(member_ids), _ptr(winner), _ptr(token),
                                 _ptr(out_logits), _stream()), "dd_aggregate")
    return member_ids, winner, token, out_logits


def token_stats(logits: torch.Tensor, vocab: int, stats: torch.Tensor, member: Optional[torch.Tensor] = None,
                slot: Optional[torch.Tensor] = None, slot_offset: int = 0, slot_per_seq: bool = False) -> torch.Tensor:
    """logits fp32 [B, K, ld] (or [B, ld]); stats fp32 [n_slots, B, 3] <- {1/max, entropy bits, varentropy} of row
    member[b] (default 0) of every sequence, written at slot index *slot + slot_offset."""
    L = _lib.load()
    _need(logits, torch.float32, "logits"); _need(stats, torch.float32, "stats")
    if logits.dim() == 2:
        logits = logits[:, None, :]
    B = logits.shape[0]
    if member is not None:
        _need(member, torch.int32, "member")
    if slot is not None:
        _need(slot, torch.int32, "slot")
    fn = L.dd_token_stats_seq if slot_per_seq else L.dd_token_stats
    check(fn(_ptr(logits), logits.stride(0), logits.stride(1), _ptr(member), B, int(vocab), _ptr(slot),
             int(slot_offset), stats.shape[0], _ptr(stats), _stream()), "dd_token_stats")
    return stats


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.load()
    _need(x, torch.bfloat16, "x"); _need(w, torch.bfloat16, "w")
    M, d = x.shape
    out = torch.empty_like(x) if out is None else out
    check(L.dd_rmsnorm(_ptr(x), _ptr(w), M, d, float(eps), _ptr(out), _stream()), "dd_rmsnorm")
    return out


def skinny_gemm(x: torch.Tensor, w: torch.Tensor, epilogue: int = EPI_NONE, residual: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [M,K] (M <= 32), w bf16 [N,K] ([2N,K] for SwiGLU) -> y [M,N] (fp32 for EPI_F32, else bf16)."""
    L = _lib.load()
    _need(x, torch.bfloat16, "x"); _need(w, torch.bfloat16, "w")
    M, K = x.shape
    N = w.shape[0] // 2 if epilogue == EPI_SWIGLU else w.shape[0]
    if w.shape[1] != K:
        raise DropDecError("skinny_gemm: x and w disagree on K")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16, device=x.device)
    if residual is not None:
        _need(residual, torch.bfloat16, "residual")
    check(L.dd_skinny_gemm(_ptr(x), _ptr(w), _ptr(residual), M, N, K, epilogue, _ptr(out), _stream()),
          "dd_skinny_gemm")
    return out


def stream_gemm_workspace(M: int, shapes: Sequence[Tuple[int, int, int]], device) -> torch.Tensor:
    """One zeroed workspace large enough for dd_stream_gemm at M rows for every (N, K, epilogue) in `shapes`."""
    L = _lib.load()
    need = max(L.dd_stream_gemm_workspace_bytes(M, n, k, e) for n, k, e in shapes)
    if need == 0:
        raise DropDecError(f"stream_gemm_workspace: unsupported shape among M={M}, {list(shapes)}")
    return torch.zeros(need, dtype=torch.uint8, device=device)


def stream_gemm(x: torch.Tensor, w: torch.Tensor, epilogue: int = EPI_NONE, residual: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tcgen05 weight-streaming GEMM: x bf16 [M,K] (M <= 256), w bf16 [N,K] ([2N,K] for SwiGLU) -> y [M,N]
    (fp32 for EPI_F32, else bf16)."""
    L = _lib.load()
    _need(x, torch.bfloat16, "x"); _need(w, torch.bfloat16, "w")
    M, K = x.shape
    N = w.shape[0] // 2 if epilogue == EPI_SWIGLU else w.shape[0]
    if w.shape[1] != K:
        raise DropDecError("stream_gemm: x and w disagree on K")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16, device=x.device)
    if residual is not None:
        _need(residual, torch.bfloat16, "residual")
    if workspace is None:
        workspace = stream_gemm_workspace(M, [(N, K, epilogue)], x.device)
    check(L.dd_stream_gemm(_ptr(x), _ptr(w), _ptr(residual), M, N, K, epilogue, _ptr(out), _ptr(workspace),
                           workspace.numel(), _stream()), "dd_stream_gemm")
    return out


def decoder_linear(x: torch.Tensor, w: torch.Tensor, epilogue: int = EPI_NONE, *, norm_w: Optional[torch.Tensor] = None,
                   eps: float = 1e-5, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                   rope: Optional[dict] = None) -> Optional[torch.Tensor]:
    """dd_decoder_linear: skinny GEMM with an optional fused RMSNorm of x (norm_w) and, for EPI_ROPE, the
    RoPE + member staging epilogue.  rope = dict(cos, sin, seq_lens, members, H, H_kv, D, q_out, k_new, v_new)."""
    L = _lib.load()
    _need(x, torch.bfloat16, "x"); _need(w, torch.bfloat16, "w")
    M, K = x.shape
    N = w.shape[0] // 2 if epilogue == EPI_SWIGLU else w.shape[0]
    if w.shape[1] != K:
        raise DropDecError("decoder_linear: x and w disagree on K")
    if norm_w is not None:
        _need(norm_w, torch.bfloat16, "norm_w")
    if residual is not None:
        _need(residual, torch.bfloat16, "residual")
    ra = None
    if epilogue == EPI_ROPE:
        if rope is None:
            raise DropDecError("decoder_linear: EPI_ROPE needs rope=dict(...)")
        _need(rope["cos"], torch.float32, "cos"); _need(rope["sin"], torch.float32, "sin")
        _need(rope["seq_lens"], torch.int32, "seq_lens")
        for k in ("q_out", "k_new", "v_new"):
            _need(rope[k], torch.bfloat16, k)
        ra = _lib.RopeArgs(_ptr(rope["cos"]), _ptr(rope["sin"]), _ptr(rope["seq_lens"]), int(rope["members"]),
                           int(rope["H"]), int(rope["H_kv"]), int(rope["D"]), _ptr(rope["q_out"]),
                           _ptr(rope["k_new"]), _ptr(rope["v_new"]))
        if rope.get("guess_winner") is not None:      # commit ahead of the lookahead pass (dd_rope_args)
            _need(rope["guess_winner"], torch.int32, "guess_winner"); _need(rope["guess_pos"], torch.int32, "guess_pos")
            _need(rope["k_cache"], torch.bfloat16, "k_cache"); _need(rope["v_cache"], torch.bfloat16, "v_cache")
            ra.guess_winner, ra.guess_pos = _ptr(rope["guess_winner"]), _ptr(rope["guess_pos"])
            ra.guess_members, ra.S_max = int(rope["guess_members"]), rope["k_cache"].shape[-2]
            ra.k_cache, ra.v_cache = _ptr(rope["k_cache"]), _ptr(rope["v_cache"])
    elif out is None:
        out = torch.empty(M, N, dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16, device=x.device)
    import ctypes
    check(L.dd_decoder_linear(_ptr(x), _ptr(norm_w), float(eps), _ptr(w), _ptr(residual), M, N, K, epilogue,
                              _ptr(out), ctypes.byref(ra) if ra is not None else None, _stream()),
          "dd_decoder_linear")
    return out


def rope_stage(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_lens: torch.Tensor, B: int, M: int,
               H: int, H_kv: int, D: int, q_out: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor) -> None:
    L = _lib.load()
    _need(qkv, torch.bfloat16, "qkv"); _need(cos, torch.float32, "cos"); _need(sin, torch.float32, "sin")
    check(L.dd_rope_stage(_ptr(qkv), _ptr(cos), _ptr(sin), _ptr(seq_lens), B, M, H, H_kv, D, _ptr(q_out),
                          _ptr(k_new), _ptr(v_new), _stream()), "dd_rope_stage")


def embed_gather(embed: torch.Tensor, token: torch.Tensor, M: int, out: torch.Tensor) -> torch.Tensor:
    L = _lib.load()
    _need(embed, torch.bfloat16, "embed"); _need(token, torch.int32, "token")
    check(L.dd_embed_gather(_ptr(embed), _ptr(token), token.shape[0], M, embed.shape[1], embed.shape[0], _ptr(out), _stream()),
          "dd_embed_gather")
    return out


def step_advance_seq(token: torch.Tensor, advance: torch.Tensor, tokens_out: Optional[torch.Tensor],
                     seq_lens: torch.Tensor, seq_steps: torch.Tensor) -> None:
    """Per-sequence progress: sequences with advance[b] != 0 record their token and move on."""
    L = _lib.load()
    for t, n in ((token, "token"), (advance, "advance"), (seq_lens, "seq_lens"), (seq_steps, "seq_steps")):
        _need(t, torch.int32, n)
    max_steps = 0 if tokens_out is None else tokens_out.shape[1]
    check(L.dd_step_advance_seq(_ptr(token), _ptr(advance), _ptr(tokens_out), max_steps, _ptr(seq_lens),
                                _ptr(seq_steps), token.shape[0], _stream()), "dd_step_advance_seq")


def step_advance(token: torch.Tensor, tokens_out: Optional[torch.Tensor], seq_lens: torch.Tensor,
                 step: torch.Tensor) -> None:
    L = _lib.load()
    max_steps = 0 if tokens_out is None else tokens_out.shape[1]
    check(L.dd_step_advance(_ptr(token), _ptr(tokens_out), max_steps, _ptr(seq_lens), _ptr(step), token.shape[0],
                            _stream()), "dd_step_advance")


def member_finish(recv: torch.Tensor, rank_stride: int, ids_off: int, stats_off: int, k_off: int, v_off: int, G: int,
                  mpr: int, seq_lens: torch.Tensor, step: torch.Tensor, member_ids: torch.Tensor, winner: torch.Tensor,
                  token: torch.Tensor, token_stats: Optional[torch.Tensor], k_cache: torch.Tensor, v_cache: torch.Tensor,
                  forced_winner: int = -1, forced_token: Optional[torch.Tensor] = None) -> None:
    """The finish step of the member-sharded engine after its per-step all-gather (dd_member_finish): vote over the
    gathered ids, publish ids / winner / token / statistics, commit the winner's K/V rows from `recv` (uint8, G
    exchange records of rank_stride bytes) into k_cache / v_cache [L, B, H_kv, S_max, D]."""
    L_ = _lib.load()
    _need(recv, torch.uint8, "recv")
    for t, n in ((seq_lens, "seq_lens"), (step, "step"), (member_ids, "member_ids"), (winner, "winner"), (token, "token")):
        _need(t, torch.int32, n)
    _need(k_cache, torch.bfloat16, "k_cache"); _need(v_cache, torch.bfloat16, "v_cache")
    nl, B, H_kv, S_max, D = k_cache.shape
    if recv.numel() < G * rank_stride:
        raise DropDecError("member_finish: recv is smaller than G records")
    n_slots = 0 if token_stats is None else token_stats.shape[0]
    check(L_.dd_member_finish(_ptr(recv), rank_stride, ids_off, stats_off, k_off, v_off, G, mpr, nl, B, H_kv, D, S_max,
                              forced_winner, _ptr(forced_token), _ptr(seq_lens), _ptr(step), n_slots, _ptr(member_ids),
                              _ptr(winner), _ptr(token), _ptr(token_stats), _ptr(k_cache), _ptr(v_cache), _stream()),
          "dd_member_finish")


# --------------------------------------------------------------------------------------------
class DecoderPassPlan:
    """Device-side plan of dd_decoder_pass for one set of pass buffers: per-layer weights, staging rows, caches and the
    caches' TMA descriptors (dd_pass_plan_fill), plus the kernel's zeroed workspace."""

    def __init__(self, layers, k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                 members_total: int, n_heads: int, num_splits: int):
        L = _lib.load()
        n_layers, B, H_kv, S_max, D = k_cache.shape
        for t in (k_new, v_new, k_cache, v_cache):
            _need(t, torch.bfloat16, "cache / staging")
        host = (C.c_uint8 * L.dd_pass_plan_bytes(n_layers))()
        for l, lw in enumerate(layers):
            check(L.dd_pass_plan_fill(host, l, _ptr(lw["ln1"]), _ptr(lw["wqkv"]), _ptr(lw["wo"]), _ptr(lw["ln2"]),
                                      _ptr(lw["wgu"]), _ptr(lw["wd"]), _ptr(k_new[l]), _ptr(v_new[l]), _ptr(k_cache[l]),
                                      _ptr(v_cache[l]), B, H_kv, S_max, D), "dd_pass_plan_fill")
        self.plan = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(k_cache.device)
        self.n_layers, self.B, self.H_kv, self.S_max, self.D = n_layers, B, H_kv, S_max, D
        self.members_total, self.H, self.num_splits = members_total, n_heads, num_splits
        need = L.dd_pass_workspace_bytes(B, members_total, n_heads, D, num_splits)
        self.workspace = torch.zeros(need, dtype=torch.uint8, device=k_cache.device)
        self._keep = (layers, k_new, v_new, k_cache, v_cache)


def decoder_pass(plan: DecoderPassPlan, *, d_model: int, d_ff: int, vocab: int, eps: float, scale: float,
                 embed: torch.Tensor, tok: torch.Tensor, tok_div: int, pos: torch.Tensor, pos_div: int,
                 cos: torch.Tensor, sin: torch.Tensor, x: torch.Tensor, x2: torch.Tensor, act: torch.Tensor,
                 q: torch.Tensor, attn: torch.Tensor, final_norm: torch.Tensor, lm_head: Optional[torch.Tensor],
                 logits: Optional[torch.Tensor], windows, vis_start: Optional[torch.Tensor] = None,
                 bitmask: Optional[torch.Tensor] = None, n_vis: int = 0, xn_out: Optional[torch.Tensor] = None,
                 guess_winner: Optional[torch.Tensor] = None, guess_pos: Optional[torch.Tensor] = None,
                 prefetch_gate_up_bytes: int = 0, timeline: Optional[torch.Tensor] = None) -> None:
    """dd_decoder_pass: the whole single-token forward of rows = x.shape[0] <= 16 rows in one persistent kernel.
    windows: [(first_member, members, seq_lens int32 [B]), ...] (1 or 2, tiling the members of a sequence)."""
    L = _lib.load()
    a = _lib.PassArgs()
    rows = x.shape[0]
    a.plan, a.n_layers, a.rows, a.members_total = _ptr(plan.plan), plan.n_layers, rows, plan.members_total
    a.d_model, a.d_ff, a.H, a.H_kv, a.D, a.S_max = d_model, d_ff, plan.H, plan.H_kv, plan.D, plan.S_max
    a.vocab, a.vocab_padded = vocab, (lm_head.shape[0] if lm_head is not None else 0)
    a.eps, a.scale = float(eps), float(scale)
    for t, name in ((embed, "embed"), (x, "x"), (x2, "x2"), (act, "act"), (q, "q"), (attn, "attn"),
                    (final_norm, "final_norm")):
        _need(t, torch.bfloat16, name)
    _need(tok, torch.int32, "tok"); _need(pos, torch.int32, "pos")
    _need(cos, torch.float32, "cos"); _need(sin, torch.float32, "sin")
    a.embed, a.tok, a.tok_div, a.pos, a.pos_div = _ptr(embed), _ptr(tok), tok_div, _ptr(pos), pos_div
    a.cos_tab, a.sin_tab = _ptr(cos), _ptr(sin)
    a.x, a.x2, a.act, a.q, a.attn, a.xn_out = _ptr(x), _ptr(x2), _ptr(act), _ptr(q), _ptr(attn), _ptr(xn_out)
    a.final_norm, a.lm_head, a.logits = _ptr(final_norm), _ptr(lm_head), _ptr(logits)
    if logits is not None:
        _need(logits, torch.float32, "logits"); _need(lm_head, torch.bfloat16, "lm_head")
    if bitmask is not None:
        _need(bitmask, torch.int32, "bitmask"); _need(vis_start, torch.int32, "vis_start")
        a.mask_words = bitmask.shape[-1]
    a.vis_start, a.bitmask, a.n_vis, a.num_splits = _ptr(vis_start), _ptr(bitmask), n_vis, plan.num_splits
    a.n_windows = len(windows)
    for i, (first, members, lens) in enumerate(windows):
        _need(lens, torch.int32, "seq_lens")
        a.win_first[i], a.win_members[i], a.win_seq_lens[i] = first, members, _ptr(lens)
    a.guess_winner, a.guess_pos = _ptr(guess_winner), _ptr(guess_pos)
    a.prefetch_gate_up_bytes = int(prefetch_gate_up_bytes)
    a.workspace, a.workspace_bytes = _ptr(plan.workspace), plan.workspace.numel()
    if timeline is not None:
        _need(timeline, torch.int64, "timeline")
        if timeline.numel() < 2 * (5 * plan.n_layers + 2):
            raise DropDecError("decoder_pass: timeline needs 2 * (5 * n_layers + 2) int64 slots")
        a.timeline = _ptr(timeline)
    check(L.dd_decoder_pass(C.byref(a), _stream()), "dd_decoder_pass")
