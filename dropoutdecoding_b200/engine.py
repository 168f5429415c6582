"""Decode engine of the Dropout-Decoding hot path on one B200.

What the reference does per generated token (models/llava.py:218-388): a
preliminary unmasked forward, then K masked forwards run sequentially, each on a deep copy of the
whole KV cache, with >= 3K host syncs.  Here one ensemble step is two batched passes over a
single shared cache, with no host sync at all, replayed as one CUDA graph:

    pass A  (rows = B)      embed -> L x [RMSNorm, QKV, RoPE, attention(M=1), O, RMSNorm, SwiGLU, down]
                            -> lm_head -> argmax = y_init                       (dd_aggregate, K=1)
    masks                   relevant set + K uncertainty-guided bitmasks         (dd_sample_masks)
    pass B  (rows = B*K)    same layers; attention reads the prefix K/V ONCE for all K members and
                            applies each member's bitmask in registers; every member keeps its own
                            new-token K/V row per layer in a staging buffer       (dd_decode_attention)
    vote / mean             member argmax ids, winner, next token                 (dd_aggregate)
    commit                  winner's staged K/V rows -> cache row seq_len, seq_len += 1

Weight GEMMs: <= 16 rows dd_skinny_gemm (mma.sync weight stream), 17..256 rows dd_stream_gemm (tcgen05 / TMEM / TMA,
one-wave tiles, cluster split-K); only passes of more than 256 rows (none of the BASELINE configurations) and the
prompt prefill, which is not part of the hot path, use library GEMMs.  All tensors are bf16 with fp32 accumulation; rounding points follow the HF
Llama decoder so that the CPU oracle's 'bf16-faithful' mode is comparable token for token.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import ops
from .schedules import SpeculativeSchedules


@dataclass
class DecoderConfig:
    n_layers: int
    d_model: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    d_ff: int
    vocab: int
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0

    @staticmethod
    def llava15_7b() -> "DecoderConfig":
        return DecoderConfig(32, 4096, 32, 32, 128, 11008, 32064, 1e-5, 10000.0)

    @staticmethod
    def llava_next_mistral_7b() -> "DecoderConfig":
        return DecoderConfig(32, 4096, 32, 8, 128, 14336, 32064, 1e-5, 1000000.0)

    @staticmethod
    def tiny() -> "DecoderConfig":
        return DecoderConfig(2, 256, 4, 4, 64, 512, 32064, 1e-5, 10000.0)


def rope_table(head_dim: int, theta: float, max_pos: int):
    """cos/sin [max_pos, head_dim/2], computed on the CPU in fp32 and rounded to bf16 like
    transformers' `cos.to(x.dtype)`; returned as fp32 holding bf16 values."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().to(torch.bfloat16).float(), ang.sin().to(torch.bfloat16).float()


class DecoderWeights:
    """bf16 device tensors; QKV and gate/up are stored concatenated ([q;k;v], [gate;up]) so that one
    weight stream feeds one fused GEMM."""

    def __init__(self, cfg: DecoderConfig, embed, layers: List[Dict[str, torch.Tensor]], norm, lm_head):
        self.cfg, self.embed, self.layers, self.norm = cfg, embed, layers, norm
        # the weight-streaming GEMM wants N % 16 == 0: pad the head with zero rows (InstructBLIP: V = 32001);
        # `lm_head` is the exact [V, d] view, `lm_head_padded` what the GEMM streams
        V = lm_head.shape[0]
        self.vocab_padded = -(-V // 16) * 16
        if self.vocab_padded != V:
            pad = torch.zeros(self.vocab_padded, lm_head.shape[1], dtype=lm_head.dtype, device=lm_head.device)
            pad[:V] = lm_head
            self.lm_head_padded, self.lm_head = pad, pad[:V]
        else:
            self.lm_head_padded = self.lm_head = lm_head

    @staticmethod
    def from_named(cfg: DecoderConfig, w: Dict[str, torch.Tensor], device) -> "DecoderWeights":
        """w uses the oracle's naming (embed, l{i}.ln1/wq/wk/wv/wo/ln2/wg/wu/wd, norm, lm_head)."""
        cv = lambda t: t.to(device=device, dtype=torch.bfloat16).contiguous()
        layers = []
        for l in range(cfg.n_layers):
            p = f"l{l}."
            layers.append(dict(
                ln1=cv(w[p + "ln1"]), ln2=cv(w[p + "ln2"]),
                wqkv=cv(torch.cat([w[p + "wq"], w[p + "wk"], w[p + "wv"]], 0)),
                wo=cv(w[p + "wo"]), wgu=cv(torch.cat([w[p + "wg"], w[p + "wu"]], 0)), wd=cv(w[p + "wd"])))
        return DecoderWeights(cfg, cv(w["embed"]), layers, cv(w["norm"]), cv(w["lm_head"]))

    @staticmethod
    def random(cfg: DecoderConfig, device, seed: int = 0, std: float = 0.02, head_std: float = 0.02
               ) -> "DecoderWeights":
        """Seeded random init directly on the device (benchmarks: no checkpoint is available)."""
        g = torch.Generator(device=device).manual_seed(seed)

        def n(*shape, s=std):
            return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(torch.bfloat16)

        ones = lambda k: torch.ones(k, dtype=torch.bfloat16, device=device)
        qkv_rows = (cfg.n_heads + 2 * cfg.n_kv_heads) * cfg.head_dim
        layers = [dict(ln1=ones(cfg.d_model), ln2=ones(cfg.d_model), wqkv=n(qkv_rows, cfg.d_model),
                       wo=n(cfg.d_model, cfg.n_heads * cfg.head_dim), wgu=n(2 * cfg.d_ff, cfg.d_model),
                       wd=n(cfg.d_model, cfg.d_ff)) for _ in range(cfg.n_layers)]
        return DecoderWeights(cfg, n(cfg.vocab, cfg.d_model), layers, ones(cfg.d_model),
                              n(cfg.vocab, cfg.d_model, s=head_std))

    def nbytes(self) -> int:
        tot = self.embed.numel() + self.norm.numel() + self.lm_head.numel()
        for l in self.layers:
            tot += sum(t.numel() for t in l.values())
        return 2 * tot

    def stream_bytes_per_pass(self) -> int:
        """Weight bytes one single-token pass has to read (embedding rows excluded)."""
        return self.nbytes() - 2 * self.embed.numel()


@dataclass
class _PassBuffers:
    M: int
    x: torch.Tensor
    x2: torch.Tensor
    xn: torch.Tensor
    qkv: torch.Tensor
    q: torch.Tensor
    k_new: torch.Tensor     # [L, B, M, H_kv, D]
    v_new: torch.Tensor
    attn: torch.Tensor
    act: torch.Tensor
    logits: torch.Tensor    # fp32 [B, M, V]
    ws: torch.Tensor
    splits: int
    plan: Optional[object] = None      # ops.DecoderPassPlan of the persistent pass kernel (built at first use)
    ws2: Optional[torch.Tensor] = None  # workspace of the two-window attention launch of the lookahead pass (first use)


class DropoutDecodeEngine(SpeculativeSchedules):
    """Static-shape engine for a batch of B sequences with the same number of visual tokens."""

    def __init__(self, weights: DecoderWeights, batch: int, n_vis: int, max_seq: int,
                 probs: Sequence[float] = (0.3, 0.5, 0.7), mode: str = "cumulative", keep_relevant: bool = True,
                 topk: int = 5, aggregate: str = "vote", seed: int = 24, max_steps: int = 1024,
                 round_logits: bool = True, est_seq: Optional[int] = None, vote_on: str = "logits",
                 seq_base: int = 0, gemm_backend: Optional[str] = None, speculative: Optional[str] = None,
                 attn_splits: Optional[int] = None):
        cfg = weights.cfg
        if cfg.d_model % 256 or cfg.d_ff % 256:
            raise ops.DropDecError("engine needs d_model % 256 == 0 and d_ff % 256 == 0")
        if vote_on not in ("logits", "hidden"):
            raise ValueError(vote_on)
        self.w, self.cfg = weights, cfg
        self.B, self.K, self.n_vis, self.S_max = batch, len(probs), n_vis, max_seq
        self.probs, self.mode, self.keep_relevant = [float(p) for p in probs], mode, keep_relevant
        self.topk, self.aggregate, self.seed, self.round_logits = topk, aggregate, seed, round_logits
        self.fuse_norm = os.environ.get("DD_FUSE_NORM", "0") == "1"
        # L2 prefetch riding on the latency-bound kernels of a <= 16-row pass (dd_hint_l2_prefetch): MB of the
        # following GEMM's weights hinted to {RMSNorm before QKV, attention (after all of o_proj), RMSNorm before gate/up};
        # off by default (DD_PF_MB="24,48,24" turns it on): measured neutral to -3 % on the launch chain -- profiles/r02/l2_prefetch_sweep.jsonl
        pf = [int(float(v) * (1 << 20)) for v in os.environ.get("DD_PF_MB", "0").split(",")]
        self.l2_prefetch = tuple(pf) if len(pf) == 3 and any(pf) else None
        # <= 16-row passes as ONE persistent kernel (dd_decoder_pass): DD_PASS=1.  Bit-identical to the launch chain but
        # measured slower on B200 so far (4.9 vs 3.4 ms per 7B pass, profiles/README.md), so the chain is the default.
        # DD_PASS_PF_MB: MB of o_proj + gate/up weights prefetched into L2 by the idle warps of its attention phase
        self.pass_kernel = os.environ.get("DD_PASS", "0") == "1"
        self.pass_pf_bytes = int(float(os.environ.get("DD_PASS_PF_MB", "0")) * (1 << 20))
        self.prefill_graph = os.environ.get("DD_PREFILL_GRAPH", "1") == "1"
        self.prefill_native = os.environ.get("DD_PREFILL_NATIVE", "1") == "1"
        self._prefill_graphs: dict = {}
        # Weight GEMMs by row count (rows = sequences x members of a pass); measured on B200 (profiles/README.md 4):
        #   rows <= 16        dd_skinny_gemm / dd_decoder_linear (mma.sync, RoPE fused into the QKV epilogue)
        #   16 < rows <= 256  dd_stream_gemm (tcgen05; profiles/r02/stream_gemm_v4.jsonl: 0.85-1.0x cuBLAS up to 128 rows,
        #                     0.5-0.9x at 192-256 rows where the GEMM sits on the HBM / tensor ridge)
        #   rows > 256        library GEMM
        #   DD_GEMM=library (gemm_backend="library") sends every pass of more than 32 rows to the library instead;
        #   DD_GEMM=stream forces dd_stream_gemm for every pass of up to 256 rows (also <= 16 rows)
        self.gemm_backend = gemm_backend or os.environ.get("DD_GEMM", "auto")
        if self.gemm_backend not in ("auto", "stream", "library", "skinny"):
            raise ValueError(self.gemm_backend)
        # programmatic dependent launch of the pass kernels: None = per pass (on for dd_stream_gemm passes); DD_PDL=0/1 pins it
        self.chain_launch = {"0": 0, "1": 1, "2": 2, "3": 3}.get(os.environ.get("DD_PDL", ""), None)
        # mode of the <= 16-row passes (skinny GEMMs): DD_PDL_SKINNY = 0 | 1 | 2 | 3.  3 (default): only the SMALL kernels
        # (RMSNorm, attention, vote ...) are programmatic dependents -- their launch latency overlaps the drain of the GEMM
        # in front of them, while no GEMM's CTAs queue up behind anything: 3.232 -> 3.208 ms per token, 2.993 -> 2.966 ms
        # per 4-row pass in back-to-back pairs; every kernel (1): 3.60 vs 3.45, only the GEMMs (2): 3.66 (profiles/README.md)
        self.chain_launch_skinny = int(os.environ.get("DD_PDL_SKINNY", "3"))
        self.attn_splits = attn_splits  # None: chosen per pass from the measured sweep (_alloc_pass); tests pin it
        self.seq_base = int(seq_base)   # global index of sequence 0 (batch sharded over GPUs)
        self.vote_on = vote_on          # "hidden": InstructBLIP votes on the argmax of the final hidden state
        self.Vp = weights.vocab_padded
        self.dev = weights.embed.device
        dev, B, L = self.dev, batch, cfg.n_layers
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        bf = dict(dtype=torch.bfloat16, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.k_cache = torch.zeros(L, B, Hkv, max_seq, D, **bf)
        self.v_cache = torch.zeros(L, B, Hkv, max_seq, D, **bf)
        self.seq_lens = torch.zeros(B, **i32)
        self.vis_start = torch.zeros(B, **i32)
        self.step_ctr = torch.zeros(1, **i32)
        self.max_steps = max_steps
        self.tokens_out = torch.zeros(B, max_steps, **i32)
        # a13 bookkeeping per generated token: {1/max(logits), entropy bits, varentropy} of the returned logits
        self.token_stats = torch.zeros(max_steps, B, 3, dtype=torch.float32, device=dev)
        self.cur_token = torch.zeros(B, **i32)
        self.next_token = torch.zeros(B, **i32)
        self.y_init = torch.zeros(B, **i32)
        self.winner = torch.zeros(B, **i32)
        self.winner_a = torch.zeros(B, **i32)
        self.member_ids = torch.zeros(B, self.K, **i32)
        self.member_ids_a = torch.zeros(B, 1, **i32)
        self.words = ops.mask_words(n_vis)
        self.bitmask = torch.zeros(B, self.K, self.words, **i32)
        self.n_masked = torch.zeros(B, self.K, **i32)
        self.out_logits = torch.zeros(B, cfg.vocab, dtype=torch.float32, device=dev)   # mean mode / forward() API
        self.uniforms: Optional[torch.Tensor] = None      # parity mode: [steps, B, K, n_vis]
        self._u_step: Optional[torch.Tensor] = None
        self.epi = torch.zeros(B, n_vis, dtype=torch.float32, device=dev)
        self.alea = torch.zeros_like(self.epi)
        self.var = torch.zeros_like(self.epi)
        self.topk_ids = torch.zeros(B, n_vis, topk, **i32)
        self._unc_overflow = torch.zeros(1, **i32)      # rows whose top-k candidate list overflowed in stage (1)
        self._unc_ws = None                             # scratch of stage (1) (logits stash), allocated at first use
        cos, sin = rope_table(D, cfg.rope_theta, max_seq + 1)
        self.cos, self.sin = cos.to(dev), sin.to(dev)
        self.scale = 1.0 / math.sqrt(D)
        est = est_seq or max_seq
        self.pa = self._alloc_pass(1, est)
        self.pb = self._alloc_pass(self.K, est)
        if vote_on == "hidden":
            self._hid32 = torch.zeros(B, self.K, cfg.d_model, dtype=torch.float32, device=dev)
            self._hid_tok = torch.zeros(B, **i32)
            self._win_h = torch.zeros(B, cfg.d_model, **bf)
            self._arange = torch.arange(B, device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.launch_count = 0           # kernels of this library launched by decode() so far (bench.py: gpu_launches)
        self._init_schedules(speculative, est)
        self._sg_ws = None
        max_rows = B * (self.K + 1 if (self.spec_active or self.spec_batch) else self.K)
        if self.batch_la:
            max_rows = B * (2 * self.K + 1)
        if any(self._use_stream(r) for r in (B, max_rows)):
            d, dff, qkv_rows = cfg.d_model, cfg.d_ff, (H + 2 * Hkv) * D
            shapes = [(qkv_rows, d, ops.EPI_NONE), (d, H * D, ops.EPI_RESIDUAL), (dff, d, ops.EPI_SWIGLU),
                      (d, dff, ops.EPI_RESIDUAL), (self.Vp, d, ops.EPI_F32)]
            self._sg_ws = ops.stream_gemm_workspace(min(max_rows, 256), shapes, dev)

    # ------------------------------------------------------------------------------------
    def _alloc_pass(self, M: int, est_seq: int) -> _PassBuffers:
        cfg, B, dev = self.cfg, self.B, self.dev
        rows = B * M
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        bf = dict(dtype=torch.bfloat16, device=dev)
        # split-KV factor (measured sweeps, profiles/README.md 0 and 2): as many splits as keep the grid in ONE wave of one
        # CTA per SM (floor, not round: 4 x 32 = 128 CTAs beat 3 x 32 and 5 x 32 for one 7B sequence, 3.22 vs 3.36 ms per
        # token), never fewer than ~2 key tiles per CTA; (sequence, kv head) pairs that already fill the SMs stay unsplit
        # (B >= 5 at 32 kv heads), and a long context on fewer pairs than SMs is cut at least in two so that two CTAs
        # share an SM
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        pairs, tiles = B * Hkv, max(1, -(-est_seq // 64))
        splits = max(1, min(sms // pairs, max(1, tiles // 2), 32))
        if splits == 1 and pairs < sms and tiles >= 16:
            splits = 2
        if self.attn_splits is not None:
            splits = int(self.attn_splits)
        elif os.environ.get("DD_ATTN_SPLITS"):          # experiments (tools/stepbench.py sweeps)
            splits = int(os.environ["DD_ATTN_SPLITS"])
        return _PassBuffers(
            M=M, x=torch.empty(rows, cfg.d_model, **bf), x2=torch.empty(rows, cfg.d_model, **bf),
            xn=torch.empty(rows, cfg.d_model, **bf), qkv=torch.empty(rows, (H + 2 * Hkv) * D, **bf),
            q=torch.empty(B, M, H, D, **bf), k_new=torch.empty(cfg.n_layers, B, M, Hkv, D, **bf),
            v_new=torch.empty(cfg.n_layers, B, M, Hkv, D, **bf), attn=torch.zeros(B, M, H, D, **bf),
            act=torch.empty(rows, cfg.d_ff, **bf), logits=torch.empty(B, M, self.Vp, dtype=torch.float32, device=dev),
            ws=ops.attention_workspace(B, M, H, D, splits, dev), splits=splits)

    # weight GEMM with the HF rounding points; <= 32 rows: our weight-streaming kernel
    def _use_stream(self, rows: int) -> bool:
        if self.gemm_backend == "stream":
            return rows <= 256
        if self.gemm_backend == "auto":
            return 16 < rows <= 256
        if self.gemm_backend == "library":
            return 16 < rows <= 32
        return False

    def _set_chain(self, rows: int) -> None:
        """Programmatic dependent launch for the kernels of the pass that follows: it pays where dd_stream_gemm can
        prefetch its ring of weight tiles under the previous kernel's tail (7B, 56 rows: 5.01 -> 4.57 ms per pass) and
        costs 3-4 % on the register-streamed skinny GEMMs of a 4-row pass."""
        if self.chain_launch is not None:
            ops.set_chain_launch(self.chain_launch)
        else:
            ops.set_chain_launch(1 if self._use_stream(rows) else self.chain_launch_skinny)

    def _linear(self, x, w, epi, out, residual=None):
        self.launches_per_step += 1
        rows = x.shape[0]
        if self._use_stream(rows):
            return ops.stream_gemm(x, w, epi, residual=residual, out=out, workspace=self._sg_ws)
        if rows <= 32:
            return ops.skinny_gemm(x, w, epi, residual=residual, out=out)
        if epi == ops.EPI_NONE:
            return torch.matmul(x, w.T, out=out)
        if epi == ops.EPI_RESIDUAL:
            return torch.add(F.linear(x, w), residual, out=out)
        if epi == ops.EPI_F32:
            out.copy_(F.linear(x, w))
            return out
        gu = F.linear(x, w)
        n = gu.shape[1] // 2
        return torch.mul(F.silu(gu[:, :n]), gu[:, n:], out=out)

    def _attention(self, pb: _PassBuffers, l: int, bitmask, seq_active) -> None:
        """dd_decode_attention of layer l; a pass with more than 32 query rows per kv head (LLaVA-NeXT: 9 contexts x
        4 query heads) is served in member groups, each reading the K/V once."""
        G = self.cfg.n_heads // self.cfg.n_kv_heads
        per = max(1, 32 // G)
        for m0 in range(0, pb.M, per):
            ops.decode_attention(pb.q, pb.k_new[l], pb.v_new[l], self.k_cache[l], self.v_cache[l], self.seq_lens,
                                 self.vis_start, bitmask, self.n_vis, self.scale, pb.splits, pb.attn, pb.ws,
                                 seq_active=seq_active, member_offset=m0, members=min(per, pb.M - m0))
            if m0:
                self.launches_per_step += 1 + (pb.splits == 1)

    def _forward(self, pb: _PassBuffers, bitmask: Optional[torch.Tensor], head: bool = True,
                 seq_active: Optional[torch.Tensor] = None) -> torch.Tensor:
        cfg, B, M = self.cfg, self.B, pb.M
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        rows = B * M
        self._set_chain(rows)
        if self._pass_kernel_ok(pb, rows, M) and head and seq_active is None:
            self._run_pass_kernel(pb, tok=self.cur_token, tok_div=M, pos=self.seq_lens, pos_div=M, bitmask=bitmask,
                                  windows=[(0, M, self.seq_lens)])
            return pb.logits
        ops.embed_gather(self.w.embed, self.cur_token, M, pb.x)
        x, x2 = pb.x, pb.x2
        if rows <= 32 and not self._use_stream(rows):
            return self._forward_fused(pb, bitmask, head, seq_active)
        for l, lw in enumerate(self.w.layers):
            ops.rmsnorm(x, lw["ln1"], cfg.rms_eps, pb.xn)
            self._linear(pb.xn, lw["wqkv"], ops.EPI_NONE, pb.qkv)
            ops.rope_stage(pb.qkv, self.cos, self.sin, self.seq_lens, B, M, H, Hkv, D, pb.q, pb.k_new[l], pb.v_new[l])
            self._attention(pb, l, bitmask, seq_active)
            if rows <= 32 and self.gemm_backend == "auto":      # 33 MB o_proj: the skinny kernel is as fast up to 32 rows
                ops.skinny_gemm(pb.attn.view(rows, H * D), lw["wo"], ops.EPI_RESIDUAL, residual=x, out=x2)
                self.launches_per_step += 1
            else:
                self._linear(pb.attn.view(rows, H * D), lw["wo"], ops.EPI_RESIDUAL, x2, residual=x)
            ops.rmsnorm(x2, lw["ln2"], cfg.rms_eps, pb.xn)
            self._linear(pb.xn, lw["wgu"], ops.EPI_SWIGLU, pb.act)
            self._linear(pb.act, lw["wd"], ops.EPI_RESIDUAL, x, residual=x2)
            self.launches_per_step += 4 + (pb.splits == 1)     # 2 rmsnorm, rope_stage, attention (+ merge when unsplit)
        ops.rmsnorm(x, self.w.norm, cfg.rms_eps, pb.xn)
        self.launches_per_step += 2         # embed gather, final norm
        if head:
            self._linear(pb.xn, self.w.lm_head_padded, ops.EPI_F32, pb.logits.view(rows, self.Vp))
        return pb.logits

    def _pass_kernel_ok(self, pb: _PassBuffers, rows: int, window_members: int) -> bool:
        """dd_decoder_pass serves passes of <= 16 rows whose attention is split (small batches) and whose member windows
        fit one 16-row query tile."""
        cfg = self.cfg
        return (self.pass_kernel and not self.fuse_norm and rows <= 16 and not self._use_stream(rows) and pb.splits >= 2
                and window_members * (cfg.n_heads // cfg.n_kv_heads) <= 16 and cfg.head_dim in (64, 128))

    def _run_pass_kernel(self, pb: _PassBuffers, *, tok, tok_div, pos, pos_div, bitmask, windows, guess_winner=None,
                         guess_pos=None) -> None:
        cfg = self.cfg
        if pb.plan is None:
            pb.plan = ops.DecoderPassPlan(self.w.layers, pb.k_new, pb.v_new, self.k_cache, self.v_cache, pb.M,
                                          cfg.n_heads, pb.splits)
        rows = self.B * pb.M
        ops.decoder_pass(pb.plan, d_model=cfg.d_model, d_ff=cfg.d_ff, vocab=cfg.vocab, eps=cfg.rms_eps, scale=self.scale,
                         embed=self.w.embed, tok=tok, tok_div=tok_div, pos=pos, pos_div=pos_div, cos=self.cos,
                         sin=self.sin, x=pb.x, x2=pb.x2, act=pb.act, q=pb.q, attn=pb.attn, final_norm=self.w.norm,
                         lm_head=self.w.lm_head_padded, logits=pb.logits.view(rows, self.Vp), windows=windows,
                         vis_start=self.vis_start if bitmask is not None else None, bitmask=bitmask, n_vis=self.n_vis,
                         xn_out=pb.xn, guess_winner=guess_winner, guess_pos=guess_pos,
                         prefetch_gate_up_bytes=self.pass_pf_bytes, timeline=getattr(self, "pass_timeline", None))
        self.launches_per_step += 1

    def _forward_fused(self, pb: _PassBuffers, bitmask: Optional[torch.Tensor], head: bool,
                       seq_active: Optional[torch.Tensor] = None) -> torch.Tensor:
        """<= 32 rows: RoPE + member staging ride in the epilogue of the QKV projection (dd_decoder_linear,
        DD_EPI_ROPE) -- 7 launches per layer.  With fuse_norm the two RMSNorms also move into the prologue of
        the GEMM that consumes them (5 launches per layer); measured slower on B200 because every one of the
        ~1400 CTAs of a GEMM then re-normalises the same rows, so it is off by default (DD_FUSE_NORM=1)."""
        cfg, B, M = self.cfg, self.B, pb.M
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        rows = B * M
        x, x2 = pb.x, pb.x2
        fuse = self.fuse_norm
        pf = self.l2_prefetch
        for l, lw in enumerate(self.w.layers):
            rope = dict(cos=self.cos, sin=self.sin, seq_lens=self.seq_lens, members=M, H=H, H_kv=Hkv, D=D, q_out=pb.q,
                        k_new=pb.k_new[l], v_new=pb.v_new[l])
            if fuse:
                ops.decoder_linear(x, lw["wqkv"], ops.EPI_ROPE, norm_w=lw["ln1"], eps=cfg.rms_eps, rope=rope)
            else:
                if pf:
                    ops.hint_l2_prefetch(lw["wqkv"], 0, pf[0])
                ops.rmsnorm(x, lw["ln1"], cfg.rms_eps, pb.xn)
                ops.decoder_linear(pb.xn, lw["wqkv"], ops.EPI_ROPE, rope=rope)
            if pf:
                ops.hint_l2_prefetch(lw["wo"], 0, 1 << 40, lw["wgu"], 0, pf[1])
            self._attention(pb, l, bitmask, seq_active)
            ops.decoder_linear(pb.attn.view(rows, H * D), lw["wo"], ops.EPI_RESIDUAL, residual=x, out=x2)
            if fuse:
                ops.decoder_linear(x2, lw["wgu"], ops.EPI_SWIGLU, norm_w=lw["ln2"], eps=cfg.rms_eps, out=pb.act)
            else:
                if pf:
                    ops.hint_l2_prefetch(lw["wgu"], pf[1], pf[2])
                ops.rmsnorm(x2, lw["ln2"], cfg.rms_eps, pb.xn)
                ops.decoder_linear(pb.xn, lw["wgu"], ops.EPI_SWIGLU, out=pb.act)
            ops.decoder_linear(pb.act, lw["wd"], ops.EPI_RESIDUAL, residual=x2, out=x)
            self.launches_per_step += (5 if fuse else 7) + (pb.splits == 1)
        self.launches_per_step += 1         # embed gather
        if head and fuse:
            ops.decoder_linear(x, self.w.lm_head_padded, ops.EPI_F32, norm_w=self.w.norm, eps=cfg.rms_eps,
                               out=pb.logits.view(rows, self.Vp))
            self.launches_per_step += 1
        else:
            if pf and head:
                ops.hint_l2_prefetch(self.w.lm_head_padded, 0, pf[0])
            ops.rmsnorm(x, self.w.norm, cfg.rms_eps, pb.xn)     # (InstructBLIP votes on this final-norm hidden state)
            self.launches_per_step += 1
            if head:
                ops.decoder_linear(pb.xn, self.w.lm_head_padded, ops.EPI_F32, out=pb.logits.view(rows, self.Vp))
                self.launches_per_step += 1
        return pb.logits

    # ------------------------------------------------------------------------------------
    def _step_uniforms(self, ahead: int = 0) -> Optional[torch.Tensor]:
        if self.uniforms is None:
            return None
        # parity mode: this step's pre-drawn numbers (device-side gather keeps the graph static)
        buf = self._u_step if ahead == 0 else self._u_step1
        buf.copy_(self.uniforms.index_select(0, (self.step_ctr.long() + ahead).clamp(max=self.uniforms.shape[0] - 1))[0])
        return buf

    def _sample(self, u, keep_relevant: bool, out, n_masked, per_seq: bool = False, step=None, y=None) -> None:
        if step is None:
            step = self.seq_steps if per_seq else self.step_ctr
        y = self.y_init if y is None else y
        ops.sample_masks(self.epi, self.topk_ids if keep_relevant else None, y if keep_relevant else None,
                         self.probs, mode=self.mode, keep_relevant=keep_relevant, uniforms=u, seed=self.seed, offset=0,
                         step=step, out=out, n_masked=n_masked, seq_base=self.seq_base, step_per_seq=per_seq)
        self.launches_per_step += 1

    def step(self) -> None:
        """One ensemble step for every sequence; asynchronous, no host sync, graph-capturable."""
        if self.spec_batch:
            return self._spec_batch_la_step() if self.batch_la else self._spec_batch_step()
        return self._two_pass_step()

    def _two_pass_step(self) -> None:
        """The reference's own order: preliminary pass, masks, K masked contexts, vote, commit."""
        self.launches_per_step = 0
        logits_a = self._forward(self.pa, None)                                    # models/llava.py:294-305
        ops.aggregate(logits_a, "vote", member_ids=self.member_ids_a, winner=self.winner_a, token=self.y_init,
                      vocab=self.cfg.vocab)
        self.launches_per_step += 1
        self._sample(self._step_uniforms(), self.keep_relevant, self.bitmask, self.n_masked)      # :342-346
        self._masked_pass_and_commit()

    def _masked_pass_and_commit(self) -> None:
        """K masked contexts with self.bitmask, aggregation, commit of the winner's K/V rows, step bookkeeping."""
        if self.vote_on == "logits":
            logits_b = self._forward(self.pb, self.bitmask)                        # :350-359
            ops.aggregate(logits_b, self.aggregate, member_ids=self.member_ids, winner=self.winner,   # :361
                          token=self.next_token, out_logits=self.out_logits if self.aggregate == "mean" else None,
                          vocab=self.cfg.vocab)
        else:
            # InstructBLIP (models/instructblip.py:125-140): members run the decoder WITHOUT the head, the
            # vote is taken on argmax of the final hidden state [.., d], lm_head is applied to the winner only
            self._forward(self.pb, self.bitmask, head=False)
            B, K, d = self.B, self.K, self.cfg.d_model
            self._hid32.copy_(self.pb.xn.view(B, K, d))
            ops.aggregate(self._hid32, "vote", member_ids=self.member_ids, winner=self.winner, token=self._hid_tok)
            self._win_h.copy_(self.pb.xn.view(B, K, d)[self._arange, self.winner.long()])
            self._linear(self._win_h, self.w.lm_head_padded, ops.EPI_F32, self.pa.logits.view(B, self.Vp))
            ops.aggregate(self.pa.logits, "vote", member_ids=self.member_ids_a, winner=self.winner_a,
                          token=self.next_token, vocab=self.cfg.vocab)
            self.launches_per_step += 2
        # a13: 1/max, entropy, varentropy of the logits this step returns (llavanext.py:584, instructblip.py:141-144)
        if self.vote_on == "hidden":
            ops.token_stats(self.pa.logits, self.cfg.vocab, self.token_stats, slot=self.step_ctr, slot_offset=1)
        elif self.aggregate == "mean":
            ops.token_stats(self.out_logits, self.cfg.vocab, self.token_stats, slot=self.step_ctr, slot_offset=1)
        else:
            ops.token_stats(self.pb.logits, self.cfg.vocab, self.token_stats, member=self.winner, slot=self.step_ctr,
                            slot_offset=1)
        ops.commit_winner(self.pb.k_new, self.pb.v_new, self.winner, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
        self.cur_token.copy_(self.next_token)
        self.launches_per_step += 4

    def capture(self) -> None:
        """Warm up once eagerly (restoring the state), then capture one step into a CUDA graph (two-pass path) or
        its three pieces -- speculative pass, finish, fallback -- into three graphs (speculative path)."""
        snap = self._snapshot()
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            if self.spec_active:
                self._spec_head(); self._spec_fin()
                self._restore(snap)
                self._spec_head(); self._spec_fallback()
                if self.lookahead:
                    self._restore(snap)
                    self._spec_head(); self._spec_chain(); self._spec_fin_chain()
            else:
                self.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        if self.spec_active:
            graphs = {}
            pieces = [("head", self._spec_head), ("fin", self._spec_fin), ("fallback", self._spec_fallback)]
            if self.lookahead:
                pieces += [("chain", self._spec_chain), ("fin_chain", self._spec_fin_chain)]
            for name, fn in pieces:
                self._restore(snap)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                graphs[name] = (g, self.launches_per_step)
            torch.cuda.synchronize()
            self._restore(snap)
            self._graphs = graphs
            return
        self._capture_two_pass(snap)

    def _capture_two_pass(self, snap=None) -> None:
        snap = snap or self._snapshot()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step()
        torch.cuda.synchronize()
        self._restore(snap)
        self._graph = g

    def _snapshot(self):
        pos = self.seq_lens.clone()
        rows_k = [self.k_cache[:, b, :, int(pos[b])].clone() for b in range(self.B)]
        rows_v = [self.v_cache[:, b, :, int(pos[b])].clone() for b in range(self.B)]
        return dict(seq=pos, step=self.step_ctr.clone(), cur=self.cur_token.clone(), tok=self.tokens_out.clone(),
                    rows_k=rows_k, rows_v=rows_v)

    def _restore(self, s) -> None:
        for b in range(self.B):
            self.k_cache[:, b, :, int(s["seq"][b])] = s["rows_k"][b]
            self.v_cache[:, b, :, int(s["seq"][b])] = s["rows_v"][b]
        self.seq_lens.copy_(s["seq"]); self.step_ctr.copy_(s["step"]); self.cur_token.copy_(s["cur"])
        self.tokens_out.copy_(s["tok"])

    def decode(self, n_steps: int, use_graph: bool = True, stop_token: Optional[int] = None) -> int:
        """Run n_steps ensemble steps; returns the number done.  Two-pass path and batches: fully asynchronous, always
        n_steps.  One sequence on the speculative path: one stream synchronisation per pass (the repeat decision), and
        with stop_token the loop ends early once that token has been generated."""
        if int(self.seq_lens.max()) + n_steps > self.S_max:
            raise ops.DropDecError("decode would overflow the KV cache")
        if use_graph and (self._graphs is None if self.spec_active else self._graph is None):
            self.capture()
        if self.spec_batch:
            self._decode_async(n_steps, use_graph)
            return n_steps
        if self.spec_active:
            return self._spec_steps(n_steps, self._graphs if use_graph else None, stop_token)
        for _ in range(n_steps):
            if use_graph:
                if self._graph is None:
                    self.capture()
                self._graph.replay()
            else:
                self.step()
            self.launch_count += self.launches_per_step
        return n_steps

    # ------------------------------------------------------------------------------------
    # prefill: library ops (not the hot path)
    # ------------------------------------------------------------------------------------
    def _norm_t(self, x, w):
        xf = x.float()
        xn = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.cfg.rms_eps)).to(torch.bfloat16)
        return xn * w

    def prefill(self, embeds: torch.Tensor, vis_start: Sequence[int]) -> torch.Tensor:
        """embeds bf16 [B, S0, d] (merged text|visual|text).  Fills the cache and returns the
        final-norm hidden states [B, S0, d]."""
        if self.prefill_graph and embeds.is_cuda:
            # the ~1500 library launches of a prefill are host-bound when issued eagerly (7B, 593 positions: 16.0 ms
            # eager, 13.0 ms replayed): one graph per prompt length, replayed on a static input buffer
            ent = self._prefill_graphs.get(embeds.shape[1])
            if ent is None:
                if len(self._prefill_graphs) >= 4:
                    self._prefill_graphs.clear()
                buf = embeds.to(torch.bfloat16).clone()
                s = torch.cuda.Stream(device=self.dev)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._prefill_layers(buf)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self._prefill_layers(buf)
                ent = self._prefill_graphs[embeds.shape[1]] = (g, buf, out)
            g, buf, out = ent
            buf.copy_(embeds)
            g.replay()
            hidden = out.clone()
        else:
            hidden = self._prefill_layers(embeds)
        self.seq_lens.fill_(embeds.shape[1])
        self.vis_start.copy_(torch.as_tensor(list(vis_start), dtype=torch.int32))
        self.step_ctr.zero_()
        return hidden

    def _prefill_layers(self, embeds: torch.Tensor) -> torch.Tensor:
        cfg, B = self.cfg, self.B
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        S0 = embeds.shape[1]
        cos, sin = self.cos[:S0, None, :], self.sin[:S0, None, :]

        def rope(t):                                   # [B, S0, heads, D]
            t = t.float()
            x1, x2 = t[..., :D // 2], t[..., D // 2:]
            return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], -1).to(torch.bfloat16)

        if self.prefill_native:
            return self._prefill_layers_native(embeds)
        x = embeds.to(torch.bfloat16)
        for l, lw in enumerate(self.w.layers):
            h = self._norm_t(x, lw["ln1"])
            qkv = F.linear(h, lw["wqkv"])
            q, k, v = qkv.split([H * D, Hkv * D, Hkv * D], dim=-1)
            q, k = rope(q.view(B, S0, H, D)), rope(k.view(B, S0, Hkv, D))
            v = v.view(B, S0, Hkv, D)
            self.k_cache[l, :, :, :S0] = k.transpose(1, 2)
            self.v_cache[l, :, :, :S0] = v.transpose(1, 2)
            a = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                               is_causal=True, enable_gqa=(H != Hkv))
            x = x + F.linear(a.transpose(1, 2).reshape(B, S0, H * D), lw["wo"])
            h = self._norm_t(x, lw["ln2"])
            gu = F.linear(h, lw["wgu"])
            x = x + F.linear(F.silu(gu[..., :cfg.d_ff]) * gu[..., cfg.d_ff:], lw["wd"])
        return self._norm_t(x, self.w.norm)

    def _prefill_layers_native(self, embeds: torch.Tensor) -> torch.Tensor:
        """The same prefill with the elementwise chains on this library's kernels: dd_rmsnorm for the 2L+1 norms and
        dd_rope_stage (every prompt position = a one-member sequence with its own RoPE position) for split + RoPE of
        q / k / v -- the rounding points of the decode path, ~25 instead of ~60 launches per layer (7B, 593 positions:
        13.0 -> 11 ms replayed).  GEMMs and the causal attention stay library calls."""
        cfg, B = self.cfg, self.B
        H, Hkv, D, d, eps = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cfg.d_model, cfg.rms_eps
        S0 = embeds.shape[1]
        rows = B * S0
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        pos = torch.arange(S0, dtype=torch.int32, device=self.dev).repeat(B)
        q = torch.empty(rows, 1, H, D, **bf)
        k, v = torch.empty(rows, 1, Hkv, D, **bf), torch.empty(rows, 1, Hkv, D, **bf)
        x = embeds.to(torch.bfloat16).reshape(rows, d).contiguous()
        for l, lw in enumerate(self.w.layers):
            qkv = F.linear(ops.rmsnorm(x, lw["ln1"], eps), lw["wqkv"])
            ops.rope_stage(qkv, self.cos, self.sin, pos, rows, 1, H, Hkv, D, q, k, v)
            kt, vt = k.view(B, S0, Hkv, D).transpose(1, 2), v.view(B, S0, Hkv, D).transpose(1, 2)
            self.k_cache[l, :, :, :S0] = kt
            self.v_cache[l, :, :, :S0] = vt
            a = F.scaled_dot_product_attention(q.view(B, S0, H, D).transpose(1, 2), kt, vt, is_causal=True,
                                               enable_gqa=(H != Hkv))
            x = x + F.linear(a.transpose(1, 2).reshape(rows, H * D), lw["wo"])
            gu = F.linear(ops.rmsnorm(x, lw["ln2"], eps), lw["wgu"])
            x = x + F.linear(F.silu(gu[:, :cfg.d_ff]) * gu[:, cfg.d_ff:], lw["wd"])
        return ops.rmsnorm(x, self.w.norm, eps).view(B, S0, d)

    def load_prefill(self, k: torch.Tensor, v: torch.Tensor, vis_start: Sequence[int]) -> None:
        """Inject an externally computed prefix cache: k, v [L, B, H_kv, S0, D] (parity tests)."""
        S0 = k.shape[3]
        self.k_cache[:, :, :, :S0] = k.to(self.dev, torch.bfloat16)
        self.v_cache[:, :, :, :S0] = v.to(self.dev, torch.bfloat16)
        self.seq_lens.fill_(S0)
        self.vis_start.copy_(torch.as_tensor(list(vis_start), dtype=torch.int32))
        self.step_ctr.zero_()

    def uncertainty_stage(self, vis_hidden: torch.Tensor) -> None:
        """vis_hidden bf16 [B, n_vis, d] (final-norm hidden states of the visual rows)."""
        if self._unc_ws is None:
            self._unc_ws = ops.uncertainty_workspace(self.B, self.n_vis, self.w.lm_head.shape[0], self.topk, self.epi.device)
        *_, ovf = ops.uncertainty_project(vis_hidden.contiguous(), self.w.lm_head, topk=self.topk,
                                          round_logits=self.round_logits, return_overflow=True, workspace=self._unc_ws,
                                          out=(self.epi, self.alea, self.var, self.topk_ids))
        self._unc_overflow.copy_(ovf)      # checked at the next host synchronisation (generated())

    def first_token(self, last_hidden: torch.Tensor) -> None:
        """The first generated token is the plain argmax of the unmasked prefill
        (models/llava.py:382-388); last_hidden bf16 [B, d], final-norm."""
        B = self.B
        logits = self.pa.logits.view(B, self.Vp)
        self._linear(last_hidden.contiguous(), self.w.lm_head_padded, ops.EPI_F32, logits)
        ops.aggregate(self.pa.logits, "vote", member_ids=self.member_ids_a, winner=self.winner_a, token=self.cur_token,
                      vocab=self.cfg.vocab)
        self.tokens_out[:, 0] = self.cur_token
        ops.token_stats(self.pa.logits, self.cfg.vocab, self.token_stats)          # slot 0: the prefill token
        self.step_ctr.zero_()          # ensemble steps done; also the sampler's Philox call index / K
        # a new batch of sequences: re-arm the speculative schedule and its adaptive switch
        self.spec_active = self._spec_fits and self.speculative != "off"
        self._spec_base = (self.n_spec_steps, self.n_fallbacks)

    def set_uniforms(self, u: Optional[torch.Tensor]) -> None:
        """Parity mode: u [steps, B, K, n_vis] replaces the Philox draws (step s uses u[s])."""
        if u is None:
            if self.uniforms is not None:
                self._graph = self._graphs = None   # the captured step reads the uniforms buffer: re-capture without it
            self.uniforms = self._u_step = self._u_step1 = None
            return
        self._graph = self._graphs = None
        self.uniforms = u.to(self.dev, torch.float32).contiguous()
        self._u_step = torch.empty(self.B, self.K, self.n_vis, dtype=torch.float32, device=self.dev)
        self._u_step1 = torch.empty_like(self._u_step)

    def step_logits(self) -> torch.Tensor:
        """fp32 [B, V]: what the reference's forward() returns after a decode step -- the winning
        member's logits (models/llava.py:370-376) or the member mean (select_by_average)."""
        V = self.cfg.vocab
        if self._last_src == "mixed":       # asynchronous batch schedule: each sequence's own advancing pass
            return self.ret_logits
        if self.aggregate == "mean":
            return self.out_logits
        if self.vote_on == "hidden":
            return self.pa.logits[:, 0, :V]
        if self._last_src == "pc":          # the speculative pass was the real one: members are rows 1..K
            return self.pc.logits[self._rows(), self.winner.long() + 1, :V]
        if self._last_src == "pd0":         # lookahead pass: the step's members are rows 0..K-1 ...
            return self.pd.logits[self._rows(), self.winner.long(), :V]
        if self._last_src == "pd1":         # ... and the following step's are rows K+1..2K
            return self.pd.logits[self._rows(), self.winner.long() + self.K + 1, :V]
        return self.pb.logits[self._rows(), self.winner.long(), :V]

    def _rows(self) -> torch.Tensor:
        if not hasattr(self, "_row_idx"):
            self._row_idx = torch.arange(self.B, device=self.dev)
        return self._row_idx

    def token_bookkeeping(self, b: int = 0):
        """(logits_mask_prob, token_entropies, token_ventropies) of sequence b as Python lists, one entry per
        generated token -- what the reference appends per forward (models/instructblip.py:141-157)."""
        n = int(self.step_ctr.item()) + 1
        t = self.token_stats[:n, b].cpu()
        return t[:, 0].tolist(), t[:, 1].tolist(), t[:, 2].tolist()

    def generated(self) -> torch.Tensor:
        """Tokens generated so far, int32 [B, n].  (A host synchronisation point: also where an overflow of the
        uncertainty stage's candidate lists surfaces as an error.)"""
        n = int(self.step_ctr.item()) + 1
        ops.raise_on_overflow(self._unc_overflow)
        return self.tokens_out[:, :n]
