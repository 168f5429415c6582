"""CPU restatement of the Dropout-Decoding hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines (relative to /root/reference) it
follows.  The reference is pure Python on top of torch, so the restatement is
plain torch on CPU: the arithmetic the parity gates care about (fp32 softmax /
entropy sums, fp32 affine map + compare for the mask draw, integer argmax /
vote) is done with exactly the reference's operator sequence, so index sets
and votes are bit-identical to the reference functions (checked by
tests/golden/make_golden.py against the imported reference code).

Only the decoder forward is *restated from the published algorithm* (Llama /
Mistral decoder of transformers==4.44.0, environment.yml:106, not vendored in
the reference): RMSNorm -> QKV -> RoPE (rotate-half) -> cache append -> masked
softmax attention -> O -> SwiGLU MLP, final RMSNorm, lm_head, ``.float()``
(models/llama.py:27-47).  It is pinned against the installed transformers
LlamaForCausalLM in tests/test_oracle_lm.py.

Product code must never import this module.
"""
from __future__ import annotations

import copy
import math
from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

__all__ = [
    "LMConfig", "init_lm_weights", "OracleLM", "OracleCache",
    "calculate_vision_uncertainty", "topk_token_ids", "relevant_visual_tokens",
    "drop_probabilities", "sample_mask_epis", "quantile_threshold",
    "sample_mask_quantile", "lowest_percent_kl_indices", "sample_mask_epis_kl", "select_by_vote", "select_by_average",
    "entropy_varentropy", "dropout_decode", "ensemble_steps", "DecodeTrace", "bf16r",
    "rope_table",
]


# --------------------------------------------------------------------------------------
# numerics helpers
# --------------------------------------------------------------------------------------

def bf16r(x: torch.Tensor) -> torch.Tensor:
    """Round an fp32 tensor to bf16 (RNE) and return it as fp32."""
    return x.to(torch.bfloat16).to(torch.float32)


def _ident(x: torch.Tensor) -> torch.Tensor:
    return x


# --------------------------------------------------------------------------------------
# (a7) visual-token uncertainty            models/llava.py:710-756
# --------------------------------------------------------------------------------------

def calculate_vision_uncertainty(logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """logits fp32 [B, N_vis, V] -> dict (models/llava.py:722-756, same op order)."""
    probs = F.softmax(logits, dim=-1)                                   # :722
    variance_per_token = torch.var(probs, dim=-1)                        # :728 (unbiased)
    variance = torch.mean(variance_per_token, dim=-1)                    # :729
    p_avg = torch.mean(probs, dim=1)                                     # :732
    epi_per_token = torch.sum(
        probs * (torch.log(probs + 1e-10) - torch.log(p_avg.unsqueeze(1) + 1e-10)), dim=-1)  # :735
    alea_per_token = -torch.sum(probs * torch.log(probs + 1e-10), dim=-1)  # :739
    return {
        "variance_per_token": variance_per_token,
        "epis_uncert_per_token": epi_per_token,
        "alea_uncert_per_token": alea_per_token,
        "variance": variance,
        "epis_uncert": torch.mean(epi_per_token, dim=-1),                # :743
        "alea_uncert": torch.mean(alea_per_token, dim=-1),               # :744
    }


# (a6) models/llava.py:428-441 (k=5); llavanext.py:652, instructblip.py:187 (k=10)
def topk_token_ids(image_logits: torch.Tensor, topk: int = 5):
    return torch.topk(image_logits, topk, dim=-1)


# --------------------------------------------------------------------------------------
# (a8) relevant ("overlap") visual tokens   models/llava.py:443-482
# --------------------------------------------------------------------------------------

def relevant_visual_tokens(prelim_logits: torch.Tensor, topk_ids: torch.Tensor) -> torch.Tensor:
    """prelim_logits [1,1,V] (decode step), topk_ids [1,N_vis,k] -> sorted int64
    indices i (0-based inside the visual span) with argmax(prelim) in topk_ids[i].
    The reference returns them offset by start_image_pos and possibly 0-d
    (``.squeeze()``, :476); callers here add the offset themselves."""
    max_ids = torch.argmax(prelim_logits, dim=-1)       # :459  [1,1]
    matches = topk_ids[0] == max_ids[0]                 # :466-470 (batch element 0 only)
    return torch.nonzero(matches.any(dim=1)).reshape(-1)  # :473-476


# --------------------------------------------------------------------------------------
# (a9) stochastic "epis" mask               models/llava.py:638-662
# --------------------------------------------------------------------------------------

def drop_probabilities(epis: torch.Tensor, prob: float) -> torch.Tensor:
    """Per-token drop probability, fp32, exactly the reference's operator order
    (models/llava.py:641-647).  quantile(.,0)/quantile(.,1) are min/max."""
    q_low = torch.quantile(epis, 0)
    q_high = torch.quantile(epis, 1)
    return 0.1 + (prob - 0.1) * (epis.clamp(min=q_low, max=q_high) - q_low) / (q_high - q_low)


def sample_mask_epis(epis: torch.Tensor, prob: float, uniforms: torch.Tensor,
                     relevant: Optional[torch.Tensor], dropped_in: Optional[torch.Tensor] = None
                     ) -> torch.Tensor:
    """Returns a bool tensor [N_vis], True = visual token hidden from the new query.

    uniforms stands for ``torch.rand_like(epis)`` (:650).  ``dropped_in`` carries
    the previous members' zeros: LLaVA-1.5 does not restore the mask between
    members (:344 is commented out) so masks accumulate; LLaVA-NeXT restores it
    (llavanext.py:546) -> pass None.  ``relevant`` = None is "epis_no_overlap"
    (:663-683)."""
    drop = uniforms < drop_probabilities(epis, prob)     # :653
    out = drop if dropped_in is None else (dropped_in | drop)   # :657 on a non-restored mask
    if relevant is not None and relevant.numel() > 0:
        out = out.clone()
        out[relevant] = False                             # :660
    return out


# (a9') deterministic quantile mask         models/instructblip.py:447-460
def quantile_threshold(epis: torch.Tensor, prob: float) -> torch.Tensor:
    return torch.quantile(epis, 1 - prob)                # :450


def sample_mask_quantile(epis: torch.Tensor, prob: float, relevant: Optional[torch.Tensor]
                         ) -> torch.Tensor:
    out = epis >= quantile_threshold(epis, prob)          # :453
    if relevant is not None and relevant.numel() > 0:
        out = out.clone()
        out[relevant] = False                             # :460
    return out


# (f2) "epis_kl" mask (a dead-code method of the reference, kept as an ablation knob)
#      models/llava.py:758-776, models/instructblip.py:464-485, 559-578
def lowest_percent_kl_indices(image_logits: torch.Tensor, logits: torch.Tensor, percent: float = 0.1) -> torch.Tensor:
    """image_logits [1, N, V], logits [1, 1, V] -> indices of the int(percent*N) visual tokens whose projected
    distribution is closest to the current one: F.kl_div(log_softmax(image), softmax(cur).expand, 'none').sum(-1),
    then topk(largest=False) (models/llava.py:763-776)."""
    il = image_logits.squeeze(0)
    cur = logits.squeeze(0)
    kl = F.kl_div(F.log_softmax(il, dim=-1), F.softmax(cur, dim=-1).expand_as(il), reduction='none').sum(dim=-1)
    n = int(percent * kl.numel())
    return torch.topk(kl, n, largest=False).indices


def sample_mask_epis_kl(epis: torch.Tensor, prob: float, uniforms: torch.Tensor, image_logits: torch.Tensor,
                        logits: torch.Tensor) -> torch.Tensor:
    """models/instructblip.py:464-485: the 'epis' draw, then the lowest-KL tenth of the visual tokens restored."""
    out = (uniforms < drop_probabilities(epis, prob)).clone()
    out[lowest_percent_kl_indices(image_logits, logits)] = False
    return out


# --------------------------------------------------------------------------------------
# (a10/a11) ensemble aggregation            models/llava.py:22-52
# --------------------------------------------------------------------------------------

def select_by_vote(member_last_logits: Sequence[torch.Tensor]):
    """member_last_logits: K tensors [V] (each ``output[0][-1]`` flattened, :27).
    Returns (winner_member_index, winning_token_id).  Counter.most_common breaks
    ties by first insertion = earliest member (:29)."""
    ids = [int(torch.argmax(l).item()) for l in member_last_logits]
    c = Counter()
    for t in ids:
        c[t] += 1
    win = c.most_common(1)[0][0]
    for i, t in enumerate(ids):
        if t == win:
            return i, win
    raise AssertionError


def select_by_average(member_last_logits: Sequence[torch.Tensor]) -> torch.Tensor:
    """fp32 mean over members via numpy on CPU (:42-47); the reference writes it
    into member 0's output, so member 0's cache is the one that survives (:39,50)."""
    stacked = torch.stack(list(member_last_logits)).cpu()
    return torch.tensor(stacked.numpy().mean(axis=0))


# (a13) models/llava.py:687-708
def entropy_varentropy(logits: torch.Tensor):
    log_probs = F.log_softmax(logits, dim=-1)
    probs = torch.exp(log_probs)
    entropy = -torch.sum(probs * log_probs) / math.log(2)
    varentropy = torch.sum(probs * (log_probs / math.log(2) + entropy) ** 2)
    return entropy.item(), varentropy.item()


# --------------------------------------------------------------------------------------
# (a2/a3) decoder forward, restated         [upstream transformers 4.44 Llama]
# --------------------------------------------------------------------------------------

@dataclass
class LMConfig:
    n_layers: int = 2
    d_model: int = 256
    n_heads: int = 4
    n_kv_heads: int = 4
    head_dim: int = 64
    d_ff: int = 512
    vocab: int = 32064
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0

    @staticmethod
    def llava15_7b() -> "LMConfig":
        return LMConfig(32, 4096, 32, 32, 128, 11008, 32064, 1e-5, 10000.0)

    @staticmethod
    def tiny() -> "LMConfig":            # BASELINE.json configs[0]
        return LMConfig(2, 256, 4, 4, 64, 512, 32064, 1e-5, 10000.0)


def init_lm_weights(cfg: LMConfig, seed: int = 0, std: float = 0.02,
                    head_std: Optional[float] = None, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random init; values are bf16-representable so that the bf16 GPU
    path and the fp32 oracle see identical weights."""
    g = torch.Generator().manual_seed(seed)

    def n(*shape, s=std):
        return bf16r(torch.randn(*shape, generator=g) * s).to(dtype)

    w: Dict[str, torch.Tensor] = {"embed": n(cfg.vocab, cfg.d_model)}
    for l in range(cfg.n_layers):
        p = f"l{l}."
        w[p + "ln1"] = torch.ones(cfg.d_model, dtype=dtype)
        w[p + "wq"] = n(cfg.n_heads * cfg.head_dim, cfg.d_model)
        w[p + "wk"] = n(cfg.n_kv_heads * cfg.head_dim, cfg.d_model)
        w[p + "wv"] = n(cfg.n_kv_heads * cfg.head_dim, cfg.d_model)
        w[p + "wo"] = n(cfg.d_model, cfg.n_heads * cfg.head_dim)
        w[p + "ln2"] = torch.ones(cfg.d_model, dtype=dtype)
        w[p + "wg"] = n(cfg.d_ff, cfg.d_model)
        w[p + "wu"] = n(cfg.d_ff, cfg.d_model)
        w[p + "wd"] = n(cfg.d_model, cfg.d_ff)
    w["norm"] = torch.ones(cfg.d_model, dtype=dtype)
    w["lm_head"] = n(cfg.vocab, cfg.d_model, s=head_std if head_std is not None else std)
    return w


def rope_table(head_dim: int, theta: float, max_pos: int, faithful: bool) -> (torch.Tensor, torch.Tensor):
    """cos/sin [max_pos, head_dim/2] fp32, computed on CPU (the GPU engine copies
    this exact table so both sides use identical values).  In bf16-faithful mode
    the table is rounded to bf16 as transformers does (cos.to(x.dtype))."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None, :]
    cos, sin = ang.cos(), ang.sin()
    if faithful:
        cos, sin = bf16r(cos), bf16r(sin)
    return cos, sin


class OracleCache:
    """Per-layer K/V lists, batch 1.  ``copy.deepcopy`` of it is what the
    reference does once per member per step (models/llava.py:292,343)."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers   # [H_kv, S, D]
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    def seq_len(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[1]

    def append(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=1)
            self.v[layer] = torch.cat([self.v[layer], v], dim=1)
        return self.k[layer], self.v[layer]


class OracleLM:
    """Batch-1 decoder.  ``faithful=False``: everything fp32 (pins against HF fp32).
    ``faithful=True``: fp32 arithmetic with a bf16 rounding wherever the GPU
    engine stores bf16 (GEMM outputs, norm outputs, RoPE outputs, activations,
    residual stream) -- the 'fp32-accumulate' mode of the parity gate."""

    def __init__(self, cfg: LMConfig, weights: Dict[str, torch.Tensor], faithful: bool = False,
                 max_pos: int = 8192):
        self.cfg, self.w, self.faithful = cfg, weights, faithful
        self.r = bf16r if faithful else _ident
        self.cos, self.sin = rope_table(cfg.head_dim, cfg.rope_theta, max_pos, faithful)

    # -- pieces --------------------------------------------------------------------
    def rmsnorm(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        var = x.pow(2).mean(-1, keepdim=True)
        xn = self.r(x * torch.rsqrt(var + self.cfg.rms_eps))
        return self.r(xn * w)

    def rope(self, x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        # x [T, H, D]; rotate-half convention (pairs i, i + D/2)
        half = x.shape[-1] // 2
        cos, sin = self.cos[pos][:, None, :], self.sin[pos][:, None, :]
        x1, x2 = x[..., :half], x[..., half:]
        return self.r(torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1))

    def embed(self, ids: torch.Tensor) -> torch.Tensor:
        return self.w["embed"][ids]

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        """lm_head + ``.float()`` (models/llama.py:46-47)."""
        return self.r(hidden @ self.w["lm_head"].T)

    # -- full forward ----------------------------------------------------------------
    def forward(self, embeds: torch.Tensor, cache: OracleCache, mask: torch.Tensor,
                pos: torch.Tensor) -> torch.Tensor:
        """embeds [T,d]; mask [S_past+T] of 0/1 over keys; pos [T].  Appends the new
        K/V to ``cache`` in place (as DynamicCache.update does) and returns the
        final-norm hidden states [T,d]."""
        cfg, r = self.cfg, self.r
        T = embeds.shape[0]
        x = embeds.to(torch.float32)
        S_past = cache.seq_len()
        S = S_past + T
        assert mask.shape[-1] == S
        # causal + key padding mask, additive
        qpos = torch.arange(S_past, S)[:, None]
        kpos = torch.arange(S)[None, :]
        allowed = (kpos <= qpos) & (mask.reshape(1, S) != 0)
        bias = torch.zeros(T, S).masked_fill(~allowed, float("-inf"))
        g = cfg.n_heads // cfg.n_kv_heads
        scale = 1.0 / math.sqrt(cfg.head_dim)
        for l in range(cfg.n_layers):
            p = f"l{l}."
            h = self.rmsnorm(x, self.w[p + "ln1"])
            q = r(h @ self.w[p + "wq"].T).view(T, cfg.n_heads, cfg.head_dim)
            k = r(h @ self.w[p + "wk"].T).view(T, cfg.n_kv_heads, cfg.head_dim)
            v = r(h @ self.w[p + "wv"].T).view(T, cfg.n_kv_heads, cfg.head_dim)
            q, k = self.rope(q, pos), self.rope(k, pos)
            K, V = cache.append(l, k.transpose(0, 1).contiguous(), v.transpose(0, 1).contiguous())
            Kx = K.repeat_interleave(g, dim=0) if g > 1 else K        # [H, S, D]
            Vx = V.repeat_interleave(g, dim=0) if g > 1 else V
            s = torch.einsum("thd,hsd->hts", q, Kx) * scale + bias[None]
            pr = torch.softmax(s, dim=-1)
            o = r(torch.einsum("hts,hsd->thd", pr, Vx).reshape(T, cfg.n_heads * cfg.head_dim))
            x = r(x + r(o @ self.w[p + "wo"].T))
            h = self.rmsnorm(x, self.w[p + "ln2"])
            gate = r(h @ self.w[p + "wg"].T)
            up = r(h @ self.w[p + "wu"].T)
            act = r(r(F.silu(gate)) * up)
            x = r(x + r(act @ self.w[p + "wd"].T))
        return self.rmsnorm(x, self.w["norm"])


# --------------------------------------------------------------------------------------
# (a1) the per-token ensemble               models/llava.py:218-388 (+ HF greedy loop)
# --------------------------------------------------------------------------------------

@dataclass
class DecodeTrace:
    tokens: List[int] = field(default_factory=list)           # new tokens, first = unmasked prefill argmax
    y_init: List[int] = field(default_factory=list)           # per ensemble step
    y_init_top2: List[tuple] = field(default_factory=list)    # per step (id1, logit1, id2, logit2) of the preliminary logits
    relevant: List[List[int]] = field(default_factory=list)   # per step, 0-based visual indices
    dropped: List[List[List[int]]] = field(default_factory=list)   # [step][member] -> sorted indices
    masked_numbers: List[List[int]] = field(default_factory=list)  # models/llava.py:661-662
    member_ids: List[List[int]] = field(default_factory=list)      # member argmax ids
    winner: List[int] = field(default_factory=list)                # winning member index
    step_logits: List[torch.Tensor] = field(default_factory=list)  # returned logits [V] per step
    member_logits: List[List[torch.Tensor]] = field(default_factory=list)  # [step][member] vote inputs (keep_logits)
    step_kv: List[tuple] = field(default_factory=list)   # [step] (k [L,H_kv,D], v [L,H_kv,D]): the committed (winner's) new rows
    uncert: Optional[Dict[str, torch.Tensor]] = None
    topk_ids: Optional[torch.Tensor] = None
    vis_hidden: Optional[torch.Tensor] = None
    prefill_cache: Optional[OracleCache] = None
    last_hidden: Optional[torch.Tensor] = None


def dropout_decode(lm: OracleLM, prompt_embeds: torch.Tensor, vis_start: int, n_vis: int,
                   max_new_tokens: int, probs: Sequence[float], uniforms: Optional[torch.Tensor],
                   mode: str = "cumulative", keep_relevant: bool = True, topk: int = 5,
                   aggregate: str = "vote", eos_token_id: Optional[int] = None,
                   keep_logits: bool = False, keep_prefill: bool = False, vote_on: str = "logits") -> DecodeTrace:
    """One image, batch 1: prefill + uncertainty stage + (T-1) ensemble steps.

    prompt_embeds [S0, d] is the merged [text | visual | text] embedding sequence
    (what _merge_input_ids_with_image_features produces, models/llava.py:74-153);
    visual tokens occupy positions vis_start .. vis_start+n_vis-1 (:75-78).

    mode: 'cumulative' (LLaVA-1.5, mask not restored between members,
    models/llava.py:342-346) | 'independent' (LLaVA-NeXT, llavanext.py:544-551) |
    'quantile' (InstructBLIP, instructblip.py:447-460).
    uniforms [steps, K, n_vis] replaces the K ``rand_like`` draws per step (:650).
    The greedy loop is HF 4.44 ``_sample``: argmax of logits[:, -1], append, stop at
    max_new_tokens or EOS."""
    K = len(probs)
    tr = DecodeTrace()
    S0 = prompt_embeds.shape[0]
    cache = OracleCache(lm.cfg.n_layers)

    # ---- prefill (models/llava.py:218-224, 285-314, 382-388) -------------------------
    hidden = lm.forward(prompt_embeds, cache, torch.ones(S0, dtype=torch.long), torch.arange(S0))
    vis_hidden = hidden[vis_start:vis_start + n_vis]
    vis_logits = lm.logits(vis_hidden)[None]                 # get_image_logits :412-426
    tr.uncert = calculate_vision_uncertainty(vis_logits)     # :314
    _, tr.topk_ids = topk_token_ids(vis_logits, topk)        # :310 -> :408
    epis = tr.uncert["epis_uncert_per_token"][0]
    last_logits = lm.logits(hidden[-1:])[0]
    tok = int(torch.argmax(last_logits).item())              # first token: unmasked (:382-388)
    tr.tokens.append(tok)
    if keep_logits:
        tr.step_logits.append(last_logits)
    if keep_prefill:
        tr.vis_hidden, tr.last_hidden = vis_hidden.clone(), hidden[-1].clone()
        tr.prefill_cache = copy.deepcopy(cache)

    ensemble_steps(lm, cache, tok, epis, tr.topk_ids, vis_start, n_vis, max_new_tokens - 1, probs, uniforms,
                   mode=mode, keep_relevant=keep_relevant, aggregate=aggregate, eos_token_id=eos_token_id,
                   keep_logits=keep_logits, trace=tr, vote_on=vote_on)
    return tr


def ensemble_steps(lm: OracleLM, cache: OracleCache, tok: int, epis: torch.Tensor, topk_ids: torch.Tensor,
                   vis_start: int, n_vis: int, n_steps: int, probs: Sequence[float],
                   uniforms: Optional[torch.Tensor], mode: str = "cumulative", keep_relevant: bool = True,
                   aggregate: str = "vote", eos_token_id: Optional[int] = None, keep_logits: bool = False,
                   trace: Optional[DecodeTrace] = None, vote_on: str = "logits") -> DecodeTrace:
    """The decode branch of forward() (models/llava.py:254-283, 292-376) driven by the HF greedy
    loop: per token one preliminary forward, K masked forwards each on a deep copy of the whole
    cache, then the vote.  Starts from `cache` (prefix already filled) and the last token `tok`."""
    tr = trace if trace is not None else DecodeTrace(tokens=[tok])
    step = 0
    while step < n_steps and (eos_token_id is None or tok != eos_token_id):
        S = cache.seq_len()
        x = lm.embed(torch.tensor([tok]))                    # :229
        ones = torch.ones(S + 1, dtype=torch.long)           # :266-283
        pos = torch.tensor([S])                              # position_ids = sum(mask) - 1
        original = copy.deepcopy(cache)                      # :292
        prelim = lm.logits(lm.forward(x, cache, ones, pos))  # :294-305 (cache mutated, then dropped)
        y0 = int(torch.argmax(prelim[0]).item())
        tr.y_init.append(y0)
        t2 = torch.topk(prelim[0].reshape(-1), 2)
        tr.y_init_top2.append((int(t2.indices[0]), float(t2.values[0]), int(t2.indices[1]), float(t2.values[1])))
        rel = relevant_visual_tokens(prelim[None], topk_ids) if keep_relevant else None
        tr.relevant.append([] if rel is None else rel.tolist())

        member_logits, member_caches, dropped_sets, masked_n = [], [], [], []
        dropped = None
        for m, prob in enumerate(probs):                     # :342
            cache_m = copy.deepcopy(original)                # :343
            if mode == "quantile":
                d = sample_mask_quantile(epis, prob, rel)
            else:
                u = uniforms[step, m]
                d = sample_mask_epis(epis, prob, u, rel, dropped if mode == "cumulative" else None)
            if mode == "cumulative":
                dropped = d
            mask = ones.clone()
            mask[vis_start:vis_start + n_vis][d] = 0         # :657-660
            masked_n.append(int((mask == 0).sum().item()))   # :661
            h = lm.forward(x, cache_m, mask, pos)            # :350-359
            member_logits.append(h[0] if vote_on == "hidden" else lm.logits(h)[0])
            member_caches.append(cache_m)
            dropped_sets.append(torch.nonzero(d).reshape(-1).tolist())
        tr.dropped.append(dropped_sets)
        tr.masked_numbers.append(masked_n)
        tr.member_ids.append([int(torch.argmax(l).item()) for l in member_logits])

        if aggregate == "vote":
            win, _ = select_by_vote(member_logits)           # :361
            out_logits = member_logits[win]
            if vote_on == "hidden":                          # instructblip.py:137-140
                out_logits = lm.logits(out_logits[None])[0]
        elif aggregate == "mean":
            win = 0                                          # :39 (member 0's output / cache)
            out_logits = select_by_average(member_logits)
        else:
            raise ValueError(aggregate)
        tr.winner.append(win)
        cache = member_caches[win]                           # :373 winner's cache becomes live
        tok = int(torch.argmax(out_logits).item())           # HF greedy
        tr.tokens.append(tok)
        if keep_logits:
            tr.step_logits.append(out_logits)
            tr.member_logits.append(member_logits)
            tr.step_kv.append((torch.stack([k[:, -1] for k in cache.k]), torch.stack([v[:, -1] for v in cache.v])))
        step += 1
    return tr
