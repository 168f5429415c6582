#!/usr/bin/env python
"""Mint the golden fixtures under tests/golden/ and pin the oracle.

Runs ONLY in the build container (it imports the reference from
/root/reference, which does not exist on the GPU box).  For every pure function
of the hot path it calls the *reference's own code* (unbound methods on a stub
``self``, exactly as SURVEY.md section 8c describes) and the oracle's
restatement on the same seeded inputs, asserts bit-equality, and stores the
reference's outputs as fixtures.  For the whole per-token ensemble it drives the
reference's real ``CustomLlavaForConditionalGeneration.forward`` (models/llava.py:
155-388) with a stub ``self`` whose ``language_model`` is the oracle decoder, inside
a restated HF-4.44 greedy loop, and checks that ``oracle.dropout_decode`` produces the
same tokens, mask sets, votes and logits.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz
"""
import copy
import importlib
import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def load_reference():
    """Import the reference's modules under their own names, isolated from this
    repo's drop-in ``models`` package."""
    assert os.path.isdir(REF), "reference tree not present (run in the build container)"
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ref_llava = importlib.import_module("models.llava")
        ref_next = importlib.import_module("models.llavanext")
        # models/instructblip.py:3 imports turtledemo (needs tkinter) -> strip that line
        src = open(os.path.join(REF, "models", "instructblip.py")).read().splitlines()
        assert "turtledemo" in src[2]
        src[2] = ""
        ref_blip = types.ModuleType("models.instructblip")
        ref_blip.__file__ = os.path.join(REF, "models", "instructblip.py")
        exec(compile("\n".join(src), ref_blip.__file__, "exec"), ref_blip.__dict__)
    finally:
        sys.path.remove(REF)
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                    if k == "models" or k.startswith("models.")}
    return ref_llava, ref_next, ref_blip, ref_mods


ref_llava, ref_next, ref_blip, _REF_MODS = load_reference()
sys.path.insert(0, REPO)
import oracle  # noqa: E402
from oracle import (LMConfig, OracleLM, OracleCache, init_lm_weights, dropout_decode)  # noqa: E402

RL = ref_llava.CustomLlavaForConditionalGeneration
RN = ref_next.CustomLlavaNextForConditionalGeneration
RB = ref_blip.CustomLlamaForCausalLM


def eq(a, b, what):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape and torch.equal(a, b), f"oracle != reference: {what}"


def close(a, b, what, rtol=1e-5, atol=1e-6):
    """For values that pass through a GEMM whose shape differs between the two sides
    (the reference projects all 593 rows at once, the oracle only the rows it needs):
    BLAS blocking changes the fp32 summation order, so only closeness can be asked."""
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape and torch.allclose(a, b, rtol=rtol, atol=atol), \
        f"oracle !~ reference: {what} (max abs diff {(a - b).abs().max().item():.3e})"


# ----------------------------------------------------------------------------------------
# 1. uncertainty stage: h_vis, lm_head from seeds -> logits -> reference functions
# ----------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(REPO, "tests"))
from golden_cases import (UNCERT_CASES, uncert_inputs, aggregate_inputs, tiny_decode_inputs,  # noqa: E402
                          tiny_uniforms, TINY_N_VIS, TINY_T, TINY_PROBS, TINY_SEED)


def gold_uncertainty():
    out = {}
    for name, n_vis, d, V, hs, ws, k in UNCERT_CASES:
        h, w = uncert_inputs(name, n_vis, d, V, hs, ws)
        logits = (h @ w.T)[None]                       # fp32 [1, n_vis, V]
        stub = SimpleNamespace()
        ref = RL.calculate_vision_uncertainty(stub, logits)
        mine = oracle.calculate_vision_uncertainty(logits)
        for key in ref:
            eq(ref[key], mine[key], f"calculate_vision_uncertainty[{key}] {name}")
        eq(RN.calculate_vision_uncertainty(stub, logits)["epis_uncert_per_token"],
           mine["epis_uncert_per_token"], "llavanext twin")
        eq(RB.calculate_vision_uncertainty(stub, logits)["epis_uncert_per_token"],
           mine["epis_uncert_per_token"], "instructblip twin")
        rv, ri = RL.get_topk_token_id(stub, logits, topk=k)
        mv, mi = oracle.topk_token_ids(logits, k)
        eq(ri, mi, f"topk ids {name}")
        # gap between the k-th and (k+1)-th logit: tells the GPU test which rows have a
        # near-tie at the set boundary
        srt = torch.sort(logits[0], dim=-1, descending=True).values
        out[f"{name}.epi"] = ref["epis_uncert_per_token"][0].numpy()
        out[f"{name}.alea"] = ref["alea_uncert_per_token"][0].numpy()
        out[f"{name}.var"] = ref["variance_per_token"][0].numpy()
        out[f"{name}.topk_ids"] = ri[0].numpy().astype(np.int32)
        out[f"{name}.topk_gap"] = (srt[:, k - 1] - srt[:, k]).numpy()
        out[f"{name}.row_max"] = srt[:, 0].numpy()
        print(f"[uncert] {name}: epi [{ref['epis_uncert_per_token'].min():.4f}, "
              f"{ref['epis_uncert_per_token'].max():.4f}] alea mean {ref['alea_uncert'].item():.4f}")
    np.savez_compressed(os.path.join(HERE, "uncertainty.npz"), **out)


# ----------------------------------------------------------------------------------------
# 2. mask sampler: reference get_image_attention_mask on a stub self
# ----------------------------------------------------------------------------------------

def mask_stub(cls, epis, topk_ids, start):
    s = SimpleNamespace(
        vision_uncert_dict={"epis_uncert_per_token": epis[None]},
        image_features=(None, topk_ids[None]),
        start_image_pos=[start], end_image_pos=[start + epis.numel() - 1],
        masked_numbers=[])
    s.get_overlap_image_tokens = lambda logits, id=-1: cls.get_overlap_image_tokens(s, logits, id)
    return s


def gold_masks():
    out, meta = {}, []
    g = torch.Generator().manual_seed(99)
    cases = [  # name, n_vis, k, V, start, S, probs, steps
        ("llava", 576, 5, 32064, 5, 600, [0.3, 0.5, 0.7], 6),
        ("llava4", 576, 5, 32064, 5, 650, [0.1, 0.3, 0.5, 0.7], 3),
        ("next", 1176, 10, 32064, 7, 1300, [0.3, 0.5, 0.7], 3),
        ("small", 33, 5, 50, 3, 40, [0.3, 0.5, 0.7], 8),       # tiny vocab: relevant set often non-empty
        ("const", 40, 5, 50, 0, 41, [0.3, 0.5, 0.7], 2),       # max == min -> NaN probs -> nothing dropped
    ]
    for name, n_vis, k, V, start, S, probs, steps in cases:
        epis = (torch.rand(n_vis, generator=g) * 0.2 + 0.75).float()
        if name == "const":
            epis = torch.full((n_vis,), 0.8)
        topk_ids = torch.stack([torch.randperm(V, generator=g)[:k] for _ in range(n_vis)])
        prelim = torch.randn(steps, 1, 1, V, generator=g)
        useed = 24                                     # models/llava.py:16
        # --- reference, LLaVA-1.5 flavour: cumulative (no restore, :342-346)
        stub = mask_stub(RL, epis, topk_ids, start)
        torch.manual_seed(useed)
        ref_cum, ref_cnt = [], []
        for t in range(steps):
            mask = torch.ones(1, S, dtype=torch.long)
            stub.masked_numbers = []
            row = []
            for p in probs:
                mask = RL.get_image_attention_mask(stub, prelim[t], mask, method="epis", prob=p)
                row.append((mask[0, start:start + n_vis] == 0).clone())
            ref_cum.append(torch.stack(row))
            ref_cnt.append(list(stub.masked_numbers))
        # --- reference, NeXT flavour: restored between members (llavanext.py:546), epis + no_overlap
        stubn = mask_stub(RN, epis, topk_ids, start)
        ref_ind, ref_noov = [], []
        for method, dst in (("epis", ref_ind), ("epis_no_overlap", ref_noov)):
            torch.manual_seed(useed)
            for t in range(steps):
                row = []
                for p in probs:
                    mask = torch.ones(1, S, dtype=torch.long)
                    mask = RN.get_image_attention_mask(stubn, prelim[t], mask, method=method, prob=p)
                    row.append((mask[0, start:start + n_vis] == 0).clone())
                dst.append(torch.stack(row))
        # --- reference, InstructBLIP flavour: quantile threshold (instructblip.py:447-460)
        stubb = mask_stub(RB, epis, topk_ids, start)
        ref_q = []
        for t in range(steps):
            row = []
            for p in probs:
                mask = torch.ones(1, S, dtype=torch.long)
                mask = RB.get_image_attention_mask(stubb, prelim[t], mask, method="epis", prob=p)
                row.append((mask[0, start:start + n_vis] == 0).clone())
            ref_q.append(torch.stack(row))
        # --- oracle on pre-drawn uniforms (K rand_like calls == one stream)
        torch.manual_seed(useed)
        U = torch.rand(steps, len(probs), n_vis)
        for t in range(steps):
            rel = oracle.relevant_visual_tokens(prelim[t], topk_ids[None])
            ref_rel = RL.get_overlap_image_tokens(stub, prelim[t]) - start
            eq(ref_rel.reshape(-1), rel, f"relevant set {name}[{t}]")
            d = None
            for m, p in enumerate(probs):
                d = oracle.sample_mask_epis(epis, p, U[t, m], rel, d)
                eq(ref_cum[t][m], d, f"cumulative mask {name}[{t},{m}]")
                assert ref_cnt[t][m] == int(d.sum()), "masked_numbers"
                eq(ref_ind[t][m], oracle.sample_mask_epis(epis, p, U[t, m], rel, None),
                   f"independent mask {name}[{t},{m}]")
                eq(ref_noov[t][m], oracle.sample_mask_epis(epis, p, U[t, m], None, None),
                   f"no_overlap mask {name}[{t},{m}]")
                eq(ref_q[t][m], oracle.sample_mask_quantile(epis, p, rel), f"quantile mask {name}[{t},{m}]")
        out[f"{name}.epis"] = epis.numpy()
        out[f"{name}.topk_ids"] = topk_ids.numpy().astype(np.int32)
        out[f"{name}.y_init"] = torch.argmax(prelim[:, 0, 0], dim=-1).numpy().astype(np.int32)
        out[f"{name}.uniforms"] = U.numpy()
        out[f"{name}.cumulative"] = torch.stack(ref_cum).numpy()
        out[f"{name}.independent"] = torch.stack(ref_ind).numpy()
        out[f"{name}.no_overlap"] = torch.stack(ref_noov).numpy()
        out[f"{name}.quantile"] = torch.stack(ref_q).numpy()
        out[f"{name}.probs"] = np.asarray(probs, dtype=np.float64)
        meta.append(dict(name=name, n_vis=n_vis, k=k, V=V, start=start, S=S, steps=steps))
        print(f"[masks] {name}: cumulative counts step0 {ref_cnt[0]}, "
              f"relevant sizes {[int(oracle.relevant_visual_tokens(prelim[t], topk_ids[None]).numel()) for t in range(steps)]}")
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)


# ----------------------------------------------------------------------------------------
# 3. aggregation
# ----------------------------------------------------------------------------------------

def gold_aggregate():
    rows = []
    V = 32064
    for K, l in aggregate_inputs():
        outs = [(l[m][None, None],) for m in range(K)]       # output[0][-1] -> [1, V]
        ref_out, ref_idx = ref_llava.select_by_vote(outs)
        idx, tok = oracle.select_by_vote([l[m] for m in range(K)])
        assert idx == ref_idx and tok == int(torch.argmax(ref_out[0][-1]).item())
        assert ref_next.select_by_vote(outs) is outs[idx]
        avg_in = [[l[m][None].clone()] for m in range(K)]    # output[0][-1] assignable
        ref_avg = ref_llava.select_by_average(avg_in)[0][-1]
        eq(ref_avg, oracle.select_by_average([l[m] for m in range(K)]), "select_by_average")
        rows.append((K, idx, tok, int(torch.argmax(ref_avg).item())))
        e, v = RL.calculate_entropy_varentropy(SimpleNamespace(), l[0])
        e2, v2 = oracle.entropy_varentropy(l[0])
        assert e == e2 and v == v2
    np.savez_compressed(os.path.join(HERE, "aggregate.npz"),
                        cases=np.asarray(rows, dtype=np.int64), seed=np.int64(7), V=np.int64(V))
    print(f"[aggregate] {len(rows)} cases")


# ----------------------------------------------------------------------------------------
# 4. the per-token ensemble: real reference forward() on a stub self vs oracle.dropout_decode
# ----------------------------------------------------------------------------------------

class _RefCache:
    """Stand-in for the 4.44 DynamicCache: subscriptable ([layer][0] -> K [1,H,S,D],
    models/llava.py:257) and deep-copyable (:292,343)."""

    def __init__(self, n_layers):
        self.oc = OracleCache(n_layers)

    def __getitem__(self, l):
        return (self.oc.k[l][None], self.oc.v[l][None])


class _RefOut:
    def __init__(self, logits, cache):
        self.logits, self.past_key_values = logits, cache
        self.hidden_states, self.attentions = None, None

    def __getitem__(self, i):
        return self.logits if i in (0, "logits") else (_ for _ in ()).throw(KeyError(i))


def make_ref_model(lm: OracleLM, image_features: torch.Tensor, image_token: int, pad_token: int):
    """A stub ``self`` that satisfies everything forward() touches."""
    s = SimpleNamespace()
    s.config = SimpleNamespace(output_attentions=False, output_hidden_states=False, use_return_dict=True,
                               vision_feature_layer=-2, vision_feature_select_strategy="full",
                               image_token_index=image_token, ignore_index=-100)
    s.pad_token_id = pad_token
    for a in ("start_generation_pos", "is_first_generation", "image_features", "vision_uncert_dict",
              "ag_mask_ids", "image_logits"):
        setattr(s, a, None)
    for a in ("start_image_pos", "end_image_pos", "logits_mask_prob", "all_outputs", "int_array",
              "token_entropies", "token_ventropies", "masked_numbers"):
        setattr(s, a, [])
    s.get_input_embeddings = lambda: (lambda ids: lm.w["embed"][ids.clamp(max=lm.cfg.vocab - 1)])
    s.vision_tower = lambda pv, output_hidden_states=True: SimpleNamespace(
        hidden_states=[None, image_features, None])           # [-2] -> features
    s.multi_modal_projector = lambda x: x

    def language_model(attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                       **kw):
        h = lm.forward(inputs_embeds[0], past_key_values.oc, attention_mask[0], position_ids[0])
        return _RefOut(lm.logits(h)[None], past_key_values)

    s.language_model = language_model
    for name in ("_merge_input_ids_with_image_features", "get_image_features", "get_image_logits",
                 "get_topk_token_id", "get_overlap_image_tokens", "get_image_attention_mask",
                 "calculate_vision_uncertainty"):
        setattr(s, name, types.MethodType(getattr(RL, name), s))
    s.forward = types.MethodType(RL.forward, s)
    return s


def ref_generate(model, lm, input_ids, attention_mask, max_new_tokens, seed):
    """HF 4.44 greedy ``_sample`` restated around the reference's real forward()
    (empty DynamicCache up front, pixel_values every step,
    mask grows by one, cache <- outputs.past_key_values)."""
    torch.manual_seed(seed)
    cache = _RefCache(lm.cfg.n_layers)
    pixel_values = torch.zeros(1, 3, 4, 4)
    ids = input_ids.clone()
    rec = dict(tokens=[], masked_numbers=[], logits=[])
    cur, mask = ids, attention_mask.clone()
    for t in range(max_new_tokens):
        out = model.forward(input_ids=cur, pixel_values=pixel_values, attention_mask=mask,
                            past_key_values=cache, use_cache=True)
        nxt = torch.argmax(out.logits[:, -1, :], dim=-1)
        rec["tokens"].append(int(nxt))
        rec["logits"].append(out.logits[0, -1].clone())
        rec["masked_numbers"].append(list(model.masked_numbers))
        cache = out.past_key_values
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        cur = nxt[:, None]
        mask = torch.cat([mask, mask.new_ones(1, 1)], dim=-1)
    return rec


def gold_decode():
    cfg = LMConfig.tiny()
    n_vis, T, probs, seed = TINY_N_VIS, TINY_T, TINY_PROBS, TINY_SEED
    out = {}
    summary = {}
    for faithful in (False, True):
        tag = "bf16" if faithful else "fp32"
        w = init_lm_weights(cfg, seed=0, std=0.02, head_std=0.08)
        lm = OracleLM(cfg, w, faithful=faithful)
        prompt, feats, image_token = tiny_decode_inputs(cfg.vocab, cfg.d_model)
        model = make_ref_model(lm, feats, image_token, pad_token=-1)
        ref = ref_generate(model, lm, prompt, torch.ones_like(prompt), T, seed)
        assert model.start_image_pos == [5] and model.end_image_pos == [5 + n_vis - 1]
        # the merged prompt embedding sequence, as _merge builds it (models/llava.py:74-153)
        emb = lm.w["embed"][prompt[0].clamp(max=cfg.vocab - 1)]
        merged = torch.cat([emb[:5], feats[0], emb[6:]], dim=0)
        U = tiny_uniforms()
        tr = dropout_decode(lm, merged, 5, n_vis, T, probs, U, mode="cumulative", topk=5, keep_logits=True)
        assert tr.tokens == ref["tokens"], (tr.tokens, ref["tokens"])
        for t in range(T - 1):
            assert tr.masked_numbers[t] == ref["masked_numbers"][t + 1], t
        for t in range(T):
            # bf16-faithful mode: a different GEMM blocking can flip a bf16 rounding of an
            # intermediate, so logits agree only to a few bf16 ulps there
            close(tr.step_logits[t], ref["logits"][t], f"step logits {t}",
                  **(dict(rtol=0, atol=2e-2) if faithful else {}))
        close(tr.uncert["epis_uncert_per_token"], model.vision_uncert_dict["epis_uncert_per_token"], "epis",
              **(dict(rtol=1e-3, atol=0) if faithful else {}))
        eq(tr.topk_ids, model.image_features[1], "topk ids")
        # margins: how safe is each argmax on this path (for the GPU greedy-parity test)
        gaps = []
        for l in tr.step_logits:
            top2 = torch.topk(l, 2).values
            gaps.append(float(top2[0] - top2[1]))
        disagree = sum(len(set(m)) > 1 for m in tr.member_ids)
        print(f"[decode/{tag}] tokens {tr.tokens[:8]}... winners {tr.winner} disagree {disagree}/{T-1} "
              f"masked {tr.masked_numbers[0]} min top-2 gap {min(gaps):.2e}")
        out[f"{tag}.tokens"] = np.asarray(tr.tokens, dtype=np.int32)
        out[f"{tag}.y_init"] = np.asarray(tr.y_init, dtype=np.int32)
        out[f"{tag}.member_ids"] = np.asarray(tr.member_ids, dtype=np.int32)
        out[f"{tag}.winner"] = np.asarray(tr.winner, dtype=np.int32)
        out[f"{tag}.masked_numbers"] = np.asarray(tr.masked_numbers, dtype=np.int32)
        out[f"{tag}.epi"] = tr.uncert["epis_uncert_per_token"][0].numpy()
        out[f"{tag}.alea"] = tr.uncert["alea_uncert_per_token"][0].numpy()
        out[f"{tag}.topk_ids"] = tr.topk_ids[0].numpy().astype(np.int32)
        out[f"{tag}.dropped_bits"] = np.packbits(np.asarray(
            [[np.isin(np.arange(n_vis), np.asarray(d, dtype=np.int64)) for d in step] for step in tr.dropped]),
            axis=-1)
        out[f"{tag}.gaps"] = np.asarray(gaps, dtype=np.float32)
        summary[tag] = dict(tokens=tr.tokens, min_gap=min(gaps))
        # independent (NeXT) and mean-aggregation flavours: oracle-only fixtures
        tr_i = dropout_decode(lm, merged, 5, n_vis, 12, probs, U, mode="independent", topk=10)
        out[f"{tag}.independent.tokens"] = np.asarray(tr_i.tokens, dtype=np.int32)
        out[f"{tag}.independent.masked_numbers"] = np.asarray(tr_i.masked_numbers, dtype=np.int32)
        tr_m = dropout_decode(lm, merged, 5, n_vis, 12, probs, U, mode="cumulative", aggregate="mean")
        out[f"{tag}.mean.tokens"] = np.asarray(tr_m.tokens, dtype=np.int32)
    out["prompt_ids"] = prompt.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "decode_tiny.npz"), **out)
    json.dump(summary, open(os.path.join(HERE, "decode_tiny.json"), "w"), indent=1)


# ----------------------------------------------------------------------------------------
# 5. InstructBLIP flavour: the reference's real patched LM forward (models/instructblip.py:43-165)
#    on a stub self vs oracle.dropout_decode(mode="quantile", vote_on="hidden", topk=10)
# ----------------------------------------------------------------------------------------
from golden_cases import BLIP_N_VIS, BLIP_T, BLIP_VOCAB, blip_merged  # noqa: E402


def blip_inputs():
    cfg = LMConfig(2, 256, 4, 4, 64, 512, BLIP_VOCAB)
    w = init_lm_weights(cfg, seed=2, std=0.02, head_std=0.08)
    return cfg, w, blip_merged(cfg.d_model)


class _BlipModelOut:
    def __init__(self, hidden, cache):
        self.last_hidden_state, self.past_key_values = hidden, cache
        self.hidden_states, self.attentions = None, None

    def __getitem__(self, i):
        assert i == 0
        return self.last_hidden_state


def gold_decode_blip():
    cfg, w, merged = blip_inputs()
    out = {}
    for faithful in (False, True):
        tag = "bf16" if faithful else "fp32"
        lm = OracleLM(cfg, w, faithful=faithful)
        s = SimpleNamespace()
        for a in ("start_image_pos", "end_image_pos", "token_entropies", "token_ventropies", "logits_mask_prob"):
            setattr(s, a, [])
        s.start_generation_pos, s.image_features, s.first_generation = -1, None, True
        s.vision_uncert_dict, s.image_logits = None, None

        def model(input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                  inputs_embeds=None, **kw):
            x = inputs_embeds[0] if inputs_embeds is not None else lm.embed(input_ids[0])
            h = lm.forward(x, past_key_values.oc, attention_mask[0], position_ids[0])
            return _BlipModelOut(h[None], past_key_values)

        s.model = model
        s.lm_head = lambda h: lm.logits(h)
        for name in ("get_image_features", "get_image_logits", "get_topk_token_id", "get_overlap_image_tokens",
                     "get_image_attention_mask", "calculate_vision_uncertainty", "calculate_entropy_varentropy"):
            setattr(s, name, types.MethodType(getattr(RB, name), s))
        # the reference passes the image span through module globals (models/instructblip.py:23-27, 599-651)
        S0 = merged.shape[0]
        ref_blip.start_img_pos, ref_blip.end_img_pos = 0, BLIP_N_VIS - 1
        ref_blip.start_generation_pos, ref_blip.first_generation = S0, True
        cache = _RefCache(cfg.n_layers)
        toks, masked = [], []
        o = RB.forward(s, input_ids=None, attention_mask=torch.ones(1, S0, dtype=torch.long),
                       position_ids=torch.arange(S0)[None], past_key_values=cache, inputs_embeds=merged[None],
                       use_cache=True)
        tok = int(torch.argmax(o.logits[0, -1]))
        toks.append(tok)
        for t in range(BLIP_T - 1):
            S = cache.oc.seq_len()
            # a fresh all-ones mask every step: the HF-4.44 aliasing quirk (SURVEY 8a note) is not reproduced
            o = RB.forward(s, input_ids=torch.tensor([[tok]]), attention_mask=torch.ones(1, S + 1, dtype=torch.long),
                           position_ids=torch.tensor([[S]]), past_key_values=cache, use_cache=True)
            cache = o.past_key_values
            tok = int(torch.argmax(o.logits[0, -1]))
            toks.append(tok)
        tr = dropout_decode(lm, merged, 0, BLIP_N_VIS, BLIP_T, [0.3, 0.5, 0.7], None, mode="quantile", topk=10,
                            vote_on="hidden", keep_logits=True)
        assert tr.tokens == toks, (tr.tokens, toks)
        close(tr.uncert["epis_uncert_per_token"], s.vision_uncert_dict["epis_uncert_per_token"], "blip epis",
              **(dict(rtol=1e-3, atol=0) if faithful else {}))
        print(f"[decode_blip/{tag}] tokens {toks[:8]}... winners {tr.winner} masked {tr.masked_numbers[0]}")
        out[f"{tag}.tokens"] = np.asarray(tr.tokens, dtype=np.int32)
        out[f"{tag}.winner"] = np.asarray(tr.winner, dtype=np.int32)
        out[f"{tag}.member_ids"] = np.asarray(tr.member_ids, dtype=np.int32)
        out[f"{tag}.masked_numbers"] = np.asarray(tr.masked_numbers, dtype=np.int32)
        out[f"{tag}.epi"] = tr.uncert["epis_uncert_per_token"][0].numpy()
    np.savez_compressed(os.path.join(HERE, "decode_blip_tiny.npz"), **out)


def gold_epis_kl():
    """lowest_percent_kl_indices (models/llava.py:758-776; instructblip.py:559-578) and the 'epis_kl' mask method
    (instructblip.py:464-485) of the reference itself on seeded inputs."""
    out = {}
    for name, n_vis, V, start in (("llava", 576, 4096, 5), ("blip", 32, 1000, 0)):
        g = torch.Generator().manual_seed(77 + n_vis)
        image_logits = torch.randn(1, n_vis, V, generator=g) * 1.3
        logits = torch.randn(1, 1, V, generator=g) * 1.3
        epis = torch.rand(n_vis, generator=g) * 0.2 + 0.75
        u = torch.rand(n_vis, generator=g)
        ref_idx = RL.lowest_percent_kl_indices(SimpleNamespace(), image_logits, logits)
        eq(oracle.lowest_percent_kl_indices(image_logits, logits), ref_idx, f"epis_kl/{name} indices")
        eq(RB.lowest_percent_kl_indices(SimpleNamespace(), image_logits, logits), ref_idx, f"epis_kl/{name} blip twin")
        # the whole mask method: InstructBLIP's get_image_attention_mask(method="epis_kl") with rand_like pinned to u
        stub = SimpleNamespace(vision_uncert_dict={"epis_uncert_per_token": epis[None]}, image_logits=image_logits,
                               start_image_pos=[start], masked_numbers=[],
                               image_features=(None, torch.zeros(1, n_vis, 10, dtype=torch.long)))
        stub.lowest_percent_kl_indices = lambda il, l, percent=0.1: RB.lowest_percent_kl_indices(stub, il, l, percent)
        stub.get_overlap_image_tokens = lambda *a, **k: torch.zeros(0, dtype=torch.long)
        mask = torch.ones(1, start + n_vis + 12, dtype=torch.long)
        real_rand = torch.rand_like
        torch.rand_like = lambda t, *a, **k: u.to(t.dtype)
        try:
            got = RB.get_image_attention_mask(stub, logits, mask, method="epis_kl", prob=0.5)
        finally:
            torch.rand_like = real_rand
        dropped_ref = (got[0, start:start + n_vis] == 0)
        eq(oracle.sample_mask_epis_kl(epis, 0.5, u, image_logits, logits), dropped_ref, f"epis_kl/{name} mask")
        print(f"[epis_kl/{name}] {int(ref_idx.numel())} lowest-KL tokens restored, {int(dropped_ref.sum())} dropped")
        out[f"{name}.indices"] = ref_idx.numpy().astype(np.int64)
        out[f"{name}.dropped"] = dropped_ref.numpy()
    np.savez_compressed(os.path.join(HERE, "epis_kl.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["uncertainty", "masks", "aggregate", "decode", "decode_blip", "epis_kl"]
    with torch.no_grad():
        for wname in which:
            globals()["gold_" + wname]()
    print("golden fixtures written to", HERE)
