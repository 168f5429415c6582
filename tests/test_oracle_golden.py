"""The oracle against the committed golden fixtures (minted with the reference's own code by
tests/golden/make_golden.py).  CPU only; this is what keeps the checker honest on boxes where
/root/reference does not exist."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from golden_cases import (TINY_N_VIS, TINY_PROBS, TINY_T, TINY_VIS_START, UNCERT_CASES, aggregate_inputs,
                          tiny_decode_inputs, tiny_uniforms, uncert_inputs)


@pytest.mark.parametrize("case", UNCERT_CASES, ids=[c[0] for c in UNCERT_CASES])
def test_uncertainty(case, golden_dir):
    name, n_vis, d, V, hs, ws, k = case
    z = np.load(os.path.join(golden_dir, "uncertainty.npz"))
    h, w = uncert_inputs(name, n_vis, d, V, hs, ws)
    logits = (h @ w.T)[None]
    u = oracle.calculate_vision_uncertainty(logits)
    np.testing.assert_allclose(u["epis_uncert_per_token"][0].numpy(), z[f"{name}.epi"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(u["alea_uncert_per_token"][0].numpy(), z[f"{name}.alea"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(u["variance_per_token"][0].numpy(), z[f"{name}.var"], rtol=1e-4, atol=1e-12)
    _, ids = oracle.topk_token_ids(logits, k)
    assert np.array_equal(ids[0].numpy(), z[f"{name}.topk_ids"])


def test_masks_all_modes(golden_dir):
    z = np.load(os.path.join(golden_dir, "masks.npz"))
    for c in json.loads(bytes(z["meta"]).decode()):
        name, steps = c["name"], c["steps"]
        epis = torch.from_numpy(z[f"{name}.epis"])
        ids = torch.from_numpy(z[f"{name}.topk_ids"]).long()[None]
        U = torch.from_numpy(z[f"{name}.uniforms"])
        probs = z[f"{name}.probs"].tolist()
        for t in range(steps):
            prelim = torch.full((1, 1, c["V"]), -1.0)
            prelim[0, 0, int(z[f"{name}.y_init"][t])] = 1.0
            rel = oracle.relevant_visual_tokens(prelim, ids)
            d = None
            for m, p in enumerate(probs):
                d = oracle.sample_mask_epis(epis, p, U[t, m], rel, d)
                assert np.array_equal(d.numpy(), z[f"{name}.cumulative"][t, m])
                assert np.array_equal(oracle.sample_mask_epis(epis, p, U[t, m], rel).numpy(),
                                      z[f"{name}.independent"][t, m])
                assert np.array_equal(oracle.sample_mask_epis(epis, p, U[t, m], None).numpy(),
                                      z[f"{name}.no_overlap"][t, m])
                assert np.array_equal(oracle.sample_mask_quantile(epis, p, rel).numpy(), z[f"{name}.quantile"][t, m])


def test_quantile_mask_counts_match_survey():
    """SURVEY 8a-a9': 10/16/22 of 32 tokens dropped for 0.3/0.5/0.7 when nothing is relevant."""
    g = torch.Generator().manual_seed(0)
    epis = torch.rand(32, generator=g)
    assert [int(oracle.sample_mask_quantile(epis, p, None).sum()) for p in (0.3, 0.5, 0.7)] == [10, 16, 22]


def test_aggregate(golden_dir):
    gold = np.load(os.path.join(golden_dir, "aggregate.npz"))["cases"]
    for (K, l), (gK, gidx, gtok, gavg) in zip(aggregate_inputs(), gold):
        idx, tok = oracle.select_by_vote([l[m] for m in range(K)])
        assert (K, idx, tok) == (gK, gidx, gtok)
        assert int(torch.argmax(oracle.select_by_average([l[m] for m in range(K)]))) == gavg


def test_vote_tie_goes_to_earliest_member():
    l = torch.zeros(4, 10)
    l[0, 3] = l[1, 7] = l[2, 7] = l[3, 3] = 1.0
    assert oracle.select_by_vote(list(l)) == (0, 3)


@pytest.mark.parametrize("tag,faithful", [("fp32", False), ("bf16", True)])
def test_decode_tiny_config1(tag, faithful, golden_dir):
    """BASELINE.json configs[0]: tiny LLaVA-1.5, 576 visual tokens, K=3, 32 new tokens, on CPU."""
    z = np.load(os.path.join(golden_dir, "decode_tiny.npz"))
    cfg = oracle.LMConfig.tiny()
    lm = oracle.OracleLM(cfg, oracle.init_lm_weights(cfg, seed=0, std=0.02, head_std=0.08), faithful=faithful)
    prompt, feats, _ = tiny_decode_inputs(cfg.vocab, cfg.d_model)
    assert np.array_equal(prompt.numpy(), z["prompt_ids"])
    emb = lm.w["embed"][prompt[0].clamp(max=cfg.vocab - 1)]
    merged = torch.cat([emb[:TINY_VIS_START], feats[0], emb[TINY_VIS_START + 1:]], dim=0)
    with torch.no_grad():
        tr = oracle.dropout_decode(lm, merged, TINY_VIS_START, TINY_N_VIS, TINY_T, TINY_PROBS, tiny_uniforms())
    assert tr.tokens == z[f"{tag}.tokens"].tolist()
    assert tr.y_init == z[f"{tag}.y_init"].tolist()
    assert tr.member_ids == z[f"{tag}.member_ids"].tolist()
    assert tr.winner == z[f"{tag}.winner"].tolist()
    assert tr.masked_numbers == z[f"{tag}.masked_numbers"].tolist()
    bits = np.unpackbits(z[f"{tag}.dropped_bits"], axis=-1)[..., :TINY_N_VIS].astype(bool)
    for s, step in enumerate(tr.dropped):
        for m, d in enumerate(step):
            assert np.nonzero(bits[s, m])[0].tolist() == d
    # cumulative masks only grow from member to member (models/llava.py:344 is commented out)
    assert all(a <= b <= c for a, b, c in tr.masked_numbers)
    np.testing.assert_allclose(tr.uncert["epis_uncert_per_token"][0].numpy(), z[f"{tag}.epi"], rtol=1e-4)


def test_epis_kl_restatement_and_wrapper_helper_match_the_reference(golden_dir):
    """`lowest_percent_kl_indices` / the 'epis_kl' mask (models/llava.py:758-776, instructblip.py:464-485, 559-578): the
    oracle restatement and the wrapper's public helper against what the reference's own functions returned
    (tests/golden/make_golden.py epis_kl)."""
    import oracle
    from dropoutdecoding_b200.models._mixin import DropoutDecodingMixin
    z = np.load(os.path.join(golden_dir, "epis_kl.npz"))
    for name, n_vis, V in (("llava", 576, 4096), ("blip", 32, 1000)):
        g = torch.Generator().manual_seed(77 + n_vis)
        image_logits = torch.randn(1, n_vis, V, generator=g) * 1.3
        logits = torch.randn(1, 1, V, generator=g) * 1.3
        epis = torch.rand(n_vis, generator=g) * 0.2 + 0.75
        u = torch.rand(n_vis, generator=g)
        want = z[f"{name}.indices"]
        assert oracle.lowest_percent_kl_indices(image_logits, logits).numpy().tolist() == want.tolist()
        got = DropoutDecodingMixin.lowest_percent_kl_indices(None, image_logits, logits)
        assert sorted(got.tolist()) == sorted(want.tolist())
        assert np.array_equal(oracle.sample_mask_epis_kl(epis, 0.5, u, image_logits, logits).numpy(), z[f"{name}.dropped"])
