"""Pin the oracle's restated decoder (transformers 4.44 Llama, not vendored by the reference)
against the installed transformers LlamaForCausalLM: fp32, eager attention, same weights."""
import pytest
import torch

import oracle


@pytest.fixture(scope="module")
def pair():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = oracle.LMConfig(n_layers=2, d_model=128, n_heads=4, n_kv_heads=2, head_dim=32, d_ff=256, vocab=500,
                          rms_eps=1e-5, rope_theta=10000.0)
    w = oracle.init_lm_weights(cfg, seed=3, std=0.05)
    hf = LlamaForCausalLM(LlamaConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.d_model, intermediate_size=cfg.d_ff, num_hidden_layers=cfg.n_layers,
        num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_kv_heads, rms_norm_eps=cfg.rms_eps,
        rope_theta=cfg.rope_theta, max_position_embeddings=256, attn_implementation="eager",
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False)).float().eval()
    sd = {"model.embed_tokens.weight": w["embed"], "model.norm.weight": w["norm"], "lm_head.weight": w["lm_head"]}
    for l in range(cfg.n_layers):
        p, q = f"l{l}.", f"model.layers.{l}."
        sd.update({q + "input_layernorm.weight": w[p + "ln1"], q + "post_attention_layernorm.weight": w[p + "ln2"],
                   q + "self_attn.q_proj.weight": w[p + "wq"], q + "self_attn.k_proj.weight": w[p + "wk"],
                   q + "self_attn.v_proj.weight": w[p + "wv"], q + "self_attn.o_proj.weight": w[p + "wo"],
                   q + "mlp.gate_proj.weight": w[p + "wg"], q + "mlp.up_proj.weight": w[p + "wu"],
                   q + "mlp.down_proj.weight": w[p + "wd"]})
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not [m for m in missing if "rotary" not in m] and not unexpected
    return cfg, oracle.OracleLM(cfg, w, faithful=False), hf


def test_prefill_logits_match_hf(pair):
    cfg, lm, hf = pair
    ids = torch.randint(0, cfg.vocab, (1, 37), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        want = hf(input_ids=ids).logits[0]
        cache = oracle.OracleCache(cfg.n_layers)
        got = lm.logits(lm.forward(lm.embed(ids[0]), cache, torch.ones(37, dtype=torch.long), torch.arange(37)))
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()


def test_masked_decode_step_matches_hf(pair):
    """One cached single-token step whose 2-D attention mask hides some prefix positions -- the
    operation Dropout Decoding performs K times per token (models/llava.py:350-359)."""
    cfg, lm, hf = pair
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, cfg.vocab, (1, 30), generator=g)
    nxt = torch.randint(0, cfg.vocab, (1, 1), generator=g)
    mask = torch.ones(1, 31, dtype=torch.long)
    mask[0, 5:20] = (torch.rand(15, generator=g) > 0.5).long()
    with torch.no_grad():
        out = hf(input_ids=ids, use_cache=True)
        want = hf(input_ids=nxt, past_key_values=out.past_key_values, attention_mask=mask,
                  position_ids=torch.tensor([[30]])).logits[0, -1]
        cache = oracle.OracleCache(cfg.n_layers)
        lm.forward(lm.embed(ids[0]), cache, torch.ones(30, dtype=torch.long), torch.arange(30))
        got = lm.logits(lm.forward(lm.embed(nxt[0]), cache, mask[0], torch.tensor([30])))[0]
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()
    assert cache.seq_len() == 31


def test_faithful_mode_is_close_to_fp32(pair):
    cfg, lm, _ = pair
    lmf = oracle.OracleLM(cfg, lm.w, faithful=True)
    ids = torch.randint(0, cfg.vocab, (20,), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        a = lm.logits(lm.forward(lm.embed(ids), oracle.OracleCache(cfg.n_layers), torch.ones(20), torch.arange(20)))
        b = lmf.logits(lmf.forward(lmf.embed(ids), oracle.OracleCache(cfg.n_layers), torch.ones(20), torch.arange(20)))
    assert (a - b).abs().max() < 5e-2 and torch.equal(b, oracle.bf16r(b))
