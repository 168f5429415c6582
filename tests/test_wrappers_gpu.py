"""The reference-facing wrappers on the GPU with tiny random-init transformers models: generate(),
the reference-style forward() loop, the introspection attributes, EOS handling, and the three model
families (reference: models/llava.py, models/llavanext.py, models/instructblip.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def llava(cuda):
    import bench
    return bench.build_llava15_7b(cuda, tiny=True)


def _inputs(B):
    import bench
    return bench.synthetic_inputs(B, seed0=1)


def test_llava_generate_and_attributes(llava, cuda):
    from models.config import settings
    ids, px = _inputs(1)
    out = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=12, num_beams=1, pad_token_id=2)
    assert out.shape == (1, 18 + 12) and torch.equal(out[:, :18].cpu(), ids)
    u = llava.vision_uncert_dict
    assert set(u) == {"variance_per_token", "epis_uncert_per_token", "alea_uncert_per_token", "variance",
                      "epis_uncert", "alea_uncert"}                                   # models/llava.py:747-756
    assert u["epis_uncert_per_token"].shape == (1, 576) and u["epis_uncert"].shape == (1,)
    assert bool((u["epis_uncert_per_token"] > 0).all()) and bool((u["alea_uncert_per_token"] > 0).all())
    assert llava.image_features[1].shape == (1, 576, 5)
    assert llava.start_image_pos == [5] and llava.end_image_pos == [580] and llava.start_generation_pos == 593
    n = llava.masked_numbers
    assert len(n) == len(settings["voting_numbers"]) and n[0] <= n[1] <= n[2] <= 576   # cumulative masks
    # deterministic: same seed, same stream
    again = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=12)
    assert torch.equal(out, again)


def test_short_generations_pope_style(llava, cuda):
    """pope_test/pope_test.py:242 asks for ONE new token (yes/no): no ensemble step runs, the token is the argmax of
    the unmasked prefill (models/llava.py:382-388); two and three tokens are prefixes of the longer generation."""
    ids, px = _inputs(1)
    want = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=6)[0, 18:].tolist()
    for n in (1, 2, 3):
        out = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=n)
        assert out.shape == (1, 18 + n) and out[0, 18:].tolist() == want[:n]
    assert llava.vision_uncert_dict["epis_uncert_per_token"].shape == (1, 576)


def test_forward_loop_equals_generate(llava, cuda):
    """Driving the reference's per-token forward() by hand (prefill iff input_ids.shape[-1] != 1,
    models/llava.py:218-226) gives the tokens generate() gives."""
    ids, px = _inputs(1)
    want = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=8)[0, 18:].tolist()
    out = llava(input_ids=ids.to(cuda), pixel_values=px.to(cuda), attention_mask=torch.ones_like(ids).to(cuda))
    assert out.logits.shape == (1, 593, 32064) and out.logits.dtype == torch.float32
    assert llava.image_logits.shape == (1, 576, 32064)
    toks = [int(out.logits[0, -1].argmax())]
    cache = out.past_key_values
    for _ in range(7):
        out = llava(input_ids=torch.tensor([[toks[-1]]], device=cuda), pixel_values=px.to(cuda), past_key_values=cache)
        assert out.logits.shape == (1, 1, 32064)
        toks.append(int(out.logits[0, -1].argmax()))
        cache = out.past_key_values
    assert toks == want
    assert cache.get_seq_length() == 593 + 7


def test_batched_generate_matches_single_sequences(llava, cuda):
    """B > 1 (not possible in the reference) decodes each sequence as if it were alone, given the
    same per-sequence uniforms."""
    ids, px = _inputs(3)
    T = 6
    U = torch.rand(T - 1, 3, 3, 576, generator=torch.Generator().manual_seed(9))
    both = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=T, uniforms=U)
    for b in range(3):
        one = llava.generate(input_ids=ids[b:b + 1], pixel_values=px[b:b + 1], max_new_tokens=T,
                             uniforms=U[:, b:b + 1].contiguous())
        assert torch.equal(one[0], both[b])


def test_eos_stops_and_pads(llava, cuda):
    ids, px = _inputs(1)
    free = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=20)[0, 18:].tolist()
    eos = free[4]
    first = free.index(eos)
    out = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=20, eos_token_id=eos, pad_token_id=0)
    new = out[0, 18:].tolist()
    assert new[:first + 1] == free[:first + 1] and len(new) == first + 1
    # one sequence: the host sees every pass's input token, so decoding stops within one pass of EOS (not at the end
    # of a 16-step chunk)
    eng = next(iter(llava._engines.values()))
    assert eng.spec_active and first <= int(eng.step_ctr) <= first + 1


def test_settings_are_read_at_call_time(llava, cuda):
    """voting_numbers / use_avg are mutated by the drivers (chair_test/chair_test.py:163-175)."""
    from models.config import settings
    ids, px = _inputs(1)
    old = dict(settings)
    try:
        settings["voting_numbers"] = [0.5, 0.3]                          # CLI --voting-numbers 2
        a = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=5)
        assert len(llava.masked_numbers) == 2
        settings["voting_numbers"] = [0.3, 0.5, 0.7]
        settings["use_avg"] = True
        b = llava.generate(input_ids=ids, pixel_values=px, max_new_tokens=5)
        assert a.shape == b.shape == (1, 23)
    finally:
        settings.update(old)


def test_llava_next_wrapper(cuda):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaNextConfig
    from models.llavanext import CustomLlavaNextForConditionalGeneration
    from models.config import settings
    vc = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                          image_size=56, patch_size=14)
    tc = LlamaConfig(vocab_size=32064, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-5, max_position_embeddings=4096)
    cfg = LlavaNextConfig(vision_config=vc, text_config=tc, image_token_index=32000, vision_feature_layer=-2,
                          vision_feature_select_strategy="default", image_grid_pinpoints=[[112, 112]])
    torch.manual_seed(0)
    with torch.device(cuda):
        m = CustomLlavaNextForConditionalGeneration(cfg)
    m = m.to(cuda, torch.bfloat16).eval()
    g = torch.Generator().manual_seed(3)
    px = torch.randn(1, 5, 3, 56, 56, generator=g)                       # base + 2x2 anyres patches
    sizes = torch.tensor([[112, 112]])
    ids = torch.randint(3, 30000, (1, 12), generator=g)
    ids[0, 4] = 32000
    out = m.generate(input_ids=ids, pixel_values=px, image_sizes=sizes, max_new_tokens=6)
    n_vis = m.end_image_pos[0] - m.start_image_pos[0] + 1
    assert out.shape == (1, 18) and n_vis > 16 and m.image_features[1].shape == (1, n_vis, 10)   # top-10
    n = m.masked_numbers
    assert len(n) == 3 and max(n) <= n_vis
    settings["use_random"] = [True]                                       # epis_no_overlap, llavanext.py:547-550
    try:
        m.generate(input_ids=ids, pixel_values=px, image_sizes=sizes, max_new_tokens=4)
    finally:
        settings["use_random"] = [False]


def test_instructblip_wrapper(cuda):
    from transformers import InstructBlipConfig, InstructBlipQFormerConfig, InstructBlipVisionConfig, LlamaConfig
    from models.instructblip import CustomInstructBlipForConditionalGeneration, CustomLlamaForCausalLM
    vc = InstructBlipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                                  image_size=56, patch_size=14)
    qc = InstructBlipQFormerConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                                   encoder_hidden_size=64, vocab_size=1000)
    tc = LlamaConfig(vocab_size=32001, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=4, rms_norm_eps=1e-5, max_position_embeddings=2048,
                     bos_token_id=1, eos_token_id=2)
    cfg = InstructBlipConfig(vision_config=vc.to_dict(), qformer_config=qc.to_dict(), text_config=tc.to_dict(),
                             num_query_tokens=32)
    torch.manual_seed(0)
    with torch.device(cuda):
        m = CustomInstructBlipForConditionalGeneration(cfg)
    m = m.to(cuda, torch.bfloat16).eval()
    # HF initialises the Q-Former query tokens to zero: 32 identical visual tokens -> identical uncertainties
    # -> the quantile threshold drops everything.  Give the random-init model distinct queries.
    m.query_tokens.data.normal_(std=1.0, generator=torch.Generator(device=cuda).manual_seed(3))
    assert isinstance(m.language_model, CustomLlamaForCausalLM)
    g = torch.Generator().manual_seed(4)
    px = torch.randn(2, 3, 56, 56, generator=g)
    qids = torch.randint(0, 1000, (2, 7), generator=g)
    ids = torch.randint(3, 30000, (2, 9), generator=g)
    out = m.generate(pixel_values=px, qformer_input_ids=qids, qformer_attention_mask=torch.ones_like(qids),
                     input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=7, eos_token_id=None)
    assert out.shape[0] == 2 and out.shape[1] <= 8 and bool((out[:, 0] == 1).all())   # BOS first, instructblip.py:680-695
    assert m.start_image_pos == [0] and m.end_image_pos == [31] and m.start_generation_pos == 41
    assert m.image_features[1].shape == (2, 32, 10)
    assert m.masked_numbers[0] <= m.masked_numbers[1] <= m.masked_numbers[2] <= 22      # quantile masks: <= 10/16/22
    assert m.language_model.vision_uncert_dict is m.vision_uncert_dict
    n_new = out.shape[1] - 1                                   # a13 bookkeeping: one entry per generated token
    assert len(m.logits_mask_prob) == len(m.token_entropies) == len(m.token_ventropies) >= n_new >= 1
    assert all(0.0 < e < 15.0 for e in m.token_entropies)


def test_reference_driver_call_sequence(cuda, tmp_path):
    """The calls chair_test/chair_test.py makes (185-214: from_pretrained(path, torch_dtype=torch.float16,
    device_map="auto"); 163-175: settings[...] = ...; 341-354: generate(**inputs, max_new_tokens, num_beams=1,
    pad_token_id=eos) and slicing off the prompt) against the drop-in alias `models/` at the repo root.  The fp16
    checkpoint is converted to bf16 as a whole (one warning), so the stock text-only path still works afterwards."""
    import bench
    from models.config import settings
    from models.llava import CustomLlavaForConditionalGeneration
    bench.build_llava15_7b("cpu", tiny=True).save_pretrained(tmp_path)
    model = CustomLlavaForConditionalGeneration.from_pretrained(tmp_path, torch_dtype=torch.float16, device_map="auto")
    assert model.lm_head.weight.device == cuda and model.lm_head.weight.dtype == torch.float16
    old = dict(settings)
    try:
        settings['voting_numbers'] = [0.5, 0.3]           # --voting-numbers 2 (chair_test/chair_test.py:167-168)
        settings['use_avg'] = False
        ids, px = _inputs(1)
        inputs = {"input_ids": ids.to(cuda), "attention_mask": torch.ones_like(ids).to(cuda), "pixel_values": px.to(cuda).half()}
        with pytest.warns(UserWarning, match="bfloat16"):
            out = model.generate(**inputs, max_new_tokens=10, num_beams=1, pad_token_id=2)
        new = out[0][inputs["input_ids"].shape[1]:]      # chair_test.py:348-354 decodes only the new tokens
        assert out.shape[0] == 1 and 1 <= new.numel() <= 10
        assert len(model.masked_numbers) == 2
        assert {p.dtype for p in model.parameters()} == {torch.bfloat16}            # one dtype everywhere
        # the stock HF path (no pixel_values) still runs on the converted model
        txt = model.generate(input_ids=ids.to(cuda)[:, :5], max_new_tokens=3, do_sample=False, pad_token_id=2)
        assert txt.shape == (1, 8)
    finally:
        settings.clear(); settings.update(old)


def test_vision_tower_graph_replay_equals_eager(llava, cuda, monkeypatch):
    """The vision tower + projector replayed from a CUDA graph give the eager features bit for bit, for a second image
    through the same graph and for another batch size (a second graph)."""
    import bench
    with torch.no_grad():
        for B, seed in ((1, 1), (1, 7), (2, 3)):
            _, px = bench.synthetic_inputs(B, seed0=seed)
            px = px.to(cuda).to(llava.dtype)
            monkeypatch.setenv("DD_VISION_GRAPH", "0")
            want = llava._visual_features(px).clone()
            monkeypatch.setenv("DD_VISION_GRAPH", "1")
            got = llava._visual_features(px)
            assert got.shape == (B, 576, llava.config.text_config.hidden_size) and torch.equal(got, want)
    graphs = llava.__dict__["_vision_graphs"]
    assert len(graphs) >= 2 and all(e is not False for e in graphs.values()), "the tower was not captured"
