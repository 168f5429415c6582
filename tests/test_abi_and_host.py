"""C-ABI surface and host-side logic; no GPU needed (no compute calls)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dropoutdecoding_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "dropdec.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    from dropoutdecoding_b200 import _lib
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dropdec.h but not exported by libdropdec.so"
    assert lib.dd_version() == 100


def test_argument_validation_never_touches_the_gpu(lib):
    """Bad arguments return DD_ERR_INVALID_ARG with a message before any CUDA call."""
    probs = (C.c_double * 3)(0.3, 0.5, 0.7)
    one = C.c_void_p(16)     # never dereferenced
    rc = lib.dd_sample_masks(one, one, 5, one, probs, 0, None, 0, 0, None, 0, 1, 1, 576, 0, one, 18, one, None)
    assert rc == -1 and b"K=0" in lib.dd_last_error()
    rc = lib.dd_sample_masks(one, one, 5, one, probs, 3, None, 0, 0, None, 0, 1, 1, 576, 0, one, 4, one, None)
    assert rc == -1 and b"mask_words" in lib.dd_last_error()
    rc = lib.dd_decode_attention(one, one, one, one, one, one, one, None, 1, 3, 32, 32, 96, 100, 0, 0, 0.1, 1, one,
                                 one, 1 << 30, None)
    assert rc == -1 and b"D=96" in lib.dd_last_error()
    rc = lib.dd_decode_attention(one, one, one, one, one, one, one, None, 1, 16, 32, 8, 128, 100, 0, 0, 0.1, 1, one,
                                 one, 1 << 30, None)
    assert rc == -1 and b"query rows" in lib.dd_last_error()
    rc = lib.dd_decode_attention(one, one, one, one, one, one, one, None, 1, 3, 32, 32, 128, 100, 0, 0, 0.1, 4, one,
                                 one, 8, None)
    assert rc == -3 and b"workspace" in lib.dd_last_error()
    rc = lib.dd_skinny_gemm(one, one, None, 33, 4096, 4096, 0, one, None)
    assert rc == -1 and b"M=33" in lib.dd_last_error()
    rc = lib.dd_skinny_gemm(one, one, None, 3, 4096, 4100, 0, one, None)
    assert rc == -1 and b"K=4100" in lib.dd_last_error()
    rc = lib.dd_stream_gemm(one, one, None, 257, 4096, 4096, 0, one, one, 1 << 30, None)
    assert rc == -1 and b"M=257" in lib.dd_last_error()
    rc = lib.dd_stream_gemm(one, one, None, 64, 4096, 4100, 0, one, one, 1 << 30, None)
    assert rc == -1 and b"K=4100" in lib.dd_last_error()
    rc = lib.dd_stream_gemm(one, one, None, 64, 4096, 4096, 0, one, one, 64, None)
    assert rc == -3 and b"workspace" in lib.dd_last_error()
    assert lib.dd_stream_gemm_workspace_bytes(64, 4096, 4096, 0) >= 16384
    assert lib.dd_stream_gemm_workspace_bytes(300, 4096, 4096, 0) == 0
    rc = lib.dd_uncertainty_project(one, one, 1, 9000, 4096, 32064, 5, 0, one, one, one, one, one, 1 << 30, None)
    assert rc == -1 and b"n_vis=9000" in lib.dd_last_error()
    rc = lib.dd_hint_l2_prefetch(C.c_void_p(24), 64, None, 0)
    assert rc == -1 and b"16-byte aligned" in lib.dd_last_error()
    rc = lib.dd_hint_l2_prefetch(None, 64, None, 0)
    assert rc == -1 and b"null range" in lib.dd_last_error()
    assert lib.dd_hint_l2_prefetch(None, 0, None, 0) == 0
    rc = lib.dd_aggregate(one, 1, 17, 100, 100, 0, one, one, one, None, None)
    assert rc == -1 and b"K=17" in lib.dd_last_error()
    assert lib.dd_attention_workspace_bytes(2, 3, 32, 128, 4) == 2 * 3 * 32 * 4 * 130 * 4 + 2 * 32 * 4   # partials + tickets
    assert lib.dd_uncertainty_workspace_bytes(1, 576, 32064, 5) > 0


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the front end fails loudly instead of computing on the host."""
    from dropoutdecoding_b200 import ops
    with pytest.raises(ops.DropDecError, match="CUDA"):
        ops.aggregate(torch.zeros(1, 3, 100))
    with pytest.raises(ops.DropDecError, match="CUDA"):
        ops.uncertainty_project(torch.zeros(1, 8, 64, dtype=torch.bfloat16), torch.zeros(100, 64, dtype=torch.bfloat16))


def test_product_code_never_imports_the_oracle():
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); import dropoutdecoding_b200.ops, "
                          "dropoutdecoding_b200.engine, dropoutdecoding_b200.models.llava, models.llava, models.config;"
                          "print('oracle' in sys.modules)" % ROOT], capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stderr
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dropoutdecoding_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "import oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_settings_and_wrapper_api():
    """Same knobs / names as the reference (models/config.py:1-4, models/llava.py:22-72)."""
    import models.config as mc
    import models.llava as ml
    assert mc.settings == {"voting_numbers": [0.3, 0.5, 0.7], "use_avg": False, "use_random": [False]}
    assert ml.seed == 24
    for name in ("CustomLlavaForConditionalGeneration", "select_by_vote", "select_by_average"):
        assert hasattr(ml, name)
    cls = ml.CustomLlavaForConditionalGeneration
    for name in ("generate", "forward", "calculate_vision_uncertainty", "get_image_attention_mask",
                 "get_overlap_image_tokens", "get_topk_token_id", "get_image_logits"):
        assert callable(getattr(cls, name))


def test_wrapper_state_merge_and_loud_cpu_failure():
    import bench
    from dropoutdecoding_b200 import ops
    torch.Tensor.pin_memory = lambda self: self          # no driver in the CPU container
    m = bench.build_llava15_7b("cpu", tiny=True)
    for attr in ("start_generation_pos", "start_image_pos", "end_image_pos", "is_first_generation", "image_features",
                 "logits_mask_prob", "token_entropies", "token_ventropies", "vision_uncert_dict", "image_logits",
                 "masked_numbers"):
        assert hasattr(m, attr)
    ids, px = bench.synthetic_inputs(2, 1)
    with torch.no_grad():
        merged, start, n_vis = m._merge(ids, px.to(torch.bfloat16))
        assert merged.shape == (2, 593, 256) and (start, n_vis) == (5, 576)
        assert m.start_image_pos == [5] and m.end_image_pos == [580]          # models/llava.py:75-78
        # pre-expanded placeholders (transformers >= 4.45 processors) give the same merge
        ids2 = torch.cat([ids[:, :5], torch.full((2, 576), 32000), ids[:, 6:]], dim=1)
        merged2, start2, _ = m._merge(ids2, px.to(torch.bfloat16))
        assert start2 == 5 and torch.equal(merged, merged2)
        bad = ids.clone(); bad[0, 7] = 32000
        with pytest.raises(ValueError, match="image tokens"):
            m._merge(bad, px.to(torch.bfloat16))
        with pytest.raises(ops.DropDecError, match="no CPU path"):
            m.generate(input_ids=ids, pixel_values=px, max_new_tokens=4)


def test_calculate_vision_uncertainty_helper_matches_oracle():
    import bench
    import oracle
    m = bench.build_llava15_7b("cpu", tiny=True)
    logits = torch.randn(1, 40, 300, generator=torch.Generator().manual_seed(0))
    a, b = m.calculate_vision_uncertainty(logits), oracle.calculate_vision_uncertainty(logits)
    for k in b:
        assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-7), k


def test_device_map_and_dtype_of_the_reference_drivers_are_resolved_explicitly(tmp_path):
    """chair_test/chair_test.py:185-214 loads `Custom*.from_pretrained(path, torch_dtype=torch.float16,
    device_map="auto")`: device_map forms that name one device are honoured, maps over several devices are refused
    loudly (no silent single-device fallback), torch_dtype is passed through to transformers."""
    import bench
    from dropoutdecoding_b200.models._mixin import resolve_device_map
    from models.llava import CustomLlavaForConditionalGeneration       # the drop-in alias at the repo root
    assert resolve_device_map(None) is None
    assert resolve_device_map("auto") == torch.device("cuda:0") and resolve_device_map(1) == torch.device("cuda:1")
    assert resolve_device_map({"": 0}) == torch.device("cuda:0") and resolve_device_map("cuda:3") == torch.device("cuda:3")
    with pytest.raises(ValueError, match="2 devices"):
        resolve_device_map({"model.vision_tower": 0, "model.language_model": 1})
    with pytest.raises(ValueError):
        resolve_device_map(3.5)
    tiny = bench.build_llava15_7b("cpu", tiny=True)
    tiny.save_pretrained(tmp_path)
    m = CustomLlavaForConditionalGeneration.from_pretrained(tmp_path, torch_dtype=torch.float16)
    assert m.lm_head.weight.dtype == torch.float16 and m.lm_head.weight.device.type == "cpu"
    assert m.masked_numbers == [] and m.vision_uncert_dict is None          # state of models/llava.py:57-72 is there
    with pytest.raises(Exception, match="CUDA"):                              # no CPU path: loud, not a fallback
        m.decoder_weights()


def test_text_only_calls_fall_through_to_the_stock_model():
    """generate() / forward() without pixel_values are the parent class's (models/llava.py keeps the HF behaviour for
    text-only input); transformers' generate() chooses its kwargs from the wrapper's forward() signature."""
    import bench
    m = bench.build_llava15_7b("cpu", tiny=True)
    ids = torch.randint(3, 30000, (1, 5), generator=torch.Generator().manual_seed(0))
    out = m.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=3, do_sample=False, pad_token_id=2)
    assert out.shape == (1, 8) and torch.equal(out[:, :5], ids)
    assert m(input_ids=ids).logits.shape == (1, 5, 32064)
