"""Seeded input recipes shared by tests/golden/make_golden.py (which mints the expected
outputs with the reference's own code) and the tests (which regenerate the inputs)."""
import torch

# name, n_vis, d, V, h_scale, head_std, topk
UNCERT_CASES = [
    ("llava_tiny", 576, 256, 32064, 1.0, 0.08, 5),       # config 1 shapes
    ("blip_small", 32, 256, 32001, 1.0, 0.08, 10),       # InstructBLIP: 32 tokens, V=32001, k=10
    ("peaked", 200, 128, 4096, 4.0, 0.25, 5),            # sharp distributions, p < 1e-10 present
    ("ragged", 77, 64, 1000, 1.0, 0.3, 10),              # nothing a multiple of anything
]


def bf16r(x):
    return x.to(torch.bfloat16).to(torch.float32)


def uncert_inputs(name, n_vis, d, V, h_scale, head_std, seed=1234):
    g = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    h = bf16r(torch.randn(n_vis, d, generator=g) * h_scale)
    w = bf16r(torch.randn(V, d, generator=g) * head_std)
    return h, w


AGG_SEED, AGG_V, AGG_CASES = 7, 32064, 40


def aggregate_inputs():
    """Yields (K, logits [K, V]) with forced vote patterns: unanimous, 2-1, all different, 2-2 ties."""
    g = torch.Generator().manual_seed(AGG_SEED)
    for case in range(AGG_CASES):
        K = [1, 2, 3, 3, 4, 8][case % 6]
        l = torch.randn(K, AGG_V, generator=g)
        pat = case % 5
        if pat == 0:
            l[:, 123] += 20
        elif pat == 1 and K >= 3:
            l[0, 5] += 20; l[1:, 9] += 20
        elif pat == 2 and K >= 2:
            l[0, 5] += 20; l[1, 9] += 20           # 1-1(-1) tie: earliest member wins
        elif pat == 3 and K >= 4:
            l[0, 1] += 20; l[1, 2] += 20; l[2, 2] += 20; l[3, 1] += 20   # 2-2 tie -> id of member 0
        yield K, l


# ---- config 1 (BASELINE.json configs[0]): tiny LLaVA-1.5, one image, K=3, 32 new tokens -------------
TINY_N_VIS, TINY_T, TINY_PROBS, TINY_SEED, TINY_VIS_START = 576, 32, [0.3, 0.5, 0.7], 24, 5


def tiny_decode_inputs(cfg_vocab=32064, d_model=256):
    """prompt ids [1,18] (BOS first, <image> at index 5) and projected image features [1,576,d]."""
    g = torch.Generator().manual_seed(1)
    image_token = cfg_vocab - 1
    prompt = torch.randint(3, 30000, (1, 18), generator=g)
    prompt[0, 0] = 1
    prompt[0, TINY_VIS_START] = image_token
    feats = bf16r(torch.randn(1, TINY_N_VIS, d_model, generator=g))
    return prompt, feats, image_token


def tiny_uniforms():
    """The numbers K x 31 consecutive torch.rand_like(epis) calls draw after torch.manual_seed(24)
    (models/llava.py:16-18, 650)."""
    g = torch.Generator().manual_seed(TINY_SEED)
    return torch.rand(TINY_T - 1, len(TINY_PROBS), TINY_N_VIS, generator=g)


# ---- InstructBLIP flavour, tiny: 32 visual tokens first, V = 32001, quantile masks, hidden-state vote ----
BLIP_N_VIS, BLIP_T, BLIP_VOCAB = 32, 16, 32001


def blip_merged(d_model=256):
    g = torch.Generator().manual_seed(5)
    return bf16r(torch.randn(BLIP_N_VIS + 9, d_model, generator=g))   # 32 Q-Former tokens + 9 prompt embeds
