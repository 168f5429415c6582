"""Parity of every CUDA kernel, called through the C ABI, against the golden fixtures minted
from the reference's own code (tests/golden/*.npz) and against fp32 restatements.

Bars (BASELINE.json north_star): mask index sets / votes / argmax bit-exact; per-token epistemic
uncertainty within 1e-3 relative; bf16 tensors within bf16 rounding of an fp32 computation.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from golden_cases import UNCERT_CASES, aggregate_inputs, uncert_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from dropoutdecoding_b200 import ops as _ops
    return _ops


def unpack_bits(words: torch.Tensor, n: int) -> np.ndarray:
    """int32 words [..., W] -> bool [..., n] (bit i of word i//32)."""
    w = words.cpu().numpy().astype(np.uint32)
    bits = ((w[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return bits.reshape(*w.shape[:-1], -1)[..., :n]


# ---------------------------------------------------------------------------------------------
# (2) mask sampler vs the reference's get_image_attention_mask (golden)
# ---------------------------------------------------------------------------------------------
def test_sampler_matches_reference_masks(ops, cuda, golden_dir):
    z = np.load(os.path.join(golden_dir, "masks.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    for c in meta:
        name, n_vis, steps = c["name"], c["n_vis"], c["steps"]
        epis = torch.from_numpy(z[f"{name}.epis"]).to(cuda)[None].contiguous()
        ids = torch.from_numpy(z[f"{name}.topk_ids"]).to(cuda)[None].contiguous()
        probs = z[f"{name}.probs"].tolist()
        for t in range(steps):
            y = torch.tensor([int(z[f"{name}.y_init"][t])], dtype=torch.int32, device=cuda)
            U = torch.from_numpy(z[f"{name}.uniforms"][t]).to(cuda)[None].contiguous()
            for mode, key, keep in (("cumulative", "cumulative", True), ("independent", "independent", True),
                                    ("independent", "no_overlap", False), ("quantile", "quantile", True)):
                bm, cnt = ops.sample_masks(epis, ids, y, probs, mode=mode, keep_relevant=keep, uniforms=U)
                got = unpack_bits(bm[0], n_vis)
                want = z[f"{name}.{key}"][t]
                assert np.array_equal(got, want), f"{name} step {t} mode {key}: mask index sets differ"
                assert cnt[0].tolist() == want.sum(-1).tolist()


def test_sampler_batched_and_padded_words(ops, cuda, golden_dir):
    """B > 1 with per-sequence y_init, and a mask_words larger than needed (padding must be zero)."""
    z = np.load(os.path.join(golden_dir, "masks.npz"))
    name, n_vis = "small", 33
    epis = torch.from_numpy(z[f"{name}.epis"]).to(cuda)
    ids = torch.from_numpy(z[f"{name}.topk_ids"]).to(cuda)
    B = 4
    y = torch.from_numpy(z[f"{name}.y_init"][:B].astype(np.int32)).to(cuda)
    U = torch.from_numpy(z[f"{name}.uniforms"][:B]).to(cuda).contiguous()
    out = torch.full((B, 3, 4), -1, dtype=torch.int32, device=cuda)
    bm, _ = ops.sample_masks(epis[None].repeat(B, 1).contiguous(), ids[None].repeat(B, 1, 1).contiguous(), y,
                             z[f"{name}.probs"].tolist(), mode="cumulative", uniforms=U, out=out)
    got = unpack_bits(bm, 128)
    for b in range(B):
        assert np.array_equal(got[b][:, :n_vis], z[f"{name}.cumulative"][b])
    assert not got[:, :, n_vis:].any()


def test_philox_stream_equals_torch_cuda_rand(ops, cuda):
    """The in-kernel Philox reproduces torch.rand on the device, so drop-in mode
    draws the numbers the reference's rand_like would draw on a GPU."""
    for n in (576, 32, 2880):
        torch.cuda.manual_seed(24)
        a = torch.rand(n, device=cuda)
        b = torch.rand(n, device=cuda)
        assert torch.equal(ops.philox_uniform(24, 0, n), a)
        assert torch.equal(ops.philox_uniform(24, 4, n), b)


def test_sampler_philox_mode_consistent(ops, cuda):
    """Philox mode == uniforms mode fed with the same Philox numbers, step counter included."""
    n_vis, K = 576, 3
    g = torch.Generator().manual_seed(3)
    epi = (torch.rand(1, n_vis, generator=g) * 0.3 + 0.6).to(cuda)
    ids = torch.randint(0, 32064, (1, n_vis, 5), generator=g, dtype=torch.int32).to(cuda)
    y = torch.tensor([17], dtype=torch.int32, device=cuda)
    probs = [0.3, 0.5, 0.7]
    step = torch.tensor([5], dtype=torch.int32, device=cuda)
    bm, _ = ops.sample_masks(epi, ids, y, probs, uniforms=None, seed=24, offset=8, step=step)
    U = torch.stack([ops.philox_uniform(24, 8 + 4 * (5 * K + m), n_vis) for m in range(K)])[None].contiguous()
    bm2, _ = ops.sample_masks(epi, ids, y, probs, uniforms=U)
    assert torch.equal(bm, bm2)


# ---------------------------------------------------------------------------------------------
# (4) aggregation vs select_by_vote / select_by_average (golden)
# ---------------------------------------------------------------------------------------------
def test_aggregate_matches_reference(ops, cuda, golden_dir):
    gold = np.load(os.path.join(golden_dir, "aggregate.npz"))["cases"]
    for (K, l), (gK, gidx, gtok, gavg) in zip(aggregate_inputs(), gold):
        assert K == gK
        lg = l.to(cuda)[None].contiguous()
        mids, win, tok, out = ops.aggregate(lg, "vote", want_logits=True)
        assert mids[0].tolist() == torch.argmax(l, dim=-1).tolist()
        assert int(win[0]) == gidx and int(tok[0]) == gtok
        assert torch.equal(out[0].cpu(), l[gidx])
        _, win_m, tok_m, out_m = ops.aggregate(lg, "mean")
        want = torch.tensor(l.numpy().mean(axis=0))
        assert torch.equal(out_m[0].cpu(), want), "mean logits must equal numpy's fp32 mean bit for bit"
        assert int(tok_m[0]) == gavg and int(win_m[0]) == 0


def test_token_stats_match_reference_bookkeeping(ops, cuda):
    """a13: 1/max(logits), base-2 entropy and varentropy of the chosen row == calculate_entropy_varentropy
    (models/llava.py:687-708) and the logits_mask_prob expression (models/llavanext.py:584)."""
    from oracle import entropy_varentropy
    g = torch.Generator().manual_seed(9)
    B, K, V, ld = 3, 4, 32001, 32016
    logits = torch.randn(B, K, ld, generator=g) * 1.5
    logits[1, 2, 77] = 14.0                                  # one sharply peaked row
    member = torch.tensor([1, 2, 0], dtype=torch.int32)
    stats = torch.zeros(5, B, 3, device=cuda)
    slot = torch.tensor([2], dtype=torch.int32, device=cuda)
    ops.token_stats(logits.to(cuda), V, stats, member=member.to(cuda), slot=slot, slot_offset=1)
    got = stats.cpu()
    assert torch.count_nonzero(got[[0, 1, 2, 4]]) == 0       # only slot 2 + 1 was written
    for b in range(B):
        row = logits[b, member[b], :V]
        en, ven = entropy_varentropy(row)
        want = torch.tensor([1.0 / row.max().item(), en, ven])
        assert torch.allclose(got[3, b], want, rtol=2e-4, atol=1e-5), (b, got[3, b], want)


def test_aggregate_batched_unaligned_vocab(ops, cuda):
    g = torch.Generator().manual_seed(5)
    l = torch.randn(5, 3, 32001, generator=g)
    l[2, :, 77] = 9.0
    l[2, :, 31] = 9.0            # exact tie inside a row: first index wins, like torch.argmax
    mids, win, tok, _ = ops.aggregate(l.to(cuda))
    assert torch.equal(mids.cpu().long(), torch.argmax(l, dim=-1))
    assert tok[2].item() == 31


@pytest.mark.parametrize("B,K,V,ld", [(3, 2, 7, 7), (1, 16, 5, 12), (70, 4, 1000, 1000), (2, 3, 32064, 32064)])
@pytest.mark.parametrize("mode", ["vote", "mean"])
def test_aggregate_cluster_slices_edges(ops, cuda, B, K, V, ld, mode):
    """The 8-CTA cluster per sequence: vocabularies smaller than the cluster (empty slices), 16 members, padded rows,
    more sequences than SMs, ties that straddle two slices, the winner's row / the mean gathered from all slices, and
    run-to-run bit equality."""
    g = torch.Generator().manual_seed(B * 131 + K * 17 + V)
    full = torch.randn(B, K, ld, generator=g)
    full[..., V:] = 50.0                                   # padding columns must never win
    if V >= 1000:
        per = -(-(V // 4 if V % 4 == 0 and ld % 4 == 0 else V) // 8) * (4 if V % 4 == 0 and ld % 4 == 0 else 1)
        full[0, 0, per - 1] = full[0, 0, per] = 20.0       # the same maximum on both sides of a slice boundary
    dev = full.to(cuda)
    mids, win, tok, out = ops.aggregate(dev, mode, want_logits=True, vocab=V)
    l = full[..., :V]
    assert torch.equal(mids.cpu().long(), torch.argmax(l, dim=-1))
    if V >= 1000:
        assert mids[0, 0].item() == per - 1
    if mode == "vote":
        for b in range(B):
            ids = mids[b].tolist()
            best = max(ids, key=lambda t: (ids.count(t), -ids.index(t)))
            assert tok[b].item() == best and win[b].item() == ids.index(best)
            assert torch.equal(out[b].cpu(), l[b, win[b].item()])
    else:
        mean = l[:, 0].clone()
        for m in range(1, K):
            mean = mean + l[:, m]
        mean = mean / K
        assert torch.equal(out.cpu(), mean) and torch.equal(tok.cpu().long(), torch.argmax(mean, dim=-1))
    again = ops.aggregate(dev, mode, want_logits=True, vocab=V)
    assert all(torch.equal(x, y) for x, y in zip((mids, win, tok, out), again))


# ---------------------------------------------------------------------------------------------
# (3) decode attention vs an fp32 restatement of K+1 masked single-query SDPA calls
# ---------------------------------------------------------------------------------------------
def ref_attention(q, k_new, v_new, kc, vc, seq_lens, vis_start, dropped, scale):
    """All fp32 on CPU.  dropped: bool [B, M, n_vis] or None."""
    B, M, H, D = q.shape
    H_kv = kc.shape[1]
    G = H // H_kv
    out = torch.zeros(B, M, H, D)
    for b in range(B):
        S = int(seq_lens[b])
        for m in range(M):
            keep = torch.ones(S + 1, dtype=torch.bool)
            if dropped is not None:
                n_vis = dropped.shape[-1]
                keep[vis_start[b]:vis_start[b] + n_vis] = ~dropped[b, m]
            for h in range(H):
                K = torch.cat([kc[b, h // G, :S], k_new[b, m, h // G][None]], 0)
                V = torch.cat([vc[b, h // G, :S], v_new[b, m, h // G][None]], 0)
                s = (K @ q[b, m, h]) * scale
                s = s.masked_fill(~keep, float("-inf"))
                out[b, m, h] = torch.softmax(s, -1) @ V
    return out


@pytest.mark.parametrize("B,M,H,H_kv,D,S_max,lens,n_vis,splits", [
    (1, 3, 32, 32, 128, 700, [593], 576, 1),          # LLaVA-1.5-7B heads, K=3
    (1, 3, 32, 32, 128, 700, [650], 576, 9),          # split-KV
    (1, 1, 32, 32, 128, 700, [593], 0, 4),            # preliminary pass: M=1, no mask
    (3, 3, 8, 8, 128, 300, [57, 200, 64], 40, 3),     # ragged batch, short sequences, empty splits
    (2, 8, 32, 8, 128, 400, [333, 290], 250, 2),      # Mistral GQA, K=8 -> 32 query rows (2 row tiles)
    (2, 3, 4, 4, 64, 700, [593, 624], 576, 2),        # tiny config: D=64
    (1, 4, 4, 2, 64, 128, [1], 1, 1),                 # single cached key
])
def test_decode_attention(ops, cuda, B, M, H, H_kv, D, S_max, lens, n_vis, splits):
    g = torch.Generator().manual_seed(B * 1000 + M * 100 + D + splits)
    bf = lambda *s: (torch.randn(*s, generator=g)).to(torch.bfloat16)
    q, k_new, v_new = bf(B, M, H, D), bf(B, M, H_kv, D), bf(B, M, H_kv, D)
    kc, vc = bf(B, H_kv, S_max, D), bf(B, H_kv, S_max, D)
    seq_lens = torch.tensor(lens, dtype=torch.int32)
    scale = 1.0 / math.sqrt(D)
    if n_vis > 0:
        vis_start = torch.tensor([min(5, max(0, l - n_vis)) for l in lens], dtype=torch.int32)
        dropped = torch.rand(B, M, n_vis, generator=g) < 0.5
        dropped[0, 0, :] = True                      # a member that hides the whole image
        words = ops.mask_words(n_vis)
        bits = np.zeros((B, M, words * 32), dtype=bool)
        bits[..., :n_vis] = dropped.numpy()
        packed = (bits.reshape(B, M, words, 32) * (1 << np.arange(32, dtype=np.uint64))).sum(-1).astype(np.uint32)
        bm = torch.from_numpy(packed.view(np.int32)).to(cuda)
        vs = vis_start.to(cuda)
    else:
        vis_start, dropped, bm, vs = torch.zeros(B, dtype=torch.int32), None, None, None
    out = ops.decode_attention(q.to(cuda), k_new.to(cuda), v_new.to(cuda), kc.to(cuda), vc.to(cuda),
                               seq_lens.to(cuda), vs, bm, n_vis, scale, num_splits=splits)
    want = ref_attention(q.float(), k_new.float(), v_new.float(), kc.float(), vc.float(), seq_lens, vis_start,
                         dropped, scale)
    err = (out.float().cpu() - want).abs().max().item()
    # output is rounded to bf16 (2^-9 relative); everything upstream is fp32-accurate
    assert err <= 2 ** -8 * want.abs().max().item() + 1e-6, f"max abs err {err}"


def test_decode_attention_member_windows_and_waiting_sequences(ops, cuda):
    """dd_decode_attention_active: 9 contexts x 4 query heads per kv head (36 query rows > the kernel's 32) served as
    member windows [0, 8) and [8, 9) equal the restated attention; a sequence flagged inactive keeps its output
    rows; per-sequence step variants of the bookkeeping kernels."""
    B, M, H, H_kv, D, S_max, n_vis, splits = 3, 9, 8, 2, 128, 200, 96, 2
    lens = [150, 97, 192]
    g = torch.Generator().manual_seed(77)
    bf = lambda *s: (torch.randn(*s, generator=g)).to(torch.bfloat16)
    q, k_new, v_new = bf(B, M, H, D), bf(B, M, H_kv, D), bf(B, M, H_kv, D)
    kc, vc = bf(B, H_kv, S_max, D), bf(B, H_kv, S_max, D)
    seq_lens = torch.tensor(lens, dtype=torch.int32)
    vis_start = torch.tensor([1, 0, 5], dtype=torch.int32)
    dropped = torch.rand(B, M, n_vis, generator=g) < 0.5
    dropped[:, 0] = False                                  # member 0 = the unmasked context
    words = ops.mask_words(n_vis)
    bits = np.zeros((B, M, words * 32), dtype=bool)
    bits[..., :n_vis] = dropped.numpy()
    packed = (bits.reshape(B, M, words, 32) * (1 << np.arange(32, dtype=np.uint64))).sum(-1).astype(np.uint32)
    bm = torch.from_numpy(packed.view(np.int32)).to(cuda)
    scale = 1.0 / math.sqrt(D)
    out = torch.full((B, M, H, D), 7.0, dtype=torch.bfloat16, device=cuda)
    active = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
    args = (q.to(cuda), k_new.to(cuda), v_new.to(cuda), kc.to(cuda), vc.to(cuda), seq_lens.to(cuda), vis_start.to(cuda),
            bm, n_vis, scale)
    ws = ops.attention_workspace(B, M, H, D, splits, cuda)
    for m0, cnt in ((0, 8), (8, 1)):
        ops.decode_attention(*args, num_splits=splits, out=out, workspace=ws, seq_active=active, member_offset=m0,
                             members=cnt)
    want = ref_attention(q.float(), k_new.float(), v_new.float(), kc.float(), vc.float(), seq_lens, vis_start, dropped,
                         scale)
    got = out.float().cpu()
    for b in (0, 2):
        err = (got[b] - want[b]).abs().max().item()
        assert err <= 2 ** -8 * want.abs().max().item() + 1e-6, f"sequence {b}: max abs err {err}"
    assert bool((got[1] == 7.0).all())                     # the waiting sequence was not touched
    with pytest.raises(ops.DropDecError):                  # 9 members x 4 heads in one call: too many query rows
        ops.decode_attention(*args, num_splits=splits, out=out, workspace=ws)
    # per-sequence progress
    tok = torch.tensor([11, 22, 33], dtype=torch.int32, device=cuda)
    adv = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
    hist = torch.zeros(B, 6, dtype=torch.int32, device=cuda)
    seq = torch.tensor([5, 9, 2], dtype=torch.int32, device=cuda)
    steps = torch.tensor([0, 3, 4], dtype=torch.int32, device=cuda)
    ops.step_advance_seq(tok, adv, hist, seq, steps)
    assert seq.tolist() == [6, 9, 3] and steps.tolist() == [1, 3, 5]
    assert hist[0].tolist() == [0, 11, 0, 0, 0, 0] and hist[1].sum().item() == 0 and hist[2].tolist() == [0, 0, 0, 0, 0, 33]


def test_commit_winner_and_step_advance(ops, cuda):
    L, B, M, H_kv, D, S_max = 3, 2, 3, 4, 64, 16
    g = torch.Generator().manual_seed(0)
    kn = torch.randn(L, B, M, H_kv, D, generator=g).to(torch.bfloat16).to(cuda)
    vn = torch.randn(L, B, M, H_kv, D, generator=g).to(torch.bfloat16).to(cuda)
    kc = torch.zeros(L, B, H_kv, S_max, D, dtype=torch.bfloat16, device=cuda)
    vc = torch.zeros_like(kc)
    winner = torch.tensor([2, 0], dtype=torch.int32, device=cuda)
    seq = torch.tensor([5, 9], dtype=torch.int32, device=cuda)
    ops.commit_winner(kn, vn, winner, seq, kc, vc)
    for b, (w, s) in enumerate([(2, 5), (0, 9)]):
        assert torch.equal(kc[:, b, :, s], kn[:, b, w]) and torch.equal(vc[:, b, :, s], vn[:, b, w])
    assert kc.float().abs().sum(dim=(0, 2, 4))[0].count_nonzero() == 1   # only one row per sequence written
    tok = torch.tensor([11, 22], dtype=torch.int32, device=cuda)
    hist = torch.zeros(B, 4, dtype=torch.int32, device=cuda)
    step = torch.tensor([1], dtype=torch.int32, device=cuda)
    ops.step_advance(tok, hist, seq, step)
    assert seq.tolist() == [6, 10] and step.item() == 2 and hist[:, 2].tolist() == [11, 22]


# ---------------------------------------------------------------------------------------------
# decoder support kernels vs fp32 restatements with the same bf16 rounding points
# ---------------------------------------------------------------------------------------------
def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K", [(1, 4096, 4096), (3, 4096, 4096), (4, 12288, 4096), (8, 4096, 11008),
                                   (12, 256, 256), (32, 512, 256), (3, 32064, 256), (17, 1024, 512)])
def test_skinny_gemm_all_epilogues(ops, cuda, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = r16(torch.randn(M, K, generator=g))
    w = r16(torch.randn(N, K, generator=g) / math.sqrt(K))
    res = r16(torch.randn(M, N, generator=g))
    acc = (x.double() @ w.double().T).float()
    xg, wg, rg = x.to(cuda).bfloat16(), w.to(cuda).bfloat16(), res.to(cuda).bfloat16()
    tol = 2 ** -7        # one bf16 ulp of the largest value class; fp32 accumulation order is free

    def close(got, want):
        assert (got.float().cpu() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())

    close(ops.skinny_gemm(xg, wg), r16(acc))
    close(ops.skinny_gemm(xg, wg, ops.EPI_RESIDUAL, residual=rg), r16(r16(acc) + res))
    y32 = ops.skinny_gemm(xg, wg, ops.EPI_F32)
    assert y32.dtype == torch.float32
    close(y32, r16(acc))
    if N % 32 == 0:
        gate, up = r16(acc[:, :N // 2]), r16(acc[:, N // 2:])
        close(ops.skinny_gemm(xg, wg, ops.EPI_SWIGLU), r16(r16(torch.nn.functional.silu(gate)) * up))


def test_skinny_gemm_exact_small_integers(ops, cuda):
    """Integer-valued inputs make every product and partial sum exact: any lane / k-permutation
    mistake shows up as a wrong integer, not as a tolerance question."""
    g = torch.Generator().manual_seed(1)
    M, N, K = 5, 64, 512
    x = torch.randint(-3, 4, (M, K), generator=g).float()
    w = torch.randint(-3, 4, (N, K), generator=g).float()
    y = ops.skinny_gemm(x.to(cuda).bfloat16(), w.to(cuda).bfloat16(), ops.EPI_F32)
    assert torch.equal(y.cpu(), x @ w.T)


@pytest.mark.parametrize("M,N,K", [(64, 4096, 4096), (192, 12288, 4096), (192, 4096, 11008), (33, 1024, 512),
                                   (256, 768, 256), (24, 32064, 256), (17, 2048, 4096), (100, 32064, 1024),
                                   (1, 256, 64), (192, 22016, 4096)])
def test_stream_gemm_all_epilogues(ops, cuda, M, N, K):
    """tcgen05 weight-streaming GEMM (dd_stream_gemm) vs an fp64 restatement with the HF rounding points; the
    shapes cover whole tiles, stream-K splits (few tiles, long K), a ragged last tile (32064 = 250.5 groups),
    row padding (M % 16 != 0) and the single-unit grid."""
    g = torch.Generator().manual_seed(M + N + K)
    x = r16(torch.randn(M, K, generator=g))
    w = r16(torch.randn(N, K, generator=g) / math.sqrt(K))
    res = r16(torch.randn(M, N, generator=g))
    acc = (x.double() @ w.double().T).float()
    xg, wg, rg = x.to(cuda).bfloat16(), w.to(cuda).bfloat16(), res.to(cuda).bfloat16()
    tol = 2 ** -7
    ws = ops.stream_gemm_workspace(M, [(N, K, ops.EPI_NONE), (N // 2, K, ops.EPI_SWIGLU)] if N % 256 == 0
                                   else [(N, K, ops.EPI_NONE)], cuda)

    def close(got, want):
        assert (got.float().cpu() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())

    for _ in range(2):      # the second round runs on the counters the first one left behind
        close(ops.stream_gemm(xg, wg, workspace=ws), r16(acc))
        close(ops.stream_gemm(xg, wg, ops.EPI_RESIDUAL, residual=rg, workspace=ws), r16(r16(acc) + res))
        y32 = ops.stream_gemm(xg, wg, ops.EPI_F32, workspace=ws)
        assert y32.dtype == torch.float32
        close(y32, r16(acc))
        if N % 256 == 0:
            gate, up = r16(acc[:, :N // 2]), r16(acc[:, N // 2:])
            close(ops.stream_gemm(xg, wg, ops.EPI_SWIGLU, workspace=ws), r16(r16(torch.nn.functional.silu(gate)) * up))
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0       # counters re-armed


def test_stream_gemm_exact_and_deterministic(ops, cuda):
    """Integer-valued inputs: every product and partial sum is exact, so a descriptor / swizzle / split-reduction
    mistake is a wrong integer.  Random inputs: two runs agree bit for bit (fixed reduction order), and the result
    equals dd_skinny_gemm's up to fp32 summation order."""
    g = torch.Generator().manual_seed(3)
    for M, N, K in ((40, 384, 512), (192, 4096, 4096), (20, 640, 11008)):
        x = torch.randint(-3, 4, (M, K), generator=g).float()
        w = torch.randint(-3, 4, (N, K), generator=g).float()
        y = ops.stream_gemm(x.to(cuda).bfloat16(), w.to(cuda).bfloat16(), ops.EPI_F32)
        want = (x @ w.T)
        assert torch.equal(y.cpu(), want.to(torch.bfloat16).float()), (M, N, K)
    x = torch.randn(24, 4096, generator=g).to(cuda).bfloat16()
    w = (torch.randn(4096, 4096, generator=g) / 64).to(cuda).bfloat16()
    a = ops.stream_gemm(x, w, ops.EPI_F32)
    b = ops.stream_gemm(x, w, ops.EPI_F32)
    assert torch.equal(a, b)
    c = ops.skinny_gemm(x, w, ops.EPI_F32)
    assert (a - c).abs().max().item() <= 2 ** -7 * max(1.0, c.abs().max().item())


def test_rmsnorm_rope_embed(ops, cuda):
    g = torch.Generator().manual_seed(2)
    M, d = 6, 4096
    x, w = r16(torch.randn(M, d, generator=g) * 3), r16(torch.rand(d, generator=g) + 0.5)
    y = ops.rmsnorm(x.to(cuda).bfloat16(), w.to(cuda).bfloat16(), 1e-5)
    want = r16(r16(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)) * w)
    assert (y.float().cpu() - want).abs().max().item() <= 2 ** -7 * want.abs().max().item()

    B, Mm, H, H_kv, D = 2, 3, 8, 4, 128
    from oracle import rope_table
    cos, sin = rope_table(D, 10000.0, 64, faithful=True)
    qkv = r16(torch.randn(B * Mm, (H + 2 * H_kv) * D, generator=g))
    seq = torch.tensor([7, 33], dtype=torch.int32)
    q_out = torch.empty(B, Mm, H, D, dtype=torch.bfloat16, device=cuda)
    k_new = torch.empty(B, Mm, H_kv, D, dtype=torch.bfloat16, device=cuda)
    v_new = torch.empty_like(k_new)
    ops.rope_stage(qkv.to(cuda).bfloat16(), cos.to(cuda), sin.to(cuda), seq.to(cuda), B, Mm, H, H_kv, D, q_out,
                   k_new, v_new)
    rows = qkv.view(B, Mm, H + 2 * H_kv, D)

    def rope(xh, pos):
        x1, x2 = xh[..., :D // 2], xh[..., D // 2:]
        c, s = cos[pos], sin[pos]
        return r16(torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1))

    for b in range(B):
        assert (q_out[b].float().cpu() - rope(rows[b, :, :H], int(seq[b]))).abs().max() <= 2 ** -6
        assert (k_new[b].float().cpu() - rope(rows[b, :, H:H + H_kv], int(seq[b]))).abs().max() <= 2 ** -6
        assert torch.equal(v_new[b].float().cpu(), rows[b, :, H + H_kv:])

    emb = r16(torch.randn(100, 256, generator=g)).to(cuda).bfloat16()
    tok = torch.tensor([3, 99], dtype=torch.int32, device=cuda)
    out = torch.empty(2 * 3, 256, dtype=torch.bfloat16, device=cuda)
    ops.embed_gather(emb, tok, 3, out)
    assert torch.equal(out.view(2, 3, 256), emb[tok.long()][:, None].expand(2, 3, 256))


@pytest.mark.parametrize("B,Mm,H,H_kv,D,K", [(1, 1, 4, 4, 64, 256), (1, 3, 32, 32, 128, 4096), (2, 3, 8, 2, 128, 512),
                                             (4, 8, 4, 4, 64, 256)])
def test_decoder_linear_fused_norm_and_rope(ops, cuda, B, Mm, H, H_kv, D, K):
    """dd_decoder_linear = rmsnorm -> skinny GEMM -> rope/staging in one launch: equal to the chain of
    separate kernels (one bf16 ulp where the 1/rms summation order moves a rounding) and to the fp32
    restatement with the HF rounding points."""
    from oracle import rope_table
    g = torch.Generator().manual_seed(B * 100 + Mm * 10 + H)
    rows, N = B * Mm, (H + 2 * H_kv) * D
    x = r16(torch.randn(rows, K, generator=g) * 2).to(cuda).bfloat16()
    nw = r16(torch.rand(K, generator=g) + 0.5).to(cuda).bfloat16()
    w = r16(torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    cos, sin = rope_table(D, 10000.0, 64, faithful=True)
    cos, sin = cos.to(cuda), sin.to(cuda)
    seq = torch.randint(3, 60, (B,), generator=g).int().to(cuda)

    def bufs():
        return (torch.zeros(B, Mm, H, D, dtype=torch.bfloat16, device=cuda),
                torch.zeros(B, Mm, H_kv, D, dtype=torch.bfloat16, device=cuda),
                torch.zeros(B, Mm, H_kv, D, dtype=torch.bfloat16, device=cuda))

    q1, k1, v1 = bufs()
    xn = ops.rmsnorm(x, nw, 1e-5)
    qkv = ops.skinny_gemm(xn, w)
    ops.rope_stage(qkv, cos, sin, seq, B, Mm, H, H_kv, D, q1, k1, v1)
    q2, k2, v2 = bufs()
    ops.decoder_linear(x, w, ops.EPI_ROPE, norm_w=nw, eps=1e-5,
                       rope=dict(cos=cos, sin=sin, seq_lens=seq, members=Mm, H=H, H_kv=H_kv, D=D, q_out=q2, k_new=k2,
                                 v_new=v2))
    for a, b in ((q1, q2), (k1, k2), (v1, v2)):
        d = (a.float() - b.float()).abs()
        assert d.max().item() <= 2 ** -6 * max(1.0, a.float().abs().max().item())
        assert (d > 0).float().mean().item() < 0.02          # nearly always bit-identical
    # fused norm + plain / SwiGLU / residual / fp32 epilogues against the unfused chain
    w2 = r16(torch.randn(512, K, generator=g) / math.sqrt(K)).to(cuda).bfloat16()
    res = r16(torch.randn(rows, 512, generator=g)).to(cuda).bfloat16()
    for epi, kw in ((ops.EPI_NONE, {}), (ops.EPI_RESIDUAL, dict(residual=res)), (ops.EPI_F32, {}), (ops.EPI_SWIGLU, {})):
        want = ops.skinny_gemm(xn, w2, epi, **kw)
        got = ops.decoder_linear(x, w2, epi, norm_w=nw, eps=1e-5, **kw)
        d = (want.float() - got.float()).abs()
        assert d.max().item() <= 2 ** -6 * max(1.0, want.float().abs().max().item())
        assert torch.equal(ops.decoder_linear(xn, w2, epi, **kw), want)      # no norm: the very same kernel


# ---------------------------------------------------------------------------------------------
# (1) uncertainty stage (tcgen05) vs calculate_vision_uncertainty / get_topk_token_id (golden)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", UNCERT_CASES, ids=[c[0] for c in UNCERT_CASES])
def test_uncertainty_matches_reference(ops, cuda, golden_dir, case):
    name, n_vis, d, V, hs, ws, k = case
    z = np.load(os.path.join(golden_dir, "uncertainty.npz"))
    h, w = uncert_inputs(name, n_vis, d, V, hs, ws)
    epi, alea, var, ids = ops.uncertainty_project(h.to(cuda).bfloat16()[None].contiguous(),
                                                  w.to(cuda).bfloat16().contiguous(), topk=k)
    torch.cuda.synchronize()
    # epistemic: the north-star bar, 1e-3 relative.  aleatoric / variance: same relative bar plus an
    # absolute floor of a few fp32 ulps of the O(1) terms they are sums of -- for a peaked row
    # (p_max ~ 1 - 1e-5) alea is ~1e-5 and the reference's own fp32 softmax is only that accurate.
    for got, key, rtol, atol in ((epi, "epi", 1e-3, 0.0), (alea, "alea", 1e-3, 2e-6), (var, "var", 1e-3, 1e-9)):
        want = z[f"{name}.{key}"]
        err = np.abs(got[0].cpu().numpy() - want) - atol
        rel = err / np.maximum(np.abs(want), 1e-30)
        assert rel.max() <= rtol, f"{name}.{key}: max rel err {rel.max():.3e}"
    # top-k id SETS are exact wherever the k-th / (k+1)-th logits are not a numerical tie
    safe = z[f"{name}.topk_gap"] > 1e-4 * np.maximum(1.0, np.abs(z[f"{name}.row_max"]))
    got_ids, want_ids = np.sort(ids[0].cpu().numpy(), -1), np.sort(z[f"{name}.topk_ids"], -1)
    assert safe.mean() > 0.95
    assert np.array_equal(got_ids[safe], want_ids[safe]), f"{name}: top-{k} id sets differ"


def test_uncertainty_batched_images_and_bf16_rounding(ops, cuda):
    """Several images in one launch give the per-image results; round_logits reproduces the
    half-precision lm_head + .float() of the reference (models/llama.py:46-47)."""
    from oracle import calculate_vision_uncertainty
    g = torch.Generator().manual_seed(11)
    n_img, n_vis, d, V = 3, 150, 128, 5000
    h = r16(torch.randn(n_img, n_vis, d, generator=g))
    w = r16(torch.randn(V, d, generator=g) * 0.2)
    for rnd in (False, True):
        epi, alea, var, ids = ops.uncertainty_project(h.to(cuda).bfloat16(), w.to(cuda).bfloat16(), topk=5,
                                                      round_logits=rnd)
        for i in range(n_img):
            logits = h[i] @ w.T
            if rnd:
                logits = r16(logits)
            ref = calculate_vision_uncertainty(logits[None])
            for got, key in ((epi, "epis_uncert_per_token"), (alea, "alea_uncert_per_token")):
                want = ref[key][0]
                assert ((got[i].cpu() - want).abs() / want.abs()).max() <= 1e-3


@pytest.mark.parametrize("n_img,n_vis,d,V,k", [(1, 1500, 128, 3000, 10), (2, 700, 64, 1111, 5)],
                         ids=["next_1500_rows_3_chunks", "two_images_700_rows"])
def test_uncertainty_more_rows_than_tmem_holds(ops, cuda, n_img, n_vis, d, V, k):
    """LLaVA-NeXT anyres yields up to ~2.9k visual tokens: several 256-row GEMM tiles and 128-row entropy tiles per
    image, a last row group that fills only one accumulator, a last vocabulary strip that is partly padding."""
    from oracle import calculate_vision_uncertainty, topk_token_ids
    g = torch.Generator().manual_seed(n_vis)
    h = r16(torch.randn(n_img, n_vis, d, generator=g))
    w = r16(torch.randn(V, d, generator=g) * 0.25)
    epi, alea, var, ids = ops.uncertainty_project(h.to(cuda).bfloat16(), w.to(cuda).bfloat16(), topk=k)
    for i in range(n_img):
        logits = h[i] @ w.T
        ref = calculate_vision_uncertainty(logits[None])
        for got, key, atol in ((epi, "epis_uncert_per_token", 0.0), (alea, "alea_uncert_per_token", 2e-6),
                               (var, "variance_per_token", 1e-9)):
            want = ref[key][0]
            rel = ((got[i].cpu() - want).abs() - atol) / want.abs()
            assert rel.max() <= 1e-3, f"{key}: {rel.max():.2e}"
        srt = torch.sort(logits, dim=-1, descending=True).values
        safe = (srt[:, k - 1] - srt[:, k]) > 1e-4 * srt[:, 0].abs().clamp(min=1)
        want_ids = torch.sort(topk_token_ids(logits[None], k)[1][0], -1).values
        got_ids = torch.sort(ids[i].cpu().long(), -1).values
        assert torch.equal(got_ids[safe], want_ids[safe])


@pytest.mark.parametrize("n_img,n_vis,d,V,k,rnd", [(1, 1, 64, 100, 5, False), (1, 129, 64, 257, 3, True),
                                                   (5, 1200, 64, 700, 5, True), (3, 2000, 64, 320, 4, False),
                                                   (2, 128, 128, 32001, 10, True)],
                         ids=["one_row", "129_rows_257_cols", "two_image_groups_bf16", "three_image_groups_fp32",
                              "instructblip_vocab"])
def test_uncertainty_groups_edges_and_determinism(ops, cuda, n_img, n_vis, d, V, k, rnd):
    """Image groups (the stash holds <= 4608 rows: 5 x 1200 rows = 3 + 2 images, 3 x 2000 rows = 2 + 1), a single
    row, tiles with one valid row, a vocabulary that ends inside a 64-column strip; results written into caller
    tensors (out=) equal the returned ones and a second run is bit-identical (every partial is summed in a fixed
    order)."""
    from oracle import calculate_vision_uncertainty, topk_token_ids
    g = torch.Generator().manual_seed(1000 * n_img + n_vis)
    h = r16(torch.randn(n_img, n_vis, d, generator=g))
    w = r16(torch.randn(V, d, generator=g) * 0.3)
    hd, wd = h.to(cuda).bfloat16(), w.to(cuda).bfloat16()
    epi, alea, var, ids = ops.uncertainty_project(hd, wd, topk=k, round_logits=rnd)
    out = (torch.empty_like(epi), torch.empty_like(alea), torch.empty_like(var), torch.empty_like(ids))
    ws = ops.uncertainty_workspace(n_img, n_vis, V, k, cuda)
    ws.fill_(0x7f)                                             # stale scratch must not matter
    ops.uncertainty_project(hd, wd, topk=k, round_logits=rnd, out=out, workspace=ws)
    for a, b in zip((epi, alea, var, ids), out):
        assert torch.equal(a, b)
    for i in range(n_img):
        logits = h[i] @ w.T
        if rnd:
            logits = r16(logits)
        ref = calculate_vision_uncertainty(logits[None])
        # (a single row is its own average: epi = A - C is a difference of two O(1) fp32 sums around zero)
        for got, key, atol in ((epi, "epis_uncert_per_token", 5e-6), (alea, "alea_uncert_per_token", 2e-6),
                               (var, "variance_per_token", 1e-9)):
            want = ref[key][0]
            rel = ((got[i].cpu() - want).abs() - atol) / want.abs().clamp_min(1e-30)
            assert rel.max() <= 1e-3, f"image {i} {key}: {rel.max():.2e}"
        srt = torch.sort(logits, dim=-1, descending=True).values
        safe = (srt[:, k - 1] - srt[:, k]) > 1e-4 * srt[:, 0].abs().clamp(min=1)
        want_ids = torch.sort(topk_token_ids(logits[None], k)[1][0], -1).values
        got_ids = torch.sort(ids[i].cpu().long(), -1).values
        assert torch.equal(got_ids[safe], want_ids[safe])


def test_uncertainty_candidate_overflow_is_an_error(ops, cuda):
    """All-equal logits tie every vocabulary entry at the k-th largest value: the per-row candidate list cannot hold
    them, and that is reported (DD_ERR_OVERFLOW through the workspace counter), never a silently truncated list."""
    h = torch.zeros(1, 64, 128, dtype=torch.bfloat16, device=cuda)
    w = torch.randn(4096, 128, generator=torch.Generator().manual_seed(0)).to(cuda).bfloat16()
    with pytest.raises(ops.DropDecError, match="overflow"):
        ops.uncertainty_project(h, w, topk=5, round_logits=True)
    *_, ovf = ops.uncertainty_project(h, w, topk=5, round_logits=True, return_overflow=True)
    assert int(ovf.item()) == 64
    hr = torch.randn(1, 64, 128, generator=torch.Generator().manual_seed(1)).to(cuda).bfloat16()
    *_, ovf = ops.uncertainty_project(hr, w, topk=5, round_logits=True, return_overflow=True)
    assert int(ovf.item()) == 0


def test_aggregate_of_nan_and_minus_inf_rows_yields_valid_ids(ops, cuda):
    """torch.argmax treats NaN as the maximum and returns index 0 for a row of -inf; the kernel used to return its
    0x7fffffff start value for such rows, which the next step's embedding gather would have read out of bounds."""
    V = 1000
    l = torch.randn(2, 3, V, generator=torch.Generator().manual_seed(3))
    l[0, 0] = float("nan")
    l[0, 1] = float("-inf")
    l[1, 2, 17] = float("nan")
    ids, winner, token, _ = ops.aggregate(l.to(cuda), "vote")
    want = torch.argmax(l, dim=-1)
    assert ids.cpu().tolist() == want.tolist()
    assert bool((token >= 0).all()) and bool((token < V).all())
    # an out-of-range id handed to the embedding gather is clamped, not dereferenced
    emb = torch.arange(8 * 16, dtype=torch.float32).reshape(8, 16).to(cuda).bfloat16()
    out = torch.empty(2, 16, dtype=torch.bfloat16, device=cuda)
    ops.embed_gather(emb, torch.tensor([0x7fffffff, -5], dtype=torch.int32, device=cuda), 1, out)
    assert torch.equal(out[0], emb[7]) and torch.equal(out[1], emb[0])


def test_l2_prefetch_hint_leaves_results_untouched(ops, cuda):
    """dd_hint_l2_prefetch: the hinted rmsnorm / decode_attention launch also issues bulk L2 prefetches over the given
    ranges (odd sizes, a range shorter than one chunk, a second range) -- outputs are bit-identical to the unhinted call
    and the hint is one-shot."""
    g = torch.Generator().manual_seed(11)
    w_big = torch.randn(3 * 1024 * 1024 + 40, generator=g).to(cuda).bfloat16()      # ~6 MB
    w_small = torch.randn(1000, generator=g).to(cuda).bfloat16()
    M, d = 5, 4096
    x, w = torch.randn(M, d, generator=g).to(cuda).bfloat16(), torch.rand(d, generator=g).to(cuda).bfloat16()
    base = ops.rmsnorm(x, w, 1e-5)
    ops.hint_l2_prefetch(w_big, 48, 5 * 1024 * 1024 + 16, w_small, 0, 1 << 30)
    assert torch.equal(ops.rmsnorm(x, w, 1e-5), base)
    assert torch.equal(ops.rmsnorm(x, w, 1e-5), base)            # the hint is gone: plain launch again

    B, Mm, H, D, S_max = 2, 4, 8, 128, 300
    bf = lambda *s: torch.randn(*s, generator=g).to(cuda).bfloat16()
    q, k_new, v_new, kc, vc = bf(B, Mm, H, D), bf(B, Mm, H, D), bf(B, Mm, H, D), bf(B, H, S_max, D), bf(B, H, S_max, D)
    lens = torch.tensor([257, 64], dtype=torch.int32, device=cuda)
    for splits in (1, 3):
        want = ops.decode_attention(q, k_new, v_new, kc, vc, lens, None, None, 0, 0.088, num_splits=splits)
        ops.hint_l2_prefetch(w_big, 0, 1 << 40, w_big, 1 << 20, 4096)
        got = ops.decode_attention(q, k_new, v_new, kc, vc, lens, None, None, 0, 0.088, num_splits=splits)
        assert torch.equal(got, want)
    torch.cuda.synchronize()
    from dropoutdecoding_b200 import _lib
    with pytest.raises(ops.DropDecError):                        # misaligned range
        ops.check_rc(_lib.load().dd_hint_l2_prefetch(w_big.data_ptr() + 2, 32, None, 0), "dd_hint_l2_prefetch")


def test_two_window_attention_equals_two_calls(ops, cuda):
    """dd_decode_attention_two_windows (the lookahead pass: members [0, Ma) over keys [0, S), members [Ma, M) over keys
    [0, S + 1), split key ranges, ONE launch) is bit-identical to two dd_decode_attention_active calls, masks included."""
    g = torch.Generator().manual_seed(21)
    B, Ma, Mb, H, D, S_max, n_vis, splits = 2, 3, 4, 8, 128, 700, 576, 3
    M = Ma + Mb
    bf = lambda *s: torch.randn(*s, generator=g).to(cuda).bfloat16()
    q, k_new, v_new, kc, vc = bf(B, M, H, D), bf(B, M, H, D), bf(B, M, H, D), bf(B, H, S_max, D), bf(B, H, S_max, D)
    lens_a = torch.tensor([601, 640], dtype=torch.int32, device=cuda)
    lens_b = lens_a + 1
    vs = torch.tensor([5, 9], dtype=torch.int32, device=cuda)
    bm = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, M, ops.mask_words(n_vis)), generator=g, dtype=torch.int64).to(torch.int32).to(cuda)
    scale = 1.0 / math.sqrt(D)
    want = torch.zeros_like(q)
    ws = ops.attention_workspace(B, M, H, D, splits, cuda)
    ops.decode_attention(q, k_new, v_new, kc, vc, lens_a, vs, bm, n_vis, scale, splits, want, ws, member_offset=0, members=Ma)
    ops.decode_attention(q, k_new, v_new, kc, vc, lens_b, vs, bm, n_vis, scale, splits, want, ws, member_offset=Ma, members=Mb)
    got = torch.zeros_like(q)
    ws2 = ops.two_window_attention_workspace(B, M, H, D, splits, cuda)
    for _ in range(2):                                            # the tickets are left at zero: callable again
        got.zero_()
        ops.decode_attention_two_windows(q, k_new, v_new, kc, vc, lens_a, lens_b, Ma, vs, bm, n_vis, scale, splits, got, ws2)
        assert torch.equal(got, want)
    with pytest.raises(ops.DropDecError):                         # unsplit ranges take the len_extra path instead
        ops.decode_attention_two_windows(q, k_new, v_new, kc, vc, lens_a, lens_b, Ma, vs, bm, n_vis, scale, 1, got, ws2)


def test_rope_epilogue_commits_the_guessed_row(ops, cuda):
    """dd_rope_args.guess_winner: the QKV projection's epilogue also writes the new K/V rows of member guess_winner[b] to
    cache row guess_pos[b] -- exactly what dd_commit_winner of that layer writes; a winner outside the members or a position
    past the cache writes nothing; everything else of the projection is unchanged."""
    from oracle import rope_table
    g = torch.Generator().manual_seed(4)
    B, Mm, H, H_kv, D, K, S_max = 2, 7, 8, 8, 128, 1024, 40
    rows, N = B * Mm, (H + 2 * H_kv) * D
    cos, sin = rope_table(D, 10000.0, 64, faithful=True)
    x = (torch.randn(rows, K, generator=g) * 0.5).to(cuda).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda).bfloat16()
    pos = torch.randint(0, 30, (rows,), generator=g, dtype=torch.int32).to(cuda)
    mk = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=cuda)
    outs = []
    for guess in (None, torch.tensor([2, 9], dtype=torch.int32, device=cuda)):      # 9: outside the 7 members
        q, kn, vn, kc, vc = mk(B, Mm, H, D), mk(B, Mm, H_kv, D), mk(B, Mm, H_kv, D), mk(B, H_kv, S_max, D), mk(B, H_kv, S_max, D)
        rope = dict(cos=cos.to(cuda), sin=sin.to(cuda), seq_lens=pos, members=1, H=H, H_kv=H_kv, D=D, q_out=q, k_new=kn, v_new=vn)
        if guess is not None:
            rope.update(guess_winner=guess, guess_pos=torch.tensor([17, 3], dtype=torch.int32, device=cuda),
                        guess_members=Mm, k_cache=kc, v_cache=vc)
        ops.decoder_linear(x, w, ops.EPI_ROPE, rope=rope)
        outs.append((q, kn, vn, kc, vc))
    (q0, kn0, vn0, kc0, vc0), (q1, kn1, vn1, kc1, vc1) = outs
    assert torch.equal(q0, q1) and torch.equal(kn0, kn1) and torch.equal(vn0, vn1)
    assert int(kc0.abs().sum()) == 0 and int(vc0.abs().sum()) == 0
    want_k, want_v = torch.zeros_like(kc1), torch.zeros_like(vc1)
    want_k[0, :, 17], want_v[0, :, 17] = kn1[0, 2], vn1[0, 2]                       # sequence 1's winner is out of range
    assert torch.equal(kc1, want_k) and torch.equal(vc1, want_v)
    assert float(kn1[0, 2].abs().sum()) > 0


def test_schedule_flags(ops, cuda):
    a = torch.randint(0, 2 ** 31 - 1, (1, 3, 18), dtype=torch.int64).to(torch.int32).to(cuda)
    b = a.clone()
    tok, win = torch.tensor([77], dtype=torch.int32, device=cuda), torch.tensor([1], dtype=torch.int32, device=cuda)
    f = torch.full((2,), -1, dtype=torch.int32, device=cuda)
    assert ops.schedule_flags(f, mask_a=a, mask_b=b).tolist() == [1, 0]
    b[0, 2, 17] ^= 1
    assert ops.schedule_flags(f, mask_a=a, mask_b=b, tok_a=tok, tok_b=tok.clone(), win_a=win, win_b=win.clone()).tolist() == [1, 1]
    assert ops.schedule_flags(f, mask_a=a, mask_b=a, tok_a=tok, tok_b=tok + 1, win_a=win, win_b=win).tolist() == [0, 0]
    assert ops.schedule_flags(f, tok_a=tok, tok_b=tok, win_a=win, win_b=win + 1).tolist() == [0, 0]
