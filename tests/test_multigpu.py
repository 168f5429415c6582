"""N > 1 on real GPUs: torchrun launches tools/multigpu_check.py (NCCL) when the box has at least two GPUs.
The host-side logic of the same module is covered on CPU by tests/test_distributed_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_member_and_sequence_sharding_on_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"ok": true' in r.stdout


def test_sequence_sharding_is_independent_of_the_split_single_gpu(cuda):
    """seq_base keys the Philox stream: sequences 2..3 decoded alone equal rows 2..3 of the batch of 4."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True,
                       text=True, timeout=900, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"sequence_sharding_tokens_equal": true' in r.stdout


@pytest.mark.parametrize("aggregate", ["vote", "mean"])
@pytest.mark.parametrize("use_graph", [True, False], ids=["cuda_graph", "eager"])
def test_member_exchange_record_and_finish_kernel_on_one_gpu(cuda, aggregate, use_graph):
    """The member-sharded engine with a world of one rank runs the whole mechanism -- kernels writing ids / statistics /
    staged K/V rows into the exchange record, the (degenerate) gather, dd_member_finish voting and committing from the
    gathered buffer -- and must reproduce the plain two-pass engine bit for bit: tokens, caches, per-token statistics."""
    from dropoutdecoding_b200 import distributed as dd_dist
    from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine
    cfg = DecoderConfig.tiny()
    w = DecoderWeights.random(cfg, cuda, seed=0, head_std=0.08)
    probs, n_vis, S0, T, B = [0.1, 0.3, 0.5, 0.7], 576, 593, 10, 2
    g = torch.Generator(device=cuda).manual_seed(5)
    kc = torch.randn(cfg.n_layers, B, cfg.n_kv_heads, S0, cfg.head_dim, generator=g, device=cuda)
    vc = torch.randn(cfg.n_layers, B, cfg.n_kv_heads, S0, cfg.head_dim, generator=g, device=cuda)
    epi = torch.rand(B, n_vis, generator=g, device=cuda) * 0.2 + 0.75
    ids = torch.randint(0, cfg.vocab, (B, n_vis, 5), generator=g, device=cuda, dtype=torch.int32)

    def prime(eng):
        eng.load_prefill(kc, vc, [5] * B)
        eng.epi.copy_(epi); eng.topk_ids.copy_(ids)
        eng.cur_token.fill_(1234); eng.tokens_out[:, 0] = 1234

    ref = DropoutDecodeEngine(w, B, n_vis, S0 + T + 4, probs=probs, max_steps=32, seed=24, aggregate=aggregate,
                              speculative="off")
    prime(ref)
    ref.decode(T, use_graph=False)
    sh = dd_dist.make_member_sharded_engine(w, B, n_vis, S0 + T + 4, probs, max_steps=32, seed=24, aggregate=aggregate)
    prime(sh)
    assert sh.decode(T, use_graph=use_graph) == T
    torch.cuda.synchronize()
    assert torch.equal(ref.generated(), sh.generated())
    assert torch.equal(ref.k_cache, sh.k_cache) and torch.equal(ref.v_cache, sh.v_cache)
    assert torch.equal(ref.member_ids, sh.member_ids) and torch.equal(ref.winner, sh.winner)
    assert torch.equal(ref.token_stats[:T + 1], sh.token_stats[:T + 1])
    a, b, c = sh.token_bookkeeping(0)
    assert len(a) == T + 1 and any(x != 0 for x in b[1:])
