#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200 (CUDA events, L2 flushed between iterations).

    python tools/microbench.py [attention|gemm|uncertainty|all]

Prints one JSON line per measurement with the achieved fraction of the measured peak
(MEASURED_PEAKS.json, else the fallback of the profiling guide).
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dropoutdecoding_b200 import ops  # noqa: E402


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


HBM, TF, SRC = peaks()
dev = torch.device("cuda:0")
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _flush.zero_()


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_attention():
    for (B, M, H, H_kv, D, S, n_vis, splits) in [
            (1, 3, 32, 32, 128, 849, 576, 9), (1, 1, 32, 32, 128, 849, 0, 9),
            (8, 3, 32, 32, 128, 849, 576, 2), (64, 3, 32, 32, 128, 849, 576, 1),
            (64, 1, 32, 32, 128, 849, 0, 1), (64, 3, 32, 32, 128, 849, 576, 2),
            (16, 8, 32, 8, 128, 3000, 2880, 4)]:
        S_max = S + 8
        q = torch.randn(B, M, H, D, device=dev).bfloat16()
        kn = torch.randn(B, M, H_kv, D, device=dev).bfloat16()
        vn = torch.randn_like(kn)
        kc = torch.randn(B, H_kv, S_max, D, device=dev).bfloat16()
        vc = torch.randn_like(kc)
        sl = torch.full((B,), S, dtype=torch.int32, device=dev)
        vs = torch.full((B,), 5, dtype=torch.int32, device=dev)
        bm = None
        if n_vis:
            bm = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, M, ops.mask_words(n_vis)), dtype=torch.int32, device=dev)
        ws = ops.attention_workspace(B, M, H, D, splits, dev)
        out = torch.empty_like(q)
        fn = lambda: ops.decode_attention(q, kn, vn, kc, vc, sl, vs, bm, n_vis, 1 / math.sqrt(D), splits, out, ws)
        med, best = timeit(fn)
        algo = 2 * S * H_kv * D * 2 * B + 2 * M * H * D * 2 * B + (M * ((n_vis + 7) // 8) * B)
        print(json.dumps(dict(kernel="decode_attention(+merge)", B=B, M=M, H=H, H_kv=H_kv, D=D, S=S, splits=splits,
                              us=round(med * 1e6, 2), best_us=round(best * 1e6, 2), GBps=round(algo / med / 1e9, 1),
                              frac_hbm=round(algo / med / 1e9 / HBM, 3), peak=SRC)))


def bench_attention_sweep():
    """num_splits sweep at small / medium batch (CUDA-graph replay of 8 calls over rotating caches)."""
    for (B, M, H, H_kv, D, S, n_vis) in [(1, 3, 32, 32, 128, 849, 576), (1, 1, 32, 32, 128, 849, 0),
                                         (8, 3, 32, 32, 128, 849, 576), (8, 1, 32, 32, 128, 849, 0),
                                         (16, 3, 32, 32, 128, 849, 576), (2, 8, 32, 8, 128, 3000, 2880),
                                         (16, 8, 32, 8, 128, 3000, 2880)]:
        S_max = S + 8
        q = torch.randn(B, M, H, D, device=dev).bfloat16()
        kn = torch.randn(B, M, H_kv, D, device=dev).bfloat16()
        vn = torch.randn_like(kn)
        R = max(2, -(-(300 << 20) // (2 * B * H_kv * S_max * D * 2)))
        R = min(R, 64)
        kcs = [torch.randn(B, H_kv, S_max, D, device=dev).bfloat16() for _ in range(R)]
        vcs = [torch.randn(B, H_kv, S_max, D, device=dev).bfloat16() for _ in range(R)]
        sl = torch.full((B,), S, dtype=torch.int32, device=dev)
        vs = torch.full((B,), 5, dtype=torch.int32, device=dev)
        bm = None
        if n_vis:
            bm = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, M, ops.mask_words(n_vis)), dtype=torch.int32, device=dev)
        out = torch.empty_like(q)
        algo = 2 * S * H_kv * D * 2 * B + 2 * M * H * D * 2 * B + (M * ((n_vis + 7) // 8) * B)
        res = {}
        for splits in (1, 2, 3, 4, 6, 8, 12, 16, 24):
            if splits > max(1, S // 64):
                continue
            ws = ops.attention_workspace(B, M, H, D, splits, dev)

            def run():
                for kc, vc in zip(kcs, vcs):
                    ops.decode_attention(q, kn, vn, kc, vc, sl, vs, bm, n_vis, 1 / math.sqrt(D), splits, out, ws)
            run(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            g.replay(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                g.replay()
            b.record(); torch.cuda.synchronize()
            res[splits] = round(a.elapsed_time(b) * 1e3 / (3 * R), 2)
        best = min(res, key=res.get)
        print(json.dumps(dict(kernel="decode_attention sweep", B=B, M=M, H_kv=H_kv, S=S, us_by_splits=res, best_splits=best,
                              best_frac_hbm=round(algo / (res[best] * 1e-6) / 1e9 / HBM, 3))), flush=True)


def bench_gemm():
    for (M, N, K, epi) in [(1, 4096, 4096, ops.EPI_NONE), (3, 4096, 4096, ops.EPI_NONE), (1, 12288, 4096, ops.EPI_NONE),
                           (3, 12288, 4096, ops.EPI_NONE), (3, 22016, 4096, ops.EPI_SWIGLU),
                           (3, 4096, 11008, ops.EPI_RESIDUAL), (3, 32064, 4096, ops.EPI_F32),
                           (8, 12288, 4096, ops.EPI_NONE), (24, 12288, 4096, ops.EPI_NONE),
                           (32, 4096, 11008, ops.EPI_RESIDUAL)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        Nout = N // 2 if epi == ops.EPI_SWIGLU else N
        res = torch.randn(M, Nout, device=dev).bfloat16()
        out = torch.empty(M, Nout, dtype=torch.float32 if epi == ops.EPI_F32 else torch.bfloat16, device=dev)
        fn = lambda: ops.skinny_gemm(x, w, epi, residual=res, out=out)
        med, best = timeit(fn)
        algo = N * K * 2
        tfn = lambda: torch.matmul(x, w.T)
        tmed, _ = timeit(tfn)
        print(json.dumps(dict(kernel="skinny_gemm", M=M, N=N, K=K, epi=epi, us=round(med * 1e6, 2),
                              best_us=round(best * 1e6, 2), GBps=round(algo / med / 1e9, 1),
                              frac_hbm=round(algo / med / 1e9 / HBM, 3), cublas_us=round(tmed * 1e6, 2), peak=SRC)))


def bench_stream_gemm():
    """dd_stream_gemm (tcgen05, stream-K) at the batched-pass shapes of LLaVA-1.5-7B: rows = 24 (B=8, K=3),
    64 / 192 (B=64: preliminary / masked pass); library GEMM and (M <= 32) dd_skinny_gemm beside it."""
    shapes = [(12288, 4096, ops.EPI_NONE), (4096, 4096, ops.EPI_RESIDUAL), (22016, 4096, ops.EPI_SWIGLU),
              (4096, 11008, ops.EPI_RESIDUAL), (32064, 4096, ops.EPI_F32)]
    quick = os.environ.get("DD_SG_QUICK")
    for M in ((192,) if quick else (24, 64, 192)):
        for (N, K, epi) in (shapes[:2] if quick else shapes):
            x = torch.randn(M, K, device=dev).bfloat16()
            R = max(2, -(-(400 << 20) // (N * K * 2)))      # rotate over > 3 x L2 worth of weight copies
            wr = [(torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16() for _ in range(R)]
            w = wr[0]
            Nout = N // 2 if epi == ops.EPI_SWIGLU else N
            res = torch.randn(M, Nout, device=dev).bfloat16()
            out = torch.empty(M, Nout, dtype=torch.float32 if epi == ops.EPI_F32 else torch.bfloat16, device=dev)
            ws = ops.stream_gemm_workspace(M, [(Nout, K, epi)], dev)

            def burst(f, reps=3):
                """R launches over rotating weight copies captured in one CUDA graph, replayed: us per launch
                (no host gaps, no L2 reuse of the weights)"""
                for wi in wr:
                    f(wi)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for wi in wr:
                        f(wi)
                g.replay()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    g.replay()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) * 1e-3 / (R * reps)

            med = best = burst(lambda wi: ops.stream_gemm(x, wi, epi, residual=res, out=out, workspace=ws))
            tout = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            tmed = burst(lambda wi: torch.matmul(x, wi.T, out=tout))
            rec = dict(kernel="stream_gemm", dbg=os.environ.get("DD_SG_DBG", "0"), M=M, N=N, K=K, epi=epi, us=round(med * 1e6, 2),
                       GBps=round(N * K * 2 / med / 1e9, 1), frac_hbm=round(N * K * 2 / med / 1e9 / HBM, 3),
                       cublas_us=round(tmed * 1e6, 2), peak=SRC, timing="CUDA-graph replay of launches over rotating weight copies")
            if M <= 32:
                smed = burst(lambda wi: ops.skinny_gemm(x, wi, epi, residual=res, out=out))
                rec["skinny_us"] = round(smed * 1e6, 2)
            print(json.dumps(rec), flush=True)


def bench_uncertainty(only_first=False):
    cases = [(1, 576, 4096, 32064, 5), (4, 576, 4096, 32064, 5), (8, 576, 4096, 32064, 5), (64, 576, 4096, 32064, 5),
             (1, 2880, 4096, 32064, 10), (32, 32, 4096, 32001, 10)]
    for (n_img, n_vis, d, V, k) in cases[:1] if only_first else cases:
        h = torch.randn(n_img, n_vis, d, device=dev).bfloat16()
        w = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
        # the engine's call: bf16-rounded logits (the reference's lm_head output), overflow flag read later (no host sync)
        one = lambda: ops.uncertainty_project(h, w, topk=k, round_logits=True, return_overflow=True)
        med1, _ = timeit(one, iters=10)                      # one call: includes the host latency before the first launch

        def fn():                                            # 4 calls back to back: the GPU never waits for the host
            for _ in range(4):
                one()
        med, best = timeit(fn, iters=6)
        med, best = med / 4, best / 4
        flops = 2.0 * n_img * n_vis * d * V
        tfn = lambda: torch.matmul(h.view(-1, d), w.T)
        tmed, _ = timeit(tfn, iters=10)
        print(json.dumps(dict(kernel="uncertainty_project", lib=os.path.basename(os.environ.get("DD_LIB", "default")), n_img=n_img, n_vis=n_vis, d=d, V=V,
                              us=round(med * 1e6, 1), best_us=round(best * 1e6, 1), single_call_us=round(med1 * 1e6, 1),
                              TFLOPs=round(flops / med / 1e12, 1),
                              frac_tensor=round(flops / med / 1e12 / TF, 3), cublas_gemm_only_us=round(tmed * 1e6, 1),
                              peak=SRC)))


def one_shot(which):
    """Single launches for ncu (no timing loop)."""
    if which == "uncertainty1":
        h = torch.randn(1, 576, 4096, device=dev).bfloat16()
        w = (torch.randn(32064, 4096, device=dev) * 0.02).bfloat16()
        for _ in range(2):
            ops.uncertainty_project(h, w, topk=5, round_logits=True, return_overflow=True)
    elif which == "attention64":
        B, M, H, D, S = 64, 3, 32, 128, 849
        q = torch.randn(B, M, H, D, device=dev).bfloat16()
        kn = torch.randn(B, M, H, D, device=dev).bfloat16(); vn = torch.randn_like(kn)
        kc = torch.randn(B, H, S + 8, D, device=dev).bfloat16(); vc = torch.randn_like(kc)
        sl = torch.full((B,), S, dtype=torch.int32, device=dev); vs = torch.full((B,), 5, dtype=torch.int32, device=dev)
        bm = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, M, 18), dtype=torch.int32, device=dev)
        for _ in range(3):
            ops.decode_attention(q, kn, vn, kc, vc, sl, vs, bm, 576, 1 / math.sqrt(D), 1)
    elif which == "sgemm1":
        for (M, N, K, epi) in [(192, 12288, 4096, ops.EPI_NONE), (192, 4096, 4096, ops.EPI_NONE),
                               (192, 22016, 4096, ops.EPI_SWIGLU), (192, 4096, 11008, ops.EPI_NONE),
                               (64, 12288, 4096, ops.EPI_NONE), (24, 12288, 4096, ops.EPI_NONE)]:
            x = torch.randn(M, K, device=dev).bfloat16()
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
            for _ in range(2):
                flush_l2()
                ops.stream_gemm(x, w, epi)
    elif which == "gemm1":
        for (M, N, K, epi) in [(3, 4096, 4096, ops.EPI_NONE), (3, 12288, 4096, ops.EPI_NONE),
                               (3, 22016, 4096, ops.EPI_SWIGLU), (3, 4096, 11008, ops.EPI_NONE)]:
            x = torch.randn(M, K, device=dev).bfloat16()
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
            for _ in range(2):
                flush_l2()
                ops.skinny_gemm(x, w, epi)
    torch.cuda.synchronize()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("uncertainty1", "attention64", "gemm1", "sgemm1"):
        with torch.no_grad():
            one_shot(which)
        sys.exit(0)
    with torch.no_grad():
        if which in ("attention", "all"):
            bench_attention()
        if which == "attention_sweep":
            bench_attention_sweep()
        if which in ("gemm", "all"):
            bench_gemm()
        if which in ("stream_gemm", "all"):
            bench_stream_gemm()
        if which in ("uncertainty", "all"):
            bench_uncertainty()
        if which == "uncertainty_first":
            bench_uncertainty(only_first=True)
