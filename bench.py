#!/usr/bin/env python
"""Headline benchmark: LLaVA-1.5-7B Dropout Decoding (K=3 masks) tokens/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): LLaVA-1.5-7B shapes, random-init bf16, one synthetic 336px
image per sequence, 18-token prompt (593 positions after the 576 visual tokens are spliced in),
K=3 masked contexts with rates [0.3,0.5,0.7], 64 new tokens, EOS ignored.  Under N GPUs every rank
decodes its own sequences (batch sharding, SURVEY.md 8e: no per-step collective) -> weak scaling.

One "step" = one complete dropout-decode of the rank's batch: uncertainty stage + first token +
63 ensemble steps (each: preliminary pass, mask sampling, K masked passes over one shared KV cache,
vote, commit).  `value` times that with the prefix KV cache and the visual hidden states resident
in HBM (the prompt prefill is outside the hot path, reported as prefill_ms); `e2e` times the public
API `model.generate(...)` from pinned host tensors to host tokens: vision tower, projector, prefill,
hot path, and both copies.

--impl reference times the reference's algorithm (the CPU oracle: sequential members, one deep
copy of the KV cache per member, fp32) on the host cores; each step is a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT_LEN, N_VIS, NEW_TOKENS, PROBS = 18, 576, 64, [0.3, 0.5, 0.7]
IMAGE_POS = 5


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_llava15_7b(device, tiny: bool = False):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig
    from dropoutdecoding_b200.models.llava import CustomLlavaForConditionalGeneration
    if tiny:
        vc = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                              image_size=336, patch_size=14)
        tc = LlamaConfig(vocab_size=32064, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=4, rms_norm_eps=1e-5, max_position_embeddings=4096)
    else:
        vc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                              image_size=336, patch_size=14)
        tc = LlamaConfig(vocab_size=32064, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                         num_attention_heads=32, num_key_value_heads=32, rms_norm_eps=1e-5, max_position_embeddings=4096)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=32000, vision_feature_layer=-2,
                      vision_feature_select_strategy="default")
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = CustomLlavaForConditionalGeneration(cfg)
    finally:
        torch.set_default_dtype(old)
    return model.to(device=device, dtype=torch.bfloat16).eval()


def synthetic_inputs(batch: int, seed0: int):
    """Pinned host tensors, like a processor would hand over (one <image> id per prompt)."""
    ids, px = [], []
    for b in range(batch):
        g = torch.Generator().manual_seed(seed0 + b)
        row = torch.randint(3, 30000, (PROMPT_LEN,), generator=g)
        row[0] = 1
        row[IMAGE_POS] = 32000
        ids.append(row)
        img = torch.rand(3, 336, 336, generator=g)
        mean = torch.tensor([0.4815, 0.4578, 0.4082])[:, None, None]
        std = torch.tensor([0.2686, 0.2613, 0.2758])[:, None, None]
        px.append((img - mean) / std)
    return torch.stack(ids).pin_memory(), torch.stack(px).pin_memory()


def run_ours(args):
    from dropoutdecoding_b200 import ops
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL may print its version banner on stdout while the communicator is created; stdout carries exactly one
        # JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            torch.cuda.set_device(local)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    hbm, tf, peak_src = measured_peaks()
    B, T = args.batch, args.new_tokens

    model = build_llava15_7b(dev, tiny=args.tiny)
    ids_h, px_h = synthetic_inputs(B, seed0=1 + rank * B)

    with torch.no_grad():
        # ---- setup for `value`: prefill once, keep the hot path's inputs resident ------------------
        t0 = time.perf_counter()
        merged, start, n_vis = model._merge(ids_h.to(dev), px_h.to(dev, torch.bfloat16))
        S0 = merged.shape[1]
        eng = model._engine(B, n_vis, S0 + T, est_seq=S0 + T // 2)
        hidden = eng.prefill(merged.to(torch.bfloat16), [start] * B)
        vis_hidden = hidden[:, start:start + n_vis].contiguous()
        last_hidden = hidden[:, -1].contiguous()
        torch.cuda.synchronize()
        prefill_ms = (time.perf_counter() - t0) * 1e3

        def hot_path():
            eng.seq_lens.fill_(S0)
            eng.uncertainty_stage(vis_hidden)
            eng.first_token(last_hidden)
            eng.decode(T - 1, use_graph=True)

        for _ in range(args.warmup):
            hot_path()
        torch.cuda.synchronize()
        eng.launch_count = eng.n_spec_steps = eng.n_fallbacks = eng.n_chain = eng.n_chain_ok = 0
        if eng.spec_batch:
            eng._fb_total.zero_()
            eng._seq_steps_done = 0

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        clocks = ClockSampler(local)
        barrier()
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            hot_path()
        e1.record()
        barrier()
        dt = e0.elapsed_time(e1) * 1e-3
        clk = clocks.stop()
        tokens_check = eng.generated().clone()
        ens_launches = eng.launch_count          # kernels of libdropdec launched by the timed ensemble steps
        spec = {"mode": "speculative single pass (K+1 contexts per weight stream; masked pass re-run only when y_init "
                        "restores a drawn visual token)" if (eng.spec_active or eng.spec_batch) else "two passes per token",
                "steps": eng.n_spec_steps, "fallback_rate": round(eng.fallback_rate(), 4)}
        if getattr(eng, "lookahead", False):
            spec["lookahead"] = {"what": "the fallback pass of step t also carries the K+1 speculative contexts of step "
                                         "t+1 (2K+1 contexts), built on the guessed vote of step t",
                                 "passes": eng.n_chain, "guess_confirmed": eng.n_chain_ok}

        # ---- e2e: the public API, host tensors in, host tokens out ---------------------------------
        for _ in range(max(1, args.warmup - 1)):
            model.generate(input_ids=ids_h, pixel_values=px_h, max_new_tokens=T).cpu()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            out = model.generate(input_ids=ids_h, pixel_values=px_h, max_new_tokens=T).cpu()
        f1.record()
        barrier()
        dt_e2e = f0.elapsed_time(f1) * 1e-3
        assert out.shape == (B, PROMPT_LEN + T)
        # ---- a POPE-style answer (pope_test/pope_test.py:242 asks for ONE new token): all prompt side -- vision tower,
        #      prefill, uncertainty stage, first token; host tensors in, host tokens out
        pope = None
        if rank == 0 and not args.tiny:
            for _ in range(2):
                model.generate(input_ids=ids_h, pixel_values=px_h, max_new_tokens=1).cpu()
            torch.cuda.synchronize()
            t_p = time.perf_counter()
            for _ in range(5):
                one = model.generate(input_ids=ids_h, pixel_values=px_h, max_new_tokens=1).cpu()
            pope = {"what": "model.generate(max_new_tokens=1), host tensors in / out, mean of 5 calls",
                    "e2e_ms": round((time.perf_counter() - t_p) * 1e3 / 5, 3), "new_tokens": 1,
                    "token_equals_first_of_the_long_answer": bool(torch.equal(one[:, -1], out[:, PROMPT_LEN]))}

        # ---- roofline of the dominant kernel, measured live: one eager ensemble step with CUDA
        #      events around every dd_skinny_gemm launch on the launching stream
        roof = profile_dominant_kernel(eng, hbm, peak_src)
        stages = profile_stages(eng, vis_hidden, S0, hbm, tf)
        sched = schedule_check(model, eng, vis_hidden, last_hidden, S0, T, start, tokens_check)
        del eng
        model._engines.clear()
        torch.cuda.empty_cache()
    target = members = iblip = None
    if not args.no_target and not args.tiny:
        target = run_target(model, args, rank, world, local, dist, hbm, tf)
        if world == 1:
            iblip = run_instructblip(args, local, hbm)
    if (args.shard_members == "on" or (args.shard_members == "auto" and world > 1)) and 8 % world == 0 and not args.tiny:
        del model
        torch.cuda.empty_cache()
        members = run_member_sharded(args, rank, world, local, dist, hbm)

    if dist is not None:
        t = torch.tensor([dt, dt_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = t.tolist()
    total_tokens = B * T * args.steps * world
    line = {
        "metric": "LLaVA-1.5-7B dropout-decode tokens/s (K=3 masks)", "value": round(total_tokens / dt, 2),
        "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"llava-1.5-7b random-init bf16, {B} synthetic 336px image(s) per GPU, K=3 masks "
                               f"[0.3,0.5,0.7], {T} new tokens, prompt 18 ids -> {S0} positions" +
                               (" [TINY SHAPES]" if args.tiny else ""),
                   "batch_per_gpu": B, "new_tokens": T, "K": 3, "n_vis": n_vis,
                   "l2": "inputs larger than L2 (13.2 GB of weights streamed twice per token)",
                   "parallelism": f"batch-sharded x{world}, no per-step collective"},
        "e2e": {"value": round(total_tokens / dt_e2e, 2), "unit": "tokens/s",
                "h2d_bytes_per_step": int(ids_h.numel() * ids_h.element_size() + px_h.numel() * px_h.element_size()),
                "d2h_bytes_per_step": int(out.numel() * out.element_size())},
        "gpu_launches": int(ens_launches + args.steps * 6), "decode_schedule": spec,
        "prefill_ms": round(prefill_ms, 2), "pope_style": pope,
        "ms_per_token": round(dt / args.steps / T * 1e3, 4),
        "clocks": clk, "roofline": roof, "stage_rooflines": stages,
        "token_checksum": int(tokens_check.long().sum().item()),
        "schedule_check": sched,
    }
    if target is not None:
        line["target_config"] = target
    if iblip is not None:
        line["instructblip_config"] = iblip
    if members is not None:
        line["member_sharded_config"] = members
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, sample_steps=args.cpu_sample_steps)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _fresh_engine(model, eng, B, n_vis, max_seq, est_seq, speculative=None, seq_base=0, gemm_backend=None):
    """An engine on the model's weights with `eng`'s algorithm settings (own KV cache)."""
    from dropoutdecoding_b200.engine import DropoutDecodeEngine
    return DropoutDecodeEngine(model.decoder_weights(), B, n_vis, max_seq, probs=eng.probs, mode=eng.mode,
                               keep_relevant=eng.keep_relevant, topk=eng.topk, aggregate=eng.aggregate, seed=eng.seed,
                               max_steps=max_seq, est_seq=est_seq, vote_on=eng.vote_on, seq_base=seq_base,
                               speculative=speculative, gemm_backend=gemm_backend)


def schedule_check(model, eng, vis_hidden, last_hidden, S0, T, start, tokens_timed, reps=2):
    """Re-decode the timed inputs in the REFERENCE'S OWN ORDER (preliminary pass, then the K masked contexts:
    models/llava.py:294-305, 336-361; speculative="off") on a second engine fed the same prefix cache, and compare the
    token sequences with what the timed (speculative + lookahead) schedule produced."""
    B, n_vis = eng.B, eng.n_vis
    e2 = _fresh_engine(model, eng, B, n_vis, S0 + T, S0 + T // 2, speculative="off", seq_base=eng.seq_base)
    e2.load_prefill(eng.k_cache[:, :, :, :S0], eng.v_cache[:, :, :, :S0], [start] * B)

    def run():
        e2.seq_lens.fill_(S0)
        e2.uncertainty_stage(vis_hidden)
        e2.first_token(last_hidden)
        e2.decode(T - 1, use_graph=True)

    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / reps
    tok2 = e2.generated()
    equal = bool(torch.equal(tok2, tokens_timed))
    first_diff = None
    if not equal:
        d = (tok2 != tokens_timed).nonzero()
        first_diff = [int(d[0, 0]), int(d[0, 1])]
    del e2
    return {"what": "the same inputs decoded again in the reference's two-pass order (speculative off) on a second engine",
            "two_pass_tokens_equal": equal, "first_difference_seq_pos": first_diff,
            "two_pass_value": round(B * T / dt, 2), "two_pass_ms_per_token": round(dt / T * 1e3, 4)}


def run_target(model, args, rank, world, local, dist, hbm, tf):
    """north_star target / BASELINE configs[2]: 64 synthetic images (64 / N per GPU), K=3, 512 new tokens, EOS ignored.
    Sequences are sharded over the ranks with no per-step collective (the Philox stream is keyed by the global
    sequence index); the final token matrix is all-gathered inside the timed region."""
    from dropoutdecoding_b200 import ops
    dev = torch.device(f"cuda:{local}")
    total, T = args.target_batch, args.target_tokens
    B = total // world
    ids_h, px_h = synthetic_inputs(B, seed0=1001 + rank * B)
    out = {}
    with torch.no_grad():
        merged, start, n_vis = model._merge(ids_h.to(dev), px_h.to(dev, torch.bfloat16))
        S0 = merged.shape[1]
        probe = model._engine(1, n_vis, S0 + 8, est_seq=S0 + 4)          # algorithm settings of the wrapper
        eng = _fresh_engine(model, probe, B, n_vis, S0 + T, S0 + T // 2, seq_base=rank * B)
        model._engines.clear()
        del probe
        hidden = eng.prefill(merged.to(torch.bfloat16), [start] * B)
        vis_hidden = hidden[:, start:start + n_vis].contiguous()
        last_hidden = hidden[:, -1].contiguous()
        del hidden, merged
        gathered = [torch.zeros(B, T, dtype=torch.int32, device=dev) for _ in range(world)] if dist is not None else None

        def hot_path():
            eng.seq_lens.fill_(S0)
            eng.uncertainty_stage(vis_hidden)
            eng.first_token(last_hidden)
            eng.decode(T - 1, use_graph=True)
            if dist is not None:
                dist.all_gather(gathered, eng.generated().contiguous())

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        hot_path()                               # warm-up (captures the step graph)
        barrier()
        eng.launch_count = eng.n_spec_steps = 0
        if eng.spec_batch:
            eng._fb_total.zero_(); eng._seq_steps_done = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hot_path()
        e1.record()
        barrier()
        dt_rank = e0.elapsed_time(e1) * 1e-3
        times = torch.tensor([dt_rank], device=dev, dtype=torch.float64)
        per_rank = [times.clone() for _ in range(world)]
        if dist is not None:
            dist.all_gather(per_rank, times)
        per_rank_ms = [round(float(t.item()) * 1e3, 2) for t in per_rank]
        dt = max(per_rank_ms) * 1e-3
        toks = eng.generated()
        cfg = eng.cfg
        rows = B * ((2 * eng.K + 1) if eng.batch_la else (eng.K + 1 if (eng.spec_batch or eng.spec_active) else eng.K))
        passes = eng.n_spec_steps
        # per-pass roofline: weights once + K/V of every sequence once, at the measured HBM peak
        S_mean = S0 + T / 2
        pass_bytes = eng.w.stream_bytes_per_pass() + B * S_mean * 2 * cfg.n_layers * cfg.n_kv_heads * cfg.head_dim * 2
        out = {"workload": f"llava-1.5-7b random-init bf16, {total} synthetic images ({B} per GPU), K=3 masks, {T} new "
                           f"tokens, EOS ignored; north_star target at 1 GPU, BASELINE configs[2] at 2/4/8",
               "value": round(total * T / dt, 2), "unit": "tokens/s", "n_gpus": world, "batch_per_gpu": B,
               "new_tokens": T, "ms_per_token": round(dt / T * 1e3, 4), "per_rank_ms": per_rank_ms,
               "rows_per_pass": rows, "passes": passes,
               "gemm": (("dd::pair_gemm_kernel (qkv, gate/up, lm_head) + dd::stream_gemm_kernel (o_proj, down)" if rows > 128
                         else "dd::stream_gemm_kernel") if eng._use_stream(rows)
                        else ("dd::skinny_gemm_kernel" if rows <= 32 else "library")),
               "fallback_rate": round(eng.fallback_rate(), 4), "gpu_launches": int(eng.launch_count),
               "pass_roofline": {"bound": "hbm", "bytes_per_pass": int(pass_bytes), "passes": passes,
                                 "achieved": round(pass_bytes * passes / dt_rank / 1e9, 1), "peak": hbm, "unit": "GB/s",
                                 "frac": round(pass_bytes * passes / dt_rank / 1e9 / hbm, 4)},
               "token_checksum": int(toks.long().sum().item())}
        # ---- stage rooflines at this batch: decode attention, the widest GEMM of the pass, the uncertainty stage
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timeit(fn, iters=5):
            fn(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) * 1e-3)
            return best

        pbuf = eng.pe if eng.batch_la else (eng.pc if (eng.spec_batch or eng.spec_active) else eng.pb)
        S_att = int(S_mean)
        eng.seq_lens.fill_(S_att)
        mask = eng.la_mask if eng.batch_la else (eng.spec_mask if (eng.spec_batch or eng.spec_active) else eng.bitmask)
        G = cfg.n_heads // cfg.n_kv_heads
        mem = min(pbuf.M, max(1, 32 // G))
        t_att = timeit(lambda: ops.decode_attention(pbuf.q, pbuf.k_new[0], pbuf.v_new[0], eng.k_cache[0], eng.v_cache[0],
                                                    eng.seq_lens, eng.vis_start, mask, eng.n_vis, eng.scale, pbuf.splits,
                                                    pbuf.attn, pbuf.ws, member_offset=0, members=mem))
        kv = 2 * S_att * cfg.n_kv_heads * cfg.head_dim * 2 * B
        out["decode_attention"] = {"bound": "hbm", "kernel": "dd::attn_partial_kernel", "us": round(t_att * 1e6, 2),
                                   "members": mem, "seq_len": S_att, "achieved": round(kv / t_att / 1e9, 1), "peak": hbm,
                                   "unit": "GB/s", "frac": round(kv / t_att / 1e9 / hbm, 4), "algorithmic_bytes": kv,
                                   "traffic_note": "ncu dram read+write = 1.02 x the K/V bytes at B=64, S=849 "
                                                   "(profiles/r02/ncu_full_summary.json)"}
        if eng._use_stream(rows):
            lw = eng.w.layers[0]
            x = pbuf.xn[:rows]
            t_g = timeit(lambda: ops.stream_gemm(x, lw["wgu"], ops.EPI_SWIGLU, out=pbuf.act[:rows], workspace=eng._sg_ws))
            wb = lw["wgu"].numel() * 2
            out["stream_gemm_gate_up"] = {"bound": "hbm",
                                          "kernel": "dd::pair_gemm_kernel" if rows > 128 else "dd::stream_gemm_kernel",
                                          "rows": rows,
                                          "us": round(t_g * 1e6, 2), "achieved": round(wb / t_g / 1e9, 1), "peak": hbm,
                                          "unit": "GB/s", "frac": round(wb / t_g / 1e9 / hbm, 4), "algorithmic_bytes": wb,
                                          "traffic": int(1.03 * wb),
                                          "traffic_note": "ncu dram read+write at 192 rows = 1.03 x the weight bytes "
                                                          "(185.5 MB vs 180.4 MB; profiles/r02/ncu_full_summary.json)"}
        t_unc = timeit(lambda: eng.uncertainty_stage(vis_hidden), iters=3)
        flops = 2.0 * B * n_vis * cfg.d_model * cfg.vocab
        out["uncertainty_stage"] = {"bound": "tensor", "images": B, "ms": round(t_unc * 1e3, 3),
                                    "achieved": round(flops / t_unc / 1e12, 1), "peak": tf, "unit": "TFLOP/s",
                                    "frac": round(flops / t_unc / 1e12 / tf, 4), "algorithmic_flops": flops}
        del eng
        torch.cuda.empty_cache()
    return out


def run_instructblip(args, local, hbm):
    """BASELINE configs[3]: InstructBLIP-Vicuna-7B decoder shapes (vocab 32001, 32 Q-Former visual tokens), K=3, batch 32,
    256 new tokens on one GPU -- the flavour of models/instructblip.py: deterministic quantile masks, the vote on the
    argmax of the final hidden state, lm_head only for the winner (:447-460, :349-376).  Visual embeddings and the prompt are
    synthetic (the Q-Former is not part of the path); prefill, uncertainty stage and decode are inside the timed region."""
    from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine
    dev = torch.device(f"cuda:{local}")
    cfg = DecoderConfig(32, 4096, 32, 32, 128, 11008, 32001, 1e-5, 10000.0)
    B, n_vis, T, S0 = 32, 32, 256, 52
    with torch.no_grad():
        w = DecoderWeights.random(cfg, dev, seed=0, std=0.02, head_std=0.02)
        eng = DropoutDecodeEngine(w, B, n_vis, S0 + T, mode="quantile", topk=10, seed=506, max_steps=T + 8,
                                  est_seq=S0 + T // 2, vote_on="hidden")
        g = torch.Generator(device=dev).manual_seed(11)
        embeds = (torch.randn(B, S0, cfg.d_model, generator=g, device=dev) * 0.02).to(torch.bfloat16)

        def hot_path():
            hidden = eng.prefill(embeds, [10] * B)
            eng.uncertainty_stage(hidden[:, 10:10 + n_vis].contiguous())
            eng.first_token(hidden[:, -1].contiguous())
            eng.decode(T - 1, use_graph=True)

        hot_path()
        torch.cuda.synchronize()
        eng.launch_count = eng.n_spec_steps = 0
        if eng.spec_batch:
            eng._fb_total.zero_(); eng._seq_steps_done = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hot_path()
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        toks = eng.generated()
        rows = B * ((2 * eng.K + 1) if eng.batch_la else (eng.K + 1 if (eng.spec_batch or eng.spec_active) else eng.K))
        passes = eng.n_spec_steps
        pass_bytes = w.stream_bytes_per_pass() + B * (S0 + T / 2) * 2 * cfg.n_layers * cfg.n_kv_heads * cfg.head_dim * 2
        out = {"workload": f"instructblip-vicuna-7b decoder shapes random-init bf16, {B} sequences x {n_vis} synthetic "
                           f"Q-Former tokens, K=3 quantile masks, hidden-state vote, {T} new tokens, EOS ignored "
                           "(BASELINE configs[3])",
               "value": round(B * T / dt, 2), "unit": "tokens/s", "n_gpus": 1, "batch_per_gpu": B, "new_tokens": T,
               "ms_per_token": round(dt / T * 1e3, 4), "rows_per_pass": rows, "passes": passes,
               "gemm": "dd::stream_gemm_kernel" if eng._use_stream(rows) else ("dd::skinny_gemm_kernel" if rows <= 32 else "library"),
               "fallback_rate": round(eng.fallback_rate(), 4), "gpu_launches": int(eng.launch_count),
               "pass_roofline": {"bound": "hbm", "bytes_per_pass": int(pass_bytes), "passes": passes,
                                 "achieved": round(pass_bytes * passes / dt / 1e9, 1), "peak": hbm, "unit": "GB/s",
                                 "frac": round(pass_bytes * passes / dt / 1e9 / hbm, 4)},
               "token_checksum": int(toks.long().sum().item())}
    del eng, w
    torch.cuda.empty_cache()
    return out


def run_member_sharded(args, rank, world, local, dist, hbm):
    """BASELINE configs[4]: LLaVA-NeXT-Mistral-7B decoder shapes (GQA, 8 kv heads), up to 2880 anyres visual tokens, K=8
    masks, batch 16, the ensemble MEMBERS sharded over the GPUs (K/N members per rank; preliminary pass replicated), one
    NCCL all-gather per step, the whole step one CUDA graph.  The reference defines no K=8 mask rates
    (chair_test/chair_test.py:165-174 supports 1-4): linspace(0.1, 0.8, 8).  Visual embeddings are synthetic (the anyres
    vision tower is not part of the path)."""
    from dropoutdecoding_b200 import distributed as dd_dist
    from dropoutdecoding_b200.engine import DecoderConfig, DecoderWeights
    dev = torch.device(f"cuda:{local}")
    cfg = DecoderConfig.llava_next_mistral_7b()
    K, B, n_vis, T = 8, args.member_batch, 2880, args.member_tokens
    probs = [round(0.1 + 0.1 * i, 2) for i in range(K)]
    S0 = n_vis + 20
    with torch.no_grad():
        w = DecoderWeights.random(cfg, dev, seed=0, std=0.02, head_std=0.02)
        eng = dd_dist.make_member_sharded_engine(w, B, n_vis, S0 + T, probs, mode="independent", topk=10, seed=506,
                                                 max_steps=S0 + T, est_seq=S0 + T // 2)
        g = torch.Generator(device=dev).manual_seed(7)               # same seed on every rank: replicated inputs
        embeds = (torch.randn(B, S0, cfg.d_model, generator=g, device=dev) * 0.02).to(torch.bfloat16)
        hidden = eng.prefill(embeds, [10] * B)
        vis_hidden = hidden[:, 10:10 + n_vis].contiguous()
        last_hidden = hidden[:, -1].contiguous()
        del hidden, embeds

        def hot_path():
            eng.seq_lens.fill_(S0)
            eng.uncertainty_stage(vis_hidden)
            eng.first_token(last_hidden)
            eng.decode(T - 1)

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        hot_path()
        barrier()
        eng.launch_count = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hot_path()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        toks = eng.generated()
        # the collective alone (eager, back to back): microseconds per all-gather of the exchange records
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            eng._exchange()
        c0.record()
        for _ in range(50):
            eng._exchange()
        c1.record()
        barrier()
        coll_us = c0.elapsed_time(c1) * 1e3 / 50
        S_mean = S0 + T / 2
        kv_tok = 2 * cfg.n_layers * cfg.n_kv_heads * cfg.head_dim * 2
        step_bytes = 2 * eng.w.stream_bytes_per_pass() + 2 * B * S_mean * kv_tok     # two passes per token on every rank
        out = {"workload": f"llava-next-mistral-7b decoder shapes random-init bf16, {B} sequences x {n_vis} synthetic visual "
                           f"tokens, K={K} masks {probs} (independent), {T} new tokens; ensemble members sharded over {world} "
                           f"GPU(s), {K // world} per rank",
               "value": round(B * T / dt, 2), "unit": "tokens/s", "n_gpus": world, "ms_per_token": round(dt / T * 1e3, 4),
               "members_per_rank": K // world, "record_bytes_per_rank": int(eng._rec),
               "collective": {"what": "all_gather_into_tensor of one exchange record per rank (member ids, token statistics, "
                                      "staged K/V rows), consumed by dd_member_finish", "us_per_step_alone": round(coll_us, 1),
                              "bytes_gathered_per_step": int(eng._rec * world)},
               "step_in_cuda_graph": bool(eng.use_graph), "gpu_launches": int(eng.launch_count),
               "step_roofline": {"bound": "hbm", "bytes_per_step_per_gpu": int(step_bytes),
                                 "achieved": round(step_bytes * (T - 1) / dt / 1e9, 1), "peak": hbm, "unit": "GB/s",
                                 "frac": round(step_bytes * (T - 1) / dt / 1e9 / hbm, 4)},
               "token_checksum": int(toks.long().sum().item())}
        del eng, w
        torch.cuda.empty_cache()
    return out


def profile_dominant_kernel(eng, hbm, peak_src):
    """Time every weight-streaming GEMM launch (dd::skinny_gemm_kernel, through dd_skinny_gemm /
    dd_decoder_linear) of one eager ensemble step with CUDA events on torch's current stream (the stream
    the C ABI launches on)."""
    from dropoutdecoding_b200 import ops
    recs, calls = [], []
    orig_sg, orig_dl = ops.skinny_gemm, ops.decoder_linear

    def timed(orig):
        def f(x, w, *args, **kw):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = orig(x, w, *args, **kw)
            b.record()
            recs.append((a, b, w.numel() * w.element_size()))
            calls.append((orig, x, w, args, kw))
            return r
        return f

    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    snap = eng._snapshot()
    ops.skinny_gemm, ops.decoder_linear = timed(orig_sg), timed(orig_dl)
    try:
        one_step = eng.step if eng.spec_batch else eng._two_pass_step
        one_step()                      # warm
        recs.clear(); calls.clear()
        torch.cuda.synchronize()
        t_all0.record()
        one_step()
        t_all1.record()
        torch.cuda.synchronize()
    finally:
        ops.skinny_gemm, ops.decoder_linear = orig_sg, orig_dl
    # the same launches back to back, replayed from a CUDA graph that holds nothing else (how the product runs them)
    graph_ms = None
    if calls:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for orig, x, w, a_, kw in calls:
                    orig(x, w, *a_, **kw)
            g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                g.replay()
            g1.record()
            torch.cuda.synchronize()
            graph_ms = g0.elapsed_time(g1) / 3
        except Exception as e:          # noqa: BLE001 - the eager numbers stand on their own
            print(f"[bench] graph replay of the GEMM launches failed: {e}", file=sys.stderr)
    eng._restore(snap)
    if not recs:
        return {"bound": "hbm", "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None, "traffic": None,
                "note": "no dd::skinny_gemm_kernel launches (rows > 32: library GEMMs)"}
    secs = sum(a.elapsed_time(b) for a, b, _ in recs) * 1e-3
    byts = sum(n for _, _, n in recs)
    ach = byts / secs / 1e9
    eager = {"what": "one CUDA event pair per launch in an eager step (includes the host launch gap of every launch)",
             "achieved": round(ach, 1), "frac": round(ach / hbm, 4), "avg_launch_us": round(secs / len(recs) * 1e6, 2),
             "share_of_eager_step": round(secs / (t_all0.elapsed_time(t_all1) * 1e-3), 3)}
    out = {"bound": "hbm", "kernel": "dd::skinny_gemm_kernel (all weight GEMMs of one ensemble step)",
           "peak": hbm, "unit": "GB/s",
           # ncu --set full (profiles/r01/ncu_full_summary.json, "gemm"): dram__bytes_read + write per launch are
           # 33.6 / 105.2 / 184.3 / 94.0 MB for the o_proj / qkv / gate-up / down shapes = 1.03 x their weight bytes
           "traffic": int(1.03 * byts / len(recs)),
           "traffic_note": "bytes per launch, averaged over the step's shapes: ncu dram__bytes_read+write = 1.00-1.04 x "
                           "the weight bytes per shape (profiles/r01/ncu_full_summary.json)",
           "peak_source": peak_src, "launches": len(recs), "algorithmic_bytes_per_step": int(byts)}
    if graph_ms:
        # the product runs these launches from a CUDA graph: the average launch duration is taken from CUDA events
        # around a graph that holds exactly the step's GEMM launches, back to back (three replays)
        g_ach = byts / (graph_ms * 1e-3) / 1e9
        out.update({"achieved": round(g_ach, 1), "frac": round(g_ach / hbm, 4),
                    "avg_launch_us": round(graph_ms * 1e3 / len(recs), 2),
                    "how": "CUDA events around a CUDA graph holding the step's GEMM launches only, replayed three times",
                    "eager": eager})
    else:
        out.update({"achieved": eager["achieved"], "frac": eager["frac"], "avg_launch_us": eager["avg_launch_us"],
                    "how": eager["what"], "eager": eager})
    return out


def profile_stages(eng, vis_hidden, S0, hbm, tf):
    """The other two north_star kernels at this workload's shapes, CUDA events, L2 flushed before each:
    the tcgen05 uncertainty stage (tensor roofline, FLOPs counted once) and one decode-attention call
    (HBM roofline, K and V of every (sequence, kv head) read once)."""
    from dropoutdecoding_b200 import ops
    cfg, B = eng.cfg, eng.B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)

    def timeit(fn, iters=5):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        return best

    t_unc = timeit(lambda: eng.uncertainty_stage(vis_hidden))
    flops = 2.0 * B * vis_hidden.shape[1] * cfg.d_model * cfg.vocab
    pb = eng.pb
    eng.seq_lens.fill_(S0)
    t_att = timeit(lambda: ops.decode_attention(pb.q, pb.k_new[0], pb.v_new[0], eng.k_cache[0], eng.v_cache[0],
                                                eng.seq_lens, eng.vis_start, eng.bitmask, eng.n_vis, eng.scale,
                                                pb.splits, pb.attn, pb.ws))
    kv = 2 * S0 * cfg.n_kv_heads * cfg.head_dim * 2 * B
    return {"uncertainty_stage": {"bound": "tensor", "ms": round(t_unc * 1e3, 4), "achieved": round(flops / t_unc / 1e12, 1),
                                  "peak": tf, "unit": "TFLOP/s", "frac": round(flops / t_unc / 1e12 / tf, 4),
                                  "algorithmic_flops": flops},
            "decode_attention_K3": {"bound": "hbm", "us": round(t_att * 1e6, 2), "achieved": round(kv / t_att / 1e9, 1),
                                    "peak": hbm, "unit": "GB/s", "frac": round(kv / t_att / 1e9 / hbm, 4),
                                    "algorithmic_bytes": kv,
                                    "note": "latency-bound at this batch; the HBM-bound shape (B=64) is in profiles/"}}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def _oracle_setup(tiny: bool):
    import oracle
    cfg = oracle.LMConfig.tiny() if tiny else oracle.LMConfig.llava15_7b()
    torch.manual_seed(0)
    w = {"embed": torch.empty(cfg.vocab, cfg.d_model).normal_(std=0.02)}
    for l in range(cfg.n_layers):
        p = f"l{l}."
        w[p + "ln1"] = torch.ones(cfg.d_model); w[p + "ln2"] = torch.ones(cfg.d_model)
        for name, shape in (("wq", (cfg.n_heads * cfg.head_dim, cfg.d_model)),
                            ("wk", (cfg.n_kv_heads * cfg.head_dim, cfg.d_model)),
                            ("wv", (cfg.n_kv_heads * cfg.head_dim, cfg.d_model)),
                            ("wo", (cfg.d_model, cfg.n_heads * cfg.head_dim)),
                            ("wg", (cfg.d_ff, cfg.d_model)), ("wu", (cfg.d_ff, cfg.d_model)),
                            ("wd", (cfg.d_model, cfg.d_ff))):
            w[p + name] = torch.empty(*shape).normal_(std=0.02)
    w["norm"] = torch.ones(cfg.d_model)
    w["lm_head"] = torch.empty(cfg.vocab, cfg.d_model).normal_(std=0.02)
    lm = oracle.OracleLM(cfg, w, faithful=False, max_pos=2048)
    S0 = PROMPT_LEN - 1 + N_VIS
    cache = oracle.OracleCache(cfg.n_layers)
    g = torch.Generator().manual_seed(1)
    for l in range(cfg.n_layers):       # a prefilled prefix: random K/V of the right shape
        cache.k[l] = torch.randn(cfg.n_kv_heads, S0, cfg.head_dim, generator=g)
        cache.v[l] = torch.randn(cfg.n_kv_heads, S0, cfg.head_dim, generator=g)
    epis = torch.rand(N_VIS, generator=g) * 0.2 + 0.75
    topk_ids = torch.randint(0, cfg.vocab, (1, N_VIS, 5), generator=g)
    return oracle, lm, cache, epis, topk_ids


def _oracle_sample(oracle, lm, cache, epis, topk_ids, n_tokens: int) -> float:
    import copy
    g = torch.Generator().manual_seed(24)
    U = torch.rand(n_tokens, len(PROBS), N_VIS, generator=g)
    c = copy.deepcopy(cache)
    t0 = time.perf_counter()
    with torch.no_grad():
        oracle.ensemble_steps(lm, c, 1234, epis, topk_ids, IMAGE_POS, N_VIS, n_tokens, PROBS, U)
    return time.perf_counter() - t0


def cpu_baseline(args, sample_steps: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle, lm, cache, epis, topk_ids = _oracle_setup(args.tiny)
    _oracle_sample(oracle, lm, cache, epis, topk_ids, 1)              # warm-up token
    dt = _oracle_sample(oracle, lm, cache, epis, topk_ids, sample_steps)
    return {"value": round(sample_steps / dt, 4), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"{sample_steps} ensemble steps (1 preliminary + 3 masked fp32 forwards, 4 KV deep copies "
                      f"each) of the same workload from a 593-token prefix, decode phase only, torch CPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle, lm, cache, epis, topk_ids = _oracle_setup(args.tiny)
    n_tok = args.cpu_sample_steps
    for _ in range(min(args.warmup, 1)):
        _oracle_sample(oracle, lm, cache, epis, topk_ids, 1)
    dt = 0.0
    for _ in range(args.steps):
        dt += _oracle_sample(oracle, lm, cache, epis, topk_ids, n_tok)
    val = round(n_tok * args.steps / dt, 4)
    sample = (f"each step = {n_tok} ensemble tokens (1 preliminary + 3 masked fp32 forwards, 4 KV deep copies per "
              f"token) of the same workload from a 593-token prefix, decode phase only")
    print(json.dumps({
        "impl": "reference", "metric": "LLaVA-1.5-7B dropout-decode tokens/s (K=3 masks)", "value": val,
        "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "llava-1.5-7b random-init, 1 synthetic image, K=3 masks [0.3,0.5,0.7], reference "
                               "algorithm on the host CPU" + (" [TINY SHAPES]" if args.tiny else "")},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="sequences per GPU")
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS)
    ap.add_argument("--tiny", action="store_true", help="tiny shapes (debugging; not a valid bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-target", action="store_true", help="skip the 64-image / 512-token target configuration")
    ap.add_argument("--target-batch", type=int, default=64, help="images of the target configuration, all GPUs together")
    ap.add_argument("--target-tokens", type=int, default=512)
    ap.add_argument("--shard-members", default="auto", choices=["auto", "on", "off"],
                    help="also run BASELINE configs[4] with the K=8 ensemble members sharded over the GPUs (auto: when N > 1)")
    ap.add_argument("--member-batch", type=int, default=16)
    ap.add_argument("--member-tokens", type=int, default=48)
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
