#!/usr/bin/env python
"""Where does model.generate() spend its time?  (bench.py's e2e number: host tensors in, host tokens out.)"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dropoutdecoding_b200.engine import DropoutDecodeEngine

dev = torch.device("cuda:0")
model = bench.build_llava15_7b(dev)
ids_h, px_h = bench.synthetic_inputs(1, 1)
times = {}

def wrap(obj, name):
    orig = getattr(obj, name)
    def f(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize(); times[name] = times.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    setattr(obj, name, f)

for n in ("prefill", "uncertainty_stage", "first_token", "decode", "capture", "set_uniforms"):
    wrap(DropoutDecodeEngine, n)
wrap(type(model), "_merge"); wrap(type(model), "_engine")
with torch.no_grad():
    for i in range(3):
        times.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = model.generate(input_ids=ids_h, pixel_values=px_h, max_new_tokens=64).cpu()
        total = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"call": i, "total_ms": round(total, 1), **{k: round(v, 1) for k, v in times.items()}}))
