"""Build libdropdec.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m dropoutdecoding_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the working tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "_build")
LIB = os.path.join(OUT, "libdropdec.so")
SOURCES = ["dd_api.cu", "dd_sampler.cu", "dd_aggregate.cu", "dd_decoder.cu", "dd_skinny_gemm.cu", "dd_stream_gemm.cu",
           "dd_attention.cu", "dd_uncertainty.cu", "dd_pass.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    names = [n for n in sorted(os.listdir(CSRC)) if n.endswith((".cu", ".cuh", ".h"))]
    for name in names + ["../../include/dropdec.h"]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra_flags=(), variant: str = "") -> str:
    """variant != "" builds an experiment library (libdropdec_<variant>.so) with extra -D flags;
    select it at run time with DD_LIB=<path>."""
    out = os.path.join(OUT, variant) if variant else OUT
    lib = os.path.join(OUT, f"libdropdec_{variant}.so") if variant else LIB
    os.makedirs(out, exist_ok=True)
    stamp = os.path.join(out, "stamp")
    dig = _digest() + " ".join(extra_flags)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(out, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--force"]
    variant = next((a.split("=", 1)[1] for a in args if a.startswith("--variant=")), "")
    flags = [a for a in args if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=tuple(flags), variant=variant))
