"""ctypes binding of libdropdec.so (include/dropdec.h).

There is no CPU fallback: if the library cannot be built or loaded, importing the ops fails
loudly.  Loading works without a GPU (the library links cudart statically and resolves the
driver lazily), so the CPU test-suite can check the exported symbols.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_i32p = C.c_void_p  # raw device pointers travel as integers
_P = C.c_void_p

class RopeArgs(C.Structure):
    """dd_rope_args of include/dropdec.h"""
    _fields_ = [("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p), ("seq_lens", C.c_void_p),
                ("members", C.c_int), ("H", C.c_int), ("H_kv", C.c_int), ("D", C.c_int),
                ("q_out", C.c_void_p), ("k_new", C.c_void_p), ("v_new", C.c_void_p),
                ("guess_winner", C.c_void_p), ("guess_pos", C.c_void_p), ("guess_members", C.c_int), ("S_max", C.c_int),
                ("k_cache", C.c_void_p), ("v_cache", C.c_void_p)]


class PassArgs(C.Structure):
    """dd_pass_args of include/dropdec.h"""
    _fields_ = [("plan", C.c_void_p), ("n_layers", C.c_int), ("rows", C.c_int), ("members_total", C.c_int),
                ("d_model", C.c_int), ("d_ff", C.c_int), ("H", C.c_int), ("H_kv", C.c_int), ("D", C.c_int),
                ("S_max", C.c_int), ("vocab", C.c_int), ("vocab_padded", C.c_int), ("eps", C.c_float),
                ("scale", C.c_float), ("embed", C.c_void_p), ("tok", C.c_void_p), ("tok_div", C.c_int),
                ("pos", C.c_void_p), ("pos_div", C.c_int), ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
                ("x", C.c_void_p), ("x2", C.c_void_p), ("act", C.c_void_p), ("q", C.c_void_p), ("attn", C.c_void_p),
                ("xn_out", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p), ("logits", C.c_void_p),
                ("vis_start", C.c_void_p), ("bitmask", C.c_void_p), ("n_vis", C.c_int), ("mask_words", C.c_int),
                ("num_splits", C.c_int), ("n_windows", C.c_int), ("win_first", C.c_int * 2),
                ("win_members", C.c_int * 2), ("win_seq_lens", C.c_void_p * 2), ("guess_winner", C.c_void_p),
                ("guess_pos", C.c_void_p), ("prefetch_gate_up_bytes", C.c_size_t), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("timeline", C.c_void_p)]


_SIGNATURES = {
    "dd_version": (C.c_int, []),
    "dd_last_error": (C.c_char_p, []),
    "dd_set_chain_launch": (C.c_int, [C.c_int]),
    "dd_hint_l2_prefetch": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t]),
    "dd_uncertainty_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "dd_uncertainty_overflow_offset": (C.c_size_t, [C.c_int] * 4),
    "dd_uncertainty_project": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "dd_sample_masks": (C.c_int, [_P, _P, C.c_int, _P, C.POINTER(C.c_double), C.c_int, _P, C.c_uint64,
                                  C.c_uint64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "dd_sample_masks_seq": (C.c_int, [_P, _P, C.c_int, _P, C.POINTER(C.c_double), C.c_int, _P, C.c_uint64,
                                      C.c_uint64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "dd_schedule_flags": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P, _P]),
    "dd_philox_uniform": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, _P, _P]),
    "dd_attention_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "dd_decode_attention": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P,
                                      C.c_size_t, _P]),
    "dd_decode_attention_active": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                             _P, _P, C.c_size_t, _P]),
    "dd_decode_attention_windows": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                              _P, _P, C.c_size_t, _P]),
    "dd_decode_attention_two_windows": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P,
                                                  C.c_size_t, _P]),
    "dd_commit_winner": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, _P]),
    "dd_aggregate": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dd_aggregate_strided": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, _P, _P, _P, _P, _P]),
    "dd_token_stats": (C.c_int, [_P, C.c_longlong, C.c_longlong, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "dd_token_stats_seq": (C.c_int, [_P, C.c_longlong, C.c_longlong, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "dd_step_advance_seq": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "dd_rmsnorm": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P]),
    "dd_skinny_gemm": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dd_stream_gemm_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "dd_stream_gemm": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "dd_decoder_linear": (C.c_int, [_P, _P, C.c_float, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                    C.POINTER(RopeArgs), _P]),
    "dd_rope_stage": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "dd_embed_gather": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dd_step_advance": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "dd_pass_plan_bytes": (C.c_size_t, [C.c_int]),
    "dd_pass_plan_fill": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                    C.c_int]),
    "dd_pass_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "dd_decoder_pass": (C.c_int, [C.POINTER(PassArgs), _P]),
    "dd_member_finish": (C.c_int, [_P, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int,
                                   _P, _P, _P, _P, _P, _P, _P]),
}

EXPORTED = tuple(_SIGNATURES)


class DropDecError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) libdropdec.so.  Raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("DD_LIB") or _build.LIB      # DD_LIB: an experiment build (build.py --variant=)
    if not os.path.exists(path):
        if not build_if_missing:
            raise DropDecError(f"{path} is missing: run `python -m dropoutdecoding_b200.build`")
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dd_last_error().decode(errors="replace")
        raise DropDecError(f"{what} failed (code {rc}): {msg}")
