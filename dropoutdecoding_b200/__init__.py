"""dropoutdecoding_b200 -- B200-native kernels and host code for the Dropout-Decoding hot path.

Layout:
  csrc/     hand-written sm_100a CUDA kernels + the C ABI (include/dropdec.h)
  _lib.py   ctypes binding of libdropdec.so (no CPU fallback)
  ops.py    torch-tensor front end of the C ABI
  engine.py decode engine (static KV cache, two-pass ensemble step, CUDA graphs)
  models/   host-side mirror of the reference's model-wrapper API
"""
__version__ = "0.1.0"
