// Shared helpers for the libdropdec kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/dropdec.h"

namespace dd {

void set_error(const char* fmt, ...);

// cuTensorMapEncodeTiled for a dense bf16 tensor with SWIZZLE_128B boxes (driver entry point resolved through
// the runtime, so the library does not link libcuda).  dims / box: innermost first; strides_bytes: rank - 1
// entries.  Returns DD_OK or DD_ERR_CUDA (message set).
int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const char* what);

#define DD_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::dd::set_error(__VA_ARGS__);        \
      return DD_ERR_INVALID_ARG;           \
    }                                      \
  } while (0)

#define DD_CHECK_CUDA(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::dd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DD_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define DD_CHECK_LAUNCH() DD_CHECK_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  One ensemble step is a chain of ~600 short kernels; with
// plain stream order every boundary costs a launch + ramp-up bubble in which HBM idles.  All chain
// kernels are launched with programmaticStreamSerialization: a kernel's CTAs may become resident
// while its predecessor is still running, do work that does not depend on it (address setup, the
// first batch of weight loads) and then block in pdl_wait() until the predecessor has completed
// and flushed.  Rule: nothing written by an earlier kernel is touched before pdl_wait(), and
// nothing is written before it either.  The host chooses per pass (dd_set_chain_launch): it pays where the
// dependent kernel has a deep prefetch to run ahead with -- dd_stream_gemm fills its whole shared-memory ring
// with weight tiles before the wait (7B pass of 56 rows: 4.99 -> 4.55 ms) -- and costs 1-3 % on the
// register-streamed skinny GEMMs of a 4-row pass, where it stays off (profiles/README.md).
// ---------------------------------------------------------------------------------------------
// weight_gemm: the launch is a weight-streaming GEMM (mode 2 of dd_set_chain_launch chains only those)
bool pdl_enabled(bool weight_gemm = false);

template <bool kWeightGemm = false, typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(kWeightGemm) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define DD_LAUNCH_CHAIN(...) DD_CHECK_CUDA(::dd::launch_chain(__VA_ARGS__))
#define DD_LAUNCH_CHAIN_GEMM(...) DD_CHECK_CUDA(::dd::launch_chain<true>(__VA_ARGS__))
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// DD_PDL_TRIGGER: where kernels let their dependents start (build-time knob for experiments)
//   0 = never explicitly (dependents start when the grid has drained)
//   1 = at kernel entry everywhere (deep pile-up of waiting grids: measured slower)
//   2 = small kernels at entry, streaming kernels (GEMM, attention) after their main loop  [default]
#ifndef DD_PDL_TRIGGER
#define DD_PDL_TRIGGER 2
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger_now() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// entry of a small kernel
__device__ __forceinline__ void pdl_trigger() {
  if (DD_PDL_TRIGGER >= 1) pdl_trigger_now();
}
// entry / end-of-main-loop of a streaming kernel
__device__ __forceinline__ void pdl_trigger_stream_entry() {
  if (DD_PDL_TRIGGER == 1) pdl_trigger_now();
}
__device__ __forceinline__ void pdl_trigger_stream_tail() {
  if (DD_PDL_TRIGGER == 2) pdl_trigger_now();
}

// ---------------------------------------------------------------------------------------------
// L2 prefetch riding on the latency-bound kernels of a small pass (dd_hint_l2_prefetch, include/dropdec.h).
// With 4-16 rows the pass is one long HBM stream of weights, interrupted by kernels that leave HBM idle (decode
// attention: ~14 us on ~10 MB of K/V; RMSNorm: ~4 us on a few rows).  Those kernels carry up to two address
// ranges -- the head of the weight stream of the GEMMs that follow -- and their otherwise idle threads issue
// cp.async.bulk.prefetch.L2 over them (fire and forget: no registers, no shared memory, no completion to wait
// for), so the stream keeps running into the 126 MB L2 and the next GEMM starts on L2 hits.
// ---------------------------------------------------------------------------------------------
struct L2Prefetch {
  const char* ptr[2];
  unsigned long long bytes[2];  // multiples of 16; 0 = unused
};
// the hint left by dd_hint_l2_prefetch on this host thread; reading it clears it
L2Prefetch take_l2_prefetch();

constexpr unsigned kPrefetchChunk = 16384;  // bytes per prefetch instruction
// The instruction takes its operands from uniform registers, so a warp is the natural worker (one lane issues; per-lane
// addresses would be serialised by the compiler anyway): warp w of n_warps across the grid covers chunks
// w, w + n_warps, ... of both ranges.  Call with all lanes of the warp.
__device__ __forceinline__ void l2_prefetch_slice(const L2Prefetch& pf, unsigned warp_idx, unsigned n_warps) {
  if ((threadIdx.x & 31) != 0) return;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const unsigned long long bytes = pf.bytes[r];
    for (unsigned long long off = (unsigned long long)warp_idx * kPrefetchChunk; off < bytes;
         off += (unsigned long long)n_warps * kPrefetchChunk) {
      const unsigned long long left = bytes - off;
      const unsigned sz = left < kPrefetchChunk ? (unsigned)left : kPrefetchChunk;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf.ptr[r] + off), "r"(sz) : "memory");
    }
  }
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 128-bit streaming load (weights / KV are read once: keep them out of L1)
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
  float2 f;
  f.x = __uint_as_float(v << 16);
  f.y = __uint_as_float(v & 0xffff0000u);
  return f;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                               uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

}  // namespace dd
