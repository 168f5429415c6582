// Skinny GEMM for the single-token forwards of the ensemble: y[m,n] = sum_k x[m,k] w[n,k] with
// M = (#sequences x #members) <= 32 rows.  At these M the GEMM is a pure weight stream
// (intensity = M FLOP/byte << the B200 ridge of ~250), so the kernel is organised around the
// HBM stream of W, not around the tensor cores:
//   * one CTA owns 16 output features (32 in SwiGLU mode: 16 gate + 16 up rows) and all of K;
//   * its 8 warps split K; every lane issues 128-bit streaming loads (ld.global.nc.L1::no_allocate)
//     of W -- a warp load covers 8 rows x 64 contiguous bytes -- several k-steps ahead;
//   * the loaded registers ARE the mma.sync A fragments: within each 32-wide k block lane (g,t)
//     owns physical k = 8t..8t+7 and feeds elements {0,1},{2,3} to one m16n8k16 MMA and {4,5},{6,7}
//     to the next; the activations are loaded with the same per-lane permutation, so the dot
//     products are exact and no shuffle / smem transpose of W is ever needed;
//   * fp32 partial sums of the 8 warps are reduced through shared memory; the epilogue fuses the
//     bf16 rounding points of the HF decoder (residual add, SwiGLU, fp32 logits).
// Algorithmic bytes per launch: N*K*2 (x2 for SwiGLU) -- the weights, read exactly once.
#include "dd_common.cuh"

#ifndef DD_GEMM_PIPE
#define DD_GEMM_PIPE 0  // 1: two register batches in flight (measured slower: occupancy)
#endif

namespace dd {

constexpr int kGemmWarps = 8;
constexpr int kGemmThreads = kGemmWarps * 32;
// k-steps (of 32) per register batch; two batches are in flight per warp (software pipeline)


// kUnroll k-steps of this warp's weight rows -> registers (guards the tail of the k-slice)
template <int NW, int kUnroll>
__device__ __forceinline__ void load_w_batch(uint4 (&dst)[kUnroll][NW][2], const __nv_bfloat16* const (&wrow)[NW][2],
                                             int s0, int steps) {
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (s0 + u < steps) {
      const int ko = (s0 + u) * 32;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        dst[u][i][0] = ld_stream_128(wrow[i][0] + ko);
        dst[u][i][1] = ld_stream_128(wrow[i][1] + ko);
      }
    }
  }
}

template <int MT, int NW, int kUnroll>
__device__ __forceinline__ void mma_batch(float (&acc)[NW][MT][4], const uint4 (&wa)[kUnroll][NW][2],
                                          const __nv_bfloat16* const (&xrow)[MT], const bool (&xok)[MT], int s0,
                                          int steps) {
  uint4 xv[kUnroll][MT];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (s0 + u < steps) {
      const int ko = (s0 + u) * 32;
#pragma unroll
      for (int j = 0; j < MT; ++j)
        xv[u][j] = xok[j] ? *reinterpret_cast<const uint4*>(xrow[j] + ko) : make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (s0 + u < steps) {
#pragma unroll
      for (int i = 0; i < NW; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          mma_bf16_16816(acc[i][j], wa[u][i][0].x, wa[u][i][1].x, wa[u][i][0].y, wa[u][i][1].y, xv[u][j].x, xv[u][j].y);
          mma_bf16_16816(acc[i][j], wa[u][i][0].z, wa[u][i][1].z, wa[u][i][0].w, wa[u][i][1].w, xv[u][j].z, xv[u][j].w);
        }
    }
  }
}

template <int MT, int EPI>
__global__ void __launch_bounds__(kGemmThreads) skinny_gemm_kernel(const __nv_bfloat16* __restrict__ x,
                                                                   const __nv_bfloat16* __restrict__ w,
                                                                   const __nv_bfloat16* __restrict__ residual,
                                                                   int M, int N, int K, void* __restrict__ y) {
  constexpr int NW = (EPI == DD_EPI_SWIGLU) ? 2 : 1;  // weight row tiles per CTA
  // k-steps (of 32) whose weight loads are in flight per warp: 8 for a single token tile (measured best
  // on B200: 6.43 vs 6.57 ms per 7B ensemble step), fewer when more accumulators compete for registers
#ifdef DD_GEMM_UNROLL
  constexpr int kUnroll = DD_GEMM_UNROLL;
#else
  constexpr int kUnroll = DD_GEMM_PIPE ? ((MT * NW >= 3) ? 2 : 4) : (MT == 1 ? 8 : 4);
#endif
  __shared__ float s_red[kGemmWarps][NW][MT][16][8 + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 16;
  const int kslice = K / kGemmWarps;  // multiple of 32
  const int kbeg = warp * kslice;
  const int steps = kslice / 32;

  // per-lane row pointers (element offset 8t inside each 32-wide k block)
  const __nv_bfloat16* wrow[NW][2];
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    const size_t r0 = (size_t)(i * N + n0 + g);  // SwiGLU: up rows start at row N
    wrow[i][0] = w + r0 * K + kbeg + 8 * t;
    wrow[i][1] = w + (r0 + 8) * K + kbeg + 8 * t;
  }
  const __nv_bfloat16* xrow[MT];
  bool xok[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j) {
    const int m = 8 * j + g;
    xok[j] = m < M;
    xrow[j] = x + (size_t)(xok[j] ? m : 0) * K + kbeg + 8 * t;
  }

  float acc[NW][MT][4];
#pragma unroll
  for (int i = 0; i < NW; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  // PDL: the weights do not depend on the previous kernel -- put the first batch of the stream in
  // flight before waiting for our activations.
#if DD_GEMM_PIPE
  uint4 wa0[kUnroll][NW][2], wa1[kUnroll][NW][2];
  load_w_batch<NW, kUnroll>(wa0, wrow, 0, steps);
  pdl_trigger_stream_entry();
  pdl_wait();
  for (int s0 = 0; s0 < steps; s0 += 2 * kUnroll) {  // software-pipelined: next batch loads under the MMAs
    if (s0 + kUnroll < steps) load_w_batch<NW, kUnroll>(wa1, wrow, s0 + kUnroll, steps);
    mma_batch<MT, NW, kUnroll>(acc, wa0, xrow, xok, s0, steps);
    if (s0 + kUnroll < steps) {
      if (s0 + 2 * kUnroll < steps) load_w_batch<NW, kUnroll>(wa0, wrow, s0 + 2 * kUnroll, steps);
      mma_batch<MT, NW, kUnroll>(acc, wa1, xrow, xok, s0 + kUnroll, steps);
    }
  }
#else
  uint4 wa[kUnroll][NW][2];
  load_w_batch<NW, kUnroll>(wa, wrow, 0, steps);
  pdl_trigger_stream_entry();
  pdl_wait();
  for (int s0 = 0; s0 < steps; s0 += kUnroll) {
    if (s0 > 0) load_w_batch<NW, kUnroll>(wa, wrow, s0, steps);
    mma_batch<MT, NW, kUnroll>(acc, wa, xrow, xok, s0, steps);
  }
#endif
  pdl_trigger_stream_tail();

  // C fragment: c0,c1 = C[row g][tok 2t,2t+1]; c2,c3 = C[row g+8][tok 2t,2t+1]
#pragma unroll
  for (int i = 0; i < NW; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      s_red[warp][i][j][g][2 * t] = acc[i][j][0];
      s_red[warp][i][j][g][2 * t + 1] = acc[i][j][1];
      s_red[warp][i][j][g + 8][2 * t] = acc[i][j][2];
      s_red[warp][i][j][g + 8][2 * t + 1] = acc[i][j][3];
    }
  __syncthreads();

  // one thread per (token, feature): 8*MT tokens x 16 features
  for (int o = tid; o < MT * 8 * 16; o += kGemmThreads) {
    const int r = o & 15, tok = o >> 4;  // feature fastest: coalesced 32B stores per token
    if (tok >= M) continue;
    const int j = tok >> 3, tc = tok & 7;
    float v[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float s = 0.f;
#pragma unroll
      for (int wp = 0; wp < kGemmWarps; ++wp) s += s_red[wp][i][j][r][tc];
      v[i] = bf16_round(s);
    }
    const size_t oidx = (size_t)tok * N + n0 + r;
    if constexpr (EPI == DD_EPI_NONE) {
      reinterpret_cast<__nv_bfloat16*>(y)[oidx] = __float2bfloat16_rn(v[0]);
    } else if constexpr (EPI == DD_EPI_RESIDUAL) {
      reinterpret_cast<__nv_bfloat16*>(y)[oidx] = __float2bfloat16_rn(v[0] + __bfloat162float(residual[oidx]));
    } else if constexpr (EPI == DD_EPI_F32) {
      reinterpret_cast<float*>(y)[oidx] = v[0];
    } else {  // SwiGLU: bf16( bf16(silu(g)) * u )
      const float gte = v[0];
      const float act = bf16_round(gte / (1.f + __expf(-gte)));
      reinterpret_cast<__nv_bfloat16*>(y)[oidx] = __float2bfloat16_rn(act * v[NW - 1]);
    }
  }
}

template <int MT>
static int launch_mt(int epi, const void* x, const void* w, const void* residual, int M, int N, int K, void* y,
                     cudaStream_t st) {
  const dim3 grid(N / 16), block(kGemmThreads);
  auto X = (const __nv_bfloat16*)x;
  auto W = (const __nv_bfloat16*)w;
  auto R = (const __nv_bfloat16*)residual;
  switch (epi) {
    case DD_EPI_NONE: DD_LAUNCH_CHAIN(skinny_gemm_kernel<MT, DD_EPI_NONE>, grid, block, 0, st, X, W, R, M, N, K, y); break;
    case DD_EPI_RESIDUAL: DD_LAUNCH_CHAIN(skinny_gemm_kernel<MT, DD_EPI_RESIDUAL>, grid, block, 0, st, X, W, R, M, N, K, y); break;
    case DD_EPI_F32: DD_LAUNCH_CHAIN(skinny_gemm_kernel<MT, DD_EPI_F32>, grid, block, 0, st, X, W, R, M, N, K, y); break;
    case DD_EPI_SWIGLU: DD_LAUNCH_CHAIN(skinny_gemm_kernel<MT, DD_EPI_SWIGLU>, grid, block, 0, st, X, W, R, M, N, K, y); break;
    default: set_error("dd_skinny_gemm: bad epilogue %d", epi); return DD_ERR_INVALID_ARG;
  }
  return DD_OK;
}

}  // namespace dd

extern "C" int dd_skinny_gemm(const void* x, const void* w, const void* residual, int M, int N, int K,
                              int epilogue, void* y, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(x && w && y, "dd_skinny_gemm: null pointer");
  DD_CHECK_ARG(M >= 0 && M <= 32, "dd_skinny_gemm: M=%d out of [0,32]", M);
  DD_CHECK_ARG(N > 0 && N % 16 == 0, "dd_skinny_gemm: N=%d must be a positive multiple of 16", N);
  DD_CHECK_ARG(K > 0 && K % (32 * kGemmWarps) == 0, "dd_skinny_gemm: K=%d must be a multiple of %d", K,
               32 * kGemmWarps);
  DD_CHECK_ARG(epilogue != DD_EPI_RESIDUAL || residual, "dd_skinny_gemm: residual epilogue needs residual");
  if (M == 0) return DD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int mt = (M + 7) / 8;
  switch (mt) {
    case 1: return launch_mt<1>(epilogue, x, w, residual, M, N, K, y, st);
    case 2: return launch_mt<2>(epilogue, x, w, residual, M, N, K, y, st);
    case 3: return launch_mt<3>(epilogue, x, w, residual, M, N, K, y, st);
    default: return launch_mt<4>(epilogue, x, w, residual, M, N, K, y, st);
  }
}
