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
//
// dd_decoder_linear adds the two fusions that remove the small kernels between the GEMMs of a layer:
//   * NORM prologue: the RMSNorm of the input rows is applied while the B fragments are built (every
//     CTA recomputes the row's 1/rms from L2-resident x while its first weight batch is in flight);
//   * ROPE epilogue (QKV projection): a CTA owns features {8j..8j+7} and {D/2+8j..D/2+8j+7} of one head,
//     i.e. complete rotate-half pairs, applies RoPE and writes q / k_new / v_new in the staging layout.
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

// bf16( bf16(x * rstd) * w ) on a packed pair -- the two rounding points of LlamaRMSNorm
__device__ __forceinline__ uint32_t norm_pair(uint32_t x, uint32_t w, float rstd) {
  const float2 f = bf16x2_to_float2(x), ww = bf16x2_to_float2(w);
  return pack_bf16x2(bf16_round(f.x * rstd) * ww.x, bf16_round(f.y * rstd) * ww.y);
}

template <int MT, int NW, int kUnroll, bool NORM>
__device__ __forceinline__ void mma_batch(float (&acc)[NW][MT][4], const uint4 (&wa)[kUnroll][NW][2],
                                          const __nv_bfloat16* const (&xrow)[MT], const bool (&xok)[MT], int s0,
                                          int steps, const __nv_bfloat16* nrow, const float (&rstd)[MT]) {
  uint4 xv[kUnroll][MT];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (s0 + u < steps) {
      const int ko = (s0 + u) * 32;
#pragma unroll
      for (int j = 0; j < MT; ++j)
        xv[u][j] = xok[j] ? *reinterpret_cast<const uint4*>(xrow[j] + ko) : make_uint4(0, 0, 0, 0);
      if (NORM) {
        const uint4 nw = *reinterpret_cast<const uint4*>(nrow + ko);
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          xv[u][j].x = norm_pair(xv[u][j].x, nw.x, rstd[j]);
          xv[u][j].y = norm_pair(xv[u][j].y, nw.y, rstd[j]);
          xv[u][j].z = norm_pair(xv[u][j].z, nw.z, rstd[j]);
          xv[u][j].w = norm_pair(xv[u][j].w, nw.w, rstd[j]);
        }
      }
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

struct GemmArgs {
  const __nv_bfloat16 *x, *w, *residual, *norm_w;
  float eps;
  int M, N, K;
  void* y;
  // DD_EPI_ROPE
  const float *cos_tab, *sin_tab;
  const int32_t* seq_lens;
  int members, H, H_kv, D;
  __nv_bfloat16 *q_out, *k_new, *v_new;
  // commit ahead (lookahead pass): row b * guess_members + guess_winner[b] also goes to cache row guess_pos[b]
  const int32_t *guess_winner, *guess_pos;
  int guess_members, S_max;
  __nv_bfloat16 *k_cache, *v_cache;
};

template <int MT, int EPI, bool NORM>
__global__ void __launch_bounds__(kGemmThreads) skinny_gemm_kernel(const GemmArgs a) {
  const __nv_bfloat16* __restrict__ x = a.x;
  const __nv_bfloat16* __restrict__ w = a.w;
  const __nv_bfloat16* __restrict__ residual = a.residual;
  const int M = a.M, N = a.N, K = a.K;
  void* __restrict__ y = a.y;
  constexpr int NW = (EPI == DD_EPI_SWIGLU) ? 2 : 1;  // weight row tiles per CTA
  // k-steps (of 32) whose weight loads are in flight per warp: 8 for a single token tile (measured best
  // on B200: 6.43 vs 6.57 ms per 7B ensemble step), fewer when more accumulators compete for registers
#ifdef DD_GEMM_UNROLL
  constexpr int kUnroll = DD_GEMM_UNROLL;
#else
  constexpr int kUnroll = DD_GEMM_PIPE ? ((MT * NW >= 3) ? 2 : 4) : (MT == 1 ? 8 : 4);
#endif
  __shared__ float s_red[kGemmWarps][NW][MT][16][8 + 1];
  __shared__ float s_ss[NORM ? kGemmWarps : 1][MT][8];

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
    size_t r0 = (size_t)(i * N + n0 + g);  // SwiGLU: up rows start at row N
    size_t r1 = r0 + 8;
    if (EPI == DD_EPI_ROPE && n0 < (a.H + a.H_kv) * a.D) {
      // q / k head: rows 8j+g and D/2+8j+g of the head -> the CTA holds whole rotate-half pairs
      const int tph = a.D / 16, head = blockIdx.x / tph, jj = blockIdx.x % tph;
      r0 = (size_t)head * a.D + 8 * jj + g;
      r1 = r0 + a.D / 2;
    }
    wrow[i][0] = w + r0 * K + kbeg + 8 * t;
    wrow[i][1] = w + r1 * K + kbeg + 8 * t;
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

  const __nv_bfloat16* nrow = NORM ? a.norm_w + kbeg + 8 * t : nullptr;
  float rstd[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j) rstd[j] = 1.f;
  // 1/rms of the token rows, computed by the whole CTA with every load independent (one L2 round trip,
  // hidden under the first weight batch): thread -> 16-byte pieces tid, tid+256, ... of each row.
  auto row_rstd = [&]() {
    if (!NORM) return;
    const int nvec = K / 8;
    for (int m0 = 0; m0 < M; m0 += 4) {  // 4 rows at a time bounds the registers
      float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (m0 + r < M) {
          const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)(m0 + r) * K);
#pragma unroll 4
          for (int v = tid; v < nvec; v += kGemmThreads) {
            const uint4 q = xr[v];
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = bf16x2_to_float2(u[k]);
              ss[r] = fmaf(f.x, f.x, ss[r]);
              ss[r] = fmaf(f.y, f.y, ss[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float tot = warp_sum(ss[r]);
        if (lane == 0 && m0 + r < M) s_ss[warp][(m0 + r) >> 3][(m0 + r) & 7] = tot;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      float tot = 0.f;
#pragma unroll
      for (int wp = 0; wp < kGemmWarps; ++wp) tot += s_ss[wp][j][g];
      rstd[j] = xok[j] ? rsqrtf(tot / (float)K + a.eps) : 0.f;
    }
  };
  // PDL: the weights do not depend on the previous kernel -- put the first batch of the stream in
  // flight before waiting for our activations.
#if DD_GEMM_PIPE
  uint4 wa0[kUnroll][NW][2], wa1[kUnroll][NW][2];
  load_w_batch<NW, kUnroll>(wa0, wrow, 0, steps);
  pdl_trigger_stream_entry();
  pdl_wait();
  row_rstd();
  for (int s0 = 0; s0 < steps; s0 += 2 * kUnroll) {  // software-pipelined: next batch loads under the MMAs
    if (s0 + kUnroll < steps) load_w_batch<NW, kUnroll>(wa1, wrow, s0 + kUnroll, steps);
    mma_batch<MT, NW, kUnroll, NORM>(acc, wa0, xrow, xok, s0, steps, nrow, rstd);
    if (s0 + kUnroll < steps) {
      if (s0 + 2 * kUnroll < steps) load_w_batch<NW, kUnroll>(wa0, wrow, s0 + 2 * kUnroll, steps);
      mma_batch<MT, NW, kUnroll, NORM>(acc, wa1, xrow, xok, s0 + kUnroll, steps, nrow, rstd);
    }
  }
#else
  uint4 wa[kUnroll][NW][2];
  load_w_batch<NW, kUnroll>(wa, wrow, 0, steps);
  pdl_trigger_stream_entry();
  pdl_wait();
  row_rstd();
  for (int s0 = 0; s0 < steps; s0 += kUnroll) {
    if (s0 > 0) load_w_batch<NW, kUnroll>(wa, wrow, s0, steps);
    mma_batch<MT, NW, kUnroll, NORM>(acc, wa, xrow, xok, s0, steps, nrow, rstd);
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

  if constexpr (EPI == DD_EPI_ROPE) {
    auto red = [&](int j, int r, int tc) {
      float s = 0.f;
#pragma unroll
      for (int wp = 0; wp < kGemmWarps; ++wp) s += s_red[wp][0][j][r][tc];
      return bf16_round(s);  // the projection's own bf16 output
    };
    if (n0 < (a.H + a.H_kv) * a.D) {
      // one thread per (token, pair): features i = 8jj+r and i + D/2 of one head
      const int tph = a.D / 16, head = blockIdx.x / tph, jj = blockIdx.x % tph, half = a.D / 2;
      for (int o = tid; o < MT * 8 * 8; o += kGemmThreads) {
        const int r = o & 7, tok = o >> 3;
        if (tok >= M) continue;
        const int j = tok >> 3, tc = tok & 7, i = 8 * jj + r;
        const float x1 = red(j, r, tc), x2 = red(j, r + 8, tc);
        const int pos = a.seq_lens[tok / a.members];
        const float c = a.cos_tab[(size_t)pos * half + i], sn = a.sin_tab[(size_t)pos * half + i];
        __nv_bfloat16* dst = head < a.H ? a.q_out + ((size_t)tok * a.H + head) * a.D
                                        : a.k_new + ((size_t)tok * a.H_kv + (head - a.H)) * a.D;
        const __nv_bfloat16 lo = __float2bfloat16_rn(x1 * c - x2 * sn), hi = __float2bfloat16_rn(x2 * c + x1 * sn);
        dst[i] = lo;
        dst[i + half] = hi;
        if (a.guess_winner && head >= a.H) {
          const int b = tok / a.guess_members, gp = a.guess_pos[b];
          if (tok - b * a.guess_members == a.guess_winner[b] && gp < a.S_max) {
            __nv_bfloat16* c2 = a.k_cache + (((size_t)b * a.H_kv + (head - a.H)) * a.S_max + gp) * a.D;
            c2[i] = lo;
            c2[i + half] = hi;
          }
        }
      }
    } else {
      for (int o = tid; o < MT * 8 * 16; o += kGemmThreads) {
        const int r = o & 15, tok = o >> 4;
        if (tok >= M) continue;
        const int n = n0 + r - (a.H + a.H_kv) * a.D;  // feature inside the v block
        const __nv_bfloat16 v = __float2bfloat16_rn(red(tok >> 3, r, tok & 7));
        a.v_new[(size_t)tok * a.H_kv * a.D + n] = v;
        if (a.guess_winner) {
          const int b = tok / a.guess_members, gp = a.guess_pos[b];
          if (tok - b * a.guess_members == a.guess_winner[b] && gp < a.S_max)
            a.v_cache[(((size_t)b * a.H_kv + n / a.D) * a.S_max + gp) * a.D + n % a.D] = v;
        }
      }
    }
    return;
  }
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

template <int MT, bool NORM>
static int launch_mt(const GemmArgs& a, int epi, cudaStream_t st) {
  const dim3 grid(a.N / 16), block(kGemmThreads);
  switch (epi) {
    case DD_EPI_NONE: DD_LAUNCH_CHAIN_GEMM(skinny_gemm_kernel<MT, DD_EPI_NONE, NORM>, grid, block, 0, st, a); break;
    case DD_EPI_RESIDUAL: DD_LAUNCH_CHAIN_GEMM(skinny_gemm_kernel<MT, DD_EPI_RESIDUAL, NORM>, grid, block, 0, st, a); break;
    case DD_EPI_F32: DD_LAUNCH_CHAIN_GEMM(skinny_gemm_kernel<MT, DD_EPI_F32, NORM>, grid, block, 0, st, a); break;
    case DD_EPI_SWIGLU: DD_LAUNCH_CHAIN_GEMM(skinny_gemm_kernel<MT, DD_EPI_SWIGLU, NORM>, grid, block, 0, st, a); break;
    case DD_EPI_ROPE: DD_LAUNCH_CHAIN_GEMM(skinny_gemm_kernel<MT, DD_EPI_ROPE, NORM>, grid, block, 0, st, a); break;
    default: set_error("dd_skinny_gemm: bad epilogue %d", epi); return DD_ERR_INVALID_ARG;
  }
  return DD_OK;
}

template <bool NORM>
static int launch_any(const GemmArgs& a, int epi, cudaStream_t st) {
  switch ((a.M + 7) / 8) {
    case 1: return launch_mt<1, NORM>(a, epi, st);
    case 2: return launch_mt<2, NORM>(a, epi, st);
    case 3: return launch_mt<3, NORM>(a, epi, st);
    default: return launch_mt<4, NORM>(a, epi, st);
  }
}

static int check_gemm_shape(const char* fn, const void* x, const void* w, int M, int N, int K) {
  DD_CHECK_ARG(x && w, "%s: null pointer", fn);
  DD_CHECK_ARG(M >= 0 && M <= 32, "%s: M=%d out of [0,32]", fn, M);
  DD_CHECK_ARG(N > 0 && N % 16 == 0, "%s: N=%d must be a positive multiple of 16", fn, N);
  DD_CHECK_ARG(K > 0 && K % (32 * kGemmWarps) == 0, "%s: K=%d must be a multiple of %d", fn, K, 32 * kGemmWarps);
  return DD_OK;
}

}  // namespace dd

extern "C" int dd_skinny_gemm(const void* x, const void* w, const void* residual, int M, int N, int K,
                              int epilogue, void* y, dd_stream_t stream) {
  using namespace dd;
  if (int rc = check_gemm_shape("dd_skinny_gemm", x, w, M, N, K)) return rc;
  DD_CHECK_ARG(y, "dd_skinny_gemm: null pointer");
  DD_CHECK_ARG(epilogue >= DD_EPI_NONE && epilogue <= DD_EPI_SWIGLU, "dd_skinny_gemm: bad epilogue %d", epilogue);
  DD_CHECK_ARG(epilogue != DD_EPI_RESIDUAL || residual, "dd_skinny_gemm: residual epilogue needs residual");
  if (M == 0) return DD_OK;
  GemmArgs a{};
  a.x = (const __nv_bfloat16*)x; a.w = (const __nv_bfloat16*)w; a.residual = (const __nv_bfloat16*)residual;
  a.M = M; a.N = N; a.K = K; a.y = y;
  return launch_any<false>(a, epilogue, (cudaStream_t)stream);
}

extern "C" int dd_decoder_linear(const void* x, const void* norm_w, float eps, const void* w, const void* residual,
                                 int M, int N, int K, int epilogue, void* y, const dd_rope_args* rope,
                                 dd_stream_t stream) {
  using namespace dd;
  if (int rc = check_gemm_shape("dd_decoder_linear", x, w, M, N, K)) return rc;
  DD_CHECK_ARG(epilogue >= DD_EPI_NONE && epilogue <= DD_EPI_ROPE, "dd_decoder_linear: bad epilogue %d", epilogue);
  DD_CHECK_ARG(epilogue == DD_EPI_ROPE || y, "dd_decoder_linear: null output");
  DD_CHECK_ARG(epilogue != DD_EPI_RESIDUAL || residual, "dd_decoder_linear: residual epilogue needs residual");
  GemmArgs a{};
  a.x = (const __nv_bfloat16*)x; a.w = (const __nv_bfloat16*)w; a.residual = (const __nv_bfloat16*)residual;
  a.norm_w = (const __nv_bfloat16*)norm_w; a.eps = eps;
  a.M = M; a.N = N; a.K = K; a.y = y;
  if (epilogue == DD_EPI_ROPE) {
    DD_CHECK_ARG(rope && rope->cos_tab && rope->sin_tab && rope->seq_lens && rope->q_out && rope->k_new && rope->v_new,
                 "dd_decoder_linear: DD_EPI_ROPE needs a complete dd_rope_args");
    DD_CHECK_ARG(rope->members >= 1 && rope->D % 16 == 0 && rope->H >= 1 && rope->H_kv >= 1 &&
                     N == (rope->H + 2 * rope->H_kv) * rope->D,
                 "dd_decoder_linear: N=%d is not (H + 2 H_kv) * D", N);
    a.cos_tab = rope->cos_tab; a.sin_tab = rope->sin_tab; a.seq_lens = rope->seq_lens;
    a.members = rope->members; a.H = rope->H; a.H_kv = rope->H_kv; a.D = rope->D;
    a.q_out = (__nv_bfloat16*)rope->q_out; a.k_new = (__nv_bfloat16*)rope->k_new; a.v_new = (__nv_bfloat16*)rope->v_new;
    if (rope->guess_winner) {
      DD_CHECK_ARG(rope->guess_pos && rope->k_cache && rope->v_cache && rope->guess_members >= 1 && rope->S_max >= 1 &&
                       M % rope->guess_members == 0,
                   "dd_decoder_linear: the commit ahead needs guess_pos, the layer's caches, S_max and guess_members "
                   "dividing M=%d", M);
      a.guess_winner = rope->guess_winner; a.guess_pos = rope->guess_pos; a.guess_members = rope->guess_members;
      a.S_max = rope->S_max; a.k_cache = (__nv_bfloat16*)rope->k_cache; a.v_cache = (__nv_bfloat16*)rope->v_cache;
    }
  }
  if (M == 0) return DD_OK;
  return norm_w ? launch_any<true>(a, epilogue, (cudaStream_t)stream)
                : launch_any<false>(a, epilogue, (cudaStream_t)stream);
}
