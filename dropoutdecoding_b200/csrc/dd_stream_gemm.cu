// Weight-streaming GEMM on the tcgen05 tensor cores for the batched single-token forwards of the ensemble:
//   y[m,n] = sum_k x[m,k] w[n,k],   16 < M = (#sequences x #members) <= 256 rows.
// (dd_skinny_gemm covers M <= 16..32 with mma.sync; above that every CTA of that kernel re-reads more
// activation bytes from L2 than weight bytes from HBM.)
//
// The GEMM is still a pure weight stream (intensity = M FLOP/byte, below the B200 ridge), so it is organised
// around HBM, with the tensor core as the only unit that can multiply a 32 KB weight block by up to 256 token
// rows in the time the block takes to arrive:
//   * swap-AB: the WEIGHTS are the M=128 operand of tcgen05.mma (128 output features per accumulator), the token
//     rows are the N operand (Np = M rounded up to 16, <= 256).  D[feature, row] lives in TMEM.
//   * a tile = two accumulators = 256 output features (SwiGLU: 128 gate + the matching 128 up features, so the
//     activation is applied in the epilogue); per 64-wide k-block the CTA ingests 32 KB of W (HBM, L2 evict-first)
//     and Np x 128 B of x (L2-resident) through TMA into a 3..6-stage mbarrier ring.
//   * stream-K: the (tile, k-block) units are cut into 148 equal contiguous ranges, one persistent CTA per SM, so
//     every SM pulls the same number of weight bytes whatever N and K are (o_proj has 16 tiles, gate/up 86).
//     A tile that is split over several CTAs is reduced deterministically: every contributor writes its fp32
//     partial to its own workspace slot and bumps the tile's counter; at the END of its range each contributor
//     waits for the tile's other partials and reduces + finishes 1/n of the tile's token rows, always summing the
//     slots in the same order.  (All CTAs are co-resident -- grid <= #SMs, one CTA per SM -- and every partial is
//     written before its owner waits, so the wait cannot deadlock; it is bounded and traps.)
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-9 = epilogue (tcgen05.ld 32x32b:
//     one thread = one output feature, 16 token rows per load) with the HF rounding points of dd_skinny_gemm.
// Algorithmic bytes per launch: N*K*2 (x2 for SwiGLU) -- the weights, read exactly once.
#include <cstdlib>

#include "dd_common.cuh"
#include "dd_tcgen05.cuh"

namespace dd {

constexpr int SG_BF = 128;                     // output features per accumulator (UMMA M)
constexpr int SG_BK = 64;                      // one 128-byte swizzle row of bf16
constexpr int SG_WBYTES = SG_BF * SG_BK * 2;   // 16 KB
constexpr int SG_EPI_WARPS = 8;
constexpr int SG_EPI_THREADS = SG_EPI_WARPS * 32;
constexpr int SG_THREADS = 64 + SG_EPI_THREADS;
constexpr int SG_MAX_STAGES = 6;
constexpr int SG_MAX_TILES = 2048;
constexpr size_t SG_COUNTER_BYTES = (size_t)SG_MAX_TILES * 2 * sizeof(int);
constexpr int SG_SMEM_LIMIT = 227 * 1024;

struct SgParams {
  int M, Np, N, K, epi;
  int w_rows;          // rows of the weight matrix (2N for SwiGLU)
  int T, KB, G;        // tiles, k-blocks per tile, CTAs
  int stages, acc_stages, tmem_cols;
  int dbg;             // DD_SG_DBG experiment bits: 1 no direct epilogue, 2 no partial flush, 4 no reduce, 8 no MMA
  const __nv_bfloat16* residual;
  void* y;
  float* ws;           // [G][2 slots][2 acc][Np][128] fp32 partials
  int* counters;       // [T][2]: partials arrived, slices finished
  unsigned long long* ts;  // DD_SG_DBG & 16: [G][8] timestamps (experiment)
};

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SG_EPI_THREADS) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SG_TS(slot) do { if (p.ts) p.ts[(size_t)cta * 8 + (slot)] = gtime(); } while (0)

// first unit of CTA c; the CTA that owns unit u
__device__ __forceinline__ long long sg_ubeg(int c, long long U, int G) { return (long long)c * U / G; }
__device__ __forceinline__ int sg_owner(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }

// 128-feature group of accumulator a of tile t (row block of the weight matrix)
__device__ __forceinline__ int sg_group(const SgParams& p, int tile, int a) {
  return p.epi == DD_EPI_SWIGLU ? tile + a * (p.N / SG_BF) : 2 * tile + a;
}

// one finished fp32 dot product -> the output, with the rounding points of the HF decoder (see dd_skinny_gemm)
__device__ __forceinline__ void sg_store(const SgParams& p, int m, int f, float v0, float v1) {
  const size_t o = (size_t)m * p.N + f;
  const float r0 = bf16_round(v0);
  if (p.epi == DD_EPI_NONE) {
    reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(r0);
  } else if (p.epi == DD_EPI_RESIDUAL) {
    reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(r0 + __bfloat162float(p.residual[o]));
  } else if (p.epi == DD_EPI_F32) {
    reinterpret_cast<float*>(p.y)[o] = r0;
  } else {  // SwiGLU: bf16( bf16(silu(g)) * u )
    const float act = bf16_round(r0 / (1.f + __expf(-r0)));
    reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(act * bf16_round(v1));
  }
}

__global__ void __launch_bounds__(SG_THREADS, 1)
stream_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   const SgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = 2 * SG_WBYTES + p.Np * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  // bars: full[SG_MAX_STAGES] | empty[SG_MAX_STAGES] | tmem_full[2] | tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SG_MAX_STAGES + 4);
  int* s_slot = reinterpret_cast<int*>(tmem_slot + 4);  // [<= 256] workspace slots of a split tile's contributors

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const long long U = (long long)p.T * p.KB;
  const long long ubeg = sg_ubeg(cta, U, p.G), uend = sg_ubeg(cta + 1, U, p.G);

  if (threadIdx.x == 0) SG_TS(0);
  const uint32_t ring = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * SG_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * SG_MAX_STAGES, bar_tempty = bar_tfull + 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SG_MAX_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, SG_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SG_TS(1);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    const uint64_t pol = l2_policy_evict_first();
    int stage = 0;
    uint32_t phase = 0;
    for (long long u = ubeg; u < uend;) {
      const int tile = (int)(u / p.KB), kb0 = (int)(u % p.KB);
      const int kb1 = (int)min((long long)p.KB, kb0 + (uend - u));
      const int g0 = sg_group(p, tile, 0);
      int g1 = sg_group(p, tile, 1);
      if (g1 * SG_BF >= p.w_rows) g1 = g0;  // odd number of groups: the second accumulator is padding
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t dst = ring + stage * stage_bytes, bar = bar_full + 8 * stage;
          mbar_expect_tx(bar, (uint32_t)stage_bytes);
          tma_load_2d_hint(dst, &tmap_w, bar, kb * SG_BK, g0 * SG_BF, pol);
          tma_load_2d_hint(dst + SG_WBYTES, &tmap_w, bar, kb * SG_BK, g1 * SG_BF, pol);
          tma_load_2d(dst + 2 * SG_WBYTES, &tmap_x, bar, kb * SG_BK, 0);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      u += kb1 - kb0;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16(SG_BF, p.Np);
    int stage = 0, acc_stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (long long u = ubeg; u < uend;) {
      const int kb0 = (int)(u % p.KB);
      const int kb1 = (int)min((long long)p.KB, kb0 + (uend - u));
      mbar_wait(bar_tempty + 8 * acc_stage, acc_phase ^ 1);  // the epilogue has drained this accumulator pair
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc_stage * 2 * p.Np;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = ring + stage * stage_bytes;
          const uint64_t b_desc = make_smem_desc(base + 2 * SG_WBYTES);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const uint64_t a_desc = make_smem_desc(base + a * SG_WBYTES);
#pragma unroll
            for (int k = 0; k < SG_BK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
              if (!(p.dbg & 8)) umma_bf16(d0 + a * p.Np, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (kb == kb1 - 1) umma_commit(bar_tfull + 8 * acc_stage);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
      u += kb1 - kb0;
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;          // which half of the 16-row chunks
    const int fl = q * 32 + lane;           // feature inside the 128-feature group
    const int et = threadIdx.x - 64;        // 0..255
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks = p.Np / 16;
    const bool swiglu = p.epi == DD_EPI_SWIGLU;
    int ptile[2], np = 0;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (long long u = ubeg; u < uend;) {
      const int tile = (int)(u / p.KB), kb0 = (int)(u % p.KB);
      const int kb1 = (int)min((long long)p.KB, kb0 + (uend - u));
      const bool complete = kb0 == 0 && kb1 == p.KB;
      const int f0 = sg_group(p, tile, 0) * SG_BF + fl, f1 = sg_group(p, tile, 1) * SG_BF + fl;
      const int fout0 = swiglu ? tile * SG_BF + fl : f0;  // SwiGLU writes feature tile*128+fl of [M, N]
      mbar_wait(bar_tfull + 8 * acc_stage, acc_phase);
      tc_fence_after();
      if (et == 0) SG_TS(2);
      const uint32_t t0 = lane_addr + acc_stage * 2 * p.Np;
      if (complete) {
        for (int c = h; c < n_chunks && !(p.dbg & 1); c += 2) {
          float v0[16], v1[16];
          tmem_ld16(t0 + c * 16, v0);
          tmem_ld16(t0 + p.Np + c * 16, v1);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int m = c * 16 + i;
            if (m < p.M) {
              if (swiglu) {
                if (fout0 < p.N) sg_store(p, m, fout0, v0[i], v1[i]);
              } else {
                if (f0 < p.N) sg_store(p, m, f0, v0[i], 0.f);
                if (f1 < p.N) sg_store(p, m, f1, v1[i], 0.f);
              }
            }
          }
        }
      } else {
        // partial: slot 0 when this is the tile the CTA's range starts in, slot 1 otherwise
        const int which = (tile == (int)(ubeg / p.KB)) ? 0 : 1;
        // slot layout [accumulator][row][feature]: a warp store = 32 consecutive features of one row (128 B)
        float* slot = p.ws + (size_t)(cta * 2 + which) * 2 * p.Np * SG_BF;
        for (int c = h; c < n_chunks && !(p.dbg & 2); c += 2) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            float v[16];
            tmem_ld16(t0 + a * p.Np + c * 16, v);
            float* dst = slot + ((size_t)a * p.Np + c * 16) * SG_BF + fl;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c * 16 + i < p.M) __stcg(dst + i * SG_BF, v[i]);
          }
        }
        ptile[np++] = tile;
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc_stage);
      if (!complete) {
        __threadfence();
        epi_bar_sync();
        if (et == 0) atomicAdd(&p.counters[tile * 2], 1);
      }
      if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
      u += kb1 - kb0;
    }

    if (et == 0) SG_TS(3);
    // ---- split tiles: wait for the other contributors, then reduce and finish this CTA's share of the rows ----
    for (int i = 0; i < np; ++i) {
      const int tile = ptile[i];
      const int c_first = sg_owner((long long)tile * p.KB, U, p.G);
      const int c_last = sg_owner((long long)(tile + 1) * p.KB - 1, U, p.G);
      const int nseg = c_last - c_first + 1, j = cta - c_first;
      epi_bar_sync();  // the previous tile's readers are done with s_slot
      if (et < nseg) {  // workspace slot of every contributor: 2*cta + (0: the tile its range starts in, 1: otherwise)
        const int cs = c_first + et;
        s_slot[et] = cs * 2 + ((tile == (int)(sg_ubeg(cs, U, p.G) / p.KB)) ? 0 : 1);
      }
      if (et == 0) {
        const long long t_start = clock64();
        while (ld_acquire_gpu(&p.counters[tile * 2]) < nseg) {
          __nanosleep(64);
          if (clock64() - t_start > (1ll << 33)) __trap();  // a protocol bug must not hang the GPU
        }
      }
      epi_bar_sync();
      __threadfence();
      if (et == 0) SG_TS(4 + i);
      // this CTA's share: rows [m_beg, m_end) of the tile.  256 threads = 32 feature quads x 8 rows; a thread owns
      // (quad, row) and (quad, row + 8) at once and sums the nseg slots in slot order, the loads of up to 8 slots
      // (32 x 16 B) issued before the first add -- they are independent, only the adds are ordered.
      const int m_beg = (int)((long long)j * p.M / nseg), m_end = (int)((long long)(j + 1) * p.M / nseg);
      const int fq = (et & 31) * 4, rsub = et >> 5;
      const int gbase0 = sg_group(p, tile, 0) * SG_BF, gbase1 = sg_group(p, tile, 1) * SG_BF;
      const size_t slot_stride = (size_t)2 * p.Np * SG_BF, acc_stride = (size_t)p.Np * SG_BF;
      for (int m = m_beg + rsub; m < m_end && !(p.dbg & 4); m += 16) {
        const bool two = m + 8 < m_end;
        float4 s[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) s[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sb = 0; sb < nseg; sb += 8) {
          float4 t[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (sb + i < nseg) {
              const float* src = p.ws + (size_t)s_slot[sb + i] * slot_stride + (size_t)m * SG_BF + fq;
              t[i][0] = __ldcg(reinterpret_cast<const float4*>(src));
              t[i][1] = __ldcg(reinterpret_cast<const float4*>(src + acc_stride));
              if (two) {
                t[i][2] = __ldcg(reinterpret_cast<const float4*>(src + 8 * SG_BF));
                t[i][3] = __ldcg(reinterpret_cast<const float4*>(src + acc_stride + 8 * SG_BF));
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (sb + i < nseg) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (e < 2 || two) {
                  s[e].x += t[i][e].x; s[e].y += t[i][e].y; s[e].z += t[i][e].z; s[e].w += t[i][e].w;
                }
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (r == 1 && !two) break;
          const int mm = m + 8 * r;
          const float a0[4] = {s[2 * r].x, s[2 * r].y, s[2 * r].z, s[2 * r].w};
          const float a1[4] = {s[2 * r + 1].x, s[2 * r + 1].y, s[2 * r + 1].z, s[2 * r + 1].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (swiglu) {
              if (tile * SG_BF + fq + e < p.N) sg_store(p, mm, tile * SG_BF + fq + e, a0[e], a1[e]);
            } else {
              if (gbase0 + fq + e < p.N) sg_store(p, mm, gbase0 + fq + e, a0[e], 0.f);
              if (gbase1 + fq + e < p.N) sg_store(p, mm, gbase1 + fq + e, a1[e], 0.f);
            }
          }
        }
      }
      epi_bar_sync();
      if (et == 0) {
        const int old = atomicAdd(&p.counters[tile * 2 + 1], 1);
        if (old == nseg - 1) {  // every contributor has passed its wait: re-arm the tile for the next launch
          p.counters[tile * 2] = 0;
          p.counters[tile * 2 + 1] = 0;
        }
      }
    }
  }

  if (threadIdx.x == 64) SG_TS(6);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SG_TS(7);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int sg_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct SgPlan {
  int Np, T, KB, G, stages, acc_stages, tmem_cols;
  size_t smem, ws_bytes;
};
static SgPlan sg_plan(int M, int N, int K, int epi) {
  SgPlan s;
  s.Np = (M + 15) / 16 * 16;
  const int groups = ceil_div(N, SG_BF);
  s.T = epi == DD_EPI_SWIGLU ? groups : ceil_div(groups, 2);
  s.KB = K / SG_BK;
  const long long U = (long long)s.T * s.KB;
  const int sms = sg_num_sms();
  s.G = (int)(U < sms ? U : sms);
  // Few tiles: cut every tile into the same number of K slices (G = T * slices), so that a CTA holds exactly ONE
  // partial and the tail of the kernel is one flush + one reduction, as long as that idles < 8 % of the SMs.
  // Otherwise (e.g. 86 gate/up tiles) plain stream-K over all SMs: some CTAs then straddle two tiles.
  if (s.T <= sms && s.KB >= sms / s.T) {
    const int slices = sms / s.T;
    if ((sms - s.T * slices) * 100 <= 8 * sms && !getenv("DD_SG_NOALIGN")) s.G = s.T * slices;
  }
  const int stage_bytes = 2 * SG_WBYTES + s.Np * 128;
  const int fixed = 1024 + (2 * SG_MAX_STAGES + 4) * 8 + 16 + 256 * 4;
  int st = (SG_SMEM_LIMIT - fixed) / stage_bytes;
  s.stages = st > SG_MAX_STAGES ? SG_MAX_STAGES : st;
  s.smem = (size_t)fixed + (size_t)s.stages * stage_bytes;
  s.acc_stages = 4 * s.Np <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < s.acc_stages * 2 * s.Np) cols *= 2;
  s.tmem_cols = cols;
  s.ws_bytes = SG_COUNTER_BYTES + (size_t)s.G * 2 * 2 * s.Np * SG_BF * sizeof(float) + (size_t)s.G * 8 * 8;
  return s;
}

static int sg_check(const char* fn, int M, int N, int K, int epi) {
  DD_CHECK_ARG(M >= 1 && M <= 256, "%s: M=%d out of [1,256]", fn, M);
  DD_CHECK_ARG(K >= SG_BK && K % SG_BK == 0, "%s: K=%d must be a positive multiple of %d", fn, K, SG_BK);
  DD_CHECK_ARG(N >= 8 && N % 8 == 0, "%s: N=%d must be a positive multiple of 8", fn, N);
  DD_CHECK_ARG(epi >= DD_EPI_NONE && epi <= DD_EPI_SWIGLU, "%s: bad epilogue %d", fn, epi);
  DD_CHECK_ARG(epi != DD_EPI_SWIGLU || N % SG_BF == 0, "%s: SwiGLU needs N %% %d == 0 (N=%d)", fn, SG_BF, N);
  DD_CHECK_ARG(ceil_div(N, SG_BF) <= SG_MAX_TILES, "%s: N=%d too large", fn, N);
  return DD_OK;
}

}  // namespace dd

extern "C" size_t dd_stream_gemm_workspace_bytes(int M, int N, int K, int epilogue) {
  using namespace dd;
  if (sg_check("dd_stream_gemm_workspace_bytes", M, N, K, epilogue)) return 0;
  return sg_plan(M, N, K, epilogue).ws_bytes;
}

extern "C" int dd_stream_gemm(const void* x, const void* w, const void* residual, int M, int N, int K, int epilogue,
                              void* y, void* workspace, size_t workspace_bytes, dd_stream_t stream) {
  using namespace dd;
  if (M == 0) return DD_OK;
  if (int rc = sg_check("dd_stream_gemm", M, N, K, epilogue)) return rc;
  DD_CHECK_ARG(x && w && y && workspace, "dd_stream_gemm: null pointer");
  DD_CHECK_ARG(epilogue != DD_EPI_RESIDUAL || residual, "dd_stream_gemm: residual epilogue needs residual");
  DD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
               "dd_stream_gemm: operands must be 16-byte aligned");
  const SgPlan s = sg_plan(M, N, K, epilogue);
  if (workspace_bytes < s.ws_bytes) {
    set_error("dd_stream_gemm: workspace %zu < %zu bytes", workspace_bytes, s.ws_bytes);
    return DD_ERR_WORKSPACE;
  }
  DD_CHECK_ARG(s.stages >= 2, "dd_stream_gemm: M=%d leaves no room for a 2-stage ring", M);
  const int w_rows = epilogue == DD_EPI_SWIGLU ? 2 * N : N;
  CUtensorMap tw, tx;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {SG_BK, SG_BF};
    if (int rc = make_tmap_bf16_sw128(&tw, w, 2, dims, strides, box, "dd_stream_gemm(w)")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {SG_BK, (uint32_t)s.Np};
    if (int rc = make_tmap_bf16_sw128(&tx, x, 2, dims, strides, box, "dd_stream_gemm(x)")) return rc;
  }
  SgParams p{};
  p.M = M; p.Np = s.Np; p.N = N; p.K = K; p.epi = epilogue; p.w_rows = w_rows;
  p.T = s.T; p.KB = s.KB; p.G = s.G; p.stages = s.stages; p.acc_stages = s.acc_stages; p.tmem_cols = s.tmem_cols;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = y;
  {
    static const int dbg = [] { const char* e = getenv("DD_SG_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    p.ts = (dbg & 16) ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + s.ws_bytes - (size_t)s.G * 8 * 8) : nullptr;
  }
  p.counters = static_cast<int*>(workspace);
  p.ws = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + SG_COUNTER_BYTES);
  static bool attr_set = false;
  if (!attr_set) {
    DD_CHECK_CUDA(cudaFuncSetAttribute(stream_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_LIMIT));
    attr_set = true;
  }
  stream_gemm_kernel<<<s.G, SG_THREADS, s.smem, (cudaStream_t)stream>>>(tw, tx, p);
  DD_CHECK_LAUNCH();
  return DD_OK;
}
