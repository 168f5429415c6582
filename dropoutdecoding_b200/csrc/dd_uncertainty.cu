// (1) Visual-token uncertainty stage on the 5th-gen tensor cores.
// Replaces lm_head over the visual rows + get_image_logits + get_topk_token_id +
// calculate_vision_uncertainty (models/llava.py:412-441, 710-756), see include/dropdec.h.
//
// logits[i,v] = <h_vis[i,:], w_lm[v,:]> is a [n_vis x vocab] GEMM (151 GFLOP per LLaVA-1.5 image)
// whose result the reference materialises in fp32 (74 MB + >=6 same-size temporaries).  Here the
// logits only ever exist as TMEM accumulators:
//
//   work item = (image, vocab strip of BN=96 columns); the CTA holds the strip for ALL visual rows
//   of the image at once -- NUM_MT row tiles of 128 -> NUM_MT*96 <= 480 of the 512 TMEM columns.
//   That is what makes the epistemic term single-sweep: it needs p_avg[v] = mean_i p[i,v], a
//   column reduction over every row of the image, before sum_v p[i,v] log(p_avg[v]) can be formed.
//
//   warps 8-10: TMA producers.  Per k-block (64 bf16 = one 128B-swizzle row) the CTA ingests the W strip tile
//              [96 x 64] (HBM stream, 4-deep ring) and NUM_MT activation tiles [128 x 64] (L2-resident, 2 k-blocks
//              deep): 92 KB per k-block, 82 B/clk at the rate the tensor core consumes them.  Round 1 fed four of the
//              five activation tiles through cp.async because "an SM ingests 35 B/clk through TMA"; that ceiling was
//              the issue loop of ONE producer thread (a barrier wait + a copy per box: ~350-670 cycles each,
//              tools/tma_probe2.cu) -- three producer warps with two boxes each per k-block keep the ring full
//              (measured: 97-126 B/clk per SM from four issuing warps).  Every ring slot belongs to one warp.
//   warp 11  : MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128 N=96 K=16,
//              both operands K-major SWIZZLE_128B; slots are released with tcgen05.commit.
//   warps 0-7: epilogue, two warps per TMEM lane quarter: warp w reads lanes 32*(w&3).. and the 48 columns
//              [48*(w>>2), +48) of the strip.  tcgen05.ld 32x32b: one thread = one visual row, so softmax
//              statistics are thread-local; the column sums are a butterfly over the warp plus 4 smem rows.  The two
//              halves never talk to each other: every per-row partial simply exists once per (CTA, half).
//
// Two sweeps of the same pipeline (the row normaliser Z_i must be complete before p exists):
//   STATS : per row running max / sum-exp, and the row's maximum inside every half strip  (sweep 1)
//   FUSED : p = exp(l - m)/Z;  A_i = sum p log(p+eps), Q_i = sum (p-1/V)^2, column sums ->
//           log(p_avg+eps) -> C_i = sum p log(p_avg+eps)                        (sweep 2)
//   finalize: alea = -A, epi = A - C, var = Q/(V-1).
// Top-k ids without per-thread sorted lists in the hot loop: tau_i = the k-th largest of row i's per-half-strip
// maxima is a lower bound of its k-th largest logit, so sweep 2 (which recomputes bit-identical logits)
// only has to append the few elements >= tau_i to a per-row candidate list (one predictable compare per
// element); the finalize kernel sorts the candidates (value descending, id ascending).  Without ties at
// most 48*k elements can pass the threshold; the list holds 96*k + 64 and an overflow is reported.
// Algorithmic FLOPs = 2*n_vis*d*V per image, counted once although two sweeps run.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "dd_common.cuh"
#include "dd_tcgen05.cuh"

#ifdef DD_UNC_EXPERIMENT
long long* g_unc_dbg = nullptr;
#endif

namespace dd {

constexpr int BM = 128, BN = 96, BK = 64, UMMA_K = 16;
constexpr int BNH = BN / 2;              // columns per epilogue half
constexpr int kSlotBytes = BM * BK * 2;  // 16 KB: one [128 x 64] bf16 activation tile
constexpr int kWBytes = BN * BK * 2;     // 12 KB: one [96 x 64] bf16 W tile
constexpr int kMaxMT = 5;                // row tiles of 128 per image (n_vis <= 640)
constexpr int kActStages = 2;            // activation ring: k-blocks in flight
constexpr int kWStages = 4;              // W ring
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kHalfThreads = kEpiThreads / 2;
constexpr int kProdWarp0 = kEpiWarps, kProdWarps = 3;
constexpr int kMmaWarp = kProdWarp0 + kProdWarps;
constexpr int kUncThreads = (kMmaWarp + 1) * 32;
// barriers: act_full[kActStages] | act_empty[kActStages][kMaxMT] | w_full[kWStages] | w_empty[kWStages] | tmem_full |
//           tmem_empty[kMaxMT]
constexpr int kNumBars = kActStages + kActStages * kMaxMT + 2 * kWStages + 1 + kMaxMT;
constexpr int kTmemCols = 512;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kEps = 1e-10f;  // models/llava.py:735,739

constexpr uint32_t kIdesc = make_idesc_bf16(BM, BN);

// The epilogue warps are alone on their scheduler (two warps per sub-partition), so the softmax sweeps are bound by
// instructions issued per element.  exp2f / __logf carry a range fix-up for denormal results / arguments (4 extra
// instructions each); here 2^x below 2^-126 may flush to zero (a probability < 1e-38) and the log argument is
// >= 1e-10, so the bare MUFU forms are exact enough and identical in both sweeps.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ln_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}
__device__ __forceinline__ void half_bar_sync(int half) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(kHalfThreads) : "memory");
}

// ------------------------------------------------------------------------------------------
struct UncParams {
  int n_images, n_vis, d, vocab, topk, round_bf16;
  int n_strips, parts;     // strips per image; CTAs per image.  Per-row partials exist per (CTA, half): 2*parts;
                           // strip maxima per half strip: 2*n_strips
  float* part_stats;       // [img][2*parts][n_vis][2]    (m, z)
  float* strip_max;        // [img][2*n_strips][n_vis]    STATS output: row maximum inside each half strip
  const float* tau;        // [img][n_vis]                candidate threshold (FUSED / COLSUM input)
  int32_t* cand_cnt;       // [img][n_vis]
  float* cand_val;         // [img][n_vis][cap]
  int32_t* cand_idx;       // [img][n_vis][cap]
  int cap;
  const float* row_stats;  // [img][n_vis][2]             (m, 1/Z)  FUSED input
  float* part_acq;         // [img][2*parts][n_vis][3]    (A, Q, C) FUSED output
  int n_chunks;            // row chunks of kMaxMT*128 rows per image (> 1 when n_vis > 640)
  float* colpart;          // [img][chunk][n_strips*BN]   column sums of p over one row chunk (COLSUM -> logpavg)
  float* logpavg;          // [img][n_strips*BN]          log(p_avg + eps)                    (CROSS input)
#ifdef DD_UNC_EXPERIMENT
  long long* dbg;          // [ctas][8] cycle counters (experiment builds)
#endif
  int32_t* overflow;       // [1] rows whose top-k candidate list overflowed (ties at the threshold): the ids of such a
                           //     row may miss a tied candidate, so the host treats a non-zero count as an error
};

// Rare path of sweep 2 (whole warp, warp-uniform call): some row of this warp has a top-k candidate inside the
// 16 columns at `taddr`.  Re-reads them from TMEM 8 at a time and appends every logit >= the row's threshold.
struct CandSink {
  int32_t* cnt;
  float* val;
  int32_t* idx;
  int cap, vocab, round_bf16;
};
__device__ __noinline__ void append_candidates(const CandSink p, uint32_t taddr, float tau, int col_base, size_t r) {
#pragma unroll 1
  for (int j = 0; j < 2; ++j) {
    float v[8];
    tmem_ld8(taddr + j * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = p.round_bf16 ? bf16_round(v[i]) : v[i];
      const int col = col_base + j * 8 + i;
      if (col < p.vocab && x >= tau) {
        const int slot = atomicAdd(&p.cnt[r], 1);
        if (slot < p.cap) {
          p.val[r * p.cap + slot] = x;
          p.idx[r * p.cap + slot] = col;
        }
      }
    }
  }
}

// n_vis <= 640: STATS, FUSED.  Larger images (LLaVA-NeXT anyres, up to ~2.9k rows) do not fit the 512
// TMEM columns, so the column mean needs its own sweep: STATS, COLSUM (p, A, Q, per-chunk column sums),
// a tiny log(p_avg) kernel, CROSS (C = sum_v p log(p_avg+eps)).
enum { MODE_STATS = 0, MODE_FUSED = 1, MODE_COLSUM = 2, MODE_CROSS = 3 };

template <int NUM_MT, int MODE, int KT>
__global__ void __launch_bounds__(kUncThreads, 1)
uncertainty_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const UncParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles (the dynamic smem base is only 16-byte aligned by contract)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_ring_p = smem + kActStages * kMaxMT * kSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring_p + kWStages * kWBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* s_col = reinterpret_cast<float*>(tmem_slot + 4);  // [4][BN]
  float* s_logp = s_col + 4 * BN;                            // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.y / p.n_chunks, chunk = blockIdx.y % p.n_chunks, part = blockIdx.x;
  const int row0 = chunk * (kMaxMT * BM);  // first row of this CTA's row chunk
  const int spc = (p.n_strips + p.parts - 1) / p.parts;
  const int strip_beg = part * spc;
  const int strip_end = min(p.n_strips, strip_beg + spc);
  const int num_kb = p.d / BK;

  const uint32_t act_ring = smem_u32(smem), w_ring = smem_u32(w_ring_p);
  const uint32_t bar_afull = smem_u32(bars), bar_aempty = bar_afull + 8 * kActStages;
  const uint32_t bar_wfull = bar_aempty + 8 * kActStages * kMaxMT, bar_wempty = bar_wfull + 8 * kWStages;
  const uint32_t bar_tfull = bar_wempty + 8 * kWStages, bar_tempty = bar_tfull + 8;
  // producer warp pw owns activation tiles 2*pw and 2*pw + 1 (those below NUM_MT); the last producer also owns W
  constexpr int kActProducers = (NUM_MT + 1) / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kActStages; ++i) {
      mbar_init(bar_afull + 8 * i, kActProducers);  // one arrive.expect_tx per producer warp that owns activation tiles
      for (int m = 0; m < kMaxMT; ++m) mbar_init(bar_aempty + 8 * (i * kMaxMT + m), 1);
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(bar_wfull + 8 * i, 1);
      mbar_init(bar_wempty + 8 * i, 1);
    }
    mbar_init(bar_tfull, 1);
    for (int i = 0; i < kMaxMT; ++i) mbar_init(bar_tempty + 8 * i, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {  // TMEM: 512 columns, whole warp, then give the permit back
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kProdWarp0 && warp < kMmaWarp) {
    // =========================== TMA producers (warp-converged, one elected lane issues) =========
    const int pw = warp - kProdWarp0;
    const int mt0 = 2 * pw, n_mine = max(0, min(NUM_MT - mt0, 2));
    const bool w_mine = pw == kProdWarps - 1;
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    uint32_t g = 0;
    for (int strip = strip_beg; strip < strip_end; ++strip) {
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const uint32_t as = g % kActStages, aphase = (g / kActStages) & 1;
        if (w_mine) {
          const uint32_t ws = g % kWStages, wphase = (g / kWStages) & 1;
          mbar_wait(bar_wempty + 8 * ws, wphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_wfull + 8 * ws, kWBytes);
            tma_load_2d(w_ring + ws * kWBytes, &tmap_w, bar_wfull + 8 * ws, kb * BK, strip * BN);
          }
          __syncwarp();
        }
        if (n_mine > 0) {
          for (int j = 0; j < n_mine; ++j) mbar_wait(bar_aempty + 8 * (as * kMaxMT + mt0 + j), aphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_afull + 8 * as, (uint32_t)(n_mine * kSlotBytes));
            for (int j = 0; j < n_mine; ++j)   // rows past n_vis are zero-filled by the tensor map
              tma_load_3d(act_ring + (as * kMaxMT + mt0 + j) * kSlotBytes, &tmap_a, bar_afull + 8 * as, kb * BK,
                          row0 + (mt0 + j) * BM, img);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (warp-converged, one elected lane issues) ==========
    // The issuing thread must stay ahead of a tensor pipe that retires one M128 x N96 x K16 MMA every ~57
    // cycles, so the body of a k-block is instantiated for each (activation stage, W stage) combination:
    // every shared-memory descriptor is then "ring base + compile-time constant", one elected region issues
    // all NUM_MT x 4 MMAs of the k-block, and the per-k-block cost is two mbarrier waits.
#ifdef DD_UNC_EXPERIMENT
    long long dbg_w = 0, dbg_a = 0, dbg_t = 0;
    const long long dbg_t0 = clock64();
#endif
    auto issue_group = [&](auto AS, auto WS, uint32_t aphase, uint32_t wphase, bool first_kb, bool last_kb,
                           uint32_t tempty_parity) {
      constexpr uint32_t as = decltype(AS)::value, ws = decltype(WS)::value;
#ifdef DD_UNC_EXPERIMENT
      const long long c0 = clock64();
#endif
      mbar_wait(bar_wfull + 8 * ws, wphase);
#ifdef DD_UNC_EXPERIMENT
      const long long c1 = clock64();
#endif
      mbar_wait(bar_afull + 8 * as, aphase);
#ifdef DD_UNC_EXPERIMENT
      const long long c2 = clock64();
#endif
      if (first_kb) {       // the epilogue must have drained the previous strip's accumulators
#pragma unroll
        for (int mt = 0; mt < NUM_MT; ++mt) mbar_wait(bar_tempty + 8 * mt, tempty_parity);
      }
#ifdef DD_UNC_EXPERIMENT
      dbg_w += c1 - c0; dbg_a += c2 - c1; dbg_t += clock64() - c2;
#endif
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b_desc = make_smem_desc(w_ring + ws * kWBytes);
        const uint32_t acc = first_kb ? 0u : 1u;
#pragma unroll
        for (int mt = 0; mt < NUM_MT; ++mt) {
          const uint64_t a_desc = make_smem_desc(act_ring + (as * kMaxMT + mt) * kSlotBytes);
          umma_bf16(tmem_base + mt * BN, a_desc, b_desc, kIdesc, acc);
#pragma unroll
          for (int k = 1; k < BK / UMMA_K; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
            umma_bf16(tmem_base + mt * BN, a_desc + 2 * k, b_desc + 2 * k, kIdesc, 1u);
          umma_commit(bar_aempty + 8 * (as * kMaxMT + mt));  // activation slot free once read
        }
        umma_commit(bar_wempty + 8 * ws);       // W slot free
        if (last_kb) umma_commit(bar_tfull);    // strip accumulators complete
      }
      __syncwarp();
    };
    using std::integral_constant;
    static_assert(kActStages == 2 && kWStages == 4, "the switch below enumerates g % 4");
    uint32_t it = 0, g = 0;
    for (int strip = strip_beg; strip < strip_end; ++strip, ++it) {
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const uint32_t aphase = (g / kActStages) & 1, wphase = (g / kWStages) & 1, tp = (it & 1) ^ 1;
        const bool f = kb == 0, l = kb == num_kb - 1;
        switch (g & 3) {
          case 0: issue_group(integral_constant<uint32_t, 0>{}, integral_constant<uint32_t, 0>{}, aphase, wphase, f, l, tp); break;
          case 1: issue_group(integral_constant<uint32_t, 1>{}, integral_constant<uint32_t, 1>{}, aphase, wphase, f, l, tp); break;
          case 2: issue_group(integral_constant<uint32_t, 0>{}, integral_constant<uint32_t, 2>{}, aphase, wphase, f, l, tp); break;
          default: issue_group(integral_constant<uint32_t, 1>{}, integral_constant<uint32_t, 3>{}, aphase, wphase, f, l, tp); break;
        }
      }
    }
#ifdef DD_UNC_EXPERIMENT
    if (p.dbg && lane == 0) {
      long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
      o[0] = clock64() - dbg_t0; o[1] = dbg_w; o[2] = dbg_a; o[3] = dbg_t;
    }
#endif
  } else {
    // =========================== epilogue (warps 0-7) ===========================
    const int q = warp & 3, half = warp >> 2;
    const int tid = q * 32 + lane;       // 0..127 = TMEM lane = row inside each 128-row tile
    const int hb = half * BNH;           // first column of this warp's half of the strip
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + hb;
    const float inv_v = 1.0f / (float)p.vocab;
    const int parts_e = 2 * p.parts, part_e = 2 * part + half;

    float st_m[NUM_MT], st_z[NUM_MT];          // STATS: running max, sum-exp | FUSED: m, 1/Z
    float acc_a[NUM_MT], acc_q[NUM_MT], acc_c[NUM_MT];
    float st_tau[NUM_MT];                      // FUSED / COLSUM: top-k candidate threshold of the row
#pragma unroll
    for (int mt = 0; mt < NUM_MT; ++mt) {
      const int row = row0 + mt * BM + tid;
      acc_a[mt] = acc_q[mt] = acc_c[mt] = 0.f;
      st_tau[mt] = INFINITY;
      if (MODE == MODE_STATS) {
        st_m[mt] = -INFINITY;
        st_z[mt] = 0.f;
      } else {
        const bool ok = row < p.n_vis;
        st_m[mt] = ok ? p.row_stats[((size_t)img * p.n_vis + row) * 2] : 0.f;
        st_z[mt] = ok ? p.row_stats[((size_t)img * p.n_vis + row) * 2 + 1] : 0.f;
        if (ok && MODE != MODE_CROSS) st_tau[mt] = p.tau[(size_t)img * p.n_vis + row];
      }
    }

#ifdef DD_UNC_EXPERIMENT
    long long dbg_ewait = 0;
    const long long dbg_e0 = clock64();
#endif
    uint32_t it = 0;
    for (int strip = strip_beg; strip < strip_end; ++strip, ++it) {
      const int col0 = strip * BN + hb;   // first vocabulary column of this warp's half strip
#ifdef DD_UNC_EXPERIMENT
      const long long e0 = clock64();
#endif
      mbar_wait(bar_tfull, it & 1);
#ifdef DD_UNC_EXPERIMENT
      const long long e1 = clock64();
      dbg_ewait += e1 - e0;
#endif
      tc_fence_after();

      if (MODE == MODE_STATS) {
#pragma unroll
        for (int mt = 0; mt < NUM_MT; ++mt) {
          const int row = row0 + mt * BM + tid;
          const bool ok = row < p.n_vis;
          float smax = -INFINITY;  // this row's maximum inside the half strip
#pragma unroll 1
          for (int c = 0; c < BNH / 16; ++c) {
            float v[16];
            tmem_ld16(lane_addr + mt * BN + c * 16, v);
            if (ok) {
              const int cbase = col0 + c * 16;
              float cmax = -INFINITY;
              if (cbase + 16 <= p.vocab) {  // (CTA-uniform) all but the last strip: no per-element bound check
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (p.round_bf16) v[i] = bf16_round(v[i]);
                  cmax = fmaxf(cmax, v[i]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (p.round_bf16) v[i] = bf16_round(v[i]);
                  if (cbase + i >= p.vocab) v[i] = -INFINITY;
                  cmax = fmaxf(cmax, v[i]);
                }
              }
              if (cmax > -INFINITY) {   // (a chunk wholly past the vocabulary contributes nothing)
                smax = fmaxf(smax, cmax);
                if (cmax > st_m[mt]) {
                  st_z[mt] *= exp2f((st_m[mt] - cmax) * kLog2e);
                  st_m[mt] = cmax;
                }
                const float mneg = -st_m[mt] * kLog2e;
                float zs0 = 0.f, zs1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  zs0 += ex2_ftz(fmaf(v[i], kLog2e, mneg));
                  zs1 += ex2_ftz(fmaf(v[i + 1], kLog2e, mneg));
                }
                st_z[mt] += zs0 + zs1;
              }
            }
          }
          if (ok) p.strip_max[((size_t)img * 2 * p.n_strips + 2 * strip + half) * p.n_vis + row] = smax;
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * mt);  // this row tile's accumulators may be overwritten
        }
      } else {
        if (MODE == MODE_FUSED || MODE == MODE_COLSUM) {
          // ---- pass 1: p, A, Q and the column sums (16 columns at a time over all row tiles, so that
          //      the running column sums stay in 16 registers) ------------------------------------
#pragma unroll 1
          for (int c = 0; c < BNH / 16; ++c) {
            float colsum[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) colsum[i] = 0.f;
#pragma unroll
            for (int mt = 0; mt < NUM_MT; ++mt) {
              const bool ok = row0 + mt * BM + tid < p.n_vis;
              const float mneg = -st_m[mt] * kLog2e, iz = st_z[mt];
              float v[16];
              tmem_ld16(lane_addr + mt * BN + c * 16, v);
              float cm = -INFINITY;  // chunk maximum of this row: any top-k candidate in here?
              if (ok) {
                auto body = [&](auto padded) {  // padded: the chunk holds vocab columns past the end (last strip only)
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    float x = v[i];
                    if (p.round_bf16) x = bf16_round(x);
                    if constexpr (decltype(padded)::value) {
                      if (col0 + c * 16 + i < p.vocab) cm = fmaxf(cm, x);
                    } else {
                      cm = fmaxf(cm, x);
                    }
                    float pr = ex2_ftz(fmaf(x, kLog2e, mneg)) * iz;
                    float dv = pr - inv_v;
                    if constexpr (decltype(padded)::value)
                      if (col0 + c * 16 + i >= p.vocab) pr = dv = 0.f;
                    colsum[i] += pr;
                    acc_a[mt] = fmaf(pr, ln_ftz(pr + kEps), acc_a[mt]);
                    acc_q[mt] = fmaf(dv, dv, acc_q[mt]);
                  }
                };
                if (col0 + c * 16 + 16 <= p.vocab) body(std::false_type{});
                else body(std::true_type{});
              }
              if (__any_sync(0xffffffffu, cm >= st_tau[mt]))  // warp-uniform, rare
                append_candidates(CandSink{p.cand_cnt, p.cand_val, p.cand_idx, p.cap, p.vocab, p.round_bf16},
                                  lane_addr + mt * BN + c * 16, st_tau[mt], col0 + c * 16,
                                  (size_t)img * p.n_vis + row0 + mt * BM + tid);
              if (MODE == MODE_COLSUM && c == BNH / 16 - 1) {  // no second pass over TMEM in this mode
                tc_fence_before();
                mbar_arrive(bar_tempty + 8 * mt);
              }
            }
            // column sums over the rows held by this CTA: warp butterfly, then 4 smem rows
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float s = colsum[i];
              s += __shfl_xor_sync(0xffffffffu, s, 16);
              s += __shfl_xor_sync(0xffffffffu, s, 8);
              s += __shfl_xor_sync(0xffffffffu, s, 4);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              if (lane == i) s_col[q * BN + hb + c * 16 + i] = s;
            }
          }
          half_bar_sync(half);
          if (tid < BNH) {
            const int cc = hb + tid;
            const float tot = s_col[cc] + s_col[BN + cc] + s_col[2 * BN + cc] + s_col[3 * BN + cc];
            if (MODE == MODE_FUSED)
              s_logp[cc] = __logf(tot / (float)p.n_vis + kEps);  // log(p_avg + eps), models/llava.py:732-735
            else
              p.colpart[((size_t)img * p.n_chunks + chunk) * ((size_t)p.n_strips * BN) + col0 + tid] = tot;
          }
          half_bar_sync(half);
        }
        if (MODE == MODE_CROSS) {
          if (tid < BNH) s_logp[hb + tid] = p.logpavg[(size_t)img * ((size_t)p.n_strips * BN) + col0 + tid];
          half_bar_sync(half);
        }
        if (MODE == MODE_FUSED || MODE == MODE_CROSS) {
          // ---- pass 2: C_i = sum_v p log(p_avg + eps) ---------------------------------------
#pragma unroll
          for (int mt = 0; mt < NUM_MT; ++mt) {
            const bool ok = row0 + mt * BM + tid < p.n_vis;
            const float mneg = -st_m[mt] * kLog2e, iz = st_z[mt];
#pragma unroll 1
            for (int c = 0; c < BNH / 16; ++c) {
              float v[16];
              tmem_ld16(lane_addr + mt * BN + c * 16, v);
              if (ok) {
                auto body = [&](auto padded) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    float x = v[i];
                    if (p.round_bf16) x = bf16_round(x);
                    float pr = ex2_ftz(fmaf(x, kLog2e, mneg)) * iz;
                    if constexpr (decltype(padded)::value)
                      if (col0 + c * 16 + i >= p.vocab) pr = 0.f;
                    acc_c[mt] = fmaf(pr, s_logp[hb + c * 16 + i], acc_c[mt]);
                  }
                };
                if (col0 + c * 16 + 16 <= p.vocab) body(std::false_type{});
                else body(std::true_type{});
              }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * mt);  // this row tile's accumulators may be overwritten
          }
          half_bar_sync(half);  // s_col / s_logp reusable
        }
      }
    }

#ifdef DD_UNC_EXPERIMENT
    if (p.dbg && warp == 0 && lane == 0) {
      long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
      o[4] = clock64() - dbg_e0; o[5] = dbg_ewait;
    }
#endif
    // ---- per-(CTA, half) partials ------------------------------------------------------------
#pragma unroll
    for (int mt = 0; mt < NUM_MT; ++mt) {
      const int row = row0 + mt * BM + tid;
      if (row >= p.n_vis) continue;
      const size_t pr = ((size_t)img * parts_e + part_e) * p.n_vis + row;
      if (MODE == MODE_STATS) {
        p.part_stats[pr * 2] = st_m[mt];
        p.part_stats[pr * 2 + 1] = st_z[mt];
      } else {
        if (MODE != MODE_CROSS) {
          p.part_acq[pr * 3] = acc_a[mt];
          p.part_acq[pr * 3 + 1] = acc_q[mt];
        }
        if (MODE != MODE_COLSUM) p.part_acq[pr * 3 + 2] = acc_c[mt];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// The two finalize kernels run ONE WARP PER ROW (lanes split the partials): a thread per row would walk
// hundreds of dependent L2 loads and take longer than a GEMM sweep.
constexpr int kFinWarps = 8;

// merge the per-CTA (m, z) partials of one row: row_stats = (m, 1/Z); tau = k-th largest per-strip maximum
// (a lower bound of the row's k-th largest logit; -inf when there are fewer than k strips); reset the
// candidate counter for sweep 2
__global__ void __launch_bounds__(kFinWarps * 32)
stats_finalize_kernel(const UncParams p, float* __restrict__ row_stats, float* __restrict__ tau) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kFinWarps + (threadIdx.x >> 5), img = blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.overflow = 0;
  if (row >= p.n_vis) return;  // warp-uniform
  const int parts_e = 2 * p.parts, strips_e = 2 * p.n_strips;   // per-row partials / strip maxima exist per half strip
  float m = -INFINITY;
  for (int q = lane; q < parts_e; q += 32)
    m = fmaxf(m, p.part_stats[(((size_t)img * parts_e + q) * p.n_vis + row) * 2]);
  m = warp_max(m);
  float z = 0.f;
  for (int q = lane; q < parts_e; q += 32) {
    const size_t pr = ((size_t)img * parts_e + q) * p.n_vis + row;
    const float pm = p.part_stats[pr * 2];
    if (pm > -INFINITY) z += p.part_stats[pr * 2 + 1] * exp2f((pm - m) * kLog2e);
  }
  z = warp_sum(z);
  // k-th largest strip maximum: every lane keeps its own values sorted (descending), then k rounds of
  // "largest head over the warp, its owner pops"
  float tv[DD_MAX_TOPK];
#pragma unroll
  for (int j = 0; j < DD_MAX_TOPK; ++j) tv[j] = -INFINITY;
  for (int s = lane; s < strips_e; s += 32) {
    float v = p.strip_max[((size_t)img * strips_e + s) * p.n_vis + row];
#pragma unroll
    for (int j = 0; j < DD_MAX_TOPK; ++j) {  // insertion by a chain of swaps keeps tv[] in registers
      const float hi = fmaxf(tv[j], v);
      v = fminf(tv[j], v);
      tv[j] = hi;
    }
  }
  float kth = -INFINITY;
  for (int j = 0; j < p.topk; ++j) {
    kth = warp_max(tv[0]);
    const unsigned owners = __ballot_sync(0xffffffffu, tv[0] == kth);
    if (lane == __ffs(owners) - 1) {  // one owner pops its head
#pragma unroll
      for (int k = 0; k + 1 < DD_MAX_TOPK; ++k) tv[k] = tv[k + 1];
      tv[DD_MAX_TOPK - 1] = -INFINITY;
    }
  }
  if (lane == 0) {
    const size_t o = (size_t)img * p.n_vis + row;
    row_stats[o * 2] = m;
    row_stats[o * 2 + 1] = 1.0f / z;
    tau[o] = kth;
    p.cand_cnt[o] = 0;
  }
}

// log(p_avg + eps) from the per-chunk column sums (n_vis > 640 path)
__global__ void logpavg_kernel(const UncParams p) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
  const int vp = p.n_strips * BN;
  if (v >= vp) return;
  float tot = 0.f;
  for (int c = 0; c < p.n_chunks; ++c) tot += p.colpart[((size_t)img * p.n_chunks + c) * vp + v];
  p.logpavg[(size_t)img * vp + v] = __logf(tot / (float)p.n_vis + kEps);
}

__global__ void __launch_bounds__(kFinWarps * 32)
fused_finalize_kernel(const UncParams p, float* __restrict__ epi, float* __restrict__ alea, float* __restrict__ var,
                      int32_t* __restrict__ topk_ids) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kFinWarps + (threadIdx.x >> 5), img = blockIdx.y;
  if (row >= p.n_vis) return;  // warp-uniform
  const size_t r = (size_t)img * p.n_vis + row;
  {  // top-k ids from the candidate list: k selection rounds, value descending, id ascending on ties
    const int n = min(p.cand_cnt[r], p.cap);
    if (lane == 0 && p.cand_cnt[r] > p.cap) atomicAdd(p.overflow, 1);
    const float* cv = p.cand_val + r * p.cap;
    const int32_t* ci = p.cand_idx + r * p.cap;
    float pv = INFINITY;
    int pi = -1;
    for (int j = 0; j < p.topk; ++j) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int e = lane; e < n; e += 32) {
        const float v = cv[e];
        const int id = ci[e];
        const bool after_prev = v < pv || (v == pv && id > pi);   // strictly after the previous pick
        if (after_prev && (v > bv || (v == bv && id < bi))) { bv = v; bi = id; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) topk_ids[r * p.topk + j] = bi == 0x7fffffff ? 0 : bi;
      pv = bv; pi = bi;
    }
  }
  float a = 0.f, q = 0.f, c = 0.f;
  for (int s = lane; s < 2 * p.parts; s += 32) {
    const size_t pr = ((size_t)img * 2 * p.parts + s) * p.n_vis + row;
    a += p.part_acq[pr * 3];
    q += p.part_acq[pr * 3 + 1];
    c += p.part_acq[pr * 3 + 2];
  }
  a = warp_sum(a); q = warp_sum(q); c = warp_sum(c);
  if (lane == 0) {
    alea[r] = -a;                         // models/llava.py:739
    epi[r] = a - c;                       // models/llava.py:735
    var[r] = q / (float)(p.vocab - 1);    // torch.var (unbiased), models/llava.py:728
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct UncLayout {
  int n_strips, parts, n_chunks;
  int cap;
  size_t off_stats, off_smax, off_tau, off_ccnt, off_cval, off_cidx, off_rowstats, off_acq, off_colpart, off_logpavg, off_ovf, total;
};
static UncLayout unc_layout(int n_images, int n_vis, int vocab, int topk) {
  UncLayout L;
  L.n_strips = ceil_div(vocab, BN);
  L.n_chunks = ceil_div(n_vis, kMaxMT * BM);
  // CTAs per (image, row chunk): one CTA owns the whole TMEM, so a grid larger than the SM count runs in waves and
  // 152 CTAs cost twice what 148 do.  Pick the split that minimises waves x strips-per-CTA (ties: fewer partials).
  const int groups = (n_images > 0 ? n_images : 1) * L.n_chunks, sms = num_sms();
  int best = 1;
  long best_cost = -1;
  for (int parts = 1; parts <= L.n_strips && parts <= sms; ++parts) {
    const long cost = (long)ceil_div(parts * groups, sms) * ceil_div(L.n_strips, parts);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = parts; }
  }
  L.parts = best;
  const size_t rows = (size_t)n_images * 2 * L.parts * n_vis;   // partials per (CTA, epilogue half)
  size_t o = 0;
  L.off_stats = o; o = align_up(o + rows * 2 * sizeof(float), 256);
  const size_t nrow = (size_t)n_images * n_vis;
  L.cap = BN * topk + 64;  // without ties at most BN*topk elements reach the threshold
  L.off_smax = o; o = align_up(o + nrow * 2 * L.n_strips * sizeof(float), 256);
  L.off_tau = o; o = align_up(o + nrow * sizeof(float), 256);
  L.off_ccnt = o; o = align_up(o + nrow * sizeof(int32_t), 256);
  L.off_cval = o; o = align_up(o + nrow * L.cap * sizeof(float), 256);
  L.off_cidx = o; o = align_up(o + nrow * L.cap * sizeof(int32_t), 256);
  L.off_rowstats = o; o = align_up(o + (size_t)n_images * n_vis * 2 * sizeof(float), 256);
  L.off_acq = o; o = align_up(o + rows * 3 * sizeof(float), 256);
  const size_t vp = (size_t)L.n_strips * BN;
  L.off_colpart = o; o = align_up(o + (L.n_chunks > 1 ? (size_t)n_images * L.n_chunks * vp * sizeof(float) : 0), 256);
  L.off_logpavg = o; o = align_up(o + (L.n_chunks > 1 ? (size_t)n_images * vp * sizeof(float) : 0), 256);
  L.off_ovf = o; o = align_up(o + sizeof(int32_t), 256);
  L.total = o;
  return L;
}

constexpr size_t kUncSmem = 1024 + (size_t)kActStages * kMaxMT * kSlotBytes + (size_t)kWStages * kWBytes + kNumBars * 8 + 16 +
                            5 * BN * sizeof(float);

template <int NUM_MT, int MODE, int KT>
static int launch_unc(const CUtensorMap& ta, const CUtensorMap& tw, const UncParams& p, cudaStream_t st) {
  auto kern = uncertainty_kernel<NUM_MT, MODE, KT>;
  DD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUncSmem));
  kern<<<dim3(p.parts, p.n_images * p.n_chunks), kUncThreads, kUncSmem, st>>>(ta, tw, p);
  DD_CHECK_LAUNCH();
  return DD_OK;
}

template <int NUM_MT>
static int launch_unc_mt(const CUtensorMap& ta, const CUtensorMap& tw, const UncParams& p, int mode,
                         cudaStream_t st) {
  if (mode == MODE_FUSED) return launch_unc<NUM_MT, MODE_FUSED, 1>(ta, tw, p, st);
  if (mode == MODE_COLSUM) return launch_unc<NUM_MT, MODE_COLSUM, 1>(ta, tw, p, st);
  if (mode == MODE_CROSS) return launch_unc<NUM_MT, MODE_CROSS, 1>(ta, tw, p, st);
  return launch_unc<NUM_MT, MODE_STATS, 1>(ta, tw, p, st);
}

static int launch_unc_any(int num_mt, const CUtensorMap& ta, const CUtensorMap& tw, const UncParams& p, int mode,
                          cudaStream_t st) {
  if (num_mt == 1) return launch_unc_mt<1>(ta, tw, p, mode, st);
  if (num_mt <= 3) return launch_unc_mt<3>(ta, tw, p, mode, st);
  return launch_unc_mt<5>(ta, tw, p, mode, st);
}

}  // namespace dd

extern "C" size_t dd_uncertainty_workspace_bytes(int n_images, int n_vis, int vocab, int topk) {
  if (n_images <= 0 || n_vis <= 0 || vocab <= 0 || topk <= 0) return 0;
  return dd::unc_layout(n_images, n_vis, vocab, topk).total;
}

#ifdef DD_UNC_EXPERIMENT
// host copy of the cycle counters of the last sweep: [ctas][8] = mma {loop, wait W, wait act, wait tmem-empty}, epilogue {loop, wait tmem-full}
extern "C" int dd_uncertainty_debug_read(long long* host, int n_ctas) {
  long long* buf = nullptr;
  cudaDeviceSynchronize();
  // the static in dd_uncertainty_project is not visible here: re-derive through a second static set by a first call
  buf = g_unc_dbg;
  if (!buf) return -1;
  return cudaMemcpy(host, buf, (size_t)n_ctas * 64, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" size_t dd_uncertainty_overflow_offset(int n_images, int n_vis, int vocab, int topk) {
  if (n_images <= 0 || n_vis <= 0 || vocab <= 0 || topk <= 0) return 0;
  return dd::unc_layout(n_images, n_vis, vocab, topk).off_ovf;
}

extern "C" int dd_uncertainty_project(const void* h_vis, const void* w_lm, int n_images, int n_vis, int d,
                                      int vocab, int topk, int round_logits_bf16, float* epi, float* alea,
                                      float* var, int32_t* topk_ids, void* workspace, size_t workspace_bytes,
                                      dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(h_vis && w_lm && epi && alea && var && topk_ids && workspace, "dd_uncertainty_project: null pointer");
  DD_CHECK_ARG(n_images >= 0 && n_vis >= 1 && n_vis <= 8192, "dd_uncertainty_project: n_vis=%d out of [1,8192]",
               n_vis);
  DD_CHECK_ARG(d >= BK && d % BK == 0, "dd_uncertainty_project: d=%d must be a multiple of %d", d, BK);
  DD_CHECK_ARG(vocab >= 2, "dd_uncertainty_project: vocab=%d", vocab);
  DD_CHECK_ARG(topk >= 1 && topk <= DD_MAX_TOPK && topk <= vocab, "dd_uncertainty_project: topk=%d out of [1,%d]",
               topk, DD_MAX_TOPK);
  DD_CHECK_ARG(((uintptr_t)h_vis & 15) == 0 && ((uintptr_t)w_lm & 15) == 0,
               "dd_uncertainty_project: operands must be 16-byte aligned");
  if (n_images == 0) return DD_OK;
  const UncLayout L = unc_layout(n_images, n_vis, vocab, topk);
  if (workspace_bytes < L.total) {
    set_error("dd_uncertainty_project: workspace %zu < %zu bytes", workspace_bytes, L.total);
    return DD_ERR_WORKSPACE;
  }
  CUtensorMap ta, tw;
  {
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)n_vis, (uint64_t)n_images};
    const uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)n_vis * d * 2};
    const uint32_t box[3] = {BK, BM, 1};
    if (int rc = make_tmap_bf16_sw128(&ta, h_vis, 3, dims, strides, box, "dd_uncertainty_project(h_vis)")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)vocab};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {BK, BN};
    if (int rc = make_tmap_bf16_sw128(&tw, w_lm, 2, dims, strides, box, "dd_uncertainty_project(w_lm)")) return rc;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  UncParams p{};
  p.n_images = n_images; p.n_vis = n_vis; p.d = d; p.vocab = vocab; p.topk = topk; p.round_bf16 = round_logits_bf16;
  p.n_strips = L.n_strips; p.parts = L.parts; p.n_chunks = L.n_chunks;
  p.colpart = reinterpret_cast<float*>(ws + L.off_colpart);
  p.logpavg = reinterpret_cast<float*>(ws + L.off_logpavg);
  p.part_stats = reinterpret_cast<float*>(ws + L.off_stats);
  p.strip_max = reinterpret_cast<float*>(ws + L.off_smax);
  float* tau = reinterpret_cast<float*>(ws + L.off_tau);
  p.tau = tau;
  p.cand_cnt = reinterpret_cast<int32_t*>(ws + L.off_ccnt);
  p.cand_val = reinterpret_cast<float*>(ws + L.off_cval);
  p.cand_idx = reinterpret_cast<int32_t*>(ws + L.off_cidx);
  p.cap = L.cap;
  float* row_stats = reinterpret_cast<float*>(ws + L.off_rowstats);
  p.row_stats = row_stats;
  p.part_acq = reinterpret_cast<float*>(ws + L.off_acq);
  p.overflow = reinterpret_cast<int32_t*>(ws + L.off_ovf);
#ifdef DD_UNC_EXPERIMENT
  if (getenv("DD_UNC_DBG") && !g_unc_dbg) cudaMalloc(&g_unc_dbg, 4096 * 64);
  p.dbg = g_unc_dbg;
#endif
  cudaStream_t st = (cudaStream_t)stream;
  const int num_mt = L.n_chunks > 1 ? kMaxMT : ceil_div(n_vis, BM);
  const dim3 fgrid(ceil_div(n_vis, kFinWarps), n_images);
  int rc = launch_unc_any(num_mt, ta, tw, p, MODE_STATS, st);
  if (rc) return rc;
  stats_finalize_kernel<<<fgrid, kFinWarps * 32, 0, st>>>(p, row_stats, tau);
  DD_CHECK_LAUNCH();
  if (L.n_chunks == 1) {
    rc = launch_unc_any(num_mt, ta, tw, p, MODE_FUSED, st);
    if (rc) return rc;
  } else {
    rc = launch_unc_any(num_mt, ta, tw, p, MODE_COLSUM, st);
    if (rc) return rc;
    logpavg_kernel<<<dim3(ceil_div(L.n_strips * BN, 256), n_images), 256, 0, st>>>(p);
    DD_CHECK_LAUNCH();
    rc = launch_unc_any(num_mt, ta, tw, p, MODE_CROSS, st);
    if (rc) return rc;
  }
  fused_finalize_kernel<<<fgrid, kFinWarps * 32, 0, st>>>(p, epi, alea, var, topk_ids);
  DD_CHECK_LAUNCH();
  return DD_OK;
}
