// (1) Visual-token uncertainty stage on the 5th-gen tensor cores.
// Replaces lm_head over the visual rows + get_image_logits + get_topk_token_id +
// calculate_vision_uncertainty (models/llava.py:412-441, 710-756), see include/dropdec.h.
//
// logits[i,v] = <h_vis[i,:], w_lm[v,:]> is a [rows x vocab] GEMM (151 GFLOP per LLaVA-1.5 image) whose result the
// reference materialises in fp32 (74 MB + >=6 same-size temporaries per image).  The statistics need the logits
// three times -- the row normaliser Z_i must be complete before p = exp(l - m)/Z exists, and p_avg[v] = mean_i p[i,v]
// (a column reduction over every row of the image) must be complete before sum_v p[i,v] log p_avg[v] can be
// formed -- but the GEMM is needed ONCE: round 1 / early round 2 recomputed it per sweep (two tensor-core sweeps,
// 0.20 of the tensor peak); now the one GEMM sweep parks its logits in a scratch ("stash", bf16 when the caller asks
// for the reference's bf16-rounded logits, else fp32 -- lossless either way) and the later passes are element-wise
// kernels over that stash, which is L2-sized for a few images (37 MB per LLaVA-1.5 image in bf16).
//
//   logits_stats_kernel   persistent tcgen05 GEMM, CTA tile 256 rows x 256 vocabulary columns (two 128-lane TMEM
//                         accumulators of 256 columns = the whole TMEM), K in 64-wide blocks through a 3-stage TMA
//                         ring (A tile 32 KB + W tile 32 KB per block: 62 B/clk per SM at the tensor rate, under the
//                         ~100 B/clk an SM ingests).  M=128 N=256 K=16 MMAs run at the full tensor rate (N=96 of
//                         the old strip design issued at 56 instead of 48 clk).  Rows of all images of a group are
//                         one M dimension, so a batch has no per-image tile waste.  Epilogue (8 warps, one thread =
//                         one row): running max / sum-exp of the row over the tile's 256 columns, the row's maximum
//                         inside every 64-column chunk, and the logits themselves -> stash.
//   stats_finalize        per row: merge the per-tile (m, sum-exp) partials -> (m, 1/Z); tau = k-th largest chunk
//                         maximum (a lower bound of the row's k-th largest logit).
//   entropy_kernel        element-wise over the stash, CTA = (image, strip of 256 columns) over ALL rows of the image:
//                         pass A: p, A_i = sum p log(p+eps), Q_i = sum (p-1/V)^2, the column sums of p, and the few
//                         logits >= tau_i appended to the row's top-k candidate list (at most 64*k without ties; the
//                         list holds 96*k + 64, an overflow is reported); then log(p_avg + eps) of the strip's
//                         columns (the CTA has seen every row); pass B over the same (now L2-hot) logits:
//                         C_i = sum_v p log(p_avg+eps).  MUFU-bound (3 per logit).
//   fused_finalize        alea = -A, epi = A - C, var = Q/(V-1); top-k ids = the candidates sorted (value
//                         descending, id ascending).
// Images are processed in groups whose stash stays <= ~600 MB.  Every partial is summed in a fixed order: results
// are bit-identical run to run.  Algorithmic FLOPs = 2*n_vis*d*V per image.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "dd_common.cuh"
#include "dd_tcgen05.cuh"

namespace dd {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int GM = 256;                  // rows per CTA tile (two accumulators of BM)
constexpr int GN = 256;                  // vocabulary columns per CTA tile (MMA N)
constexpr int kChunk = 64;               // columns per "row maximum" chunk (top-k threshold granularity)
constexpr int kATile = GM * BK * 2;      // 32 KB
constexpr int kWTile = GN * BK * 2;      // 32 KB
constexpr int kStageBytes = kATile + kWTile;
constexpr int kStages = 3;
constexpr int kProdWarpA = 0, kProdWarpW = 1, kMmaWarp = 2, kEpiWarp0 = 4, kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kNumBars = 2 * kStages + 2;   // full[kStages] | empty[kStages] | tmem_full | tmem_empty
constexpr int kTmemCols = 512;
constexpr size_t kGemmSmem = 1024 + (size_t)kStages * kStageBytes + kNumBars * 8 + 16;
constexpr int kRowsCap = 4608;           // rows per image group (8 LLaVA-1.5 images): bounds the stash
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kEps = 1e-10f;  // models/llava.py:735,739

constexpr uint32_t kIdesc = make_idesc_bf16(BM, GN);

// exp2f / __logf carry a range fix-up for denormal results / arguments (4 extra instructions each); here 2^x below
// 2^-126 may flush to zero (a probability < 1e-38) and the log argument is >= 1e-10, so the bare MUFU forms are
// exact enough and identical in every pass.
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

// ------------------------------------------------------------------------------------------
struct UncParams {
  int R;                   // rows of this image group (n_img * n_vis)
  int n_img, n_vis, d, vocab, topk, round_bf16;
  int n_cs, n_rg;          // column strips of GN, row groups of GM (over R)
  int n_es;                // column strips of EC (element-wise passes)
  int Vs;                  // stash row length = n_cs * GN
  void* stash;             // [R][Vs] bf16 (round_bf16) or fp32
  float* part_stats;       // [R][n_cs][2]            (m, sum-exp) of the row inside the strip
  float* strip_max;        // [R][n_cs * GN/kChunk]   row maximum inside each 64-column chunk
  float* row_stats;        // [R][4]                  (m, 1/Z, tau = top-k candidate threshold, -)
  int32_t* cand_cnt;       // [R]
  float* cand_val;         // [R][cap]
  int32_t* cand_idx;       // [R][cap]
  int cap;
  float* part_aq;          // [n_es][R][2]            (A, Q) of the row inside the strip
  float* part_c;           // [n_es][R]               C of the row inside the strip
  int32_t* overflow;       // [1] rows whose top-k candidate list overflowed (ties at the threshold): the ids of such a
                           //     row may miss a tied candidate, so the host treats a non-zero count as an error
  int reset_overflow;      // first group of a call
};

// ------------------------------------------------------------------------------------------
// The GEMM sweep
// ------------------------------------------------------------------------------------------
// Tile order: tile i = (column strip i / n_rg, row group i % n_rg), dealt round-robin to the CTAs.  The row groups of
// one W strip are adjacent, so they run at the same time on neighbouring CTAs: the strip is read from HBM once and
// from L2 after that (ncu: 273 MB of DRAM reads for the 263 MB lm_head).  A row group whose second 128 rows lie past R
// costs half the MMAs (one accumulator); dealing those half tiles last, in snake order, looked better on paper for one
// image (2.25 instead of 2.5 tile times per SM) but re-streams all of W for them: 160 vs 152 us, 533 MB of DRAM reads.
struct TileOrder {
  int n_tiles, n_rg;
  __device__ __forceinline__ bool get(int round, int& cs, int& rg) const {
    const int i = round * gridDim.x + blockIdx.x;
    if (i >= n_tiles) return false;
    cs = i / n_rg;
    rg = i - cs * n_rg;
    return true;
  }
  __device__ __forceinline__ int rounds() const { return (n_tiles + gridDim.x - 1) / gridDim.x; }
};

template <bool BF16_STASH>
__global__ void __launch_bounds__(kGemmThreads, 1)
logits_stats_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const UncParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles (the dynamic smem base is only 16-byte aligned by contract)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.d / BK;
  const TileOrder order{p.n_cs * p.n_rg, p.n_rg};
  const int n_rounds = order.rounds();
  const uint32_t ring = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 8;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 2);    // one arrive.expect_tx by each of the two producer warps
      mbar_init(bar_empty + 8 * i, 1);   // tcgen05.commit
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, kEpiThreads);
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

  if (warp == kProdWarpA || warp == kProdWarpW) {
    // =========================== TMA producers (one elected lane issues) ===========================
    const bool is_a = warp == kProdWarpA;
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(is_a ? &tmap_a : &tmap_w) : "memory");
    uint32_t g = 0;
    for (int round = 0; round < n_rounds; ++round) {
      int cs, rg;
      if (!order.get(round, cs, rg)) continue;
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const uint32_t s = g % kStages, phase = (g / kStages) & 1;
        mbar_wait(bar_empty + 8 * s, phase ^ 1);
        if (elect_one()) {
          const uint32_t dst = ring + s * kStageBytes, bar = bar_full + 8 * s;
          if (is_a) {      // rows past R are zero-filled by the tensor map
            mbar_expect_tx(bar, kATile);
            tma_load_2d(dst, &tmap_a, bar, kb * BK, rg * GM);
          } else {         // rows past the vocabulary are zero-filled
            mbar_expect_tx(bar, kWTile);
            tma_load_2d(dst + kATile, &tmap_w, bar, kb * BK, cs * GN);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (one elected lane issues) ===========================
    uint32_t g = 0, it = 0;
    for (int round = 0; round < n_rounds; ++round) {
      int cs, rg;
      if (!order.get(round, cs, rg)) continue;
      const int nacc = (rg * GM + BM < p.R) ? 2 : 1;   // the second accumulator's rows may all lie past R
      mbar_wait(bar_tempty, (it & 1) ^ 1);             // the epilogue has drained the previous tile
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const uint32_t s = g % kStages, phase = (g / kStages) & 1;
        mbar_wait(bar_full + 8 * s, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = ring + s * kStageBytes;
          const uint64_t b_desc = make_smem_desc(base + kATile);
          for (int a = 0; a < nacc; ++a) {
            const uint64_t a_desc = make_smem_desc(base + a * BM * 128);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
              umma_bf16(tmem_base + a * GN, a_desc + 2 * k, b_desc + 2 * k, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * s);
          if (kb == num_kb - 1) umma_commit(bar_tfull);
        }
        __syncwarp();
      }
      ++it;
    }
  } else if (warp >= kEpiWarp0) {
    // =========================== epilogue (warps 4-11): one thread = one row ===========================
    const int q = warp & 3, a = (warp - kEpiWarp0) >> 2;   // TMEM lane quarter, accumulator
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * GN;
    uint32_t it = 0;
    for (int round = 0; round < n_rounds; ++round) {
      int cs, rg;
      if (!order.get(round, cs, rg)) continue;
      const int nacc = (rg * GM + BM < p.R) ? 2 : 1;
      const int row = rg * GM + a * BM + q * 32 + lane;
      const bool ok = row < p.R;
      mbar_wait(bar_tfull, it & 1);
      tc_fence_after();
      if (a < nacc) {   // (warp-uniform)
        float m = -INFINITY, z = 0.f, smax = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < GN / 32; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if (ok) {
            const int cbase = cs * GN + c * 32;
            float cmax = -INFINITY;
            if (cbase + 32 <= p.vocab) {  // all but the last strip: no per-element bound check
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (p.round_bf16) v[i] = bf16_round(v[i]);
                cmax = fmaxf(cmax, v[i]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (p.round_bf16) v[i] = bf16_round(v[i]);
                if (cbase + i >= p.vocab) v[i] = -INFINITY;
                cmax = fmaxf(cmax, v[i]);
              }
            }
            smax = fmaxf(smax, cmax);
            if (c & 1) {   // a 64-column chunk is complete
              p.strip_max[(size_t)row * (p.n_cs * (GN / kChunk)) + cs * (GN / kChunk) + (c >> 1)] = smax;
              smax = -INFINITY;
            }
            if (cmax > -INFINITY) {   // (32 columns wholly past the vocabulary contribute nothing)
              if (cmax > m) {
                z *= ex2_ftz((m - cmax) * kLog2e);
                m = cmax;
              }
              const float mneg = -m * kLog2e;
              float zs0 = 0.f, zs1 = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                zs0 += ex2_ftz(fmaf(v[i], kLog2e, mneg));
                zs1 += ex2_ftz(fmaf(v[i + 1], kLog2e, mneg));
              }
              z += zs0 + zs1;
            }
            // the logits themselves (columns past the vocabulary hold -inf; the readers mask them anyway)
            if constexpr (BF16_STASH) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.stash) + (size_t)row * p.Vs + cbase);
#pragma unroll
              for (int h = 0; h < 4; ++h)
                dst[h] = make_uint4(pack_bf16x2(v[8 * h], v[8 * h + 1]), pack_bf16x2(v[8 * h + 2], v[8 * h + 3]),
                                    pack_bf16x2(v[8 * h + 4], v[8 * h + 5]), pack_bf16x2(v[8 * h + 6], v[8 * h + 7]));
            } else {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.stash) + (size_t)row * p.Vs + cbase);
#pragma unroll
              for (int h = 0; h < 8; ++h) dst[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
            }
          }
        }
        if (ok) {
          reinterpret_cast<float2*>(p.part_stats)[(size_t)row * p.n_cs + cs] = make_float2(m, z);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty);   // the accumulators may be overwritten
      ++it;
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

// merge the per-strip (m, z) partials of one row: row_stats = (m, 1/Z); tau = k-th largest chunk maximum
// (a lower bound of the row's k-th largest logit; -inf when there are fewer than k chunks); reset the
// candidate counter
__global__ void __launch_bounds__(kFinWarps * 32) stats_finalize_kernel(const UncParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kFinWarps + (threadIdx.x >> 5);
  if (p.reset_overflow && blockIdx.x == 0 && threadIdx.x == 0) *p.overflow = 0;
  if (row >= p.R) return;  // warp-uniform
  const float2* ps = reinterpret_cast<const float2*>(p.part_stats) + (size_t)row * p.n_cs;
  float m = -INFINITY;
  for (int s = lane; s < p.n_cs; s += 32) m = fmaxf(m, ps[s].x);
  m = warp_max(m);
  float z = 0.f;
  for (int s = lane; s < p.n_cs; s += 32) {
    const float2 pr = ps[s];
    if (pr.x > -INFINITY) z += pr.y * exp2f((pr.x - m) * kLog2e);
  }
  z = warp_sum(z);
  // k-th largest chunk maximum: every lane keeps its own values sorted (descending), then k rounds of
  // "largest head over the warp, its owner pops"
  float tv[DD_MAX_TOPK];
#pragma unroll
  for (int j = 0; j < DD_MAX_TOPK; ++j) tv[j] = -INFINITY;
  const int n_chunks = p.n_cs * (GN / kChunk);
  for (int s = lane; s < n_chunks; s += 32) {
    float v = p.strip_max[(size_t)row * n_chunks + s];
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
    reinterpret_cast<float4*>(p.row_stats)[row] = make_float4(m, 1.0f / z, kth, 0.f);
    p.cand_cnt[row] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// The element-wise passes over the stash
// ------------------------------------------------------------------------------------------
// CTA = (image, strip of EC = 64 vocabulary columns), 128 threads.  The strip is walked in tiles of ER = 64 rows: a
// tile is copied stash -> shared memory with cp.async (coalesced 128-byte rows, double buffered), then TWO THREADS OWN
// ONE ROW (32 columns each; lanes 0-15 / 16-31 of a warp hold the two halves of 16 rows, so a quarter-warp touches 8
// consecutive padded rows: conflict-free 128-bit accesses): the logits come out of shared memory as 128-bit loads and
// the row sums A, Q, C are thread-local plus one shuffle (a warp per row spends as many instructions on the shuffle
// reductions and per-row bookkeeping as on the math: measured 40 instructions per logit, 65 us per image).  The column
// sums are the transposed reduction: p goes to a padded fp32 tile in shared memory and thread (column, row half)
// adds it up in row order -- fixed order, deterministic.  36 KB of shared memory: six CTAs per SM, so the 501 CTAs of one
// LLaVA-1.5 image are resident at once (with 256-thread CTAs at three per SM they ran as a full wave plus an eighth).
constexpr int EC = 64;                       // columns per strip
constexpr int ER = 64;                       // rows per tile
constexpr int kEwThreads = 2 * ER;
constexpr int kPLd = EC + 4;                 // padded row of the p tile (floats): 128-bit row stores and column reads
                                             // are both conflict-free
template <bool BF16_STASH>
struct EwCfg {
  static constexpr int kElem = BF16_STASH ? 2 : 4;
  static constexpr int kRowBytes = EC * kElem;           // 128 / 256
  static constexpr int kLd = kRowBytes + 16;             // padded row of the logits tile (bytes)
  static constexpr int kTileBytes = ER * kLd;
  static constexpr int kChunksPerRow = kRowBytes / 16;   // 8 / 16
  static constexpr size_t kSmem = 2 * (size_t)kTileBytes + (size_t)ER * kPLd * 4 + 2 * EC * 4 + EC * 4 + 16;
  static constexpr int kCtasPerSm = BF16_STASH ? 6 : 4;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [t0, t0 + ER) of the strip -> tile buffer (rows past n_vis are not touched)
template <bool BF16_STASH>
__device__ __forceinline__ void load_tile(const UncParams& p, uint32_t buf, size_t row_base, int t0, int c0) {
  using C = EwCfg<BF16_STASH>;
  const uint8_t* src0 = static_cast<const uint8_t*>(p.stash) + ((row_base + t0) * p.Vs + c0) * C::kElem;
#pragma unroll
  for (int i = 0; i < ER * C::kChunksPerRow / kEwThreads; ++i) {
    const int q = i * kEwThreads + threadIdx.x;
    const int r = q / C::kChunksPerRow, part = q % C::kChunksPerRow;
    if (t0 + r < p.n_vis) cp_async_16(buf + r * C::kLd + part * 16, src0 + (size_t)r * p.Vs * C::kElem + part * 16);
  }
  cp_async_commit();
}

// 8 consecutive logits (column group g of the strip) of a row of the tile
template <bool BF16_STASH>
__device__ __forceinline__ void tile_logits8(const uint8_t* row, int g, float (&x)[8]) {
  if constexpr (BF16_STASH) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + g * 16);
    const float2 a = bf16x2_to_float2(u.x), b = bf16x2_to_float2(u.y), c = bf16x2_to_float2(u.z), d = bf16x2_to_float2(u.w);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(row + g * 32);
    const float4 b = *reinterpret_cast<const float4*>(row + g * 32 + 16);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
}

// Rare path of pass A (noinline keeps it out of the hot loop): some of the 8 logits at column `col` reach the row's
// top-k threshold.
__device__ __noinline__ void append_candidates8(int32_t* cand_cnt, float* cand_val, int32_t* cand_idx, int cap, int vocab,
                                                size_t grow, float4 xa, float4 xb, float tau, int col) {
  const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (col + i < vocab && x[i] >= tau) {
      const int slot = atomicAdd(&cand_cnt[grow], 1);
      if (slot < cap) {
        cand_val[grow * cap + slot] = x[i];
        cand_idx[grow * cap + slot] = col + i;
      }
    }
  }
}

template <bool BF16_STASH>
__global__ void __launch_bounds__(kEwThreads, EwCfg<BF16_STASH>::kCtasPerSm) entropy_kernel(const UncParams p) {
  using C = EwCfg<BF16_STASH>;
  extern __shared__ __align__(16) uint8_t ew_smem[];
  uint8_t* s_tile = ew_smem;                                              // [2][ER][kLd]
  float* s_p = reinterpret_cast<float*>(ew_smem + 2 * C::kTileBytes);     // [ER][kPLd]
  float* s_red = s_p + ER * kPLd;                                         // [2][EC]
  float* s_lp = s_red + 2 * EC;                                           // [EC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int trow = warp * 16 + (lane & 15);    // row of the tile
  const int half = lane >> 4;                  // columns [32*half, 32*half + 32) of the strip
  const int es = blockIdx.x, img = blockIdx.y;
  const int c0 = es * EC;
  const size_t row_base = (size_t)img * p.n_vis;
  const int n_tiles = (p.n_vis + ER - 1) / ER;
  const float inv_v = 1.0f / (float)p.vocab;
  const bool full = c0 + EC <= p.vocab;        // (CTA-uniform) no column of the strip lies past the vocabulary
  const uint32_t tile_u32 = smem_u32(s_tile);
  const size_t R = (size_t)p.R;
  float* prow = s_p + trow * kPLd + half * 32;

  // ---- pass A: p, A, Q, top-k candidates, column sums ----
  float colsum = 0.f;                          // threads 0..EC-1: column tid, over the tiles in order
  load_tile<BF16_STASH>(p, tile_u32, row_base, 0, c0);
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) {
      load_tile<BF16_STASH>(p, tile_u32 + ((t + 1) & 1) * C::kTileBytes, row_base, (t + 1) * ER, c0);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                           // tile t has landed
    const int r = t * ER + trow;
    const uint8_t* xrow = s_tile + (t & 1) * C::kTileBytes + trow * C::kLd;
    const size_t grow = row_base + r;
    float acc_a = 0.f, acc_q = 0.f;            // acc_a in units of log2
    if (r < p.n_vis) {
      const float4 st = reinterpret_cast<const float4*>(p.row_stats)[grow];
      const float mneg = -st.x * kLog2e, iz = st.y, tau = st.z;
      auto body = [&](auto padded) {           // padded: the strip holds columns past the vocabulary (last strip only)
#pragma unroll 2
        for (int gi = 0; gi < 4; ++gi) {
          const int g = half * 4 + gi;
          float x[8], pr[8];
          tile_logits8<BF16_STASH>(xrow, g, x);
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pr[i] = ex2_ftz(fmaf(x[i], kLog2e, mneg)) * iz;
            float dv = pr[i] - inv_v;
            if constexpr (decltype(padded)::value) {
              if (c0 + g * 8 + i >= p.vocab) pr[i] = dv = 0.f;
              else cm = fmaxf(cm, x[i]);
            } else {
              cm = fmaxf(cm, x[i]);
            }
            float lg;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(pr[i] + kEps));
            acc_a = fmaf(pr[i], lg, acc_a);
            acc_q = fmaf(dv, dv, acc_q);
          }
          *reinterpret_cast<float4*>(prow + gi * 8) = make_float4(pr[0], pr[1], pr[2], pr[3]);
          *reinterpret_cast<float4*>(prow + gi * 8 + 4) = make_float4(pr[4], pr[5], pr[6], pr[7]);
          if (cm >= tau)   // rare
            append_candidates8(p.cand_cnt, p.cand_val, p.cand_idx, p.cap, p.vocab, grow, make_float4(x[0], x[1], x[2], x[3]),
                               make_float4(x[4], x[5], x[6], x[7]), tau, c0 + g * 8);
        }
      };
      if (full) body(std::false_type{});
      else body(std::true_type{});
    } else {
#pragma unroll
      for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(prow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 16);     // the row's two halves (whole warp: converged here)
    acc_q += __shfl_xor_sync(0xffffffffu, acc_q, 16);
    if (half == 0 && r < p.n_vis)
      reinterpret_cast<float2*>(p.part_aq)[(size_t)es * R + grow] = make_float2(acc_a * 0.6931471805599453f, acc_q);
    __syncthreads();
    {  // column sums of the tile: thread (column, row half) in row order, then the halves in order
      const int c = tid & (EC - 1), rh = tid / EC;
      const float* col = s_p + (rh * (ER / 2)) * kPLd + c;
      float sum = 0.f;
#pragma unroll 16
      for (int i = 0; i < ER / 2; ++i) sum += col[i * kPLd];
      s_red[rh * EC + c] = sum;
    }
    __syncthreads();                           // (also: every read of this logits tile and of s_p is done)
    if (tid < EC) colsum += s_red[tid] + s_red[EC + tid];
  }
  // ---- log(p_avg + eps) of the strip's columns, models/llava.py:732-735 ----
  if (tid < EC) s_lp[tid] = __logf(colsum / (float)p.n_vis + kEps);
  __syncthreads();

  // ---- pass B: C = sum_v p log(p_avg + eps); one tile: its logits are still in shared memory ----
  if (n_tiles > 1) load_tile<BF16_STASH>(p, tile_u32, row_base, 0, c0);
  for (int t = 0; t < n_tiles; ++t) {
    if (n_tiles > 1) {
      if (t + 1 < n_tiles) {
        load_tile<BF16_STASH>(p, tile_u32 + ((t + 1) & 1) * C::kTileBytes, row_base, (t + 1) * ER, c0);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
    }
    const int r = t * ER + trow;
    const uint8_t* xrow = s_tile + (t & 1) * C::kTileBytes + trow * C::kLd;
    const size_t grow = row_base + r;
    float acc_c = 0.f, iz = 0.f;
    if (r < p.n_vis) {
      const float4 st = reinterpret_cast<const float4*>(p.row_stats)[grow];
      const float mneg = -st.x * kLog2e;
      iz = st.y;
      auto body = [&](auto padded) {
#pragma unroll 2
        for (int gi = 0; gi < 4; ++gi) {
          const int g = half * 4 + gi;
          float x[8];
          tile_logits8<BF16_STASH>(xrow, g, x);
          const float4 l0 = *reinterpret_cast<const float4*>(s_lp + g * 8), l1 = *reinterpret_cast<const float4*>(s_lp + g * 8 + 4);
          const float lp[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float pr = ex2_ftz(fmaf(x[i], kLog2e, mneg));
            if constexpr (decltype(padded)::value)
              if (c0 + g * 8 + i >= p.vocab) pr = 0.f;
            acc_c = fmaf(pr, lp[i], acc_c);
          }
        }
      };
      if (full) body(std::false_type{});
      else body(std::true_type{});
    }
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, 16);
    if (half == 0 && r < p.n_vis) p.part_c[(size_t)es * R + grow] = acc_c * iz;
    if (n_tiles > 1) __syncthreads();          // the buffer of tile t is refilled by the load issued next iteration
  }
}

// alea / epi / var from the per-strip partials, top-k ids from the candidate lists.  CTA = 8 rows: thread
// (group g of 32, row r of 8) adds the strips g, g+32, ... of its row, the groups are added in order; then each warp
// sorts the candidates of one row.
constexpr int kFinRows = 8, kFinGroups = 32;
__global__ void __launch_bounds__(kFinRows * kFinGroups)
fused_finalize_kernel(const UncParams p, float* __restrict__ epi, float* __restrict__ alea, float* __restrict__ var,
                      int32_t* __restrict__ topk_ids) {
  __shared__ float s_acc[kFinGroups][kFinRows][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rr = threadIdx.x & (kFinRows - 1), g = threadIdx.x / kFinRows;
  const int row0 = blockIdx.x * kFinRows;
  {
    const int row = row0 + rr;
    float a = 0.f, q = 0.f, c = 0.f;
    if (row < p.R) {
#pragma unroll 4
      for (int s = g; s < p.n_es; s += kFinGroups) {
        const float2 aq = reinterpret_cast<const float2*>(p.part_aq)[(size_t)s * p.R + row];
        a += aq.x;
        q += aq.y;
        c += p.part_c[(size_t)s * p.R + row];
      }
    }
    s_acc[g][rr][0] = a; s_acc[g][rr][1] = q; s_acc[g][rr][2] = c;
  }
  __syncthreads();
  if (threadIdx.x < kFinRows && row0 + threadIdx.x < p.R) {
    float a = 0.f, q = 0.f, c = 0.f;
#pragma unroll
    for (int k = 0; k < kFinGroups; ++k) { a += s_acc[k][threadIdx.x][0]; q += s_acc[k][threadIdx.x][1]; c += s_acc[k][threadIdx.x][2]; }
    const size_t r = (size_t)(row0 + threadIdx.x);
    alea[r] = -a;                         // models/llava.py:739
    epi[r] = a - c;                       // models/llava.py:735
    var[r] = q / (float)(p.vocab - 1);    // torch.var (unbiased), models/llava.py:728
  }
  // top-k ids from the candidate list: k selection rounds, value descending, id ascending on ties
  const int row = row0 + warp;
  if (row >= p.R) return;   // warp-uniform
  const size_t r = (size_t)row;
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  static int cache[64];
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

struct UncLayout {
  int group;        // images per group
  int n_cs, n_es, Vs, cap;
  size_t off_stash, off_stats, off_smax, off_ccnt, off_cval, off_cidx, off_rowstats, off_aq, off_c, off_ovf, total;
};
static UncLayout unc_layout(int n_images, int n_vis, int vocab, int topk) {
  UncLayout L;
  L.group = kRowsCap / n_vis;
  if (L.group < 1) L.group = 1;
  if (L.group > n_images) L.group = n_images > 0 ? n_images : 1;
  L.n_cs = ceil_div(vocab, GN);
  L.n_es = ceil_div(vocab, 64);
  L.Vs = L.n_cs * GN;
  L.cap = 96 * topk + 64;  // without ties at most kChunk*topk elements reach the threshold
  const size_t R = (size_t)L.group * n_vis;
  size_t o = 0;
  L.off_stash = o; o = align_up(o + R * L.Vs * sizeof(float), 256);   // (bf16 logits use half of it)
  L.off_stats = o; o = align_up(o + (size_t)L.n_cs * R * 2 * sizeof(float), 256);
  L.off_smax = o; o = align_up(o + (size_t)L.n_cs * (GN / kChunk) * R * sizeof(float), 256);
  L.off_ccnt = o; o = align_up(o + R * sizeof(int32_t), 256);
  L.off_cval = o; o = align_up(o + R * L.cap * sizeof(float), 256);
  L.off_cidx = o; o = align_up(o + R * L.cap * sizeof(int32_t), 256);
  L.off_rowstats = o; o = align_up(o + R * 4 * sizeof(float), 256);
  L.off_aq = o; o = align_up(o + (size_t)L.n_es * R * 2 * sizeof(float), 256);
  L.off_c = o; o = align_up(o + (size_t)L.n_es * R * sizeof(float), 256);
  L.off_ovf = o; o = align_up(o + sizeof(int32_t), 256);
  L.total = o;
  return L;
}

template <bool BF16_STASH>
static int run_group(const CUtensorMap& ta, const CUtensorMap& tw, const UncParams& p, float* epi, float* alea,
                     float* var, int32_t* topk_ids, cudaStream_t st) {
  {
    // the opt-in for > 48 KB of dynamic shared memory is a per-device attribute
    static bool attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      DD_CHECK_CUDA(cudaFuncSetAttribute(logits_stats_kernel<BF16_STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kGemmSmem));
      DD_CHECK_CUDA(cudaFuncSetAttribute(entropy_kernel<BF16_STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)EwCfg<BF16_STASH>::kSmem));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  const int tiles = p.n_cs * p.n_rg, sms = num_sms();
  logits_stats_kernel<BF16_STASH><<<tiles < sms ? tiles : sms, kGemmThreads, kGemmSmem, st>>>(ta, tw, p);
  DD_CHECK_LAUNCH();
  const dim3 fgrid(ceil_div(p.R, kFinWarps));
  stats_finalize_kernel<<<fgrid, kFinWarps * 32, 0, st>>>(p);
  DD_CHECK_LAUNCH();
  entropy_kernel<BF16_STASH><<<dim3(p.n_es, p.n_img), kEwThreads, EwCfg<BF16_STASH>::kSmem, st>>>(p);
  DD_CHECK_LAUNCH();
  fused_finalize_kernel<<<ceil_div(p.R, kFinRows), kFinRows * kFinGroups, 0, st>>>(p, epi, alea, var, topk_ids);
  DD_CHECK_LAUNCH();
  return DD_OK;
}

}  // namespace dd

extern "C" size_t dd_uncertainty_workspace_bytes(int n_images, int n_vis, int vocab, int topk) {
  if (n_images <= 0 || n_vis <= 0 || vocab <= 0 || topk <= 0) return 0;
  return dd::unc_layout(n_images, n_vis, vocab, topk).total;
}

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
  CUtensorMap tw;
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)vocab};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {BK, GN};
    if (int rc = make_tmap_bf16_sw128(&tw, w_lm, 2, dims, strides, box, "dd_uncertainty_project(w_lm)")) return rc;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  UncParams p{};
  p.n_vis = n_vis; p.d = d; p.vocab = vocab; p.topk = topk; p.round_bf16 = round_logits_bf16;
  p.n_cs = L.n_cs; p.n_es = L.n_es; p.Vs = L.Vs; p.cap = L.cap;
  p.stash = ws + L.off_stash;
  p.part_stats = reinterpret_cast<float*>(ws + L.off_stats);
  p.strip_max = reinterpret_cast<float*>(ws + L.off_smax);
  p.cand_cnt = reinterpret_cast<int32_t*>(ws + L.off_ccnt);
  p.cand_val = reinterpret_cast<float*>(ws + L.off_cval);
  p.cand_idx = reinterpret_cast<int32_t*>(ws + L.off_cidx);
  p.row_stats = reinterpret_cast<float*>(ws + L.off_rowstats);
  p.part_aq = reinterpret_cast<float*>(ws + L.off_aq);
  p.part_c = reinterpret_cast<float*>(ws + L.off_c);
  p.overflow = reinterpret_cast<int32_t*>(ws + L.off_ovf);
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n_images; i0 += L.group) {
    const int gi = n_images - i0 < L.group ? n_images - i0 : L.group;
    p.n_img = gi;
    p.R = gi * n_vis;
    p.n_rg = ceil_div(p.R, GM);
    p.reset_overflow = i0 == 0;
    const size_t row0 = (size_t)i0 * n_vis;
    CUtensorMap ta;
    {
      const uint64_t dims[2] = {(uint64_t)d, (uint64_t)p.R};
      const uint64_t strides[1] = {(uint64_t)d * 2};
      const uint32_t box[2] = {BK, GM};
      if (int rc = make_tmap_bf16_sw128(&ta, static_cast<const __nv_bfloat16*>(h_vis) + row0 * d, 2, dims, strides, box,
                                        "dd_uncertainty_project(h_vis)"))
        return rc;
    }
    const int rc = round_logits_bf16
                       ? run_group<true>(ta, tw, p, epi + row0, alea + row0, var + row0, topk_ids + row0 * topk, st)
                       : run_group<false>(ta, tw, p, epi + row0, alea + row0, var + row0, topk_ids + row0 * topk, st);
    if (rc) return rc;
  }
  return DD_OK;
}
