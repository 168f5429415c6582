// Weight-streaming GEMM on the tcgen05 tensor cores for the batched single-token forwards of the ensemble:
//   y[m,n] = sum_k x[m,k] w[n,k],   M = (#sequences x #contexts of a pass) <= 256 rows.
// Replaces the weight GEMMs inside the reference's 1+K `self.language_model(...)` calls per token
// (models/llava.py:294-303, 350-359; models/llama.py:27-47).
//
// The GEMM is a pure weight stream (intensity = M FLOP/byte, at or below the B200 ridge), so it is organised
// around HBM: every SM must pull the same number of weight bytes, exactly once, with nothing serial behind the
// stream.  Round 1 cut 256-feature tiles into K slices (stream-K) and paid for it with a reduction tail as long
// as the main loop (fp32 partials through L2: flush, wait, reduce).  This version never sends a partial through L2:
//   * the token rows are the M operand of tcgen05.mma (one accumulator per 128 rows, D[row, feature] in TMEM), the
//     WEIGHTS are the N operand, so the feature width F of a tile is free in steps of 16 and is chosen per shape
//     such that there are (just under) as many tiles as SMs: F = ceil16(N / #SMs), T = ceil(N / F) CTAs, each
//     streaming its F x K slab of W once over the full K.  (qkv 12288 -> 128 x 96, gate|up 2 x 11008 -> 138 x
//     (80+80), lm_head 32064 -> 144 x 224, o/down 4096 -> 128 x 32.)
//   * per 64-wide k-block a CTA ingests F x 128 B of W (HBM, L2 evict-first) and Mp x 128 B of x (L2-resident,
//     Mp = M rounded up to 16) through TMA into a 3..8-stage mbarrier ring.  FOUR producer warps: measured on
//     B200 (tools/tma_probe.cu, profiles/r02/tma_probe.jsonl) one warp issues one bulk-tensor copy per ~600-670
//     cycles whatever its size and warps scale linearly (1 warp x 16 KB boxes: 24 B/clk per SM, 4 warps: 97
//     B/clk), so the copies of the ring are dealt round-robin to the producer warps.
//   * a tcgen05.mma with operands in shared memory costs max(N/2, 32 + N/4) cycles (the tensor pipe, or reading the
//     128-row A tile + the N-row B tile from shared memory; measured ~56 cycles at N=96), so a thin tile (o_proj /
//     down: 4096 features over 148 SMs = 32 per SM) would be bound by MMA issue, not by HBM.  Shapes whose
//     one-wave tile would be narrower than 128 features are therefore cut along K inside a thread-block CLUSTER of
//     2 or 4 CTAs (qkv 12288 -> 70 tiles x 176 x 2 K-halves, o/down -> 32 x 128 x 4 K-quarters); after the main
//     loop the CTAs of a cluster exchange their fp32 partials through distributed shared memory (each CTA owns a
//     range of the tile's 16-feature chunks, the peers write their partial of those chunks into the idle operand
//     ring of the owner, two cluster barriers), and the owner sums the ranks in rank order: deterministic, no
//     global-memory round trip, no spinning on counters.
//   * SwiGLU: the tile's W operand is [Fg gate rows | Fg up rows] (two boxes), so a thread of the epilogue holds
//     gate and up of the same feature and applies the activation itself.
//   * warps 0-3 = TMA producers, warp 4 = MMA issuer (one elected lane), warps 5-12 = epilogue (tcgen05.ld
//     32x32b: one thread = one token row, 16 features per load = one 32-byte store) with the HF rounding points
//     of dd_skinny_gemm.
//   * 129..256 rows on the wide shapes: CTA pairs (cta_group::2), see pair_gemm_kernel below.
// Every output element is one fp32 accumulation over k in ascending order (per K slice), whatever M is: results are
// row-independent and bit-reproducible.  No workspace is used (the parameter is kept for ABI stability).
// Algorithmic bytes per launch: N*K*2 (x2 for SwiGLU) -- the weights, read exactly once.
#include <cstdlib>

#include "dd_common.cuh"
#include "dd_tcgen05.cuh"

namespace dd {

constexpr int SG_BM = 128;                     // token rows per accumulator (UMMA M)
constexpr int SG_BK = 64;                      // one 128-byte swizzle row of bf16
constexpr int SG_PROD_WARPS = 4;
constexpr int SG_MMA_WARP = SG_PROD_WARPS;
constexpr int SG_EPI_WARP0 = SG_MMA_WARP + 1;
constexpr int SG_EPI_WARPS = 8;
constexpr int SG_EPI_THREADS = SG_EPI_WARPS * 32;
constexpr int SG_THREADS = (SG_EPI_WARP0 + SG_EPI_WARPS) * 32;
constexpr int SG_MAX_STAGES = 8;
constexpr int SG_SMEM_LIMIT = 227 * 1024;
constexpr size_t SG_WS_BYTES = 16384;          // nominal (unused) workspace

struct SgParams {
  int M, Mp, N, K, epi;
  int F;               // output features per tile (SwiGLU: Fg gate + Fg up columns, F = Fg)
  int Fw;              // N extent of the MMA = W rows per k-block (2F for SwiGLU)
  int nacc;            // accumulators (128-row groups): 1 or 2
  int T, KB, G;        // tiles, k-blocks (of this CTA's K slice), CTAs
  int ring_bytes;      // operand ring (also the exchange buffer of the cluster reduction)
  int S;               // CTAs per cluster = K slices per tile (1, 2, 4)
  int max_own;         // 16-feature chunks a cluster rank owns at most
  int bks;             // k-blocks per pipeline stage (one mbarrier round trip per stage)
  int x_kb, x_bytes, w_kb, stage_bytes;   // smem: bytes per k-block of x / per stage of x / per k-block of W / per stage
  int stages, acc_stages, tmem_cols;
  const __nv_bfloat16* residual;
  void* y;
#ifdef DD_SG_EXPERIMENT
  long long* dbg;      // [G][8]: producer-0 {loop, wait} cycles, MMA {loop, wait} cycles, epilogue {wait, work} cycles
#endif
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

#ifdef DD_SG_EXPERIMENT
__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
  return t;
}
#define SG_STAMP(i) do { if (p.dbg && (threadIdx.x == SG_EPI_WARP0 * 32) && rank == 0) p.dbg[cta * 16 + 8 + (i)] = gtime_ns(); } while (0)
#else
#define SG_STAMP(i) do { } while (0)
#endif
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// 16 finished fp32 dot products of one token row (features f .. f+15) -> the output, with the rounding points of
// the HF decoder (see dd_skinny_gemm): bf16(acc), + residual, fp32 logits, bf16(bf16(silu(g)) * u)
__device__ __forceinline__ void sg_store16(const SgParams& p, int m, int f, const float (&v0)[16], const float (&v1)[16]) {
  const size_t o = (size_t)m * p.N + f;
  const int nvalid = min(16, p.N - f);   // N % 8 == 0: 16 or 8
  if (p.epi == DD_EPI_F32) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + o);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nvalid)
        dst[i] = make_float4(bf16_round(v0[4 * i]), bf16_round(v0[4 * i + 1]), bf16_round(v0[4 * i + 2]),
                             bf16_round(v0[4 * i + 3]));
    return;
  }
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = bf16_round(v0[i]);
  if (p.epi == DD_EPI_RESIDUAL) {
    const uint4* rs = reinterpret_cast<const uint4*>(p.residual + o);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (8 * h < nvalid) {
        const uint4 q = rs[h];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = bf16x2_to_float2(w[i]);
          r[8 * h + 2 * i] += t.x;
          r[8 * h + 2 * i + 1] += t.y;
        }
      }
    }
  } else if (p.epi == DD_EPI_SWIGLU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = bf16_round(r[i] / (1.f + __expf(-r[i]))) * bf16_round(v1[i]);
  }
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + o);
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (8 * h < nvalid)
      dst[h] = make_uint4(pack_bf16x2(r[8 * h], r[8 * h + 1]), pack_bf16x2(r[8 * h + 2], r[8 * h + 3]),
                          pack_bf16x2(r[8 * h + 4], r[8 * h + 5]), pack_bf16x2(r[8 * h + 6], r[8 * h + 7]));
}

__global__ void __launch_bounds__(SG_THREADS, 1)
stream_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                   const SgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
  // bars: full[SG_MAX_STAGES] | empty[SG_MAX_STAGES] | tmem_full[2] | tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SG_MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = p.S > 1 ? (int)cluster_ctarank() : 0;   // K slice of the tile
  const int cta = blockIdx.x / p.S;                        // tile sequence index (clusters are consecutive CTAs)
  const bool swiglu = p.epi == DD_EPI_SWIGLU;
  const int n_stage_kb = p.KB / p.bks;   // pipeline stages per tile

#ifdef DD_SG_EXPERIMENT
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cta * 16 + 6] = gtime_ns();
#endif
  const uint32_t ring = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * SG_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * SG_MAX_STAGES, bar_tempty = bar_tfull + 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SG_MAX_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);      // one arrive.expect_tx (all the stage's bytes) by the stage's producer
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, SG_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == SG_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = cta < p.T ? (p.T - cta + p.G - 1) / p.G : 0;   // tiles cta, cta + G, ...

  if (warp < SG_PROD_WARPS) {
    // =========================== TMA producers ===========================
    // pipeline stage i of this CTA's tile sequence (bks k-blocks: x tile + W tile each) is filled by warp i % 4:
    // one wait on the slot's empty barrier, one expect_tx, then bks * nops copies back to back (a copy costs the
    // issuing thread ~170 cycles, a barrier wait ~190: tools/tma_probe2.cu)
    if (warp == 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    const uint64_t pol = l2_policy_evict_first();
    const int n_stage_total = my_tiles * n_stage_kb;
    const uint32_t stage_tx = (uint32_t)(p.bks * (p.Mp + p.Fw) * 128);
    // ring slot s always belongs to warp s % 4, so the uses of a slot are issued in order by ONE warp (with several
    // warps per slot a warp two uses ahead would see the slot's empty barrier in the parity it waits for)
#ifdef DD_SG_EXPERIMENT
    long long dbg_wait = 0;
    const long long dbg_t0 = clock64();
#endif
    // one pipeline stage: its expect_tx, then the W boxes and / or the x boxes of its bks k-blocks
    auto fill = [&](int i, int stage, bool arm, bool do_w, bool do_x) {
      const int t = i / n_stage_kb, sk = i - t * n_stage_kb;
      const int tile = cta + t * p.G;
      const uint32_t dst = ring + stage * stage_bytes, bar = bar_full + 8 * stage;
      if (arm) mbar_expect_tx(bar, stage_tx);
      for (int b = 0; b < p.bks; ++b) {
        const int kcol = ((rank * n_stage_kb + sk) * p.bks + b) * SG_BK;
        if (do_x) tma_load_2d(dst + b * p.x_kb, &tmap_x, bar, kcol, 0);
        if (do_w) {
          const uint32_t wdst = dst + p.x_bytes + b * p.w_kb;
          tma_load_2d_hint(wdst, &tmap_w, bar, kcol, tile * p.F, pol);      // (rows past the end of W are zero-filled)
          if (swiglu)                                                       // [F gate rows | F up rows]
            tma_load_2d_hint(wdst + p.F * 128, &tmap_w, bar, kcol, p.N + tile * p.F, pol);
        }
      }
    };
    // Round 0 (the ring is empty).  The weights do not depend on the previous kernel of the stream: under a
    // programmatic dependent launch the whole ring of W tiles is put in flight BEFORE griddepcontrol.wait, the
    // activation tiles after it.
    for (int stage = warp; stage < p.stages && stage < n_stage_total; stage += SG_PROD_WARPS) {
      if (elect_one()) fill(stage, stage, true, true, false);
      __syncwarp();
    }
    pdl_wait();
    for (int stage = warp; stage < p.stages && stage < n_stage_total; stage += SG_PROD_WARPS) {
      if (elect_one()) fill(stage, stage, false, false, true);
      __syncwarp();
    }
    for (int round = 1; round * p.stages < n_stage_total; ++round) {
      const uint32_t phase = (uint32_t)(round & 1);
      for (int stage = warp; stage < p.stages; stage += SG_PROD_WARPS) {
        const int i = round * p.stages + stage;
        if (i >= n_stage_total) break;
#ifdef DD_SG_EXPERIMENT
        const long long dbg_w0 = clock64();
#endif
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
#ifdef DD_SG_EXPERIMENT
        dbg_wait += clock64() - dbg_w0;
#endif
        if (elect_one()) fill(i, stage, true, true, true);
        __syncwarp();
      }
    }
    pdl_trigger_stream_tail();   // every load of this CTA is issued: the next kernel of the chain may take its place
#ifdef DD_SG_EXPERIMENT
    if (p.dbg && warp == 0 && lane == 0 && rank == 0) { p.dbg[cta * 16 + 0] = clock64() - dbg_t0; p.dbg[cta * 16 + 1] = dbg_wait; }
#endif
  } else if (warp == SG_MMA_WARP) {
    // =========================== MMA issuer ===========================
    pdl_wait();
    const uint32_t idesc = make_idesc_bf16(SG_BM, p.Fw);
    int stage = 0, acc_stage = 0;
    uint32_t phase = 0, acc_phase = 0;
#ifdef DD_SG_EXPERIMENT
    long long dbg_wait = 0;
    const long long dbg_t0 = clock64();
#endif
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(bar_tempty + 8 * acc_stage, acc_phase ^ 1);  // the epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc_stage * p.nacc * p.Fw;
      for (int sk = 0; sk < n_stage_kb; ++sk) {
#ifdef DD_SG_EXPERIMENT
        const long long dbg_w0 = clock64();
#endif
        mbar_wait(bar_full + 8 * stage, phase);
#ifdef DD_SG_EXPERIMENT
        dbg_wait += clock64() - dbg_w0;
#endif
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = ring + stage * stage_bytes;
          for (int b = 0; b < p.bks; ++b) {
            const uint64_t b_desc = make_smem_desc(base + p.x_bytes + b * p.w_kb);
            for (int a = 0; a < p.nacc; ++a) {
              const uint64_t a_desc = make_smem_desc(base + b * p.x_kb + a * SG_BM * 128);
#pragma unroll
              for (int k = 0; k < SG_BK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
                umma_bf16(d0 + a * p.Fw, a_desc + 2 * k, b_desc + 2 * k, idesc, (sk > 0 || b > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
          if (sk == n_stage_kb - 1) umma_commit(bar_tfull + 8 * acc_stage);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
    }
#ifdef DD_SG_EXPERIMENT
    if (p.dbg && lane == 0 && rank == 0) { p.dbg[cta * 16 + 2] = clock64() - dbg_t0; p.dbg[cta * 16 + 3] = dbg_wait; p.dbg[cta * 16 + 5] = gtime_ns(); }
#endif
  } else {
    // =========================== epilogue (warps 5..12) ===========================
    pdl_wait();   // (reads the residual, writes y)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int g = (warp - SG_EPI_WARP0) >> 2;     // warp group 0 / 1
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks = p.F / 16;
    // two accumulators: group g owns accumulator g (rows g*128 ..); one: the groups alternate over the 16-feature chunks
    const int a = p.nacc == 2 ? g : 0;
    const int c0 = p.nacc == 2 ? 0 : g, cstep = p.nacc == 2 ? 1 : 2;
    const int m = a * SG_BM + q * 32 + lane;     // token row of this thread
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = cta + t * p.G;
#ifdef DD_SG_EXPERIMENT
      const long long dbg_e0 = clock64();
#endif
      mbar_wait(bar_tfull + 8 * acc_stage, acc_phase);
#ifdef DD_SG_EXPERIMENT
      const long long dbg_e1 = clock64();
#endif
      tc_fence_after();
      SG_STAMP(0);          // accumulators complete
      if (p.S > 1) break;   // split tile (always the CTA's only one): the cluster reduction below finishes it
      const uint32_t t0 = lane_addr + acc_stage * p.nacc * p.Fw + a * p.Fw;
      for (int c = c0; c < n_chunks; c += cstep) {
        float v0[16], v1[16];
        tmem_ld16(t0 + c * 16, v0);
        if (swiglu) tmem_ld16(t0 + p.F + c * 16, v1);
        const int f = tile * p.F + c * 16;
        if (m < p.M && f < p.N) sg_store16(p, m, f, v0, swiglu ? v1 : v0);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc_stage);
#ifdef DD_SG_EXPERIMENT
      if (p.dbg && threadIdx.x == SG_EPI_WARP0 * 32) { p.dbg[cta * 16 + 4] = clock64() - dbg_e1; }
#endif
      if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  if (p.S > 1) {
    // ---- split-K inside the cluster: exchange the fp32 partials through distributed shared memory ----
    // rank r owns chunks [r*nc/S, (r+1)*nc/S) of the tile's nc 16-feature chunks.  Buffer in the (now idle) operand
    // ring of the owner: [source rank (own slot skipped)][own chunk][16 features][rows] floats, so that the 32 lanes
    // of a warp (32 token rows) write / read 128 contiguous bytes.
    const int n_chunks = p.F / 16, rows = p.nacc * SG_BM;
    const bool epi = warp >= SG_EPI_WARP0;
    const int q = warp & 3, g = (warp - SG_EPI_WARP0) >> 2;
    const int a = p.nacc == 2 ? g : 0;
    const int c0 = p.nacc == 2 ? 0 : g, cstep = p.nacc == 2 ? 1 : 2;
    const int row = a * SG_BM + q * 32 + lane;
    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + a * p.Fw;
    tc_fence_before();
    cluster_sync_all();     // #1: every CTA's accumulators are complete and its operand ring is idle
    tc_fence_after();
    SG_STAMP(1);
    if (epi) {
      for (int c = c0; c < n_chunks; c += cstep) {
        const int owner = ((c + 1) * p.S - 1) / n_chunks;          // the rank r with r*nc/S <= c < (r+1)*nc/S
        if (owner == rank) continue;
        const int lc = c - owner * n_chunks / p.S;
        const int src = rank < owner ? rank : rank - 1;
        if (a * SG_BM + q * 32 >= p.M) continue;                   // (warp-uniform) no token row in this warp
        float v[16];
        tmem_ld16(t0 + c * 16, v);
        const uint32_t dst = mapa_cluster(ring + (uint32_t)(((src * p.max_own + lc) * 16 * rows + row) * 4), (uint32_t)owner);
        if (row < p.M) {
#pragma unroll
          for (int e = 0; e < 16; ++e) st_cluster_f32(dst + e * rows * 4, v[e]);
        }
      }
    }
    SG_STAMP(2);
    cluster_sync_all();     // #2: the peers' partials of my chunks have landed in my shared memory
    SG_STAMP(3);
    if (epi) {
      const int own_b = rank * n_chunks / p.S, own_e = (rank + 1) * n_chunks / p.S;
      for (int c = own_b + c0; c < own_e; c += cstep) {
        if (a * SG_BM + q * 32 >= p.M) break;                      // (warp-uniform) no token row in this warp
        const int lc = c - own_b;
        float acc[16];
        for (int r = 0; r < p.S; ++r) {                            // rank order, whoever sums
          float v[16];
          if (r == rank) {
            tmem_ld16(t0 + c * 16, v);
          } else {
            const int src = r < rank ? r : r - 1;
            const uint32_t srca = ring + (uint32_t)(((src * p.max_own + lc) * 16 * rows + row) * 4);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = ld_shared_f32(srca + e * rows * 4);
          }
          if (r == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = v[e];
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] += v[e];
          }
        }
        const int f = cta * p.F + c * 16;
        if (row < p.M && f < p.N) sg_store16(p, row, f, acc, acc);
      }
    }
  }

  SG_STAMP(4);            // outputs stored
  tc_fence_before();
  __syncthreads();
#ifdef DD_SG_EXPERIMENT
  if (p.dbg && threadIdx.x == 0 && rank == 0) p.dbg[cta * 16 + 7] = gtime_ns();
#endif
  if (warp == SG_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int sg_num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  static int cache[64];
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

struct SgPlan {
  int Mp, F, Fw, nacc, S, max_own, T, KB, G, bks, x_kb, x_bytes, w_kb, stage_bytes, stages, acc_stages, tmem_cols;
  size_t smem;
};
static SgPlan sg_plan(int M, int N, int K, int epi) {
  SgPlan s;
  const bool swiglu = epi == DD_EPI_SWIGLU;
  const int sms = sg_num_sms();
  s.Mp = (M + 15) / 16 * 16;
  s.nacc = M > SG_BM ? 2 : 1;
  // Tile width F (multiple of 16, MMA N extent = F or 2F <= 256) and K slices S per tile (cluster size): the
  // smallest S in {1, 2, 4} whose one-wave tile is at least 128 features wide -- below that the MMA issue rate, not
  // HBM, bounds the stream.  Clusters of 4 only place on 132 of the 148 SMs (GPC boundaries).  The choice depends
  // on N and K only, never on M: a token row sums the same products in the same order whatever rides in the pass.
  const int total_kb = K / SG_BK;
  const int fmax = swiglu ? 128 : 256;
  s.S = 1;
  int force_s = 0;
#ifdef DD_SG_EXPERIMENT
  if (const char* e = getenv("DD_SG_S")) force_s = atoi(e);
#endif
  for (int S = 1; S <= 4; S *= 2) {
    if (total_kb % S) break;
    s.S = S;
    const int slots = (S == 4 ? sms / 37 * 33 : sms) / S;        // clusters that are co-resident
    int F = (ceil_div(N, slots > 0 ? slots : 1) + 15) / 16 * 16;
    if (F > fmax) F = fmax;
    s.F = F;
    if (force_s ? (S == force_s || swiglu) : ((swiglu ? 2 * F : F) >= 128 || swiglu)) break;   // (SwiGLU pairs gate / up columns: never split)
  }
  s.Fw = swiglu ? 2 * s.F : s.F;
  s.T = ceil_div(N, s.F);
  s.KB = total_kb / s.S;
  {
    const int slots = (s.S == 4 ? sms / 37 * 33 : sms) / s.S;
    if (s.S > 1 && s.T > slots) {   // more tiles than one wave of clusters: fall back to whole-K tiles
      s.S = 1;
      s.KB = total_kb;
    }
    s.G = s.T < slots || s.S > 1 ? s.T : (s.T < sms ? s.T : sms);
  }
  s.max_own = ceil_div(s.F / 16, s.S);
  // k-blocks per pipeline stage: a stage costs the producer and the MMA issuer one mbarrier round trip each
  // (~300 cycles in all), so thin operands ride several k-blocks per stage; a stage stays <= 48 KB and the ring at
  // least 4 deep
  s.x_kb = s.Mp * 128;
  s.w_kb = s.Fw * 128;
  const int slack = (s.nacc * SG_BM - s.Mp) * 128;   // the last k-block's MMA reads 128-row tiles past Mp rows
  s.bks = 1;
  for (int b = 8; b > 1; b >>= 1) {
    if (s.KB % b) continue;
    const int sb = b * (s.x_kb + s.w_kb) + slack;
    if (sb <= 48 * 1024 && (SG_SMEM_LIMIT - 2048) / sb >= 4) { s.bks = b; break; }
  }
  s.x_bytes = (s.bks * s.x_kb + slack + 1023) & ~1023;
  s.stage_bytes = s.x_bytes + s.bks * s.w_kb;
  const int stage_bytes = s.stage_bytes;
  const int fixed = 1024 + (2 * SG_MAX_STAGES + 4) * 8 + 16;
  int st = (SG_SMEM_LIMIT - fixed) / stage_bytes;
  s.stages = st > SG_MAX_STAGES ? SG_MAX_STAGES : st;
  // the cluster reduction parks (S - 1) x max_own chunks of fp32 partials in the ring
  const size_t xchg = s.S > 1 ? (size_t)(s.S - 1) * s.max_own * 16 * s.nacc * SG_BM * 4 : 0;
  size_t ring = (size_t)s.stages * stage_bytes;
  if (ring < xchg) ring = xchg;
  s.smem = (size_t)fixed + ring;
  s.acc_stages = (s.S == 1 && s.T > s.G && 2 * s.nacc * s.Fw <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < s.acc_stages * s.nacc * s.Fw) cols *= 2;
  s.tmem_cols = cols;
  return s;
}

// =============================================================================================
// 129..256 rows, wide shapes (qkv, gate|up, lm_head): CTA PAIRS (tcgen05 cta_group::2).
// At 256 rows a single CTA needs two accumulators, i.e. two MMAs per W tile and k-step: the stream is bound by the
// tensor pipe of the SM (qkv: 704 cycles per 22 KB of W), not by HBM, and every CTA re-reads all 256 rows of x.  A pair
// of CTAs (two SMs of one TPC) shares one tile instead: MMA M = 256 with rows [0,128) in the leader's TMEM and
// [128,256) in its peer's, the W tile (B operand) split by rows between the two shared memories, x split by rows.  Per
// k-block a CTA now ingests 128 rows of x (16 KB) + HALF the W tile and its tensor pipe does one MMA per k-step: twice
// the W bytes per tensor-pipe cycle and per ingested byte.  74 pairs, tile width F = ceil16(N / (74 r)) <= 256.
//   * both CTAs run TMA producers; the peer's copies signal the LEADER's full barrier (cp.async.bulk.tensor
//     .cta_group::2, the barrier address mapped to rank 0), the leader arms it with the bytes of both;
//   * the leader's MMA warp issues tcgen05.mma.cta_group::2; tcgen05.commit.multicast releases the ring slot / hands
//     the accumulators over in BOTH CTAs; the peer's epilogue threads arrive on the leader's tmem_empty barrier
//     through the cluster;
//   * SwiGLU: one MMA for the gate rows and one for the up rows of a tile (each split between the CTAs), gate and up
//     columns of a feature in the same TMEM lane as before.
// Same per-element arithmetic as the single-CTA kernel: one fp32 accumulation over k in ascending order.
// =============================================================================================
constexpr int PG_PAIRS_MAX = 74;
struct PgParams {
  int M, N, K, epi;
  int F;               // output features per tile (SwiGLU: gate features = up features = F)
  int Fw;              // accumulator columns per tile (2F for SwiGLU)
  int T, KB, G;        // tiles, k-blocks, pairs
  int w_half;          // W bytes per k-block and CTA
  int bks;             // k-blocks per pipeline stage (one mbarrier round trip per stage)
  int stage_bytes, stages, acc_stages;
  int S;               // K slices accumulated separately and added in slice order (1, 2, 4): the same summation order as
                       // the single-CTA kernel's cluster split-K of this shape, so a row's result does not depend on M
  const __nv_bfloat16* residual;
  void* y;
};

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far
// have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__device__ __forceinline__ void pg_store16(const PgParams& p, int m, int f, const float (&v0)[16], const float (&v1)[16]) {
  SgParams q{};
  q.N = p.N; q.epi = p.epi; q.residual = p.residual; q.y = p.y;
  sg_store16(q, m, f, v0, v1);
}

__global__ void __launch_bounds__(SG_THREADS, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const PgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  // bars: full[SG_MAX_STAGES] | empty[SG_MAX_STAGES] | tmem_full[2] | tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SG_MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();       // 0 = leader (rows 0..127, first half of the W tile)
  const int pair = blockIdx.x >> 1;
  const bool swiglu = p.epi == DD_EPI_SWIGLU;
  constexpr int kXBytes = SG_BM * 128;           // 128 rows x 64 bf16

  const uint32_t ring = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * SG_MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * SG_MAX_STAGES, bar_tempty = bar_tfull + 16;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SG_MAX_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);      // (leader's is used) one arrive.expect_tx with the bytes of both CTAs
      mbar_init(bar_empty + 8 * i, 1);     // tcgen05.commit (multicast)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 2 * SG_EPI_THREADS);   // (leader's is used) the epilogue threads of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == SG_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = pair < p.T ? (p.T - pair + p.G - 1) / p.G : 0;   // tiles pair, pair + G, ...

  if (warp < SG_PROD_WARPS) {
    // =========================== TMA producers (both CTAs; slot s belongs to warp s % 4) ===========================
    if (warp == 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    const uint64_t pol_w = l2_policy_evict_first(), pol_x = l2_policy_evict_last();
    const uint32_t stage_tx = (uint32_t)(2 * p.bks * (kXBytes + p.w_half));
    const int half_rows = p.F / 2;               // W rows of the tile held by one CTA (per gate / up part)
    const int n_stage_kb = p.KB / p.bks;         // pipeline stages per tile
    const int n_stage_total = my_tiles * n_stage_kb;
    const uint32_t sub = (uint32_t)(kXBytes + p.w_half);
    // one pipeline stage: (leader) its expect_tx, then the W boxes and / or the x boxes of its bks k-blocks
    auto fill = [&](int i, int stage, bool arm, bool do_w, bool do_x) {
      const int t = i / n_stage_kb, sk = i - t * n_stage_kb;
      const int tile = pair + t * p.G;
      const uint32_t dst = ring + stage * p.stage_bytes;
      const uint32_t bar = mapa_cluster(bar_full + 8 * stage, 0);
      if (arm && rank == 0) mbar_expect_tx(bar_full + 8 * stage, stage_tx);
      const int wrow = tile * p.F + rank * half_rows;
      for (int b = 0; b < p.bks; ++b) {
        const int kcol = (sk * p.bks + b) * SG_BK;
        if (do_x) tma_load_2d_pair(dst + b * sub, &tmap_x, bar, kcol, rank * SG_BM, pol_x);   // rows past M: zero-filled
        if (do_w) {
          tma_load_2d_pair(dst + b * sub + kXBytes, &tmap_w, bar, kcol, wrow, pol_w);   // (rows past the end of W too)
          if (swiglu) tma_load_2d_pair(dst + b * sub + kXBytes + half_rows * 128, &tmap_w, bar, kcol, p.N + wrow, pol_w);
        }
      }
    };
    // round 0 (empty ring): the weights do not depend on the previous kernel of the stream -- under a programmatic
    // dependent launch the whole ring of W tiles is in flight BEFORE griddepcontrol.wait, the activation tiles after
    for (int stage = warp; stage < p.stages && stage < n_stage_total; stage += SG_PROD_WARPS) {
      if (elect_one()) fill(stage, stage, true, true, false);
      __syncwarp();
    }
    pdl_wait();
    for (int stage = warp; stage < p.stages && stage < n_stage_total; stage += SG_PROD_WARPS) {
      if (elect_one()) fill(stage, stage, false, false, true);
      __syncwarp();
    }
    for (int round = 1; round * p.stages < n_stage_total; ++round) {
      const uint32_t phase = (uint32_t)(round & 1);
      for (int stage = warp; stage < p.stages; stage += SG_PROD_WARPS) {
        const int i = round * p.stages + stage;
        if (i >= n_stage_total) break;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) fill(i, stage, true, true, true);
        __syncwarp();
      }
    }
    pdl_trigger_stream_tail();
  } else if (warp == SG_MMA_WARP) {
    // =========================== MMA issuer (leader CTA only) ===========================
    pdl_wait();
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * SG_BM, p.F);
      const int half_rows = p.F / 2;
      int stage = 0, acc_stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        mbar_wait(bar_tempty + 8 * acc_stage, acc_phase ^ 1);  // both epilogues have drained this accumulator set
        tc_fence_after();
        const int kb_slice = p.KB / p.S;
        const int n_stage_kb = p.KB / p.bks;
        const uint32_t sub = (uint32_t)(kXBytes + p.w_half);
        for (int sk = 0; sk < n_stage_kb; ++sk) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t base = ring + stage * p.stage_bytes;
            for (int b = 0; b < p.bks; ++b) {
              const int kb = sk * p.bks + b;
              const int slice = kb / kb_slice;
              const bool first = kb - slice * kb_slice == 0;      // first k-block of its K slice: overwrite
              const uint32_t d0 = tmem_base + (acc_stage * p.S + slice) * p.Fw;
              const uint64_t a_desc = make_smem_desc(base + b * sub);
              const uint64_t b_desc = make_smem_desc(base + b * sub + kXBytes);
#pragma unroll
              for (int k = 0; k < SG_BK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
                umma_bf16_pair(d0, a_desc + 2 * k, b_desc + 2 * k, idesc, (!first || k > 0) ? 1u : 0u);
              if (swiglu) {
                const uint64_t u_desc = make_smem_desc(base + b * sub + kXBytes + half_rows * 128);
#pragma unroll
                for (int k = 0; k < SG_BK / 16; ++k)
                  umma_bf16_pair(d0 + p.F, a_desc + 2 * k, u_desc + 2 * k, idesc, (!first || k > 0) ? 1u : 0u);
              }
            }
            umma_commit_pair(bar_empty + 8 * stage);
            if (sk == n_stage_kb - 1) umma_commit_pair(bar_tfull + 8 * acc_stage);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =========================== epilogue (warps 5..12 of both CTAs) ===========================
    pdl_wait();   // (reads the residual, writes y)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int g = (warp - SG_EPI_WARP0) >> 2;     // warp group 0 / 1: the groups alternate over the 16-feature chunks
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks = p.F / 16;
    const int m = rank * SG_BM + q * 32 + lane;   // token row of this thread
    const uint32_t leader_tempty = mapa_cluster(bar_tempty, 0);
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = pair + t * p.G;
      mbar_wait(bar_tfull + 8 * acc_stage, acc_phase);
      tc_fence_after();
      const uint32_t t0 = lane_addr + acc_stage * p.S * p.Fw;
      for (int c = g; c < n_chunks; c += 2) {
        float v0[16], v1[16];
        tmem_ld16(t0 + c * 16, v0);
        if (swiglu) tmem_ld16(t0 + p.F + c * 16, v1);
        for (int sl = 1; sl < p.S; ++sl) {   // (never SwiGLU) the further K slices, added in slice order
          tmem_ld16(t0 + sl * p.Fw + c * 16, v1);
#pragma unroll
          for (int e = 0; e < 16; ++e) v0[e] += v1[e];
        }
        const int f = tile * p.F + c * 16;
        if (m < p.M && f < p.N) pg_store16(p, m, f, v0, swiglu ? v1 : v0);
      }
      tc_fence_before();
      mbar_arrive_cluster(leader_tempty + 8 * acc_stage);
      if (++acc_stage == p.acc_stages) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();      // nobody still reads the peer's shared memory / TMEM
  if (warp == SG_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

struct PgPlan {
  int F, Fw, T, G, w_half, bks, stage_bytes, stages, acc_stages, S;
  size_t smem;
  bool ok;
};
static PgPlan pg_plan(int M, int N, int K, int epi) {
  PgPlan s{};
  const bool swiglu = epi == DD_EPI_SWIGLU;
  const int pairs = sg_num_sms() / 2 < PG_PAIRS_MAX ? sg_num_sms() / 2 : PG_PAIRS_MAX;
  // tile width: the fewest rounds r of `pairs` tiles whose tiles are at most 256 features wide (an MMA's N), and
  // narrow enough that the accumulators fit the TMEM (SwiGLU: gate + up)
  const int fmax = 256;
  int r = 1, F = 0;
  for (;; ++r) {
    F = (ceil_div(N, pairs * r) + 15) / 16 * 16;
    if (F <= fmax && (swiglu ? 2 * F : F) <= 512) break;
  }
  if (F < 32) F = 32;
  s.F = F;
  s.Fw = swiglu ? 2 * F : F;
  s.T = ceil_div(N, F);
  s.G = s.T < pairs ? s.T : pairs;
  s.w_half = (s.Fw / 2) * 128;
  s.S = sg_plan(M, N, K, epi).S;                   // the K slicing of the single-CTA kernel for this (N, K)
  const int fixed = 1024 + (2 * SG_MAX_STAGES + 4) * 8 + 16;
  const int kb_bytes = SG_BM * 128 + s.w_half;     // (a multiple of 1024: F % 16 == 0)
  // two k-blocks per stage when that still leaves a ring of >= 3 stages: the MMA issuer pays one barrier round trip
  // (~300 cycles) per stage, as much as the 4 MMAs of a narrow k-block take
  s.bks = 1;
  if ((K / SG_BK / (s.S > 0 ? s.S : 1)) % 2 == 0 && (SG_SMEM_LIMIT - fixed) / (2 * kb_bytes) >= 3) s.bks = 2;
  s.stage_bytes = s.bks * kb_bytes;
  int st = (SG_SMEM_LIMIT - fixed) / s.stage_bytes;
  s.stages = st > SG_MAX_STAGES ? SG_MAX_STAGES : st;
  s.acc_stages = (s.T > s.G && 2 * s.S * s.Fw <= 512) ? 2 : 1;
  s.smem = (size_t)fixed + (size_t)s.stages * s.stage_bytes;
  // when M needs the second accumulator and the shape is wide: the 4096-wide o_proj / down tiles would be 64 features
  // (x dominates the ingest, N=64 MMAs are issue-bound) -- measured 22.1 / 60.7 us against 22.9 / 36.5 us on the cluster
  // split-K path, where they stay
  s.ok = M > SG_BM && N >= 8192 && s.stages >= 3 && (s.F / 2) % 8 == 0 && s.S * s.Fw <= 512 && (K / SG_BK) % s.S == 0;
  return s;
}

static int sg_check(const char* fn, int M, int N, int K, int epi) {
  DD_CHECK_ARG(M >= 1 && M <= 256, "%s: M=%d out of [1,256]", fn, M);
  DD_CHECK_ARG(K >= SG_BK && K % SG_BK == 0, "%s: K=%d must be a positive multiple of %d", fn, K, SG_BK);
  DD_CHECK_ARG(N >= 8 && N % 8 == 0, "%s: N=%d must be a positive multiple of 8", fn, N);
  DD_CHECK_ARG(epi >= DD_EPI_NONE && epi <= DD_EPI_SWIGLU, "%s: bad epilogue %d", fn, epi);
  return DD_OK;
}

}  // namespace dd

#ifdef DD_SG_EXPERIMENT
extern "C" int dd_stream_gemm_debug_ctas(int M, int N, int K, int epilogue) { return dd::sg_plan(M, N, K, epilogue).G; }
#endif

extern "C" size_t dd_stream_gemm_workspace_bytes(int M, int N, int K, int epilogue) {
  using namespace dd;
  if (sg_check("dd_stream_gemm_workspace_bytes", M, N, K, epilogue)) return 0;
  return SG_WS_BYTES;
}

extern "C" int dd_stream_gemm(const void* x, const void* w, const void* residual, int M, int N, int K, int epilogue,
                              void* y, void* workspace, size_t workspace_bytes, dd_stream_t stream) {
  using namespace dd;
  if (M == 0) return DD_OK;
  if (int rc = sg_check("dd_stream_gemm", M, N, K, epilogue)) return rc;
  DD_CHECK_ARG(x && w && y && workspace, "dd_stream_gemm: null pointer");
  DD_CHECK_ARG(epilogue != DD_EPI_RESIDUAL || residual, "dd_stream_gemm: residual epilogue needs residual");
  DD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                   ((uintptr_t)residual & 15) == 0,
               "dd_stream_gemm: operands must be 16-byte aligned");
  if (workspace_bytes < SG_WS_BYTES) {
    set_error("dd_stream_gemm: workspace %zu < %zu bytes", workspace_bytes, SG_WS_BYTES);
    return DD_ERR_WORKSPACE;
  }
  const int w_rows = epilogue == DD_EPI_SWIGLU ? 2 * N : N;
  {
    static const bool pair_on = [] { const char* e = getenv("DD_PAIR_GEMM"); return !(e && e[0] == '0'); }();
    const PgPlan g = pg_plan(M, N, K, epilogue);
    if (pair_on && g.ok) {
      CUtensorMap tw, tx;
      {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows};
        const uint64_t strides[1] = {(uint64_t)K * 2};
        const uint32_t box[2] = {SG_BK, (uint32_t)(g.F / 2)};
        if (int rc = make_tmap_bf16_sw128(&tw, w, 2, dims, strides, box, "dd_stream_gemm(w, pair)")) return rc;
      }
      {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        const uint64_t strides[1] = {(uint64_t)K * 2};
        const uint32_t box[2] = {SG_BK, (uint32_t)SG_BM};
        if (int rc = make_tmap_bf16_sw128(&tx, x, 2, dims, strides, box, "dd_stream_gemm(x, pair)")) return rc;
      }
      PgParams p{};
      p.M = M; p.N = N; p.K = K; p.epi = epilogue; p.F = g.F; p.Fw = g.Fw; p.T = g.T; p.KB = K / SG_BK; p.G = g.G;
      p.w_half = g.w_half; p.bks = g.bks; p.stage_bytes = g.stage_bytes; p.stages = g.stages; p.acc_stages = g.acc_stages; p.S = g.S;
      p.residual = static_cast<const __nv_bfloat16*>(residual);
      p.y = y;
      {
        static bool attr_set[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
          DD_CHECK_CUDA(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_LIMIT));
          if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * g.G);
      cfg.blockDim = dim3(SG_THREADS);
      cfg.dynamicSmemBytes = g.smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled() ? 2 : 1;
      DD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pair_gemm_kernel, tw, tx, p));
      return DD_OK;
    }
  }
  const SgPlan s = sg_plan(M, N, K, epilogue);
  DD_CHECK_ARG(s.stages >= 2, "dd_stream_gemm: M=%d, N=%d leave no room for a 2-stage ring", M, N);
  CUtensorMap tw, tx;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {SG_BK, (uint32_t)s.F};
    if (int rc = make_tmap_bf16_sw128(&tw, w, 2, dims, strides, box, "dd_stream_gemm(w)")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {SG_BK, (uint32_t)s.Mp};
    if (int rc = make_tmap_bf16_sw128(&tx, x, 2, dims, strides, box, "dd_stream_gemm(x)")) return rc;
  }
  SgParams p{};
  p.M = M; p.Mp = s.Mp; p.N = N; p.K = K; p.epi = epilogue; p.F = s.F; p.Fw = s.Fw; p.nacc = s.nacc;
  p.bks = s.bks; p.x_kb = s.x_kb; p.x_bytes = s.x_bytes; p.w_kb = s.w_kb; p.stage_bytes = s.stage_bytes;
  p.S = s.S; p.max_own = s.max_own; p.ring_bytes = (int)(s.smem - (1024 + (2 * SG_MAX_STAGES + 4) * 8 + 16));
  p.T = s.T; p.KB = s.KB; p.G = s.G; p.stages = s.stages; p.acc_stages = s.acc_stages; p.tmem_cols = s.tmem_cols;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = y;
#ifdef DD_SG_EXPERIMENT
  p.dbg = (getenv("DD_SG_DBG") && workspace_bytes >= (size_t)s.G * 128 && s.G * 128 <= 16384) ? static_cast<long long*>(workspace) : nullptr;
  if (getenv("DD_SG_DBG")) fprintf(stderr, "[sg] M=%d N=%d K=%d epi=%d: F=%d T=%d S=%d bks=%d stage=%d B x %d stages, tmem %d\n", M, N, K, epilogue, s.F, s.T, s.S, s.bks, s.stage_bytes, s.stages, s.tmem_cols);
#endif
  {
    // the opt-in for > 48 KB of dynamic shared memory is a per-device attribute
    static bool attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      DD_CHECK_CUDA(cudaFuncSetAttribute(stream_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_LIMIT));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(s.G * s.S);
  cfg.blockDim = dim3(SG_THREADS);
  cfg.dynamicSmemBytes = s.smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (s.S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = s.S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stream_gemm_kernel, tw, tx, p));
  return DD_OK;
}
