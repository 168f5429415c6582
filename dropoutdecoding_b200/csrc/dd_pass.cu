// dd_decoder_pass: ONE persistent kernel for a whole single-token decoder pass of <= 16 rows
// (embedding gather, L x {RMSNorm, QKV + RoPE, masked decode attention, o_proj, RMSNorm, gate/up + SwiGLU, down},
// final RMSNorm, lm_head) -- the forward the reference runs 1 + K times per token through
// `self.language_model(...)` (models/llava.py:294-303, 350-359; models/llama.py:27-47).
//
// Why: with 4-7 rows the pass is a 13 GB weight stream (2.07 ms at the measured HBM peak) that the launch chain cuts
// into ~230 kernels; every boundary drains the stream (launch gap, pipeline ramp, tail), and the latency-bound
// kernels between the GEMMs (attention, RMSNorm) leave HBM idle: 3.4 ms per pass.  Here one CTA of 8 warps per SM stays
// resident for the pass (cooperative launch):
//   * phases are separated by a grid barrier (one release atomic + one acquire spin per CTA, bounded) instead of a
//     kernel boundary;
//   * a CTA owns 16 output features (16 + 16 for SwiGLU) at a time and runs the arithmetic of dd_skinny_gemm (same k
//     split over the 8 warps, same MMA order, same cross-warp sum, same epilogue rounding points), so the pass is
//     bit-identical to the launch chain; the flat (tile, batch) sequence of a phase is software-pipelined over two
//     register batches per thread, so a tile's reduction / epilogue runs under the next tile's loads;
//   * the RMSNorms disappear into the GEMM that consumes them: every CTA normalises the (<= 16) rows once into shared
//     memory (rmsnorm_kernel's arithmetic, all rows in one L2 round trip) and the B fragments of all its tiles come
//     from there;
//   * attention: warps 0-3 of a CTA run the work items of dd_attention.cu's body (split x kv head x sequence x member
//     window; the chain pass of the lookahead schedule has two windows with different lengths);
//   * the lookahead's "guessed winner" K/V row goes to its cache row from the QKV epilogue (no commit launch).
// Algorithmic bytes per pass: the decoder weights once + the K/V prefix once per attention window.
// STATUS (profiles/README.md 0): correct and bit-identical to the chain (tests), but measured SLOWER --
// 4.8 vs 3.2 ms per 7B pass: 128 KB of registers in flight per SM bound the steady state at 4.5-5.5 TB/s and every phase
// opens with a 13-16 us bubble (row staging + the first tile's latency with all SMs starting cold at once).  Bulk L2
// prefetches (kept behind DD_PASS_TILES_AHEAD / DD_PASS_TILES_PRE / prefetch_gate_up_bytes, all 0) and cp.async
// shared-memory slots made it slower still.  The engine uses it only with DD_PASS=1.
#include <cstring>

#include "dd_attention_body.cuh"

namespace dd {

constexpr int kPassGroups = 1;          // (one 8-warp group per CTA: 255 registers per thread, two weight batches in flight)
constexpr int kGroupWarps = 8;
constexpr int kGroupThreads = kGroupWarps * 32;
constexpr int kPassThreads = kPassGroups * kGroupThreads;
constexpr int kXPad = 32;          // bf16 elements (64 B) appended to a staged row: conflict-free 128-bit reads
constexpr int kWaRegs = 16;        // uint4 registers of one weight batch (kUnroll x NW x 2); two batches are in flight

// one decoder layer of the plan (device memory, filled on the host by dd_pass_plan_fill)
struct alignas(128) PassLayer {
  CUtensorMap tmap_k, tmap_v;                        // this layer's caches as [B*H_kv][S_max][D]
  const __nv_bfloat16 *ln1, *wqkv, *wo, *ln2, *wgu, *wd;
  __nv_bfloat16 *k_new, *v_new;                      // staging rows of this layer [rows, H_kv, D]
  __nv_bfloat16 *k_cache, *v_cache;                  // [B, H_kv, S_max, D]
};
static_assert(sizeof(PassLayer) == 384, "PassLayer layout (dd_pass_plan_bytes)");

struct PassParams {
  const PassLayer* layers;
  int L, rows, d, dff, H, H_kv, S_max, Vp, vocab;
  float eps;
  const __nv_bfloat16* embed;
  const int32_t *tok, *pos;            // row r: token tok[r / tok_div], RoPE position pos[r / pos_div]
  int tok_div, pos_div;
  const float *cos_tab, *sin_tab;
  __nv_bfloat16 *x, *x2, *act, *q, *attn, *xn_out;
  const __nv_bfloat16 *final_norm, *lm_head;
  float* logits;
  AttnParams win[2];                   // member windows (k/v pointers of the layer are filled in per layer)
  int n_windows;
  const int32_t *guess_winner, *guess_pos;   // chain pass: row b*Mt + guess_winner[b] is also written to cache row guess_pos[b]
  unsigned* sync;                      // {barrier counter, exit counter}: zero on entry, zero again on exit
  unsigned long long pf_gu_bytes;      // bytes of gate/up prefetched to L2 during the attention phase (after o_proj's)
  int xs_bytes, red_bytes;             // shared-memory layout
  unsigned long long* timeline;        // or NULL: CTA 0 stamps %globaltimer at {start, end of its work} of every phase
};

// ---------------------------------------------------------------------------------------------
// grid barrier (all CTAs co-resident: cooperative launch).  arrive = release, wait = acquire; bounded spin.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stamp(unsigned long long* timeline, unsigned slot) {
  if (timeline && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    timeline[slot] = t;
  }
}
__device__ __forceinline__ void grid_arrive(unsigned* ctr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (guess row) before later TMA reads
    __threadfence();
    atomicAdd(ctr, 1u);
  }
}
__device__ __forceinline__ void grid_wait(unsigned* ctr, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned v = 0;
    for (unsigned spin = 0;; ++spin) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (spin > (1u << 24)) __trap();   // a protocol bug must surface as a trapped kernel, never as a hung GPU
      __nanosleep(32);
    }
    __threadfence();                     // (invalidates this SM's L1: plain loads below see the other CTAs' stores)
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// GEMM tiles (the arithmetic of dd_skinny_gemm.cu; see its header for the fragment trick)
// ---------------------------------------------------------------------------------------------
template <int EPI>
struct TileShape {
  static constexpr int NW = (EPI == DD_EPI_SWIGLU) ? 2 : 1;
  static constexpr int kUnroll = kWaRegs / (2 * NW);    // 8 k-steps of 32 in flight (4 for the two SwiGLU row tiles)
};

// the two weight rows (g, g + 8 of the tile; RoPE: a rotate-half pair) this lane streams, per row tile
template <int EPI>
__device__ __forceinline__ void tile_rows(const __nv_bfloat16* (&wrow)[TileShape<EPI>::NW][2], const __nv_bfloat16* w,
                                          int tile, int N, int K, int kbeg, int g, int t, int H, int H_kv, int D) {
  const int n0 = tile * 16;
#pragma unroll
  for (int i = 0; i < TileShape<EPI>::NW; ++i) {
    size_t r0 = (size_t)(i * N + n0 + g), r1 = r0 + 8;
    if (EPI == DD_EPI_ROPE && n0 < (H + H_kv) * D) {
      const int tph = D / 16, head = tile / tph, jj = tile % tph;
      r0 = (size_t)head * D + 8 * jj + g;
      r1 = r0 + D / 2;
    }
    wrow[i][0] = w + r0 * K + kbeg + 8 * t;
    wrow[i][1] = w + r1 * K + kbeg + 8 * t;
  }
}

// One batch of kUnroll k-steps into registers.  Loads are unconditional (a k-step past the end of the slice re-reads
// the last one: an L2 hit whose MMAs are skipped) -- predicated partial updates of the batch registers make the
// register allocator keep the whole batch in local memory.
template <int EPI>
__device__ __forceinline__ void load_w(uint4 (&wa)[kWaRegs], const __nv_bfloat16* const (&wrow)[TileShape<EPI>::NW][2],
                                       int s0, int steps) {
  constexpr int NW = TileShape<EPI>::NW, kU = TileShape<EPI>::kUnroll;
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int ko = min(s0 + u, steps - 1) * 32;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      wa[(u * NW + i) * 2] = ld_stream_128(wrow[i][0] + ko);
      wa[(u * NW + i) * 2 + 1] = ld_stream_128(wrow[i][1] + ko);
    }
  }
}

// L2 prefetch of the weight rows of tile `tile` (2 * NW blocks of 8 contiguous rows), spread over the CTA's 8 warps.
// Used (a) before a grid barrier for the first tiles of the NEXT phase: the weights never depend on the phase before,
// so the HBM stream restarts while the barrier resolves and the phase starts on L2 hits; (b) inside a phase, two
// tiles ahead of the register pipeline, which keeps ~40 MB in flight chip-wide without holding registers.
template <int EPI>
__device__ __forceinline__ void tile_prefetch(const __nv_bfloat16* w, int N, int K, const PassParams& p, int D, int tile,
                                              int warp) {
  if (tile >= N / 16) return;
  constexpr int NW = TileShape<EPI>::NW;
  const __nv_bfloat16* wrow[NW][2];
  tile_rows<EPI>(wrow, w, tile, N, K, 0, 0, 0, p.H, p.H_kv, D);     // g = t = 0: the first row of each block
  L2Prefetch pf{};
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    pf.ptr[0] = reinterpret_cast<const char*>(wrow[i][0]); pf.bytes[0] = (unsigned long long)8 * K * 2;
    pf.ptr[1] = reinterpret_cast<const char*>(wrow[i][1]); pf.bytes[1] = (unsigned long long)8 * K * 2;
    l2_prefetch_slice(pf, warp, kGroupWarps);
  }
}
#ifndef DD_PASS_TILES_AHEAD
#define DD_PASS_TILES_AHEAD 0
#endif
#ifndef DD_PASS_TILES_PRE
#define DD_PASS_TILES_PRE 0
#endif
constexpr int kTilesAhead = DD_PASS_TILES_AHEAD;   // L2 prefetch distance inside a phase, in tiles of this CTA (0: off)
constexpr int kTilesPre = DD_PASS_TILES_PRE;       // tiles of the next phase prefetched before the grid barrier
// before a grid barrier: the first tiles of this CTA in the next phase
template <int EPI>
__device__ __forceinline__ void gemm_prefetch(const __nv_bfloat16* w, int N, int K, const PassParams& p, int D, int cta,
                                              int warp) {
#pragma unroll
  for (int a = 0; a < kTilesPre; ++a) tile_prefetch<EPI>(w, N, K, p, D, cta + a * (int)gridDim.x, warp);
}

// RMSNorm of the pass's rows into shared memory with rmsnorm_kernel's arithmetic: thread tid owns the 16-byte pieces
// tid, tid + 256, ... of EVERY row (same per-thread partial sums, same shuffle tree, same order over the warps as one
// rmsnorm CTA per row), eight rows at a time so that all their loads are one L2 round trip.  nw == NULL: plain copy.
__device__ __forceinline__ void stage_rows(const __nv_bfloat16* x, const __nv_bfloat16* nw, int rows, int d, float eps,
                                           __nv_bfloat16* xs, float* s_part, int tid) {
  const int nvec = d / 8, ld = d + kXPad;
  const uint4* w4 = reinterpret_cast<const uint4*>(nw);
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float ss[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) ss[r] = 0.f;
    if (nw) {
      for (int j = tid; j < nvec; j += kGroupThreads) {
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r0 + r < rows) v[r] = reinterpret_cast<const uint4*>(x + (size_t)(r0 + r) * d)[j];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r0 + r < rows) {
            const uint32_t u[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = bf16x2_to_float2(u[k]);
              ss[r] = fmaf(f.x, f.x, ss[r]);
              ss[r] = fmaf(f.y, f.y, ss[r]);
            }
          }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float tot = warp_sum(ss[r]);
        if ((tid & 31) == 0) s_part[r * kGroupWarps + (tid >> 5)] = tot;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < kGroupWarps; ++k) tot += s_part[r * kGroupWarps + k];
        ss[r] = rsqrtf(tot / (float)d + eps);     // 1/rms of row r0 + r
      }
    }
    for (int j = tid; j < nvec; j += kGroupThreads) {
      uint4 gw4 = make_uint4(0, 0, 0, 0);
      if (nw) gw4 = w4[j];
      const uint32_t gw[4] = {gw4.x, gw4.y, gw4.z, gw4.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < rows) {
          uint4 v = reinterpret_cast<const uint4*>(x + (size_t)(r0 + r) * d)[j];
          if (nw) {
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = bf16x2_to_float2(u[k]), ww = bf16x2_to_float2(gw[k]);
              o[k] = pack_bf16x2(bf16_round(f.x * ss[r]) * ww.x, bf16_round(f.y * ss[r]) * ww.y);
            }
            v = make_uint4(o[0], o[1], o[2], o[3]);
          }
          *reinterpret_cast<uint4*>(xs + (size_t)(r0 + r) * ld + j * 8) = v;
        }
    }
    __syncthreads();   // s_part is reused by the next eight rows; the staged rows are complete
  }
}

struct GemmIO {
  const __nv_bfloat16* w;
  int N, K;
  const __nv_bfloat16* x_global;     // activations in global memory [rows, K] (phases without a norm)
  const __nv_bfloat16* residual;
  void* y;
  // RoPE epilogue
  __nv_bfloat16 *k_new, *v_new, *k_cache, *v_cache;
};

// One GEMM phase for this CTA: its 8 warps walk the tiles cta, cta + G, ...; every warp streams its k-slice of a tile in
// batches of kUnroll k-steps.  The flat sequence of (tile, batch) pairs is software-pipelined over two register
// buffers: while the MMAs of one batch run, the next batch -- of this tile or already of the next one -- is in flight,
// so a warp always has 8-16 KB of weights outstanding and the reduction / epilogue of a tile hides under the next
// tile's loads.  kStaged: activations come from the shared-memory rows `xs` (else from global memory).
// Not inlined: ptxas allocates a phase on its own without spills (166-194 registers); inlined into the whole pass it
// keeps the weight batches in local memory.
template <int MT, int EPI, bool kStaged, int D>
__device__ __forceinline__ void gemm_phase(const GemmIO& io, const PassParams& p, const __nv_bfloat16* xs, float* red,
                                           int tid, int cta, int n_ctas, unsigned long long* trace = nullptr) {
#ifdef DD_PASS_TRACE
  int trace_n = 0;
  auto tr = [&]() { if (trace && trace_n < 16) stamp(trace, trace_n++); };
#else
  auto tr = [&]() {};
#endif
  tr();
  constexpr int NW = TileShape<EPI>::NW, kU = TileShape<EPI>::kUnroll;
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int M = p.rows, N = io.N, K = io.K;
  const int kslice = K / kGroupWarps, kbeg = warp * kslice, steps = kslice / 32;
  const int ntiles = N / 16;
  const int ldx = kStaged ? K + kXPad : K;
  const __nv_bfloat16* xrow[MT];
  uint32_t xrow_s[MT];
  bool xok[MT];
#pragma unroll
  for (int j = 0; j < MT; ++j) {
    const int m = 8 * j + g;
    xok[j] = m < M;
    xrow[j] = io.x_global + (size_t)(xok[j] ? m : 0) * ldx + kbeg + 8 * t;
    xrow_s[j] = smem_u32(xs + (size_t)(xok[j] ? m : 0) * ldx + kbeg + 8 * t);
  }
  auto red_at = [&](int wp, int i, int j, int r, int tc) -> float& {   // s_red[warp][NW][MT][16][9]
    return red[((((size_t)wp * NW + i) * MT + j) * 16 + r) * 9 + tc];
  };
  float acc[NW][MT][4];
#pragma unroll
  for (int i = 0; i < NW; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  int ld_tile = cta, ld_s0 = 0;    // the next batch to put in flight
  int cp_tile = cta, cp_s0 = 0;    // the next batch to consume
  auto issue = [&](uint4 (&buf)[kWaRegs]) {
    if (ld_tile >= ntiles) return;
    const __nv_bfloat16* wrow[NW][2];
    tile_rows<EPI>(wrow, io.w, ld_tile, N, K, kbeg, g, t, p.H, p.H_kv, D);
    if (kTilesAhead > 0 && ld_s0 == 0) tile_prefetch<EPI>(io.w, N, K, p, D, ld_tile + kTilesAhead * n_ctas, warp);
    load_w<EPI>(buf, wrow, ld_s0, steps);
    ld_s0 += kU;
    if (ld_s0 >= steps) { ld_s0 = 0; ld_tile += n_ctas; }
  };
  auto finish_tile = [&](int tile) {
    tr();
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        red_at(warp, i, j, g, 2 * t) = acc[i][j][0];
        red_at(warp, i, j, g, 2 * t + 1) = acc[i][j][1];
        red_at(warp, i, j, g + 8, 2 * t) = acc[i][j][2];
        red_at(warp, i, j, g + 8, 2 * t + 1) = acc[i][j][3];
      }
    __syncthreads();
    const int n0 = tile * 16;
    if constexpr (EPI == DD_EPI_ROPE) {
      auto rsum = [&](int j, int r, int tc) {
        float s = 0.f;
#pragma unroll
        for (int wp = 0; wp < kGroupWarps; ++wp) s += red_at(wp, 0, j, r, tc);
        return bf16_round(s);  // the projection's own bf16 output
      };
      const int Mt = p.win[0].Mt;
      if (n0 < (p.H + p.H_kv) * D) {
        const int tph = D / 16, head = tile / tph, jj = tile % tph, half = D / 2;
        for (int o = tid; o < MT * 8 * 8; o += kGroupThreads) {
          const int r = o & 7, tok = o >> 3;
          if (tok >= M) continue;
          const int j = tok >> 3, tc = tok & 7, i = 8 * jj + r;
          const float x1 = rsum(j, r, tc), x2 = rsum(j, r + 8, tc);
          const int pos = p.pos[tok / p.pos_div];
          const float c = p.cos_tab[(size_t)pos * half + i], sn = p.sin_tab[(size_t)pos * half + i];
          const __nv_bfloat16 lo = __float2bfloat16_rn(x1 * c - x2 * sn), hi = __float2bfloat16_rn(x2 * c + x1 * sn);
          if (head < p.H) {
            __nv_bfloat16* dst = p.q + ((size_t)tok * p.H + head) * D;
            dst[i] = lo;
            dst[i + half] = hi;
          } else {
            const int hk = head - p.H;
            __nv_bfloat16* dst = io.k_new + ((size_t)tok * p.H_kv + hk) * D;
            dst[i] = lo;
            dst[i + half] = hi;
            if (p.guess_winner) {
              const int b = tok / Mt, gw = p.guess_winner[b], gp = p.guess_pos[b];
              if (tok - b * Mt == gw && gp < p.S_max) {
                __nv_bfloat16* c2 = io.k_cache + (((size_t)b * p.H_kv + hk) * p.S_max + gp) * D;
                c2[i] = lo;
                c2[i + half] = hi;
              }
            }
          }
        }
      } else {
        for (int o = tid; o < MT * 8 * 16; o += kGroupThreads) {
          const int r = o & 15, tok = o >> 4;
          if (tok >= M) continue;
          const int n = n0 + r - (p.H + p.H_kv) * D;  // feature inside the v block
          const __nv_bfloat16 v = __float2bfloat16_rn(rsum(tok >> 3, r, tok & 7));
          io.v_new[(size_t)tok * p.H_kv * D + n] = v;
          if (p.guess_winner) {
            const int b = tok / Mt, gw = p.guess_winner[b], gp = p.guess_pos[b];
            if (tok - b * Mt == gw && gp < p.S_max)
              io.v_cache[(((size_t)b * p.H_kv + n / D) * p.S_max + gp) * D + n % D] = v;
          }
        }
      }
    } else {
      for (int o = tid; o < MT * 8 * 16; o += kGroupThreads) {
        const int r = o & 15, tok = o >> 4;  // feature fastest: coalesced 32B stores per token
        if (tok >= M) continue;
        const int j = tok >> 3, tc = tok & 7;
        float v[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          float s = 0.f;
#pragma unroll
          for (int wp = 0; wp < kGroupWarps; ++wp) s += red_at(wp, i, j, r, tc);
          v[i] = bf16_round(s);
        }
        const size_t oidx = (size_t)tok * N + n0 + r;
        if constexpr (EPI == DD_EPI_NONE) {
          reinterpret_cast<__nv_bfloat16*>(io.y)[oidx] = __float2bfloat16_rn(v[0]);
        } else if constexpr (EPI == DD_EPI_RESIDUAL) {
          reinterpret_cast<__nv_bfloat16*>(io.y)[oidx] =
              __float2bfloat16_rn(v[0] + __bfloat162float(io.residual[oidx]));
        } else if constexpr (EPI == DD_EPI_F32) {
          reinterpret_cast<float*>(io.y)[oidx] = v[0];
        } else {  // SwiGLU: bf16( bf16(silu(g)) * u )
          const float gte = v[0];
          const float act = bf16_round(gte / (1.f + __expf(-gte)));
          reinterpret_cast<__nv_bfloat16*>(io.y)[oidx] = __float2bfloat16_rn(act * v[NW - 1]);
        }
      }
    }
    __syncthreads();   // the next tile overwrites s_red
    tr();
  };
  // consume one batch from `buf`; false when this CTA has no tile left
  auto consume = [&](uint4 (&buf)[kWaRegs]) -> bool {
    if (cp_tile >= ntiles) return false;
    if (cp_s0 == 0) {
#pragma unroll
      for (int i = 0; i < NW; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (cp_s0 + u < steps) {
        const int ko = (cp_s0 + u) * 32;
        uint4 xv[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          if (!xok[j]) {
            xv[j] = make_uint4(0, 0, 0, 0);
          } else if (kStaged) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(xv[j].x), "=r"(xv[j].y), "=r"(xv[j].z), "=r"(xv[j].w)
                         : "r"(xrow_s[j] + ko * 2));
          } else {
            xv[j] = *reinterpret_cast<const uint4*>(xrow[j] + ko);
          }
        }
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const uint4 a0 = buf[(u * NW + i) * 2], a1 = buf[(u * NW + i) * 2 + 1];
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            mma_bf16_16816(acc[i][j], a0.x, a1.x, a0.y, a1.y, xv[j].x, xv[j].y);
            mma_bf16_16816(acc[i][j], a0.z, a1.z, a0.w, a1.w, xv[j].z, xv[j].w);
          }
        }
      }
    }
    issue(buf);      // its registers are free: the batch after next flies under the reduction / epilogue below
    const int tile = cp_tile;
    cp_s0 += kU;
    const bool done = cp_s0 >= steps;
    if (done) { cp_s0 = 0; cp_tile += n_ctas; }
    if (done) finish_tile(tile);
    return true;
  };

  // (zero-initialised: a register that may be read before it is written on SOME path -- the batch of an exhausted
  // tile list -- is live from the kernel's entry for ptxas, and six phases' worth of those leave no registers)
  uint4 wa0[kWaRegs], wa1[kWaRegs];
#pragma unroll
  for (int i = 0; i < kWaRegs; ++i) wa0[i] = wa1[i] = make_uint4(0, 0, 0, 0);
  issue(wa0);
  issue(wa1);
  for (;;) {
    if (!consume(wa0)) break;
    if (!consume(wa1)) break;
  }
}

// The attention phase of one layer for warps 0-3 of a CTA.
template <int D>
__device__ __forceinline__ void attention_phase(const PassParams& p, const PassLayer& ly, uint8_t* smem, uint64_t* s_full,
                                             int* s_ext, int* s_last) {
  const int tid = threadIdx.x, G = gridDim.x;
  int base = 0;
  for (int wi = 0; wi < p.n_windows; ++wi) {
    AttnParams ap = p.win[wi];
    ap.k_cache = ly.k_cache; ap.v_cache = ly.v_cache; ap.k_new = ly.k_new; ap.v_new = ly.v_new;
    const int items = ap.B * ap.H_kv * ap.num_splits;
    // items of window wi are dealt to CTAs base, base + 1, ... (mod G), so two windows do not pile up on CTA 0..
    for (int it = ((int)blockIdx.x - base % G + G) % G; it < items; it += G) {
      const int split = it % ap.num_splits, hkv = (it / ap.num_splits) % ap.H_kv, b = it / (ap.num_splits * ap.H_kv);
      attn_partial_body<D, 1, false>(&ly.tmap_k, &ly.tmap_v, ap, L2Prefetch{}, split, hkv, b, tid, 0u, 1u, smem, s_full,
                                     s_ext, s_last,
                                     [] { asm volatile("bar.sync 1, %0;" ::"n"(kAttnThreads) : "memory"); }, [] {});
      asm volatile("bar.sync 1, %0;" ::"n"(kAttnThreads) : "memory");   // s_last / the ring are free again
    }
    base += items;
  }
}

// ---------------------------------------------------------------------------------------------
// -DDD_PASS_TRACE: layer 1's GEMM phases also stamp {phase entry, start / end of every tile's reduction + epilogue} of
// CTA 0 behind the phase stamps of the timeline (16 slots per phase kind)
#ifdef DD_PASS_TRACE
#define TRACE_PTR(k) ((p.timeline && l == 1) ? p.timeline + 2 * (5 * p.L + 2) + 16 * (k) : nullptr)
#else
#define TRACE_PTR(k) nullptr
#endif

template <int MT, int D>
__global__ void __launch_bounds__(kPassThreads, 1) decoder_pass_kernel(const __grid_constant__ PassParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);            // staged rows; the attention ring aliases it
  float* s_red = reinterpret_cast<float*>(smem + p.xs_bytes);
  __shared__ uint64_t s_full[kStages];
  __shared__ int s_ext[DD_MAX_MEMBERS];
  __shared__ int s_last;
  __shared__ float s_part[8 * kGroupWarps];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int d = p.d, HD = p.H * D, Nqkv = (p.H + 2 * p.H_kv) * D;
  unsigned epoch = 0;
  stamp(p.timeline, 0);

  // ---- embedding gather: x[r,:] = embed[tok[r / tok_div], :]
  for (int r = cta; r < p.rows; r += G) {
    const int tk = min(max(p.tok[r / p.tok_div], 0), p.vocab - 1);
    const uint4* src = reinterpret_cast<const uint4*>(p.embed + (size_t)tk * d);
    uint4* dst = reinterpret_cast<uint4*>(p.x + (size_t)r * d);
    for (int j = tid; j < d / 8; j += kPassThreads) dst[j] = src[j];
  }
  gemm_prefetch<DD_EPI_ROPE>(p.layers[0].wqkv, Nqkv, d, p, D, cta, warp);
  stamp(p.timeline, 1);
  grid_arrive(p.sync);

  for (int l = 0; l < p.L; ++l) {
    const PassLayer& ly = p.layers[l];
    GemmIO io{};
    // ---- RMSNorm + QKV projection + RoPE + staging (+ the guessed winner's row into the cache)
    grid_wait(p.sync, ++epoch * G);
    stamp(p.timeline, 2 * epoch);
    stage_rows(p.x, ly.ln1, p.rows, d, p.eps, xs, s_part, tid);
    io.w = ly.wqkv; io.N = Nqkv; io.K = d;
    io.k_new = ly.k_new; io.v_new = ly.v_new; io.k_cache = ly.k_cache; io.v_cache = ly.v_cache;
    gemm_phase<MT, DD_EPI_ROPE, true, D>(io, p, xs, s_red, tid, cta, G, TRACE_PTR(0));
    stamp(p.timeline, 2 * epoch + 1);
    grid_arrive(p.sync);

    // ---- attention: warps 0-3 run the work items, the others keep HBM busy with the next weights
    grid_wait(p.sync, ++epoch * G);
    stamp(p.timeline, 2 * epoch);
    if (tid < kAttnThreads) {
      attention_phase<D>(p, ly, smem, s_full, s_ext, &s_last);
    } else {
      L2Prefetch pf{};
      pf.ptr[0] = reinterpret_cast<const char*>(ly.wo); pf.bytes[0] = p.pf_gu_bytes ? (unsigned long long)d * HD * 2 : 0;
      pf.ptr[1] = reinterpret_cast<const char*>(ly.wgu); pf.bytes[1] = p.pf_gu_bytes;
      constexpr int kIdleWarps = (kPassThreads - kAttnThreads) / 32;
      l2_prefetch_slice(pf, cta * kIdleWarps + (tid - kAttnThreads) / 32, G * kIdleWarps);
    }
    stamp(p.timeline, 2 * epoch + 1);
    grid_arrive(p.sync);

    // ---- o_proj + residual: x2 = x + attn . wo^T
    grid_wait(p.sync, ++epoch * G);
    stamp(p.timeline, 2 * epoch);
    stage_rows(p.attn, nullptr, p.rows, HD, 0.f, xs, s_part, tid);
    io.w = ly.wo; io.N = d; io.K = HD; io.residual = p.x; io.y = p.x2;
    gemm_phase<MT, DD_EPI_RESIDUAL, true, D>(io, p, xs, s_red, tid, cta, G, TRACE_PTR(1));
    gemm_prefetch<DD_EPI_SWIGLU>(ly.wgu, p.dff, d, p, D, cta, warp);
    stamp(p.timeline, 2 * epoch + 1);
    grid_arrive(p.sync);

    // ---- RMSNorm + gate/up + SwiGLU
    grid_wait(p.sync, ++epoch * G);
    stamp(p.timeline, 2 * epoch);
    stage_rows(p.x2, ly.ln2, p.rows, d, p.eps, xs, s_part, tid);
    io.w = ly.wgu; io.N = p.dff; io.K = d; io.y = p.act;
    gemm_phase<MT, DD_EPI_SWIGLU, true, D>(io, p, xs, s_red, tid, cta, G, TRACE_PTR(2));
    gemm_prefetch<DD_EPI_RESIDUAL>(ly.wd, d, p.dff, p, D, cta, warp);
    stamp(p.timeline, 2 * epoch + 1);
    grid_arrive(p.sync);

    // ---- down + residual: x = x2 + act . wd^T
    grid_wait(p.sync, ++epoch * G);
    stamp(p.timeline, 2 * epoch);
    io.w = ly.wd; io.N = d; io.K = p.dff; io.x_global = p.act; io.residual = p.x2; io.y = p.x;
    gemm_phase<MT, DD_EPI_RESIDUAL, false, D>(io, p, xs, s_red, tid, cta, G, TRACE_PTR(3));
    if (l + 1 < p.L)
      gemm_prefetch<DD_EPI_ROPE>(p.layers[l + 1].wqkv, Nqkv, d, p, D, cta, warp);
    else if (p.logits)
      gemm_prefetch<DD_EPI_F32>(p.lm_head, p.Vp, d, p, D, cta, warp);
    stamp(p.timeline, 2 * epoch + 1);
    grid_arrive(p.sync);
  }

  // ---- final RMSNorm (+ its hidden state for the caller) + lm_head
  grid_wait(p.sync, ++epoch * G);
  stamp(p.timeline, 2 * epoch);
  if (p.logits || p.xn_out) {
    stage_rows(p.x, p.final_norm, p.rows, d, p.eps, xs, s_part, tid);
    if (p.xn_out && cta == 0)
      for (int idx = tid; idx < p.rows * (d / 8); idx += kPassThreads) {
        const int r = idx / (d / 8), j = idx % (d / 8);
        reinterpret_cast<uint4*>(p.xn_out + (size_t)r * d)[j] =
            *reinterpret_cast<const uint4*>(xs + (size_t)r * (d + kXPad) + j * 8);
      }
  }
  if (p.logits) {
    GemmIO io{};
    io.w = p.lm_head; io.N = p.Vp; io.K = d; io.y = p.logits;
    gemm_phase<MT, DD_EPI_F32, true, D>(io, p, xs, s_red, tid, cta, G);
  }

  stamp(p.timeline, 2 * epoch + 1);
  // ---- leave the counters at zero for the next launch: the last CTA to get here knows every wait has returned
  __syncthreads();
  if (tid == 0) {
    if (atomicAdd(p.sync + 1, 1u) == (unsigned)G - 1) {
      p.sync[0] = 0;
      p.sync[1] = 0;
      __threadfence();
    }
  }
}

static int pass_grid(int* sms_out) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  *sms_out = sms;
  return dev;
}

template <int MT, int D>
static int launch_pass(const PassParams& p, size_t smem, cudaStream_t st) {
  int sms = 0;
  const int dev = pass_grid(&sms);
  if (dev < 0) {
    set_error("dd_decoder_pass: no CUDA device");
    return DD_ERR_CUDA;
  }
  {
    static size_t configured[64];
    if (dev >= 64 || smem > configured[dev]) {
      DD_CHECK_CUDA(cudaFuncSetAttribute(decoder_pass_kernel<MT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
      if (dev < 64) configured[dev] = smem;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms);       // one CTA per SM, all resident: the grid barrier needs it (cooperative launch)
  cfg.blockDim = dim3(kPassThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decoder_pass_kernel<MT, D>, p));
  return DD_OK;
}

}  // namespace dd

extern "C" size_t dd_pass_plan_bytes(int n_layers) { return (size_t)(n_layers > 0 ? n_layers : 0) * sizeof(dd::PassLayer); }

extern "C" int dd_pass_plan_fill(void* host_plan, int layer, const void* ln1, const void* wqkv, const void* wo,
                                 const void* ln2, const void* wgu, const void* wd, void* k_new, void* v_new,
                                 void* k_cache, void* v_cache, int B, int H_kv, int S_max, int D) {
  using namespace dd;
  DD_CHECK_ARG(host_plan && layer >= 0, "dd_pass_plan_fill: null plan / bad layer");
  DD_CHECK_ARG(ln1 && wqkv && wo && ln2 && wgu && wd && k_new && v_new && k_cache && v_cache,
               "dd_pass_plan_fill: null pointer");
  DD_CHECK_ARG(D == 64 || D == 128, "dd_pass_plan_fill: D=%d unsupported (64 or 128)", D);
  DD_CHECK_ARG(B >= 1 && H_kv >= 1 && S_max >= 1, "dd_pass_plan_fill: bad cache shape");
  DD_CHECK_ARG(((uintptr_t)k_cache & 15) == 0 && ((uintptr_t)v_cache & 15) == 0,
               "dd_pass_plan_fill: the caches must be 16-byte aligned (TMA)");
  PassLayer ly{};
  const uint64_t dims[3] = {(uint64_t)D, (uint64_t)S_max, (uint64_t)B * H_kv};
  const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)S_max * D * 2};
  const uint32_t box[3] = {64, (uint32_t)kTileKeys, 1};
  if (int rc = make_tmap_bf16_sw128(&ly.tmap_k, k_cache, 3, dims, strides, box, "dd_pass_plan_fill(k_cache)")) return rc;
  if (int rc = make_tmap_bf16_sw128(&ly.tmap_v, v_cache, 3, dims, strides, box, "dd_pass_plan_fill(v_cache)")) return rc;
  ly.ln1 = (const __nv_bfloat16*)ln1; ly.wqkv = (const __nv_bfloat16*)wqkv; ly.wo = (const __nv_bfloat16*)wo;
  ly.ln2 = (const __nv_bfloat16*)ln2; ly.wgu = (const __nv_bfloat16*)wgu; ly.wd = (const __nv_bfloat16*)wd;
  ly.k_new = (__nv_bfloat16*)k_new; ly.v_new = (__nv_bfloat16*)v_new;
  ly.k_cache = (__nv_bfloat16*)k_cache; ly.v_cache = (__nv_bfloat16*)v_cache;
  memcpy(reinterpret_cast<uint8_t*>(host_plan) + (size_t)layer * sizeof(PassLayer), &ly, sizeof(PassLayer));
  return DD_OK;
}

extern "C" size_t dd_pass_workspace_bytes(int B, int members_total, int H, int D, int num_splits) {
  // {barrier counter, exit counter} (256 B) + the split-KV partials and tickets of two member windows
  return 256 + 2 * ((dd_attention_workspace_bytes(B, members_total, H, D, num_splits) + 255) / 256 * 256);
}

extern "C" int dd_decoder_pass(const dd_pass_args* a, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(a && a->plan && a->n_layers >= 1, "dd_decoder_pass: null arguments / plan");
  DD_CHECK_ARG(a->rows >= 1 && a->rows <= 16 && a->members_total >= 1 && a->rows % a->members_total == 0,
               "dd_decoder_pass: rows=%d out of [1,16] or not a multiple of members_total=%d", a->rows,
               a->members_total);
  DD_CHECK_ARG(a->D == 64 || a->D == 128, "dd_decoder_pass: D=%d unsupported (64 or 128)", a->D);
  DD_CHECK_ARG(a->H_kv >= 1 && a->H % a->H_kv == 0, "dd_decoder_pass: H=%d not a multiple of H_kv=%d", a->H, a->H_kv);
  DD_CHECK_ARG(a->d_model % 256 == 0 && a->d_ff % 256 == 0 && (a->H * a->D) % 256 == 0 && a->d_ff % 16 == 0 &&
                   a->d_model % 16 == 0,
               "dd_decoder_pass: d_model=%d, d_ff=%d, H*D=%d must be multiples of 256", a->d_model, a->d_ff,
               a->H * a->D);
  DD_CHECK_ARG(!a->logits || (a->lm_head && a->vocab_padded % 16 == 0), "dd_decoder_pass: lm_head / vocab_padded");
  DD_CHECK_ARG(a->embed && a->tok && a->pos && a->tok_div >= 1 && a->pos_div >= 1 && a->cos_tab && a->sin_tab &&
                   a->x && a->x2 && a->act && a->q && a->attn && a->final_norm && a->workspace,
               "dd_decoder_pass: null pointer");
  DD_CHECK_ARG(a->n_windows >= 1 && a->n_windows <= 2, "dd_decoder_pass: n_windows=%d out of [1,2]", a->n_windows);
  DD_CHECK_ARG(a->num_splits >= 2 && a->num_splits <= 256,
               "dd_decoder_pass: num_splits=%d out of [2,256] (the unsplit merge is a separate kernel of the launch chain)",
               a->num_splits);
  DD_CHECK_ARG(!a->bitmask || (a->vis_start && a->mask_words * 32 >= a->n_vis && a->mask_words <= kMaxMaskWords),
               "dd_decoder_pass: bitmask needs vis_start and n_vis <= 32*mask_words <= %d", 32 * kMaxMaskWords);
  DD_CHECK_ARG(!a->guess_winner || a->guess_pos, "dd_decoder_pass: guess_winner needs guess_pos");
  const int B = a->rows / a->members_total, Gq = a->H / a->H_kv;
  int covered = 0, max_m = 0;
  for (int w = 0; w < a->n_windows; ++w) {
    DD_CHECK_ARG(a->win_members[w] >= 1 && a->win_first[w] == covered && a->win_seq_lens[w],
                 "dd_decoder_pass: member windows must tile [0, members_total) in order");
    DD_CHECK_ARG(a->win_members[w] * Gq <= 16, "dd_decoder_pass: a window of %d members x %d query heads per kv head "
                 "exceeds 16 query rows", a->win_members[w], Gq);
    covered += a->win_members[w];
    max_m = a->win_members[w] > max_m ? a->win_members[w] : max_m;
  }
  DD_CHECK_ARG(covered == a->members_total, "dd_decoder_pass: member windows cover %d of %d members", covered,
               a->members_total);
  if (a->workspace_bytes < dd_pass_workspace_bytes(B, a->members_total, a->H, a->D, a->num_splits)) {
    set_error("dd_decoder_pass: workspace %zu < %zu bytes", a->workspace_bytes,
              dd_pass_workspace_bytes(B, a->members_total, a->H, a->D, a->num_splits));
    return DD_ERR_WORKSPACE;
  }

  PassParams p{};
  p.layers = reinterpret_cast<const PassLayer*>(a->plan);
  p.L = a->n_layers; p.rows = a->rows; p.d = a->d_model; p.dff = a->d_ff; p.H = a->H; p.H_kv = a->H_kv;
  p.S_max = a->S_max; p.Vp = a->vocab_padded; p.vocab = a->vocab; p.eps = a->eps;
  p.embed = (const __nv_bfloat16*)a->embed; p.tok = a->tok; p.pos = a->pos; p.tok_div = a->tok_div; p.pos_div = a->pos_div;
  p.cos_tab = a->cos_tab; p.sin_tab = a->sin_tab;
  p.x = (__nv_bfloat16*)a->x; p.x2 = (__nv_bfloat16*)a->x2; p.act = (__nv_bfloat16*)a->act; p.q = (__nv_bfloat16*)a->q;
  p.attn = (__nv_bfloat16*)a->attn; p.xn_out = (__nv_bfloat16*)a->xn_out;
  p.final_norm = (const __nv_bfloat16*)a->final_norm; p.lm_head = (const __nv_bfloat16*)a->lm_head; p.logits = a->logits;
  p.n_windows = a->n_windows; p.guess_winner = a->guess_winner; p.guess_pos = a->guess_pos;
  p.sync = reinterpret_cast<unsigned*>(a->workspace);
  p.pf_gu_bytes = (unsigned long long)(a->prefetch_gate_up_bytes & ~(size_t)15);
  p.timeline = reinterpret_cast<unsigned long long*>(a->timeline);
  const size_t win_ws = (dd_attention_workspace_bytes(B, a->members_total, a->H, a->D, a->num_splits) + 255) / 256 * 256;
  for (int w = 0; w < a->n_windows; ++w) {
    AttnParams& ap = p.win[w];
    ap.q = p.q; ap.seq_lens = a->win_seq_lens[w]; ap.vis_start = a->vis_start; ap.bitmask = a->bitmask;
    float* ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + 256 + (size_t)w * win_ws);
    ap.part_o = ws;
    ap.part_ml = ap.part_o + (size_t)B * a->members_total * a->H * a->num_splits * a->D;
    ap.tickets = reinterpret_cast<int*>(ap.part_ml + (size_t)B * a->members_total * a->H * a->num_splits * 2);
    ap.out = p.attn;
    ap.B = B; ap.M = a->win_members[w]; ap.Mt = a->members_total; ap.m0 = a->win_first[w]; ap.H = a->H; ap.H_kv = a->H_kv;
    ap.S_max = a->S_max; ap.n_vis = a->n_vis; ap.mask_words = a->mask_words; ap.num_splits = a->num_splits;
    ap.scale_log2 = a->scale * 1.4426950408889634f;
  }
  const int MT = (a->rows + 7) / 8;
  const int kmax = a->d_model > a->H * a->D ? a->d_model : a->H * a->D;
  const size_t xs = (size_t)MT * 8 * (kmax + kXPad) * 2;
  const size_t attn_b = (a->D == 128 ? attn_smem_bytes<128, 1>(max_m, a->bitmask ? a->mask_words : 0)
                                     : attn_smem_bytes<64, 1>(max_m, a->bitmask ? a->mask_words : 0)) - 1024;
  p.xs_bytes = (int)(((xs > attn_b ? xs : attn_b) + 127) / 128 * 128);
  p.red_bytes = kGroupWarps * 2 * MT * 16 * 9 * (int)sizeof(float);
  const size_t smem = 1024 + (size_t)p.xs_bytes + p.red_bytes;
  DD_CHECK_ARG(smem <= 227 * 1024, "dd_decoder_pass: %zu bytes of shared memory needed (d_model=%d, rows=%d)", smem,
               a->d_model, a->rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->D == 128) return MT == 1 ? launch_pass<1, 128>(p, smem, st) : launch_pass<2, 128>(p, smem, st);
  return MT == 1 ? launch_pass<1, 64>(p, smem, st) : launch_pass<2, 64>(p, smem, st);
}
