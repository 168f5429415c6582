// Device-side body of the decode attention (shared by dd_attention.cu and the persistent pass kernel of dd_pass.cu).
// See dd_attention.cu for the design notes.
#pragma once
#include <cuda.h>

#include "dd_common.cuh"

namespace dd {

constexpr int kAttnWarps = 4;
constexpr int kAttnThreads = kAttnWarps * 32;
constexpr int kTileKeys = 16 * kAttnWarps;  // 64 keys per pipeline stage
constexpr int kStages = 3;
constexpr int kMaxMaskWords = 256;  // n_vis <= 8192

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

struct AttnParams {
  const __nv_bfloat16 *q, *k_cache, *v_cache;
  const int32_t *seq_lens, *vis_start;
  const int32_t* seq_active;  // [B] or NULL: sequences with 0 are skipped (their output rows stay untouched) ...
  const int32_t* seq_active2; // ... unless this second flag (or NULL) is set: a pass with two member windows
  const int32_t* len_extra;   // [B, Mt] or NULL: member m of sequence b attends keys [0, seq_lens[b] + len_extra[b,m])
                              //   (0 / 1: the lookahead contexts of a pass also see the row committed ahead)
  const uint32_t* bitmask;
  float *part_o, *part_ml;  // [B,H_kv,splits,R,D], [B,H_kv,splits,R,2]
  int* tickets;             // [B,H_kv] zero on entry, zero again on exit
  const __nv_bfloat16 *k_new, *v_new;
  __nv_bfloat16* out;
  int B, M, H, H_kv, S_max, n_vis, mask_words, num_splits;
  int Mt, m0;  // the M members of this call are members m0..m0+M-1 of the Mt stored per sequence (q, k_new, v_new, out, bitmask)
  float scale_log2;  // scale * log2(e)
  // second member window of the same launch (dd_decode_attention_two_windows; M2 == 0: none): members m02..m02+M2-1
  // attend keys [0, seq_lens2[b]) with their own split partials -- the grid's z index runs over (window, sequence)
  int M2, m02;
  const int32_t* seq_lens2;
  float *part_o2, *part_ml2;
  int* tickets2;
};

// Q rows (staged by the threads): one row of D bf16 = D/8 chunks of 16 B; chunk c of row r lives at
// physical chunk c ^ (r & 7): ldmatrix's 8 row addresses then hit 8 distinct 16-byte bank groups.
template <int D>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}
// K / V tiles (written by TMA, SWIZZLE_128B): a tile of 64 keys is D/64 boxes of [64 keys x 128 B]; inside a
// box the 16-byte chunk c of row r lives at chunk c ^ (r & 7) of the row -- the same conflict-free pattern.
template <int D>
__device__ __forceinline__ uint32_t swz_kv(int row, int chunk) {
  return (uint32_t)((chunk >> 3) * (kTileKeys * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// One work item of the decode attention: split `split` of (sequence b, kv head hkv), run by kAttnThreads threads
// (tid = 0..127) that synchronise through `sync` (the block barrier in attn_partial_kernel, a named barrier inside the
// persistent pass kernel of dd_pass.cu).  smem: 1024-byte aligned, attn_smem_bytes<D, QT>() - 1024 bytes; s_full /
// s_ext / s_last_p: small shared scratch owned by the caller.  May be called again (next item) after it returns: the
// mbarriers are re-initialised on entry, every thread has passed the last sync of the previous item by then.
// Warps 1.. issue the L2 prefetch hint `pf` behind the ring's first tiles (cta_lin / n_cta: this CTA in the grid).
// wait_for_predecessor: griddepcontrol.wait of a programmatic dependent launch (a no-op functor elsewhere).  It is
// called AFTER the mbarrier set-up and after the loads of the first tiles that end before the range's last key have
// been issued: those cache rows were final before the kernel in front of this one started (the only cache row a pass
// writes ahead of its attention is the lookahead's commit at the LAST key of a window), while q, k_new, v_new and the
// masks are only touched after the wait.
template <int D, int QT, bool kChainTrigger, typename Sync, typename Wait>
__device__ __forceinline__ void attn_partial_body(const CUtensorMap* tmap_k_p, const CUtensorMap* tmap_v_p,
                                                  const AttnParams& p, const L2Prefetch& pf, int split, int hkv, int b,
                                                  int tid, unsigned cta_lin, unsigned n_cta, uint8_t* smem,
                                                  uint64_t* s_full, int* s_ext, int* s_last_p, Sync sync,
                                                  Wait wait_for_predecessor) {
  constexpr int CH = D / 8;            // 16-byte chunks per row
  constexpr int ROW_B = D * 2;         // bytes per row
  constexpr int TILE_B = kTileKeys * ROW_B;
  uint8_t* sK = smem;                           // [kStages][64][D]
  uint8_t* sV = smem + kStages * TILE_B;        // [kStages][64][D]
  uint8_t* sQ = smem + 2 * kStages * TILE_B;    // [16*QT][D]
  uint32_t* sMask = reinterpret_cast<uint32_t*>(sQ + 16 * QT * ROW_B);  // [M][mask_words]
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int G = p.H / p.H_kv, R = p.M * G;
  if (p.seq_active && p.seq_active[b] == 0 && !(p.seq_active2 && p.seq_active2[b] != 0)) {
    if (warp) l2_prefetch_slice(pf, cta_lin * (kAttnWarps - 1) + (warp - 1), n_cta * (kAttnWarps - 1));
    return;  // CTA-uniform: nothing of this sequence is read
  }
  const int S0 = p.seq_lens[b];
  int any_extra = 0;
  if (p.len_extra)
    for (int m = 0; m < p.M; ++m) any_extra |= p.len_extra[(size_t)b * p.Mt + p.m0 + m];
  const int S = S0 + (any_extra ? 1 : 0);   // keys streamed; a member without the extra key masks it below
  const int vis0 = p.vis_start ? p.vis_start[b] : 0;
  const int vis1 = p.bitmask ? vis0 + p.n_vis : vis0;  // empty span when unmasked

  // this CTA's key range
  const int per = ((S + p.num_splits - 1) / p.num_splits + kTileKeys - 1) / kTileKeys * kTileKeys;
  const int k_beg = split * per;
  const int k_end = min(S, k_beg + per);
  const int ntiles = k_end > k_beg ? (k_end - k_beg + kTileKeys - 1) / kTileKeys : 0;

  const uint32_t bar_full = smem_u32(s_full);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(bar_full + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  sync();

  // one elected thread issues the D/64 + D/64 boxes of a tile; the stage's mbarrier counts the bytes
  auto load_tile = [&](int tile, int stage) {
    const int key0 = k_beg + tile * kTileKeys;
    const uint32_t kb = smem_u32(sK + stage * TILE_B), vb = smem_u32(sV + stage * TILE_B);
    const uint32_t bar = bar_full + 8 * stage;
    mbar_expect_tx(bar, 2 * TILE_B);
#pragma unroll
    for (int h = 0; h < D / 64; ++h) {
      tma_load_3d(kb + h * (kTileKeys * 128), tmap_k_p, bar, h * 64, key0, b * p.H_kv + hkv);
      tma_load_3d(vb + h * (kTileKeys * 128), tmap_v_p, bar, h * 64, key0, b * p.H_kv + hkv);
    }
  };

  // ---- prologue: start the ring (early tiles before the dependency wait), stage Q and the member bitmasks --------
  int early = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s)
      if (early == s && s < ntiles && k_beg + (s + 1) * kTileKeys <= S - 1) {
        load_tile(s, s);
        early = s + 1;
      }
  }
  wait_for_predecessor();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s)
      if (s >= early && s < ntiles) load_tile(s, s);
  }
  // L2 prefetch hint (dd_hint_l2_prefetch): warps 1.. issue it behind the ring's first K/V tiles
  if (warp) l2_prefetch_slice(pf, cta_lin * (kAttnWarps - 1) + (warp - 1), n_cta * (kAttnWarps - 1));
  for (int idx = tid; idx < 16 * QT * CH; idx += kAttnThreads) {
    const int r = idx / CH, c = idx % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < R) {
      const int m = r / G, h = hkv * G + (r % G);
      v = *reinterpret_cast<const uint4*>(p.q + (((size_t)b * p.Mt + p.m0 + m) * p.H + h) * D + c * 8);
    }
    *reinterpret_cast<uint4*>(sQ + swz<D>(r, c)) = v;
  }
  if (p.bitmask)
    for (int idx = tid; idx < p.M * p.mask_words; idx += kAttnThreads)
      sMask[idx] = p.bitmask[((size_t)b * p.Mt + p.m0) * p.mask_words + idx];
  if (tid < p.M) s_ext[tid] = p.len_extra ? (p.len_extra[(size_t)b * p.Mt + p.m0 + tid] != 0) : 0;

  float o[QT][D / 8][4];
  float m_run[QT][2], l_run[QT][2];
#pragma unroll
  for (int qt = 0; qt < QT; ++qt) {
    m_run[qt][0] = m_run[qt][1] = -INFINITY;
    l_run[qt][0] = l_run[qt][1] = 0.f;
#pragma unroll
    for (int n = 0; n < D / 8; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[qt][n][c] = 0.f;
  }

  sync();  // Q and the masks are staged
  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile % kStages;
    sync();  // everyone is done with the stage that is refilled next (it held tile - 1)
    if (tid == 0) {
      const int nt = tile + kStages - 1;
      if (nt < ntiles) load_tile(nt, nt % kStages);
    }
    mbar_wait(bar_full + 8 * stage, (tile / kStages) & 1);  // tile `tile` has landed
    if (tile == ntiles - 1 && (k_end - k_beg) % kTileKeys != 0) {
      // cache rows past seq_len are never written by anyone and may hold anything: a masked key has p = 0, but
      // 0 * NaN would still poison P.V, so clear the dead V rows of the last tile
      const int first_dead = k_end - (k_beg + tile * kTileKeys);
      uint8_t* vt = sV + stage * TILE_B;
      for (int idx = tid; idx < (kTileKeys - first_dead) * CH; idx += kAttnThreads) {
        const int row = first_dead + idx / CH, c = idx % CH;
        *reinterpret_cast<uint4*>(vt + swz_kv<D>(row, c)) = make_uint4(0, 0, 0, 0);
      }
      sync();
    }
    const uint32_t kb = smem_u32(sK + stage * TILE_B), vb = smem_u32(sV + stage * TILE_B);
    const int key0 = k_beg + tile * kTileKeys + warp * 16;  // this warp's 16 keys
    if (key0 >= k_end) continue;                            // warp-uniform
    const bool need_mask = (key0 + 16 > (any_extra ? min(k_end, S0) : k_end)) || (key0 < vis1 && key0 + 16 > vis0);

#pragma unroll
    for (int qt = 0; qt < QT; ++qt) {
      // ---- S = Q K^T for 16 query rows x 16 keys --------------------------------------------
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        uint32_t a[4], kf[4];
        ldmatrix_x4(a, smem_u32(sQ) + swz<D>(qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)));
        ldmatrix_x4(kf, kb + swz_kv<D>(warp * 16 + (lane & 7) + ((lane >> 4) << 3), 2 * ks + ((lane >> 3) & 1)));
        mma_bf16_16816(s[0], a[0], a[1], a[2], a[3], kf[0], kf[1]);
        mma_bf16_16816(s[1], a[0], a[1], a[2], a[3], kf[2], kf[3]);
      }
      // ---- scale, mask (registers), online softmax ---------------------------------------------
      // fragment: s[j][0..1] -> row g, keys 8j+2t,+1 ; s[j][2..3] -> row g+8
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[j][c] *= p.scale_log2;
      if (need_mask) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int r = qt * 16 + g + hr * 8;
          const int mrow = r < R ? r / G : 0;
          const uint32_t* mw = sMask + mrow * p.mask_words;
          const int k_row = any_extra ? min(k_end, S0 + s_ext[mrow]) : k_end;   // this member's last key + 1
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = key0 + 8 * j + 2 * t + e;
              bool dead = key >= k_row;
              const int vp = key - vis0;
              if (key < vis1 && vp >= 0) dead |= (mw[vp >> 5] >> (vp & 31)) & 1u;
              if (dead) s[j][hr * 2 + e] = -INFINITY;
            }
        }
      }
      float alpha[2];
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        float mx = fmaxf(fmaxf(s[0][hr * 2], s[0][hr * 2 + 1]), fmaxf(s[1][hr * 2], s[1][hr * 2 + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run[qt][hr], mx);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;  // all keys masked so far
        alpha[hr] = exp2f(m_run[qt][hr] - m_safe);
        m_run[qt][hr] = m_new;
        float rs = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = exp2f(s[j][hr * 2 + e] - m_safe);
            s[j][hr * 2 + e] = pv;
            rs += pv;
          }
        l_run[qt][hr] = l_run[qt][hr] * alpha[hr] + rs;  // per-thread partial; quad-reduced at the end
      }
      // ---- O = alpha * O + P V with P = P_hi + P_lo (both bf16) -----------------------------
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const float p0 = s[j][hr * 2], p1 = s[j][hr * 2 + 1];
          const float h0 = bf16_round(p0), h1 = bf16_round(p1);
          ph[j * 2 + hr] = pack_bf16x2(h0, h1);  // a0a1: row g keys 2t.. (j=0) | a2a3: row g+8 | a4a5: j=1 ...
          pl[j * 2 + hr] = pack_bf16x2(p0 - h0, p1 - h1);
        }
#pragma unroll
      for (int n = 0; n < D / 8; ++n) {
        o[qt][n][0] *= alpha[0];
        o[qt][n][1] *= alpha[0];
        o[qt][n][2] *= alpha[1];
        o[qt][n][3] *= alpha[1];
      }
#pragma unroll
      for (int n2 = 0; n2 < D / 16; ++n2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, vb + swz_kv<D>(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * n2 + (lane >> 4)));
        mma_bf16_16816(o[qt][2 * n2], ph[0], ph[1], ph[2], ph[3], vf[0], vf[1]);
        mma_bf16_16816(o[qt][2 * n2 + 1], ph[0], ph[1], ph[2], ph[3], vf[2], vf[3]);
        mma_bf16_16816(o[qt][2 * n2], pl[0], pl[1], pl[2], pl[3], vf[0], vf[1]);
        mma_bf16_16816(o[qt][2 * n2 + 1], pl[0], pl[1], pl[2], pl[3], vf[2], vf[3]);
      }
    }
  }
  if (kChainTrigger) pdl_trigger_stream_tail();
  sync();  // ring is idle: reuse it for the cross-warp merge

  // ---- merge the 4 warps (disjoint key subsets) and write the split partial ---------------------
  float* sO = reinterpret_cast<float*>(smem);                       // [warps][16*QT][D]
  float* sML = sO + kAttnWarps * 16 * QT * D;                        // [warps][16*QT][2]
#pragma unroll
  for (int qt = 0; qt < QT; ++qt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      float l = l_run[qt][hr];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const int r = qt * 16 + g + hr * 8;
      if (t == 0) {
        sML[(warp * 16 * QT + r) * 2] = m_run[qt][hr];
        sML[(warp * 16 * QT + r) * 2 + 1] = l;
      }
#pragma unroll
      for (int n = 0; n < D / 8; ++n) {
        float* dst = sO + ((size_t)(warp * 16 * QT + r)) * D + n * 8 + 2 * t;
        dst[0] = o[qt][n][hr * 2];
        dst[1] = o[qt][n][hr * 2 + 1];
      }
    }
  sync();
  const size_t pbase = (((size_t)b * p.H_kv + hkv) * p.num_splits + split) * R;
  for (int idx = tid; idx < R * D; idx += kAttnThreads) {
    const int r = idx / D, dcol = idx % D;
    float mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) mx = fmaxf(mx, sML[(w * 16 * QT + r) * 2]);
    const float ms = mx == -INFINITY ? 0.f : mx;
    float acc = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float f = exp2f(sML[(w * 16 * QT + r) * 2] - ms);
      acc += f * sO[((size_t)(w * 16 * QT + r)) * D + dcol];
      l += f * sML[(w * 16 * QT + r) * 2 + 1];
    }
    p.part_o[(pbase + r) * D + dcol] = acc;
    if (dcol == 0) {
      p.part_ml[(pbase + r) * 2] = mx;
      p.part_ml[(pbase + r) * 2 + 1] = l;
    }
  }

  // Unsplit (large batches: the grid already fills the GPU): the one-partial "merge" with the member's own row
  // is a separate, fully parallel launch -- measured faster at B=64 than a serial tail on every streaming CTA.
  if (p.num_splits == 1) return;
  // ---- split KV (small batches, latency-bound): the last CTA of this (sequence, kv head) merges the splits and
  //      each member's own new-token row, saving a launch ----
  __threadfence();
  sync();
  if (tid == 0) *s_last_p = atomicAdd(&p.tickets[b * p.H_kv + hkv], 1) == p.num_splits - 1;
  sync();
  if (!*s_last_p) return;
  __threadfence();
  constexpr int DL = D / 32;  // dims per lane: one warp per query row
  for (int r = warp; r < R; r += kAttnWarps) {
    const int m = r / G, h = hkv * G + (r % G);
    const size_t qoff = (((size_t)b * p.Mt + p.m0 + m) * p.H + h) * D + lane * DL;
    const size_t kvoff = (((size_t)b * p.Mt + p.m0 + m) * p.H_kv + hkv) * D + lane * DL;
    // score of the member's own new token (always attended: it is the query position itself)
    float dot = 0.f, vself[DL];
#pragma unroll
    for (int i = 0; i < DL; ++i) {
      dot = fmaf(__bfloat162float(p.q[qoff + i]), __bfloat162float(p.k_new[kvoff + i]), dot);
      vself[i] = __bfloat162float(p.v_new[kvoff + i]);
    }
    const float s_self = warp_sum(dot) * p.scale_log2;
    const size_t rbase = (((size_t)b * p.H_kv + hkv) * p.num_splits) * R + r;
    float mx = s_self;
    for (int sp = 0; sp < p.num_splits; ++sp) mx = fmaxf(mx, __ldcg(&p.part_ml[(rbase + (size_t)sp * R) * 2]));
    const float f_self = exp2f(s_self - mx);
    float acc[DL], l = f_self;
#pragma unroll
    for (int i = 0; i < DL; ++i) acc[i] = f_self * vself[i];
    for (int sp = 0; sp < p.num_splits; ++sp) {
      const size_t i0 = rbase + (size_t)sp * R;
      const float f = exp2f(__ldcg(&p.part_ml[i0 * 2]) - mx);
      l += f * __ldcg(&p.part_ml[i0 * 2 + 1]);
#pragma unroll
      for (int i = 0; i < DL; ++i) acc[i] += f * __ldcg(&p.part_o[i0 * D + lane * DL + i]);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < DL; ++i) p.out[qoff + i] = __float2bfloat16_rn(acc[i] * inv);
  }
  if (tid == 0) p.tickets[b * p.H_kv + hkv] = 0;  // ready for the next call on this workspace
}

template <int D, int QT>
static inline size_t attn_smem_bytes(int M, int mask_words) {
  const size_t ring = 2ull * kStages * kTileKeys * D * 2;
  const size_t merge = (size_t)kAttnWarps * 16 * QT * (D + 2) * sizeof(float);
  const size_t main_b = ring + 16ull * QT * D * 2 + (size_t)M * mask_words * 4;
  return 1024 + (main_b > merge ? main_b : merge);  // + slack to align the SWIZZLE_128B tiles to 1024 B
}


}  // namespace dd
