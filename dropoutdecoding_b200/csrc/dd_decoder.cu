// Decoder support kernels for the single-token forwards of the ensemble (RMSNorm, RoPE + member
// staging, embedding gather, winner commit, step bookkeeping).  They follow the HF Llama decoder
// the reference calls through `self.language_model(...)` (models/llava.py:294-303,350-359;
// models/llama.py:27-47).  All are small and latency-bound; the weight-streaming GEMM lives in
// dd_skinny_gemm.cu and the attention in dd_attention.cu.
#include "dd_common.cuh"

namespace dd {

// ---------------------------------------------------------------------------------------------
// RMSNorm: y = bf16( bf16(x * rsqrt(mean(x^2) + eps)) * w )   (LlamaRMSNorm rounding points)
// ---------------------------------------------------------------------------------------------
constexpr int kNormThreads = 256;
constexpr int kPrefetchCtas = 144;

__global__ void __launch_bounds__(kNormThreads) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ w, int d,
                                                              float eps, __nv_bfloat16* __restrict__ y, int M,
                                                              const L2Prefetch pf) {
  __shared__ float s_part[kNormThreads / 32];
  pdl_trigger();
  if ((int)blockIdx.x >= M) {  // prefetch-only CTAs (dd_hint_l2_prefetch): the weights do not depend on earlier kernels
    l2_prefetch_slice(pf, (blockIdx.x - M) * (kNormThreads / 32) + (threadIdx.x >> 5), (gridDim.x - M) * (kNormThreads / 32));
    return;
  }
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)row * d);
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  uint4* y4 = reinterpret_cast<uint4*>(y + (size_t)row * d);
  const int nvec = d / 8;
  // rows of up to kNormKeep * 256 * 8 elements (d <= 8192) stay in registers between the two passes, and the norm
  // weights are fetched together with them: one memory round trip instead of two dependent ones
  constexpr int kNormKeep = 4;
  const bool keep = nvec <= kNormKeep * kNormThreads;
  uint4 xv[kNormKeep], gv[kNormKeep];
  float ss = 0.f;
  if (keep) {
#pragma unroll
    for (int i = 0; i < kNormKeep; ++i) {
      const int j = tid + i * kNormThreads;
      if (j < nvec) { xv[i] = x4[j]; gv[i] = w4[j]; }
    }
#pragma unroll
    for (int i = 0; i < kNormKeep; ++i) {
      if (tid + i * kNormThreads < nvec) {
        const uint32_t u[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = bf16x2_to_float2(u[k]);
          ss = fmaf(f.x, f.x, ss);
          ss = fmaf(f.y, f.y, ss);
        }
      }
    }
  } else {
    for (int j = tid; j < nvec; j += kNormThreads) {
      const uint4 v = x4[j];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = bf16x2_to_float2(u[k]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_part[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < kNormThreads / 32; ++k) tot += s_part[k];
  const float rstd = rsqrtf(tot / (float)d + eps);
  auto scale = [&](const uint4& v, const uint4& g) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = bf16x2_to_float2(u[k]), ww = bf16x2_to_float2(gw[k]);
      o[k] = pack_bf16x2(bf16_round(f.x * rstd) * ww.x, bf16_round(f.y * rstd) * ww.y);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  };
  if (keep) {
#pragma unroll
    for (int i = 0; i < kNormKeep; ++i) {
      const int j = tid + i * kNormThreads;
      if (j < nvec) y4[j] = scale(xv[i], gv[i]);
    }
  } else {
    for (int j = tid; j < nvec; j += kNormThreads) y4[j] = scale(x4[j], w4[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// RoPE (rotate-half) + member staging.  One CTA per (row = b*M + m, head of q|k|v), D/2 threads:
// a thread owns the pair (i, i + D/2).  (The <= 32-row decode path fuses this into the QKV GEMM.)
// ---------------------------------------------------------------------------------------------
__global__ void rope_stage_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ cos_tab,
                                  const float* __restrict__ sin_tab, const int32_t* __restrict__ seq_lens,
                                  int M, int H, int H_kv, int D, __nv_bfloat16* __restrict__ q_out,
                                  __nv_bfloat16* __restrict__ k_new, __nv_bfloat16* __restrict__ v_new) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x, head = blockIdx.y, i = threadIdx.x;  // b*M + m
  const int half = D / 2;
  const __nv_bfloat16* src = qkv + ((size_t)row * (H + 2 * H_kv) + head) * D;
  if (head >= H + H_kv) {
    __nv_bfloat16* dst = v_new + ((size_t)row * H_kv + (head - H - H_kv)) * D;
    dst[i] = src[i];
    dst[i + half] = src[i + half];
    return;
  }
  const int pos = seq_lens[row / M];
  const float c = cos_tab[(size_t)pos * half + i], s = sin_tab[(size_t)pos * half + i];
  const float x1 = __bfloat162float(src[i]), x2 = __bfloat162float(src[i + half]);
  __nv_bfloat16* dst = head < H ? q_out + ((size_t)row * H + head) * D : k_new + ((size_t)row * H_kv + (head - H)) * D;
  dst[i] = __float2bfloat16_rn(x1 * c - x2 * s);
  dst[i + half] = __float2bfloat16_rn(x2 * c + x1 * s);
}

__global__ void embed_gather_kernel(const uint4* __restrict__ embed, const int32_t* __restrict__ token, int M,
                                    int dvec, int vocab, uint4* __restrict__ x) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;  // b*M + m
  const int t = min(max(token[row / M], 0), vocab - 1);   // a caller-supplied id never indexes outside the table
  for (int j = threadIdx.x; j < dvec; j += blockDim.x) x[(size_t)row * dvec + j] = embed[(size_t)t * dvec + j];
}

__global__ void step_advance_kernel(const int32_t* __restrict__ token, int32_t* tokens_out, int max_steps,
                                    int32_t* seq_lens, int32_t* step, int B) {
  pdl_trigger();
  pdl_wait();
  const int st = *step;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (tokens_out && st + 1 < max_steps) tokens_out[(size_t)b * max_steps + st + 1] = token[b];
    seq_lens[b] += 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) *step = st + 1;
}

// per-sequence progress (asynchronous batch schedule): only sequences with advance[b] != 0 move on
__global__ void step_advance_seq_kernel(const int32_t* __restrict__ token, const int32_t* __restrict__ advance,
                                        int32_t* tokens_out, int max_steps, int32_t* seq_lens, int32_t* seq_steps,
                                        int B) {
  pdl_trigger();
  pdl_wait();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    if (!advance[b]) continue;
    const int st = seq_steps[b];
    if (tokens_out && st + 1 < max_steps) tokens_out[(size_t)b * max_steps + st + 1] = token[b];
    seq_lens[b] += 1;
    seq_steps[b] = st + 1;
  }
}

// winner's new-token K/V rows of every layer -> cache row seq_lens[b].  grid (L, B), 16B per thread.
__global__ void commit_winner_kernel(const uint4* __restrict__ k_new, const uint4* __restrict__ v_new,
                                     const int32_t* __restrict__ winner, const int32_t* __restrict__ seq_lens,
                                     uint4* __restrict__ k_cache, uint4* __restrict__ v_cache, int B, int M,
                                     int H_kv, int dvec, int S_max) {
  pdl_trigger();
  pdl_wait();
  const int l = blockIdx.x, b = blockIdx.y;
  const int w = winner[b], pos = seq_lens[b];
  if (pos >= S_max || w < 0 || w >= M) return;  // w < 0: this sequence commits from another staging buffer
  const size_t src = (((size_t)l * B + b) * M + w) * H_kv * dvec;
  for (int idx = threadIdx.x; idx < H_kv * dvec; idx += blockDim.x) {
    const int h = idx / dvec, j = idx % dvec;
    const size_t dst = ((((size_t)l * B + b) * H_kv + h) * S_max + pos) * dvec + j;
    k_cache[dst] = k_new[src + idx];
    v_cache[dst] = v_new[src + idx];
  }
}

// Member-sharded decode (ensemble members on different GPUs): ONE launch after the per-step all-gather.
// Every rank has written one exchange record -- { member argmax ids [B, mpr] | token statistics [mpr, B, 3] |
// staged K rows [L, B, mpr, H_kv, D] | staged V rows } -- straight from its kernels into the collective's send buffer;
// `recv` holds the G records.  This kernel votes over the gathered ids (select_by_vote, models/llava.py:22-36:
// most frequent id, ties to the id seen first, winner = first member holding it), publishes ids / winner / token and
// the winner's token statistics, and commits the winner's K/V rows from the gathered buffer into the cache row
// seq_lens[b] -- the caches of all ranks stay bit-identical.  forced_winner >= 0 (logit-mean mode: member 0's cache
// survives, models/llava.py:39,50) skips the vote and takes the token from forced_token.
__global__ void member_finish_kernel(const uint8_t* __restrict__ recv, long long rank_stride, long long ids_off,
                                     long long stats_off, long long k_off, long long v_off, int G, int mpr, int B, int H_kv,
                                     int dvec, int S_max, int forced_winner, const int32_t* __restrict__ forced_token,
                                     const int32_t* __restrict__ seq_lens, const int32_t* __restrict__ step, int n_slots,
                                     int32_t* member_ids, int32_t* winner, int32_t* token, float* token_stats,
                                     uint4* __restrict__ k_cache, uint4* __restrict__ v_cache) {
  __shared__ int s_w;
  pdl_trigger();
  pdl_wait();
  const int l = blockIdx.x, b = blockIdx.y, K = G * mpr;
  if (threadIdx.x == 0) {
    int ids[DD_MAX_MEMBERS];
    for (int m = 0; m < K; ++m)
      ids[m] = reinterpret_cast<const int32_t*>(recv + (size_t)(m / mpr) * rank_stride + ids_off)[b * mpr + m % mpr];
    int w = 0, tok = ids[0];
    if (forced_winner >= 0) {
      w = forced_winner;
      tok = forced_token[b];
    } else {
      int best_cnt = 0;
      for (int m = 0; m < K; ++m) {
        int c = 0;
        for (int j = 0; j < K; ++j) c += (ids[j] == ids[m]);
        if (c > best_cnt) { best_cnt = c; tok = ids[m]; }
      }
      while (ids[w] != tok) ++w;
    }
    s_w = w;
    if (l == 0) {
      for (int m = 0; m < K; ++m) member_ids[b * K + m] = ids[m];
      winner[b] = w;
      token[b] = tok;
      const int slot = *step + 1;
      if (token_stats && forced_winner < 0 && slot >= 0 && slot < n_slots) {
        const float* st = reinterpret_cast<const float*>(recv + (size_t)(w / mpr) * rank_stride + stats_off) +
                          ((size_t)(w % mpr) * B + b) * 3;
        float* o = token_stats + ((size_t)slot * B + b) * 3;
        o[0] = st[0]; o[1] = st[1]; o[2] = st[2];
      }
    }
  }
  __syncthreads();
  const int w = s_w, pos = seq_lens[b];
  if (pos >= S_max) return;
  const size_t row = ((((size_t)l * B + b) * mpr + w % mpr) * H_kv) * dvec;
  const uint4* ks = reinterpret_cast<const uint4*>(recv + (size_t)(w / mpr) * rank_stride + k_off) + row;
  const uint4* vs = reinterpret_cast<const uint4*>(recv + (size_t)(w / mpr) * rank_stride + v_off) + row;
  for (int idx = threadIdx.x; idx < H_kv * dvec; idx += blockDim.x) {
    const int h = idx / dvec, j = idx % dvec;
    const size_t dst = ((((size_t)l * B + b) * H_kv + h) * S_max + pos) * dvec + j;
    k_cache[dst] = ks[idx];
    v_cache[dst] = vs[idx];
  }
}

}  // namespace dd

extern "C" int dd_rmsnorm(const void* x, const void* w, int M, int d, float eps, void* y, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(x && w && y, "dd_rmsnorm: null pointer");
  DD_CHECK_ARG(M >= 0 && d > 0 && d % 8 == 0, "dd_rmsnorm: d=%d must be a positive multiple of 8", d);
  const L2Prefetch pf = take_l2_prefetch();
  if (M == 0) return DD_OK;
  // with a prefetch hint: ~one extra CTA per SM that only issues the bulk prefetches (the row CTAs stay as they are)
  const int extra = (pf.bytes[0] | pf.bytes[1]) ? kPrefetchCtas : 0;
  DD_LAUNCH_CHAIN(rmsnorm_kernel, dim3(M + extra), dim3(kNormThreads), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x,
                  (const __nv_bfloat16*)w, d, eps, (__nv_bfloat16*)y, M, pf);
  return DD_OK;
}

extern "C" int dd_rope_stage(const void* qkv, const float* cos_tab, const float* sin_tab,
                             const int32_t* seq_lens, int B, int M, int H, int H_kv, int D, void* q_out,
                             void* k_new, void* v_new, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(qkv && cos_tab && sin_tab && seq_lens && q_out && k_new && v_new, "dd_rope_stage: null pointer");
  DD_CHECK_ARG(D % 2 == 0 && H >= 1 && H_kv >= 1, "dd_rope_stage: bad head shape");
  if (B * M == 0) return DD_OK;
  DD_LAUNCH_CHAIN(rope_stage_kernel, dim3(B * M, H + 2 * H_kv), dim3(D / 2), 0, (cudaStream_t)stream, (const __nv_bfloat16*)qkv,
                  cos_tab, sin_tab, seq_lens, M, H, H_kv, D, (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_new,
                  (__nv_bfloat16*)v_new);
  return DD_OK;
}

extern "C" int dd_embed_gather(const void* embed, const int32_t* token, int B, int M, int d, int vocab, void* x,
                               dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(embed && token && x, "dd_embed_gather: null pointer");
  DD_CHECK_ARG(d % 8 == 0, "dd_embed_gather: d=%d must be a multiple of 8", d);
  DD_CHECK_ARG(vocab >= 1, "dd_embed_gather: vocab=%d", vocab);
  if (B * M == 0) return DD_OK;
  DD_LAUNCH_CHAIN(embed_gather_kernel, dim3(B * M), dim3(128), 0, (cudaStream_t)stream, (const uint4*)embed, token,
                  M, d / 8, vocab, (uint4*)x);
  return DD_OK;
}

extern "C" int dd_step_advance(const int32_t* token, int32_t* tokens_out, int max_steps, int32_t* seq_lens,
                               int32_t* step, int B, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(token && seq_lens && step, "dd_step_advance: null pointer");
  if (B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(step_advance_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, token, tokens_out, max_steps,
                  seq_lens, step, B);
  return DD_OK;
}

extern "C" int dd_step_advance_seq(const int32_t* token, const int32_t* advance, int32_t* tokens_out, int max_steps,
                                   int32_t* seq_lens, int32_t* seq_steps, int B, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(token && advance && seq_lens && seq_steps, "dd_step_advance_seq: null pointer");
  if (B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(step_advance_seq_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, token, advance, tokens_out,
                  max_steps, seq_lens, seq_steps, B);
  return DD_OK;
}

extern "C" int dd_commit_winner(const void* k_new_all, const void* v_new_all, const int32_t* winner,
                                const int32_t* seq_lens, void* k_cache_all, void* v_cache_all, int L, int B,
                                int M, int H_kv, int D, int S_max, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(k_new_all && v_new_all && winner && seq_lens && k_cache_all && v_cache_all,
               "dd_commit_winner: null pointer");
  DD_CHECK_ARG(D % 8 == 0, "dd_commit_winner: D=%d must be a multiple of 8", D);
  if (L * B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(commit_winner_kernel, dim3(L, B), dim3(128), 0, (cudaStream_t)stream, (const uint4*)k_new_all,
                  (const uint4*)v_new_all, winner, seq_lens, (uint4*)k_cache_all, (uint4*)v_cache_all, B, M, H_kv,
                  D / 8, S_max);
  return DD_OK;
}

extern "C" int dd_member_finish(const void* recv, long long rank_stride, long long ids_off, long long stats_off,
                                long long k_off, long long v_off, int G, int mpr, int L, int B, int H_kv, int D, int S_max,
                                int forced_winner, const int32_t* forced_token, const int32_t* seq_lens,
                                const int32_t* step, int n_slots, int32_t* member_ids, int32_t* winner, int32_t* token,
                                float* token_stats, void* k_cache_all, void* v_cache_all, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(recv && seq_lens && step && member_ids && winner && token && k_cache_all && v_cache_all,
               "dd_member_finish: null pointer");
  DD_CHECK_ARG(G >= 1 && mpr >= 1 && G * mpr <= DD_MAX_MEMBERS, "dd_member_finish: %d ranks x %d members out of [1,%d]",
               G, mpr, DD_MAX_MEMBERS);
  DD_CHECK_ARG(D % 8 == 0 && L >= 1 && H_kv >= 1 && S_max >= 1, "dd_member_finish: bad shape");
  DD_CHECK_ARG((rank_stride | ids_off | stats_off | k_off | v_off) % 16 == 0 && ((uintptr_t)recv & 15) == 0,
               "dd_member_finish: the exchange record must be 16-byte aligned");
  DD_CHECK_ARG(forced_winner < G * mpr && (forced_winner < 0 || forced_token), "dd_member_finish: bad forced winner");
  if (B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(member_finish_kernel, dim3(L, B), dim3(128), 0, (cudaStream_t)stream, (const uint8_t*)recv, rank_stride,
                  ids_off, stats_off, k_off, v_off, G, mpr, B, H_kv, D / 8, S_max, forced_winner, forced_token, seq_lens,
                  step, n_slots, member_ids, winner, token, token_stats, (uint4*)k_cache_all, (uint4*)v_cache_all);
  return DD_OK;
}
