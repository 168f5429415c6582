// (4) Ensemble aggregation: fused per-member argmax + max-vote (or logit mean) in one launch,
// no host sync.  Replaces select_by_vote / select_by_average (models/llava.py:22-52), whose
// 2K `.item()` calls stall the host every token.
//
// HBM-bound: reads B*K*V fp32 once (+ writes B*V when out_logits is requested).
#include "dd_common.cuh"

namespace dd {

constexpr int kAggThreads = 512;

struct ArgMax {
  float v;
  int i;
};
// torch.argmax: the first index of the maximum; NaN counts as the maximum (so a row of NaNs, like a row of -inf,
// still yields a valid index and never the 0x7fffffff start value)
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  const bool gt = bn ? !an : (!an && b.v > a.v);
  const bool eq = (an && bn) || b.v == a.v;
  return (gt || (eq && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* s_warp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
    x = better(x, y);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // s_warp may still be read from the previous call
  if (lane == 0) s_warp[warp] = x;
  __syncthreads();
  ArgMax r = s_warp[0];
  for (int w = 1; w < kAggThreads / 32; ++w) r = better(r, s_warp[w]);
  return r;
}

__global__ void __launch_bounds__(kAggThreads) aggregate_kernel(const float* __restrict__ logits, int K,
                                                                int V, int ld, long long seq_stride, int mode,
                                                                int32_t* member_ids,
                                                                int32_t* winner, int32_t* token,
                                                                float* out_logits) {
  __shared__ ArgMax s_warp[kAggThreads / 32];
  __shared__ int s_ids[DD_MAX_MEMBERS];
  __shared__ int s_win;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* base = logits + (size_t)b * seq_stride;
  const bool vec = (V % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);

  // per-member argmax, four rows at a time: the four loads of an iteration are independent, so a single CTA (one
  // sequence) keeps four times the bytes in flight of a row-by-row scan
  for (int m0 = 0; m0 < K; m0 += 4) {
    ArgMax best[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) best[r] = ArgMax{-INFINITY, 0x7fffffff};
    if (vec) {
      for (int j = tid; j < V / 4; j += kAggThreads) {
        float4 x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (m0 + r < K) x[r] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(m0 + r) * ld) + j);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (m0 + r < K) {
            best[r] = better(best[r], ArgMax{x[r].x, 4 * j});
            best[r] = better(best[r], ArgMax{x[r].y, 4 * j + 1});
            best[r] = better(best[r], ArgMax{x[r].z, 4 * j + 2});
            best[r] = better(best[r], ArgMax{x[r].w, 4 * j + 3});
          }
      }
    } else {
      for (int j = tid; j < V; j += kAggThreads)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (m0 + r < K) best[r] = better(best[r], ArgMax{base[(size_t)(m0 + r) * ld + j], j});
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (m0 + r < K) {  // block-uniform
        const ArgMax b4 = block_argmax(best[r], s_warp);
        if (tid == 0) {
          s_ids[m0 + r] = b4.i;
          if (member_ids) member_ids[b * K + m0 + r] = b4.i;
        }
      }
    }
  }
  __syncthreads();

  if (mode == DD_AGG_VOTE) {
    if (tid == 0) {
      // Counter.most_common(1): highest count, ties -> the id inserted first (models/llava.py:29)
      int best_cnt = 0, best_id = s_ids[0];
      for (int m = 0; m < K; ++m) {
        int c = 0;
        for (int j = 0; j < K; ++j) c += (s_ids[j] == s_ids[m]);
        if (c > best_cnt) { best_cnt = c; best_id = s_ids[m]; }
      }
      int w = 0;
      while (s_ids[w] != best_id) ++w;  // first member with that id (models/llava.py:31-33)
      s_win = w;
      winner[b] = w;
      token[b] = best_id;
    }
    if (out_logits) {
      __syncthreads();
      const float* row = base + (size_t)s_win * ld;
      for (int j = tid; j < V; j += kAggThreads) out_logits[(size_t)b * V + j] = row[j];
    }
  } else {  // DD_AGG_MEAN: numpy mean(axis=0) in fp32 = sequential row sum, then / K
    ArgMax best{-INFINITY, 0x7fffffff};
    const float inv_k = (float)K;
    for (int j = tid; j < V; j += kAggThreads) {
      float s = base[j];
      for (int m = 1; m < K; ++m) s = __fadd_rn(s, base[(size_t)m * ld + j]);
      s = __fdiv_rn(s, inv_k);
      out_logits[(size_t)b * V + j] = s;
      best = better(best, ArgMax{s, j});
    }
    best = block_argmax(best, s_warp);
    if (tid == 0) {
      winner[b] = 0;  // member 0's output carries the mean (models/llava.py:39,50)
      token[b] = best.i;
    }
  }
}

// a13: per-step bookkeeping of the reference on the chosen logits row of every sequence.
//   stats = { 1 / max(logits)                                     (models/llavanext.py:584,593)
//             entropy (bits), varentropy                          (calculate_entropy_varentropy, models/llava.py:687-708) }
// Three passes over the row (L2-resident): max + sum-exp, entropy, varentropy -- the reference's own order of
// operations (log_softmax, p = exp(lp), H = -sum p lp / ln2, sum p (lp/ln2 + H)^2).
__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kAggThreads / 32; ++w) t += s_red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int w = 0; w < kAggThreads / 32; ++w) t = fmaxf(t, s_red[w]);
  return t;
}

__global__ void __launch_bounds__(kAggThreads) token_stats_kernel(const float* __restrict__ logits, long long seq_stride,
                                                                  long long member_stride,
                                                                  const int32_t* __restrict__ member, int V,
                                                                  const int32_t* __restrict__ slot, int slot_per_seq,
                                                                  int slot_offset, int n_slots, int B,
                                                                  float* __restrict__ stats) {
  __shared__ float s_red[kAggThreads / 32];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int row = (slot ? (slot_per_seq ? slot[b] : *slot) : 0) + slot_offset;
  if (row < 0 || row >= n_slots) return;
  if (member && member[b] < 0) return;  // block-uniform: this sequence's row lives in another buffer
  const float* x = logits + (size_t)b * seq_stride + (size_t)(member ? member[b] : 0) * member_stride;
  float mx = -INFINITY;
  for (int j = tid; j < V; j += kAggThreads) mx = fmaxf(mx, x[j]);
  mx = block_max(mx, s_red);
  float z = 0.f;
  for (int j = tid; j < V; j += kAggThreads) z += expf(x[j] - mx);
  z = block_sum(z, s_red);
  const float lse = mx + logf(z);
  constexpr float kInvLn2 = 1.4426950408889634f;
  float h = 0.f;
  for (int j = tid; j < V; j += kAggThreads) {
    const float lp = x[j] - lse;
    h += expf(lp) * lp;
  }
  const float entropy = -block_sum(h, s_red) * kInvLn2;
  float ve = 0.f;
  for (int j = tid; j < V; j += kAggThreads) {
    const float lp = x[j] - lse, t = lp * kInvLn2 + entropy;
    ve += expf(lp) * t * t;
  }
  ve = block_sum(ve, s_red);
  if (tid == 0) {
    float* o = stats + ((size_t)row * B + b) * 3;
    o[0] = 1.0f / mx;
    o[1] = entropy;
    o[2] = ve;
  }
}

}  // namespace dd

static int token_stats_impl(const float* logits, long long seq_stride, long long member_stride, const int32_t* member,
                            int B, int V, const int32_t* slot, int slot_per_seq, int slot_offset, int n_slots,
                            float* stats, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(logits && stats, "dd_token_stats: null pointer");
  DD_CHECK_ARG(B >= 0 && V >= 1 && n_slots >= 1, "dd_token_stats: bad shape B=%d V=%d n_slots=%d", B, V, n_slots);
  if (B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(token_stats_kernel, dim3(B), dim3(kAggThreads), 0, (cudaStream_t)stream, logits, seq_stride,
                  member_stride, member, V, slot, slot_per_seq, slot_offset, n_slots, B, stats);
  return DD_OK;
}

extern "C" int dd_token_stats(const float* logits, long long seq_stride, long long member_stride, const int32_t* member,
                              int B, int V, const int32_t* slot, int slot_offset, int n_slots, float* stats,
                              dd_stream_t stream) {
  return token_stats_impl(logits, seq_stride, member_stride, member, B, V, slot, 0, slot_offset, n_slots, stats, stream);
}

extern "C" int dd_token_stats_seq(const float* logits, long long seq_stride, long long member_stride,
                                  const int32_t* member, int B, int V, const int32_t* slots, int slot_offset, int n_slots,
                                  float* stats, dd_stream_t stream) {
  return token_stats_impl(logits, seq_stride, member_stride, member, B, V, slots, 1, slot_offset, n_slots, stats, stream);
}

extern "C" int dd_aggregate_strided(const float* logits, int B, int K, int V, int ld, long long seq_stride, int mode,
                                    int32_t* member_ids, int32_t* winner, int32_t* token, float* out_logits,
                                    dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(logits && winner && token, "dd_aggregate: null pointer");
  DD_CHECK_ARG(K >= 1 && K <= DD_MAX_MEMBERS, "dd_aggregate: K=%d out of [1,%d]", K, DD_MAX_MEMBERS);
  DD_CHECK_ARG(V >= 1 && B >= 0 && ld >= V, "dd_aggregate: bad shape B=%d V=%d ld=%d", B, V, ld);
  DD_CHECK_ARG(seq_stride >= (long long)K * ld, "dd_aggregate: seq_stride=%lld < K*ld", seq_stride);
  DD_CHECK_ARG(mode == DD_AGG_VOTE || mode == DD_AGG_MEAN, "dd_aggregate: bad mode %d", mode);
  DD_CHECK_ARG(mode != DD_AGG_MEAN || out_logits, "dd_aggregate: mean mode needs out_logits");
  if (B == 0) return DD_OK;
  DD_LAUNCH_CHAIN(aggregate_kernel, dim3(B), dim3(kAggThreads), 0, (cudaStream_t)stream, logits, K, V, ld, seq_stride,
                  mode, member_ids, winner, token, out_logits);
  return DD_OK;
}

extern "C" int dd_aggregate(const float* logits, int B, int K, int V, int ld, int mode, int32_t* member_ids,
                            int32_t* winner, int32_t* token, float* out_logits, dd_stream_t stream) {
  return dd_aggregate_strided(logits, B, K, V, ld, (long long)K * ld, mode, member_ids, winner, token, out_logits,
                              stream);
}
