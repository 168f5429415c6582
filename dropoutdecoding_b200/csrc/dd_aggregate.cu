// (4) Ensemble aggregation: fused per-member argmax + max-vote (or logit mean) in one launch,
// no host sync.  Replaces select_by_vote / select_by_average (models/llava.py:22-52), whose
// 2K `.item()` calls stall the host every token.
//
// HBM-bound: reads B*K*V fp32 once (+ writes B*V when out_logits is requested).
//
// A sequence is served by a thread-block CLUSTER of kAggCluster CTAs, each scanning one slice of the vocabulary (a single
// CTA per sequence took 13-22 us for K x 32064 logits -- latency-bound on one SM).  The per-slice results meet in the
// shared memory of the cluster's CTAs (distributed shared memory, one cluster barrier per exchange): no workspace, no
// second launch, results independent of the cluster size for the argmax / vote (ties go to the lower index whatever the
// merge order) and combined in rank order for the sums.
#include "dd_common.cuh"

namespace dd {

constexpr int kAggThreads = 512;
constexpr int kAggCluster = 8;

__device__ __forceinline__ uint32_t agg_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void agg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared-memory object of this CTA) in CTA `cta` of the cluster
template <typename T>
__device__ __forceinline__ uint32_t agg_remote(const T* local, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_remote(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_remote(uint32_t addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_remote_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// launch one cluster of kAggCluster CTAs per sequence: grid (kAggCluster, B)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_agg_cluster(void (*kernel)(KArgs...), int B, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kAggCluster, B);
  cfg.blockDim = dim3(kAggThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kAggCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
  __shared__ float s_pv[kAggCluster][DD_MAX_MEMBERS];   // rank 0: the slices' per-member partial maxima ...
  __shared__ int s_pi[kAggCluster][DD_MAX_MEMBERS];     // ... and their indices
  __shared__ int s_ids[DD_MAX_MEMBERS];
  __shared__ int s_win;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, tid = threadIdx.x;
  const uint32_t rank = agg_rank();
  const float* base = logits + (size_t)b * seq_stride;
  const bool vec = (V % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  // this CTA's slice of the vocabulary: [j0, j1) in units of 4 elements (vec) or 1 element
  const int units = vec ? V / 4 : V, per = (units + kAggCluster - 1) / kAggCluster;
  const int j0 = min(units, (int)rank * per), j1 = min(units, j0 + per);
  const uint32_t pv0 = agg_remote(&s_pv[0][0], 0), pi0 = agg_remote(&s_pi[0][0], 0);
  agg_cluster_sync();   // every CTA of the cluster is running before anyone touches distributed shared memory

  {
    // per-member argmax over the slice (both modes publish the member ids), four rows at a time: the four loads of
    // an iteration are independent
    for (int m0 = 0; m0 < K; m0 += 4) {
      ArgMax best[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) best[r] = ArgMax{-INFINITY, 0x7fffffff};
      if (vec) {
        for (int j = j0 + tid; j < j1; j += kAggThreads) {
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
        for (int j = j0 + tid; j < j1; j += kAggThreads)
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (m0 + r < K) best[r] = better(best[r], ArgMax{base[(size_t)(m0 + r) * ld + j], j});
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (m0 + r < K) {  // block-uniform
          const ArgMax b4 = block_argmax(best[r], s_warp);
          if (tid == 0) {
            st_remote(pv0 + (rank * DD_MAX_MEMBERS + m0 + r) * 4, b4.v);
            st_remote(pi0 + (rank * DD_MAX_MEMBERS + m0 + r) * 4, b4.i);
          }
        }
      }
    }
    agg_cluster_sync();   // #1: every slice's partials are in rank 0's shared memory
    if (rank == 0 && tid == 0) {
      for (int m = 0; m < K; ++m) {
        ArgMax r{s_pv[0][m], s_pi[0][m]};
        for (int c = 1; c < kAggCluster; ++c) r = better(r, ArgMax{s_pv[c][m], s_pi[c][m]});
        s_ids[m] = r.i;
        if (member_ids) member_ids[b * K + m] = r.i;
      }
    }
  }
  if (mode == DD_AGG_VOTE) {
    if (rank == 0 && tid == 0) {
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
    if (out_logits) {     // (kernel-uniform) the winner's row, every CTA its slice
      agg_cluster_sync(); // #2: rank 0 has published the winner
      const int w = ld_remote_s32(agg_remote(&s_win, 0));
      const float* row = base + (size_t)w * ld;
      const int e0 = vec ? 4 * j0 : j0, e1 = vec ? 4 * j1 : j1;
      for (int j = e0 + tid; j < e1; j += kAggThreads) out_logits[(size_t)b * V + j] = row[j];
      agg_cluster_sync(); // #3: nobody reads rank 0's shared memory after it may have exited
    }
  } else {  // DD_AGG_MEAN: numpy mean(axis=0) in fp32 = sequential row sum, then / K
    ArgMax best{-INFINITY, 0x7fffffff};
    const float inv_k = (float)K;
    const int e0 = vec ? 4 * j0 : j0, e1 = vec ? 4 * j1 : j1;
    for (int j = e0 + tid; j < e1; j += kAggThreads) {
      float sum = base[j];
      for (int m = 1; m < K; ++m) sum = __fadd_rn(sum, base[(size_t)m * ld + j]);
      sum = __fdiv_rn(sum, inv_k);
      out_logits[(size_t)b * V + j] = sum;
      best = better(best, ArgMax{sum, j});
    }
    best = block_argmax(best, s_warp);
    agg_cluster_sync();   // rank 0 has consumed the member partials: their slots carry the mean's partials now
    if (tid == 0) {
      st_remote(pv0 + (rank * DD_MAX_MEMBERS) * 4, best.v);
      st_remote(pi0 + (rank * DD_MAX_MEMBERS) * 4, best.i);
    }
    agg_cluster_sync();
    if (rank == 0 && tid == 0) {
      ArgMax r{s_pv[0][0], s_pi[0][0]};
      for (int c = 1; c < kAggCluster; ++c) r = better(r, ArgMax{s_pv[c][0], s_pi[c][0]});
      winner[b] = 0;  // member 0's output carries the mean (models/llava.py:39,50)
      token[b] = r.i;
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

// cluster-wide reduction: every CTA's block result goes to slot `rank` of ALL CTAs' `s_x` (distributed shared memory),
// one cluster barrier, then each CTA combines the kAggCluster values in rank order -- deterministic.
template <bool kMax>
__device__ __forceinline__ float cluster_reduce(float v, float* s_red, float* s_x, uint32_t rank) {
  v = kMax ? block_max(v, s_red) : block_sum(v, s_red);
  if (threadIdx.x < kAggCluster) st_remote(agg_remote(s_x + rank, threadIdx.x), v);
  agg_cluster_sync();
  float t = s_x[0];
#pragma unroll
  for (int c = 1; c < kAggCluster; ++c) t = kMax ? fmaxf(t, s_x[c]) : t + s_x[c];
  return t;
}

__global__ void __launch_bounds__(kAggThreads) token_stats_kernel(const float* __restrict__ logits, long long seq_stride,
                                                                  long long member_stride,
                                                                  const int32_t* __restrict__ member, int V,
                                                                  const int32_t* __restrict__ slot, int slot_per_seq,
                                                                  int slot_offset, int n_slots, int B,
                                                                  float* __restrict__ stats) {
  __shared__ float s_red[kAggThreads / 32];
  __shared__ float s_x[4][kAggCluster];   // one exchange slot per reduction
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, tid = threadIdx.x;
  const uint32_t rank = agg_rank();
  const int row = (slot ? (slot_per_seq ? slot[b] : *slot) : 0) + slot_offset;
  if (row < 0 || row >= n_slots) return;          // cluster-uniform
  if (member && member[b] < 0) return;            // cluster-uniform: this sequence's row lives in another buffer
  const float* x = logits + (size_t)b * seq_stride + (size_t)(member ? member[b] : 0) * member_stride;
  agg_cluster_sync();   // every CTA of the cluster is running before anyone touches distributed shared memory
  const int per = (V + kAggCluster - 1) / kAggCluster, j0 = min(V, (int)rank * per), j1 = min(V, j0 + per);
  float mx = -INFINITY;
  for (int j = j0 + tid; j < j1; j += kAggThreads) mx = fmaxf(mx, x[j]);
  mx = cluster_reduce<true>(mx, s_red, s_x[0], rank);
  float z = 0.f;
  for (int j = j0 + tid; j < j1; j += kAggThreads) z += expf(x[j] - mx);
  z = cluster_reduce<false>(z, s_red, s_x[1], rank);
  const float lse = mx + logf(z);
  constexpr float kInvLn2 = 1.4426950408889634f;
  float h = 0.f;
  for (int j = j0 + tid; j < j1; j += kAggThreads) {
    const float lp = x[j] - lse;
    h += expf(lp) * lp;
  }
  const float entropy = -cluster_reduce<false>(h, s_red, s_x[2], rank) * kInvLn2;
  float ve = 0.f;
  for (int j = j0 + tid; j < j1; j += kAggThreads) {
    const float lp = x[j] - lse, t = lp * kInvLn2 + entropy;
    ve += expf(lp) * t * t;
  }
  ve = cluster_reduce<false>(ve, s_red, s_x[3], rank);
  if (rank == 0 && tid == 0) {
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
  DD_CHECK_CUDA(launch_agg_cluster(token_stats_kernel, B, (cudaStream_t)stream, logits, seq_stride, member_stride,
                                   member, V, slot, slot_per_seq, slot_offset, n_slots, B, stats));
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
  DD_CHECK_CUDA(launch_agg_cluster(aggregate_kernel, B, (cudaStream_t)stream, logits, K, V, ld, seq_stride, mode,
                                   member_ids, winner, token, out_logits));
  return DD_OK;
}

extern "C" int dd_aggregate(const float* logits, int B, int K, int V, int ld, int mode, int32_t* member_ids,
                            int32_t* winner, int32_t* token, float* out_logits, dd_stream_t stream) {
  return dd_aggregate_strided(logits, B, K, V, ld, (long long)K * ld, mode, member_ids, winner, token, out_logits,
                              stream);
}
