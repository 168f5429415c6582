// (2) Uncertainty-guided mask sampling: one CTA per sequence writes the visual-token
// bitmasks of all K ensemble members.  Replaces models/llava.py:443-482 + 589-683 (and the
// llavanext.py / instructblip.py twins), see include/dropdec.h.
//
// HBM traffic is tiny (B*n_vis*(4 + 4k) bytes in, B*K*n_vis/8 out); the point of the kernel is
// to take ~12 micro-kernels, two sorts and 1-2 host syncs per member off the critical path and
// to produce the masks as bits that the attention kernel applies in registers.
#include "dd_common.cuh"

namespace dd {

constexpr int kSamplerThreads = 256;
constexpr int kSamplerMaxChunks = 32;  // n_vis <= 8192

struct SamplerParams {
  const float* epi;
  const int32_t* topk_ids;
  const int32_t* y_init;
  const float* uniforms;
  const int32_t* step;
  int step_per_seq;  // step points at [B] per-sequence counters instead of one scalar
  uint32_t* bitmask;
  int32_t* n_masked;
  uint64_t seed, offset0;
  float coef[DD_MAX_MEMBERS];  // (float)(prob - 0.1)      models/llava.py:646
  float q32[DD_MAX_MEMBERS];   // (float)(1 - prob)        models/instructblip.py:450
  int topk, K, mode, keep_relevant, B, n_vis, mask_words, seq_base;
};

// Philox4x32-10, bit-compatible with cuRAND's curand_init(seed, subsequence, offset) +
// curand_uniform4 as ATen uses them for torch.rand on CUDA.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// first uniform of (seed, subsequence, offset); offset in cuRAND units (4 per Philox block)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t subsequence, uint64_t offset) {
  const uint64_t blk = offset >> 2;
  const uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)subsequence,
                               (uint32_t)(subsequence >> 32));
  const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t lane = (uint32_t)(offset & 3);
  const uint32_t x = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);  // _curand_uniform: (0, 1]
  return u == 1.0f ? 0.0f : u;                                     // ATen: (0,1] -> [0,1)
}

__global__ void __launch_bounds__(kSamplerThreads) sampler_kernel(const SamplerParams p) {
  extern __shared__ float s_epi[];  // [n_vis], quantile mode only
  __shared__ float s_red[2][kSamplerThreads / 32];
  __shared__ float s_below[DD_MAX_MEMBERS], s_above[DD_MAX_MEMBERS];
  __shared__ int s_count[DD_MAX_MEMBERS];

  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n_vis;
  const int chunks = (n + kSamplerThreads - 1) / kSamplerThreads;
  const float* epi = p.epi + (size_t)b * n;

  if (tid < p.K) s_count[tid] = 0;

  // ---- min / max of epi  (torch.quantile(.,0) and (.,1), models/llava.py:641-642) ----
  float lo = INFINITY, hi = -INFINITY;
  for (int i = tid; i < n; i += kSamplerThreads) {
    const float e = epi[i];
    lo = fminf(lo, e);
    hi = fmaxf(hi, e);
    if (p.mode == DD_MASK_QUANTILE) s_epi[i] = e;
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
  __syncthreads();
  lo = s_red[0][0];
  hi = s_red[1][0];
#pragma unroll
  for (int w = 1; w < kSamplerThreads / 32; ++w) {
    lo = fminf(lo, s_red[0][w]);
    hi = fmaxf(hi, s_red[1][w]);
  }
  const float range = __fsub_rn(hi, lo);

  // ---- relevant set: y_init in the token's top-k ids (models/llava.py:459-476) ----
  uint32_t relevant = 0;  // bit c: token c*256+tid is relevant
  if (p.keep_relevant) {
    const int y = p.y_init[b];
    for (int c = 0; c < chunks; ++c) {
      const int i = c * kSamplerThreads + tid;
      if (i < n) {
        const int32_t* ids = p.topk_ids + ((size_t)b * n + i) * p.topk;
        bool hit = false;
        for (int j = 0; j < p.topk; ++j) hit |= (ids[j] == y);
        relevant |= (uint32_t)hit << c;
      }
    }
  }

  // ---- quantile thresholds (models/instructblip.py:450, torch.quantile 'linear') ----
  if (p.mode == DD_MASK_QUANTILE) {
    if (tid < p.K) { s_below[tid] = 0.f; s_above[tid] = 0.f; }
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
      const int i = c * kSamplerThreads + tid;
      if (i >= n) break;
      const float e = s_epi[i];
      int rank = 0;  // stable ascending rank
      for (int j = 0; j < n; ++j) {
        const float f = s_epi[j];
        rank += (f < e) || (f == e && j < i);
      }
      for (int m = 0; m < p.K; ++m) {
        const float r = __fmul_rn(p.q32[m], (float)(n - 1));
        if (rank == (int)r) s_below[m] = e;
        if (rank == (int)ceilf(r)) s_above[m] = e;
      }
    }
    __syncthreads();
  }

  const int step = p.step ? (p.step_per_seq ? p.step[blockIdx.x] : *p.step) : 0;
  uint32_t cumulative = 0;  // bit c: token c*256+tid drawn by this or an earlier member
  for (int m = 0; m < p.K; ++m) {
    float thr = 0.f;
    if (p.mode == DD_MASK_QUANTILE) {
      const float r = __fmul_rn(p.q32[m], (float)(n - 1));
      const float w = __fsub_rn(r, (float)(int)r);
      const float a = s_below[m], z = s_above[m];
      // at::lerp: a + w (z - a) for w < 0.5, else z - (z - a)(1 - w)
      thr = (fabsf(w) < 0.5f) ? fmaf(w, __fsub_rn(z, a), a) : fmaf(-__fsub_rn(z, a), __fsub_rn(1.f, w), z);
    }
    int cnt = 0;
    for (int c = 0; c < chunks; ++c) {
      const int i = c * kSamplerThreads + tid;
      bool drop = false;
      if (i < n) {
        const float e = epi[i];
        if (p.mode == DD_MASK_QUANTILE) {
          drop = e >= thr;
        } else {
          // 0.1 + (prob - 0.1) * (e - min) / (max - min), one rounding per op (no FMA contraction)
          const float pr = __fadd_rn(0.1f, __fdiv_rn(__fmul_rn(p.coef[m], __fsub_rn(e, lo)), range));
          float u;
          if (p.uniforms) {
            u = p.uniforms[((size_t)b * p.K + m) * n + i];
          } else {
            const uint64_t call = (uint64_t)step * p.K + m;  // one rand_like call per member per step
            u = philox_uniform(p.seed, ((uint64_t)(p.seq_base + b) << 32) | (uint32_t)i, p.offset0 + 4ull * call);
          }
          drop = u < pr;  // NaN (max == min) compares false: nothing dropped
        }
        if (p.mode == DD_MASK_CUMULATIVE) {
          cumulative |= (uint32_t)drop << c;
          drop = (cumulative >> c) & 1u;
        }
        drop = drop && !((relevant >> c) & 1u);
      }
      const uint32_t word = __ballot_sync(0xffffffffu, drop);
      const int widx = c * (kSamplerThreads / 32) + warp;
      if (lane == 0 && widx < p.mask_words) {
        p.bitmask[((size_t)b * p.K + m) * p.mask_words + widx] = word;
        cnt += __popc(word);
      }
    }
    // words past the last chunk (mask_words may be padded by the caller)
    for (int widx = chunks * (kSamplerThreads / 32) + tid; widx < p.mask_words; widx += kSamplerThreads)
      p.bitmask[((size_t)b * p.K + m) * p.mask_words + widx] = 0u;
    if (lane == 0 && cnt) atomicAdd(&s_count[m], cnt);
  }
  __syncthreads();
  if (tid < p.K && p.n_masked) p.n_masked[b * p.K + tid] = s_count[tid];
}

__global__ void philox_uniform_kernel(uint64_t seed, uint64_t offset, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_uniform(seed, (uint64_t)i, offset);
}


// Host-decision flags of the one-sequence speculative schedule (schedules.py), one launch instead of ~9 elementwise ones:
//   flags[0] = (tok_a == tok_b) && (win_a == win_b) for every sequence  -- the lookahead's guess was right (1 when NULL)
//   flags[1] = any word of mask_a differs from mask_b                   -- the speculated masks were not the real ones
__global__ void schedule_flags_kernel(const int32_t* __restrict__ tok_a, const int32_t* __restrict__ tok_b,
                                      const int32_t* __restrict__ win_a, const int32_t* __restrict__ win_b, int B,
                                      const int32_t* __restrict__ mask_a, const int32_t* __restrict__ mask_b, int n_words,
                                      int32_t* __restrict__ flags) {
  __shared__ int s_bad, s_diff;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) s_bad = s_diff = 0;
  __syncthreads();
  int bad = 0, diff = 0;
  if (tok_a)
    for (int b = threadIdx.x; b < B; b += blockDim.x) bad |= (tok_a[b] != tok_b[b]) | (win_a[b] != win_b[b]);
  if (mask_a)
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) diff |= mask_a[i] != mask_b[i];
  if (bad) s_bad = 1;
  if (diff) s_diff = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[0] = s_bad ? 0 : 1;
    flags[1] = s_diff;
  }
}

}  // namespace dd

static int sample_masks_impl(const float* epi, const int32_t* topk_ids, int topk, const int32_t* y_init,
                             const double* probs_host, int K, const float* uniforms, uint64_t seed,
                             uint64_t offset0, const int32_t* step, int step_per_seq, int mode, int keep_relevant,
                             int B, int n_vis, int seq_base, uint32_t* bitmask, int mask_words, int32_t* n_masked,
                             dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(epi && bitmask && probs_host, "dd_sample_masks: null pointer");
  DD_CHECK_ARG(K >= 1 && K <= DD_MAX_MEMBERS, "dd_sample_masks: K=%d out of [1,%d]", K, DD_MAX_MEMBERS);
  DD_CHECK_ARG(B >= 0 && n_vis >= 1 && n_vis <= kSamplerThreads * kSamplerMaxChunks,
               "dd_sample_masks: n_vis=%d out of [1,%d]", n_vis, kSamplerThreads * kSamplerMaxChunks);
  DD_CHECK_ARG(mask_words * 32 >= n_vis, "dd_sample_masks: mask_words=%d too small for n_vis=%d", mask_words,
               n_vis);
  DD_CHECK_ARG(mode >= DD_MASK_CUMULATIVE && mode <= DD_MASK_QUANTILE, "dd_sample_masks: bad mode %d", mode);
  DD_CHECK_ARG(!keep_relevant || (topk_ids && y_init && topk >= 1 && topk <= DD_MAX_TOPK),
               "dd_sample_masks: keep_relevant needs topk_ids, y_init and 1 <= topk <= %d", DD_MAX_TOPK);
  if (B == 0) return DD_OK;
  SamplerParams p{};
  p.epi = epi; p.topk_ids = topk_ids; p.y_init = y_init; p.uniforms = uniforms; p.step = step; p.step_per_seq = step_per_seq;
  p.bitmask = bitmask; p.n_masked = n_masked; p.seed = seed; p.offset0 = offset0;
  for (int m = 0; m < K; ++m) {
    p.coef[m] = (float)(probs_host[m] - 0.1);  // python double arithmetic, then the fp32 tensor op
    p.q32[m] = (float)(1.0 - probs_host[m]);
  }
  p.topk = topk; p.K = K; p.mode = mode; p.keep_relevant = keep_relevant; p.B = B; p.n_vis = n_vis;
  p.mask_words = mask_words; p.seq_base = seq_base;
  const size_t smem = mode == DD_MASK_QUANTILE ? (size_t)n_vis * sizeof(float) : 0;
  DD_LAUNCH_CHAIN(sampler_kernel, dim3(B), dim3(kSamplerThreads), smem, (cudaStream_t)stream, p);
  return DD_OK;
}

extern "C" int dd_sample_masks(const float* epi, const int32_t* topk_ids, int topk, const int32_t* y_init,
                               const double* probs_host, int K, const float* uniforms, uint64_t seed,
                               uint64_t offset0, const int32_t* step, int mode, int keep_relevant, int B,
                               int n_vis, int seq_base, uint32_t* bitmask, int mask_words, int32_t* n_masked,
                               dd_stream_t stream) {
  return sample_masks_impl(epi, topk_ids, topk, y_init, probs_host, K, uniforms, seed, offset0, step, 0, mode,
                           keep_relevant, B, n_vis, seq_base, bitmask, mask_words, n_masked, stream);
}

extern "C" int dd_sample_masks_seq(const float* epi, const int32_t* topk_ids, int topk, const int32_t* y_init,
                                   const double* probs_host, int K, const float* uniforms, uint64_t seed,
                                   uint64_t offset0, const int32_t* seq_steps, int mode, int keep_relevant, int B,
                                   int n_vis, int seq_base, uint32_t* bitmask, int mask_words, int32_t* n_masked,
                                   dd_stream_t stream) {
  return sample_masks_impl(epi, topk_ids, topk, y_init, probs_host, K, uniforms, seed, offset0, seq_steps, 1, mode,
                           keep_relevant, B, n_vis, seq_base, bitmask, mask_words, n_masked, stream);
}

// Test hook: the first n uniforms of the (seed, offset) Philox stream, laid out as torch.rand(n)
// lays them out for n <= 256 * #blocks (element i = first output of subsequence i).
extern "C" int dd_philox_uniform(uint64_t seed, uint64_t offset, int n, float* out, dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(out && n >= 0, "dd_philox_uniform: bad arguments");
  if (n == 0) return DD_OK;
  philox_uniform_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, n, out);
  DD_CHECK_LAUNCH();
  return DD_OK;
}

extern "C" int dd_schedule_flags(const int32_t* tok_a, const int32_t* tok_b, const int32_t* win_a, const int32_t* win_b,
                                 int B, const int32_t* mask_a, const int32_t* mask_b, int n_words, int32_t* flags,
                                 dd_stream_t stream) {
  using namespace dd;
  DD_CHECK_ARG(flags, "dd_schedule_flags: null flags");
  DD_CHECK_ARG(!tok_a || (tok_b && win_a && win_b && B >= 1), "dd_schedule_flags: the guess check needs both pairs");
  DD_CHECK_ARG(!mask_a || (mask_b && n_words >= 1), "dd_schedule_flags: the mask check needs both masks");
  DD_LAUNCH_CHAIN(schedule_flags_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, tok_a, tok_b, win_a, win_b, B, mask_a,
                  mask_b, n_words, flags);
  return DD_OK;
}
