// (3) Decode attention for M masked contexts over ONE shared prefix KV cache.
// Replaces the K+1 KV deep copies and the K+1 masked single-query SDPA calls of the reference
// (models/llava.py:292,343 and :294-303,350-359), see include/dropdec.h.
//
// HBM-bound design (algorithmic bytes per sequence/layer/pass = 2*S*H_kv*D*2):
//   grid = (kv splits, H_kv, B); a CTA streams its slice of K and V exactly once through a
//   3-stage ring of TMA tiles (cp.async.bulk.tensor, [64 keys x 64 dims] boxes, SWIZZLE_128B, completion
//   on an mbarrier: one elected thread issues, no thread spends issue slots on addresses) into
//   shared memory and scores it against ALL R = M * (H/H_kv) query rows of that
//   (sequence, kv head) -- members x GQA group -- with mma.sync m16n8k16 (queries are the M=16
//   operand, 16 keys per warp per tile).  Member m's visual-token bitmask is applied to the score
//   fragments in registers; masked keys get -inf exactly like the reference's additive
//   finfo.min mask (their softmax weight is exactly 0).
//   P is split into bf16 hi + lo parts so that P.V is accurate to ~2^-17 relative: the result is
//   then independent of the online-softmax rescaling order and matches an fp32 softmax.
//   Split-KV partials (m, l, unnormalised O) go to a workspace; the LAST CTA of a (sequence, kv head) to
//   finish (ticket counter in the workspace) merges them and folds in each member's own new-token row
//   (k_new / v_new, never written to the cache before the vote) -- one launch per attention call.
#include "dd_attention_body.cuh"

namespace dd {

template <int D, int QT>
__global__ void __launch_bounds__(kAttnThreads) attn_partial_kernel(const __grid_constant__ CUtensorMap tmap_k,
                                                                    const __grid_constant__ CUtensorMap tmap_v,
                                                                    const AttnParams p, const L2Prefetch pf) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (the dynamic smem base is only 16-byte aligned by contract)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t s_full[kStages];
  __shared__ int s_ext[DD_MAX_MEMBERS];                                  // per member: 1 = also attends key S0
  __shared__ int s_last;
  pdl_trigger_stream_entry();
  if ((int)blockIdx.x >= p.num_splits) {
    // prefetch-only CTAs (dd_hint_l2_prefetch): they land on SMs the split grid leaves idle, issue their share of the
    // bulk L2 prefetches and leave -- the TMA queues of the SMs that stream K/V stay free.  (The weights do not
    // depend on earlier kernels: no griddepcontrol.wait.)
    const unsigned extra = gridDim.x - p.num_splits;
    const unsigned idx = (blockIdx.z * gridDim.y + blockIdx.y) * extra + (blockIdx.x - p.num_splits);
    l2_prefetch_slice(pf, idx * kAttnWarps + (threadIdx.x >> 5), extra * gridDim.y * gridDim.z * kAttnWarps);
    return;
  }
  auto wait = [] { pdl_wait(); };
  if (p.M2 > 0 && (int)blockIdx.z >= p.B) {
    // second member window: its own members, key range and split partials; everything else is shared
    AttnParams w = p;
    w.M = p.M2; w.m0 = p.m02; w.seq_lens = p.seq_lens2;
    w.part_o = p.part_o2; w.part_ml = p.part_ml2; w.tickets = p.tickets2;
    attn_partial_body<D, QT, true>(&tmap_k, &tmap_v, w, L2Prefetch{}, blockIdx.x, blockIdx.y, blockIdx.z - p.B,
                                   threadIdx.x, 0u, 1u, smem, s_full, s_ext, &s_last, [] { __syncthreads(); }, wait);
    return;
  }
  attn_partial_body<D, QT, true>(&tmap_k, &tmap_v, p, L2Prefetch{}, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, 0u,
                                 1u, smem, s_full, s_ext, &s_last, [] { __syncthreads(); }, wait);
}

// Unsplit case: out[b,m,h,:] = merge(the partial, own new-token row).  One CTA of D threads per (b,m,h).
template <int D>
__global__ void __launch_bounds__(D) attn_merge_kernel(const AttnParams p) {
  __shared__ float s_part[D / 32];
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, m = blockIdx.y, b = blockIdx.z, d = threadIdx.x;
  if (p.seq_active && p.seq_active[b] == 0 && !(p.seq_active2 && p.seq_active2[b] != 0)) return;
  const int G = p.H / p.H_kv, hkv = h / G, r = m * G + (h % G), R = p.M * G;
  const size_t qoff = (((size_t)b * p.Mt + p.m0 + m) * p.H + h) * D;
  const size_t kvoff = (((size_t)b * p.Mt + p.m0 + m) * p.H_kv + hkv) * D;
  // score of the member's own new token (always attended: it is the query position itself)
  float dot = __bfloat162float(p.q[qoff + d]) * __bfloat162float(p.k_new[kvoff + d]);
  dot = warp_sum(dot);
  if ((d & 31) == 0) s_part[d >> 5] = dot;
  __syncthreads();
  float s_self = 0.f;
#pragma unroll
  for (int w = 0; w < D / 32; ++w) s_self += s_part[w];
  s_self *= p.scale_log2;
  const size_t i = ((size_t)b * p.H_kv + hkv) * R + r;
  const float pm = p.part_ml[i * 2];
  const float mx = fmaxf(s_self, pm);
  const float f_self = exp2f(s_self - mx), f = exp2f(pm - mx);
  const float acc = f_self * __bfloat162float(p.v_new[kvoff + d]) + f * p.part_o[i * D + d];
  const float l = f_self + f * p.part_ml[i * 2 + 1];
  p.out[qoff + d] = __float2bfloat16_rn(acc / l);
}

template <int D, int QT>
static int launch_attn(const AttnParams& p, const L2Prefetch& pf, const void* k_new, const void* v_new, void* out,
                       cudaStream_t st) {
  // K / V of one layer as [B*H_kv][S_max][D]; a box is 64 keys x 64 dims (128 B rows, SWIZZLE_128B)
  CUtensorMap tmap_k, tmap_v;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)p.S_max, (uint64_t)p.B * p.H_kv};
    const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)p.S_max * D * 2};
    const uint32_t box[3] = {64, (uint32_t)kTileKeys, 1};
    if (int rc = make_tmap_bf16_sw128(&tmap_k, p.k_cache, 3, dims, strides, box, "dd_decode_attention(k_cache)")) return rc;
    if (int rc = make_tmap_bf16_sw128(&tmap_v, p.v_cache, 3, dims, strides, box, "dd_decode_attention(v_cache)")) return rc;
  }
  const size_t smem = attn_smem_bytes<D, QT>(p.M > p.M2 ? p.M : p.M2, p.bitmask ? p.mask_words : 0);
  {
    // opt in to > 48 KB dynamic smem: a per-DEVICE attribute, cached per device and instantiation (grow-only)
    static size_t configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > configured[dev]) {
      DD_CHECK_CUDA(cudaFuncSetAttribute(attn_partial_kernel<D, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
      if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
  }
  // with a prefetch hint: extra CTAs along x (about one per SM the split grid leaves idle, at least one per (b, kv head))
  int extra = 0;
  if (pf.bytes[0] | pf.bytes[1]) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int pairs = p.H_kv * p.B, idle = sms - (p.num_splits * pairs) % sms;
    extra = (idle + pairs - 1) / pairs;
    if (extra < 1) extra = 1;
    if (extra > 8) extra = 8;
  }
  DD_LAUNCH_CHAIN(attn_partial_kernel<D, QT>, dim3(p.num_splits + extra, p.H_kv, p.B * (p.M2 > 0 ? 2 : 1)),
                  dim3(kAttnThreads), smem, st, tmap_k, tmap_v, p, pf);
  (void)k_new; (void)v_new; (void)out;
  if (p.num_splits == 1) DD_LAUNCH_CHAIN(attn_merge_kernel<D>, dim3(p.H, p.M, p.B), dim3(D), 0, st, p);
  return DD_OK;
}

}  // namespace dd

extern "C" size_t dd_attention_workspace_bytes(int B, int M, int H, int D, int num_splits) {
  // part_o [B,H_kv,splits,R,D] + part_ml [..,2] with H_kv * R = M * H, + one ticket counter per (b, kv head)
  return (size_t)B * M * H * num_splits * (D + 2) * sizeof(float) + (size_t)B * H * sizeof(int);
}

extern "C" int dd_decode_attention(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                   const void* v_cache, const int32_t* seq_lens, const int32_t* vis_start,
                                   const uint32_t* bitmask, int B, int M, int H, int H_kv, int D, int S_max,
                                   int n_vis, int mask_words, float scale, int num_splits, void* out,
                                   void* workspace, size_t workspace_bytes, dd_stream_t stream) {
  return dd_decode_attention_active(q, k_new, v_new, k_cache, v_cache, seq_lens, vis_start, bitmask, nullptr, B, M, 0,
                                    M, H, H_kv, D, S_max, n_vis, mask_words, scale, num_splits, out, workspace,
                                    workspace_bytes, stream);
}

extern "C" int dd_decode_attention_active(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                          const void* v_cache, const int32_t* seq_lens, const int32_t* vis_start,
                                          const uint32_t* bitmask, const int32_t* seq_active, int B, int M,
                                          int member_offset, int members_total, int H, int H_kv, int D, int S_max,
                                          int n_vis, int mask_words, float scale, int num_splits, void* out,
                                          void* workspace, size_t workspace_bytes, dd_stream_t stream) {
  return dd_decode_attention_windows(q, k_new, v_new, k_cache, v_cache, seq_lens, vis_start, bitmask, seq_active, nullptr,
                                     nullptr, B, M, member_offset, members_total, H, H_kv, D, S_max, n_vis, mask_words,
                                     scale, num_splits, out, workspace, workspace_bytes, stream);
}

extern "C" int dd_decode_attention_windows(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                           const void* v_cache, const int32_t* seq_lens, const int32_t* vis_start,
                                           const uint32_t* bitmask, const int32_t* seq_active,
                                           const int32_t* seq_active2, const int32_t* len_extra, int B, int M,
                                           int member_offset, int members_total, int H, int H_kv, int D, int S_max,
                                           int n_vis, int mask_words, float scale, int num_splits, void* out,
                                           void* workspace, size_t workspace_bytes, dd_stream_t stream) {
  using namespace dd;
  const L2Prefetch pf = take_l2_prefetch();  // one-shot: consumed even when the call is refused below
  DD_CHECK_ARG(!len_extra || num_splits == 1,
               "dd_decode_attention: per-member key windows (len_extra) need num_splits == 1 (the split points of a "
               "sequence depend on its length, so members of different lengths cannot share split partials bit-exactly)");
  DD_CHECK_ARG(member_offset >= 0 && member_offset + M <= members_total,
               "dd_decode_attention: members [%d, %d) outside the %d stored per sequence", member_offset,
               member_offset + M, members_total);
  DD_CHECK_ARG(q && k_new && v_new && k_cache && v_cache && seq_lens && out && workspace,
               "dd_decode_attention: null pointer");
  DD_CHECK_ARG(D == 64 || D == 128, "dd_decode_attention: D=%d unsupported (64 or 128)", D);
  DD_CHECK_ARG(((uintptr_t)k_cache & 15) == 0 && ((uintptr_t)v_cache & 15) == 0,
               "dd_decode_attention: the caches must be 16-byte aligned (TMA)");
  DD_CHECK_ARG(H_kv >= 1 && H % H_kv == 0, "dd_decode_attention: H=%d not a multiple of H_kv=%d", H, H_kv);
  DD_CHECK_ARG(M >= 1 && M <= DD_MAX_MEMBERS && M * (H / H_kv) <= 32,
               "dd_decode_attention: M*(H/H_kv)=%d exceeds 32 query rows per kv head", M * (H / H_kv));
  DD_CHECK_ARG(num_splits >= 1 && num_splits <= 256, "dd_decode_attention: num_splits=%d out of [1,256]",
               num_splits);
  DD_CHECK_ARG(!bitmask || (vis_start && mask_words * 32 >= n_vis && mask_words <= kMaxMaskWords),
               "dd_decode_attention: bitmask needs vis_start and n_vis <= 32*mask_words <= %d",
               32 * kMaxMaskWords);
  // laid out for all members_total members, so the ticket counters sit at the same place for every member window
  if (workspace_bytes < dd_attention_workspace_bytes(B, members_total, H, D, num_splits)) {
    set_error("dd_decode_attention: workspace %zu < %zu bytes", workspace_bytes,
              dd_attention_workspace_bytes(B, members_total, H, D, num_splits));
    return DD_ERR_WORKSPACE;
  }
  if (B == 0) return DD_OK;
  AttnParams p{};
  p.q = (const __nv_bfloat16*)q; p.k_cache = (const __nv_bfloat16*)k_cache; p.v_cache = (const __nv_bfloat16*)v_cache;
  p.seq_lens = seq_lens; p.vis_start = vis_start; p.bitmask = bitmask; p.seq_active = seq_active;
  p.seq_active2 = seq_active2; p.len_extra = len_extra;
  p.part_o = (float*)workspace;
  p.part_ml = p.part_o + (size_t)B * members_total * H * num_splits * D;
  p.tickets = reinterpret_cast<int*>(p.part_ml + (size_t)B * members_total * H * num_splits * 2);
  p.k_new = (const __nv_bfloat16*)k_new; p.v_new = (const __nv_bfloat16*)v_new; p.out = (__nv_bfloat16*)out;
  p.B = B; p.M = M; p.Mt = members_total; p.m0 = member_offset; p.H = H; p.H_kv = H_kv; p.S_max = S_max; p.n_vis = n_vis; p.mask_words = mask_words;
  p.num_splits = num_splits;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int R = M * (H / H_kv);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128)
    return R <= 16 ? launch_attn<128, 1>(p, pf, k_new, v_new, out, st) : launch_attn<128, 2>(p, pf, k_new, v_new, out, st);
  return R <= 16 ? launch_attn<64, 1>(p, pf, k_new, v_new, out, st) : launch_attn<64, 2>(p, pf, k_new, v_new, out, st);
}

extern "C" int dd_decode_attention_two_windows(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                               const void* v_cache, const int32_t* seq_lens_a, const int32_t* seq_lens_b,
                                               const int32_t* vis_start, const uint32_t* bitmask, int B, int members_a,
                                               int members_b, int H, int H_kv, int D, int S_max, int n_vis,
                                               int mask_words, float scale, int num_splits, void* out, void* workspace,
                                               size_t workspace_bytes, dd_stream_t stream) {
  using namespace dd;
  const L2Prefetch pf = take_l2_prefetch();
  const int Mt = members_a + members_b;
  DD_CHECK_ARG(q && k_new && v_new && k_cache && v_cache && seq_lens_a && seq_lens_b && out && workspace,
               "dd_decode_attention_two_windows: null pointer");
  DD_CHECK_ARG(D == 64 || D == 128, "dd_decode_attention_two_windows: D=%d unsupported (64 or 128)", D);
  DD_CHECK_ARG(((uintptr_t)k_cache & 15) == 0 && ((uintptr_t)v_cache & 15) == 0,
               "dd_decode_attention_two_windows: the caches must be 16-byte aligned (TMA)");
  DD_CHECK_ARG(H_kv >= 1 && H % H_kv == 0, "dd_decode_attention_two_windows: H=%d not a multiple of H_kv=%d", H, H_kv);
  DD_CHECK_ARG(members_a >= 1 && members_b >= 1 && Mt <= DD_MAX_MEMBERS && members_a * (H / H_kv) <= 32 &&
                   members_b * (H / H_kv) <= 32,
               "dd_decode_attention_two_windows: windows of %d and %d members (at most 32 query rows per kv head each)",
               members_a, members_b);
  DD_CHECK_ARG(num_splits >= 2 && num_splits <= 256,
               "dd_decode_attention_two_windows: num_splits=%d out of [2,256] (unsplit: dd_decode_attention_windows with "
               "len_extra)", num_splits);
  DD_CHECK_ARG(!bitmask || (vis_start && mask_words * 32 >= n_vis && mask_words <= kMaxMaskWords),
               "dd_decode_attention_two_windows: bitmask needs vis_start and n_vis <= 32*mask_words <= %d",
               32 * kMaxMaskWords);
  const size_t one = (dd_attention_workspace_bytes(B, Mt, H, D, num_splits) + 255) / 256 * 256;
  if (workspace_bytes < 2 * one) {
    set_error("dd_decode_attention_two_windows: workspace %zu < %zu bytes", workspace_bytes, 2 * one);
    return DD_ERR_WORKSPACE;
  }
  if (B == 0) return DD_OK;
  AttnParams p{};
  p.q = (const __nv_bfloat16*)q; p.k_cache = (const __nv_bfloat16*)k_cache; p.v_cache = (const __nv_bfloat16*)v_cache;
  p.seq_lens = seq_lens_a; p.vis_start = vis_start; p.bitmask = bitmask;
  const size_t n_o = (size_t)B * Mt * H * num_splits * D, n_ml = (size_t)B * Mt * H * num_splits * 2;
  p.part_o = (float*)workspace;
  p.part_ml = p.part_o + n_o;
  p.tickets = reinterpret_cast<int*>(p.part_ml + n_ml);
  p.part_o2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + one);
  p.part_ml2 = p.part_o2 + n_o;
  p.tickets2 = reinterpret_cast<int*>(p.part_ml2 + n_ml);
  p.k_new = (const __nv_bfloat16*)k_new; p.v_new = (const __nv_bfloat16*)v_new; p.out = (__nv_bfloat16*)out;
  p.B = B; p.M = members_a; p.Mt = Mt; p.m0 = 0; p.H = H; p.H_kv = H_kv; p.S_max = S_max; p.n_vis = n_vis;
  p.mask_words = mask_words; p.num_splits = num_splits;
  p.M2 = members_b; p.m02 = members_a; p.seq_lens2 = seq_lens_b;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int R = (members_a > members_b ? members_a : members_b) * (H / H_kv);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128)
    return R <= 16 ? launch_attn<128, 1>(p, pf, k_new, v_new, out, st) : launch_attn<128, 2>(p, pf, k_new, v_new, out, st);
  return R <= 16 ? launch_attn<64, 1>(p, pf, k_new, v_new, out, st) : launch_attn<64, 2>(p, pf, k_new, v_new, out, st);
}
