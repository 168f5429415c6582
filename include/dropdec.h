/*
 * dropdec.h -- C ABI of libdropdec.so: the B200 (sm_100a) kernels of the
 * Dropout-Decoding inference hot path.
 *
 * The reference (kigb/DropoutDecoding) has no FFI: its hot path is Python inside
 * three HuggingFace model wrappers.  The entry points below are what a maintainer
 * would bind (ctypes) to replace the cited reference code.
 * All `file:line` citations are relative to the reference tree.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - bf16 tensors are passed as `const void*` (2 bytes/element), row-major, dense;
 *   - the caller allocates all outputs and workspaces (query the *_workspace_bytes
 *     function first); functions are stateless and re-entrant;
 *   - work is enqueued on `stream` (a cudaStream_t); nothing synchronises the host,
 *     so every entry point is CUDA-graph capturable;
 *   - return value: 0 on success, a negative DD_ERR_* code otherwise (never throws);
 *     dd_last_error() gives the message for the calling thread.
 */
#ifndef DROPDEC_H
#define DROPDEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dd_stream_t; /* cudaStream_t */

#define DD_OK 0
#define DD_ERR_INVALID_ARG (-1)
#define DD_ERR_UNSUPPORTED (-2)
#define DD_ERR_WORKSPACE (-3)
#define DD_ERR_CUDA (-4)
#define DD_ERR_OVERFLOW (-5) /* a bounded on-device list overflowed (dd_uncertainty_project candidate lists) */

#define DD_MAX_MEMBERS 16  /* K, the number of masked ensemble members        */
#define DD_MAX_TOPK 16     /* k of the per-visual-token top-k id table         */

/* mask sampler modes */
#define DD_MASK_CUMULATIVE 0  /* LLaVA-1.5: mask not restored between members (models/llava.py:342-346) */
#define DD_MASK_INDEPENDENT 1 /* LLaVA-NeXT: restored between members (models/llavanext.py:544-551)     */
#define DD_MASK_QUANTILE 2    /* InstructBLIP: deterministic quantile threshold (models/instructblip.py:447-460) */

/* aggregation modes */
#define DD_AGG_VOTE 0 /* select_by_vote    models/llava.py:22-36 */
#define DD_AGG_MEAN 1 /* select_by_average models/llava.py:37-52 */

int dd_version(void);
const char* dd_last_error(void);

/* Programmatic dependent launch for the kernels of a decoder pass (host choice, process-wide; returns the previous
 * mode).  1: every dd_* kernel launched from now on may become resident while its predecessor in the stream is still
 * draining and run its independent prologue -- dd_stream_gemm puts its whole ring of weight tiles in flight --
 * before it waits for the predecessor's results (griddepcontrol.wait).  2: only the skinny weight GEMMs
 * (dd_skinny_gemm / dd_decoder_linear) are launched that way: their first weight batch and launch latency overlap the
 * small kernel in front of them (RMSNorm, attention), while no small kernel queues up behind a GEMM's tail.  3: the
 * opposite -- every kernel EXCEPT the skinny weight GEMMs: the small kernels' launch latency overlaps the drain of the
 * GEMM in front of them (the engine's choice for passes of <= 16 rows).  0: off.
 * Results are identical in every mode.  Default: the DD_PDL environment variable ("1" / "2" / "3"), else off.  No reference
 * counterpart (the reference's forward is a sequence of stream-ordered torch kernels, models/llama.py:16-69). */
int dd_set_chain_launch(int mode);

/* L2 prefetch hint for the NEXT dd_rmsnorm / dd_decode_attention* call of this host thread (one-shot: that call
 * consumes it, also when it fails).  In a pass of a few rows the decoder is one long HBM stream of weights that the
 * latency-bound kernels interrupt (attention over ~10 MB of K/V, RMSNorm over a few rows): the hinted call's kernel
 * also issues cp.async.bulk.prefetch.L2 over up to two address ranges -- the head of the weights of the GEMMs that
 * follow it -- so the stream keeps running into the 126 MB L2 while HBM would idle.  Results are unaffected.
 * Ranges: 16-byte aligned, a multiple of 16 bytes long, readable device memory; bytes = 0 disables a range.
 * No reference counterpart (models/llama.py:16-69 is a sequence of stream-ordered torch kernels). */
int dd_hint_l2_prefetch(const void* range0, size_t bytes0, const void* range1, size_t bytes1);

/* ------------------------------------------------------------------------------------------
 * (1) Visual-token uncertainty stage.
 * Replaces: lm_head over the visual rows + get_image_logits (models/llava.py:412-426),
 * get_topk_token_id (models/llava.py:428-441) and calculate_vision_uncertainty
 * (models/llava.py:710-756; twins llavanext.py:878-924, instructblip.py:511-557).
 *
 * logits[i,v] = <h_vis[img,i,:], w_lm[v,:]> is produced ONCE, tile by tile on the tcgen05 tensor
 * cores (TMA -> smem -> UMMA -> TMEM); the epilogue keeps the softmax statistics and parks the logits
 * (bf16 if round_logits_bf16, else fp32) in a scratch inside the workspace that element-wise kernels
 * then walk for the entropy / KL terms -- no fp32 [n_vis x vocab] temporaries, 37 MB per
 * LLaVA-1.5 image (L2-sized), at most ~0.6 GB per call whatever the batch.  Per image:
 *   p = softmax(logits), p_avg[v] = mean_i p[i,v]
 *   epi[i]  = sum_v p (log(p+1e-10) - log(p_avg+1e-10))
 *   alea[i] = -sum_v p log(p+1e-10)
 *   var[i]  = unbiased variance of p[i,:]
 *   topk_ids[i,:] = ids of the topk largest logits of row i (descending)
 * round_logits_bf16 != 0 rounds each logit to bf16 before the softmax, which is what the
 * reference's `lm_head(...)` in half precision followed by `.float()` does
 * (models/llama.py:46-47).
 * Limits: d % 64 == 0, n_vis <= 8192 per image, topk <= DD_MAX_TOPK.
 * Top-k ids come from a per-row candidate list of bounded size; a row whose list overflows (more than
 * 96*topk + 64 logits tied at or above its threshold) is counted in an int32 at byte offset
 * dd_uncertainty_overflow_offset(...) of the workspace, valid once the stream has reached the end of the call.
 * Callers must treat a non-zero count as an error (DD_ERR_OVERFLOW): the ids of such a row may miss a tied id.
 * ------------------------------------------------------------------------------------------ */
size_t dd_uncertainty_workspace_bytes(int n_images, int n_vis, int vocab, int topk);
size_t dd_uncertainty_overflow_offset(int n_images, int n_vis, int vocab, int topk);
int dd_uncertainty_project(const void* h_vis /* bf16 [n_images, n_vis, d] */,
                           const void* w_lm /* bf16 [vocab, d] */, int n_images, int n_vis, int d,
                           int vocab, int topk, int round_logits_bf16,
                           float* epi /* [n_images, n_vis] */, float* alea, float* var,
                           int32_t* topk_ids /* [n_images, n_vis, topk] */, void* workspace,
                           size_t workspace_bytes, dd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) Uncertainty-guided mask sampling.
 * Replaces: get_overlap_image_tokens (models/llava.py:443-482) + get_image_attention_mask
 * method "epis" / "epis_no_overlap" (models/llava.py:589-683, llavanext.py:779-829) and the
 * InstructBLIP quantile variant (models/instructblip.py:447-460), for all K members at once.
 *
 * Output bit (b,m,i) = 1  <=>  visual token i of sequence b is hidden from member m.
 *   p_i    = 0.1f + (float)(prob_m - 0.1) * (epi_i - min) / (max - min)     (fp32, op by op)
 *   draw_m = { i : u[b,m,i] < p_i }
 *   cumulative : bits_m = (draw_0 | .. | draw_m) \ relevant
 *   independent: bits_m = draw_m \ relevant
 *   quantile   : bits_m = { i : epi_i >= quantile(epi, 1 - prob_m) } \ relevant
 *   relevant   = { i : y_init[b] in topk_ids[b,i,:] }   (skipped when keep_relevant == 0)
 * u comes from `uniforms` when it is non-NULL (parity mode: the caller supplies the same
 * numbers the reference's torch.rand_like would draw), otherwise from Philox4x32-10 keyed
 * by (seed, offset0 + *step * K, seq_base + sequence, member, token): seq_base is the global index of this
 * call's first sequence, so a batch sharded over several GPUs draws the same numbers as on one.
 * n_masked[b,m] = popcount(bits_m)   (the reference's masked_numbers, models/llava.py:661).
 * ------------------------------------------------------------------------------------------ */
int dd_sample_masks(const float* epi /* [B, n_vis] */, const int32_t* topk_ids /* [B, n_vis, topk] */,
                    int topk, const int32_t* y_init /* [B] */, const double* probs_host /* [K] */,
                    int K, const float* uniforms /* [B, K, n_vis] or NULL */, uint64_t seed,
                    uint64_t offset0, const int32_t* step /* device scalar or NULL */, int mode,
                    int keep_relevant, int B, int n_vis, int seq_base, uint32_t* bitmask /* [B, K, mask_words] */,
                    int mask_words, int32_t* n_masked /* [B, K] */, dd_stream_t stream);

/* The same with one step counter PER SEQUENCE (seq_steps [B]): the asynchronous batch schedule lets sequences that
 * had to repeat their masked contexts fall behind the others, and sequence b then draws the numbers of ITS step. */
int dd_sample_masks_seq(const float* epi, const int32_t* topk_ids, int topk, const int32_t* y_init,
                        const double* probs_host, int K, const float* uniforms, uint64_t seed, uint64_t offset0,
                        const int32_t* seq_steps /* [B] */, int mode, int keep_relevant, int B, int n_vis, int seq_base,
                        uint32_t* bitmask, int mask_words, int32_t* n_masked, dd_stream_t stream);

/* Test hook: out[i] = first uniform of Philox subsequence i at `offset` -- the layout
 * torch.rand(n, device="cuda") uses for n <= 256 * #blocks. */
int dd_philox_uniform(uint64_t seed, uint64_t offset, int n, float* out, dd_stream_t stream);

/* Host-decision flags of the one-sequence speculative schedule in one launch (no reference counterpart: the reference
 * takes these decisions with `.item()` calls, models/llava.py:22-36, 443-482):
 *   flags[0] = 1 iff tok_a[b] == tok_b[b] and win_a[b] == win_b[b] for every b < B (NULL tok_a: 1) -- the guessed vote of
 *              the lookahead pass was the real one;
 *   flags[1] = 1 iff any of the n_words mask words differ (NULL mask_a: 0) -- the speculated masks were not the real
 *              ones (y_init restored a drawn visual token), i.e. the step has to be repeated. */
int dd_schedule_flags(const int32_t* tok_a, const int32_t* tok_b, const int32_t* win_a, const int32_t* win_b, int B,
                      const int32_t* mask_a, const int32_t* mask_b, int n_words, int32_t* flags /* [2] */,
                      dd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (3) Decode attention for the M masked contexts over one shared prefix KV cache.
 * Replaces: the K+1 `copy.deepcopy(past_key_values)` (models/llava.py:292,343) and the masked
 * SDPA of the K+1 single-token forwards (models/llava.py:294-303,350-359 -> HF Llama attention).
 *
 * The prefix K/V of (sequence b, kv head h) is read from HBM ONCE and scored against the
 * queries of all M members (and all H/H_kv query heads of the group); member m's bitmask is
 * applied to the scores of the visual span [vis_start[b], vis_start[b]+n_vis) in registers.
 * Each member also attends to its own new-token row (k_new/v_new), which is not in the cache.
 *   out[b,m,h,:] = softmax( [q.K_prefix^T masked_m , q.k_new[m]] * scale ) [V_prefix ; v_new[m]]
 * bitmask == NULL means no masking (the preliminary pass, M = 1).
 * Limits: D in {64,128}; M * (H / H_kv) <= 32.
 * Workspace: split-KV partials plus one ticket counter per (sequence, kv head) at its end.  The caller zeroes
 * the workspace ONCE after allocating it; every call leaves the counters at zero again.
 * ------------------------------------------------------------------------------------------ */
size_t dd_attention_workspace_bytes(int B, int M, int H, int D, int num_splits);
int dd_decode_attention(const void* q /* bf16 [B, M, H, D] */, const void* k_new /* bf16 [B, M, H_kv, D] */,
                        const void* v_new, const void* k_cache /* bf16 [B, H_kv, S_max, D] */,
                        const void* v_cache, const int32_t* seq_lens /* [B] committed tokens */,
                        const int32_t* vis_start /* [B] */, const uint32_t* bitmask /* [B, M, mask_words] */,
                        int B, int M, int H, int H_kv, int D, int S_max, int n_vis, int mask_words,
                        float scale, int num_splits, void* out /* bf16 [B, M, H, D] */, void* workspace,
                        size_t workspace_bytes, dd_stream_t stream);

/* The same with (a) a per-sequence activity flag: sequences with seq_active[b] == 0 are skipped entirely (no K/V is
 * read, their rows of `out` stay untouched; used by the asynchronous batch schedule for sequences that wait), and
 * (b) a member window: q, k_new, v_new, out and bitmask hold `members_total` members per sequence and this call
 * serves members [member_offset, member_offset + M) -- a pass with more than 32 query rows per kv head (LLaVA-NeXT:
 * 9 contexts x 4 query heads) is served by two calls; the workspace is sized (and laid out) for members_total:
 * dd_attention_workspace_bytes(B, members_total, ...).  seq_active == NULL: all active. */
int dd_decode_attention_active(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                               const void* v_cache, const int32_t* seq_lens, const int32_t* vis_start,
                               const uint32_t* bitmask, const int32_t* seq_active /* [B] or NULL */, int B, int M,
                               int member_offset, int members_total, int H, int H_kv, int D, int S_max, int n_vis,
                               int mask_words, float scale, int num_splits, void* out, void* workspace,
                               size_t workspace_bytes, dd_stream_t stream);
/* dd_decode_attention_active with TWO member windows in one launch (the lookahead schedules: "verify" contexts of a
 * repeated step next to the speculative contexts of the following step): the sequence is processed when
 * seq_active[b] or seq_active2[b] is set, and member m attends keys [0, seq_lens[b] + len_extra[b, m]) with
 * len_extra int32 [B, members_total] in {0, 1} -- the speculative contexts also see the row committed ahead at
 * seq_lens[b].  The prefix K/V is still read ONCE for all members.  Needs num_splits == 1. */
int dd_decode_attention_windows(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                const void* v_cache, const int32_t* seq_lens, const int32_t* vis_start,
                                const uint32_t* bitmask, const int32_t* seq_active, const int32_t* seq_active2,
                                const int32_t* len_extra, int B, int M, int member_offset, int members_total, int H,
                                int H_kv, int D, int S_max, int n_vis, int mask_words, float scale, int num_splits,
                                void* out, void* workspace, size_t workspace_bytes, dd_stream_t stream);

/* Two member windows in ONE launch with a SPLIT key range (the lookahead pass of one sequence: members [0, members_a) --
 * the repeated step -- attend keys [0, seq_lens_a[b]), members [members_a, members_a + members_b) -- the next step's
 * contexts -- attend keys [0, seq_lens_b[b])).  Each window keeps the split points, partials and merge of its own
 * dd_decode_attention_active call (bit-identical results), but the two grids run side by side instead of one after the
 * other: at one sequence both are latency-bound and a quarter of the lookahead pass.  workspace: 2 x
 * dd_attention_workspace_bytes(B, members_a + members_b, H, D, num_splits), each rounded up to 256 bytes.  num_splits >= 2
 * (unsplit ranges: dd_decode_attention_windows with len_extra). */
int dd_decode_attention_two_windows(const void* q, const void* k_new, const void* v_new, const void* k_cache,
                                    const void* v_cache, const int32_t* seq_lens_a, const int32_t* seq_lens_b,
                                    const int32_t* vis_start, const uint32_t* bitmask, int B, int members_a, int members_b,
                                    int H, int H_kv, int D, int S_max, int n_vis, int mask_words, float scale,
                                    int num_splits, void* out, void* workspace, size_t workspace_bytes, dd_stream_t stream);

/* Commit the winning member's new-token K/V rows of every layer into the shared cache at
 * position seq_lens[b] (replaces "the winner's cache becomes the live cache",
 * models/llava.py:370-376).  seq_lens is NOT advanced here (see dd_step_advance).  winner[b] < 0 skips sequence b
 * (it commits from another staging buffer). */
int dd_commit_winner(const void* k_new_all /* bf16 [L, B, M, H_kv, D] */, const void* v_new_all,
                     const int32_t* winner /* [B] */, const int32_t* seq_lens /* [B] */,
                     void* k_cache_all /* bf16 [L, B, H_kv, S_max, D] */, void* v_cache_all, int L, int B,
                     int M, int H_kv, int D, int S_max, dd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (4) Ensemble aggregation.
 * Replaces: select_by_vote (models/llava.py:22-36, llavanext.py:26-39) and select_by_average
 * (models/llava.py:37-52); with K = 1 it is the plain argmax that yields y_init.
 *   member_ids[b,m] = argmax_v logits[b,m,:]          (first index on ties, as torch.argmax)
 *   vote: winner[b] = first member whose id is the most frequent one (ties between ids ->
 *         the id seen first); token[b] = that id; out_logits[b,:] = logits[b,winner,:]
 *   mean: out_logits[b,:] = mean_m logits[b,m,:] (fp32); token = argmax; winner = 0
 * out_logits may be NULL.  ld >= V is the row pitch of `logits` in elements (a vocabulary padded to a
 * GEMM-friendly size: only the first V entries of each row take part).
 * ------------------------------------------------------------------------------------------ */
int dd_aggregate(const float* logits /* [B, K, ld] */, int B, int K, int V, int ld, int mode,
                 int32_t* member_ids /* [B, K] */, int32_t* winner /* [B] */, int32_t* token /* [B] */,
                 float* out_logits /* [B, V] or NULL */, dd_stream_t stream);

/* The same over a strided view: sequence b's K rows start at logits + b * seq_stride (elements), seq_stride >= K * ld.
 * Used when the K member rows are a slice of a larger [B, K+1, ld] buffer (speculative single pass: row 0 is the
 * unmasked preliminary context, rows 1..K the members). */
int dd_aggregate_strided(const float* logits, int B, int K, int V, int ld, long long seq_stride, int mode,
                         int32_t* member_ids /* [B, K] */, int32_t* winner /* [B] */, int32_t* token /* [B] */,
                         float* out_logits /* [B, V] or NULL */, dd_stream_t stream);

/* Per-token bookkeeping the reference keeps on the wrapper (models/llavanext.py:584,593;
 * models/instructblip.py:141-144,154-157; calculate_entropy_varentropy models/llava.py:687-708), computed on the
 * chosen logits row of every sequence without a host sync:
 *   row(b)  = logits + b * seq_stride + (member ? member[b] : 0) * member_stride      (elements)
 *   stats[r, b, :] = { 1 / max(row), entropy in bits, varentropy }     with r = (slot ? *slot : 0) + slot_offset
 * (nothing is written when r is outside [0, n_slots), nor for sequences with member[b] < 0). */
int dd_token_stats(const float* logits, long long seq_stride, long long member_stride,
                   const int32_t* member /* [B] or NULL */, int B, int V, const int32_t* slot /* device scalar or NULL */,
                   int slot_offset, int n_slots, float* stats /* [n_slots, B, 3] */, dd_stream_t stream);

/* dd_token_stats with a per-sequence slot: r = slots[b] + slot_offset. */
int dd_token_stats_seq(const float* logits, long long seq_stride, long long member_stride,
                       const int32_t* member /* [B] or NULL */, int B, int V, const int32_t* slots /* [B] */,
                       int slot_offset, int n_slots, float* stats /* [n_slots, B, 3] */, dd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Decoder support for the single-token forwards of the ensemble (models/llama.py:27-47 and the
 * HF Llama decoder layer it calls).  Small-M, weight-streaming (HBM-bound) kernels.
 * ------------------------------------------------------------------------------------------ */

/* y[m,:] = bf16( bf16(x[m,:] * rsqrt(mean(x^2)+eps)) * w ) */
int dd_rmsnorm(const void* x /* bf16 [M, d] */, const void* w /* bf16 [d] */, int M, int d, float eps,
               void* y /* bf16 [M, d] */, dd_stream_t stream);

/* Skinny GEMM  y[m,n] = sum_k x[m,k] * w[n,k]   (M <= 32; N % 16 == 0; K % 256 == 0)
 * epilogue: DD_EPI_NONE     y            = bf16(acc)
 *           DD_EPI_RESIDUAL y            = bf16(float(bf16(acc)) + residual[m,n])
 *           DD_EPI_F32      y (float32)  = float(bf16(acc))      (lm_head + .float())
 *           DD_EPI_SWIGLU   w is [2N, K] = [gate ; up];  y = bf16( bf16(silu(bf16(g))) * bf16(u) )
 */
#define DD_EPI_NONE 0
#define DD_EPI_RESIDUAL 1
#define DD_EPI_F32 2
#define DD_EPI_SWIGLU 3
int dd_skinny_gemm(const void* x /* bf16 [M, K] */, const void* w /* bf16 [N(,2N), K] */,
                   const void* residual /* bf16 [M, N] or NULL */, int M, int N, int K, int epilogue,
                   void* y, dd_stream_t stream);

/* The same GEMM and epilogues (DD_EPI_NONE / RESIDUAL / F32 / SWIGLU) for 1 <= M <= 256 rows on the tcgen05
 * tensor cores -- the batched passes of the ensemble (rows = sequences x members; batch 64 with K = 3 members is
 * 64 rows in the preliminary pass and 192 in the masked pass), which the reference runs as 1+K separate
 * `self.language_model(...)` calls (models/llava.py:294-303, 350-359).  Swap-AB (the weights are the M=128 UMMA
 * operand, the token rows the N operand, accumulators in TMEM), TMA-fed, stream-K over (256-feature tile,
 * 64-wide k-block) units so that every SM streams the same number of weight bytes; split tiles are reduced
 * deterministically through `workspace`.
 * Limits: K % 64 == 0, N % 8 == 0 (SwiGLU: N % 128 == 0), x / w 16-byte aligned.
 * Workspace: dd_stream_gemm_workspace_bytes(M, N, K, epilogue) bytes; it starts with the tile counters, which the
 * caller zeroes ONCE after allocating (every call leaves them at zero again); one workspace sized for the largest
 * shape can serve every call on a stream. */
size_t dd_stream_gemm_workspace_bytes(int M, int N, int K, int epilogue);
int dd_stream_gemm(const void* x /* bf16 [M, K] */, const void* w /* bf16 [N(,2N), K] */,
                   const void* residual /* bf16 [M, N] or NULL */, int M, int N, int K, int epilogue, void* y,
                   void* workspace, size_t workspace_bytes, dd_stream_t stream);

/* Fused decoder GEMM: dd_skinny_gemm plus the two fusions that remove the small kernels between the
 * GEMMs of a decoder layer (same rounding points as the separate kernels):
 *   norm_w != NULL : x is replaced on the fly by dd_rmsnorm(x, norm_w, eps)   (input_layernorm /
 *                    post_attention_layernorm / final norm of the HF Llama decoder);
 *   DD_EPI_ROPE    : w = [q ; k ; v] rows, N = (H + 2 H_kv) * D; the bf16 projection is rotated and
 *                    staged exactly as dd_rope_stage does (q_out, k_new, v_new); y is ignored.
 * Row r of x is (sequence r / members, member r % members). */
#define DD_EPI_ROPE 4
typedef struct dd_rope_args {
  const float* cos_tab; /* fp32 [max_pos, D/2] */
  const float* sin_tab;
  const int32_t* seq_lens; /* [B]: position of the new token */
  int members, H, H_kv, D;
  void* q_out; /* bf16 [B, members, H, D]    */
  void* k_new; /* bf16 [B, members, H_kv, D] */
  void* v_new;
  /* optional "commit ahead" of the lookahead pass (guess_winner NULL: off): the new K/V rows of row
   * b * guess_members + guess_winner[b] are ALSO written to row guess_pos[b] of this layer's cache [B, H_kv, S_max, D] --
   * dd_commit_winner of one layer folded into the projection's epilogue (guess_winner[b] outside [0, guess_members) or
   * guess_pos[b] >= S_max: nothing is written for that sequence) */
  const int32_t* guess_winner; /* [B] */
  const int32_t* guess_pos;    /* [B] */
  int guess_members, S_max;
  void* k_cache; /* bf16 [B, H_kv, S_max, D] */
  void* v_cache;
} dd_rope_args;
int dd_decoder_linear(const void* x /* bf16 [M, K] */, const void* norm_w /* bf16 [K] or NULL */, float eps,
                      const void* w /* bf16 [N(,2N), K] */, const void* residual /* bf16 [M, N] or NULL */,
                      int M, int N, int K, int epilogue, void* y, const dd_rope_args* rope /* or NULL */,
                      dd_stream_t stream);

/* RoPE (rotate-half) on the new token's q and k, and staging of the member rows:
 *   qkv [B*M, (H + 2 H_kv) * D] (q | k | v) -> q_out [B,M,H,D], k_new [B,M,H_kv,D], v_new [B,M,H_kv,D]
 * position of sequence b = seq_lens[b]; cos/sin tables are fp32 [max_pos, D/2]. */
int dd_rope_stage(const void* qkv, const float* cos_tab, const float* sin_tab, const int32_t* seq_lens,
                  int B, int M, int H, int H_kv, int D, void* q_out, void* k_new, void* v_new,
                  dd_stream_t stream);

/* x[b*M + m, :] = embed[token[b], :]  for m < M */
int dd_embed_gather(const void* embed /* bf16 [V, d] */, const int32_t* token /* [B] */, int B, int M,
                    int d, int vocab /* rows of embed: ids are clamped to [0, vocab) */,
                    void* x /* bf16 [B*M, d] */, dd_stream_t stream);

/* End of one ensemble step (*step = ensemble steps done so far, 0-based):
 * tokens_out[b, *step + 1] = token[b]  (column 0 holds the un-ensembled first token,
 * models/llava.py:382-388); seq_lens[b] += 1; *step += 1. */
int dd_step_advance(const int32_t* token /* [B] */, int32_t* tokens_out /* [B, max_steps] */,
                    int max_steps, int32_t* seq_lens /* [B] */, int32_t* step /* scalar */, int B,
                    dd_stream_t stream);

/* Per-sequence variant: for every sequence with advance[b] != 0:
 * tokens_out[b, seq_steps[b] + 1] = token[b]; seq_lens[b] += 1; seq_steps[b] += 1.  Others are left alone. */
int dd_step_advance_seq(const int32_t* token /* [B] */, const int32_t* advance /* [B] */,
                        int32_t* tokens_out /* [B, max_steps] */, int max_steps, int32_t* seq_lens /* [B] */,
                        int32_t* seq_steps /* [B] */, int B, dd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Persistent decoder pass: ONE kernel for a whole single-token forward of <= 16 rows (rows = sequences x members) --
 * embedding gather, n_layers x {RMSNorm, QKV + RoPE + staging, masked decode attention, o_proj, RMSNorm, gate/up +
 * SwiGLU, down}, final RMSNorm, lm_head -- i.e. one `self.language_model(...)` call of the reference
 * (models/llava.py:294-303, 350-359; models/llama.py:27-47) for all members at once.  One CTA per SM stays resident;
 * grid barriers replace the ~230 kernel boundaries of the launch chain, the first weight batch of every GEMM phase is
 * put in flight before the barrier, the RMSNorms are computed per CTA into shared memory, idle warps prefetch the next
 * weights into L2 during the attention phase.  Results are bit-identical to the chain dd_embed_gather / dd_rmsnorm /
 * dd_decoder_linear / dd_decode_attention (same arithmetic, same order).
 *
 * Plan: one record per layer in DEVICE memory (weights, staging rows, caches and the caches' TMA descriptors);
 * dd_pass_plan_fill writes record `layer` of a HOST copy (dd_pass_plan_bytes(n_layers) bytes) that the caller then
 * copies to the device.  k_new / v_new: that layer's staging rows [rows, H_kv, D]; k_cache / v_cache: that layer's
 * cache [B, H_kv, S_max, D].
 *
 * Rows: row r embeds token tok[r / tok_div] and rotates with position pos[r / pos_div] (tok_div = pos_div =
 * members_total with per-sequence arrays, 1 with per-row arrays).  Attention: the members of a sequence form 1 or 2
 * windows [win_first, win_first + win_members) that attend keys [0, win_seq_lens[w][b]) of the shared cache under
 * their bitmask rows, plus their own new row (as dd_decode_attention_active with member_offset).  guess_winner (or
 * NULL): the new K/V rows of member guess_winner[b] are also written to cache row guess_pos[b] before the attention
 * (the lookahead's guessed commit, dd_commit_winner of one layer).  logits NULL: no lm_head; xn_out (or NULL) receives
 * the final-norm hidden state [rows, d_model].  workspace: dd_pass_workspace_bytes, ZEROED once by the caller (the
 * kernel leaves its counters at zero).  Needs num_splits >= 2 and win_members * (H / H_kv) <= 16.
 * ------------------------------------------------------------------------------------------ */
typedef struct dd_pass_args {
  const void* plan; int n_layers;
  int rows, members_total;
  int d_model, d_ff, H, H_kv, D, S_max, vocab, vocab_padded;
  float eps, scale;
  const void* embed /* bf16 [vocab, d_model] */; const int32_t* tok; int tok_div; const int32_t* pos; int pos_div;
  const float *cos_tab, *sin_tab;
  void *x, *x2 /* bf16 [rows, d_model] */, *act /* [rows, d_ff] */, *q /* [rows, H, D] */, *attn /* [rows, H, D] */,
      *xn_out;
  const void *final_norm, *lm_head /* bf16 [vocab_padded, d_model] */; float* logits /* [rows, vocab_padded] */;
  const int32_t* vis_start; const uint32_t* bitmask /* [B, members_total, mask_words] or NULL */;
  int n_vis, mask_words, num_splits;
  int n_windows; int win_first[2], win_members[2]; const int32_t* win_seq_lens[2];
  const int32_t *guess_winner, *guess_pos;
  size_t prefetch_gate_up_bytes; /* bytes of gate/up prefetched to L2 during the attention phase (o_proj always is) */
  void* workspace; size_t workspace_bytes;
  void* timeline; /* or NULL: uint64 [2 * (5 * n_layers + 2)] device buffer; CTA 0 stamps %globaltimer (ns) at the start
                     and the end of its work in every phase {embed, L x {qkv, attention, o_proj, gate/up, down}, head} */
} dd_pass_args;
size_t dd_pass_plan_bytes(int n_layers);
int dd_pass_plan_fill(void* host_plan, int layer, const void* ln1, const void* wqkv, const void* wo, const void* ln2,
                      const void* wgu, const void* wd, void* k_new, void* v_new, void* k_cache, void* v_cache, int B,
                      int H_kv, int S_max, int D);
size_t dd_pass_workspace_bytes(int B, int members_total, int H, int D, int num_splits);
int dd_decoder_pass(const dd_pass_args* args, dd_stream_t stream);

/* Member-sharded decode (ensemble members on different GPUs; the reference runs them sequentially on one,
 * models/llava.py:342-361): the finish step after the per-step all-gather.  `recv` holds G exchange records of
 * rank_stride bytes; a record = member argmax ids int32 [B, mpr] at ids_off | token statistics fp32 [mpr, B, 3] at
 * stats_off | staged K rows bf16 [L, B, mpr, H_kv, D] at k_off | staged V rows at v_off, all written by the rank's
 * kernels directly into the collective's send buffer.  Votes over the gathered ids (select_by_vote, models/llava.py:
 * 22-36), writes member_ids [B, G*mpr] / winner [B] / token [B] and token_stats[*step + 1] of the winner, and commits the
 * winner's K/V rows into cache row seq_lens[b] of k_cache_all / v_cache_all [L, B, H_kv, S_max, D].
 * forced_winner >= 0: no vote (logit-mean mode: member 0's cache survives, models/llava.py:39,50), token = forced_token. */
int dd_member_finish(const void* recv, long long rank_stride, long long ids_off, long long stats_off, long long k_off,
                     long long v_off, int G, int mpr, int L, int B, int H_kv, int D, int S_max, int forced_winner,
                     const int32_t* forced_token, const int32_t* seq_lens, const int32_t* step, int n_slots,
                     int32_t* member_ids, int32_t* winner, int32_t* token, float* token_stats, void* k_cache_all,
                     void* v_cache_all, dd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DROPDEC_H */
