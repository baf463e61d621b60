/* mrmt3_b200 -- C ABI of the B200-native MR-MT3 transcription hot path.
 *
 * The reference (gudgud96/MR-MT3) is pure Python and has no FFI layer; its boundary for this
 * path is the Python object API listed in SURVEY.md section 8(b).  Each entry point below names
 * the reference interface it stands behind (file:line in the reference checkout).  The Python
 * classes in `mr-mt3_b200/` (same names, arguments and state-dict keys as the reference's) bind
 * these with ctypes; see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `mrmt3_last_error` gives the
 *     message.  No C++ exception crosses this boundary.
 *   - one handle per device; a handle is NOT thread-safe (the reference is single-threaded).
 *   - the caller owns every input/output buffer; the library owns its packed bf16 weight copy,
 *     the KV-cache pages and its workspaces (inside the handle).
 *   - pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it.  Functions that run
 *     the greedy loop synchronise that stream internally when they poll for early exit and
 *     before they return (the reference syncs once per token, models/t5.py:294).
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef MRMT3_B200_H_
#define MRMT3_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrmt3_handle mrmt3_handle;

/* mem_variant */
#define MRMT3_MEM_NONE 0        /* models/t5.py T5ForConditionalGeneration               */
#define MRMT3_MEM_V1_PREPEND 1  /* models/t5_segmem.py T5SegMem.generate_2                */
#define MRMT3_MEM_V2_APPEND 2   /* models/t5_segmem_v2_with_prev.py T5SegMemV2WithPrev    */

/* flags of mrmt3_logmel / mrmt3_transcribe_host */
#define MRMT3_MEL_NORM 1        /* inference.py:113-118 clip [-12,5] -> [0,1]             */

typedef struct mrmt3_config {
    int32_t d_model;          /* 512  (pretrained/config.json)                             */
    int32_t n_heads;          /* 6                                                         */
    int32_t d_kv;             /* 64                                                        */
    int32_t d_ff;             /* 1024                                                      */
    int32_t vocab;            /* 1536                                                      */
    int32_t n_enc_layers;     /* 8                                                         */
    int32_t n_dec_layers;     /* 8                                                         */
    int32_t mem_variant;      /* MRMT3_MEM_*                                               */
    int32_t n_mem_layers;     /* segmem_num_layers (models/t5_segmem.py:52), 0 when NONE   */
    int32_t mem_len;          /* segmem_length L_agg (models/t5_segmem.py:53)              */
    int32_t start_id;         /* decoder_start_token_id = 0                                */
    int32_t eos_id;           /* 1                                                         */
    int32_t pad_id;           /* 0                                                         */
    float ln_eps;             /* layer_norm_epsilon = 1e-6                                 */
} mrmt3_config;

/* ---- lifetime ------------------------------------------------------------------------- */
/* Replaces: T5ForConditionalGeneration(config) / T5SegMem*(config, segmem_num_layers,
 * segmem_length) + .cuda()  (models/t5.py:46-80, models/t5_segmem.py:47-66,
 * inference.py:56,183). */
int mrmt3_create(const mrmt3_config* cfg, int device, mrmt3_handle** out);
void mrmt3_destroy(mrmt3_handle* h);
/* message of the last failure on this handle (h == NULL: last mrmt3_create failure) */
const char* mrmt3_last_error(const mrmt3_handle* h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t mrmt3_launch_count(const mrmt3_handle* h);

/* Tuning knobs (all have working defaults).  key: "group_lanes" = lanes per concurrently
 * decoding lane group (0 = one group, negative = by batch size: the default), "use_graphs" = replay the decode step as a CUDA graph,
 * "attn_variant" = decode attention kernel (1 = persistent TMA ring + mma.sync, the default;
 * 0 = one CTA per (lane, head) on CUDA cores), "attn_ring_stages" = ring depth of variant 1
 * (2, 3, 4 or 6 stages of 16 KB per warp quartet), "attn_ring_quartets" = 1 or 2 math-warp
 * quartets per CTA, "attn_ring_ctas" = its persistent CTAs per SM (1..8, 0 = by launch size).  The
 * attention settings are process-wide.  "attn_part_keys_self" / "attn_part_keys_cross" = split-key work
 * units of variant 1: an item's keys are cut at fixed multiples of this many keys (0 = off, else a
 * multiple of 128) and the parts are merged in order by the last finisher.  "fuse_greedy" = 1 (default):
 * the vocabulary projection, the arg-max, the EOS bookkeeping and the next step's embedding lookup run as
 * ONE kernel and the logits never reach memory; 0 = separate lm_head / arg-max / embed kernels.
 * "hooks_fast_path" = 1 sends calls that use the parity hooks of
 * mrmt3_generate / mrmt3_generate_segmem (forced_ids, logits_out) through the production decode path
 * (CUDA-graph replay, concurrent lane groups) instead of the eager single-group debug path.
 * "gemm_2cta" = 1 (default): the large-M tcgen05 GEMMs (encoder, cross-K/V, teacher-forced decoder,
 * fine-tune forward and data gradients) run on two-CTA clusters with tcgen05.mma.cta_group::2 tiles of
 * 256 x BN; 0 = single-CTA 128 x BN tiles (bit-identical results); -1 = the library default.
 * "attn_full_tc" = 1 selects the tcgen05 whole-sequence attention forward (validated, slower; default 0).
 * Both are process-wide. */
int mrmt3_set_option(mrmt3_handle* h, const char* key, int value);

/* ---- per-kernel timing (bench.py's roofline leg) ---------------------------------------
 * While enabled, the decode loop launches eagerly (no CUDA graph) and brackets every kernel
 * with CUDA events on the launching stream; mrmt3_profile_read returns, per kernel class, the
 * summed device time in ms and the launch count since mrmt3_profile_enable(h, 1), and resets
 * them.  Never enabled inside a timed region. */
#define MRMT3_PROF_EMBED 0
#define MRMT3_PROF_RMSNORM 1
#define MRMT3_PROF_GEMM_QKV 2
#define MRMT3_PROF_ATTN_SELF 3
#define MRMT3_PROF_GEMM_O 4
#define MRMT3_PROF_GEMM_CQ 5
#define MRMT3_PROF_ATTN_CROSS 6
#define MRMT3_PROF_GEMM_CO 7
#define MRMT3_PROF_GEMM_WI 8
#define MRMT3_PROF_GEMM_WFF 9
#define MRMT3_PROF_LM_HEAD 10
#define MRMT3_PROF_ARGMAX 11
#define MRMT3_PROF_NCAT 12
int mrmt3_profile_enable(mrmt3_handle* h, int on);
int mrmt3_profile_read(mrmt3_handle* h, double* ms_out, int64_t* launches_out, int n_cat);

/* ---- in-graph timeline trace (profiles/, never inside a timed region) --------------------
 * With tracing on, every decode-step kernel stamps the GPU global timer (ns) when its CTA 0
 * starts and when its last CTA ends; the stamps of the most recent step are read back as
 * out[2*i] = begin, out[2*i+1] = end for kernel i of the step (embed, then per layer: qkv,
 * self-attention, o, cross-q, cross-attention, o, ffn-in, ffn-out, then lm_head, arg-max).
 * Enabling/disabling drops the captured step graphs.  Returns the number of slots written. */
int mrmt3_trace_enable(mrmt3_handle* h, int on);
int mrmt3_trace_read(mrmt3_handle* h, uint64_t* out, int max_slots);

/* Test hook: C (M,N) fp32 = A (M,K) bf16 * W (N,K)^T bf16 through one of the library's GEMM
 * kernels: which = 0 mma.sync pipeline, 1 = TMA + tcgen05/TMEM, 2 = decode-step single-shot
 * kernel (K in {384, 512, 1024}), 3 = tcgen05 with the bf16 store epilogue (c holds M*N bf16).
 * The two operand forms of the fine-tune backward, on ROW-major operands as they lie in memory:
 * which = 4: C (M,N) = A^T W with A (K,M), W (K,N) -- both MN-major, reduction over the K rows,
 * split-K chosen as the weight-gradient path chooses it; which = 5: C (M,N) = A W with A (M,K),
 * W (K,N) -- the data-gradient form (B operand MN-major); which = 6: kernel 3 with the epilogue's
 * global stores dropped (c is not written; measurement only).  Used by tests/test_kernels_gpu.py and
 * scripts/gpu_gemm_2cta_check.py only. */
int mrmt3_test_gemm(mrmt3_handle* h, const void* a_bf16, const void* w_bf16, int M, int N, int K,
                    float* c_f32, int which, void* stream);

/* ---- weights -------------------------------------------------------------------------- */
/* Replaces: model.load_state_dict(sd) (test.py:106-110, train.py:80-83).  `name` is the
 * reference's state-dict key (weight contract); `data` is fp32 row-major
 * (rows, cols) on the host or on the device (norm weights / inv_freq: rows = 1).  The library
 * converts to its packed bf16 layout.  Returns 3 for a key it does not know. */
int mrmt3_set_weight(mrmt3_handle* h, const char* name, const float* data, int rows, int cols);
/* after the last mrmt3_set_weight: checks that every tensor the configuration needs is there */
int mrmt3_commit_weights(mrmt3_handle* h);
/* Optional: dense (1025, 512) fp32 HOST mel filterbank; default is the exact-arithmetic HTK
 * table, the Python binding passes torchaudio's fp32 table (contrib/spectrograms.py:130-139). */
int mrmt3_set_mel_filterbank(mrmt3_handle* h, const float* fb_host);

/* ---- frontend --------------------------------------------------------------------------
 * Replaces: InferenceHandler._compute_spectrograms + the pad zeroing of _preprocess
 * (inference.py:97-127) = spectrograms.compute_spectrogram per 256-frame segment
 * (contrib/spectrograms.py:105-145).
 * Segment i reads audio[seg_start[i] : seg_start[i]+seg_len[i]]; samples of its 32768+1920
 * sample window at or past seg_len read as zeros (seg_len <= 32768 reproduces the reference's
 * per-segment transform; up to 34688 lets the last frames see the following samples, which is
 * how compute_spectrogram on a longer signal is tiled).  valid_frames[i] = `paddings[i]`; rows at or
 * past it are zero.  Outputs (n_seg, 256, 512): out_f32 and/or out_bf16 (either may be NULL). */
int mrmt3_logmel(mrmt3_handle* h, const float* audio, const int64_t* seg_start,
                 const int32_t* seg_len, const int32_t* valid_frames, int n_seg, int flags,
                 float* out_f32, void* out_bf16, void* stream);

/* ---- encoder ---------------------------------------------------------------------------
 * Replaces: self.proj + self.encoder(...) (models/t5.py:253-258, T5Stack :507-702).
 * mel (B,256,512) fp32 -> enc_out (B,256,512) fp32 (final-normed encoder states). */
int mrmt3_encode(mrmt3_handle* h, const float* mel, int B, float* enc_out, void* stream);

/* ---- greedy transcription --------------------------------------------------------------
 * Replaces: T5ForConditionalGeneration.generate (models/t5.py:251-302).
 * mel (B,256,512) fp32 -> out_ids (B, max_length+1) int64, col 0 = start token, rows padded
 * with pad_id after their EOS; *steps_host = number of decode steps the reference loop would
 * have run (its output is out_ids[:, :1+steps]).
 * Debug / parity hooks (both may be NULL): forced_ids (B, max_length+1) int64 feeds these
 * tokens instead of the arg-max (teacher forcing through the KV-cached step kernels);
 * logits_out (B, max_length, vocab) fp32 receives every step's logits. */
int mrmt3_generate(mrmt3_handle* h, const float* mel, int B, int max_length, int64_t* out_ids,
                   int32_t* steps_host, const int64_t* forced_ids, float* logits_out,
                   void* stream);

/* Replaces: T5SegMemV2WithPrev.generate (models/t5_segmem_v2_with_prev.py:226-297) and
 * T5SegMem.generate_2 (models/t5_segmem.py:172-252), batched ACROSS tracks: the segments of
 * one track are decoded in order (each one's memory block is built from the previous output
 * row), and all tracks advance one segment per round so the decode batch is n_tracks wide.
 * mel (S_total,256,512) fp32, tracks concatenated; seg_counts_host[n_tracks] segments per
 * track (sum = S_total).  out_ids (S_total, max_length) int64 in the same order.
 * n_tracks = 1 is exactly the reference call on one (S,256,512) batch.
 * logits_out (optional): (S_total, max_length, vocab) fp32. */
int mrmt3_generate_segmem(mrmt3_handle* h, const float* mel, const int32_t* seg_counts_host,
                          int n_tracks, int max_length, int64_t* out_ids, float* logits_out,
                          void* stream);

/* Parity hook of mrmt3_generate_segmem (tests only): forced_ids (S_total, max_length+1) int64 feeds
 * these tokens instead of the arg-max, row by row, so that every segment's memory block is built
 * from the caller's (the oracle's) previous row and each step's logits can be compared at every
 * position of every segment even after a low-margin token flip.  Rows are run for all max_length
 * steps (no early exit).  Same reference lines as mrmt3_generate_segmem. */
int mrmt3_generate_segmem_forced(mrmt3_handle* h, const float* mel, const int32_t* seg_counts_host,
                                 int n_tracks, int max_length, const int64_t* forced_ids,
                                 int64_t* out_ids, float* logits_out, void* stream);

/* ---- teacher-forced forward ------------------------------------------------------------
 * Replaces: model.forward / get_model_outputs (models/t5.py:99-249,
 * models/t5_segmem_v2_with_prev.py:60-224): logits (B, L, vocab) fp32 for
 * decoder_input_ids (B, L) int64 (= _shift_right(labels)); targets_prev (B, Lp) int64 with
 * -100 already replaced by pad (NULL when mem_variant == NONE).
 * MRMT3_MEM_V1_PREPEND handles (T5SegMem, models/t5_segmem.py:68-170): targets_prev holds every row's
 * segmem ids (row i = row i-1's decoder input without its start token + a trailing 0; row 0 = [1, 0, ...],
 * t5_segmem.py:123-131); the min(mem_len, Lp) memory rows are prepended to the row's decoder input and the
 * logits of the L token rows are returned; L + memory rows must fit the 1024-entry positional table. */
int mrmt3_forward_logits(mrmt3_handle* h, const float* mel, int B, const int64_t* decoder_input_ids,
                         int L, const int64_t* targets_prev, int Lp, float* logits_out,
                         void* stream);

/* memory block alone (models/t5_segmem_v2_with_prev.py:121-123): prev_ids (B, Lp) int64 ->
 * mem_out (B, min(mem_len, Lp), 512) fp32 */
int mrmt3_memory_block(mrmt3_handle* h, const int64_t* prev_ids, int B, int Lp, float* mem_out,
                       void* stream);

/* ---- fine-tune step ---------------------------------------------------------------------
 * Replaces: one `training_step` of tasks/mt3_net.py / mt3_net_segmem_v2_with_prev.py (forward
 * models/t5.py:182-249 or models/t5_segmem_v2_with_prev.py:155-224 -> CrossEntropyLoss(
 * ignore_index=-100) -> loss.backward() -> AdamW.step()), for the MT3 and V2WithPrev models.
 * Dropout (the reference's config.dropout_rate at its six sites: stack input, attention weights,
 * sublayer outputs, FFN inner, final norm output; never in the memory encoder,
 * models/t5_segmem.py:64) is off until mrmt3_train_set_dropout(p, seed); the masks come from a
 * counter-based hash of (seed, site, element index), regenerated in the backward pass, a new seed
 * being derived after every forward.  Every trainable tensor lives in ONE flat fp32 order
 * (the library's packed layouts); the caller owns the flat gradient buffer, so data-parallel
 * training is train_forward + train_backward on every rank, one all-reduce (mean) of the flat
 * buffer, train_apply on every rank.
 *   mrmt3_train_init     allocates fp32 masters and Adam moments; *n_params = flat length
 *   mrmt3_train_locate   where a reference state-dict tensor `name` (rows, cols) sits in the flat
 *                        order: element (r, c) at flat[offset + ((r*row_mul + row_off)*cols + c)]
 *   mrmt3_train_forward  mel (B,256,512) fp32, decoder_input_ids / labels (B,L) int64, for
 *                        V2WithPrev also targets_prev (B,Lp) int64 with -100 already replaced by
 *                        pad (else NULL, 0) -> logits_out (B,L,vocab) fp32, mean loss in *loss_host;
 *                        the label count and the loss are reduced on the device: loss_host != NULL
 *                        costs the call's only stream synchronisation, NULL keeps it asynchronous
 *   mrmt3_train_backward gradient w.r.t. every trainable tensor -> grad_flat (fp32): of the built-in
 *                        loss when dlogits == NULL, else back-propagates the caller's dlogits
 *                        (B,L,vocab) fp32 (torch autograd over the logits, any loss)
 *   mrmt3_train_apply    AdamW (torch.optim.AdamW semantics) with grad_flat; refreshes the bf16
 *                        weights used by every other entry point
 *   mrmt3_train_read_master  flat fp32 copy of the current parameters */
int mrmt3_train_init(mrmt3_handle* h, int64_t* n_params);
int mrmt3_train_set_dropout(mrmt3_handle* h, float p, uint64_t seed);
/* Host only, no GPU, no handle: keep_out[i] = 1 iff flat element i of the tensor `tensor_id`
 * ((stack << 16) | (layer << 8) | site) survives dropout p under step seed `seed` -- the mask
 * definition the kernels evaluate, exposed so that CPU tests can pin the oracle's mirror of it. */
int mrmt3_dropout_keep_host(float p, uint64_t seed, uint32_t tensor_id, int64_t n, uint8_t* keep_out);
int mrmt3_train_locate(mrmt3_handle* h, const char* name, int64_t* offset, int32_t* rows, int32_t* cols,
                       int32_t* row_mul, int32_t* row_off);
int mrmt3_train_forward(mrmt3_handle* h, const float* mel, int B, const int64_t* decoder_input_ids,
                        const int64_t* labels, int L, const int64_t* targets_prev, int Lp, float* logits_out,
                        float* loss_host, void* stream);
int mrmt3_train_backward(mrmt3_handle* h, float* grad_flat, const float* dlogits, void* stream);
int mrmt3_train_apply(mrmt3_handle* h, const float* grad_flat, float lr, float beta1, float beta2, float eps,
                      float weight_decay, void* stream);
int mrmt3_train_read_master(mrmt3_handle* h, float* out_flat, void* stream);
/* Data-parallel overlap (reference: DDP's bucketed all-reduce under `train.py`'s Lightning trainer,
 * config/config.yaml:45-46).  The flat order is laid out in the order the backward FINISHES the
 * gradients, so it splits into contiguous buckets (lm_head | decoder layers last to first | stacked
 * cross K/V | memory encoder + segmem_proj | encoder layers last to first | proj, embedding, norms).
 * mrmt3_train_backward records one event per bucket on its stream; mrmt3_train_wait_bucket makes
 * another stream wait for bucket i of the most recent backward, so the caller can all-reduce
 * grad_flat[offset : offset + count] on a side stream while the backward is still running.
 * mrmt3_train_loss fetches the mean loss of the last forward (synchronises `stream`). */
int mrmt3_train_bucket_count(mrmt3_handle* h, int32_t* n);
int mrmt3_train_bucket(mrmt3_handle* h, int i, int64_t* offset, int64_t* count);
int mrmt3_train_wait_bucket(mrmt3_handle* h, int i, void* stream);
int mrmt3_train_loss(mrmt3_handle* h, float* loss_host, void* stream);

/* ---- end to end from host buffers ------------------------------------------------------
 * Replaces: InferenceHandler.inference up to the token rows (inference.py:149-191): pinned or
 * pageable HOST audio in, HOST token rows out; H2D, log-mel, encode, greedy decode and D2H all
 * inside the call.  Segment tables as in mrmt3_logmel but on the host.  When the handle has a
 * memory variant, seg_counts_host/n_tracks describe the tracks (as mrmt3_generate_segmem) and
 * out_ids_host is (n_seg, max_length); otherwise seg_counts_host may be NULL and out_ids_host
 * is (n_seg, max_length+1) with *steps_host set. */
int mrmt3_transcribe_host(mrmt3_handle* h, const float* audio_host, int64_t n_samples,
                          const int64_t* seg_start_host, const int32_t* seg_len_host,
                          const int32_t* valid_frames_host, int n_seg,
                          const int32_t* seg_counts_host, int n_tracks, int flags, int max_length,
                          int64_t* out_ids_host, int32_t* steps_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRMT3_B200_H_ */
