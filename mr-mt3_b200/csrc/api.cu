// extern "C" boundary (include/mrmt3_b200.h).  Status -> int + message; nothing throws across.
#include <algorithm>
#include <new>

#include "model.cuh"

namespace mrmt3 {
Status handle_init(mrmt3_handle* h);
void handle_destroy(mrmt3_handle* h);
Status set_weight(mrmt3_handle* h, const std::string& name, const float* data, int rows, int cols);
Status commit_weights(mrmt3_handle* h);
Status api_encode(mrmt3_handle* h, const float* mel, int B, float* enc_out, cudaStream_t s);
Status generate_base(mrmt3_handle* h, const float* mel_f32, const bf16* mel_bf16, int B, int max_length,
                     long long* out_ids, int* steps_host, const long long* forced, float* logits_out,
                     cudaStream_t s);
Status generate_segmem(mrmt3_handle* h, const float* mel_f32, const bf16* mel_bf16, const int* seg_counts,
                       int n_tracks, int max_length, long long* out_ids, float* logits_out, cudaStream_t s,
                       const long long* forced);
Status api_memory_block(mrmt3_handle* h, const long long* prev_ids, int B, int Lp, float* mem_out, cudaStream_t s);
Status api_forward_logits(mrmt3_handle* h, const float* mel, int B, const long long* dec_ids, int L,
                          const long long* targets_prev, int Lp, float* logits_out, cudaStream_t s);
Status api_logmel(mrmt3_handle* h, const float* audio, const long long* seg_start, const int* seg_len,
                  const int* valid_frames, int n_seg, int flags, float* out_f32, bf16* out_bf16, cudaStream_t s);
Status api_transcribe_host(mrmt3_handle* h, const float* audio_host, long long n_samples,
                           const long long* seg_start_host, const int* seg_len_host,
                           const int* valid_frames_host, int n_seg, const int* seg_counts_host, int n_tracks,
                           int flags, int max_length, long long* out_ids_host, int* steps_host, cudaStream_t s);
Status train_init(mrmt3_handle* h);
size_t train_param_count(mrmt3_handle* h);
Status train_read_master(mrmt3_handle* h, float* out, cudaStream_t s);
Status train_set_dropout(mrmt3_handle* h, float p, unsigned long long seed);
Status train_set_dropout_sites(mrmt3_handle* h, int mask);
Status train_locate(mrmt3_handle* h, const std::string& name, long long* offset, int* rows, int* cols, int* row_mul,
                    int* row_off);
Status train_forward(mrmt3_handle* h, const float* mel, int B, const long long* dec_ids, const long long* labels, int L,
                     const long long* targets_prev, int Lp, float* logits_out, float* loss_host, cudaStream_t s);
Status train_backward(mrmt3_handle* h, float* grad, const float* dlogits_f32, cudaStream_t s);
Status train_apply(mrmt3_handle* h, const float* grad, float lr, float beta1, float beta2, float adam_eps, float wd,
                   cudaStream_t s);
int train_bucket_count(mrmt3_handle* h);
Status train_bucket(mrmt3_handle* h, int i, long long* offset, long long* count);
Status train_wait_bucket(mrmt3_handle* h, int i, cudaStream_t stream);
Status train_loss(mrmt3_handle* h, float* loss_host, cudaStream_t s);
Status profile_collect(mrmt3_handle* h);
Status trace_enable(mrmt3_handle* h, bool on);
void drop_graphs(mrmt3_handle* h);
Status test_gemm(mrmt3_handle* h, const bf16* A, const bf16* W, int M, int N, int K, float* C, int which, cudaStream_t s);
}  // namespace mrmt3

using namespace mrmt3;

static std::string g_create_error;

static int finish(mrmt3_handle* h, const Status& st) {
    if (st.ok()) return 0;
    if (h) h->err = st.msg;
    // leave the CUDA error state clean for the caller's next call
    cudaGetLastError();
    return st.code ? st.code : 1;
}

#define GUARD(h)                         \
    if (!(h)) return 1;                  \
    try {
#define END_GUARD(h)                                             \
    }                                                            \
    catch (const std::exception& e) {                            \
        (h)->err = std::string("C++ exception: ") + e.what();    \
        return 1;                                                \
    }                                                            \
    catch (...) {                                                \
        (h)->err = "unknown C++ exception";                      \
        return 1;                                                \
    }

#pragma GCC visibility push(default)
extern "C" {

int mrmt3_create(const mrmt3_config* cfg, int device, mrmt3_handle** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return 1;
    }
    *out = nullptr;
    mrmt3_handle* h = new (std::nothrow) mrmt3_handle();
    if (!h) {
        g_create_error = "out of host memory";
        return 1;
    }
    try {
        h->cfg = *cfg;
        h->device = device;
        Status st = handle_init(h);
        if (!st.ok()) {
            g_create_error = st.msg;
            cudaGetLastError();
            handle_destroy(h);
            delete h;
            return st.code ? st.code : 1;
        }
    } catch (const std::exception& e) {
        g_create_error = std::string("C++ exception: ") + e.what();
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

void mrmt3_destroy(mrmt3_handle* h) {
    if (!h) return;
    try {
        handle_destroy(h);
    } catch (...) {
    }
    delete h;
}

const char* mrmt3_last_error(const mrmt3_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t mrmt3_launch_count(const mrmt3_handle* h) { return h ? h->launches : 0; }

int mrmt3_set_option(mrmt3_handle* h, const char* key, int value) {
    GUARD(h)
    if (!key) return finish(h, Error(1, "null key"));
    std::string k(key);
    if (k == "group_lanes") h->group_lanes = value;
    else if (k == "use_graphs") h->use_graphs = value != 0;
    else if (k == "group_serial") h->group_serial = value != 0;
    else if (k == "hooks_fast_path") h->hooks_fast_path = value != 0;
    else if (k == "fuse_greedy") {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        drop_graphs(h);
        h->fuse_greedy = value != 0;
    }
    else if (k == "attn_full_tc") attn_full_configure(value);
    else if (k == "gemm_2cta") {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        drop_graphs(h);  // the encoder GEMMs of a captured graph keep the kernel they were captured with
        gemm_set_2cta(value);
    }

    else if (k == "attn_part_keys_self" || k == "attn_part_keys_cross") {
        const bool self = k == "attn_part_keys_self";
        if (value < 0) value = self ? kDefaultPartKeysSelf : kDefaultPartKeysCross;   // -1: the library default
        if (value % 128 != 0) return finish(h, Error(2, k + " must be 0 (off) or a multiple of 128"));
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        drop_graphs(h);  // baked into the captured step graphs
        (self ? h->attn_part_keys_self : h->attn_part_keys_cross) = value;
    }
    else if (k == "train_dropout_sites") return finish(h, train_set_dropout_sites(h, value));
    else if (k == "attn_variant" || k == "attn_ring_stages" || k == "attn_ring_ctas" || k == "attn_ring_quartets") {
        // the kernel choice is baked into the captured step graphs
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        drop_graphs(h);
        if (k == "attn_variant") attn_decode_configure(value ? 1 : 0, 0, -1, 0);
        else if (k == "attn_ring_quartets") {
            if (value != 1 && value != 2) return finish(h, Error(2, "attn_ring_quartets must be 1 or 2"));
            attn_decode_configure(-1, 0, -1, value);
        }
        else if (k == "attn_ring_stages") {
            if (value != 2 && value != 3 && value != 4 && value != 6 && value != 8)
                return finish(h, Error(2, "attn_ring_stages must be 2, 3, 4, 6 or 8"));
            attn_decode_configure(-1, value, -1, 0);
        } else {
            if (value < 0 || value > 8) return finish(h, Error(2, "attn_ring_ctas must be in 0..8 (0 = automatic)"));
            attn_decode_configure(-1, 0, value, 0);
        }
    }
    else return finish(h, Error(3, "unknown option: " + k));
    return 0;
    END_GUARD(h)
}

int mrmt3_trace_enable(mrmt3_handle* h, int on) {
    GUARD(h)
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    Status st = trace_enable(h, on != 0);
    return finish(h, st);
    END_GUARD(h)
}

int mrmt3_trace_read(mrmt3_handle* h, uint64_t* out, int max_slots) {
    GUARD(h)
    if (!h->trace_buf.p || !out) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    int n = std::min(max_slots, 512);
    if (cudaMemcpy(out, h->trace_buf.p, (size_t)n * 16, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
    END_GUARD(h)
}

int mrmt3_test_gemm(mrmt3_handle* h, const void* a_bf16, const void* w_bf16, int M, int N, int K, float* c_f32,
                    int which, void* stream) {
    GUARD(h)
    return finish(h, test_gemm(h, (const bf16*)a_bf16, (const bf16*)w_bf16, M, N, K, c_f32, which, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_profile_enable(mrmt3_handle* h, int on) {
    GUARD(h)
    h->prof_on = on != 0;
    for (int i = 0; i < MRMT3_PROF_NCAT; ++i) {
        h->prof_ms[i] = 0;
        h->prof_n[i] = 0;
    }
    return 0;
    END_GUARD(h)
}

int mrmt3_profile_read(mrmt3_handle* h, double* ms_out, int64_t* launches_out, int n_cat) {
    GUARD(h)
    Status st = profile_collect(h);
    if (!st.ok()) return finish(h, st);
    for (int i = 0; i < n_cat && i < MRMT3_PROF_NCAT; ++i) {
        if (ms_out) ms_out[i] = h->prof_ms[i];
        if (launches_out) launches_out[i] = h->prof_n[i];
        h->prof_ms[i] = 0;
        h->prof_n[i] = 0;
    }
    return 0;
    END_GUARD(h)
}

int mrmt3_set_weight(mrmt3_handle* h, const char* name, const float* data, int rows, int cols) {
    GUARD(h)
    if (!name || !data) return finish(h, Error(1, "null argument"));
    return finish(h, set_weight(h, name, data, rows, cols));
    END_GUARD(h)
}

int mrmt3_commit_weights(mrmt3_handle* h) {
    GUARD(h)
    return finish(h, commit_weights(h));
    END_GUARD(h)
}

int mrmt3_set_mel_filterbank(mrmt3_handle* h, const float* fb_host) {
    GUARD(h)
    if (!fb_host) return finish(h, Error(1, "null argument"));
    cudaSetDevice(h->device);
    return finish(h, h->frontend.set_filterbank(fb_host));
    END_GUARD(h)
}

int mrmt3_logmel(mrmt3_handle* h, const float* audio, const int64_t* seg_start, const int32_t* seg_len,
                 const int32_t* valid_frames, int n_seg, int flags, float* out_f32, void* out_bf16,
                 void* stream) {
    GUARD(h)
    return finish(h, api_logmel(h, audio, (const long long*)seg_start, seg_len, valid_frames, n_seg, flags,
                                out_f32, (bf16*)out_bf16, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_encode(mrmt3_handle* h, const float* mel, int B, float* enc_out, void* stream) {
    GUARD(h)
    return finish(h, api_encode(h, mel, B, enc_out, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_generate(mrmt3_handle* h, const float* mel, int B, int max_length, int64_t* out_ids,
                   int32_t* steps_host, const int64_t* forced_ids, float* logits_out, void* stream) {
    GUARD(h)
    return finish(h, generate_base(h, mel, nullptr, B, max_length, (long long*)out_ids, steps_host,
                                   (const long long*)forced_ids, logits_out, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_generate_segmem(mrmt3_handle* h, const float* mel, const int32_t* seg_counts_host, int n_tracks,
                          int max_length, int64_t* out_ids, float* logits_out, void* stream) {
    GUARD(h)
    if (!seg_counts_host) return finish(h, Error(1, "null seg_counts_host"));
    return finish(h, generate_segmem(h, mel, nullptr, seg_counts_host, n_tracks, max_length,
                                     (long long*)out_ids, logits_out, (cudaStream_t)stream, nullptr));
    END_GUARD(h)
}

int mrmt3_generate_segmem_forced(mrmt3_handle* h, const float* mel, const int32_t* seg_counts_host, int n_tracks,
                                 int max_length, const int64_t* forced_ids, int64_t* out_ids, float* logits_out,
                                 void* stream) {
    GUARD(h)
    if (!seg_counts_host || !forced_ids) return finish(h, Error(1, "null argument"));
    return finish(h, generate_segmem(h, mel, nullptr, seg_counts_host, n_tracks, max_length, (long long*)out_ids,
                                     logits_out, (cudaStream_t)stream, (const long long*)forced_ids));
    END_GUARD(h)
}

int mrmt3_forward_logits(mrmt3_handle* h, const float* mel, int B, const int64_t* decoder_input_ids, int L,
                         const int64_t* targets_prev, int Lp, float* logits_out, void* stream) {
    GUARD(h)
    return finish(h, api_forward_logits(h, mel, B, (const long long*)decoder_input_ids, L,
                                        (const long long*)targets_prev, Lp, logits_out, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_memory_block(mrmt3_handle* h, const int64_t* prev_ids, int B, int Lp, float* mem_out, void* stream) {
    GUARD(h)
    return finish(h, api_memory_block(h, (const long long*)prev_ids, B, Lp, mem_out, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_init(mrmt3_handle* h, int64_t* n_params) {
    GUARD(h)
    Status st = train_init(h);
    if (st.ok() && n_params) *n_params = (int64_t)train_param_count(h);
    return finish(h, st);
    END_GUARD(h)
}

int mrmt3_train_set_dropout(mrmt3_handle* h, float p, uint64_t seed) {
    GUARD(h)
    return finish(h, train_set_dropout(h, p, (unsigned long long)seed));
    END_GUARD(h)
}

int mrmt3_dropout_keep_host(float p, uint64_t seed, uint32_t tensor_id, int64_t n, uint8_t* keep_out) {
    if (n < 0 || (n > 0 && !keep_out) || !(p >= 0.f && p < 1.f)) return 1;
    const DropSpec d = make_drop_spec(p, (unsigned long long)seed, tensor_id);
    for (int64_t i = 0; i < n; ++i) keep_out[i] = drop_factor(d, (unsigned long long)i) != 0.f;
    return 0;
}

int mrmt3_train_locate(mrmt3_handle* h, const char* name, int64_t* offset, int32_t* rows, int32_t* cols,
                       int32_t* row_mul, int32_t* row_off) {
    GUARD(h)
    if (!name || !offset || !rows || !cols || !row_mul || !row_off) return finish(h, Error(1, "null argument"));
    long long off = 0;
    int r = 0, c = 0, mul = 1, ro = 0;
    Status st = train_locate(h, name, &off, &r, &c, &mul, &ro);
    *offset = off; *rows = r; *cols = c; *row_mul = mul; *row_off = ro;
    return finish(h, st);
    END_GUARD(h)
}

int mrmt3_train_forward(mrmt3_handle* h, const float* mel, int B, const int64_t* decoder_input_ids,
                        const int64_t* labels, int L, const int64_t* targets_prev, int Lp, float* logits_out,
                        float* loss_host, void* stream) {
    GUARD(h)
    return finish(h, train_forward(h, mel, B, reinterpret_cast<const long long*>(decoder_input_ids),
                                   reinterpret_cast<const long long*>(labels), L,
                                   reinterpret_cast<const long long*>(targets_prev), Lp, logits_out, loss_host,
                                   (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_backward(mrmt3_handle* h, float* grad_flat, const float* dlogits, void* stream) {
    GUARD(h)
    return finish(h, train_backward(h, grad_flat, dlogits, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_apply(mrmt3_handle* h, const float* grad_flat, float lr, float beta1, float beta2, float eps,
                      float weight_decay, void* stream) {
    GUARD(h)
    return finish(h, train_apply(h, grad_flat, lr, beta1, beta2, eps, weight_decay, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_bucket_count(mrmt3_handle* h, int32_t* n) {
    GUARD(h)
    if (!n) return finish(h, Error(1, "null argument"));
    *n = train_bucket_count(h);
    return 0;
    END_GUARD(h)
}

int mrmt3_train_bucket(mrmt3_handle* h, int i, int64_t* offset, int64_t* count) {
    GUARD(h)
    if (!offset || !count) return finish(h, Error(1, "null argument"));
    long long o = 0, c = 0;
    Status st = train_bucket(h, i, &o, &c);
    *offset = o;
    *count = c;
    return finish(h, st);
    END_GUARD(h)
}

int mrmt3_train_wait_bucket(mrmt3_handle* h, int i, void* stream) {
    GUARD(h)
    return finish(h, train_wait_bucket(h, i, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_loss(mrmt3_handle* h, float* loss_host, void* stream) {
    GUARD(h)
    if (!loss_host) return finish(h, Error(1, "null argument"));
    return finish(h, train_loss(h, loss_host, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_train_read_master(mrmt3_handle* h, float* out_flat, void* stream) {
    GUARD(h)
    return finish(h, train_read_master(h, out_flat, (cudaStream_t)stream));
    END_GUARD(h)
}

int mrmt3_transcribe_host(mrmt3_handle* h, const float* audio_host, int64_t n_samples,
                          const int64_t* seg_start_host, const int32_t* seg_len_host,
                          const int32_t* valid_frames_host, int n_seg, const int32_t* seg_counts_host,
                          int n_tracks, int flags, int max_length, int64_t* out_ids_host, int32_t* steps_host,
                          void* stream) {
    GUARD(h)
    return finish(h, api_transcribe_host(h, audio_host, n_samples, (const long long*)seg_start_host, seg_len_host,
                                         valid_frames_host, n_seg, seg_counts_host, n_tracks, flags, max_length,
                                         (long long*)out_ids_host, steps_host, (cudaStream_t)stream));
    END_GUARD(h)
}

}  // extern "C"
#pragma GCC visibility pop
