// The fine-tune step (BASELINE configs[4]; reference tasks/mt3_net*.py `training_step` over
// models/t5.py:99-249): teacher-forced forward with saved activations, cross-entropy with
// ignore_index -100, hand-written backward through every op of the path, AdamW on fp32 masters.
// The gradient lives in ONE flat fp32 buffer owned by the caller (packed-weight order), so data
// parallel training is a single all-reduce of that buffer between `train_backward` and
// `train_apply`.  Dropout at the reference's sites is a counter-based hash the oracle
// mirrors (mrmt3_train_set_dropout; p = 0 is the reference's eval()-mode arithmetic).  Models: plain MT3 (models/t5.py) and MR-MT3 V2WithPrev
// (models/t5_segmem_v2_with_prev.py:60-153: memory block appended to the encoder output).
//
// GEMMs are the tcgen05 kernel with the operands as they lie in memory (no transposed copies):
//   dgrad  dX[M,Kin] = dY[M,N] . W[N,Kin]        A = dY K-major, B = W as stored, MN-major (gemm_tcgen05 MODE 2)
//   wgrad  dW[N,Kin] = dY^T[N,M] . X[M,Kin]      A = dY, B = X, both MN-major (MODE 1), split over the rows
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gemm_tcgen05.cuh"
#include "model.cuh"
#include "train.cuh"

namespace mrmt3 {

constexpr int kMaxTrainBatch = kMaxLanes;  // one cross-cache lane per sample

struct ParamSlot {
    bf16* w16;      // packed bf16 weight in the arena (nullptr for fp32-only parameters)
    float* w32;     // fp32 parameter in the arena (norm weights, embedding) or nullptr
    int rows, cols; // packed shape
    size_t off;     // offset in the flat gradient / master / moment buffers
};

struct LayerSlots {
    int wqkv, wo, ln_self, cq, co, ln_cross, wi, wff, ln_ff;
};

struct LayerStash {
    float *h_in, *h_mid2, *h_mid;
    bf16 *n1, *qkv, *ctx, *nc, *qc, *ctx_c, *n2, *raw, *ff;
    float *lse, *lse_c;
    unsigned short *keep, *keep_c;  // attention-dropout keep bits (AttnFullParams::keep)
};

struct TrainState {
    std::vector<ParamSlot> slots;
    std::vector<LayerSlots> enc, dec, mem;
    int proj = -1, emb = -1, lm_head = -1, cross_kv = -1, enc_final = -1, dec_final = -1, mem_final = -1, segmem_proj = -1;
    size_t n_total = 0;
    DeviceBuffer master, m, v, stash, scratch;
    DeviceBuffer adam_table;   // AdamSlot per tensor (train_apply)
    int adam_slots = 0;
    unsigned int adam_blocks = 0;
    int step = 0;
    // the last forward's saved state
    int B = 0, L = 0, Lp = 0, n_mem = 0;
    std::vector<LayerStash> enc_st, dec_st, mem_st;
    float *enc_h_final = nullptr, *dec_h_final = nullptr, *mem_h_final = nullptr;
    bf16 *mel16 = nullptr, *enc_n_final = nullptr, *dec_n_final = nullptr, *dlogits = nullptr;
    bf16 *mem_emb16 = nullptr, *mem_n_final = nullptr, *kv_in = nullptr;  // kv_in: (B, T_k, d) rows fed to the K/V projection
    const long long* prev_ids = nullptr;
    DeviceBuffer prev_copy;
    // dropout (reference config dropout_rate, 0 in the memory encoder): p, the seed of the next
    // forward and the seed the last forward used (the backward regenerates its masks from it)
    float drop_p = 0.f;
    unsigned long long drop_seed = 0, drop_seed_used = 0;
    int drop_sites = 0x1ff;  // bit i = site i enabled (tests isolate one site at a time)
    float* row_loss = nullptr;
    float* loss_pinned = nullptr;  // pinned landing spot of the loss scalar
    const long long* dec_ids = nullptr;
    DeviceBuffer ids_copy;
    // gradient buckets: contiguous ranges of the flat order, in the order the backward completes them
    // (the flat order is laid out for that), each with an event recorded when its last kernel is queued
    struct Bucket { size_t off, count; cudaEvent_t done; };
    std::vector<Bucket> buckets;
    std::vector<int> bucket_first_slot;  // slot index where each bucket starts
    float* scal = nullptr;               // [1 / #labels, mean loss] of the last forward (device)
    bool reloaded = false;  // weights were set after train_init: the optimizer state restarts at the next commit
};

static TrainState* state(mrmt3_handle* h) { return reinterpret_cast<TrainState*>(h->train); }

// dropout sites; stack 0 encoder, 1 decoder, 2 memory encoder (never dropped: models/t5_segmem.py:64)
enum { kSiteInput = 1, kSiteSelfProbs = 2, kSiteSelfOut = 3, kSiteFfnInner = 4, kSiteFfnOut = 5, kSiteFinal = 6,
       kSiteCrossProbs = 7, kSiteCrossOut = 8 };
static DropSpec drop_spec(const TrainState* t, unsigned long long seed, int stack, int layer, int site) {
    if (t->drop_p <= 0.f || stack == 2 || !((t->drop_sites >> site) & 1)) return DropSpec{0ull, 0u, 1.f};
    return make_drop_spec(t->drop_p, seed, (unsigned)((stack << 16) | (layer << 8) | site));
}

Status train_set_dropout_sites(mrmt3_handle* h, int mask) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    t->drop_sites = mask;
    return OkStatus();
}

Status train_set_dropout(mrmt3_handle* h, float p, unsigned long long seed) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    if (!(p >= 0.f && p < 1.f)) return Error(2, "dropout probability must be in [0, 1)");
    t->drop_p = p;
    t->drop_seed = seed;
    return OkStatus();
}

void train_destroy(mrmt3_handle* h) {
    TrainState* t = state(h);
    if (!t) return;
    t->master.release(); t->m.release(); t->v.release();
    t->stash.release(); t->scratch.release(); t->ids_copy.release(); t->prev_copy.release(); t->adam_table.release();
    for (auto& b : t->buckets)
        if (b.done) cudaEventDestroy(b.done);
    if (t->loss_pinned) cudaFreeHost(t->loss_pinned);
    delete t;
    h->train = nullptr;
}

static int add_slot(TrainState* t, bf16* w16, float* w32, int rows, int cols) {
    ParamSlot s{w16, w32, rows, cols, t->n_total};
    t->n_total += (size_t)rows * cols;
    t->slots.push_back(s);
    return (int)t->slots.size() - 1;
}

// The flat order is the order in which the backward pass FINISHES the gradients, so that every bucket
// handed to the all-reduce is one contiguous range: lm_head | decoder layers last to first | stacked
// cross K/V | memory encoder layers + segmem_proj | encoder layers last to first | proj, embedding and
// every norm weight (the embedding gradient gets its second contribution from the memory block and the
// norm-weight gradients are reduced at the very end of the backward).
static void begin_bucket(TrainState* t) { t->bucket_first_slot.push_back((int)t->slots.size()); }

// matrices of one layer, one bucket per layer, layers last to first
static void add_stack_matrices(TrainState* t, StackW& st, bool decoder, std::vector<LayerSlots>& out, bool bucket_per_layer) {
    out.assign(st.layers.size(), LayerSlots{});
    for (int li = (int)st.layers.size() - 1; li >= 0; --li) {
        LayerW& L = st.layers[li];
        LayerSlots& ls = out[li];
        if (bucket_per_layer) begin_bucket(t);
        ls.wff = add_slot(t, L.wff, nullptr, kDModel, kDFF);
        ls.wi = add_slot(t, L.wi, nullptr, 2 * kDFF, kDModel);
        ls.cq = ls.co = ls.ln_cross = -1;
        if (decoder) {
            ls.co = add_slot(t, L.co, nullptr, kDModel, kInner);
            ls.cq = add_slot(t, L.cq, nullptr, kInner, kDModel);
        }
        ls.wo = add_slot(t, L.wo, nullptr, kDModel, kInner);
        ls.wqkv = add_slot(t, L.wqkv, nullptr, 3 * kInner, kDModel);
    }
}

static void add_stack_norms(TrainState* t, StackW& st, bool decoder, std::vector<LayerSlots>& out, int& final_slot) {
    for (size_t li = 0; li < st.layers.size(); ++li) {
        LayerW& L = st.layers[li];
        out[li].ln_self = add_slot(t, nullptr, L.ln_self, 1, kDModel);
        if (decoder) out[li].ln_cross = add_slot(t, nullptr, L.ln_cross, 1, kDModel);
        out[li].ln_ff = add_slot(t, nullptr, L.ln_ff, 1, kDModel);
    }
    final_slot = add_slot(t, nullptr, st.final_ln, 1, kDModel);
}

Status train_init(mrmt3_handle* h) {
    if (h->train) return OkStatus();
    if (!h->committed) return Error(5, "weights not committed");
    if (h->cfg.mem_variant == MRMT3_MEM_V1_PREPEND)
        return Error(2, "the fine-tune step is implemented for the MT3 and V2WithPrev models");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    TrainState* t = new TrainState();
    h->train = t;
    const bool with_mem = h->cfg.mem_variant == MRMT3_MEM_V2_APPEND;
    begin_bucket(t);
    t->lm_head = add_slot(t, h->lm_head, nullptr, kVocab, kDModel);
    add_stack_matrices(t, h->dec, true, t->dec, true);
    begin_bucket(t);
    t->cross_kv = add_slot(t, h->cross_kv_w, nullptr, h->cfg.n_dec_layers * 2 * kInner, kDModel);
    if (with_mem) {
        begin_bucket(t);
        add_stack_matrices(t, h->mem, false, t->mem, false);
        t->segmem_proj = add_slot(t, h->segmem_proj, nullptr, kDModel, kDModel);
    }
    add_stack_matrices(t, h->enc, false, t->enc, true);
    begin_bucket(t);
    t->proj = add_slot(t, h->proj, nullptr, kDModel, kDModel);
    t->emb = add_slot(t, nullptr, h->emb, kVocab, kDModel);
    add_stack_norms(t, h->dec, true, t->dec, t->dec_final);
    if (with_mem) add_stack_norms(t, h->mem, false, t->mem, t->mem_final);
    add_stack_norms(t, h->enc, false, t->enc, t->enc_final);
    for (size_t i = 0; i < t->bucket_first_slot.size(); ++i) {
        const int s0 = t->bucket_first_slot[i];
        const int s1 = i + 1 < t->bucket_first_slot.size() ? t->bucket_first_slot[i + 1] : (int)t->slots.size();
        const size_t off = t->slots[s0].off;
        const size_t end = t->slots[s1 - 1].off + (size_t)t->slots[s1 - 1].rows * t->slots[s1 - 1].cols;
        TrainState::Bucket bk{off, end - off, nullptr};
        MRMT3_CUDA_TRY(cudaEventCreateWithFlags(&bk.done, cudaEventDisableTiming));
        t->buckets.push_back(bk);
    }
    const size_t n = t->n_total;
    MRMT3_TRY(t->master.reserve(n * 4));
    MRMT3_TRY(t->m.reserve(n * 4));
    MRMT3_TRY(t->v.reserve(n * 4));
    MRMT3_CUDA_TRY(cudaMemset(t->m.p, 0, n * 4));
    MRMT3_CUDA_TRY(cudaMemset(t->v.p, 0, n * 4));
    // fp32 masters: the lm_head and the norm-folded decoder weights kept theirs from set_weight;
    // the others restart from the bf16 copies (2^-9 relative, once)
    for (auto& s : t->slots) {
        float* dst = t->master.as<float>() + s.off;
        const size_t cnt = (size_t)s.rows * s.cols;
        if (s.w32) {
            MRMT3_CUDA_TRY(cudaMemcpy(dst, s.w32, cnt * 4, cudaMemcpyDeviceToDevice));
        } else {
            RUN(h, launch_bf16_to_f32(s.w16, dst, cnt, 0));
        }
    }
    auto copy_master = [&](int slot, const float* src) -> Status {
        const ParamSlot& s = t->slots[slot];
        MRMT3_CUDA_TRY(cudaMemcpy(t->master.as<float>() + s.off, src, (size_t)s.rows * s.cols * 4, cudaMemcpyDeviceToDevice));
        return OkStatus();
    };
    MRMT3_TRY(copy_master(t->lm_head, h->m_lm_head));
    for (size_t i = 0; i < h->dec.layers.size(); ++i) {
        MRMT3_TRY(copy_master(t->dec[i].wqkv, h->dec.layers[i].m_wqkv));
        MRMT3_TRY(copy_master(t->dec[i].cq, h->dec.layers[i].m_cq));
        MRMT3_TRY(copy_master(t->dec[i].wi, h->dec.layers[i].m_wi));
    }
    MRMT3_CUDA_TRY(cudaDeviceSynchronize());
    return OkStatus();
}

size_t train_param_count(mrmt3_handle* h) { return state(h) ? state(h)->n_total : 0; }

Status train_locate(mrmt3_handle* h, const std::string& name, long long* offset, int* rows, int* cols,
                    int* row_mul, int* row_off);

// mrmt3_set_weight after mrmt3_train_init (a checkpoint loaded into the module, a torch optimizer
// step on the mirror parameters): the tensor's fp32 master takes the exact fp32 value that was just
// staged, so that the next AdamW step starts from the loaded weights and not from stale masters.
// Tensors whose parameter IS the fp32 arena copy (embedding, norms) were already overwritten in place.
Status train_on_set_weight(mrmt3_handle* h, const std::string& name, const float* staged_f32) {
    TrainState* t = state(h);
    if (!t) return OkStatus();
    long long off = 0;
    int rows = 0, cols = 0, mul = 1, ro = 0;
    Status st = train_locate(h, name, &off, &rows, &cols, &mul, &ro);
    if (!st.ok()) return OkStatus();  // not a trainable tensor (aliases, inv_freq)
    MRMT3_TRY(launch_pack_weight_f32(staged_f32, t->master.as<float>() + off, rows, cols, mul, ro, 0));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(0));
    t->reloaded = true;
    return OkStatus();
}

// mrmt3_commit_weights after such a reload: Adam moments and the step count restart (the loaded
// weights are a new starting point; torch keeps optimizer state outside the model the same way)
void train_on_commit(mrmt3_handle* h) {
    TrainState* t = state(h);
    if (!t || !t->reloaded) return;
    cudaMemset(t->m.p, 0, t->n_total * 4);
    cudaMemset(t->v.p, 0, t->n_total * 4);
    t->step = 0;
    t->reloaded = false;
}

// flat fp32 copy of every trainable tensor (packed order), e.g. to rebuild a state dict after training
Status train_read_master(mrmt3_handle* h, float* out, cudaStream_t s) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    for (auto& sl : t->slots) {
        const float* src = sl.w32 ? sl.w32 : t->master.as<float>() + sl.off;
        MRMT3_CUDA_TRY(cudaMemcpyAsync(out + sl.off, src, (size_t)sl.rows * sl.cols * 4, cudaMemcpyDeviceToDevice, s));
    }
    return OkStatus();
}

// where a reference state-dict tensor lives inside the flat buffers: element (r, c) of the tensor
// is flat[offset + ((r * row_mul + row_off) * cols + c)]
Status train_locate(mrmt3_handle* h, const std::string& name, long long* offset, int* rows, int* cols,
                    int* row_mul, int* row_off) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    auto set = [&](int slot, int r, int c, int mul, int off) {
        *offset = (long long)t->slots[slot].off;
        *rows = r; *cols = c; *row_mul = mul; *row_off = off;
        return OkStatus();
    };
    const int d = kDModel, in = kInner;
    if (name == "proj.weight") return set(t->proj, d, d, 1, 0);
    if (name == "decoder_embed_tokens.weight") return set(t->emb, kVocab, d, 1, 0);
    if (name == "lm_head.weight") return set(t->lm_head, kVocab, d, 1, 0);
    if (name == "segmem_proj.weight" && t->segmem_proj >= 0) return set(t->segmem_proj, d, d, 1, 0);
    struct { const char* n; std::vector<LayerSlots>* ls; int* fin; bool dec; } stacks[] = {
        {"encoder.", &t->enc, &t->enc_final, false}, {"decoder.", &t->dec, &t->dec_final, true},
        {"segmem_encoder.", &t->mem, &t->mem_final, false}};
    for (auto& sk : stacks) {
        size_t pl = strlen(sk.n);
        if (name.compare(0, pl, sk.n) != 0) continue;
        std::string rest = name.substr(pl);
        if (rest == "final_layer_norm.weight") return set(*sk.fin, 1, d, 1, 0);
        int bi = -1, li = -1, consumed = 0;
        if (sscanf(rest.c_str(), "block.%d.layer.%d.%n", &bi, &li, &consumed) < 2 || consumed == 0) break;
        if (bi < 0 || bi >= (int)sk.ls->size()) break;
        const LayerSlots& L = (*sk.ls)[bi];
        std::string sub = rest.substr(consumed);
        const int ffi = sk.dec ? 2 : 1;
        if (li == 0 && sub == "SelfAttention.q.weight") return set(L.wqkv, in, d, 1, 0);
        if (li == 0 && sub == "SelfAttention.k.weight") return set(L.wqkv, in, d, 1, in);
        if (li == 0 && sub == "SelfAttention.v.weight") return set(L.wqkv, in, d, 1, 2 * in);
        if (li == 0 && sub == "SelfAttention.o.weight") return set(L.wo, d, in, 1, 0);
        if (li == 0 && sub == "layer_norm.weight") return set(L.ln_self, 1, d, 1, 0);
        if (sk.dec && li == 1 && sub == "EncDecAttention.q.weight") return set(L.cq, in, d, 1, 0);
        if (sk.dec && li == 1 && sub == "EncDecAttention.k.weight") return set(t->cross_kv, in, d, 1, bi * 2 * in);
        if (sk.dec && li == 1 && sub == "EncDecAttention.v.weight") return set(t->cross_kv, in, d, 1, bi * 2 * in + in);
        if (sk.dec && li == 1 && sub == "EncDecAttention.o.weight") return set(L.co, d, in, 1, 0);
        if (sk.dec && li == 1 && sub == "layer_norm.weight") return set(L.ln_cross, 1, d, 1, 0);
        if (li == ffi && sub == "DenseReluDense.wi_0.weight") return set(L.wi, kDFF, d, 2, 0);
        if (li == ffi && sub == "DenseReluDense.wi_1.weight") return set(L.wi, kDFF, d, 2, 1);
        if (li == ffi && sub == "DenseReluDense.wo.weight") return set(L.wff, d, kDFF, 1, 0);
        if (li == ffi && sub == "layer_norm.weight") return set(L.ln_ff, 1, d, 1, 0);
        break;
    }
    return Error(3, "no trainable tensor named " + name);
}

// dW (N, Kin) fp32 = dY^T . X with dY (M, N) pitch ldy and X (M, Kin) pitch ldx, both row-major and
// used as they lie (MN-major tcgen05 operands).  The output has only (N / 128) x (Kin / BN) tiles:
// the long reduction over M is split over the SMs (partials in `wpart`, 16 x the largest weight).
Status wgrad_gemm(mrmt3_handle* h, const bf16* dY, int ldy, int N, const bf16* X, int ldx, int Kin, float* dW, size_t M,
                  float* wpart, cudaStream_t s, int* splits_out) {
    const int bn = Kin % 256 == 0 ? 256 : (Kin % 192 == 0 ? 192 : (Kin % 128 == 0 ? 128 : 64));
    const int out_tiles = ceil_div(N, 128) * (Kin / bn);
    const int blocks = (int)((M + 63) / 64);
    int splits = 1;
    while (splits < 16 && out_tiles * splits * 2 <= 160 && splits * 2 <= blocks) splits *= 2;
    if (splits_out) *splits_out = splits;
    if (splits == 1) {
        RUN(h, launch_gemm_tc_mn(*h->tma, dY, ldy, N, X, ldx, Kin, (int)M, EpiStoreF32{dW, Kin}, s));
    } else {
        RUN(h, launch_gemm_tc_mn(*h->tma, dY, ldy, N, X, ldx, Kin, (int)M, EpiStoreF32{wpart, Kin}, s, splits));
        RUN(h, launch_reduce_splits(wpart, dW, (size_t)N * Kin, splits, s));
    }
    return OkStatus();
}

// test hook (mrmt3_test_gemm which = 4 / 5): the backward's two GEMM forms on caller data
Status test_gemm_train(mrmt3_handle* h, const bf16* A, const bf16* W, int M, int N, int K, float* C, int which,
                       cudaStream_t s) {
    if (which == 4) {  // C (M, N) = A^T W: A (K, M), W (K, N) row-major, reduction over the K rows (split-K as wgrad)
        DeviceBuffer part;
        MRMT3_TRY(part.reserve((size_t)16 * M * N * 4));
        Status st = wgrad_gemm(h, A, M, M, W, N, N, C, (size_t)K, part.as<float>(), s, nullptr);
        cudaStreamSynchronize(s);
        part.release();
        return st;
    }
    // which == 5: C (M, N) = A W: A (M, K) K-major, W (K, N) row-major taken MN-major (the dgrad form)
    RUN(h, launch_gemm_tc_nn(*h->tma, A, K, M, W, N, N, K, EpiStoreF32{C, N}, s));
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
struct Bump {
    char* p;
    size_t used = 0, cap;
    bool overflow = false;
    // a request past the reservation sets `overflow` and returns the base (valid memory): callers
    // check the flag BEFORE launching anything that would write through the pointer
    template <class T>
    T* take(size_t n) {
        used = (used + 255) & ~size_t(255);
        T* r = reinterpret_cast<T*>(p + used);
        used += n * sizeof(T);
        if (used > cap) {
            overflow = true;
            return reinterpret_cast<T*>(p);
        }
        return r;
    }
};
#define BUMP_CHECK(bp)                                                                                    \
    do {                                                                                                  \
        if ((bp).overflow) return Error(2, "internal: activation stash / scratch reservation too small"); \
    } while (0)

static size_t keep_bytes(size_t nb, size_t Tq, size_t Tk) { return nb * kHeads * Tq * ((Tk + 63) / 64) * 8; }

static size_t stash_bytes(int n_enc, int n_dec, int n_mem_layers, size_t Me, size_t Md, size_t Mm, int B, int L, int Lp,
                          int tk, bool drop) {
    auto layer = [&](size_t M, size_t T, bool dec) {
        size_t b = M * kDModel * 4 * (dec ? 3 : 2) + M * kDModel * 2 * (dec ? 3 : 2) + M * 3 * kInner * 2 +
                   M * kInner * 2 * (dec ? 3 : 1) + M * 2 * kDFF * 2 + M * kDFF * 2 + (size_t)B * kHeads * T * 4 * (dec ? 2 : 1);
        if (drop) b += keep_bytes(B, T, T) + (dec ? keep_bytes(B, T, tk) : 0);
        return b + 32 * 256;
    };
    return n_enc * layer(Me, kSegFrames, false) + n_dec * layer(Md, L, true) + n_mem_layers * layer(Mm, Lp, false) +
           Me * kDModel * (2 + 4 + 2) + Mm * kDModel * (2 + 4 + 2) + (Me + Mm) * kDModel * 2 +
           Md * kDModel * (4 + 2) + Md * (size_t)kVocab * 2 + Md * 4 + (1 << 16);
}

// forward over B segments with teacher forcing; logits (B, L, V) fp32 to the caller, loss to *loss_host
Status train_forward(mrmt3_handle* h, const float* mel, int B, const long long* dec_ids, const long long* labels,
                     int L, const long long* targets_prev, int Lp, float* logits_out, float* loss_host, cudaStream_t s) {
    MRMT3_TRY(train_init(h));
    TrainState* t = state(h);
    if (B <= 0 || L <= 0 || B > kMaxTrainBatch) return Error(2, "bad batch or length");
    if (L > h->n_pos || Lp > h->n_pos) return Error(2, "L / Lp exceed the positional table (reference FixedPositionalEmbedding max_length 5000)");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const int n_enc = h->cfg.n_enc_layers, n_dec = h->cfg.n_dec_layers;
    const bool with_mem = h->cfg.mem_variant == MRMT3_MEM_V2_APPEND;
    if (with_mem && (!targets_prev || Lp <= 0)) return Error(2, "targets_prev required for the memory variant");
    if (!with_mem) Lp = 0;
    const int n_mem = with_mem ? std::min(h->cfg.mem_len, Lp) : 0;
    const int tk = kSegFrames + n_mem;
    const size_t Me = (size_t)B * kSegFrames, Md = (size_t)B * L, Mm = (size_t)B * Lp;
    const float eps = h->cfg.ln_eps;
    const ARowMap id{nullptr, 1};
    MRMT3_TRY(t->stash.reserve(stash_bytes(n_enc, n_dec, with_mem ? h->cfg.n_mem_layers : 0, Me, Md, Mm, B, L, Lp, tk,
                                           t->drop_p > 0.f)));
    t->Lp = Lp;
    t->n_mem = n_mem;
    const unsigned long long seed = t->drop_seed;
    t->drop_seed_used = seed;
    t->drop_seed = seed * 6364136223846793005ull + 1442695040888963407ull;  // the next step draws new masks
    auto mk = [&](int stack, int layer, int site) { return drop_spec(t, seed, stack, layer, site); };
    Bump bp{reinterpret_cast<char*>(t->stash.p), 0, t->stash.cap};
    t->B = B;
    t->L = L;
    MRMT3_TRY(t->ids_copy.reserve(Md * sizeof(long long)));
    RUN(h, launch_sanitize_ids(dec_ids, t->ids_copy.as<long long>(), Md, h->cfg.pad_id, s));
    t->dec_ids = t->ids_copy.as<long long>();

    // ---- encoder (reference models/t5.py:253-258) ----
    t->mel16 = bp.take<bf16>(Me * kMels);
    float* H = bp.take<float>(Me * kDModel);  // residual stream: every sublayer writes its output into a fresh buffer
    BUMP_CHECK(bp);
    RUN(h, launch_cast_bf16(mel, t->mel16, Me * kMels, s));
    RUN(h, launch_gemm_tc(*h->tma, t->mel16, kDModel, Me, id, h->proj, kDModel, (int)Me, kDModel, kDModel,
                          EpiPosAdd{H, kDModel, h->pe, kSegFrames, 0}, s));
    auto attn_fwd = [&](const bf16* Q, long qb, int qr, const bf16* K, const bf16* V, long kb, long kh, int kr, bf16* O,
                        int Tq, int Tk, int causal, float* lse, int nb, DropSpec drop, unsigned short*& keep) -> Status {
        AttnFullParams ap{};
        keep = drop.on() ? bp.take<unsigned short>(keep_bytes(nb, Tq, Tk) / 2) : nullptr;
        BUMP_CHECK(bp);
        ap.keep = keep;
        ap.Q = Q; ap.q_batch_stride = qb; ap.q_head_stride = kDKV; ap.q_row_stride = qr;
        ap.K = K; ap.V = V;
        ap.k_batch_stride = ap.v_batch_stride = kb;
        ap.k_head_stride = ap.v_head_stride = kh;
        ap.k_row_stride = ap.v_row_stride = kr;
        ap.O = O; ap.o_batch_stride = (long)Tq * kInner; ap.o_head_stride = kDKV; ap.o_row_stride = kInner;
        ap.Tq = Tq; ap.Tk = Tk; ap.causal = causal; ap.causal_offset = 0; ap.lse2 = lse; ap.drop = drop;
        RUN(h, launch_attn_full_auto(*h->tma, ap, nb, s));
        return OkStatus();
    };
    // H' = H + dropout(A . W^T)   (sublayer output) into a fresh buffer: the old H is exactly the
    // activation the backward needs for this sublayer's norm, so it stays in the stash as it is
    auto out_proj = [&](const bf16* A, int K, const bf16* W, float*& Hres, size_t M, DropSpec drop) -> Status {
        float* Hnew = bp.take<float>(M * kDModel);
        BUMP_CHECK(bp);
        RUN(h, launch_gemm_tc(*h->tma, A, K, M, id, W, K, (int)M, kDModel, K, EpiResidualTo{Hres, Hnew, kDModel, drop}, s));
        Hres = Hnew;
        return OkStatus();
    };
    auto ffn_fwd = [&](const LayerW& Lw, LayerStash& st, float*& Hres, size_t M, int stack, int li) -> Status {
        st.h_mid = Hres;
        st.n2 = bp.take<bf16>(M * kDModel);
        st.raw = bp.take<bf16>(M * 2 * kDFF);
        st.ff = bp.take<bf16>(M * kDFF);
        BUMP_CHECK(bp);
        RUN(h, launch_rmsnorm(Hres, Lw.ln_ff, eps, st.n2, nullptr, (int)M, nullptr, 1, s));
        RUN(h, launch_gemm_tc(*h->tma, st.n2, kDModel, M, id, Lw.wi, kDModel, (int)M, 2 * kDFF, kDModel,
                              EpiStoreBf16{st.raw, 2 * kDFF}, s));
        RUN(h, launch_gated_gelu_fwd(st.raw, st.ff, M, mk(stack, li, kSiteFfnInner), s));
        MRMT3_TRY(out_proj(st.ff, kDFF, Lw.wff, Hres, M, mk(stack, li, kSiteFfnOut)));
        return OkStatus();
    };
    auto self_fwd = [&](const LayerW& Lw, LayerStash& st, float*& Hres, size_t M, int T, int causal, int stack, int li) -> Status {
        const int nb = (int)(M / T);
        st.h_in = Hres;
        st.n1 = bp.take<bf16>(M * kDModel);
        st.qkv = bp.take<bf16>(M * 3 * kInner);
        st.ctx = bp.take<bf16>(M * kInner);
        st.lse = bp.take<float>((size_t)nb * kHeads * T);
        BUMP_CHECK(bp);
        RUN(h, launch_rmsnorm(Hres, Lw.ln_self, eps, st.n1, nullptr, (int)M, nullptr, 1, s));
        RUN(h, launch_gemm_tc(*h->tma, st.n1, kDModel, M, id, Lw.wqkv, kDModel, (int)M, 3 * kInner, kDModel,
                              EpiStoreBf16{st.qkv, 3 * kInner}, s));
        MRMT3_TRY(attn_fwd(st.qkv, (long)T * 3 * kInner, 3 * kInner, st.qkv + kInner, st.qkv + 2 * kInner,
                           (long)T * 3 * kInner, kDKV, 3 * kInner, st.ctx, T, T, causal, st.lse, nb, mk(stack, li, kSiteSelfProbs),
                           st.keep));
        MRMT3_TRY(out_proj(st.ctx, kInner, Lw.wo, Hres, M, mk(stack, li, kSiteSelfOut)));
        return OkStatus();
    };
    RUN(h, launch_dropout_f32(H, Me * kDModel, mk(0, 0, kSiteInput), s));
    t->enc_st.assign(n_enc, LayerStash{});
    for (int li = 0; li < n_enc; ++li) {
        MRMT3_TRY(self_fwd(h->enc.layers[li], t->enc_st[li], H, Me, kSegFrames, 0, 0, li));
        MRMT3_TRY(ffn_fwd(h->enc.layers[li], t->enc_st[li], H, Me, 0, li));
    }
    t->enc_h_final = H;
    t->enc_n_final = bp.take<bf16>(Me * kDModel);
    BUMP_CHECK(bp);
    RUN(h, launch_rmsnorm(H, h->enc.final_ln, eps, t->enc_n_final, nullptr, (int)Me, nullptr, 1, s));
    RUN(h, launch_dropout_bf16(t->enc_n_final, Me * kDModel, mk(0, 0, kSiteFinal), s));

    // ---- MR-MT3 memory block (models/t5_segmem_v2_with_prev.py:119-128): Emb[prev] -> segmem_proj
    //      (+PE) -> memory encoder over all Lp positions -> final norm -> first n_mem rows ----
    if (with_mem) {
        MRMT3_TRY(t->prev_copy.reserve(Mm * sizeof(long long)));
        RUN(h, launch_sanitize_ids(targets_prev, t->prev_copy.as<long long>(), Mm, h->cfg.pad_id, s));
        t->prev_ids = t->prev_copy.as<long long>();
        t->mem_emb16 = bp.take<bf16>(Mm * kDModel);
        float* Hm = bp.take<float>(Mm * kDModel);
        BUMP_CHECK(bp);
        RUN(h, launch_embed_bf16(t->prev_ids, Lp, Lp, B, nullptr, h->emb, t->mem_emb16, s));
        RUN(h, launch_gemm_tc(*h->tma, t->mem_emb16, kDModel, Mm, id, h->segmem_proj, kDModel, (int)Mm, kDModel, kDModel,
                              EpiPosAdd{Hm, kDModel, h->pe, Lp, 0}, s));
        t->mem_st.assign(h->cfg.n_mem_layers, LayerStash{});
        for (int li = 0; li < h->cfg.n_mem_layers; ++li) {
            MRMT3_TRY(self_fwd(h->mem.layers[li], t->mem_st[li], Hm, Mm, Lp, 0, 2, li));
            MRMT3_TRY(ffn_fwd(h->mem.layers[li], t->mem_st[li], Hm, Mm, 2, li));
        }
        t->mem_h_final = Hm;
        t->mem_n_final = bp.take<bf16>(Mm * kDModel);
        BUMP_CHECK(bp);
        RUN(h, launch_rmsnorm(Hm, h->mem.final_ln, eps, t->mem_n_final, nullptr, (int)Mm, nullptr, 1, s));
    }

    // ---- cross K/V of all decoder layers: rows [encoder states ; memory rows] per sample ----
    MRMT3_TRY(ensure_decode_capacity(h, B, tk, 1));
    t->kv_in = bp.take<bf16>((size_t)B * tk * kDModel);
    BUMP_CHECK(bp);
    MRMT3_CUDA_TRY(cudaMemcpy2DAsync(t->kv_in, (size_t)tk * kDModel * 2, t->enc_n_final, (size_t)kSegFrames * kDModel * 2,
                                     (size_t)kSegFrames * kDModel * 2, B, cudaMemcpyDeviceToDevice, s));
    if (n_mem)
        MRMT3_CUDA_TRY(cudaMemcpy2DAsync(t->kv_in + (size_t)kSegFrames * kDModel, (size_t)tk * kDModel * 2, t->mem_n_final,
                                         (size_t)Lp * kDModel * 2, (size_t)n_mem * kDModel * 2, B, cudaMemcpyDeviceToDevice, s));
    RUN(h, launch_gemm_tc(*h->tma, t->kv_in, kDModel, (long)B * tk, id, h->cross_kv_w, kDModel, B * tk,
                          n_dec * 2 * kInner, kDModel,
                          EpiCrossKV{h->cross_cache.as<bf16>(), tk, 0, n_dec, h->tk_cap, nullptr}, s));

    // ---- decoder, teacher forced (reference models/t5.py:99-180) ----
    float* Hd = bp.take<float>(Md * kDModel);
    BUMP_CHECK(bp);
    RUN(h, launch_embed_tokens(t->dec_ids, h->emb, h->pe, Hd, B, L, 0, s));
    RUN(h, launch_dropout_f32(Hd, Md * kDModel, mk(1, 0, kSiteInput), s));
    t->dec_st.assign(n_dec, LayerStash{});
    const size_t lane_sz = (size_t)n_dec * 2 * kHeads * h->tk_cap * kDKV;
    for (int li = 0; li < n_dec; ++li) {
        const LayerW& Lw = h->dec.layers[li];
        LayerStash& st = t->dec_st[li];
        MRMT3_TRY(self_fwd(Lw, st, Hd, Md, L, 1, 1, li));
        st.h_mid2 = Hd;
        st.nc = bp.take<bf16>(Md * kDModel);
        st.qc = bp.take<bf16>(Md * kInner);
        st.ctx_c = bp.take<bf16>(Md * kInner);
        st.lse_c = bp.take<float>((size_t)B * kHeads * L);
        BUMP_CHECK(bp);
        RUN(h, launch_rmsnorm(Hd, Lw.ln_cross, eps, st.nc, nullptr, (int)Md, nullptr, 1, s));
        RUN(h, launch_gemm_tc(*h->tma, st.nc, kDModel, Md, id, Lw.cq, kDModel, (int)Md, kInner, kDModel,
                              EpiStoreBf16{st.qc, kInner}, s));
        const bf16* kbase = h->cross_cache.as<bf16>() + (size_t)li * 2 * kHeads * h->tk_cap * kDKV;
        MRMT3_TRY(attn_fwd(st.qc, (long)L * kInner, kInner, kbase, kbase + (size_t)kHeads * h->tk_cap * kDKV, (long)lane_sz,
                           (long)h->tk_cap * kDKV, kDKV, st.ctx_c, L, tk, 0, st.lse_c, B, mk(1, li, kSiteCrossProbs), st.keep_c));
        MRMT3_TRY(out_proj(st.ctx_c, kInner, Lw.co, Hd, Md, mk(1, li, kSiteCrossOut)));
        MRMT3_TRY(ffn_fwd(Lw, st, Hd, Md, 1, li));
    }
    t->dec_h_final = Hd;
    t->dec_n_final = bp.take<bf16>(Md * kDModel);
    t->dlogits = bp.take<bf16>(Md * kVocab);
    t->row_loss = bp.take<float>(Md);
    float* scal = bp.take<float>(2);
    t->scal = scal;
    BUMP_CHECK(bp);
    RUN(h, launch_rmsnorm(Hd, h->dec.final_ln, eps, t->dec_n_final, nullptr, (int)Md, nullptr, 1, s));
    RUN(h, launch_dropout_bf16(t->dec_n_final, Md * kDModel, mk(1, 0, kSiteFinal), s));
    RUN(h, launch_gemm_tc(*h->tma, t->dec_n_final, kDModel, Md, id, h->lm_head, kDModel, (int)Md, kVocab, kDModel,
                          EpiStoreF32{logits_out, kVocab}, s));

    // ---- loss (tasks/mt3_net.py: CrossEntropyLoss(ignore_index=-100) over (B*L, V)) ----
    RUN(h, launch_xent(logits_out, labels, (int)Md, kVocab, scal, t->row_loss, t->dlogits, s));
    h->launches += 2;
    // the only host round trip of the forward, and only when the caller wants the number now
    if (loss_host) {
        if (!t->loss_pinned) MRMT3_CUDA_TRY(cudaMallocHost(&t->loss_pinned, sizeof(float)));
        MRMT3_CUDA_TRY(cudaMemcpyAsync(t->loss_pinned, scal + 1, sizeof(float), cudaMemcpyDeviceToHost, s));
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
        *loss_host = *t->loss_pinned;
    }
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// backward of the last train_forward into the caller's flat fp32 gradient buffer (overwritten)
Status train_backward(mrmt3_handle* h, float* grad, const float* dlogits_f32, cudaStream_t s) {
    TrainState* t = state(h);
    if (!t || !t->B) return Error(5, "mrmt3_train_forward first");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    // an external loss (autograd on the logits) replaces the built-in cross-entropy gradient
    if (dlogits_f32) RUN(h, launch_cast_f32_bf16(dlogits_f32, t->dlogits, (size_t)t->B * t->L * kVocab, s));
    const int B = t->B, L = t->L, Lp = t->Lp, n_mem = t->n_mem, n_enc = h->cfg.n_enc_layers, n_dec = h->cfg.n_dec_layers;
    const int tk = kSegFrames + n_mem;
    const size_t Me = (size_t)B * kSegFrames, Md = (size_t)B * L, Mm = (size_t)B * Lp, Mk = (size_t)B * tk;
    const size_t Mmax = std::max(std::max(Me, Md), std::max(Mm, Mk));
    const float eps = h->cfg.ln_eps;
    MRMT3_CUDA_TRY(cudaMemsetAsync(grad, 0, t->n_total * 4, s));
    const unsigned long long seed = t->drop_seed_used;
    auto mk = [&](int stack, int layer, int site) { return drop_spec(t, seed, stack, layer, site); };

    // scratch
    const size_t kvN = (size_t)n_dec * 2 * kInner;
    size_t need = Mmax * kDModel * 4 * 2 + Mmax * kDModel * 2 * 2 + Mmax * 2 * kDFF * 2 + Mmax * kDFF * 2 +
                  Mmax * 3 * kInner * 2 + Mmax * kInner * 2 * 2 +
                  (size_t)16 * 2 * kDFF * kDModel * 4 + Mk * kvN * 2 + (size_t)B * kHeads * std::max(std::max(L, Lp), kSegFrames) * 4 + Mmax * kDModel * 2 +
                  embed_bwd_scratch_bytes((int)Mmax) + (size_t)kNormBwdMaxNorms * kNormBwdMaxParts * kDModel * 4 + (1 << 16);
    MRMT3_TRY(t->scratch.reserve(need));
    Bump bp{reinterpret_cast<char*>(t->scratch.p), 0, t->scratch.cap};
    float* dH = bp.take<float>(Mmax * kDModel);
    bf16* dHb = bp.take<bf16>(Mmax * kDModel);
    bf16* dn = bp.take<bf16>(Mmax * kDModel);
    bf16* draw = bp.take<bf16>(Mmax * 2 * kDFF);
    bf16* dff = bp.take<bf16>(Mmax * kDFF);
    bf16* dqkv = bp.take<bf16>(Mmax * 3 * kInner);
    bf16* dctx = bp.take<bf16>(Mmax * kInner);
    bf16* dqc = bp.take<bf16>(Mmax * kInner);
    float* wpart = bp.take<float>((size_t)16 * 2 * kDFF * kDModel);  // split-K partials (a split needs <= 160 / splits tiles)
    bf16* dkv = bp.take<bf16>(Mk * kvN);
    bf16* dsplit = bp.take<bf16>(Mmax * kDModel);  // rows of the K/V-input gradient regrouped per consumer
    float* delta = bp.take<float>((size_t)B * kHeads * std::max(std::max(L, Lp), kSegFrames));
    void* emb_scratch = bp.take<char>(embed_bwd_scratch_bytes((int)Mmax));
    float* norm_parts = bp.take<float>((size_t)kNormBwdMaxNorms * kNormBwdMaxParts * kDModel);
    NormDgList norm_list{};
    BUMP_CHECK(bp);

    // optional per-category timing (MRMT3_TRAIN_PROFILE=1): CUDA events around every backward op
    static const bool prof = getenv("MRMT3_TRAIN_PROFILE") && atoi(getenv("MRMT3_TRAIN_PROFILE")) != 0;
    struct Rec { const char* cat; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    auto tic = [&](const char* cat) {
        if (!prof) return;
        Rec r{cat, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    };
    auto toc = [&]() {
        if (prof) cudaEventRecord(recs.back().b, s);
    };
    auto G = [&](int slot) { return grad + t->slots[slot].off; };
    // dX (M, Kin) = dY (M, N) . W (N, Kin): the weight as stored is the MN-major B operand of the
    // tcgen05 GEMM (no transposed weight copies)
    auto dgrad_to = [&](const bf16* dY, int N, int slot, auto epi, size_t M) -> Status {
        const ParamSlot& sl = t->slots[slot];
        tic("dgrad");
        RUN(h, launch_gemm_tc_nn(*h->tma, dY, N, (int)M, sl.w16, sl.cols, sl.cols, N, epi, s));
        toc();
        return OkStatus();
    };
    auto dgrad = [&](const bf16* dY, int N, int slot, bf16* dX, size_t M) -> Status {
        return dgrad_to(dY, N, slot, EpiStoreBf16{dX, t->slots[slot].cols}, M);
    };
    // dW (N, Kin) fp32 = dY^T . X ; dY (M, N) with pitch ldy, X (M, Kin) with pitch ldx
    auto wgrad = [&](const bf16* dY, int ldy, int N, const bf16* X, int ldx, int Kin, float* dW, size_t M) -> Status {
        tic("wgrad gemm");
        MRMT3_TRY(wgrad_gemm(h, dY, ldy, N, X, ldx, Kin, dW, M, wpart, s, nullptr));
        toc();
        return OkStatus();
    };
    // dHb mirrors dH (every update of dH goes through norm_bwd, which writes both); a sublayer
    // whose output was dropped back-propagates dH * mask / (1 - p) instead
    auto cast_dH = [&](size_t M, DropSpec drop) -> Status {
        if (drop.on()) RUN(h, launch_dropout_cast(dH, dHb, M * kDModel, drop, s));
        return OkStatus();
    };
    auto norm_bwd = [&](const float* x, const float* g, const bf16* dy, size_t M, float* dg) -> Status {
        tic("rmsnorm bwd");
        if (norm_list.n >= kNormBwdMaxNorms) return Error(2, "internal: too many norms in one backward pass");
        float* part = norm_parts + (size_t)norm_list.n * kNormBwdMaxParts * kDModel;
        norm_list.dst[norm_list.n] = dg;
        norm_list.n_parts[norm_list.n] = rmsnorm_bwd_parts((int)M);
        norm_list.n += 1;
        RUN(h, launch_rmsnorm_bwd(x, g, eps, dy, (int)M, dH, dHb, part, s));
        toc();
        return OkStatus();
    };
    auto attn_bwd = [&](const bf16* Q, long qb, int qr, const bf16* K, const bf16* V, long kb, long kh, int kr,
                        const bf16* O, const bf16* dO, bf16* dQ, bf16* dK, bf16* dV, long dkb, long dkh, int dkr,
                        const float* lse, int Tq, int Tk, int causal, DropSpec drop, const unsigned short* keep) -> Status {
        AttnBwdParams ap{};
        ap.drop = drop;
        ap.keep = keep;
        if (drop.on() && !keep) return Error(3, "attention dropout keep bits missing (forward ran without dropout?)");
        ap.Q = Q; ap.K = K; ap.V = V; ap.O = O; ap.dO = dO; ap.dQ = dQ; ap.dK = dK; ap.dV = dV;
        ap.q_batch_stride = qb; ap.q_head_stride = kDKV; ap.q_row_stride = qr;
        ap.k_batch_stride = ap.v_batch_stride = kb;
        ap.k_head_stride = ap.v_head_stride = kh;
        ap.k_row_stride = ap.v_row_stride = kr;
        ap.o_batch_stride = (long)Tq * kInner; ap.o_head_stride = kDKV; ap.o_row_stride = kInner;
        ap.dk_batch_stride = dkb; ap.dk_head_stride = dkh; ap.dk_row_stride = dkr;
        ap.lse2 = lse; ap.delta = delta; ap.Tq = Tq; ap.Tk = Tk; ap.causal = causal; ap.causal_offset = 0;
        tic(causal ? "attention bwd (causal self)" : (Tk == Tq ? "attention bwd (self)" : "attention bwd (cross)"));
        RUN(h, launch_attn_bwd(ap, B, s));
        toc();
        h->launches += 1;
        return OkStatus();
    };
    auto ffn_bwd = [&](const LayerSlots& ls, const LayerW& Lw, const LayerStash& st, size_t M, int stack, int li) -> Status {
        MRMT3_TRY(cast_dH(M, mk(stack, li, kSiteFfnOut)));
        MRMT3_TRY(dgrad(dHb, kDModel, ls.wff, dff, M));
        MRMT3_TRY(wgrad(dHb, kDModel, kDModel, st.ff, kDFF, kDFF, G(ls.wff), M));
        tic("gated gelu bwd");
        RUN(h, launch_gated_gelu_bwd(st.raw, dff, draw, M, mk(stack, li, kSiteFfnInner), s));
        toc();
        MRMT3_TRY(dgrad(draw, 2 * kDFF, ls.wi, dn, M));
        MRMT3_TRY(wgrad(draw, 2 * kDFF, 2 * kDFF, st.n2, kDModel, kDModel, G(ls.wi), M));
        MRMT3_TRY(norm_bwd(st.h_mid, Lw.ln_ff, dn, M, G(ls.ln_ff)));
        return OkStatus();
    };
    auto self_bwd = [&](const LayerSlots& ls, const LayerW& Lw, const LayerStash& st, size_t M, int T, int causal,
                        int stack, int li) -> Status {
        MRMT3_TRY(cast_dH(M, mk(stack, li, kSiteSelfOut)));
        MRMT3_TRY(dgrad(dHb, kDModel, ls.wo, dctx, M));
        MRMT3_TRY(wgrad(dHb, kDModel, kDModel, st.ctx, kInner, kInner, G(ls.wo), M));
        const long tb = (long)T * 3 * kInner;
        MRMT3_TRY(attn_bwd(st.qkv, tb, 3 * kInner, st.qkv + kInner, st.qkv + 2 * kInner, tb, kDKV, 3 * kInner, st.ctx, dctx,
                           dqkv, dqkv + kInner, dqkv + 2 * kInner, tb, kDKV, 3 * kInner, st.lse, T, T, causal,
                           mk(stack, li, kSiteSelfProbs), st.keep));
        MRMT3_TRY(dgrad(dqkv, 3 * kInner, ls.wqkv, dn, M));
        MRMT3_TRY(wgrad(dqkv, 3 * kInner, 3 * kInner, st.n1, kDModel, kDModel, G(ls.wqkv), M));
        MRMT3_TRY(norm_bwd(st.h_in, Lw.ln_self, dn, M, G(ls.ln_self)));
        return OkStatus();
    };

    // bucket i of the flat gradient is final once everything queued so far has run
    const int bk_dec0 = 1, bk_cross = 1 + n_dec, bk_mem = t->mem.empty() ? -1 : 2 + n_dec;
    const int bk_enc0 = (t->mem.empty() ? 2 : 3) + n_dec, bk_last = bk_enc0 + n_enc;
    if ((int)t->buckets.size() != bk_last + 1) return Error(2, "internal: gradient bucket table out of step with the backward");
    auto bucket_done = [&](int i) -> Status {
        MRMT3_CUDA_TRY(cudaEventRecord(t->buckets[i].done, s));
        return OkStatus();
    };

    // ---- head ----
    MRMT3_CUDA_TRY(cudaMemsetAsync(dH, 0, Md * kDModel * 4, s));
    MRMT3_TRY(dgrad(t->dlogits, kVocab, t->lm_head, dn, Md));
    MRMT3_TRY(wgrad(t->dlogits, kVocab, kVocab, t->dec_n_final, kDModel, kDModel, G(t->lm_head), Md));
    MRMT3_TRY(bucket_done(0));
    RUN(h, launch_dropout_bf16(dn, Md * kDModel, mk(1, 0, kSiteFinal), s));
    MRMT3_TRY(norm_bwd(t->dec_h_final, h->dec.final_ln, dn, Md, G(t->dec_final)));

    // ---- decoder layers, last to first ----
    MRMT3_CUDA_TRY(cudaMemsetAsync(dkv, 0, Mk * kvN * 2, s));
    const size_t lane_sz = (size_t)n_dec * 2 * kHeads * h->tk_cap * kDKV;
    for (int li = n_dec - 1; li >= 0; --li) {
        const LayerW& Lw = h->dec.layers[li];
        const LayerSlots& ls = t->dec[li];
        const LayerStash& st = t->dec_st[li];
        MRMT3_TRY(ffn_bwd(ls, Lw, st, Md, 1, li));
        // cross-attention sublayer
        MRMT3_TRY(cast_dH(Md, mk(1, li, kSiteCrossOut)));
        MRMT3_TRY(dgrad(dHb, kDModel, ls.co, dctx, Md));
        MRMT3_TRY(wgrad(dHb, kDModel, kDModel, st.ctx_c, kInner, kInner, G(ls.co), Md));
        const bf16* kbase = h->cross_cache.as<bf16>() + (size_t)li * 2 * kHeads * h->tk_cap * kDKV;
        // dK / dV of this layer go straight into the (B*256, n_dec*768) operand of the stacked K/V weight
        bf16* dk_l = dkv + (size_t)li * 2 * kInner;
        MRMT3_TRY(attn_bwd(st.qc, (long)L * kInner, kInner, kbase, kbase + (size_t)kHeads * h->tk_cap * kDKV, (long)lane_sz,
                           (long)h->tk_cap * kDKV, kDKV, st.ctx_c, dctx, dqc, dk_l, dk_l + kInner,
                           (long)tk * (long)kvN, kDKV, (int)kvN, st.lse_c, L, tk, 0, mk(1, li, kSiteCrossProbs), st.keep_c));
        MRMT3_TRY(dgrad(dqc, kInner, ls.cq, dn, Md));
        MRMT3_TRY(wgrad(dqc, kInner, kInner, st.nc, kDModel, kDModel, G(ls.cq), Md));
        MRMT3_TRY(norm_bwd(st.h_mid2, Lw.ln_cross, dn, Md, G(ls.ln_cross)));
        MRMT3_TRY(self_bwd(ls, Lw, st, Md, L, 1, 1, li));
        MRMT3_TRY(bucket_done(bk_dec0 + (n_dec - 1 - li)));
    }
    RUN(h, launch_dropout_f32(dH, Md * kDModel, mk(1, 0, kSiteInput), s));
    tic("embedding bwd");
    RUN(h, launch_embed_bwd(t->dec_ids, dH, G(t->emb), (int)Md, emb_scratch, s));
    h->launches += 1;
    toc();

    // ---- cross K/V projection -> [encoder output ; memory rows] ----
    MRMT3_TRY(dgrad(dkv, (int)kvN, t->cross_kv, dn, Mk));
    MRMT3_TRY(wgrad(dkv, (int)kvN, (int)kvN, t->kv_in, kDModel, kDModel, G(t->cross_kv), Mk));
    MRMT3_TRY(bucket_done(bk_cross));

    // ---- memory block backward (MR-MT3) ----
    if (n_mem) {
        // rows 256.. of every sample are the first n_mem rows of its memory sequence; the other
        // Lp - n_mem rows of the memory encoder output receive no gradient
        MRMT3_CUDA_TRY(cudaMemsetAsync(dsplit, 0, Mm * kDModel * 2, s));
        MRMT3_CUDA_TRY(cudaMemcpy2DAsync(dsplit, (size_t)Lp * kDModel * 2, dn + (size_t)kSegFrames * kDModel,
                                         (size_t)tk * kDModel * 2, (size_t)n_mem * kDModel * 2, B, cudaMemcpyDeviceToDevice, s));
        MRMT3_CUDA_TRY(cudaMemsetAsync(dH, 0, Mm * kDModel * 4, s));
        MRMT3_TRY(norm_bwd(t->mem_h_final, h->mem.final_ln, dsplit, Mm, G(t->mem_final)));
        // dsplit is free again; keep the encoder rows of dn for later in it
        MRMT3_CUDA_TRY(cudaMemcpy2DAsync(dsplit, (size_t)kSegFrames * kDModel * 2, dn, (size_t)tk * kDModel * 2,
                                         (size_t)kSegFrames * kDModel * 2, B, cudaMemcpyDeviceToDevice, s));
        for (int li = h->cfg.n_mem_layers - 1; li >= 0; --li) {
            MRMT3_TRY(ffn_bwd(t->mem[li], h->mem.layers[li], t->mem_st[li], Mm, 2, li));
            MRMT3_TRY(self_bwd(t->mem[li], h->mem.layers[li], t->mem_st[li], Mm, Lp, 0, 2, li));
        }
        // stack input = segmem_proj(Emb[prev]) + PE
        MRMT3_TRY(cast_dH(Mm, mk(2, 0, kSiteInput)));
        MRMT3_TRY(wgrad(dHb, kDModel, kDModel, t->mem_emb16, kDModel, kDModel, G(t->segmem_proj), Mm));
        MRMT3_TRY(dgrad_to(dHb, kDModel, t->segmem_proj, EpiStoreF32{dH, kDModel}, Mm));
        tic("embedding bwd");
        RUN(h, launch_embed_bwd(t->prev_ids, dH, G(t->emb), (int)Mm, emb_scratch, s));
        h->launches += 1;
        toc();
        MRMT3_TRY(bucket_done(bk_mem));
    } else {
        MRMT3_CUDA_TRY(cudaMemcpyAsync(dsplit, dn, Me * kDModel * 2, cudaMemcpyDeviceToDevice, s));
    }

    // ---- encoder ----
    RUN(h, launch_dropout_bf16(dsplit, Me * kDModel, mk(0, 0, kSiteFinal), s));
    MRMT3_CUDA_TRY(cudaMemsetAsync(dH, 0, Me * kDModel * 4, s));
    MRMT3_TRY(norm_bwd(t->enc_h_final, h->enc.final_ln, dsplit, Me, G(t->enc_final)));
    for (int li = n_enc - 1; li >= 0; --li) {
        MRMT3_TRY(ffn_bwd(t->enc[li], h->enc.layers[li], t->enc_st[li], Me, 0, li));
        MRMT3_TRY(self_bwd(t->enc[li], h->enc.layers[li], t->enc_st[li], Me, kSegFrames, 0, 0, li));
        MRMT3_TRY(bucket_done(bk_enc0 + (n_enc - 1 - li)));
    }
    // proj: h0 = dropout(mel . Wproj^T + PE)
    MRMT3_TRY(cast_dH(Me, mk(0, 0, kSiteInput)));
    MRMT3_TRY(wgrad(dHb, kDModel, kDModel, t->mel16, kMels, kMels, G(t->proj), Me));
    // norm-weight gradients: the partial rows of every norm of the step, added in a fixed order
    RUN(h, launch_norm_dg_reduce(norm_list, norm_parts, s));
    MRMT3_TRY(bucket_done(bk_last));
    if (prof) {
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
        std::map<std::string, std::pair<double, int>> agg;
        for (auto& r : recs) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            agg[r.cat].first += ms;
            agg[r.cat].second += 1;
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (auto& kv : agg) fprintf(stderr, "[train profile] %-28s %8.3f ms in %d calls\n", kv.first.c_str(), kv.second.first, kv.second.second);
    }
    return OkStatus();
}

// ---- gradient buckets (data-parallel overlap: all-reduce bucket i while the backward still runs) ----
int train_bucket_count(mrmt3_handle* h) { return state(h) ? (int)state(h)->buckets.size() : 0; }
Status train_bucket(mrmt3_handle* h, int i, long long* offset, long long* count) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    if (i < 0 || i >= (int)t->buckets.size()) return Error(2, "bucket index out of range");
    *offset = (long long)t->buckets[i].off;
    *count = (long long)t->buckets[i].count;
    return OkStatus();
}
Status train_wait_bucket(mrmt3_handle* h, int i, cudaStream_t stream) {
    TrainState* t = state(h);
    if (!t || !t->B) return Error(5, "mrmt3_train_backward first");
    if (i < 0 || i >= (int)t->buckets.size()) return Error(2, "bucket index out of range");
    MRMT3_CUDA_TRY(cudaStreamWaitEvent(stream, t->buckets[i].done, 0));
    return OkStatus();
}
// mean loss of the last train_forward (synchronises `s`)
Status train_loss(mrmt3_handle* h, float* loss_host, cudaStream_t s) {
    TrainState* t = state(h);
    if (!t || !t->scal) return Error(5, "mrmt3_train_forward first");
    if (!t->loss_pinned) MRMT3_CUDA_TRY(cudaMallocHost(&t->loss_pinned, sizeof(float)));
    MRMT3_CUDA_TRY(cudaMemcpyAsync(t->loss_pinned, t->scal + 1, sizeof(float), cudaMemcpyDeviceToHost, s));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    *loss_host = *t->loss_pinned;
    return OkStatus();
}

// AdamW step on the (possibly all-reduced) flat gradient; refreshes the bf16 weights and the
// norm-folded decoder copies
Status train_apply(mrmt3_handle* h, const float* grad, float lr, float beta1, float beta2, float adam_eps, float wd,
                   cudaStream_t s) {
    TrainState* t = state(h);
    if (!t) return Error(5, "mrmt3_train_init first");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    ++t->step;
    if (!t->adam_table.p) {  // one table entry per tensor, built once (slots and buffers are fixed after train_init)
        std::vector<AdamSlot> tab;
        unsigned int blocks = 0;
        for (auto& sl : t->slots) {
            const size_t n = (size_t)sl.rows * sl.cols;
            if (!n) continue;
            tab.push_back(AdamSlot{sl.w32 ? sl.w32 : t->master.as<float>() + sl.off, sl.w16, sl.off, n, blocks, 0u});
            blocks += (unsigned int)((n + 255) / 256);
        }
        MRMT3_TRY(t->adam_table.reserve(tab.size() * sizeof(AdamSlot)));
        MRMT3_CUDA_TRY(cudaMemcpyAsync(t->adam_table.p, tab.data(), tab.size() * sizeof(AdamSlot), cudaMemcpyHostToDevice, s));
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));  // `tab` is a local
        t->adam_slots = (int)tab.size();
        t->adam_blocks = blocks;
    }
    RUN(h, launch_adamw_multi(t->adam_table.as<AdamSlot>(), t->adam_slots, t->adam_blocks, grad, t->m.as<float>(),
                              t->v.as<float>(), lr, beta1, beta2, adam_eps, wd, t->step, s));
    // masters of the norm-folded decode weights follow their training masters
    auto sync_master = [&](float* dst, int slot) -> Status {
        const ParamSlot& sl = t->slots[slot];
        MRMT3_CUDA_TRY(cudaMemcpyAsync(dst, t->master.as<float>() + sl.off, (size_t)sl.rows * sl.cols * 4,
                                       cudaMemcpyDeviceToDevice, s));
        return OkStatus();
    };
    MRMT3_TRY(sync_master(h->m_lm_head, t->lm_head));
    for (size_t i = 0; i < h->dec.layers.size(); ++i) {
        LayerW& Lw = h->dec.layers[i];
        MRMT3_TRY(sync_master(Lw.m_wqkv, t->dec[i].wqkv));
        MRMT3_TRY(sync_master(Lw.m_cq, t->dec[i].cq));
        MRMT3_TRY(sync_master(Lw.m_wi, t->dec[i].wi));
        RUN(h, launch_fold_norm(Lw.m_wqkv, Lw.ln_self, Lw.wqkv_f, 3 * kInner, kDModel, s));
        RUN(h, launch_fold_norm(Lw.m_cq, Lw.ln_cross, Lw.cq_f, kInner, kDModel, s));
        RUN(h, launch_fold_norm(Lw.m_wi, Lw.ln_ff, Lw.wi_f, 2 * kDFF, kDModel, s));
    }
    RUN(h, launch_fold_norm(h->m_lm_head, h->dec.final_ln, h->lm_head_f, kVocab, kDModel, s));
    return OkStatus();
}

}  // namespace mrmt3
