// Host-side drivers of the path: weight packing, encoder, memory block, cross-K/V projection,
// the KV-cached greedy decode loop (CUDA-graph replay per step, masked early exit) and the
// cross-track segment scheduler of MR-MT3.  Mirrors, stage by stage, what the reference does in
// models/t5.py:251-302 and models/t5_segmem_v2_with_prev.py:226-297 -- with a KV cache instead
// of the reference's full-prefix recompute.
#include "model.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "gemm_skinny.cuh"
#include "gemm_tcgen05.cuh"

using namespace mrmt3;

namespace mrmt3 {

// ---------------------------------------------------------------------------------------------
Status DeviceBuffer::reserve(size_t bytes, bool* moved) {
    if (moved) *moved = false;
    if (bytes <= cap) return OkStatus();
    if (p) MRMT3_CUDA_TRY(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t want = (bytes + 255) & ~size_t(255);
    MRMT3_CUDA_TRY(cudaMalloc(&p, want));
    cap = want;
    if (moved) *moved = true;
    return OkStatus();
}
void DeviceBuffer::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

Status RowWorkspace::reserve(int r) {
    if (r <= rows) return OkStatus();
    size_t n = (size_t)r;
    MRMT3_TRY(x_bf16.reserve(n * kDModel * sizeof(bf16)));
    MRMT3_TRY(h32.reserve(n * kDModel * sizeof(float)));
    MRMT3_TRY(n_bf16.reserve(n * kDModel * sizeof(bf16)));
    MRMT3_TRY(qkv.reserve(n * 3 * kInner * sizeof(bf16)));
    MRMT3_TRY(ctx.reserve(n * kInner * sizeof(bf16)));
    MRMT3_TRY(ff.reserve(n * kDFF * sizeof(bf16)));
    MRMT3_TRY(qc.reserve(n * kInner * sizeof(bf16)));
    rows = r;
    return OkStatus();
}
void RowWorkspace::release() {
    x_bf16.release(); h32.release(); n_bf16.release(); qkv.release(); ctx.release();
    ff.release(); qc.release(); hc32.release(); ctx_c.release();
    rows = 0;
}

constexpr int kEncChunk = 256;    // encoder segments per pass (bounds the row workspace)
constexpr int kNPos = 5000;       // FixedPositionalEmbedding max_length, reference models/t5.py:706
constexpr int kPollEvery = 8;     // decode steps between early-exit polls
constexpr int kMaxGroups = 16;    // lane groups decoding concurrently on their own streams
constexpr int kGroupScalars = 16; // ints per group in the state header: [0] step, [2] ticket
constexpr int kStateHeader = kMaxGroups * kGroupScalars;

static cudaEvent_t prof_event(mrmt3_handle* h) {
    if (!h->prof_pool.empty()) {
        cudaEvent_t e = h->prof_pool.back();
        h->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// decode-step launch tagged with its kernel class; event-bracketed when profiling is on
#define RUNC(h, cat, s, call)                                   \
    do {                                                        \
        if ((h)->prof_on) {                                     \
            mrmt3_handle::ProfRec _r{cat, prof_event(h), prof_event(h)}; \
            cudaEventRecord(_r.a, s);                           \
            MRMT3_TRY(call);                                    \
            cudaEventRecord(_r.b, s);                           \
            (h)->prof_recs.push_back(_r);                       \
        } else {                                                \
            MRMT3_TRY(call);                                    \
        }                                                       \
        ++(h)->launches;                                        \
    } while (0)

Status profile_collect(mrmt3_handle* h) {
    for (auto& r : h->prof_recs) {
        MRMT3_CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0.f;
        MRMT3_CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        h->prof_ms[r.cat] += ms;
        h->prof_n[r.cat] += 1;
        h->prof_pool.push_back(r.a);
        h->prof_pool.push_back(r.b);
    }
    h->prof_recs.clear();
    return OkStatus();
}

static size_t page_elems(const mrmt3_handle* h) {
    return (size_t)h->cfg.n_dec_layers * 2 * kHeads * kKVPage * kDKV;
}

// ---------------------------------------------------------------------------------------------
// weights
static Status arena_take(mrmt3_handle* h, size_t bytes, void** out) {
    size_t off = (h->arena_used + 255) & ~size_t(255);
    if (off + bytes > h->arena.cap) return Error(2, "weight arena overflow");
    *out = (char*)h->arena.p + off;
    h->arena_used = off + bytes;
    return OkStatus();
}

static Status alloc_stack(mrmt3_handle* h, StackW& st, int n_layers, bool decoder) {
    st.layers.resize(n_layers);
    for (auto& L : st.layers) {
        MRMT3_TRY(arena_take(h, (size_t)3 * kInner * kDModel * 2, (void**)&L.wqkv));
        MRMT3_TRY(arena_take(h, (size_t)kDModel * kInner * 2, (void**)&L.wo));
        MRMT3_TRY(arena_take(h, kDModel * 4, (void**)&L.ln_self));
        if (decoder) {
            MRMT3_TRY(arena_take(h, (size_t)3 * kInner * kDModel * 2, (void**)&L.wqkv_f));
            MRMT3_TRY(arena_take(h, (size_t)kInner * kDModel * 2, (void**)&L.cq_f));
            MRMT3_TRY(arena_take(h, (size_t)2 * kDFF * kDModel * 2, (void**)&L.wi_f));
            MRMT3_TRY(arena_take(h, (size_t)3 * kInner * kDModel * 4, (void**)&L.m_wqkv));
            MRMT3_TRY(arena_take(h, (size_t)kInner * kDModel * 4, (void**)&L.m_cq));
            MRMT3_TRY(arena_take(h, (size_t)2 * kDFF * kDModel * 4, (void**)&L.m_wi));
            MRMT3_TRY(arena_take(h, (size_t)kInner * kDModel * 2, (void**)&L.cq));
            MRMT3_TRY(arena_take(h, (size_t)kDModel * kInner * 2, (void**)&L.co));
            MRMT3_TRY(arena_take(h, kDModel * 4, (void**)&L.ln_cross));
        }
        MRMT3_TRY(arena_take(h, (size_t)2 * kDFF * kDModel * 2, (void**)&L.wi));
        MRMT3_TRY(arena_take(h, (size_t)kDModel * kDFF * 2, (void**)&L.wff));
        MRMT3_TRY(arena_take(h, kDModel * 4, (void**)&L.ln_ff));
    }
    MRMT3_TRY(arena_take(h, kDModel * 4, (void**)&st.final_ln));
    return OkStatus();
}

Status handle_init(mrmt3_handle* h) {
    const mrmt3_config& c = h->cfg;
    if (c.d_model != kDModel || c.n_heads != kHeads || c.d_kv != kDKV || c.d_ff != kDFF ||
        c.vocab != kVocab)
        return Error(2, "this build is specialised for the MT3 shape: d_model 512, 6 heads x 64, "
                        "d_ff 1024, vocab 1536 (reference pretrained/config.json)");
    if (c.n_enc_layers < 1 || c.n_dec_layers < 1 || c.n_enc_layers > 64 || c.n_dec_layers > 64)
        return Error(2, "bad layer count");
    if (c.mem_variant < 0 || c.mem_variant > 2) return Error(2, "bad mem_variant");
    if (c.mem_variant != MRMT3_MEM_NONE && (c.n_mem_layers < 1 || c.mem_len < 1 || c.mem_len > 512))
        return Error(2, "memory variant needs n_mem_layers >= 1 and 1 <= mem_len <= 512");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    int n_mem = c.mem_variant ? c.n_mem_layers : 0;
    size_t per_enc = (size_t)(3 * kInner * kDModel + kDModel * kInner + 2 * kDFF * kDModel + kDModel * kDFF) * 2 + 3 * 4096;
    size_t per_dec = per_enc + (size_t)(2 * kInner * kDModel) * 2 + 2 * 4096 +
                     (size_t)(3 * kInner * kDModel + kInner * kDModel + 2 * kDFF * kDModel) * (2 + 4) + 6 * 4096;
    size_t total = (size_t)kDModel * kDModel * 2 * 2 + (size_t)kVocab * kDModel * (4 + 2 + 2 + 4) +
                   (size_t)c.n_dec_layers * 2 * kInner * kDModel * 2 +
                   per_enc * (c.n_enc_layers + n_mem) + per_dec * c.n_dec_layers +
                   (size_t)kNPos * kDModel * 4 + (1 << 20);
    MRMT3_TRY(h->arena.reserve(total));
    MRMT3_CUDA_TRY(cudaMemset(h->arena.p, 0, h->arena.cap));
    MRMT3_TRY(arena_take(h, (size_t)kDModel * kDModel * 2, (void**)&h->proj));
    MRMT3_TRY(arena_take(h, (size_t)kVocab * kDModel * 4, (void**)&h->emb));
    MRMT3_TRY(arena_take(h, (size_t)kVocab * kDModel * 2, (void**)&h->lm_head));
    MRMT3_TRY(arena_take(h, (size_t)kVocab * kDModel * 2, (void**)&h->lm_head_f));
    MRMT3_TRY(arena_take(h, (size_t)kVocab * kDModel * 4, (void**)&h->m_lm_head));
    MRMT3_TRY(arena_take(h, (size_t)c.n_dec_layers * 2 * kInner * kDModel * 2, (void**)&h->cross_kv_w));
    MRMT3_TRY(alloc_stack(h, h->enc, c.n_enc_layers, false));
    MRMT3_TRY(alloc_stack(h, h->dec, c.n_dec_layers, true));
    if (n_mem) {
        MRMT3_TRY(arena_take(h, (size_t)kDModel * kDModel * 2, (void**)&h->segmem_proj));
        MRMT3_TRY(alloc_stack(h, h->mem, n_mem, false));
    }
    MRMT3_TRY(arena_take(h, (size_t)kNPos * kDModel * 4, (void**)&h->pe));
    h->n_pos = kNPos;
    MRMT3_TRY(h->stage.reserve((size_t)2 * kDFF * kDModel * 4));
    MRMT3_TRY(h->frontend.init());
    h->tma = new TmaCache();
    MRMT3_CUDA_TRY(cudaMallocHost((void**)&h->h_pinned, (kMaxLanes + 64) * sizeof(int)));
    MRMT3_CUDA_TRY(cudaEventCreateWithFlags(&h->poll_ev[0], cudaEventDisableTiming));
    MRMT3_CUDA_TRY(cudaEventCreateWithFlags(&h->poll_ev[1], cudaEventDisableTiming));
    const char* ng = getenv("MRMT3_NO_GRAPH");
    h->use_graphs = !(ng && ng[0] == '1');
    if (const char* gl = getenv("MRMT3_GROUP_LANES")) h->group_lanes = atoi(gl);
    if (const char* e = getenv("MRMT3_FUSE_GREEDY")) h->fuse_greedy = atoi(e) != 0;
    if (const char* e = getenv("MRMT3_ATTN_PART_SELF")) h->attn_part_keys_self = atoi(e);
    if (const char* e = getenv("MRMT3_ATTN_PART_CROSS")) h->attn_part_keys_cross = atoi(e);
    {   // tuning sweeps (scripts/): decode-attention kernel selection
        const char* av = getenv("MRMT3_ATTN_VARIANT");
        const char* as = getenv("MRMT3_ATTN_STAGES");
        const char* ac = getenv("MRMT3_ATTN_CTAS");
        const char* aq = getenv("MRMT3_ATTN_QUARTETS");
        attn_decode_configure(av ? atoi(av) : -1, as ? atoi(as) : 0, ac ? atoi(ac) : -1, aq ? atoi(aq) : 0);
    }
    return OkStatus();
}

static void destroy_graphs(mrmt3_handle* h) {
    for (auto& kv : h->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
}

void drop_graphs(mrmt3_handle* h) { destroy_graphs(h); }

void gemm_set_2cta(int on) { gemm_configure_2cta(on); }

Status test_gemm_train(mrmt3_handle* h, const bf16* A, const bf16* W, int M, int N, int K, float* C, int which,
                       cudaStream_t s);  // train.cu

Status test_gemm(mrmt3_handle* h, const bf16* A, const bf16* W, int M, int N, int K, float* C, int which,
                 cudaStream_t s) {
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    if (which == 4 || which == 5) return test_gemm_train(h, A, W, M, N, K, C, which, s);
    const ARowMap id{nullptr, 1};
    if (which == 0) return launch_gemm_mma(A, K, id, W, K, M, N, K, EpiStoreF32{C, N}, s);
    if (which == 1) return launch_gemm_tc(*h->tma, A, K, M, id, W, K, M, N, K, EpiStoreF32{C, N}, s);
    if (which == 3)  // tcgen05 with the bf16 store epilogue the encoder uses (C holds M*N bf16)
        return launch_gemm_tc(*h->tma, A, K, M, id, W, K, M, N, K, EpiStoreBf16{reinterpret_cast<bf16*>(C), N}, s);
    if (which == 6)  // measurement: the same kernel with the epilogue's stores dropped (C is not written)
        return launch_gemm_tc(*h->tma, A, K, M, id, W, K, M, N, K, EpiDiscard{}, s);
    if (which == 2) {
        if (K == 384) return launch_gemm_skinny<32, 384, false>(*h->tma, A, K, W, K, M, N, 0.f, EpiStoreF32{C, N}, s);
        if (K == 512) return launch_gemm_skinny<32, 512, false>(*h->tma, A, K, W, K, M, N, 0.f, EpiStoreF32{C, N}, s);
        if (K == 1024) return launch_gemm_skinny<32, 1024, false>(*h->tma, A, K, W, K, M, N, 0.f, EpiStoreF32{C, N}, s);
    }
    return Error(2, "test_gemm: unsupported kernel / shape");
}

Status trace_enable(mrmt3_handle* h, bool on) {
    destroy_graphs(h);
    h->trace_on = on;
    if (on) {
        MRMT3_TRY(h->trace_buf.reserve(512 * 16));
        MRMT3_CUDA_TRY(cudaMemset(h->trace_buf.p, 0, 512 * 16));
    }
    return OkStatus();
}

// buffers whose address is baked into captured decode-step graphs
static Status reserve_graph_visible(mrmt3_handle* h, DeviceBuffer& b, size_t bytes) {
    bool moved = false;
    MRMT3_TRY(b.reserve(bytes, &moved));
    if (moved) destroy_graphs(h);
    return OkStatus();
}

void handle_destroy(mrmt3_handle* h) {
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    destroy_graphs(h);
    delete h->tma;
    h->tma = nullptr;
    h->frontend.destroy();
    h->arena.release(); h->stage.release(); h->rows.release(); h->enc_bf16.release();
    h->mem_bf16.release(); h->mem_f32.release(); h->mel_f32.release(); h->mel_bf16.release(); h->v1_logits.release();
    h->audio.release(); h->seg_tab.release(); h->ids_dev.release(); h->tok_out.release(); h->dummy_ids.release();
    h->d_h32.release(); h->d_n_bf16.release(); h->d_qkv.release(); h->d_ctx.release();
    h->d_qc.release(); h->d_ff.release(); h->d_logits.release(); h->d_state.release();
    train_destroy(h);
    h->kv_pool.release(); h->block_table.release(); h->cross_cache.release(); h->lane_tab.release();
    h->attn_parts.release(); h->attn_tickets.release(); h->greedy_cand.release(); h->greedy_tickets.release();
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->poll_ev[0]) cudaEventDestroy(h->poll_ev[0]);
    if (h->poll_ev[1]) cudaEventDestroy(h->poll_ev[1]);
    for (int g = 0; g < 16; ++g) {
        if (h->gstream[g]) cudaStreamDestroy(h->gstream[g]);
        if (h->gdone[g]) cudaEventDestroy(h->gdone[g]);
    }
    if (h->gfork) cudaEventDestroy(h->gfork);
}

// train.cu: keeps the fine-tune state consistent with weights (re)loaded after mrmt3_train_init
Status train_on_set_weight(mrmt3_handle* h, const std::string& name, const float* staged_f32);
void train_on_commit(mrmt3_handle* h);

static Status pack(mrmt3_handle* h, const float* src, int rows, int cols, int want_rows,
                   int want_cols, bf16* dst, int row_mul, int row_off, float* master = nullptr) {
    if (rows != want_rows || cols != want_cols) {
        char b[160];
        snprintf(b, sizeof(b), "shape mismatch: got (%d,%d), expected (%d,%d)", rows, cols, want_rows, want_cols);
        return Error(2, b);
    }
    size_t bytes = (size_t)rows * cols * 4;
    MRMT3_TRY(h->stage.reserve(bytes));
    MRMT3_CUDA_TRY(cudaMemcpy(h->stage.p, src, bytes, cudaMemcpyDefault));
    MRMT3_TRY(launch_pack_weight(h->stage.as<float>(), dst, rows, cols, row_mul, row_off, 0));
    if (master) MRMT3_TRY(launch_pack_weight_f32(h->stage.as<float>(), master, rows, cols, row_mul, row_off, 0));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(0));
    h->stage_holds_tensor = true;
    return OkStatus();
}

static Status copy_f32(const float* src, int rows, int cols, int want, float* dst) {
    if (rows * cols != want) return Error(2, "shape mismatch for fp32 vector");
    MRMT3_CUDA_TRY(cudaMemcpy(dst, src, (size_t)want * 4, cudaMemcpyDefault));
    return OkStatus();
}

Status set_weight(mrmt3_handle* h, const std::string& name, const float* data, int rows, int cols) {
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const int d = kDModel, in = kInner;
    Status st = OkStatus();
    bool known = true;
    h->stage_holds_tensor = false;
    if (name == "proj.weight") st = pack(h, data, rows, cols, d, d, h->proj, 1, 0);
    else if (name == "decoder_embed_tokens.weight") st = copy_f32(data, rows, cols, kVocab * d, h->emb);
    else if (name == "lm_head.weight") st = pack(h, data, rows, cols, kVocab, d, h->lm_head, 1, 0, h->m_lm_head);
    else if (name == "segmem_proj.weight") {
        if (!h->segmem_proj) return Error(3, "segmem_proj.weight given to a model without a memory variant");
        st = pack(h, data, rows, cols, d, d, h->segmem_proj, 1, 0);
    } else if (name == "decoder.block.0.layer.1.EncDecAttention.relative_attention_bias.weight") {
        return OkStatus();  // ignored by the reference too (models/t5.py:43-45)
    } else {
        struct { const char* n; StackW* s; bool dec; } stacks[] = {
            {"encoder.", &h->enc, false}, {"decoder.", &h->dec, true}, {"segmem_encoder.", &h->mem, false}};
        known = false;
        for (auto& sk : stacks) {
            size_t pl = strlen(sk.n);
            if (name.compare(0, pl, sk.n) != 0) continue;
            std::string rest = name.substr(pl);
            if (sk.s == &h->mem && !h->segmem_proj) return Error(3, "segmem weights given to a model without a memory variant");
            if (rest == "embed_tokens.weight") return OkStatus();  // alias of proj / embedding / segmem_proj
            if (rest == "pos_emb.inv_freq") {
                if (rows * cols != 256) return Error(2, "inv_freq must have 256 entries");
                MRMT3_CUDA_TRY(cudaMemcpy(h->inv_freq, data, 256 * 4, cudaMemcpyDefault));
                h->have_inv_freq = true;
                return OkStatus();
            }
            if (rest == "final_layer_norm.weight") { known = true; st = copy_f32(data, rows, cols, d, sk.s->final_ln); break; }
            int bi = -1, li = -1, consumed = 0;
            if (sscanf(rest.c_str(), "block.%d.layer.%d.%n", &bi, &li, &consumed) < 2 || consumed == 0) break;
            if (bi < 0 || bi >= (int)sk.s->layers.size()) return Error(3, "block index out of range: " + name);
            LayerW& L = sk.s->layers[bi];
            std::string sub = rest.substr(consumed);
            const int ff_idx = sk.dec ? 2 : 1;
            known = true;
            if (li == 0 && sub == "SelfAttention.q.weight") st = pack(h, data, rows, cols, in, d, L.wqkv, 1, 0, L.m_wqkv);
            else if (li == 0 && sub == "SelfAttention.k.weight") st = pack(h, data, rows, cols, in, d, L.wqkv, 1, in, L.m_wqkv);
            else if (li == 0 && sub == "SelfAttention.v.weight") st = pack(h, data, rows, cols, in, d, L.wqkv, 1, 2 * in, L.m_wqkv);
            else if (li == 0 && sub == "SelfAttention.o.weight") st = pack(h, data, rows, cols, d, in, L.wo, 1, 0);
            else if (li == 0 && sub == "layer_norm.weight") st = copy_f32(data, rows, cols, d, L.ln_self);
            else if (sk.dec && li == 1 && sub == "EncDecAttention.q.weight") st = pack(h, data, rows, cols, in, d, L.cq, 1, 0, L.m_cq);
            else if (sk.dec && li == 1 && sub == "EncDecAttention.k.weight") st = pack(h, data, rows, cols, in, d, h->cross_kv_w, 1, bi * 2 * in);
            else if (sk.dec && li == 1 && sub == "EncDecAttention.v.weight") st = pack(h, data, rows, cols, in, d, h->cross_kv_w, 1, bi * 2 * in + in);
            else if (sk.dec && li == 1 && sub == "EncDecAttention.o.weight") st = pack(h, data, rows, cols, d, in, L.co, 1, 0);
            else if (sk.dec && li == 1 && sub == "layer_norm.weight") st = copy_f32(data, rows, cols, d, L.ln_cross);
            else if (li == ff_idx && sub == "DenseReluDense.wi_0.weight") st = pack(h, data, rows, cols, kDFF, d, L.wi, 2, 0, L.m_wi);
            else if (li == ff_idx && sub == "DenseReluDense.wi_1.weight") st = pack(h, data, rows, cols, kDFF, d, L.wi, 2, 1, L.m_wi);
            else if (li == ff_idx && sub == "DenseReluDense.wo.weight") st = pack(h, data, rows, cols, d, kDFF, L.wff, 1, 0);
            else if (li == ff_idx && sub == "layer_norm.weight") st = copy_f32(data, rows, cols, d, L.ln_ff);
            else known = false;
            break;
        }
    }
    if (!known) return Error(3, "unknown state-dict key: " + name);
    if (!st.ok()) return Error(st.code, name + ": " + st.msg);
    // a (re)load after mrmt3_train_init: the fp32 master of this tensor takes the exact value too
    if (h->train && h->stage_holds_tensor) MRMT3_TRY(train_on_set_weight(h, name, h->stage.as<float>()));
    h->seen.insert(name);
    h->committed = false;
    return OkStatus();
}

static void required_stack(std::vector<std::string>& req, const std::string& s, int n, bool dec) {
    for (int i = 0; i < n; ++i) {
        std::string p = s + ".block." + std::to_string(i) + ".layer.";
        for (const char* w : {"q", "k", "v", "o"}) req.push_back(p + "0.SelfAttention." + w + ".weight");
        req.push_back(p + "0.layer_norm.weight");
        int f = 1;
        if (dec) {
            for (const char* w : {"q", "k", "v", "o"}) req.push_back(p + "1.EncDecAttention." + w + ".weight");
            req.push_back(p + "1.layer_norm.weight");
            f = 2;
        }
        for (const char* w : {"wi_0", "wi_1", "wo"})
            req.push_back(p + std::to_string(f) + ".DenseReluDense." + w + ".weight");
        req.push_back(p + std::to_string(f) + ".layer_norm.weight");
    }
    req.push_back(s + ".final_layer_norm.weight");
}

Status commit_weights(mrmt3_handle* h) {
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    std::vector<std::string> req = {"proj.weight", "decoder_embed_tokens.weight", "lm_head.weight"};
    required_stack(req, "encoder", h->cfg.n_enc_layers, false);
    required_stack(req, "decoder", h->cfg.n_dec_layers, true);
    if (h->cfg.mem_variant) {
        req.push_back("segmem_proj.weight");
        required_stack(req, "segmem_encoder", h->cfg.n_mem_layers, false);
    }
    for (auto& r : req)
        if (!h->seen.count(r)) return Error(4, "missing weight: " + r);
    // sinusoid table, reference FixedPositionalEmbedding (models/t5.py:705-719): fp32 product
    // position * inv_freq, then sin | cos halves
    if (!h->have_inv_freq)
        for (int j = 0; j < 256; ++j) h->inv_freq[j] = 1.0f / powf(10000.0f, (float)(2 * j) / (float)kDModel);
    std::vector<float> pe((size_t)kNPos * kDModel);
    for (int p = 0; p < kNPos; ++p)
        for (int j = 0; j < 256; ++j) {
            float ang = (float)p * h->inv_freq[j];
            pe[(size_t)p * kDModel + j] = (float)std::sin((double)ang);
            pe[(size_t)p * kDModel + 256 + j] = (float)std::cos((double)ang);
        }
    MRMT3_CUDA_TRY(cudaMemcpy(h->pe, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice));
    // norm-folded decoder weights for the fused-RMSNorm decode GEMMs (gemm_skinny.cuh)
    for (auto& L : h->dec.layers) {
        MRMT3_TRY(launch_fold_norm(L.m_wqkv, L.ln_self, L.wqkv_f, 3 * kInner, kDModel, 0));
        MRMT3_TRY(launch_fold_norm(L.m_cq, L.ln_cross, L.cq_f, kInner, kDModel, 0));
        MRMT3_TRY(launch_fold_norm(L.m_wi, L.ln_ff, L.wi_f, 2 * kDFF, kDModel, 0));
    }
    MRMT3_TRY(launch_fold_norm(h->m_lm_head, h->dec.final_ln, h->lm_head_f, kVocab, kDModel, 0));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(0));
    h->committed = true;
    train_on_commit(h);
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// encoder-type stack over `rows` rows organised as sequences of T rows.  Input: rows.h32 holds
// embed + PE.  q_rows < T restricts the LAST layer's queries (and everything after it) to the
// first q_rows rows of each sequence (memory block: only L_agg outputs are consumed, and a
// row's output depends on other rows only through K/V, so this is exact).  Output: final-normed
// states as bf16 (out_bf16) and/or fp32 (out_f32), (n_seq * q_rows, d).
static Status run_encoder_stack(mrmt3_handle* h, const StackW& st, int n_seq, int T, int q_rows,
                                bf16* out_bf16, float* out_f32, cudaStream_t s) {
    RowWorkspace& w = h->rows;
    const int M = n_seq * T;
    float* hcur = w.h32.as<float>();
    int Mcur = M;
    const float eps = h->cfg.ln_eps;
    for (size_t li = 0; li < st.layers.size(); ++li) {
        const LayerW& L = st.layers[li];
        const bool reduce = (li + 1 == st.layers.size()) && q_rows < T;
        RUN(h, launch_rmsnorm(hcur, L.ln_self, eps, w.n_bf16.as<bf16>(), nullptr, M, nullptr, 1, s));
        RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, ARowMap{nullptr, 1}, L.wqkv, kDModel, M,
                               3 * kInner, kDModel, EpiStoreBf16{w.qkv.as<bf16>(), 3 * kInner}, s));
        AttnFullParams ap{};
        ap.Q = w.qkv.as<bf16>();
        ap.K = ap.Q + kInner;
        ap.V = ap.Q + 2 * kInner;
        ap.q_batch_stride = ap.k_batch_stride = ap.v_batch_stride = (long)T * 3 * kInner;
        ap.q_head_stride = ap.k_head_stride = ap.v_head_stride = kDKV;
        ap.q_row_stride = ap.k_row_stride = ap.v_row_stride = 3 * kInner;
        ap.Tk = T;
        ap.causal = 0;
        ap.causal_offset = 0;
        ap.o_head_stride = kDKV;
        ap.o_row_stride = kInner;
        if (reduce) {
            MRMT3_TRY(w.ctx_c.reserve((size_t)n_seq * q_rows * kInner * sizeof(bf16)));
            MRMT3_TRY(w.hc32.reserve((size_t)n_seq * q_rows * kDModel * sizeof(float)));
            ap.Tq = q_rows;
            ap.O = w.ctx_c.as<bf16>();
            ap.o_batch_stride = (long)q_rows * kInner;
            RUN(h, launch_attn_full_auto(*h->tma, ap, n_seq, s));
            RUN(h, launch_gather_rows(hcur, w.hc32.as<float>(), n_seq, q_rows, T, s));
            hcur = w.hc32.as<float>();
            Mcur = n_seq * q_rows;
            RUN(h, launch_gemm_tc(*h->tma, w.ctx_c.as<bf16>(), kInner, Mcur, ARowMap{nullptr, 1}, L.wo, kInner, Mcur,
                                   kDModel, kInner, EpiResidual{hcur, kDModel}, s));
        } else {
            ap.Tq = T;
            ap.O = w.ctx.as<bf16>();
            ap.o_batch_stride = (long)T * kInner;
            RUN(h, launch_attn_full_auto(*h->tma, ap, n_seq, s));
            RUN(h, launch_gemm_tc(*h->tma, w.ctx.as<bf16>(), kInner, M, ARowMap{nullptr, 1}, L.wo, kInner, M, kDModel,
                                   kInner, EpiResidual{hcur, kDModel}, s));
        }
        RUN(h, launch_rmsnorm(hcur, L.ln_ff, eps, w.n_bf16.as<bf16>(), nullptr, Mcur, nullptr, 1, s));
        RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, Mcur, ARowMap{nullptr, 1}, L.wi, kDModel, Mcur,
                               2 * kDFF, kDModel, EpiGatedGelu{w.ff.as<bf16>(), kDFF}, s));
        RUN(h, launch_gemm_tc(*h->tma, w.ff.as<bf16>(), kDFF, Mcur, ARowMap{nullptr, 1}, L.wff, kDFF, Mcur, kDModel,
                               kDFF, EpiResidual{hcur, kDModel}, s));
    }
    if (Mcur == M && q_rows < T) {
        // no layer reduced (cannot happen with >= 1 layer), keep the contract anyway
        MRMT3_TRY(w.hc32.reserve((size_t)n_seq * q_rows * kDModel * sizeof(float)));
        RUN(h, launch_gather_rows(hcur, w.hc32.as<float>(), n_seq, q_rows, T, s));
        hcur = w.hc32.as<float>();
        Mcur = n_seq * q_rows;
    }
    RUN(h, launch_rmsnorm(hcur, st.final_ln, eps, out_bf16, out_f32, Mcur, nullptr, 1, s));
    return OkStatus();
}

// proj + encoder over n_seg segments (reference models/t5.py:253-258).  Exactly one of
// mel_f32 / mel_bf16 is given.
static Status encode_segments(mrmt3_handle* h, const float* mel_f32, const bf16* mel_bf16, int n_seg,
                              bf16* enc_out_bf16, float* enc_out_f32, cudaStream_t s) {
    for (int c0 = 0; c0 < n_seg; c0 += kEncChunk) {
        const int n = std::min(kEncChunk, n_seg - c0);
        const int M = n * kSegFrames;
        MRMT3_TRY(h->rows.reserve(M));
        const bf16* x;
        if (mel_bf16) {
            x = mel_bf16 + (size_t)c0 * kSegFrames * kMels;
        } else {
            RUN(h, launch_cast_bf16(mel_f32 + (size_t)c0 * kSegFrames * kMels, h->rows.x_bf16.as<bf16>(),
                                    (size_t)M * kMels, s));
            x = h->rows.x_bf16.as<bf16>();
        }
        RUN(h, launch_gemm_tc(*h->tma, x, kDModel, M, ARowMap{nullptr, 1}, h->proj, kDModel, M, kDModel, kDModel,
                               EpiPosAdd{h->rows.h32.as<float>(), kDModel, h->pe, kSegFrames, 0}, s));
        MRMT3_TRY(run_encoder_stack(
            h, h->enc, n, kSegFrames, kSegFrames,
            enc_out_bf16 ? enc_out_bf16 + (size_t)c0 * kSegFrames * kDModel : nullptr,
            enc_out_f32 ? enc_out_f32 + (size_t)c0 * kSegFrames * kDModel : nullptr, s));
    }
    return OkStatus();
}

// MR-MT3 memory block (reference models/t5_segmem_v2_with_prev.py:263-266):
// Emb[ids] -> segmem_proj -> +PE -> memory encoder over all Lp positions -> first n_mem rows.
// ids of lane l: ids[(src_row ? src_row[l] : l) * ids_stride + 0..Lp)
static Status memory_block(mrmt3_handle* h, const long long* ids, long ids_stride, const int* src_row,
                           int n_lanes, int Lp, int n_mem, bf16* out_bf16, float* out_f32,
                           cudaStream_t s) {
    const int chunk = std::max(1, (kEncChunk * kSegFrames) / Lp);
    for (int c0 = 0; c0 < n_lanes; c0 += chunk) {
        const int n = std::min(chunk, n_lanes - c0);
        const int M = n * Lp;
        MRMT3_TRY(h->rows.reserve(M));
        RUN(h, launch_embed_bf16(src_row ? ids : ids + (size_t)c0 * ids_stride, ids_stride, Lp, n,
                                 src_row ? src_row + c0 : nullptr, h->emb, h->rows.x_bf16.as<bf16>(), s));
        RUN(h, launch_gemm_tc(*h->tma, h->rows.x_bf16.as<bf16>(), kDModel, M, ARowMap{nullptr, 1}, h->segmem_proj,
                              kDModel, M, kDModel, kDModel,
                              EpiPosAdd{h->rows.h32.as<float>(), kDModel, h->pe, Lp, 0}, s));
        MRMT3_TRY(run_encoder_stack(h, h->mem, n, Lp, n_mem,
                                    out_bf16 ? out_bf16 + (size_t)c0 * n_mem * kDModel : nullptr,
                                    out_f32 ? out_f32 + (size_t)c0 * n_mem * kDModel : nullptr, s));
    }
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// decode lanes
struct LaneArrays {
    int *step, *n_active, *ticket, *tok, *active, *finish_step;
    int *out_row, *prev_row, *init_active, *seg_index;
};

static LaneArrays lane_arrays(const mrmt3_handle* h) {
    LaneArrays a;
    int* s = h->d_state.as<int>();
    a.step = s;
    a.n_active = s + 1;
    a.ticket = s + 2;
    a.tok = s + kStateHeader;
    a.active = a.tok + h->lane_cap;
    a.finish_step = a.active + h->lane_cap;
    int* t = h->lane_tab.as<int>();
    a.out_row = t;
    a.prev_row = t + h->lane_cap;
    a.init_active = t + 2 * h->lane_cap;
    a.seg_index = t + 3 * h->lane_cap;
    return a;
}

Status ensure_decode_capacity(mrmt3_handle* h, int n_lanes, int tk, int max_positions) {
    const int pages = ceil_div(max_positions, kKVPage);
    if (max_positions > attn_decode_max_keys())
        return Error(2, "max_length (+ memory prefix) exceeds the decode attention key capacity");
    if (n_lanes <= h->lane_cap && tk <= h->tk_cap && pages <= h->page_cap) return OkStatus();
    MRMT3_CUDA_TRY(cudaDeviceSynchronize());
    destroy_graphs(h);
    const int cap = std::max(n_lanes, h->lane_cap);
    const int tkc = std::max(tk, h->tk_cap);
    const int pgc = std::max(pages, h->page_cap);
    const size_t c = (size_t)cap;
    MRMT3_TRY(h->d_h32.reserve(c * kDModel * 4));
    MRMT3_TRY(h->d_n_bf16.reserve(c * kDModel * 2));
    MRMT3_TRY(h->d_qkv.reserve(c * 3 * kInner * 2));
    MRMT3_TRY(h->d_ctx.reserve(c * kInner * 2));
    MRMT3_TRY(h->d_qc.reserve(c * kInner * 2));
    MRMT3_TRY(h->d_ff.reserve(c * kDFF * 2));
    MRMT3_TRY(h->d_logits.reserve(c * kVocab * 4));
    MRMT3_TRY(h->d_state.reserve((kStateHeader + 3 * c) * sizeof(int)));
    MRMT3_TRY(h->lane_tab.reserve(4 * c * sizeof(int)));
    MRMT3_TRY(h->kv_pool.reserve(c * pgc * page_elems(h) * sizeof(bf16)));
    MRMT3_TRY(h->cross_cache.reserve(c * h->cfg.n_dec_layers * 2 * kHeads * (size_t)tkc * kDKV * sizeof(bf16)));
    MRMT3_TRY(h->block_table.reserve(c * pgc * sizeof(int)));
    h->attn_max_parts = std::max(pgc * kKVPage, tkc) / 128 + 1;   // parts are >= 128 keys
    MRMT3_TRY(h->attn_parts.reserve(c * kHeads * h->attn_max_parts * 66 * sizeof(float)));
    MRMT3_TRY(h->attn_tickets.reserve(c * kHeads * sizeof(int)));
    MRMT3_CUDA_TRY(cudaMemset(h->attn_tickets.p, 0, h->attn_tickets.cap));
    MRMT3_TRY(h->greedy_cand.reserve(c * (kVocab / 64) * sizeof(float2)));
    MRMT3_TRY(h->greedy_tickets.reserve(c * sizeof(int)));
    MRMT3_CUDA_TRY(cudaMemset(h->greedy_tickets.p, 0, h->greedy_tickets.cap));

    MRMT3_CUDA_TRY(cudaMemset(h->d_h32.p, 0, h->d_h32.cap));
    // the TMA attention kernel loads whole 32-row boxes and masks the rows past the valid keys
    // with p = 0: cache memory must never hold a NaN/Inf bit pattern
    MRMT3_CUDA_TRY(cudaMemset(h->kv_pool.p, 0, h->kv_pool.cap));
    MRMT3_CUDA_TRY(cudaMemset(h->cross_cache.p, 0, h->cross_cache.cap));
    MRMT3_CUDA_TRY(cudaMemset(h->d_state.p, 0, h->d_state.cap));
    MRMT3_CUDA_TRY(cudaMemset(h->lane_tab.p, 0, h->lane_tab.cap));
    // static page assignment: lane l owns pages [l*pgc, (l+1)*pgc)
    std::vector<int> bt(c * pgc);
    for (size_t i = 0; i < bt.size(); ++i) bt[i] = (int)i;
    MRMT3_CUDA_TRY(cudaMemcpy(h->block_table.p, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->lane_cap = cap;
    h->tk_cap = tkc;
    h->page_cap = pgc;
    return OkStatus();
}

struct StepPlan {
    int lane0;             // first lane of this group (all per-lane buffers are offset by it)
    int n_lanes, tk;
    DecodeState st;
    const float* prefix;   // V1 memory prefix rows (lanes, prefix_len, d) fp32 or nullptr
    float* ext_logits;     // optional (rows, max_tokens, V) sink
};

// all kernels of one decode step.  kind 0: token step (lm_head + arg-max); kind 1: memory-prefix
// step of the V1 variant (no token is produced).
static Status enqueue_step(mrmt3_handle* h, const StepPlan& pl, int kind, cudaStream_t s) {
    const int n = pl.n_lanes;
    const float eps = h->cfg.ln_eps;
    const size_t l0 = (size_t)pl.lane0;
    float* H = h->d_h32.as<float>() + l0 * kDModel;      // fp32 residual stream
    bf16* Hb = h->d_n_bf16.as<bf16>() + l0 * kDModel;    // its bf16 copy (input of the NORM GEMMs)
    bf16* qkv = h->d_qkv.as<bf16>() + l0 * 3 * kInner;
    bf16* ctx = h->d_ctx.as<bf16>() + l0 * kInner;
    bf16* qc = h->d_qc.as<bf16>() + l0 * kInner;
    bf16* ff = h->d_ff.as<bf16>() + l0 * kDFF;
    const size_t cross_lane = (size_t)h->cfg.n_dec_layers * 2 * kHeads * h->tk_cap * kDKV;
    // tensor maps of the two caches viewed as (rows, 64) bf16 for the TMA attention kernel
    const CUtensorMap *self_map = nullptr, *cross_map = nullptr;
    MRMT3_TRY(h->tma->get(h->kv_pool.p, (long)(h->kv_pool.cap / (kDKV * sizeof(bf16))), kDKV, kDKV, 32, &self_map));
    MRMT3_TRY(h->tma->get(h->cross_cache.p, (long)(h->cross_cache.cap / (kDKV * sizeof(bf16))), kDKV, kDKV, 32, &cross_map));
    // column-tile width of the wide projections (qkv, ffn-in): a CTA ingests (32 + BN) x K x 2
    // bytes at the ~70 GB/s one SM gets from L2, so narrow tiles on more SMs win while the launch
    // has fewer CTAs than the GPU has SMs
    static const int narrow_env = getenv("MRMT3_SKINNY_NARROW") ? atoi(getenv("MRMT3_SKINNY_NARROW")) : -1;
    const bool narrow = narrow_env >= 0 ? narrow_env != 0 : n <= 64;
    int tslot = 0;
    auto next_trace = [&]() { return TraceSlot{h->trace_on ? h->trace_buf.as<unsigned long long>() : nullptr, tslot++}; };
    // Plain token steps (no memory prefix, no logits sink) end in the fused greedy head, which also
    // writes the next step's input rows: such a step has no embed and no arg-max kernel (the rows of
    // step 0 come from one embed launch ahead of the loop, run_decode).
    const bool fused_head = h->fuse_greedy && kind == 0 && !pl.ext_logits;
    const bool fused_embed = fused_head && pl.st.prefix_len == 0 && !pl.prefix;
    DecodeState st_embed = pl.st;
    st_embed.trace = next_trace();
    if (!fused_embed)
        RUNC(h, MRMT3_PROF_EMBED, s, launch_decode_embed(st_embed, h->emb, h->pe,
                                                        pl.prefix ? pl.prefix + l0 * pl.st.prefix_len * kDModel : nullptr,
                                                        pl.st.prefix_len * kDModel, H, Hb, n, s));
    for (int li = 0; li < h->cfg.n_dec_layers; ++li) {
        const LayerW& L = h->dec.layers[li];
        if (narrow)
            RUNC(h, MRMT3_PROF_GEMM_QKV, s, (launch_gemm_skinny<32, kDModel, true>(
                     *h->tma, Hb, kDModel, L.wqkv_f, kDModel, n, 3 * kInner, eps, EpiStoreBf16{qkv, 3 * kInner}, s, next_trace())));
        else
            RUNC(h, MRMT3_PROF_GEMM_QKV, s, (launch_gemm_skinny<64, kDModel, true>(
                     *h->tma, Hb, kDModel, L.wqkv_f, kDModel, n, 3 * kInner, eps, EpiStoreBf16{qkv, 3 * kInner}, s, next_trace())));
        AttnDecodeParams ap{};
        ap.q = qkv;
        ap.q_stride = 3 * kInner;
        ap.out = ctx;
        ap.out_stride = kInner;
        ap.kv_pool = h->kv_pool.as<bf16>();
        ap.layer = li;
        ap.n_layers = h->cfg.n_dec_layers;
        ap.step_ptr = pl.st.step;
        ap.pos_offset = 0;
        ap.block_table = h->block_table.as<int>() + l0 * h->page_cap;
        ap.max_pages = h->page_cap;
        ap.page_stride = page_elems(h);
        ap.active = pl.st.active;
        ap.tmap = self_map;
        ap.tmap_row0 = 0;
        ap.part_keys = h->attn_part_keys_self;
        ap.max_parts = h->attn_max_parts;
        ap.part_scratch = h->attn_parts.as<float>() + l0 * kHeads * h->attn_max_parts * 66;
        ap.part_counter = h->attn_tickets.as<int>() + l0 * kHeads;
        ap.trace = next_trace();
        RUNC(h, MRMT3_PROF_ATTN_SELF, s, launch_attn_decode(ap, n, true, s));
        RUNC(h, MRMT3_PROF_GEMM_O, s, (launch_gemm_skinny<32, kInner, false>(
                 *h->tma, ctx, kInner, L.wo, kInner, n, kDModel, eps, EpiResidualBoth{H, Hb, kDModel}, s, next_trace())));

        RUNC(h, MRMT3_PROF_GEMM_CQ, s, (launch_gemm_skinny<32, kDModel, true>(
                 *h->tma, Hb, kDModel, L.cq_f, kDModel, n, kInner, eps, EpiStoreBf16{qc, kInner}, s, next_trace())));
        AttnDecodeParams cp{};
        cp.q = qc;
        cp.q_stride = kInner;
        cp.out = ctx;
        cp.out_stride = kInner;
        cp.kv_pool = h->cross_cache.as<bf16>() + l0 * cross_lane;
        cp.layer = li;
        cp.n_layers = h->cfg.n_dec_layers;
        cp.tk_cap = h->tk_cap;
        cp.n_keys = pl.tk;
        cp.active = pl.st.active;
        cp.tmap = cross_map;
        cp.tmap_row0 = (long long)(l0 * cross_lane / kDKV);
        cp.part_keys = h->attn_part_keys_cross;
        cp.max_parts = h->attn_max_parts;
        cp.part_scratch = ap.part_scratch;
        cp.part_counter = ap.part_counter;
        cp.trace = next_trace();
        RUNC(h, MRMT3_PROF_ATTN_CROSS, s, launch_attn_decode(cp, n, false, s));
        RUNC(h, MRMT3_PROF_GEMM_CO, s, (launch_gemm_skinny<32, kInner, false>(
                 *h->tma, ctx, kInner, L.co, kInner, n, kDModel, eps, EpiResidualBoth{H, Hb, kDModel}, s, next_trace())));

        if (narrow)
            RUNC(h, MRMT3_PROF_GEMM_WI, s, (launch_gemm_skinny<32, kDModel, true>(
                     *h->tma, Hb, kDModel, L.wi_f, kDModel, n, 2 * kDFF, eps, EpiGatedGelu{ff, kDFF}, s, next_trace())));
        else
            RUNC(h, MRMT3_PROF_GEMM_WI, s, (launch_gemm_skinny<64, kDModel, true>(
                     *h->tma, Hb, kDModel, L.wi_f, kDModel, n, 2 * kDFF, eps, EpiGatedGelu{ff, kDFF}, s, next_trace())));
        RUNC(h, MRMT3_PROF_GEMM_WFF, s, (launch_gemm_skinny<32, kDFF, false>(
                 *h->tma, ff, kDFF, L.wff, kDFF, n, kDModel, eps, EpiResidualBoth{H, Hb, kDModel}, s, next_trace())));
    }
    if (kind == 1) {
        RUNC(h, MRMT3_PROF_ARGMAX, s, launch_advance_only(pl.st, s));
        return OkStatus();
    }
    // final norm is folded into lm_head_f
    if (fused_head) {
        EpiGreedy eg{};
        eg.st = pl.st;
        eg.cand = h->greedy_cand.as<float2>() + l0 * (kVocab / 64);
        eg.tile_ticket = h->greedy_tickets.as<int>() + l0;
        eg.vocab = kVocab;
        eg.emb = fused_embed ? h->emb : nullptr;
        eg.pe = h->pe;
        eg.H = H;
        eg.Hb = Hb;
        RUNC(h, MRMT3_PROF_LM_HEAD, s, (launch_gemm_skinny<64, kDModel, true>(
                 *h->tma, Hb, kDModel, h->lm_head_f, kDModel, n, kVocab, eps, eg, s, next_trace())));
        next_trace();  // the arg-max slot stays empty
        return OkStatus();
    }
    if (pl.ext_logits) {
        const size_t lane_stride = (size_t)pl.st.max_tokens * kVocab;
        RUNC(h, MRMT3_PROF_LM_HEAD, s, (launch_gemm_skinny<64, kDModel, true>(
                 *h->tma, Hb, kDModel, h->lm_head_f, kDModel, n, kVocab, eps,
                 EpiStoreF32Step{pl.ext_logits, kVocab, lane_stride, pl.st.step, pl.st.prefix_len, pl.st.out_row}, s, next_trace())));
        RUNC(h, MRMT3_PROF_ARGMAX, s, launch_argmax_advance(pl.st, pl.ext_logits, lane_stride, kVocab, n, kVocab, s));
    } else {
        float* lg = h->d_logits.as<float>() + l0 * kVocab;
        RUNC(h, MRMT3_PROF_LM_HEAD, s, (launch_gemm_skinny<64, kDModel, true>(
                 *h->tma, Hb, kDModel, h->lm_head_f, kDModel, n, kVocab, eps, EpiStoreF32{lg, kVocab}, s, next_trace())));
        DecodeState st_arg = pl.st;
        st_arg.trace = next_trace();
        RUNC(h, MRMT3_PROF_ARGMAX, s, launch_argmax_advance(st_arg, lg, kVocab, 0, n, kVocab, s));
    }
    return OkStatus();
}

static Status get_graph(mrmt3_handle* h, const StepPlan& pl, int kind, StepGraph** out) {
    StepGraphKey key{pl.lane0, pl.n_lanes, pl.tk, pl.st.max_tokens, pl.st.prefix_len, kind, pl.st.forced, pl.ext_logits};
    auto it = h->graphs.find(key);
    if (it != h->graphs.end()) {
        *out = &it->second;
        return OkStatus();
    }
    cudaStream_t cs;
    MRMT3_CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    StepGraph g;
    int64_t before = h->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    Status st = OkStatus();
    if (e == cudaSuccess) {
        st = enqueue_step(h, pl, kind, cs);
        e = cudaStreamEndCapture(cs, &graph);
    }
    g.n_kernels = (int)(h->launches - before);
    h->launches = before;  // capture launches nothing; replays are counted when launched
    if (st.ok() && e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cs);
    if (!st.ok()) return st;
    if (e != cudaSuccess) return Error((int)e, std::string("decode-step graph capture failed: ") + cudaGetErrorString(e));
    auto ins = h->graphs.emplace(key, g);
    *out = &ins.first->second;
    return OkStatus();
}

// view of the decode state for lanes [lane0, lane0 + n): per-lane arrays are offset, the step and
// ticket scalars are the group's own, n_active is shared by all groups
static DecodeState group_state(const DecodeState& st, int lane0, int group) {
    DecodeState g = st;
    g.step = st.step + group * kGroupScalars;
    g.ticket = st.ticket + group * kGroupScalars;
    g.tok = st.tok + lane0;
    g.active = st.active + lane0;
    g.finish_step = st.finish_step + lane0;
    g.out_row = st.out_row + lane0;
    if (st.forced && !st.forced_by_row) g.forced = st.forced + (size_t)lane0 * st.forced_stride;
    return g;
}

// Greedy loop over lanes whose state was set by launch_decode_init.  Runs n_prefix prefix steps,
// then up to max_tokens token steps, stopping early once every lane has emitted EOS (polled
// every kPollEvery steps without draining the queue).
//
// The lanes are split into groups that run the same loop independently, each replaying its own
// step graph on its own stream: the chain of one step is ~90 dependent kernels, half of them
// small latency-bound projections, so a single chain leaves HBM idle most of the time; several
// chains in flight let one group's attention (HBM-bound) overlap the others' projections.
static Status run_decode(mrmt3_handle* h, StepPlan pl, int n_prefix, bool debug_mode, cudaStream_t s) {
    if (h->hooks_fast_path) debug_mode = false;
    if (h->graphs.size() > 256) {  // hook pointers are part of the key and change per call: bound the cache
        MRMT3_CUDA_TRY(cudaDeviceSynchronize());
        destroy_graphs(h);
    }
    const bool graphs = h->use_graphs && !debug_mode && !h->prof_on;
    int G = 1;
    // default group size (group_lanes < 0): 32 lanes, 16 for small batches (MR-MT3 with one lane per
    // track: 64 tracks decode fastest as 4 groups; more groups than that only add launches)
    const int gl = h->group_lanes >= 0 ? h->group_lanes : (pl.n_lanes <= 64 ? 16 : 32);
    if (!debug_mode && !h->prof_on && gl > 0)
        G = std::min(kMaxGroups, ceil_div(pl.n_lanes, gl));
    const int per = ceil_div(pl.n_lanes, G);
    G = ceil_div(pl.n_lanes, per);
    std::vector<StepPlan> gp(G);
    for (int g = 0; g < G; ++g) {
        gp[g] = pl;
        gp[g].lane0 = pl.lane0 + g * per;
        gp[g].n_lanes = std::min(per, pl.n_lanes - g * per);
        gp[g].st = group_state(pl.st, g * per, g);
    }
    std::vector<cudaStream_t> gs(G, s);
    if (G > 1 && !h->group_serial) {
        if (!h->gfork) MRMT3_CUDA_TRY(cudaEventCreateWithFlags(&h->gfork, cudaEventDisableTiming));
        MRMT3_CUDA_TRY(cudaEventRecord(h->gfork, s));
        for (int g = 0; g < G; ++g) {
            if (!h->gstream[g]) {
                MRMT3_CUDA_TRY(cudaStreamCreateWithFlags(&h->gstream[g], cudaStreamNonBlocking));
                MRMT3_CUDA_TRY(cudaEventCreateWithFlags(&h->gdone[g], cudaEventDisableTiming));
            }
            gs[g] = h->gstream[g];
            MRMT3_CUDA_TRY(cudaStreamWaitEvent(gs[g], h->gfork, 0));
        }
    }
    auto step_all = [&](int kind) -> Status {
        for (int g = 0; g < G; ++g) {
            if (graphs) {
                StepGraph* sg = nullptr;
                MRMT3_TRY(get_graph(h, gp[g], kind, &sg));
                MRMT3_CUDA_TRY(cudaGraphLaunch(sg->exec, gs[g]));
                h->launches += sg->n_kernels;
            } else {
                MRMT3_TRY(enqueue_step(h, gp[g], kind, gs[g]));
            }
        }
        return OkStatus();
    };
    for (int i = 0; i < n_prefix; ++i) MRMT3_TRY(step_all(1));
    if (h->fuse_greedy && !pl.ext_logits && pl.st.prefix_len == 0 && !pl.prefix) {
        // the fused greedy head writes the input rows of step t + 1; step 0's rows come from here
        for (int g = 0; g < G; ++g) {
            const size_t l0 = (size_t)gp[g].lane0;
            RUNC(h, MRMT3_PROF_EMBED, gs[g], launch_decode_embed(gp[g].st, h->emb, h->pe, nullptr, 0,
                                                               h->d_h32.as<float>() + l0 * kDModel,
                                                               h->d_n_bf16.as<bf16>() + l0 * kDModel, gp[g].n_lanes, gs[g]));
        }
    }
    int polls = 0;
    bool pending[2] = {false, false};
    bool all_done = false;
    const bool may_exit_early = pl.st.forced == nullptr;
    for (int t = 0; t < pl.st.max_tokens && !all_done; ++t) {
        MRMT3_TRY(step_all(0));
        if (may_exit_early && (t + 1) % kPollEvery == 0 && t + 1 < pl.st.max_tokens) {
            // n_active is decremented by every group; a read through any stream can only be
            // stale-high, which delays the exit but never cuts a lane short
            const int slot = polls & 1;
            MRMT3_CUDA_TRY(cudaMemcpyAsync(&h->h_pinned[slot], pl.st.n_active, sizeof(int),
                                           cudaMemcpyDeviceToHost, gs[G - 1]));
            MRMT3_CUDA_TRY(cudaEventRecord(h->poll_ev[slot], gs[G - 1]));
            pending[slot] = true;
            const int prev = slot ^ 1;
            if (pending[prev]) {  // look at the poll issued kPollEvery steps ago
                MRMT3_CUDA_TRY(cudaEventSynchronize(h->poll_ev[prev]));
                pending[prev] = false;
                if (h->h_pinned[prev] == 0) all_done = true;
            }
            ++polls;
        }
    }
    if (G > 1 && !h->group_serial) {
        for (int g = 0; g < G; ++g) {
            MRMT3_CUDA_TRY(cudaEventRecord(h->gdone[g], gs[g]));
            MRMT3_CUDA_TRY(cudaStreamWaitEvent(s, h->gdone[g], 0));
        }
    }
    MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    if (h->prof_on) MRMT3_TRY(profile_collect(h));
    return OkStatus();
}

static DecodeState make_state(mrmt3_handle* h, long long* out, int out_stride, int max_tokens,
                              int prefix_len, const long long* forced) {
    LaneArrays a = lane_arrays(h);
    DecodeState st{};
    st.step = a.step;
    st.tok = a.tok;
    st.active = a.active;
    st.finish_step = a.finish_step;
    st.n_active = a.n_active;
    st.ticket = a.ticket;
    st.out = out;
    st.out_row = a.out_row;
    st.out_stride = out_stride;
    st.prefix_len = prefix_len;
    st.forced = forced;
    st.forced_stride = out_stride;
    st.eos_id = h->cfg.eos_id;
    st.pad_id = h->cfg.pad_id;
    st.max_tokens = max_tokens;
    return st;
}

// cross K/V of every decoder layer for `n_lanes` lanes: encoder rows (gathered through
// seg_index when given) and, for the V2 variant, the memory rows appended at key 256.
static Status project_cross_kv(mrmt3_handle* h, const bf16* enc, long enc_rows, const int* seg_index, int n_lanes,
                               const bf16* mem, int n_mem, cudaStream_t s) {
    const int N = h->cfg.n_dec_layers * 2 * kInner;
    RUN(h, launch_gemm_tc(*h->tma, enc, kDModel, enc_rows, ARowMap{seg_index, kSegFrames}, h->cross_kv_w, kDModel,
                          n_lanes * kSegFrames, N, kDModel,
                           EpiCrossKV{h->cross_cache.as<bf16>(), kSegFrames, 0, h->cfg.n_dec_layers,
                                      h->tk_cap, nullptr}, s));
    if (mem && n_mem > 0)
        RUN(h, launch_gemm_tc(*h->tma, mem, kDModel, n_lanes * n_mem, ARowMap{nullptr, 1}, h->cross_kv_w, kDModel,
                              n_lanes * n_mem, N, kDModel,
                               EpiCrossKV{h->cross_cache.as<bf16>(), n_mem, kSegFrames,
                                          h->cfg.n_dec_layers, h->tk_cap, nullptr}, s));
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// public drivers
Status api_encode(mrmt3_handle* h, const float* mel, int B, float* enc_out, cudaStream_t s) {
    if (!h->committed) return Error(5, "weights not committed");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    return encode_segments(h, mel, nullptr, B, nullptr, enc_out, s);
}

// shared by mrmt3_generate and the e2e path (mel as fp32 or bf16)
Status generate_base(mrmt3_handle* h, const float* mel_f32, const bf16* mel_bf16, int B, int max_length,
                     long long* out_ids, int* steps_host, const long long* forced, float* logits_out,
                     cudaStream_t s) {
    if (!h->committed) return Error(5, "weights not committed");
    // a handle with a memory variant may run this loop too: T5SegMem.generate (reference
    // models/t5_segmem.py:254-311) is the plain batched loop, its memory weights unused
    if (B <= 0 || max_length <= 0) return Error(2, "B and max_length must be positive");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const int stride = max_length + 1;
    MRMT3_TRY(reserve_graph_visible(h, h->tok_out, (size_t)B * stride * sizeof(long long)));
    long long* tok = h->tok_out.as<long long>();
    MRMT3_CUDA_TRY(cudaMemsetAsync(tok, 0, (size_t)B * stride * sizeof(long long), s));
    int steps = 0;
    for (int c0 = 0; c0 < B; c0 += kMaxLanes) {
        const int n = std::min(kMaxLanes, B - c0);
        MRMT3_TRY(ensure_decode_capacity(h, n, kSegFrames, max_length));
        MRMT3_TRY(h->enc_bf16.reserve((size_t)n * kSegFrames * kDModel * sizeof(bf16)));
        MRMT3_TRY(encode_segments(h, mel_f32 ? mel_f32 + (size_t)c0 * kSegFrames * kMels : nullptr,
                                  mel_bf16 ? mel_bf16 + (size_t)c0 * kSegFrames * kMels : nullptr, n,
                                  h->enc_bf16.as<bf16>(), nullptr, s));
        MRMT3_TRY(project_cross_kv(h, h->enc_bf16.as<bf16>(), (long)n * kSegFrames, nullptr, n, nullptr, 0, s));
        LaneArrays a = lane_arrays(h);
        std::vector<int> rows(n);
        for (int i = 0; i < n; ++i) rows[i] = c0 + i;
        MRMT3_CUDA_TRY(cudaMemcpyAsync(a.out_row, rows.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));  // `rows` is pageable host memory
        StepPlan pl{};
        pl.n_lanes = n;
        pl.tk = kSegFrames;
        pl.st = make_state(h, tok, stride, max_length, 0, forced);
        pl.prefix = nullptr;
        pl.ext_logits = logits_out;
        RUN(h, launch_decode_init(pl.st, n, nullptr, n, h->cfg.start_id, kMaxGroups, kGroupScalars, s));
        MRMT3_TRY(run_decode(h, pl, 0, forced != nullptr || logits_out != nullptr, s));
        MRMT3_CUDA_TRY(cudaMemcpyAsync(h->h_pinned + 8, a.finish_step, n * sizeof(int), cudaMemcpyDeviceToHost, s));
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) steps = std::max(steps, h->h_pinned[8 + i]);
    }
    MRMT3_CUDA_TRY(cudaMemcpyAsync(out_ids, tok, (size_t)B * stride * sizeof(long long),
                                   cudaMemcpyDeviceToDevice, s));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    if (steps_host) *steps_host = steps;
    return OkStatus();
}

// MR-MT3 greedy transcription batched across tracks (see include/mrmt3_b200.h).
Status generate_segmem(mrmt3_handle* h, const float* mel_f32, const bf16* mel_bf16, const int* seg_counts,
                       int n_tracks, int max_length, long long* out_ids, float* logits_out,
                       cudaStream_t s, const long long* forced) {
    if (!h->committed) return Error(5, "weights not committed");
    if (h->cfg.mem_variant == MRMT3_MEM_NONE) return Error(2, "model has no memory variant");
    if (n_tracks <= 0 || max_length <= 0) return Error(2, "n_tracks and max_length must be positive");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const bool v1 = h->cfg.mem_variant == MRMT3_MEM_V1_PREPEND;
    const int n_mem = std::min(h->cfg.mem_len, max_length);
    if (v1 && max_length < h->cfg.mem_len)
        return Error(2, "V1 (generate_2) needs max_length >= segmem_length (reference models/t5_segmem.py:213)");
    std::vector<long long> seg_base(n_tracks);
    long long total = 0;
    int max_segs = 0;
    for (int t = 0; t < n_tracks; ++t) {
        if (seg_counts[t] < 0) return Error(2, "negative segment count");
        seg_base[t] = total;
        total += seg_counts[t];
        max_segs = std::max(max_segs, seg_counts[t]);
    }
    if (total == 0) return OkStatus();
    const int S = (int)total;
    const int stride = max_length + 1;

    // 1. encoder over every segment of every track (batched; no sequential dependency)
    MRMT3_TRY(h->enc_bf16.reserve((size_t)S * kSegFrames * kDModel * sizeof(bf16)));
    MRMT3_TRY(encode_segments(h, mel_f32, mel_bf16, S, h->enc_bf16.as<bf16>(), nullptr, s));

    // 2. token rows (S, max_length+1) + the dummy memory ids of the first segment
    MRMT3_TRY(reserve_graph_visible(h, h->tok_out, (size_t)S * stride * sizeof(long long)));
    long long* tok = h->tok_out.as<long long>();
    MRMT3_CUDA_TRY(cudaMemsetAsync(tok, 0, (size_t)S * stride * sizeof(long long), s));
    MRMT3_TRY(h->dummy_ids.reserve((size_t)max_length * sizeof(long long)));
    {
        std::vector<long long> dummy(max_length, 0);
        if (v1) {
            dummy[0] = 1;                                  // reference models/t5_segmem.py:194
        } else {
            dummy[0] = 1134;                               // tie token, t5_segmem_v2_with_prev.py:257
            if (max_length > 1) dummy[1] = 1;
        }
        MRMT3_CUDA_TRY(cudaMemcpyAsync(h->dummy_ids.p, dummy.data(), dummy.size() * sizeof(long long),
                                       cudaMemcpyHostToDevice, s));
        MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    }

    // 3. waves of at most kMaxLanes tracks; inside a wave every track advances one segment per round
    for (int w0 = 0; w0 < n_tracks; w0 += kMaxLanes) {
        const int n = std::min(kMaxLanes, n_tracks - w0);
        int wave_segs = 0;
        for (int i = 0; i < n; ++i) wave_segs = std::max(wave_segs, seg_counts[w0 + i]);
        const int tk = v1 ? kSegFrames : kSegFrames + n_mem;
        const int prefix = v1 ? n_mem : 0;
        MRMT3_TRY(ensure_decode_capacity(h, n, tk, max_length + prefix));
        MRMT3_TRY(h->mem_bf16.reserve((size_t)n * n_mem * kDModel * sizeof(bf16)));
        MRMT3_TRY(reserve_graph_visible(h, h->mem_f32, (size_t)n * n_mem * kDModel * sizeof(float)));
        LaneArrays a = lane_arrays(h);
        // lanes in order of decreasing track length: the tracks still running in round r are then
        // the lane PREFIX [0, n_r), and the round launches n_r lanes instead of masking the rest --
        // in the ragged tail the step shrinks to fewer lane groups instead of idling whole groups
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = w0 + i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return seg_counts[x] > seg_counts[y]; });
        const int n_wave = n;
        std::vector<int> tab(4 * (size_t)n_wave);
        for (int r = 0; r < wave_segs; ++r) {
            int n_active = 0;
            while (n_active < n_wave && seg_counts[order[n_active]] > r) ++n_active;
            const int n = n_active;                                   // shadows the wave size: this round's lanes
            for (int i = 0; i < n; ++i) {
                const int seg = (int)seg_base[order[i]] + r;
                tab[i] = seg;                                        // out_row
                tab[n + i] = r > 0 ? seg - 1 : seg;                  // prev_row (r == 0: unused)
                tab[2 * n + i] = 1;                                  // init_active
                tab[3 * n + i] = seg;                                // seg_index
            }
            MRMT3_CUDA_TRY(cudaMemcpyAsync(a.out_row, tab.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
            MRMT3_CUDA_TRY(cudaMemcpyAsync(a.prev_row, tab.data() + n, n * sizeof(int), cudaMemcpyHostToDevice, s));
            MRMT3_CUDA_TRY(cudaMemcpyAsync(a.init_active, tab.data() + 2 * n, n * sizeof(int), cudaMemcpyHostToDevice, s));
            MRMT3_CUDA_TRY(cudaMemcpyAsync(a.seg_index, tab.data() + 3 * n, n * sizeof(int), cudaMemcpyHostToDevice, s));
            MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
            // memory block from the previous output row (or the dummy ids)
            if (r == 0)
                MRMT3_TRY(memory_block(h, h->dummy_ids.as<long long>(), 0, nullptr, n, max_length, n_mem,
                                       h->mem_bf16.as<bf16>(), h->mem_f32.as<float>(), s));
            else
                MRMT3_TRY(memory_block(h, tok, stride, a.prev_row, n, max_length, n_mem,
                                       h->mem_bf16.as<bf16>(), h->mem_f32.as<float>(), s));
            MRMT3_TRY(project_cross_kv(h, h->enc_bf16.as<bf16>(), (long)S * kSegFrames, a.seg_index, n,
                                       v1 ? nullptr : h->mem_bf16.as<bf16>(), v1 ? 0 : n_mem, s));
            StepPlan pl{};
            pl.n_lanes = n;
            pl.tk = tk;
            pl.st = make_state(h, tok, stride, max_length, prefix, nullptr);
            if (forced) {
                // parity hook: lane i is fed row out_row[i] of `forced` (S_total, max_length + 1), so
                // every segment's memory block is built from the caller's tokens, not the arg-max
                pl.st.forced = forced;
                pl.st.forced_by_row = 1;
            }
            pl.prefix = v1 ? h->mem_f32.as<float>() : nullptr;
            pl.ext_logits = logits_out;
            RUN(h, launch_decode_init(pl.st, n, a.init_active, n_active, h->cfg.start_id, kMaxGroups, kGroupScalars, s));
            MRMT3_TRY(run_decode(h, pl, prefix, logits_out != nullptr || forced != nullptr, s));
        }
    }
    // 4. (S, max_length+1) -> (S, max_length): F.pad incl. the negative pad that drops the last
    //    token of a row that never emitted EOS (reference t5_segmem_v2_with_prev.py:287-291)
    MRMT3_CUDA_TRY(cudaMemcpy2DAsync(out_ids, (size_t)max_length * sizeof(long long), tok,
                                     (size_t)stride * sizeof(long long), (size_t)max_length * sizeof(long long),
                                     S, cudaMemcpyDeviceToDevice, s));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    return OkStatus();
}

Status api_memory_block(mrmt3_handle* h, const long long* prev_ids, int B, int Lp, float* mem_out,
                        cudaStream_t s) {
    if (!h->committed) return Error(5, "weights not committed");
    if (h->cfg.mem_variant == MRMT3_MEM_NONE) return Error(2, "model has no memory variant");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const int n_mem = std::min(h->cfg.mem_len, Lp);
    return memory_block(h, prev_ids, Lp, nullptr, B, Lp, n_mem, nullptr, mem_out, s);
}

// teacher-forced forward (reference models/t5.py:99-249, t5_segmem_v2_with_prev.py:60-224)
Status api_forward_logits(mrmt3_handle* h, const float* mel, int B, const long long* dec_ids, int L,
                          const long long* targets_prev, int Lp, float* logits_out, cudaStream_t s) {
    if (!h->committed) return Error(5, "weights not committed");
    // V1 (T5SegMem, models/t5_segmem.py:68-170): `targets_prev` holds the segmem ids of every row (the caller
    // builds them from the previous row's decoder input, as the reference does) and the memory rows are
    // PREPENDED to the decoder input instead of joining the cross-attention keys
    const bool v1 = h->cfg.mem_variant == MRMT3_MEM_V1_PREPEND;
    if (B <= 0 || L <= 0 || L > kNPos) return Error(2, "bad B or L");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const bool mem = h->cfg.mem_variant == MRMT3_MEM_V2_APPEND;
    if ((mem || v1) && (!targets_prev || Lp <= 0)) return Error(2, "targets_prev required for the memory variants");
    const int n_mem = (mem || v1) ? std::min(h->cfg.mem_len, Lp) : 0;
    const int n_pre = v1 ? n_mem : 0;             // memory rows in front of every decoder sequence
    const int Lt = L + n_pre;                     // decoder rows per sequence
    if (Lt > kNPos) return Error(2, "L + memory length exceeds the positional table");
    const int tk = kSegFrames + (mem ? n_mem : 0);
    const float eps = h->cfg.ln_eps;
    for (int c0 = 0; c0 < B; c0 += kMaxLanes) {
        const int n = std::min(kMaxLanes, B - c0);
        MRMT3_TRY(ensure_decode_capacity(h, n, tk, 1));
        MRMT3_TRY(h->enc_bf16.reserve((size_t)n * kSegFrames * kDModel * sizeof(bf16)));
        MRMT3_TRY(encode_segments(h, mel + (size_t)c0 * kSegFrames * kMels, nullptr, n, h->enc_bf16.as<bf16>(), nullptr, s));
        if (mem || v1) {
            MRMT3_TRY(h->mem_bf16.reserve((size_t)n * n_mem * kDModel * sizeof(bf16)));
            if (v1) MRMT3_TRY(reserve_graph_visible(h, h->mem_f32, (size_t)n * n_mem * kDModel * sizeof(float)));
            MRMT3_TRY(memory_block(h, targets_prev + (size_t)c0 * Lp, Lp, nullptr, n, Lp, n_mem,
                                   h->mem_bf16.as<bf16>(), v1 ? h->mem_f32.as<float>() : nullptr, s));
        }
        MRMT3_TRY(project_cross_kv(h, h->enc_bf16.as<bf16>(), (long)n * kSegFrames, nullptr, n, mem ? h->mem_bf16.as<bf16>() : nullptr,
                                   mem ? n_mem : 0, s));
        // decoder over n*Lt rows, in row chunks of whole sequences
        const int seq_chunk = std::max(1, (kEncChunk * kSegFrames) / Lt);
        if (v1) MRMT3_TRY(h->v1_logits.reserve((size_t)std::min(seq_chunk, n) * Lt * kVocab * sizeof(float)));
        for (int b0 = 0; b0 < n; b0 += seq_chunk) {
            const int nb = std::min(seq_chunk, n - b0);
            const int M = nb * Lt;
            const int L_tok = L;   // token rows per sequence (the caller's L)
            const int L = Lt;      // from here on a "sequence" is memory rows + token rows
            MRMT3_TRY(h->rows.reserve(M));
            RowWorkspace& w = h->rows;
            float* H = w.h32.as<float>();
            const ARowMap id{nullptr, 1};
            if (v1)
                RUN(h, launch_embed_tokens_prefixed(dec_ids + (size_t)(c0 + b0) * L_tok, h->emb, h->pe,
                                                    h->mem_f32.as<float>() + (size_t)b0 * n_mem * kDModel, H, nb, L_tok, n_pre, s));
            else
                RUN(h, launch_embed_tokens(dec_ids + (size_t)(c0 + b0) * L, h->emb, h->pe, H, nb, L, 0, s));
            for (int li = 0; li < h->cfg.n_dec_layers; ++li) {
                const LayerW& Lw = h->dec.layers[li];
                RUN(h, launch_rmsnorm(H, Lw.ln_self, eps, w.n_bf16.as<bf16>(), nullptr, M, nullptr, 1, s));
                RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, id, Lw.wqkv, kDModel, M, 3 * kInner, kDModel,
                                       EpiStoreBf16{w.qkv.as<bf16>(), 3 * kInner}, s));
                AttnFullParams ap{};
                ap.Q = w.qkv.as<bf16>();
                ap.K = ap.Q + kInner;
                ap.V = ap.Q + 2 * kInner;
                ap.q_batch_stride = ap.k_batch_stride = ap.v_batch_stride = (long)L * 3 * kInner;
                ap.q_head_stride = ap.k_head_stride = ap.v_head_stride = kDKV;
                ap.q_row_stride = ap.k_row_stride = ap.v_row_stride = 3 * kInner;
                ap.O = w.ctx.as<bf16>();
                ap.o_batch_stride = (long)L * kInner;
                ap.o_head_stride = kDKV;
                ap.o_row_stride = kInner;
                ap.Tq = L;
                ap.Tk = L;
                ap.causal = 1;
                ap.causal_offset = 0;
                RUN(h, launch_attn_full_auto(*h->tma, ap, nb, s));
                RUN(h, launch_gemm_tc(*h->tma, w.ctx.as<bf16>(), kInner, M, id, Lw.wo, kInner, M, kDModel, kInner, EpiResidual{H, kDModel}, s));

                RUN(h, launch_rmsnorm(H, Lw.ln_cross, eps, w.n_bf16.as<bf16>(), nullptr, M, nullptr, 1, s));
                RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, id, Lw.cq, kDModel, M, kInner, kDModel,
                                       EpiStoreBf16{w.qc.as<bf16>(), kInner}, s));
                AttnFullParams cp{};
                const size_t lane_sz = (size_t)h->cfg.n_dec_layers * 2 * kHeads * h->tk_cap * kDKV;
                const bf16* kbase = h->cross_cache.as<bf16>() + (size_t)b0 * lane_sz +
                                    (size_t)li * 2 * kHeads * h->tk_cap * kDKV;
                cp.Q = w.qc.as<bf16>();
                cp.q_batch_stride = (long)L * kInner;
                cp.q_head_stride = kDKV;
                cp.q_row_stride = kInner;
                cp.K = kbase;
                cp.V = kbase + (size_t)kHeads * h->tk_cap * kDKV;
                cp.k_batch_stride = cp.v_batch_stride = (long)lane_sz;
                cp.k_head_stride = cp.v_head_stride = (long)h->tk_cap * kDKV;
                cp.k_row_stride = cp.v_row_stride = kDKV;
                cp.O = w.ctx.as<bf16>();
                cp.o_batch_stride = (long)L * kInner;
                cp.o_head_stride = kDKV;
                cp.o_row_stride = kInner;
                cp.Tq = L;
                cp.Tk = tk;
                cp.causal = 0;
                RUN(h, launch_attn_full_auto(*h->tma, cp, nb, s));
                RUN(h, launch_gemm_tc(*h->tma, w.ctx.as<bf16>(), kInner, M, id, Lw.co, kInner, M, kDModel, kInner, EpiResidual{H, kDModel}, s));

                RUN(h, launch_rmsnorm(H, Lw.ln_ff, eps, w.n_bf16.as<bf16>(), nullptr, M, nullptr, 1, s));
                RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, id, Lw.wi, kDModel, M, 2 * kDFF, kDModel,
                                       EpiGatedGelu{w.ff.as<bf16>(), kDFF}, s));
                RUN(h, launch_gemm_tc(*h->tma, w.ff.as<bf16>(), kDFF, M, id, Lw.wff, kDFF, M, kDModel, kDFF, EpiResidual{H, kDModel}, s));
            }
            RUN(h, launch_rmsnorm(H, h->dec.final_ln, eps, w.n_bf16.as<bf16>(), nullptr, M, nullptr, 1, s));
            if (v1) {
                // logits of all rows into a scratch, then the token rows of every sequence to the caller
                // (reference: sequence_output[:, segmem_length:], models/t5_segmem.py:158-159)
                float* scratch = h->v1_logits.as<float>();
                RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, id, h->lm_head, kDModel, M, kVocab, kDModel,
                                       EpiStoreF32{scratch, kVocab}, s));
                MRMT3_CUDA_TRY(cudaMemcpy2DAsync(logits_out + (size_t)(c0 + b0) * L_tok * kVocab, (size_t)L_tok * kVocab * sizeof(float),
                                                 scratch + (size_t)n_pre * kVocab, (size_t)L * kVocab * sizeof(float),
                                                 (size_t)L_tok * kVocab * sizeof(float), nb, cudaMemcpyDeviceToDevice, s));
            } else {
                RUN(h, launch_gemm_tc(*h->tma, w.n_bf16.as<bf16>(), kDModel, M, id, h->lm_head, kDModel, M, kVocab, kDModel,
                                       EpiStoreF32{logits_out + (size_t)(c0 + b0) * L * kVocab, kVocab}, s));
            }
        }
    }
    return OkStatus();
}

Status api_logmel(mrmt3_handle* h, const float* audio, const long long* seg_start, const int* seg_len,
                  const int* valid_frames, int n_seg, int flags, float* out_f32, bf16* out_bf16,
                  cudaStream_t s) {
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    if (!out_f32 && !out_bf16) return Error(2, "logmel needs at least one output buffer");
    RUN(h, h->frontend.run(audio, seg_start, seg_len, valid_frames, n_seg, (flags & MRMT3_MEL_NORM) ? 1 : 0,
                           out_f32, out_bf16, s));
    return OkStatus();
}

Status api_transcribe_host(mrmt3_handle* h, const float* audio_host, long long n_samples,
                           const long long* seg_start_host, const int* seg_len_host,
                           const int* valid_frames_host, int n_seg, const int* seg_counts_host, int n_tracks,
                           int flags, int max_length, long long* out_ids_host, int* steps_host,
                           cudaStream_t s) {
    if (!h->committed) return Error(5, "weights not committed");
    if (n_seg <= 0) return Error(2, "n_seg must be positive");
    MRMT3_CUDA_TRY(cudaSetDevice(h->device));
    const bool mem = h->cfg.mem_variant != MRMT3_MEM_NONE;
    if (mem && (!seg_counts_host || n_tracks <= 0)) return Error(2, "memory variants need seg_counts_host/n_tracks");
    MRMT3_TRY(h->audio.reserve((size_t)n_samples * sizeof(float)));
    const size_t tab_bytes = (size_t)n_seg * (sizeof(long long) + 2 * sizeof(int));
    MRMT3_TRY(h->seg_tab.reserve(tab_bytes));
    long long* d_start = h->seg_tab.as<long long>();
    int* d_len = reinterpret_cast<int*>(d_start + n_seg);
    int* d_valid = d_len + n_seg;
    MRMT3_CUDA_TRY(cudaMemcpyAsync(h->audio.p, audio_host, (size_t)n_samples * sizeof(float), cudaMemcpyHostToDevice, s));
    MRMT3_CUDA_TRY(cudaMemcpyAsync(d_start, seg_start_host, n_seg * sizeof(long long), cudaMemcpyHostToDevice, s));
    MRMT3_CUDA_TRY(cudaMemcpyAsync(d_len, seg_len_host, n_seg * sizeof(int), cudaMemcpyHostToDevice, s));
    if (valid_frames_host)
        MRMT3_CUDA_TRY(cudaMemcpyAsync(d_valid, valid_frames_host, n_seg * sizeof(int), cudaMemcpyHostToDevice, s));
    MRMT3_TRY(h->mel_bf16.reserve((size_t)n_seg * kSegFrames * kMels * sizeof(bf16)));
    MRMT3_TRY(api_logmel(h, h->audio.as<float>(), d_start, d_len, valid_frames_host ? d_valid : nullptr, n_seg,
                         flags, nullptr, h->mel_bf16.as<bf16>(), s));
    const int width = mem ? max_length : max_length + 1;
    DeviceBuffer& ids = h->ids_dev;
    MRMT3_TRY(ids.reserve((size_t)n_seg * width * sizeof(long long)));
    if (mem)
        MRMT3_TRY(generate_segmem(h, nullptr, h->mel_bf16.as<bf16>(), seg_counts_host, n_tracks, max_length,
                                  ids.as<long long>(), nullptr, s, nullptr));
    else
        MRMT3_TRY(generate_base(h, nullptr, h->mel_bf16.as<bf16>(), n_seg, max_length, ids.as<long long>(),
                                steps_host, nullptr, nullptr, s));
    MRMT3_CUDA_TRY(cudaMemcpyAsync(out_ids_host, ids.p, (size_t)n_seg * width * sizeof(long long),
                                   cudaMemcpyDeviceToHost, s));
    MRMT3_CUDA_TRY(cudaStreamSynchronize(s));
    return OkStatus();
}

}  // namespace mrmt3
