// The handle behind include/mrmt3_b200.h: packed weights, workspaces, KV-cache pages, CUDA
// graphs of the decode step, and the host-side drivers of the path (model.cu).
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/mrmt3_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "frontend.cuh"
#include "layers.cuh"

namespace mrmt3 {

class TmaCache;  // gemm_tcgen05.cuh
void gemm_set_2cta(int on);  // option gemm_2cta: CTA-pair tiles in the tcgen05 GEMM (-1 = library default)

struct DeviceBuffer {
    void* p = nullptr;
    size_t cap = 0;
    // grow-only; *moved set when the address changed
    Status reserve(size_t bytes, bool* moved = nullptr);
    void release();
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct LayerW {
    bf16* wqkv = nullptr;     // (3*inner, d)  rows [q; k; v]
    bf16* wo = nullptr;       // (d, inner)
    float* ln_self = nullptr; // (d)
    bf16* cq = nullptr;       // (inner, d)    decoder only
    bf16* co = nullptr;       // (d, inner)    decoder only
    float* ln_cross = nullptr;
    bf16* wi = nullptr;       // (2*d_ff, d)   rows interleaved wi_0[j], wi_1[j]
    bf16* wff = nullptr;      // (d, d_ff)
    float* ln_ff = nullptr;
    // decoder only: norm-folded copies for the fused-RMSNorm decode GEMMs (W * diag(ln)) and the
    // fp32 masters they are folded from at commit time
    bf16 *wqkv_f = nullptr, *cq_f = nullptr, *wi_f = nullptr;
    float *m_wqkv = nullptr, *m_cq = nullptr, *m_wi = nullptr;
};

struct StackW {
    std::vector<LayerW> layers;
    float* final_ln = nullptr;
};

// row-parallel scratch (encoder / memory block / teacher-forced decoder)
struct RowWorkspace {
    DeviceBuffer x_bf16, h32, n_bf16, qkv, ctx, ff, qc, hc32, ctx_c;
    int rows = 0;
    Status reserve(int rows_needed);
    void release();
};

struct StepGraphKey {
    int lane0, n_lanes, tk, max_tokens, prefix_len, kind;  // kind 0 = token step, 1 = prefix step
    // parity hooks baked into the step (forced tokens, per-step logits sink); null on the plain path
    const void *forced, *ext_logits;
    bool operator<(const StepGraphKey& o) const {
        if (forced != o.forced) return forced < o.forced;
        if (ext_logits != o.ext_logits) return ext_logits < o.ext_logits;
        if (lane0 != o.lane0) return lane0 < o.lane0;
        if (n_lanes != o.n_lanes) return n_lanes < o.n_lanes;
        if (tk != o.tk) return tk < o.tk;
        if (max_tokens != o.max_tokens) return max_tokens < o.max_tokens;
        if (prefix_len != o.prefix_len) return prefix_len < o.prefix_len;
        return kind < o.kind;
    }
};

struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int n_kernels = 0;
};

constexpr int kMaxLanes = 512;    // decode lanes per wave
// split-key decode attention defaults (0 = one work unit per (lane, head)); see attention.cuh
constexpr int kDefaultPartKeysSelf = 0;
constexpr int kDefaultPartKeysCross = 0;

}  // namespace mrmt3

#define RUN(h, call)          \
    do {                      \
        MRMT3_TRY(call);      \
        ++(h)->launches;      \
    } while (0)

struct mrmt3_handle {
    mrmt3_config cfg;
    int device = 0;
    mutable std::string err;
    int64_t launches = 0;
    bool committed = false;
    bool use_graphs = true;

    mrmt3::Frontend frontend;
    mrmt3::TmaCache* tma = nullptr;      // cuTensorMap cache of the tcgen05 GEMMs

    // ---- weights ----
    mrmt3::DeviceBuffer arena;       // every packed tensor lives here
    size_t arena_used = 0;
    mrmt3::bf16* proj = nullptr;         // (d, d)
    float* emb = nullptr;                // (V, d) fp32
    mrmt3::bf16* lm_head = nullptr;      // (V, d)
    mrmt3::bf16* lm_head_f = nullptr;    // lm_head * diag(decoder.final_layer_norm)
    float* m_lm_head = nullptr;
    mrmt3::bf16* segmem_proj = nullptr;  // (d, d)
    mrmt3::bf16* cross_kv_w = nullptr;   // (n_dec * 2 * inner, d): layer-major [k; v]
    mrmt3::StackW enc, dec, mem;
    float* pe = nullptr;                 // (n_pos, d) fp32 sinusoid table
    int n_pos = 0;
    float inv_freq[256];
    bool have_inv_freq = false;
    std::set<std::string> seen;
    mrmt3::DeviceBuffer stage;           // fp32 staging for set_weight
    bool stage_holds_tensor = false;     // the last set_weight left its fp32 tensor in `stage`

    // ---- workspaces ----
    mrmt3::RowWorkspace rows;
    mrmt3::DeviceBuffer enc_bf16;        // (n_seg, 256, d) encoder states of the current call
    mrmt3::DeviceBuffer mem_bf16, mem_f32;  // (lanes, n_mem, d)
    mrmt3::DeviceBuffer v1_logits;          // V1 teacher forcing: logits of memory + token rows before the slice
    mrmt3::DeviceBuffer mel_f32, mel_bf16, audio, seg_tab, ids_dev;  // e2e path
    mrmt3::DeviceBuffer tok_out;         // (rows, max_length+1) int64 token rows of the call
    mrmt3::DeviceBuffer dummy_ids;       // (max_length) int64 first-segment memory ids

    // decode lanes
    int lane_cap = 0, tk_cap = 0, page_cap = 0;  // capacities the buffers below were sized for
    mrmt3::DeviceBuffer d_h32, d_n_bf16, d_qkv, d_ctx, d_qc, d_ff, d_logits;
    mrmt3::DeviceBuffer d_state;         // ints: step, n_active, ticket, then per-lane arrays
    mrmt3::DeviceBuffer kv_pool, block_table, cross_cache;
    // split-key decode attention (attention.cuh): partial softmax states + per-item tickets
    mrmt3::DeviceBuffer attn_parts, attn_tickets;
    // fused greedy head (gemm_skinny.cuh EpiGreedy): per-lane arg-max candidates of the column tiles
    mrmt3::DeviceBuffer greedy_cand, greedy_tickets;
    bool fuse_greedy = true;             // lm_head + arg-max + next-step embedding in one kernel
    int attn_max_parts = 0;
    int attn_part_keys_self = mrmt3::kDefaultPartKeysSelf, attn_part_keys_cross = mrmt3::kDefaultPartKeysCross;
    mrmt3::DeviceBuffer lane_tab;        // per-lane int tables (seg index, prev row, active)
    int* h_pinned = nullptr;             // pinned host ints for polling / finish steps

    std::map<mrmt3::StepGraphKey, mrmt3::StepGraph> graphs;

    // per-kernel-class timing of the decode step (mrmt3_profile_*): eager launches bracketed by
    // CUDA events on the launching stream
    bool prof_on = false;
    struct ProfRec { int cat; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[MRMT3_PROF_NCAT] = {0};
    int64_t prof_n[MRMT3_PROF_NCAT] = {0};
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    // lane groups: independent greedy loops on their own streams so that one group's
    // latency-bound projections overlap another group's HBM-bound attention
    int group_lanes = -1;            // < 0: by batch size (run_decode)
    bool group_serial = false;
    // parity hooks (forced tokens / per-step logits) normally run eagerly as one lane group; with
    // this set they go through the production path (CUDA-graph replay, concurrent lane groups)
    bool hooks_fast_path = false;
    mrmt3::DeviceBuffer trace_buf;       // 2 x u64 per decode-step kernel slot (mrmt3_trace_*)
    bool trace_on = false;           // debugging: run the groups one after another on one stream
    cudaStream_t gstream[16] = {nullptr};
    cudaEvent_t gdone[16] = {nullptr};
    cudaEvent_t gfork = nullptr;
    void* train = nullptr;               // mrmt3::TrainState (train.cu), created by the first train call
};

namespace mrmt3 {
// sizes the decode-lane buffers (KV pages, cross cache, ...) for n_lanes lanes, tk cross keys
Status ensure_decode_capacity(mrmt3_handle* h, int n_lanes, int tk, int max_positions);
void train_destroy(mrmt3_handle* h);
}  // namespace mrmt3
