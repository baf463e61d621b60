// Parameter blocks and launchers of the attention kernels (attention.cu).
#pragma once
#include "common.cuh"

namespace mrmt3 {

// whole-sequence attention; strides in elements: element (b, head, row, d) of X lives at
// X + b*x_batch_stride + head*x_head_stride + row*x_row_stride + d
struct AttnFullParams {
    const bf16* Q;
    long q_batch_stride;
    long q_head_stride;
    int q_row_stride;
    const bf16* K;
    long k_batch_stride;
    long k_head_stride;
    int k_row_stride;
    const bf16* V;
    long v_batch_stride;
    long v_head_stride;
    int v_row_stride;
    bf16* O;
    long o_batch_stride;
    long o_head_stride;
    int o_row_stride;
    int Tq, Tk;
    int causal;         // key j visible to query i iff j <= i + causal_offset
    int causal_offset;
    float* lse2;        // optional (batch, heads, Tq): m*log2(e) + log2(l), saved for the backward pass
    DropSpec drop;      // training only: dropout on the attention weights, index ((b*H + h)*Tq + q)*Tk + k
    unsigned short* keep;  // optional with `drop`: the keep bits, saved for the backward pass
                           // [(b*H + h)*Tq + q][ceil(Tk/64)][4] words; word (kt, j) bit 2*ni + e is key
                           // 64 kt + 8 ni + 2 j + e  (the mma C-fragment order of a 64-key tile)
};
Status launch_attn_full(const AttnFullParams& p, int batch, cudaStream_t stream);
// tcgen05 / TMEM version of the same contract (attention_tc.cu): S = Q K^T and O = P V as tcgen05.mma with
// the accumulators in TMEM, one softmax thread per query row.  Selected by launch_attn_full_auto.
class TmaCache;
Status launch_attn_full_tc(TmaCache& tc, const AttnFullParams& p, int batch, cudaStream_t stream);
// the tcgen05 kernel when enabled (MRMT3_ATTN_FULL_TC / option "attn_full_tc") and the problem has at
// least one full 128-row query tile, else the mma.sync kernel
Status launch_attn_full_auto(TmaCache& tc, const AttnFullParams& p, int batch, cudaStream_t stream);
void attn_full_configure(int use_tc);

// decode-step attention (one query per lane and head)
struct AttnDecodeParams {
    const bf16* q;        // (lanes, q_stride): self -> fused [q|k|v] row; cross -> q row
    int q_stride;
    bf16* out;            // (lanes, out_stride) context, heads concatenated
    int out_stride;
    const bf16* kv_pool;  // self: page pool; cross: cross cache
    int layer;
    int n_layers;
    // paged self-attention
    const int* step_ptr;     // device scalar: position of the token being decoded
    int pos_offset;
    const int* block_table;  // (lanes, max_pages) page ids
    int max_pages;
    size_t page_stride;      // elements per page (all layers, k and v, all heads)
    // cross-attention
    int tk_cap;
    int n_keys;              // keys per lane when n_keys_ptr == nullptr
    const int* n_keys_ptr;   // optional per-lane key count
    // early-exit mask: lanes with active[lane] == 0 are skipped
    const int* active;
    // TMA variant: host pointer to a CUtensorMap over the whole cache viewed as (rows, 64) bf16,
    // box 32 x 64, 128-byte swizzle; tmap_row0 = row of kv_pool inside that view (cross cache of
    // a lane group), 0 for the page pool
    const void* tmap;
    long long tmap_row0;
    // TMA variant, split-key work units: when part_keys > 0 (a multiple of 64) an item's keys are cut
    // at FIXED multiples of part_keys, every part is a work unit of its own (partial softmax state
    // to part_scratch), and the CTA that finishes an item's last outstanding part merges them in part
    // order.  The cut points depend on the key count only, never on the batch: a row's result does
    // not depend on its batch neighbours.  part_scratch: (lanes, heads, max_parts, 66) fp32 of this
    // lane group; part_counter: (lanes, heads) ints, zero between launches (the merger resets them).
    int part_keys;
    int max_parts;
    float* part_scratch;
    int* part_counter;
    TraceSlot trace;
};
Status launch_attn_decode(const AttnDecodeParams& p, int n_lanes, bool paged, cudaStream_t stream);
int attn_decode_max_keys();
// process-wide kernel selection: variant 0 = one CTA per (lane, head), 1 = persistent bulk-copy
// ring (default); negative / zero arguments leave a setting unchanged
void attn_decode_configure(int variant, int stages, int ctas_per_sm, int quartets);

}  // namespace mrmt3
