// Row-wise / elementwise kernels of the path (layers.cu).
#pragma once
#include "common.cuh"

namespace mrmt3 {

// fp32 -> bf16, n elements (n % 4 == 0)
Status launch_cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t s);

// fp32 (rows, cols) -> bf16 written at dst[(row * row_mul + row_off) * cols + c]
// (weight packing: fused QKV = row offsets, gated FFN = interleaved rows)
Status launch_pack_weight(const float* src, bf16* dst, int rows, int cols, int row_mul, int row_off,
                          cudaStream_t s);

// same row mapping, fp32 destination (masters of the norm-folded decoder weights)
Status launch_pack_weight_f32(const float* src, float* dst, int rows, int cols, int row_mul, int row_off,
                              cudaStream_t s);
// dst[n][k] = bf16(master[n][k] * g[k])
Status launch_fold_norm(const float* master, const float* g, bf16* dst, int rows, int cols, cudaStream_t s);

// T5LayerNorm (RMSNorm, fp32 statistics): y = w * x * rsqrt(mean(x^2) + eps), rows of 512.
// active (optional): per-lane flag, row r belongs to lane r / rows_per_lane.
Status launch_rmsnorm(const float* x, const float* w, float eps, bf16* out_bf16, float* out_f32,
                      int rows, const int* active, int rows_per_lane, cudaStream_t s);

// decoder input for teacher forcing: H[b*L + l] = Emb[ids[b*L + l]] + PE[pos0 + l]
Status launch_embed_tokens(const long long* ids, const float* emb, const float* pe, float* H, int B,
                           int L, int pos0, cudaStream_t s);
// rows of sequence b: n_mem memory rows (mem (B, n_mem, d) fp32) then L token rows, + PE[0 .. n_mem + L)
Status launch_embed_tokens_prefixed(const long long* ids, const float* emb, const float* pe, const float* mem, float* H,
                                    int B, int L, int n_mem, cudaStream_t s);

// memory-block input: out_bf16[r] = bf16(Emb[ids[r]])   (ids int64, rows = n)
Status launch_embed_bf16(const long long* ids, long ids_row_stride, int rows_per_lane, int n_lanes,
                         const int* lane_src_row, const float* emb, bf16* out, cudaStream_t s);

// dst[b*n + j][:] = src[b*src_block + j][:]   (fp32 rows of 512)
Status launch_gather_rows(const float* src, float* dst, int n_lanes, int n, int src_block,
                          cudaStream_t s);

// ---- decode-step kernels -------------------------------------------------------------------
struct DecodeState {
    int* step;            // device scalar: absolute decoder position of the token being fed
    int* tok;             // (lanes) token being fed at this step
    int* active;          // (lanes) 1 while the lane is still decoding
    int* finish_step;     // (lanes) number of emitted tokens when the lane finished
    int* n_active;        // device scalar
    int* ticket;          // device scalar used to elect the last CTA of the argmax kernel
    long long* out;       // (rows, out_stride) int64 token rows (col 0 = BOS)
    const int* out_row;   // (lanes) row of `out` each lane writes
    int out_stride;
    int prefix_len;       // V1 memory prefix length (0 otherwise): out col = step - prefix + 1
    const long long* forced;  // optional teacher forcing: (lanes, forced_stride) next-token ids
    int forced_stride;
    int forced_by_row;        // 0: row = lane index; 1: row = out_row[lane] (MR-MT3: lanes change rows per round)
    int eos_id, pad_id;
    int max_tokens;       // stop a lane after this many emitted tokens
    TraceSlot trace;      // timeline trace slot of the kernel this state is passed to
};

// H[lane] = Emb[tok[lane]] + PE[step]     (prefix == nullptr)
// H[lane] = prefix[lane][step] + PE[step] (V1 memory prefix, step < prefix_len)
Status launch_decode_embed(const DecodeState& st, const float* emb, const float* pe,
                           const float* prefix, int prefix_stride, float* H, bf16* Hb, int n_lanes,
                           cudaStream_t s);

// greedy head: argmax over V logits (lowest index wins ties, as torch.argmax), EOS bookkeeping
// (reference models/t5.py:286-291), token write-out, step advance.
// logits of lane l at logits[out_row(l)*lane_stride + (step-prefix)*step_stride] when
// step_stride != 0, else logits[l*lane_stride].
Status launch_argmax_advance(const DecodeState& st, const float* logits, size_t lane_stride,
                             size_t step_stride, int n_lanes, int vocab, cudaStream_t s);

// start-of-segment state: tok = start_id, active = init_active (or 1), step = 0
Status launch_decode_init(const DecodeState& st, int n_lanes, const int* init_active, int n_active,
                          int start_id, int n_groups, int group_stride, cudaStream_t s);

// prefix steps produce no token: just advance the position
Status launch_advance_only(const DecodeState& st, cudaStream_t s);

}  // namespace mrmt3
