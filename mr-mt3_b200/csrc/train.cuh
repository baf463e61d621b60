// Backward-pass and optimizer kernels of the fine-tune step (train_kernels.cu) and the training
// state kept in the handle (train.cu).
#pragma once
#include "attention.cuh"
#include "common.cuh"

namespace mrmt3 {

Status launch_xent(const float* logits, const long long* labels, int rows, int V, float* scal /* [1/count, mean loss] */,
                   float* row_loss, bf16* dlogits, cudaStream_t s);
// the norm-weight gradient is left as rmsnorm_bwd_parts(rows) partial rows in dg_part; after the
// last norm of the step, launch_norm_dg_reduce adds them up in a fixed order into the gradient
constexpr int kNormBwdMaxParts = 592;
constexpr int kNormBwdMaxNorms = 64;
struct NormDgList {
    float* dst[kNormBwdMaxNorms];
    int n_parts[kNormBwdMaxNorms];
    int n;
};
int rmsnorm_bwd_parts(int rows);
Status launch_norm_dg_reduce(const NormDgList& list, const float* parts /* [n][kNormBwdMaxParts][512] */, cudaStream_t s);
Status launch_rmsnorm_bwd(const float* x, const float* g, float eps, const bf16* dy, int rows, float* dres,
                          bf16* dres_bf16, float* dg_part, cudaStream_t s);
Status launch_gated_gelu_fwd(const bf16* raw, bf16* ff, size_t rows, DropSpec drop, cudaStream_t s);
Status launch_gated_gelu_bwd(const bf16* raw, const bf16* dff, bf16* draw, size_t rows, DropSpec drop, cudaStream_t s);
// x[i] *= keep(i) / (1 - p)   (dropout forward on a value, or backward on its gradient)
Status launch_dropout_f32(float* x, size_t n, DropSpec drop, cudaStream_t s);
Status launch_dropout_bf16(bf16* x, size_t n, DropSpec drop, cudaStream_t s);
// out_bf16[i] = bf16(in[i] * keep(i) / (1 - p))
Status launch_dropout_cast(const float* in, bf16* out, size_t n, DropSpec drop, cudaStream_t s);
size_t embed_bwd_scratch_bytes(int rows);
Status launch_embed_bwd(const long long* ids, const float* dH, float* dEmb, int rows, void* scratch, cudaStream_t s);
Status launch_cast_f32_bf16(const float* in, bf16* out, size_t n, cudaStream_t s);
Status launch_sanitize_ids(const long long* in, long long* out, size_t n, int pad, cudaStream_t s);
Status launch_bf16_to_f32(const bf16* in, float* out, size_t n, cudaStream_t s);
Status launch_reduce_splits(const float* partials, float* out, size_t n, int splits, cudaStream_t s);
Status launch_adamw(float* p, const float* g, float* m, float* v, bf16* p_bf16, size_t n, float lr, float beta1,
                    float beta2, float eps, float wd, int step, cudaStream_t s);
// One launch over every parameter tensor: entry i covers 256-element blocks [first_block, next entry's) of
// the flat order; p is the tensor's own fp32 storage (arena parameter or master slice), off its offset in the
// flat gradient / moment buffers
struct AdamSlot {
    float* p;
    bf16* p_bf16;
    unsigned long long off, n;
    unsigned int first_block, pad;
};
Status launch_adamw_multi(const AdamSlot* table, int n_slots, unsigned int n_blocks, const float* g, float* m, float* v,
                          float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t s);

// attention backward; Q/K/V/O layouts as AttnFullParams, dQ in Q's layout, dO in O's layout,
// dK/dV in their own (dk_*) layout; lse2 / delta are (batch, heads, Tq) fp32
struct AttnBwdParams {
    const bf16 *Q, *K, *V, *O, *dO;
    bf16 *dQ, *dK, *dV;
    long q_batch_stride, q_head_stride;
    int q_row_stride;
    long k_batch_stride, k_head_stride;
    int k_row_stride;
    long v_batch_stride, v_head_stride;
    int v_row_stride;
    long o_batch_stride, o_head_stride;
    int o_row_stride;
    long dk_batch_stride, dk_head_stride;
    int dk_row_stride;
    const float* lse2;
    float* delta;
    int Tq, Tk, causal, causal_offset;
    DropSpec drop;  // the forward's attention-weight dropout (only on() and scale are used here)
    const unsigned short* keep;  // its keep bits as the forward saved them (AttnFullParams::keep); required when drop.on()
};

// Hout (fp32) = Hin + dropout(acc): the sublayer output of the training forward, mask index
// row*ld + col.  Out of place, so that the sublayer's input stays behind as the activation the
// backward needs (no snapshot copy of the residual stream); drop may be off.
#ifndef MRMT3_EPI_WIDE
#define MRMT3_EPI_WIDE 1
#endif
struct EpiResidualTo {
    const float* Hin;
    float* Hout;
    int ldh;
    DropSpec drop;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        const unsigned long long i = (unsigned long long)row * ldh + col;
        float2 h = *reinterpret_cast<const float2*>(Hin + (size_t)row * ldh + col);
        float f0 = 1.f, f1 = 1.f;
        if (drop.on()) drop_factor2(drop, i, f0, f1);
        h.x += v0 * f0;
        h.y += v1 * f1;
        *reinterpret_cast<float2*>(Hout + (size_t)row * ldh + col) = h;
    }
};
// 8 consecutive columns of one row (col % 8 == 0, ldh % 4 == 0): two aligned mask groups, 16-byte accesses
__device__ __forceinline__ void epi_store8(const EpiResidualTo& e, int row, int col, const float (&v)[8]) {
    float f0[4] = {1.f, 1.f, 1.f, 1.f}, f1[4] = {1.f, 1.f, 1.f, 1.f};
    if (e.drop.on()) {
        const unsigned long long g = ((unsigned long long)row * e.ldh + col) >> 2;
        drop_factor4(e.drop, g, f0);
        drop_factor4(e.drop, g + 1, f1);
    }
    const float* hin = e.Hin + (size_t)row * e.ldh + col;
    float* hout = e.Hout + (size_t)row * e.ldh + col;
    if (MRMT3_EPI_WIDE && ((reinterpret_cast<size_t>(e.Hin) | reinterpret_cast<size_t>(e.Hout) | ((size_t)e.ldh * 4)) & 31) == 0) {
        float a[8];   // one whole 32-byte sector per access (STG.E.256 / LDG.E.256), see gemm_tcgen05.cuh
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                     : "l"(hin));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] += v[j] * f0[j];
            a[4 + j] += v[4 + j] * f1[j];
        }
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(hout), "f"(a[0]), "f"(a[1]), "f"(a[2]),
                     "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
                     : "memory");
        return;
    }
    const float4* p = reinterpret_cast<const float4*>(hin);
    float4* q = reinterpret_cast<float4*>(hout);
    const float4 a = p[0], b = p[1];
    q[0] = make_float4(a.x + v[0] * f0[0], a.y + v[1] * f0[1], a.z + v[2] * f0[2], a.w + v[3] * f0[3]);
    q[1] = make_float4(b.x + v[4] * f1[0], b.y + v[5] * f1[1], b.z + v[6] * f1[2], b.w + v[7] * f1[3]);
}
// prefetching interface of the tcgen05 GEMM epilogue (gemm_tcgen05.cuh: EpiPrefetch)
template <class Epi> struct EpiPrefetch;
template <> struct EpiPrefetch<EpiResidualTo> { static constexpr bool value = true; };
__device__ __forceinline__ void epi_fetch32(const EpiResidualTo& e, int row, int col, float (&h)[32]) {
    const float* p = e.Hin + (size_t)row * e.ldh + col;
    if (MRMT3_EPI_WIDE && ((reinterpret_cast<size_t>(e.Hin) | ((size_t)e.ldh * 4)) & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=f"(h[8 * j]), "=f"(h[8 * j + 1]), "=f"(h[8 * j + 2]), "=f"(h[8 * j + 3]), "=f"(h[8 * j + 4]),
                           "=f"(h[8 * j + 5]), "=f"(h[8 * j + 6]), "=f"(h[8 * j + 7])
                         : "l"(p + 8 * j));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(&h[4 * j]) = *reinterpret_cast<const float4*>(p + 4 * j);
    }
}
__device__ __forceinline__ void epi_store32_fetched(const EpiResidualTo& e, int row, int col, const float (&v)[32], const float (&h)[32]) {
    float o[32];
    if (e.drop.on()) {
        const unsigned long long g = ((unsigned long long)row * e.ldh + col) >> 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float f[4];
            drop_factor4(e.drop, g + j, f);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[4 * j + k] = h[4 * j + k] + v[4 * j + k] * f[k];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = h[j] + v[j];
    }
    float* q = e.Hout + (size_t)row * e.ldh + col;
    if (MRMT3_EPI_WIDE && ((reinterpret_cast<size_t>(e.Hout) | ((size_t)e.ldh * 4)) & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(q + 8 * j), "f"(o[8 * j]), "f"(o[8 * j + 1]),
                         "f"(o[8 * j + 2]), "f"(o[8 * j + 3]), "f"(o[8 * j + 4]), "f"(o[8 * j + 5]), "f"(o[8 * j + 6]), "f"(o[8 * j + 7])
                         : "memory");
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(q + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
}
Status launch_attn_bwd(const AttnBwdParams& p, int batch, cudaStream_t s);

}  // namespace mrmt3
