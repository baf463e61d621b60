// Backward-pass and optimizer kernels of the fine-tune step (BASELINE configs[4]; reference
// tasks/mt3_net*.py training_step: logits -> cross-entropy(ignore_index=-100) -> autograd ->
// AdamW).  The reference gets all of this from PyTorch autograd; here every op of the forward path
// has a hand-written backward:
//
//   xent_kernel            softmax cross-entropy forward + d(logits), rows with label -100 ignored
//   rmsnorm_bwd_kernel     T5LayerNorm backward (dx accumulated into the residual gradient; dg as
//                          per-CTA partial rows, norm_dg_reduce_kernel adds them in a fixed order)
//   gated_gelu_*           gated-GELU forward from the saved raw ffn-in output, and its backward
//   dropout_*              the elementwise dropout sites (counter-based hash, common.cuh)
//   embed_bwd_*            the stack-input gradient summed per token id into the embedding table,
//                          two levels, no atomics on values
//   attn_bwd_*             flash-style attention backward (no 1/sqrt(d) scale, no bias): dQ per
//                          query tile (first a pass forming delta = sum P dP), dK/dV per key tile,
//                          P recomputed from the saved log-sum-exp, dropout from the saved keep bits
//   (dgrad / wgrad are the tcgen05 GEMM of gemm_tcgen05.cuh in its MN-major operand modes)
//   adamw_kernel           AdamW on fp32 masters, refreshed bf16 copy
#include "train.cuh"

#include <type_traits>

#include <algorithm>

namespace mrmt3 {

// ---------------------------------------------------------------------------------------------
// cross-entropy, all on the device so that the forward needs no host round trip:
//   scal[0] = 1 / #(labels != ignore_index), row losses + logits gradient, scal[1] = mean loss
__global__ void __launch_bounds__(1024)
    xent_count_kernel(const long long* __restrict__ labels, int rows, int V, float* __restrict__ scal) {
    __shared__ int part[32];
    int c = 0;
    for (int i = threadIdx.x; i < rows; i += 1024) c += labels[i] >= 0 && labels[i] < V;
    c = (int)warp_sum((float)c);  // <= 32 * ceil(rows / 1024): exact in fp32 for any batch the stash can hold
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int w = 0; w < 32; ++w) n += part[w];
        scal[0] = n ? 1.0f / (float)n : 0.f;
    }
}

__global__ void __launch_bounds__(1024)
    xent_mean_kernel(const float* __restrict__ row_loss, int rows, float* __restrict__ scal) {
    __shared__ double part[1024];
    double a = 0.0;
    for (int i = threadIdx.x; i < rows; i += 1024) a += (double)row_loss[i];
    part[threadIdx.x] = a;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {  // fixed tree: the same bits every run
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) scal[1] = (float)(part[0] * (double)scal[0]);
}

// one warp per row of V logits
__global__ void __launch_bounds__(256)
    xent_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int rows, int V,
                const float* __restrict__ scal, float* __restrict__ row_loss, bf16* __restrict__ dlogits) {
    const float inv_count = scal[0];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = logits + (size_t)row * V;
    bf16* dx = dlogits + (size_t)row * V;
    const long long label = labels[row];
    if (label < 0 || label >= V) {  // ignore_index (and anything that is not a class: never indexed)
        if (lane == 0) row_loss[row] = 0.f;
        for (int j = lane * 2; j < V; j += 64) *reinterpret_cast<uint32_t*>(dx + j) = 0u;
        return;
    }
    float mx = -INFINITY;
    for (int j = lane; j < V; j += 32) mx = fmaxf(mx, x[j]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < V; j += 32) se += __expf(x[j] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    if (lane == 0) row_loss[row] = lse - x[label];
    const float inv_se = 1.f / se;
    for (int j = lane * 2; j < V; j += 64) {
        float p0 = __expf(x[j] - mx) * inv_se - (j == label ? 1.f : 0.f);
        float p1 = __expf(x[j + 1] - mx) * inv_se - (j + 1 == label ? 1.f : 0.f);
        *reinterpret_cast<uint32_t*>(dx + j) = pack_bf16(p0 * inv_count, p1 * inv_count);
    }
}

Status launch_xent(const float* logits, const long long* labels, int rows, int V, float* scal,
                   float* row_loss, bf16* dlogits, cudaStream_t s) {
    if (rows <= 0) return OkStatus();
    xent_count_kernel<<<1, 1024, 0, s>>>(labels, rows, V, scal);
    MRMT3_CHECK_LAUNCH();
    xent_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(logits, labels, rows, V, scal, row_loss, dlogits);
    MRMT3_CHECK_LAUNCH();
    xent_mean_kernel<<<1, 1024, 0, s>>>(row_loss, rows, scal);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// RMSNorm backward.  y = g * x * r, r = rsqrt(mean(x^2) + eps):
//   dx = r * (g * dy) - x * r^3 / d * sum_j (g_j dy_j x_j);  dg_j = sum_rows dy_j x_j r
// dx is ADDED to dres (the gradient of the residual stream the norm branched from); the updated
// rows are also written as bf16 (dres_bf16, optional) for the GEMMs that consume them next.
// dg: every CTA combines its eight warps in a fixed order and stores ONE partial row,
// dg_part[blockIdx.x][512]; norm_dg_reduce_kernel adds the partial rows of all norms of the step in
// CTA order at the end of the backward pass -- no float atomics, the same bits every run.
__global__ void __launch_bounds__(256)
    rmsnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float eps,
                       const bf16* __restrict__ dy, int rows, float* __restrict__ dres, bf16* __restrict__ dres_bf16,
                       float* __restrict__ dg_part) {
    __shared__ float s_dg[8][kDModel];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc_dg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc_dg[i] = 0.f;
    const float4* gr = reinterpret_cast<const float4*>(g);
    float4 gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gv[i] = gr[lane + i * 32];
    for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * kDModel);
        const uint2* dyr = reinterpret_cast<const uint2*>(dy + (size_t)row * kDModel);
        float4 xv[4], dv[4];
        float ss = 0.f, dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xv[i] = xr[lane + i * 32];
            uint2 raw = dyr[lane + i * 32];
            float2 a = __bfloat1622float2(*reinterpret_cast<bf162*>(&raw.x));
            float2 b = __bfloat1622float2(*reinterpret_cast<bf162*>(&raw.y));
            dv[i] = make_float4(a.x, a.y, b.x, b.y);
            ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
            dot += gv[i].x * dv[i].x * xv[i].x + gv[i].y * dv[i].y * xv[i].y + gv[i].z * dv[i].z * xv[i].z +
                   gv[i].w * dv[i].w * xv[i].w;
        }
        ss = warp_sum(ss);
        dot = warp_sum(dot);
        const float r = rsqrtf(ss * (1.0f / kDModel) + eps);
        const float c = dot * r * r * r * (1.0f / kDModel);
        float4* dr = reinterpret_cast<float4*>(dres + (size_t)row * kDModel);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 d = dr[lane + i * 32];
            d.x += r * gv[i].x * dv[i].x - xv[i].x * c;
            d.y += r * gv[i].y * dv[i].y - xv[i].y * c;
            d.z += r * gv[i].z * dv[i].z - xv[i].z * c;
            d.w += r * gv[i].w * dv[i].w - xv[i].w * c;
            dr[lane + i * 32] = d;
            if (dres_bf16)  // the next backward GEMMs take the updated residual gradient as a bf16 operand
                *reinterpret_cast<uint2*>(dres_bf16 + (size_t)row * kDModel + (lane + i * 32) * 4) =
                    make_uint2(pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
            acc_dg[i * 4 + 0] += dv[i].x * xv[i].x * r;
            acc_dg[i * 4 + 1] += dv[i].y * xv[i].y * r;
            acc_dg[i * 4 + 2] += dv[i].z * xv[i].z * r;
            acc_dg[i * 4 + 3] += dv[i].w * xv[i].w * r;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&s_dg[warp][(lane + i * 32) * 4]) =
            make_float4(acc_dg[i * 4 + 0], acc_dg[i * 4 + 1], acc_dg[i * 4 + 2], acc_dg[i * 4 + 3]);
    __syncthreads();
    for (int i = threadIdx.x; i < kDModel; i += 256) {
        float a = s_dg[0][i];
#pragma unroll
        for (int w = 1; w < 8; ++w) a += s_dg[w][i];
        dg_part[(size_t)blockIdx.x * kDModel + i] = a;
    }
}

int rmsnorm_bwd_parts(int rows) { return std::min(ceil_div(rows, 8), kNormBwdMaxParts); }

Status launch_rmsnorm_bwd(const float* x, const float* g, float eps, const bf16* dy, int rows, float* dres,
                          bf16* dres_bf16, float* dg_part, cudaStream_t s) {
    if (rows <= 0) return OkStatus();
    rmsnorm_bwd_kernel<<<rmsnorm_bwd_parts(rows), 256, 0, s>>>(x, g, eps, dy, rows, dres, dres_bf16, dg_part);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// dst[i][col] = sum over the n_parts[i] partial rows of norm i, in order: 8 row groups per column
// run in parallel and are combined in a fixed order
__global__ void __launch_bounds__(1024)
    norm_dg_reduce_kernel(NormDgList list, const float* __restrict__ parts) {
    __shared__ float s[8][128];
    const int i = blockIdx.x, col = blockIdx.y * 128 + (threadIdx.x & 127), grp = threadIdx.x >> 7;
    const int n = list.n_parts[i];
    const float* p = parts + (size_t)i * kNormBwdMaxParts * kDModel + col;
    const int per = (n + 7) / 8, b0 = grp * per, b1 = min(n, b0 + per);
    float a = 0.f;
    for (int b = b0; b < b1; ++b) a += p[(size_t)b * kDModel];
    s[grp][threadIdx.x & 127] = a;
    __syncthreads();
    if (grp == 0) {
        float t = s[0][threadIdx.x];
#pragma unroll
        for (int g2 = 1; g2 < 8; ++g2) t += s[g2][threadIdx.x];
        list.dst[i][col] = t;
    }
}

Status launch_norm_dg_reduce(const NormDgList& list, const float* parts, cudaStream_t s) {
    if (list.n <= 0) return OkStatus();
    norm_dg_reduce_kernel<<<dim3(list.n, kDModel / 128), 1024, 0, s>>>(list, parts);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// gated GELU from the raw ffn-in output (columns interleaved: 2j -> wi_0, 2j+1 -> wi_1)
__global__ void gated_gelu_fwd_kernel(const uint4* __restrict__ raw, uint2* __restrict__ ff, size_t n4, DropSpec drop) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // four outputs per thread = one mask group
    if (g >= n4) return;
    const uint4 v = raw[g];
    float f[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop.on()) drop_factor4(drop, g, f);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 ab = __bfloat1622float2(*reinterpret_cast<const bf162*>(&w[k]));
        o[k] = gelu_new(ab.x) * ab.y * f[k];
    }
    ff[g] = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
}

__global__ void gated_gelu_bwd_kernel(const uint4* __restrict__ raw, const uint2* __restrict__ dff,
                                      uint4* __restrict__ draw, size_t n4, DropSpec drop) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    const uint4 v = raw[g];
    const uint2 dv = dff[g];
    float f[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop.on()) drop_factor4(drop, g, f);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float2 d01 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&dv.x));
    float2 d23 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&dv.y));
    const float dd[4] = {d01.x * f[0], d01.y * f[1], d23.x * f[2], d23.y * f[3]};
    uint32_t out[4];
    const float k = 0.7978845608028654f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 ab = __bfloat1622float2(*reinterpret_cast<const bf162*>(&w[q]));
        const float a = ab.x, b = ab.y, d = dd[q];
        const float t = tanhf(k * (a + 0.044715f * a * a * a));
        const float gelu = 0.5f * a * (1.0f + t);
        const float dgelu = 0.5f * (1.0f + t) + 0.5f * a * (1.0f - t * t) * k * (1.0f + 3.0f * 0.044715f * a * a);
        out[q] = pack_bf16(d * b * dgelu, d * gelu);
    }
    draw[g] = make_uint4(out[0], out[1], out[2], out[3]);
}

Status launch_gated_gelu_fwd(const bf16* raw, bf16* ff, size_t rows, DropSpec drop, cudaStream_t s) {
    size_t n = rows * kDFF;
    if (!n) return OkStatus();
    gated_gelu_fwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(raw),
                                                                        reinterpret_cast<uint2*>(ff), n / 4, drop);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}
Status launch_gated_gelu_bwd(const bf16* raw, const bf16* dff, bf16* draw, size_t rows, DropSpec drop, cudaStream_t s) {
    size_t n = rows * kDFF;
    if (!n) return OkStatus();
    gated_gelu_bwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(raw),
                                                                        reinterpret_cast<const uint2*>(dff),
                                                                        reinterpret_cast<uint4*>(draw), n / 4, drop);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// elementwise dropout (forward on a value, backward on its gradient: the same factor)
__global__ void dropout_f32_kernel(float4* __restrict__ x, size_t n4, DropSpec drop) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    float f[4];
    drop_factor4(drop, g, f);
    float4 v = x[g];
    x[g] = make_float4(v.x * f[0], v.y * f[1], v.z * f[2], v.w * f[3]);
}
__global__ void dropout_bf16_kernel(uint2* __restrict__ x, size_t n4, DropSpec drop) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    float f[4];
    drop_factor4(drop, g, f);
    uint2 v = x[g];
    float2 a = __bfloat1622float2(*reinterpret_cast<bf162*>(&v.x)), b = __bfloat1622float2(*reinterpret_cast<bf162*>(&v.y));
    x[g] = make_uint2(pack_bf16(a.x * f[0], a.y * f[1]), pack_bf16(b.x * f[2], b.y * f[3]));
}
__global__ void dropout_cast_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4, DropSpec drop) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    float f[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop.on()) drop_factor4(drop, g, f);
    float4 v = in[g];
    out[g] = make_uint2(pack_bf16(v.x * f[0], v.y * f[1]), pack_bf16(v.z * f[2], v.w * f[3]));
}
Status launch_dropout_f32(float* x, size_t n, DropSpec drop, cudaStream_t s) {
    if (!n || !drop.on()) return OkStatus();
    dropout_f32_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<float4*>(x), n / 4, drop);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}
Status launch_dropout_bf16(bf16* x, size_t n, DropSpec drop, cudaStream_t s) {
    if (!n || !drop.on()) return OkStatus();
    dropout_bf16_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint2*>(x), n / 4, drop);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}
Status launch_dropout_cast(const float* in, bf16* out, size_t n, DropSpec drop, cudaStream_t s) {
    if (!n) return OkStatus();
    dropout_cast_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(in),
                                                                      reinterpret_cast<uint2*>(out), n / 4, drop);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// dEmb[ids[r]] += dH[r]   (rows of 512 fp32).  Token ids repeat heavily (pad = 0 fills half of a
// batch), so a plain scatter serialises 16 K x 512 atomics on one row -- and any atomic float
// accumulation makes the step irreproducible.  Two deterministic levels instead:
//   1. a CTA owns kEmbIds consecutive ids and one slice of kEmbSliceRows rows: it scans the slice's
//      ids from shared memory, each warp sums the matching rows it is responsible for in registers
//      (four independent 16-byte loads per lane per row), the four warps are combined in a fixed
//      order, and the partial row goes to a compact scratch list; only its SLOT is drawn from an
//      atomic counter, and slot_tab[slice][id] remembers it (-1: the slice has no such id)
//   2. one CTA per id adds that id's partial rows in slice order and updates dEmb (plain RMW: calls
//      are ordered by the stream)
// The values are therefore a fixed function of the inputs; only their scratch location varies.
constexpr int kEmbIds = 4;
constexpr int kEmbSliceRows = 256;
__global__ void __launch_bounds__(128)
    embed_bwd_partial_kernel(const long long* __restrict__ ids, const float* __restrict__ dH, float4* __restrict__ part,
                             int* __restrict__ slot_tab, int* __restrict__ counter, int rows, int vocab) {
    __shared__ int s_ids[kEmbSliceRows];
    __shared__ float4 s_acc[4][kEmbIds][128];
    __shared__ int s_has[kEmbIds], s_slot[kEmbIds];
    const int id0 = blockIdx.x * kEmbIds;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * kEmbSliceRows, n = min(kEmbSliceRows, rows - r0);
    if (threadIdx.x < kEmbIds) s_has[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < n; i += 128) s_ids[i] = (int)ids[r0 + i];
    __syncthreads();
    float4 acc[kEmbIds][4];
#pragma unroll
    for (int k = 0; k < kEmbIds; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
    for (int i = warp; i < n; i += 4) {
        const int k = s_ids[i] - id0;
        if (k >= 0 && k < kEmbIds) {  // uniform across the warp
            any = true;
            const float4* src = reinterpret_cast<const float4*>(dH + (size_t)(r0 + i) * kDModel) + lane;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = src[j * 32];
            if (lane == 0) s_has[k] = 1;  // same value from every writer
#pragma unroll
            for (int q = 0; q < kEmbIds; ++q)
                if (q == k) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[q][j].x += v[j].x; acc[q][j].y += v[j].y; acc[q][j].z += v[j].z; acc[q][j].w += v[j].w;
                    }
                }
        }
    }
    const bool cta_any = __syncthreads_or(any);
    if (!cta_any) {  // the common case: none of these ids occurs in the slice
        if (threadIdx.x < kEmbIds && id0 + threadIdx.x < vocab)
            slot_tab[(size_t)blockIdx.y * vocab + id0 + threadIdx.x] = -1;
        return;
    }
#pragma unroll
    for (int k = 0; k < kEmbIds; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_acc[warp][k][j * 32 + lane] = acc[k][j];
    if (threadIdx.x < kEmbIds) {
        const int k = threadIdx.x;
        const int slot = s_has[k] ? atomicAdd(counter, 1) : -1;
        s_slot[k] = slot;
        if (id0 + k < vocab) slot_tab[(size_t)blockIdx.y * vocab + id0 + k] = slot;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEmbIds; ++k) {
        const int slot = s_slot[k];
        if (slot < 0) continue;
        float4 a = s_acc[0][k][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
            const float4 b = s_acc[w][k][threadIdx.x];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        part[(size_t)slot * 128 + threadIdx.x] = a;
    }
}

__global__ void __launch_bounds__(128)
    embed_bwd_combine_kernel(const float4* __restrict__ part, const int* __restrict__ slot_tab, float4* __restrict__ dEmb,
                             int n_slices, int vocab) {
    const int id = blockIdx.x;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
    for (int sl = 0; sl < n_slices; ++sl) {
        const int slot = slot_tab[(size_t)sl * vocab + id];
        if (slot < 0) continue;
        any = true;
        const float4 b = part[(size_t)slot * 128 + threadIdx.x];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (!any) return;
    float4* d = dEmb + (size_t)id * 128 + threadIdx.x;
    float4 c = *d;
    c.x += a.x; c.y += a.y; c.z += a.z; c.w += a.w;
    *d = c;
}

size_t embed_bwd_scratch_bytes(int rows) {
    const size_t slices = ceil_div(rows, kEmbSliceRows);
    // every partial row belongs to a distinct (slice, id) pair that owns at least one input row
    return (size_t)rows * kDModel * 4 + slices * kVocab * 4 + 256;
}

Status launch_embed_bwd(const long long* ids, const float* dH, float* dEmb, int rows, void* scratch, cudaStream_t s) {
    if (rows <= 0) return OkStatus();
    static_assert(kDModel == 512, "one float4 column per thread of a 128-thread CTA");
    const int slices = ceil_div(rows, kEmbSliceRows);
    float4* part = reinterpret_cast<float4*>(scratch);
    int* slot_tab = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + (size_t)rows * kDModel * 4);
    int* counter = slot_tab + (size_t)slices * kVocab;
    MRMT3_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), s));
    embed_bwd_partial_kernel<<<dim3(ceil_div(kVocab, kEmbIds), slices), 128, 0, s>>>(ids, dH, part, slot_tab, counter, rows, kVocab);
    MRMT3_CHECK_LAUNCH();
    embed_bwd_combine_kernel<<<kVocab, 128, 0, s>>>(part, slot_tab, reinterpret_cast<float4*>(dEmb), slices, kVocab);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}
Status launch_bf16_to_f32(const bf16* in, float* out, size_t n, cudaStream_t s) {
    if (!n) return OkStatus();
    bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// out[i] = sum_s partials[s * n + i]   (split-K wgrad)
__global__ void reduce_splits_kernel(const float4* __restrict__ partials, float4* __restrict__ out, size_t n4, int splits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = partials[i];
    for (int s = 1; s < splits; ++s) {
        float4 b = partials[(size_t)s * n4 + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
}
Status launch_reduce_splits(const float* partials, float* out, size_t n, int splits, cudaStream_t s) {
    if (!n) return OkStatus();
    reduce_splits_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(partials),
                                                                      reinterpret_cast<float4*>(out), n / 4, splits);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// fp32 (n) -> bf16
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) *reinterpret_cast<uint32_t*>(out + i) = pack_bf16(in[i], in[i + 1]);
    else if (i < n) out[i] = __float2bfloat16(in[i]);
}
// token ids as the embedding kernels read them: anything outside [0, vocab) -- the -100 padding of
// targets_prev in particular (reference models/t5_segmem_v2_with_prev.py:119 masks it to pad) --
// becomes the pad id, so that the forward gather and the backward scatter agree on the row
__global__ void sanitize_ids_kernel(const long long* __restrict__ in, long long* __restrict__ out, size_t n, int pad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = in[i];
    out[i] = (v < 0 || v >= kVocab) ? (long long)pad : v;
}
Status launch_sanitize_ids(const long long* in, long long* out, size_t n, int pad, cudaStream_t s) {
    if (n == 0) return OkStatus();
    sanitize_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n, pad);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

Status launch_cast_f32_bf16(const float* in, bf16* out, size_t n, cudaStream_t s) {
    if (!n) return OkStatus();
    cast_f32_bf16_kernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, s>>>(in, out, n);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, bf16* __restrict__ p_bf16, size_t n, float lr, float beta1,
                             float beta2, float eps, float wd, float bc1, float bc2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pi = p[i], gi = g[i];
    pi *= 1.f - lr * wd;
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
    if (p_bf16) p_bf16[i] = __float2bfloat16(pi);
}

Status launch_adamw(float* p, const float* g, float* m, float* v, bf16* p_bf16, size_t n, float lr, float beta1,
                    float beta2, float eps, float wd, int step, cudaStream_t s) {
    if (!n) return OkStatus();
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, wd, bc1, bc2);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// every tensor in one launch (the per-tensor form spends 134 launches of ~4 us on tensors most of which are
// a few thousand elements): block -> its tensor by binary search in the table, then the same update
__global__ void __launch_bounds__(256)
    adamw_multi_kernel(const AdamSlot* __restrict__ table, int n_slots, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2) {
    int lo = 0, hi = n_slots - 1;
    while (lo < hi) {  // last entry whose first_block <= blockIdx.x
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].first_block <= blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const AdamSlot sl = table[lo];
    const unsigned long long i = (unsigned long long)(blockIdx.x - sl.first_block) * 256 + threadIdx.x;
    if (i >= sl.n) return;
    const unsigned long long j = sl.off + i;
    float pi = sl.p[i], gi = g[j];
    pi *= 1.f - lr * wd;
    const float mi = beta1 * m[j] + (1.f - beta1) * gi;
    const float vi = beta2 * v[j] + (1.f - beta2) * gi * gi;
    m[j] = mi;
    v[j] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    sl.p[i] = pi;
    if (sl.p_bf16) sl.p_bf16[i] = __float2bfloat16(pi);
}

Status launch_adamw_multi(const AdamSlot* table, int n_slots, unsigned int n_blocks, const float* g, float* m, float* v,
                          float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t s) {
    if (!n_blocks) return OkStatus();
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_multi_kernel<<<n_blocks, 256, 0, s>>>(table, n_slots, g, m, v, lr, beta1, beta2, eps, wd, bc1, bc2);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// attention backward.  Layout conventions are those of AttnFullParams (attention.cuh): element
// (b, head, row, d) of X at X + b*x_batch_stride + head*x_head_stride + row*x_row_stride + d.
// lse2[b][head][row] = m*log2(e) + log2(l) from the forward kernel, so p = exp2(s*log2(e) - lse2).
// The dQ kernel runs first and leaves delta[b][head][row] = sum_k P dP for the dK/dV kernel.
constexpr int kBwdT = 64;  // query and key tile

__device__ __forceinline__ void bwd_load_tile(bf16* dst, const bf16* src, int row_stride, int r0, int rmax) {
    // 64 rows x 8 chunks of 16 B, 128 threads -> 4 chunks each, XOR-swizzled like attn_full_kernel
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = threadIdx.x + i * 128;
        int row = c >> 3, ch = c & 7;
        bool pred = (r0 + row) < rmax;
        const bf16* g = src + (size_t)(pred ? r0 + row : 0) * row_stride + ch * 8;
        cp_async16(dst + row * kDKV + ((ch ^ (row & 7)) << 3), g, pred);
    }
}

// A fragments (16 rows of this warp x 64 cols) of a [64][64] swizzled tile
__device__ __forceinline__ void bwd_a_frags(uint32_t (&f)[4][4], const bf16* tile, int warp, int lane) {
    const uint32_t base = smem_u32(tile);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        int row = warp * 16 + (lane & 15);
        int ch = kk * 2 + (lane >> 4);
        ldmatrix_x4(f[kk][0], f[kk][1], f[kk][2], f[kk][3], base + row * 128 + ((ch ^ (row & 7)) << 4));
    }
}

// acc(16 x 64) += A(16 x 64 over k) * T^T where T is a [64 n][64 k] swizzled tile ("K-style" operand)
__device__ __forceinline__ void bwd_mma_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* tile, int lane) {
    const uint32_t base = smem_u32(tile);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            int row = nj * 16 + (lane & 7) + ((lane >> 4) << 3);
            int ch = kk * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, base + row * 128 + ((ch ^ (row & 7)) << 4));
            mma_bf16_16816(acc[nj * 2], a[kk], b0, b1);
            mma_bf16_16816(acc[nj * 2 + 1], a[kk], b2, b3);
        }
    }
}

// acc(16 x 64) += A(16 x 64 over k) * T where T is a [64 k][64 n] swizzled tile ("V-style" operand)
__device__ __forceinline__ void bwd_mma_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* tile, int lane) {
    const uint32_t base = smem_u32(tile);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            int ch = nj * 2 + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(b0, b1, b2, b3, base + row * 128 + ((ch ^ (row & 7)) << 4));
            mma_bf16_16816(acc[nj * 2], a[kk], b0, b1);
            mma_bf16_16816(acc[nj * 2 + 1], a[kk], b2, b3);
        }
    }
}

// C fragment (16 x 64 fp32) -> A fragments (bf16), the FlashAttention register trick
__device__ __forceinline__ void bwd_c_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        a[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16(c[ni][0], c[ni][1]);
        a[ni >> 1][(ni & 1) * 2 + 1] = pack_bf16(c[ni][2], c[ni][3]);
    }
}

__device__ __forceinline__ void bwd_zero(float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) c[i][r] = 0.f;
}

// dQ: one CTA per (query tile, head, batch); loops over the key tiles.
//   S = Q K^T, P = exp2(S log2e - lse2[row]), dP = dO V^T, dS = P * (dP - delta[row]), dQ += dS K
// K/V tiles are double-buffered: once the Q / dO fragments are in registers their shared-memory
// tiles become the second stage, so the next tile's cp.async overlaps this tile's mma.
template <int CTAS>  // resident CTAs per SM the register budget is cut for (2: 255, 3: 168 registers)
__global__ void __launch_bounds__(128, CTAS)
    attn_bwd_dq_kernel(AttnBwdParams p) {
    __shared__ __align__(128) bf16 sA[2][kBwdT * kDKV];  // stage s: K tile   (stage 1 holds Q first)
    __shared__ __align__(128) bf16 sB[2][kBwdT * kDKV];  // stage s: V tile   (stage 1 holds dO first)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * kBwdT, head = blockIdx.y, b = blockIdx.z;
    const bf16* Q = p.Q + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
    const bf16* K = p.K + (size_t)b * p.k_batch_stride + head * p.k_head_stride;
    const bf16* V = p.V + (size_t)b * p.v_batch_stride + head * p.v_head_stride;
    const bf16* dO = p.dO + (size_t)b * p.o_batch_stride + head * p.o_head_stride;
    const float kLog2e = 1.4426950408889634f;

    int n_kt = (p.Tk + kBwdT - 1) / kBwdT;
    if (p.causal) n_kt = min(n_kt, max(0, (q0 + kBwdT - 1 + p.causal_offset) / kBwdT + 1));

    bwd_load_tile(sA[1], Q, p.q_row_stride, q0, p.Tq);
    bwd_load_tile(sB[1], dO, p.o_row_stride, q0, p.Tq);
    cp_async_commit();
    if (n_kt > 0) {
        bwd_load_tile(sA[0], K, p.k_row_stride, 0, p.Tk);
        bwd_load_tile(sB[0], V, p.v_row_stride, 0, p.Tk);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    uint32_t qf[4][4], dof[4][4];
    bwd_a_frags(qf, sA[1], warp, lane);
    bwd_a_frags(dof, sB[1], warp, lane);
    __syncthreads();  // stage 1 may now be overwritten

    const int row_lo = q0 + warp * 16 + (lane >> 2);
    const unsigned long long bh_row0 = (unsigned long long)(b * kHeads + head) * p.Tq;
    const int keep_words = ((p.Tk + kBwdT - 1) / kBwdT) * 4;
    const float fscale = p.drop.on() ? p.drop.scale : 1.f;
    float lse[2], dl[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int r = row_lo + h * 8;
        lse[h] = p.lse2[((size_t)b * kHeads + head) * p.Tq + min(r, p.Tq - 1)];
    }
    float dq[8][4];
    bwd_zero(dq);
    // pass 0: delta[row] = sum_k P dP with the SAME recomputed P and dP that pass 1 uses, so that
    // dS = P (dP - delta) sums to zero over every row exactly as in exact arithmetic (taking
    // delta = sum_d dO O from the bf16 output instead leaves a 2^-9 mismatch that dominates dS on
    // sharply peaked rows).  pass 1: dQ.  The tile sequence runs through both passes: item
    // i = pass * n_kt + kt lives in stage i & 1, and item i + 1 is requested before item i is used.
    const int n_items = 2 * n_kt;
    for (int i = 0; i < n_items; ++i) {
        const int st = i & 1, kt = i % n_kt, pass = i / n_kt;
        if (i + 1 < n_items) {
            const int kn = (i + 1) % n_kt;
            bwd_load_tile(sA[st ^ 1], K, p.k_row_stride, kn * kBwdT, p.Tk);
            bwd_load_tile(sB[st ^ 1], V, p.v_row_stride, kn * kBwdT, p.Tk);
        }
        cp_async_commit();
        uint32_t kb[2] = {0u, 0u};
        if (p.drop.on()) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                kb[h] = __ldg(p.keep + (bh_row0 + min(row_lo + h * 8, p.Tq - 1)) * keep_words + kt * 4 + (lane & 3));
        }
        cp_async_wait<1>();
        __syncthreads();
        float s[8][4], dp[8][4];
        bwd_zero(s);
        bwd_zero(dp);
        bwd_mma_nt(s, qf, sA[st], lane);
        bwd_mma_nt(dp, dof, sB[st], lane);
        // dropout factor of element (ni, r) from the forward's keep bits
        auto fdrop = [&](int ni, int r) -> float {
            return !p.drop.on() || ((kb[r >> 1] >> (2 * ni + (r & 1))) & 1u) ? fscale : 0.f;
        };
        // only the tiles on the key-length / causal boundary need per-element predicates
        const bool need_mask = ((kt + 1) * kBwdT > p.Tk) || (p.causal && ((kt + 1) * kBwdT - 1 > q0 + p.causal_offset));
        auto visible = [&](int ni, int r) -> bool {
            const int key = kt * kBwdT + ni * 8 + (lane & 3) * 2 + (r & 1);
            const int row = row_lo + ((r >> 1) << 3);
            return key < p.Tk && (!p.causal || key <= row + p.causal_offset);
        };
        if (pass == 0) {
            if (need_mask) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (visible(ni, r)) dl[r >> 1] += fast_exp2(s[ni][r] * kLog2e - lse[r >> 1]) * dp[ni][r] * fdrop(ni, r);
            } else {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int r = 0; r < 4; ++r) dl[r >> 1] += fast_exp2(s[ni][r] * kLog2e - lse[r >> 1]) * dp[ni][r] * fdrop(ni, r);
            }
            if (kt == n_kt - 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    dl[h] += __shfl_xor_sync(0xffffffffu, dl[h], 1);
                    dl[h] += __shfl_xor_sync(0xffffffffu, dl[h], 2);
                    int r = row_lo + h * 8;
                    if ((lane & 3) == 0 && r < p.Tq) p.delta[((size_t)b * kHeads + head) * p.Tq + r] = dl[h];
                }
            }
        } else {
            auto ds_tile = [&](auto masked) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        bool ok = true;
                        if constexpr (decltype(masked)::value) ok = visible(ni, r);
                        const float pr = ok ? fast_exp2(s[ni][r] * kLog2e - lse[r >> 1]) : 0.f;
                        // dropout sits between softmax and P V: dP = dP' * m / (1 - p)
                        const float dpe = dp[ni][r] * fdrop(ni, r);
                        s[ni][r] = pr * (dpe - dl[r >> 1]);  // dS
                    }
                }
            };
            if (need_mask) ds_tile(std::true_type{});
            else ds_tile(std::false_type{});
            uint32_t dsf[4][4];
            bwd_c_to_a(dsf, s);
            bwd_mma_nn(dq, dsf, sA[st], lane);
        }
        __syncthreads();  // every warp is done with stage st before it is refilled
    }
    cp_async_wait<0>();
    bf16* dQ = p.dQ + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        int col = ni * 8 + (lane & 3) * 2;
        if (row_lo < p.Tq)
            *reinterpret_cast<uint32_t*>(dQ + (size_t)row_lo * p.q_row_stride + col) = pack_bf16(dq[ni][0], dq[ni][1]);
        if (row_lo + 8 < p.Tq)
            *reinterpret_cast<uint32_t*>(dQ + (size_t)(row_lo + 8) * p.q_row_stride + col) = pack_bf16(dq[ni][2], dq[ni][3]);
    }
}

// dQ in ONE pass over the key tiles (default; MRMT3_ATTN_BWD_DQ_PASSES=2 selects the kernel above).
// The two-pass kernel exists because delta must be consistent with the recomputed P and dP: taking
// delta~ = sum_d dO O from the forward's bf16 output leaves an error eps = delta - delta~ of relative size
// 2^-9, which on sharply peaked rows is as large as dS itself.  That error is, however, a per-row SCALAR,
// and dQ is linear in it:
//     dQ = sum_k P (dP - delta) K = sum_k P (dP - delta~) K  -  eps * sum_k P K
// so one pass accumulates U = sum_k [P (dP - delta~)] K and W = sum_k P K on the tensor cores and
// delta = sum_k P dP in fp32 on the side, and the end applies dQ = U - (delta - delta~) W.  The rounding of
// P (dP - delta~) to bf16 now costs 2^-9 * P * |eps| ~ 2^-18 |delta|: the accuracy of the two-pass form at
// 4 MMA groups (S, dP, U, W) and ONE sweep of exponentials / masks / keep bits instead of 5 and two.
// delta (exact, for the dK/dV kernel) is written out as before.  Q / dO stay in shared memory and their A
// fragments are re-read per tile (8 ldmatrix), which keeps the kernel at three CTAs per SM.
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS)
    attn_bwd_dq1_kernel(AttnBwdParams p) {
    __shared__ __align__(128) bf16 sQ[kBwdT * kDKV];
    __shared__ __align__(128) bf16 sdO[kBwdT * kDKV];
    __shared__ __align__(128) bf16 sA[2][kBwdT * kDKV];  // stage s: K tile   (stage 1 holds O first)
    __shared__ __align__(128) bf16 sB[2][kBwdT * kDKV];  // stage s: V tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * kBwdT, head = blockIdx.y, b = blockIdx.z;
    const bf16* Q = p.Q + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
    const bf16* K = p.K + (size_t)b * p.k_batch_stride + head * p.k_head_stride;
    const bf16* V = p.V + (size_t)b * p.v_batch_stride + head * p.v_head_stride;
    const bf16* O = p.O + (size_t)b * p.o_batch_stride + head * p.o_head_stride;
    const bf16* dO = p.dO + (size_t)b * p.o_batch_stride + head * p.o_head_stride;
    const float kLog2e = 1.4426950408889634f;

    int n_kt = (p.Tk + kBwdT - 1) / kBwdT;
    if (p.causal) n_kt = min(n_kt, max(0, (q0 + kBwdT - 1 + p.causal_offset) / kBwdT + 1));

    bwd_load_tile(sQ, Q, p.q_row_stride, q0, p.Tq);
    bwd_load_tile(sdO, dO, p.o_row_stride, q0, p.Tq);
    bwd_load_tile(sA[1], O, p.o_row_stride, q0, p.Tq);
    cp_async_commit();
    if (n_kt > 0) {
        bwd_load_tile(sA[0], K, p.k_row_stride, 0, p.Tk);
        bwd_load_tile(sB[0], V, p.v_row_stride, 0, p.Tk);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const int row_lo = q0 + warp * 16 + (lane >> 2);
    const unsigned long long bh_row0 = (unsigned long long)(b * kHeads + head) * p.Tq;
    const int keep_words = ((p.Tk + kBwdT - 1) / kBwdT) * 4;
    const float fscale = p.drop.on() ? p.drop.scale : 1.f;
    // delta~[row] = sum_d dO O from the A fragments of the two tiles (this thread: rows row_lo, row_lo + 8)
    float dt[2] = {0.f, 0.f};
    {
        uint32_t of[4][4], dof[4][4];
        bwd_a_frags(of, sA[1], warp, lane);
        bwd_a_frags(dof, sdO, warp, lane);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // registers 0, 2: row lane / 4; 1, 3: row lane / 4 + 8
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&of[kk][j]));
                const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dof[kk][j]));
                dt[j & 1] += a.x * g.x + a.y * g.y;
            }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            dt[h] += __shfl_xor_sync(0xffffffffu, dt[h], 1);
            dt[h] += __shfl_xor_sync(0xffffffffu, dt[h], 2);
        }
    }
    __syncthreads();  // stage 1 (O) may now be overwritten

    float lse[2], dl[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) lse[h] = p.lse2[((size_t)b * kHeads + head) * p.Tq + min(row_lo + h * 8, p.Tq - 1)];
    float u[8][4], w[8][4];
    bwd_zero(u);
    bwd_zero(w);
    for (int kt = 0; kt < n_kt; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < n_kt) {
            bwd_load_tile(sA[st ^ 1], K, p.k_row_stride, (kt + 1) * kBwdT, p.Tk);
            bwd_load_tile(sB[st ^ 1], V, p.v_row_stride, (kt + 1) * kBwdT, p.Tk);
        }
        cp_async_commit();
        uint32_t kb[2] = {0u, 0u};
        if (p.drop.on()) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                kb[h] = __ldg(p.keep + (bh_row0 + min(row_lo + h * 8, p.Tq - 1)) * keep_words + kt * 4 + (lane & 3));
        }
        cp_async_wait<1>();
        __syncthreads();
        float s[8][4], dp[8][4];
        bwd_zero(s);
        bwd_zero(dp);
        {
            uint32_t af[4][4];
            bwd_a_frags(af, sQ, warp, lane);
            bwd_mma_nt(s, af, sA[st], lane);
            bwd_a_frags(af, sdO, warp, lane);
            bwd_mma_nt(dp, af, sB[st], lane);
        }
        const bool need_mask = ((kt + 1) * kBwdT > p.Tk) || (p.causal && ((kt + 1) * kBwdT - 1 > q0 + p.causal_offset));
        auto tile = [&](auto masked) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    bool ok = true;
                    if constexpr (decltype(masked)::value) {
                        const int key = kt * kBwdT + ni * 8 + (lane & 3) * 2 + (r & 1);
                        const int row = row_lo + ((r >> 1) << 3);
                        ok = key < p.Tk && (!p.causal || key <= row + p.causal_offset);
                    }
                    const float pr = ok ? fast_exp2(s[ni][r] * kLog2e - lse[r >> 1]) : 0.f;
                    const float mk = !p.drop.on() || ((kb[r >> 1] >> (2 * ni + (r & 1))) & 1u) ? fscale : 0.f;
                    const float dpe = dp[ni][r] * mk;       // dropout sits between softmax and P V
                    dl[r >> 1] += pr * dpe;
                    s[ni][r] = pr;                          // P
                    dp[ni][r] = pr * (dpe - dt[r >> 1]);    // P (dP - delta~)
                }
            }
        };
        if (need_mask) tile(std::true_type{});
        else tile(std::false_type{});
        uint32_t af[4][4];
        bwd_c_to_a(af, dp);
        bwd_mma_nn(u, af, sA[st], lane);
        bwd_c_to_a(af, s);
        bwd_mma_nn(w, af, sA[st], lane);
        __syncthreads();  // every warp is done with stage st before it is refilled
    }
    cp_async_wait<0>();
    float eps[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        dl[h] += __shfl_xor_sync(0xffffffffu, dl[h], 1);
        dl[h] += __shfl_xor_sync(0xffffffffu, dl[h], 2);
        eps[h] = dl[h] - dt[h];
        const int r = row_lo + h * 8;
        if ((lane & 3) == 0 && r < p.Tq) p.delta[((size_t)b * kHeads + head) * p.Tq + r] = dl[h];
    }
    bf16* dQ = p.dQ + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        int col = ni * 8 + (lane & 3) * 2;
        if (row_lo < p.Tq)
            *reinterpret_cast<uint32_t*>(dQ + (size_t)row_lo * p.q_row_stride + col) =
                pack_bf16(u[ni][0] - eps[0] * w[ni][0], u[ni][1] - eps[0] * w[ni][1]);
        if (row_lo + 8 < p.Tq)
            *reinterpret_cast<uint32_t*>(dQ + (size_t)(row_lo + 8) * p.q_row_stride + col) =
                pack_bf16(u[ni][2] - eps[1] * w[ni][2], u[ni][3] - eps[1] * w[ni][3]);
    }
}

// dK, dV: one CTA per (key tile, head, batch); loops over the query tiles, everything transposed
// (rows = keys):  S^T = K Q^T, P^T = exp2(S^T log2e - lse2[col]), dV += P^T dO,
//                 dP^T = V dO^T, dS^T = P^T * (dP^T - delta[col]), dK += dS^T Q
// Q / dO tiles are double-buffered the same way (the K / V tiles become stage 1).
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS)
    attn_bwd_dkv_kernel(AttnBwdParams p) {
    __shared__ __align__(128) bf16 sA[2][kBwdT * kDKV];  // stage s: Q tile   (stage 1 holds K first)
    __shared__ __align__(128) bf16 sB[2][kBwdT * kDKV];  // stage s: dO tile  (stage 1 holds V first)
    __shared__ float s_lse[2][kBwdT], s_dl[2][kBwdT];
    __shared__ __align__(8) unsigned short s_keep[2][kBwdT][4];   // keep words of (query row, this key tile)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.x * kBwdT, head = blockIdx.y, b = blockIdx.z;
    const bf16* Q = p.Q + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
    const bf16* K = p.K + (size_t)b * p.k_batch_stride + head * p.k_head_stride;
    const bf16* V = p.V + (size_t)b * p.v_batch_stride + head * p.v_head_stride;
    const bf16* dO = p.dO + (size_t)b * p.o_batch_stride + head * p.o_head_stride;
    const float kLog2e = 1.4426950408889634f;

    const int n_qt = (p.Tq + kBwdT - 1) / kBwdT;
    int qt0 = 0;
    if (p.causal) qt0 = max(0, (k0 - p.causal_offset) / kBwdT);  // first query tile that can see key k0
    // per-row statistics of a query tile ride in the same cp.async group as the tile itself: a plain
    // load + shared store here would expose one global-load latency per tile in front of the barrier
    auto load_stats = [&](int stg, int qt) {
        if (threadIdx.x < kBwdT) {
            int r = qt * kBwdT + threadIdx.x;
            size_t idx = ((size_t)b * kHeads + head) * p.Tq + min(r, p.Tq - 1);
            cp_async4(&s_lse[stg][threadIdx.x], p.lse2 + idx);
            cp_async4(&s_dl[stg][threadIdx.x], p.delta + idx);
            // the tile's 64 x 4 keep words too (8 bytes per row): sixteen scattered 2-byte global loads per
            // thread and tile otherwise (ncu: long_scoreboard 1.44 per issue in this kernel, 0.6 in dQ)
            if (p.drop.on())
                cp_async8(&s_keep[stg][threadIdx.x][0],
                          p.keep + idx * (size_t)(((p.Tk + kBwdT - 1) / kBwdT) * 4) + blockIdx.x * 4);
        }
    };

    bwd_load_tile(sA[1], K, p.k_row_stride, k0, p.Tk);
    bwd_load_tile(sB[1], V, p.v_row_stride, k0, p.Tk);
    cp_async_commit();
    if (qt0 < n_qt) {
        bwd_load_tile(sA[0], Q, p.q_row_stride, qt0 * kBwdT, p.Tq);
        bwd_load_tile(sB[0], dO, p.o_row_stride, qt0 * kBwdT, p.Tq);
        load_stats(0, qt0);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    bwd_a_frags(kf, sA[1], warp, lane);
    bwd_a_frags(vf, sB[1], warp, lane);
    __syncthreads();

    const int key_lo = k0 + warp * 16 + (lane >> 2);  // this thread's keys: key_lo, key_lo + 8
    const unsigned long long bh_row0 = (unsigned long long)(b * kHeads + head) * p.Tq;
    const int keep_words = ((p.Tk + kBwdT - 1) / kBwdT) * 4;
    // key_lo = 64 kt + 8 (2 warp) + 2 (lane >> 3) + ((lane >> 2) & 1): word lane >> 3, bit 2 (2 warp) + e; key_lo + 8 two bits up
    const int keep_shift = 4 * warp + ((lane >> 2) & 1);
    const float fscale = p.drop.on() ? p.drop.scale : 1.f;
    float dk[8][4], dv[8][4];
    bwd_zero(dk);
    bwd_zero(dv);
    for (int qt = qt0; qt < n_qt; ++qt) {
        const int st = (qt - qt0) & 1;
        if (qt + 1 < n_qt) {
            bwd_load_tile(sA[st ^ 1], Q, p.q_row_stride, (qt + 1) * kBwdT, p.Tq);
            bwd_load_tile(sB[st ^ 1], dO, p.o_row_stride, (qt + 1) * kBwdT, p.Tq);
            load_stats(st ^ 1, qt + 1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        // keep bits of this thread's 16 query rows x 2 keys: both keys sit in the same saved word
        uint32_t kbw[8][2];
        if (p.drop.on()) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) kbw[ni][e] = (uint32_t)s_keep[st][ni * 8 + (lane & 3) * 2 + e][lane >> 3] >> keep_shift;
        }
        float stt[8][4], dpt[8][4];
        bwd_zero(stt);
        bwd_zero(dpt);
        bwd_mma_nt(stt, kf, sA[st], lane);   // S^T = K Q^T
        bwd_mma_nt(dpt, vf, sB[st], lane);   // dP^T = V dO^T
        float pt[8][4];
        // only the tiles on a length / causal boundary need per-element predicates (uniform over the CTA)
        const bool need_mask = ((qt + 1) * kBwdT > p.Tq) || (k0 + kBwdT > p.Tk) ||
                               (p.causal && (k0 + kBwdT - 1 > qt * kBwdT + p.causal_offset));
        auto pds_tile = [&](auto masked) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int qc = ni * 8 + (lane & 3) * 2 + (r & 1);
                    bool ok = true;
                    if constexpr (decltype(masked)::value) {
                        const int row = qt * kBwdT + qc;              // query index
                        const int key = key_lo + ((r >> 1) << 3);
                        ok = row < p.Tq && key < p.Tk && (!p.causal || key <= row + p.causal_offset);
                    }
                    const float pr = ok ? fast_exp2(stt[ni][r] * kLog2e - s_lse[st][qc]) : 0.f;
                    const float mk = !p.drop.on() || ((kbw[ni][r & 1] >> ((r >> 1) << 1)) & 1u) ? fscale : 0.f;
                    pt[ni][r] = pr * mk;                                   // P'^T = dropout(P)^T
                    stt[ni][r] = pr * (dpt[ni][r] * mk - s_dl[st][qc]);    // dS^T
                }
            }
        };
        if (need_mask) pds_tile(std::true_type{});
        else pds_tile(std::false_type{});
        uint32_t af[4][4];
        bwd_c_to_a(af, pt);
        bwd_mma_nn(dv, af, sB[st], lane);  // dV += P^T dO
        bwd_c_to_a(af, stt);
        bwd_mma_nn(dk, af, sA[st], lane);  // dK += dS^T Q
        __syncthreads();
    }
    cp_async_wait<0>();
    bf16* dK = p.dK + (size_t)b * p.dk_batch_stride + head * p.dk_head_stride;
    bf16* dV = p.dV + (size_t)b * p.dk_batch_stride + head * p.dk_head_stride;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        int col = ni * 8 + (lane & 3) * 2;
        if (key_lo < p.Tk) {
            *reinterpret_cast<uint32_t*>(dK + (size_t)key_lo * p.dk_row_stride + col) = pack_bf16(dk[ni][0], dk[ni][1]);
            *reinterpret_cast<uint32_t*>(dV + (size_t)key_lo * p.dk_row_stride + col) = pack_bf16(dv[ni][0], dv[ni][1]);
        }
        if (key_lo + 8 < p.Tk) {
            *reinterpret_cast<uint32_t*>(dK + (size_t)(key_lo + 8) * p.dk_row_stride + col) = pack_bf16(dk[ni][2], dk[ni][3]);
            *reinterpret_cast<uint32_t*>(dV + (size_t)(key_lo + 8) * p.dk_row_stride + col) = pack_bf16(dv[ni][2], dv[ni][3]);
        }
    }
}

Status launch_attn_bwd(const AttnBwdParams& p, int batch, cudaStream_t s) {
    if (batch <= 0 || p.Tq <= 0 || p.Tk <= 0) return OkStatus();
    // Register budgets (resident CTAs per SM), chosen by measurement on the fine-tune step
    // (profiles/r3b_bench_finetune_*.json: dQ/dKV at 2/2 CTAs 25.9 ms per step, 2/3 26.3, 3/3 26.6): both kernels
    // at 2 CTAs per SM (244 / 255 registers, no spills) beat 3 CTAs with spills now that the element-wise work
    // is small; MRMT3_ATTN_BWD_DQ_CTAS / MRMT3_ATTN_BWD_DKV_CTAS / MRMT3_ATTN_BWD_DQ_PASSES for A/B runs
    static const int dq_ctas = [] {
        const char* e = getenv("MRMT3_ATTN_BWD_DQ_CTAS");
        return e && atoi(e) == 3 ? 3 : 2;
    }();
    static const int dkv_ctas = [] {
        const char* e = getenv("MRMT3_ATTN_BWD_DKV_CTAS");
        return e && atoi(e) == 3 ? 3 : 2;
    }();
    static const int dq_passes = [] {
        const char* e = getenv("MRMT3_ATTN_BWD_DQ_PASSES");
        return e && atoi(e) == 2 ? 2 : 1;
    }();
    const dim3 gq(ceil_div(p.Tq, kBwdT), kHeads, batch), gk(ceil_div(p.Tk, kBwdT), kHeads, batch);
    if (dq_passes == 1) {
        if (dq_ctas == 2) attn_bwd_dq1_kernel<2><<<gq, 128, 0, s>>>(p);
        else attn_bwd_dq1_kernel<3><<<gq, 128, 0, s>>>(p);
    } else {
        if (dq_ctas == 2) attn_bwd_dq_kernel<2><<<gq, 128, 0, s>>>(p);
        else attn_bwd_dq_kernel<3><<<gq, 128, 0, s>>>(p);
    }
    MRMT3_CHECK_LAUNCH();
    if (dkv_ctas == 3) attn_bwd_dkv_kernel<3><<<gk, 128, 0, s>>>(p);
    else attn_bwd_dkv_kernel<2><<<gk, 128, 0, s>>>(p);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

}  // namespace mrmt3
