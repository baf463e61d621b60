#include "layers.cuh"

namespace mrmt3 {

// ---------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 v = src[i];
        dst[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

Status launch_cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t s) {
    if (n == 0) return OkStatus();
    if (n % 4) return Error(2, "cast_bf16: n must be a multiple of 4");
    size_t n4 = n / 4;
    int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
    cast_bf16_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                            reinterpret_cast<uint2*>(dst), n4);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows,
                                   int cols, int row_mul, int row_off) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)rows * cols;
    if (i >= n) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    dst[((size_t)r * row_mul + row_off) * cols + c] = __float2bfloat16(src[i]);
}

Status launch_pack_weight(const float* src, bf16* dst, int rows, int cols, int row_mul, int row_off,
                          cudaStream_t s) {
    size_t n = (size_t)rows * cols;
    pack_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, rows, cols, row_mul, row_off);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void pack_weight_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                       int cols, int row_mul, int row_off) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)rows * cols;
    if (i >= n) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    dst[((size_t)r * row_mul + row_off) * cols + c] = src[i];
}

Status launch_pack_weight_f32(const float* src, float* dst, int rows, int cols, int row_mul, int row_off,
                              cudaStream_t s) {
    size_t n = (size_t)rows * cols;
    pack_weight_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, rows, cols, row_mul, row_off);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void fold_norm_kernel(const float* __restrict__ master, const float* __restrict__ g,
                                 bf16* __restrict__ dst, int rows, int cols) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    dst[i] = __float2bfloat16(master[i] * g[i % cols]);
}

Status launch_fold_norm(const float* master, const float* g, bf16* dst, int rows, int cols, cudaStream_t s) {
    size_t n = (size_t)rows * cols;
    fold_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(master, g, dst, rows, cols);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
// one warp per row of 512
__global__ void __launch_bounds__(256)
    rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float eps,
                   bf16* __restrict__ out_bf16, float* __restrict__ out_f32, int rows,
                   const int* __restrict__ active, int rows_per_lane) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (active && !active[row / rows_per_lane]) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * kDModel);
    float4 v[4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = xr[lane + i * 32];
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss * (1.0f / kDModel) + eps);
    const float4* wr = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 g = wr[lane + i * 32];
        float4 y = make_float4(g.x * (v[i].x * r), g.y * (v[i].y * r), g.z * (v[i].z * r),
                               g.w * (v[i].w * r));
        size_t o = (size_t)row * kDModel + (lane + i * 32) * 4;
        if (out_bf16)
            *reinterpret_cast<uint2*>(out_bf16 + o) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = y;
    }
}

Status launch_rmsnorm(const float* x, const float* w, float eps, bf16* out_bf16, float* out_f32,
                      int rows, const int* active, int rows_per_lane, cudaStream_t s) {
    if (rows <= 0) return OkStatus();
    rmsnorm_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, w, eps, out_bf16, out_f32, rows, active,
                                                     rows_per_lane > 0 ? rows_per_lane : 1);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ emb,
                                    const float* __restrict__ pe, float* __restrict__ H, int L,
                                    int pos0, int rows) {
    int row = blockIdx.x * 2 + (threadIdx.x >> 7);
    if (row >= rows) return;
    int c = (threadIdx.x & 127) * 4;
    long long id = ids[row];
    if (id < 0 || id >= kVocab) id = 0;  // -100 (ignore_index) and junk read the pad row, never out of bounds
    int pos = pos0 + row % L;
    float4 e = *reinterpret_cast<const float4*>(emb + (size_t)id * kDModel + c);
    float4 p = *reinterpret_cast<const float4*>(pe + (size_t)pos * kDModel + c);
    *reinterpret_cast<float4*>(H + (size_t)row * kDModel + c) =
        make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
}

// V1 (T5SegMem) teacher forcing: the decoder's input rows of sequence b are [memory rows of b ; token rows of
// b], all with the positional encoding of their place in the (n_mem + L)-long sequence
// (reference models/t5_segmem.py:136-139 + the stack input of models/t5.py:596-598)
__global__ void embed_tokens_prefixed_kernel(const long long* __restrict__ ids, const float* __restrict__ emb,
                                             const float* __restrict__ pe, const float* __restrict__ mem,
                                             float* __restrict__ H, int L, int n_mem, int rows) {
    int row = blockIdx.x * 2 + (threadIdx.x >> 7);
    if (row >= rows) return;
    const int c = (threadIdx.x & 127) * 4;
    const int Lt = L + n_mem, b = row / Lt, t = row - b * Lt;
    float4 e;
    if (t < n_mem) {
        e = *reinterpret_cast<const float4*>(mem + ((size_t)b * n_mem + t) * kDModel + c);
    } else {
        long long id = ids[(size_t)b * L + (t - n_mem)];
        if (id < 0 || id >= kVocab) id = 0;
        e = *reinterpret_cast<const float4*>(emb + (size_t)id * kDModel + c);
    }
    const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)t * kDModel + c);
    *reinterpret_cast<float4*>(H + (size_t)row * kDModel + c) = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
}

Status launch_embed_tokens_prefixed(const long long* ids, const float* emb, const float* pe, const float* mem, float* H,
                                    int B, int L, int n_mem, cudaStream_t s) {
    const int rows = B * (L + n_mem);
    if (rows <= 0) return OkStatus();
    embed_tokens_prefixed_kernel<<<ceil_div(rows, 2), 256, 0, s>>>(ids, emb, pe, mem, H, L, n_mem, rows);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

Status launch_embed_tokens(const long long* ids, const float* emb, const float* pe, float* H, int B,
                           int L, int pos0, cudaStream_t s) {
    int rows = B * L;
    if (rows <= 0) return OkStatus();
    embed_tokens_kernel<<<ceil_div(rows, 2), 256, 0, s>>>(ids, emb, pe, H, L, pos0, rows);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void embed_bf16_kernel(const long long* __restrict__ ids, long ids_row_stride,
                                  int rows_per_lane, const int* __restrict__ lane_src_row,
                                  const float* __restrict__ emb, bf16* __restrict__ out, int rows) {
    int row = blockIdx.x * 2 + (threadIdx.x >> 7);
    if (row >= rows) return;
    int c = (threadIdx.x & 127) * 4;
    int lane = row / rows_per_lane, j = row % rows_per_lane;
    long src_row = lane_src_row ? lane_src_row[lane] : lane;
    long long id = ids[src_row * ids_row_stride + j];
    if (id < 0 || id >= kVocab) id = 0;  // targets_prev is padded with -100 (reference masks it to pad)
    float4 e = *reinterpret_cast<const float4*>(emb + (size_t)id * kDModel + c);
    *reinterpret_cast<uint2*>(out + (size_t)row * kDModel + c) =
        make_uint2(pack_bf16(e.x, e.y), pack_bf16(e.z, e.w));
}

Status launch_embed_bf16(const long long* ids, long ids_row_stride, int rows_per_lane, int n_lanes,
                         const int* lane_src_row, const float* emb, bf16* out, cudaStream_t s) {
    int rows = rows_per_lane * n_lanes;
    if (rows <= 0) return OkStatus();
    embed_bf16_kernel<<<ceil_div(rows, 2), 256, 0, s>>>(ids, ids_row_stride, rows_per_lane,
                                                        lane_src_row, emb, out, rows);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int n,
                                   int src_block, int rows) {
    int row = blockIdx.x * 2 + (threadIdx.x >> 7);
    if (row >= rows) return;
    int c = (threadIdx.x & 127) * 4;
    int b = row / n, j = row % n;
    *reinterpret_cast<float4*>(dst + (size_t)row * kDModel + c) =
        *reinterpret_cast<const float4*>(src + ((size_t)b * src_block + j) * kDModel + c);
}

Status launch_gather_rows(const float* src, float* dst, int n_lanes, int n, int src_block,
                          cudaStream_t s) {
    int rows = n_lanes * n;
    if (rows <= 0) return OkStatus();
    gather_rows_kernel<<<ceil_div(rows, 2), 256, 0, s>>>(src, dst, n, src_block, rows);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// ---------------------------------------------------------------------------------------------
__global__ void decode_embed_kernel(DecodeState st, const float* __restrict__ emb,
                                    const float* __restrict__ pe, const float* __restrict__ prefix,
                                    int prefix_stride, float* __restrict__ H, bf16* __restrict__ Hb,
                                    int n_lanes) {
    trace_begin(st.trace);
    pdl_wait();  // tok / active / step come from the previous step's arg-max kernel
    pdl_launch_dependents();
    int lane = blockIdx.x * 2 + (threadIdx.x >> 7);
    if (lane >= n_lanes) return;
    if (!st.active[lane]) return;
    int c = (threadIdx.x & 127) * 4;
    const int pos = st.step[0];
    float4 e;
    if (prefix && pos < st.prefix_len)
        e = *reinterpret_cast<const float4*>(prefix + (size_t)lane * prefix_stride + (size_t)pos * kDModel + c);
    else
        e = *reinterpret_cast<const float4*>(emb + (size_t)st.tok[lane] * kDModel + c);
    float4 p = *reinterpret_cast<const float4*>(pe + (size_t)pos * kDModel + c);
    const float4 hv = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
    *reinterpret_cast<float4*>(H + (size_t)lane * kDModel + c) = hv;
    *reinterpret_cast<uint2*>(Hb + (size_t)lane * kDModel + c) =
        make_uint2(pack_bf16(hv.x, hv.y), pack_bf16(hv.z, hv.w));
    trace_end(st.trace);
}

Status launch_decode_embed(const DecodeState& st, const float* emb, const float* pe,
                           const float* prefix, int prefix_stride, float* H, bf16* Hb, int n_lanes,
                           cudaStream_t s) {
    MRMT3_TRY(launch_pdl(decode_embed_kernel, dim3(ceil_div(n_lanes, 2)), dim3(256), 0, s, st, emb, pe, prefix,
                         prefix_stride, H, Hb, n_lanes));
    return OkStatus();
}

// one warp per lane; 8 lanes per CTA.  The last CTA to finish advances the step counter.
__global__ void __launch_bounds__(256)
    argmax_advance_kernel(DecodeState st, const float* __restrict__ logits, size_t lane_stride,
                          size_t step_stride, int n_lanes, int vocab) {
    trace_begin(st.trace);
    pdl_wait();
    pdl_launch_dependents();
    const int lane = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int t = threadIdx.x & 31;
    const int step = st.step[0];
    if (lane < n_lanes && st.active[lane]) {
        size_t off = step_stride ? (size_t)st.out_row[lane] * lane_stride +
                                       (size_t)(step - st.prefix_len) * step_stride
                                 : (size_t)lane * lane_stride;
        const float4* row = reinterpret_cast<const float4*>(logits + off);
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        for (int i = t; i < vocab / 4; i += 32) {
            float4 v = row[i];
            int b = i * 4;
            // strict '>' keeps the lowest index among equal values within this thread
            if (v.x > best) { best = v.x; best_i = b; }
            if (v.y > best) { best = v.y; best_i = b + 1; }
            if (v.z > best) { best = v.z; best_i = b + 2; }
            if (v.w > best) { best = v.w; best_i = b + 3; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
        }
        if (t == 0) {
            const int n_emitted = step - st.prefix_len + 1;  // tokens emitted incl. this one
            int next = best_i;
            if (st.forced) {
                const size_t frow = st.forced_by_row ? (size_t)st.out_row[lane] : (size_t)lane;
                const long long f = st.forced[frow * st.forced_stride + n_emitted];
                next = (f < 0 || f >= vocab) ? st.pad_id : (int)f;
            }
            st.out[(size_t)st.out_row[lane] * st.out_stride + n_emitted] = next;
            st.tok[lane] = next;
            bool done = (!st.forced && next == st.eos_id) || n_emitted >= st.max_tokens;
            if (done) {
                st.active[lane] = 0;
                st.finish_step[lane] = n_emitted;
                atomicSub(st.n_active, 1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int tk = atomicAdd(st.ticket, 1);
        if (tk == (int)gridDim.x - 1) {
            st.ticket[0] = 0;
            st.step[0] = step + 1;
        }
    }
    trace_end(st.trace);
}

Status launch_argmax_advance(const DecodeState& st, const float* logits, size_t lane_stride,
                             size_t step_stride, int n_lanes, int vocab, cudaStream_t s) {
    MRMT3_TRY(launch_pdl(argmax_advance_kernel, dim3(ceil_div(n_lanes, 8)), dim3(256), 0, s, st, logits,
                         lane_stride, step_stride, n_lanes, vocab));
    return OkStatus();
}

__global__ void decode_init_kernel(DecodeState st, int n_lanes, const int* __restrict__ init_active,
                                   int n_active, int start_id, int n_groups, int group_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lanes) {
        int a = init_active ? init_active[i] : 1;
        st.active[i] = a;
        st.tok[i] = start_id;
        st.finish_step[i] = a ? st.max_tokens : 0;
        if (a) st.out[(size_t)st.out_row[i] * st.out_stride] = start_id;
    }
    if (i == 0) st.n_active[0] = n_active;
    if (i < n_groups) {
        st.step[i * group_stride] = 0;
        st.ticket[i * group_stride] = 0;
    }
}

Status launch_decode_init(const DecodeState& st, int n_lanes, const int* init_active, int n_active,
                          int start_id, int n_groups, int group_stride, cudaStream_t s) {
    decode_init_kernel<<<ceil_div(std::max(n_lanes, n_groups), 256), 256, 0, s>>>(
        st, n_lanes, init_active, n_active, start_id, n_groups, group_stride);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

__global__ void advance_only_kernel(DecodeState st) {
    pdl_wait();
    pdl_launch_dependents();
    st.step[0] += 1;
}

Status launch_advance_only(const DecodeState& st, cudaStream_t s) {
    MRMT3_TRY(launch_pdl(advance_only_kernel, dim3(1), dim3(1), 0, s, st));
    return OkStatus();
}

}  // namespace mrmt3
