// Whole-sequence attention forward on tcgen05 / TMEM (plan: DESIGN.md appendix A).  Same contract as
// attn_full_kernel (attention.cu): O = softmax(Q K^T) V without scale or bias, optional causal mask,
// optional log-sum-exp output, optional dropout on the weights with the keep bits saved in the layout
// the backward kernels read.
//
// One CTA per (128 query rows, head, sample), 192 threads, one CTA per SM:
//   warp 0     TMA producer: Q once, then (K_j, V_j) tiles of 128 keys through a 3-stage ring
//   warp 1     allocates TMEM; one lane issues   S_j = Q K_j^T     (K-major A and B, N = 128)
//                                                O_j = P_j V_j     (A = P from shared memory,
//                                                                   B = V as TMA lands it = MN-major, N = 64)
//              S is double-buffered in TMEM so S_{j+1} is computed while the softmax warps work on S_j
//   warps 2-5  one thread per query row: tcgen05.ld the row of S, mask, online softmax, dropout, P as
//              bf16 into a SWIZZLE_128B tile, then O_j from TMEM into the running fp32 row in registers
//              (o = o * scale + O_j: TMEM is never read-modify-written)
#include "attention.cuh"
#include "gemm_tcgen05.cuh"

namespace mrmt3 {

namespace {

constexpr int kQT = 128;                 // query rows per CTA
constexpr int kKT = 128;                 // keys per tile
constexpr int kStagesKV = 3;
constexpr int kTileBytes = 128 * kDKV * 2;             // 16 KB: 128 rows x 64 bf16
constexpr int kSmemQ = 0;
constexpr int kSmemKV = kSmemQ + kTileBytes;           // stage s: K at +0, V at +16 KB
constexpr int kSmemP = kSmemKV + kStagesKV * 2 * kTileBytes;   // two 128 x 64 sub-tiles
constexpr int kSmemBar = kSmemP + 2 * kTileBytes;
constexpr int kSmemTotal = kSmemBar + 256 + 1024;
constexpr int kTmemCols = 512;           // S[2] at 0 / 128, O tile at 256
constexpr int kTmemO = 256;

struct TcCoords {
    int row0, col;  // tensor-map coordinates of (sample, head, row 0)
};
struct AttnFullTcArgs {
    AttnFullParams p;
    // element (b, head, r, 0) of X is row x_rows_per_batch*b + x_rows_per_head*head + r, column
    // x_cols_per_head*head of X's tensor map
    int q_rows_per_batch, q_rows_per_head, q_cols_per_head;
    int k_rows_per_batch, k_rows_per_head, k_cols_per_head;
    int v_rows_per_batch, v_rows_per_head, v_cols_per_head;
};

__global__ void __launch_bounds__(192, 1)
    attn_full_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, AttnFullTcArgs a) {
    const AttnFullParams& p = a.p;
    extern __shared__ unsigned char attn_tc_smem[];
    const uint32_t raw = smem_u32(attn_tc_smem);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = attn_tc_smem + (base - raw);
    const uint32_t bar_q = base + kSmemBar;                 // 8 B
    const uint32_t bar_kv_full = bar_q + 8;                 // kStagesKV x 8 B
    const uint32_t bar_kv_empty = bar_kv_full + kStagesKV * 8;
    const uint32_t bar_s_full = bar_kv_empty + kStagesKV * 8;   // 2 x 8 B
    const uint32_t bar_s_free = bar_s_full + 16;                // 2 x 8 B, 128 arrivals
    const uint32_t bar_p_full = bar_s_free + 16;                // 128 arrivals
    const uint32_t bar_p_free = bar_p_full + 8;                 // tcgen05.commit
    const uint32_t bar_o_full = bar_p_free + 8;                 // tcgen05.commit
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + kSmemBar + 192);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQT, head = blockIdx.y, b = blockIdx.z;
    int n_kt = (p.Tk + kKT - 1) / kKT;
    if (p.causal) n_kt = min(n_kt, max(0, (q0 + kQT - 1 + p.causal_offset) / kKT + 1));

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < kStagesKV; ++s) {
            mbar_init(bar_kv_full + s * 8, 1);
            mbar_init(bar_kv_empty + s * 8, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_s_full + i * 8, 1);
            mbar_init(bar_s_free + i * 8, 128);
        }
        mbar_init(bar_p_full, 128);
        mbar_init(bar_p_free, 1);
        mbar_init(bar_o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         smem_u32((const void*)tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && n_kt > 0) {
            mbar_expect_tx(bar_q, kTileBytes);
            tma_load_2d(base + kSmemQ, &tmap_q, a.q_cols_per_head * head,
                        a.q_rows_per_batch * b + a.q_rows_per_head * head + q0, bar_q);
            for (int j = 0; j < n_kt; ++j) {
                const int s = j % kStagesKV;
                mbar_wait(bar_kv_empty + s * 8, ((j / kStagesKV) & 1) ^ 1);
                mbar_expect_tx(bar_kv_full + s * 8, 2 * kTileBytes);
                const uint32_t dst = base + kSmemKV + s * 2 * kTileBytes;
                tma_load_2d(dst, &tmap_k, a.k_cols_per_head * head,
                            a.k_rows_per_batch * b + a.k_rows_per_head * head + j * kKT, bar_kv_full + s * 8);
                tma_load_2d(dst + kTileBytes, &tmap_v, a.v_cols_per_head * head,
                            a.v_rows_per_batch * b + a.v_rows_per_head * head + j * kKT, bar_kv_full + s * 8);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_kt > 0) {
            constexpr uint32_t idesc_s = tc_idesc(kKT, false, false);   // 128 x 128, both K-major
            constexpr uint32_t idesc_o = tc_idesc(kDKV, false, true);   // 128 x 64, B = V MN-major
            const uint64_t dq = tc_smem_desc(base + kSmemQ);
            auto issue_s = [&](int j) {
                const int s = j % kStagesKV, buf = j & 1;
                mbar_wait(bar_kv_full + s * 8, (j / kStagesKV) & 1);
                mbar_wait(bar_s_free + buf * 8, ((j >> 1) & 1) ^ 1);    // softmax has read S[buf] of tile j - 2
                tc_fence_after();
                const uint64_t dk = tc_smem_desc(base + kSmemKV + s * 2 * kTileBytes);
#pragma unroll
                for (int k = 0; k < kDKV / 16; ++k)
                    tc_mma_f16(tmem_base + (uint32_t)(buf * kKT), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
                tc_commit(bar_s_full + buf * 8);
            };
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < n_kt; ++j) {
                if (j + 1 < n_kt) issue_s(j + 1);
                const int s = j % kStagesKV;
                mbar_wait(bar_p_full, j & 1);                           // P_j is in shared memory
                tc_fence_after();
                // A = P: two K-major 128 x 64 sub-tiles (keys 0-63, 64-127); B = V tile, MN-major: 16 keys
                // per k step = two 1024-B swizzle atoms
                const uint32_t vaddr = base + kSmemKV + s * 2 * kTileBytes + kTileBytes;
                const uint64_t dv = tc_smem_desc_mn(vaddr, kTileBytes);
#pragma unroll
                for (int k = 0; k < kKT / 16; ++k) {
                    const uint64_t dp = tc_smem_desc(base + kSmemP + (k >> 2) * kTileBytes) + (uint64_t)(2 * (k & 3));
                    tc_mma_f16(tmem_base + kTmemO, dp, dv + (uint64_t)(128 * k), idesc_o, k != 0);
                }
                tc_commit(bar_o_full);                 // O_j complete
                tc_commit(bar_p_free);                 // ... and P may be overwritten
                tc_commit(bar_kv_empty + s * 8);       // ... and the K/V stage refilled
            }
        }
    } else {
        // softmax / epilogue: thread = query row (TMEM lane quarter = warp % 4)
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;             // row inside the tile = TMEM lane
        const int row = q0 + r;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        const float kLog2e = 1.4426950408889634f;
        const unsigned long long bh = (unsigned long long)(b * kHeads + head);
        const unsigned long long row_index0 = (bh * p.Tq + (unsigned long long)min(row, p.Tq - 1)) * p.Tk;
        const bool group_aligned = (row_index0 & 3ull) == 0ull;
        const int keep_words = ((p.Tk + 63) / 64) * 4;
        float o[kDKV];
#pragma unroll
        for (int d = 0; d < kDKV; ++d) o[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kt; ++j) {
            const int buf = j & 1;
            mbar_wait(bar_s_full + buf * 8, (j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tc_ld_32x32(tmem_base + tlane + (uint32_t)(buf * kKT + c * 32), sv[c]);
            tc_fence_before();
            mbar_arrive(bar_s_free + buf * 8);         // S[buf] may be overwritten by tile j + 2
            const int key0 = j * kKT;
            const bool need_mask = (key0 + kKT > p.Tk) || (p.causal && (key0 + kKT - 1 > q0 + p.causal_offset));
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    float x = __uint_as_float(sv[c][e]);
                    if (need_mask) {
                        const int key = key0 + c * 32 + e;
                        if (!(key < p.Tk && (!p.causal || key <= row + p.causal_offset))) x = -INFINITY;
                        sv[c][e] = __float_as_uint(x);
                    }
                    mx = fmaxf(mx, x);
                }
            const float m_new = fmaxf(m_run, mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float scale_old = exp2f((m_run - m_use) * kLog2e);   // m_run = -inf -> 0
            m_run = m_new;
            l_run *= scale_old;
            // P_j -> shared memory (SWIZZLE_128B, K-major: chunk ch of row r at ((ch & 7) ^ (r & 7)) * 16
            // inside sub-tile ch >> 3), once the previous P V has finished reading the tile
            if (j > 0) mbar_wait(bar_p_free, (j - 1) & 1);
            unsigned long long keep_bits[2] = {0ull, 0ull};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {       // 8 keys = one 16-byte chunk
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        pv[e] = exp2f((__uint_as_float(sv[c][g8 * 8 + e]) - m_use) * kLog2e);
                        l_run += pv[e];                // the normaliser sums the undropped weights
                    }
                    if (p.drop.on()) {
                        const int key = key0 + c * 32 + g8 * 8;
                        float f[8];
                        if (group_aligned) {
                            drop_factor4(p.drop, (row_index0 + key) >> 2, *reinterpret_cast<float(*)[4]>(&f[0]));
                            drop_factor4(p.drop, ((row_index0 + key) >> 2) + 1, *reinterpret_cast<float(*)[4]>(&f[4]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = drop_factor(p.drop, row_index0 + key + e);
                        }
                        // keep bits in the backward's layout: 64-key tile kt, word jq = (key % 8) / 2,
                        // bit 2 * ((key % 64) / 8) + key % 2; here half = c >> 1, ni = (c & 1) * 4 + g8
                        const int ni = (c & 1) * 4 + g8;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            pv[e] *= f[e];
                            if (f[e] != 0.f) keep_bits[c >> 1] |= 1ull << ((e >> 1) * 16 + 2 * ni + (e & 1));
                        }
                    }
                    const int ch = c * 4 + g8;         // 16-byte chunk of the 256-byte row
                    const uint32_t dst = base + kSmemP + (ch >> 3) * kTileBytes + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "r"(pack_bf16(pv[0], pv[1])),
                                 "r"(pack_bf16(pv[2], pv[3])), "r"(pack_bf16(pv[4], pv[5])), "r"(pack_bf16(pv[6], pv[7]))
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_arrive(bar_p_full);
            if (p.drop.on() && p.keep && row < p.Tq) {
                // keep_bits[h] packs the four 16-bit words of 64-key tile 2 j + h (word jq in bits 16 jq ..)
                unsigned long long* w = reinterpret_cast<unsigned long long*>(
                    p.keep + ((size_t)(bh * p.Tq + row)) * keep_words + (size_t)(2 * j) * 4);
                w[0] = keep_bits[0];
                if ((2 * j + 1) * 64 < p.Tk) w[1] = keep_bits[1];
            }
            // O_j = P_j V_j from TMEM into the running row
            mbar_wait(bar_o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ov[32];
                tc_ld_32x32(tmem_base + tlane + (uint32_t)(kTmemO + c * 32), ov);
#pragma unroll
                for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * scale_old + __uint_as_float(ov[e]);
            }
            tc_fence_before();
        }
        if (row < p.Tq) {
            if (p.lse2) p.lse2[(size_t)bh * p.Tq + row] = m_run * kLog2e + log2f(l_run);
            const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
            bf16* O = p.O + (size_t)b * p.o_batch_stride + head * p.o_head_stride + (size_t)row * p.o_row_stride;
#pragma unroll
            for (int d = 0; d < kDKV; d += 8) {
                uint4 v = make_uint4(pack_bf16(o[d] * inv, o[d + 1] * inv), pack_bf16(o[d + 2] * inv, o[d + 3] * inv),
                                     pack_bf16(o[d + 4] * inv, o[d + 5] * inv), pack_bf16(o[d + 6] * inv, o[d + 7] * inv));
                *reinterpret_cast<uint4*>(O + d) = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(kTmemCols)
                     : "memory");
    }
}

// (rows_per_batch, rows_per_head, cols_per_head) of a strided (b, head, row, 64) view, and the 2-D tensor
// (rows, cols) with pitch row_stride it is a window of
bool tc_view(long batch_stride, long head_stride, int row_stride, int batch, int T, int* rows_per_batch,
             int* rows_per_head, int* cols_per_head, long* rows, int* cols) {
    if (row_stride <= 0 || batch_stride % row_stride != 0) return false;
    *rows_per_batch = (int)(batch_stride / row_stride);
    if (head_stride < row_stride) {                     // heads side by side inside a row
        *rows_per_head = 0;
        *cols_per_head = (int)head_stride;
        *cols = (int)head_stride * kHeads;
        if (*cols > row_stride) return false;
    } else {                                            // one block of rows per head
        if (head_stride % row_stride != 0) return false;
        *rows_per_head = (int)(head_stride / row_stride);
        *cols_per_head = 0;
        *cols = kDKV;
    }
    *rows = (long)*rows_per_batch * (batch - 1) + (long)*rows_per_head * (kHeads - 1) + T;
    return true;
}

}  // namespace

Status launch_attn_full_tc(TmaCache& tc, const AttnFullParams& p, int batch, cudaStream_t stream) {
    if (batch <= 0 || p.Tq <= 0) return OkStatus();
    AttnFullTcArgs a{};
    a.p = p;
    long q_rows, k_rows, v_rows;
    int q_cols, k_cols, v_cols;
    if (!tc_view(p.q_batch_stride, p.q_head_stride, p.q_row_stride, batch, p.Tq, &a.q_rows_per_batch, &a.q_rows_per_head,
                 &a.q_cols_per_head, &q_rows, &q_cols) ||
        !tc_view(p.k_batch_stride, p.k_head_stride, p.k_row_stride, batch, p.Tk, &a.k_rows_per_batch, &a.k_rows_per_head,
                 &a.k_cols_per_head, &k_rows, &k_cols) ||
        !tc_view(p.v_batch_stride, p.v_head_stride, p.v_row_stride, batch, p.Tk, &a.v_rows_per_batch, &a.v_rows_per_head,
                 &a.v_cols_per_head, &v_rows, &v_cols))
        return Error(2, "attn_full_tc: strides do not form a 2-D tensor view");
    if (p.o_row_stride % 8 != 0 || p.o_head_stride % 8 != 0 || p.o_batch_stride % 8 != 0)
        return Error(2, "attn_full_tc: output rows must be 16-byte aligned");
    const CUtensorMap* m = nullptr;
    MRMT3_TRY(tc.get(p.Q, q_rows, q_cols, p.q_row_stride, kQT, &m));
    const CUtensorMap mq = *m;
    MRMT3_TRY(tc.get(p.K, k_rows, k_cols, p.k_row_stride, kKT, &m));
    const CUtensorMap mk = *m;
    MRMT3_TRY(tc.get(p.V, v_rows, v_cols, p.v_row_stride, kKT, &m));
    const CUtensorMap mv = *m;
    MRMT3_TRY(ensure_dynamic_smem(attn_full_tc_kernel, kSmemTotal));
    dim3 grid(ceil_div(p.Tq, kQT), kHeads, batch);
    attn_full_tc_kernel<<<grid, 192, kSmemTotal, stream>>>(mq, mk, mv, a);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

constexpr int kAttnFullTcDefault = 0;
static int g_attn_full_tc = [] {
    const char* e = getenv("MRMT3_ATTN_FULL_TC");
    return e ? atoi(e) : kAttnFullTcDefault;
}();

void attn_full_configure(int use_tc) { g_attn_full_tc = use_tc < 0 ? kAttnFullTcDefault : use_tc; }

Status launch_attn_full_auto(TmaCache& tc, const AttnFullParams& p, int batch, cudaStream_t stream) {
    if (g_attn_full_tc && p.Tq >= kQT) return launch_attn_full_tc(tc, p, batch, stream);
    return launch_attn_full(p, batch, stream);
}

}  // namespace mrmt3
