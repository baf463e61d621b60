// Attention kernels of the path.  T5 attention as the reference uses it has NO 1/sqrt(d) scale
// and NO relative-position bias (reference models/t5.py:485-490 builds every block
// with has_relative_attention_bias=False), softmax in fp32.
//
//  * attn_full_kernel   : flash-style tiled attention for whole sequences (encoder self-attn
//                         256x256, memory encoder 64 queries x L keys, teacher-forced decoder
//                         self (causal) and cross).  mma.sync bf16, fp32 online softmax.
//  * attn_decode_kernel : one query per (lane, head) against a K/V stream -- the HBM-bound
//                         decode-step kernel.  Self-attention reads the paged KV cache and
//                         appends the step's K/V; cross-attention reads the static cross cache.
#include "attention.cuh"

namespace mrmt3 {

// =============================================================================================
// full attention
constexpr int kAttnBQ = 64;
constexpr int kAttnBK = 64;

__global__ void __launch_bounds__(128)
    attn_full_kernel(AttnFullParams p) {
    __shared__ __align__(128) bf16 sQ[kAttnBQ * kDKV];
    __shared__ __align__(128) bf16 sK[2][kAttnBK * kDKV];
    __shared__ __align__(128) bf16 sV[2][kAttnBK * kDKV];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * kAttnBQ;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const bf16* Q = p.Q + (size_t)b * p.q_batch_stride + head * p.q_head_stride;
    const bf16* K = p.K + (size_t)b * p.k_batch_stride + head * p.k_head_stride;
    const bf16* V = p.V + (size_t)b * p.v_batch_stride + head * p.v_head_stride;

    // number of key tiles this query tile needs
    int n_kt = (p.Tk + kAttnBK - 1) / kAttnBK;
    if (p.causal) {
        int last_key = q0 + kAttnBQ - 1 + p.causal_offset;  // largest key any row may see
        int lim = last_key / kAttnBK + 1;
        if (last_key < 0) lim = 0;
        n_kt = min(n_kt, lim);
    }

    auto load_rows = [&](bf16* dst, const bf16* src, int row_stride, int r0, int rmax) {
        // 64 rows x 8 chunks of 16 B, 128 threads -> 4 chunks each
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = tid + i * 128;
            int row = c >> 3, ch = c & 7;
            bool pred = (r0 + row) < rmax;
            const bf16* g = src + (size_t)(pred ? r0 + row : 0) * row_stride + ch * 8;
            cp_async16(dst + row * kDKV + ((ch ^ (row & 7)) << 3), g, pred);
        }
    };

    load_rows(sQ, Q, p.q_row_stride, q0, p.Tq);
    if (n_kt > 0) {
        load_rows(sK[0], K, p.k_row_stride, 0, p.Tk);
        load_rows(sV[0], V, p.v_row_stride, 0, p.Tk);
    }
    cp_async_commit();

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[i][r] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    uint32_t qf[4][4];
    const float kLog2e = 1.4426950408889634f;

    const int row_lo = q0 + warp * 16 + (lane >> 2);  // this thread's rows: row_lo, row_lo + 8

    for (int kt = 0; kt < n_kt; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < n_kt) {
            load_rows(sK[st ^ 1], K, p.k_row_stride, (kt + 1) * kAttnBK, p.Tk);
            load_rows(sV[st ^ 1], V, p.v_row_stride, (kt + 1) * kAttnBK, p.Tk);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        if (kt == 0) {
            const uint32_t bq = smem_u32(sQ);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                int row = warp * 16 + (lane & 15);
                int ch = kk * 2 + (lane >> 4);
                ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3],
                            bq + row * 128 + ((ch ^ (row & 7)) << 4));
            }
        }

        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int r = 0; r < 4; ++r) s[i][r] = 0.f;
        const uint32_t bk = smem_u32(sK[st]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                int row = nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = kk * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3, bk + row * 128 + ((ch ^ (row & 7)) << 4));
                mma_bf16_16816(s[nj * 2], qf[kk], b0, b1);
                mma_bf16_16816(s[nj * 2 + 1], qf[kk], b2, b3);
            }
        }

        // masking (keys past Tk; causal)
        const int key_base = kt * kAttnBK + (lane & 3) * 2;
        const bool need_mask = ((kt + 1) * kAttnBK > p.Tk) ||
                               (p.causal && ((kt + 1) * kAttnBK - 1 > q0 + p.causal_offset));
        if (need_mask) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int key = key_base + ni * 8 + (r & 1);
                    int row = row_lo + ((r >> 1) << 3);
                    bool ok = key < p.Tk && (!p.causal || key <= row + p.causal_offset);
                    if (!ok) s[ni][r] = -INFINITY;
                }
            }
        }

        // online softmax
        float scale_old[2], m_use[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) mx = fmaxf(mx, fmaxf(s[ni][h * 2], s[ni][h * 2 + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float m_new = fmaxf(m_run[h], mx);
            m_use[h] = (m_new == -INFINITY) ? 0.f : m_new;
            scale_old[h] = exp2f((m_run[h] - m_use[h]) * kLog2e);  // m_run=-inf -> 0
            m_run[h] = m_new;
            l_run[h] *= scale_old[h];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            float p0 = exp2f((s[ni][0] - m_use[0]) * kLog2e);
            float p1 = exp2f((s[ni][1] - m_use[0]) * kLog2e);
            float p2 = exp2f((s[ni][2] - m_use[1]) * kLog2e);
            float p3 = exp2f((s[ni][3] - m_use[1]) * kLog2e);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            // C-fragment of two adjacent n-blocks == A-fragment of one k16 step
            pf[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[ni >> 1][(ni & 1) * 2 + 1] = pack_bf16(p2, p3);
#pragma unroll
            for (int r = 0; r < 4; ++r) o[ni][r] *= scale_old[r >> 1];
        }

        // O += P V   (V tile is [key][d]; transposed ldmatrix gives the col-major B fragment)
        const uint32_t bv = smem_u32(sV[st]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {       // 16 keys per step
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {   // 16 d-columns per ldmatrix
                int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                int ch = nj * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(b0, b1, b2, b3, bv + row * 128 + ((ch ^ (row & 7)) << 4));
                mma_bf16_16816(o[nj * 2], pf[kk], b0, b1);
                mma_bf16_16816(o[nj * 2 + 1], pf[kk], b2, b3);
            }
        }
        __syncthreads();  // all warps done with stage st before it is refilled
    }
    cp_async_wait<0>();

    // finalize
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float l = l_run[h];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        l_run[h] = l > 0.f ? 1.f / l : 0.f;
    }
    bf16* O = p.O + (size_t)b * p.o_batch_stride + head * p.o_head_stride;
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        int col = ni * 8 + (lane & 3) * 2;
        if (row_lo < p.Tq)
            *reinterpret_cast<uint32_t*>(O + (size_t)row_lo * p.o_row_stride + col) =
                pack_bf16(o[ni][0] * l_run[0], o[ni][1] * l_run[0]);
        if (row_lo + 8 < p.Tq)
            *reinterpret_cast<uint32_t*>(O + (size_t)(row_lo + 8) * p.o_row_stride + col) =
                pack_bf16(o[ni][2] * l_run[1], o[ni][3] * l_run[1]);
    }
}

Status launch_attn_full(const AttnFullParams& p, int batch, cudaStream_t stream) {
    if (batch <= 0 || p.Tq <= 0) return OkStatus();
    dim3 grid(ceil_div(p.Tq, kAttnBQ), kHeads, batch);
    attn_full_kernel<<<grid, 128, 0, stream>>>(p);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// =============================================================================================
// decode-step attention: one query vector per (lane, head), single pass over the K/V stream
//
// HBM-bound: per launch the kernel must move every K and V row of every (lane, head) once and
// nothing else.  Thread mapping: 4 threads share one key (32 B = 16 dims each, one 256-bit
// LDG with L1 no-allocate / L2 evict-first so the stream does not push the L2-resident weights
// out), a warp covers 8 keys per load instruction, the 128-thread CTA 32 key slots; K and V of
// two key blocks are in flight per thread.  Each key slot keeps its own online-softmax state
// (running max, sum, 64-dim accumulator spread over its 4 threads); the 32 slots are merged once
// at the end through shared memory.  No score buffer, no CTA-wide barrier inside the stream.
//
// Programmatic dependent launch: the kernel is launched while its producer (the Q/QKV projection)
// is still running.  Everything that does not depend on the producer -- the first block of old
// K/V rows -- is requested BEFORE griddepcontrol.wait, so the stream is already in flight when
// the query arrives.  (A split-key variant, one CTA per 128-key page with a ticket merge, was
// measured slower: 58 % vs 75 % of HBM peak -- 32 KB per CTA does not amortise the CTA.)
constexpr int kDecThreads = 128;
constexpr int kDecSlots = kDecThreads / 4;
constexpr int kDecUnroll = 2;

struct __align__(32) Bf16x16 {
    uint32_t w[8];
};

__device__ __forceinline__ Bf16x16 ld_stream32(const void* p) {
    Bf16x16 r;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]),
                   "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

template <bool PAGED>
__device__ __forceinline__ const bf16* kv_row_ptr(const AttnDecodeParams& p, const bf16* base,
                                                   const int* pages, int head, int pos) {
    if (PAGED) {
        // page layout: [page][layer][k|v][head][kKVPage][64]; `base` already points at
        // (layer, k|v) inside page 0, pages are p.page_stride elements apart.
        int pg = pages[pos / kKVPage];
        return base + (size_t)pg * p.page_stride + ((size_t)head * kKVPage + (pos % kKVPage)) * kDKV;
    } else {
        // cross cache: [lane][layer][k|v][head][tk_cap][64]; `base` points at (lane, layer, k|v)
        return base + ((size_t)head * p.tk_cap + pos) * kDKV;
    }
}

template <bool PAGED>
__global__ void __launch_bounds__(kDecThreads)
    attn_decode_kernel(AttnDecodeParams p) {
    __shared__ float s_m[kDecSlots];
    __shared__ float s_l[kDecSlots];
    __shared__ __align__(16) float s_acc[kDecSlots][kDKV];

    trace_begin(p.trace);
    const int lane_id = blockIdx.y;  // decode lane (sequence)
    const int head = blockIdx.x;
    const int tid = threadIdx.x;
    const int slot = tid >> 2;  // key slot 0..31
    const int sub = tid & 3;    // 16-dim chunk 0..3

    // Every kernel of the chain triggers its dependents only AFTER its own wait has returned, so
    // at most the direct producer (the Q/QKV projection, which writes nothing but q|k|v) can
    // still be running here: `active`, `step`, the block table and the cache rows of earlier
    // positions are final and may be read ahead of the dependency wait.
    if (p.active && !p.active[lane_id]) return;  // finished lanes cost nothing

    int n_keys, n_old;
    const int* pages = nullptr;
    const bf16 *kbase, *vbase;
    if (PAGED) {
        const int pos = p.step_ptr[0] + p.pos_offset;  // position of the new token
        n_keys = pos + 1;
        n_old = pos;                                   // rows that exist before this step
        pages = p.block_table + (size_t)lane_id * p.max_pages;
        kbase = p.kv_pool + (size_t)(p.layer * 2 + 0) * kHeads * kKVPage * kDKV;
        vbase = p.kv_pool + (size_t)(p.layer * 2 + 1) * kHeads * kKVPage * kDKV;
    } else {
        n_keys = p.n_keys_ptr ? p.n_keys_ptr[lane_id] : p.n_keys;
        n_old = n_keys;
        size_t lane_off = ((size_t)lane_id * p.n_layers + p.layer) * 2 * kHeads * p.tk_cap * kDKV;
        kbase = p.kv_pool + lane_off;
        vbase = kbase + (size_t)kHeads * p.tk_cap * kDKV;
    }

    // first key block: request the rows that do not depend on the producer kernel
    Bf16x16 kr[kDecUnroll], vr[kDecUnroll];
#pragma unroll
    for (int u = 0; u < kDecUnroll; ++u) {
        const int key = u * kDecSlots + slot;
        if (key < n_old) {
            kr[u] = ld_stream32(kv_row_ptr<PAGED>(p, kbase, pages, head, key) + sub * 16);
            vr[u] = ld_stream32(kv_row_ptr<PAGED>(p, vbase, pages, head, key) + sub * 16);
        }
    }

    pdl_wait();  // the producer's q (and k, v) rows are complete and visible from here on
    pdl_launch_dependents();

    if (PAGED) {
        // append this step's K and V (they sit in the fused QKV row right after Q)
        const int pos = n_keys - 1;
        if (tid < 16) {
            const bf16* src = p.q + (size_t)lane_id * p.q_stride + kInner * (1 + (tid >> 3)) +
                              head * kDKV + (tid & 7) * 8;
            bf16* dst = const_cast<bf16*>(kv_row_ptr<true>(p, (tid >> 3) ? vbase : kbase, pages, head, pos)) +
                        (tid & 7) * 8;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        }
        __syncthreads();  // the appended row is read back below by other threads of this CTA
        if (pos < kDecSlots * kDecUnroll) {  // the new row belongs to the first key block
            const int u = pos / kDecSlots;
            if (slot == pos % kDecSlots) {
#pragma unroll
                for (int uu = 0; uu < kDecUnroll; ++uu) {
                    if (uu == u) {
                        kr[uu] = ld_stream32(kv_row_ptr<true>(p, kbase, pages, head, pos) + sub * 16);
                        vr[uu] = ld_stream32(kv_row_ptr<true>(p, vbase, pages, head, pos) + sub * 16);
                    }
                }
            }
        }
    }

    // this thread's 16 query dims, pre-multiplied by log2(e) so the softmax runs on exp2
    float qv[16];
    {
        const float kLog2e = 1.4426950408889634f;
        const uint4* qp = reinterpret_cast<const uint4*>(p.q + (size_t)lane_id * p.q_stride + head * kDKV + sub * 16);
        uint4 raw[2] = {qp[0], qp[1]};
        const bf162* h2 = reinterpret_cast<const bf162*>(raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 f = __bfloat1622float2(h2[i]);
            qv[2 * i] = f.x * kLog2e;
            qv[2 * i + 1] = f.y * kLog2e;
        }
    }

    float m_run = -INFINITY, l_run = 0.f;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;

    for (int k0 = 0; k0 < n_keys; k0 += kDecSlots * kDecUnroll) {
        if (k0 > 0) {
#pragma unroll
            for (int u = 0; u < kDecUnroll; ++u) {
                const int key = k0 + u * kDecSlots + slot;
                if (key < n_keys) {
                    kr[u] = ld_stream32(kv_row_ptr<PAGED>(p, kbase, pages, head, key) + sub * 16);
                    vr[u] = ld_stream32(kv_row_ptr<PAGED>(p, vbase, pages, head, key) + sub * 16);
                }
            }
        }
        float sc[kDecUnroll];
        float m_new = m_run;
#pragma unroll
        for (int u = 0; u < kDecUnroll; ++u) {
            const int key = k0 + u * kDecSlots + slot;
            float dot = 0.f;
            if (key < n_keys) {
                const bf162* h2 = reinterpret_cast<const bf162*>(kr[u].w);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 f = __bfloat1622float2(h2[i]);
                    dot += qv[2 * i] * f.x + qv[2 * i + 1] * f.y;
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            sc[u] = (key < n_keys) ? dot : -INFINITY;
            m_new = fmaxf(m_new, sc[u]);
        }
        if (m_new > -INFINITY) {
            const float corr = exp2f(m_run - m_new);  // m_run = -inf -> 0
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] *= corr;
#pragma unroll
            for (int u = 0; u < kDecUnroll; ++u) {
                if (sc[u] > -INFINITY) {
                    const float pk = exp2f(sc[u] - m_new);
                    l_run += pk;
                    const bf162* h2 = reinterpret_cast<const bf162*>(vr[u].w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float2 f = __bfloat1622float2(h2[i]);
                        acc[2 * i] += pk * f.x;
                        acc[2 * i + 1] += pk * f.y;
                    }
                }
            }
            m_run = m_new;
        }
    }

    // merge the 32 key slots
    if (sub == 0) {
        s_m[slot] = m_run;
        s_l[slot] = l_run;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&s_acc[slot][sub * 16 + i * 4]) =
            make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    __syncthreads();
    if (tid < kDKV / 2) {
        float mx = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < kDecSlots; ++s2) mx = fmaxf(mx, s_m[s2]);
        float den = 0.f, o0 = 0.f, o1 = 0.f;
        const int d = tid * 2;
#pragma unroll 8
        for (int s2 = 0; s2 < kDecSlots; ++s2) {
            const float w = exp2f(s_m[s2] - mx);  // empty slot: exp2(-inf) = 0
            den += s_l[s2] * w;
            o0 += s_acc[s2][d] * w;
            o1 += s_acc[s2][d + 1] * w;
        }
        const float inv = 1.f / den;
        *reinterpret_cast<uint32_t*>(p.out + (size_t)lane_id * p.out_stride + head * kDKV + d) =
            pack_bf16(o0 * inv, o1 * inv);
    }
    trace_end(p.trace);
}

Status launch_attn_decode(const AttnDecodeParams& p, int n_lanes, bool paged, cudaStream_t stream) {
    if (n_lanes <= 0) return OkStatus();
    dim3 grid(kHeads, n_lanes);
    if (paged)
        MRMT3_TRY(launch_pdl(attn_decode_kernel<true>, grid, dim3(kDecThreads), 0, stream, p));
    else
        MRMT3_TRY(launch_pdl(attn_decode_kernel<false>, grid, dim3(kDecThreads), 0, stream, p));
    return OkStatus();
}

int attn_decode_max_keys() { return 1 << 20; }

}  // namespace mrmt3
