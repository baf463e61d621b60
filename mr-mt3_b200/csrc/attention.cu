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

#include <cuda.h>

#include <algorithm>

namespace mrmt3 {

// =============================================================================================
// full attention
constexpr int kAttnBQ = 64;
constexpr int kAttnBK = 64;

template <int CTAS>  // resident CTAs per SM the register budget is cut for (3: 168 registers, 4: 128, no spills)
__global__ void __launch_bounds__(128, CTAS)
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
        uint32_t keep_lo = 0, keep_hi = 0;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            float p0 = fast_exp2((s[ni][0] - m_use[0]) * kLog2e);
            float p1 = fast_exp2((s[ni][1] - m_use[0]) * kLog2e);
            float p2 = fast_exp2((s[ni][2] - m_use[1]) * kLog2e);
            float p3 = fast_exp2((s[ni][3] - m_use[1]) * kLog2e);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            if (p.drop.on()) {  // dropout(softmax): the normaliser above sums the undropped weights
                const unsigned long long r0i = ((unsigned long long)(b * kHeads + head) * p.Tq + row_lo) * p.Tk;
                const unsigned long long r1i = r0i + 8ull * p.Tk;
                const int key = kt * kAttnBK + ni * 8 + (lane & 3) * 2;
                float f0, f1, f2, f3;
                bool k0, k1, k2, k3;   // keep decisions (the bits the backward reads)
                if ((p.Tk & 3) == 0) {
                    // One hash covers four consecutive keys of a row; a lane pair (lane ^ 1) holds exactly
                    // those four keys for two rows.  The even lane hashes the group of row_lo, the odd lane
                    // that of row_lo + 8, and each hands the partner the 32 bits (two draws) it needs:
                    // 8 hashes per tile and thread instead of 16, same mask bit for bit.
                    const bool odd = lane & 1;
                    const unsigned long long h = drop_hash4(p.drop, ((odd ? r1i : r0i) + (unsigned long long)(key & ~3)) >> 2);
                    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
                    const uint32_t d0 = odd ? got : lo;   // draws of this lane's two keys in row_lo
                    const uint32_t d1 = odd ? hi : got;   // ... in row_lo + 8
                    k0 = (d0 & 0xFFFFu) >= p.drop.threshold;
                    k1 = (d0 >> 16) >= p.drop.threshold;
                    k2 = (d1 & 0xFFFFu) >= p.drop.threshold;
                    k3 = (d1 >> 16) >= p.drop.threshold;
                    f0 = k0 ? p.drop.scale : 0.f;
                    f1 = k1 ? p.drop.scale : 0.f;
                    f2 = k2 ? p.drop.scale : 0.f;
                    f3 = k3 ? p.drop.scale : 0.f;
                } else {
                    drop_factor2(p.drop, r0i + key, f0, f1);
                    drop_factor2(p.drop, r1i + key, f2, f3);
                    k0 = f0 != 0.f;
                    k1 = f1 != 0.f;
                    k2 = f2 != 0.f;
                    k3 = f3 != 0.f;
                }
                p0 *= f0;
                p1 *= f1;
                p2 *= f2;
                p3 *= f3;
                keep_lo |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (2 * ni);
                keep_hi |= ((k2 ? 1u : 0u) | (k3 ? 2u : 0u)) << (2 * ni);
            }
            // C-fragment of two adjacent n-blocks == A-fragment of one k16 step
            pf[ni >> 1][(ni & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[ni >> 1][(ni & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
        // rescale the running output only when some row of the warp saw a new maximum (after the first
        // tiles of a row that is the exception: 32 multiplies per tile saved, same arithmetic otherwise)
        if (__any_sync(0xffffffffu, scale_old[0] != 1.f || scale_old[1] != 1.f)) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                for (int r = 0; r < 4; ++r) o[ni][r] *= scale_old[r >> 1];
        }

        if (p.drop.on() && p.keep) {
            const int n_words = ((p.Tk + kAttnBK - 1) / kAttnBK) * 4;
            unsigned short* w = p.keep + ((size_t)(b * kHeads + head) * p.Tq + row_lo) * n_words + kt * 4 + (lane & 3);
            if (row_lo < p.Tq) w[0] = (unsigned short)keep_lo;
            if (row_lo + 8 < p.Tq) w[(size_t)8 * n_words] = (unsigned short)keep_hi;
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
        if (p.lse2 && (lane & 3) == 0 && row_lo + h * 8 < p.Tq)
            p.lse2[((size_t)b * kHeads + head) * p.Tq + row_lo + h * 8] = m_run[h] * kLog2e + log2f(l);
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
    // MRMT3_ATTN_FWD_CTAS=3 for A/B runs against the 4-CTA register budget
    static const int ctas = [] {
        const char* e = getenv("MRMT3_ATTN_FWD_CTAS");
        return e && atoi(e) == 3 ? 3 : 4;
    }();
    if (ctas == 4) attn_full_kernel<4><<<grid, 128, 0, stream>>>(p);
    else attn_full_kernel<3><<<grid, 128, 0, stream>>>(p);
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

// =============================================================================================
// decode-step attention, persistent TMA-ring + tensor-core variant (the default)
//
// The per-item kernel above is limited by wave quantisation (1536 CTAs in 1.73 waves of 888
// register-limited slots), by the ramp of every CTA, and by the CUDA-core instruction stream
// (bf16 -> fp32 unpacking plus one FMA per cache element).  This kernel removes all three:
//
//   * persistent grid: `ctas_per_sm` CTAs per SM walk the work items (lane, head) round-robin;
//   * a producer warp streams K and V in 64-key chunks (two 32-row TMA boxes each, 128-byte
//     swizzle, L2 evict-first) into a Q x S-stage shared-memory ring guarded by full/empty
//     mbarriers.  It runs ahead across item boundaries, so Q S x 16 KB per CTA are in flight at
//     all times whatever the register pressure or occupancy of the math warps;
//   * Q = 1 or 2 quartets of math warps work on Q chunks at a time (chunk c of an item goes to
//     quartet c % Q, which has its own S-stage sub-ring, so every warp observes every phase of the
//     stage barriers it waits on).  Inside a quartet warp w owns keys [16w, 16w + 16) of the chunk and
//     computes s = q K^T and o += p V with mma.sync m16n8k16 (row 0 of the 16-row A tile carries
//     the single query, the FlashAttention register trick turns the score fragment into the P
//     fragment), fp32 online softmax in registers: ~250 warp instructions per 16 KB chunk
//     instead of ~800, and two warps per scheduler hide the ldmatrix -> mma -> shuffle chain;
//   * the eight partial (m, l, o) states of an item are merged in fixed warp order through shared
//     memory -- the result of a (lane, head) does not depend on what else is in the batch.
//
// The step's own K/V row (self-attention) is patched into the shared-memory tile from the fused
// QKV row (the TMA box that covers it reads whatever the page holds at that row) and appended to
// the cache page with plain stores.  Rows of a box past the valid keys are masked to p = 0; they
// are cache memory that is zero-initialised at allocation and only ever holds finite values, so
// 0 * v cannot produce a NaN.
constexpr int kMmaChunk = 64;     // keys per ring stage
constexpr int kMmaBox = 32;       // rows per TMA box
constexpr int kMmaMaxQuartets = 2;  // warp quartets (template parameter Q), each working on its own chunk
constexpr int kMmaTileBytes = kMmaChunk * kDKV * (int)sizeof(bf16);  // 8 KB, K or V
constexpr int kMmaStageBytes = 2 * kMmaTileBytes;
constexpr int kMmaPartFloats = 2 + kDKV;                              // m, l, o[64]
constexpr int kMmaMergeBytes = 2 * 4 * kMmaMaxQuartets * kMmaPartFloats * (int)sizeof(float);

__device__ __forceinline__ void tma_box_load(uint32_t smem_dst, const CUtensorMap* map, int row, uint32_t bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(smem_dst),
        "l"(map), "r"(0), "r"(row), "r"(bar), "l"(policy)
        : "memory");
}

struct MmaItem {
    int n_keys, n_chunks;
    long long k_row, v_row;  // tensor-map row of key 0 (cross) / of row 0 inside page 0 (paged)
    const int* pages;
};

template <bool PAGED>
__device__ __forceinline__ MmaItem mma_item(const AttnDecodeParams& p, int lane_id, int head, int pos) {
    MmaItem it;
    if (PAGED) {
        it.n_keys = pos + 1;
        it.pages = p.block_table + (size_t)lane_id * p.max_pages;
        it.k_row = ((long long)(p.layer * 2 + 0) * kHeads + head) * kKVPage;
        it.v_row = ((long long)(p.layer * 2 + 1) * kHeads + head) * kKVPage;
    } else {
        it.n_keys = p.n_keys_ptr ? p.n_keys_ptr[lane_id] : p.n_keys;
        it.pages = nullptr;
        it.k_row = p.tmap_row0 + ((((long long)lane_id * p.n_layers + p.layer) * 2 + 0) * kHeads + head) * p.tk_cap;
        it.v_row = it.k_row + (long long)kHeads * p.tk_cap;
    }
    it.n_chunks = (it.n_keys + kMmaChunk - 1) / kMmaChunk;
    return it;
}

template <bool PAGED, int S, int Q>
__global__ void __launch_bounds__((4 * Q + 1) * 32)
    attn_decode_mma_kernel(const __grid_constant__ CUtensorMap tmap, AttnDecodeParams p, int n_lanes) {
    extern __shared__ unsigned char mma_smem_raw[];
    const uint32_t raw = smem_u32(mma_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-B alignment
    unsigned char* base_ptr = mma_smem_raw + (base - raw);
    constexpr int kMmaQuartets = Q, kMmaWarps = 4 * Q;
    constexpr int NS = S * kMmaQuartets;  // S stages per quartet
    float* merge = reinterpret_cast<float*>(base_ptr + NS * kMmaStageBytes);
    const uint32_t full0 = base + NS * kMmaStageBytes + kMmaMergeBytes;
    const uint32_t empty0 = full0 + 8 * NS;

    trace_begin(p.trace);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 4);
        }
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    const int n_items = n_lanes * kHeads;
    // `active`, `step` and the block table were written by kernels that completed before this one
    // was launched (see attn_decode_kernel), so both roles may read them ahead of the PDL wait.
    const int pos = PAGED ? p.step_ptr[0] + p.pos_offset : 0;
    const long long rows_per_page = (long long)(p.page_stride / kDKV);
    // work units: (item, part); all items of a launch have the same key count (self: pos + 1; cross:
    // n_keys -- the host disables the split when key counts are per lane)
    const int part_chunks = p.part_keys / kMmaChunk;             // 0: one unit per item
    const int launch_keys = PAGED ? pos + 1 : p.n_keys;
    const int n_parts = part_chunks ? max(1, (launch_keys + p.part_keys - 1) / p.part_keys) : 1;
    const int n_units = n_items * n_parts;

    if (warp == kMmaWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            int cnt0 = 0, cnt1 = 0;  // chunks handed to each quartet so far
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const int item = unit / n_parts, part = unit - item * n_parts;
                const int lane_id = item / kHeads, head = item - lane_id * kHeads;
                if (p.active && !p.active[lane_id]) continue;
                const MmaItem it = mma_item<PAGED>(p, lane_id, head, pos);
                const int c_begin = part * part_chunks;
                const int c_end = part_chunks ? min(it.n_chunks, c_begin + part_chunks) : it.n_chunks;
                for (int c = c_begin; c < c_end; ++c) {
                    const int qt = c & (kMmaQuartets - 1);
                    const int n = qt ? cnt1++ : cnt0++;
                    const int s = qt * S + n % S;
                    mbar_wait(empty0 + 8 * s, ((n / S) & 1) ^ 1);
                    const int first = c * kMmaChunk;
                    const int n_box = (min(kMmaChunk, it.n_keys - first) + kMmaBox - 1) / kMmaBox;
                    long long kr = it.k_row + first, vr = it.v_row + first;
                    if (PAGED) {
                        const long long pg = (long long)it.pages[first / kKVPage] * rows_per_page - (first / kKVPage) * kKVPage;
                        kr += pg;
                        vr += pg;
                    }
                    const uint32_t dst = base + s * kMmaStageBytes;
                    const uint32_t bar = full0 + 8 * s;
                    mbar_expect_tx(bar, (uint32_t)n_box * 2 * kMmaBox * kDKV * sizeof(bf16));
                    for (int b = 0; b < n_box; ++b) {
                        tma_box_load(dst + b * kMmaBox * kDKV * 2, &tmap, (int)(kr + b * kMmaBox), bar, policy);
                        tma_box_load(dst + kMmaTileBytes + b * kMmaBox * kDKV * 2, &tmap, (int)(vr + b * kMmaBox), bar, policy);
                    }
                }
            }
        }
        return;
    }

    // ---------------- math warps ----------------
    pdl_wait();  // the producer kernel's q (and k, v) rows are complete and visible from here on
    trace_mark(p.trace, 0);
    pdl_launch_dependents();
    const float kLog2e = 1.4426950408889634f;
    const int quad = lane & 3;

    int cnt = 0, buf = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int item = unit / n_parts, part = unit - item * n_parts;
        const int lane_id = item / kHeads, head = item - lane_id * kHeads;
        if (p.active && !p.active[lane_id]) continue;
        const MmaItem it = mma_item<PAGED>(p, lane_id, head, pos);
        const int c_begin = part * part_chunks;   // part_chunks is even: chunk parity == ring-quartet parity
        const int c_end = part_chunks ? min(it.n_chunks, c_begin + part_chunks) : it.n_chunks;

        // A fragments of q: row 0 of the 16 x 64 tile is the query, rows 1..15 are zero
        const bf16* qrow = p.q + (size_t)lane_id * p.q_stride + head * kDKV;
        uint32_t qf[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            qf[kk][0] = lane < 4 ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + quad * 2) : 0u;
            qf[kk][1] = 0u;
            qf[kk][2] = lane < 4 ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + 8 + quad * 2) : 0u;
            qf[kk][3] = 0u;
        }

        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int r = 0; r < 4; ++r) o[i][r] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;

        const int qt = warp >> 2;          // this warp's quartet: chunks with (c % 2) == qt
        const int kb = (warp & 3) * 16;    // its 16 keys inside every such chunk
        for (int c = c_begin + qt; c < c_end; c += kMmaQuartets, ++cnt) {
            const int s = qt * S + cnt % S;
            mbar_wait(full0 + 8 * s, (cnt / S) & 1);
            const int k0 = c * kMmaChunk;
            const int n_valid = min(kMmaChunk, it.n_keys - k0);
            if (kb < n_valid) {
                const uint32_t bk = base + s * kMmaStageBytes;
                const uint32_t bv = bk + kMmaTileBytes;
                if (PAGED && pos >= k0 + kb && pos < k0 + kb + 16) {
                    // the step's K and V sit in the fused QKV row right after Q: patch them into
                    // the tile and append them to the cache page for the steps to come
                    const int r = pos - k0;
                    if (lane < 16) {
                        const int kv = lane >> 3, ch = lane & 7;
                        const uint4 val = *reinterpret_cast<const uint4*>(qrow + kInner * (1 + kv) + ch * 8);
                        unsigned char* tile = base_ptr + s * kMmaStageBytes + kv * kMmaTileBytes;
                        *reinterpret_cast<uint4*>(tile + r * 128 + ((ch ^ (r & 7)) << 4)) = val;
                        const size_t off = ((size_t)it.pages[pos / kKVPage] * rows_per_page +
                                            (size_t)(kv ? it.v_row : it.k_row) + (pos % kKVPage)) * kDKV + ch * 8;
                        *reinterpret_cast<uint4*>(const_cast<bf16*>(p.kv_pool) + off) = val;
                    }
                    __syncwarp();
                }
                // all eight fragment loads of this warp's 16 keys up front: the V loads do not
                // depend on the softmax and overlap the score mma chain
                uint32_t kf[4][4], vf[4][4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    int row = kb + (lane & 7) + ((lane >> 4) << 3);
                    int ch = kk * 2 + ((lane >> 3) & 1);
                    ldmatrix_x4(kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3], bk + row * 128 + ((ch ^ (row & 7)) << 4));
                }
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    int row = kb + (lane & 7) + (((lane >> 3) & 1) << 3);
                    int ch = nj * 2 + (lane >> 4);
                    ldmatrix_x4_trans(vf[nj][0], vf[nj][1], vf[nj][2], vf[nj][3],
                                      bv + row * 128 + ((ch ^ (row & 7)) << 4));
                }
                // s = q K^T for this warp's 16 keys (row 0 of two 16 x 8 blocks); two independent
                // accumulator pairs (dims 0-15|32-47 and 16-31|48-63) halve the dependent mma chain
                float sc[2][4], sd[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int r = 0; r < 4; ++r) sc[i][r] = sd[i][r] = 0.f;
#pragma unroll
                for (int kk = 0; kk < 4; kk += 2) {
                    mma_bf16_16816(sc[0], qf[kk], kf[kk][0], kf[kk][1]);
                    mma_bf16_16816(sc[1], qf[kk], kf[kk][2], kf[kk][3]);
                    mma_bf16_16816(sd[0], qf[kk + 1], kf[kk + 1][0], kf[kk + 1][1]);
                    mma_bf16_16816(sd[1], qf[kk + 1], kf[kk + 1][2], kf[kk + 1][3]);
                }
                // mask + online softmax on row 0 (registers [0], [1]; the row lives in lanes 0..3,
                // the other lanes carry the all-zero query rows and are never read)
                float mx = -INFINITY;
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int key = kb + ni * 8 + quad * 2;
                    sc[ni][0] = key < n_valid ? (sc[ni][0] + sd[ni][0]) * kLog2e : -INFINITY;
                    sc[ni][1] = key + 1 < n_valid ? (sc[ni][1] + sd[ni][1]) * kLog2e : -INFINITY;
                    mx = fmaxf(mx, fmaxf(sc[ni][0], sc[ni][1]));
                }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(m_run, mx);  // finite: key kb of the chunk is valid
                const float corr = exp2f(m_run - m_new);
                m_run = m_new;
                const float p00 = exp2f(sc[0][0] - m_new), p01 = exp2f(sc[0][1] - m_new);
                const float p10 = exp2f(sc[1][0] - m_new), p11 = exp2f(sc[1][1] - m_new);
                l_run = l_run * corr + ((p00 + p01) + (p10 + p11));
                uint32_t pf[4] = {pack_bf16(p00, p01), 0u, pack_bf16(p10, p11), 0u};
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) {
                    o[ni][0] *= corr;
                    o[ni][1] *= corr;
                }
                // o += p V  (V tile is [key][d]; the transposed ldmatrix gave the col-major B fragment)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    mma_bf16_16816(o[nj * 2], pf, vf[nj][0], vf[nj][1]);
                    mma_bf16_16816(o[nj * 2 + 1], pf, vf[nj][2], vf[nj][3]);
                }
            }
            // every ldmatrix of this stage has been consumed by an mma: hand the stage back
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }

        // merge the eight warps' partial states in fixed order
        if (warp == 0 && unit == (int)blockIdx.x) trace_mark(p.trace, 2);  // CTA 0: first unit's chunks consumed
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
        float* wpart = merge + (buf * kMmaWarps + warp) * kMmaPartFloats;
        if (lane < 4) {
            if (lane == 0) {
                wpart[0] = m_run;
                wpart[1] = l_run;
            }
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                *reinterpret_cast<float2*>(wpart + 2 + ni * 8 + quad * 2) = make_float2(o[ni][0], o[ni][1]);
        }
        named_bar_sync(1, kMmaWarps * 32);
        if (warp == 0) {
            const float* pb = merge + buf * kMmaWarps * kMmaPartFloats;
            float mx = -INFINITY;
#pragma unroll
            for (int w = 0; w < kMmaWarps; ++w) mx = fmaxf(mx, pb[w * kMmaPartFloats]);
            float den = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int w = 0; w < kMmaWarps; ++w) {
                const float wgt = exp2f(pb[w * kMmaPartFloats] - mx);  // idle warp: exp2(-inf) = 0
                den += pb[w * kMmaPartFloats + 1] * wgt;
                const float2 ov = *reinterpret_cast<const float2*>(pb + w * kMmaPartFloats + 2 + lane * 2);
                o0 += ov.x * wgt;
                o1 += ov.y * wgt;
            }
            bool write_out = true;
            if (n_parts > 1) {
                // publish this part's state; whoever completes the item merges the parts in part order
                // (fixed order: the result is a function of the item alone)
                float* mine = p.part_scratch + ((size_t)item * p.max_parts + part) * kMmaPartFloats;
                if (lane == 0) {
                    mine[0] = mx;
                    mine[1] = den;
                }
                *reinterpret_cast<float2*>(mine + 2 + lane * 2) = make_float2(o0, o1);
                __threadfence();
                __syncwarp();
                int ticket = 0;
                if (lane == 0) ticket = atomicAdd(p.part_counter + item, 1);
                ticket = __shfl_sync(0xffffffffu, ticket, 0);
                write_out = ticket == n_parts - 1;
                if (write_out) {
                    __threadfence();
                    const float* all = p.part_scratch + (size_t)item * p.max_parts * kMmaPartFloats;
                    float mall = -INFINITY;
                    for (int q = 0; q < n_parts; ++q) mall = fmaxf(mall, __ldcg(all + q * kMmaPartFloats));
                    den = 0.f, o0 = 0.f, o1 = 0.f;
                    for (int q = 0; q < n_parts; ++q) {
                        const float* pq = all + q * kMmaPartFloats;
                        const float wgt = exp2f(__ldcg(pq) - mall);
                        den += __ldcg(pq + 1) * wgt;
                        const float2 ov = __ldcg(reinterpret_cast<const float2*>(pq + 2 + lane * 2));
                        o0 += ov.x * wgt;
                        o1 += ov.y * wgt;
                    }
                    if (lane == 0) p.part_counter[item] = 0;  // ready for the next launch (stream-ordered)
                }
            }
            if (write_out) {
                const float inv = 1.f / den;
                *reinterpret_cast<uint32_t*>(p.out + (size_t)lane_id * p.out_stride + head * kDKV + lane * 2) =
                    pack_bf16(o0 * inv, o1 * inv);
            }
        }
        if (warp == 0 && unit == (int)blockIdx.x) trace_mark(p.trace, 3);  // ... and merged / written
        buf ^= 1;  // the other half is free again once every warp has passed the next item's barrier
    }
    trace_end(p.trace);
}

static int g_attn_variant = 1;      // 0: one CTA per (lane, head), CUDA cores; 1: TMA ring + mma.sync
static int g_ring_stages = 4;       // per warp quartet
static int g_ring_ctas_per_sm = 0;  // 0 = by launch size (see launch_ring)
static int g_ring_quartets = 1;

void attn_decode_configure(int variant, int stages, int ctas_per_sm, int quartets) {
    if (variant >= 0) g_attn_variant = variant;
    if (stages > 0) g_ring_stages = stages;
    if (ctas_per_sm >= 0) g_ring_ctas_per_sm = ctas_per_sm;
    if (quartets > 0) g_ring_quartets = quartets;
}

template <bool PAGED, int S, int Q>
static Status launch_ring(const AttnDecodeParams& p, int n_lanes, cudaStream_t stream) {
    if (!p.tmap) return Error(2, "attn_decode: the TMA variant needs a tensor map of the cache");
    auto kern = attn_decode_mma_kernel<PAGED, S, Q>;
    constexpr int smem = Q * S * kMmaStageBytes + kMmaMergeBytes + 2 * Q * S * 8 + 1024;
    static int n_sm_of[64] = {0};  // per device: SM count
    int dev = 0;
    MRMT3_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return Error(2, "device index out of range");
    MRMT3_TRY(ensure_dynamic_smem(kern, smem));
    if (!n_sm_of[dev]) MRMT3_CUDA_TRY(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
    // CTAs per SM: a launch of a 32-lane group (192 items, other groups' kernels running beside
    // it) wants one light CTA per SM so that the projections of the other groups stay resident; a
    // launch that has the GPU to itself (one group of 256 lanes = 1536 items) needs 2-3 CTAs per
    // SM to keep enough bytes in flight
    const int n_items = n_lanes * kHeads;
    const int per_sm = g_ring_ctas_per_sm > 0 ? g_ring_ctas_per_sm
                                              : std::max(1, std::min(3, n_items / (2 * n_sm_of[dev])));
    AttnDecodeParams q = p;
    if (q.part_keys > 0) {
        if (q.part_keys % (kMmaChunk * Q) != 0 || !q.part_scratch || !q.part_counter || q.max_parts < 1)
            return Error(2, "attn ring: part_keys must be a multiple of 64 x quartets, with scratch attached");
        if (!PAGED && q.n_keys_ptr) q.part_keys = 0;  // per-lane key counts: units would not be uniform
    }
    // with split keys the number of units depends on the position (a device scalar): size the grid for
    // the SMs, CTAs that find no unit exit at once
    const int max_units = q.part_keys > 0 ? n_items * q.max_parts : n_items;
    const int grid = std::min(max_units, n_sm_of[dev] * per_sm);
    MRMT3_TRY(launch_pdl(kern, dim3(grid), dim3((4 * Q + 1) * 32), smem, stream,
                         *reinterpret_cast<const CUtensorMap*>(q.tmap), q, n_lanes));
    return OkStatus();
}

Status launch_attn_decode(const AttnDecodeParams& p, int n_lanes, bool paged, cudaStream_t stream) {
    if (n_lanes <= 0) return OkStatus();
    if (g_attn_variant == 1) {
#define MRMT3_RING(S_, Q_) \
    return paged ? launch_ring<true, S_, Q_>(p, n_lanes, stream) : launch_ring<false, S_, Q_>(p, n_lanes, stream)
        switch (g_ring_stages * 10 + g_ring_quartets) {
            case 21: MRMT3_RING(2, 1);
            case 31: MRMT3_RING(3, 1);
            case 41: MRMT3_RING(4, 1);
            case 61: MRMT3_RING(6, 1);
            case 81: MRMT3_RING(8, 1);
            case 22: MRMT3_RING(2, 2);
            case 32: MRMT3_RING(3, 2);
            case 42: MRMT3_RING(4, 2);
            default: return Error(2, "attn ring: stages per quartet must be 2, 3, 4 (or 6 / 8 with one quartet)");
        }
#undef MRMT3_RING
    }
    dim3 grid(kHeads, n_lanes);
    if (paged)
        MRMT3_TRY(launch_pdl(attn_decode_kernel<true>, grid, dim3(kDecThreads), 0, stream, p));
    else
        MRMT3_TRY(launch_pdl(attn_decode_kernel<false>, grid, dim3(kDecThreads), 0, stream, p));
    return OkStatus();
}

int attn_decode_max_keys() { return 1 << 20; }

}  // namespace mrmt3
