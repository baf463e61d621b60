// Decode-step projections: C[M,N] = A[M,K] * W[N,K]^T with M = number of decode lanes (<= 512).
//
// These GEMMs are latency-bound, not tensor-bound (0.1-0.5 GFLOP, operands L2-resident): what
// matters is how many dependent memory round trips a CTA makes and how many CTAs share the work.
// So, unlike the pipelined kernel in gemm_mma.cuh, every CTA here issues the loads for its WHOLE
// K extent up front (K is a compile-time 384/512/1024: at most 128 KB of shared memory), one
// cp.async group per 64-wide k-tile, and consumes the tiles as they land: one memory latency per
// CTA instead of K/64.  Tiles are 32 x 32 or 32 x 64 so that M = 256 gives 128-256 CTAs.
//
// NORM = true fuses T5LayerNorm (reference via HF modeling_t5.py:55-70) into the GEMM: A holds
// the un-normalised bf16 copy of the residual stream, W has the norm weight folded into its
// columns (W'[n,k] = W[n,k] * g[k], packed at commit time), and because K == d_model each CTA
// sees complete rows, computes sum(x^2) from its shared A tiles and scales the accumulators by
// rsqrt(mean + eps) in the epilogue:  (x * r * g) W^T == r * (x (W g)^T).
#pragma once
#include <type_traits>

#include "gemm_mma.cuh"
#include "tma.cuh"

namespace mrmt3 {

// residual stream update that also refreshes the bf16 copy the next NORM GEMM reads
struct EpiResidualBoth {
    float* H;
    bf16* Hb;
    int ldh;
    // the residual values can be fetched before the k loop: the L2 round trip of H then overlaps
    // the operand loads instead of sitting between the last mma and the store
    static constexpr bool kPrefetch = true;
    __device__ __forceinline__ float2 load(int row, int col) const {
        return *reinterpret_cast<const float2*>(H + (size_t)row * ldh + col);
    }
    __device__ __forceinline__ void store(int row, int col, float2 h, float v0, float v1) const {
        size_t o = (size_t)row * ldh + col;
        h.x += v0;
        h.y += v1;
        *reinterpret_cast<float2*>(H + o) = h;
        *reinterpret_cast<uint32_t*>(Hb + o) = pack_bf16(h.x, h.y);
    }
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        size_t o = (size_t)row * ldh + col;
        float2* p = reinterpret_cast<float2*>(H + o);
        float2 h = *p;
        h.x += v0;
        h.y += v1;
        *p = h;
        *reinterpret_cast<uint32_t*>(Hb + o) = pack_bf16(h.x, h.y);
    }
};

template <class Epi, class = void>
struct EpiPrefetches : std::false_type {};
template <class Epi>
struct EpiPrefetches<Epi, std::enable_if_t<Epi::kPrefetch>> : std::true_type {};

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        case 7: cp_async_wait<7>(); break;
        case 8: cp_async_wait<8>(); break;
        case 9: cp_async_wait<9>(); break;
        case 10: cp_async_wait<10>(); break;
        case 11: cp_async_wait<11>(); break;
        case 12: cp_async_wait<12>(); break;
        case 13: cp_async_wait<13>(); break;
        case 14: cp_async_wait<14>(); break;
        default: cp_async_wait<15>(); break;
    }
}

template <int BN, int K, bool NORM, class Epi>
__global__ void __launch_bounds__(128)
    gemm_skinny_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, int M,
                       float eps, Epi epi, TraceSlot trace, const bf16* __restrict__ A, int lda, int a_mode) {
    constexpr int BM = 32;
    trace_begin(trace);
    constexpr int KT = K / 64;
    constexpr int NI = BN / 16;  // n-blocks of 8 per warp (warp tile 16 x BN/2)
    static_assert(KT <= 16, "K too large for the single-shot pipeline");
    extern __shared__ unsigned char smem_raw[];
    // 128-byte-swizzled TMA tiles need 1024-byte alignment
    unsigned char* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    bf16* sA = reinterpret_cast<bf16*>(smem_al);    // [KT][BM*64]
    bf16* sW = sA + KT * BM * 64;                   // [KT][BN*64]
    const uint32_t bars = smem_u32(sW + KT * BN * 64);  // KT mbarriers, one per k-tile
    __shared__ float s_scale[BM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp >> 1) * 16;
    const int wn0 = (warp & 1) * (BN / 2);
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    // Every operand byte of this CTA is requested up front with TMA box loads (one 64-wide k-tile
    // of W and of A per mbarrier): a handful of bulk requests instead of ~50 cp.async per thread,
    // whose per-SM request tracking limits a CTA to ~30 GB/s (measured: 2-2.5 us for 100 KB; see
    // DESIGN.md).  The weight tiles do not depend on the producer kernel, so they are requested
    // before the programmatic-dependency wait and their latency overlaps the producer's tail; the
    // activation tiles follow the wait.  Rows past M are zero-filled by the tensor map.
    if (tid == 0) {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) mbar_init(bars + 8 * kt, 2);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            mbar_expect_tx(bars + 8 * kt, BN * 128);
            tma_load_2d(smem_u32(sW + kt * BN * 64), &tm_w, kt * 64, n0, bars + 8 * kt);
        }
    }
    pdl_wait();
    trace_mark(trace, 0);
    pdl_launch_dependents();
    if (a_mode == 0) {
        if (tid == 0) {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                mbar_expect_tx(bars + 8 * kt, BM * 128);
                tma_load_2d(smem_u32(sA + kt * BM * 64), &tm_a, kt * 64, m0, bars + 8 * kt);
            }
        }
    } else {
        // The activation tile is the only operand on the critical path (it exists once the producer
        // kernel has finished; the weights were requested before the wait).  One thread's KT box loads
        // land one after another (~0.15 us per box, measured with the trace stamps: 1.5 us for K = 512);
        // issued as 16-byte cp.async from all 128 threads at once they cost one L2 round trip.  Rows
        // at or past M are zero-filled without touching memory.  Same 128-byte swizzle as the TMA tiles.
        constexpr int CPR = K / 8;  // 16-byte chunks per row
        for (int c = tid; c < BM * CPR; c += 128) {
            const int row = c / CPR, cc = c - row * CPR;
            const int kt = cc >> 3, ch = cc & 7;
            const bool pred = m0 + row < M;
            cp_async16(sA + kt * BM * 64 + row * 64 + ((ch ^ (row & 7)) << 3),
                       A + (size_t)(pred ? m0 + row : 0) * lda + cc * 8, pred);
        }
        cp_async_commit();
        if (tid == 0) {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) mbar_arrive(bars + 8 * kt);  // stands in for the A half of each barrier
        }
        cp_async_wait<0>();
    }
    __syncthreads();  // barrier initialisation (and the cp.async tile) visible to every waiter

    // four independent accumulator sets, one per 16-wide k step of a k-tile: the dependent
    // mma -> mma chain of an output fragment is K/64 long instead of K/16 (the chain, not the
    // tensor pipe, is what a single 4-warp CTA waits on); they are summed in fixed order at the end
    float accs[4][NI][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) accs[q][j][r] = 0.f;
    float ss = 0.f;  // NORM: partial sum of squares of row tid/4
    float2 pre[NI][2];
    if constexpr (EpiPrefetches<Epi>::value) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int row = m0 + wm0 + (lane >> 2);
            const int col = n0 + wn0 + ni * 8 + (lane & 3) * 2;
            pre[ni][0] = row < M ? epi.load(row, col) : make_float2(0.f, 0.f);
            pre[ni][1] = row + 8 < M ? epi.load(row + 8, col) : make_float2(0.f, 0.f);
        }
    }

#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        mbar_wait(bars + 8 * kt, 0);
        if (kt == 0) trace_mark(trace, 2);
        if (kt == KT / 2) trace_mark(trace, 3);
        if (kt == KT - 1) trace_mark(trace, 1);
        const uint32_t baseA = smem_u32(sA + kt * BM * 64);
        const uint32_t baseW = smem_u32(sW + kt * BN * 64);
        // all fragment loads of the k-tile first, then the math: issued back to back the ldmatrix
        // latencies overlap (interleaved with their mma, every 16-wide k step paid one in full:
        // 220-440 cycles per k-tile, measured with the in-kernel trace stamps)
        uint32_t af[4][4], wf[4][NI / 2][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            int row = wm0 + (lane & 15);
            int ch = kk * 2 + (lane >> 4);
            ldmatrix_x4(af[kk][0], af[kk][1], af[kk][2], af[kk][3], baseA + row * 128 + ((ch ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nj = 0; nj < NI / 2; ++nj) {
                int row = wn0 + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(wf[kk][nj][0], wf[kk][nj][1], wf[kk][nj][2], wf[kk][nj][3],
                            baseW + row * 128 + ((ch ^ (row & 7)) << 4));
            }
        }
        if (NORM) {
            // 4 threads per row, 2 of the 8 16-byte chunks each.  The chunks are addressed
            // LOGICALLY (un-swizzled), so the order of the fp32 sum -- and with it the result --
            // does not depend on where in the tile the row sits (rows must not depend on their
            // batch neighbours: tests/test_parity_gpu.py::test_lane_groups_do_not_change_tokens)
            const int row = tid >> 2;
            const bf16* rp = sA + kt * BM * 64 + row * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint4 raw = *reinterpret_cast<const uint4*>(rp + ((((tid & 3) * 2 + c) ^ (row & 7)) << 3));
                const bf162* h2 = reinterpret_cast<const bf162*>(&raw);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 f = __bfloat1622float2(h2[i]);
                    ss += f.x * f.x + f.y * f.y;
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nj = 0; nj < NI / 2; ++nj) {
                mma_bf16_16816(accs[kk][nj * 2], af[kk], wf[kk][nj][0], wf[kk][nj][1]);
                mma_bf16_16816(accs[kk][nj * 2 + 1], af[kk], wf[kk][nj][2], wf[kk][nj][3]);
            }
        }
    }

    float acc[NI][4];
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[j][r] = (accs[0][j][r] + accs[1][j][r]) + (accs[2][j][r] + accs[3][j][r]);

    float sc0 = 1.f, sc1 = 1.f;
    if (NORM) {
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        if ((tid & 3) == 0) s_scale[tid >> 2] = rsqrtf(ss * (1.0f / K) + eps);
        __syncthreads();
        sc0 = s_scale[wm0 + (lane >> 2)];
        sc1 = s_scale[wm0 + (lane >> 2) + 8];
    }
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
        int row = m0 + wm0 + (lane >> 2);
        int col = n0 + wn0 + ni * 8 + (lane & 3) * 2;
        if constexpr (EpiPrefetches<Epi>::value) {
            if (row < M) epi.store(row, col, pre[ni][0], acc[ni][0] * sc0, acc[ni][1] * sc0);
            if (row + 8 < M) epi.store(row + 8, col, pre[ni][1], acc[ni][2] * sc1, acc[ni][3] * sc1);
        } else {
            if (row < M) epi(row, col, acc[ni][0] * sc0, acc[ni][1] * sc0);
            if (row + 8 < M) epi(row + 8, col, acc[ni][2] * sc1, acc[ni][3] * sc1);
        }
    }
    trace_end(trace);
}

// how the activation tile reaches shared memory: 0 = TMA boxes, 1 = cp.async from all threads (default);
// MRMT3_SKINNY_A_MODE overrides (A/B measurements)
inline int skinny_a_mode() {
    static const int mode = [] {
        const char* e = getenv("MRMT3_SKINNY_A_MODE");
        return e ? atoi(e) : 1;
    }();
    return mode;
}

template <int BN, int K, bool NORM, class Epi>
Status launch_gemm_skinny(TmaCache& tc, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, float eps,
                          const Epi& epi, cudaStream_t stream, TraceSlot trace = TraceSlot{nullptr, 0}) {
    if (M <= 0) return OkStatus();
    if (N % BN != 0) return Error(2, "gemm_skinny: N must be a multiple of the column tile");
    static_assert(!NORM || K == kDModel, "fused RMSNorm needs complete rows: K == d_model");
    auto kern = gemm_skinny_kernel<BN, K, NORM, Epi>;
    constexpr int smem = (32 + BN) * K * (int)sizeof(bf16) + (K / 64) * 8 + 1024;
    MRMT3_TRY(ensure_dynamic_smem(kern, smem));
    const CUtensorMap *ma = nullptr, *mw = nullptr;
    MRMT3_TRY(tc.get(A, M, K, lda, 32, &ma));
    const CUtensorMap a_copy = *ma;  // the second lookup may rotate the cache
    MRMT3_TRY(tc.get(W, N, K, ldw, BN, &mw));
    dim3 grid(N / BN, ceil_div(M, 32));
    MRMT3_TRY(launch_pdl(kern, grid, dim3(128), smem, stream, a_copy, *mw, M, eps, epi, trace, A, lda,
                         skinny_a_mode()));
    return OkStatus();
}

}  // namespace mrmt3
