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
#include "layers.cuh"
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

// Greedy head fused into the lm_head GEMM (the logits are never materialised).  Every CTA
// reduces its 64 logit columns to one (value, index) per row; the last CTA of a 32-row tile (atomic
// ticket) reduces the per-tile candidates in column order -- the lowest index wins ties, as
// torch.argmax and the stand-alone arg-max kernel do -- and then does what argmax_advance_kernel does
// (token write-out, EOS bookkeeping of reference models/t5.py:286-291, forced tokens) and, when `emb`
// is given, the NEXT step's input row H = Emb[token] + PE[step + 1] (decode_embed_kernel); the last
// tile to finish advances the step counter.
struct EpiGreedy {
    static constexpr bool kGreedy = true;
    DecodeState st;
    float2* cand;        // (lanes of the group, n column tiles): value, index bits
    int* tile_ticket;    // one per 32-row tile of the group, zero between launches
    int vocab;
    const float* emb;    // nullptr: leave the next step's embedding to decode_embed_kernel
    const float* pe;
    float* H;
    bf16* Hb;
};
template <class Epi, class = void>
struct EpiIsGreedy : std::false_type {};
template <class Epi>
struct EpiIsGreedy<Epi, std::enable_if_t<Epi::kGreedy>> : std::true_type {};

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
                       float eps, Epi epi, TraceSlot trace, const bf16* __restrict__ A, int lda, int a_mode,
                       int k_mode) {
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

    auto do_tile = [&](int kt) __attribute__((always_inline)) {
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
    };
    if (k_mode == 1) {
        // Every tile was requested up front and they land within one L2 round trip of each other, so
        // waiting tile by tile buys no overlap -- but a barrier wait between two k-tiles (volatile asm)
        // stops the compiler from overlapping the fragment loads of tile i + 1 with the mma of tile i:
        // measured with the trace stamps, the loop then costs 0.14-0.19 us PER k-tile (1.5 us for
        // K = 512, 2.2 us for K = 1024), most of a decode-step projection.  Wait for all of them, then
        // run the K loop as one straight-line block.
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) mbar_wait(bars + 8 * kt, 0);
        trace_mark(trace, 1);
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) do_tile(kt);
        trace_mark(trace, 3);
    } else {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(bars + 8 * kt, 0);
            if (kt == 0) trace_mark(trace, 2);
            if (kt == KT / 2) trace_mark(trace, 3);
            if (kt == KT - 1) trace_mark(trace, 1);
            do_tile(kt);
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
    if constexpr (EpiIsGreedy<Epi>::value) {
        __shared__ float s_val[2][BM];
        __shared__ int s_idx[2][BM];
        __shared__ int s_tok[BM];
        __shared__ int s_last;
        // this thread's best over its columns of rows r and r + 8 (ascending columns, strict '>': the
        // lowest index wins among equals)
        float b0 = -INFINITY, b1 = -INFINITY;
        int i0 = 0x7fffffff, i1 = 0x7fffffff;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int col = n0 + wn0 + ni * 8 + (lane & 3) * 2;
            const float v00 = acc[ni][0] * sc0, v01 = acc[ni][1] * sc0, v10 = acc[ni][2] * sc1, v11 = acc[ni][3] * sc1;
            if (v00 > b0) { b0 = v00; i0 = col; }
            if (v01 > b0) { b0 = v01; i0 = col + 1; }
            if (v10 > b1) { b1 = v10; i1 = col; }
            if (v11 > b1) { b1 = v11; i1 = col + 1; }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, b0, o);
            int oi = __shfl_xor_sync(0xffffffffu, i0, o);
            if (ob > b0 || (ob == b0 && oi < i0)) { b0 = ob; i0 = oi; }
            ob = __shfl_xor_sync(0xffffffffu, b1, o);
            oi = __shfl_xor_sync(0xffffffffu, i1, o);
            if (ob > b1 || (ob == b1 && oi < i1)) { b1 = ob; i1 = oi; }
        }
        if ((lane & 3) == 0) {
            const int r = wm0 + (lane >> 2);
            s_val[warp & 1][r] = b0;
            s_idx[warp & 1][r] = i0;
            s_val[warp & 1][r + 8] = b1;
            s_idx[warp & 1][r + 8] = i1;
        }
        __syncthreads();
        if (tid < BM && m0 + tid < M) {
            float v = s_val[0][tid];
            int i = s_idx[0][tid];
            if (s_val[1][tid] > v) { v = s_val[1][tid]; i = s_idx[1][tid]; }  // half 1 holds the higher columns
            epi.cand[(size_t)(m0 + tid) * gridDim.x + blockIdx.x] = make_float2(v, __int_as_float(i));
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(epi.tile_ticket + blockIdx.y, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const DecodeState& st = epi.st;
            const int step = st.step[0];  // advanced only after every tile has passed this point
            if (tid < BM) s_tok[tid] = -1;
            if (tid < BM && m0 + tid < M && st.active[m0 + tid]) {
                const int ln = m0 + tid;
                float best = -INFINITY;
                int best_i = 0x7fffffff;
                for (int j = 0; j < (int)gridDim.x; ++j) {  // ascending column tiles
                    const float2 c = __ldcg(epi.cand + (size_t)ln * gridDim.x + j);
                    if (c.x > best) { best = c.x; best_i = __float_as_int(c.y); }
                }
                const int n_emitted = step - st.prefix_len + 1;  // tokens emitted incl. this one
                int next = best_i;
                if (st.forced) {
                    const size_t frow = st.forced_by_row ? (size_t)st.out_row[ln] : (size_t)ln;
                    const long long f = st.forced[frow * st.forced_stride + n_emitted];
                    next = (f < 0 || f >= epi.vocab) ? st.pad_id : (int)f;
                }
                st.out[(size_t)st.out_row[ln] * st.out_stride + n_emitted] = next;
                st.tok[ln] = next;
                const bool done = (!st.forced && next == st.eos_id) || n_emitted >= st.max_tokens;
                if (done) {
                    st.active[ln] = 0;
                    st.finish_step[ln] = n_emitted;
                    atomicSub(st.n_active, 1);
                } else {
                    s_tok[tid] = next;
                }
            }
            __syncthreads();
            if (epi.emb) {  // next step's input rows of the lanes that go on
                const int c = tid * 4;
                const float4 pv = *reinterpret_cast<const float4*>(epi.pe + (size_t)(step + 1) * kDModel + c);
                for (int r = 0; r < BM; ++r) {
                    const int tok = s_tok[r];
                    if (tok < 0) continue;
                    const float4 e = *reinterpret_cast<const float4*>(epi.emb + (size_t)tok * kDModel + c);
                    const float4 hv = make_float4(e.x + pv.x, e.y + pv.y, e.z + pv.z, e.w + pv.w);
                    *reinterpret_cast<float4*>(epi.H + (size_t)(m0 + r) * kDModel + c) = hv;
                    *reinterpret_cast<uint2*>(epi.Hb + (size_t)(m0 + r) * kDModel + c) =
                        make_uint2(pack_bf16(hv.x, hv.y), pack_bf16(hv.z, hv.w));
                }
            }
            __syncthreads();
            if (tid == 0) {
                epi.tile_ticket[blockIdx.y] = 0;
                __threadfence();
                const int tk = atomicAdd(st.ticket, 1);
                if (tk == (int)gridDim.y - 1) {
                    st.ticket[0] = 0;
                    st.step[0] = step + 1;
                }
            }
        }
    } else {
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
    }
    trace_end(trace);
}

// =============================================================================================
// Version 2 (the default): K split across the four warps.
//
// Trace stamps of version 1 on B200 (profiles/r2c_*): the operands are in shared memory 0.5-1.0 us after
// the dependency wait (one L2 round trip) and the K loop then takes another 0.7-1.7 us -- the same
// with a barrier wait per k-tile or without.  That loop is bound by shared-memory bandwidth and by
// its own serial length: every warp walks ALL K/64 k-tiles, the activation tile is read by both
// column-half warps and the weight tile by both row-half warps (16 KB of ldmatrix traffic per 8 KB
// k-tile, plus the RMSNorm row sums re-reading the activations), and with <= 16 lanes the two warps
// that own rows 16-31 multiply zeros.
//
// Here warp w owns the k-tiles kt = w, w + 4, ... and computes the WHOLE BM x BN tile over them:
//   * every byte of shared memory is read by ldmatrix exactly once (48 KB instead of 128 KB for a
//     16-lane 32-column K = 512 tile), the per-warp loop is K/256 k-tiles long instead of K/64;
//   * a warp waits only for the barriers of its own k-tiles;
//   * the RMSNorm row sums come from the A fragments the warp holds anyway (no extra reads);
//   * BM = 16 for groups of <= 16 lanes: no warp works on rows that do not exist;
//   * the four partial tiles meet in shared memory and are summed in warp order by all 128 threads,
//     thread <-> (row, column pair): the stores of the epilogue are row-contiguous.
// The order of every fp32 sum is fixed by (k-tile, warp): a row's result does not depend on its batch
// neighbours, as before.
template <int BM, int BN, int K, bool NORM, class Epi>
__global__ void __launch_bounds__(128)
    gemm_skinny2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, int M,
                        float eps, Epi epi, TraceSlot trace) {
    trace_begin(trace);
    constexpr int KT = K / 64;
    constexpr int MT = BM / 16;        // m16 tiles
    constexpr int NT = BN / 8;         // n8 tiles
    constexpr int PITCH = BN + 8;      // floats per row of a partial tile (conflict-free fragment stores)
    constexpr int EPT = BM * BN / 2 / 128;  // (row, column pair) elements per thread in the final sum
    static_assert(KT <= 16 && BN % 16 == 0 && (BM == 16 || BM == 32), "unsupported tile");
    static_assert(4 * BM * PITCH * 4 <= KT * BN * 128, "partial tiles must fit in the weight region");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    bf16* sA = reinterpret_cast<bf16*>(smem_al);    // [KT][BM*64]
    bf16* sW = sA + KT * BM * 64;                   // [KT][BN*64]
    const uint32_t bars = smem_u32(sW + KT * BN * 64);
    __shared__ float s_ss[4][BM];
    __shared__ float s_scale[BM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    if (tid == 0) {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) mbar_init(bars + 8 * kt, 2);
        mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        // weights first: they do not depend on the producer kernel (their latency hides under its tail)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            mbar_expect_tx(bars + 8 * kt, BN * 128);
            tma_load_2d(smem_u32(sW + kt * BN * 64), &tm_w, kt * 64, n0, bars + 8 * kt);
        }
    }
    pdl_wait();
    trace_mark(trace, 0);
    pdl_launch_dependents();
    if (tid == 0) {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            mbar_expect_tx(bars + 8 * kt, BM * 128);
            tma_load_2d(smem_u32(sA + kt * BM * 64), &tm_a, kt * 64, m0, bars + 8 * kt);
        }
    }
    __syncthreads();  // barrier initialisation is visible to every waiter

    // the residual values this thread will add in the epilogue: fetched now, under the operand loads
    float2 pre[EPT];
    if constexpr (EpiPrefetches<Epi>::value) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int e = tid + 128 * i, row = e / (BN / 2), cp = e - row * (BN / 2);
            pre[i] = m0 + row < M ? epi.load(m0 + row, n0 + 2 * cp) : make_float2(0.f, 0.f);
        }
    }

    float acc[MT][NT][4];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mi][j][r] = 0.f;
    float ss[MT][2];  // NORM: this lane's share of the squares of rows mi*16 + lane/4 (+ 8)
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) ss[mi][0] = ss[mi][1] = 0.f;

#pragma unroll
    for (int kt0 = 0; kt0 < KT; kt0 += 4) {
        const int kt = kt0 + warp;
        if (kt < KT) {
            mbar_wait(bars + 8 * kt, 0);
            const uint32_t baseA = smem_u32(sA + kt * BM * 64);
            const uint32_t baseW = smem_u32(sW + kt * BN * 64);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t af[MT][4], wf[NT / 2][4];
#pragma unroll
                for (int mi = 0; mi < MT; ++mi) {
                    const int row = mi * 16 + (lane & 15);
                    const int ch = kk * 2 + (lane >> 4);
                    ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], baseA + row * 128 + ((ch ^ (row & 7)) << 4));
                }
#pragma unroll
                for (int nj = 0; nj < NT / 2; ++nj) {
                    const int row = nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                    const int ch = kk * 2 + ((lane >> 3) & 1);
                    ldmatrix_x4(wf[nj][0], wf[nj][1], wf[nj][2], wf[nj][3], baseW + row * 128 + ((ch ^ (row & 7)) << 4));
                }
                if (NORM) {
                    // A fragment: regs 0 / 2 hold row lane/4 (columns 2c, 2c+1 and 2c+8, 2c+9), regs 1 / 3 row + 8
#pragma unroll
                    for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const bf162*>(&af[mi][r]));
                            ss[mi][r & 1] += f.x * f.x + f.y * f.y;
                        }
                    }
                }
#pragma unroll
                for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
                    for (int nj = 0; nj < NT / 2; ++nj) {
                        mma_bf16_16816(acc[mi][nj * 2], af[mi], wf[nj][0], wf[nj][1]);
                        mma_bf16_16816(acc[mi][nj * 2 + 1], af[mi], wf[nj][2], wf[nj][3]);
                    }
                }
            }
        }
    }
    trace_mark(trace, 1);
    __syncthreads();  // every warp is done with the operand tiles: the weight region becomes the meeting place

    float* part = reinterpret_cast<float*>(sW) + warp * BM * PITCH;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
        const int row = mi * 16 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = j * 8 + (lane & 3) * 2;
            *reinterpret_cast<float2*>(part + row * PITCH + col) = make_float2(acc[mi][j][0], acc[mi][j][1]);
            *reinterpret_cast<float2*>(part + (row + 8) * PITCH + col) = make_float2(acc[mi][j][2], acc[mi][j][3]);
        }
        if (NORM) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v = ss[mi][h];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((lane & 3) == 0) s_ss[warp][row + 8 * h] = v;
            }
        }
    }
    __syncthreads();
    if (NORM) {
        if (tid < BM) s_scale[tid] = rsqrtf(((s_ss[0][tid] + s_ss[1][tid]) + (s_ss[2][tid] + s_ss[3][tid])) * (1.0f / K) + eps);
        __syncthreads();
    }
    const float* p0 = reinterpret_cast<const float*>(sW);
    float2 val[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int e = tid + 128 * i, row = e / (BN / 2), cp = e - row * (BN / 2);
        const float* q = p0 + row * PITCH + 2 * cp;
        const float2 a = *reinterpret_cast<const float2*>(q), b = *reinterpret_cast<const float2*>(q + BM * PITCH);
        const float2 c = *reinterpret_cast<const float2*>(q + 2 * BM * PITCH), d = *reinterpret_cast<const float2*>(q + 3 * BM * PITCH);
        const float sc = NORM ? s_scale[row] : 1.f;
        val[i] = make_float2(((a.x + b.x) + (c.x + d.x)) * sc, ((a.y + b.y) + (c.y + d.y)) * sc);
    }

    if constexpr (EpiIsGreedy<Epi>::value) {
        // thread <-> (row, column pair) with BN / 2 pairs per row: a row's candidates sit in BN / 2
        // consecutive lanes, so the row arg-max is a sub-warp shuffle reduction (lowest index wins ties)
        static_assert(BN == 64, "the greedy head reduces one 64-column row per warp pass");
        __shared__ int s_tok[BM];
        __shared__ int s_last;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int row = (tid + 128 * i) / 32;  // = warp + 4 i
            float b = val[i].x;
            int bi = n0 + 2 * lane;
            if (val[i].y > b) { b = val[i].y; bi = n0 + 2 * lane + 1; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, b, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
            }
            if (lane == 0 && m0 + row < M)
                epi.cand[(size_t)(m0 + row) * gridDim.x + blockIdx.x] = make_float2(b, __int_as_float(bi));
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(epi.tile_ticket + blockIdx.y, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const DecodeState& st = epi.st;
            const int step = st.step[0];  // advanced only after every tile has passed this point
            if (tid < BM) s_tok[tid] = -1;
            if (tid < BM && m0 + tid < M && st.active[m0 + tid]) {
                const int ln = m0 + tid;
                float best = -INFINITY;
                int best_i = 0x7fffffff;
                for (int j = 0; j < (int)gridDim.x; ++j) {  // ascending column tiles
                    const float2 c = __ldcg(epi.cand + (size_t)ln * gridDim.x + j);
                    if (c.x > best) { best = c.x; best_i = __float_as_int(c.y); }
                }
                const int n_emitted = step - st.prefix_len + 1;  // tokens emitted incl. this one
                int next = best_i;
                if (st.forced) {
                    const size_t frow = st.forced_by_row ? (size_t)st.out_row[ln] : (size_t)ln;
                    const long long f = st.forced[frow * st.forced_stride + n_emitted];
                    next = (f < 0 || f >= epi.vocab) ? st.pad_id : (int)f;
                }
                st.out[(size_t)st.out_row[ln] * st.out_stride + n_emitted] = next;
                st.tok[ln] = next;
                const bool done = (!st.forced && next == st.eos_id) || n_emitted >= st.max_tokens;
                if (done) {
                    st.active[ln] = 0;
                    st.finish_step[ln] = n_emitted;
                    atomicSub(st.n_active, 1);
                } else {
                    s_tok[tid] = next;
                }
            }
            __syncthreads();
            if (epi.emb) {  // next step's input rows of the lanes that go on
                // warp w takes rows w, w + 4, ...; all loads of a row batch are issued before the first use
                // (one L2 round trip per batch of four rows instead of one per row)
                const float* pe_row = epi.pe + (size_t)(step + 1) * kDModel;
#pragma unroll
                for (int rb = 0; rb < BM / 4; rb += 4) {
                    float4 ev[4][4];
                    int toks[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = warp + 4 * (rb + u);
                        toks[u] = (rb + u < BM / 4) ? s_tok[r] : -1;
                        if (toks[u] >= 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                ev[u][q] = *reinterpret_cast<const float4*>(epi.emb + (size_t)toks[u] * kDModel + (lane + 32 * q) * 4);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (toks[u] < 0) continue;
                        const int r = warp + 4 * (rb + u);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = (lane + 32 * q) * 4;
                            const float4 pv = *reinterpret_cast<const float4*>(pe_row + c);
                            const float4 hv = make_float4(ev[u][q].x + pv.x, ev[u][q].y + pv.y, ev[u][q].z + pv.z, ev[u][q].w + pv.w);
                            *reinterpret_cast<float4*>(epi.H + (size_t)(m0 + r) * kDModel + c) = hv;
                            *reinterpret_cast<uint2*>(epi.Hb + (size_t)(m0 + r) * kDModel + c) =
                                make_uint2(pack_bf16(hv.x, hv.y), pack_bf16(hv.z, hv.w));
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                epi.tile_ticket[blockIdx.y] = 0;
                __threadfence();
                const int tk = atomicAdd(st.ticket, 1);
                if (tk == (int)gridDim.y - 1) {
                    st.ticket[0] = 0;
                    st.step[0] = step + 1;
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int e = tid + 128 * i, row = e / (BN / 2), cp = e - row * (BN / 2);
            if (m0 + row < M) {
                if constexpr (EpiPrefetches<Epi>::value)
                    epi.store(m0 + row, n0 + 2 * cp, pre[i], val[i].x, val[i].y);
                else
                    epi(m0 + row, n0 + 2 * cp, val[i].x, val[i].y);
            }
        }
    }
    trace_end(trace);
}

// how the activation tile reaches shared memory: 0 = TMA boxes (default), 1 = cp.async from all threads
// (measured on B200: 5 % slower per decode step, the TMA boxes were never the bottleneck);
// MRMT3_SKINNY_A_MODE overrides (A/B measurements)
inline int skinny_a_mode() {
    static const int mode = [] {
        const char* e = getenv("MRMT3_SKINNY_A_MODE");
        return e ? atoi(e) : 0;
    }();
    return mode;
}
// K loop: 1 = wait for every k-tile, then one straight-line block (default); 0 = barrier wait per k-tile
inline int skinny_k_mode() {
    static const int mode = [] {
        const char* e = getenv("MRMT3_SKINNY_K_MODE");
        return e ? atoi(e) : 1;
    }();
    return mode;
}

// which kernel the decode-step projections use: 2 = K split across the warps (default), 1 = version 1
inline int skinny_version() {
    static const int v = [] {
        const char* e = getenv("MRMT3_SKINNY_VERSION");
        return e ? atoi(e) : 2;
    }();
    return v;
}

template <int BM, int BN, int K, bool NORM, class Epi>
Status launch_gemm_skinny2(TmaCache& tc, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, float eps,
                           const Epi& epi, cudaStream_t stream, TraceSlot trace) {
    auto kern = gemm_skinny2_kernel<BM, BN, K, NORM, Epi>;
    constexpr int smem = (BM + BN) * K * (int)sizeof(bf16) + (K / 64) * 8 + 1024;
    MRMT3_TRY(ensure_dynamic_smem(kern, smem));
    const CUtensorMap *ma = nullptr, *mw = nullptr;
    MRMT3_TRY(tc.get(A, M, K, lda, BM, &ma));
    const CUtensorMap a_copy = *ma;  // the second lookup may rotate the cache
    MRMT3_TRY(tc.get(W, N, K, ldw, BN, &mw));
    dim3 grid(N / BN, ceil_div(M, BM));
    MRMT3_TRY(launch_pdl(kern, grid, dim3(128), smem, stream, a_copy, *mw, M, eps, epi, trace));
    return OkStatus();
}

template <int BN, int K, bool NORM, class Epi>
Status launch_gemm_skinny(TmaCache& tc, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, float eps,
                          const Epi& epi, cudaStream_t stream, TraceSlot trace = TraceSlot{nullptr, 0}) {
    if (M <= 0) return OkStatus();
    if (N % BN != 0) return Error(2, "gemm_skinny: N must be a multiple of the column tile");
    static_assert(!NORM || K == kDModel, "fused RMSNorm needs complete rows: K == d_model");
    if (skinny_version() == 2) {
        // groups of <= 16 lanes (the MR-MT3 small-batch regime) use 16-row tiles: no warp multiplies zeros
        if (M <= 16) return launch_gemm_skinny2<16, BN, K, NORM>(tc, A, lda, W, ldw, M, N, eps, epi, stream, trace);
        return launch_gemm_skinny2<32, BN, K, NORM>(tc, A, lda, W, ldw, M, N, eps, epi, stream, trace);
    }
    auto kern = gemm_skinny_kernel<BN, K, NORM, Epi>;
    constexpr int smem = (32 + BN) * K * (int)sizeof(bf16) + (K / 64) * 8 + 1024;
    MRMT3_TRY(ensure_dynamic_smem(kern, smem));
    const CUtensorMap *ma = nullptr, *mw = nullptr;
    MRMT3_TRY(tc.get(A, M, K, lda, 32, &ma));
    const CUtensorMap a_copy = *ma;  // the second lookup may rotate the cache
    MRMT3_TRY(tc.get(W, N, K, ldw, BN, &mw));
    dim3 grid(N / BN, ceil_div(M, 32));
    MRMT3_TRY(launch_pdl(kern, grid, dim3(128), smem, stream, a_copy, *mw, M, eps, epi, trace, A, lda,
                         skinny_a_mode(), skinny_k_mode()));
    return OkStatus();
}

}  // namespace mrmt3
