// C[M,N] = A[M,K] * W[N,K]^T  (both operands K-major bf16, fp32 accumulate) with a fused
// epilogue functor.  Warp-level mma.sync (m16n8k16) kernel fed by a 3-stage cp.async ring with
// 128-byte XOR-swizzled shared tiles.
//
// This is the small-M / odd-shape GEMM of the library (decode-step projections with M = batch,
// the memory block).  Large-M encoder GEMMs go through the TMA + tcgen05 kernel in
// gemm_tcgen05.cu.  nn.Linear layout: W is (out_features, in_features) row-major, exactly the
// reference's state-dict tensors (models/t5.py weight contract).
#pragma once
#include "common.cuh"

namespace mrmt3 {

constexpr int kGemmBK = 64;
constexpr int kGemmStages = 3;

// optional gather on A rows: logical row r -> source row map[r / block] * block + r % block
struct ARowMap {
    const int* map;  // nullptr => identity
    int block;
};

// ------------------------------------------------------------------------------------------
// epilogues: called once per (row, even col) with the two adjacent accumulators
struct EpiStoreBf16 {
    bf16* C;
    int ldc;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        *reinterpret_cast<uint32_t*>(C + (size_t)row * ldc + col) = pack_bf16(v0, v1);
    }
};

struct EpiStoreF32 {
    float* C;
    int ldc;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        *reinterpret_cast<float2*>(C + (size_t)row * ldc + col) = make_float2(v0, v1);
    }
};

// decode-step logits: row = lane, written at C[out_row[lane]*lane_stride + (step-prefix)*ldc + col]
// (out_row == nullptr -> lane; step == nullptr -> plain (lanes, ldc) matrix)
struct EpiStoreF32Step {
    float* C;
    int ldc;
    size_t lane_stride;
    const int* step;
    int prefix;
    const int* out_row;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        size_t r = out_row ? (size_t)out_row[row] : (size_t)row;
        size_t off = r * lane_stride + col;
        if (step) off += (size_t)(step[0] - prefix) * ldc;
        *reinterpret_cast<float2*>(C + off) = make_float2(v0, v1);
    }
};

// residual stream update: H (fp32) += acc
struct EpiResidual {
    float* H;
    int ldh;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        float2* p = reinterpret_cast<float2*>(H + (size_t)row * ldh + col);
        float2 h = *p;
        h.x += v0;
        h.y += v1;
        *p = h;
    }
};

// stack input: H = acc + PE[pos_offset + row % period]   (reference models/t5.py:596-598)
struct EpiPosAdd {
    float* H;
    int ldh;
    const float* pe;  // (n_pos, ldh)
    int period;
    int pos_offset;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        const float2 p = *reinterpret_cast<const float2*>(
            pe + (size_t)(pos_offset + row % period) * ldh + col);
        *reinterpret_cast<float2*>(H + (size_t)row * ldh + col) = make_float2(v0 + p.x, v1 + p.y);
    }
};

// gated-GELU: W rows are interleaved (2j -> wi_0[j], 2j+1 -> wi_1[j]); out[row][j] = gelu(a)*b
struct EpiGatedGelu {
    bf16* C;
    int ldc;
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        C[(size_t)row * ldc + (col >> 1)] = __float2bfloat16(gelu_new(v0) * v1);
    }
};

// cross-attention K/V for all decoder layers in one GEMM (N = n_layers * 2 * inner), scattered
// into the cross cache [lane][layer][k|v][head][tk_cap][64].
struct EpiCrossKV {
    bf16* cache;
    int rows_per_lane;  // 256 encoder rows, or mem_len memory rows
    int t_offset;       // 0 for encoder rows, 256 for memory rows
    int n_layers;
    int tk_cap;
    const int* lane_map;  // optional: logical lane (row / rows_per_lane) -> cache lane
    __device__ __forceinline__ void operator()(int row, int col, float v0, float v1) const {
        int lane = row / rows_per_lane;
        int t = row - lane * rows_per_lane + t_offset;
        if (lane_map) lane = lane_map[lane];
        int layer = col / (2 * kInner);
        int r = col - layer * (2 * kInner);
        int kv = r / kInner;
        r -= kv * kInner;
        int head = r >> 6;
        int d = r & 63;
        size_t off = ((((size_t)lane * n_layers + layer) * 2 + kv) * kHeads + head) * tk_cap + t;
        *reinterpret_cast<uint32_t*>(cache + off * kDKV + d) = pack_bf16(v0, v1);
    }
};

// ------------------------------------------------------------------------------------------
template <int BM, int BN, int WARPS_M, int WARPS_N, class Epi>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32)
    gemm_tn_mma_kernel(const bf16* __restrict__ A, int lda, ARowMap amap,
                       const bf16* __restrict__ W, int ldw, int M, int N, int K, Epi epi) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WTM = BM / WARPS_M;
    constexpr int WTN = BN / WARPS_N;
    constexpr int MI = WTM / 16;
    constexpr int NI = WTN / 8;
    static_assert(WTM % 16 == 0 && WTN % 16 == 0, "warp tile must be a multiple of 16x16");
    constexpr int A_TILE = BM * kGemmBK;  // elements
    constexpr int W_TILE = BN * kGemmBK;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    bf16* sA = reinterpret_cast<bf16*>(smem_raw);
    bf16* sW = sA + kGemmStages * A_TILE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm0 = (warp / WARPS_N) * WTM;
    const int wn0 = (warp % WARPS_N) * WTN;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int KT = K / kGemmBK;

    auto load_tile = [&](int kt, int stage) {
        const int k0 = kt * kGemmBK;
        bf16* dA = sA + stage * A_TILE;
        bf16* dW = sW + stage * W_TILE;
#pragma unroll
        for (int i = 0; i < (BM * 8 + NT - 1) / NT; ++i) {
            int c = tid + i * NT;
            if ((BM * 8) % NT == 0 || c < BM * 8) {
                int row = c >> 3, ch = c & 7;
                int grow = m0 + row;
                bool pred = grow < M;
                int srow = pred ? grow : 0;
                if (amap.map) srow = amap.map[srow / amap.block] * amap.block + srow % amap.block;
                cp_async16(dA + row * kGemmBK + ((ch ^ (row & 7)) << 3),
                           A + (size_t)srow * lda + k0 + ch * 8, pred);
            }
        }
#pragma unroll
        for (int i = 0; i < (BN * 8 + NT - 1) / NT; ++i) {
            int c = tid + i * NT;
            if ((BN * 8) % NT == 0 || c < BN * 8) {
                int row = c >> 3, ch = c & 7;
                cp_async16(dW + row * kGemmBK + ((ch ^ (row & 7)) << 3),
                           W + (size_t)(n0 + row) * ldw + k0 + ch * 8, true);
            }
        }
    };

    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

#pragma unroll
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        {
            int nk = kt + kGemmStages - 1;
            if (nk < KT) load_tile(nk, nk % kGemmStages);
            cp_async_commit();
        }
        const bf16* tA = sA + (kt % kGemmStages) * A_TILE;
        const bf16* tW = sW + (kt % kGemmStages) * W_TILE;
        const uint32_t baseA = smem_u32(tA);
        const uint32_t baseW = smem_u32(tW);
#pragma unroll
        for (int kk = 0; kk < kGemmBK / 16; ++kk) {
            uint32_t af[MI][4];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                int row = wm0 + mi * 16 + (lane & 15);
                int ch = kk * 2 + (lane >> 4);
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3],
                            baseA + row * (kGemmBK * 2) + ((ch ^ (row & 7)) << 4));
            }
#pragma unroll
            for (int nj = 0; nj < NI / 2; ++nj) {
                int row = wn0 + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                int ch = kk * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3, baseW + row * (kGemmBK * 2) + ((ch ^ (row & 7)) << 4));
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    mma_bf16_16816(acc[mi][nj * 2], af[mi], b0, b1);
                    mma_bf16_16816(acc[mi][nj * 2 + 1], af[mi], b2, b3);
                }
            }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            int row = m0 + wm0 + mi * 16 + (lane >> 2);
            int col = n0 + wn0 + ni * 8 + (lane & 3) * 2;
            if (row < M) epi(row, col, acc[mi][ni][0], acc[mi][ni][1]);
            if (row + 8 < M) epi(row + 8, col, acc[mi][ni][2], acc[mi][ni][3]);
        }
    }
}

template <int BM, int BN>
constexpr int gemm_smem_bytes() {
    return kGemmStages * (BM + BN) * kGemmBK * (int)sizeof(bf16);
}

// Host launcher.  Picks the tile by M: small tiles keep more CTAs in flight for decode-step
// shapes (M = batch), the 128x128 tile is for anything large that is not routed to tcgen05.
template <class Epi>
Status launch_gemm_mma(const bf16* A, int lda, ARowMap amap, const bf16* W, int ldw, int M, int N,
                       int K, const Epi& epi, cudaStream_t stream) {
    if (M <= 0) return OkStatus();
    if (K % kGemmBK != 0 || N % 64 != 0)
        return Error(2, "gemm_mma: K must be a multiple of 64 and N a multiple of 64");
    if (M > 512 && N % 128 == 0) {
        constexpr int BM = 128, BN = 128;
        auto kern = gemm_tn_mma_kernel<BM, BN, 2, 4, Epi>;
        constexpr int smem = gemm_smem_bytes<BM, BN>();
        MRMT3_TRY(ensure_dynamic_smem(kern, smem));
        dim3 grid(N / BN, ceil_div(M, BM));
        kern<<<grid, 256, smem, stream>>>(A, lda, amap, W, ldw, M, N, K, epi);
    } else if (M > 64) {
        constexpr int BM = 64, BN = 64;
        auto kern = gemm_tn_mma_kernel<BM, BN, 2, 2, Epi>;
        constexpr int smem = gemm_smem_bytes<BM, BN>();
        MRMT3_TRY(ensure_dynamic_smem(kern, smem));
        dim3 grid(N / BN, ceil_div(M, BM));
        kern<<<grid, 128, smem, stream>>>(A, lda, amap, W, ldw, M, N, K, epi);
    } else {
        constexpr int BM = 32, BN = 64;
        auto kern = gemm_tn_mma_kernel<BM, BN, 1, 4, Epi>;
        constexpr int smem = gemm_smem_bytes<BM, BN>();
        MRMT3_TRY(ensure_dynamic_smem(kern, smem));
        dim3 grid(N / BN, ceil_div(M, BM));
        kern<<<grid, 128, smem, stream>>>(A, lda, amap, W, ldw, M, N, K, epi);
    }
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

}  // namespace mrmt3
