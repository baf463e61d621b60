// C[M,N] = A[M,K] * W[N,K]^T for the large-M GEMMs of the path (encoder, cross-K/V projection,
// memory block, teacher-forced decoder): TMA-fed tcgen05.mma with the fp32 accumulator in TMEM.
//
//   * persistent CTAs (one per SM) walking 128 x BN output tiles (BN = 256/192/128/64, the widest
//     that divides N), BK = 64, 4-6 stage shared-memory ring, accumulator double-buffered in TMEM
//   * warp 0  : TMA producer  (cp.async.bulk.tensor.2d, 128-byte swizzle, mbarrier complete_tx)
//   * warp 1  : allocates TMEM, issues tcgen05.mma.cta_group::1.kind::f16 (one elected lane),
//               tcgen05.commit frees ring slots / publishes the accumulator
//   * warps 2-5: epilogue -- tcgen05.ld (32 lanes x 32 columns per warp), the same epilogue
//               functors as the mma.sync kernels (gemm_mma.cuh), direct global stores
//   * the epilogue warps drain accumulator buffer b while the MMA warp fills buffer b^1.
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables (the
// same fields CUTLASS's cute/arch/mma_sm100_desc.hpp names): K-major operands, SWIZZLE_128B,
// 8-row groups 1024 B apart.
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "gemm_mma.cuh"
#include "tma.cuh"

namespace mrmt3 {

constexpr int kTcBM = 128;
constexpr int kTcBK = 64;
constexpr int kTcThreads = 192;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- CTA pair (cta_group::2) forms ---------------------------------------------------------------
// One tcgen05.mma issued by the even CTA of a 2-CTA cluster multiplies a 256-row A tile (128 rows in
// each CTA's shared memory, same offset) by an N-row B tile (N / 2 rows in each CTA's shared memory)
// into 128 x N accumulators in EACH CTA's TMEM: every CTA loads half of each operand.
__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all earlier MMAs of this thread are complete) on the barrier at the same shared-memory
// offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on a barrier of the pair's even CTA
// (`bar_cluster`: a shared::cluster address from mapa)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* map, int x, int y, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_dst),
        "l"(map), "r"(x), "r"(y), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Remote arrive on a barrier of the pair's even CTA.  NOT `.release.cluster`: that form makes the thread
// wait until its earlier global stores are visible cluster-wide (ERRBAR + membar: 21 % of the pair kernel's
// stall samples in profiles/r2s_gemm_pair_stalls.txt), and nothing is published through memory here --
// the TMEM reads it orders are complete (tcgen05.wait::ld) and fenced (tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(bar_cluster) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from the other CTA of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (tile rows are 128 B, 8-row groups 1024 B)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// MN-major SWIZZLE_128B descriptor: the tile is a stack of TMA boxes of 64 reduction rows x 64
// M/N elements (128 B per row, 8-row swizzle atoms of 1024 B).  In CuTe's canonical form
// ((8,n),(8,k)):((1,LBO),(8,SBO)) (16-byte units): LBO = distance between 64-element blocks along
// M/N = one box, SBO = distance between 8-row groups along the reduction = 1024 B.
__device__ __forceinline__ uint64_t tc_smem_desc_mn(uint32_t smem_addr, uint32_t box_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((box_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: D = f32, A = B = bf16, M = 128, N = BN; bit 15 / 16 = A / B is MN-major
__host__ __device__ constexpr uint32_t tc_idesc(int bn, bool a_mn = false, bool b_mn = false, int bm = kTcBM) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}

// ---- epilogue stores: whole sectors, whole lines -----------------------------------------------------
// Measured (round 2, profiles/r2q_*): with the epilogue's global stores dropped the kernel runs at
// 1.2-1.64 PFLOP/s on the encoder shapes; with them, 0.66-1.05.  The mainloop was never the limit: the
// store path was.  An epilogue thread owns one accumulator ROW (that is how tcgen05.ld hands out TMEM
// lanes), so a warp-wide store touches 32 different rows:
//   MRMT3_EPI_WIDE=0  16-byte stores: HALF a sector each, two L2 write requests per sector
//                     (l1tex__m_l1tex2xbar_write_bytes = 302 MB for a 151 MB output)           0.66-1.05 PFLOP/s
//   MRMT3_EPI_WIDE=1  (default) 256-bit stores (STG.E.256): one whole sector per thread per instruction,
//                     still 32 lines per instruction                                           0.78-1.12
//   MRMT3_EPI_WIDE=2  the warp first transposes its 32 x 64 block through a private, swizzled
//                     shared-memory tile, so that the threads of a store instruction cover whole
//                     128-byte lines of 4-16 rows: no better for bf16 outputs (0.72-1.12), +15-30 % on
//                     fp32-output test shapes, and the fine-tune forward got SLOWER (8.7 -> 10.2 ms:
//                     the tile round trip lengthens every epilogue) -- kept as a compile-time option.
// What the three have in common: the extra time over the store-less kernel is ~ output bytes / (32 B per
// clock per SM) -- the L1 -> crossbar port, one sector-sized packet per cycle, shared with the TMA read
// requests (l1tex__m_l1tex2xbar_req_cycles_active 59-69 %).
#ifndef MRMT3_EPI_WIDE
#define MRMT3_EPI_WIDE 1
#endif
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void st_global_256f(void* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_256f(const void* p, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
// a (pointer, pitch in bytes) pair whose rows start on 32-byte boundaries
__device__ __forceinline__ bool epi_wide_ok(const void* p, size_t pitch_bytes) {
    return MRMT3_EPI_WIDE && ((reinterpret_cast<size_t>(p) | pitch_bytes) & 31) == 0;
}

// ---- 8-wide epilogue stores --------------------------------------------------------------------
// An epilogue thread owns one output row and walks its columns, so it can store 8 consecutive
// columns with one or two 16-byte transactions instead of four 4/8-byte ones.  Generic fallback:
// the pairwise functor interface of gemm_mma.cuh.
template <class Epi>
__device__ __forceinline__ void epi_store8(const Epi& epi, int row, int col, const float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) epi(row, col + 2 * j, v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void epi_store8(const EpiStoreBf16& e, int row, int col, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(e.C + (size_t)row * e.ldc + col) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void epi_store8(const EpiStoreF32& e, int row, int col, const float (&v)[8]) {
    float* c = e.C + (size_t)row * e.ldc + col;
    if (epi_wide_ok(e.C, (size_t)e.ldc * 4)) {
        st_global_256f(c, v);
        return;
    }
    float4* p = reinterpret_cast<float4*>(c);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void epi_store8(const EpiResidual& e, int row, int col, const float (&v)[8]) {
    float* h = e.H + (size_t)row * e.ldh + col;
    if (epi_wide_ok(e.H, (size_t)e.ldh * 4)) {
        float a[8];
        ld_global_256f(h, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
        st_global_256f(h, a);
        return;
    }
    float4* p = reinterpret_cast<float4*>(h);
    float4 a = p[0], b = p[1];
    p[0] = make_float4(a.x + v[0], a.y + v[1], a.z + v[2], a.w + v[3]);
    p[1] = make_float4(b.x + v[4], b.y + v[5], b.z + v[6], b.w + v[7]);
}
__device__ __forceinline__ void epi_store8(const EpiPosAdd& e, int row, int col, const float (&v)[8]) {
    const float* pe = e.pe + (size_t)(e.pos_offset + row % e.period) * e.ldh + col;
    float* h = e.H + (size_t)row * e.ldh + col;
    if (epi_wide_ok(e.H, (size_t)e.ldh * 4) && (reinterpret_cast<size_t>(e.pe) & 31) == 0) {
        float a[8];
        ld_global_256f(pe, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
        st_global_256f(h, a);
        return;
    }
    const float4* q = reinterpret_cast<const float4*>(pe);
    const float4 a = q[0], b = q[1];
    float4* p = reinterpret_cast<float4*>(h);
    p[0] = make_float4(a.x + v[0], a.y + v[1], a.z + v[2], a.w + v[3]);
    p[1] = make_float4(b.x + v[4], b.y + v[5], b.z + v[6], b.w + v[7]);
}
__device__ __forceinline__ void epi_store8(const EpiGatedGelu& e, int row, int col, const float (&v)[8]) {
    *reinterpret_cast<uint2*>(e.C + (size_t)row * e.ldc + (col >> 1)) =
        make_uint2(pack_bf16(gelu_new(v[0]) * v[1], gelu_new(v[2]) * v[3]),
                   pack_bf16(gelu_new(v[4]) * v[5], gelu_new(v[6]) * v[7]));
}
__device__ __forceinline__ void epi_store8(const EpiCrossKV& e, int row, int col, const float (&v)[8]) {
    int lane = row / e.rows_per_lane;
    int t = row - lane * e.rows_per_lane + e.t_offset;
    if (e.lane_map) lane = e.lane_map[lane];
    int layer = col / (2 * kInner);
    int r = col - layer * (2 * kInner);
    int kv = r / kInner;
    r -= kv * kInner;
    size_t off = ((((size_t)lane * e.n_layers + layer) * 2 + kv) * kHeads + (r >> 6)) * e.tk_cap + t;
    *reinterpret_cast<uint4*>(e.cache + off * kDKV + (r & 63)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// measurement only (mrmt3_test_gemm which = 6): the accumulator is read out of TMEM and dropped, which
// separates the cost of the epilogue's global stores from the rest of the kernel
struct EpiDiscard {
    __device__ __forceinline__ void operator()(int, int, float, float) const {}
};

template <class Epi>
__device__ __forceinline__ void epi_store16(const Epi& epi, int row, int col, const float (&v)[16]) {
    epi_store8(epi, row, col, *reinterpret_cast<const float(*)[8]>(&v[0]));
    epi_store8(epi, row, col + 8, *reinterpret_cast<const float(*)[8]>(&v[8]));
}
__device__ __forceinline__ void epi_store16(const EpiStoreBf16& e, int row, int col, const float (&v)[16]) {
    if (!epi_wide_ok(e.C, (size_t)e.ldc * 2)) {
        epi_store8(e, row, col, *reinterpret_cast<const float(*)[8]>(&v[0]));
        epi_store8(e, row, col + 8, *reinterpret_cast<const float(*)[8]>(&v[8]));
        return;
    }
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    st_global_256(e.C + (size_t)row * e.ldc + col, w);
}
__device__ __forceinline__ void epi_store16(const EpiCrossKV& e, int row, int col, const float (&v)[16]) {
    if (!epi_wide_ok(e.cache, 32)) {
        epi_store8(e, row, col, *reinterpret_cast<const float(*)[8]>(&v[0]));
        epi_store8(e, row, col + 8, *reinterpret_cast<const float(*)[8]>(&v[8]));
        return;
    }
    int lane = row / e.rows_per_lane;
    int t = row - lane * e.rows_per_lane + e.t_offset;
    if (e.lane_map) lane = e.lane_map[lane];
    int layer = col / (2 * kInner);
    int r = col - layer * (2 * kInner);
    int kv = r / kInner;
    r -= kv * kInner;
    size_t off = ((((size_t)lane * e.n_layers + layer) * 2 + kv) * kHeads + (r >> 6)) * e.tk_cap + t;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    st_global_256(e.cache + off * kDKV + (r & 63), w);   // 16 columns never straddle a head (16 | 64)
}

template <class Epi>
__device__ __forceinline__ void epi_store32(const Epi& epi, int row, int col, const float (&v)[32]) {
    epi_store16(epi, row, col, *reinterpret_cast<const float(*)[16]>(&v[0]));
    epi_store16(epi, row, col + 16, *reinterpret_cast<const float(*)[16]>(&v[16]));
}
// gated-GELU halves the width: 32 accumulator columns -> 16 bf16 = one sector
__device__ __forceinline__ void epi_store32(const EpiGatedGelu& e, int row, int col, const float (&v)[32]) {
    if (!epi_wide_ok(e.C, (size_t)e.ldc * 2)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) epi_store8(e, row, col + 8 * j, *reinterpret_cast<const float(*)[8]>(&v[8 * j]));
        return;
    }
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16(gelu_new(v[4 * j]) * v[4 * j + 1], gelu_new(v[4 * j + 2]) * v[4 * j + 3]);
    st_global_256(e.C + (size_t)row * e.ldc + (col >> 1), w);
}

// Epilogues that READ what they add to (the residual stream) pay one global-load round trip per 32-column
// chunk if the load is issued after the accumulator chunk has arrived: eight dependent round trips per
// tile, longer than the K = 384 mainloop.  Functors with EpiPrefetch<>::value provide
//     fetch(row, col, h[32])            -- issue the loads of 32 columns (no use of the values)
//     store32(row, col, acc[32], h[32]) -- combine and store
// and the kernel fetches chunk c + 1 (chunk 0 before it even waits for the accumulator) while chunk c is
// read out of TMEM and stored.
template <class Epi> struct EpiPrefetch { static constexpr bool value = false; };
template <> struct EpiPrefetch<EpiResidual> { static constexpr bool value = true; };
__device__ __forceinline__ void epi_fetch32(const EpiResidual& e, int row, int col, float (&h)[32]) {
    const float* p = e.H + (size_t)row * e.ldh + col;
    if (epi_wide_ok(e.H, (size_t)e.ldh * 4)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ld_global_256f(p + 8 * j, *reinterpret_cast<float(*)[8]>(&h[8 * j]));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(&h[4 * j]) = *reinterpret_cast<const float4*>(p + 4 * j);
    }
}
__device__ __forceinline__ void epi_store32_fetched(const EpiResidual& e, int row, int col, const float (&v)[32], const float (&h)[32]) {
    float* p = e.H + (size_t)row * e.ldh + col;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = h[j] + v[j];
    if (epi_wide_ok(e.H, (size_t)e.ldh * 4)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st_global_256f(p + 8 * j, &o[8 * j]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(p + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
}

// accumulator columns one thread hands to the functor at a time in the line-coalesced epilogue: 32 bytes
// of OUTPUT per thread (8 fp32, 16 bf16, 32 accumulator columns for the gated-GELU's 16 bf16)
template <class Epi> struct EpiWidth { static constexpr int value = 8; };
template <> struct EpiWidth<EpiStoreBf16> { static constexpr int value = 16; };
template <> struct EpiWidth<EpiCrossKV> { static constexpr int value = 16; };
template <> struct EpiWidth<EpiGatedGelu> { static constexpr int value = 32; };
template <class Epi, int W>
__device__ __forceinline__ void epi_emit(const Epi& epi, int row, int col, const float (&v)[W]) {
    if constexpr (W == 8) epi_store8(epi, row, col, v);
    else if constexpr (W == 16) epi_store16(epi, row, col, v);
    else epi_store32(epi, row, col, v);
}

// NCTA = 2: the CTA pair computes a 256 x BN tile; each CTA stages its 128 rows of A and BN / 2 rows of W
template <int BN, int NCTA = 1>
struct TcSmem {
    static constexpr int kStages = NCTA == 2 ? (BN > 128 ? 6 : 8) : (BN > 128 ? 4 : 6);
    static constexpr int kABytes = kTcBM * kTcBK * 2;
    static constexpr int kWBytes = BN / NCTA * kTcBK * 2;
    static constexpr int kStageBytes = kABytes + kWBytes;
    static constexpr int kEpiOffset = kStages * kStageBytes;     // 4 epilogue warps x (32 rows x 64 fp32) transpose tiles
    static constexpr int kEpiBytes = MRMT3_EPI_WIDE >= 2 ? 4 * 32 * 64 * 4 : 0;
    static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
    static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + slack for 1024-B alignment
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
    static constexpr int kAccCols = BN > 128 ? 256 : 128;   // TMEM columns per accumulator buffer
    static constexpr int kTmemCols = 2 * kAccCols;          // double-buffered accumulator
};

// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x + i * gridDim.x with
// the n-block fastest, so CTAs running at the same time share their A rows through L2.  The
// accumulator is double-buffered in TMEM: the epilogue warps drain tile i while the MMA warp
// already accumulates tile i + 1.
//
// MODE 0: D = A W^T, both operands K-major (A (M x K), W (N x K) row-major).
// MODE 1, the weight-gradient form D[m, n] = sum_r A[r, m] B[r, n] on ROW-major A (R x M) and
// B (R x N): both operands MN-major, so neither has to be transposed in memory first.
// MODE 2, the data-gradient form D = A B on row-major A (M x K) and B (K x N): A K-major, B MN-major.
// In modes 1 and 2 K is the number of reduction rows (any value: TMA zero-fills past the end) and
// the k ranges of the splits are ceil(ceil(K / 64) / k_splits) blocks each.
//
// NCTA = 2 (launched as clusters of two CTAs = one TPC): the pair works on 256 x BN output tiles with
// tcgen05.mma.cta_group::2.  Each CTA stages ITS 128 rows of A and ITS BN / 2 rows of W, so the pair
// moves (256 + BN) x 64 operand elements per 256 x BN x 64 MACs: 1.5-1.67x fewer bytes through L2 per
// flop than the single-CTA tile (the kernel sits at the L2-slice throughput cap).
//   * both producers load with the cta_group::2 TMA form, which counts the bytes on the EVEN CTA's
//     full barrier; only the even CTA arms it (expect_tx of both halves) and only its MMA warp waits;
//   * the even CTA's elected thread issues the MMAs; tcgen05.commit multicasts to the barrier at the
//     same offset in both CTAs: "ring slot free" for both producers, "accumulator complete" for both
//     epilogues (each drains the 128 x BN accumulator in its own TMEM);
//   * all 256 epilogue threads of the pair arrive on the even CTA's "accumulator drained" barrier.
template <int BN, class Epi, int MODE = 0, int NCTA = 1>
__global__ void __launch_bounds__(kTcThreads, 1)
    gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                           int M, int N, int K, ARowMap amap, Epi epi, int k_splits) {
    using S = TcSmem<BN, NCTA>;
    constexpr int kStages = S::kStages;
    constexpr bool MN = MODE != 0, kAMn = MODE == 1, kBMn = MODE != 0;
    constexpr bool kPair = NCTA == 2;
    constexpr int kTileM = kTcBM * NCTA;                          // output rows per work item
    constexpr int kWRows = BN / NCTA;                             // W rows (output columns) staged by one CTA
    static_assert(!kPair || (BN % 32 == 0 && (!kBMn || kWRows % 64 == 0)), "tile not splittable across the pair");
    extern __shared__ unsigned char tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-B alignment
    unsigned char* base_ptr = tc_smem_raw + (base - raw);
    const uint32_t bar_full = base + S::kBarOffset;               // kStages x 8 B
    const uint32_t bar_empty = bar_full + kStages * 8;            // kStages x 8 B
    const uint32_t bar_acc_full = bar_empty + kStages * 8;        // 2 x 8 B
    const uint32_t bar_acc_empty = bar_acc_full + 16;             // 2 x 8 B
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + S::kBarOffset + 192);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;         // 0 = the CTA that issues the MMAs
    const int worker = kPair ? blockIdx.x >> 1 : blockIdx.x;
    const int n_workers = kPair ? gridDim.x >> 1 : gridDim.x;
    // split-K (wgrad: small output, long reduction): work item t -> output tile t % n_out, k range
    // split t / n_out; split s stores its partial sums at rows [s*M, (s+1)*M) of the output
    const int KB = MN ? ((K + kTcBK - 1) / kTcBK + k_splits - 1) / k_splits : K / kTcBK / k_splits;
    const int tiles_n = N / BN;
    const int tiles_m = (M + kTileM - 1) / kTileM;
    const int n_out = tiles_n * tiles_m;
    const int n_tiles = n_out * k_splits;

    if constexpr (kPair) cluster_sync_all();                      // both CTAs of the pair are resident
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + b * 8, 1);
            mbar_init(bar_acc_empty + b * 8, 128 * NCTA);  // every epilogue thread (of the pair) arrives
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             smem_u32((const void*)tmem_slot)),
                         "n"(S::kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             smem_u32((const void*)tmem_slot)),
                         "n"(S::kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = worker; t < n_tiles; t += n_workers) {
                const int to = t % n_out, kb0 = (t / n_out) * KB;
                const int m0 = (to / tiles_n) * kTileM + (int)rank * kTcBM;   // this CTA's 128 rows of A
                const int n0 = (to % tiles_n) * BN + (int)rank * kWRows;      // ... and its rows of W
                // A rows may be gathered block-wise (cross-K/V projection picks each lane's segment)
                int arow = m0;
                if (amap.map) arow = amap.map[m0 / amap.block] * amap.block + m0 % amap.block;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_empty + s * 8, ((it / kStages) & 1) ^ 1);
                    const uint32_t sa = base + s * S::kStageBytes;
                    constexpr int kBox = 64 * kTcBK * 2;  // MN-major: 64 reduction rows x 64 columns
                    const int r0 = (kb0 + kb) * kTcBK;
                    uint32_t bar = bar_full + s * 8;
                    if constexpr (kPair) {
                        if (rank == 0) mbar_expect_tx(bar, 2 * S::kStageBytes);
                        bar = mapa_shared(bar, 0);
                    } else {
                        mbar_expect_tx(bar, S::kStageBytes);
                    }
                    auto load = [&](uint32_t dst, const CUtensorMap* map, int x, int y) {
                        if constexpr (kPair) tma_load_2d_2cta(dst, map, x, y, bar);
                        else tma_load_2d(dst, map, x, y, bar);
                    };
                    if constexpr (kAMn) {
#pragma unroll
                        for (int j = 0; j < kTcBM / 64; ++j) load(sa + j * kBox, &tmap_a, m0 + 64 * j, r0);
                    } else {
                        load(sa, &tmap_a, r0, arow);
                    }
                    if constexpr (kBMn) {
#pragma unroll
                        for (int j = 0; j < kWRows / 64; ++j) load(sa + S::kABytes + j * kBox, &tmap_w, n0 + 64 * j, r0);
                    } else {
                        load(sa + S::kABytes, &tmap_w, r0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = tc_idesc(BN, kAMn, kBMn, kTileM);
            int it = 0, i = 0;
            for (int t = worker; t < n_tiles; t += n_workers, ++i) {
                const int buf = i & 1;
                // the epilogue (of both CTAs) has drained this buffer
                if constexpr (kPair) mbar_wait_cluster(bar_acc_empty + buf * 8, ((i >> 1) & 1) ^ 1);
                else mbar_wait(bar_acc_empty + buf * 8, ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * S::kAccCols);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % kStages;
                    if constexpr (kPair) mbar_wait_cluster(bar_full + s * 8, (it / kStages) & 1);
                    else mbar_wait(bar_full + s * 8, (it / kStages) & 1);
                    tc_fence_after();
                    const uint32_t sa = base + s * S::kStageBytes;
                    const uint64_t da = kAMn ? tc_smem_desc_mn(sa, 64 * kTcBK * 2) : tc_smem_desc(sa);
                    const uint64_t dw = kBMn ? tc_smem_desc_mn(sa + S::kABytes, 64 * kTcBK * 2) : tc_smem_desc(sa + S::kABytes);
                    // K-major: advance 16 elements (32 B) along K inside the 128-B swizzle atom, +2
                    // in the 16-byte-granular start-address field; MN-major: 16 reduction rows = two
                    // whole 1024-B atoms, +128
                    constexpr uint64_t kStepA = kAMn ? 128 : 2, kStepB = kBMn ? 128 : 2;
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k) {
                        if constexpr (kPair) tc_mma_f16_2cta(tacc, da + kStepA * k, dw + kStepB * k, idesc, (kb | k) != 0);
                        else tc_mma_f16(tacc, da + kStepA * k, dw + kStepB * k, idesc, (kb | k) != 0);
                    }
                    // frees the ring slot (in both CTAs) once the MMAs have read it
                    if constexpr (kPair) tc_commit_2cta(bar_empty + s * 8, 3);
                    else tc_commit(bar_empty + s * 8);
                }
                // accumulator of this tile complete
                if constexpr (kPair) tc_commit_2cta(bar_acc_full + buf * 8, 3);
                else tc_commit(bar_acc_full + buf * 8);
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        int i = 0;
        for (int t = worker; t < n_tiles; t += n_workers, ++i) {
            const int to = t % n_out;
            const int m0 = (to / tiles_n) * kTileM + (int)rank * kTcBM;
            const int n0 = (to % tiles_n) * BN;
            const int out_row0 = (t / n_out) * M;  // split-K partials are stacked along the rows
            const int buf = i & 1;
            if constexpr (!(EpiPrefetch<Epi>::value && MRMT3_EPI_WIDE < 2)) {
                mbar_wait(bar_acc_full + buf * 8, (i >> 1) & 1);
                tc_fence_after();
            }
            const int row = m0 + quarter * 32 + lane;
            const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * S::kAccCols);
#if MRMT3_EPI_WIDE >= 2
            // 64 accumulator columns at a time: row-per-thread out of TMEM into the warp's swizzled tile
            // (16-byte unit u of row r at unit (u & 8) | ((u & 7) ^ (r & 7)): conflict-free both ways), then
            // W columns per thread with the threads of an instruction side by side along the row
            constexpr int W = EpiWidth<Epi>::value, kPieces = 64 / W, kRowsPerInstr = 32 / kPieces;
            float* tile = reinterpret_cast<float*>(base_ptr + S::kEpiOffset) + quarter * (32 * 64);
            const int piece = lane % kPieces, rsub = lane / kPieces;
#pragma unroll 1
            for (int c = 0; c < BN / 64; ++c) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tc_ld_32x32(tacc + (uint32_t)(c * 64 + h * 32), v);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<uint4*>(tile + lane * 64 + h * 32 + ((k ^ (lane & 7)) << 2)) =
                            make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kPieces; ++i) {
                    const int r = rsub + kRowsPerInstr * i;
                    float f[W];
#pragma unroll
                    for (int j = 0; j < W / 4; ++j) {
                        const int u = piece * (W / 4) + j;
                        const float4 x = *reinterpret_cast<const float4*>(tile + r * 64 + ((u & 8) << 2) + (((u & 7) ^ (r & 7)) << 2));
                        f[4 * j] = x.x;
                        f[4 * j + 1] = x.y;
                        f[4 * j + 2] = x.z;
                        f[4 * j + 3] = x.w;
                    }
                    const int orow = m0 + quarter * 32 + r;
                    if (orow < M) epi_emit<Epi, W>(epi, out_row0 + orow, n0 + c * 64 + piece * W, f);
                }
                __syncwarp();
            }
#else
            if constexpr (EpiPrefetch<Epi>::value) {
                float hbuf[2][32];
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) {
                    if (c == 0 && row < M) epi_fetch32(epi, out_row0 + row, n0, hbuf[0]);   // (hoisted above the accumulator wait below)
                    if (c + 1 < BN / 32 && row < M) epi_fetch32(epi, out_row0 + row, n0 + (c + 1) * 32, hbuf[(c + 1) & 1]);
                    if (c == 0) {
                        mbar_wait(bar_acc_full + buf * 8, (i >> 1) & 1);
                        tc_fence_after();
                    }
                    uint32_t v[32];
                    tc_ld_32x32(tacc + (uint32_t)(c * 32), v);
                    if (row < M) {
                        float f[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                        epi_store32_fetched(epi, out_row0 + row, n0 + c * 32, f, hbuf[c & 1]);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tc_ld_32x32(tacc + (uint32_t)(c * 32), v);
                    if (row < M) {
                        float f[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                        epi_store32(epi, out_row0 + row, n0 + c * 32, f);
                    }
                }
            }
#endif
            tc_fence_before();
            if constexpr (kPair) mbar_arrive_cluster(mapa_shared(bar_acc_empty + buf * 8, 0));
            else mbar_arrive(bar_acc_empty + buf * 8);
        }
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if constexpr (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(S::kTmemCols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(S::kTmemCols)
                         : "memory");
    }
}

// ---- launchers ----------------------------------------------------------------------------------
// CTA-pair tiles: option `gemm_2cta` / MRMT3_GEMM_2CTA (1 = use the pair kernel where the tile splits and
// the problem has at least one 256-row tile per pair; 0 = single-CTA kernel everywhere).  Measured on the six
// encoder shapes at M = 65 536 (profiles/r2t_gemm_pair2.jsonl): pair 841-1225 TFLOP/s, single 787-1116,
// cuBLAS 867-1199.  The weight-gradient form (MODE 1) stays on single-CTA tiles: its outputs are a few
// dozen tiles, split-K spreads them over the SMs, and 256-row tiles would halve the number of work items.
constexpr int kGemm2CtaDefault = 1;
inline int& gemm_2cta_flag() {
    static int flag = [] {
        const char* e = getenv("MRMT3_GEMM_2CTA");
        return e ? atoi(e) : kGemm2CtaDefault;
    }();
    return flag;
}
inline void gemm_configure_2cta(int on) { gemm_2cta_flag() = on < 0 ? kGemm2CtaDefault : on; }

inline Status tc_sm_count(int* n_sms) {
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        MRMT3_CUDA_TRY(cudaGetDevice(&dev));
        MRMT3_CUDA_TRY(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    }
    *n_sms = cached;
    return OkStatus();
}

template <int BN, class Epi, int MODE, int NCTA>
Status launch_gemm_tc_kernel(const CUtensorMap* ma, const CUtensorMap* mw, int M, int N, int K, ARowMap amap, const Epi& epi,
                             cudaStream_t stream, int k_splits) {
    int n_sms = 0;
    MRMT3_TRY(tc_sm_count(&n_sms));
    auto kern = gemm_tn_tcgen05_kernel<BN, Epi, MODE, NCTA>;
    using S = TcSmem<BN, NCTA>;
    MRMT3_TRY(ensure_dynamic_smem(kern, S::kTotal));
    const int n_tiles = (N / BN) * ceil_div(M, kTcBM * NCTA) * k_splits;
    if constexpr (NCTA == 1) {
        kern<<<std::min(n_tiles, n_sms), kTcThreads, S::kTotal, stream>>>(*ma, *mw, M, N, K, amap, epi, k_splits);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * std::min(n_tiles, n_sms / 2));
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = S::kTotal;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MRMT3_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, *ma, *mw, M, N, K, amap, epi, k_splits));
    }
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

// the pair kernel pays off when every pair has a full 256-row tile to work on
inline bool tc_use_pair(int M) { return gemm_2cta_flag() != 0 && M >= 2 * kTcBM; }

// a_rows: rows addressable through A (>= M; larger when amap gathers from a bigger tensor)
template <int BN, class Epi>
Status launch_gemm_tc_bn(TmaCache& tc, const CUtensorMap* ma, const bf16* W, int ldw, int M, int N, int K,
                         ARowMap amap, const Epi& epi, cudaStream_t stream, int k_splits) {
    const CUtensorMap* mw = nullptr;
    if constexpr (BN >= 128) {
        if (tc_use_pair(M)) {
            const CUtensorMap a_copy = *ma;  // the lookup below may rotate the cache
            MRMT3_TRY(tc.get(W, N, K, ldw, BN / 2, &mw));
            return launch_gemm_tc_kernel<BN, Epi, 0, 2>(&a_copy, mw, M, N, K, amap, epi, stream, k_splits);
        }
    }
    const CUtensorMap a_copy = *ma;
    MRMT3_TRY(tc.get(W, N, K, ldw, BN, &mw));
    return launch_gemm_tc_kernel<BN, Epi, 0, 1>(&a_copy, mw, M, N, K, amap, epi, stream, k_splits);
}

// D (M x N) = A^T B over R rows (MODE 1) or A B (MODE 2); B (R x N, pitch ldb) row-major bf16, MN-major
// operand: the pair kernel needs whole 64-column boxes per CTA, i.e. BN = 256 or 128.
// With k_splits > 1, split s stores its partial sums at rows [s*M, (s+1)*M) of the output.
template <int MODE, class Epi>
Status launch_gemm_tc_mn_any(const CUtensorMap* ma, const CUtensorMap* mb, int M, int N, int R, const Epi& epi,
                             cudaStream_t stream, int k_splits) {
    const ARowMap none{nullptr, 1};
    const bool pair = MODE == 2 && tc_use_pair(M);
    if (N % 256 == 0)
        return pair ? launch_gemm_tc_kernel<256, Epi, MODE, 2>(ma, mb, M, N, R, none, epi, stream, k_splits)
                    : launch_gemm_tc_kernel<256, Epi, MODE, 1>(ma, mb, M, N, R, none, epi, stream, k_splits);
    if (N % 192 == 0 && !pair) return launch_gemm_tc_kernel<192, Epi, MODE, 1>(ma, mb, M, N, R, none, epi, stream, k_splits);
    if (N % 128 == 0)
        return pair ? launch_gemm_tc_kernel<128, Epi, MODE, 2>(ma, mb, M, N, R, none, epi, stream, k_splits)
                    : launch_gemm_tc_kernel<128, Epi, MODE, 1>(ma, mb, M, N, R, none, epi, stream, k_splits);
    return launch_gemm_tc_kernel<64, Epi, MODE, 1>(ma, mb, M, N, R, none, epi, stream, k_splits);
}

template <class Epi>
Status launch_gemm_tc_mn(TmaCache& tc, const bf16* A, int lda, int M, const bf16* B, int ldb, int N, int R,
                         const Epi& epi, cudaStream_t stream, int k_splits = 1) {
    if (M <= 0 || R <= 0) return OkStatus();
    if (N % 64 != 0 || k_splits < 1) return Error(2, "gemm_tc_mn: N must be a multiple of 64");
    const CUtensorMap *pa = nullptr, *mb = nullptr;
    MRMT3_TRY(tc.get(A, R, M, lda, 64, &pa));
    const CUtensorMap a_copy = *pa;  // the second lookup may evict the cache
    MRMT3_TRY(tc.get(B, R, N, ldb, 64, &mb));
    return launch_gemm_tc_mn_any<1>(&a_copy, mb, M, N, R, epi, stream, k_splits);
}

// D (M x N) = A B; A (M x K, pitch lda) and B (K x N, pitch ldb) row-major bf16 (the data gradient
// dX = dY W with the weight as stored, no W^T copy)
template <class Epi>
Status launch_gemm_tc_nn(TmaCache& tc, const bf16* A, int lda, int M, const bf16* B, int ldb, int N, int K,
                         const Epi& epi, cudaStream_t stream) {
    if (M <= 0 || K <= 0) return OkStatus();
    if (N % 64 != 0) return Error(2, "gemm_tc_nn: N must be a multiple of 64");
    const CUtensorMap *pa = nullptr, *mb = nullptr;
    MRMT3_TRY(tc.get(A, M, K, lda, kTcBM, &pa));
    const CUtensorMap a_copy = *pa;
    MRMT3_TRY(tc.get(B, K, N, ldb, 64, &mb));
    return launch_gemm_tc_mn_any<2>(&a_copy, mb, M, N, K, epi, stream, 1);
}

template <class Epi>
Status launch_gemm_tc(TmaCache& tc, const bf16* A, int lda, long a_rows, ARowMap amap, const bf16* W, int ldw,
                      int M, int N, int K, const Epi& epi, cudaStream_t stream, int k_splits = 1) {
    if (M <= 0) return OkStatus();
    if (K % kTcBK != 0 || N % 64 != 0) return Error(2, "gemm_tc: K and N must be multiples of 64");
    if (k_splits < 1 || (K / kTcBK) % k_splits != 0) return Error(2, "gemm_tc: k_splits must divide K / 64");
    if (amap.map && amap.block % kTcBM != 0) return Error(2, "gemm_tc: gather block must be a multiple of 128 rows");
    const CUtensorMap* ma = nullptr;
    MRMT3_TRY(tc.get(A, a_rows, K, lda, kTcBM, &ma));
    // widest tile that divides N: wider tiles move fewer operand bytes per MAC through L2
    if (N % 256 == 0) return launch_gemm_tc_bn<256>(tc, ma, W, ldw, M, N, K, amap, epi, stream, k_splits);
    if (N % 192 == 0) return launch_gemm_tc_bn<192>(tc, ma, W, ldw, M, N, K, amap, epi, stream, k_splits);
    if (N % 128 == 0) return launch_gemm_tc_bn<128>(tc, ma, W, ldw, M, N, K, amap, epi, stream, k_splits);
    return launch_gemm_tc_bn<64>(tc, ma, W, ldw, M, N, K, amap, epi, stream, k_splits);
}

}  // namespace mrmt3
