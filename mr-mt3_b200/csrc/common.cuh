// Shared device/host helpers for the mrmt3_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <mutex>
#include <set>
#include <string>
#include <utility>

namespace mrmt3 {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// model constants of the path (reference pretrained/config.json, contrib/spectrograms.py:34-41)
constexpr int kDModel = 512;
constexpr int kHeads = 6;
constexpr int kDKV = 64;
constexpr int kInner = kHeads * kDKV;  // 384
constexpr int kDFF = 1024;
constexpr int kVocab = 1536;
constexpr int kSegFrames = 256;
constexpr int kHop = 128;
constexpr int kNFFT = 2048;
constexpr int kMels = 512;
constexpr int kSegSamples = kSegFrames * kHop;  // 32768
constexpr int kKVPage = 128;                    // positions per KV-cache page

// ---- error plumbing: no C++ exceptions cross the C ABI -------------------------------------
struct Status {
    int code;
    std::string msg;
    bool ok() const { return code == 0; }
};

#define MRMT3_CUDA_TRY(expr)                                                               \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            char _b[512];                                                                  \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                     cudaGetErrorString(_e));                                              \
            return ::mrmt3::Status{(int)_e == 0 ? 1 : (int)_e, _b};                        \
        }                                                                                  \
    } while (0)

#define MRMT3_TRY(expr)                  \
    do {                                 \
        ::mrmt3::Status _s = (expr);     \
        if (!_s.ok()) return _s;         \
    } while (0)

#define MRMT3_CHECK_LAUNCH() MRMT3_CUDA_TRY(cudaGetLastError())

inline Status OkStatus() { return Status{0, ""}; }
inline Status Error(int code, const std::string& m) { return Status{code, m}; }

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// cp.async 16 B with zero-fill when !pred (src-size 0)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool pred) {
    uint32_t d = smem_u32(smem_dst);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                            uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// gelu_new (tanh form), reference via HF ACT2FN["gelu_new"]
__device__ __forceinline__ float gelu_new(float a) {
    const float k = 0.7978845608028654f;  // sqrt(2/pi)
    float inner = k * (a + 0.044715f * a * a * a);
    return 0.5f * a * (1.0f + tanhf(inner));
}

// ---- dropout (fine-tune step) -------------------------------------------------------------------
// Counter-based: one 64-bit hash of (seed, tensor id, index / 4) yields four 16-bit draws, element
// `index` is kept iff its draw >= p * 65536.  The backward regenerates every mask instead of storing
// it, and the CPU oracle mirrors it (oracle/mt3_oracle.py:dropout_keep; mrmt3_dropout_keep_host
// exposes this definition to the CPU tests).  `scale` = 1 / (1 - p);
// p == 0 disables everything.
struct DropSpec {
    unsigned long long seed;  // already mixed with the tensor id
    unsigned int threshold;   // p * 2^16
    float scale;
    __host__ __device__ bool on() const { return threshold != 0u; }
};
__host__ __device__ __forceinline__ unsigned long long drop_mix_tid(unsigned long long seed, unsigned int tid) {
    return seed ^ ((unsigned long long)tid * 0x9E3779B97F4A7C15ull);
}
// the spec of tensor `tid` under dropout probability p and step seed `seed`
__host__ __device__ __forceinline__ DropSpec make_drop_spec(float p, unsigned long long seed, unsigned int tid) {
    if (!(p > 0.f)) return DropSpec{0ull, 0u, 1.f};
    const float th = p * 65536.0f;
    return DropSpec{drop_mix_tid(seed, tid), (unsigned int)(th > 65535.f ? 65535.f : th), 1.f / (1.f - p)};
}
// the four draws of the aligned group index / 4
__host__ __device__ __forceinline__ unsigned long long drop_hash4(const DropSpec& d, unsigned long long group) {
    unsigned long long x = d.seed + group * 0xD1B54A32D192ED03ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
__host__ __device__ __forceinline__ float drop_lane(const DropSpec& d, unsigned long long h, unsigned int lane) {
    return ((unsigned int)(h >> (16u * lane)) & 0xFFFFu) >= d.threshold ? d.scale : 0.f;
}
__host__ __device__ __forceinline__ float drop_factor(const DropSpec& d, unsigned long long index) {
    return d.on() ? drop_lane(d, drop_hash4(d, index >> 2), (unsigned int)index & 3u) : 1.f;
}
// factors of two consecutive elements (one hash when they share a group)
__host__ __device__ __forceinline__ void drop_factor2(const DropSpec& d, unsigned long long index, float& f0, float& f1) {
    const unsigned long long h0 = drop_hash4(d, index >> 2);
    const unsigned long long h1 = ((index + 1) >> 2) == (index >> 2) ? h0 : drop_hash4(d, (index + 1) >> 2);
    f0 = drop_lane(d, h0, (unsigned int)index & 3u);
    f1 = drop_lane(d, h1, (unsigned int)(index + 1) & 3u);
}
// factors of the aligned group 4g .. 4g+3
__host__ __device__ __forceinline__ void drop_factor4(const DropSpec& d, unsigned long long group, float (&f)[4]) {
    const unsigned long long h = drop_hash4(d, group);
#pragma unroll
    for (unsigned int k = 0; k < 4; ++k) f[k] = drop_lane(d, h, k);
}

// 2^x as one MUFU.EX2 (exp2f adds a range test and two predicated multiplies so that results below
// 2^-126 come out as denormals; softmax weights that small are zero for every purpose here)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// streaming 16 B load that does not allocate in L1 (KV cache / one-shot reads)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- mbarrier / bulk-copy (TMA) wrappers ------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned); completion is
// signalled on `bar` as `bytes` of transaction count
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar,
                                          uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
            smem_dst),
        "l"(gsrc), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n_threads) : "memory");
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: opt in once per
// (device, kernel), whichever handle / thread launches it first
template <class Kernel>
inline Status ensure_dynamic_smem(Kernel kern, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    MRMT3_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(dev, reinterpret_cast<const void*>(kern));
    if (done.count(key)) return OkStatus();
    MRMT3_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert(key);
    return OkStatus();
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// The decode step is a chain of ~70 short dependent kernels.  Launched with the programmatic
// stream-serialization attribute, kernel N+1 is scheduled while kernel N is still running; it may
// do anything that does not depend on N (fetch weights, old KV rows) and then blocks in
// pdl_wait() until N has completed and its writes are visible.  Works eagerly and under stream
// capture (the edge becomes a programmatic dependency in the CUDA graph).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::); }

// ---- in-graph timeline trace ------------------------------------------------------------------
// When a trace buffer is attached, every decode-step kernel stamps %globaltimer at the entry of its
// earliest CTA and at the exit of its latest CTA into slot `slot` (two u64 per slot).  The slots
// are baked into the captured graph, so after a replay the buffer holds the timeline of that step
// as it really ran inside the graph (overlaps, gaps) -- ncu serialises launches and cannot show it.
struct TraceSlot {
    unsigned long long* buf;  // nullptr = tracing off
    int slot;
};
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(const TraceSlot& t) {
    // plain store by CTA 0: a later replay overwrites an earlier one (the end stamp is an atomicMax
    // of a monotonic clock, so it also ends up holding the latest replay)
    if (t.buf && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        t.buf[2 * t.slot] = global_timer();
}
// optional intermediate stamps of CTA 0 (which = 0: dependency wait returned, 1: operands in shared
// memory); they live in the upper half of the trace buffer (slot + 128)
__device__ __forceinline__ void trace_mark(const TraceSlot& t, int which) {
    if (t.buf && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        t.buf[2 * (t.slot + 128 * (1 + (which >> 1))) + (which & 1)] = global_timer();
}
__device__ __forceinline__ void trace_end(const TraceSlot& t) {
    if (t.buf && threadIdx.x == 0) atomicMax(t.buf + 2 * t.slot + 1, global_timer());
}

template <class... KArgs, class... Args>
inline Status launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                         Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MRMT3_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    return OkStatus();
}

}  // namespace mrmt3
