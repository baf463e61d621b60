// Latency / issue-interval microbenchmark of the legacy warp-level mma.sync (HMMA.16816) on sm_100a:
// one warp, clock64 around (a) a dependent chain of N mma (latency), (b) N independent mma over 8
// accumulators (issue interval), (c) ldmatrix -> mma -> shfl -> ex2 -> mma, the decode-attention chunk chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_latency hmma_latency.cu && ./hmma_latency
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void bench(long long* out, float* sink, int n_warps_active) {
    __shared__ __align__(128) uint32_t tile[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = 0x3c003c00u + i;
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= n_warps_active) return;
    const int lane = threadIdx.x & 31;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    uint32_t b0 = 0x3f803f80u + lane, b1 = 0x3f003f00u;
    float acc[8][4];
    for (int i = 0; i < 8; ++i) for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
    // (a) dependent chain
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) mma(acc[0], a, b0, b1);
    long long t1 = clock64();
    // (b) 8 independent accumulators
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) mma(acc[j], a, b0, b1);
    long long t2 = clock64();
    // (c) ldmatrix -> mma -> shfl x2 -> ex2 -> pack -> mma, 16 times (dependent through the accumulator)
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile) + (lane & 15) * 128;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    long long t3 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        uint32_t k0, k1, k2, k3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(saddr + (i & 7) * 16));
        float s[4] = {o[0] * 1e-9f, 0.f, 0.f, 0.f};
        mma(s, a, k0, k1);
        float mx = fmaxf(s[0], s[1]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float p0 = exp2f(s[0] - mx), p1 = exp2f(s[1] - mx);
        __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        uint32_t pf[4] = {*reinterpret_cast<uint32_t*>(&pk), 0u, *reinterpret_cast<uint32_t*>(&pk), 0u};
        mma(o, pf, k2, k3);
    }
    long long t4 = clock64();
    // (d) ldmatrix latency alone (dependent addresses)
    uint32_t ad = saddr;
    long long t5 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        uint32_t k0, k1, k2, k3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(ad));
        ad = saddr + ((k0 ^ k1 ^ k2 ^ k3) & 0x70);
    }
    long long t6 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t4 - t3; out[3] = t6 - t5;
    }
    float sum = o[0] + o[1] + (float)ad;
    for (int i = 0; i < 8; ++i) sum += acc[i][0] + acc[i][3];
    sink[threadIdx.x] = sum;
}

int main() {
    long long* d; float* s; long long h[4];
    cudaMalloc(&d, 64); cudaMalloc(&s, 4096);
    for (int warps : {1, 4, 8, 16}) {
        bench<<<1, 512>>>(d, s, warps); bench<<<1, 512>>>(d, s, warps);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("{\"warps_active\": %d, \"dependent_mma_cycles_each\": %.1f, \"independent_mma_issue_cycles_each\": %.1f, "
               "\"attention_chunk_chain_cycles\": %.1f, \"ldmatrix_dependent_cycles_each\": %.1f}\n",
               warps, h[0] / 64.0, h[1] / 64.0, h[2] / 16.0, h[3] / 32.0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
