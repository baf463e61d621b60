// Log-mel frontend: framing + periodic-Hann window + 2048-point real FFT + magnitude + HTK mel
// projection + safe_log (+ optional clip/scale normalisation and pad-frame zeroing), fused in
// one kernel.  Restates reference contrib/spectrograms.py:92-145 (torch branch) and
// inference.py:100-126 per 256-frame segment:
//   x = 32768 segment samples || 1920 zeros;  frame k = x[128k : 128k+2048] * hann_periodic
//   -> rfft -> |.| (power=1.0) -> . fb (1025x512) -> log(x<=0 ? 1e-5 : x)
//   -> [mel_norm] clip(-12,5), (y+12)/17 -> rows >= valid_frames zeroed AFTER normalisation.
//
// STFT intermediates never touch HBM: the 2048-point real FFT runs as a 1024-point complex
// radix-4 Stockham FFT in shared memory (fp32), the mel projection uses the filterbank's band
// structure (<= 12 contiguous non-zeros per mel bin; the dense table is 0.37 % dense) in fp32.
// A bf16 tensor-core GEMM for the mel projection would miss the 1e-3 log-mel tolerance (two
// 2^-9 roundings per product) and move 270x more flops; see DESIGN.md.
#include "frontend.cuh"

#include <cmath>
#include <vector>

namespace mrmt3 {

constexpr int kFrThreads = 256;
constexpr int kFramesPerCta = 8;
constexpr int kSpan = kNFFT + (kFramesPerCta - 1) * kHop;  // samples one CTA touches

struct FrontendSmem {
    float2 buf0[1024];
    float2 buf1[1024];
    float2 tw1024[1024];  // exp(-2 pi i m / 1024)
    float2 tw2048[1025];  // exp(-2 pi i k / 2048)
    float window[kNFFT];
    float span[kSpan];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(kFrThreads)
    logmel_kernel(const float* __restrict__ audio, const long long* __restrict__ seg_start,
                  const int* __restrict__ seg_len, const int* __restrict__ valid_frames,
                  FrontendTables tab, int mel_norm, float* __restrict__ out_f32,
                  bf16* __restrict__ out_bf16) {
    extern __shared__ __align__(16) unsigned char fr_smem[];
    FrontendSmem& sm = *reinterpret_cast<FrontendSmem*>(fr_smem);
    const int tid = threadIdx.x;
    const int seg = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int nvalid = valid_frames ? valid_frames[seg] : kSegFrames;
    const size_t out_row0 = (size_t)seg * kSegFrames + f0;

    // frames at or past `nvalid` are zero in the output (reference inference.py:125-126)
    const int n_live = max(0, min(kFramesPerCta, nvalid - f0));
    for (int f = n_live; f < kFramesPerCta; ++f) {
        for (int m = tid; m < kMels; m += kFrThreads) {
            if (out_f32) out_f32[(out_row0 + f) * kMels + m] = 0.f;
            if (out_bf16) out_bf16[(out_row0 + f) * kMels + m] = __float2bfloat16(0.f);
        }
    }
    if (n_live == 0) return;

    // stage tables and this CTA's audio span
    for (int i = tid; i < 1024; i += kFrThreads) sm.tw1024[i] = tab.tw1024[i];
    for (int i = tid; i < 1025; i += kFrThreads) sm.tw2048[i] = tab.tw2048[i];
    for (int i = tid; i < kNFFT; i += kFrThreads) sm.window[i] = tab.window[i];
    {
        const long long base = seg_start[seg];
        const int len = seg_len[seg];
        const int s0 = f0 * kHop;
        for (int i = tid; i < kSpan; i += kFrThreads) {
            int s = s0 + i;
            sm.span[i] = (s < len) ? audio[base + s] : 0.f;  // zero tail == pad_end + segment pad
        }
    }
    __syncthreads();

    for (int f = 0; f < n_live; ++f) {
        // z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1]
        const float* x = sm.span + f * kHop;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n = tid + i * kFrThreads;
            float2 xv = *reinterpret_cast<const float2*>(x + 2 * n);
            float2 wv = *reinterpret_cast<const float2*>(sm.window + 2 * n);
            sm.buf0[n] = make_float2(xv.x * wv.x, xv.y * wv.y);
        }
        __syncthreads();

        // 5 radix-4 Stockham stages, one butterfly per thread per stage
        float2* src = sm.buf0;
        float2* dst = sm.buf1;
#pragma unroll
        for (int stage = 0; stage < 5; ++stage) {
            const int Ns = 1 << (2 * stage);
            const int k = tid & (Ns - 1);
            float2 v0 = src[tid], v1 = src[tid + 256], v2 = src[tid + 512], v3 = src[tid + 768];
            if (stage > 0) {
                const int tstep = 256 >> (2 * stage);  // 1024 / (4 Ns)
                v1 = cmul(v1, sm.tw1024[k * tstep]);
                v2 = cmul(v2, sm.tw1024[2 * k * tstep]);
                v3 = cmul(v3, sm.tw1024[3 * k * tstep]);
            }
            float2 a = make_float2(v0.x + v2.x, v0.y + v2.y);
            float2 b = make_float2(v0.x - v2.x, v0.y - v2.y);
            float2 c = make_float2(v1.x + v3.x, v1.y + v3.y);
            float2 d = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // -i (v1 - v3)
            const int j0 = ((tid >> (2 * stage)) << (2 * stage + 2)) + k;
            dst[j0] = make_float2(a.x + c.x, a.y + c.y);
            dst[j0 + Ns] = make_float2(b.x + d.x, b.y + d.y);
            dst[j0 + 2 * Ns] = make_float2(a.x - c.x, a.y - c.y);
            dst[j0 + 3 * Ns] = make_float2(b.x - d.x, b.y - d.y);
            __syncthreads();
            float2* t = src;
            src = dst;
            dst = t;
        }
        // result Z is in `src` (== buf1 after 5 stages); magnitudes go to `dst` (== buf0)
        float* mag = reinterpret_cast<float*>(dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int k = tid + i * kFrThreads;  // 0..1023
            float2 zk = src[k];
            float2 zc = src[(1024 - k) & 1023];
            zc.y = -zc.y;
            float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            float2 df = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            float2 xo = make_float2(df.y, -df.x);  // -i * df
            float2 t = cmul(sm.tw2048[k], xo);
            float re = xe.x + t.x, im = xe.y + t.y;
            mag[k] = sqrtf(re * re + im * im);
        }
        if (tid == 0) {  // k = 1024 (Nyquist): X = Re(Z0) - Im(Z0)
            float2 z0 = src[0];
            mag[1024] = fabsf(z0.x - z0.y);
        }
        __syncthreads();

        // banded mel projection + safe_log (+ normalisation)
#pragma unroll
        for (int i = 0; i < kMels / kFrThreads; ++i) {
            int m = tid + i * kFrThreads;
            int k0 = tab.band_start[m];
            int cnt = tab.band_count[m];
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc += mag[k0 + j] * tab.band_w[j * kMels + m];
            float y = logf(acc <= 0.f ? 1e-5f : acc);
            if (mel_norm) {
                y = fminf(fmaxf(y, -12.f), 5.f);
                y = (y + 12.f) / 17.f;
            }
            size_t o = (out_row0 + f) * kMels + m;
            if (out_f32) out_f32[o] = y;
            if (out_bf16) out_bf16[o] = __float2bfloat16(y);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant (the default).  The shared-memory radix-4 kernel above spends its time in
// shared memory: five passes with 7 CTA barriers per frame and 4-way bank conflicts on the strided
// butterfly writes (ncu, round 1: 94 M conflicts, 842 us for 256 segments = 2 % of the HBM roofline).
// Here ONE WARP owns a frame and the 1024-point complex FFT is a 32 x 32 decomposition held in
// registers:  n = 32 n1 + n2,  k = k1 + 32 k2,
//     Z[k1 + 32 k2] = sum_n2 W32^(n2 k2) . W1024^(n2 k1) . ( sum_n1 z[32 n1 + n2] W32^(n1 k1) )
//   1. lane n2 loads z[32 n1 + n2], n1 = 0..31 (windowed, conflict-free) and runs a 32-point FFT in
//      registers (radix-2 DIF, twiddles are compile-time constants, result in bit-reversed positions)
//   2. twiddle W1024^(n2 k1) from a lane-contiguous table, then a 32 x 32 transpose through a padded
//      per-warp tile (one store + one load per element, no CTA barrier: only __syncwarp)
//   3. lane k1 runs the second 32-point FFT over n2 -> Z[k1 + 32 k2]
//   4. Z in natural order to the warp's tile, X[k] from Z[k] and conj(Z[1024 - k]) (both accesses
//      lane-contiguous), |X[k]| to the warp's magnitude row, banded mel projection + log + clip/scale,
//      16 mel bins per lane, every store a coalesced 128-byte (fp32) / 64-byte (bf16) row piece.
// Tables (window, twiddles, filter bands) are read through the read-only path (lane-contiguous, L1
// resident: 40 KB) instead of being re-staged into shared memory by every CTA.
constexpr int kF2Frames = 16;                              // frames per CTA
constexpr int kF2Warps = 4;
constexpr int kF2Span = kNFFT + (kF2Frames - 1) * kHop;    // 3968 samples
constexpr int kF2TileFloats = 32 * 33 * 2;                 // padded 32 x 32 float2 tile (also Z in natural order)
constexpr int kF2MagFloats = 1028;                         // 1025 magnitudes, padded
constexpr int kF2SmemBytes = (kF2Span + kF2Warps * (kF2TileFloats + kF2MagFloats)) * (int)sizeof(float);

__host__ __device__ constexpr int bitrev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// d * exp(-2 pi i j / 32), j a compile-time constant after unrolling
__device__ __forceinline__ float2 mul_w32(float2 d, int j) {
    constexpr float kCos[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
                                0.f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
                                -0.70710678118654757f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
    constexpr float kSin[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                                0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                                1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f};
    if (j == 0) return d;
    if (j == 8) return make_float2(d.y, -d.x);  // -i d
    const float c = kCos[j], s = kSin[j];       // (a + i b)(c - i s) = (a c + b s) + i (b c - a s)
    return make_float2(d.x * c + d.y * s, d.y * c - d.x * s);
}

// in-place 32-point DFT, decimation in frequency: v[p] ends up holding X[bitrev5(p)]
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if ((i & h) == 0) {
                const int j = (i & (h - 1)) * (16 / h);
                const float2 a = v[i], b = v[i + h];
                v[i] = make_float2(a.x + b.x, a.y + b.y);
                v[i + h] = mul_w32(make_float2(a.x - b.x, a.y - b.y), j);
            }
        }
    }
}

__global__ void __launch_bounds__(kF2Warps * 32)
    logmel_regfft_kernel(const float* __restrict__ audio, const long long* __restrict__ seg_start,
                         const int* __restrict__ seg_len, const int* __restrict__ valid_frames,
                         FrontendTables tab, int mel_norm, float* __restrict__ out_f32,
                         bf16* __restrict__ out_bf16) {
    extern __shared__ __align__(16) unsigned char fr_smem[];
    float* span = reinterpret_cast<float*>(fr_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* tile = span + kF2Span + warp * (kF2TileFloats + kF2MagFloats);
    float* mag = tile + kF2TileFloats;
    const int seg = blockIdx.y;
    const int f0 = blockIdx.x * kF2Frames;
    const int nvalid = valid_frames ? valid_frames[seg] : kSegFrames;
    const size_t out_row0 = (size_t)seg * kSegFrames + f0;
    const int n_live = max(0, min(kF2Frames, nvalid - f0));

    // frames at or past `nvalid` are zero in the output (reference inference.py:125-126)
    for (int f = n_live + warp; f < kF2Frames; f += kF2Warps) {
#pragma unroll
        for (int j = 0; j < kMels / 128; ++j) {
            const size_t o = (out_row0 + f) * kMels + j * 128 + lane * 4;
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + o) = make_uint2(0u, 0u);
        }
    }
    if (n_live == 0) return;

    // this CTA's audio span; samples at or past seg_len read as zeros (pad_end + segment padding)
    {
        const long long base = seg_start[seg];
        const int len = seg_len[seg];
        const int s0 = f0 * kHop;
        const int need = kNFFT + (n_live - 1) * kHop;
        const float* src = audio + base + s0;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = tid * 4; i < need; i += kF2Warps * 32 * 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int s = s0 + i;
                if (s + 3 < len) {
                    v = __ldg(reinterpret_cast<const float4*>(src + i));
                } else {
                    if (s < len) v.x = src[i];
                    if (s + 1 < len) v.y = src[i + 1];
                    if (s + 2 < len) v.z = src[i + 2];
                }
                *reinterpret_cast<float4*>(span + i) = v;
            }
        } else {
            for (int i = tid; i < need; i += kF2Warps * 32) span[i] = (s0 + i < len) ? src[i] : 0.f;
        }
    }
    __syncthreads();

    // the 16 mel bins of this lane (m = lane + 32 j): band tables stay in registers across frames
    // b_cmax[j]: the widest band among the 32 bins of group j (warp-uniform): the tap loop runs to it in
    // unrolled blocks of four -- taps past a bin's own count carry zero weight in the padded table --
    // so that four independent load pairs are in flight instead of one dependent pair per iteration
    int b_start[kMels / 32], b_cmax[kMels / 32];
#pragma unroll
    for (int j = 0; j < kMels / 32; ++j) {
        b_start[j] = __ldg(tab.band_start + lane + 32 * j);
        int c = __ldg(tab.band_count + lane + 32 * j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
        b_cmax[j] = c;
    }
    const float2* win2 = reinterpret_cast<const float2*>(tab.window);
    float2* tile2 = reinterpret_cast<float2*>(tile);

    for (int f = warp; f < n_live; f += kF2Warps) {
        const float2* x2 = reinterpret_cast<const float2*>(span + f * kHop);
        float2 v[32];
        // 1. z[32 n1 + lane] = w[2n] x[2n] + i w[2n+1] x[2n+1]
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const float2 xv = x2[32 * n1 + lane];
            const float2 wv = __ldg(win2 + 32 * n1 + lane);
            v[n1] = make_float2(xv.x * wv.x, xv.y * wv.y);
        }
        fft32_dif(v);
        // 2. twiddle + transpose: element (n2 = lane, k1) -> tile[k1][lane]
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const int k1 = bitrev5(p);
            const float2 y = cmul(v[p], __ldg(tab.tw_t + k1 * 32 + lane));
            tile2[k1 * 33 + lane] = y;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = tile2[lane * 33 + n2];
        __syncwarp();
        // 3. lane = k1: second FFT over n2 -> Z[k1 + 32 k2] at position bitrev5(k2)
        fft32_dif(v);
        // 4. Z in natural order, then X[k] = Xe[k] + W2048^k Xo[k] from Z[k] and conj(Z[1024 - k])
#pragma unroll
        for (int p = 0; p < 32; ++p) tile2[lane + 32 * bitrev5(p)] = v[p];
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const int k = lane + 32 * bitrev5(p);
            const float2 zk = v[p];
            float2 zc = tile2[(1024 - k) & 1023];
            zc.y = -zc.y;
            const float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 df = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            const float2 t = cmul(__ldg(tab.tw2048 + k), make_float2(df.y, -df.x));  // W^k . (-i df)
            const float re = xe.x + t.x, im = xe.y + t.y;
            mag[k] = sqrtf(re * re + im * im);
            if (k == 0) mag[1024] = fabsf(zk.x - zk.y);  // Nyquist: Re(Z0) - Im(Z0)
        }
        __syncwarp();
        // banded mel projection + safe_log (+ normalisation); lane owns bins lane + 32 j
#pragma unroll
        for (int j = 0; j < kMels / 32; ++j) {
            const int m = lane + 32 * j;
            float acc = 0.f;
            for (int t0 = 0; t0 < b_cmax[j]; t0 += 4) {   // kMaxBand is a multiple of 4: t0 + 3 stays inside the table
                float mv[4], wv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    mv[u] = mag[min(b_start[j] + t0 + u, 1024)];  // clamped: a finite value under a zero weight
                    wv[u] = __ldg(tab.band_w + (t0 + u) * kMels + m);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += mv[u] * wv[u];  // taps in ascending order, as the reference's matmul row
            }
            float y = logf(acc <= 0.f ? 1e-5f : acc);
            if (mel_norm) {
                y = fminf(fmaxf(y, -12.f), 5.f);
                y = (y + 12.f) / 17.f;
            }
            const size_t o = (out_row0 + f) * kMels + m;
            if (out_f32) out_f32[o] = y;
            if (out_bf16) out_bf16[o] = __float2bfloat16(y);
        }
        __syncwarp();  // the tile and the magnitude row are reused by the next frame
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tables
static void default_filterbank(std::vector<float>& fb) {
    // HTK mel triangles between 20 and 7600 Hz over linspace(0, 8000, 1025), no area norm.
    // Exact-arithmetic version; the Python binding normally overrides it with the fp32 table
    // torchaudio builds (see mr-mt3_b200/spectrograms.py) so that it matches the reference's.
    const int nf = kNFFT / 2 + 1;
    auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
    auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
    std::vector<double> fpts(kMels + 2);
    double m0 = hz2mel(20.0), m1 = hz2mel(7600.0);
    for (int i = 0; i < kMels + 2; ++i) fpts[i] = mel2hz(m0 + (m1 - m0) * i / (kMels + 1));
    fb.assign((size_t)nf * kMels, 0.f);
    for (int k = 0; k < nf; ++k) {
        double f = 8000.0 * k / (nf - 1);
        for (int m = 0; m < kMels; ++m) {
            double down = (f - fpts[m]) / (fpts[m + 1] - fpts[m]);
            double up = (fpts[m + 2] - f) / (fpts[m + 2] - fpts[m + 1]);
            double v = std::fmax(0.0, std::fmin(down, up));
            fb[(size_t)k * kMels + m] = (float)v;
        }
    }
}

Status Frontend::set_filterbank(const float* fb_dense_host) {
    const int nf = kNFFT / 2 + 1;
    std::vector<int> start(kMels, 0), count(kMels, 0);
    std::vector<float> w((size_t)kMaxBand * kMels, 0.f);
    for (int m = 0; m < kMels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < nf; ++k) {
            if (fb_dense_host[(size_t)k * kMels + m] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        if (first < 0) continue;
        if (last - first + 1 > kMaxBand)
            return Error(2, "mel filterbank column has more than kMaxBand contiguous taps");
        start[m] = first;
        count[m] = last - first + 1;
        for (int k = first; k <= last; ++k)
            w[(size_t)(k - first) * kMels + m] = fb_dense_host[(size_t)k * kMels + m];
    }
    MRMT3_CUDA_TRY(cudaMemcpy(d_band_start_, start.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
    MRMT3_CUDA_TRY(cudaMemcpy(d_band_count_, count.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
    MRMT3_CUDA_TRY(cudaMemcpy(d_band_w_, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    return OkStatus();
}

Status Frontend::init() {
    const double kPi = 3.14159265358979323846;
    std::vector<float2> t1(1024), t2(1025);
    for (int i = 0; i < 1024; ++i)
        t1[i] = make_float2((float)std::cos(-2.0 * kPi * i / 1024.0), (float)std::sin(-2.0 * kPi * i / 1024.0));
    for (int i = 0; i < 1025; ++i)
        t2[i] = make_float2((float)std::cos(-2.0 * kPi * i / 2048.0), (float)std::sin(-2.0 * kPi * i / 2048.0));
    std::vector<float> win(kNFFT);
    for (int i = 0; i < kNFFT; ++i) win[i] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * i / kNFFT));
    MRMT3_CUDA_TRY(cudaMalloc(&d_tw1024_, 1024 * sizeof(float2)));
    MRMT3_CUDA_TRY(cudaMalloc(&d_tw2048_, 1025 * sizeof(float2)));
    MRMT3_CUDA_TRY(cudaMalloc(&d_window_, kNFFT * sizeof(float)));
    MRMT3_CUDA_TRY(cudaMalloc(&d_band_start_, kMels * sizeof(int)));
    MRMT3_CUDA_TRY(cudaMalloc(&d_band_count_, kMels * sizeof(int)));
    MRMT3_CUDA_TRY(cudaMalloc(&d_band_w_, (size_t)kMaxBand * kMels * sizeof(float)));
    MRMT3_CUDA_TRY(cudaMemcpy(d_tw1024_, t1.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
    MRMT3_CUDA_TRY(cudaMemcpy(d_tw2048_, t2.data(), 1025 * sizeof(float2), cudaMemcpyHostToDevice));
    MRMT3_CUDA_TRY(cudaMemcpy(d_window_, win.data(), kNFFT * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float2> tt(1024);
    for (int k1 = 0; k1 < 32; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -2.0 * kPi * (double)(n2 * k1) / 1024.0;
            tt[k1 * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    MRMT3_CUDA_TRY(cudaMalloc(&d_tw_t_, 1024 * sizeof(float2)));
    MRMT3_CUDA_TRY(cudaMemcpy(d_tw_t_, tt.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
    MRMT3_CUDA_TRY(cudaFuncSetAttribute(logmel_regfft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2SmemBytes));
    std::vector<float> fb;
    default_filterbank(fb);
    MRMT3_TRY(set_filterbank(fb.data()));
    MRMT3_CUDA_TRY(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FrontendSmem)));
    return OkStatus();
}

void Frontend::destroy() {
    cudaFree(d_tw1024_);
    cudaFree(d_tw2048_);
    cudaFree(d_window_);
    cudaFree(d_band_start_);
    cudaFree(d_band_count_);
    cudaFree(d_band_w_);
    cudaFree(d_tw_t_);
    d_tw1024_ = nullptr;
}

Status Frontend::run(const float* audio, const long long* seg_start, const int* seg_len,
                     const int* valid_frames, int n_seg, int mel_norm, float* out_f32,
                     bf16* out_bf16, cudaStream_t stream) const {
    if (n_seg <= 0) return OkStatus();
    FrontendTables tab{d_tw1024_, d_tw2048_, d_window_, d_band_start_, d_band_count_, d_band_w_, d_tw_t_};
    static const int variant = [] {   // 2 = register-resident FFT (default), 1 = shared-memory radix-4 (A/B partner)
        const char* e = getenv("MRMT3_FRONTEND_VARIANT");
        return e ? atoi(e) : 2;
    }();
    if (variant == 1) {
        dim3 grid(kSegFrames / kFramesPerCta, n_seg);
        logmel_kernel<<<grid, kFrThreads, sizeof(FrontendSmem), stream>>>(
            audio, seg_start, seg_len, valid_frames, tab, mel_norm, out_f32, out_bf16);
    } else {
        dim3 grid(kSegFrames / kF2Frames, n_seg);
        logmel_regfft_kernel<<<grid, kF2Warps * 32, kF2SmemBytes, stream>>>(
            audio, seg_start, seg_len, valid_frames, tab, mel_norm, out_f32, out_bf16);
    }
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

}  // namespace mrmt3
