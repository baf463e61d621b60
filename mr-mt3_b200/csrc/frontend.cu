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
    d_tw1024_ = nullptr;
}

Status Frontend::run(const float* audio, const long long* seg_start, const int* seg_len,
                     const int* valid_frames, int n_seg, int mel_norm, float* out_f32,
                     bf16* out_bf16, cudaStream_t stream) const {
    if (n_seg <= 0) return OkStatus();
    FrontendTables tab{d_tw1024_, d_tw2048_, d_window_, d_band_start_, d_band_count_, d_band_w_};
    dim3 grid(kSegFrames / kFramesPerCta, n_seg);
    logmel_kernel<<<grid, kFrThreads, sizeof(FrontendSmem), stream>>>(
        audio, seg_start, seg_len, valid_frames, tab, mel_norm, out_f32, out_bf16);
    MRMT3_CHECK_LAUNCH();
    return OkStatus();
}

}  // namespace mrmt3
