// Log-mel frontend (frontend.cu): tables + launcher.
#pragma once
#include "common.cuh"

namespace mrmt3 {

constexpr int kMaxBand = 12;  // max contiguous non-zero FFT bins per mel filter (measured: 10)

struct FrontendTables {
    const float2* tw1024;
    const float2* tw2048;
    const float* window;
    const int* band_start;  // (512) first FFT bin of mel filter m
    const int* band_count;  // (512) number of taps
    const float* band_w;    // (kMaxBand, 512) tap j of filter m at [j*512 + m]
    const float2* tw_t;     // (32, 32): exp(-2 pi i n2 k1 / 1024) at [k1 * 32 + n2] (register FFT kernel)
};

class Frontend {
public:
    Status init();
    void destroy();
    // dense (1025, 512) fp32 host table, e.g. torchaudio's melscale_fbanks output
    Status set_filterbank(const float* fb_dense_host);
    // audio: device fp32; segment i covers audio[seg_start[i] : seg_start[i] + seg_len[i]]
    // (seg_len <= 32768, the rest of the segment reads as zeros); valid_frames may be null.
    Status run(const float* audio, const long long* seg_start, const int* seg_len,
               const int* valid_frames, int n_seg, int mel_norm, float* out_f32, bf16* out_bf16,
               cudaStream_t stream) const;

private:
    float2* d_tw1024_ = nullptr;
    float2* d_tw2048_ = nullptr;
    float* d_window_ = nullptr;
    int* d_band_start_ = nullptr;
    int* d_band_count_ = nullptr;
    float* d_band_w_ = nullptr;
    float2* d_tw_t_ = nullptr;
};

}  // namespace mrmt3
