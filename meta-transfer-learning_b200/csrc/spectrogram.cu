// Device-side input features: SpectrogramParser.parse_audio of the reference
// (utils/data_loader.py:65-96: n_fft = win = sr * 0.02, hop = sr * 0.01, symmetric window, librosa.stft defaults =
// centred frames with reflect padding, log1p(|STFT|), utterance-level (x - mean) / std with the unbiased std).
//
// The transform is a DFT of 320-sample frames: 161 bins x 320 samples per frame.  That is 0.1 MFLOP per frame -- a
// 1-second utterance is 10 MFLOP, a 50-second one 0.5 GFLOP -- so it runs as a direct real-input DFT on the CUDA
// cores: one CTA per block of frames, the windowed frames and a sin / cos table of ONE period (n_fft entries,
// index = (bin * sample) mod n_fft, exact integer arithmetic -- no range-reduction error) in shared memory, one thread per
// (bin, frame) pair, coalesced writes along time.  Memory traffic (the wave once, the F x T plane once) is what an
// HBM roofline would count; the arithmetic hides under the H2D copy of the wave.
#include "kernels.h"
#include <math.h>

constexpr int SPEC_FRAMES = 4;         // frames per CTA (16 left a 1 s utterance on 7 CTAs: 259 us under ncu)
constexpr int SPEC_MAX_FFT = 1024;

__global__ void __launch_bounds__(256) spectrogram_kernel(const float* __restrict__ wav, int n, int n_fft, int hop,
                                                          const float* __restrict__ window, float* __restrict__ out,
                                                          int ld_out, int T) {
  extern __shared__ float sm[];
  float* cs = sm;                         // cos table, n_fft
  float* sn = sm + n_fft;                 // sin table, n_fft
  float* fr = sm + 2 * n_fft;             // SPEC_FRAMES windowed frames of n_fft samples
  const int t0 = blockIdx.x * SPEC_FRAMES, pad = n_fft / 2, F = n_fft / 2 + 1;
  for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)i / (float)n_fft, &s, &c);
    cs[i] = c; sn[i] = s;
  }
  for (int i = threadIdx.x; i < SPEC_FRAMES * n_fft; i += blockDim.x) {
    const int f = i / n_fft, k = i - f * n_fft, t = t0 + f;
    float v = 0.f;
    if (t < T) {
      int j = t * hop + k - pad;          // index into the un-padded wave; reflect (no edge repeat) outside [0, n)
      if (j < 0) j = -j;
      if (j >= n) j = 2 * (n - 1) - j;
      v = wav[j] * window[k];
    }
    fr[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F * SPEC_FRAMES; i += blockDim.x) {
    const int bin = i / SPEC_FRAMES, f = i - bin * SPEC_FRAMES, t = t0 + f;
    if (t >= T) continue;
    const float* x = fr + f * n_fft;
    float re = 0.f, im = 0.f;
    int ph = 0;                           // (bin * k) mod n_fft
    for (int k = 0; k < n_fft; ++k) {
      re = fmaf(x[k], cs[ph], re);
      im = fmaf(x[k], sn[ph], im);
      ph += bin;
      if (ph >= n_fft) ph -= n_fft;
    }
    out[(size_t)bin * ld_out + t] = log1pf(sqrtf(re * re + im * im));
  }
}

// utterance statistics over the F x T plane: stat[0] = sum, stat[1] = sum of squares (fp64 accumulators)
__global__ void __launch_bounds__(256) spec_stats_kernel(const float* __restrict__ x, int ld, int F, int T, double* stat) {
  __shared__ double rs[8], rq[8];
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)F * T; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[(i / T) * ld + i % T];
    s += v; q += (double)v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s += rs[i]; q += rq[i]; }
    atomicAdd(stat, s); atomicAdd(stat + 1, q);
  }
}
__global__ void __launch_bounds__(256) spec_norm_kernel(float* __restrict__ x, int ld, int F, int T, const double* stat) {
  const double n = (double)F * T, mean = stat[0] / n;
  const double var = (stat[1] - n * mean * mean) / (n - 1.0);       // torch.Tensor.std(): unbiased
  const float m = (float)mean, inv = (float)(1.0 / sqrt(var));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)F * T; i += (long long)gridDim.x * blockDim.x) {
    float* p = x + (i / T) * ld + i % T;
    *p = (*p - m) * inv;
  }
}

int k_spectrogram(const float* wav, int n, int n_fft, int hop, const float* window, float* out, int ld_out, int normalize,
                  double* stat2, cudaStream_t s) {
  MTL_REQUIRE(wav && window && out, "null argument");
  MTL_REQUIRE(n_fft >= 8 && n_fft <= SPEC_MAX_FFT && n_fft % 2 == 0 && hop >= 1 && n > n_fft / 2, "spectrogram geometry");
  const int T = 1 + n / hop, F = n_fft / 2 + 1;
  MTL_REQUIRE(ld_out >= T, "output row stride shorter than the number of frames");
  MTL_REQUIRE(!normalize || stat2, "normalisation needs the 2-double scratch");
  const size_t smem = sizeof(float) * (size_t)(2 + SPEC_FRAMES) * n_fft;
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(float) * (2 + SPEC_FRAMES) * SPEC_MAX_FFT)));
    configured = true;
  }
  spectrogram_kernel<<<mtl_cdiv(T, SPEC_FRAMES), 256, smem, s>>>(wav, n, n_fft, hop, window, out, ld_out, T);
  MTL_CHECK_LAUNCH();
  ++g_mtl_launches;
  if (normalize) {
    MTL_CHECK_CUDA(cudaMemsetAsync(stat2, 0, 2 * sizeof(double), s));
    const int g = (int)((((long long)F * T + 255) / 256) < 296 ? (((long long)F * T + 255) / 256) : 296);
    spec_stats_kernel<<<g, 256, 0, s>>>(out, ld_out, F, T, stat2);
    MTL_CHECK_LAUNCH();
    spec_norm_kernel<<<g, 256, 0, s>>>(out, ld_out, F, T, stat2);
    MTL_CHECK_LAUNCH();
    g_mtl_launches += 2;
  }
  return MTL_OK;
}
