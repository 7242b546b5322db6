// VGG front-end support kernels on NHWC activations ([B, F(freq), T(time), C]).
// Reference: models/asr/transformer.py:47-59 (Conv2d 3x3 s1 p1 + ReLU x2, MaxPool2d(2,2), twice) and
// :136-138 (view(B, C*F, T).transpose(1,2) -> feature index c*F'+f).
// The 64->64, 64->128, 128->128 convolutions run as GEMMs over (tap, channel)-ordered patches
// (im2col3x3 + weight re-layout here, contraction in gemm_*.cu); conv1 (C_in = 1, K = 9) is a
// direct memory-bound kernel.
#include "kernels.h"

// ------------------------------------------------------------------ conv1: 1 -> Cout, direct
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int B, int F, int T, int Cout) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sw[];               // [9][Cout] + bias[Cout]
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) { int co = i / 9, tap = i % 9; sw[tap * Cout + co] = w[i]; }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[9 * Cout + i] = bias[i];
  __syncthreads();
  const unsigned cg = (unsigned)Cout >> 2;
  const unsigned total = (unsigned)B * F * T * cg;      // < 2^31 (checked by the launcher): 32-bit divisions only
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cg) * 4;
    const unsigned p = i / cg;
    const unsigned row = p / (unsigned)T;
    const int t = (int)(p - row * (unsigned)T), f = (int)(row % (unsigned)F);
    const size_t img = (size_t)(row / (unsigned)F) * (size_t)F * T;
    float4 acc = *reinterpret_cast<const float4*>(sw + 9 * Cout + c4);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ff = f + kh - 1;
      if (ff < 0 || ff >= F) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tt = t + kw - 1;
        if (tt < 0 || tt >= T) continue;
        const float xv = x[img + (size_t)ff * T + tt];
        const float4 wv = *reinterpret_cast<const float4*>(sw + (kh * 3 + kw) * Cout + c4);
        acc.x += xv * wv.x; acc.y += xv * wv.y; acc.z += xv * wv.z; acc.w += xv * wv.w;
      }
    }
    acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    *reinterpret_cast<float4*>(out + (size_t)p * Cout + c4) = acc;
  }
}
// Cout == 64 (the model's conv.0): one warp per quarter of a (b, f) row, the two half-warps take alternate time steps, a
// lane owns 4 channels whose 36 weights + bias live in registers -- no shared memory, no division in the loop; every store
// is one coalesced float4 per lane (512 contiguous bytes per warp) and four pixels per thread are in flight.  Same
// accumulation order (bias, then taps in (kh, kw) order, FMA) as the generic kernel: bit-identical results.
__global__ void __launch_bounds__(128) conv1_fwd64_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, int B,
                                                          int F, int T) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TS = 4;
  const int unit = blockIdx.x * 4 + warp;
  const int row = unit / TS, seg = unit % TS;
  if (row >= B * F) return;
  const int tlen = (T + TS - 1) / TS, t0 = seg * tlen, t1 = min(T, t0 + tlen);
  const int half = lane >> 4, c4 = (lane & 15) * 4;
  float wr[9][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][j] = __ldg(w + (c4 + j) * 9 + tap);
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c4));
  const int f = row % F, b = row / F;
  const float* xr[3];
  bool rv[3];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ff = f + kh - 1;
    rv[kh] = ff >= 0 && ff < F;
    xr[kh] = x + ((size_t)b * F + (rv[kh] ? ff : f)) * T;
  }
  float* o = out + (size_t)row * T * 64 + c4;
  // the x taps of four pixels are fetched before the first FMA (explicit batches, see conv1_wgrad64_kernel)
  constexpr int U = 4;
  for (int t = t0 + half; t < t1; t += 2 * U) {
    float xv[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int tt = t + 2 * u + kw - 1;
          xv[u][kh * 3 + kw] = (t + 2 * u < t1 && rv[kh] && tt >= 0 && tt < T) ? __ldg(xr[kh] + tt) : 0.f;
        }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int tu = t + 2 * u;
      if (tu >= t1) break;
      float4 acc = b4;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int tt = tu + kw - 1;
          if (!rv[kh] || tt < 0 || tt >= T) continue;      // the generic kernel skips padded taps too (no +0 term)
          const float v = xv[u][kh * 3 + kw];
          acc.x += v * wr[kh * 3 + kw][0]; acc.y += v * wr[kh * 3 + kw][1];
          acc.z += v * wr[kh * 3 + kw][2]; acc.w += v * wr[kh * 3 + kw][3];
        }
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
      *reinterpret_cast<float4*>(o + (size_t)tu * 64) = acc;
    }
  }
}
int k_conv1_fwd(const float* x, const float* w, const float* b, float* out, int B, int F, int T, int Cout,
                cudaStream_t s) {
  MTL_REQUIRE(Cout % 4 == 0, "conv1 Cout % 4");
  if (Cout == 64 && B * F > 0 && T > 0 && (((uintptr_t)out) & 15u) == 0 && (((uintptr_t)b) & 15u) == 0) {
    MTL_CHECK_CUDA(mtl_launch_pdl(conv1_fwd64_kernel, dim3(B * F), dim3(128), 0, s, x, w, b, out, B, F, T));
    ++g_mtl_launches;
    return MTL_OK;
  }
  size_t total = (size_t)B * F * T * (Cout / 4);
  if (!total) return MTL_OK;
  MTL_REQUIRE(total < (1ull << 31), "conv1: B*F*T*Cout/4 must stay below 2^31");
  int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
  MTL_CHECK_CUDA(mtl_launch_pdl(conv1_fwd_kernel, dim3(grid), dim3(256), (size_t)10 * Cout * sizeof(float), s, x, w, b, out, B, F, T, Cout));
  ++g_mtl_launches;
  return MTL_OK;
}

// dw[co,tap] += sum_p dout[p,co] * x[p+tap] ; db[co] += sum_p dout[p,co].   dout is the gradient
// w.r.t. the PRE-ReLU output (relu mask already applied by the caller).
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                          float* __restrict__ dw, float* __restrict__ db, int B,
                                                          int F, int T, int Cout, int pix_per_block) {
  extern __shared__ float red[];              // [lanes][Cout][10]
  const int lanes = blockDim.x / Cout;
  const int co = threadIdx.x % Cout, sub = threadIdx.x / Cout;
  const size_t P = (size_t)B * F * T;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  const size_t p1 = p0 + pix_per_block < P ? p0 + pix_per_block : P;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  if (sub < lanes) {
    for (size_t p = p0 + sub; p < p1; p += lanes) {
      const float g = dout[p * Cout + co];
      const int t = (int)(p % T), f = (int)((p / T) % F);
      const size_t img = (p / ((size_t)F * T)) * (size_t)F * T;
      acc[9] += g;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ff = f + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int tt = t + kw - 1;
          const bool ok = ff >= 0 && ff < F && tt >= 0 && tt < T;
          const float xv = ok ? x[img + (size_t)ff * T + tt] : 0.f;
          acc[kh * 3 + kw] += g * xv;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) red[((size_t)sub * Cout + co) * 10 + i] = acc[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 10; i += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(size_t)l * Cout * 10 + i];
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, t); else atomicAdd(db + c, t);
  }
}
// Cout == 64 (the model's conv.0): one warp per quarter of a (b, f) row, the two half-warps take alternate time steps, a lane owns 4
// channels -- every dout load is one coalesced float4 (a whole pixel per half-warp), the nine x taps are L1 broadcasts,
// no division in the loop.  33 MB of dout are read once; the block's 640 partial sums leave through atomics.
__global__ void __launch_bounds__(128) conv1_wgrad64_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                            float* __restrict__ dw, float* __restrict__ db, int B, int F,
                                                            int T) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[4][640];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TS = 4;                                     // time segments per row: enough warps in flight to cover HBM latency
  const int unit = blockIdx.x * 4 + warp;
  const int row = unit / TS, seg = unit % TS;               // (b, f) row, quarter of its time axis
  const int tlen = (T + TS - 1) / TS, t0 = seg * tlen, t1 = min(T, t0 + tlen);
  const int half = lane >> 4, c4 = (lane & 15) * 4;
  float acc[10][4];
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (row < B * F) {
    const int f = row % F, b = row / F;
    const float* xr[3];
    bool rv[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ff = f + kh - 1;
      rv[kh] = ff >= 0 && ff < F;
      xr[kh] = x + ((size_t)b * F + (rv[kh] ? ff : f)) * T;
    }
    const float* g = dout + (size_t)row * T * 64 + c4;
    // Four pixels per thread are fetched before the first FMA (explicit batches: with a plain `#pragma unroll` the compiler
    // kept load -> FMA -> load order, i.e. ONE 512-byte request per warp in flight and 0.9 TB/s; ncu, profiles/r02_d_*)
    constexpr int U = 4;
    for (int t = t0 + half; t < t1; t += 2 * U) {
      float4 gv[U];
      float xv[U][9];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int tu = t + 2 * u;
        const bool ok = tu < t1;
        gv[u] = ok ? __ldg(reinterpret_cast<const float4*>(g + (size_t)tu * 64)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int tt = tu + kw - 1;
            xv[u][kh * 3 + kw] = (ok && rv[kh] && tt >= 0 && tt < T) ? __ldg(xr[kh] + tt) : 0.f;
          }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[9][0] += gv[u].x; acc[9][1] += gv[u].y; acc[9][2] += gv[u].z; acc[9][3] += gv[u].w;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          acc[k][0] += xv[u][k] * gv[u].x; acc[k][1] += xv[u][k] * gv[u].y;
          acc[k][2] += xv[u][k] * gv[u].z; acc[k][3] += xv[u][k] * gv[u].w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
      if (lane < 16) red[warp][(c4 + j) * 10 + i] = acc[i][j];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 640; i += 128) {
    const float t = red[0][i] + red[1][i] + red[2][i] + red[3][i];
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, t); else atomicAdd(db + c, t);
  }
}
int k_conv1_wgrad(const float* x, const float* dout, float* dw, float* db, int B, int F, int T, int Cout,
                  cudaStream_t s) {
  if (Cout == 64 && B * F > 0 && T > 0 && (((uintptr_t)dout) & 15u) == 0) {
    MTL_CHECK_CUDA(mtl_launch_pdl(conv1_wgrad64_kernel, dim3(B * F), dim3(128), 0, s, x, dout, dw, db, B, F, T));   // 4 warps = the 4 time segments of one row
    ++g_mtl_launches;
    return MTL_OK;
  }
  MTL_REQUIRE(Cout <= 256 && 256 % Cout == 0, "conv1 Cout must divide 256");
  size_t P = (size_t)B * F * T;
  if (!P) return MTL_OK;
  int ppb = 512;
  int grid = mtl_cdiv((long long)P, ppb);
  size_t smem = (size_t)(256 / Cout) * Cout * 10 * sizeof(float);
  conv1_wgrad_kernel<<<grid, 256, smem, s>>>(x, dout, dw, db, B, F, T, Cout, ppb);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

// ------------------------------------------------------------------ im2col (3x3, pad 1) on NHWC
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float4* __restrict__ x, float4* __restrict__ col,
                                                        int B, int F, int T, int C4) {
  const size_t total = (size_t)B * F * T * 9 * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const int tap = (int)((i / C4) % 9);
    const size_t p = i / ((size_t)9 * C4);
    const int t = (int)(p % T), f = (int)((p / T) % F);
    const size_t bimg = p / ((size_t)F * T);
    const int ff = f + tap / 3 - 1, tt = t + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ff >= 0 && ff < F && tt >= 0 && tt < T) v = x[((bimg * F + ff) * T + tt) * C4 + c];
    col[i] = v;
  }
}
int k_im2col3x3(const float* x, float* col, int B, int F, int T, int C, cudaStream_t s) {
  MTL_REQUIRE(C % 4 == 0, "im2col C % 4");
  size_t total = (size_t)B * F * T * 9 * (C / 4);
  if (!total) return MTL_OK;
  size_t g = (total + 255) / 256; if (g > 148 * 32) g = 148 * 32;
  im2col3x3_kernel<<<(int)g, 256, 0, s>>>((const float4*)x, (float4*)col, B, F, T, C / 4);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

// ------------------------------------------------------------------ weight layouts
// 3xTF32 operand halves (same rounding as the in-kernel splitter of gemm_tc.cu): hi = rna_tf32(v), lo = rna_tf32(v - hi)
__device__ __forceinline__ float tf32_rna_f(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void store_w(float* __restrict__ dst, int i, int n, float v, int split) {
  if (!split) { dst[i] = v; return; }
  const float h = tf32_rna_f(v);
  dst[i] = h;
  dst[n + i] = tf32_rna_f(v - h);
}
__global__ void conv_w_fwd_layout_kernel(const float* __restrict__ w, float* __restrict__ wg, int Cout, int Cin, int split) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;          // over wg [Cout][9][Cin]
  if (i >= Cout * 9 * Cin) return;
  int ci = i % Cin, tap = (i / Cin) % 9, co = i / (9 * Cin);
  store_w(wg, i, Cout * 9 * Cin, w[((size_t)co * Cin + ci) * 9 + tap], split);
}
__global__ void conv_w_dgrad_layout_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin, int split) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;          // over wd [Cin][9][Cout]
  if (i >= Cout * 9 * Cin) return;
  int co = i % Cout, tap = (i / Cout) % 9, ci = i / (9 * Cout);
  store_w(wd, i, Cout * 9 * Cin, w[((size_t)co * Cin + ci) * 9 + (8 - tap)], split);       // flipped taps: (2-kh, 2-kw)
}
__global__ void conv_wgrad_scatter_kernel(const float* __restrict__ dwg, float* __restrict__ dw, int Cout, int Cin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;          // over dw [Cout][Cin][9]
  if (i >= Cout * 9 * Cin) return;
  int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
  dw[i] += dwg[((size_t)co * 9 + tap) * Cin + ci];
}
__global__ void conv_wgrad_scatter_t_kernel(const float* __restrict__ dwgT, float* __restrict__ dw, int Cout, int Cin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;          // over dw [Cout][Cin][9]
  if (i >= Cout * 9 * Cin) return;
  int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
  dw[i] += dwgT[((size_t)tap * Cin + ci) * Cout + co];
}
int k_conv_wgrad_scatter_t(const float* dwgT, float* dw, int Cout, int Cin, cudaStream_t s) {
  conv_wgrad_scatter_t_kernel<<<mtl_cdiv(Cout * 9 * Cin, 256), 256, 0, s>>>(dwgT, dw, Cout, Cin);
  MTL_CHECK_LAUNCH(); return MTL_OK;
}
int k_conv_w_fwd_layout(const float* w, float* wg, int Cout, int Cin, int split, cudaStream_t s) {
  conv_w_fwd_layout_kernel<<<mtl_cdiv(Cout * 9 * Cin, 256), 256, 0, s>>>(w, wg, Cout, Cin, split);
  MTL_CHECK_LAUNCH(); return MTL_OK;
}
int k_conv_w_dgrad_layout(const float* w, float* wd, int Cout, int Cin, int split, cudaStream_t s) {
  conv_w_dgrad_layout_kernel<<<mtl_cdiv(Cout * 9 * Cin, 256), 256, 0, s>>>(w, wd, Cout, Cin, split);
  MTL_CHECK_LAUNCH(); return MTL_OK;
}
int k_conv_wgrad_scatter(const float* dwg, float* dw, int Cout, int Cin, cudaStream_t s) {
  conv_wgrad_scatter_kernel<<<mtl_cdiv(Cout * 9 * Cin, 256), 256, 0, s>>>(dwg, dw, Cout, Cin);
  MTL_CHECK_LAUNCH(); return MTL_OK;
}

// ------------------------------------------------------------------ 2x2 max-pool (floor) on NHWC
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ out,
                                                           int B, int F, int T, int C4) {
  pdl_wait();
  pdl_trigger();
  const int F2 = F / 2, T2 = T / 2;
  const size_t total = (size_t)B * F2 * T2 * C4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const size_t p = i / C4;
    const int t2 = (int)(p % T2), f2 = (int)((p / T2) % F2);
    const size_t b = p / ((size_t)F2 * T2);
    const size_t base = ((b * F + 2 * f2) * T + 2 * t2) * C4 + c;
    float4 a = x[base], bb = x[base + C4], cc = x[base + (size_t)T * C4], d = x[base + (size_t)T * C4 + C4];
    float4 r;
    r.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)); r.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
    r.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)); r.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
    out[i] = r;
  }
}
int k_maxpool2_fwd(const float* x, float* out, int B, int F, int T, int C, cudaStream_t s) {
  MTL_REQUIRE(C % 4 == 0, "pool C % 4");
  size_t total = (size_t)B * (F / 2) * (T / 2) * (C / 4);
  if (!total) return MTL_OK;
  size_t g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16;
  MTL_CHECK_CUDA(mtl_launch_pdl(maxpool2_fwd_kernel, dim3((unsigned)g), dim3(256), 0, s, (const float4*)x, (float4*)out, B, F, T, C / 4));
  ++g_mtl_launches;
  return MTL_OK;
}
// Gradient of relu -> maxpool w.r.t. the pre-ReLU conv output, given x = post-ReLU activation.
// The window's gradient goes to its first maximum in scan order (PyTorch max_pool2d picks the
// first index whose value is strictly greater); relu'(x) = [x > 0].
__device__ __forceinline__ float pool_route(float v00, float v01, float v10, float v11, int me, float g) {
  float m = v00; int am = 0;
  if (v01 > m) { m = v01; am = 1; }
  if (v10 > m) { m = v10; am = 2; }
  if (v11 > m) { m = v11; am = 3; }
  float mine = me == 0 ? v00 : (me == 1 ? v01 : (me == 2 ? v10 : v11));
  return (am == me && mine > 0.f) ? g : 0.f;
}
// One thread per 2x2 window and channel quad: the four x values are read once, the four dx values written once (the
// element-per-thread version read every window four times).  Windows that floor pooling drops (odd last row / column)
// get zeros.
__global__ void __launch_bounds__(256) maxpool2_relu_bwd_kernel(const float4* __restrict__ x,
                                                                const float4* __restrict__ dpool,
                                                                float4* __restrict__ dx, int B, int F, int T, int C4) {
  pdl_wait();
  pdl_trigger();
  const int F2 = F / 2, T2 = T / 2, Fw = (F + 1) / 2, Tw = (T + 1) / 2;
  const unsigned total = (unsigned)B * Fw * Tw * C4;      // < 2^31 (checked by the launcher): 32-bit divisions only
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)C4);
    const unsigned p = i / (unsigned)C4;
    const unsigned prow = p / (unsigned)Tw;
    const int t2 = (int)(p - prow * (unsigned)Tw), f2 = (int)(prow % (unsigned)Fw);
    const size_t b = prow / (unsigned)Fw;
    const int f = 2 * f2, t = 2 * t2;
    const bool hf = f + 1 < F, ht = t + 1 < T;            // second row / column of the window exists
    const size_t base = ((b * F + f) * T + t) * C4 + c, down = (size_t)T * C4;
    if (f2 < F2 && t2 < T2) {
      const float4 a = x[base], bb = x[base + C4], cc = x[base + down], d = x[base + down + C4];
      const float4 g = dpool[((b * F2 + f2) * T2 + t2) * C4 + c];
      float4 r0, r1, r2, r3;
      r0.x = pool_route(a.x, bb.x, cc.x, d.x, 0, g.x); r0.y = pool_route(a.y, bb.y, cc.y, d.y, 0, g.y);
      r0.z = pool_route(a.z, bb.z, cc.z, d.z, 0, g.z); r0.w = pool_route(a.w, bb.w, cc.w, d.w, 0, g.w);
      r1.x = pool_route(a.x, bb.x, cc.x, d.x, 1, g.x); r1.y = pool_route(a.y, bb.y, cc.y, d.y, 1, g.y);
      r1.z = pool_route(a.z, bb.z, cc.z, d.z, 1, g.z); r1.w = pool_route(a.w, bb.w, cc.w, d.w, 1, g.w);
      r2.x = pool_route(a.x, bb.x, cc.x, d.x, 2, g.x); r2.y = pool_route(a.y, bb.y, cc.y, d.y, 2, g.y);
      r2.z = pool_route(a.z, bb.z, cc.z, d.z, 2, g.z); r2.w = pool_route(a.w, bb.w, cc.w, d.w, 2, g.w);
      r3.x = pool_route(a.x, bb.x, cc.x, d.x, 3, g.x); r3.y = pool_route(a.y, bb.y, cc.y, d.y, 3, g.y);
      r3.z = pool_route(a.z, bb.z, cc.z, d.z, 3, g.z); r3.w = pool_route(a.w, bb.w, cc.w, d.w, 3, g.w);
      dx[base] = r0; dx[base + C4] = r1; dx[base + down] = r2; dx[base + down + C4] = r3;
    } else {
      dx[base] = zero;
      if (ht) dx[base + C4] = zero;
      if (hf) dx[base + down] = zero;
      if (hf && ht) dx[base + down + C4] = zero;
    }
  }
}
int k_maxpool2_relu_bwd(const float* x, const float* dpool, float* dx, int B, int F, int T, int C, cudaStream_t s) {
  MTL_REQUIRE(C % 4 == 0, "pool C % 4");
  size_t total = (size_t)B * ((F + 1) / 2) * ((T + 1) / 2) * (C / 4);
  if (!total) return MTL_OK;
  MTL_REQUIRE((size_t)B * F * T * (C / 4) < (1ull << 31), "maxpool bwd: B*F*T*C/4 must stay below 2^31");
  size_t g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16;
  MTL_CHECK_CUDA(mtl_launch_pdl(maxpool2_relu_bwd_kernel, dim3((unsigned)g), dim3(256), 0, s, (const float4*)x, (const float4*)dpool, (float4*)dx, B, F, T, C / 4));
  ++g_mtl_launches;
  return MTL_OK;
}

// ------------------------------------------------------------------ (B,F4,T4,C) <-> (B,T4,C*F4)
// One block per (b, t, 32-channel chunk): the F4 x 32 slice p4[b, :, t, c0:c0+32] (128-byte rows, T4*C floats apart) is
// transposed through shared memory into the contiguous run feat[b, t, c0*F4 : (c0+32)*F4] (feature index c*F4 + f) --
// both global sides move whole 128-byte lines (the element-per-thread version read with a stride of T4*C floats), every
// thread has all of its loads in flight before the first store.  The backward is the same tile read the other way.
constexpr int FT_CH = 32, FT_MAX_PER = 8;                    // channels per block; elements per thread (F4 <= 64)
template <bool BWD>
__global__ void __launch_bounds__(256) feat_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int B, int F4, int T4, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[64 * (FT_CH + 1)];                   // [F4][33]
  const int chunks = C / FT_CH;
  const int cc = blockIdx.x % chunks, bt = blockIdx.x / chunks, b = bt / T4, t = bt - b * T4;
  const int n = F4 * FT_CH;
  // NHWC side: element (f, c) at p4[((b*F4 + f)*T4 + t)*C + cc*32 + c]; feature side: feat[bt*C*F4 + (cc*32 + c)*F4 + f]
  const size_t nhwc0 = ((size_t)b * F4 * T4 + t) * C + cc * FT_CH, feat0 = (size_t)bt * C * F4 + (size_t)cc * FT_CH * F4;
  const size_t fstride = (size_t)T4 * C;
  float v[FT_MAX_PER];
  if (!BWD) {
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) v[k] = src[nhwc0 + (size_t)(i >> 5) * fstride + (i & 31)];
    }
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) tile[(i >> 5) * (FT_CH + 1) + (i & 31)] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) { const int c = i / F4, f = i - c * F4; dst[feat0 + i] = tile[f * (FT_CH + 1) + c]; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) v[k] = src[feat0 + i];
    }
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) { const int c = i / F4, f = i - c * F4; tile[f * (FT_CH + 1) + c] = v[k]; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FT_MAX_PER; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n) dst[nhwc0 + (size_t)(i >> 5) * fstride + (i & 31)] = tile[(i >> 5) * (FT_CH + 1) + (i & 31)];
    }
  }
}
// generic fallback (C % 32 != 0 or F4 > 64): one element per thread
__global__ void __launch_bounds__(256) feat_transpose_generic_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                     int B, int F4, int T4, int C, int bwd) {
  const size_t total = (size_t)B * T4 * C * F4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F4), c = (int)((i / F4) % C);
    const size_t bt = i / ((size_t)F4 * C);
    const int t = (int)(bt % T4);
    const size_t b = bt / T4;
    const size_t j = ((b * F4 + f) * T4 + t) * C + c;
    if (bwd) dst[j] = src[i]; else dst[i] = src[j];
  }
}
static int feat_transpose_launch(bool bwd, const float* src, float* dst, int B, int F4, int T4, int C, cudaStream_t s) {
  const size_t total = (size_t)B * T4 * C * F4;
  if (total == 0) return MTL_OK;
  if (C % FT_CH == 0 && F4 <= 64) {
    const dim3 grid((unsigned)((size_t)B * T4 * (C / FT_CH)));
    if (bwd) MTL_CHECK_CUDA(mtl_launch_pdl(feat_transpose_kernel<true>, grid, dim3(256), 0, s, src, dst, B, F4, T4, C));
    else MTL_CHECK_CUDA(mtl_launch_pdl(feat_transpose_kernel<false>, grid, dim3(256), 0, s, src, dst, B, F4, T4, C));
    ++g_mtl_launches;
    return MTL_OK;
  }
  size_t g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16;
  feat_transpose_generic_kernel<<<(unsigned)g, 256, 0, s>>>(src, dst, B, F4, T4, C, bwd ? 1 : 0);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
int k_feat_transpose(const float* p4, float* feat, int B, int F4, int T4, int C, cudaStream_t s) {
  return feat_transpose_launch(false, p4, feat, B, F4, T4, C, s);
}
int k_feat_transpose_bwd(const float* dfeat, float* dp4, int B, int F4, int T4, int C, cudaStream_t s) {
  return feat_transpose_launch(true, dfeat, dp4, B, F4, T4, C, s);
}
