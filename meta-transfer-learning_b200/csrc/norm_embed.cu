// LayerNorm(+dropout+residual+row-mask+PE) forward/backward, column sums, embedding gather/scatter,
// decoder preprocessing and mask builders.  Warp-per-row, shuffle reductions.
// Reference: modules/common_layers.py:129-131,303-304 (LN(dropout(branch)+x)), modules/encoder.py:72-73,
// 101-104 (stem LN + PE, "*= non_pad_mask"), modules/decoder.py:55-69,86-96,314-321.
#include "kernels.h"

#define LN_MAXV 32   // d <= 1024
#define LN_EPS 1e-5f

// One warp per row; each lane owns float4 column groups (4*(lane + 32 i) ..), so one Philox block
// serves the 4 elements of a group.
__global__ void __launch_bounds__(128) ln_fwd_kernel(
    const float* __restrict__ y, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ rowmask, const float* __restrict__ pe,
    int pe_period, MtlDrop drop, float* __restrict__ out, float* __restrict__ xhat,
    float* __restrict__ rstd_out, int M, int d) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  constexpr int NV = LN_MAXV / 4;
  const int d4 = d >> 2;
  float4 z[NV];
  float sum = 0.f;
  const size_t base = (size_t)row * d;
  const float4* y4 = reinterpret_cast<const float4*>(y + base);
  const float4* r4 = res ? reinterpret_cast<const float4*>(res + base) : nullptr;
  const unsigned long long seed = drop.p > 0.f ? mtl_eff_seed(drop) : 0ull;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d4) {
      v = y4[c];
      if (drop.p > 0.f) {
        const float4 sc = dropout_scale4(seed, drop.site, base + 4ull * c, drop.p, drop.inv_keep);
        v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
      }
      if (r4) { const float4 r = r4[c]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    }
    z[i] = v;
    sum += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(sum) / (float)d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + i * 32 < d4) {
      const float a = z[i].x - mean, b = z[i].y - mean, c = z[i].z - mean, e = z[i].w - mean;
      var += (a * a + b * b) + (c * c + e * e);
    }
  }
  var = warp_sum(var) / (float)d;
  const float rstd = rsqrtf(var + LN_EPS);
  const float rm = rowmask ? rowmask[row] : 1.f;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* p4 = pe ? reinterpret_cast<const float4*>(pe + (size_t)(row % pe_period) * d) : nullptr;
  float4* xh4 = reinterpret_cast<float4*>(xhat + base);
  float4* o4 = reinterpret_cast<float4*>(out + base);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < d4) {
      const float4 g = g4[c], b = b4[c];
      float4 xh, o;
      xh.x = (z[i].x - mean) * rstd; xh.y = (z[i].y - mean) * rstd;
      xh.z = (z[i].z - mean) * rstd; xh.w = (z[i].w - mean) * rstd;
      o.x = xh.x * g.x + b.x; o.y = xh.y * g.y + b.y; o.z = xh.z * g.z + b.z; o.w = xh.w * g.w + b.w;
      if (p4) { const float4 pp = p4[c]; o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w; }
      o.x *= rm; o.y *= rm; o.z *= rm; o.w *= rm;
      xh4[c] = xh;
      o4[c] = o;
    }
  }
  if (lane == 0) rstd_out[row] = rstd;
}

int k_ln_fwd(const float* y, const float* res, const float* gamma, const float* beta, const float* rowmask,
             const float* pe, int pe_period, MtlDrop drop, float* out, float* xhat, float* rstd, int M,
             int d, cudaStream_t s) {
  MTL_REQUIRE(d <= 32 * LN_MAXV && d % 4 == 0, "layer norm width must be a multiple of 4, <= 1024");
  if (M == 0) return MTL_OK;
  MTL_CHECK_CUDA(mtl_launch_pdl(ln_fwd_kernel, dim3(mtl_cdiv(M, 4)), dim3(128), 0, s, y, res, gamma, beta, rowmask, pe,
                                pe_period > 0 ? pe_period : 1, drop, out, xhat, rstd, M, d));
  ++g_mtl_launches;
  return MTL_OK;
}

__global__ void __launch_bounds__(128) ln_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ xhat, const float* __restrict__ rstd,
    const float* __restrict__ gamma, const float* __restrict__ rowmask, MtlDrop drop,
    float* __restrict__ dy, float* __restrict__ dres, int dres_acc, int M, int d) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  constexpr int NV = LN_MAXV / 4;
  const int d4 = d >> 2;
  const size_t base = (size_t)row * d;
  const float rm = rowmask ? rowmask[row] : 1.f;
  const float4* do4 = reinterpret_cast<const float4*>(dout + base);
  const float4* xh4 = reinterpret_cast<const float4*>(xhat + base);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 dxh[NV], xh[NV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (c < d4) {
      const float4 g = g4[c];
      a = do4[c]; b = xh4[c];
      a.x *= rm * g.x; a.y *= rm * g.y; a.z *= rm * g.z; a.w *= rm * g.w;
    }
    dxh[i] = a; xh[i] = b;
    s1 += (a.x + a.y) + (a.z + a.w);
    s2 += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
  s1 = warp_sum(s1) / (float)d;
  s2 = warp_sum(s2) / (float)d;
  const float r = rstd[row];
  const unsigned long long seed = (dy && drop.p > 0.f) ? mtl_eff_seed(drop) : 0ull;
  float4* dr4 = dres ? reinterpret_cast<float4*>(dres + base) : nullptr;
  float4* dy4 = dy ? reinterpret_cast<float4*>(dy + base) : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < d4) {
      float4 dz;
      dz.x = r * (dxh[i].x - s1 - xh[i].x * s2); dz.y = r * (dxh[i].y - s1 - xh[i].y * s2);
      dz.z = r * (dxh[i].z - s1 - xh[i].z * s2); dz.w = r * (dxh[i].w - s1 - xh[i].w * s2);
      if (dr4) {
        float4 o = dz;
        if (dres_acc) { const float4 p = dr4[c]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        dr4[c] = o;
      }
      if (dy4) {
        float4 o = dz;
        if (drop.p > 0.f) {
          const float4 sc = dropout_scale4(seed, drop.site, base + 4ull * c, drop.p, drop.inv_keep);
          o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
        }
        dy4[c] = o;
      }
    }
  }
}

// dgamma[c] += sum_m dout*rm*xhat ; dbeta[c] += sum_m dout*rm.  Block = 32 columns x 8 row-lanes.
__global__ void __launch_bounds__(256) ln_param_grad_kernel(
    const float* __restrict__ dout, const float* __restrict__ xhat, const float* __restrict__ rowmask,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int d, int rows_per_slab) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sg[8][33], sb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
  float ag = 0.f, ab = 0.f;
  if (c < d) {
#pragma unroll 4
    for (int m = m0 + ty; m < m1; m += 8) {
      float rm = rowmask ? rowmask[m] : 1.f;
      float g = dout[(size_t)m * d + c] * rm;
      ag += g * xhat[(size_t)m * d + c];
      ab += g;
    }
  }
  sg[ty][tx] = ag; sb[ty][tx] = ab;
  __syncthreads();
  if (ty == 0 && c < d) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g += sg[i][tx]; b += sb[i][tx]; }
    if (gridDim.y == 1) { dgamma[c] += g; dbeta[c] += b; }
    else { atomicAdd(dgamma + c, g); atomicAdd(dbeta + c, b); }
  }
}

int k_ln_bwd(const float* dout, const float* xhat, const float* rstd, const float* gamma, const float* rowmask,
             MtlDrop drop, float* dy, float* dres, int dres_accumulate, float* dgamma, float* dbeta, int M,
             int d, cudaStream_t s) {
  MTL_REQUIRE(d <= 32 * LN_MAXV && d % 4 == 0, "layer norm width must be a multiple of 4, <= 1024");
  if (M == 0) return MTL_OK;
  MTL_CHECK_CUDA(mtl_launch_pdl(ln_bwd_kernel, dim3(mtl_cdiv(M, 4)), dim3(128), 0, s, dout, xhat, rstd, gamma, rowmask, drop,
                                dy, dres, dres_accumulate, M, d));
  ++g_mtl_launches;
  if (!dgamma && !dbeta) return MTL_OK;             // the caller runs k_ln_param_grad itself (off the activation chain)
  return k_ln_param_grad(dout, xhat, rowmask, dgamma, dbeta, M, d, s);
}
int k_ln_param_grad(const float* dout, const float* xhat, const float* rowmask, float* dgamma, float* dbeta, int M, int d,
                    cudaStream_t s) {
  MTL_REQUIRE(dgamma && dbeta, "null argument");
  if (M == 0) return MTL_OK;
  // a (32 columns x 8 row-lanes) block walks its rows serially: slabs of 32 rows keep that walk 4 loads deep
  // (M = 264: 16 x 9 blocks instead of 16 blocks doing 33 dependent-latency steps each)
  const int slabs = M > 32 ? (mtl_cdiv(M, 32) < 512 ? mtl_cdiv(M, 32) : 512) : 1;
  MTL_CHECK_CUDA(mtl_launch_pdl(ln_param_grad_kernel, dim3(mtl_cdiv(d, 32), slabs), dim3(256), 0, s, dout, xhat, rowmask, dgamma,
                                dbeta, M, d, mtl_cdiv(M, slabs)));
  ++g_mtl_launches;
  return MTL_OK;
}

// out[n] += sum_m x[m*ld + n].  Rows are split over gridDim.y slabs when M is large (atomic merge).
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int M, int N, int ld,
                                                     float* __restrict__ out, int rows_per_slab) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per_slab;
  const int m1 = min(M, m0 + rows_per_slab);
  float a = 0.f;
  if (c < N)
    for (int m = m0 + ty; m < m1; m += 8) a += x[(size_t)m * ld + c];
  sm[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][tx];
    if (gridDim.y == 1) out[c] += t; else atomicAdd(out + c, t);
  }
}
int k_colsum_acc(const float* x, int M, int N, int ld, float* out, cudaStream_t s) {
  if (M == 0 || N == 0) return MTL_OK;
  // slabs of ~32 rows (4 loads per thread) merged with atomics; large M keeps at most 512 slabs per column block
  int slabs = M > 32 ? mtl_cdiv(M, 32) : 1;
  const int max_slabs = N >= 2048 ? 64 : 512;
  if (slabs > max_slabs) slabs = max_slabs;
  int rps = mtl_cdiv(M, slabs);
  MTL_CHECK_CUDA(mtl_launch_pdl(colsum_kernel, dim3(mtl_cdiv(N, 32), slabs), dim3(256), 0, s, x, M, N, ld, out, rps));
  ++g_mtl_launches;
  return MTL_OK;
}

// ---------------------------------------------------------------- embedding (+PE, dropout)
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int* __restrict__ tok, const float* __restrict__ E,
                                                        const float* __restrict__ pe, MtlDrop drop,
                                                        float* __restrict__ out, int rows, int n, int d) {
  pdl_wait();
  pdl_trigger();
  size_t total = (size_t)rows * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int row = (int)(i / d), c = (int)(i % d);
    int pos = row % n;
    float v = E[(size_t)tok[row] * d + c] * 1.0f + pe[(size_t)pos * d + c];   // x_logit_scale = 1 (decoder.py:53)
    if (drop.p > 0.f) v *= dropout_scale(mtl_eff_seed(drop), drop.site, i, drop.p, drop.inv_keep);
    out[i] = v;
  }
}
int k_embed_fwd(const int* tok, const float* E, const float* pe, MtlDrop drop, float* out, int B, int n, int d,
                cudaStream_t s) {
  size_t total = (size_t)B * n * d;
  if (!total) return MTL_OK;
  int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
  MTL_CHECK_CUDA(mtl_launch_pdl(embed_fwd_kernel, dim3(grid), dim3(256), 0, s, tok, E, pe, drop, out, B * n, n, d));
  ++g_mtl_launches;
  return MTL_OK;
}
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int* __restrict__ tok, const float* __restrict__ dout,
                                                        MtlDrop drop, float* __restrict__ dE, int rows, int d,
                                                        int pad_id) {
  pdl_wait();
  pdl_trigger();
  size_t total = (size_t)rows * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int row = (int)(i / d), c = (int)(i % d);
    int t = tok[row];
    if (t == pad_id) continue;                       // nn.Embedding(padding_idx=PAD) (decoder.py:39)
    float g = dout[i];
    if (drop.p > 0.f) g *= dropout_scale(mtl_eff_seed(drop), drop.site, i, drop.p, drop.inv_keep);
    atomicAdd(dE + (size_t)t * d + c, g);
  }
}
int k_embed_bwd(const int* tok, const float* dout, MtlDrop drop, float* dE, int B, int n, int d, int pad_id,
                cudaStream_t s) {
  size_t total = (size_t)B * n * d;
  if (!total) return MTL_OK;
  int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
  MTL_CHECK_CUDA(mtl_launch_pdl(embed_bwd_kernel, dim3(grid), dim3(256), 0, s, tok, dout, drop, dE, B * n, d, pad_id));
  ++g_mtl_launches;
  return MTL_OK;
}

// ---------------------------------------------------------------- decoder preprocess + masks
// Decoder.preprocess (decoder.py:55-69): strip PAD(0); seq_in = <SOS> y, padded with EOS;
// seq_out = y <EOS>, padded with PAD.  non_pad = seq_in != EOS (:86); key-pad = seq_in == EOS (:87).
__global__ void dec_preprocess_kernel(const long long* __restrict__ trg, int B, int L, int n, int* seq_in,
                                      int* seq_out, float* rowmask, unsigned char* keypad, int* overflow) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int PAD = 0, SOS = 1, EOS = 2;
  int* si = seq_in + (size_t)b * n;
  int* so = seq_out + (size_t)b * n;
  for (int i = 0; i < n; ++i) { si[i] = EOS; so[i] = PAD; }
  si[0] = SOS;
  int cnt = 0;
  for (int j = 0; j < L; ++j) {
    int t = (int)trg[(size_t)b * L + j];
    if (t == PAD) continue;
    if (cnt + 1 < n) { si[cnt + 1] = t; so[cnt] = t; }
    ++cnt;
  }
  if (cnt < n) so[cnt] = EOS; else if (overflow) atomicExch(overflow, 1);
  for (int i = 0; i < n; ++i) {
    bool e = si[i] == EOS;
    rowmask[(size_t)b * n + i] = e ? 0.f : 1.f;
    keypad[(size_t)b * n + i] = e ? 1 : 0;
  }
}
int k_dec_preprocess(const long long* trg, int B, int L, int n, int* seq_in, int* seq_out, float* rowmask,
                     unsigned char* keypad, int* overflow_flag, cudaStream_t s) {
  if (B == 0) return MTL_OK;
  dec_preprocess_kernel<<<mtl_cdiv(B, 64), 64, 0, s>>>(trg, B, L, n, seq_in, seq_out, rowmask, keypad, overflow_flag);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
__global__ void enc_masks_kernel(const int* __restrict__ lens, int B, int Tp, float* rowmask, unsigned char* keypad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Tp) return;
  int b = i / Tp, t = i % Tp;
  bool valid = t < lens[b];
  rowmask[i] = valid ? 1.f : 0.f;
  keypad[i] = valid ? 0 : 1;
}
int k_enc_masks(const int* lens, int B, int Tp, float* rowmask, unsigned char* keypad, cudaStream_t s) {
  if (B * Tp == 0) return MTL_OK;
  enc_masks_kernel<<<mtl_cdiv(B * Tp, 256), 256, 0, s>>>(lens, B, Tp, rowmask, keypad);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
