// Fused log-softmax + NLL (+label smoothing) + top-1 index over the vocabulary, and its backward.
// One CTA per (b, position) row; online max/sum so the log-probabilities are never materialised.
// Reference: utils/metrics.py:96-126 (F.cross_entropy ignore_index=PAD, mean over non-pad rows;
// label-smoothing branch :113-124), utils/metrics.py:83-89 (token accuracy),
// models/asr/transformer.py:146 (hyp = topk(pred,1)).
#include "kernels.h"
#include <float.h>

__global__ void __launch_bounds__(256) ce_row_kernel(const float* __restrict__ logits, int ld,
                                                     const int* __restrict__ gold, int V, float smoothing,
                                                     int pad_id, float* __restrict__ row_lse,
                                                     float* __restrict__ row_loss, int* __restrict__ hyp) {
  __shared__ float s_m[8], s_s[8], s_x[8];
  __shared__ int s_i[8];
  const int row = blockIdx.x;
  const float* x = logits + (size_t)row * ld;
  float m = -FLT_MAX, ssum = 0.f, xsum = 0.f;
  int am = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float t = x[v];
    xsum += t;
    if (t > m) { ssum = ssum * __expf(m - t) + 1.f; m = t; am = v; }
    else ssum += __expf(t - m);
  }
  // warp reduce (max, argmax lowest index on ties, rescaled sum)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    float s2 = __shfl_xor_sync(0xffffffffu, ssum, o);
    int a2 = __shfl_xor_sync(0xffffffffu, am, o);
    xsum += __shfl_xor_sync(0xffffffffu, xsum, o);
    float mn = fmaxf(m, m2);
    ssum = ssum * __expf(m - mn) + s2 * __expf(m2 - mn);
    if (m2 > m || (m2 == m && a2 < am)) am = a2;
    m = mn;
  }
  if (lane == 0) { s_m[w] = m; s_s[w] = ssum; s_i[w] = am; s_x[w] = xsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = s_m[0], S = s_s[0], X = s_x[0];
    int A = s_i[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      float m2 = s_m[i], s2 = s_s[i];
      int a2 = s_i[i];
      X += s_x[i];
      float mn = fmaxf(M, m2);
      S = S * __expf(M - mn) + s2 * __expf(m2 - mn);
      if (m2 > M || (m2 == M && a2 < A)) A = a2;
      M = mn;
    }
    float lse = M + logf(S);
    int g = gold[row];
    float loss = 0.f;
    if (g != pad_id) {
      float lp_g = x[g] - lse;
      if (smoothing > 0.f) {
        // -(sum_v t_v logp_v), t = onehot*(1-eps) + (1-onehot)*eps/V   (metrics.py:117-121)
        float sum_lp = X - (float)V * lse;
        loss = -((1.f - smoothing) * lp_g + (smoothing / (float)V) * (sum_lp - lp_g));
      } else {
        loss = -lp_g;
      }
    }
    row_lse[row] = lse;
    row_loss[row] = loss;
    hyp[row] = A;
  }
}

__global__ void __launch_bounds__(256) ce_finalize_kernel(const float* __restrict__ row_loss,
                                                          const int* __restrict__ gold, const int* __restrict__ hyp,
                                                          int M, int pad_id, CeOut* out) {
  __shared__ float red[32];
  float l = 0.f, nv = 0.f, nc = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    int g = gold[i];
    if (g != pad_id) { l += row_loss[i]; nv += 1.f; nc += (hyp[i] == g) ? 1.f : 0.f; }
  }
  l = block_sum(l, red);
  nv = block_sum(nv, red);
  nc = block_sum(nc, red);
  if (threadIdx.x == 0) {
    out->loss = l / nv;          // nv == 0 -> NaN, as F.cross_entropy(mean) over an all-ignored batch
    out->n_valid = nv;
    out->n_correct = nc;
  }
}

int k_ce_fwd(const float* logits, int ld, const int* gold, int M, int V, float smoothing, int pad_id,
             float* row_lse, float* row_loss, int* hyp, CeOut* out, cudaStream_t s) {
  MTL_REQUIRE(M > 0 && V > 0, "empty logits");
  ce_row_kernel<<<M, 256, 0, s>>>(logits, ld, gold, V, smoothing, pad_id, row_lse, row_loss, hyp);
  MTL_CHECK_LAUNCH();
  ce_finalize_kernel<<<1, 256, 0, s>>>(row_loss, gold, hyp, M, pad_id, out);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, int ld,
                                                     const int* __restrict__ gold,
                                                     const float* __restrict__ row_lse, const CeOut* out,
                                                     float scale, float smoothing, int pad_id,
                                                     float* __restrict__ dlogits, int V) {
  const int row = blockIdx.x;
  const int g = gold[row];
  const float* x = logits + (size_t)row * ld;
  float* dx = dlogits + (size_t)row * ld;
  if (g == pad_id) {
    for (int v = threadIdx.x; v < ld; v += blockDim.x) dx[v] = 0.f;
    return;
  }
  const float lse = row_lse[row];
  const float c = scale / out->n_valid;
  const float t_off = smoothing > 0.f ? smoothing / (float)V : 0.f;
  const float t_on = smoothing > 0.f ? 1.f - smoothing : 1.f;
  // sum_v t_v = t_on + (V-1)*t_off  (== 1 without smoothing)
  const float tsum = t_on + (float)(V - 1) * t_off;
  for (int v = threadIdx.x; v < ld; v += blockDim.x) {
    float r = 0.f;
    if (v < V) {
      float p = __expf(x[v] - lse);
      float t = (v == g) ? t_on : t_off;
      r = c * (tsum * p - t);
    }
    dx[v] = r;
  }
}
int k_ce_bwd(const float* logits, int ld, const int* gold, const float* row_lse, const CeOut* out, float scale,
             float smoothing, int pad_id, float* dlogits, int M, int V, cudaStream_t s) {
  MTL_REQUIRE(M > 0 && V > 0, "empty logits");
  ce_bwd_kernel<<<M, 256, 0, s>>>(logits, ld, gold, row_lse, out, scale, smoothing, pad_id, dlogits, V);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
