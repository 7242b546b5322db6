// Fused masked scaled-dot-product attention, forward and backward (flash-style: scores are never
// written to HBM; the backward recomputes probabilities from the saved log-sum-exp).
// Masks are predicates computed in-kernel from a (B,Tk) key-pad byte map + a causal flag instead of
// the reference's materialised (B,Tq,Tk) tensors repeated H times.
// Reference: modules/common_layers.py:291-301 (head split / merge), :317-331 (QK^T/sqrt(dk),
// masked_fill(-inf), softmax(dim=2), dropout, .V), modules/decoder.py:86-94 (masks).
// Layout: q/k/v/o stay in the (B*T, H*dk) row-major layout the projections produce; head h is the
// column block [h*dk, (h+1)*dk) -- the reference's permute+contiguous copies disappear.
#include "kernels.h"
#include <math.h>

#define ATT_TILE 32      // keys (fwd/dQ) or queries (dKV) per smem tile == warp width
#define ATT_ROWS 16      // rows per CTA (4 warps x 4 rows)
#define ATT_RPW 4

template <int DK>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnArgs a) {
  constexpr int NV = DK / 32;
  __shared__ float Ks[ATT_TILE][DK + 1];
  __shared__ float Vs[ATT_TILE][DK + 1];
  __shared__ float Qs[ATT_ROWS][DK];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_ROWS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ATT_ROWS * DK; i += 128) {
    int r = i / DK, d = i % DK, qi = q0 + r;
    Qs[r][d] = qi < a.Tq ? a.q[(size_t)(b * a.Tq + qi) * a.ldq + h * DK + d] : 0.f;
  }
  float m[ATT_RPW], l[ATT_RPW], acc[ATT_RPW][NV];
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[r][v] = 0.f;
  }
  // causal: keys beyond the last query row of this CTA never contribute
  const int k_end = a.causal ? min(a.Tk, q0 + ATT_ROWS) : a.Tk;
  for (int k0 = 0; k0 < k_end; k0 += ATT_TILE) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_TILE * DK; i += 128) {
      int j = i / DK, d = i % DK, kj = k0 + j;
      bool ok = kj < a.Tk;
      Ks[j][d] = ok ? a.k[(size_t)(b * a.Tk + kj) * a.ldk + h * DK + d] : 0.f;
      Vs[j][d] = ok ? a.v[(size_t)(b * a.Tk + kj) * a.ldv + h * DK + d] : 0.f;
    }
    __syncthreads();
    const int kj = k0 + lane;
    const bool key_ok = kj < a.Tk && !(a.keypad && a.keypad[(size_t)b * a.Tk + kj]);
#pragma unroll
    for (int r = 0; r < ATT_RPW; ++r) {
      const int rr = w * ATT_RPW + r, qi = q0 + rr;
      if (qi >= a.Tq) continue;                       // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DK; ++d) s += Qs[rr][d] * Ks[lane][d];
      s *= a.inv_temp;
      const bool ok = key_ok && !(a.causal && kj > qi);
      s = ok ? s : -INFINITY;
      const float mt = warp_max(s);
      const float mn = fmaxf(m[r], mt);
      float p = 0.f, corr = 1.f;
      if (mn != -INFINITY) {
        p = ok ? __expf(s - mn) : 0.f;
        corr = (m[r] == -INFINITY) ? 0.f : __expf(m[r] - mn);
      }
      l[r] = l[r] * corr + warp_sum(p);
      m[r] = mn;
      if (a.drop.p > 0.f && ok) {
        unsigned long long idx = (((unsigned long long)(b * a.H + h) * a.Tq + qi) * a.Tk + kj);
        p *= dropout_scale(mtl_eff_seed(a.drop), a.drop.site, idx, a.drop.p, a.drop.inv_keep);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[r][v] *= corr;
#pragma unroll
      for (int j = 0; j < ATT_TILE; ++j) {
        float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[r][v] += pj * Vs[j][lane + 32 * v];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    const int qi = q0 + w * ATT_RPW + r;
    if (qi >= a.Tq) continue;
    const float inv = 1.f / l[r];                      // fully masked row -> NaN, like softmax(-inf row)
#pragma unroll
    for (int v = 0; v < NV; ++v)
      a.o[(size_t)(b * a.Tq + qi) * a.ldo + h * DK + lane + 32 * v] = acc[r][v] * inv;
    if (lane == 0) a.lse[((size_t)b * a.H + h) * a.Tq + qi] = m[r] + logf(l[r]);
  }
}

int k_attn_fwd(const AttnArgs& a, cudaStream_t s) {
  MTL_REQUIRE(a.dk == 32 || a.dk == 64, "attention head dim must be 32 or 64");
  if (a.B * a.H * a.Tq == 0) return MTL_OK;
  dim3 grid(mtl_cdiv(a.Tq, ATT_ROWS), a.H, a.B);
  if (a.dk == 64) attn_fwd_kernel<64><<<grid, 128, 0, s>>>(a);
  else attn_fwd_kernel<32><<<grid, 128, 0, s>>>(a);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

// delta[b,h,q] = dO . O   (== sum_j P_ij dP_ij, also with dropout)
template <int DK>
__global__ void __launch_bounds__(128) attn_delta_kernel(AttnBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int total = a.f.B * a.f.Tq * a.f.H;
  if (gw >= total) return;
  const int h = gw % a.f.H, row = gw / a.f.H;          // row = b*Tq + qi
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < DK / 32; ++v) {
    size_t off = (size_t)row * a.f.ldo + h * DK + lane + 32 * v;
    s += a.d_o[off] * a.f.o[off];
  }
  s = warp_sum(s);
  if (lane == 0) {
    int b = row / a.f.Tq, qi = row % a.f.Tq;
    a.delta[((size_t)b * a.f.H + h) * a.f.Tq + qi] = s;
  }
}

// dQ: CTA = 16 query rows of one (b,h); loops over key tiles.
template <int DK>
__global__ void __launch_bounds__(128) attn_dq_kernel(AttnBwdArgs a) {
  constexpr int NV = DK / 32;
  const AttnArgs& f = a.f;
  __shared__ float Ks[ATT_TILE][DK + 1];
  __shared__ float Vs[ATT_TILE][DK + 1];
  __shared__ float Qs[ATT_ROWS][DK];
  __shared__ float Gs[ATT_ROWS][DK];      // dO rows
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_ROWS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ATT_ROWS * DK; i += 128) {
    int r = i / DK, d = i % DK, qi = q0 + r;
    bool ok = qi < f.Tq;
    Qs[r][d] = ok ? f.q[(size_t)(b * f.Tq + qi) * f.ldq + h * DK + d] : 0.f;
    Gs[r][d] = ok ? a.d_o[(size_t)(b * f.Tq + qi) * f.ldo + h * DK + d] : 0.f;
  }
  float lse[ATT_RPW], dl[ATT_RPW], acc[ATT_RPW][NV];
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    const int qi = q0 + w * ATT_RPW + r;
    size_t si = ((size_t)b * f.H + h) * f.Tq + min(qi, f.Tq - 1);
    lse[r] = f.lse[si]; dl[r] = a.delta[si];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[r][v] = 0.f;
  }
  const int k_end = f.causal ? min(f.Tk, q0 + ATT_ROWS) : f.Tk;
  for (int k0 = 0; k0 < k_end; k0 += ATT_TILE) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_TILE * DK; i += 128) {
      int j = i / DK, d = i % DK, kj = k0 + j;
      bool ok = kj < f.Tk;
      Ks[j][d] = ok ? f.k[(size_t)(b * f.Tk + kj) * f.ldk + h * DK + d] : 0.f;
      Vs[j][d] = ok ? f.v[(size_t)(b * f.Tk + kj) * f.ldv + h * DK + d] : 0.f;
    }
    __syncthreads();
    const int kj = k0 + lane;
    const bool key_ok = kj < f.Tk && !(f.keypad && f.keypad[(size_t)b * f.Tk + kj]);
#pragma unroll
    for (int r = 0; r < ATT_RPW; ++r) {
      const int rr = w * ATT_RPW + r, qi = q0 + rr;
      if (qi >= f.Tq) continue;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DK; ++d) { s += Qs[rr][d] * Ks[lane][d]; dp += Gs[rr][d] * Vs[lane][d]; }
      const bool ok = key_ok && !(f.causal && kj > qi);
      float p = ok ? __expf(s * f.inv_temp - lse[r]) : 0.f;
      if (f.drop.p > 0.f && ok) {
        unsigned long long idx = (((unsigned long long)(b * f.H + h) * f.Tq + qi) * f.Tk + kj);
        dp *= dropout_scale(mtl_eff_seed(f.drop), f.drop.site, idx, f.drop.p, f.drop.inv_keep);
      }
      const float ds = p * (dp - dl[r]);
#pragma unroll
      for (int j = 0; j < ATT_TILE; ++j) {
        float dj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[r][v] += dj * Ks[j][lane + 32 * v];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    const int qi = q0 + w * ATT_RPW + r;
    if (qi >= f.Tq) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      a.dq[(size_t)(b * f.Tq + qi) * f.ldq + h * DK + lane + 32 * v] = acc[r][v] * f.inv_temp;
  }
}

// dK,dV: CTA = 16 key rows of one (b,h); loops over query tiles.
template <int DK>
__global__ void __launch_bounds__(128) attn_dkv_kernel(AttnBwdArgs a) {
  constexpr int NV = DK / 32;
  const AttnArgs& f = a.f;
  __shared__ float Qt[ATT_TILE][DK + 1];
  __shared__ float Gt[ATT_TILE][DK + 1];
  __shared__ float Ks[ATT_ROWS][DK];
  __shared__ float Vs[ATT_ROWS][DK];
  __shared__ float Ls[ATT_TILE], Ds[ATT_TILE];
  const int b = blockIdx.z, h = blockIdx.y, kb0 = blockIdx.x * ATT_ROWS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ATT_ROWS * DK; i += 128) {
    int r = i / DK, d = i % DK, kj = kb0 + r;
    bool ok = kj < f.Tk;
    Ks[r][d] = ok ? f.k[(size_t)(b * f.Tk + kj) * f.ldk + h * DK + d] : 0.f;
    Vs[r][d] = ok ? f.v[(size_t)(b * f.Tk + kj) * f.ldv + h * DK + d] : 0.f;
  }
  float accK[ATT_RPW][NV], accV[ATT_RPW][NV];
  bool kvalid[ATT_RPW];
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    const int kj = kb0 + w * ATT_RPW + r;
    kvalid[r] = kj < f.Tk && !(f.keypad && f.keypad[(size_t)b * f.Tk + kj]);
#pragma unroll
    for (int v = 0; v < NV; ++v) { accK[r][v] = 0.f; accV[r][v] = 0.f; }
  }
  // causal: queries before the first key row of this CTA never see these keys
  const int q_begin = f.causal ? (kb0 / ATT_TILE) * ATT_TILE : 0;
  for (int q0 = q_begin; q0 < f.Tq; q0 += ATT_TILE) {
    __syncthreads();
    for (int i = threadIdx.x; i < ATT_TILE * DK; i += 128) {
      int j = i / DK, d = i % DK, qi = q0 + j;
      bool ok = qi < f.Tq;
      Qt[j][d] = ok ? f.q[(size_t)(b * f.Tq + qi) * f.ldq + h * DK + d] : 0.f;
      Gt[j][d] = ok ? a.d_o[(size_t)(b * f.Tq + qi) * f.ldo + h * DK + d] : 0.f;
    }
    if (threadIdx.x < ATT_TILE) {
      int qi = q0 + threadIdx.x;
      size_t si = ((size_t)b * f.H + h) * f.Tq + min(qi, f.Tq - 1);
      Ls[threadIdx.x] = f.lse[si];
      Ds[threadIdx.x] = a.delta[si];
    }
    __syncthreads();
    const int qi = q0 + lane;
#pragma unroll
    for (int r = 0; r < ATT_RPW; ++r) {
      const int rr = w * ATT_RPW + r, kj = kb0 + rr;
      if (kj >= f.Tk) continue;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DK; ++d) { s += Qt[lane][d] * Ks[rr][d]; dp += Gt[lane][d] * Vs[rr][d]; }
      const bool ok = kvalid[r] && qi < f.Tq && !(f.causal && kj > qi);
      float p = ok ? __expf(s * f.inv_temp - Ls[lane]) : 0.f;
      float pd = p;
      if (f.drop.p > 0.f && ok) {
        unsigned long long idx = (((unsigned long long)(b * f.H + h) * f.Tq + qi) * f.Tk + kj);
        float sc = dropout_scale(mtl_eff_seed(f.drop), f.drop.site, idx, f.drop.p, f.drop.inv_keep);
        pd = p * sc; dp *= sc;
      }
      const float ds = p * (dp - Ds[lane]);
#pragma unroll
      for (int j = 0; j < ATT_TILE; ++j) {
        float pj = __shfl_sync(0xffffffffu, pd, j);
        float dj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          accV[r][v] += pj * Gt[j][lane + 32 * v];
          accK[r][v] += dj * Qt[j][lane + 32 * v];
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATT_RPW; ++r) {
    const int kj = kb0 + w * ATT_RPW + r;
    if (kj >= f.Tk) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      a.dk[(size_t)(b * f.Tk + kj) * f.ldk + h * DK + lane + 32 * v] = accK[r][v] * f.inv_temp;
      a.dv[(size_t)(b * f.Tk + kj) * f.ldv + h * DK + lane + 32 * v] = accV[r][v];
    }
  }
}

template <int DK>
static int attn_bwd_launch(const AttnBwdArgs& a, cudaStream_t s) {
  const AttnArgs& f = a.f;
  int nw = f.B * f.Tq * f.H;
  attn_delta_kernel<DK><<<mtl_cdiv(nw, 4), 128, 0, s>>>(a);
  MTL_CHECK_LAUNCH();
  attn_dq_kernel<DK><<<dim3(mtl_cdiv(f.Tq, ATT_ROWS), f.H, f.B), 128, 0, s>>>(a);
  MTL_CHECK_LAUNCH();
  attn_dkv_kernel<DK><<<dim3(mtl_cdiv(f.Tk, ATT_ROWS), f.H, f.B), 128, 0, s>>>(a);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
int k_attn_bwd(const AttnBwdArgs& a, cudaStream_t s) {
  MTL_REQUIRE(a.f.dk == 32 || a.f.dk == 64, "attention head dim must be 32 or 64");
  if (a.f.B * a.f.H * a.f.Tq == 0) return MTL_OK;
  return a.f.dk == 64 ? attn_bwd_launch<64>(a, s) : attn_bwd_launch<32>(a, s);
}
