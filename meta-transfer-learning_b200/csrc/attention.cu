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

int k_attn_small_fwd(const AttnArgs& a, cudaStream_t s);
int k_attn_small_bwd(const AttnBwdArgs& a, cudaStream_t s);
bool k_attn_small_eligible(const AttnArgs& a);
int k_attn_flash_fwd(const AttnArgs& a, cudaStream_t s);
int k_attn_flash_bwd(const AttnBwdArgs& a, cudaStream_t s);
bool k_attn_flash_eligible(const AttnArgs& a);

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
  if (k_attn_small_eligible(a)) return k_attn_small_fwd(a, s);
  if (k_attn_flash_eligible(a)) return k_attn_flash_fwd(a, s);
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
  if (k_attn_small_eligible(a.f) && (((uintptr_t)a.d_o | (uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) & 15u) == 0)
    return k_attn_small_bwd(a, s);
  if (k_attn_flash_eligible(a.f) && (((uintptr_t)a.d_o | (uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) & 15u) == 0)
    return k_attn_flash_bwd(a, s);
  return a.f.dk == 64 ? attn_bwd_launch<64>(a, s) : attn_bwd_launch<32>(a, s);
}

// ----------------------------------------------------------------------------- short sequences (Tq, Tk <= 64)
// BASELINE cfg 2 runs attention over T' = 25 encoder frames and n = 33 decoder tokens: one (b, h) pair is a 33 x 33 x 64
// problem.  The tiled kernels above pay a whole 32-wide tile for the 33rd key / query and split the backward over three
// launches on the critical path; here ONE CTA owns one (b, h): Q, K, V (and dO) sit in shared memory once, the scores
// never leave it, and the backward (delta, dQ, dK, dV) is a single kernel.  Every product runs on a 16 x 16 thread grid
// with 4 x 4 register blocks and float4 shared-memory operands.  Same masks, same Philox dropout indices and the same
// fully-masked-row semantics (NaN, like softmax over a row of -inf) as the tiled kernels.
__device__ long long g_attn_dbg[32];   // cycle stamps of CTA (0,0): [0,8) forward phases, [16,24) backward phases
#define SA_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_attn_dbg[i] = clock64(); } while (0)
int k_attn_debug_stamps(long long* host32) {
  MTL_CHECK_CUDA(cudaMemcpyFromSymbol(host32, g_attn_dbg, sizeof(long long) * 32));
  return MTL_OK;
}
constexpr int SA_T = 64;           // rows (queries / keys) a CTA can hold
constexpr int SA_LD = 68;          // row stride in floats: 16 B aligned rows, 16 consecutive rows hit distinct bank quads

struct SaSmem {
  float q[SA_T][SA_LD], k[SA_T][SA_LD], v[SA_T][SA_LD], g[SA_T][SA_LD];   // g = dO (backward only)
  float p[SA_T][SA_LD], ds[SA_T][SA_LD];                                   // probabilities (after dropout) / dS
  float lse[SA_T], delta[SA_T];
  int kmask[SA_T];                                                         // 1 = key masked (padding) or beyond Tk
};
// key-pad flags of batch row b -> shared memory (one global round trip for the whole CTA, not one per score)
__device__ __forceinline__ void sa_load_kmask(int* dst, const unsigned char* keypad, int b, int Tk) {
  if (threadIdx.x < SA_T) {
    const int j = threadIdx.x;
    dst[j] = (j >= Tk || (keypad && keypad[(size_t)b * Tk + j])) ? 1 : 0;
  }
}
static_assert(sizeof(SaSmem) <= 227 * 1024, "shared memory budget");

// rows [0, T) of a (B*T, ld) matrix's head-h column block -> shared memory, rows [T, SA_T) zero.  Two phases (every
// global load of the CTA is in flight before the first shared-memory store) so that a kernel pays ONE global round
// trip for all of its operands: sa_fetch into registers for each matrix, then sa_store for each.
template <int DK, int NT>
struct SaTile { float4 v[SA_T * (DK / 4) / NT]; };
template <int DK, int NT>
__device__ __forceinline__ void sa_fetch(SaTile<DK, NT>& t, const float* src, int b, int T, int ld, int h) {
  constexpr int C4 = DK / 4, N = SA_T * C4 / NT;
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const int i = threadIdx.x + u * NT, r = i / C4, c4 = i % C4;
    t.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < T) t.v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(b * T + r) * ld + h * DK + c4 * 4));
  }
}
template <int DK, int NT>
__device__ __forceinline__ void sa_store(float (*dst)[SA_LD], const SaTile<DK, NT>& t) {
  constexpr int C4 = DK / 4, N = SA_T * C4 / NT;
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const int i = threadIdx.x + u * NT, r = i / C4, c4 = i % C4;
    *reinterpret_cast<float4*>(&dst[r][c4 * 4]) = t.v[u];
  }
}
__device__ __forceinline__ void fma4(float& acc, const float4& a, const float4& b) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
}
// Thread grid of the products: tx = thread % 16 walks columns, ty = thread / 16 walks rows with stride RS = NT / 16;
// a thread owns R = SA_T / RS register rows.
// acc[r][c] = sum_d X[ty + RS r][d] * Y[tx + 16 c][d]   (rows ty + RS r < rows, c < nc).  The row test is per thread: with
// 16 warps and T = 33 only the warp that owns row 32 runs the second register row at all.
template <int DK, int RS>
__device__ __forceinline__ void sa_nt(float (&acc)[SA_T / RS][4], const float (*X)[SA_LD], const float (*Y)[SA_LD], int ty,
                                      int tx, int rows, int nc) {
  constexpr int R = SA_T / RS;
  bool rv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rv[r] = ty + RS * r < rows;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  }
#pragma unroll 4
  for (int d = 0; d < DK; d += 4) {
    float4 x[R], y[4];
#pragma unroll
    for (int r = 0; r < R; ++r) if (rv[r]) x[r] = *reinterpret_cast<const float4*>(&X[ty + RS * r][d]);
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < nc) y[c] = *reinterpret_cast<const float4*>(&Y[tx + 16 * c][d]);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) if (rv[r] && c < nc) fma4(acc[r][c], x[r], y[c]);
  }
}
// out[ty + RS r][4 tx .. 4 tx + 3] = scale * sum_{j < J} W[ty + RS r][j] * Y[j][4 tx ..]   (r < nr; J padded to 4 with zeros)
template <int DK, int RS>
__device__ __forceinline__ void sa_nn(float* out, int ld, const float (*W)[SA_LD], const float (*Y)[SA_LD], int ty, int tx,
                                      int nr, int rows, int J, float scale) {
  if (tx * 4 >= DK) return;
  constexpr int R = SA_T / RS;
  float4 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < J; j += 4) {
    float4 y[4], w[R];
#pragma unroll
    for (int u = 0; u < 4; ++u) y[u] = *reinterpret_cast<const float4*>(&Y[j + u][tx * 4]);
#pragma unroll
    for (int r = 0; r < R; ++r) if (ty + RS * r < rows) w[r] = *reinterpret_cast<const float4*>(&W[ty + RS * r][j]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (ty + RS * r >= rows) continue;
      acc[r].x += w[r].x * y[0].x + w[r].y * y[1].x + w[r].z * y[2].x + w[r].w * y[3].x;
      acc[r].y += w[r].x * y[0].y + w[r].y * y[1].y + w[r].z * y[2].y + w[r].w * y[3].y;
      acc[r].z += w[r].x * y[0].z + w[r].y * y[1].z + w[r].z * y[2].z + w[r].w * y[3].z;
      acc[r].w += w[r].x * y[0].w + w[r].y * y[1].w + w[r].z * y[2].w + w[r].w * y[3].w;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = ty + RS * r;
    if (i < rows)
      *reinterpret_cast<float4*>(out + (size_t)i * ld + tx * 4) =
          make_float4(acc[r].x * scale, acc[r].y * scale, acc[r].z * scale, acc[r].w * scale);
  }
}
// out[ty + RS r][4 tx ..] = scale * sum_{i < I} W[i][ty + RS r] * Y[i][4 tx ..]   (contraction over ROWS of W)
template <int DK, int RS>
__device__ __forceinline__ void sa_tn(float* out, int ld, const float (*W)[SA_LD], const float (*Y)[SA_LD], int ty, int tx,
                                      int nr, int rows, int I, float scale) {
  if (tx * 4 >= DK) return;
  constexpr int R = SA_T / RS;
  float4 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    const float4 y = *reinterpret_cast<const float4*>(&Y[i][tx * 4]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (ty + RS * r >= rows) continue;
      const float w = W[i][ty + RS * r];
      acc[r].x += w * y.x; acc[r].y += w * y.y; acc[r].z += w * y.z; acc[r].w += w * y.w;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = ty + RS * r;
    if (j < rows)
      *reinterpret_cast<float4*>(out + (size_t)j * ld + tx * 4) =
          make_float4(acc[r].x * scale, acc[r].y * scale, acc[r].z * scale, acc[r].w * scale);
  }
}

template <int DK, int NT>
__global__ void __launch_bounds__(NT) attn_small_fwd_kernel(AttnArgs a) {
  constexpr int RS = NT / 16, R = SA_T / RS, NW = NT / 32;
  extern __shared__ __align__(16) unsigned char sa_raw[];
  SaSmem& S = *reinterpret_cast<SaSmem*>(sa_raw);
  const int h = blockIdx.x, b = blockIdx.y;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nrq = (a.Tq + RS - 1) / RS, nck = (a.Tk + 15) >> 4;
  pdl_wait();
  pdl_trigger();
  SA_STAMP(0);
  {
    SaTile<DK, NT> tq, tk, tv;
    sa_fetch<DK, NT>(tq, a.q, b, a.Tq, a.ldq, h);
    sa_fetch<DK, NT>(tk, a.k, b, a.Tk, a.ldk, h);
    sa_fetch<DK, NT>(tv, a.v, b, a.Tk, a.ldv, h);
    sa_load_kmask(S.kmask, a.keypad, b, a.Tk);
    sa_store<DK, NT>(S.q, tq); sa_store<DK, NT>(S.k, tk); sa_store<DK, NT>(S.v, tv);
  }
  __syncthreads();
  SA_STAMP(1);
  {  // masked, scaled scores -> S.p
    float acc[R][4];
    sa_nt<DK, RS>(acc, S.q, S.k, ty, tx, a.Tq, nck);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (r >= nrq || c >= nck) continue;
        const int i = ty + RS * r, j = tx + 16 * c;
        const bool ok = i < a.Tq && !S.kmask[j] && !(a.causal && j > i);
        S.p[i][j] = ok ? acc[r][c] * a.inv_temp : -INFINITY;
      }
  }
  __syncthreads();
  SA_STAMP(2);
  const unsigned long long seed = a.drop.p > 0.f ? mtl_eff_seed(a.drop) : 0ull;
  // softmax over keys, one warp per query row; the rows of a warp are independent dependency chains (shuffles, exp,
  // Philox): unrolled so that they overlap instead of running back to back
  const int Tk = a.Tk, Tq = a.Tq;
  const float drop_p = a.drop.p, inv_keep = a.drop.inv_keep;
  const uint32_t site = a.drop.site;
#pragma unroll
  for (int u = 0; u < SA_T / NW; ++u) {
    const int i = w + NW * u;
    if (i >= Tq) continue;                                              // warp-uniform
    const float s0 = lane < Tk ? S.p[i][lane] : -INFINITY, s1 = lane + 32 < Tk ? S.p[i][lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(s0, s1));
    float p0 = 0.f, p1 = 0.f;
    if (m != -INFINITY) { p0 = __expf(s0 - m); p1 = __expf(s1 - m); }   // exp(-inf) = 0 for masked keys
    const float l = warp_sum(p0 + p1);
    const float inv = 1.f / l;                                          // fully masked row -> 0 * inf = NaN, like the reference
    p0 *= inv; p1 *= inv;
    if (drop_p > 0.f) {
      const unsigned long long base = ((unsigned long long)(b * a.H + h) * Tq + i) * Tk;
      if (s0 != -INFINITY) p0 *= dropout_scale(seed, site, base + lane, drop_p, inv_keep);
      if (s1 != -INFINITY) p1 *= dropout_scale(seed, site, base + lane + 32, drop_p, inv_keep);
    }
    S.p[i][lane] = p0; S.p[i][lane + 32] = p1;
    if (lane == 0) a.lse[((size_t)b * a.H + h) * Tq + i] = m + logf(l);
  }
  __syncthreads();
  SA_STAMP(3);
  sa_nn<DK, RS>(a.o + (size_t)b * a.Tq * a.ldo + h * DK, a.ldo, S.p, S.v, ty, tx, nrq, a.Tq, (a.Tk + 3) & ~3, 1.f);
  SA_STAMP(4);
}

template <int DK, int NT>
__global__ void __launch_bounds__(NT) attn_small_bwd_kernel(AttnBwdArgs a) {
  constexpr int RS = NT / 16, R = SA_T / RS, NW = NT / 32;
  extern __shared__ __align__(16) unsigned char sa_raw[];
  SaSmem& S = *reinterpret_cast<SaSmem*>(sa_raw);
  const AttnArgs& f = a.f;
  const int h = blockIdx.x, b = blockIdx.y;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nrq = (f.Tq + RS - 1) / RS, nck = (f.Tk + 15) >> 4, nrk = (f.Tk + RS - 1) / RS;
  pdl_wait();
  pdl_trigger();
  SA_STAMP(16);
  {
    SaTile<DK, NT> tq, tk, tv, tg, to;
    sa_fetch<DK, NT>(tq, f.q, b, f.Tq, f.ldq, h);
    sa_fetch<DK, NT>(tk, f.k, b, f.Tk, f.ldk, h);
    sa_fetch<DK, NT>(tv, f.v, b, f.Tk, f.ldv, h);
    sa_fetch<DK, NT>(tg, a.d_o, b, f.Tq, f.ldo, h);
    sa_fetch<DK, NT>(to, f.o, b, f.Tq, f.ldo, h);      // O rows, only for delta (S.p is rewritten below)
    sa_load_kmask(S.kmask, f.keypad, b, f.Tk);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + SA_T) {
      const int i = threadIdx.x - 64;
      S.lse[i] = i < f.Tq ? f.lse[((size_t)b * f.H + h) * f.Tq + i] : 0.f;
    }
    sa_store<DK, NT>(S.q, tq); sa_store<DK, NT>(S.k, tk); sa_store<DK, NT>(S.v, tv); sa_store<DK, NT>(S.g, tg); sa_store<DK, NT>(S.p, to);
  }
  __syncthreads();
  SA_STAMP(17);
  for (int i = w; i < SA_T; i += NW) {         // delta_i = dO_i . O_i (== sum_j P_ij dP_ij, also with dropout)
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < DK / 32; ++v) s += S.g[i][lane + 32 * v] * S.p[i][lane + 32 * v];
    s = warp_sum(s);
    if (lane == 0) S.delta[i] = s;
  }
  __syncthreads();
  SA_STAMP(18);
  {
    float sc[R][4], dp[R][4];
    sa_nt<DK, RS>(sc, S.q, S.k, ty, tx, f.Tq, nck);
    sa_nt<DK, RS>(dp, S.g, S.v, ty, tx, f.Tq, nck);
    const unsigned long long seed = f.drop.p > 0.f ? mtl_eff_seed(f.drop) : 0ull;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (r >= nrq || c >= nck) continue;
        const int i = ty + RS * r, j = tx + 16 * c;
        const bool ok = i < f.Tq && !S.kmask[j] && !(f.causal && j > i);
        const float p = ok ? __expf(sc[r][c] * f.inv_temp - S.lse[i]) : 0.f;
        float pd = p, d = dp[r][c];
        if (f.drop.p > 0.f && ok) {
          const unsigned long long idx = (((unsigned long long)(b * f.H + h) * f.Tq + i) * f.Tk + j);
          const float s = dropout_scale(seed, f.drop.site, idx, f.drop.p, f.drop.inv_keep);
          pd = p * s; d *= s;
        }
        S.p[i][j] = pd;
        S.ds[i][j] = p * (d - S.delta[i]);
      }
  }
  __syncthreads();
  SA_STAMP(19);
  const int j4 = (f.Tk + 3) & ~3;
  sa_nn<DK, RS>(a.dq + (size_t)b * f.Tq * f.ldq + h * DK, f.ldq, S.ds, S.k, ty, tx, nrq, f.Tq, j4, f.inv_temp);
  sa_tn<DK, RS>(a.dk + (size_t)b * f.Tk * f.ldk + h * DK, f.ldk, S.ds, S.q, ty, tx, nrk, f.Tk, f.Tq, f.inv_temp);
  sa_tn<DK, RS>(a.dv + (size_t)b * f.Tk * f.ldv + h * DK, f.ldv, S.p, S.g, ty, tx, nrk, f.Tk, f.Tq, 1.f);
  SA_STAMP(20);
}

// ----------------------------------------------------------------------------- short sequences on the tensor cores
// The same two kernels with every product on mma.sync.m16n8k8 TF32 fragments, 3xTF32 in registers (hi = rna_tf32(a),
// lo = rna_tf32(a - hi); lo*hi + hi*lo + hi*hi accumulate in fp32): fp32-grade scores and gradients.  A (b, h) problem is
// 33 x 33 x 64 -- far below one tcgen05 tile (M = 128) and not worth a TMEM allocation and a TMA descriptor fetch on a
// latency-bound chain kernel -- so the warp-level MMA is the tensor-core path that fits: a warp owns 16 x 8 output tiles
// and reads its fragments straight from the row-major shared-memory operands (row stride 68 floats: the 8 rows x 4
// columns of an A / NT-B fragment fall into 32 distinct banks).  The CUDA-core version spent 6.8 k of its 17 k cycles on
// the scores alone (shared-memory wavefronts of a 4 x 4 register-blocked FFMA loop).  MTL_ATTN_MMA=0 selects it (A/B).
// round-to-nearest (ties away) fp32 -> tf32 with two full-rate integer ops (cvt.rna.tf32.f32 runs on the quarter-rate
// conversion pipe; same result for finite inputs -- see tf32_rna in gemm_tc.cu)
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3x(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}
// Fragment coordinates of lane = 4 g + t:  A (16 x 8): (g, t) (g + 8, t) (g, t + 4) (g + 8, t + 4);  B (8 x 8): (k = t, n = g)
// (k = t + 4, n = g);  C (16 x 8): (g, 2t) (g, 2t + 1) (g + 8, 2t) (g + 8, 2t + 1).
// c += X[m0.., k] * Y[n0.., k]^T over k < K8 (multiple of 8)           (both operands row-major, contraction along columns)
__device__ __forceinline__ void sa_mma_nt(float (&c)[4], const float (*X)[SA_LD], const float (*Y)[SA_LD], int m0, int n0,
                                          int K8, int g, int t) {
  for (int k0 = 0; k0 < K8; k0 += 8) {
    const float a[4] = {X[m0 + g][k0 + t], X[m0 + g + 8][k0 + t], X[m0 + g][k0 + t + 4], X[m0 + g + 8][k0 + t + 4]};
    const float b[2] = {Y[n0 + g][k0 + t], Y[n0 + g][k0 + t + 4]};
    mma_3x(c, a, b);
  }
}
// c += W[m0.., k] * Y[k, n0..] over k < K8                                (A row-major, B row-major [k][n])
__device__ __forceinline__ void sa_mma_nn(float (&c)[4], const float (*W)[SA_LD], const float (*Y)[SA_LD], int m0, int n0,
                                          int K8, int g, int t) {
  for (int k0 = 0; k0 < K8; k0 += 8) {
    const float a[4] = {W[m0 + g][k0 + t], W[m0 + g + 8][k0 + t], W[m0 + g][k0 + t + 4], W[m0 + g + 8][k0 + t + 4]};
    const float b[2] = {Y[k0 + t][n0 + g], Y[k0 + t + 4][n0 + g]};
    mma_3x(c, a, b);
  }
}
// c += W[k, m0..]^T * Y[k, n0..] over k < K8                              (contraction over the ROWS of both operands)
__device__ __forceinline__ void sa_mma_tn(float (&c)[4], const float (*W)[SA_LD], const float (*Y)[SA_LD], int m0, int n0,
                                          int K8, int g, int t) {
  for (int k0 = 0; k0 < K8; k0 += 8) {
    const float a[4] = {W[k0 + t][m0 + g], W[k0 + t][m0 + g + 8], W[k0 + t + 4][m0 + g], W[k0 + t + 4][m0 + g + 8]};
    const float b[2] = {Y[k0 + t][n0 + g], Y[k0 + t + 4][n0 + g]};
    mma_3x(c, a, b);
  }
}
// fragment -> global rows [0, rows) of a (.., ld) matrix, columns n0 + 2t, n0 + 2t + 1
__device__ __forceinline__ void sa_store_frag(float* out, int ld, const float (&c)[4], int m0, int n0, int rows, int g, int t,
                                              float scale) {
  if (m0 + g < rows) *reinterpret_cast<float2*>(out + (size_t)(m0 + g) * ld + n0 + 2 * t) = make_float2(c[0] * scale, c[1] * scale);
  if (m0 + g + 8 < rows) *reinterpret_cast<float2*>(out + (size_t)(m0 + g + 8) * ld + n0 + 2 * t) = make_float2(c[2] * scale, c[3] * scale);
}

template <int DK, int NT>
__global__ void __launch_bounds__(NT) attn_small_fwd_mma_kernel(AttnArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char sa_raw[];
  SaSmem& S = *reinterpret_cast<SaSmem*>(sa_raw);
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  pdl_wait();
  pdl_trigger();
  SA_STAMP(0);
  {
    SaTile<DK, NT> tq, tk, tv;
    sa_fetch<DK, NT>(tq, a.q, b, a.Tq, a.ldq, h);
    sa_fetch<DK, NT>(tk, a.k, b, a.Tk, a.ldk, h);
    sa_fetch<DK, NT>(tv, a.v, b, a.Tk, a.ldv, h);
    sa_load_kmask(S.kmask, a.keypad, b, a.Tk);
    sa_store<DK, NT>(S.q, tq); sa_store<DK, NT>(S.k, tk); sa_store<DK, NT>(S.v, tv);
  }
  __syncthreads();
  SA_STAMP(1);
  const int Tq = a.Tq, Tk = a.Tk;
  const int mt = (Tq + 15) >> 4, nt = (Tk + 7) >> 3;
  for (int tile = w; tile < mt * nt; tile += NW) {               // masked, scaled scores -> S.p
    const int m0 = (tile / nt) << 4, n0 = (tile % nt) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    sa_mma_nt(c, S.q, S.k, m0, n0, DK, g, t);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = m0 + g + ((e >> 1) << 3), j = n0 + 2 * t + (e & 1);
      const bool ok = i < Tq && !S.kmask[j] && !(a.causal && j > i);
      S.p[i][j] = ok ? c[e] * a.inv_temp : -INFINITY;
    }
  }
  __syncthreads();
  SA_STAMP(2);
  const unsigned long long seed = a.drop.p > 0.f ? mtl_eff_seed(a.drop) : 0ull;
  const float drop_p = a.drop.p, inv_keep = a.drop.inv_keep;
  const uint32_t site = a.drop.site;
  const int nt8 = nt << 3;                                        // columns the score tiles wrote
#pragma unroll
  for (int u = 0; u < SA_T / NW; ++u) {                           // softmax over keys, one warp per query row
    const int i = w + NW * u;
    if (i >= ((Tq + 15) & ~15)) continue;                         // warp-uniform; rows [Tq, 16 mt) become zeros for P.V
    const float s0 = lane < nt8 ? S.p[i][lane] : -INFINITY, s1 = lane + 32 < nt8 ? S.p[i][lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(s0, s1));
    float p0 = 0.f, p1 = 0.f;
    if (m != -INFINITY) { p0 = __expf(s0 - m); p1 = __expf(s1 - m); }
    const float l = warp_sum(p0 + p1);
    const float inv = i < Tq ? 1.f / l : 0.f;                     // fully masked REAL row -> 0 * inf = NaN, like the reference
    p0 *= inv; p1 *= inv;
    if (drop_p > 0.f && i < Tq) {
      const unsigned long long base = ((unsigned long long)(b * a.H + h) * Tq + i) * Tk;
      if (s0 != -INFINITY) p0 *= dropout_scale(seed, site, base + lane, drop_p, inv_keep);
      if (s1 != -INFINITY) p1 *= dropout_scale(seed, site, base + lane + 32, drop_p, inv_keep);
    }
    S.p[i][lane] = p0; S.p[i][lane + 32] = p1;
    if (lane == 0 && i < Tq) a.lse[((size_t)b * a.H + h) * Tq + i] = m + logf(l);
  }
  __syncthreads();
  SA_STAMP(3);
  float* out = a.o + (size_t)b * Tq * a.ldo + h * DK;
  for (int tile = w; tile < mt * (DK / 8); tile += NW) {          // O = P . V
    const int m0 = (tile / (DK / 8)) << 4, n0 = (tile % (DK / 8)) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    sa_mma_nn(c, S.p, S.v, m0, n0, nt8, g, t);
    sa_store_frag(out, a.ldo, c, m0, n0, Tq, g, t, 1.f);
  }
  SA_STAMP(4);
}

template <int DK, int NT>
__global__ void __launch_bounds__(NT) attn_small_bwd_mma_kernel(AttnBwdArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char sa_raw[];
  SaSmem& S = *reinterpret_cast<SaSmem*>(sa_raw);
  const AttnArgs& f = a.f;
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  pdl_wait();
  pdl_trigger();
  SA_STAMP(16);
  {
    SaTile<DK, NT> tq, tk, tv, tg, to;
    sa_fetch<DK, NT>(tq, f.q, b, f.Tq, f.ldq, h);
    sa_fetch<DK, NT>(tk, f.k, b, f.Tk, f.ldk, h);
    sa_fetch<DK, NT>(tv, f.v, b, f.Tk, f.ldv, h);
    sa_fetch<DK, NT>(tg, a.d_o, b, f.Tq, f.ldo, h);
    sa_fetch<DK, NT>(to, f.o, b, f.Tq, f.ldo, h);      // O rows, only for delta (S.p is rewritten below)
    sa_load_kmask(S.kmask, f.keypad, b, f.Tk);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + SA_T) {
      const int i = threadIdx.x - 64;
      S.lse[i] = i < f.Tq ? f.lse[((size_t)b * f.H + h) * f.Tq + i] : 0.f;
    }
    sa_store<DK, NT>(S.q, tq); sa_store<DK, NT>(S.k, tk); sa_store<DK, NT>(S.v, tv); sa_store<DK, NT>(S.g, tg); sa_store<DK, NT>(S.p, to);
  }
  __syncthreads();
  SA_STAMP(17);
  for (int i = w; i < SA_T; i += NW) {         // delta_i = dO_i . O_i (== sum_j P_ij dP_ij, also with dropout)
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < DK / 32; ++v) s += S.g[i][lane + 32 * v] * S.p[i][lane + 32 * v];
    s = warp_sum(s);
    if (lane == 0) S.delta[i] = s;
  }
  __syncthreads();
  SA_STAMP(18);
  const int Tq = f.Tq, Tk = f.Tk;
  const int mt = (Tq + 15) >> 4, nt = (Tk + 7) >> 3;
  {
    const unsigned long long seed = f.drop.p > 0.f ? mtl_eff_seed(f.drop) : 0ull;
    for (int tile = w; tile < mt * nt; tile += NW) {              // scores and dP of one 16 x 8 tile -> P (dropped), dS
      const int m0 = (tile / nt) << 4, n0 = (tile % nt) << 3;
      float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
      sa_mma_nt(sc, S.q, S.k, m0, n0, DK, g, t);
      sa_mma_nt(dp, S.g, S.v, m0, n0, DK, g, t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = m0 + g + ((e >> 1) << 3), j = n0 + 2 * t + (e & 1);
        const bool ok = i < Tq && !S.kmask[j] && !(f.causal && j > i);
        const float p = ok ? __expf(sc[e] * f.inv_temp - S.lse[i]) : 0.f;
        float pd = p, d = dp[e];
        if (f.drop.p > 0.f && ok) {
          const unsigned long long idx = (((unsigned long long)(b * f.H + h) * Tq + i) * Tk + j);
          const float s = dropout_scale(seed, f.drop.site, idx, f.drop.p, f.drop.inv_keep);
          pd = p * s; d *= s;
        }
        S.p[i][j] = pd;
        S.ds[i][j] = ok ? p * (d - S.delta[i]) : 0.f;
      }
    }
  }
  __syncthreads();
  SA_STAMP(19);
  // dQ = dS . K / temp ; dK = dS^T . Q / temp ; dV = P^T . dO -- (mt + 2 mk) x DK / 8 tiles of 16 x 8 shared by the warps
  const int mk = (Tk + 15) >> 4, ND = DK / 8;
  const int i8 = mt << 4, j8 = nt << 3;                           // extents of P / dS written above (zeros beyond Tq / Tk)
  // dK / dV tiles read COLUMNS of dS / P up to 16 mk, beyond the 8 nt written above (stale O values there): column j of the
  // transposed operand only reaches output row j, and rows >= Tk are never stored
  float* dq = a.dq + (size_t)b * Tq * f.ldq + h * DK;
  float* dk = a.dk + (size_t)b * Tk * f.ldk + h * DK;
  float* dv = a.dv + (size_t)b * Tk * f.ldv + h * DK;
  for (int tile = w; tile < (mt + 2 * mk) * ND; tile += NW) {
    const int r = tile / ND, n0 = (tile % ND) << 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < mt) {
      sa_mma_nn(c, S.ds, S.k, r << 4, n0, j8, g, t);
      sa_store_frag(dq, f.ldq, c, r << 4, n0, Tq, g, t, f.inv_temp);
    } else if (r < mt + mk) {
      sa_mma_tn(c, S.ds, S.q, (r - mt) << 4, n0, i8, g, t);
      sa_store_frag(dk, f.ldk, c, (r - mt) << 4, n0, Tk, g, t, f.inv_temp);
    } else {
      sa_mma_tn(c, S.p, S.g, (r - mt - mk) << 4, n0, i8, g, t);
      sa_store_frag(dv, f.ldv, c, (r - mt - mk) << 4, n0, Tk, g, t, 1.f);
    }
  }
  SA_STAMP(20);
}

// ----------------------------------------------------------------------------- long sequences on the tensor cores
// Flash-style attention for Tq or Tk > 64 (every utterance longer than 2.6 s: T' = frames / 4; BASELINE configs[3] runs
// T' = 1250) on the same mma.sync.m16n8k8 TF32 fragments as the short-sequence kernels.  A CTA of 4 warps owns 64 query
// rows (forward, dQ) or 64 key rows (dK / dV) of one (b, h); the opposite side streams through shared memory in 64-row
// tiles; a warp owns 16 rows, keeps its 16 x 64 score / probability tile in a warp-private shared-memory strip between the
// two products, and the running max / sum / output accumulators in registers (online softmax).  Scores never reach HBM;
// the backward recomputes probabilities from the saved log-sum-exp.  The CUDA-core tiled kernels above (16 rows per
// CTA, shuffle-broadcast FFMA) took 62 % of the kernel time of a cfg-4 pass: 1.44 / 1.92 / 2.48 ms per launch.
// X3 = 3xTF32 (fp32-grade, default) or single-pass TF32 (AttnArgs.prec = 1: the session's TF32 engine).
constexpr int FA_T = 64;
// pre-split operand fragment: hi / lo tf32 words (lo unused in single-pass TF32)
template <bool X3, int N>
__device__ __forceinline__ void fa_split(const float (&x)[N], uint32_t (&hi)[N], uint32_t (&lo)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    hi[i] = f2tf32(x[i]);
    lo[i] = X3 ? f2tf32(x[i] - __uint_as_float(hi[i])) : 0u;
  }
}
template <bool X3>
__device__ __forceinline__ void fa_mma(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                       const uint32_t (&bl)[2]) {
  if (X3) { mma_tf32(c, al, bh); mma_tf32(c, ah, bl); }
  mma_tf32(c, ah, bh);
}
// c[nt] += X[m0.., :K8] * Y[8 nt.., :K8]^T for the eight 8-column tiles nt: the A fragment is loaded and split ONCE per k-step
template <bool X3>
__device__ __forceinline__ void fa_nt8(float (&c)[8][4], const float (*X)[SA_LD], const float (*Y)[SA_LD], int m0, int K8, int g,
                                       int t) {
#pragma unroll 2
  for (int k0 = 0; k0 < K8; k0 += 8) {
    const float a[4] = {X[m0 + g][k0 + t], X[m0 + g + 8][k0 + t], X[m0 + g][k0 + t + 4], X[m0 + g + 8][k0 + t + 4]};
    uint32_t ah[4], al[4];
    fa_split<X3, 4>(a, ah, al);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float b[2] = {Y[8 * nt + g][k0 + t], Y[8 * nt + g][k0 + t + 4]};
      uint32_t bh[2], bl[2];
      fa_split<X3, 2>(b, bh, bl);
      fa_mma<X3>(c[nt], ah, al, bh, bl);
    }
  }
}
// c[d] += W[0..16, :K8] * Y[:K8, 8 d..] for the ND 8-column tiles d
template <bool X3, int ND>
__device__ __forceinline__ void fa_nn8(float (&c)[ND][4], const float (*W)[SA_LD], const float (*Y)[SA_LD], int K8, int g, int t) {
#pragma unroll 2
  for (int k0 = 0; k0 < K8; k0 += 8) {
    const float a[4] = {W[g][k0 + t], W[g + 8][k0 + t], W[g][k0 + t + 4], W[g + 8][k0 + t + 4]};
    uint32_t ah[4], al[4];
    fa_split<X3, 4>(a, ah, al);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const float b[2] = {Y[k0 + t][8 * d + g], Y[k0 + t + 4][8 * d + g]};
      uint32_t bh[2], bl[2];
      fa_split<X3, 2>(b, bh, bl);
      fa_mma<X3>(c[d], ah, al, bh, bl);
    }
  }
}
// rows [row0, row0 + 64) of the head-h column block of a (B*T, ld) matrix -> tile (rows >= T zero), 128 threads
template <int DK>
__device__ __forceinline__ void fa_load(float (*dst)[SA_LD], const float* src, int b, int T, int ld, int h, int row0) {
  constexpr int C4 = DK / 4;
#pragma unroll
  for (int u = 0; u < FA_T * C4 / 128; ++u) {
    const int i = threadIdx.x + u * 128, r = i / C4, c4 = i % C4, gr = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < T) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(b * T + gr) * ld + h * DK + c4 * 4));
    *reinterpret_cast<float4*>(&dst[r][c4 * 4]) = v;
  }
}
// dropout scales of two ADJACENT elements idx, idx + 1 (one Philox block when they share it)
__device__ __forceinline__ void drop2(unsigned long long seed, uint32_t site, unsigned long long idx, float p, float inv_keep,
                                      float& s0, float& s1) {
  if ((idx & 3ull) != 3ull) {
    const float4 q = dropout_scale4(seed, site, idx & ~3ull, p, inv_keep);
    const int e = (int)(idx & 3ull);
    s0 = e == 0 ? q.x : (e == 1 ? q.y : q.z);
    s1 = e == 0 ? q.y : (e == 1 ? q.z : q.w);
  } else {
    s0 = dropout_scale(seed, site, idx, p, inv_keep);
    s1 = dropout_scale(seed, site, idx + 1, p, inv_keep);
  }
}
struct FaFwdSmem {
  float q[FA_T][SA_LD], k[FA_T][SA_LD], v[FA_T][SA_LD], p[4][16][SA_LD];
  int kmask[FA_T];
};
struct FaBwdSmem {
  float q[FA_T][SA_LD], g[FA_T][SA_LD], k[FA_T][SA_LD], v[FA_T][SA_LD], p[4][16][SA_LD], ds[4][16][SA_LD];
  float lse[FA_T], delta[FA_T];
  int kmask[FA_T];
};
static_assert(sizeof(FaBwdSmem) <= 227 * 1024, "shared memory budget");

template <int DK, bool X3>
__global__ void __launch_bounds__(128) attn_flash_fwd_kernel(AttnArgs a) {
  constexpr int ND = DK / 8;
  extern __shared__ __align__(16) unsigned char fa_raw[];
  FaFwdSmem& S = *reinterpret_cast<FaFwdSmem*>(fa_raw);
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int Tq = a.Tq, Tk = a.Tk;
  fa_load<DK>(S.q, a.q, b, Tq, a.ldq, h, q0);
  float m_[2] = {-INFINITY, -INFINITY}, l_[2] = {0.f, 0.f}, o[ND][4];
#pragma unroll
  for (int d = 0; d < ND; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
  const int row[2] = {q0 + 16 * w + g, q0 + 16 * w + g + 8};
  const unsigned long long seed = a.drop.p > 0.f ? mtl_eff_seed(a.drop) : 0ull;
  const int k_end = a.causal ? min(Tk, q0 + FA_T) : Tk;         // causal: keys beyond this CTA's last query never contribute
  for (int k0 = 0; k0 < k_end; k0 += FA_T) {
    __syncthreads();
    fa_load<DK>(S.k, a.k, b, Tk, a.ldk, h, k0);
    fa_load<DK>(S.v, a.v, b, Tk, a.ldv, h, k0);
    if (threadIdx.x < FA_T) {
      const int kj = k0 + threadIdx.x;
      S.kmask[threadIdx.x] = (kj >= Tk || (a.keypad && a.keypad[(size_t)b * Tk + kj])) ? 1 : 0;
    }
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    fa_nt8<X3>(s, S.q, S.k, 16 * w, DK, g, t);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int jl = 8 * nt + 2 * t + (e & 1), i = row[e >> 1];
        const bool ok = i < Tq && !S.kmask[jl] && !(a.causal && k0 + jl > i);
        s[nt][e] = ok ? s[nt][e] * a.inv_temp : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float corr[2], sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m_[r], mx[r]);
      corr[r] = (mn == -INFINITY) ? 1.f : (m_[r] == -INFINITY ? 0.f : __expf(m_[r] - mn));
      m_[r] = mn;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = (s[nt][e] == -INFINITY) ? 0.f : __expf(s[nt][e] - m_[e >> 1]);
        s[nt][e] = p;
        sum[e >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      l_[r] = l_[r] * corr[r] + sum[r];
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) { o[d][0] *= corr[0]; o[d][1] *= corr[0]; o[d][2] *= corr[1]; o[d][3] *= corr[1]; }
    // probabilities (after dropout) -> the warp's strip
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float p0 = s[nt][2 * r], p1 = s[nt][2 * r + 1];
        if (a.drop.p > 0.f && row[r] < Tq) {
          float d0, d1;
          drop2(seed, a.drop.site, (((unsigned long long)(b * a.H + h) * Tq + row[r]) * Tk + k0 + 8 * nt + 2 * t), a.drop.p,
                a.drop.inv_keep, d0, d1);
          p0 *= d0; p1 *= d1;
        }
        *reinterpret_cast<float2*>(&S.p[w][g + 8 * r][8 * nt + 2 * t]) = make_float2(p0, p1);
      }
    __syncwarp();
    fa_nn8<X3, ND>(o, S.p[w], S.v, FA_T, g, t);
  }
  float* out = a.o + (size_t)b * Tq * a.ldo + h * DK;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (row[r] >= Tq) continue;
    const float inv = 1.f / l_[r];                               // fully masked row -> NaN, like softmax over a row of -inf
#pragma unroll
    for (int d = 0; d < ND; ++d)
      *reinterpret_cast<float2*>(out + (size_t)row[r] * a.ldo + 8 * d + 2 * t) = make_float2(o[d][2 * r] * inv, o[d][2 * r + 1] * inv);
    if (t == 0) a.lse[((size_t)b * a.H + h) * Tq + row[r]] = m_[r] + logf(l_[r]);
  }
}

// dQ: 64 query rows per CTA, key tiles stream by
template <int DK, bool X3>
__global__ void __launch_bounds__(128) attn_flash_dq_kernel(AttnBwdArgs a) {
  constexpr int ND = DK / 8;
  extern __shared__ __align__(16) unsigned char fa_raw[];
  FaBwdSmem& S = *reinterpret_cast<FaBwdSmem*>(fa_raw);
  const AttnArgs& f = a.f;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int Tq = f.Tq, Tk = f.Tk;
  fa_load<DK>(S.q, f.q, b, Tq, f.ldq, h, q0);
  fa_load<DK>(S.g, a.d_o, b, Tq, f.ldo, h, q0);
  const int row[2] = {q0 + 16 * w + g, q0 + 16 * w + g + 8};
  float lse[2], dl[2], acc[ND][4];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const size_t si = ((size_t)b * f.H + h) * Tq + min(row[r], Tq - 1);
    lse[r] = f.lse[si]; dl[r] = a.delta[si];
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) acc[d][0] = acc[d][1] = acc[d][2] = acc[d][3] = 0.f;
  const unsigned long long seed = f.drop.p > 0.f ? mtl_eff_seed(f.drop) : 0ull;
  const int k_end = f.causal ? min(Tk, q0 + FA_T) : Tk;
  for (int k0 = 0; k0 < k_end; k0 += FA_T) {
    __syncthreads();
    fa_load<DK>(S.k, f.k, b, Tk, f.ldk, h, k0);
    fa_load<DK>(S.v, f.v, b, Tk, f.ldv, h, k0);
    if (threadIdx.x < FA_T) {
      const int kj = k0 + threadIdx.x;
      S.kmask[threadIdx.x] = (kj >= Tk || (f.keypad && f.keypad[(size_t)b * Tk + kj])) ? 1 : 0;
    }
    __syncthreads();
    float sc8[8][4], dp8[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { sc8[nt][0] = sc8[nt][1] = sc8[nt][2] = sc8[nt][3] = 0.f; dp8[nt][0] = dp8[nt][1] = dp8[nt][2] = dp8[nt][3] = 0.f; }
    fa_nt8<X3>(sc8, S.q, S.k, 16 * w, DK, g, t);
    fa_nt8<X3>(dp8, S.g, S.v, 16 * w, DK, g, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float (&sc)[4] = sc8[nt];
      const float (&dp)[4] = dp8[nt];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float ds2[2];
        float d0 = 1.f, d1 = 1.f;
        if (f.drop.p > 0.f && row[r] < Tq)
          drop2(seed, f.drop.site, (((unsigned long long)(b * f.H + h) * Tq + row[r]) * Tk + k0 + 8 * nt + 2 * t), f.drop.p,
                f.drop.inv_keep, d0, d1);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int jl = 8 * nt + 2 * t + c, i = row[r];
          const bool ok = i < Tq && !S.kmask[jl] && !(f.causal && k0 + jl > i);
          const float p = ok ? __expf(sc[2 * r + c] * f.inv_temp - lse[r]) : 0.f;
          ds2[c] = ok ? p * (dp[2 * r + c] * (c ? d1 : d0) - dl[r]) : 0.f;
        }
        *reinterpret_cast<float2*>(&S.ds[w][g + 8 * r][8 * nt + 2 * t]) = make_float2(ds2[0], ds2[1]);
      }
    }
    __syncwarp();
    fa_nn8<X3, ND>(acc, S.ds[w], S.k, FA_T, g, t);
  }
  float* dq = a.dq + (size_t)b * Tq * f.ldq + h * DK;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (row[r] >= Tq) continue;
#pragma unroll
    for (int d = 0; d < ND; ++d)
      *reinterpret_cast<float2*>(dq + (size_t)row[r] * f.ldq + 8 * d + 2 * t) =
          make_float2(acc[d][2 * r] * f.inv_temp, acc[d][2 * r + 1] * f.inv_temp);
  }
}

// dK, dV: 64 key rows per CTA, query tiles stream by; the warp's tiles are TRANSPOSED (rows = its 16 keys, columns = queries)
template <int DK, bool X3>
__global__ void __launch_bounds__(128) attn_flash_dkv_kernel(AttnBwdArgs a) {
  constexpr int ND = DK / 8;
  extern __shared__ __align__(16) unsigned char fa_raw[];
  FaBwdSmem& S = *reinterpret_cast<FaBwdSmem*>(fa_raw);
  const AttnArgs& f = a.f;
  const int b = blockIdx.z, h = blockIdx.y, kb0 = blockIdx.x * FA_T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int Tq = f.Tq, Tk = f.Tk;
  fa_load<DK>(S.k, f.k, b, Tk, f.ldk, h, kb0);
  fa_load<DK>(S.v, f.v, b, Tk, f.ldv, h, kb0);
  const int key[2] = {kb0 + 16 * w + g, kb0 + 16 * w + g + 8};
  bool kvalid[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) kvalid[r] = key[r] < Tk && !(f.keypad && f.keypad[(size_t)b * Tk + key[r]]);
  float accK[ND][4], accV[ND][4];
#pragma unroll
  for (int d = 0; d < ND; ++d) { accK[d][0] = accK[d][1] = accK[d][2] = accK[d][3] = 0.f; accV[d][0] = accV[d][1] = accV[d][2] = accV[d][3] = 0.f; }
  const unsigned long long seed = f.drop.p > 0.f ? mtl_eff_seed(f.drop) : 0ull;
  const int q_begin = f.causal ? kb0 : 0;                       // causal: earlier queries never see these keys (kb0 % 64 == 0)
  for (int q0 = q_begin; q0 < Tq; q0 += FA_T) {
    __syncthreads();
    fa_load<DK>(S.q, f.q, b, Tq, f.ldq, h, q0);
    fa_load<DK>(S.g, a.d_o, b, Tq, f.ldo, h, q0);
    if (threadIdx.x < FA_T) {
      const size_t si = ((size_t)b * f.H + h) * Tq + min(q0 + (int)threadIdx.x, Tq - 1);
      S.lse[threadIdx.x] = f.lse[si];
      S.delta[threadIdx.x] = a.delta[si];
    }
    __syncthreads();
    float sc8[8][4], dp8[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { sc8[nt][0] = sc8[nt][1] = sc8[nt][2] = sc8[nt][3] = 0.f; dp8[nt][0] = dp8[nt][1] = dp8[nt][2] = dp8[nt][3] = 0.f; }
    fa_nt8<X3>(sc8, S.k, S.q, 16 * w, DK, g, t);                   // S^T: rows = keys, columns = queries
    fa_nt8<X3>(dp8, S.v, S.g, 16 * w, DK, g, t);                   // dP^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float (&sc)[4] = sc8[nt];
      const float (&dp)[4] = dp8[nt];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float pd2[2], ds2[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int il = 8 * nt + 2 * t + c, i = q0 + il, j = key[r];
          const bool ok = kvalid[r] && i < Tq && !(f.causal && j > i);
          const float p = ok ? __expf(sc[2 * r + c] * f.inv_temp - S.lse[il]) : 0.f;
          float scale = 1.f;
          if (f.drop.p > 0.f && ok)
            scale = dropout_scale(seed, f.drop.site, (((unsigned long long)(b * f.H + h) * Tq + i) * Tk + j), f.drop.p, f.drop.inv_keep);
          pd2[c] = p * scale;
          ds2[c] = ok ? p * (dp[2 * r + c] * scale - S.delta[il]) : 0.f;
        }
        *reinterpret_cast<float2*>(&S.p[w][g + 8 * r][8 * nt + 2 * t]) = make_float2(pd2[0], pd2[1]);
        *reinterpret_cast<float2*>(&S.ds[w][g + 8 * r][8 * nt + 2 * t]) = make_float2(ds2[0], ds2[1]);
      }
    }
    __syncwarp();
    fa_nn8<X3, ND>(accV, S.p[w], S.g, FA_T, g, t);                 // dV += P_dropped^T . dO
    fa_nn8<X3, ND>(accK, S.ds[w], S.q, FA_T, g, t);                // dK += dS^T . Q
  }
  float* dk = a.dk + (size_t)b * Tk * f.ldk + h * DK;
  float* dv = a.dv + (size_t)b * Tk * f.ldv + h * DK;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (key[r] >= Tk) continue;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      *reinterpret_cast<float2*>(dk + (size_t)key[r] * f.ldk + 8 * d + 2 * t) =
          make_float2(accK[d][2 * r] * f.inv_temp, accK[d][2 * r + 1] * f.inv_temp);
      *reinterpret_cast<float2*>(dv + (size_t)key[r] * f.ldv + 8 * d + 2 * t) = make_float2(accV[d][2 * r], accV[d][2 * r + 1]);
    }
  }
}

// MTL_ATTN_FLASH=0: the CUDA-core tiled kernels for long sequences (A/B measurements; also AttnArgs.prec = 2, the fp32 engine)
static bool attn_flash_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_ATTN_FLASH"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
static bool attn_flash_ok(const AttnArgs& a) {
  auto al = [](const void* p) { return (((uintptr_t)p) & 15u) == 0; };
  return attn_flash_enabled() && a.prec != 2 && a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0 &&
         al(a.q) && al(a.k) && al(a.v) && al(a.o);
}
template <int DK, bool X3>
static int attn_flash_fwd_launch(const AttnArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_fwd_kernel<DK, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FaFwdSmem)));
    configured = true;
  }
  attn_flash_fwd_kernel<DK, X3><<<dim3(mtl_cdiv(a.Tq, FA_T), a.H, a.B), 128, sizeof(FaFwdSmem), s>>>(a);
  MTL_CHECK_LAUNCH();
  ++g_mtl_launches;
  return MTL_OK;
}
template <int DK, bool X3>
static int attn_flash_bwd_launch(const AttnBwdArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_dq_kernel<DK, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FaBwdSmem)));
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_flash_dkv_kernel<DK, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FaBwdSmem)));
    configured = true;
  }
  const AttnArgs& f = a.f;
  attn_delta_kernel<DK><<<mtl_cdiv(f.B * f.Tq * f.H, 4), 128, 0, s>>>(a);
  MTL_CHECK_LAUNCH();
  attn_flash_dq_kernel<DK, X3><<<dim3(mtl_cdiv(f.Tq, FA_T), f.H, f.B), 128, sizeof(FaBwdSmem), s>>>(a);
  MTL_CHECK_LAUNCH();
  attn_flash_dkv_kernel<DK, X3><<<dim3(mtl_cdiv(f.Tk, FA_T), f.H, f.B), 128, sizeof(FaBwdSmem), s>>>(a);
  MTL_CHECK_LAUNCH();
  g_mtl_launches += 3;
  return MTL_OK;
}
int k_attn_flash_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.prec == 1) return a.dk == 64 ? attn_flash_fwd_launch<64, false>(a, s) : attn_flash_fwd_launch<32, false>(a, s);
  return a.dk == 64 ? attn_flash_fwd_launch<64, true>(a, s) : attn_flash_fwd_launch<32, true>(a, s);
}
int k_attn_flash_bwd(const AttnBwdArgs& a, cudaStream_t s) {
  if (a.f.prec == 1) return a.f.dk == 64 ? attn_flash_bwd_launch<64, false>(a, s) : attn_flash_bwd_launch<32, false>(a, s);
  return a.f.dk == 64 ? attn_flash_bwd_launch<64, true>(a, s) : attn_flash_bwd_launch<32, true>(a, s);
}
bool k_attn_flash_eligible(const AttnArgs& a) { return attn_flash_ok(a); }

// MTL_ATTN_SMALL=0 keeps the tiled kernels for short sequences too (A/B measurements)
static bool attn_small_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_ATTN_SMALL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
static bool attn_small_ok(const AttnArgs& a) {
  auto al = [](const void* p) { return (((uintptr_t)p) & 15u) == 0; };
  return attn_small_enabled() && a.Tq <= SA_T && a.Tk <= SA_T && a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 &&
         a.ldo % 4 == 0 && al(a.q) && al(a.k) && al(a.v) && al(a.o);
}
// MTL_ATTN_THREADS=256 runs the short-sequence kernels with 8 warps per CTA instead of 16 (A/B measurements)
static int attn_small_threads() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_ATTN_THREADS"); v = (e && atoi(e) == 256) ? 256 : 512; }
  return v;
}
// MTL_ATTN_MMA=0: the CUDA-core FFMA versions of the short-sequence kernels (A/B measurements)
static bool attn_mma_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_ATTN_MMA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
template <int DK, int NT>
static int attn_small_fwd_launch(const AttnArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<DK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaSmem)));
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_small_fwd_mma_kernel<DK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaSmem)));
    configured = true;
  }
  auto kern = attn_mma_enabled() ? attn_small_fwd_mma_kernel<DK, NT> : attn_small_fwd_kernel<DK, NT>;
  MTL_CHECK_CUDA(mtl_launch_pdl(kern, dim3(a.H, a.B), dim3(NT), sizeof(SaSmem), s, a));
  ++g_mtl_launches;
  return MTL_OK;
}
template <int DK, int NT>
static int attn_small_bwd_launch(const AttnBwdArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<DK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaSmem)));
    MTL_CHECK_CUDA(cudaFuncSetAttribute(attn_small_bwd_mma_kernel<DK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SaSmem)));
    configured = true;
  }
  auto kern = attn_mma_enabled() ? attn_small_bwd_mma_kernel<DK, NT> : attn_small_bwd_kernel<DK, NT>;
  MTL_CHECK_CUDA(mtl_launch_pdl(kern, dim3(a.f.H, a.f.B), dim3(NT), sizeof(SaSmem), s, a));
  ++g_mtl_launches;
  return MTL_OK;
}
int k_attn_small_fwd(const AttnArgs& a, cudaStream_t s) {
  if (attn_small_threads() == 256) return a.dk == 64 ? attn_small_fwd_launch<64, 256>(a, s) : attn_small_fwd_launch<32, 256>(a, s);
  return a.dk == 64 ? attn_small_fwd_launch<64, 512>(a, s) : attn_small_fwd_launch<32, 512>(a, s);
}
int k_attn_small_bwd(const AttnBwdArgs& a, cudaStream_t s) {
  if (attn_small_threads() == 256) return a.f.dk == 64 ? attn_small_bwd_launch<64, 256>(a, s) : attn_small_bwd_launch<32, 256>(a, s);
  return a.f.dk == 64 ? attn_small_bwd_launch<64, 512>(a, s) : attn_small_bwd_launch<32, 512>(a, s);
}
bool k_attn_small_eligible(const AttnArgs& a) { return attn_small_ok(a); }
