// Exact-fp32 CUDA-core GEMM with generic operand majors, fused epilogue and atomic split-K.
// This is the bit-faithful fp32 path of the library (parity mode, odd shapes, tiny problems);
// the tensor-core path for the same GemmArgs contract lives in gemm_tc.cu (tcgen05 + TMA + TMEM).
// C[M,N] = epi(alpha * op(A) op(B) + bias) + beta*C   -- see GemmArgs in kernels.h.
#include "kernels.h"

#define SG_BM 128
#define SG_BN 64
#define SG_BK 16
#define SG_PAD 4

struct SimtParams {
  GemmArgs g;
  long long sam, sak, sbk, sbn;   // element strides
  int vecA, vecB, vecC;           // 16B-vectorisable along the contiguous dim
  int k_chunk;                    // K range per blockIdx.z
};

__device__ __forceinline__ void ldg_tile_a(const SimtParams& P, int m0, int k0, int kmax, float (&r)[8]) {
  const GemmArgs& g = P.g;
  const int tid = threadIdx.x;
  if (P.sak == 1) {              // K contiguous: 4 consecutive k per load, 64 rows per pass
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int m = m0 + (tid >> 2) + ps * 64, k = k0 + (tid & 3) * 4;
      const float* p = g.A + (long long)m * P.sam + k;
      if (m < g.M && P.vecA && k + 3 < kmax) {
        float4 v = *reinterpret_cast<const float4*>(p);
        r[ps * 4 + 0] = v.x; r[ps * 4 + 1] = v.y; r[ps * 4 + 2] = v.z; r[ps * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) r[ps * 4 + j] = (m < g.M && k + j < kmax) ? p[j] : 0.f;
      }
    }
  } else {                       // M contiguous: 4 consecutive m per load, 8 k-rows per pass
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int k = k0 + (tid >> 5) + ps * 8, m = m0 + (tid & 31) * 4;
      const float* p = g.A + (long long)k * P.sak + (long long)m * P.sam;
      if (k < kmax && P.vecA && P.sam == 1 && m + 3 < g.M) {
        float4 v = *reinterpret_cast<const float4*>(p);
        r[ps * 4 + 0] = v.x; r[ps * 4 + 1] = v.y; r[ps * 4 + 2] = v.z; r[ps * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) r[ps * 4 + j] = (k < kmax && m + j < g.M) ? p[(long long)j * P.sam] : 0.f;
      }
    }
  }
}
__device__ __forceinline__ void sts_tile_a(const SimtParams& P, float (*As)[SG_BM + SG_PAD], const float (&r)[8]) {
  const int tid = threadIdx.x;
  if (P.sak == 1) {
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int m = (tid >> 2) + ps * 64, k = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[k + j][m] = r[ps * 4 + j];
    }
  } else {
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int k = (tid >> 5) + ps * 8, m = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&As[k][m]) = make_float4(r[ps * 4], r[ps * 4 + 1], r[ps * 4 + 2], r[ps * 4 + 3]);
    }
  }
}
__device__ __forceinline__ void ldg_tile_b(const SimtParams& P, int n0, int k0, int kmax, float (&r)[4]) {
  const GemmArgs& g = P.g;
  const int tid = threadIdx.x;
  if (P.sbn == 1) {              // N contiguous: [K,N] storage
    const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
    const float* p = g.B + (long long)k * P.sbk + n;
    if (k < kmax && P.vecB && n + 3 < g.N) {
      float4 v = *reinterpret_cast<const float4*>(p);
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (k < kmax && n + j < g.N) ? p[j] : 0.f;
    }
  } else {                       // K contiguous: [N,K] storage
    const int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 4;
    const float* p = g.B + (long long)n * P.sbn + (long long)k * P.sbk;
    if (n < g.N && P.vecB && P.sbk == 1 && k + 3 < kmax) {
      float4 v = *reinterpret_cast<const float4*>(p);
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (n < g.N && k + j < kmax) ? p[(long long)j * P.sbk] : 0.f;
    }
  }
}
__device__ __forceinline__ void sts_tile_b(const SimtParams& P, float (*Bs)[SG_BN + SG_PAD], const float (&r)[4]) {
  const int tid = threadIdx.x;
  if (P.sbn == 1) {
    const int k = tid >> 4, n = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&Bs[k][n]) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    const int n = tid >> 2, k = (tid & 3) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[k + j][n] = r[j];
  }
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams P) {
  const GemmArgs& g = P.g;
  __shared__ __align__(16) float As[SG_BK][SG_BM + SG_PAD];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + SG_PAD];
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int kbeg = blockIdx.z * P.k_chunk;
  const int kmax = min(g.K, kbeg + P.k_chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];
  if (kbeg < kmax) {
    ldg_tile_a(P, m0, kbeg, kmax, ra);
    ldg_tile_b(P, n0, kbeg, kmax, rb);
  }
  for (int k0 = kbeg; k0 < kmax; k0 += SG_BK) {
    __syncthreads();
    sts_tile_a(P, As, ra);
    sts_tile_b(P, Bs, rb);
    __syncthreads();
    if (k0 + SG_BK < kmax) {
      ldg_tile_a(P, m0, k0 + SG_BK, kmax, ra);
      ldg_tile_b(P, n0, k0 + SG_BK, kmax, rb);
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  // ---- epilogue
  const int n = n0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.bias)
#pragma unroll
    for (int j = 0; j < 4; ++j) if (n + j < g.N) bias[j] = g.bias[n + j];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    float* c = g.C + (long long)m * g.ldc + n;
    if (g.split_k > 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < g.N) atomicAdd(c + j, g.alpha * acc[i][j]);
      continue;
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = g.alpha * acc[i][j] + bias[j];
      if (g.epi == EPI_RELU) v[j] = fmaxf(v[j], 0.f);
    }
    if (g.epi == EPI_RELU_BWD) {
      const float* ax = g.aux + (long long)m * g.ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] = ax[j] > 0.f ? v[j] : 0.f;
    }
    if (P.vecC && n + 3 < g.N) {
      if (g.beta != 0.f) {
        float4 o = *reinterpret_cast<const float4*>(c);
        v[0] += g.beta * o.x; v[1] += g.beta * o.y; v[2] += g.beta * o.z; v[3] += g.beta * o.w;
      }
      *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) c[j] = (g.beta != 0.f) ? v[j] + g.beta * c[j] : v[j];
    }
  }
}

static inline bool al16p(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int k_gemm_simt(const GemmArgs& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return MTL_OK;
  MTL_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0, "gemm dims");
  SimtParams P;
  P.g = g;
  if (g.transA) { P.sam = 1; P.sak = g.lda; } else { P.sam = g.lda; P.sak = 1; }
  if (g.transB) { P.sbk = 1; P.sbn = g.ldb; } else { P.sbk = g.ldb; P.sbn = 1; }
  P.vecA = al16p(g.A) && (g.lda % 4 == 0);
  P.vecB = al16p(g.B) && (g.ldb % 4 == 0);
  P.vecC = al16p(g.C) && (g.ldc % 4 == 0) && (g.epi != EPI_RELU_BWD || al16p(g.aux));
  int split = g.split_k > 1 ? g.split_k : 1;
  if (split > 1) {
    MTL_REQUIRE(g.beta == 1.f && g.epi == EPI_NONE && g.bias == nullptr, "split-K needs beta=1, no epilogue");
    int chunk = mtl_cdiv(mtl_cdiv(g.K, split), SG_BK) * SG_BK;
    if (chunk < SG_BK) chunk = SG_BK;
    split = mtl_cdiv(g.K, chunk);
    P.k_chunk = chunk;
    P.g.split_k = split > 1 ? split : 2;   // keep the atomic epilogue even if it collapsed to one slab
  } else {
    P.k_chunk = g.K > 0 ? g.K : 1;
    P.g.split_k = 1;
  }
  dim3 grid(mtl_cdiv(g.M, SG_BM), mtl_cdiv(g.N, SG_BN), split);
  MTL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm grid too large");
  gemm_simt_kernel<<<grid, 256, 0, s>>>(P);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
