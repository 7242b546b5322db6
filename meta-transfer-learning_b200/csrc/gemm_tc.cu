// Blackwell tensor-core GEMM for the GemmArgs contract (kernels.h): TMA (cp.async.bulk.tensor, 128B
// swizzle) stages fp32 operand tiles into shared memory, one elected thread issues
// tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8 per instruction) with the accumulator in
// tensor memory, and four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue
// (alpha, bias, ReLU, ReLU-backward mask, beta*C, or atomic split-K accumulation).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two CTAs are co-resident per SM so one
// tile's epilogue overlaps the other's main loop.
//
// Operand majors map onto UMMA shared-memory descriptors (SWIZZLE_128B, sm_100 descriptor version 1):
//   K-major  (A stored [M,K], B stored [N,K]):  TMA box {32 fp32 of K, rows}; rows are 128 B, 8-row
//            swizzle atoms stacked every 1024 B (SBO = 1024); each K=8 step advances the start by 32 B.
//   MN-major (A stored [K,M], B stored [K,N]):  TMA boxes {32 fp32 of M/N, 32 rows of K} with the
//            128B_ATOM_32B swizzle (the only legal MN-major layout for 32-bit operands is
//            SWIZZLE_128B_BASE32B), one 4096 B box per 32 columns (LBO = 4096), 4-row atoms every 512 B
//            (SBO = 512); each K=8 step advances the start by 1024 B.
// Out-of-range rows/columns/K are zero-filled by TMA, so ragged M, N, K need no special casing.
#include "kernels.h"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <unordered_map>

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per k-block = one 128 B swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must trap (error reported to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset=0, layout_type=SWIZZLE_128B(2) [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major tile: SWIZZLE_128B (2), 8-row atoms every 1024 B.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return umma_desc(saddr, 16, 1024, 2); }
// MN-major fp32/tf32 tile: the only legal layout is SWIZZLE_128B_BASE32B (1) -- 32 B swizzle granules,
// atoms of 32 (MN) x 4 (K) elements = 512 B stacked along K (SBO = 512), 32-column groups 4096 B apart (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) { return umma_desc(saddr, 4096, 512, 1); }

struct TcParams {
  GemmArgs g;
  int kb_total;      // number of 32-wide k-blocks
  int kb_per_split;  // k-blocks per blockIdx.z
  int vecC;          // C (and aux) rows are 16 B aligned
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const TcParams P) {
  constexpr int B_STAGE_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const GemmArgs& g = P.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * P.kb_per_split;
  const int kb1 = min(P.kb_total, kb0 + P.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_STAGE_BYTES;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, &full[s], kb * BK, m0);
        } else {
#pragma unroll
          for (int i = 0; i < BM / 32; ++i) tma_load_2d(sa + i * 4096, &tmA, &full[s], m0 + i * 32, kb * BK);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, &full[s], kb * BK, n0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 32; ++i) tma_load_2d(sb + i * 4096, &tmB, &full[s], n0 + i * 32, kb * BK);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t da = A_MN ? desc_mnmajor(sa + k * 1024) : desc_kmajor(sa + k * 32);
          const uint64_t db = B_MN ? desc_mnmajor(sb + k * 1024) : desc_kmajor(sb + k * 32);
          tc_mma_tf32(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);          // frees the smem slot once these MMAs have read it
      }
      tc_commit(tmem_full);            // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      const int col0 = n0 + c * 32;
      if (row < g.M && col0 < g.N) {
        float* crow = g.C + (long long)row * g.ldc + col0;
        if (g.split_k > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < g.N) atomicAdd(crow + j, g.alpha * __uint_as_float(r[j]));
        } else {
          const float* arow = g.epi == EPI_RELU_BWD ? g.aux + (long long)row * g.ldc + col0 : nullptr;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = col0 + j4 + j;
              float t = g.alpha * __uint_as_float(r[j4 + j]);
              if (g.bias && col < g.N) t += g.bias[col];
              if (g.epi == EPI_RELU) t = fmaxf(t, 0.f);
              v[j] = t;
            }
            if (P.vecC && col0 + j4 + 3 < g.N) {
              if (arow) {
                const float4 a = *reinterpret_cast<const float4*>(arow + j4);
                v[0] = a.x > 0.f ? v[0] : 0.f; v[1] = a.y > 0.f ? v[1] : 0.f;
                v[2] = a.z > 0.f ? v[2] : 0.f; v[3] = a.w > 0.f ? v[3] : 0.f;
              }
              if (g.beta != 0.f) {
                const float4 o = *reinterpret_cast<const float4*>(crow + j4);
                v[0] += g.beta * o.x; v[1] += g.beta * o.y; v[2] += g.beta * o.z; v[3] += g.beta * o.w;
              }
              *reinterpret_cast<float4*>(crow + j4) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col0 + j4 + j < g.N) {
                  float t = v[j];
                  if (arow) t = arow[j4 + j] > 0.f ? t : 0.f;
                  if (g.beta != 0.f) t += g.beta * crow[j4 + j];
                  crow[j4 + j] = t;
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* ptr; long long inner, outer, ld; int box_outer, dtype;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer && dtype == o.dtype;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ (size_t)k.inner; h = h * 1000003u ^ (size_t)k.outer; h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)(k.box_outer * 4 + k.dtype);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// TMA data type for the fp32 operands: TFLOAT32 (default) lets the copy engine deliver TF32-formatted
// values; MTL_TMA_FP32=1 selects plain FLOAT32 (the tensor core then ignores the low mantissa bits).
int tma_dtype() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_TMA_FP32"); v = (e && e[0] == '1') ? 0 : 1; }
  return v;
}

// 2-D map over a row-major fp32 matrix: `inner` contiguous elements, `outer` rows of stride ld; box {32, box_outer}.
int make_map(const float* ptr, long long inner, long long outer, long long ld, int box_outer, bool mn_major,
             CUtensorMap* out) {
  MapKey key{ptr, inner, outer, ld, box_outer, tma_dtype() + (mn_major ? 2 : 0)};
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return MTL_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { mtl_set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return MTL_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, tma_dtype() ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr,
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mtl_set_error("cuTensorMapEncodeTiled failed (%d) for ptr=%p inner=%lld outer=%lld ld=%lld", (int)r, (const void*)ptr,
                  inner, outer, ld);
    return MTL_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps[key] = *out;
  return MTL_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& P, dim3 grid, cudaStream_t s) {
  constexpr int SMEM = STAGES * (A_STAGE_BYTES + BN * BK * 4) + 1024 /*align slack*/ + 256 /*barriers*/;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN>;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kern<<<grid, 192, SMEM, s>>>(ta, tb, P);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

template <int BN, int STAGES>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& P, dim3 grid,
                   cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, STAGES, false, false>(ta, tb, P, grid, s);
  if (!a_mn && b_mn) return launch<BN, STAGES, false, true>(ta, tb, P, grid, s);
  if (a_mn && !b_mn) return launch<BN, STAGES, true, false>(ta, tb, P, grid, s);
  return launch<BN, STAGES, true, true>(ta, tb, P, grid, s);
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

}  // namespace

bool k_gemm_tc_eligible(const GemmArgs& g) {
  if (g.M < 1 || g.N < 1 || g.K < 1) return false;
  if (!al16(g.A) || !al16(g.B)) return false;
  if (g.lda % 4 != 0 || g.ldb % 4 != 0) return false;
  return true;
}

int k_gemm_tc(const GemmArgs& g, int precision_mode, cudaStream_t s) {
  (void)precision_mode;
  MTL_REQUIRE(k_gemm_tc_eligible(g), "operands not TMA-compatible (16 B aligned base, leading dims % 4 == 0)");
  const bool a_mn = g.transA != 0;     // A stored [K,M]: M contiguous
  const bool b_mn = g.transB == 0;     // B stored [K,N]: N contiguous
  const int bn = g.N <= 64 ? 64 : 128;
  CUtensorMap ta, tb;
  if (!a_mn) MTL_TRY(make_map(g.A, g.K, g.M, g.lda, BM, false, &ta)); else MTL_TRY(make_map(g.A, g.M, g.K, g.lda, 32, true, &ta));
  if (!b_mn) MTL_TRY(make_map(g.B, g.K, g.N, g.ldb, bn, false, &tb)); else MTL_TRY(make_map(g.B, g.N, g.K, g.ldb, 32, true, &tb));
  TcParams P;
  P.g = g;
  P.kb_total = mtl_cdiv(g.K, BK);
  int split = g.split_k > 1 ? g.split_k : 1;
  if (split > 1) {
    MTL_REQUIRE(g.beta == 1.f && g.epi == EPI_NONE && g.bias == nullptr, "split-K needs beta=1, no epilogue");
    if (split > P.kb_total) split = P.kb_total;
    P.kb_per_split = mtl_cdiv(P.kb_total, split);
    split = mtl_cdiv(P.kb_total, P.kb_per_split);
    P.g.split_k = 2;                   // atomic epilogue (also when the split collapsed to one slab)
  } else {
    P.kb_per_split = P.kb_total;
    P.g.split_k = 1;
  }
  P.vecC = al16(g.C) && (g.ldc % 4 == 0) && (g.epi != EPI_RELU_BWD || al16(g.aux));
  dim3 grid(mtl_cdiv(g.M, BM), mtl_cdiv(g.N, bn), split);
  MTL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm grid too large");
  if (bn == 64) return dispatch_major<64, 4>(a_mn, b_mn, ta, tb, P, grid, s);
  return dispatch_major<128, 3>(a_mn, b_mn, ta, tb, P, grid, s);
}
