// Blackwell tensor-core contraction for the GemmArgs contract (kernels.h) and for the VGG 3x3
// convolutions as IMPLICIT GEMMs (no im2col buffer): TMA (cp.async.bulk.tensor, 128B swizzle) stages
// fp32 operand tiles into shared memory, one elected thread issues tcgen05.mma.cta_group::1.kind::tf32
// (M=128, N=BN, K=8 per instruction) with the accumulator in tensor memory, and four epilogue warps
// drain TMEM with tcgen05.ld and apply the fused epilogue (alpha, bias, ReLU, ReLU-backward mask,
// beta*C, or vectorised-atomic split-K accumulation).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = operand splitter (3xTF32 only) during the main loop, then epilogue
// (TMEM lane quarter = warp_id % 4).
//
// Precision modes
//   TF32   : TMA delivers TFLOAT32-typed (rounded) tiles, one MMA per k-step.
//   3xTF32 : TMA delivers raw fp32; the splitter warps rewrite every staged tile in place as
//            hi = rna_tf32(a) and write lo = rna_tf32(a - hi) to a sibling buffer; three MMAs per k-step
//            (lo*hi + hi*lo + hi*hi) accumulate in fp32 TMEM.  Per-operand residual <= 2^-22 |a| (unbiased)
//            and the dropped lo*lo term is <= 2^-22 |a||b|: fp32-grade products.
//
// Operand majors map onto UMMA shared-memory descriptors (SWIZZLE_128B family, sm_100 descriptor v1):
//   K-major  (A stored [M,K], B stored [N,K]):  TMA box {32 fp32 of K, rows}; rows are 128 B, 8-row
//            swizzle atoms stacked every 1024 B (SBO = 1024); each K=8 step advances the start by 32 B.
//   MN-major (A stored [K,M], B stored [K,N]):  TMA boxes {32 fp32 of M/N, 32 rows of K} with the
//            128B_ATOM_32B swizzle (the only legal MN-major layout for 32-bit operands is
//            SWIZZLE_128B_BASE32B), one 4096 B box per 32 columns (LBO = 4096), 4-row atoms every 512 B
//            (SBO = 512); each K=8 step advances the start by 1024 B.
// Out-of-range rows/columns/K are zero-filled by TMA, so ragged M, N, K need no special casing.
//
// Implicit convolution (NHWC activations [B,F,T,C], 3x3, stride 1, zero pad 1 -- models/asr/transformer.py:47-59)
//   CONV_FWD  : C[pixel, co] = sum_{tap,ci} X[pixel + tap, ci] * Wg[co, tap*Cin + ci]   (also the dgrad, on dY with
//               flipped taps).  The A tile of k-block (tap, 32-channel chunk) is ONE 4-D TMA box
//               {32 c, BT t, 128/BT f, 1} fetched at the tap-shifted coordinate; the zero padding is TMA's
//               out-of-bounds fill.  Tile rows are pixels in (f, t) order.
//   CONV_WGRAD: dWgT[tap*Cin + ci, co] = sum_pixel X[pixel + tap, ci] * dY[pixel, co].  K runs over 32-pixel
//               boxes {BT t, 32/BT f}; A boxes (MN-major) are tap-shifted X boxes, B boxes are dY boxes.
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug must trap (error reported to the host), never hang the GPU.
// debug stamps (MTL_GEMM_DBG=1): SM cycle counter at the main events of CTA (0,0,0)
__device__ long long g_dbg[160];   // [0,32): phase stamps; [32,160): per-k-block pipeline stamps of CTA (0,0,0), 4 per k-block
__device__ unsigned long long g_cta_span[512];   // MTL_GEMM_DBG=99: globaltimer at entry / exit of the first 256 CTAs of the last launch
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DBG_SPAN(slot) do { if (P.dbg == 99 && threadIdx.x == 0) { const unsigned id = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (id < 256) g_cta_span[2 * id + (slot)] = globaltimer_ns(); } } while (0)
#define DBG_STAMP(i) do { if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_dbg[i] = clock64(); } while (0)
// role r (0 producer issued, 1 splitter saw full, 2 MMA saw ready/full, 3 MMA issued+committed) of k-block iteration it
#define DBG_KB(it, r) do { if (P.dbg && (it) < 32 && blockIdx.x == P.dbg - 1 && blockIdx.y == 0 && blockIdx.z == 0) g_dbg[32 + (it) * 4 + (r)] = clock64(); } while (0)
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
// Non-blocking probe of a phase (the kw-box producer polls two rings from one thread).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA stores of a staged tile: plain, or reduce-add (C += tile, performed by the L2 -- replaces read-modify-write and atomics)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3,
                                             bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
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
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld32_issue(taddr, v);
  tmem_ld_wait();
}
// round-to-nearest (ties away from zero) fp32 -> tf32: 10 explicit mantissa bits, low 13 bits zero.  Same result as
// cvt.rna.tf32.f32 for finite inputs, but two full-rate integer ops instead of a quarter-rate conversion-pipe
// instruction: the splitter converts 12288 values per k-block, which at 16 conversions/clk/SM was ~770 cycles.
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// Truncating split: hi = the top 19 bits of the fp32 word -- exactly what the tensor core reads from a raw fp32 tile under
// kind::tf32 (it ignores the low 13 mantissa bits), so the hi operand needs NO rewrite; lo = rna_tf32(a - hi) carries the
// next 11 bits.  a - hi is exact, so the only rounding is lo's (<= 2^-21 |a|, unbiased); the dropped lo*lo term is
// <= 2^-20 |a||b|.  The splitter then moves 1 load + 1 store per element instead of 1 + 2 (shared-memory traffic is
// what bounds it).
__device__ __forceinline__ float tf32_lo_trunc(float x) {
  return tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}
// Pins a loop-invariant value in a register (opaque to the optimiser, so it cannot be rematerialised from the
// constant bank at every use).
__device__ __forceinline__ int keep_reg(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ float keep_reg(float v) { asm volatile("" : "+f"(v)); return v; }
template <typename T>
__device__ __forceinline__ T* keep_reg(T* p) { asm volatile("" : "+l"(p)); return p; }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- thread-block-cluster helpers (cluster = the split-K group of one output tile)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_saddr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(cta));
  return ra;
}
// 16-byte load from a shared::cluster address (own CTA's shared memory when the address was not mapa'd)
__device__ __forceinline__ float4 ld_cluster_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset=0, layout_type [61,64).
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
// MN-major fp32/tf32 tile: SWIZZLE_128B_BASE32B (1) -- 32 B swizzle granules, atoms of 32 (MN) x 4 (K)
// elements = 512 B stacked along K (SBO = 512), 32-column groups 4096 B apart (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) { return umma_desc(saddr, 4096, 512, 1); }

enum { CONV_NONE = 0, CONV_FWD = 1, CONV_WGRAD = 2 };

struct TcParams {
  GemmArgs g;
  int kb_total;      // number of 32-wide k-blocks
  int kb_per_split;  // k-blocks per blockIdx.z
  int vecC;          // C (and aux) rows are 16 B aligned
  // implicit-convolution geometry (conv_mode != CONV_NONE)
  int conv_mode;
  int cF, cT, cCin;  // activation extents of the tap-shifted operand
  int bt_log2;       // pixel box: (1 << bt_log2) time steps wide
  int tiles_t, tiles_f;
  int dbg;
  int aux_off;       // TMA epilogue: byte offset (from the aligned smem base) of the ReLU-mask tile
  int tma_epi;       // epilogue through TMA: 1 = store, 2 = reduce-add (beta == 1 or split-K slabs)
  int epi_test;      // measurement only (MTL_EPI_TEST): 1 = conv epilogue computes but does not store
  int cluster_k;     // > 1: grid.z CTAs form one cluster that splits K and reduces over distributed shared memory
  int split_trunc;   // 3xTF32 splitter: 1 = truncating split (hi stays the raw tile), 0 = round-to-nearest hi written in place
  int lin_stages;    // > 0: plain GEMM on the small-footprint pipeline (this many stages; see lin_small_stages)
  int pool;          // kw-box forward convolution: also store the 2x2 max-pooled tile (ReLU + MaxPool2d(2, 2)) through the tmX map
};

// ---- TMA epilogue (shared by the GEMM / tap-box convolution kernel and the kw-box convolution kernel).
// Runs in the four epilogue warps (threads 64..191): each thread owns one accumulator row (TMEM lane); per 32-column
// chunk it drains TMEM, applies alpha / bias / ReLU / ReLU-backward mask in registers and parks the row in a
// 128B-swizzled staging box; one thread then hands every box to the TMA unit (store, or reduce-add for beta == 1
// and split-K slabs).  Row / column tails and the ragged edges of convolution pixel boxes are clipped by the TMA
// unit, so there is no per-row address arithmetic at all.  stg_u / aux_u: 1024-aligned shared-memory addresses of
// the staging tile and of the mask tile (both BN/32 boxes of 128 rows x 128 B); they alias dead operand buffers.
template <int BN, bool SPLIT3, bool POOL = false>
__device__ __forceinline__ void tma_epilogue_rows(const TcParams& P, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                                  uint32_t stg_u, uint32_t aux_u, uint64_t* tmem_full,
                                                  uint64_t* aux_full, const float* bias_s, uint32_t tmem_base, int warp,
                                                  int lane, int m0, int n0, bool conv, int ct0, int cf0, int cb,
                                                  uint32_t aux_par = 0) {
  constexpr int CHUNKS = BN / 32, CHUNK_BYTES = BM * 128;
  const GemmArgs& g = P.g;
  const int q = warp & 3;                                        // TMEM lane quarter this warp may access
  const bool mask = g.epi == EPI_RELU_BWD;
  const int nch = min(CHUNKS, (g.N - n0 + 31) / 32);            // chunks with at least one valid column
  mbar_wait(tmem_full, 0);                                       // accumulator complete => operand buffers are dead
  tc_fence_after();
  // The staging tile aliases operand stages these same four warps read / rewrote as splitters.  The mbarrier chain
  // (splitter -> ready -> MMA -> commit -> tmem_full) already orders those accesses before the stores below; the named
  // barrier states the same order in a form compute-sanitizer's racecheck can see (it does not model mbarriers).
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 64) DBG_STAMP(5);
  if (mask && threadIdx.x == 64) {
    mbar_expect_tx(aux_full, (uint32_t)(nch * CHUNK_BYTES));
    for (int c = 0; c < nch; ++c) {
      if (conv) tma_load_4d(aux_u + c * CHUNK_BYTES, tmX, aux_full, n0 + c * 32, ct0, cf0, cb);
      else tma_load_2d(aux_u + c * CHUNK_BYTES, tmX, aux_full, n0 + c * 32, m0);
    }
  }
  const int row = q * 32 + lane;
  const uint32_t row_u = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
  const float alpha = g.alpha;
  const float lo = g.epi == EPI_RELU ? 0.f : -INFINITY;          // branch-free ReLU
  // Fused 2x2 max-pool (models/asr/transformer.py:51,58): tile rows are pixels in (f, t) order, t fastest over 8, so the
  // window partners of a row are lane ^ 1 (t) and lane ^ 8 (f); lanes with even t and even f park the pooled pixel in a
  // second staging tile (32 rows = 8 f' x 4 t', where the unused mask tile would be) that goes out through tmX.  Floor
  // pooling and ragged edges need no code: windows that reach past the image have pooled coordinates past F/2 or T/2 and
  // are clipped by the TMA unit.
  const bool pool = POOL && conv && !mask;       // compile-time: the GEMM / weight-gradient instantiations carry none of it
  const uint32_t prow = (uint32_t)((((row >> 3) & 15) >> 1) * 4 + ((row & 7) >> 1));   // pooled row f' * 4 + t'
  const bool pool_owner = (lane & 9) == 0;
  if (mask) mbar_wait(aux_full, aux_par);
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint32_t v[32];
    if (SPLIT3) {                                                // 3xTF32: columns [BN, 2BN) hold the a_hi*b_lo products
      uint32_t w[32];
      // both TMEM loads in flight before the one wait (each round trip is ~300 cycles of a latency-bound epilogue)
      tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c * 32), w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    } else {
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j * 4);
      float4 o;
      o.x = fmaxf(fmaf(alpha, __uint_as_float(v[4 * j + 0]), b4.x), lo);
      o.y = fmaxf(fmaf(alpha, __uint_as_float(v[4 * j + 1]), b4.y), lo);
      o.z = fmaxf(fmaf(alpha, __uint_as_float(v[4 * j + 2]), b4.z), lo);
      o.w = fmaxf(fmaf(alpha, __uint_as_float(v[4 * j + 3]), b4.w), lo);
      const uint32_t slot = (uint32_t)c * CHUNK_BYTES + row_u + ((((uint32_t)j) ^ sw) << 4);
      if (mask) {
        float4 a;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(aux_u + slot));
        o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f;
        o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
      }
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_u + slot), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
      if (pool) {
        float4 m = o;
        m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, 1)); m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, 1));
        m.z = fmaxf(m.z, __shfl_xor_sync(0xffffffffu, m.z, 1)); m.w = fmaxf(m.w, __shfl_xor_sync(0xffffffffu, m.w, 1));
        m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, 8)); m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, 8));
        m.z = fmaxf(m.z, __shfl_xor_sync(0xffffffffu, m.z, 8)); m.w = fmaxf(m.w, __shfl_xor_sync(0xffffffffu, m.w, 8));
        if (pool_owner) {
          const uint32_t pslot = (uint32_t)c * 4096u + prow * 128u + ((((uint32_t)j) ^ (prow & 7u)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aux_u + pslot), "f"(m.x), "f"(m.y), "f"(m.z), "f"(m.w) : "memory");
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the TMA unit
  if (threadIdx.x == 64) DBG_STAMP(9);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 64) {
    const bool add = P.tma_epi == 2;
    for (int c = 0; c < nch; ++c) {
      if (conv) tma_store_4d(tmC, stg_u + c * CHUNK_BYTES, n0 + c * 32, ct0, cf0, cb, add);
      else tma_store_2d(tmC, stg_u + c * CHUNK_BYTES, n0 + c * 32, m0, add);
      if (pool) tma_store_4d(tmX, aux_u + c * 4096, n0 + c * 32, ct0 >> 1, cf0 >> 1, cb, false);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging boxes must outlive the reads
    DBG_STAMP(6);
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool SPLIT3>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                             const CUtensorMap& tmX, const TcParams& P) {
  constexpr int B_STAGE_BYTES = BN * BK * 4;
  constexpr int HI_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int STAGE_BYTES = HI_BYTES * (SPLIT3 ? 2 : 1);
  // stage layout: TF32 [A | B];  3xTF32 [A_hi | A_lo | B_hi | B_lo] -- B_hi and B_lo are adjacent so that ONE MMA with
  // N = 2*BN multiplies A_hi by both (accumulator columns [0,BN) and [BN,2BN) are summed in the epilogue)
  constexpr int B_OFF = SPLIT3 ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  constexpr int TMEM_COLS = SPLIT3 ? 2 * BN : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* ready = empty + STAGES;       // SPLIT3: splitter warps -> MMA issuer
  uint64_t* tmem_full = ready + STAGES;
  uint64_t* aux_full = tmem_full + 1;     // TMA epilogue: the ReLU-mask tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // BN floats (TMA epilogue)

  const GemmArgs& g = P.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * P.kb_per_split;
  const int kb1 = min(P.kb_total, kb0 + P.kb_per_split);
  // cluster split-K (P.cluster_k > 1): the CTAs of one cluster (grid.z) own disjoint K slabs of ONE output tile;
  // partial accumulators are pushed over DSMEM to the CTA that owns the row and summed there in a fixed order.

  // CONV_FWD: this CTA's 128 output pixels are the box (b, f0.., t0..)
  int cb = 0, cf0 = 0, ct0 = 0;
  if (P.conv_mode == CONV_FWD) {
    const int tt = blockIdx.x % P.tiles_t, rem = blockIdx.x / P.tiles_t;
    ct0 = tt << P.bt_log2;
    cf0 = (rem % P.tiles_f) * (BM >> P.bt_log2);
    cb = rem / P.tiles_f;
  }

  DBG_SPAN(0);
  if (threadIdx.x == 0) DBG_STAMP(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 4); }
    mbar_init(tmem_full, 1);
    mbar_init(aux_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                          // prologue above overlapped the previous kernel's tail; global memory from here on
  if (threadIdx.x == 0) DBG_STAMP(1);
  if (P.tma_epi && warp >= 2) {
    // bias slice of this tile -> shared memory now, so the epilogue never waits on a global load
    const int t = threadIdx.x - 64, cn = blockIdx.y * BN + t;
    // K slabs (split_k > 1) accumulate into a zeroed / running C: slab 0 alone adds the bias
    if (t < BN) bias_s[t] = (P.g.bias && (P.g.split_k <= 1 || blockIdx.z == 0) && cn < P.g.N) ? __ldg(P.g.bias + cn) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int conv_mode = P.conv_mode, cCin = P.cCin, btl = P.bt_log2, tiles_t = P.tiles_t, tiles_f = P.tiles_f;
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (it == 0) DBG_STAMP(2);
        mbar_expect_tx(&full[s], HI_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + B_OFF;
        if (conv_mode == CONV_NONE) {
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
        } else if (conv_mode == CONV_FWD) {
          const int cpb = cCin >> 5;                         // 32-channel chunks per tap
          const int tap = kb / cpb, c0 = (kb - tap * cpb) << 5;
          const int kh = tap / 3, kw = tap - kh * 3;
          tma_load_4d(sa, &tmA, &full[s], c0, ct0 + kw - 1, cf0 + kh - 1, cb);
          tma_load_2d(sb, &tmB, &full[s], kb * BK, n0);
          DBG_KB(it, 0);
        } else {                                             // CONV_WGRAD: k-block = 32-pixel box
          const int tt = kb % tiles_t, rem = kb / tiles_t;
          const int t0 = tt << btl, f0 = (rem % tiles_f) * (32 >> btl), b = rem / tiles_f;
#pragma unroll
          for (int i = 0; i < BM / 32; ++i) {
            const int mg = m0 + i * 32;
            int tap = mg / cCin, c0 = mg - tap * cCin;
            if (tap > 8) { tap = 0; c0 = cCin; }             // rows past 9*Cin: fully out of bounds -> zeros
            const int kh = tap / 3, kw = tap - kh * 3;
            tma_load_4d(sa + i * 4096, &tmA, &full[s], c0, t0 + kw - 1, f0 + kh - 1, b);
          }
#pragma unroll
          for (int i = 0; i < BN / 32; ++i) tma_load_4d(sb + i * 4096, &tmB, &full[s], n0 + i * 32, t0, f0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * BN) >> 3) << 17);   // same, N = 2*BN
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(SPLIT3 ? &ready[s] : &full[s], ph);
        if (it == 0) DBG_STAMP(3);
        DBG_KB(it, 2);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + B_OFF;
        // descriptors of this stage; a K=8 step advances the 16-byte-granular start-address field (bits [0,14)) by
        // 32 B (K-major) or 1024 B (MN-major) -- no carry out of the field below 256 KB of shared memory
        constexpr uint64_t STEP_A = A_MN ? 64 : 2, STEP_B = B_MN ? 64 : 2;
        const uint64_t da0 = A_MN ? desc_mnmajor(sa) : desc_kmajor(sa);
        const uint64_t db0 = B_MN ? desc_mnmajor(sb) : desc_kmajor(sb);
        const uint64_t la0 = A_MN ? desc_mnmajor(sa + A_STAGE_BYTES) : desc_kmajor(sa + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t da = da0 + k * STEP_A, db = db0 + k * STEP_B;
          const uint32_t acc0 = (it > 0 || k > 0) ? 1u : 0u;
          if (SPLIT3) {
            // a_hi * [b_hi ; b_lo] over all 2*BN accumulator columns (also initialises them), then a_lo * b_hi into the
            // first BN: the three 3xTF32 products in two instructions, A_hi read from shared memory once
            tc_mma_tf32(tmem_base, da, db, idesc2, acc0);
            tc_mma_tf32(tmem_base, la0 + k * STEP_A, db, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_base, da, db, idesc, acc0);
          }
        }
        tc_commit(&empty[s]);          // frees the smem slot once these MMAs have read it
        DBG_KB(it, 3);
      }
      tc_commit(tmem_full);            // accumulator complete
      // main loop issued: the next kernel may start its prologue.  Not earlier -- a waiting grid holds shared memory
      // and TMEM, so at most one per stream, and only once this grid owns its own TMEM.
      pdl_trigger();
      DBG_STAMP(4);
    }
  } else {
    if (SPLIT3) {
      // ===================== operand splitter: a -> (hi in place, lo beside it) =====================
      const int tid = threadIdx.x - 64;                      // 0..127
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        if (threadIdx.x == 64) DBG_KB(it, 1);
        // One warp per scheduler: a load -> convert -> store chain per element would expose the shared-memory latency
        // once per element.  Issue every load of an operand first, then convert and store (hi in place, lo beside it).
        float4* a_hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES) + tid;
        float4* b_hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + B_OFF) + tid;
        constexpr int PER_A = A_STAGE_BYTES / 16 / 128, PER_B = B_STAGE_BYTES / 16 / 128;
        static_assert(A_STAGE_BYTES % (16 * 128) == 0 && B_STAGE_BYTES % (16 * 128) == 0, "whole float4 columns per thread");
        float4 a[PER_A], b[PER_B];
#pragma unroll
        for (int j = 0; j < PER_A; ++j) a[j] = a_hi[j * 128];
#pragma unroll
        for (int j = 0; j < PER_B; ++j) b[j] = b_hi[j * 128];
#pragma unroll
        for (int j = 0; j < PER_A; ++j) {
          float4 h, l;
          if (P.split_trunc) {
            l.x = tf32_lo_trunc(a[j].x); l.y = tf32_lo_trunc(a[j].y); l.z = tf32_lo_trunc(a[j].z); l.w = tf32_lo_trunc(a[j].w);
          } else {
            h.x = tf32_rna(a[j].x); l.x = tf32_rna(a[j].x - h.x);
            h.y = tf32_rna(a[j].y); l.y = tf32_rna(a[j].y - h.y);
            h.z = tf32_rna(a[j].z); l.z = tf32_rna(a[j].z - h.z);
            h.w = tf32_rna(a[j].w); l.w = tf32_rna(a[j].w - h.w);
            a_hi[j * 128] = h;
          }
          a_hi[j * 128 + A_STAGE_BYTES / 16] = l;
        }
#pragma unroll
        for (int j = 0; j < PER_B; ++j) {
          float4 h, l;
          if (P.split_trunc) {
            l.x = tf32_lo_trunc(b[j].x); l.y = tf32_lo_trunc(b[j].y); l.z = tf32_lo_trunc(b[j].z); l.w = tf32_lo_trunc(b[j].w);
          } else {
            h.x = tf32_rna(b[j].x); l.x = tf32_rna(b[j].x - h.x);
            h.y = tf32_rna(b[j].y); l.y = tf32_rna(b[j].y - h.y);
            h.z = tf32_rna(b[j].z); l.z = tf32_rna(b[j].z - h.z);
            h.w = tf32_rna(b[j].w); l.w = tf32_rna(b[j].w - h.w);
            b_hi[j * 128] = h;
          }
          b_hi[j * 128 + B_STAGE_BYTES / 16] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
    }
  }

  // ===================== epilogue =====================
  // Phase 1: every epilogue thread owns one accumulator row (TMEM lane) and parks it in a shared-memory staging
  //          tile (row-major, padded).
  // Phase 2: lanes run along COLUMNS (float4 each), so every warp instruction touches whole contiguous rows:
  //          apply alpha / bias / ReLU / ReLU-mask / beta*C and write coalesced rows (or vectorised atomics).
  //          With cluster split-K each CTA finishes rows [rank*128/CS, ...) of the tile and first sums the CS
  //          partial tiles, pulling its peers' rows over distributed shared memory in a fixed order (deterministic).
  constexpr int LDS = BN + 4;                       // staging row stride (floats): conflict-free float4 rows
  float* stg = reinterpret_cast<float*>(smem);      // aliases the operand stages (dead once the accumulator is complete)
  const int CS = P.cluster_k;
  const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int rows_per = BM / CS;
  const int q = warp & 3;                           // TMEM lane quarter this warp may access
  if (P.tma_epi) {
    // ---- TMA epilogue (no cluster split); the staging tile aliases the (dead) operand stages
    static_assert((BN / 32) * BM * 128 <= STAGES * STAGE_BYTES, "staging tile must fit in the operand stages");
    if (warp >= 2)
      tma_epilogue_rows<BN, SPLIT3>(P, &tmC, &tmX, base, base + (uint32_t)P.aux_off, tmem_full, aux_full, bias_s,
                                    tmem_base, warp, lane, m0, n0, P.conv_mode == CONV_FWD, ct0, cf0, cb);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) DBG_STAMP(7);
    if (warp == 1) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    DBG_SPAN(1);
    return;
  }
  if (warp >= 2) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");   // see tma_epilogue_rows: staging aliases the splitter's operand stages
    if (threadIdx.x == 64) DBG_STAMP(5);
    float* dst = stg + (q * 32 + lane) * LDS;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (SPLIT3) {
        uint32_t w[32];
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c * 32), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c * 32 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    if (threadIdx.x == 64) DBG_STAMP(9);
  }
  if (CS > 1) cluster_sync_all();                   // every CTA's partial tile is staged
  else if (warp >= 2) asm volatile("bar.sync 1, 128;" ::: "memory");
  if (threadIdx.x == 64) DBG_STAMP(10);
  if (warp >= 2) {
    // Everything the row loop needs is pinned in registers (keep_reg): the compiler otherwise re-reads kernel
    // parameters from the constant bank inside the loop, a dependent ~70-cycle load per field that nothing hides
    // with one warp per scheduler (measured: 640 cycles per 2-row iteration, 11k cycles per conv tile).
    constexpr int LPR = BN / 4;                     // lanes per output row (float4 each)
    constexpr int RPI = 32 / LPR;                   // rows per warp iteration
    constexpr int STEP = 4 * RPI;                   // rows per iteration of the four epilogue warps
    constexpr int UNR = 4;                          // rows in flight per thread
    const int col_t = (lane % LPR) * 4, col = n0 + col_t;
    const int N = keep_reg(g.N), M = keep_reg(g.M), ldc = keep_reg(g.ldc);
    const float alpha = keep_reg(g.alpha), beta = keep_reg(g.beta);
    float* const Cp = keep_reg(g.C);
    const float* const auxp = keep_reg(g.aux);
    const bool atomic = keep_reg((int)(g.split_k > 1)) != 0;
    const bool mask = keep_reg((int)(g.epi == EPI_RELU_BWD)) != 0;
    const bool has_beta = keep_reg((int)(beta != 0.f)) != 0;
    const bool vec = P.vecC && col + 3 < N;
    const bool conv = keep_reg((int)(P.conv_mode == CONV_FWD)) != 0;
    const int btl = keep_reg(P.bt_log2), btm = (1 << btl) - 1, cF = keep_reg(P.cF), cT = keep_reg(P.cT);
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.bias && !atomic) {
      if (col + 0 < N) bias4.x = g.bias[col + 0];
      if (col + 1 < N) bias4.y = g.bias[col + 1];
      if (col + 2 < N) bias4.z = g.bias[col + 2];
      if (col + 3 < N) bias4.w = g.bias[col + 3];
    }
    const float lo = g.epi == EPI_RELU ? 0.f : -INFINITY;   // branch-free ReLU
    if (lane == 0) DBG_STAMP(11 + (warp - 2));      // per-warp: bias loaded, row loop starts
    const int lr0 = q * RPI + lane / LPR;
    const int r_first = rank * rows_per;            // first tile row this CTA finishes
    uint32_t peer[8];                               // staging address of (row r_first + lr0, col_t) in every peer CTA
    {
      const uint32_t a0 = smem_u32(stg + (r_first + lr0) * LDS + col_t);
#pragma unroll
      for (int s2 = 0; s2 < 8; ++s2) peer[s2] = (CS > 1 && s2 < CS) ? mapa_u32(a0, (uint32_t)s2) : a0;
    }
    // element offset of tile row r_t in C / aux, or -1 when the row is outside the problem
    auto row_off = [&](int r_t) -> long long {
      if (conv) {
        const int f = cf0 + (r_t >> btl), t = ct0 + (r_t & btm);
        return (f < cF && t < cT) ? ((long long)(cb * cF + f) * cT + t) * ldc + col : -1;
      }
      return (m0 + r_t < M) ? (long long)(m0 + r_t) * ldc + col : -1;
    };
    if (col < N) {
      if (!conv && vec && !atomic && !mask && !has_beta) {
        // hot path (every nn.Linear forward / dgrad): valid rows are a prefix of the tile
        const int lim = min(rows_per, M - m0 - r_first);
        float* cp = Cp + (long long)(m0 + r_first + lr0) * ldc + col;
        const long long cstep = (long long)STEP * ldc;
        uint32_t off = 0;
#pragma unroll 4
        for (int lr = lr0; lr < lim; lr += STEP, cp += cstep, off += STEP * LDS * 4) {
          float4 acc = CS > 1 ? ld_cluster_v4(peer[0] + off)
                              : *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(stg + lr0 * LDS + col_t) + off);
          if (CS > 1) {
            float4 t[7];
#pragma unroll
            for (int s2 = 1; s2 < 8; ++s2) if (s2 < CS) t[s2 - 1] = ld_cluster_v4(peer[s2] + off);
#pragma unroll
            for (int s2 = 1; s2 < 8; ++s2)
              if (s2 < CS) { acc.x += t[s2 - 1].x; acc.y += t[s2 - 1].y; acc.z += t[s2 - 1].z; acc.w += t[s2 - 1].w; }
          }
          float4 o;
          o.x = fmaxf(fmaf(alpha, acc.x, bias4.x), lo); o.y = fmaxf(fmaf(alpha, acc.y, bias4.y), lo);
          o.z = fmaxf(fmaf(alpha, acc.z, bias4.z), lo); o.w = fmaxf(fmaf(alpha, acc.w, bias4.w), lo);
          *reinterpret_cast<float4*>(cp) = o;
        }
      } else if (vec && CS == 1) {
        // every other 16-byte-aligned case without a cluster (implicit conv rows, ReLU-backward mask, beta*C,
        // split-K atomics): UNR rows per thread are in flight at once -- all loads first, then arithmetic and stores
#pragma unroll 1
        for (int lr = lr0; lr < BM; lr += STEP * UNR) {
          long long o[UNR];
          float4 acc[UNR], ax[UNR], cc[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int l2 = lr + u * STEP;
            o[u] = l2 < BM ? row_off(l2) : -1;
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (o[u] < 0) continue;
            if (mask) ax[u] = *reinterpret_cast<const float4*>(auxp + o[u]);
            if (has_beta) cc[u] = *reinterpret_cast<const float4*>(Cp + o[u]);
            acc[u] = *reinterpret_cast<const float4*>(stg + (lr + u * STEP) * LDS + col_t);
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (o[u] < 0) continue;
            float4 v;
            v.x = fmaxf(fmaf(alpha, acc[u].x, bias4.x), lo); v.y = fmaxf(fmaf(alpha, acc[u].y, bias4.y), lo);
            v.z = fmaxf(fmaf(alpha, acc[u].z, bias4.z), lo); v.w = fmaxf(fmaf(alpha, acc[u].w, bias4.w), lo);
            float* crow = Cp + o[u];
            if (atomic) { red_add_v4(crow, v.x, v.y, v.z, v.w); continue; }   // split-K slabs without a cluster
            if (mask) {
              v.x = ax[u].x > 0.f ? v.x : 0.f; v.y = ax[u].y > 0.f ? v.y : 0.f;
              v.z = ax[u].z > 0.f ? v.z : 0.f; v.w = ax[u].w > 0.f ? v.w : 0.f;
            }
            if (has_beta) {
              v.x = fmaf(beta, cc[u].x, v.x); v.y = fmaf(beta, cc[u].y, v.y);
              v.z = fmaf(beta, cc[u].z, v.z); v.w = fmaf(beta, cc[u].w, v.w);
            }
            if (P.epi_test != 1) *reinterpret_cast<float4*>(crow) = v;
          }
        }
      } else if (vec) {
        // the same with a cluster split-K: the CS partial tiles are summed over distributed shared memory first
#pragma unroll 1
        for (int lr = lr0; lr < rows_per; lr += STEP) {
          const long long ro = row_off(r_first + lr);
          if (ro < 0) continue;
          float4 ax = make_float4(1.f, 1.f, 1.f, 1.f), cc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mask) ax = *reinterpret_cast<const float4*>(auxp + ro);
          if (has_beta) cc = *reinterpret_cast<const float4*>(Cp + ro);
          const uint32_t off = (uint32_t)((lr - lr0) * LDS * 4);
          float4 t[8];
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2) if (s2 < CS) t[s2] = ld_cluster_v4(peer[s2] + off);
          float4 acc = t[0];
#pragma unroll
          for (int s2 = 1; s2 < 8; ++s2)
            if (s2 < CS) { acc.x += t[s2].x; acc.y += t[s2].y; acc.z += t[s2].z; acc.w += t[s2].w; }
          float4 v;
          v.x = fmaxf(fmaf(alpha, acc.x, bias4.x), lo); v.y = fmaxf(fmaf(alpha, acc.y, bias4.y), lo);
          v.z = fmaxf(fmaf(alpha, acc.z, bias4.z), lo); v.w = fmaxf(fmaf(alpha, acc.w, bias4.w), lo);
          if (mask) {
            v.x = ax.x > 0.f ? v.x : 0.f; v.y = ax.y > 0.f ? v.y : 0.f;
            v.z = ax.z > 0.f ? v.z : 0.f; v.w = ax.w > 0.f ? v.w : 0.f;
          }
          v.x = fmaf(beta, cc.x, v.x); v.y = fmaf(beta, cc.y, v.y);
          v.z = fmaf(beta, cc.z, v.z); v.w = fmaf(beta, cc.w, v.w);
          *reinterpret_cast<float4*>(Cp + ro) = v;
        }
      } else {
        // ragged / unaligned columns: scalar tail (rare: N % 4 != 0 or an unaligned C)
#pragma unroll 1
        for (int lr = lr0; lr < rows_per; lr += STEP) {
          const int r_t = r_first + lr;
          const long long ro = row_off(r_t);
          if (ro < 0) continue;
          const uint32_t off = (uint32_t)((lr - lr0) * LDS * 4);
          float4 acc = CS > 1 ? ld_cluster_v4(peer[0] + off) : *reinterpret_cast<const float4*>(stg + r_t * LDS + col_t);
#pragma unroll
          for (int s2 = 1; s2 < 8; ++s2)
            if (s2 < CS) { const float4 t = ld_cluster_v4(peer[s2] + off); acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
          const float v[4] = {fmaxf(fmaf(alpha, acc.x, bias4.x), lo), fmaxf(fmaf(alpha, acc.y, bias4.y), lo),
                              fmaxf(fmaf(alpha, acc.z, bias4.z), lo), fmaxf(fmaf(alpha, acc.w, bias4.w), lo)};
          float* crow = Cp + ro;
          const float* arow = mask ? auxp + ro : nullptr;
          for (int j = 0; j < 4; ++j) {
            if (col + j >= N) continue;
            if (atomic) { atomicAdd(crow + j, v[j]); continue; }
            float t = v[j];
            if (arow) t = arow[j] > 0.f ? t : 0.f;
            if (has_beta) t += beta * crow[j];
            crow[j] = t;
          }
        }
      }
    }
    if (threadIdx.x == 64) DBG_STAMP(6);
    if (lane == 0) DBG_STAMP(15 + (warp - 2));      // per-warp: row loop done
  }
  if (CS > 1) cluster_sync_all();                   // nobody leaves while a peer may still read its staged rows
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG_STAMP(7);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  DBG_SPAN(1);
}

// Two names for one body, so that a kernel timeline tells the nn.Linear GEMMs from the implicit convolutions that share it
// (bench.py attributes kernel time to the attention / FFN family by name).
template <int BN, int STAGES, bool A_MN, bool B_MN, bool SPLIT3>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB,
                                                      const __grid_constant__ CUtensorMap tmC,
                                                      const __grid_constant__ CUtensorMap tmX, const TcParams P) {
  gemm_tc_body<BN, STAGES, A_MN, B_MN, SPLIT3>(tmA, tmB, tmC, tmX, P);
}
template <int BN, int STAGES, bool A_MN, bool B_MN, bool SPLIT3>
__global__ void __launch_bounds__(192) conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmC,
                                                           const __grid_constant__ CUtensorMap tmX, const TcParams P) {
  gemm_tc_body<BN, STAGES, A_MN, B_MN, SPLIT3>(tmA, tmB, tmC, tmX, P);
}

// ----------------------------------------------------------------------------- kw-box 3x3 convolution
// Second-generation implicit GEMM for the forward / dgrad convolutions.  The tap-box kernel above fetches (and, in
// 3xTF32, re-splits) every input pixel nine times, once per tap.  Here a tile is a fixed 8 (t) x 16 (f) pixel box and
// an A stage is the 8 x 18 box of ONE horizontal tap offset kw and one 32-channel chunk (144 rows of 128 B, rows in
// (f, t) order): its three vertical taps kh = 0, 1, 2 are the row windows [8*kh, 8*kh + 128) of the same box, i.e.
// UMMA descriptors that start kh * 1024 B further -- whole swizzle atoms, so no unaligned descriptor is needed.
// Activations are therefore fetched and split 3x instead of 9x.  Weights arrive PRE-SPLIT (hi / lo matrices written
// once per pass by the layout kernel) through their own ring, one (tap, chunk) 32-wide k-block per slot, so the
// splitter warps only touch activations.
//   warp 0: TMA producer (A ring + B ring); warp 1: TMEM allocator + MMA issuer; warps 2..5: A splitter (3xTF32),
//   then the TMA epilogue.
constexpr int KW_BT = 8, KW_BF = 16;                       // pixel box of one 128-row accumulator tile
#ifndef KW_TF32_NA64
#define KW_TF32_NA64 2
#endif
#ifndef KW_TF32_NA128
#define KW_TF32_NA128 2
#endif

// MT = accumulator tiles per CTA.  The forward / dgrad convolutions are L2 -> SM bandwidth bound (ncu, conv.2 forward:
// 464 MB through the L2 for 34 MB of DRAM reads, 6.8 TB/s): every CTA streams the whole weight matrix (hi + lo) for its 128
// pixels.  With MT = 2 the pixel box is 8 (t) x 32 (f): ONE A box of 8 x 34 rows per (chunk, kw) whose row windows
// [8 kh, 8 kh + 128) and [128 + 8 kh, 256 + 8 kh) feed two accumulators from the same weight k-block, so the weights are
// fetched once per 256 pixels (and the f halo once per 32 rows instead of once per 16).
template <int BN, bool SPLIT3, int MT>
struct KwCfg {
  static constexpr int A_ROWS = (KW_BF * MT + 2) * KW_BT;                  // 144 / 272 rows per A box
  static constexpr int A_BYTES = A_ROWS * 128;                             // 18 / 34 KB
  static constexpr int B_BYTES = BN * 128;                                 // one (tap, chunk) weight k-block
  static constexpr int A_SLOT = A_BYTES * (SPLIT3 ? 2 : 1);                // [A_hi | A_lo]
  static constexpr int B_SLOT = B_BYTES * (SPLIT3 ? 2 : 1);                // [B_hi | B_lo]
  // weight slots: 3xTF32 Cout<=64 keeps the one-tile CTA at ~106 KB so that two CTAs share an SM (one tile's epilogue
  // overlaps the other's main loop)
  static constexpr int NB = BN == 64 ? (SPLIT3 ? 2 : 4) : (SPLIT3 && MT == 2 ? 2 : 3);
  // A slots.  3xTF32: two (the [hi | lo] slots are 36 / 68 KB each).  Single-pass TF32: two as well -- four were measured
  // for Cout = 128 (conv.7 dgrad 53.6 -> 52.0 us: not latency bound); KW_TF32_NA64 / KW_TF32_NA128 for A/B builds.
  static constexpr int NA = SPLIT3 ? 2 : (BN == 64 ? KW_TF32_NA64 : KW_TF32_NA128);
  static constexpr int RING = NA * A_SLOT + NB * B_SLOT;
  static constexpr int TILE = BM * BN * 4;                                 // staged output tile (aliases the rings)
  static constexpr int HEAD = 1024;                                        // barriers + bias slice, in front of the data
  static constexpr int DATA = RING > TILE ? RING : TILE;
  static constexpr int DATA_MASK = RING > 2 * TILE ? RING : 2 * TILE;      // + the ReLU-mask tile (EPI_RELU_BWD)
  static constexpr int SMEM = HEAD + DATA + 1024 /*align slack*/;
  static constexpr int SMEM_MASK = HEAD + DATA_MASK + 1024;
  static constexpr int ACC_COLS = SPLIT3 ? 2 * BN : BN;                    // TMEM columns of one accumulator tile
  static constexpr int TMEM_COLS = MT * ACC_COLS;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation: power of two");
  static_assert(A_BYTES % (16 * 128) == 0, "splitter: whole float4 columns per thread");
};

template <int BN, bool SPLIT3, int MT>
__global__ void __launch_bounds__(192) conv3x3_kw_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmBh,
                                                         const __grid_constant__ CUtensorMap tmBl,
                                                         const __grid_constant__ CUtensorMap tmC,
                                                         const __grid_constant__ CUtensorMap tmX, const TcParams P) {
  using Cfg = KwCfg<BN, SPLIT3, MT>;
  constexpr int NB = Cfg::NB;
  constexpr int KW_NA = Cfg::NA;
  constexpr int KW_A_BYTES = Cfg::A_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t head = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_head = smem_raw + (head - smem_u32(smem_raw));
  const uint32_t base = head + Cfg::HEAD;                    // rings / staging tiles (1024 B aligned)
  uint8_t* smem = smem_head + Cfg::HEAD;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_head);
  uint64_t* a_empty = a_full + KW_NA;
  uint64_t* a_ready = a_empty + KW_NA;
  uint64_t* b_full = a_ready + KW_NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tmem_full = b_empty + NB;
  uint64_t* aux_full = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem_head + 256);   // BN floats
  const uint32_t a_ring = base, b_ring = base + KW_NA * Cfg::A_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int tt = blockIdx.x % P.tiles_t, rem = blockIdx.x / P.tiles_t;
  const int ct0 = tt * KW_BT, cf0 = (rem % P.tiles_f) * (KW_BF * MT), cb = rem / P.tiles_f;
  const int cpb = P.cCin >> 5;                               // 32-channel chunks
  const int n_stage = 3 * cpb;                               // (chunk, kw) A boxes per tile

  if (threadIdx.x == 0) DBG_STAMP(0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < KW_NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], 4); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(aux_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (threadIdx.x == 0) DBG_STAMP(1);
  if (warp >= 2) {
    const int t = threadIdx.x - 64, cn = n0 + t;
    if (t < BN) bias_s[t] = (P.g.bias && cn < P.g.N) ? __ldg(P.g.bias + cn) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    // One thread feeds both rings and must never block on one while the other has a free slot (an A box is needed
    // one whole stage ahead of its weights), so it polls the two "empty" barriers instead of waiting on either.
    if (lane == 0) {
      const int n_b = 3 * n_stage;
      int ia = 0, ib = 0;
      const long long t0 = clock64();
      while (ia < n_stage || ib < n_b) {
        bool progress = false;
        if (ia < n_stage) {
          const int sa = ia % KW_NA;
          if (mbar_test(&a_empty[sa], ((ia / KW_NA) & 1) ^ 1)) {
            const int c = ia / 3, kw = ia - c * 3;
            mbar_expect_tx(&a_full[sa], KW_A_BYTES);
            tma_load_4d(a_ring + sa * Cfg::A_SLOT, &tmA, &a_full[sa], c << 5, ct0 + kw - 1, cf0 - 1, cb);
            if (ia == 0) DBG_STAMP(2);
            ++ia;
            progress = true;
          }
        }
        if (ib < n_b) {
          const int sb = ib % NB;
          if (mbar_test(&b_empty[sb], ((ib / NB) & 1) ^ 1)) {
            const int st = ib / 3, kh = ib - st * 3;
            const int c = st / 3, kw = st - c * 3;
            const int kcol = ((kh * 3 + kw) * cpb + c) << 5;        // k-block (tap, chunk) of Wg[co, tap*Cin + ci]
            const uint32_t dst = b_ring + sb * Cfg::B_SLOT;
            mbar_expect_tx(&b_full[sb], Cfg::B_SLOT);
            tma_load_2d(dst, &tmBh, &b_full[sb], kcol, n0);
            if (SPLIT3) tma_load_2d(dst + Cfg::B_BYTES, &tmBl, &b_full[sb], kcol, n0);
            ++ib;
            progress = true;
          }
        }
        if (!progress && clock64() - t0 > 4000000000LL) __trap();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * BN) >> 3) << 17);
      int ib = 0;
      for (int st = 0; st < n_stage; ++st) {
        const int sa = st % KW_NA;
        mbar_wait(SPLIT3 ? &a_ready[sa] : &a_full[sa], (st / KW_NA) & 1);
        tc_fence_after();
        const uint32_t a_u = a_ring + sa * Cfg::A_SLOT;
        for (int kh = 0; kh < 3; ++kh, ++ib) {
          const int sb = ib % NB;
          mbar_wait(&b_full[sb], (ib / NB) & 1);
          tc_fence_after();
          const uint64_t db0 = desc_kmajor(b_ring + sb * Cfg::B_SLOT);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            // accumulator tile mt = rows [128 mt + 8 kh, 128 mt + 8 kh + 128) of the box (whole 1024 B swizzle atoms)
            const uint64_t da0 = desc_kmajor(a_u + mt * (BM * 128) + kh * 1024);
            const uint64_t la0 = desc_kmajor(a_u + KW_A_BYTES + mt * (BM * 128) + kh * 1024);
            const uint32_t acc = tmem_base + (uint32_t)(mt * Cfg::ACC_COLS);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc0 = (st > 0 || kh > 0 || k > 0) ? 1u : 0u;
              if (SPLIT3) {
                tc_mma_tf32(acc, da0 + 2 * k, db0 + 2 * k, idesc2, acc0);   // a_hi * [b_hi ; b_lo]
                tc_mma_tf32(acc, la0 + 2 * k, db0 + 2 * k, idesc, 1u);      // a_lo * b_hi
              } else {
                tc_mma_tf32(acc, da0 + 2 * k, db0 + 2 * k, idesc, acc0);
              }
            }
          }
          tc_commit(&b_empty[sb]);
        }
        tc_commit(&a_empty[sa]);
      }
      tc_commit(tmem_full);
      pdl_trigger();
      DBG_STAMP(4);
    }
  } else if (SPLIT3) {
    // ===================== activation splitter: a -> (hi in place, lo beside it) =====================
    const int tid = threadIdx.x - 64;
    for (int st = 0; st < n_stage; ++st) {
      const int sa = st % KW_NA;
      mbar_wait(&a_full[sa], (st / KW_NA) & 1);
      float4* hi = reinterpret_cast<float4*>(smem + sa * Cfg::A_SLOT) + tid;
      constexpr int PER = KW_A_BYTES / 16 / 128;               // 9 / 17 float4 per thread
      float4 a[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) a[j] = hi[j * 128];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        float4 h, l;
        if (P.split_trunc) {
          l.x = tf32_lo_trunc(a[j].x); l.y = tf32_lo_trunc(a[j].y); l.z = tf32_lo_trunc(a[j].z); l.w = tf32_lo_trunc(a[j].w);
        } else {
          h.x = tf32_rna(a[j].x); l.x = tf32_rna(a[j].x - h.x);
          h.y = tf32_rna(a[j].y); l.y = tf32_rna(a[j].y - h.y);
          h.z = tf32_rna(a[j].z); l.z = tf32_rna(a[j].z - h.z);
          h.w = tf32_rna(a[j].w); l.w = tf32_rna(a[j].w - h.w);
          hi[j * 128] = h;
        }
        hi[j * 128 + KW_A_BYTES / 16] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[sa]);
    }
  }

  // ===================== epilogue =====================
  if (warp >= 2) {
    int n_mask = 0;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      if (mt > 0) {
        if (cf0 + mt * KW_BF >= P.cF) break;                  // the second 16-row tile lies outside the image (odd tile count)
        // the staging / mask tiles are reused: their TMA reads are done (wait_group.read in the call above), order the
        // generic-proxy accesses of every thread before the next tile's TMA mask load and staging stores
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (P.pool)
        tma_epilogue_rows<BN, SPLIT3, true>(P, &tmC, &tmX, base, base + Cfg::TILE, tmem_full, aux_full, bias_s,
                                            tmem_base + (uint32_t)(mt * Cfg::ACC_COLS), warp, lane, 0, n0, true, ct0, cf0 + mt * KW_BF,
                                            cb, (uint32_t)(n_mask & 1));
      else
        tma_epilogue_rows<BN, SPLIT3, false>(P, &tmC, &tmX, base, base + Cfg::TILE, tmem_full, aux_full, bias_s,
                                             tmem_base + (uint32_t)(mt * Cfg::ACC_COLS), warp, lane, 0, n0, true, ct0, cf0 + mt * KW_BF,
                                             cb, (uint32_t)(n_mask & 1));
      ++n_mask;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG_STAMP(7);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------- kw-box 3x3 weight gradient
// dWgT[(kh, kw, ci), co] += sum_pixels x[pixel + (kh - 1, kw - 1), ci] * dy[pixel, co] for a slab of 32-pixel k-blocks
// (8 t x 4 f).  K runs over pixels (MN-major operands, SWIZZLE_128B_BASE32B): an A stage is four boxes {32 ci, 8 t,
// 4 + 2 f} = 48 K-rows each (LBO = 6144 between the 32-channel groups of the M = 128 tile), and the vertical tap kh is the
// K window that starts 8 rows = 1024 B = one K-step later -- so the x boxes are fetched (and, in 3xTF32, split) once for
// three taps instead of once per tap, and the dY boxes once for three taps instead of once per M tile.
//   Cin = 128: the four groups are the channel chunks of ONE horizontal tap kw = blockIdx.x (grid.x = 3).
//   Cin =  64: the four groups are two horizontal taps x two chunks; blockIdx.x = 0 -> taps kw 0, 1 (128 rows of dWgT per
//              kh), blockIdx.x = 1 -> tap kw 2 loaded twice (rows 64..127 of the tile are duplicates: a 64-row store box).
// Three accumulators (kh) of BN columns live in TMEM; 3xTF32 issues a_hi*b_hi + a_lo*b_hi + a_hi*b_lo into the same
// columns (no N-concatenated operand: 3 * BN columns are already taken).  Epilogue: one tap row after the other through
// the TMA reduce-add staging path.
//   warp 0: TMA producer; warp 1: TMEM allocator + MMA issuer; warps 2..5: splitter (3xTF32, truncating), then epilogue.
constexpr int WK_ROWS_A = 48, WK_A_BOX = WK_ROWS_A * 128, WK_B_BOX = 32 * 128;      // 6144 B / 4096 B per 32-column box
constexpr int WK_A_BYTES = 4 * WK_A_BOX;
constexpr int WK_HEAD = 1024;
#ifndef WK_TF32_STAGES
#define WK_TF32_STAGES 4
#endif
template <int BN, bool SPLIT3>
struct WkCfg {
  static constexpr int B_BYTES = (BN / 32) * WK_B_BOX;
  static constexpr int STAGE = (WK_A_BYTES + B_BYTES) * (SPLIT3 ? 2 : 1);            // [A | A_lo | B | B_lo]
  static constexpr int B_OFF = WK_A_BYTES * (SPLIT3 ? 2 : 1);
  // pipeline depth: 3xTF32 stages are 64-96 KB (two fit); single-pass TF32 stages are half that, and with one CTA per SM
  // (one wave) only the ring hides the ~1 us TMA round trip of a 0.3-0.6 us stage: four stages in flight
  static constexpr int STAGES = SPLIT3 ? 2 : WK_TF32_STAGES;
  static constexpr int DATA = STAGES * STAGE > BM * BN * 4 ? STAGES * STAGE : BM * BN * 4;
  static constexpr int SMEM = WK_HEAD + DATA + 1024;
  static constexpr int TMEM_COLS = BN == 128 ? 512 : 256;                            // 3 * BN rounded up to a power of two
};
template <int CIN, int BN, bool SPLIT3>
__global__ void __launch_bounds__(192) conv3x3_wgrad_kw_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmC,
                                                               const __grid_constant__ CUtensorMap tmC64, const TcParams P) {
  using Cfg = WkCfg<BN, SPLIT3>;
  constexpr int WK_STAGES = Cfg::STAGES;
  constexpr int NCH = CIN / 32;                                  // 32-channel chunks per tap: 4 or 2
  static_assert(NCH == 4 || NCH == 2, "Cin must be 64 or 128");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t head = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_head = smem_raw + (head - smem_u32(smem_raw));
  const uint32_t base = head + WK_HEAD;
  uint8_t* smem = smem_head + WK_HEAD;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_head);
  uint64_t* empty = full + WK_STAGES;
  uint64_t* ready = empty + WK_STAGES;
  uint64_t* tmem_full = ready + WK_STAGES;
  uint64_t* aux_full = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem_head + 256);    // zeros (the shared epilogue adds a bias slice)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.y * P.kb_per_split, kb1 = min(P.kb_total, kb0 + P.kb_per_split);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WK_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&ready[i], 4); }
    mbar_init(tmem_full, 1);
    mbar_init(aux_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    if (t < BN) bias_s[t] = 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s2 = it % WK_STAGES;
        mbar_wait(&empty[s2], ((it / WK_STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s2], WK_A_BYTES + Cfg::B_BYTES);
        const int tt = kb % P.tiles_t, rem = kb / P.tiles_t;
        const int t0 = tt * 8, f0 = (rem % P.tiles_f) * 4, b = rem / P.tiles_f;
        const uint32_t sa = base + s2 * Cfg::STAGE, sb = sa + Cfg::B_OFF;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          // group g of the M tile = (horizontal tap, 32-channel chunk)
          const int kw = NCH == 4 ? (int)blockIdx.x : (blockIdx.x == 0 ? g / 2 : 2);
          const int c = NCH == 4 ? g : g % 2;
          tma_load_4d(sa + g * WK_A_BOX, &tmA, &full[s2], c * 32, t0 + kw - 1, f0 - 1, b);
        }
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tma_load_4d(sb + c * WK_B_BOX, &tmB, &full[s2], c * 32, t0, f0, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s2 = it % WK_STAGES;
        mbar_wait(SPLIT3 ? &ready[s2] : &full[s2], (it / WK_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = base + s2 * Cfg::STAGE, sb = sa + Cfg::B_OFF;
        const uint64_t da0 = umma_desc(sa, WK_A_BOX, 512, 1), la0 = umma_desc(sa + WK_A_BYTES, WK_A_BOX, 512, 1);
        const uint64_t db0 = desc_mnmajor(sb), lb0 = desc_mnmajor(sb + Cfg::B_BYTES);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t d = tmem_base + (uint32_t)(kh * BN);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t da = da0 + (uint64_t)((kh + j) * 64), db = db0 + (uint64_t)(j * 64);
            tc_mma_tf32(d, da, db, idesc, (it > 0 || j > 0) ? 1u : 0u);
            if (SPLIT3) {
              tc_mma_tf32(d, la0 + (uint64_t)((kh + j) * 64), db, idesc, 1u);          // a_lo * b_hi
              tc_mma_tf32(d, da, lb0 + (uint64_t)(j * 64), idesc, 1u);                 // a_hi * b_lo
            }
          }
        }
        tc_commit(&empty[s2]);
      }
      tc_commit(tmem_full);
      pdl_trigger();
    }
  } else if (SPLIT3) {
    // ===================== splitter: lo = rna_tf32(a - trunc_tf32(a)), hi stays the raw tile =====================
    const int tid = threadIdx.x - 64;
    int it = 0;
    for (int kb = kb0; kb < kb1; ++kb, ++it) {
      const int s2 = it % WK_STAGES;
      mbar_wait(&full[s2], (it / WK_STAGES) & 1);
      float4* a4 = reinterpret_cast<float4*>(smem + s2 * Cfg::STAGE) + tid;
      float4* b4 = reinterpret_cast<float4*>(smem + s2 * Cfg::STAGE + Cfg::B_OFF) + tid;
      constexpr int PA = WK_A_BYTES / 16 / 128, PB = Cfg::B_BYTES / 16 / 128;
      float4 a[PA], bq[PB];
#pragma unroll
      for (int j = 0; j < PA; ++j) a[j] = a4[j * 128];
#pragma unroll
      for (int j = 0; j < PB; ++j) bq[j] = b4[j * 128];
#pragma unroll
      for (int j = 0; j < PA; ++j) {
        float4 l;
        l.x = tf32_lo_trunc(a[j].x); l.y = tf32_lo_trunc(a[j].y); l.z = tf32_lo_trunc(a[j].z); l.w = tf32_lo_trunc(a[j].w);
        a4[j * 128 + WK_A_BYTES / 16] = l;
      }
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        float4 l;
        l.x = tf32_lo_trunc(bq[j].x); l.y = tf32_lo_trunc(bq[j].y); l.z = tf32_lo_trunc(bq[j].z); l.w = tf32_lo_trunc(bq[j].w);
        b4[j * 128 + Cfg::B_BYTES / 16] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[s2]);
    }
  }

  // ===================== epilogue: tap row kh -> rows of dWgT, reduce-add =====================
  if (warp >= 2) {
    // Cin = 128: tile = tap (kh, kw = blockIdx.x), 128 rows.  Cin = 64: blockIdx.x = 0 -> taps (kh, 0), (kh, 1) = 128 rows
    // from (kh * 3) * 64; blockIdx.x = 1 -> tap (kh, 2), 64 rows (the tile's upper half repeats them: 64-row store box).
    const CUtensorMap* mapC = (NCH == 2 && blockIdx.x == 1) ? &tmC64 : &tmC;
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int m0 = NCH == 4 ? (kh * 3 + (int)blockIdx.x) * 128 : (kh * 3 + (blockIdx.x == 0 ? 0 : 2)) * 64;
      tma_epilogue_rows<BN, false>(P, mapC, mapC, base, base, tmem_full, aux_full, bias_s, tmem_base + (uint32_t)(kh * BN), warp,
                                   lane, m0, 0, false, 0, 0, 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");            // the staging tile is free again (its TMA reads are done)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------- fused low-rank projection pair
// y (+)= (x . W1) . W2 (+ bias) for the rank-r factorised nn.Linear pairs of the attention blocks
// (modules/common_layers.py:250-257,287-289,303: q / k / v / output = linear_b(linear_a(x))) in ONE kernel: the rank-r
// intermediate never makes a round trip through global memory between the two contractions.
//   forward  (B_MN = false): x [M, K1], W1 = A [r, K1], W2 = Bw [N2, r] (nn.Linear weights: K-major B operands), bias [N2]
//   backward (B_MN = true) : x = dy [M, K1], W1 = Bw [K1, r], W2 = A [r, N2] (the same weights read as [K, N]: MN-major)
// Phase 1 accumulates the [128, r <= 128] tile of a = x . W1 over this CTA's K slab in TMEM; the epilogue warps drain
// it, split it into tf32 hi / lo in registers and park it in shared memory in the K-major 128B-swizzled layout the UMMA
// descriptors of phase 2 read -- which is also the layout of a TMA store box, so the same tile is reduce-added to the
// global `a` (the weight-gradient contractions need it).  Phase 2 multiplies the tile by this CTA's 64 columns of W2
// (all of its <= 4 k-blocks were fetched, and split, while phase 1 ran) into a second accumulator; the result goes out
// through the TMA reduce-add staging path.  K slabs (blockIdx.z) are merged in L2 by linearity of BOTH phases:
// y = sum_s (x_s . W1_s) . W2, so y and a must be zero (zero pool) or hold the running sum (dx of a backward).
// Up to three problems that share the shapes (q | k | v of a self-attention, k | v of a cross-attention) run as one
// launch: blockIdx.z = problem * slabs + slab.
//   warp 0: TMA producer; warp 1: TMEM allocator + MMA issuer (both phases); warps 2..5: splitter, then both epilogues.
constexpr int LR_BN1 = 128;                    // rank, padded (TMA zero-fills rows / columns >= r)
constexpr int LR_BN2 = 64;                     // output columns per CTA
constexpr int LR_K2B = 4;                      // k-blocks of phase 2 (r <= 128)
constexpr int LR_MAXG = 3;
struct LrMaps { CUtensorMap x, w1, w2, y, a; };
struct LrParams {
  LrMaps tm[LR_MAXG];
  const float* bias[LR_MAXG];
  int M, r, N2;
  int kb_total, kb_per_slab, slabs;
  int k2b;                                     // ceil(r / 32)
  int store_a;
};
template <bool SPLIT3>
struct LrCfg {
  static constexpr int STAGES = 2;
  static constexpr int B1_BYTES = LR_BN1 * BK * 4;                             // 16 KB
  static constexpr int HI1 = A_STAGE_BYTES + B1_BYTES;                         // bytes TMA delivers per stage
  static constexpr int STAGE1 = HI1 * (SPLIT3 ? 2 : 1);                        // [A | A_lo | B1 | B1_lo]
  static constexpr int B1_OFF = SPLIT3 ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  static constexpr int RING = STAGES * STAGE1;                                 // 128 KB / 64 KB
  static constexpr int ATILE = LR_K2B * A_STAGE_BYTES;                         // a as the A operand of phase 2: 4 k-blocks
  static constexpr int B2_KB = LR_BN2 * BK * 4;                                // one k-block of W2 (hi)
  static constexpr int B2_SLOT = B2_KB * (SPLIT3 ? 2 : 1);                     // [hi | lo]
  static constexpr int B2_BYTES = LR_K2B * B2_SLOT;                            // 64 KB / 32 KB (later: the y staging tile)
  static constexpr int TM1 = SPLIT3 ? 2 * LR_BN1 : LR_BN1;                     // accumulator columns of phase 1
  static constexpr int TMEM_COLS = SPLIT3 ? 512 : 256;                         // TM1 + (2 x) LR_BN2, power of two
  static constexpr int SMEM = RING + B2_BYTES + 1024 /*barriers + bias*/ + 1024 /*align slack*/;
  static_assert(ATILE * (SPLIT3 ? 2 : 1) <= RING, "the a tile (hi | lo) aliases the phase-1 ring");
  static_assert((LR_BN2 / 32) * BM * 128 <= B2_BYTES, "the y staging tile aliases the W2 k-blocks");
};

template <bool B_MN, bool SPLIT3>
__global__ void __launch_bounds__(192) lowrank_pair_kernel(const __grid_constant__ LrParams P) {
  using Cfg = LrCfg<SPLIT3>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t b2_u = base + Cfg::RING;
  uint8_t* b2 = smem + Cfg::RING;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING + Cfg::B2_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* ready = empty + STAGES;
  uint64_t* b2_full = ready + STAGES;
  uint64_t* b2_ready = b2_full + 1;
  uint64_t* tmem_full1 = b2_ready + 1;
  uint64_t* a_ready = tmem_full1 + 1;
  uint64_t* tmem_full2 = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full2 + 1);
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::RING + Cfg::B2_BYTES + 256);   // LR_BN2 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.z / P.slabs, slab = blockIdx.z - prob * P.slabs;
  const LrMaps& tm = P.tm[prob];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * LR_BN2;
  const int kb0 = slab * P.kb_per_slab, kb1 = min(P.kb_total, kb0 + P.kb_per_slab);
  const int k2b = P.k2b;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 4); }
    mbar_init(b2_full, 1); mbar_init(b2_ready, 4);
    mbar_init(tmem_full1, 1); mbar_init(a_ready, 4); mbar_init(tmem_full2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem1 = *tmem_slot, tmem2 = tmem1 + (uint32_t)Cfg::TM1;
  pdl_wait();
  if (warp >= 2) {
    const int t = threadIdx.x - 64, cn = n0 + t;
    const float* bp = P.bias[prob];
    if (t < LR_BN2) bias_s[t] = (bp && slab == 0 && cn < P.N2) ? __ldg(bp + cn) : 0.f;   // slab 0 alone adds the bias
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // W2 first: it does not depend on phase 1 and is needed (split) the moment the a tile is parked
      mbar_expect_tx(b2_full, (uint32_t)(k2b * Cfg::B2_KB));
      for (int kb = 0; kb < k2b; ++kb) {
        const uint32_t dst = b2_u + kb * Cfg::B2_SLOT;
        if (!B_MN) {
          tma_load_2d(dst, &tm.w2, b2_full, kb * BK, n0);
        } else {
#pragma unroll
          for (int i = 0; i < LR_BN2 / 32; ++i) tma_load_2d(dst + i * 4096, &tm.w2, b2_full, n0 + i * 32, kb * BK);
        }
      }
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s], Cfg::HI1);
        const uint32_t sa = base + s * Cfg::STAGE1, sb = sa + Cfg::B1_OFF;
        tma_load_2d(sa, &tm.x, &full[s], kb * BK, m0);
        if (!B_MN) {
          tma_load_2d(sb, &tm.w1, &full[s], kb * BK, 0);
        } else {
#pragma unroll
          for (int i = 0; i < LR_BN1 / 32; ++i) tma_load_2d(sb + i * 4096, &tm.w1, &full[s], i * 32, kb * BK);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t IDB = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t id1 = IDB | ((uint32_t)(LR_BN1 >> 3) << 17), id1x2 = IDB | ((uint32_t)((2 * LR_BN1) >> 3) << 17);
      constexpr uint32_t id2 = IDB | ((uint32_t)(LR_BN2 >> 3) << 17), id2x2 = IDB | ((uint32_t)((2 * LR_BN2) >> 3) << 17);
      constexpr uint64_t STEP_B = B_MN ? 64 : 2;
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(SPLIT3 ? &ready[s] : &full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * Cfg::STAGE1, sb = sa + Cfg::B1_OFF;
        const uint64_t da0 = desc_kmajor(sa), la0 = desc_kmajor(sa + A_STAGE_BYTES);
        const uint64_t db0 = B_MN ? desc_mnmajor(sb) : desc_kmajor(sb);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint32_t acc0 = (it > 0 || k > 0) ? 1u : 0u;
          if (SPLIT3) {
            tc_mma_tf32(tmem1, da0 + 2 * k, db0 + k * STEP_B, id1x2, acc0);   // x_hi * [w_hi ; w_lo]
            tc_mma_tf32(tmem1, la0 + 2 * k, db0 + k * STEP_B, id1, 1u);       // x_lo * w_hi
          } else {
            tc_mma_tf32(tmem1, da0 + 2 * k, db0 + k * STEP_B, id1, acc0);
          }
        }
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full1);
      // ---- phase 2: a tile (parked over the ring by the epilogue warps) x W2 k-blocks
      mbar_wait(SPLIT3 ? b2_ready : b2_full, 0);
      mbar_wait(a_ready, 0);
      tc_fence_after();
      for (int kb = 0; kb < k2b; ++kb) {
        const uint64_t da0 = desc_kmajor(base + kb * A_STAGE_BYTES);
        const uint64_t la0 = desc_kmajor(base + Cfg::ATILE + kb * A_STAGE_BYTES);
        const uint32_t sb = b2_u + kb * Cfg::B2_SLOT;
        const uint64_t db0 = B_MN ? desc_mnmajor(sb) : desc_kmajor(sb);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
          if (SPLIT3) {
            tc_mma_tf32(tmem2, da0 + 2 * k, db0 + k * STEP_B, id2x2, acc0);
            tc_mma_tf32(tmem2, la0 + 2 * k, db0 + k * STEP_B, id2, 1u);
          } else {
            tc_mma_tf32(tmem2, da0 + 2 * k, db0 + k * STEP_B, id2, acc0);
          }
        }
      }
      tc_commit(tmem_full2);
      pdl_trigger();
    }
  } else {
    const int tid = threadIdx.x - 64;
    if (SPLIT3) {
      // ===================== operand splitter (truncating: hi stays the raw tile, lo = rna_tf32(a - trunc(a))) =====================
      int it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        float4* a_hi = reinterpret_cast<float4*>(smem + s * Cfg::STAGE1) + tid;
        float4* b_hi = reinterpret_cast<float4*>(smem + s * Cfg::STAGE1 + Cfg::B1_OFF) + tid;
        constexpr int PER_A = A_STAGE_BYTES / 16 / 128, PER_B = Cfg::B1_BYTES / 16 / 128;
        float4 a[PER_A], b[PER_B];
#pragma unroll
        for (int j = 0; j < PER_A; ++j) a[j] = a_hi[j * 128];
#pragma unroll
        for (int j = 0; j < PER_B; ++j) b[j] = b_hi[j * 128];
#pragma unroll
        for (int j = 0; j < PER_A; ++j) {
          float4 l;
          l.x = tf32_lo_trunc(a[j].x); l.y = tf32_lo_trunc(a[j].y); l.z = tf32_lo_trunc(a[j].z); l.w = tf32_lo_trunc(a[j].w);
          a_hi[j * 128 + A_STAGE_BYTES / 16] = l;
        }
#pragma unroll
        for (int j = 0; j < PER_B; ++j) {
          float4 l;
          l.x = tf32_lo_trunc(b[j].x); l.y = tf32_lo_trunc(b[j].y); l.z = tf32_lo_trunc(b[j].z); l.w = tf32_lo_trunc(b[j].w);
          b_hi[j * 128 + Cfg::B1_BYTES / 16] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
      // W2 k-blocks (landed long ago)
      mbar_wait(b2_full, 0);
      constexpr int PER_2 = Cfg::B2_KB / 16 / 128;
      for (int kb = 0; kb < k2b; ++kb) {
        float4* h = reinterpret_cast<float4*>(b2 + kb * Cfg::B2_SLOT) + tid;
        float4 v[PER_2];
#pragma unroll
        for (int j = 0; j < PER_2; ++j) v[j] = h[j * 128];
#pragma unroll
        for (int j = 0; j < PER_2; ++j) {
          float4 l;
          l.x = tf32_lo_trunc(v[j].x); l.y = tf32_lo_trunc(v[j].y); l.z = tf32_lo_trunc(v[j].z); l.w = tf32_lo_trunc(v[j].w);
          h[j * 128 + Cfg::B2_KB / 16] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(b2_ready);
    }
    // ===================== epilogue 1: a tile TMEM -> (hi | lo) K-major operand tiles over the dead ring =====================
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t row_u = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    mbar_wait(tmem_full1, 0);                                      // phase-1 MMAs complete => the ring is dead
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");                 // (racecheck-visible form of the same order)
#pragma unroll 1
    for (int c = 0; c < k2b; ++c) {
      uint32_t v[32];
      if (SPLIT3) {
        uint32_t w[32];
        tmem_ld32_issue(tmem1 + lane_addr + (uint32_t)(c * 32), v);
        tmem_ld32_issue(tmem1 + lane_addr + (uint32_t)(LR_BN1 + c * 32), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
        tmem_ld32(tmem1 + lane_addr + (uint32_t)(c * 32), v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t slot = (uint32_t)c * A_STAGE_BYTES + row_u + ((((uint32_t)j) ^ sw) << 4);
        const float x0 = __uint_as_float(v[4 * j + 0]), x1 = __uint_as_float(v[4 * j + 1]);
        const float x2 = __uint_as_float(v[4 * j + 2]), x3 = __uint_as_float(v[4 * j + 3]);
        // the hi tile is the raw fp32 sum (kind::tf32 reads its top 19 bits); it is also what `a` receives
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + slot), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
        if (SPLIT3)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)Cfg::ATILE + slot), "f"(tf32_lo_trunc(x0)),
                       "f"(tf32_lo_trunc(x1)), "f"(tf32_lo_trunc(x2)), "f"(tf32_lo_trunc(x3)) : "memory");
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // parked tiles -> visible to the MMA and TMA units
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
    if (P.store_a && blockIdx.y == 0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        for (int c = 0; c < k2b; ++c) tma_store_2d(&tm.a, base + c * A_STAGE_BYTES, c * 32, m0, true);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    // ===================== epilogue 2: y tile -> staging (over the dead W2 k-blocks) -> TMA reduce-add =====================
    const int nch = min(LR_BN2 / 32, (P.N2 - n0 + 31) / 32);
    mbar_wait(tmem_full2, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32];
      if (SPLIT3) {
        uint32_t w[32];
        tmem_ld32_issue(tmem2 + lane_addr + (uint32_t)(c * 32), v);
        tmem_ld32_issue(tmem2 + lane_addr + (uint32_t)(LR_BN2 + c * 32), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
        tmem_ld32(tmem2 + lane_addr + (uint32_t)(c * 32), v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j * 4);
        const uint32_t slot = (uint32_t)c * (BM * 128) + row_u + ((((uint32_t)j) ^ sw) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(b2_u + slot), "f"(__uint_as_float(v[4 * j + 0]) + b4.x),
                     "f"(__uint_as_float(v[4 * j + 1]) + b4.y), "f"(__uint_as_float(v[4 * j + 2]) + b4.z),
                     "f"(__uint_as_float(v[4 * j + 3]) + b4.w) : "memory");
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
      for (int c = 0; c < nch; ++c) tma_store_2d(&tm.y, b2_u + c * (BM * 128), n0 + c * 32, m0, true);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // both staged tiles (a, y) must outlive their reads
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem1), "r"(Cfg::TMEM_COLS) : "memory");
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
  const void* ptr; long long d0, d1, d2, d3, ld; int b1, b2, flags;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && ld == o.ld && b1 == o.b1 &&
           b2 == o.b2 && flags == o.flags;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ (size_t)k.d0; h = h * 1000003u ^ (size_t)k.d1; h = h * 1000003u ^ (size_t)k.d2;
    h = h * 1000003u ^ (size_t)k.d3; h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)(k.b1 * 1024 + k.b2 * 8 + k.flags);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

int encode_cached(const MapKey& key, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                  bool tf32_type, bool mn_major, CUtensorMap* out) {
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return MTL_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { mtl_set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return MTL_ERR_CUDA; }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                   (void*)key.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mtl_set_error("cuTensorMapEncodeTiled failed (%d) rank=%d ptr=%p dims=%lld,%lld,%lld,%lld ld=%lld box=%d,%d", (int)r,
                  rank, key.ptr, key.d0, key.d1, key.d2, key.d3, key.ld, key.b1, key.b2);
    return MTL_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps[key] = *out;
  return MTL_OK;
}

// 2-D map over a row-major fp32 matrix: `inner` contiguous elements, `outer` rows of stride ld; box {32, box_outer}.
int make_map(const float* ptr, long long inner, long long outer, long long ld, int box_outer, bool mn_major,
             bool tf32_type, CUtensorMap* out) {
  MapKey key{ptr, inner, outer, 0, 0, ld, box_outer, 0, (tf32_type ? 1 : 0) + (mn_major ? 2 : 0)};
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
  return encode_cached(key, 2, dims, strides, box, tf32_type, mn_major, out);
}
// 4-D map over an NHWC activation [B,F,T,C]: dims (C,T,F,B), box {32, bt, bf, 1}.
int make_map_nhwc(const float* ptr, int B, int F, int T, int C, int bt, int bf, bool mn_major, bool tf32_type,
                  CUtensorMap* out) {
  MapKey key{ptr, C, T, F, B, 0, bt, bf, 4 + (tf32_type ? 1 : 0) + (mn_major ? 2 : 0)};
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)T * C * 4, (cuuint64_t)F * T * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)bt, (cuuint32_t)bf, 1};
  return encode_cached(key, 4, dims, strides, box, tf32_type, mn_major, out);
}

struct Maps { CUtensorMap a, b, c, x; };

template <int BN, int STAGES, bool A_MN, bool B_MN, bool SPLIT3>
int launch(const Maps& tm, const TcParams& P_in, dim3 grid, cudaStream_t s) {
  constexpr int STAGE_MEM = STAGES * (A_STAGE_BYTES + BN * BK * 4) * (SPLIT3 ? 2 : 1);
  constexpr int TILE = BM * BN * 4;                       // one staged output / mask tile
  constexpr int SMEM = STAGE_MEM + 1024 /*align slack*/ + 1024 /*barriers + bias slice*/;
  constexpr bool AUX_INSIDE = 2 * TILE <= STAGE_MEM;
  constexpr int SMEM_MAX = AUX_INSIDE ? SMEM : SMEM + TILE;
  static_assert(SMEM_MAX <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, A_MN, B_MN, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    MTL_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES, A_MN, B_MN, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    configured = true;
  }
  auto kern = P_in.conv_mode != CONV_NONE ? conv_gemm_tc_kernel<BN, STAGES, A_MN, B_MN, SPLIT3> : gemm_tc_kernel<BN, STAGES, A_MN, B_MN, SPLIT3>;
  // ReLU-mask tile of the TMA epilogue: beside the staging tile inside the (dead) operand stages when they are
  // large enough, else in extra dynamic shared memory behind the barrier block
  TcParams P = P_in;
  int smem = SMEM;
  if (P.tma_epi && P.g.epi == EPI_RELU_BWD) {
    if (AUX_INSIDE) P.aux_off = TILE;
    else { P.aux_off = STAGE_MEM + 1024; smem = SMEM + TILE; }
  }
  static_assert(SMEM - 1280 >= BM * (BN + 4) * 4, "staging tile must fit in the operand stages");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(192, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[3];
  int na = 0;
  if (g_mtl_launch_prio != 0) {
    at[na].id = cudaLaunchAttributePriority;
    at[na].val.priority = g_mtl_launch_prio;
    ++na;
  }
  if (P.cluster_k > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = (unsigned)P.cluster_k;
    ++na;
  }
  if (mtl_pdl_enabled() && (g_mtl_launch_prio != 0 || !mtl_pdl_chain_only())) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  MTL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.c, tm.x, P));
  ++g_mtl_launches;
  return MTL_OK;
}

// 3xTF32 Cout=64 convolutions run with 2 pipeline stages (2 CTAs per SM: one tile's epilogue and prologue overlap the
// other's main loop); MTL_CONV_STAGES=4 restores 4 stages / 1 CTA per SM (A/B measurements)
bool conv_two_cta() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_CONV_STAGES"); v = (e && e[0] == '4') ? 0 : 1; }
  return v != 0;
}

// Cout = 64-wide conv weight-gradient tiles run with 2 pipeline stages / 2 CTAs per SM; MTL_CONV_WGRAD_2CTA=0 restores 4 / 1
bool conv_wgrad_two_cta() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_CONV_WGRAD_2CTA"); v = (e && e[0] == '0') ? 0 : 1; }   // 8.16 -> 7.99 ms/step
  return v != 0;
}

// Small-footprint pipeline for the nn.Linear GEMMs (M = 200-264 rows).  Those CTAs stream 2-16 k-blocks and spend most of
// their life waiting (first TMA, splitter, commit, epilogue); with 3 x 64 KB of stages each one owned a whole SM, and with
// three task lanes in flight they crowd each other out.  Two 48 KB stages (3xTF32, 64-wide tile: A hi|lo 32 KB + B hi|lo
// 16 KB; the staged output tile aliases them) let two CTAs share an SM -- 128 TMEM columns each -- and fit beside a
// convolution CTA.  Measured (3 lanes, ms / meta-step): 3 x 64 KB stages 7.95, 2 x 48 KB 7.70, 1 x 48 KB (four CTAs per SM,
// no TMA / MMA overlap inside a CTA) 8.49, 2 stages with 128-wide tiles 8.22.  MTL_LIN_STAGES: 0 = the 3 / 4-stage
// pipelines, 1, 2 (default).
int lin_small_stages() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_LIN_STAGES"); v = e ? atoi(e) : 2; if (v < 0 || (v > 2 && v != 4)) v = 2; }   // 4: 64-wide tiles, 4 x 48 KB
  return v;
}
int lin_small_bn() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_LIN_BN"); v = (e && atoi(e) == 128) ? 128 : 64; }
  return v;
}
template <int BN, bool SPLIT3>
int dispatch_major(bool a_mn, bool b_mn, const Maps& tm, const TcParams& P, dim3 grid, cudaStream_t s) {
  constexpr int STAGES = SPLIT3 ? (BN == 64 ? 4 : 3) : (BN == 64 ? 4 : 3);
  if (P.lin_stages > 0 && P.lin_stages <= 2 && P.conv_mode == CONV_NONE) {
    // smallest legal ring: the staged output tile (TMA or padded row-major) has to fit in the operand stages
    constexpr int S1 = SPLIT3 ? (BN == 64 ? 1 : 2) : (BN == 64 ? 2 : 3);
    constexpr int S2 = SPLIT3 ? 2 : 3;
    if (P.lin_stages == 1) {
      if (!a_mn && !b_mn) return launch<BN, S1, false, false, SPLIT3>(tm, P, grid, s);
      if (!a_mn && b_mn) return launch<BN, S1, false, true, SPLIT3>(tm, P, grid, s);
      if (a_mn && !b_mn) return launch<BN, S1, true, false, SPLIT3>(tm, P, grid, s);
      return launch<BN, S1, true, true, SPLIT3>(tm, P, grid, s);
    }
    if (!a_mn && !b_mn) return launch<BN, S2, false, false, SPLIT3>(tm, P, grid, s);
    if (!a_mn && b_mn) return launch<BN, S2, false, true, SPLIT3>(tm, P, grid, s);
    if (a_mn && !b_mn) return launch<BN, S2, true, false, SPLIT3>(tm, P, grid, s);
    return launch<BN, S2, true, true, SPLIT3>(tm, P, grid, s);
  }
  if (BN == 64 && SPLIT3 && !a_mn && !b_mn && P.conv_mode == CONV_FWD && conv_two_cta()) {
    // 2 stages x 48 KB: two CTAs share an SM, so one tile's epilogue overlaps the other's main loop
    return launch<BN, (BN == 64 && SPLIT3) ? 2 : STAGES, false, false, SPLIT3>(tm, P, grid, s);
  }
  if (BN == 64 && SPLIT3 && a_mn && b_mn && P.conv_mode == CONV_WGRAD && conv_wgrad_two_cta()) {
    // conv.2 weight gradient: 2 stages x 48 KB so that two CTAs share an SM (their TMA / split / MMA phases interleave)
    return launch<BN, (BN == 64 && SPLIT3) ? 2 : STAGES, true, true, SPLIT3>(tm, P, grid, s);
  }
  if (!a_mn && !b_mn) return launch<BN, STAGES, false, false, SPLIT3>(tm, P, grid, s);
  if (!a_mn && b_mn) return launch<BN, STAGES, false, true, SPLIT3>(tm, P, grid, s);
  if (a_mn && !b_mn) return launch<BN, STAGES, true, false, SPLIT3>(tm, P, grid, s);
  return launch<BN, STAGES, true, true, SPLIT3>(tm, P, grid, s);
}
int dispatch(int bn, bool split3, bool a_mn, bool b_mn, const Maps& tm, const TcParams& P, dim3 grid, cudaStream_t s) {
  if (bn == 64) return split3 ? dispatch_major<64, true>(a_mn, b_mn, tm, P, grid, s)
                              : dispatch_major<64, false>(a_mn, b_mn, tm, P, grid, s);
  return split3 ? dispatch_major<128, true>(a_mn, b_mn, tm, P, grid, s)
                : dispatch_major<128, false>(a_mn, b_mn, tm, P, grid, s);
}

struct KwMaps { CUtensorMap a, bh, bl, c, x; };
template <int BN, bool SPLIT3, int MT>
int launch_kw(const KwMaps& tm, const TcParams& P, dim3 grid, cudaStream_t s) {
  using Cfg = KwCfg<BN, SPLIT3, MT>;
  static_assert(Cfg::SMEM_MASK <= 227 * 1024, "shared memory budget");
  static_assert(256 + BN * 4 <= Cfg::HEAD, "bias slice must fit in the head block");
  static bool configured = false;
  auto kern = conv3x3_kw_kernel<BN, SPLIT3, MT>;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_MASK));
    configured = true;
  }
  const int smem = P.g.epi == EPI_RELU_BWD ? Cfg::SMEM_MASK : Cfg::SMEM;
  MTL_CHECK_CUDA(mtl_launch_pdl(kern, grid, dim3(192, 1, 1), (size_t)smem, s, tm.a, tm.bh, tm.bl, tm.c, tm.x, P));
  ++g_mtl_launches;
  return MTL_OK;
}

// MTL_CONV_WGRAD_KW=0 keeps the tap-box weight-gradient kernel for Cin = Cout = 128 too (A/B measurements:
// conv.4 wgrad call 85.5 -> 65.8 us in 3xTF32, 75.6 -> 44.5 us in TF32)
bool conv_wgrad_kw_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_CONV_WGRAD_KW"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// the Cin = 64 variants (conv.2, conv.3) of the kw-box weight gradient.  In 3xTF32 they were no faster than the tap-box
// kernel (three MMAs per product without the N-concatenated operand); under the default single-pass TF32 policy for the
// VGG weight gradients they are: 6.54 -> 6.41 ms / meta-step (3 lanes).  MTL_CONV_WGRAD_KW64=0 restores the tap-box kernel.
bool conv_wgrad_kw64_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_CONV_WGRAD_KW64"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
template <int CIN, int BN, bool SPLIT3>
int launch_wgrad_kw(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc64,
                    const TcParams& P, dim3 grid, cudaStream_t s) {
  using Cfg = WkCfg<BN, SPLIT3>;
  static_assert(Cfg::SMEM <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  auto kern = conv3x3_wgrad_kw_kernel<CIN, BN, SPLIT3>;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  MTL_CHECK_CUDA(mtl_launch_pdl(kern, grid, dim3(192, 1, 1), (size_t)Cfg::SMEM, s, ta, tb, tc, tc64, P));
  ++g_mtl_launches;
  return MTL_OK;
}

// MTL_TMA_EPI=0 keeps the register/shared-memory epilogue everywhere (A/B measurements)
bool tma_epi_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_TMA_EPI"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// Picks the epilogue: TMA store / reduce-add when there is no cluster split, C (and the mask) are TMA-addressable
// and beta is 0 or 1.  Fills tm.c / tm.x through `mk` (2-D matrices or NHWC pixel boxes).
template <typename MakeMap>
int plan_epilogue(TcParams& P, Maps& tm, MakeMap mk) {
  const GemmArgs& g = P.g;
  P.tma_epi = 0;
  tm.c = tm.a; tm.x = tm.a;                                        // placeholders (never dereferenced)
  const bool slab = g.split_k > 1;
  // the TMA unit clips the innermost dimension at 16-byte granularity: N % 4 != 0 would spill into padding columns
  if (!tma_epi_enabled() || P.cluster_k > 1 || !P.vecC || g.N % 4 != 0) return MTL_OK;
  if (!(g.beta == 0.f || g.beta == 1.f)) return MTL_OK;
  MTL_TRY(mk(g.C, &tm.c));
  if (g.epi == EPI_RELU_BWD) MTL_TRY(mk(g.aux, &tm.x));
  P.tma_epi = (slab || g.beta == 1.f) ? 2 : 1;
  return MTL_OK;
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
// MTL_SPLIT_TRUNC=0 restores the round-to-nearest in-place split (A/B measurements)
int split_trunc_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_SPLIT_TRUNC"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}
// MTL_CLUSTER_SPLITK=0 disables the cluster split (A/B measurements)
bool cluster_split_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTL_CLUSTER_SPLITK"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// Splits kb_total k-blocks over `want` CTAs (blockIdx.z); returns the grid depth and fills kb_per_split.
int plan_split(TcParams& P, int want) {
  int split = want > 1 ? want : 1;
  if (split > 1) {
    if (split > P.kb_total) split = P.kb_total;
    P.kb_per_split = mtl_cdiv(P.kb_total, split);
    split = mtl_cdiv(P.kb_total, P.kb_per_split);
    P.g.split_k = 2;                   // atomic epilogue (also when the split collapsed to one slab)
  } else {
    P.kb_per_split = P.kb_total;
    P.g.split_k = 1;
  }
  return split;
}

// pixel-box width (log2) minimising padded work for a rows-pixel box over an F x T plane
int pick_bt_log2(int rows, int F, int T, int* tiles_t, int* tiles_f) {
  int best = -1;
  long long best_cost = 0;
  for (int l = 3; (1 << l) <= rows; ++l) {            // bt >= 8: TMA inner extents stay >= 16 B anyway (channels are inner)
    const int bt = 1 << l, bf = rows / bt;
    if (bf > 256 || bt > 256) continue;
    const long long cost = (long long)mtl_cdiv(T, bt) * bt * mtl_cdiv(F, bf) * bf;
    if (best < 0 || cost < best_cost) { best = l; best_cost = cost; }
  }
  *tiles_t = mtl_cdiv(T, 1 << best);
  *tiles_f = mtl_cdiv(F, rows >> best);
  return best;
}

}  // namespace

bool k_gemm_tc_eligible(const GemmArgs& g) {
  if (g.M < 1 || g.N < 1 || g.K < 1) return false;
  if (!al16(g.A) || !al16(g.B)) return false;
  if (g.lda % 4 != 0 || g.ldb % 4 != 0) return false;
  return true;
}

int k_gemm_tc(const GemmArgs& g, int precision_mode, cudaStream_t s) {
  MTL_REQUIRE(k_gemm_tc_eligible(g), "operands not TMA-compatible (16 B aligned base, leading dims % 4 == 0)");
  const bool split3 = precision_mode == 2;
  const bool tf = !split3;
  const bool a_mn = g.transA != 0;     // A stored [K,M]: M contiguous
  const bool b_mn = g.transB == 0;     // B stored [K,N]: N contiguous
  const int lin_stages = lin_small_stages();
  const int bn = g.N <= 64 ? 64 : (lin_stages > 0 ? lin_small_bn() : 128);
  Maps tm;
  CUtensorMap& ta = tm.a;
  CUtensorMap& tb = tm.b;
  if (!a_mn) MTL_TRY(make_map(g.A, g.K, g.M, g.lda, BM, false, tf, &ta)); else MTL_TRY(make_map(g.A, g.M, g.K, g.lda, 32, true, tf, &ta));
  if (!b_mn) MTL_TRY(make_map(g.B, g.K, g.N, g.ldb, bn, false, tf, &tb)); else MTL_TRY(make_map(g.B, g.N, g.K, g.ldb, 32, true, tf, &tb));
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.split_trunc = split_trunc_enabled();
  P.g = g;
  P.conv_mode = CONV_NONE;
  P.lin_stages = lin_stages;
  P.kb_total = mtl_cdiv(g.K, BK);
  // K slabs accumulate into C: the epilogue has to be linear in the partial sum -- none, or the ReLU-backward mask (an
  // element-wise 0 / 1 factor: every slab masks its partial; TMA epilogue only)
  if (g.split_k > 1) MTL_REQUIRE(g.beta == 1.f && (g.epi == EPI_NONE || g.epi == EPI_RELU_BWD), "split-K needs beta=1 and a linear epilogue");
  int split = plan_split(P, g.split_k);
  P.cluster_k = 1;
  if (g.split_k <= 1 && cluster_split_enabled()) {
    // Few output tiles (M ~ 264 rows): one CTA per tile would stream its whole K extent through a single SM at
    // one-TMA-latency per 3 stages.  Spread K over a cluster of up to 8 CTAs (deterministic DSMEM reduction).
    const long long tiles = (long long)mtl_cdiv(g.M, BM) * mtl_cdiv(g.N, bn);
    int cs = 1;
    // K <= 128 (e.g. the rank-100 side of the low-rank projections: 4 k-blocks) is not worth a cluster: measured as a
    // graph node, 264 x 512 x 100 runs 9.2 us alone vs 10.9 us split 4 ways (the reduction costs more than it saves)
    // a cluster CTA owns a whole SM (192 KB of stages) for its ~8 us life: with several task lanes in flight SM-time is
    // the scarce resource, so clusters stay at <= 4 CTAs and <= 64 CTAs per GEMM there (measured 9.80 -> 9.48 ms/step
    // with three lanes); a pass running alone takes 8 / 160 (14.7 vs 15.9 ms/step with one lane).
    // MTL_CLUSTER_MAX / MTL_CLUSTER_CTAS override both.
    static int env_cs = -1, env_cta = -1;
    if (env_cs < 0) { const char* e = getenv("MTL_CLUSTER_MAX"); env_cs = e ? atoi(e) : 0; }
    if (env_cta < 0) { const char* e = getenv("MTL_CLUSTER_CTAS"); env_cta = e ? atoi(e) : 0; }
    const int cs_max = env_cs > 0 ? env_cs : (g_mtl_concurrency >= 2 ? 4 : 8);
    const int cta_max = env_cta > 0 ? env_cta : (g_mtl_concurrency >= 2 ? 64 : 160);
    while (P.kb_total > 4 && cs < cs_max && tiles * (cs * 2) <= cta_max && P.kb_total >= cs * 2) cs *= 2;
    while (cs > 1 && (long long)(cs - 1) * mtl_cdiv(P.kb_total, cs) >= P.kb_total) cs /= 2;   // every CTA gets >= 1 k-block
    if (cs > 1) {
      P.cluster_k = cs;
      P.kb_per_split = mtl_cdiv(P.kb_total, cs);
      split = cs;
    }
  }
  P.vecC = al16(g.C) && (g.ldc % 4 == 0) && (g.epi != EPI_RELU_BWD || al16(g.aux));
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MTL_GEMM_DBG"); dbg = e ? atoi(e) : 0; } P.dbg = dbg; }
  dim3 grid(mtl_cdiv(g.M, BM), mtl_cdiv(g.N, bn), split);
  MTL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm grid too large");
  const int M = g.M, N = g.N, ldc = g.ldc;
  MTL_TRY(plan_epilogue(P, tm, [&](const float* ptr, CUtensorMap* out) { return make_map(ptr, N, M, ldc, BM, false, false, out); }));
  MTL_REQUIRE(!(g.split_k > 1 && (g.bias || g.epi == EPI_RELU_BWD)) || P.tma_epi,
              "split-K with bias / ReLU mask needs the TMA epilogue (16 B aligned C, N % 4 == 0)");
  return dispatch(bn, split3, a_mn, b_mn, tm, P, grid, s);
}

// ----------------------------------------------------------------------------- fused low-rank pair: host side
namespace {
template <bool B_MN, bool SPLIT3>
int launch_lowrank(const LrParams& P, dim3 grid, cudaStream_t s) {
  using Cfg = LrCfg<SPLIT3>;
  static_assert(Cfg::SMEM <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  auto kern = lowrank_pair_kernel<B_MN, SPLIT3>;
  if (!configured) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  MTL_CHECK_CUDA(mtl_launch_pdl(kern, grid, dim3(192, 1, 1), (size_t)Cfg::SMEM, s, P));
  ++g_mtl_launches;
  return MTL_OK;
}
}  // namespace

bool k_lowrank_pair_eligible(const LrPairArgs& a) {
  if (a.G < 1 || a.G > LR_MAXG || a.M < 1 || a.K1 < 1 || a.N2 < 1) return false;
  if (a.r < 4 || a.r > LR_BN1 || a.r % 4 != 0 || a.K1 % 4 != 0 || a.N2 % 4 != 0) return false;
  for (int g = 0; g < a.G; ++g) {
    if (!al16(a.x[g]) || !al16(a.w1[g]) || !al16(a.w2[g]) || !al16(a.y[g]) || !al16(a.a[g])) return false;
    if (a.ldx[g] % 4 != 0 || a.ldy[g] % 4 != 0) return false;
  }
  return true;
}

int k_lowrank_pair(const LrPairArgs& a, int precision_mode, cudaStream_t s) {
  MTL_REQUIRE(k_lowrank_pair_eligible(a), "low-rank pair: rank <= 128, dims % 4 == 0, 16 B aligned operands");
  const bool split3 = precision_mode == 2;
  const bool tf = !split3;
  LrParams P;
  memset(&P, 0, sizeof(P));
  P.M = a.M; P.r = a.r; P.N2 = a.N2;
  P.kb_total = mtl_cdiv(a.K1, BK);
  P.k2b = mtl_cdiv(a.r, BK);
  P.store_a = 1;
  const int tiles = mtl_cdiv(a.M, BM) * mtl_cdiv(a.N2, LR_BN2) * a.G;
  int slabs = a.ctas > 0 ? a.ctas / tiles : 1;
  if (slabs > P.kb_total) slabs = P.kb_total;
  if (slabs < 1) slabs = 1;
  P.kb_per_slab = mtl_cdiv(P.kb_total, slabs);
  P.slabs = mtl_cdiv(P.kb_total, P.kb_per_slab);               // every slab owns at least one k-block
  for (int g = 0; g < a.G; ++g) {
    LrMaps& m = P.tm[g];
    MTL_TRY(make_map(a.x[g], a.K1, a.M, a.ldx[g], BM, false, tf, &m.x));
    if (!a.bwd) {
      MTL_TRY(make_map(a.w1[g], a.K1, a.r, a.K1, LR_BN1, false, tf, &m.w1));      // A  [r, K1]: K-major B operand
      MTL_TRY(make_map(a.w2[g], a.r, a.N2, a.r, LR_BN2, false, tf, &m.w2));       // Bw [N2, r]
    } else {
      MTL_TRY(make_map(a.w1[g], a.r, a.K1, a.r, 32, true, tf, &m.w1));            // Bw [K1, r]: MN-major B operand
      MTL_TRY(make_map(a.w2[g], a.N2, a.r, a.N2, 32, true, tf, &m.w2));           // A  [r, N2]
    }
    MTL_TRY(make_map(a.y[g], a.N2, a.M, a.ldy[g], BM, false, false, &m.y));
    MTL_TRY(make_map(a.a[g], a.r, a.M, a.r, BM, false, false, &m.a));
    P.bias[g] = a.bias[g];
  }
  dim3 grid(mtl_cdiv(a.M, BM), mtl_cdiv(a.N2, LR_BN2), a.G * P.slabs);
  MTL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "low-rank pair grid too large");
  if (split3) return a.bwd ? launch_lowrank<true, true>(P, grid, s) : launch_lowrank<false, true>(P, grid, s);
  return a.bwd ? launch_lowrank<true, false>(P, grid, s) : launch_lowrank<false, false>(P, grid, s);
}

// MTL_CONV_KW=0 keeps the first-generation tap-box kernel for the forward / dgrad convolutions (A/B measurements)
bool k_conv3x3_kw_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTL_CONV_KW");
    const char* t = getenv("MTL_TMA_EPI");
    v = ((e && e[0] == '0') || (t && t[0] == '0')) ? 0 : 1;
  }
  return v != 0;
}

int k_conv3x3_w_split(int precision_mode, int Cout) {
  return (precision_mode == 2 && k_conv3x3_kw_enabled() && Cout % 32 == 0) ? 1 : 0;
}

// y[pixel, co] = epi(sum_{tap,ci} x[pixel+tap, ci] * wg[co, tap*Cin+ci] + bias[co])   (x NHWC [B,F,T,Cin], y [B*F*T, Cout])
// w_split != 0: wg holds TWO matrices, hi = tf32(w) followed by lo = tf32(w - hi) (k_conv_w_*_layout with lo != null);
// required by the kw-box kernel in 3xTF32, ignored otherwise.
// ReLU + MaxPool2d(2, 2) from the convolution's epilogue.  MEASURED (cfg 2, 3 lanes): 2696 -> 2684 launches
// per step and 6.22 -> 6.28 ms: the 64 shuffles + pooled staging per 32-column chunk lengthen every convolution tile's
// epilogue (the part of a 3xTF32 tile that nothing overlaps), while the stand-alone pooling kernels (8 + 6 us per pass) hide
// under the other lanes.  Bit-identical to the two-kernel path (tests/test_gpu_ops.py::test_conv3x3_relu_pool_fused);
// MTL_CONV_POOL_FUSE=1 / mtl_conv3x3_relu_pool_fwd select it.
static bool conv_pool_fuse_env() {
  static int en = -1;
  if (en < 0) { const char* e = getenv("MTL_CONV_POOL_FUSE"); en = (e && e[0] == '1') ? 1 : 0; }
  return en != 0;
}
bool k_conv3x3_pool_fuse_default() { return conv_pool_fuse_env(); }
// can this convolution run on the kw-box kernel (the only one with the pooling epilogue)?
bool k_conv3x3_pool_fusable(int precision_mode, int Cout, int w_split) {
  const bool split3 = precision_mode == 2;
  return precision_mode != 0 && k_conv3x3_kw_enabled() && Cout % 32 == 0 && (!split3 || w_split);
}
int k_conv3x3_tc(const float* x, const float* wg, const float* bias, float* y, int B, int F, int T, int Cin, int Cout,
                 int epi, const float* aux, int precision_mode, int w_split, cudaStream_t s, float* pool_out) {
  MTL_REQUIRE(Cin % 32 == 0 && Cout % 4 == 0 && al16(x) && al16(wg) && al16(y), "conv3x3_tc: Cin % 32, Cout % 4, 16 B alignment");
  MTL_REQUIRE(epi != EPI_RELU_BWD || (aux && al16(aux)), "conv3x3_tc: aux");
  const bool split3 = precision_mode == 2, tf = !split3;
  const int bn = Cout <= 64 ? 64 : 128;   // (64-wide tiles / 2 CTAs per SM for Cout = 128 measured no better: 8.03 vs 7.99 ms/step)
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.split_trunc = split_trunc_enabled();
  P.conv_mode = CONV_FWD;
  P.cluster_k = 1;
  P.cF = F; P.cT = T; P.cCin = Cin;
  GemmArgs& g = P.g;
  g.A = x; g.B = wg; g.C = y; g.M = B * F * T; g.N = Cout; g.K = 9 * Cin; g.lda = Cin; g.ldb = 9 * Cin; g.ldc = Cout;
  g.transA = 0; g.transB = 1; g.alpha = 1.f; g.beta = 0.f; g.bias = bias; g.epi = epi; g.aux = aux; g.split_k = 1;
  P.kb_total = 9 * (Cin / 32);
  P.kb_per_split = P.kb_total;
  P.vecC = 1;
  { const char* e = getenv("MTL_GEMM_DBG"); P.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("MTL_EPI_TEST"); P.epi_test = e ? atoi(e) : 0; }
  MTL_REQUIRE(!(split3 && w_split) || k_conv3x3_w_split(precision_mode, Cout), "conv3x3_tc: pre-split weights need the kw-box kernel");
  if (k_conv3x3_kw_enabled() && Cout % 32 == 0 && (!split3 || w_split)) {
    // kw-box kernel: 8 (t) x 16 (f) pixel tiles, one per CTA (A box 8 x 18).  MTL_CONV_KW_MT=2: two per CTA (A box 8 x 34,
    // two accumulators fed by the same weight k-block: half the weight traffic per pixel).  MEASURED NEGATIVE RESULT (cfg 2,
    // one lane, us per launch, MT = 1 -> 2): conv.2 forward 3xTF32 62 -> 96, conv.5 / conv.7 forward 3xTF32 41 -> 69, conv.7
    // dgrad TF32 52 -> 64, conv.2 / conv.5 dgrad TF32 43 -> 41.5; step 6.28 -> 6.49 ms.  The 3xTF32 CTAs are bound by their four
    // splitter warps and the shared-memory traffic of the split, not by the L2: one 174-205 KB CTA per SM halves the
    // splitter warps per SM; at 80 x 50 pixels the 8 x 32 boxes pad 34 % (168 CTAs = two waves on 148 SMs).
    static int mt = -1;
    if (mt < 0) { const char* e = getenv("MTL_CONV_KW_MT"); mt = (e && e[0] == '2') ? 2 : 1; }
    P.bt_log2 = 3;
    P.tiles_t = mtl_cdiv(T, KW_BT);
    P.tiles_f = mtl_cdiv(F, KW_BF * mt);
    P.tma_epi = 1;
    KwMaps km;
    MTL_TRY(make_map_nhwc(x, B, F, T, Cin, KW_BT, KW_BF * mt + 2, false, tf, &km.a));
    MTL_TRY(make_map(wg, 9LL * Cin, Cout, 9LL * Cin, bn, false, tf, &km.bh));
    km.bl = km.bh;
    if (split3) MTL_TRY(make_map(wg + (size_t)Cout * 9 * Cin, 9LL * Cin, Cout, 9LL * Cin, bn, false, false, &km.bl));
    MTL_TRY(make_map_nhwc(y, B, F, T, Cout, KW_BT, KW_BF, false, false, &km.c));
    km.x = km.c;
    if (epi == EPI_RELU_BWD) MTL_TRY(make_map_nhwc(aux, B, F, T, Cout, KW_BT, KW_BF, false, false, &km.x));
    if (pool_out) {
      // 2x2 max-pooled copy [B, F/2, T/2, Cout] from the same epilogue: 4 (t') x 8 (f') pooled pixels per 128-pixel tile
      MTL_REQUIRE(epi == EPI_RELU && al16(pool_out) && F >= 2 && T >= 2, "conv3x3_tc: fused pooling follows the ReLU forward");
      MTL_TRY(make_map_nhwc(pool_out, B, F / 2, T / 2, Cout, KW_BT / 2, KW_BF / 2, false, false, &km.x));
      P.pool = 1;
    }
    dim3 grid(B * P.tiles_f * P.tiles_t, mtl_cdiv(Cout, bn), 1);
    if (mt == 1) {
      if (bn == 64) return split3 ? launch_kw<64, true, 1>(km, P, grid, s) : launch_kw<64, false, 1>(km, P, grid, s);
      return split3 ? launch_kw<128, true, 1>(km, P, grid, s) : launch_kw<128, false, 1>(km, P, grid, s);
    }
    if (bn == 64) return split3 ? launch_kw<64, true, 2>(km, P, grid, s) : launch_kw<64, false, 2>(km, P, grid, s);
    return split3 ? launch_kw<128, true, 2>(km, P, grid, s) : launch_kw<128, false, 2>(km, P, grid, s);
  }
  MTL_REQUIRE(!pool_out, "conv3x3_tc: fused pooling needs the kw-box kernel (k_conv3x3_pool_fusable)");
  P.bt_log2 = pick_bt_log2(BM, F, T, &P.tiles_t, &P.tiles_f);
  Maps tm;
  CUtensorMap& ta = tm.a;
  CUtensorMap& tb = tm.b;
  MTL_TRY(make_map_nhwc(x, B, F, T, Cin, 1 << P.bt_log2, BM >> P.bt_log2, false, tf, &ta));
  MTL_TRY(make_map(wg, 9LL * Cin, Cout, 9LL * Cin, bn, false, tf, &tb));
  dim3 grid(B * P.tiles_f * P.tiles_t, mtl_cdiv(Cout, bn), 1);
  const int bt = 1 << P.bt_log2, bf = BM >> P.bt_log2;
  MTL_TRY(plan_epilogue(P, tm, [&](const float* ptr, CUtensorMap* out) { return make_map_nhwc(ptr, B, F, T, Cout, bt, bf, false, false, out); }));
  return dispatch(bn, split3, false, false, tm, P, grid, s);
}

// dwgT[tap*Cin+ci, co] += sum_pixel x[pixel+tap, ci] * dy[pixel, co]   (vectorised atomics; dwgT must be pre-zeroed or hold a running sum)
int k_conv3x3_wgrad_tc(const float* x, const float* dy, float* dwgT, int B, int F, int T, int Cin, int Cout,
                       int precision_mode, cudaStream_t s) {
  MTL_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0 && al16(x) && al16(dy) && al16(dwgT), "conv3x3_wgrad_tc: channels % 32, alignment");
  const bool split3 = precision_mode == 2, tf = !split3;
  const int bn = Cout <= 64 ? 64 : 128;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.split_trunc = split_trunc_enabled();
  P.conv_mode = CONV_WGRAD;
  P.cluster_k = 1;
  P.cF = F; P.cT = T; P.cCin = Cin;
  P.bt_log2 = pick_bt_log2(32, F, T, &P.tiles_t, &P.tiles_f);
  const int bt = 1 << P.bt_log2, bf = 32 >> P.bt_log2;
  Maps tm;
  CUtensorMap& ta = tm.a;
  CUtensorMap& tb = tm.b;
  MTL_TRY(make_map_nhwc(x, B, F, T, Cin, bt, bf, true, tf, &ta));
  MTL_TRY(make_map_nhwc(dy, B, F, T, Cout, bt, bf, true, tf, &tb));
  GemmArgs& g = P.g;
  g.A = x; g.B = dy; g.C = dwgT; g.M = 9 * Cin; g.N = Cout; g.K = B * F * T; g.lda = Cin; g.ldb = Cout; g.ldc = Cout;
  g.transA = 1; g.transB = 0; g.alpha = 1.f; g.beta = 1.f; g.bias = nullptr; g.epi = EPI_NONE; g.aux = nullptr;
  const bool kw_shape = (Cin == 128 && Cout == 128) || (conv_wgrad_kw64_enabled() && Cin == 64 && (Cout == 64 || Cout == 128));
  if (conv_wgrad_kw_enabled() && kw_shape && (!split3 || P.split_trunc)) {
    // kw-box weight gradient: k-blocks of 8 (t) x 4 (f) pixels, one wave of CTAs
    const int gx = Cin == 128 ? 3 : 2;
    P.bt_log2 = 3;
    P.tiles_t = mtl_cdiv(T, 8); P.tiles_f = mtl_cdiv(F, 4);
    P.kb_total = B * P.tiles_f * P.tiles_t;
    int want = 148 / gx;
    if (want > P.kb_total) want = P.kb_total;
    P.kb_per_split = mtl_cdiv(P.kb_total, want);
    const int split = mtl_cdiv(P.kb_total, P.kb_per_split);
    P.g.split_k = 2;
    P.tma_epi = 2;
    P.vecC = 1;
    CUtensorMap ka, kb, kc, kc64;
    MTL_TRY(make_map_nhwc(x, B, F, T, Cin, 8, 6, true, tf, &ka));
    MTL_TRY(make_map_nhwc(dy, B, F, T, Cout, 8, 4, true, tf, &kb));
    MTL_TRY(make_map(dwgT, Cout, 9LL * Cin, Cout, BM, false, false, &kc));
    MTL_TRY(make_map(dwgT, Cout, 9LL * Cin, Cout, 64, false, false, &kc64));
    dim3 grid(gx, split, 1);
    if (Cin == 128) return split3 ? launch_wgrad_kw<128, 128, true>(ka, kb, kc, kc64, P, grid, s)
                                  : launch_wgrad_kw<128, 128, false>(ka, kb, kc, kc64, P, grid, s);
    if (Cout == 128) return split3 ? launch_wgrad_kw<64, 128, true>(ka, kb, kc, kc64, P, grid, s)
                                   : launch_wgrad_kw<64, 128, false>(ka, kb, kc, kc64, P, grid, s);
    return split3 ? launch_wgrad_kw<64, 64, true>(ka, kb, kc, kc64, P, grid, s)
                  : launch_wgrad_kw<64, 64, false>(ka, kb, kc, kc64, P, grid, s);
  }
  P.kb_total = B * P.tiles_f * P.tiles_t;
  const int mt = mtl_cdiv(g.M, BM), nt = mtl_cdiv(Cout, bn);
  // ONE full wave of CTAs at one CTA per SM (192 KB of operand stages).  History: rounding the slab count UP to ~2 waves
  // left a third, nearly empty wave behind (conv.2: 5 x 60 = 300 CTAs on 148 SMs); two exact waves: 9.91 -> 9.80 ms/step;
  // one wave (longer K per CTA, half the reduce-add epilogues): 8.67 -> 8.49 ms/step.  MTL_CONV_WGRAD_WAVES overrides.
  static int waves = -1;
  if (waves < 0) { const char* e = getenv("MTL_CONV_WGRAD_WAVES"); waves = e && atoi(e) > 0 ? atoi(e) : 1; }
  const int w_eff = (bn == 64 && split3 && conv_wgrad_two_cta()) ? 2 * waves : waves;     // two co-resident CTAs per SM
  int want = (w_eff * 148) / (mt * nt);
  if (want > P.kb_total / 8) want = P.kb_total / 8 > 0 ? P.kb_total / 8 : 1;
  const int split = plan_split(P, want > 1 ? want : 2);
  P.g.split_k = 2;
  P.vecC = 1;
  dim3 grid(mt, nt, split);
  MTL_REQUIRE(grid.z <= 65535, "conv wgrad grid too large");
  MTL_TRY(plan_epilogue(P, tm, [&](const float* ptr, CUtensorMap* out) { return make_map(ptr, Cout, 9LL * Cin, Cout, BM, false, false, out); }));
  return dispatch(bn, split3, true, true, tm, P, grid, s);
}

// debug (MTL_GEMM_DBG=99): per-CTA globaltimer entry / exit pairs of the last GEMM launch
int k_gemm_tc_debug_span(unsigned long long* host512) {
  MTL_CHECK_CUDA(cudaMemcpyFromSymbol(host512, g_cta_span, sizeof(unsigned long long) * 512));
  return MTL_OK;
}
// debug: copies the 32 cycle stamps of the last instrumented GEMM launch to the host
int k_gemm_tc_debug_stamps(long long* host32) {
  MTL_CHECK_CUDA(cudaMemcpyFromSymbol(host32, g_dbg, sizeof(long long) * 160));
  return MTL_OK;
}
