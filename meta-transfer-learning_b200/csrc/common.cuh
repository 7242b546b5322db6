// Shared device/host helpers for libmtl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MTL_OK 0
#define MTL_ERR_CUDA -1
#define MTL_ERR_ARG -2
#define MTL_ERR_WORKSPACE -3

void mtl_set_error(const char* fmt, ...);

#define MTL_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      mtl_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),      \
                    cudaGetErrorString(_e));                                                  \
      return MTL_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

extern int g_mtl_concurrency;                 // task lanes being enqueued together (1 outside mtl_meta_tasks)
bool mtl_pdl_chain_only();                    // MTL_PDL_SIDE=0: programmatic dependent launch only for chain kernels
extern int g_mtl_launch_prio;                 // CUDA launch priority of the kernels being enqueued (0 = default / lowest; < 0 = more urgent)
extern unsigned long long g_mtl_launches;   // kernels enqueued by this library (bench.py reports it)
#define MTL_CHECK_LAUNCH()               \
  do {                                   \
    ++g_mtl_launches;                    \
    MTL_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

#define MTL_REQUIRE(cond, msg)                                                \
  do {                                                                        \
    if (!(cond)) {                                                            \
      mtl_set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg);   \
      return MTL_ERR_ARG;                                                     \
    }                                                                         \
  } while (0)

#define MTL_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != MTL_OK) return _r; \
  } while (0)

// MTL_PDL=0 disables programmatic dependent launch (A/B measurements); defined in engine.cu
bool mtl_pdl_enabled();

static inline int mtl_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// Programmatic dependent launch (kernels launched through mtl_launch_pdl; both are no-ops otherwise).
// pdl_wait: everything the previous kernel in the stream wrote is visible after it -- nothing before it may touch
// global memory.  pdl_trigger: the next kernel's CTAs may become resident and run their prologue (barrier init, TMEM
// allocation, descriptor fetch, index math) while this one finishes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef MTL_PDL_LATE
#define MTL_PDL_LATE 0
#endif
// MTL_PDL_LATE (build-time A/B): 1 = no explicit trigger (dependents become resident when this grid's CTAs exit);
// 0 = trigger where the kernel says.  A dependent grid that is resident but parked in pdl_wait holds its shared memory
// and TMEM: free with one task lane, SM-time with several.
__device__ __forceinline__ void pdl_trigger() {
#if !MTL_PDL_LATE
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// <<<grid, block, smem, s>>> with the programmatic-stream-serialization attribute (the kernel MUST call pdl_wait)
template <typename... KArgs, typename... Args>
static inline cudaError_t mtl_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (mtl_pdl_enabled() && (g_mtl_launch_prio != 0 || !mtl_pdl_chain_only())) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_mtl_launch_prio != 0) {              // latency-critical chain kernel: ahead of GPU-filling / deferred work
    at[na].id = cudaLaunchAttributePriority;
    at[na].val.priority = g_mtl_launch_prio;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

// ---- Philox4x32-10 counter-based RNG (dropout masks; recomputed in backward) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Keep-decision for element `idx` of dropout site `site` (seed identifies the pass).
// One Philox block serves 4 consecutive elements.  Returns the scale to apply (0 or 1/(1-p)).
__device__ __forceinline__ float dropout_scale(unsigned long long seed, uint32_t site,
                                               unsigned long long idx, float p, float inv_keep) {
  uint4 c = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), site, 0u);
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t sel = (uint32_t)(idx & 3ull);
  uint32_t v = sel == 0 ? r.x : (sel == 1 ? r.y : (sel == 2 ? r.z : r.w));
  // uniform in [0,1): keep iff u >= p
  float u = (float)(v >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}
// Same decisions for the 4 consecutive elements idx4 .. idx4+3 (idx4 % 4 == 0) from ONE Philox block.
__device__ __forceinline__ float4 dropout_scale4(unsigned long long seed, uint32_t site,
                                                 unsigned long long idx4, float p, float inv_keep) {
  uint4 c = make_uint4((uint32_t)(idx4 >> 2), (uint32_t)(idx4 >> 34), site, 0u);
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 1.0f / 16777216.0f;
  float4 o;
  o.x = (float)(r.x >> 8) * k >= p ? inv_keep : 0.f;
  o.y = (float)(r.y >> 8) * k >= p ? inv_keep : 0.f;
  o.z = (float)(r.z >> 8) * k >= p ? inv_keep : 0.f;
  o.w = (float)(r.w >> 8) * k >= p ? inv_keep : 0.f;
  return o;
}
#endif

// Dropout descriptor passed to kernels (p == 0 disables).  The effective Philox seed is
// seed (+ *seed_dev * seed_mul when seed_dev != null): a device-resident base lets a captured CUDA
// graph draw fresh masks on every replay.
struct MtlDrop {
  float p;
  float inv_keep;
  unsigned long long seed;
  uint32_t site;
  const unsigned long long* seed_dev;
  unsigned long long seed_mul;
};
static inline MtlDrop mtl_nodrop() {
  MtlDrop d; d.p = 0.f; d.inv_keep = 1.f; d.seed = 0; d.site = 0; d.seed_dev = nullptr; d.seed_mul = 0; return d;
}
static inline MtlDrop mtl_drop(float p, unsigned long long seed, uint32_t site,
                               const unsigned long long* seed_dev = nullptr, unsigned long long seed_mul = 0) {
  MtlDrop d; d.p = p; d.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f; d.seed = seed; d.site = site;
  d.seed_dev = seed_dev; d.seed_mul = seed_mul; return d;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long mtl_eff_seed(const MtlDrop& d) {
  return d.seed_dev ? __ldg(d.seed_dev) * d.seed_mul + d.seed : d.seed;
}
#endif
