// Probe: where does the fixed ~10 us per tcgen05-GEMM launch come from?  Times back-to-back launches of stub
// kernels that add one ingredient at a time.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_floor launch_floor.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
}

template <int V>
__global__ void __launch_bounds__(192) stub(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                            float* out) {
  extern __shared__ uint8_t smem_raw[];
  if (V == 0) { if (out == nullptr) out[0] = 1.f; return; }
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* tfull = full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (V >= 2) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(tfull)), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (V >= 3 && warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  uint32_t tmem = V >= 3 ? *slot : 0;
  if (V >= 4) {
    if (warp == 0 && lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full)), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&tmA), "r"(smem_u32(full)), "r"(0), "r"(0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(smem) + 16384), "l"((uint64_t)&tmB), "r"(smem_u32(full)), "r"(0), "r"(0) : "memory");
    }
    if (V == 4) { mbar_wait(full, 0); }
  }
  if (V >= 5) {
    if (warp == 1 && lane == 0) {
      mbar_wait(full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int k = 0; k < 4; ++k) {
        uint64_t da = 0, db = 0;
        uint32_t sa = smem_u32(smem) + k * 32, sb = sa + 16384;
        da = (uint64_t)((sa & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        db = (uint64_t)((sb & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(tfull)) : "memory");
    }
    if (warp >= 2) {
      mbar_wait(tfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (V >= 6) {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                     : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[(blockIdx.x * 128 + (warp & 3) * 32 + lane) * 4] = __uint_as_float(r0) + __uint_as_float(r1) + __uint_as_float(r2) + __uint_as_float(r3);
        if (V == 7) {          // 32 coalesced 512 B row stores per warp straight from registers (64 KB per CTA)
          float4 v = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
          for (int i = 0; i < 32; ++i)
            *reinterpret_cast<float4*>(out + 4096 + ((size_t)(blockIdx.x * 128 + (warp & 3) * 32 + i) * 128 + lane * 4)) = v;
        }
        if (V == 8) {          // same bytes through a shared-memory staging tile that aliases the TMA-written operand buffer
          float* stg = reinterpret_cast<float*>(smem);
          const int r = (warp & 3) * 32 + lane;
          for (int j = 0; j < 128; j += 4)
            *reinterpret_cast<float4*>(&stg[r * 132 + j]) = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int i = (warp & 3); i < 128; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(&stg[i * 132 + lane * 4]);
            *reinterpret_cast<float4*>(out + 4096 + ((size_t)(blockIdx.x * 128 + i) * 128 + lane * 4)) = v;
          }
        }
        if (V == 9) {          // V8 with a staging tile that does NOT alias the operand buffer
          float* stg = reinterpret_cast<float*>(smem + 70 * 1024);
          const int r = (warp & 3) * 32 + lane;
          for (int j = 0; j < 128; j += 4)
            *reinterpret_cast<float4*>(&stg[r * 132 + j]) = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int i = (warp & 3); i < 128; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(&stg[i * 132 + lane * 4]);
            *reinterpret_cast<float4*>(out + 4096 + ((size_t)(blockIdx.x * 128 + i) * 128 + lane * 4)) = v;
          }
        }
      }
    }
  }
  if (V >= 2) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (V >= 3 && warp == 1) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
int run(const char* what, const CUtensorMap& ta, const CUtensorMap& tb, float* out, int smem, int grid, int reps = 2000) {
  auto k = stub<V>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 50; ++i) k<<<grid, 192, smem>>>(ta, tb, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<<<grid, 192, smem>>>(ta, tb, out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  // same through a graph of 200 dependent nodes (no host launch cost)
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < 200; ++i) k<<<grid, 192, smem, st>>>(ta, tb, out);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float msg; CK(cudaEventElapsedTime(&msg, e0, e1));
  printf("V%d %-44s smem %6d grid %3d: stream %7.2f us/launch   graph %7.2f us/node\n", V, what, smem, grid, ms * 1e3 / reps,
         msg * 1e3 / 2000);
  return 0;
}

int main() {
  float *A, *B, *out;
  CK(cudaMalloc(&A, 4096 * 512 * 4)); CK(cudaMalloc(&B, 4096 * 512 * 4)); CK(cudaMalloc(&out, 64 << 20));
  CK(cudaMemset(A, 0, 4096 * 512 * 4)); CK(cudaMemset(B, 0, 4096 * 512 * 4));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap ta, tb;
  cuuint64_t dims[2] = {512, 4096}; cuuint64_t strides[1] = {512 * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
  if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  enc(&tb, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, B, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  for (int grid : {1, 12, 148}) {
    if (run<0>("empty kernel", ta, tb, out, 0, grid)) return 1;
    if (run<0>("empty kernel + 99 KB dynamic smem", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<0>("empty kernel + 193 KB dynamic smem", ta, tb, out, 193 * 1024, grid)) return 1;
    if (run<2>("+ mbarrier init, syncthreads", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<3>("+ TMEM alloc/dealloc", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<4>("+ 2 TMA loads (32 KB) + wait", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<5>("+ 4 MMAs + commit + epilogue wait", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<6>("+ tcgen05.ld + global store", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<6>("same, 193 KB smem", ta, tb, out, 193 * 1024, grid)) return 1;
    if (run<7>("+ 64 KB of coalesced stores from registers", ta, tb, out, 99 * 1024, grid)) return 1;
    if (run<8>("+ 64 KB via smem staging (aliasing operands)", ta, tb, out, 193 * 1024, grid)) return 1;
    if (run<9>("+ 64 KB via smem staging (separate region)", ta, tb, out, 193 * 1024, grid)) return 1;
  }
  return 0;
}
