// Probe: how fast does a B200 retire CUDA-graph kernel nodes?  L independent chains of N kernels (each kernel: `ctas`
// CTAs spinning `us` microseconds), optionally with a side branch forked / joined every `fork` nodes (like the
// parameter-gradient side streams of a pass).  Prints us per node = graph time / (L * N).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o graph_node_rate graph_node_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(err_), __LINE__); exit(1); } } while (0)

__global__ void spin(long long ns, float* sink) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
  if (ns < 0) sink[0] = 1.f;
}

static float run(int L, int N, int ctas, double us, int fork, bool pdl) {
  std::vector<cudaStream_t> st(L), side(L);
  for (int l = 0; l < L; ++l) { CK(cudaStreamCreateWithFlags(&st[l], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&side[l], cudaStreamNonBlocking)); }
  cudaStream_t cap; CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> ev(4 * L + 2);
  for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
  CK(cudaEventRecord(ev[0], cap));
  auto launch = [&](cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(192); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, spin, (long long)(us * 1000), (float*)nullptr));
  };
  for (int l = 0; l < L; ++l) CK(cudaStreamWaitEvent(st[l], ev[0], 0));
  for (int i = 0; i < N; ++i)
    for (int l = 0; l < L; ++l) {
      launch(st[l]);
      if (fork > 0 && i % fork == fork - 1) {             // side branch: one kernel that depends on the chain so far
        CK(cudaEventRecord(ev[1 + 4 * l], st[l]));
        CK(cudaStreamWaitEvent(side[l], ev[1 + 4 * l], 0));
        launch(side[l]);
      }
    }
  for (int l = 0; l < L; ++l) {
    if (fork > 0) { CK(cudaEventRecord(ev[2 + 4 * l], side[l])); CK(cudaStreamWaitEvent(st[l], ev[2 + 4 * l], 0)); }
    CK(cudaEventRecord(ev[3 + 4 * l], st[l])); CK(cudaStreamWaitEvent(cap, ev[3 + 4 * l], 0));
  }
  cudaGraph_t g; CK(cudaStreamEndCapture(cap, &g));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ex, cap));
  CK(cudaStreamSynchronize(cap));
  CK(cudaEventRecord(e0, cap));
  const int reps = 10;
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ex, cap));
  CK(cudaEventRecord(e1, cap));
  CK(cudaStreamSynchronize(cap));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  size_t nn = 0; CK(cudaGraphGetNodes(g, nullptr, &nn));
  cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  printf("chains %d x %4d nodes (%3d CTAs, %4.1f us each, side branch every %d, pdl %d): %8.1f us / replay = %5.2f us per chain node, %5.2f us per graph node (%zu nodes)\n",
         L, N, ctas, us, fork, (int)pdl, 1e3 * ms / reps, 1e3 * ms / reps / N, 1e3 * ms / reps / nn, nn);
  return ms;
}

int main() {
  for (int pdl = 0; pdl < 2; ++pdl)
    for (double us : {0.0, 4.0, 8.0})
      for (int fork : {0, 3}) {
        run(1, 400, 48, us, fork, pdl);
        run(3, 400, 48, us, fork, pdl);
        run(6, 400, 48, us, fork, pdl);
      }
  return 0;
}
