// Word-level LSTM language model and its first-order meta-step (BASELINE configs[4]).
//
// Reference wiring restated by this file (relative to the reference tree):
//   lm/model/rnn_model.py:12-62     RNNModel: Embedding -> Dropout -> nn.LSTM(ninp, nhid, nlayers, dropout) -> Dropout ->
//                                   Linear(nhid, ntoken); hidden = (h, c) of shape [nlayers, B, nhid]
//   lm/main_meta_transfer.py:268-275 forward_one_batch: detached hidden in, zero_grad, forward
//   lm/main_meta_transfer.py:277-372 meta loop body: snapshot, per task [train fwd/bwd, clip, SGD(lr / meta_lr_factor),
//                                   val fwd with the train pass's hidden, weighted val loss], reset, ONE backward of the
//                                   weighted sum, clip, SGD(lr)
//   lm/util/data.py:36-67           (T = bptt, B) token blocks; targets = the block shifted by one, flattened
// The reference's `batch_loss.backward()` runs AFTER `load_state_dict(weights_original)` overwrote the adapted weights in
// place, which autograd rejects on torch >= 1.5; the meta-gradient implemented here is the first-order one SURVEY
// prescribes: d(val loss)/d(theta) evaluated at each task's adapted weights, weighted and summed.
//
// Execution: the input-side gate GEMMs ([T*B, ninp] x [ninp, 4H], all time steps at once), the decoder and every
// weight-gradient contraction go through the tensor-core GEMM (k_gemm); the recurrence is one small kernel per time step
// (gates + cell update fused forward; recurrent dgrad + gate gradients fused backward).  Nothing synchronises or allocates.
#include "kernels.h"
#include "../../include/mtl_b200.h"

namespace {

struct LmLayout {
  size_t enc;                       // encoder.weight [V, ninp]
  size_t w_ih[8], w_hh[8], b_ih[8], b_hh[8];
  size_t dec_w, dec_b;              // decoder.weight [V, H], decoder.bias [V]
  size_t off[64], numel[64];
  int n = 0;
  size_t total = 0;
  size_t add(size_t k) {
    const size_t o = (total + 63) & ~(size_t)63;
    off[n] = o; numel[n] = k; ++n;
    total = o + k;
    return o;
  }
};
int build_lm_layout(const mtl_lm_cfg& c, LmLayout& L) {
  MTL_REQUIRE(c.vocab > 1 && c.ninp > 0 && c.nhid > 0 && c.nlayers >= 1 && c.nlayers <= 8, "lm cfg");
  MTL_REQUIRE(c.ninp % 4 == 0 && c.nhid % 4 == 0, "ninp and nhid must be multiples of 4");
  const size_t H = c.nhid;
  L.enc = L.add((size_t)c.vocab * c.ninp);
  for (int l = 0; l < c.nlayers; ++l) {            // nn.LSTM registration order: weight_ih, weight_hh, bias_ih, bias_hh per layer
    const size_t in = l == 0 ? c.ninp : c.nhid;
    L.w_ih[l] = L.add(4 * H * in);
    L.w_hh[l] = L.add(4 * H * H);
    L.b_ih[l] = L.add(4 * H);
    L.b_hh[l] = L.add(4 * H);
  }
  L.dec_w = L.add((size_t)c.vocab * H);
  L.dec_b = L.add(c.vocab);
  L.total = (L.total + 63) & ~(size_t)63;
  return MTL_OK;
}

// ----------------------------------------------------------------------------- kernels
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// out[row, k] = drop(E[tok[row], k]);  rows = T*B
__global__ void lm_embed_fwd_kernel(const long long* __restrict__ tok, const float* __restrict__ E, MtlDrop drop,
                                    float* __restrict__ out, int rows, int d, int V) {
  const unsigned long long seed = drop.p > 0.f ? mtl_eff_seed(drop) : 0ull;
  const size_t total = (size_t)rows * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / d;
    long long t = tok[row];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    float v = E[(size_t)t * d + (i - row * d)];
    if (drop.p > 0.f) v *= dropout_scale(seed, drop.site, i, drop.p, drop.inv_keep);
    out[i] = v;
  }
}
// dE[tok[row], k] += drop_mask * dx[row, k]
__global__ void lm_embed_bwd_kernel(const long long* __restrict__ tok, const float* __restrict__ dx, MtlDrop drop,
                                    float* __restrict__ dE, int rows, int d, int V) {
  const unsigned long long seed = drop.p > 0.f ? mtl_eff_seed(drop) : 0ull;
  const size_t total = (size_t)rows * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / d;
    long long t = tok[row];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    float v = dx[i];
    if (drop.p > 0.f) v *= dropout_scale(seed, drop.site, i, drop.p, drop.inv_keep);
    atomicAdd(dE + (size_t)t * d + (i - row * d), v);
  }
}
// y = x * mask (forward and backward of a dropout site are the same map)
__global__ void lm_dropout_kernel(const float* __restrict__ x, MtlDrop drop, float* __restrict__ y, size_t n) {
  const unsigned long long seed = mtl_eff_seed(drop);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, drop.site, i, drop.p, drop.inv_keep);
}
// tokens (int64, any value) -> int32 gold indices for the CE kernels
__global__ void lm_gold_kernel(const long long* __restrict__ trg, int* __restrict__ gold, int n, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { long long t = trg[i]; gold[i] = (int)(t < 0 ? 0 : (t >= V ? V - 1 : t)); }
}
// W [4H, H] -> WT [H, 4H] (so that the backward step's warp reads row j of W_hh^T coalesced)
__global__ void lm_transpose_kernel(const float* __restrict__ W, float* __restrict__ WT, int rows, int cols) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[i][threadIdx.x] = W[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) WT[(size_t)c * rows + r] = t[threadIdx.x][i];
  }
}

// One LSTM time step (torch.nn.LSTM cell; gate order i, f, g, o).  One WARP per hidden unit j: its lanes stride the
// K = H reduction of the four gate rows W_hh[g*H + j, :] (coalesced), the previous hidden state of a chunk of LM_BCH batch
// rows sits in shared memory, partial sums are reduced by shuffles and lane b finishes unit (b, j):
//   pre = xg[b, g*H + j] (x W_ih^T + b_ih, precomputed) + b_hh[g*H + j] + sum_k h_prev[b, k] * W_hh[g*H + j, k]
//   c = sig(f) * c_prev + sig(i) * tanh(g);  h = sig(o) * tanh(c);  the four ACTIVATED gates are kept for the backward
// (the first version, one thread per (b, j) walking K alone, ran 40 CTAs at ~25 us per step.)
// One bulk copy (TMA, cp.async.bulk) of `bytes` (multiple of 16) global -> shared, completion on an mbarrier: the 64 KB gate-
// gradient block of a backward step arrives in ONE round trip instead of four rounds of float4 loads per thread.
__device__ __forceinline__ uint32_t lm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lm_bulk_load(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = lm_smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(lm_smem_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void lm_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = lm_smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// global -> shared copy of n floats (n % 4 == 0, both 16 B aligned) with eight float4 loads per thread in flight at once:
// the copy is one round trip to the L2 instead of one per element (a scalar loop here was most of the step kernels' time)
__device__ __forceinline__ void fill_smem(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i0 = 0; i0 < n4; i0 += 8 * (int)blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * (int)blockDim.x + (int)threadIdx.x; if (i < n4) v[u] = __ldg(s4 + i); }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * (int)blockDim.x + (int)threadIdx.x; if (i < n4) d4[i] = v[u]; }
  }
}
#ifndef LM_WARPS_PER_CTA
#define LM_WARPS_PER_CTA 4   // 8: 9.03, 4: 8.41, 2: 9.45 ms per meta-iteration (fewer shuffle reductions per SM vs. more state re-fills)
#endif
// Warp reduce-scatter by recursive halving: v[] holds N = 32 * PER per-lane partial sums; afterwards v[0 .. PER) of lane L is
// the 32-lane total of entries [L * PER, (L + 1) * PER).  N - PER shuffles instead of 5 * N for N butterfly all-reduces
// (124 instead of 480 for the 4 gates x 24 batch rows of a forward step).
template <int PER>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[32 * PER], int lane) {
#pragma unroll
  for (int s = 16, half = 16 * PER; s >= 1; s >>= 1, half >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < 16 * PER; ++i) {
      if (i < half) {
        const float send = up ? v[i] : v[i + half];
        const float keep = up ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
  }
}
constexpr int LM_BCH = 24, LM_WARPS = LM_WARPS_PER_CTA;      // batch rows per pass over the weights (the script's batch of 20 in one), warps per CTA
constexpr int LM_KPF = 8, LM_KPB = 16;        // weight values per lane fetched together (all loads in flight before the first FMA)
__global__ void __launch_bounds__(LM_WARPS * 32) lstm_step_fwd_kernel(const float* __restrict__ xg, const float* __restrict__ h_prev,
                                                                      const float* __restrict__ c_prev, const float* __restrict__ Whh,
                                                                      const float* __restrict__ b_hh, float* __restrict__ gates,
                                                                      float* __restrict__ c_out, float* __restrict__ h_out, int B,
                                                                      int H) {
  extern __shared__ float hs[];                   // h_prev[b0 .. b0 + LM_BCH, :]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * LM_WARPS + warp;
  const int H4 = 4 * H;
  for (int b0 = 0; b0 < B; b0 += LM_BCH) {
    const int nb = min(LM_BCH, B - b0);
    __syncthreads();
    fill_smem(hs, h_prev + (size_t)b0 * H, nb * H);
    __syncthreads();
    if (j >= H) continue;
    float acc[4][LM_BCH];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int b = 0; b < LM_BCH; ++b) acc[g][b] = 0.f;
    for (int k0 = 0; k0 < H; k0 += 32 * LM_KPF) {
      float w[4][LM_KPF];
#pragma unroll
      for (int i = 0; i < LM_KPF; ++i) {
        const int k = k0 + i * 32 + lane;
#pragma unroll
        for (int g = 0; g < 4; ++g) w[g][i] = k < H ? __ldg(Whh + (size_t)(g * H + j) * H + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < LM_KPF; ++i) {
        const int k = k0 + i * 32 + lane;
        if (k0 + i * 32 >= H) break;               // warp-uniform
#pragma unroll
        for (int b = 0; b < LM_BCH; ++b) {
          const float hv = (b < nb && k < H) ? hs[b * H + k] : 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) acc[g][b] = fmaf(hv, w[g][i], acc[g][b]);
        }
      }
    }
    // lane b finishes batch row b0 + b: reduce-scatter with entry index b * 4 + gate, so lane b ends with its four gate sums
    float v[128];
#pragma unroll
    for (int b = 0; b < 32; ++b)
#pragma unroll
      for (int g = 0; g < 4; ++g) v[b * 4 + g] = b < LM_BCH ? acc[g][b < LM_BCH ? b : 0] : 0.f;
    warp_reduce_scatter<4>(v, lane);
    const float a4[4] = {v[0], v[1], v[2], v[3]};
    if (lane < nb) {
      const int b = b0 + lane;
      const float* x = xg + (size_t)b * H4;
      const float gi = sigmoidf_(x[j] + b_hh[j] + a4[0]);
      const float gf = sigmoidf_(x[H + j] + b_hh[H + j] + a4[1]);
      const float gg = tanhf(x[2 * H + j] + b_hh[2 * H + j] + a4[2]);
      const float go = sigmoidf_(x[3 * H + j] + b_hh[3 * H + j] + a4[3]);
      const float c = gf * c_prev[(size_t)b * H + j] + gi * gg;
      float* g = gates + (size_t)b * H4;
      g[j] = gi; g[H + j] = gf; g[2 * H + j] = gg; g[3 * H + j] = go;
      c_out[(size_t)b * H + j] = c;
      h_out[(size_t)b * H + j] = go * tanhf(c);
    }
  }
}
// Backward of one time step, one WARP per hidden unit j (lanes stride the 4H reduction over row j of W_hh^T):
//   dh = dh_above[b, j] + sum_m dgates_next[b, m] * W_hh[m, j]      (recurrent input gradient of step t + 1; absent at t = T - 1)
//   dc = dc_carry + dh * o * (1 - tanh(c)^2);  pre-activation gate gradients -> dgates[b, :];  dc_carry = dc * f
__global__ void __launch_bounds__(LM_WARPS * 32) lstm_step_bwd_kernel(const float* __restrict__ dh_above, const float* __restrict__ dg_next,
                                                                      const float* __restrict__ WhhT, const float* __restrict__ gates,
                                                                      const float* __restrict__ c_t, const float* __restrict__ c_prev,
                                                                      float* __restrict__ dc_carry, float* __restrict__ dgates, int B,
                                                                      int H) {
  extern __shared__ __align__(128) float ds[];    // dgates_next[b0 .. b0 + LM_BCH, :] (4H each)
  __shared__ uint64_t fill_bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * LM_WARPS + warp;
  const int H4 = 4 * H;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&fill_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t fill_phase = 0;
  for (int b0 = 0; b0 < B; b0 += LM_BCH) {
    const int nb = min(LM_BCH, B - b0);
    float rec = 0.f;
    if (dg_next) {
      __syncthreads();                               // barrier initialised / the previous chunk's readers are done
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of ds before the async write
        lm_bulk_load(ds, dg_next + (size_t)b0 * H4, (uint32_t)(nb * H4 * sizeof(float)), &fill_bar);
      }
      lm_bar_wait(&fill_bar, fill_phase);
      fill_phase ^= 1u;
      if (j < H) {
        float acc[LM_BCH];
#pragma unroll
        for (int b = 0; b < LM_BCH; ++b) acc[b] = 0.f;
        const float* w = WhhT + (size_t)j * H4;
        for (int m0 = 0; m0 < H4; m0 += 32 * LM_KPB) {
          float wv[LM_KPB];
#pragma unroll
          for (int i = 0; i < LM_KPB; ++i) { const int m = m0 + i * 32 + lane; wv[i] = m < H4 ? __ldg(w + m) : 0.f; }
#pragma unroll
          for (int i = 0; i < LM_KPB; ++i) {
            const int m = m0 + i * 32 + lane;
            if (m0 + i * 32 >= H4) break;          // warp-uniform
#pragma unroll
            for (int b = 0; b < LM_BCH; ++b) acc[b] = fmaf((b < nb && m < H4) ? ds[b * H4 + m] : 0.f, wv[i], acc[b]);
          }
        }
        float v[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) v[b] = b < LM_BCH ? acc[b < LM_BCH ? b : 0] : 0.f;
        warp_reduce_scatter<1>(v, lane);
        rec = v[0];
      }
    }
    if (j >= H || lane >= nb) continue;
    const int b = b0 + lane;
    const float dh = dh_above[(size_t)b * H + j] + rec;
    const float* g = gates + (size_t)b * H4;
    const float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
    const float tc = tanhf(c_t[(size_t)b * H + j]);
    const float dc = dc_carry[(size_t)b * H + j] + dh * go * (1.f - tc * tc);
    float* d = dgates + (size_t)b * H4;
    d[j] = dc * gg * gi * (1.f - gi);
    d[H + j] = dc * c_prev[(size_t)b * H + j] * gf * (1.f - gf);
    d[2 * H + j] = dc * gi * (1.f - gg * gg);
    d[3 * H + j] = dh * tc * go * (1.f - go);
    dc_carry[(size_t)b * H + j] = dc * gf;
  }
}

// ---- persistent variants: ALL time steps of one layer in ONE launch.  A warp keeps its unit's recurrent weights in
// registers across the whole sequence (the step kernels re-read the 640 KB W_hh from the L2 at every step), the CTAs meet at
// a global-memory barrier between steps (h_t / dgates_t of every unit must be visible before step t +- 1 starts) and the
// cell-gradient carry never leaves its thread.  35 launches and launch gaps per layer become one; eligible when the batch
// fits one chunk (B <= LM_BCH) and the weights fit the register budget (nhid <= 256); MTL_LM_SEQ=0 keeps the step kernels.
constexpr int LM_KPS = 32;                    // W_hh^T values per lane held by the backward sequence kernel (4 * nhid <= 1024)
__device__ __forceinline__ void fill_smem_cg(float* __restrict__ dst, const float* src, int n) {
  // as fill_smem, but through the L2 (ld.global.cg): the data was written by other CTAs of THIS launch
  const int n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i0 = 0; i0 < n4; i0 += 8 * (int)blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * (int)blockDim.x + (int)threadIdx.x; if (i < n4) v[u] = __ldcg(s4 + i); }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * (int)blockDim.x + (int)threadIdx.x; if (i < n4) d4[i] = v[u]; }
  }
}
// all CTAs of the grid have finished step `round` (1-based); bounded spin: a scheduling bug must trap, not hang the GPU
__device__ __forceinline__ void lm_grid_barrier(unsigned* bar, unsigned round) {
  __threadfence();                               // this thread's h / dgates stores -> device scope
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(bar, 1u);
    const unsigned target = round * gridDim.x;
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(LM_WARPS * 32) lstm_seq_fwd_kernel(const float* __restrict__ xg, float* hseq, float* cseq,
                                                                     const float* __restrict__ Whh, const float* __restrict__ b_hh,
                                                                     float* __restrict__ gates, int B, int H, int T, unsigned* bar) {
  extern __shared__ float hs[];                   // h_{t-1}[0 .. B, :]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * LM_WARPS + warp;
  const int H4 = 4 * H;
  const size_t BH = (size_t)B * H;
  float w[4][LM_KPF];
  float bh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < LM_KPF; ++i) {
    const int k = i * 32 + lane;
#pragma unroll
    for (int g = 0; g < 4; ++g) w[g][i] = (j < H && k < H) ? __ldg(Whh + (size_t)(g * H + j) * H + k) : 0.f;
  }
  if (j < H) { bh[0] = b_hh[j]; bh[1] = b_hh[H + j]; bh[2] = b_hh[2 * H + j]; bh[3] = b_hh[3 * H + j]; }
  for (int t = 0; t < T; ++t) {
    fill_smem_cg(hs, hseq + (size_t)t * BH, B * H);
    __syncthreads();
    if (j < H) {
      // two half-chunks of 12 batch rows: 48 accumulators live beside the 32 weight registers (96 spilled)
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int bb = 0; bb < LM_BCH; bb += LM_BCH / 2) {
        float acc[4][LM_BCH / 2];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int b = 0; b < LM_BCH / 2; ++b) acc[g][b] = 0.f;
#pragma unroll
        for (int i = 0; i < LM_KPF; ++i) {
          const int k = i * 32 + lane;
          if (i * 32 >= H) break;                  // warp-uniform
#pragma unroll
          for (int b = 0; b < LM_BCH / 2; ++b) {
            const float hv = (bb + b < B && k < H) ? hs[(bb + b) * H + k] : 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[g][b] = fmaf(hv, w[g][i], acc[g][b]);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int b = 0; b < LM_BCH / 2; ++b) acc[g][b] = warp_sum(acc[g][b]);
#pragma unroll
        for (int b = 0; b < LM_BCH / 2; ++b)
          if (lane == bb + b) { a4[0] = acc[0][b]; a4[1] = acc[1][b]; a4[2] = acc[2][b]; a4[3] = acc[3][b]; }
      }
      if (lane < B) {
        const int b = lane;
        const float* x = xg + ((size_t)t * B + b) * H4;
        const float gi = sigmoidf_(x[j] + bh[0] + a4[0]);
        const float gf = sigmoidf_(x[H + j] + bh[1] + a4[1]);
        const float gg = tanhf(x[2 * H + j] + bh[2] + a4[2]);
        const float go = sigmoidf_(x[3 * H + j] + bh[3] + a4[3]);
        const float c = gf * __ldcg(cseq + (size_t)t * BH + (size_t)b * H + j) + gi * gg;
        float* g = gates + ((size_t)t * B + b) * H4;
        g[j] = gi; g[H + j] = gf; g[2 * H + j] = gg; g[3 * H + j] = go;
        cseq[(size_t)(t + 1) * BH + (size_t)b * H + j] = c;
        hseq[(size_t)(t + 1) * BH + (size_t)b * H + j] = go * tanhf(c);
      }
    }
    if (t + 1 < T) lm_grid_barrier(bar, (unsigned)(t + 1));   // also protects hs against the next step's refill
  }
}
__global__ void __launch_bounds__(LM_WARPS * 32) lstm_seq_bwd_kernel(const float* __restrict__ dh_above, const float* __restrict__ WhhT,
                                                                     const float* __restrict__ gates, const float* __restrict__ cseq,
                                                                     float* dgates, int B, int H, int T, unsigned* bar) {
  extern __shared__ float ds[];                   // dgates_{t+1}[0 .. B, :] (4H each)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * LM_WARPS + warp;
  const int H4 = 4 * H;
  const size_t BH = (size_t)B * H;
  float wv[LM_KPS];
#pragma unroll
  for (int i = 0; i < LM_KPS; ++i) { const int m = i * 32 + lane; wv[i] = (j < H && m < H4) ? __ldg(WhhT + (size_t)j * H4 + m) : 0.f; }
  float dc_carry = 0.f;                            // of (b = lane, j): never leaves this thread
  for (int t = T - 1; t >= 0; --t) {
    float rec = 0.f;
    if (t + 1 < T) {
      fill_smem_cg(ds, dgates + (size_t)(t + 1) * B * H4, B * H4);
      __syncthreads();
      if (j < H) {
        float acc[LM_BCH];
#pragma unroll
        for (int b = 0; b < LM_BCH; ++b) acc[b] = 0.f;
#pragma unroll
        for (int i = 0; i < LM_KPS; ++i) {
          const int m = i * 32 + lane;
          if (i * 32 >= H4) break;                 // warp-uniform
#pragma unroll
          for (int b = 0; b < LM_BCH; ++b) acc[b] = fmaf((b < B && m < H4) ? ds[b * H4 + m] : 0.f, wv[i], acc[b]);
        }
#pragma unroll
        for (int b = 0; b < LM_BCH; ++b) acc[b] = warp_sum(acc[b]);
#pragma unroll
        for (int b = 0; b < LM_BCH; ++b)
          if (lane == b) rec = acc[b];
      }
    }
    if (j < H && lane < B) {
      const int b = lane;
      const float dh = dh_above[(size_t)t * BH + (size_t)b * H + j] + rec;
      const float* g = gates + ((size_t)t * B + b) * H4;
      const float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
      const float tc = tanhf(cseq[(size_t)(t + 1) * BH + (size_t)b * H + j]);
      const float dc = dc_carry + dh * go * (1.f - tc * tc);
      float* d = dgates + ((size_t)t * B + b) * H4;
      d[j] = dc * gg * gi * (1.f - gi);
      d[H + j] = dc * cseq[(size_t)t * BH + (size_t)b * H + j] * gf * (1.f - gf);
      d[2 * H + j] = dc * gi * (1.f - gg * gg);
      d[3 * H + j] = dh * tc * go * (1.f - go);
      dc_carry = dc * gf;
    }
    if (t > 0) lm_grid_barrier(bar, (unsigned)(T - t));
  }
}

// ----------------------------------------------------------------------------- host helpers
int gemm(int mode, const float* A, int lda, int tA, const float* B, int ldb, int tB, float* C, int ldc, int M, int N, int K,
         float beta, const float* bias, int split, cudaStream_t s) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.transA = tA; g.B = B; g.ldb = ldb; g.transB = tB; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.alpha = 1.f; g.beta = beta; g.bias = bias; g.epi = EPI_NONE; g.split_k = split;
  return k_gemm(g, mode, s);
}
inline int ew_grid(size_t n) { size_t g = (n + 255) / 256; return (int)(g > 148 * 8 ? 148 * 8 : (g ? g : 1)); }

struct LmWs {
  uintptr_t base; size_t off = 0;
  void* raw(size_t bytes) { off = (off + 255) & ~(size_t)255; void* p = (void*)(base + off); off += bytes; return p; }
  float* f(size_t n) { return (float*)raw(n * sizeof(float)); }
};
struct LmPlan {
  float *emb, *xg, *dx, *zeros;
  float *hseq[8], *cseq[8], *gates[8], *yd[8], *whhT[8];
  float *logits, *dlogits, *row_lse, *row_loss, *dy, *dgates, *dc;
  int *gold, *hyp;
  unsigned* bar;                     // grid-barrier counters of the sequence kernels: 2 per layer
  CeOut* ce;
  size_t bytes;
};
void plan_ws(const mtl_lm_cfg& c, int T, int B, uintptr_t base, LmPlan& P) {
  LmWs w; w.base = base;
  const size_t R = (size_t)T * B, H = c.nhid, in_max = (size_t)(c.ninp > c.nhid ? c.ninp : c.nhid);
  const int ldp = (c.vocab + 3) & ~3;
  P.emb = w.f(R * c.ninp);
  P.xg = w.f(R * 4 * H);
  P.dx = w.f(R * in_max);
  P.zeros = w.f((size_t)B * H);
  for (int l = 0; l < c.nlayers; ++l) {
    P.hseq[l] = w.f((size_t)(T + 1) * B * H);      // slot 0 = h0, slot t + 1 = h_t
    P.cseq[l] = w.f((size_t)(T + 1) * B * H);
    P.gates[l] = w.f(R * 4 * H);
    P.yd[l] = w.f(R * H);                          // dropout(h sequence) = the next consumer's input
    P.whhT[l] = w.f(4 * H * H);
  }
  P.logits = w.f(R * ldp);
  P.dlogits = w.f(R * ldp);
  P.row_lse = w.f(R); P.row_loss = w.f(R);
  P.dy = w.f(R * H);
  P.dgates = w.f(R * 4 * H);
  P.dc = w.f((size_t)B * H);
  P.gold = (int*)w.raw(R * sizeof(int)); P.hyp = (int*)w.raw(R * sizeof(int));
  P.ce = (CeOut*)w.f(8);
  P.bar = (unsigned*)w.raw(64 * sizeof(unsigned));
  P.bytes = w.off + 256;
}

struct LmSeed { unsigned long long seed; const unsigned long long* dev; unsigned long long mul; };
MtlDrop lm_drop(float p, const LmSeed& sd, uint32_t site) { return p > 0.f ? mtl_drop(p, sd.seed, site, sd.dev, sd.mul) : mtl_nodrop(); }

// forward (+ backward when grad != null: grad += loss_scale * dCE/dtheta); hidden in: h0 / c0 [L, B, H] (null = zeros),
// hidden out: hT / cT (nullable, may alias h0 / c0)
int lm_pass(const mtl_lm_cfg& c, const LmLayout& L, int mode, const float* theta, float* grad, const long long* tokens,
            const long long* targets, int T, int B, const float* h0, const float* c0, float* hT, float* cT, float p_drop,
            LmSeed sd, float loss_scale, void* ws, long long ws_bytes, float* loss_out, float* logits_out, cudaStream_t s) {
  MTL_REQUIRE(theta && tokens && ws && T >= 1 && B >= 1, "null argument");
  MTL_REQUIRE((((uintptr_t)ws) & 255u) == 0, "workspace must be 256B aligned");
  MTL_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout in [0,1)");
  MTL_REQUIRE(targets || (!grad && !loss_out), "targets are required for the loss / backward");
  LmPlan P;
  plan_ws(c, T, B, (uintptr_t)ws, P);
  MTL_REQUIRE((long long)P.bytes <= ws_bytes, "workspace too small (mtl_lm_workspace_bytes)");
  const int H = c.nhid, H4 = 4 * H, V = c.vocab, ldp = (V + 3) & ~3;
  const int R = T * B;
  const size_t BH = (size_t)B * H;
  const dim3 sgrid(mtl_cdiv(H, LM_WARPS));
  const size_t fwd_smem = (size_t)LM_BCH * H * sizeof(float), bwd_smem = (size_t)LM_BCH * H4 * sizeof(float);
  MTL_REQUIRE(bwd_smem <= 200 * 1024, "nhid too large for the recurrent step kernels (16 * 4 * nhid floats of shared memory)");
  {
    static size_t configured = 0;                  // opt in to > 48 KB of dynamic shared memory once per size
    if (bwd_smem > configured) {
      MTL_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
      MTL_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
      configured = bwd_smem;
    }
  }

  // whole-sequence kernels (one launch per layer and direction) when the batch is one chunk and the weights fit registers
  static int seq_env = -1;
  if (seq_env < 0) { const char* e = getenv("MTL_LM_SEQ"); seq_env = (e && e[0] == '1') ? 1 : 0; }
  const bool seq = seq_env && B <= LM_BCH && H <= 32 * LM_KPF && H4 <= 32 * LM_KPS && (int)sgrid.x <= 148;
  if (seq) MTL_CHECK_CUDA(cudaMemsetAsync(P.bar, 0, 64 * sizeof(unsigned), s));
  {
    static size_t configured_seq = 0;
    if (bwd_smem > configured_seq) {
      MTL_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
      MTL_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
      configured_seq = bwd_smem;
    }
  }

  // ---- forward
  lm_embed_fwd_kernel<<<ew_grid((size_t)R * c.ninp), 256, 0, s>>>(tokens, theta + L.enc, lm_drop(p_drop, sd, 0), P.emb, R, c.ninp, V);
  MTL_CHECK_LAUNCH();
  MTL_TRY(k_zero(P.zeros, BH, s));
  const float* x = P.emb;
  int in = c.ninp;
  for (int l = 0; l < c.nlayers; ++l) {
    lm_transpose_kernel<<<dim3(mtl_cdiv(H, 32), mtl_cdiv(H4, 32)), dim3(32, 8), 0, s>>>(theta + L.w_hh[l], P.whhT[l], H4, H);
    MTL_CHECK_LAUNCH();
    MTL_TRY(gemm(mode, x, in, 0, theta + L.w_ih[l], in, 1, P.xg, H4, R, H4, in, 0.f, theta + L.b_ih[l], 1, s));
    MTL_TRY(k_copy(P.hseq[l], h0 ? h0 + (size_t)l * BH : P.zeros, BH, s));
    MTL_TRY(k_copy(P.cseq[l], c0 ? c0 + (size_t)l * BH : P.zeros, BH, s));
    if (seq) {
      lstm_seq_fwd_kernel<<<sgrid, LM_WARPS * 32, fwd_smem, s>>>(P.xg, P.hseq[l], P.cseq[l], theta + L.w_hh[l], theta + L.b_hh[l],
                                                                 P.gates[l], B, H, T, P.bar + 2 * l);
      MTL_CHECK_LAUNCH();
    }
    for (int t = 0; t < T && !seq; ++t) {
      lstm_step_fwd_kernel<<<sgrid, LM_WARPS * 32, fwd_smem, s>>>(P.xg + (size_t)t * B * H4, P.hseq[l] + (size_t)t * BH,
                                                                  P.cseq[l] + (size_t)t * BH, theta + L.w_hh[l], theta + L.b_hh[l],
                                                                  P.gates[l] + (size_t)t * B * H4, P.cseq[l] + (size_t)(t + 1) * BH,
                                                                  P.hseq[l] + (size_t)(t + 1) * BH, B, H);
      MTL_CHECK_LAUNCH();
    }
    const float* y = P.hseq[l] + BH;               // [T*B, H]
    if (p_drop > 0.f) {
      lm_dropout_kernel<<<ew_grid((size_t)R * H), 256, 0, s>>>(y, lm_drop(p_drop, sd, 1 + l), P.yd[l], (size_t)R * H);
      MTL_CHECK_LAUNCH();
      x = P.yd[l];
    } else {
      x = y;
    }
    in = H;
  }
  // hidden out: read the final slots before anything may alias them (hT / cT may be h0 / c0: those were copied above)
  for (int l = 0; l < c.nlayers; ++l) {
    if (hT) MTL_TRY(k_copy(hT + (size_t)l * BH, P.hseq[l] + (size_t)T * BH, BH, s));
    if (cT) MTL_TRY(k_copy(cT + (size_t)l * BH, P.cseq[l] + (size_t)T * BH, BH, s));
  }
  const float* y_last = x;
  MTL_TRY(gemm(mode, y_last, H, 0, theta + L.dec_w, H, 1, P.logits, ldp, R, V, H, 0.f, theta + L.dec_b, 1, s));
  if (logits_out) MTL_CHECK_CUDA(cudaMemcpy2DAsync(logits_out, sizeof(float) * V, P.logits, sizeof(float) * ldp, sizeof(float) * V, R,
                                                  cudaMemcpyDeviceToDevice, s));
  if (!targets) return MTL_OK;
  lm_gold_kernel<<<mtl_cdiv(R, 256), 256, 0, s>>>(targets, P.gold, R, V);
  MTL_CHECK_LAUNCH();
  MTL_TRY(k_ce_fwd(P.logits, ldp, P.gold, R, V, 0.f, -1, P.row_lse, P.row_loss, P.hyp, P.ce, s));   // nn.CrossEntropyLoss(): no ignore index
  if (loss_out) MTL_CHECK_CUDA(cudaMemcpyAsync(loss_out, P.ce, sizeof(CeOut), cudaMemcpyDeviceToDevice, s));
  if (!grad) return MTL_OK;

  // ---- backward
  MTL_TRY(k_ce_bwd(P.logits, ldp, P.gold, P.row_lse, P.ce, loss_scale, 0.f, -1, P.dlogits, R, V, s));
  // K slabs of the weight-gradient contractions: about two waves of CTAs, at least one 32-row k-block each
  auto slabs = [&](int M_, int N_) { const long long tiles = (long long)mtl_cdiv(M_, 128) * mtl_cdiv(N_, 64);
                                     long long sp = (296 + tiles - 1) / tiles; const int kb = mtl_cdiv(R, 32);
                                     return (int)(sp > kb ? kb : (sp > 1 ? sp : 2)); };
  const int wsplit = slabs(H4, H);
  MTL_TRY(gemm(mode, P.dlogits, ldp, 1, y_last, H, 0, grad + L.dec_w, H, V, H, R, 1.f, nullptr, slabs(V, H), s));   // dW_dec += dlogits^T y
  MTL_TRY(k_colsum_acc(P.dlogits, R, V, ldp, grad + L.dec_b, s));
  MTL_TRY(gemm(mode, P.dlogits, ldp, 0, theta + L.dec_w, H, 0, P.dy, H, R, H, V, 0.f, nullptr, 1, s));            // d(y_last) = dlogits W_dec
  for (int l = c.nlayers - 1; l >= 0; --l) {
    if (p_drop > 0.f) {
      lm_dropout_kernel<<<ew_grid((size_t)R * H), 256, 0, s>>>(P.dy, lm_drop(p_drop, sd, 1 + l), P.dy, (size_t)R * H);
      MTL_CHECK_LAUNCH();
    }
    MTL_TRY(k_zero(P.dc, BH, s));
    if (seq) {
      lstm_seq_bwd_kernel<<<sgrid, LM_WARPS * 32, bwd_smem, s>>>(P.dy, P.whhT[l], P.gates[l], P.cseq[l], P.dgates, B, H, T,
                                                                 P.bar + 2 * l + 1);
      MTL_CHECK_LAUNCH();
    }
    for (int t = T - 1; t >= 0 && !seq; --t) {
      lstm_step_bwd_kernel<<<sgrid, LM_WARPS * 32, bwd_smem, s>>>(P.dy + (size_t)t * BH,
                                                                  t + 1 < T ? P.dgates + (size_t)(t + 1) * B * H4 : nullptr,
                                                                  P.whhT[l], P.gates[l] + (size_t)t * B * H4,
                                                                  P.cseq[l] + (size_t)(t + 1) * BH, P.cseq[l] + (size_t)t * BH, P.dc,
                                                                  P.dgates + (size_t)t * B * H4, B, H);
      MTL_CHECK_LAUNCH();
    }
    const int lin = l == 0 ? c.ninp : H;
    const float* xin = l == 0 ? P.emb : (p_drop > 0.f ? P.yd[l - 1] : P.hseq[l - 1] + BH);
    MTL_TRY(gemm(mode, P.dgates, H4, 1, P.hseq[l], H, 0, grad + L.w_hh[l], H, H4, H, R, 1.f, nullptr, wsplit, s));   // dW_hh += dgates^T h_prev
    MTL_TRY(gemm(mode, P.dgates, H4, 1, xin, lin, 0, grad + L.w_ih[l], lin, H4, lin, R, 1.f, nullptr, wsplit, s));   // dW_ih += dgates^T x
    MTL_TRY(k_colsum_acc(P.dgates, R, H4, H4, grad + L.b_ih[l], s));
    MTL_TRY(k_colsum_acc(P.dgates, R, H4, H4, grad + L.b_hh[l], s));
    float* dxo = l == 0 ? P.dx : P.dy;            // input gradient: the layer below's dy (its steps are done with P.dy), or the embedding's
    MTL_TRY(gemm(mode, P.dgates, H4, 0, theta + L.w_ih[l], lin, 0, dxo, lin, R, lin, H4, 0.f, nullptr, 1, s));
  }
  lm_embed_bwd_kernel<<<ew_grid((size_t)R * c.ninp), 256, 0, s>>>(tokens, P.dx, lm_drop(p_drop, sd, 0), grad + L.enc, R, c.ninp, V);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

}  // namespace

// ----------------------------------------------------------------------------- C ABI
extern "C" long long mtl_lm_param_floats(const mtl_lm_cfg* cfg) {
  LmLayout L;
  if (!cfg || build_lm_layout(*cfg, L) != MTL_OK) return -1;
  return (long long)L.total;
}
extern "C" int mtl_lm_param_count(const mtl_lm_cfg* cfg) {
  LmLayout L;
  if (!cfg || build_lm_layout(*cfg, L) != MTL_OK) return -1;
  return L.n;
}
extern "C" int mtl_lm_param_info(const mtl_lm_cfg* cfg, int idx, long long* off, long long* numel) {
  LmLayout L;
  MTL_REQUIRE(cfg && off && numel, "null argument");
  MTL_TRY(build_lm_layout(*cfg, L));
  MTL_REQUIRE(idx >= 0 && idx < L.n, "param index");
  *off = (long long)L.off[idx]; *numel = (long long)L.numel[idx];
  return MTL_OK;
}
extern "C" long long mtl_lm_workspace_bytes(const mtl_lm_cfg* cfg, int T, int B) {
  LmLayout L;
  if (!cfg || T < 1 || B < 1 || build_lm_layout(*cfg, L) != MTL_OK) return -1;
  LmPlan P;
  plan_ws(*cfg, T, B, 0, P);
  return (long long)P.bytes;
}
extern "C" int mtl_lm_pass(const mtl_lm_cfg* cfg, int gemm_mode, const float* theta, float* grad, const long long* tokens,
                           const long long* targets, int T, int B, const float* h0, const float* c0, float* hT, float* cT,
                           float dropout, unsigned long long seed, float loss_scale, void* workspace, long long workspace_bytes,
                           float* loss_out8, float* logits_out, void* stream) {
  MTL_REQUIRE(cfg, "null argument");
  MTL_REQUIRE(gemm_mode >= 0 && gemm_mode <= 2, "gemm mode");
  LmLayout L;
  MTL_TRY(build_lm_layout(*cfg, L));
  LmSeed sd = {seed, nullptr, 0};
  return lm_pass(*cfg, L, gemm_mode, theta, grad, tokens, targets, T, B, h0, c0, hT, cT, dropout, sd, loss_scale, workspace,
                 workspace_bytes, loss_out8, logits_out, (cudaStream_t)stream);
}

// One meta-step (lm/main_meta_transfer.py:293-372), first-order:
//   for task i: theta_i = theta0 - (lr / meta_lr_factor) * clip(dCE(theta0; train_i, hidden));  hidden <- hidden after that train forward
//               G += w_i * dCE(theta_i; val, hidden)
//   theta0 <- theta0 - lr * clip(G)
// theta_work / grad / meta_grad are caller arenas of mtl_lm_param_floats floats; hidden_h / hidden_c [L, B, H] are carried
// across calls (the reference never re-initialises them); results: 16 floats per task (train CeOut, val CeOut).
// seed_slot (nullable): device word holding the step's seed -- nothing in the enqueued work then depends on a host value that
// changes per step, so the caller may capture the call in a CUDA graph and replay it.
extern "C" int mtl_lm_meta_step(const mtl_lm_cfg* cfg, int gemm_mode, float* theta, float* theta_work, float* grad,
                                float* meta_grad, float* hidden_h, float* hidden_c, int n_tasks,
                                const long long* const* train_tokens, const long long* const* train_targets,
                                const long long* val_tokens, const long long* val_targets, int T, int B,
                                const float* task_weights, float lr, float meta_lr_factor, float clip, float dropout,
                                unsigned long long seed, const unsigned long long* seed_slot, void* workspace,
                                long long workspace_bytes, float* results, float* scratch1032, void* stream) {
  MTL_REQUIRE(cfg && theta && theta_work && grad && meta_grad && hidden_h && hidden_c && train_tokens && train_targets &&
                  val_tokens && val_targets && task_weights && scratch1032,
              "null argument");
  MTL_REQUIRE(n_tasks >= 1 && n_tasks <= 64 && meta_lr_factor > 0.f, "1 <= n_tasks <= 64, meta_lr_factor > 0");
  MTL_REQUIRE(gemm_mode >= 0 && gemm_mode <= 2, "gemm mode");
  LmLayout L;
  MTL_TRY(build_lm_layout(*cfg, L));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = L.total;
  MTL_TRY(k_zero(meta_grad, n, s));                                                  // model.zero_grad() of the last forward_one_batch
  for (int i = 0; i < n_tasks; ++i) {
    MTL_TRY(k_copy(theta_work, theta, n, s));                                        // weights_original / load_state_dict
    MTL_TRY(k_zero(grad, n, s));
    // dropout keys: seed * 128 + 2 * task + pass; with a device seed slot (a captured CUDA graph patches it per replay) the
    // seed is read on the device
    LmSeed s0 = {seed * 128ull + 2ull * i, nullptr, 0}, s1 = {seed * 128ull + 2ull * i + 1ull, nullptr, 0};
    if (seed_slot) { s0 = {2ull * i, seed_slot, 128ull}; s1 = {2ull * i + 1ull, seed_slot, 128ull}; }
    MTL_TRY(lm_pass(*cfg, L, gemm_mode, theta_work, grad, train_tokens[i], train_targets[i], T, B, hidden_h, hidden_c, hidden_h,
                    hidden_c, dropout, s0, 1.f, workspace, workspace_bytes, results ? results + 16 * i : nullptr, nullptr, s));
    if (clip > 0.f) {
      MTL_TRY(k_clip_coef(grad, n, clip, scratch1032, scratch1032 + MTL_NORM_PARTIALS, s));
      MTL_TRY(k_scale_by_dev(grad, scratch1032 + MTL_NORM_PARTIALS + 1, n, s));
    }
    MTL_TRY(k_sgd(theta_work, grad, lr / meta_lr_factor, n, s));                     // inner_opt.step()
    MTL_TRY(lm_pass(*cfg, L, gemm_mode, theta_work, meta_grad, val_tokens, val_targets, T, B, hidden_h, hidden_c, nullptr, nullptr,
                    dropout, s1, task_weights[i], workspace, workspace_bytes, results ? results + 16 * i + 8 : nullptr, nullptr, s));
  }
  if (clip > 0.f) {
    MTL_TRY(k_clip_coef(meta_grad, n, clip, scratch1032, scratch1032 + MTL_NORM_PARTIALS, s));
    MTL_TRY(k_scale_by_dev(meta_grad, scratch1032 + MTL_NORM_PARTIALS + 1, n, s));
  }
  MTL_TRY(k_sgd(theta, meta_grad, lr, n, s));                                        // outer_opt.step()
  return MTL_OK;
}
