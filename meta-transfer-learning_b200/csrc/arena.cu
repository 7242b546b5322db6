// Flat-arena optimizer kernels (HBM-bound streaming; float4 vectorised, grid-stride over a
// multiple of the SM count).  Reference: trainer/asr/transient_trainer.py:160,165,207,229,237,248,255
// and models/asr/transformer.py:204-240 (copy_grad API), torch.optim.SGD/Adam, clip_grad_norm_.
#include "kernels.h"
#include <math.h>

static inline int arena_grid(size_t n4) {
  long long b = (long long)((n4 + 255) / 256);
  const long long cap = 148LL * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename F>
__global__ void __launch_bounds__(256) ew_kernel(size_t n, F f) {
  size_t n4 = n >> 2;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) f.vec(i);
  // tail
  size_t t = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) f.one(t);
}

struct ZeroF { float* p;
  __device__ void vec(size_t i) { reinterpret_cast<float4*>(p)[i] = make_float4(0, 0, 0, 0); }
  __device__ void one(size_t i) { p[i] = 0.f; } };
struct CopyF { float* d; const float* s;
  __device__ void vec(size_t i) { reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i]; }
  __device__ void one(size_t i) { d[i] = s[i]; } };
struct AxpyF { float* y; const float* x; float a;
  __device__ void vec(size_t i) {
    float4 u = reinterpret_cast<float4*>(y)[i]; float4 w = reinterpret_cast<const float4*>(x)[i];
    u.x += a * w.x; u.y += a * w.y; u.z += a * w.z; u.w += a * w.w;
    reinterpret_cast<float4*>(y)[i] = u; }
  __device__ void one(size_t i) { y[i] += a * x[i]; } };
struct ScaleDevF { float* y; const float* c;
  __device__ void vec(size_t i) { float a = *c; float4 u = reinterpret_cast<float4*>(y)[i];
    u.x *= a; u.y *= a; u.z *= a; u.w *= a; reinterpret_cast<float4*>(y)[i] = u; }
  __device__ void one(size_t i) { y[i] *= *c; } };
struct SgdF { float* p; const float* g; float lr;
  __device__ void vec(size_t i) {
    float4 u = reinterpret_cast<float4*>(p)[i]; float4 w = reinterpret_cast<const float4*>(g)[i];
    u.x -= lr * w.x; u.y -= lr * w.y; u.z -= lr * w.z; u.w -= lr * w.w;
    reinterpret_cast<float4*>(p)[i] = u; }
  __device__ void one(size_t i) { p[i] -= lr * g[i]; } };

// Inner SGD step out of place, with the clip coefficient folded in: g *= *coef (kept: the validation backward accumulates
// onto the clipped gradient), out = in - lr * g.  Same fp32 operations as scale_by_dev + copy + sgd, two arena passes fewer.
struct SgdOutF { float* out; const float* in; float* g; const float* coef; float lr;
  __device__ void vec(size_t i) {
    float4 w = reinterpret_cast<float4*>(g)[i];
    if (coef) { const float c = *coef; w.x *= c; w.y *= c; w.z *= c; w.w *= c; reinterpret_cast<float4*>(g)[i] = w; }
    float4 u = reinterpret_cast<const float4*>(in)[i];
    u.x -= lr * w.x; u.y -= lr * w.y; u.z -= lr * w.z; u.w -= lr * w.w;
    reinterpret_cast<float4*>(out)[i] = u; }
  __device__ void one(size_t i) { float w = g[i]; if (coef) { w *= *coef; g[i] = w; } out[i] = in[i] - lr * w; } };

struct AdamF {
  float* p; const float* g; float* m; float* v; const float* coef; float b2, omb1, omb2, eps;
  __device__ __forceinline__ void upd(float& pp, float gg, float& mm, float& vv, float step_size, float bc2s) {
    mm = mm + omb1 * (gg - mm);                       // exp_avg.lerp_(grad, 1-beta1)
    vv = vv * b2 + omb2 * gg * gg;                   // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    float denom = sqrtf(vv) / bc2s + eps;
    pp = pp - step_size * (mm / denom);
  }
  __device__ void vec(size_t i) {
    float ss = coef[0], bc = coef[1];
    float4 P = reinterpret_cast<float4*>(p)[i]; float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 Mv = reinterpret_cast<float4*>(m)[i]; float4 V = reinterpret_cast<float4*>(v)[i];
    upd(P.x, G.x, Mv.x, V.x, ss, bc); upd(P.y, G.y, Mv.y, V.y, ss, bc);
    upd(P.z, G.z, Mv.z, V.z, ss, bc); upd(P.w, G.w, Mv.w, V.w, ss, bc);
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = Mv; reinterpret_cast<float4*>(v)[i] = V;
  }
  __device__ void one(size_t i) { upd(p[i], g[i], m[i], v[i], coef[0], coef[1]); }
};

#define LAUNCH_EW(F_, n, s)                                        \
  do {                                                             \
    if ((n) == 0) return MTL_OK;                                   \
    ew_kernel<<<arena_grid((n) >> 2), 256, 0, s>>>((n), F_);       \
    MTL_CHECK_LAUNCH();                                            \
    return MTL_OK;                                                 \
  } while (0)

static bool al16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int k_zero(float* p, size_t n, cudaStream_t s) { MTL_REQUIRE(al16(p), "arena not 16B aligned"); ZeroF f{p}; LAUNCH_EW(f, n, s); }
int k_copy(float* d, const float* src, size_t n, cudaStream_t s) { MTL_REQUIRE(al16(d) && al16(src), "arena not 16B aligned"); CopyF f{d, src}; LAUNCH_EW(f, n, s); }
int k_axpy(float* y, const float* x, float a, size_t n, cudaStream_t s) { MTL_REQUIRE(al16(y) && al16(x), "arena not 16B aligned"); AxpyF f{y, x, a}; LAUNCH_EW(f, n, s); }
int k_scale_by_dev(float* y, const float* c, size_t n, cudaStream_t s) { MTL_REQUIRE(al16(y), "arena not 16B aligned"); ScaleDevF f{y, c}; LAUNCH_EW(f, n, s); }
int k_sgd(float* p, const float* g, float lr, size_t n, cudaStream_t s) { MTL_REQUIRE(al16(p) && al16(g), "arena not 16B aligned"); SgdF f{p, g, lr}; LAUNCH_EW(f, n, s); }
int k_sgd_out(float* out, const float* in, float* g, const float* coef_dev, float lr, size_t n, cudaStream_t s) {
  MTL_REQUIRE(al16(out) && al16(in) && al16(g), "arena not 16B aligned");
  SgdOutF f{out, in, g, coef_dev, lr};
  LAUNCH_EW(f, n, s);
}
int k_adam(float* p, const float* g, float* m, float* v, const float* coef, double b1, double b2, double eps,
           size_t n, cudaStream_t s) {
  MTL_REQUIRE(al16(p) && al16(g) && al16(m) && al16(v), "arena not 16B aligned");
  // scalars are rounded to fp32 from double exactly as torch does for Python-float hyper-parameters
  AdamF f{p, g, m, v, coef, (float)b2, (float)(1.0 - b1), (float)(1.0 - b2), (float)eps};
  LAUNCH_EW(f, n, s);
}

__global__ void adam_prep_kernel(int* step, float* coef, double lr, double b1, double b2) {
  int t = *step + 1;
  *step = t;
  double bc1 = 1.0 - pow(b1, (double)t);
  double bc2 = 1.0 - pow(b2, (double)t);
  coef[0] = (float)(lr / bc1);
  coef[1] = (float)sqrt(bc2);
}
int k_adam_prep(int* step_dev, float* coef2_dev, double lr, double b1, double b2, cudaStream_t s) {
  adam_prep_kernel<<<1, 1, 0, s>>>(step_dev, coef2_dev, lr, b1, b2);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}

// ---- L2 norm (deterministic two-stage) + clip coefficient
__global__ void __launch_bounds__(256) sumsq_partial(const float* __restrict__ g, size_t n, float* partial) {
  __shared__ float red[32];
  size_t n4 = n >> 2;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 u = reinterpret_cast<const float4*>(g)[i];
    acc += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
  }
  size_t t = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) acc += g[t] * g[t];
  float tot = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) clip_finalize(const float* partial, int np, float max_norm, float* out2) {
  __shared__ float red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) acc += (double)partial[i];
  // double block reduce via two float halves is overkill; np<=1024 so do it in float pairs
  float hi = (float)acc, lo = (float)(acc - (double)hi);
  float shi = block_sum(hi, red);
  float slo = block_sum(lo, red);
  if (threadIdx.x == 0) {
    float norm = sqrtf(shi + slo);
    out2[0] = norm;
    float c = max_norm / (norm + 1e-6f);
    out2[1] = c < 1.f ? c : 1.f;
  }
}
int k_clip_coef(const float* g, size_t n, float max_norm, float* partial, float* out2, cudaStream_t s) {
  MTL_REQUIRE(al16(g), "arena not 16B aligned");
  int grid = arena_grid(n >> 2);
  if (grid > MTL_NORM_PARTIALS) grid = MTL_NORM_PARTIALS;
  sumsq_partial<<<grid, 256, 0, s>>>(g, n, partial);
  MTL_CHECK_LAUNCH();
  clip_finalize<<<1, 256, 0, s>>>(partial, grid, max_norm, out2);
  MTL_CHECK_LAUNCH();
  return MTL_OK;
}
