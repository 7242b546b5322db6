// Internal (C++ linkage) launch wrappers.  Every function enqueues work on `s` and returns
// MTL_OK or a negative error code; none of them synchronises or allocates.
#pragma once
#include "common.cuh"

// ----------------------------------------------------------------------------- arena.cu
// Flat fp32 "arena" ops: the parameter set, its gradient, the copy-grad accumulator, the theta0
// snapshot and the Adam moments are sibling flat buffers, so every optimizer-side operation of
// the meta-step is one streaming kernel (replaces the per-tensor loops at
// models/asr/transformer.py:204-240 and torch.optim.SGD/Adam).
int k_zero(float* p, size_t n, cudaStream_t s);
int k_copy(float* dst, const float* src, size_t n, cudaStream_t s);
int k_axpy(float* y, const float* x, float a, size_t n, cudaStream_t s);            // y += a*x
int k_scale_by_dev(float* y, const float* coef_dev, size_t n, cudaStream_t s);      // y *= *coef
int k_sgd(float* p, const float* g, float lr, size_t n, cudaStream_t s);            // p -= lr*g
// g *= *coef_dev (when non-null) ; out = in - lr*g   (inner step from the shared theta0 into a lane's adapted copy)
int k_sgd_out(float* out, const float* in, float* g, const float* coef_dev, float lr, size_t n, cudaStream_t s);
// clip_grad_norm_: writes total L2 norm to out[0] and coef=min(1,max_norm/(norm+1e-6)) to out[1].
// `partial` must hold MTL_NORM_PARTIALS floats.
#define MTL_NORM_PARTIALS 1024
int k_clip_coef(const float* g, size_t n, float max_norm, float* partial, float* out2, cudaStream_t s);
// Adam (torch.optim.Adam defaults).  state3 = {int step, float step_size, float bc2_sqrt} on
// device; k_adam_prep increments step and refreshes the two floats so the step is graph-safe.
int k_adam_prep(int* step_dev, float* coef2_dev, double lr, double b1, double b2, cudaStream_t s);
int k_adam(float* p, const float* g, float* m, float* v, const float* coef2_dev, double b1, double b2,
           double eps, size_t n, cudaStream_t s);

// ----------------------------------------------------------------------------- gemm_simt.cu / gemm_tc.cu
enum { EPI_NONE = 0, EPI_RELU = 1, EPI_RELU_BWD = 2 };
struct GemmArgs {
  // C[M,N] = epi(alpha * op(A)[M,K] * op(B)[K,N] + bias[N]) + beta*C
  // A is stored [M,K] (transA=0, row stride lda) or [K,M] (transA=1, row stride lda);
  // B is stored [K,N] (transB=0, row stride ldb) or [N,K] (transB=1, row stride ldb).
  const float* A; const float* B; float* C;
  int M, N, K;
  int lda, ldb, ldc;
  int transA, transB;
  float alpha, beta;
  const float* bias;   // nullable, length N
  int epi;             // EPI_*
  const float* aux;    // EPI_RELU_BWD: same shape/ld as C; output zeroed where aux<=0
  int split_k;         // >1: partial sums are atomically added to C (requires beta==1, epi NONE, no bias)
};
int k_gemm_simt(const GemmArgs& g, cudaStream_t s);
int k_gemm_tc(const GemmArgs& g, int precision_mode, cudaStream_t s);   // tcgen05 path
bool k_gemm_tc_eligible(const GemmArgs& g);
// Dispatcher: mode 0 = SIMT fp32 (exact), 1 = tcgen05 TF32, 2 = tcgen05 3xTF32.
int k_gemm(const GemmArgs& g, int mode, cudaStream_t s);
// Fused rank-r projection pair (tcgen05; modules/common_layers.py:250-257,287-289,303): for each of the G <= 3 problems
//   a (+)= x . W1 ;  y (+)= (x . W1) . W2 (+ bias)          -- K slabs are merged by TMA reduce-add, so a and y must be
// zero (or, for y, hold the sum being accumulated) when the kernel starts.
//   bwd == 0: x [M, K1] (ldx), W1 = linear_a.weight [r, K1], W2 = linear_b.weight [N2, r], bias [N2] or null
//   bwd == 1: x = dy [M, K1] (ldx), W1 = linear_b.weight [K1, r], W2 = linear_a.weight [r, N2]
// a is [M, r] (row stride r), y is [M, N2] (ldy).  ctas: CTA budget that sizes the K slabs (0 = no K split).
struct LrPairArgs {
  int G;
  const float* x[3]; int ldx[3];
  const float* w1[3]; const float* w2[3]; const float* bias[3];
  float* a[3]; float* y[3]; int ldy[3];
  int M, K1, r, N2, bwd, ctas;
};
bool k_lowrank_pair_eligible(const LrPairArgs& a);
int k_lowrank_pair(const LrPairArgs& a, int precision_mode, cudaStream_t s);
// Implicit-GEMM 3x3 convolutions on NHWC activations (tcgen05; precision_mode 1 = TF32, 2 = 3xTF32):
//   y[pixel,co] = epi(sum x[pixel+tap,ci] * wg[co, tap*Cin+ci] + bias[co])          (forward, and dgrad on dY with flipped taps)
//   dwgT[tap*Cin+ci, co] += sum_pixel x[pixel+tap,ci] * dy[pixel,co]                  (weight gradient, atomics)
// w_split: wg holds [hi | lo] tf32 halves (2 x Cout*9*Cin floats, written by k_conv_w_*_layout with split = 1); the
// 3xTF32 kw-box kernel needs it -- k_conv3x3_w_split(mode, Cout) says whether this build / environment uses it.
// pool_out (nullable, forward with EPI_RELU only, when k_conv3x3_pool_fusable): the 2x2 floor-max-pooled output
// [B, F/2, T/2, Cout] is written from the same epilogue (models/asr/transformer.py:51,58: ReLU -> MaxPool2d(2, 2))
int k_conv3x3_tc(const float* x, const float* wg, const float* bias, float* y, int B, int F, int T, int Cin, int Cout,
                 int epi, const float* aux, int precision_mode, int w_split, cudaStream_t s, float* pool_out = nullptr);
bool k_conv3x3_pool_fusable(int precision_mode, int Cout, int w_split);
bool k_conv3x3_pool_fuse_default();   // MTL_CONV_POOL_FUSE=1: the engine uses the fused epilogue (measured slower: default off)
int k_conv3x3_w_split(int precision_mode, int Cout);
int k_conv3x3_wgrad_tc(const float* x, const float* dy, float* dwgT, int B, int F, int T, int Cin, int Cout,
                       int precision_mode, cudaStream_t s);

// ----------------------------------------------------------------------------- norm_embed.cu
// z = drop(y)*? + res ; xhat=(z-mean)*rstd ; out = (xhat*gamma+beta) [+ pe[row % pe_period]] ; out *= rowmask[row]
int k_ln_fwd(const float* y, const float* res, const float* gamma, const float* beta,
             const float* rowmask, const float* pe, int pe_period, MtlDrop drop,
             float* out, float* xhat, float* rstd, int M, int d, cudaStream_t s);
// dz -> dres_out (= or +=), dy = dz*dropmask ; dgamma/dbeta accumulated (+=)
int k_ln_bwd(const float* dout, const float* xhat, const float* rstd, const float* gamma,
             const float* rowmask, MtlDrop drop, float* dy, float* dres, int dres_accumulate,
             float* dgamma, float* dbeta, int M, int d, cudaStream_t s);
// dgamma / dbeta == null: only the activation gradient (the caller enqueues k_ln_param_grad where it wants it)
int k_ln_param_grad(const float* dout, const float* xhat, const float* rowmask, float* dgamma, float* dbeta, int M, int d,
                    cudaStream_t s);
int k_colsum_acc(const float* x, int M, int N, int ld, float* out, cudaStream_t s);  // out[n] += sum_m x[m,n]
int k_embed_fwd(const int* tok, const float* E, const float* pe, MtlDrop drop, float* out,
                int B, int n, int d, cudaStream_t s);
int k_embed_bwd(const int* tok, const float* dout, MtlDrop drop, float* dE, int B, int n, int d,
                int pad_id, cudaStream_t s);
// decoder.py:55-69 on device: trg (B,L) int64 -> seq_in/seq_out (B,n) int32, rowmask (B,n), keypad (B,n)
int k_dec_preprocess(const long long* trg, int B, int L, int n, int* seq_in, int* seq_out,
                     float* rowmask, unsigned char* keypad, int* overflow_flag, cudaStream_t s);
// encoder masks from RAW lengths against the T' axis (encoder.py:64-66)
int k_enc_masks(const int* lens, int B, int Tp, float* rowmask, unsigned char* keypad, cudaStream_t s);

// ----------------------------------------------------------------------------- ce.cu
struct CeOut {        // device-resident result block (8 floats)
  float loss;         // mean CE over non-pad rows
  float n_valid;      // number of non-pad rows
  float n_correct;    // argmax == gold on non-pad rows
  float pad_[5];
};
int k_ce_fwd(const float* logits, int ld, const int* gold, int M, int V, float smoothing, int pad_id,
             float* row_lse, float* row_loss, int* hyp, CeOut* out, cudaStream_t s);
// dlogits = scale_host * (*scale_dev or 1) / n_valid * (softmax - target) on non-pad rows, 0 elsewhere
int k_ce_bwd(const float* logits, int ld, const int* gold, const float* row_lse, const CeOut* out,
             float scale, float smoothing, int pad_id, float* dlogits, int M, int V, cudaStream_t s);

// ----------------------------------------------------------------------------- attention.cu
struct AttnArgs {
  const float *q, *k, *v;     // q [B*Tq, ldq], k/v [B*Tk, ldk]; head h at column h*dk
  float* o;                    // [B*Tq, ldo]
  float* lse;                  // [B,H,Tq]
  const unsigned char* keypad; // [B,Tk], 1 = masked
  int B, H, Tq, Tk, dk;        // dk == dv
  int ldq, ldk, ldv, ldo;
  int causal;
  float inv_temp;              // 1/sqrt(dk)
  MtlDrop drop;
  int prec;                    // tensor-core kernels: 0 = 3xTF32 (fp32-grade), 1 = single-pass TF32; 2 = CUDA-core fp32 kernels
};
int k_attn_fwd(const AttnArgs& a, cudaStream_t s);
struct AttnBwdArgs {
  AttnArgs f;
  const float* d_o;            // [B*Tq, ldo]
  float* delta;                // [B,H,Tq] scratch
  float *dq, *dk, *dv;         // same strides as q,k,v; overwritten
};
int k_attn_bwd(const AttnBwdArgs& a, cudaStream_t s);

// ----------------------------------------------------------------------------- conv.cu  (NHWC activations)
int k_conv1_fwd(const float* x, const float* w, const float* b, float* out, int B, int F, int T, int Cout, cudaStream_t s);
int k_conv1_wgrad(const float* x, const float* dout, float* dw, float* db, int B, int F, int T, int Cout, cudaStream_t s);
int k_im2col3x3(const float* x, float* col, int B, int F, int T, int C, cudaStream_t s);   // col [B*F*T, 9*C], (tap,c) order
// conv weight [Cout,Cin,3,3] -> GEMM layouts: fwd Wg[co,(tap,ci)], dgrad Wd[ci,(tap',co)] with flipped taps
// split != 0: the output holds two matrices, hi = tf32(w) and, Cout*9*Cin floats later, lo = tf32(w - hi)
int k_conv_w_fwd_layout(const float* w, float* wg, int Cout, int Cin, int split, cudaStream_t s);
int k_conv_w_dgrad_layout(const float* w, float* wd, int Cout, int Cin, int split, cudaStream_t s);
int k_conv_wgrad_scatter(const float* dwg, float* dw, int Cout, int Cin, cudaStream_t s);  // dw[co,ci,kh,kw] += dwg[co,(tap,ci)]
int k_conv_wgrad_scatter_t(const float* dwgT, float* dw, int Cout, int Cin, cudaStream_t s);  // dw[co,ci,kh,kw] += dwgT[(tap,ci),co]
int k_maxpool2_fwd(const float* x, float* out, int B, int F, int T, int C, cudaStream_t s);
// dx = (x>0 && x is the first max of its 2x2 window) ? dpool : 0
int k_maxpool2_relu_bwd(const float* x, const float* dpool, float* dx, int B, int F, int T, int C, cudaStream_t s);
// p4 [B,F4,T4,C] -> feat [B,T4,C*F4] (feature index c*F4+f), and back
int k_feat_transpose(const float* p4, float* feat, int B, int F4, int T4, int C, cudaStream_t s);
int k_feat_transpose_bwd(const float* dfeat, float* dp4, int B, int F4, int T4, int C, cudaStream_t s);

// ----------------------------------------------------------------------------- spectrogram.cu
// log1p(|STFT|) of a wave (centred, reflect-padded frames) into out[bin * ld_out + frame], optionally normalised over
// the utterance (unbiased std); stat2 = 2 doubles of device scratch
int k_spectrogram(const float* wav, int n, int n_fft, int hop, const float* window, float* out, int ld_out, int normalize,
                  double* stat2, cudaStream_t s);
