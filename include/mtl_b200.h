/* libmtl_b200 -- C ABI of the B200-native meta-transfer training step.
 *
 * The reference (audioku/meta-transfer-learning) has no FFI layer: its hot path is Python calling
 * PyTorch.  This header is therefore the boundary a maintainer would bind with ctypes (see
 * INTEGRATION.md) to replace, per entry point, the reference code cited beside it.  Paths are
 * relative to the reference tree.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; mtl_last_error() gives the message
 *     (thread-local).  Nothing here falls back to the CPU: without a CUDA device every compute
 *     entry point fails with MTL_ERR_CUDA.
 *   - all data pointers are DEVICE pointers owned by the caller (fp32 unless stated); the library
 *     never allocates device memory and never synchronises: every call only enqueues kernels on
 *     the given stream (cudaStream_t passed as void*), so a whole meta-step can be captured in a
 *     CUDA graph.
 *   - activations are NHWC inside the VGG front-end; (rows, features) row-major elsewhere.
 */
#ifndef MTL_B200_H
#define MTL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTL_ABI_VERSION 2

/* GEMM engines (mtl_session_set_gemm_mode / mtl_gemm `mode`) */
#define MTL_GEMM_SIMT_FP32 0 /* CUDA-core fp32, exact                        */
#define MTL_GEMM_TC_TF32 1   /* tcgen05.mma kind::tf32, TMA + TMEM           */
#define MTL_GEMM_TC_3XTF32 2 /* tcgen05 with hi/lo split operands (~fp32)    */

const char* mtl_last_error(void);
int mtl_abi_version(void);
/* number of kernels this library has enqueued so far in this process */
unsigned long long mtl_launch_count(void);

/* ------------------------------------------------------------------ model layout / session
 * utils/functions.py:307-351 (init_transformer_model), modules/encoder.py:20-51,
 * modules/decoder.py:19-53, models/asr/transformer.py:22-76: the parameter set.  A session fixes the
 * hyper-parameters, derives the flat parameter arena (tensors in model.parameters() order, each
 * start padded to 64 floats) and remembers the activation record of the last forward pass. */
typedef struct mtl_model_cfg {
  int n_enc, n_dec, d_model, n_heads, d_k, d_v, d_inner, rank, vocab, n_freq;
} mtl_model_cfg;
typedef struct mtl_session mtl_session;

int mtl_session_create(const mtl_model_cfg* cfg, mtl_session** out);
void mtl_session_destroy(mtl_session* s);
int mtl_session_set_gemm_mode(mtl_session* s, int mode);
/* Per-operation-class precision policy (tensor-core engines only; the fp32 CUDA-core engine is all-or-nothing).
 * Every contraction of the pass belongs to one class; its engine is op_mode[class] if set, else the session mode.
 * mode: MTL_GEMM_TC_TF32, MTL_GEMM_TC_3XTF32, or -1 = follow the session mode.  The classes name the
 * reference's nn.Conv2d / nn.Linear call sites: VGG convs (models/asr/transformer.py:47-59) forward, input
 * gradient and weight gradient; the nn.Linear layers of the attention / FFN blocks
 * (modules/common_layers.py:117-118,250-257); the encoder input Linear (modules/encoder.py:44) and the
 * vocabulary projection (modules/decoder.py:50) in all three directions. */
#define MTL_OP_CONV_FWD 0
#define MTL_OP_CONV_DGRAD 1
#define MTL_OP_CONV_WGRAD 2
#define MTL_OP_LIN_FWD 3
#define MTL_OP_LIN_DGRAD 4
#define MTL_OP_LIN_WGRAD 5
#define MTL_OP_STEM 6
#define MTL_OP_VOCAB 7
#define MTL_OP_ATTN 8 /* QK^T / PV contractions of sequences longer than 64 (modules/common_layers.py:321-329) */
#define MTL_OP_CLASSES 9
int mtl_session_set_op_mode(mtl_session* s, int op_class, int mode);
/* Engine variants kept for A/B measurements.  "merge_lowrank" (default 0): run each low-rank projection pair
 * B(A x) of modules/common_layers.py:287-289,303 as ONE GEMM against W = B.A formed once per pass. */
int mtl_session_set_flag(mtl_session* s, const char* name, int value);
long long mtl_param_arena_floats(const mtl_session* s);
int mtl_param_count(const mtl_session* s);
int mtl_param_info(const mtl_session* s, int idx, long long* offset_floats, long long* numel);
/* bytes of workspace one forward+backward of a (B, T frames, n = max target length + 1) batch needs */
long long mtl_workspace_bytes(mtl_session* s, int B, int T, int n);

/* One batch as SpectrogramDataset.sample / AudioDataLoader deliver it (utils/data_loader.py:245-321):
 * x (B,1,F,T) fp32 zero padded, lens (B) int32 RAW frame counts, trg (B,L) int64 PAD=0.
 * n = 1 + max_b #non-PAD targets (the reference derives it inside Decoder.preprocess,
 * modules/decoder.py:55-69; the host knows it before the H2D copy).
 * Optional outputs (device, may be NULL): hyp_out/gold_out (B*n int32), ce_out (8 floats:
 * loss, n_valid, n_correct, ...). */
typedef struct mtl_batch {
  const float* x;
  const int* lens;
  const long long* trg;
  int B, T, L, n;
  int* hyp_out;
  int* gold_out;
  float* ce_out;
} mtl_batch;

/* Transformer.forward (models/asr/transformer.py:120-149) + calculate_metrics CE
 * (utils/metrics.py:68-126).  pe_enc / pe_dec are the PositionalEncoding buffers
 * (modules/common_layers.py:86-108), row-major (max_len, d_model).
 * On return *pred_out points at the logits inside the workspace, (B*n) rows of *ldp_out floats,
 * the first `vocab` of each row valid. */
int mtl_asr_forward(mtl_session* s, const float* theta, const float* pe_enc, const float* pe_dec,
                    void* workspace, long long workspace_bytes, const mtl_batch* batch, float dropout,
                    unsigned long long seed, float label_smoothing, void* stream, float** pred_out,
                    int* ldp_out);
/* loss.backward() for the pass last run by mtl_asr_forward on this session; gradients are
 * ACCUMULATED into `grad` (same layout as theta), scaled by loss_scale
 * (trainer/asr/transient_trainer.py:198-199,226-227).  If dpred_ext != NULL it is used as the
 * gradient w.r.t. the logits ((B*n) x ld_ext) instead of the fused CE gradient. */
int mtl_asr_backward(mtl_session* s, const float* theta, float* grad, float loss_scale,
                     const float* dpred_ext, int ld_ext, void* stream);

/* ------------------------------------------------------------------ input features
 * SpectrogramParser.parse_audio (utils/data_loader.py:65-96): log1p(|STFT|) of one utterance, n_fft = window =
 * sample_rate * window_size samples, hop = sample_rate * window_stride, centred frames with reflect padding (the
 * librosa.stft defaults the reference relies on), optional utterance-level (x - mean) / std (unbiased).
 * wav: n_samples floats (device); window: n_fft floats (device); out: (n_fft / 2 + 1) rows of ld_out floats, the first
 * 1 + n_samples / hop of each row are written -- so an utterance can be written straight into its slice
 * x[b, 0, :, :T_b] of the zero-padded batch tensor (ld_out = T_max).  stat2: 2 doubles of device scratch. */
int mtl_spectrogram(const float* wav, int n_samples, int n_fft, int hop, const float* window, float* out, int ld_out,
                    int normalize, double* stat2, void* stream);

/* ------------------------------------------------------------------ inference: encode + greedy search
 * mtl_asr_encode: Transformer.encode (models/asr/transformer.py:151-160): VGG front-end, flatten, Encoder.forward;
 * enc_out (device) receives (B * T') x d_model rows, T' = (T / 2) / 2.
 * mtl_asr_greedy: Decoder.greedy_search (modules/decoder.py:131-184) on an encoder output: out_tokens (device,
 * B x max_steps int32) receives the arg-max token of every step (the caller cuts each row at the first EOS, :175-183).
 * Masks as in the reference: all-ones non-pad mask, subsequent-only self-attention mask, no encoder-side key mask. */
long long mtl_encode_workspace_bytes(mtl_session* s, int B, int T);
int mtl_asr_encode(mtl_session* s, const float* theta, const float* pe_enc, void* workspace,
                   long long workspace_bytes, const mtl_batch* batch, float* enc_out, void* stream);
long long mtl_greedy_workspace_bytes(mtl_session* s, int B, int Tp, int max_steps);
int mtl_asr_greedy(mtl_session* s, const float* theta, const float* pe_dec, void* workspace,
                   long long workspace_bytes, const float* enc_out, int B, int Tp, int start_token, int max_steps,
                   int* out_tokens, void* stream);

/* ------------------------------------------------------------------ meta-step pieces
 * trainer/asr/transient_trainer.py:178-237 for ONE task, is_copy_grad=True:
 *   grad <- d CE(theta; train)            (:188-199)
 *   [clip_grad_norm_(grad, max_norm)]     (:205-206)
 *   theta <- theta - lr*grad              (:207, SGD)
 *   grad += d [CE(theta; val) * val_scale]   (:215-227; NO zero_grad in between, so the train
 *                                          gradient stays in .grad -- the reference's behaviour)
 *   copy_grad += grad                     (:229, models/asr/transformer.py:219-224)
 *   theta <- theta0                       (:237)
 * results (device, 16 floats): [0..7] train CE block, [8..15] val CE block. */
typedef struct mtl_meta_hparams {
  float lr;            /* inner SGD lr (args.lr)                 */
  float val_scale;     /* 1/N                                    */
  int clip;            /* args.clip                              */
  float max_norm;      /* args.max_norm                          */
  float dropout;       /* args.dropout                           */
  float label_smoothing;
  unsigned long long seed;
} mtl_meta_hparams;
int mtl_meta_task(mtl_session* s, float* theta, const float* theta0, float* grad, float* copy_grad,
                  const float* pe_enc, const float* pe_dec, void* workspace, long long workspace_bytes,
                  const mtl_batch* train, const mtl_batch* val, const mtl_meta_hparams* hp,
                  float* results16, void* stream);
/* ALL tasks of one meta-step (transient_trainer.py:155-237), tasks running concurrently:
 *   copy_grad <- 0
 *   task t (on lane t % n_lanes, own stream): theta_l <- theta (replaces deepcopy(state_dict) +
 *     load_state_dict: theta itself is never modified); the mtl_meta_task body at theta_l;
 *     copy_grad += grad_l, ordered task 0, 1, 2, ... by events so the fp32 sum has the reference's order.
 * theta is left untouched; the caller all-reduces copy_grad when the tasks are sharded over GPUs
 * and then calls mtl_meta_finish.  Dropout seed of task t, pass p (0 train, 1 val): hp.seed*128 + 2t + p.
 * use_graph != 0 (needs seed_slot, 8 device bytes): the second call with identical pointers, shapes and
 * hyper-parameters captures the whole step into a CUDA graph; later calls replay it (only hp.seed may
 * differ -- it is patched into the graph), cutting ~1000 launches per task to one. */
typedef struct mtl_lane {
  float* theta;             /* adapted weights of the task running on this lane (arena-sized) */
  float* grad;              /* its gradient (arena-sized)                                     */
  void* workspace;          /* >= mtl_workspace_bytes(...) + 8 KiB, 256 B aligned              */
  long long workspace_bytes;
} mtl_lane;
typedef struct mtl_meta_step_args {
  const float* theta;
  float* copy_grad;
  const float* pe_enc;
  const float* pe_dec;
  int n_tasks;
  const mtl_batch* train;   /* n_tasks training batches                                       */
  const mtl_batch* val;     /* the shared validation batch (transient_trainer.py:168-169)     */
  int n_lanes;
  const mtl_lane* lanes;
  mtl_meta_hparams hp;
  float* results;           /* n_tasks x 16 floats (device), may be NULL                      */
  unsigned long long* seed_slot;
  int use_graph;
} mtl_meta_step_args;
int mtl_meta_tasks(mtl_session* s, const mtl_meta_step_args* args, void* stream);
int mtl_graph_stats(const mtl_session* s, unsigned long long* captures, unsigned long long* replays);
/* Early exchange of the outer gradient (trainer/asr/transient_trainer.py:229,248 across GPUs).  copy_grad is accumulated
 * in two regions: A = the first mtl_region_a_floats() floats (every parameter except the VGG front-end, 98 % of the
 * arena) is complete when the LAST task's validation pass reaches its VGG backward; B = the VGG tail at the end.
 * mtl_stream_wait_region_a makes `stream` wait for region A of the most recent mtl_meta_tasks call (graph replay or
 * eager), so a collective on region A can run under the remaining convolution backward. */
long long mtl_region_a_floats(const mtl_session* s);
int mtl_stream_wait_region_a(mtl_session* s, void* stream);
/* transient_trainer.py:248-255: grad <- copy_grad; [clip]; theta <- Adam(theta, grad) with the outer optimizer's
 * betas / eps (torch.optim.Adam defaults 0.9, 0.999, 1e-8 at :109; a resumed optimizer keeps its own).
 * adam_state (device): int step, float step_size, float bc2_sqrt, pad (16 bytes). */
int mtl_meta_finish(float* theta, float* grad, const float* copy_grad, float* adam_m, float* adam_v,
                    void* adam_state, double meta_lr, double beta1, double beta2, double eps, int clip,
                    float max_norm, float* scratch1032, long long n, void* stream);

/* ------------------------------------------------------------------ flat-arena ops
 * models/asr/transformer.py:204-240 (zero/add/from_copy_grad), deepcopy(state_dict) /
 * load_state_dict (transient_trainer.py:160,237), torch.optim.SGD / Adam, clip_grad_norm_. */
int mtl_arena_zero(float* p, long long n, void* stream);
int mtl_arena_copy(float* dst, const float* src, long long n, void* stream);
int mtl_arena_axpy(float* y, const float* x, float a, long long n, void* stream);
int mtl_arena_sgd(float* p, const float* g, float lr, long long n, void* stream);
int mtl_arena_clip(float* g, long long n, float max_norm, float* scratch1032, void* stream);
int mtl_arena_adam(float* p, const float* g, float* m, float* v, void* adam_state, double lr, double b1,
                   double b2, double eps, long long n, void* stream);

/* ------------------------------------------------------------------ LSTM language-model meta-loop
 * Replaces lm/model/rnn_model.py:12-62 (RNNModel: Embedding -> Dropout -> nn.LSTM -> Dropout -> Linear) and the loop
 * body of lm/main_meta_transfer.py:277-372.  Parameters live in ONE flat fp32 arena in model.parameters() order
 * (encoder.weight, rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l<k>, decoder.weight, decoder.bias); every tensor start is
 * padded to 64 floats.  tokens are (T, B) int64 row-major (time-major, as LMDataset.get_batch returns them,
 * lm/util/data.py:36-44), targets (T*B) int64; hidden state h / c are (nlayers, B, nhid) fp32. */
typedef struct mtl_lm_cfg { int vocab, ninp, nhid, nlayers; } mtl_lm_cfg;
long long mtl_lm_param_floats(const mtl_lm_cfg* cfg);
int mtl_lm_param_count(const mtl_lm_cfg* cfg);
int mtl_lm_param_info(const mtl_lm_cfg* cfg, int idx, long long* offset_floats, long long* numel);
long long mtl_lm_workspace_bytes(const mtl_lm_cfg* cfg, int T, int B);
/* RNNModel.forward (+ nn.CrossEntropyLoss + .backward() when grad != NULL: grad += loss_scale * dCE/dtheta).
 * h0 / c0 NULL = zeros (init_hidden); hT / cT nullable, may alias h0 / c0.  loss_out8: device block {mean CE, n tokens,
 * n correct, ...}; logits_out: optional (T*B, vocab) copy of the decoder output.  targets NULL = logits only. */
int mtl_lm_pass(const mtl_lm_cfg* cfg, int gemm_mode, const float* theta, float* grad, const long long* tokens,
                const long long* targets, int T, int B, const float* h0, const float* c0, float* hT, float* cT,
                float dropout, unsigned long long seed, float loss_scale, void* workspace, long long workspace_bytes,
                float* loss_out8, float* logits_out, void* stream);
/* One meta-iteration (lm/main_meta_transfer.py:293-372), first-order: per task an inner SGD(lr / meta_lr_factor) step on
 * the clipped train gradient at a working copy of theta, then the val gradient at the adapted weights scaled by
 * task_weights[i] ((1-ratio)/2, (1-ratio)/2, ratio) accumulated in meta_grad; finally theta -= lr * clip(meta_grad).
 * The hidden state is threaded exactly as the script does: every train forward starts from the previous train forward's
 * (detached) state, the val forward from its own task's.  clip <= 0 disables clipping.  results: 16 floats per task
 * (train loss block, val loss block), device memory, nullable.  scratch1032: 1032 floats of device scratch.
 * seed_slot (nullable): device word that holds the step's dropout seed instead of `seed`; with it the enqueued work
 * depends on no per-step host value, so the call can be captured in a CUDA graph and replayed (no sync, no allocation). */
int mtl_lm_meta_step(const mtl_lm_cfg* cfg, int gemm_mode, float* theta, float* theta_work, float* grad, float* meta_grad,
                     float* hidden_h, float* hidden_c, int n_tasks, const long long* const* train_tokens,
                     const long long* const* train_targets, const long long* val_tokens, const long long* val_targets,
                     int T, int B, const float* task_weights, float lr, float meta_lr_factor, float clip, float dropout,
                     unsigned long long seed, const unsigned long long* seed_slot, void* workspace,
                     long long workspace_bytes, float* results, float* scratch1032, void* stream);

/* ------------------------------------------------------------------ single operators (unit-test surface) */
/* nn.Linear / its two backward contractions: C = epi(alpha*op(A)op(B)+bias)+beta*C, row-major.
 * transA: A stored (K,M); transB: B stored (N,K).  epi: 0 none, 1 relu, 2 relu-backward (aux). */
int mtl_gemm(int mode, int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, const float* bias, int epi,
             const float* aux, int split_k, void* stream);
/* The rank-r factorised nn.Linear pair of the attention blocks, linear_b(linear_a(x)) (modules/common_layers.py:250-257,
 * 287-289, 303), as ONE kernel -- forward (bwd = 0) or its input gradient (bwd = 1) -- for G <= 3 problems of equal shape
 * (q | k | v).  Pointer arrays hold G entries.  K slabs are merged by reduce-add: a [M, r] and y [M, N2] (row stride ldy)
 * must be zero (y may instead hold a sum to accumulate into) before the call.
 *   bwd = 0: x [M, K1] (ldx), w1 = linear_a.weight [r, K1], w2 = linear_b.weight [N2, r], bias[g] [N2] or null
 *   bwd = 1: x = dy [M, K1] (ldx), w1 = linear_b.weight [K1, r], w2 = linear_a.weight [r, N2], bias ignored
 * mode: MTL_GEMM_TC_TF32 or MTL_GEMM_TC_3XTF32.  ctas: CTA budget sizing the K split (0 = none). */
int mtl_lowrank_pair(int mode, int bwd, int G, int M, int K1, int r, int N2, const float* const* x, int ldx,
                     const float* const* w1, const float* const* w2, const float* const* bias, float* const* a,
                     float* const* y, int ldy, int ctas, void* stream);
/* measurement hook: `reps` back-to-back launches of the same GEMM (alpha = 1, no bias / epilogue) */
int mtl_gemm_repeat(int reps, int mode, int transA, int transB, int M, int N, int K, const float* A, int lda,
                    const float* B, int ldb, float beta, float* C, int ldc, int split_k, void* stream);
/* debug hook (MTL_GEMM_DBG=n): SM-cycle stamps of the last tcgen05 GEMM / conv launch into 160 host slots:
 * [0,32) phase stamps of CTA (0,0,0); [32,160) per-k-block pipeline stamps (4 per k-block) of CTA (n-1,0,0) */
int mtl_debug_gemm_stamps(long long* host160);
/* clock64 stamps of CTA (0,0) of the last short-sequence attention launches: [0,8) forward phases, [16,24) backward */
int mtl_debug_attn_stamps(long long* host32);
/* MTL_GEMM_DBG=99: %globaltimer (ns) at entry / exit of the first 256 CTAs of the last tcgen05 GEMM launch */
int mtl_debug_gemm_span(unsigned long long* host512);
/* debug: device pointers (inside the caller's workspace) of the VGG front-end intermediates of the last
 * mtl_asr_forward + mtl_asr_backward: c1, c2, p2, c3, c4, p4, feat, dfeat, dp4, dc4, dc3, dp2, dc2, dc1, dh, NULL */
int mtl_debug_pass_buffers(mtl_session* s, const float** out16);
/* LayerNorm(dropout(y)+res)*rowmask (+pe)  -- modules/common_layers.py:129-131,303-304 */
int mtl_ln_fwd(const float* y, const float* res, const float* gamma, const float* beta,
               const float* rowmask, const float* pe, int pe_period, float drop_p,
               unsigned long long seed, unsigned site, float* out, float* xhat, float* rstd, int M,
               int d, void* stream);
int mtl_ln_bwd(const float* dout, const float* xhat, const float* rstd, const float* gamma,
               const float* rowmask, float drop_p, unsigned long long seed, unsigned site, float* dy,
               float* dres, int dres_accumulate, float* dgamma, float* dbeta, int M, int d,
               void* stream);
/* ScaledDotProductAttention (modules/common_layers.py:317-331) on (B*T, H*dk) projections */
int mtl_attn_fwd(const float* q, const float* k, const float* v, const unsigned char* keypad, int B,
                 int H, int Tq, int Tk, int dk, int causal, float drop_p, unsigned long long seed,
                 unsigned site, float* o, float* lse, void* stream);
int mtl_attn_bwd(const float* q, const float* k, const float* v, const unsigned char* keypad,
                 const float* o, const float* lse, const float* d_o, int B, int H, int Tq, int Tk,
                 int dk, int causal, float drop_p, unsigned long long seed, unsigned site,
                 float* delta, float* dq, float* dkk, float* dv, void* stream);
/* F.cross_entropy(ignore_index=0) + topk(1)  -- utils/metrics.py:96-126, transformer.py:146 */
int mtl_ce_fwd(const float* logits, int ld, const int* gold, int M, int V, float smoothing,
               float* row_lse, float* row_loss, int* hyp, float* out8, void* stream);
int mtl_ce_bwd(const float* logits, int ld, const int* gold, const float* row_lse, const float* out8,
               float scale, float smoothing, float* dlogits, int M, int V, void* stream);
/* VGG front-end pieces (models/asr/transformer.py:47-59,136-138), NHWC */
int mtl_conv1_fwd(const float* x, const float* w, const float* b, float* out, int B, int F, int T,
                  int Cout, void* stream);
/* relu(conv3x3(x) + b): mode 0 = im2col (col: B*F*T*9*Cin floats) + fp32 GEMM; modes 1/2 = tcgen05 implicit GEMM
 * through 4-D TMA boxes (col unused, may be NULL).  wg: 2*Cout*9*Cin floats of scratch (GEMM-layout weights; in
 * 3xTF32 their tf32 hi and lo halves). */
int mtl_conv3x3_relu_fwd(int mode, const float* x, const float* w, const float* b, float* col,
                         float* wg, float* out, int B, int F, int T, int Cin, int Cout, void* stream);
/* the same with the following MaxPool2d(2, 2) (models/asr/transformer.py:51,58) written from the convolution's epilogue:
 * out [B,F,T,Cout] as above (the backward needs it) and pool_out [B,F/2,T/2,Cout]; tensor-core engines, Cout % 32 == 0 */
int mtl_conv3x3_relu_pool_fwd(int mode, const float* x, const float* w, const float* b, float* wg, float* out,
                              float* pool_out, int B, int F, int T, int Cin, int Cout, void* stream);
/* Backward of y = conv3x3(x) + b given dy (gradient w.r.t. the pre-ReLU output): dw += , db += ,
 * dx = (dgrad) masked by relu_aux > 0 when relu_aux != NULL (dx may be NULL). */
long long mtl_conv3x3_bwd_scratch_floats(int mode, int B, int F, int T, int Cin, int Cout);
int mtl_conv3x3_bwd(int mode, const float* x, const float* w, const float* dy, const float* relu_aux,
                    float* dw, float* db, float* dx, float* scratch, int B, int F, int T, int Cin, int Cout,
                    void* stream);
/* conv.0 weight / bias gradient (models/asr/transformer.py:48, Cin = 1): dw (Cout,1,3,3) and db accumulate */
int mtl_conv1_wgrad(const float* x, const float* dout, float* dw, float* db, int B, int F, int T, int Cout, void* stream);
/* (B,F4,T4,C) NHWC <-> (B,T4,C*F4) encoder features (models/asr/transformer.py:133-138); backward != 0: the inverse map */
int mtl_feat_transpose(const float* src, float* dst, int B, int F4, int T4, int C, int backward, void* stream);
/* decoder input embedding + positional encoding + dropout (modules/decoder.py:96) and its scatter-add gradient;
 * out (nullable): forward; dout + dE (nullable): backward into dE (vocab x d, accumulated; PAD row 0 skipped) */
int mtl_embed(const int* tok, const float* E, const float* pe, float drop_p, unsigned long long seed, unsigned site,
              float* out, const float* dout, float* dE, int B, int n, int d, void* stream);
int mtl_maxpool2_fwd(const float* x, float* out, int B, int F, int T, int C, void* stream);
int mtl_maxpool2_relu_bwd(const float* x, const float* dpool, float* dx, int B, int F, int T, int C,
                          void* stream);
int mtl_dec_preprocess(const long long* trg, int B, int L, int n, int* seq_in, int* seq_out,
                       float* rowmask, unsigned char* keypad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTL_B200_H */
