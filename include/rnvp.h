/*
 * rnvp.h -- C-ABI of the B200-native RealNVP coupling-stack hot path.
 *
 * The reference (alisher-turubayev/dl-normalizing-flows) has no FFI: its
 * boundary is the Python class API in flow_realnvp.py / modules_realnvp.py /
 * utils.py, all of whose arithmetic runs inside torch.  This header is the
 * boundary a maintainer binds (ctypes) to move that
 * arithmetic onto sm_100a kernels.  Each entry point names the reference code
 * it replaces (file:line into the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, a negative rnvp_status otherwise;
 *     rnvp_last_error() gives the message for the calling thread.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; nothing is cached per thread,
 *     so autograd's backward thread may call in (SURVEY.md 8b "Threading").
 *   - the library allocates no persistent device memory except the small
 *     per-plan tables created by rnvp_plan_create / rnvp_plan_bind; all
 *     activations, saved tensors and scratch live in a caller-owned workspace.
 *   - activations inside the library are NHWC fp32 with the channel stride of
 *     conv operands padded to a multiple of 32; the public entry points take and
 *     return the reference's NCHW fp32 tensors.
 */
#ifndef RNVP_H_
#define RNVP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RNVP_OK = 0,
  RNVP_ERR_INVALID = -1,     /* bad argument / unsupported configuration   */
  RNVP_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed        */
  RNVP_ERR_WORKSPACE = -3,   /* workspace too small                        */
  RNVP_ERR_STATE = -4,       /* call order violated (e.g. backward w/o fwd)*/
  RNVP_ERR_NCCL = -5
} rnvp_status;

/* conv arithmetic tier (SURVEY.md 7 "precision tiers") */
typedef enum {
  RNVP_MATH_FP32 = 0,        /* CUDA-core fp32 FMA implicit GEMM (1e-5 tier, reconstruction gate) */
  RNVP_MATH_TF32 = 1,        /* tcgen05 kind::tf32, fp32 accumulate in TMEM (1e-3 tier, throughput) */
  RNVP_MATH_TF32X3 = 2       /* the same tensor-core kernels with split operands ("3xTF32": hi*hi + lo*hi + hi*lo, three
                                MMAs per K step, nothing rounded): fp32-class accuracy (1e-5 tier) on tcgen05 */
} rnvp_math;

/* RealNVP(channels, image_size, prior, hps): flow_realnvp.py:36-95, utils.py:78-93.
 * num_scales = 5 reproduces the in-tree model; other values give the same
 * doubling rule truncated (BASELINE config 3).                              */
typedef struct {
  int32_t channels;
  int32_t image_size;
  int32_t base_dim;
  int32_t res_blocks;
  int32_t num_scales;
  float prior_loc;           /* torch.distributions.Normal(loc, scale), train.py:109 */
  float prior_scale;
} rnvp_config;

typedef struct rnvp_plan rnvp_plan;

const char* rnvp_last_error(void);
const char* rnvp_version(void);
int rnvp_device_ok(void);                      /* 0 when cuda:current is sm_100 */
/* number of kernels this library has launched in the process so far (bench.py's gpu_launches) */
unsigned long long rnvp_launch_count(void);
/* Measurement hooks: with profiling on, every conv / dgrad / wgrad / batch-norm launch is bracketed by
 * CUDA events on its own stream.  rnvp_prof_collect synchronises, sums the intervals per kernel class
 * into rows of 7 doubles (kind 0 conv 1 dgrad 2 wgrad 3 bn 4 bn-bwd, S, taps, cin, cout, launches,
 * total_ms), clears the records and returns the row count.  While profiling is on, the backward pass
 * keeps every kernel on the caller's stream (no side-stream overlap), so an interval is one kernel.   */
int rnvp_prof_enable(int on);
int rnvp_prof_collect(double* rows_host, int max_rows);

/* ---- plan ------------------------------------------------------------- */
int rnvp_plan_create(const rnvp_config* cfg, rnvp_plan** out);
/* a plan holding ONE stand-alone coupling module, as constructed by
 * CheckerboardAffineCoupling(in_out_dim=C, mid_dim=D, size=S, mask_config, hps)   (kind 0,
 * modules_realnvp.py:240) or ChannelwiseAffineCoupling(in_out_dim=C, mid_dim=D, mask_config, hps)
 * (kind 1, modules_realnvp.py:305); only the rnvp_coupling_* entry points accept it.           */
int rnvp_plan_create_single(int kind, int C, int S, int D, int mask_cfg, int res_blocks, rnvp_plan** out);
int rnvp_plan_destroy(rnvp_plan* plan);
int rnvp_plan_num_couplings(const rnvp_plan* plan);
/* number of pointer slots per coupling in the parameter table (see
 * rnvp_plan_slot_name) and in total                                        */
int rnvp_plan_slots_per_coupling(const rnvp_plan* plan);
/* name of slot `i` relative to its coupling module, e.g.
 * "block.1.core_block.0.res_block.3.conv.weight_v" -- the reference's
 * state-dict key suffix.  Returns NULL when out of range.    */
const char* rnvp_plan_slot_name(const rnvp_plan* plan, int slot);
/* coupling `i`: name ("s1_ckbd.0"), kind (0 ckbd / 1 chan), C, S, D, mask cfg */
int rnvp_plan_coupling_info(const rnvp_plan* plan, int i, char* name, int name_len,
                            int* kind, int* C, int* S, int* D, int* mask_cfg);

/* Bind parameter / gradient device pointers.  `params` and `grads` are HOST
 * arrays of num_couplings*slots_per_coupling device pointers; grads entries
 * are NULL for buffers and frozen parameters (weight_g of scale=False convs,
 * modules_realnvp.py:57-59).  Re-bind whenever a pointer changes.  Not
 * capturable in a CUDA graph (it uploads tables).                          */
int rnvp_plan_bind(rnvp_plan* plan, void* const* params_host, void* const* grads_host, void* stream);

/* bytes of workspace the calls below need for batch B.
 * mode: 0 = inference forward / inverse,
 *       1 = training forward + backward, lean (normalised activations are recomputed in the backward),
 *       2 = training, fast (they are kept: ~1.9x the activation memory, 13 fewer passes per coupling) */
size_t rnvp_plan_workspace_bytes(const rnvp_plan* plan, int batch, int mode);

int rnvp_plan_set_math(rnvp_plan* plan, int math /* rnvp_math */);
/* Counter bumped by every rnvp_flow_forward / rnvp_flow_inverse / rnvp_coupling_forward / _inverse call: the
 * activations a backward call needs live in the caller's workspace and belong to the LAST forward.  A binding that
 * keeps several forwards alive (autograd) records the value after its forward and refuses a backward whose
 * record is stale (rnvp_engine.FlowLogProb does).                                                          */
unsigned long long rnvp_plan_forward_generation(const rnvp_plan* plan);

/* ---- the flow (flow_realnvp.py:196-370) -------------------------------- */
/* log_prob / forward (flow_realnvp.py:329-340, 354-370).
 *   x_nchw       (B,C,H,W) logit-space input
 *   ll           (B)  log prior + log det                       [out]
 *   logdet       (B)  log det only, may be NULL                 [out]
 *   z_nchw       (B,C,H,W) latent, may be NULL                  [out]
 *   weight_scale (1)  sum p^2 over trainable weight_g / scale, may be NULL [out]
 *   training     1 or 2: batch statistics, running stats updated, tensors saved in
 *                   the workspace (laid out for that mode, see rnvp_plan_workspace_bytes) for
 *                   rnvp_flow_backward; 0: running statistics                 */
int rnvp_flow_forward(rnvp_plan* plan, const float* x_nchw, float* ll, float* logdet, float* z_nchw,
                      float* weight_scale, int batch, int training,
                      void* workspace, size_t workspace_bytes, void* stream);

/* backward of rnvp_flow_forward(training=1) (autograd of flow_realnvp.py:252-340
 * and of the modules).  dll (B) is dLoss/dll, dweight_scale a
 * DEVICE scalar holding dLoss/dweight_scale (NULL = 0).  Parameter gradients are ADDED
 * into the bound grads; dx_nchw (may be NULL) receives dLoss/dx.            */
int rnvp_flow_backward(rnvp_plan* plan, const float* dll, const float* dweight_scale, float* dx_nchw,
                       int batch, void* workspace, size_t workspace_bytes, void* stream);

/* g: z -> x (flow_realnvp.py:196-249).  training selects batch vs running
 * statistics for in_bn exactly like nn.BatchNorm2d; the un-normalisation of
 * out_bn always uses running statistics (modules_realnvp.py:285-291).       */
int rnvp_flow_inverse(rnvp_plan* plan, const float* z_nchw, float* x_nchw, int batch, int training,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- one coupling (modules_realnvp.py:264-302, 324-370) ----------------- */
/* forward(x, reverse=False): y and the full log_diag_J tensor (both NCHW,
 * (B,C,S,S)); logdet (B) optional per-sample sum.  With training=1 the
 * tensors needed by rnvp_coupling_backward stay in the workspace.            */
int rnvp_coupling_forward(rnvp_plan* plan, int coupling, const float* x_nchw, float* y_nchw,
                          float* logJ_nchw, int batch, int training,
                          void* workspace, size_t workspace_bytes, void* stream);
/* forward(x, reverse=True): x and, optionally, the tensor the reference returns beside it -- log_rescale
 * (modules_realnvp.py:283, 302; zeros on the untouched half of a channelwise coupling), logJ_nchw may be NULL */
int rnvp_coupling_inverse(rnvp_plan* plan, int coupling, const float* y_nchw, float* x_nchw, float* logJ_nchw,
                          int batch, int training,
                          void* workspace, size_t workspace_bytes, void* stream);
/* VJP of rnvp_coupling_forward(training=1): dy, dlogJ (B,C,S,S) NCHW upstream
 * grads -> dx; parameter grads are added into the bound grads.
 * dlogJ must be constant within a sample (it is dLoss/dll broadcast) -- the
 * per-sample value is read from element 0 of each sample.                    */
int rnvp_coupling_backward(rnvp_plan* plan, int coupling, const float* dy_nchw, const float* dlogJ_nchw,
                           float* dx_nchw, int batch,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- logit dequantisation (utils.py:33-72) ------------------------------ */
/* forward: x in [0,1] (fp32, n_per_sample = C*H*W), noise in [0,1) or NULL
 * (then Philox(seed, offset) noise is drawn in-kernel, utils.py:47);
 * y and per-sample log-det out.                                             */
int rnvp_logit_forward(const float* x, const float* noise, float* y, float* logdet,
                       int batch, int n_per_sample, float constraint,
                       uint64_t seed, uint64_t offset, void* stream);
/* same from uint8 pixels (value/255 is what ToTensor yields, train.py:69) */
int rnvp_logit_forward_u8(const uint8_t* x, const float* noise, float* y, float* logdet,
                          int batch, int n_per_sample, float constraint,
                          uint64_t seed, uint64_t offset, void* stream);
/* reverse=True branch, utils.py:34-42 */
int rnvp_logit_inverse(const float* y, float* x, size_t n, float constraint, void* stream);

/* ---- layout transforms (flow_realnvp.py:121-193), NCHW in / NCHW out ---- */
int rnvp_squeeze(const float* x, float* y, int B, int C, int H, int W, void* stream);
int rnvp_undo_squeeze(const float* x, float* y, int B, int C, int H, int W, void* stream);
int rnvp_factor_out(const float* x, float* on, float* off, int B, int C, int H, int W, void* stream);
int rnvp_restore(const float* on, const float* off, float* x, int B, int C, int H, int W, void* stream);

/* ---- building blocks exposed for per-op parity tests --------------------- */
/* w = g * v / ||v|| (modules_realnvp.py:53-56) into the padded GEMM layouts:
 *   wf [taps][pad16(Cout)][pad32(Cin)]   forward operand
 *   wb [taps][pad16(Cin)][pad32(Cout)]   dgrad operand (taps flipped, transposed); may be NULL */
int rnvp_weightnorm_forward(const float* v, const float* g, float* wf, float* wb,
                            int cout, int cin, int ksize, void* stream);
/* dwf (same layout as wf) -> dv (+=), dg (+=, may be NULL when g is frozen) */
int rnvp_weightnorm_backward(const float* v, const float* g, const float* dwf, float* dv, float* dg,
                             int cout, int cin, int ksize, void* stream);
/* stride-1 "same" conv as implicit GEMM over NHWC:
 *   y[p, n] = sum_tap sum_k x[p + tap, k] * wf[tap][n][k] (+ bias[n]) (+ res[p, n])
 * x [B,S,S,kpad], wf [taps][npad][kpad], y/res row stride ldy.  stats (2*n
 * doubles, optional) accumulates per-channel sum and sum of squares of y.
 * math = RNVP_MATH_TF32X3: wf holds the weights followed by their lo halves,
 * [w | w - trunc_tf32(w)], taps*npad*kpad floats each (inside the flow entry
 * points the weight-norm kernel writes both); rnvp_conv_wgrad needs no extra
 * operand in that tier (it splits x and dy in shared memory).                 */
int rnvp_conv_forward(const float* x, const float* wf, const float* bias, const float* res, float* y,
                      double* stats, int B, int S, int kpad, int n, int npad, int ksize, int ldy,
                      int math, void* stream);
/* dwf[tap][n][k] += sum_p dy[p, n] * x[p + tap, k];  dbias[n] += sum_p dy[p, n] (optional) */
int rnvp_conv_wgrad(const float* x, const float* dy, float* dwf, float* dbias,
                    int B, int S, int kpad, int n, int npad, int ksize, int lddy,
                    int math, void* stream);

/* BatchNorm2d -> ReLU -> conv as ONE tensor-core kernel (modules_realnvp.py:83-97, 139-143; tier RNVP_MATH_TF32):
 *   y = conv(tf32(relu(x_raw * scale + shift)), wf) (+ bias) (+ res), zero padding applied AFTER the activation;
 * x_raw [B,S,S,kpad] is the raw pre-BN tensor (bn_C real channels), the BN is applied to the operand tiles in
 * shared memory, so relu(bn(x)) never exists in HBM.  bn_mode / sums / count / gamma / beta / running
 * statistics / save as in rnvp_bn_relu_forward (mode 1 also saves the coefficients and updates the running
 * statistics).  round_out: y is rounded to nearest TF32 (it is read raw by another conv MMA).                 */
int rnvp_conv_forward_bn(const float* x_raw, const float* wf, const float* bias, const float* res, float* y,
                         double* stats, int B, int S, int kpad, int n, int npad, int ksize, int ldy, int bn_mode,
                         int bn_C, const double* bn_sums, double bn_count, const float* gamma, const float* beta,
                         float* run_mean, float* run_var, float* save, int round_out, void* stream);
/* weight gradient of that conv from the RAW x: dwf[tap][n][k] += sum_p dy[p, n] * tf32(relu(bn(x_raw)))[p + tap, k]
 * with the coefficients (mean, rstd, scale, shift)[bn_C] the forward saved.                                    */
int rnvp_conv_wgrad_bn(const float* x_raw, const float* dy, float* dwf, float* dbias, int B, int S, int kpad, int n,
                       int npad, int ksize, int lddy, const float* bn_save, int bn_C, void* stream);

/* BatchNorm2d + ReLU (modules_realnvp.py:83-85, 139-141) on a [P, ld] NHWC trunk tensor with C real channels:
 *   mode 1 (training): batch statistics from `sums` (2C doubles: per-channel sum, sum of squares over `count`
 *           values, as accumulated by a conv epilogue's `stats`); save[4C] = (mean, rstd, scale, shift); running
 *           statistics updated (momentum 0.1, unbiased variance)
 *   mode 0 (eval): running statistics;   mode 2: coefficients read back from `save`
 * tf32_round: h is rounded to nearest TF32 (it is a tensor-core operand).                                   */
int rnvp_bn_relu_forward(const float* x, float* h, int P, int C, int ld, const double* sums, double count,
                         const float* gamma, const float* beta, float* run_mean, float* run_var, float* save,
                         int mode, int tf32_round, void* stream);
/* The fused backward unit of conv -> BN -> ReLU (autograd of modules_realnvp.py:83-97), tensor-core tier:
 *   gm[p, n] = (sum_tap sum_k dy[p + tap, k] * wb[tap][n][k]) * 1[bn_x[p, n] * scale[n] + shift[n] > 0]
 *   sums2[0:n] += sum_p gm,   sums2[n:2n] += sum_p gm * bn_x        (the two BN-backward reductions)
 * bn_x = the raw pre-BN activations, bn_save = (mean, rstd, scale, shift)[n] of that BN.                     */
int rnvp_conv_dgrad_bn(const float* dy, const float* wb, const float* bn_x, const float* bn_save, float* gm,
                       double* sums2, int B, int S, int kpad, int n, int npad, int ksize, int ldy, void* stream);
/* dx = gamma * rstd * (gm - mean(gm) - xhat * mean(gm * xhat)) (+ add); dgamma / dbeta are added to.
 * raw_x_sums = 1 when sums2[C:2C] holds sum gm * x (rnvp_conv_dgrad_bn) instead of sum gm * xhat.           */
int rnvp_bn_backward_apply(const float* gm, const float* x, float* dx, const float* add, int P, int C, int ld,
                           const float* save, const double* sums2, double count, const float* gamma, float* dgamma,
                           float* dbeta, int raw_x_sums, int tf32_round, void* stream);

/* ---- data parallel (absent from the reference) ------------- */
/* 128-byte NCCL unique id, created on rank 0 and shipped by the caller      */
int rnvp_dp_unique_id(void* id128_host);
int rnvp_dp_init(rnvp_plan* plan, const void* id128_host, int rank, int world);
/* Tell the plan where the gradients live so that rnvp_flow_backward can all-reduce (average) them in
 * buckets while the rest of the backward runs: `flat` is the contiguous fp32 buffer that the bound
 * grads point into (couplings in forward order), offsets_host[i] .. offsets_host[i+1] the element
 * range of coupling i (num_couplings + 1 entries).  Buckets are whole couplings, at least
 * bucket_elems elements (0 = default 1 Mi).                                                  */
int rnvp_dp_set_grad_layout(rnvp_plan* plan, float* flat, const int64_t* offsets_host, int64_t bucket_elems);
/* One-shot NVLink exchange for the batch-norm statistic vectors (optional; without it they go through
 * ncclAllReduce).  Every rank allocates an inbox of `cap_doubles` per (peer, slot) and returns its 64-byte
 * CUDA-IPC handle; after an all-gather of the handles (rank order, world * 64 bytes) rnvp_dp_xchg_open maps
 * the peers' inboxes.  From then on every statistic vector of at most `cap_doubles` is reduced by one
 * single-CTA kernel that pushes it into all inboxes and adds the arrivals up in rank order.
 * rnvp_dp_xchg_errors returns the exchange's STICKY error word (mapped host memory, no synchronisation): 0 healthy,
 * 1 a peer never arrived (~2 min), 2 the ranks issued different call sequences, 3 the ranks' local batch sizes
 * differ (every training forward reduces (B, B^2) first); once set, every compute entry point of the plan fails
 * with RNVP_ERR_STATE and the affected reductions return NaN.                                            */
int rnvp_dp_xchg_alloc(rnvp_plan* plan, int cap_doubles, void* ipc_handle64_host);
int rnvp_dp_xchg_open(rnvp_plan* plan, const void* all_handles_host);
int rnvp_dp_xchg_errors(rnvp_plan* plan);
/* sum-all-reduce `n` doubles in place on `stream` exactly as the batch-norm statistics are (test hook)  */
int rnvp_dp_allreduce_stats(rnvp_plan* plan, double* buf, size_t n, void* stream);
int rnvp_dp_finalize(rnvp_plan* plan);
/* sum-all-reduce `n` floats in place on `stream` (gradient buckets)         */
int rnvp_dp_allreduce(rnvp_plan* plan, float* buf, size_t n, void* stream);

/* ---- optimizer step ------------------------------------- *
 * torch.optim.Adam(model.parameters(), lr, weight_decay) -- train.py:134, stepped at train.py:200 -- as one
 * launch over all trainable tensors: coupled L2 weight decay, bias-corrected moments, optional clearing of
 * the gradients.  Parameters are the caller's tensors (`params_host[i]`, `sizes_host[i]` floats each);
 * gradient and the two moment buffers share one flat layout, segment i starting at `offsets_host[i]`.   */
typedef struct rnvp_adam rnvp_adam;
int rnvp_adam_create(void* const* params_host, const int64_t* offsets_host, const int64_t* sizes_host, int n,
                     rnvp_adam** out);
int rnvp_adam_destroy(rnvp_adam* a);
/* `step` counts from 1 (the value torch keeps in state['step'] AFTER the update).                       */
int rnvp_adam_step(rnvp_adam* a, float* flat_grad, float* exp_avg, float* exp_avg_sq, int64_t flat_len, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int64_t step, int zero_grad,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* RNVP_H_ */
