// Launcher prototypes shared between the kernel translation units and the runtime.
// Everything here is internal; the public surface is include/rnvp.h.
#pragma once
#include "common.cuh"

namespace rnvp {

// 0 / 1 alternating per call (see elementwise.cu): the direction in which the next bulk kernel sweeps
int next_sweep_dir();

// ---- layout (flow_realnvp.py:121-193), all NHWC ---------------------------------
enum PermMode {
  PERM_SQUEEZE = 0,      // hi [B,2s,2s,C]            -> sq [B,s,s,4C]
  PERM_UNDO_SQUEEZE,     // sq                         -> hi
  PERM_FACTOR_OUT,       // hi                         -> on,off [B,s,s,2C]
  PERM_RESTORE,          // on,off                     -> hi
  PERM_UNSQ_FACTOR,      // sq                         -> on,off   (undo_squeeze o factor_out)
  PERM_FACTOR_SQ         // on,off                     -> sq       (restore o squeeze)
};
// s = low-res side, C = channels of the high-res tensor
int k_permute(PermMode mode, const float* hi, const float* sq, const float* on, const float* off,
              float* hi_o, float* sq_o, float* on_o, float* off_o, int B, int s, int C, cudaStream_t st);
int k_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, cudaStream_t st);
int k_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, cudaStream_t st);

// ---- logit (utils.py:33-72) -----------------------------------------------------
int k_logit_fwd(const float* xf, const uint8_t* xu8, const float* noise, float* y, float* logdet,
                int B, int n, float constraint, uint64_t seed, uint64_t offset, cudaStream_t st);
int k_logit_inv(const float* y, float* x, size_t n, float constraint, cudaStream_t st);

// ---- batch norm on [P,ld] trunk tensors (C real channels, C % 4 == 0, ld = pad32(C)) ---
// h = relu(gamma*(x-mean)*rstd+beta).
//   mode 1 (training): (sum,sumsq) over `count` values come from `sums` (2C doubles);
//           save[4C] = mean,rstd,scale,shift; running stats updated by block 0
//   mode 0 (eval): running statistics
//   mode 2 (recompute in backward): scale/shift read back from `save`
// xg: data parallel -- the cross-rank reduction of `sums` happens inside this kernel (common.cuh, DpXchg)
int k_bn_relu(const float* x, float* h, int P, int C, int ld, const double* sums, double count,
              const float* gamma, const float* beta, float* run_mean, float* run_var,
              float* save, int mode, int tf32_round, cudaStream_t st, DpXchg xg = DpXchg());
// gm = g * 1[h>0] written to gm_out; sums2[0:C] += sum gm, sums2[C:2C] += sum gm*xhat
int k_bn_bwd_reduce(const float* g, const float* x, float* gm_out, int P, int C, int ld,
                    const float* save, double* sums2, cudaStream_t st);
// dx = gamma*rstd*(gm - m1 - xhat*m2) (+ add, which may alias dx); block 0 adds dgamma/dbeta.
// raw_x_sums: sums2[C:2C] holds sum gm*x (fused dgrad epilogue) instead of sum gm*xhat
// tf32_round: dx is the dy operand of a dgrad / wgrad MMA -> round to nearest TF32
int k_bn_bwd_apply(const float* gm, const float* x, float* dx, const float* add, int P, int C, int ld,
                   const float* save, const double* sums2, double count, const float* gamma,
                   float* dgamma, float* dbeta, float inv_world, int raw_x_sums, int tf32_round, cudaStream_t st,
                   DpXchg xg = DpXchg());

// eval mode: (scale, shift) of every batch norm of a plan from its running statistics, one launch for all of them;
// job i writes save[2C:3C] = gamma * rsqrt(rv + eps), save[3C:4C] = beta - rm * scale
struct BnEvalJob { const float* gamma; const float* beta; const float* rm; const float* rv; size_t save_off; int C; };
int k_bn_eval_coefs(const BnEvalJob* jobs_dev, int njobs, int max_c, float* save_base, cudaStream_t st);

// ---- coupling pieces (modules_realnvp.py:264-302, 324-370) ------------------------
int k_cpl_in_stats(const float* x, CplGeom g, double* sums, cudaStream_t st);
int k_cpl_in_build(const float* x, CplGeom g, const double* sums, double count, const float* gamma,
                   const float* beta, float* run_mean, float* run_var, float* save, int training,
                   float* h0, int tf32_round, cudaStream_t st);
int k_cpl_fwd_a(const float* x, const float* stt, CplGeom g, const float* scale, const float* sshift,
                float* xprime, double* sums, double* logdet_acc, int training, cudaStream_t st);
int k_cpl_fwd_b(const float* xprime, const float* x, const float* stt, CplGeom g, const double* sums,
                double count, float* run_mean, float* run_var, float* save, int training,
                const float* scale, const float* sshift, float* y, float* logJ, double* logdet_acc,
                cudaStream_t st);
int k_cpl_inv(const float* y, const float* stt, CplGeom g, const float* run_mean, const float* run_var,
              const float* scale, const float* sshift, float* x, float* logJ, cudaStream_t st);
// sums2: [0:cio] sum g, [cio:2cio] sum g*xhat, [2cio] K = sum_p dll_b(p)*keep_p
int k_cpl_bwd_a(const float* dy, const float* xprime, CplGeom g, const float* save, const float* dll,
                double* sums2, cudaStream_t st);
int k_cpl_bwd_b(const float* dy, const float* xprime, const float* x, const float* stt, CplGeom g,
                const float* save, const double* sums2, double count, const float* dll,
                const float* scale, const float* sshift, float* dst, float* dxdir,
                float* dscale, float* dsshift, int tf32_round, cudaStream_t st);
int k_cpl_in_bwd_a(const float* dh0, const float* x, CplGeom g, const float* save, double* sums3,
                   cudaStream_t st);
int k_cpl_in_bwd_b(const float* dh0, const float* x, const float* dxdir, const float* dy, CplGeom g,
                   const float* save, const double* sums3, double count, const float* gamma,
                   float* dgamma, float* dbeta, float* dx, float inv_world, cudaStream_t st);

// ---- prior, per-sample sums ---------------------------------------------------------
int k_prior_ll(const float* z, int B, int n, float loc, float scale, double* acc, cudaStream_t st);
int k_prior_grad(const float* z, const float* dll, float* dz, int accumulate, int B, int n, float loc,
                 float scale, cudaStream_t st);
int k_finalize_ll(const double* logdet_acc, const double* prior_acc, float* ll, float* logdet, int B,
                  cudaStream_t st);
int k_add(float* dst, const float* src, size_t n, cudaStream_t st);
// out[b] = in[b*n]  (per-sample scalar carried by a broadcast tensor)
int k_gather_first(const float* in, float* out, int B, int n, cudaStream_t st);

// ---- weight norm (modules_realnvp.py:53-59), batched through a job table ----------
// Offsets are in floats relative to the weight arena (wf/wb) and the wgrad scratch (dw), both of
// which live in the caller's workspace, so the table survives a workspace move.
struct WnJob {
  const float* v;      // (cout, cin, k, k)
  const float* g;      // (cout)
  float* dv;           // grad of v (+=) or null
  float* dg;           // grad of g (+=) or null when frozen
  size_t wf_off;       // [taps][npad_f][kpad_f]
  size_t wb_off;       // [taps][npad_b][kpad_b]
  size_t dw_off;       // wgrad output, same layout as wf
  int cout, cin, taps;
  int npad_f, kpad_f, npad_b, kpad_b;
  // Row strides (floats) of wf and of the wgrad output; 0 = kpad_f.  The skip convs of a coupling (in_skip,
  // core_skips) are column blocks of ONE [npad][(R+1)*ldD] matrix: their sum over the trunk tensors is a single
  // GEMM with the K dimension concatenated (runtime.cu, skip_fused).
  int ld_f, ld_dw;
  // bias gradient that the fused skip wgrad left in the wgrad scratch (it is the same vector for every skip conv):
  // dbias[co] += dwbase[dbias_src_off + co] when dbias != null
  float* dbias;
  size_t dbias_src_off;
};
// writes every element of wf and wb (zero in the padding), so the arena needs no clearing
// lo_delta != 0 (3xTF32 tier): also write lo = w - trunc_tf32(w) of both layouts lo_delta floats further on
int k_weightnorm_fwd(const WnJob* jobs_dev, int njobs, int max_cout, float* wbase, int tf32_round,
                     cudaStream_t st, size_t lo_delta = 0);
int k_weightnorm_bwd(const WnJob* jobs_dev, int njobs, int max_cout, const float* wbase,
                     const float* dwbase, cudaStream_t st);

// out[wbase + out_off + c] = sum_i b[i][c]: the bias of the fused skip conv (one launch for all couplings)
constexpr int kMaxSkipConvs = 17;
struct BiasSumJob { const float* b[kMaxSkipConvs]; int n; int C; size_t out_off; };
int k_bias_sum(const BiasSumJob* jobs_dev, int njobs, float* wbase, cudaStream_t st);

struct Seg { const float* p; float* g; int n; };
int k_sumsq(const Seg* segs_dev, int nsegs, double* acc, cudaStream_t st);      // acc += sum p^2
int k_sumsq_finish(const double* acc, float* out, cudaStream_t st);
int k_sumsq_bwd(const Seg* segs_dev, int nsegs, const float* dws_dev, cudaStream_t st);   // g += 2*p*dws

// ---- data parallel state owned by a plan (dp.cu) -------------------------------------
struct DpState {
  void* comm = nullptr;              // ncclComm_t
  int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
  float* flat = nullptr;             // flat gradient buffer (all couplings, forward order)
  int64_t* off = nullptr;            // host array: n_cpl + 1 element offsets into `flat`
  int n_cpl = 0;
  int64_t pending_end = -1;          // end of the not-yet-reduced contiguous range
  int64_t bucket_elems = 1 << 20;    // reduce when at least this many elements are pending
  bool launched = false;
  // one-shot NVLink exchange of the small batch-norm statistic vectors (dp.cu): every rank owns an inbox
  // [world][2 slots][cap] doubles + [world] flags in device memory that its peers map through CUDA IPC
  void* xchg_local = nullptr;        // this rank's inbox
  void** xchg_peers_dev = nullptr;   // device array: inbox base of every rank (own entry = xchg_local)
  void* xchg_peer_host[16] = {};     // opened peer mappings (to close them)
  int xchg_cap = 0;                  // doubles per slot
  unsigned long long xchg_seq = 0;   // exchanges issued so far (all ranks call in the same order)
  int* xchg_err = nullptr;           // device alias of the sticky error word (0 healthy; see rnvp_dp_xchg_errors)
  int* xchg_err_host = nullptr;      // the word itself, in mapped host memory
  bool xchg_ready = false;
};
int dp_allreduce_doubles(DpState* dp, double* buf, size_t n, cudaStream_t st);
// descriptor for a reduction of n doubles folded into the consumer kernel (takes the next sequence number), or a
// disabled descriptor after having reduced `buf` in place the stand-alone way (exchange closed, n too large,
// RNVP_DP_FUSED=0)
int dp_fused_exchange(DpState* dp, double* buf, size_t n, cudaStream_t st, DpXchg* out);
// reduce (B, B^2) over the ranks and flag unequal local batches in the sticky error word (scratch2: 2 device doubles)
int dp_check_equal_batches(DpState* dp, int batch, double* scratch2, cudaStream_t st);
int dp_sticky_error(const DpState* dp);   // current value of the sticky error word (host read, no synchronisation)
// called after the backward of coupling `ci` was enqueued on `main` (couplings finish last -> first)
int dp_coupling_done(DpState* dp, int ci, cudaStream_t main);
// make `main` wait for every gradient bucket launched during this backward
int dp_join(DpState* dp, cudaStream_t main);

// ---- convolutions -------------------------------------------------------------------
// BatchNorm2d + ReLU folded into a tensor-core kernel's operand path: the operand in HBM is the raw pre-BN tensor
struct BnPrologue {
  int mode;            // 1 batch statistics (sums), 0 running statistics, 2 coefficients read back from `save`
  int C;               // real channels
  const double* sums;  // [2C] sum, sum of squares over `count` values (mode 1)
  double count;
  const float* gamma;
  const float* beta;
  float* run_mean;     // updated in mode 1
  float* run_var;
  float* save;         // [4C] mean, rstd, scale, shift: written in mode 1, read in mode 2
  DpXchg xg;           // data parallel, mode 1: reduce `sums` over the ranks inside the kernel's prologue
};
// The affine coupling fused into the epilogue of the s/t network's out conv (modules_realnvp.py:277-301, 339-361):
// the accumulator row of a pixel holds (t | l); the epilogue forms s = (scale*tanh(l)+shift)*(1-m), t*(1-m) and
//   mode 1 (training forward): x' = x*exp(s)+t -> out [P,cio]; sums += (sum x', sum x'^2); logdet_acc[b] += sum s
//   mode 2 (eval forward)    : y = out_bn_running(x') (masked) -> out [P,C]; logdet_acc[b] += sum (s - hl*(1-m))
//   mode 3 (reverse=True)    : x = (y*exp(hl*(1-m)) + rm*(1-m) - t)*exp(-s) -> out [P,C]
// with hl = 0.5*log(running_var + 1e-5).  The per-sample log-det is reduced with warp shuffles.
struct CplEpilogue {
  int mode = 0;
  int store_st = 0;            // also write the raw (t | l) tensor (the backward pass reads l)
  CplGeom g;
  const float* x = nullptr;    // coupling input (forward) / output (reverse), NHWC [P,C]
  float* out = nullptr;
  float* logJ = nullptr;       // modes 2, 3: optional full log_diag_J tensor [P,C] (mode 3: log_rescale)
  const float* scale = nullptr;
  const float* sshift = nullptr;
  const float* run_mean = nullptr;   // out_bn running statistics (modes 2, 3)
  const float* run_var = nullptr;
  double* sums = nullptr;      // mode 1: [2 cio]
  double* logdet_acc = nullptr;
};
struct ConvArgs {
  const float* x;      // [B,S,S,kpad] activated input
  const float* w;      // [taps][npad][kpad]
  const float* bias;   // [n] or null
  const float* res;    // [P,ldy] or null; may alias y
  float* y;            // [P,ldy]
  double* stats;       // [2n] or null
  int B, S, kpad, n, npad, taps, ldy;
  // fused ReLU+BN backward epilogue (dgrad only, tensor-core kernel only): y = acc * 1[bn_x*scale+shift > 0],
  // stats = (sum y, sum y * xhat) with the coefficients of bn_save = (mean, rstd, scale, shift)[n]
  const float* bn_x = nullptr;     // [P,ldy] raw pre-BN activations
  const float* bn_save = nullptr;  // [4n]
  // TF32 tier: y is read raw by another conv MMA (trunk a_i -> skip convs / wgrads, d a_i -> dgrads / wgrads), so
  // the epilogue rounds it to nearest TF32; tcgen05 would otherwise truncate it (biased towards zero)
  int round_out = 0;
  // tensor-core kernel only: x is the RAW pre-BN activation; relu(bn(x)) is applied to the staged operand tiles
  const BnPrologue* xf = nullptr;
  // tensor-core kernel only, eval mode: y = relu(y * post_scale[n] + post_shift[n]) after bias / residual -- the batch
  // norm (running statistics = a fixed affine map) + ReLU of the NEXT layer, folded into this conv's epilogue
  const float* post_scale = nullptr;
  const float* post_shift = nullptr;
  const CplEpilogue* cpl = nullptr;          // with xf only: the out conv of a coupling's s/t network
  // K-concatenated input: x is `segs` tensors [B,S,S,kpad] lying `seg_stride` floats apart, w has segs*kpad columns
  // (y = sum_i conv(x_i, w[:, i*kpad:(i+1)*kpad])); ldw = row stride of w in floats (0 = segs*kpad)
  int segs = 1;
  size_t seg_stride = 0;
  int ldw = 0;
  // fp32-accurate tensor-core tier ("3xTF32", conv_tc.cu): split operands, three MMAs per K step.  w_lo_delta = distance
  // in floats from w to its lo copy (w - trunc_tf32(w), same layout), written by the weight-norm kernel.
  int x3 = 0;
  size_t w_lo_delta = 0;
};
int k_conv_fwd_fp32(const ConvArgs& a, cudaStream_t st);
int k_conv_fwd_tf32(const ConvArgs& a, cudaStream_t st);
bool conv_tf32_fusable(const ConvArgs& a);   // true when k_conv_fwd_tf32 runs the tensor-core kernel for `a`
bool conv_tf32_prologue_ok(const ConvArgs& a);   // ... and can take a BnPrologue
// one member of a wgrad group (k_conv_wgrad_tf32 with njobs > 1): jobs share the shape, every pointer is per job
constexpr int kMaxWgradJobs = 8;
struct WgradJob {
  const float* x;
  const float* dy;
  float* dw;
  float* dbias;          // or null
  const float* xf_save;  // or null
};
struct WgradArgs {
  const float* x;      // [B,S,S,kpad]
  const float* dy;     // [P,lddy]
  float* dw;           // [taps][npad][kpad]  (+=)
  float* dbias;        // [n] (+=) or null
  int B, S, kpad, n, npad, taps, lddy;
  // tensor-core kernel only: x is the raw pre-BN activation of the forward conv's input; its boxes are rewritten
  // to tf32(relu(x * scale + shift)) in shared memory with (mean, rstd, scale, shift)[xf_C] = xf_save
  const float* xf_save = nullptr;
  int xf_C = 0;
  // K-concatenated x (see ConvArgs): dw has segs*kpad columns; lddw = row stride of dw in floats (0 = segs*kpad)
  int segs = 1;
  size_t seg_stride = 0;
  int lddw = 0;
  int x3 = 0;          // 3xTF32: x and dy boxes are split into hi / lo in shared memory, three MMAs per K step
  // tensor-core kernel only: njobs > 1 = a group of independent wgrads of this shape in ONE launch; x / dy / dw / dbias /
  // xf_save above are then ignored in favour of jobs[0 .. njobs)
  int njobs = 1;
  const WgradJob* jobs = nullptr;
};
bool wgrad_tf32_prologue_ok(const WgradArgs& a);
int k_conv_wgrad_fp32(const WgradArgs& a, cudaStream_t st);
int k_conv_wgrad_tf32(const WgradArgs& a, cudaStream_t st);

}  // namespace rnvp
