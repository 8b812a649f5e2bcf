// HBM-bound kernels of the RealNVP hot path: layout transforms, logit dequantisation,
// batch-norm apply / backward, the coupling maps with their per-sample log-det reductions,
// the prior term and weight normalisation.  All fp32, NHWC, coalesced; reductions go through
// warp shuffles -> shared memory -> one double atomic per block and channel.
#include <cstdlib>
#include "kernels.h"

namespace rnvp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RNVP_PDL");
    on = (e && e[0] == '0') ? 0 : 1;      // on by default (+3% on the training step); RNVP_PDL=0 disables
  }
  return on != 0;
}

// Serpentine sweeps: consecutive kernels of a chain walk their tensors in opposite directions, so each
// one starts on the part of its input that the previous kernel wrote last and that is still resident in
// the 126 MB L2 (trunk tensors are 33-134 MB at batch 256).  RNVP_SERPENTINE=0 switches it off.
int next_sweep_dir() {
  static int on = -1;
  static thread_local int dir = 0;
  if (on < 0) {
    const char* e = getenv("RNVP_SERPENTINE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on) return 0;
  dir ^= 1;
  return dir;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

constexpr int kThreads = 256;

// =====================================================================================
// layout
// =====================================================================================
// k -> (dy,dx) of factor_out: 0:(0,0) 1:(1,1) 2:(0,1) 3:(1,0)   (flow_realnvp.py:148-164)
// q = 2*dy+dx of squeeze                                      (flow_realnvp.py:121-126)
__device__ __forceinline__ int q_of_k(int k) { return (0x2130 >> (4 * k)) & 0xf; }   // {0,3,1,2}

__global__ void permute_kernel(int mode, const float* __restrict__ hi, const float* __restrict__ sq,
                               const float* __restrict__ on, const float* __restrict__ off,
                               float* __restrict__ hi_o, float* __restrict__ sq_o,
                               float* __restrict__ on_o, float* __restrict__ off_o,
                               int B, int s, int C) {
  pdl_wait();
  pdl_trigger();
  // one thread per element of the "factored" index space (b,i,j,k,c)
  int64_t total = (int64_t)B * s * s * 4 * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    int64_t r = e / C;
    int k = (int)(r % 4);
    r /= 4;
    int j = (int)(r % s);
    r /= s;
    int i = (int)(r % s);
    int b = (int)(r / s);
    int q = q_of_k(k), dy = q >> 1, dx = q & 1;
    int64_t pix = ((int64_t)b * s + i) * s + j;
    int64_t a_hi = (((int64_t)b * 2 * s + (2 * i + dy)) * 2 * s + (2 * j + dx)) * C + c;
    int64_t a_sq = pix * 4 * C + 4 * c + q;
    int64_t a_fa = pix * 2 * C + (k & 1) * C + c;       // inside on (k<2) or off (k>=2)
    switch (mode) {
      case PERM_SQUEEZE: sq_o[a_sq] = hi[a_hi]; break;
      case PERM_UNDO_SQUEEZE: hi_o[a_hi] = sq[a_sq]; break;
      case PERM_FACTOR_OUT: (k < 2 ? on_o : off_o)[a_fa] = hi[a_hi]; break;
      case PERM_RESTORE: hi_o[a_hi] = (k < 2 ? on : off)[a_fa]; break;
      case PERM_UNSQ_FACTOR: (k < 2 ? on_o : off_o)[a_fa] = sq[a_sq]; break;
      case PERM_FACTOR_SQ: sq_o[a_sq] = (k < 2 ? on : off)[a_fa]; break;
    }
  }
}

int k_permute(PermMode mode, const float* hi, const float* sq, const float* on, const float* off,
              float* hi_o, float* sq_o, float* on_o, float* off_o, int B, int s, int C, cudaStream_t st) {
  int64_t total = (int64_t)B * s * s * 4 * C;
  if (total == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(permute_kernel, dim3(grid_for(total, kThreads)), dim3(kThreads), 0, st, (int)mode, hi, sq, on, off, hi_o, sq_o,
                                                                  on_o, off_o, B, s, C));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// NCHW <-> NHWC through a 32x32 shared tile per (b, hw-tile, c-tile); C is small here
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (int64_t)b * C * HW;
  float* dst = out + (int64_t)b * C * HW;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, hw = hw0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && hw < HW) ? src[(int64_t)c * HW + hw] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int hw = hw0 + r, c = c0 + threadIdx.x;
    if (c < C && hw < HW) dst[(int64_t)hw * C + c] = tile[threadIdx.x][r];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (int64_t)b * C * HW;
  float* dst = out + (int64_t)b * C * HW;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int hw = hw0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && hw < HW) ? src[(int64_t)hw * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) dst[(int64_t)c * HW + hw] = tile[threadIdx.x][r];
  }
}
int k_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, cudaStream_t st) {
  if (B == 0) return RNVP_OK;
  RNVP_REQUIRE(B <= 65535, "batch %d exceeds grid.z limit", B);
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, st>>>(in, out, C, H * W);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
int k_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, cudaStream_t st) {
  if (B == 0) return RNVP_OK;
  RNVP_REQUIRE(B <= 65535, "batch %d exceeds grid.z limit", B);
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, st>>>(in, out, C, H * W);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// =====================================================================================
// logit dequantisation
// =====================================================================================
__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// Philox4x32-10: counter (idx, offset), key seed
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    uint32_t lo0 = mulhilo(0xD2511F53u, c0, &hi0);
    uint32_t lo1 = mulhilo(0xCD9E8D57u, c2, &hi1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// one block row per sample: grid (blocks_per_sample, B)
__global__ void logit_fwd_kernel(const float* __restrict__ xf, const uint8_t* __restrict__ xu8,
                                 const float* __restrict__ noise, float* __restrict__ y,
                                 float* __restrict__ logdet, int n, float constraint, float sp_pre,
                                 uint64_t seed, uint64_t offset) {
  int b = blockIdx.y;
  int64_t base = (int64_t)b * n;
  float acc = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int64_t idx = base + e;
    float x = xf ? xf[idx] : (float)xu8[idx] / 255.0f;
    float u;
    if (noise) {
      u = noise[idx];
    } else {
      uint32_t r[4];
      philox4(seed, (uint64_t)(idx >> 2), offset, r);
      u = (float)(r[idx & 3] >> 8) * (1.0f / 16777216.0f);      // [0,1)
    }
    // same op order as utils.py:49-64
    x = (x * 255.0f + u) / 256.0f;
    x = x * 2.0f;
    x = x - 1.0f;
    x = x * constraint;
    x = x + 1.0f;
    x = x / 2.0f;
    float v = logf(x) - logf(1.0f - x);
    y[idx] = v;
    acc += softplus_f(v) + softplus_f(-v) - sp_pre;
  }
  acc = warp_sum(acc);
  __shared__ float sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += sm[w];
    atomicAdd(&logdet[b], t);
  }
}
int k_logit_fwd(const float* xf, const uint8_t* xu8, const float* noise, float* y, float* logdet, int B,
                int n, float constraint, uint64_t seed, uint64_t offset, cudaStream_t st) {
  if (B == 0 || n == 0) return RNVP_OK;
  RNVP_REQUIRE(B <= 65535, "batch %d exceeds grid.y limit", B);
  RNVP_CUDA(cudaMemsetAsync(logdet, 0, sizeof(float) * B, st));
  double pre = log((double)constraint) - log(1.0 - (double)constraint);       // utils.py:67-68 (float64)
  float sp_pre = (float)log1p(exp(-pre));                                      // softplus(-pre)
  int bps = ceil_div(n, kThreads * 4);
  if (bps > 32) bps = 32;
  logit_fwd_kernel<<<dim3(bps, B), kThreads, 0, st>>>(xf, xu8, noise, y, logdet, n, constraint, sp_pre,
                                                       seed, offset);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void logit_inv_kernel(const float* __restrict__ y, float* __restrict__ x, size_t n, float constraint) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    float v = 1.0f / (expf(-y[e]) + 1.0f);      // utils.py:36-41
    v *= 2.0f;
    v -= 1.0f;
    v /= constraint;
    v += 1.0f;
    v /= 2.0f;
    x[e] = v;
  }
}
int k_logit_inv(const float* y, float* x, size_t n, float constraint, cudaStream_t st) {
  if (n == 0) return RNVP_OK;
  logit_inv_kernel<<<grid_for((int64_t)n, kThreads), kThreads, 0, st>>>(y, x, n, constraint);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// =====================================================================================
// batch norm on trunk tensors [P,C], C % 4 == 0, C <= 1024
// =====================================================================================
// h = relu(x*scale+shift); float4 per thread; columns >= C (padding up to ld) are left untouched
__global__ void bn_relu_kernel(const float4* __restrict__ x, float4* __restrict__ h, int64_t n4, int C, int ld,
                               const double* __restrict__ sums, double count,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* run_mean, float* run_var, float* save, int mode, int rnd, int rev, DpXchg xg) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];          // scale[C], shift[C]
  float* s_scale = sm;
  float* s_shift = sm + C;
  if (mode == 1) dp_exchange(xg, const_cast<double*>(sums), 2 * C, blockIdx.x == 0, threadIdx.x, blockDim.x, 0);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (mode == 2) {
      s_scale[c] = save[2 * C + c];
      s_shift[c] = save[3 * C + c];
      continue;
    }
    BnCoef k = mode == 1 ? bn_coef_from_sums(dp_reduced(xg, sums, c), dp_reduced(xg, sums, C + c), count, gamma[c], beta[c])
                         : bn_coef_from_running(run_mean[c], run_var[c], gamma[c], beta[c]);
    s_scale[c] = k.scale;
    s_shift[c] = k.shift;
    if (mode == 1 && blockIdx.x == 0) {
      save[c] = k.mean;
      save[C + c] = k.rstd;
      save[2 * C + c] = k.scale;
      save[3 * C + c] = k.shift;
      // nn.BatchNorm2d running update: momentum 0.1, unbiased variance
      double unb = count > 1.0 ? (double)k.var * count / (count - 1.0) : (double)k.var;
      run_mean[c] = (1.f - kBnMomentum) * run_mean[c] + kBnMomentum * k.mean;
      run_var[c] = (1.f - kBnMomentum) * run_var[c] + kBnMomentum * (float)unb;
    }
  }
  __syncthreads();
  const int l4 = ld >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;                       // independent 16-byte loads in flight per thread
  for (int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e0 < n4; e0 += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t e = e0 + u * stride;
      if (e < n4) v[u] = __ldcs(&x[rev ? n4 - 1 - e : e]);      // x is not read again before the backward
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t e = e0 + u * stride;
      if (e >= n4) break;
      if (rev) e = n4 - 1 - e;
      int c = (int)(e % l4) * 4;
      if (c >= C) continue;
      float4 w = v[u];
      w.x = maybe_round(fmaxf(fmaf(w.x, s_scale[c + 0], s_shift[c + 0]), 0.f), rnd);
      w.y = maybe_round(fmaxf(fmaf(w.y, s_scale[c + 1], s_shift[c + 1]), 0.f), rnd);
      w.z = maybe_round(fmaxf(fmaf(w.z, s_scale[c + 2], s_shift[c + 2]), 0.f), rnd);
      w.w = maybe_round(fmaxf(fmaf(w.w, s_scale[c + 3], s_shift[c + 3]), 0.f), rnd);
      h[e] = w;
    }
  }
}
int k_bn_relu(const float* x, float* h, int P, int C, int ld, const double* sums, double count,
              const float* gamma, const float* beta, float* run_mean, float* run_var, float* save, int mode,
              int tf32_round, cudaStream_t st, DpXchg xg) {
  if (P == 0) return RNVP_OK;
  RNVP_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld >= C, "bn_relu: C=%d ld=%d unsupported", C, ld);
  int64_t n4 = (int64_t)P * ld / 4;
  RNVP_CUDA(launch_pdl(bn_relu_kernel, dim3(grid_for(n4, kThreads * 4, kNumSMs * 8)), dim3(kThreads),
                       2 * C * sizeof(float), st, (const float4*)x, (float4*)h, n4, C, ld, sums, count, gamma, beta,
                       run_mean, run_var, save, mode, tf32_round, next_sweep_dir(), xg));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

__global__ void bn_eval_coef_kernel(const BnEvalJob* __restrict__ jobs, float* __restrict__ save_base) {
  const BnEvalJob j = jobs[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j.C) return;
  const BnCoef k = bn_coef_from_running(j.rm[c], j.rv[c], j.gamma[c], j.beta[c]);
  float* save = save_base + j.save_off;
  save[2 * j.C + c] = k.scale;
  save[3 * j.C + c] = k.shift;
}
int k_bn_eval_coefs(const BnEvalJob* jobs_dev, int njobs, int max_c, float* save_base, cudaStream_t st) {
  if (njobs == 0) return RNVP_OK;
  bn_eval_coef_kernel<<<dim3(ceil_div(max_c, 128), njobs), 128, 0, st>>>(jobs_dev, save_base);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// Column reduction helper: every thread owns float4-column `col` (cols = C/4) and row `row`;
// acc[v][0..3] are its partial sums of NV quantities; reduce over rows, one double atomic per channel.
template <int NV>
__device__ __forceinline__ void block_col_reduce(float (&acc)[NV][4], int col, int row, int cols, int rows,
                                                 int C, double* __restrict__ out, float* sm /*[NV*4*256]*/) {
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < 4; ++k) sm[(v * 4 + k) * kThreads + threadIdx.x] = acc[v][k];
  __syncthreads();
  if (row == 0 && col < cols) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += sm[(v * 4 + k) * kThreads + r * cols + col];
        atomicAdd(&out[v * C + col * 4 + k], (double)t);
      }
  }
}

// gm = g * 1[relu input > 0]; sums2 += (sum gm, sum gm*xhat)
__global__ void bn_bwd_reduce_kernel(const float4* g, const float4* __restrict__ x,
                                     float4* gm_out, int P, int C, int ld,
                                     const float* __restrict__ save, double* __restrict__ sums2) {
  __shared__ float sm[2 * 4 * kThreads];
  int cols = C >> 2, rows = kThreads / cols;
  int col = threadIdx.x % cols, row = threadIdx.x / cols;
  float acc[2][4] = {};
  if (row < rows) {
    int c = col * 4;
    float mean[4], rstd[4], scale[4], shift[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mean[k] = save[c + k];
      rstd[k] = save[C + c + k];
      scale[k] = save[2 * C + c + k];
      shift[k] = save[3 * C + c + k];
    }
    constexpr int U = 4;                     // rows in flight per thread (8 independent 16-byte loads)
    const int pstride = gridDim.x * rows;
    for (int p0 = blockIdx.x * rows + row; p0 < P; p0 += U * pstride) {
      float4 gv[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int p = p0 + u * pstride;
        if (p < P) {
          int64_t e = (int64_t)p * (ld >> 2) + col;
          gv[u] = g[e];
          xv[u] = x[e];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int p = p0 + u * pstride;
        if (p >= P) break;
        int64_t e = (int64_t)p * (ld >> 2) + col;
        float gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w}, xx[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float hval = fmaf(xx[k], scale[k], shift[k]);
          float m = hval > 0.f ? gg[k] : 0.f;
          gg[k] = m;
          acc[0][k] += m;
          acc[1][k] += m * ((xx[k] - mean[k]) * rstd[k]);
        }
        gm_out[e] = make_float4(gg[0], gg[1], gg[2], gg[3]);
      }
    }
  }
  block_col_reduce<2>(acc, col, row, cols, rows, C, sums2, sm);
}
int k_bn_bwd_reduce(const float* g, const float* x, float* gm_out, int P, int C, int ld, const float* save,
                    double* sums2, cudaStream_t st) {
  if (P == 0) return RNVP_OK;
  RNVP_REQUIRE(C % 4 == 0 && C / 4 <= kThreads && ld >= C && ld % 4 == 0, "bn_bwd_reduce: unsupported C=%d", C);
  int rows = kThreads / (C / 4);
  bn_bwd_reduce_kernel<<<grid_for(P, rows * 16, kNumSMs * 6), kThreads, 0, st>>>(
      (const float4*)g, (const float4*)x, (float4*)gm_out, P, C, ld, save, sums2);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

__global__ void bn_bwd_apply_kernel(const float4* gm, const float4* __restrict__ x,
                                    float4* dx, const float4* add, int64_t n4, int C, int ld,
                                    const float* __restrict__ save, const double* __restrict__ sums2,
                                    double count, const float* __restrict__ gamma, float* dgamma,
                                    float* dbeta, float inv_world, int raw_x_sums, int rnd, int rev, DpXchg xg) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];       // a[C] = gamma*rstd, m1[C], m2[C], mean[C], rstd[C]
  float *s_a = sm, *s_m1 = sm + C, *s_m2 = sm + 2 * C, *s_mean = sm + 3 * C, *s_rstd = sm + 4 * C;
  dp_exchange(xg, const_cast<double*>(sums2), 2 * C, blockIdx.x == 0, threadIdx.x, blockDim.x, 0);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float rstd = save[C + c];
    // second statistic: sum g*xhat, or (fused dgrad epilogue) sum g*x which maps to it linearly
    double s1 = dp_reduced(xg, sums2, c);
    const double s2raw = dp_reduced(xg, sums2, C + c);
    double s2 = raw_x_sums ? (double)rstd * (s2raw - (double)save[c] * s1) : s2raw;
    s_a[c] = gamma[c] * rstd;
    s_m1[c] = (float)(s1 / count);
    s_m2[c] = (float)(s2 / count);
    s_mean[c] = save[c];
    s_rstd[c] = rstd;
    if (blockIdx.x == 0) {
      // the sums are GLOBAL under data parallelism; every rank adds its 1/world share so that the
      // gradient average over ranks is the global gradient
      dbeta[c] += (float)s1 * inv_world;
      dgamma[c] += (float)s2 * inv_world;
    }
  }
  __syncthreads();
  const int l4 = ld >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e0 < n4; e0 += U * stride) {
    float4 gv[U], xv[U], ov[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t e = e0 + u * stride;
      if (e < n4) {
        if (rev) e = n4 - 1 - e;
        gv[u] = gm[e];
        xv[u] = __ldcs(&x[e]);                 // last use of the saved activation
        if (add) ov[u] = add[e];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t e = e0 + u * stride;
      if (e >= n4) break;
      if (rev) e = n4 - 1 - e;
      int c = (int)(e % l4) * 4;
      if (c >= C) continue;
      float gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w}, xx[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xh = (xx[k] - s_mean[c + k]) * s_rstd[c + k];
        r[k] = s_a[c + k] * (gg[k] - s_m1[c + k] - xh * s_m2[c + k]);
      }
      if (add) {
        r[0] += ov[u].x; r[1] += ov[u].y; r[2] += ov[u].z; r[3] += ov[u].w;
      }
      dx[e] = make_float4(maybe_round(r[0], rnd), maybe_round(r[1], rnd), maybe_round(r[2], rnd), maybe_round(r[3], rnd));
    }
  }
}
int k_bn_bwd_apply(const float* gm, const float* x, float* dx, const float* add, int P, int C, int ld,
                   const float* save, const double* sums2, double count, const float* gamma,
                   float* dgamma, float* dbeta, float inv_world, int raw_x_sums, int tf32_round, cudaStream_t st,
                   DpXchg xg) {
  if (P == 0) return RNVP_OK;
  int64_t n4 = (int64_t)P * ld / 4;
  RNVP_CUDA(launch_pdl(bn_bwd_apply_kernel, dim3(grid_for(n4, kThreads * 4, kNumSMs * 8)), dim3(kThreads),
                       5 * C * sizeof(float), st, (const float4*)gm, (const float4*)x, (float4*)dx, (const float4*)add, n4, C, ld,
                       save, sums2, count, gamma, dgamma, dbeta, inv_world, raw_x_sums, tf32_round, next_sweep_dir(), xg));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// =====================================================================================
// coupling kernels: thread per pixel, loop over the (few) coupling channels
// =====================================================================================
constexpr int kMaxCio = 256;     // coupling channels handled by the shared-memory accumulators

// add a per-pixel value into a per-sample double accumulator (pixels of a sample are contiguous)
__device__ __forceinline__ void sample_accumulate(double* acc, int p, int P, int hw, float v) {
  if ((hw & 31) == 0) {            // a warp never straddles two samples
    float t = warp_sum(p < P ? v : 0.f);
    if ((threadIdx.x & 31) == 0 && p - (int)(threadIdx.x & 31) < P) {
      int b = (p < P ? p : P - 1) / hw;
      atomicAdd(&acc[b], (double)t);
    }
  } else if (p < P) {
    atomicAdd(&acc[p / hw], (double)v);
  }
}

// per-channel accumulate of NV quantities through warp shuffles into shared floats
__device__ __forceinline__ void chan_accumulate(float* sm_acc, int idx, float v) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sm_acc[idx], v);
}

// The coupling kernels map one thread to one pixel and loop over channels; grid.y splits that loop so
// that the small deep-scale tensors (4096 pixels x 48..96 channels at S = 4) still fill the machine.
constexpr int kCplChanPerBlock = 8;
__device__ __forceinline__ void cpl_chan_range(int cio, int& c0, int& c1) {
  c0 = blockIdx.y * kCplChanPerBlock;
  c1 = min(cio, c0 + kCplChanPerBlock);
}
static inline dim3 cpl_grid(int P, int per_block, int cio, int max_x = kNumSMs * 4) {
  return dim3(grid_for(P, per_block, max_x), ceil_div(cio, kCplChanPerBlock));
}

__global__ void cpl_in_stats_kernel(const float* __restrict__ x, CplGeom g, double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  __shared__ float acc[2 * kMaxCio];
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  int P = g.P();
  int iters = ceil_div(P, gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = p < P;
    float m = ok ? g.mask_in(p) : 0.f;
    for (int c = c0; c < c1; ++c) {
      float v = ok ? x[(int64_t)p * g.C + g.in_off + c] * m : 0.f;
      chan_accumulate(acc, c, v);
      chan_accumulate(acc, g.cio + c, v * v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x)
    if (i % g.cio >= c0 && i % g.cio < c1) atomicAdd(&sums[i], (double)acc[i]);
}
int k_cpl_in_stats(const float* x, CplGeom g, double* sums, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_REQUIRE(g.cio <= kMaxCio, "coupling with %d channels unsupported (max %d)", g.cio, kMaxCio);
  RNVP_CUDA(launch_pdl(cpl_in_stats_kernel, cpl_grid(g.P(), kThreads * 2, g.cio), dim3(kThreads), 0, st, x, g, sums));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// h0 = relu(cat(u, -u[, m])), u = in_bn(x*m); one float4 of h0 per thread
__global__ void cpl_in_build_kernel(const float* __restrict__ x, CplGeom g, const double* __restrict__ sums,
                                    double count, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* run_mean, float* run_var,
                                    float* __restrict__ save, int training, float4* __restrict__ h0, int rnd) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_scale[kMaxCio], s_shift[kMaxCio];
  for (int c = threadIdx.x; c < g.cio; c += blockDim.x) {
    BnCoef k = training ? bn_coef_from_sums(sums[c], sums[g.cio + c], count, gamma[c], beta[c])
                        : bn_coef_from_running(run_mean[c], run_var[c], gamma[c], beta[c]);
    s_scale[c] = k.scale;
    s_shift[c] = k.shift;
    if (training && blockIdx.x == 0) {
      save[c] = k.mean;
      save[g.cio + c] = k.rstd;
      save[2 * g.cio + c] = k.scale;
      save[3 * g.cio + c] = k.shift;
      double unb = count > 1.0 ? (double)k.var * count / (count - 1.0) : (double)k.var;
      run_mean[c] = (1.f - kBnMomentum) * run_mean[c] + kBnMomentum * k.mean;
      run_var[c] = (1.f - kBnMomentum) * run_var[c] + kBnMomentum * (float)unb;
    }
  }
  __syncthreads();
  const uint32_t q4 = (uint32_t)g.cin_pad >> 2;
  const uint32_t total = (uint32_t)g.P() * q4;          // < 2^32: checked by the launcher (32-bit index math)
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const uint32_t pu = e / q4;
    const int p = (int)pu, q = (int)(e - pu * q4);
    float m = g.mask_in(p);
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int j = q * 4 + k;
      float v = 0.f;
      if (j < 2 * g.cio) {
        int c = j < g.cio ? j : j - g.cio;
        float u = fmaf(x[(int64_t)p * g.C + g.in_off + c] * m, s_scale[c], s_shift[c]);
        v = maybe_round(fmaxf(j < g.cio ? u : -u, 0.f), rnd);
      } else if (g.ckbd && j == 2 * g.cio) {
        v = m;                                   // relu(mask) = mask (modules_realnvp.py:275-276,259)
      }
      r[k] = v;
    }
    h0[e] = make_float4(r[0], r[1], r[2], r[3]);
  }
}
int k_cpl_in_build(const float* x, CplGeom g, const double* sums, double count, const float* gamma,
                   const float* beta, float* run_mean, float* run_var, float* save, int training,
                   float* h0, int tf32_round, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  int64_t total = (int64_t)g.P() * (g.cin_pad / 4);
  RNVP_REQUIRE(total < (int64_t)1 << 32, "cpl_in_build: %lld float4 elements exceed the 32-bit index range", (long long)total);
  RNVP_CUDA(launch_pdl(cpl_in_build_kernel, dim3(grid_for(total, kThreads * 2)), dim3(kThreads), 0, st, x, g, sums, count, gamma, beta, run_mean, run_var, save, training, (float4*)h0, tf32_round));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// x' = x*exp(s)+t on the transformed channels, s = (scale*tanh(l)+shift)*(1-m), t *= (1-m)
__global__ void cpl_fwd_a_kernel(const float* __restrict__ x, const float* __restrict__ stt, CplGeom g,
                                 const float* __restrict__ scale_p, const float* __restrict__ sshift_p,
                                 float* __restrict__ xprime, double* __restrict__ sums,
                                 double* __restrict__ logdet_acc, int training) {
  pdl_wait();
  pdl_trigger();
  __shared__ float acc[2 * kMaxCio];
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const float scale = *scale_p, sshift = *sshift_p;
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P(), hw = g.S * g.S;
  int iters = ceil_div(P, gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = p < P;
    float keep = ok ? 1.f - (g.ckbd ? g.mask_in(p) : 0.f) : 0.f;
    float ssum = 0.f;
    for (int c = c0; c < c1; ++c) {
      float xp = 0.f;
      if (ok) {
        float t = stt[(int64_t)p * g.cst_pad + c] * keep;
        float l = stt[(int64_t)p * g.cst_pad + g.cio + c];
        float s = (scale * tanhf(l) + sshift) * keep;
        xp = x[(int64_t)p * g.C + g.on_off + c] * expf(s) + t;
        xprime[(int64_t)p * g.cio + c] = xp;
        ssum += s;
      }
      if (training) {
        chan_accumulate(acc, c, xp);
        chan_accumulate(acc, g.cio + c, xp * xp);
      }
    }
    sample_accumulate(logdet_acc, p, P, hw, ssum);
  }
  if (training) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x)
      if (i % g.cio >= c0 && i % g.cio < c1) atomicAdd(&sums[i], (double)acc[i]);
  }
}
int k_cpl_fwd_a(const float* x, const float* stt, CplGeom g, const float* scale, const float* sshift,
                float* xprime, double* sums, double* logdet_acc, int training, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_fwd_a_kernel, cpl_grid(g.P(), kThreads, g.cio), dim3(kThreads), 0, st, x, stt, g, scale, sshift, xprime, sums, logdet_acc, training));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// y = out_bn(x')*(1-m) + x'*m; log-det term -0.5*log(var+eps)*(1-m); optional full logJ
__global__ void cpl_fwd_b_kernel(const float* __restrict__ xprime, const float* __restrict__ x,
                                 const float* __restrict__ stt, CplGeom g, const double* __restrict__ sums,
                                 double count, float* run_mean, float* run_var, float* __restrict__ save,
                                 int training, const float* __restrict__ scale_p,
                                 const float* __restrict__ sshift_p, float* __restrict__ y,
                                 float* __restrict__ logJ, double* __restrict__ logdet_acc) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_mean[kMaxCio], s_rstd[kMaxCio], s_hl[kMaxCio];     // hl = 0.5*log(var+eps)
  __shared__ float s_L;
  for (int c = threadIdx.x; c < g.cio; c += blockDim.x) {
    BnCoef k = training ? bn_coef_from_sums(sums[c], sums[g.cio + c], count, 1.f, 0.f)
                        : bn_coef_from_running(run_mean[c], run_var[c], 1.f, 0.f);
    s_mean[c] = k.mean;
    s_rstd[c] = k.rstd;
    s_hl[c] = 0.5f * logf(k.var + 1e-5f);
    if (training && blockIdx.x == 0 && blockIdx.y == 0) {
      save[c] = k.mean;
      save[g.cio + c] = k.rstd;
      double unb = count > 1.0 ? (double)k.var * count / (count - 1.0) : (double)k.var;
      run_mean[c] = (1.f - kBnMomentum) * run_mean[c] + kBnMomentum * k.mean;
      run_var[c] = (1.f - kBnMomentum) * run_var[c] + kBnMomentum * (float)unb;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float L = 0.f;
    for (int c = 0; c < g.cio; ++c) L += s_hl[c];
    s_L = L;
  }
  __syncthreads();
  const float scale = logJ ? *scale_p : 0.f, sshift = logJ ? *sshift_p : 0.f;
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P(), hw = g.S * g.S;
  int iters = ceil_div(P, gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = p < P;
    float keep = ok ? 1.f - (g.ckbd ? g.mask_in(p) : 0.f) : 0.f;
    if (ok) {
      for (int c = c0; c < c1; ++c) {
        float xp = xprime[(int64_t)p * g.cio + c];
        float yn = (xp - s_mean[c]) * s_rstd[c];
        y[(int64_t)p * g.C + g.on_off + c] = keep != 0.f ? yn : xp;
        if (!g.ckbd) y[(int64_t)p * g.C + g.in_off + c] = x[(int64_t)p * g.C + g.in_off + c];
        if (logJ) {
          float l = stt[(int64_t)p * g.cst_pad + g.cio + c];
          float s = (scale * tanhf(l) + sshift) * keep;
          logJ[(int64_t)p * g.C + g.on_off + c] = s - s_hl[c] * keep;
          if (!g.ckbd) logJ[(int64_t)p * g.C + g.in_off + c] = 0.f;
        }
      }
    }
    if (blockIdx.y == 0) sample_accumulate(logdet_acc, p, P, hw, -s_L * keep);
  }
}
int k_cpl_fwd_b(const float* xprime, const float* x, const float* stt, CplGeom g, const double* sums,
                double count, float* run_mean, float* run_var, float* save, int training,
                const float* scale, const float* sshift, float* y, float* logJ, double* logdet_acc,
                cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_fwd_b_kernel, cpl_grid(g.P(), kThreads, g.cio), dim3(kThreads), 0, st, xprime, x, stt, g, sums, count, run_mean, run_var, save, training, scale, sshift, y, logJ, logdet_acc));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// reverse=True branch: modules_realnvp.py:284-291, 345-351
__global__ void cpl_inv_kernel(const float* __restrict__ y, const float* __restrict__ stt, CplGeom g,
                               const float* __restrict__ run_mean, const float* __restrict__ run_var,
                               const float* __restrict__ scale_p, const float* __restrict__ sshift_p,
                               float* __restrict__ x, float* __restrict__ logJ) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_mean[kMaxCio], s_hl[kMaxCio];
  for (int c = threadIdx.x; c < g.cio; c += blockDim.x) {
    s_mean[c] = run_mean[c];
    s_hl[c] = 0.5f * logf(run_var[c] + 1e-5f);
  }
  __syncthreads();
  const float scale = *scale_p, sshift = *sshift_p;
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    float keep = 1.f - (g.ckbd ? g.mask_in(p) : 0.f);
    for (int c = c0; c < c1; ++c) {
      float yv = y[(int64_t)p * g.C + g.on_off + c];
      float xt = yv * expf(s_hl[c] * keep) + s_mean[c] * keep;
      float t = stt[(int64_t)p * g.cst_pad + c] * keep;
      float l = stt[(int64_t)p * g.cst_pad + g.cio + c];
      float s = (scale * tanhf(l) + sshift) * keep;
      x[(int64_t)p * g.C + g.on_off + c] = (xt - t) * expf(-s);
      if (!g.ckbd) x[(int64_t)p * g.C + g.in_off + c] = y[(int64_t)p * g.C + g.in_off + c];
      if (logJ) {                                  // the reference returns log_rescale (modules_realnvp.py:283, 302)
        logJ[(int64_t)p * g.C + g.on_off + c] = s;
        if (!g.ckbd) logJ[(int64_t)p * g.C + g.in_off + c] = 0.f;
      }
    }
  }
}
int k_cpl_inv(const float* y, const float* stt, CplGeom g, const float* run_mean, const float* run_var,
              const float* scale, const float* sshift, float* x, float* logJ, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_inv_kernel, cpl_grid(g.P(), kThreads, g.cio, kNumSMs * 8), dim3(kThreads), 0, st, y, stt, g, run_mean, run_var,
                                                                                      scale, sshift, x, logJ));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// backward, pass A: sums2 = (sum g, sum g*xhat') with g = dy*(1-m); slot [2cio] = K
__global__ void cpl_bwd_a_kernel(const float* __restrict__ dy, const float* __restrict__ xprime, CplGeom g,
                                 const float* __restrict__ save, const float* __restrict__ dll,
                                 double* __restrict__ sums2) {
  pdl_wait();
  pdl_trigger();
  __shared__ float acc[2 * kMaxCio];
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P();
  int iters = ceil_div(P, gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = p < P;
    float keep = ok ? 1.f - (g.ckbd ? g.mask_in(p) : 0.f) : 0.f;
    for (int c = c0; c < c1; ++c) {
      float gv = 0.f, xh = 0.f;
      if (ok) {
        gv = dy[(int64_t)p * g.C + g.on_off + c] * keep;
        xh = (xprime[(int64_t)p * g.cio + c] - save[c]) * save[g.cio + c];
      }
      chan_accumulate(acc, c, gv);
      chan_accumulate(acc, g.cio + c, gv * xh);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x)
    if (i % g.cio >= c0 && i % g.cio < c1) atomicAdd(&sums2[i], (double)acc[i]);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32) {
    // K = sum over pixels of dll_b * keep.  keep = 1 - mask covers every position of a channelwise coupling;
    // of a checkerboard one it covers the positions with (cfg + i + j) even: ceil(S*S/2) of them for cfg 0,
    // floor(S*S/2) for cfg 1 (closed form: a serial count over S*S positions cost 100 us at S = 64)
    double k = 0.0;
    for (int b = threadIdx.x; b < g.B; b += 32) k += (double)dll[b];
    k = warp_sum_d(k);
    const int ss = g.S * g.S;
    const double npos = !g.ckbd ? (double)ss : (double)(g.cfg ? ss / 2 : (ss + 1) / 2);
    if (threadIdx.x == 0) atomicAdd(&sums2[2 * g.cio], k * npos);
  }
}
int k_cpl_bwd_a(const float* dy, const float* xprime, CplGeom g, const float* save, const float* dll,
                double* sums2, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_bwd_a_kernel, cpl_grid(g.P(), kThreads * 2, g.cio), dim3(kThreads), 0, st, dy, xprime, g, save, dll, sums2));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// backward, pass B: out_bn backward (+ log-det variance term), affine map backward
__global__ void cpl_bwd_b_kernel(const float* __restrict__ dy, const float* __restrict__ xprime,
                                 const float* __restrict__ x, const float* __restrict__ stt, CplGeom g,
                                 const float* __restrict__ save, const double* __restrict__ sums2,
                                 double count, const float* __restrict__ dll,
                                 const float* __restrict__ scale_p, const float* __restrict__ sshift_p,
                                 float* __restrict__ dst, float* __restrict__ dxdir, float* dscale,
                                 float* dsshift, int rnd) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_m1[kMaxCio], s_m2[kMaxCio];
  __shared__ float s_red[2];
  float kn = (float)(sums2[2 * g.cio] / count);
  for (int c = threadIdx.x; c < g.cio; c += blockDim.x) {
    s_m1[c] = (float)(sums2[c] / count);
    s_m2[c] = (float)(sums2[g.cio + c] / count);
  }
  if (threadIdx.x < 2) s_red[threadIdx.x] = 0.f;
  __syncthreads();
  const float scale = *scale_p, sshift = *sshift_p;
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P(), hw = g.S * g.S;
  float a_scale = 0.f, a_shift = 0.f;
  // Few channels (one channel block, 32-float dst rows): a thread owns a pixel = a 128-byte dst row, which it
  // would write with 32 scattered 4-byte stores; the rows of the block's 256 pixels are staged in shared
  // memory instead and leave as one contiguous, fully coalesced 32 KB burst.
  extern __shared__ float stage[];                 // [256][33] when staged
  const bool staged = gridDim.y == 1 && g.cst_pad == 32;
  for (int pb = blockIdx.x * blockDim.x; pb < P; pb += gridDim.x * blockDim.x) {
    const int p = pb + threadIdx.x;
    if (p < P) {
      float keep = 1.f - (g.ckbd ? g.mask_in(p) : 0.f);
      float dl_b = dll[p / hw];
      float* drow = staged ? stage + threadIdx.x * 33 : dst + (int64_t)p * g.cst_pad;
      for (int c = c0; c < c1; ++c) {
        float mean = save[c], rstd = save[g.cio + c];
        float dyv = dy[(int64_t)p * g.C + g.on_off + c];
        float xp = xprime[(int64_t)p * g.cio + c];
        float xh = (xp - mean) * rstd;
        float gv = dyv * keep;
        float dxp = rstd * (gv - s_m1[c] - xh * s_m2[c]) - kn * rstd * xh + dyv * (1.f - keep);
        float l = stt[(int64_t)p * g.cst_pad + g.cio + c];
        float th = tanhf(l);
        float s = (scale * th + sshift) * keep;
        float es = expf(s);
        float xv = x[(int64_t)p * g.C + g.on_off + c];
        float ds = (dxp * xv * es + dl_b) * keep;
        drow[c] = maybe_round(dxp * keep, rnd);
        drow[g.cio + c] = maybe_round(ds * scale * (1.f - th * th), rnd);
        dxdir[(int64_t)p * g.cio + c] = dxp * es;
        a_scale += ds * th;
        a_shift += ds;
      }
      if (blockIdx.y == 0)
        for (int c = 2 * g.cio; c < g.cst_pad; ++c) drow[c] = 0.f;
    }
    if (staged) {
      __syncthreads();
      const int rows = min((int)blockDim.x, P - pb);
      float4* out = reinterpret_cast<float4*>(dst + (int64_t)pb * 32);
      for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const float* r = stage + (i >> 3) * 33 + (i & 7) * 4;
        out[i] = make_float4(r[0], r[1], r[2], r[3]);
      }
      __syncthreads();
    }
  }
  a_scale = warp_sum(a_scale);
  a_shift = warp_sum(a_shift);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_red[0], a_scale);
    atomicAdd(&s_red[1], a_shift);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(dscale, s_red[0]);
    atomicAdd(dsshift, s_red[1]);
  }
}
int k_cpl_bwd_b(const float* dy, const float* xprime, const float* x, const float* stt, CplGeom g,
                const float* save, const double* sums2, double count, const float* dll, const float* scale,
                const float* sshift, float* dst, float* dxdir, float* dscale, float* dsshift, int tf32_round,
                cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_bwd_b_kernel, cpl_grid(g.P(), kThreads, g.cio), dim3(kThreads), kThreads * 33 * sizeof(float), st, dy, xprime, x, stt, g, save, sums2, count, dll, scale, sshift, dst, dxdir, dscale, dsshift, tf32_round));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// in-branch backward, pass A: du = dh0[c]*1[u>0] - dh0[cio+c]*1[u<0]; sums3 = (sum du, sum du*un)
__global__ void cpl_in_bwd_a_kernel(const float* __restrict__ dh0, const float* __restrict__ x, CplGeom g,
                                    const float* __restrict__ save, double* __restrict__ sums3) {
  pdl_wait();
  pdl_trigger();
  __shared__ float acc[2 * kMaxCio];
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P();
  int iters = ceil_div(P, gridDim.x * blockDim.x);
  for (int it = 0; it < iters; ++it) {
    int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = p < P;
    float m = ok ? g.mask_in(p) : 0.f;
    for (int c = c0; c < c1; ++c) {
      float du = 0.f, un = 0.f;
      if (ok) {
        float v = x[(int64_t)p * g.C + g.in_off + c] * m;
        float u = fmaf(v, save[2 * g.cio + c], save[3 * g.cio + c]);
        un = (v - save[c]) * save[g.cio + c];
        float d1 = dh0[(int64_t)p * g.cin_pad + c], d2 = dh0[(int64_t)p * g.cin_pad + g.cio + c];
        du = (u > 0.f ? d1 : 0.f) - (u < 0.f ? d2 : 0.f);
      }
      chan_accumulate(acc, c, du);
      chan_accumulate(acc, g.cio + c, du * un);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.cio; i += blockDim.x)
    if (i % g.cio >= c0 && i % g.cio < c1) atomicAdd(&sums3[i], (double)acc[i]);
}
int k_cpl_in_bwd_a(const float* dh0, const float* x, CplGeom g, const float* save, double* sums3,
                   cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_in_bwd_a_kernel, cpl_grid(g.P(), kThreads * 2, g.cio), dim3(kThreads), 0, st, dh0, x, g, save, sums3));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

__global__ void cpl_in_bwd_b_kernel(const float* __restrict__ dh0, const float* __restrict__ x,
                                    const float* __restrict__ dxdir, const float* __restrict__ dy,
                                    CplGeom g, const float* __restrict__ save,
                                    const double* __restrict__ sums3, double count,
                                    const float* __restrict__ gamma, float* dgamma, float* dbeta,
                                    float* __restrict__ dx, float inv_world) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_m1[kMaxCio], s_m2[kMaxCio];
  for (int c = threadIdx.x; c < g.cio; c += blockDim.x) {
    s_m1[c] = (float)(sums3[c] / count);
    s_m2[c] = (float)(sums3[g.cio + c] / count);
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      dbeta[c] += (float)sums3[c] * inv_world;
      dgamma[c] += (float)sums3[g.cio + c] * inv_world;
    }
  }
  __syncthreads();
  int c0, c1;
  cpl_chan_range(g.cio, c0, c1);
  int P = g.P();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    float m = g.mask_in(p);
    for (int c = c0; c < c1; ++c) {
      float v = x[(int64_t)p * g.C + g.in_off + c] * m;
      float u = fmaf(v, save[2 * g.cio + c], save[3 * g.cio + c]);
      float un = (v - save[c]) * save[g.cio + c];
      float d1 = dh0[(int64_t)p * g.cin_pad + c], d2 = dh0[(int64_t)p * g.cin_pad + g.cio + c];
      float du = (u > 0.f ? d1 : 0.f) - (u < 0.f ? d2 : 0.f);
      float dv = gamma[c] * save[g.cio + c] * (du - s_m1[c] - un * s_m2[c]);
      if (g.ckbd) {
        dx[(int64_t)p * g.C + c] = dxdir[(int64_t)p * g.cio + c] + dv * m;
      } else {
        dx[(int64_t)p * g.C + g.on_off + c] = dxdir[(int64_t)p * g.cio + c];
        dx[(int64_t)p * g.C + g.in_off + c] = dy[(int64_t)p * g.C + g.in_off + c] + dv;
      }
    }
  }
}
int k_cpl_in_bwd_b(const float* dh0, const float* x, const float* dxdir, const float* dy, CplGeom g,
                   const float* save, const double* sums3, double count, const float* gamma, float* dgamma,
                   float* dbeta, float* dx, float inv_world, cudaStream_t st) {
  if (g.P() == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(cpl_in_bwd_b_kernel, cpl_grid(g.P(), kThreads, g.cio), dim3(kThreads), 0, st, dh0, x, dxdir, dy, g, save, sums3, count, gamma, dgamma, dbeta, dx, inv_world));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// =====================================================================================
// prior and per-sample bookkeeping
// =====================================================================================
__global__ void prior_ll_kernel(const float* __restrict__ z, int n, float loc, float inv_2var, float cst,
                                double* __restrict__ acc) {
  int b = blockIdx.y;
  const float* zb = z + (int64_t)b * n;
  float a = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float d = zb[e] - loc;
    a += -(d * d) * inv_2var - cst;           // Normal.log_prob
  }
  a = warp_sum(a);
  __shared__ float sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += (double)sm[w];
    atomicAdd(&acc[b], t);
  }
}
int k_prior_ll(const float* z, int B, int n, float loc, float scale, double* acc, cudaStream_t st) {
  if (B == 0 || n == 0) return RNVP_OK;
  int bps = ceil_div(n, kThreads * 4);
  if (bps > 16) bps = 16;
  float cst = logf(scale) + 0.5f * logf(2.0f * 3.14159265358979323846f);
  prior_ll_kernel<<<dim3(bps, B), kThreads, 0, st>>>(z, n, loc, 1.0f / (2.0f * scale * scale), cst, acc);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void prior_grad_kernel(const float* __restrict__ z, const float* __restrict__ dll, float* dz,
                                  int accumulate, int64_t total, int n, float loc, float inv_var) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float v = -(z[e] - loc) * inv_var * dll[e / n];
    dz[e] = accumulate ? dz[e] + v : v;
  }
}
int k_prior_grad(const float* z, const float* dll, float* dz, int accumulate, int B, int n, float loc,
                 float scale, cudaStream_t st) {
  int64_t total = (int64_t)B * n;
  if (total == 0) return RNVP_OK;
  prior_grad_kernel<<<grid_for(total, kThreads * 2), kThreads, 0, st>>>(z, dll, dz, accumulate, total, n, loc,
                                                                        1.0f / (scale * scale));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void finalize_ll_kernel(const double* __restrict__ ld, const double* __restrict__ pr, float* ll,
                                   float* logdet, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    // the reference sums in fp32: log_prior (fp32) + log_det (fp32)   flow_realnvp.py:338-340
    float l = (float)ld[b], p = (float)pr[b];
    if (ll) ll[b] = p + l;
    if (logdet) logdet[b] = l;
  }
}
int k_finalize_ll(const double* logdet_acc, const double* prior_acc, float* ll, float* logdet, int B,
                  cudaStream_t st) {
  if (B == 0) return RNVP_OK;
  finalize_ll_kernel<<<ceil_div(B, kThreads), kThreads, 0, st>>>(logdet_acc, prior_acc, ll, logdet, B);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void add_kernel(float* dst, const float* __restrict__ src, size_t n) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    dst[e] += src[e];
}
int k_add(float* dst, const float* src, size_t n, cudaStream_t st) {
  if (n == 0) return RNVP_OK;
  add_kernel<<<grid_for((int64_t)n, kThreads * 2), kThreads, 0, st>>>(dst, src, n);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

__global__ void gather_first_kernel(const float* __restrict__ in, float* out, int B, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = in[(int64_t)b * n];
}
int k_gather_first(const float* in, float* out, int B, int n, cudaStream_t st) {
  if (B == 0) return RNVP_OK;
  gather_first_kernel<<<ceil_div(B, kThreads), kThreads, 0, st>>>(in, out, B, n);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// =====================================================================================
// weight norm: w = g * v / ||v||  (norm over (cin,kh,kw) per output channel)
// =====================================================================================
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
  return t;
}

// grid (pad32(max_cout), njobs); v is (cout, cin, taps) with taps innermost (NCHW-style weight).
// Block `co` writes row co of wf and column co of wb completely, zeros included.
// One block = (job, tile of 32 output channels).  Phase 1: the 32 row norms (a warp per row, coalesced).
// Phase 2, per tile of 32 input channels: the contiguous [32 co][32 ci * taps] slab of v goes through shared
// memory so that BOTH layouts are written with full 128-byte rows -- wf[tap][co][ci] along ci, the
// transposed and tap-flipped wb[taps-1-tap][ci][co] along co.  Padding rows / columns are written as zeros.
constexpr int kWnTile = 32;
__global__ void __launch_bounds__(256) weightnorm_fwd_kernel(const WnJob* __restrict__ jobs, float* __restrict__ wbase, int rnd,
                                                             size_t lo_delta) {
  extern __shared__ float wn_sm[];               // [32][32 * taps + 1] slab, then f[32]
  const WnJob j = jobs[blockIdx.y];
  const int co0 = blockIdx.x * kWnTile;
  if (co0 >= j.kpad_b) return;                   // kpad_b = pad32(cout) >= npad_f
  const int taps = j.taps, per = j.cin * taps;
  const int pitch = kWnTile * taps + 1;
  float* slab = wn_sm;
  float* f = wn_sm + kWnTile * pitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kWnTile; r += 8) {
    const int co = co0 + r;
    float ss = 0.f;
    if (co < j.cout) {
      const float* v = j.v + (int64_t)co * per;
      for (int e = lane; e < per; e += 32) ss = fmaf(v[e], v[e], ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) f[r] = co < j.cout ? j.g[co] / sqrtf(ss) : 0.f;
  }
  float* wf = wbase + j.wf_off;
  float* wb = wbase + j.wb_off;
  const int ldf = j.ld_f ? j.ld_f : j.kpad_f;     // row stride of wf (a skip conv is a column block of a wider matrix)
  for (int ci0 = 0; ci0 < j.kpad_f; ci0 += kWnTile) {
    __syncthreads();                             // f ready / previous slab consumed
    const int ncols = min(kWnTile, j.cin - ci0) * taps;            // real floats per row in this slab (<= 0: none)
    for (int r = warp; r < kWnTile; r += 8) {
      const int co = co0 + r;
      const float* v = j.v + (int64_t)co * per + (int64_t)ci0 * taps;
      for (int e = lane; e < kWnTile * taps; e += 32)
        slab[r * pitch + e] = (co < j.cout && e < ncols) ? v[e] : 0.f;
    }
    __syncthreads();
    for (int tap = 0; tap < taps; ++tap) {
      // wf[tap][co0 + r][ci0 + lane]: a warp writes one 128-byte row
      for (int r = warp; r < kWnTile; r += 8) {
        const int co = co0 + r;
        if (co < j.npad_f) {
          const float w = maybe_round(slab[r * pitch + lane * taps + tap] * f[r], rnd);
          const int64_t o = ((int64_t)tap * j.npad_f + co) * ldf + ci0 + lane;
          wf[o] = w;
          if (lo_delta) wf[o + lo_delta] = w - trunc_tf32(w);
        }
      }
      // wb[taps-1-tap][ci0 + i][co0 + lane]
      for (int i = warp; i < kWnTile; i += 8) {
        const int ci = ci0 + i;
        if (ci < j.npad_b) {
          const float w = maybe_round(slab[lane * pitch + i * taps + tap] * f[lane], rnd);
          const int64_t o = ((int64_t)(taps - 1 - tap) * j.npad_b + ci) * j.kpad_b + co0 + lane;
          wb[o] = w;
          if (lo_delta) wb[o + lo_delta] = w - trunc_tf32(w);
        }
      }
    }
  }
}
int k_weightnorm_fwd(const WnJob* jobs_dev, int njobs, int max_cout, float* wbase, int tf32_round,
                     cudaStream_t st, size_t lo_delta) {
  if (njobs == 0) return RNVP_OK;
  // the slab is sized for 3x3 kernels (the only other tap count is 1)
  const int smem = (kWnTile * (kWnTile * 9 + 1) + kWnTile) * (int)sizeof(float);
  weightnorm_fwd_kernel<<<dim3(ceil_div(pad_to(max_cout, 32), kWnTile), njobs), 256, smem, st>>>(jobs_dev, wbase, tf32_round, lo_delta);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
// bias of the fused skip conv: sum of the biases of in_skip and the core_skips (one block per coupling)
__global__ void bias_sum_kernel(const BiasSumJob* __restrict__ jobs, float* __restrict__ wbase) {
  const BiasSumJob& j = jobs[blockIdx.x];
  for (int c = threadIdx.x; c < j.C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < j.n; ++i) s += j.b[i][c];
    wbase[j.out_off + c] = s;
  }
}
int k_bias_sum(const BiasSumJob* jobs_dev, int njobs, float* wbase, cudaStream_t st) {
  if (njobs == 0) return RNVP_OK;
  bias_sum_kernel<<<njobs, 256, 0, st>>>(jobs_dev, wbase);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
// dg = sum dw * v/||v|| ; dv = (g/||v||) * (dw - dg * v/||v||)
__global__ void weightnorm_bwd_kernel(const WnJob* __restrict__ jobs, const float* __restrict__ dwbase) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[8];
  const WnJob j = jobs[blockIdx.y];
  int co = blockIdx.x;
  if (co >= j.cout) return;
  int per = j.cin * j.taps;
  const float* v = j.v + (int64_t)co * per;
  const float* dwf = dwbase + j.dw_off;
  const int lddw = j.ld_dw ? j.ld_dw : j.kpad_f;
  if (j.dbias && threadIdx.x == 0) j.dbias[co] += dwbase[j.dbias_src_off + co];
  float ss = 0.f, dot = 0.f;
  for (int e = threadIdx.x; e < per; e += blockDim.x) {
    int ci = e / j.taps, tap = e % j.taps;
    float vv = v[e];
    ss += vv * vv;
    dot += vv * dwf[((int64_t)tap * j.npad_f + co) * lddw + ci];
  }
  ss = block_sum(ss, sm);
  dot = block_sum(dot, sm);
  float inv = 1.0f / sqrtf(ss);
  float dg = dot * inv;
  float gi = j.g[co] * inv;
  if (j.dg && threadIdx.x == 0) j.dg[co] += dg;
  if (!j.dv) return;
  float* dv = j.dv + (int64_t)co * per;
  for (int e = threadIdx.x; e < per; e += blockDim.x) {
    int ci = e / j.taps, tap = e % j.taps;
    float dw = dwf[((int64_t)tap * j.npad_f + co) * lddw + ci];
    dv[e] += gi * (dw - dg * v[e] * inv);
  }
}
int k_weightnorm_bwd(const WnJob* jobs_dev, int njobs, int max_cout, const float* wbase, const float* dwbase,
                     cudaStream_t st) {
  (void)wbase;
  if (njobs == 0) return RNVP_OK;
  RNVP_CUDA(launch_pdl(weightnorm_bwd_kernel, dim3(max_cout, njobs), dim3(128), 0, st, jobs_dev, dwbase));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// weight_scale = sum p^2 over trainable weight_g and scale (flow_realnvp.py:362-369)
__global__ void sumsq_kernel(const Seg* __restrict__ segs, int nsegs, double* acc) {
  __shared__ float sm[8];
  float a = 0.f;
  for (int s = blockIdx.x; s < nsegs; s += gridDim.x) {
    const Seg sg = segs[s];
    for (int e = threadIdx.x; e < sg.n; e += blockDim.x) a += sg.p[e] * sg.p[e];
  }
  a = block_sum(a, sm);
  if (threadIdx.x == 0) atomicAdd(acc, (double)a);
}
int k_sumsq(const Seg* segs_dev, int nsegs, double* acc, cudaStream_t st) {
  if (nsegs == 0) return RNVP_OK;
  sumsq_kernel<<<nsegs < 296 ? nsegs : 296, kThreads, 0, st>>>(segs_dev, nsegs, acc);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void sumsq_finish_kernel(const double* acc, float* out) { *out = (float)*acc; }
int k_sumsq_finish(const double* acc, float* out, cudaStream_t st) {
  sumsq_finish_kernel<<<1, 1, 0, st>>>(acc, out);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}
__global__ void sumsq_bwd_kernel(const Seg* __restrict__ segs, int nsegs, const float* __restrict__ dws_p) {
  const float dws = *dws_p;
  for (int s = blockIdx.x; s < nsegs; s += gridDim.x) {
    const Seg sg = segs[s];
    if (!sg.g) continue;
    for (int e = threadIdx.x; e < sg.n; e += blockDim.x) sg.g[e] += 2.0f * sg.p[e] * dws;
  }
}
int k_sumsq_bwd(const Seg* segs_dev, int nsegs, const float* dws, cudaStream_t st) {
  if (nsegs == 0) return RNVP_OK;
  sumsq_bwd_kernel<<<nsegs < 296 ? nsegs : 296, kThreads, 0, st>>>(segs_dev, nsegs, dws);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

}  // namespace rnvp
