// fp32 CUDA-core implicit-GEMM convolution (forward / dgrad share one kernel; wgrad is the
// second).  This is the "fp32-accurate" arithmetic tier (1e-5 parity, reconstruction gate);
// the tcgen05 TF32 tier lives in conv_tc.cu and has the same interface and operand layouts.
//
//   forward : y[p,n]      = sum_tap sum_k x[p+tap,k] * w[tap][n][k]  (+bias) (+res)
//   dgrad   : same kernel on dy with the flipped/transposed weights (rnvp_weightnorm_forward's wb)
//   wgrad   : dw[tap][n][k] += sum_p dy[p,n] * x[p+tap,k]
#include "kernels.h"

namespace rnvp {

constexpr int BM = 64, BN = 64, BK = 16, CT = 256;

__device__ __forceinline__ bool tap_pixel(int p, int S, int tap, int taps, int P, int64_t* src) {
  // returns whether pixel p shifted by `tap` stays inside its image; *src = shifted pixel index
  if (p >= P) return false;
  if (taps == 1) {
    *src = p;
    return true;
  }
  int dy = tap / 3 - 1, dx = tap % 3 - 1;
  int j = p % S, i = (p / S) % S;
  if ((unsigned)(i + dy) >= (unsigned)S || (unsigned)(j + dx) >= (unsigned)S) return false;
  *src = (int64_t)p + dy * S + dx;
  return true;
}

__global__ void __launch_bounds__(CT) conv_fwd_fp32_kernel(ConvArgs a) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int P = a.B * a.S * a.S;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  const int lr = t >> 2, lq = t & 3;          // loader: row (pixel / out-channel), k-quad
  const int tx = t & 15, ty = t >> 4;         // compute: 4 columns at tx*4, 4 rows at ty*4
  float acc[4][4] = {};

  for (int tap = 0; tap < a.taps; ++tap) {
    int64_t src = 0;
    const bool pv = tap_pixel(m0 + lr, a.S, tap, a.taps, P, &src);
    const float* xrow = a.x + src * a.kpad;
    const int nrow = n0 + lr;
    const float* wrow = a.w + ((int64_t)tap * a.npad + nrow) * (a.ldw ? a.ldw : a.kpad);
    const bool nv = nrow < a.npad;
    for (int k0 = 0; k0 < a.kpad; k0 += BK) {
      float4 av = pv ? *reinterpret_cast<const float4*>(xrow + k0 + lq * 4) : make_float4(0, 0, 0, 0);
      float4 bv = nv ? *reinterpret_cast<const float4*>(wrow + k0 + lq * 4) : make_float4(0, 0, 0, 0);
      __syncthreads();
      As[lq * 4 + 0][lr] = av.x; As[lq * 4 + 1][lr] = av.y; As[lq * 4 + 2][lr] = av.z; As[lq * 4 + 3][lr] = av.w;
      Bs[lq * 4 + 0][lr] = bv.x; Bs[lq * 4 + 1][lr] = bv.y; Bs[lq * 4 + 2][lr] = bv.z; Bs[lq * 4 + 3][lr] = bv.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 ar = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 br = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float aa[4] = {ar.x, ar.y, ar.z, ar.w}, bb[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }

  // epilogue: bias, residual, store, optional per-channel (sum, sumsq)
  float cs[4] = {}, cq[4] = {};
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int p = m0 + ty * 4 + i;
    if (p >= P) continue;
    float* yrow = a.y + (int64_t)p * a.ldy;
    const float* rrow = a.res ? a.res + (int64_t)p * a.ldy : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = nb + j;
      if (n < a.n) {
        float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f) + (rrow ? rrow[n] : 0.f);
        yrow[n] = v;
        cs[j] += v;
        cq[j] += v * v;
      }
    }
  }
  if (a.stats) {
    __syncthreads();
    float* red = &As[0][0];                       // reuse: [2][16][64] floats = 8 KB <= sizeof(As)+sizeof(Bs)
    float* red2 = &Bs[0][0];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[ty * BN + tx * 4 + j] = cs[j];
      red2[ty * BN + tx * 4 + j] = cq[j];
    }
    __syncthreads();
    if (t < BN && n0 + t < a.n) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        s += red[r * BN + t];
        q += red2[r * BN + t];
      }
      atomicAdd(&a.stats[n0 + t], (double)s);
      atomicAdd(&a.stats[a.n + n0 + t], (double)q);
    }
  }
}

int k_conv_fwd_fp32(const ConvArgs& a, cudaStream_t st) {
  const int P = a.B * a.S * a.S;
  if (P == 0) return RNVP_OK;
  RNVP_REQUIRE(a.kpad % BK == 0, "conv: kpad=%d must be a multiple of %d", a.kpad, BK);
  RNVP_REQUIRE(a.taps == 1 || a.taps == 9, "conv: taps=%d", a.taps);
  RNVP_REQUIRE(a.segs == 1, "the CUDA-core conv takes one input tensor (the runtime loops over K segments)");
  static_assert(sizeof(float) * BK * (BM + 4) >= sizeof(float) * 16 * BN, "stats scratch");
  dim3 grid(ceil_div(P, BM), ceil_div(a.n, BN));
  conv_fwd_fp32_kernel<<<grid, CT, 0, st>>>(a);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

// -------------------------------------------------------------------------------------------
// wgrad: grid.x = n-tiles * k-tiles * taps, grid.y = pixel splits
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) conv_wgrad_fp32_kernel(WgradArgs a, int ntiles, int ktiles, int pix_per_split) {
  __shared__ __align__(16) float As[BK][BN + 4];     // dy chunk  [pixel][n]
  __shared__ __align__(16) float Bs[BK][BN + 4];     // x  chunk  [pixel][k]
  const int P = a.B * a.S * a.S;
  int bx = blockIdx.x;
  const int kt = bx % ktiles; bx /= ktiles;
  const int nt = bx % ntiles; bx /= ntiles;
  const int tap = bx;
  const int n0 = nt * BN, k0 = kt * BN;
  const int p_begin = blockIdx.y * pix_per_split;
  const int p_end = min(P, p_begin + pix_per_split);
  const int t = threadIdx.x;
  const int lr = t >> 4, lq = t & 15;          // loader: pixel row 0..15, column quad 0..15
  const int tx = t & 15, ty = t >> 4;          // compute: k columns tx*4, n rows ty*4
  float acc[4][4] = {};
  float bsum = 0.f;
  const bool do_bias = a.dbias && kt == 0 && tap == 0;

  for (int pc = p_begin; pc < p_end; pc += BK) {
    int p = pc + lr;
    float4 av = make_float4(0, 0, 0, 0), bv = make_float4(0, 0, 0, 0);
    if (p < p_end) {
      int n = n0 + lq * 4;
      if (n < a.lddy) av = *reinterpret_cast<const float4*>(a.dy + (int64_t)p * a.lddy + n);
      int64_t src;
      int k = k0 + lq * 4;
      if (k < a.kpad && tap_pixel(p, a.S, tap, a.taps, P, &src))
        bv = *reinterpret_cast<const float4*>(a.x + src * a.kpad + k);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lr][lq * 4]) = av;
    *reinterpret_cast<float4*>(&Bs[lr][lq * 4]) = bv;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      float4 ar = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
      float4 br = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
      float aa[4] = {ar.x, ar.y, ar.z, ar.w}, bb[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (do_bias && t < BN) {
#pragma unroll
      for (int r = 0; r < BK; ++r) bsum += As[r][t];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= a.n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < a.kpad) atomicAdd(&a.dw[((int64_t)tap * a.npad + n) * (a.lddw ? a.lddw : a.kpad) + k], acc[i][j]);
    }
  }
  if (do_bias && t < BN && n0 + t < a.n) atomicAdd(&a.dbias[n0 + t], bsum);
}

int k_conv_wgrad_fp32(const WgradArgs& a, cudaStream_t st) {
  const int P = a.B * a.S * a.S;
  if (P == 0) return RNVP_OK;
  RNVP_REQUIRE(a.kpad % 4 == 0 && a.lddy % 4 == 0, "wgrad: strides must be multiples of 4");
  RNVP_REQUIRE(a.segs == 1, "the CUDA-core wgrad takes one input tensor (the runtime loops over K segments)");
  int ntiles = ceil_div(a.n, BN), ktiles = ceil_div(a.kpad, BN);
  int base = ntiles * ktiles * a.taps;
  int splits = ceil_div(kNumSMs * 4, base);
  int max_splits = ceil_div(P, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int pps = pad_to(ceil_div(P, splits), BK);
  splits = ceil_div(P, pps);
  conv_wgrad_fp32_kernel<<<dim3(base, splits), CT, 0, st>>>(a, ntiles, ktiles, pps);
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

}  // namespace rnvp
