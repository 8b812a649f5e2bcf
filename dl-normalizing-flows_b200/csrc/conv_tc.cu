// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a, fp32 accumulate, two operand tiers: TF32 (kind::tf32
// on fp32 words that their producers rounded to TF32) and 3xTF32 (split operands, the fp32-accurate tier, "X3" below).
//
//   forward / dgrad :  y[p,n] = sum_tap sum_k x[p+tap,k] * w[tap][n][k]  (+bias) (+res)  [+ BN stats]
//   wgrad           :  dw[tap][n][k] += sum_p dy[p,n] * x[p+tap,k]       (+ dbias[n] += sum_p dy[p,n])
//
// Operands stay fp32 in HBM (NHWC, channel stride padded to 32 floats = one 128-byte swizzle row)
// and are read by the tensor core as TF32 (kind::tf32), so no conversion pass exists.
//
// forward: CTA tile = 128 output pixels x BN channels.  The A tile of one (tap, 32-channel chunk)
// is ONE 5-D TMA box (32 ch, bw, bh, bn, 1 tensor) with bw*bh*bn = 128 whose W/H coordinates are shifted by
// the tap: out-of-image pixels are zero-filled by TMA, so the 3x3 halo costs no instructions and no
// im2col buffer; the fifth coordinate walks the tensors of a K-concatenated input (the fused skip path).
// B is a 3-D box over w[tap][n][k].  Both land in 128B-swizzled, K-major smem and
// feed tcgen05.mma (M=128, N=BN, K=8) x4 per chunk; the accumulator lives in TMEM, double buffered.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), then 4 (BN <= 64, two CTAs per
// SM) or 8 (BN = 128, one CTA per SM) epilogue warps: tcgen05.ld, bias, residual (TMA-prefetched box),
// TMA store from a swizzled staging box, per-channel sum / sum-of-squares for the next batch norm read
// back from that box -- or, as a dgrad, the ReLU mask and the two BN-backward sums -- and, in the XF
// instantiations, four transform warps that rewrite the landed A tiles (BN + ReLU prologue, 3xTF32 split).
//
// wgrad: both operands are MN-major views of the same kind of TMA tiles (pixels are the GEMM K
// dimension): A = tap-shifted x tiles (M = (tap, in channel)), B = dy tile (N = out channels); one CTA
// owns a (n-tile, k-tile, tap-group, pixel-range) slab, keeps up to 512 TMEM columns of partial dw
// and flushes them with fp32 atomics; the bias gradient comes from the staged dy boxes.  A launch may carry a
// group of up to 8 independent wgrads of one shape (blockIdx.z).
#include <cuda.h>
#include <cstdlib>
#include "kernels.h"

namespace rnvp {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(uint32_t src, const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp (warp-uniform result register on every lane: 1 on the elected lane)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Shared-memory accesses through explicit state-space instructions and 32-bit addresses.  The dynamic shared-memory
// pointer loses its address space in the 1 KB alignment arithmetic, and generic LD.E / ST.E are what the compiler
// then emits: slower, and -- since a generic store may alias anything -- serialised load -> compute -> store chains.
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive columns: thread i of the warp gets row (lane base + i), 32 columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (sm_100 version 1), 128-byte swizzle
//   K-major : 8-row x 128B atoms stacked every SBO bytes; LBO unused
//   MN-major: 32 MN-elements x 8 K-rows atoms; next 32 MN-elements at LBO, next 8 K-rows at SBO
//   MN-major 32-bit operands must use the "128B swizzle, 32B atom" layout (type 1): atoms of
//             32 MN-elements x 4 K-rows (rows 128 B apart), next 4 K-rows at SBO, next 32 MN at LBO
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2 /* SWIZZLE_128B */) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
constexpr uint32_t kLayoutSw128Base32 = 1;   // UMMA::LayoutType::SWIZZLE_128B_BASE32B
// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// column sums over the 32 lanes of a warp for 32 per-lane values: lane l ends with sum_lanes v[l]
// (recursive halving: 31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      float keep = hi ? v[j + off] : v[j];
      float send = hi ? v[j] : v[j + off];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------------------------
// forward / dgrad kernel (persistent)
//
// grid = min(#tiles, resident CTAs); every CTA walks tiles t = blockIdx.x, += gridDim.x.  The TMA
// producer runs ahead across tile boundaries through an smem ring; the accumulator is double
// buffered in TMEM so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
// Epilogue (one warp per 32-row TMEM quarter, 32 columns at a time): the residual / skip tile is
// prefetched by TMA into swizzled smem, the accumulator comes from TMEM, bias + residual are added in
// registers (BN partial sums taken there), the result is written to swizzled smem and leaves through
// a TMA store -- every global access of the epilogue is a full-line bulk transfer, and the tensor
// map clips ragged row / channel counts.
//
// XF = 1 ("BN prologue"): the A operand in HBM is the RAW pre-BN activation x; four extra warps rewrite
// every landed A tile in place to h = tf32(relu(x * scale + shift)) before the MMA warp may read it
// (TMA -> full barrier -> transform -> fence.proxy.async -> ready barrier -> tcgen05.mma), so
// BatchNorm2d + ReLU (modules_realnvp.py:83-85, 139-141) cost no pass over the tensor and relu(bn(x))
// never exists in HBM.  The per-channel coefficients are computed by every CTA in its prologue from the
// batch sums the producing conv's epilogue accumulated (or from the running statistics in eval mode);
// CTA 0 also saves (mean, rstd, scale, shift) for the backward pass and updates the running statistics.
// Out-of-image pixels of a 3x3 tap were zero-filled by TMA and must stay zero (the reference pads the
// ACTIVATED tensor), so the transform skips them.
// ---------------------------------------------------------------------------------------------
constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 32 fp32
constexpr int EPI_BOX_BYTES = 32 * 128;  // 32 rows x 32 fp32
constexpr int TC_MAX_STAGES = 8;
constexpr int XF_MAX_K = 512;            // input channels (padded) the BN prologue keeps coefficients for
// BN = 128 runs one CTA per SM: it gets eight epilogue warps (two per TMEM lane quarter, each owning half
// of the column chunks) so that twice as many residual prefetches / result stores are in flight, and a
// three-deep ring to pay for their staging buffers.  The narrower tiles keep four warps (two CTAs per SM).
template <int BN, int XF> struct TcCfg {
  // ring depth cap (the launch takes as many stages as keep MIN_CTAS resident).  The BN prologue adds a pipeline step
  // (TMA -> transform -> MMA), so those kernels want one stage more in flight.
  static constexpr int STAGES = XF ? (BN == 128 ? 5 : (BN == 64 ? 4 : 5)) : 3;
  static constexpr int EPI_WARPS = BN == 128 ? 8 : 4;
  static constexpr int XF_WARPS = XF ? 4 : 0;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * XF_WARPS;
  static constexpr int MIN_CTAS = BN == 128 ? 1 : 2;     // register cap: two CTAs of the narrow tiles per SM
};

struct ConvTcParams {
  const float* bias;
  double* stats;
  const float* bn_save;              // non-null: fused ReLU+BN backward epilogue (tmR maps the raw activations)
  int has_res;
  int round_out;                     // round the result to TF32 (cvt.rna): it is the raw operand of another conv MMA
  const float* post_scale;           // eval mode: y = relu(y * post_scale[n] + post_shift[n]) -- the NEXT layer's batch
  const float* post_shift;           //   norm (fixed affine on running statistics) + ReLU applied where y is produced
  int P, n, taps, kchunks;           // kchunks = segs * kpad / 32
  int segs, kps;                     // K-concatenated input: `segs` tensors of kps 32-channel chunks each (tmA's 5th dimension)
  int S, log2S;
  int m_tiles, n_tiles;
  int stages;                        // depth of the smem ring (<= TC_MAX_STAGES)
  int out_bufs;                      // output staging boxes per epilogue warp (1 or 2); a residual box precedes them if has_res
  int rev;                           // walk the tiles from the last to the first (L2 reuse, see next_sweep_dir)
  // BN prologue (XF kernels)
  int xf_mode;                       // 1 batch statistics from xf_sums, 0 running statistics, 2 coefficients from xf_save
  int xf_C;                          // real input channels
  const double* xf_sums;             // [2C] sum, sum of squares over xf_count values
  double xf_count;
  const float* xf_gamma;
  const float* xf_beta;
  float* xf_rm;
  float* xf_rv;
  float* xf_save;                    // [4C] mean, rstd, scale, shift
  DpXchg xf_xg;                      // data parallel: reduce xf_sums over the ranks in the prologue (mode 1)
  // coupling epilogue (CPL kernels: the s/t net's out conv, n_tiles == 1)
  CplEpilogue cpl;
};

// byte offset of logical 16-byte chunk j of row r inside a 128B-swizzled box
__device__ __forceinline__ uint32_t swz_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// BN coefficients of channels [0, kpad) into smem (scale at [c], shift at [kpad_max + c]); channels >= C get 0 / 0.
// Block `writer` also saves them for the backward pass and updates the running statistics (nn.BatchNorm2d).
// Called by `nthreads` threads numbered `tid` (the transform warps: the loads of the first stages are already in
// flight while they do this).
__device__ __forceinline__ void xf_coefficients(const ConvTcParams& prm, float* s_scale, float* s_shift, int kpad, bool writer,
                                                int tid, int nthreads) {
  const int C = prm.xf_C;
  if (prm.xf_mode == 1) dp_exchange(prm.xf_xg, const_cast<double*>(prm.xf_sums), 2 * C, writer, tid, nthreads, 2);
  for (int c = tid; c < kpad; c += nthreads) {
    float sc = 0.f, sh = 0.f;
    if (c < C) {
      if (prm.xf_mode == 2) {
        sc = prm.xf_save[2 * C + c];
        sh = prm.xf_save[3 * C + c];
      } else {
        const BnCoef k = prm.xf_mode == 1
                             ? bn_coef_from_sums(dp_reduced(prm.xf_xg, prm.xf_sums, c), dp_reduced(prm.xf_xg, prm.xf_sums, C + c),
                                                 prm.xf_count, prm.xf_gamma[c], prm.xf_beta[c])
                             : bn_coef_from_running(prm.xf_rm[c], prm.xf_rv[c], prm.xf_gamma[c], prm.xf_beta[c]);
        sc = k.scale;
        sh = k.shift;
        if (prm.xf_mode == 1 && writer) {
          prm.xf_save[c] = k.mean;
          prm.xf_save[C + c] = k.rstd;
          prm.xf_save[2 * C + c] = k.scale;
          prm.xf_save[3 * C + c] = k.shift;
          const double cnt = prm.xf_count;
          const double unb = cnt > 1.0 ? (double)k.var * cnt / (cnt - 1.0) : (double)k.var;
          prm.xf_rm[c] = (1.f - kBnMomentum) * prm.xf_rm[c] + kBnMomentum * k.mean;
          prm.xf_rv[c] = (1.f - kBnMomentum) * prm.xf_rv[c] + kBnMomentum * (float)unb;
        }
      }
    }
    s_scale[c] = sc;
    s_shift[c] = sh;
  }
}

// X3 = 1 ("3xTF32", the fp32-accurate tensor-core tier; needs XF = 1 for its transform warps): every operand is split
// into hi = its TF32 truncation (what the tensor core reads from an fp32 word anyway, so hi needs no store) and
// lo = x - hi (exact in fp32, <= 13 significant bits), and each K step issues three MMAs, lo*hi + hi*lo + hi*hi, into
// the same fp32 accumulator.  The operand error left is ~2^-21 relative (lo*lo dropped, lo itself truncated to TF32);
// what bounds the tier in practice is the tensor core's fp32 ACCUMULATION, which truncates (measured: rounding hi and lo
// to nearest instead changed neither the per-op error, 1.0e-5 vs 1.1e-5 of max at K = 1440, nor the log-likelihood
// errors, and cost 23 % of the step).  A stage holds [A hi | B hi | A lo | B lo]: A lo is written by the transform
// warps (after the BN prologue, if any -- nothing is rounded in this tier), B lo arrives by TMA from the lo copy of
// the weights that the weight-norm kernel writes next to them (tmBlo).
template <int BN, int XF, int CPL, int X3>
__global__ void __launch_bounds__(TcCfg<BN, XF>::THREADS, TcCfg<BN, XF>::MIN_CTAS)
conv_fwd_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmBlo, const ConvTcParams prm) {
  static_assert(!X3 || XF, "the 3xTF32 split needs the transform warps");
  const int STAGES = prm.stages;
  constexpr int B_TILE_BYTES = BN * 128;
  constexpr int HALF_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr int STAGE_BYTES = X3 ? 2 * HALF_BYTES : HALF_BYTES;
  // The tensor core TRUNCATES its fp32 accumulator after every accumulation step -- a bias towards zero that grows with
  // the number of steps and is what bounds the accuracy of the 3xTF32 tier.  So that tier (a) keeps the two small
  // products (lo*hi, hi*lo) in their own accumulator (its truncation is relative to values 2^-11 times smaller) and
  // (b) deals the hi*hi products of successive chunks round-robin over NACC partial accumulators; the epilogue adds
  // them up in registers (round to nearest).  One buffer = [main 0 | .. | main NACC-1 | lo]; the 128-wide tile then
  // fills the 512 TMEM columns with ONE buffer (no overlap of a tile's epilogue with the next tile's MMAs).
  constexpr int NACC = X3 ? 3 : 1;
  constexpr int ACC_COLS = X3 ? (NACC + 1) * BN : BN;
  constexpr int NBUF = 2 * ACC_COLS <= 512 ? 2 : 1;
  constexpr int TMEM_COLS = NBUF * ACC_COLS < 32 ? 32 : NBUF * ACC_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_a = smem_u32(smem);                            // shared-window address of the ring
  const uint32_t epi_smem = smem_a + (uint32_t)(STAGES * STAGE_BYTES);
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t ready_bar[XF ? TC_MAX_STAGES : 1];   // XF: A tile transformed
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  constexpr int EW = TcCfg<BN, XF>::EPI_WARPS;
  constexpr int CH_PER_WARP = (BN / 32) / (EW / 4) < 1 ? 1 : (BN / 32) / (EW / 4);   // column chunks per epilogue warp
  __shared__ __align__(8) uint64_t res_bar[EW];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_sum[EW][BN];
  __shared__ float red_sq[EW][BN];
  __shared__ __align__(16) float coef[XF ? 2 * XF_MAX_K : 2 * BN];   // XF: BN-prologue (scale | shift); else the fused
                                                                      // BN backward's (scale | shift) when n_tiles == 1

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iters = prm.taps * prm.kchunks;
  const int num_tiles = prm.m_tiles * prm.n_tiles;
  const bool bnbwd = !XF && prm.bn_save != nullptr;
  const bool coef_in_smem = bnbwd && prm.n_tiles == 1;          // else read through L1 from global

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    if (prm.has_res) prefetch_tmap(&tmR);
    if (X3) prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if (XF) mbar_init(&ready_bar[s], TcCfg<BN, XF>::XF_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], EW);       // one arrive per epilogue warp
    }
    for (int s = 0; s < EW; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  if constexpr (CPL) {
    for (int i = threadIdx.x; i < BN; i += blockDim.x) { red_sum[0][i] = 0.f; red_sq[0][i] = 0.f; }
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&tmem_base_slot);
  // everything above overlaps the tail of the previous kernel; from here on its results are needed.
  // The trigger comes AFTER the wait: the successor may then start launching (and run its own prologue on
  // SMs this grid has left), but never more than one kernel ahead -- triggering before the wait lets a whole
  // chain of successors pile up on the SMs holding shared memory and TMEM columns.
  pdl_wait();
  pdl_trigger();
  if constexpr (!XF) if (coef_in_smem) {
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
      const int k = i / BN, cidx = i % BN;
      coef[i] = cidx < prm.n ? prm.bn_save[(2 + k) * prm.n + cidx] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // ring slot and barrier phase advance incrementally: no division on the issue path
      const int hw = prm.S * prm.S;
      int s = 0;
      uint32_t ph = 1;                                   // parity of the "slot free" wait (first pass: free)
      for (int tt = blockIdx.x; tt < num_tiles; tt += gridDim.x) {
        const int t = prm.rev ? num_tiles - 1 - tt : tt;
        const int m_tile = t / prm.n_tiles, n0 = (t % prm.n_tiles) * BN;
        const int p0 = m_tile * 128;
        const int img0 = p0 / hw, row0 = (p0 % hw) / prm.S, col0 = (p0 % hw) % prm.S;
        for (int tap = 0; tap < prm.taps; ++tap) {
          int dy = 0, dx = 0;
          if (prm.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
          int kc = 0;
          for (int seg = 0; seg < prm.segs; ++seg) {
            for (int kk = 0; kk < prm.kps; ++kk, ++kc) {
              mbar_wait(&empty_bar[s], ph);
              uint8_t* a_dst = smem + s * STAGE_BYTES;
              mbar_expect_tx(&full_bar[s], HALF_BYTES + (X3 ? B_TILE_BYTES : 0));
              tma_load_5d(a_dst, &tmA, &full_bar[s], kk * 32, col0 + dx, row0 + dy, img0, seg);
              tma_load_3d(a_dst + A_TILE_BYTES, &tmB, &full_bar[s], kc * 32, n0, tap);
              if (X3) tma_load_3d(a_dst + HALF_BYTES + A_TILE_BYTES, &tmBlo, &full_bar[s], kc * 32, n0, tap);
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop in lock step and one elected lane issues: descriptors and barrier
    // addresses are then warp-uniform values (uniform registers), and the per-MMA work of the issuing
    // thread shrinks to "descriptor + 2 -> tcgen05.mma".  Computed inside an `if (lane == 0)` region every
    // MMA cost ~16 instructions of descriptor arithmetic and uniform-register election, ~100 cycles -- more
    // than the tensor pipe needs for an N <= 128 tile, i.e. the issue thread was the bound.
    constexpr uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
    const uint32_t leader = elect_one();
    const uint64_t a_desc0 = make_desc(smem_u32(smem), 16, 1024);
    const uint64_t b_desc0 = make_desc(smem_u32(smem) + A_TILE_BYTES, 16, 1024);
    int s = 0, ti = 0;
    uint32_t phase = 0;
    uint64_t soff = 0;                                       // (s * STAGE_BYTES) >> 4, added to the address field
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
      const int as = ti % NBUF;
      mbar_wait(&acc_empty[as], ((ti / NBUF) & 1) ^ 1);     // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_COLS);
      for (int it = 0; it < iters; ++it) {
        mbar_wait(XF ? &ready_bar[XF ? s : 0] : &full_bar[s], phase);   // XF: wait for the transformed tile
        tc_fence_after();
        const uint64_t ad = a_desc0 + soff, bd = b_desc0 + soff;
        if (leader) {
          if constexpr (X3) {
            constexpr uint64_t LO = (uint64_t)(HALF_BYTES >> 4);   // address-field distance of the lo tiles
            const uint32_t d_main = d_tmem + (uint32_t)((it % NACC) * BN);                  // partial accumulator
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
              umma_tf32(d_tmem + NACC * BN, ad + LO + k, bd + k, idesc, (it != 0) | (k != 0));     // lo accumulator
              umma_tf32(d_tmem + NACC * BN, ad + k, bd + LO + k, idesc, 1);
              umma_tf32(d_main, ad + k, bd + k, idesc, (it >= NACC) | (k != 0));
            }
          } else {
            umma_tf32(d_tmem, ad, bd, idesc, it != 0);        // 4 x (K = 8 fp32 = 32 bytes) per 128-byte row
            umma_tf32(d_tmem, ad + 2, bd + 2, idesc, 1);
            umma_tf32(d_tmem, ad + 4, bd + 4, idesc, 1);
            umma_tf32(d_tmem, ad + 6, bd + 6, idesc, 1);
          }
          umma_commit(&empty_bar[s]);                         // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        soff += STAGE_BYTES >> 4;
        if (++s == STAGES) { s = 0; soff = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&acc_full[as]);                // accumulator of this tile complete
      __syncwarp();
    }
  } else if (XF && warp >= 2 + EW) {
    // ===================== BN + ReLU transform of the landed A tiles =====================
    // thread t owns logical 16-byte chunk j = t & 7 (channels kc*32 + 4j .. 4j+3) of rows (t >> 3) + 16 i:
    // a warp covers four full 128-byte rows per step (conflict free), and the four coefficients of a thread
    // change only with kc.
    const int t = threadIdx.x - 32 * (2 + EW);
    const int j = t & 7, r0 = t >> 3;
    const int Smask = prm.S - 1;
    // the BN coefficients are only needed here: compute them while the producer's first loads are in flight
    const bool xf_bn = prm.xf_mode != 3;          // mode 3 (3xTF32 tier): no batch norm in front of this conv, split only
    if constexpr (XF) {
      if (xf_bn) xf_coefficients(prm, coef, coef + XF_MAX_K, prm.kchunks * 32, blockIdx.x == 0, t, 32 * TcCfg<BN, XF>::XF_WARPS);
      asm volatile("bar.sync 2, %0;" ::"n"(32 * TcCfg<BN, XF>::XF_WARPS) : "memory");
    }
    int s = 0;
    uint32_t ph = 0;
    for (int tt = blockIdx.x; tt < num_tiles; tt += gridDim.x) {
      const int tl = prm.rev ? num_tiles - 1 - tt : tt;
      const int p0 = (tl / prm.n_tiles) * 128;
      for (int tap = 0; tap < prm.taps; ++tap) {
        int dy = 0, dx = 0;
        if (prm.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        // rows of this thread whose (shifted) pixel lies inside the image and the batch
        uint32_t valid = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = p0 + r0 + 16 * i;
          const int x = (p & Smask) + dx, y = ((p >> prm.log2S) & Smask) + dy;
          if (p < prm.P && (unsigned)x < (unsigned)prm.S && (unsigned)y < (unsigned)prm.S) valid |= 1u << i;
        }
        for (int kc = 0; kc < prm.kchunks; ++kc) {
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (xf_bn) {
            sc = *reinterpret_cast<const float4*>(&coef[kc * 32 + 4 * j]);
            sh = *reinterpret_cast<const float4*>(&coef[XF_MAX_K + kc * 32 + 4 * j]);
          }
          mbar_wait(&full_bar[s], ph);
          const uint32_t a_tile = smem_a + (uint32_t)(s * STAGE_BYTES) + swz_off(r0, j);   // rows r0 + 16 i: + 2048 i
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (valid & (1u << i)) v[i] = lds128(a_tile + 2048u * i);      // (r0 + 16 i) & 7 == r0 & 7: same swizzle
          if constexpr (X3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 lo = make_float4(0.f, 0.f, 0.f, 0.f);          // zero-filled (out-of-image) rows: lo = 0 as well
              if (valid & (1u << i)) {
                float4 w = v[i];
                if (xf_bn) {
                  w.x = fmaxf(fmaf(w.x, sc.x, sh.x), 0.f); w.y = fmaxf(fmaf(w.y, sc.y, sh.y), 0.f);
                  w.z = fmaxf(fmaf(w.z, sc.z, sh.z), 0.f); w.w = fmaxf(fmaf(w.w, sc.w, sh.w), 0.f);
                }
                lo = make_float4(w.x - trunc_tf32(w.x), w.y - trunc_tf32(w.y), w.z - trunc_tf32(w.z), w.w - trunc_tf32(w.w));
                if (xf_bn) sts128(a_tile + 2048u * i, w);            // hi = the value itself: the tensor core truncates it
              }
              sts128(a_tile + (uint32_t)HALF_BYTES + 2048u * i, lo);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (valid & (1u << i)) {
                float4 w = v[i];
                w.x = round_tf32(fmaxf(fmaf(w.x, sc.x, sh.x), 0.f));
                w.y = round_tf32(fmaxf(fmaf(w.y, sc.y, sh.y), 0.f));
                w.z = round_tf32(fmaxf(fmaf(w.z, sc.z, sh.z), 0.f));
                w.w = round_tf32(fmaxf(fmaf(w.w, sc.w, sh.w), 0.f));
                sts128(a_tile + 2048u * i, w);
              }
            }
          }
          fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int ew = warp - 2;                      // epilogue warp index
    const int ci_lo = (ew >> 2) * CH_PER_WARP;    // first column chunk of this warp (two warps share a quarter)
    const int boxes = prm.out_bufs + (prm.has_res ? 1 : 0);
    const uint32_t res_buf = epi_smem + (uint32_t)(ew * boxes * EPI_BOX_BYTES);
    const uint32_t out_buf = res_buf + (prm.has_res ? EPI_BOX_BYTES : 0);   // one or two staging boxes
    float acc_s[CH_PER_WARP], acc_q[CH_PER_WARP]; // per-lane column sums (column c0 + lane), n_tiles == 1
#pragma unroll
    for (int i = 0; i < CH_PER_WARP; ++i) acc_s[i] = acc_q[i] = 0.f;
    const bool keep_stats = prm.stats != nullptr;
    // per-channel sums can stay in registers across tiles when the CTA never changes its n-tile
    const bool own_ntile = prm.n_tiles == 1;
    const int chunks_per_tile = min(BN / 32, ceil_div(prm.n, 32));   // column chunks that hold real channels
    // first tile index >= tt0 (stride gridDim.x) of this CTA in which this warp owns a real column chunk
    auto next_tile_with_work = [&](int tt0) {
      for (; tt0 < num_tiles; tt0 += gridDim.x) {
        const int tl = prm.rev ? num_tiles - 1 - tt0 : tt0;
        if ((tl % prm.n_tiles) * BN + ci_lo * 32 < prm.n) break;
      }
      return tt0;
    };
    auto prefetch_res = [&](int tt_n, int chunk) {
      const int tl = prm.rev ? num_tiles - 1 - tt_n : tt_n;
      mbar_expect_tx(&res_bar[ew], EPI_BOX_BYTES);
      tma_load_2d(res_buf, &tmR, &res_bar[ew], (tl % prm.n_tiles) * BN + chunk * 32, (tl / prm.n_tiles) * 128 + q * 32);
    };
    // residual prefetch of the first (tile, chunk) step of this warp
    int step = 0;
    if (prm.has_res && lane == 0) {
      const int tt0 = next_tile_with_work(blockIdx.x);
      if (tt0 < num_tiles) prefetch_res(tt0, ci_lo);
    }
    int ti = 0;
    for (int tt = blockIdx.x; tt < num_tiles; tt += gridDim.x, ++ti) {
      const int t = prm.rev ? num_tiles - 1 - tt : tt;
      const int as = ti % NBUF;
      const int m_tile = t / prm.n_tiles, n0 = (t % prm.n_tiles) * BN;
      const int prow0 = m_tile * 128 + q * 32;
      const bool pvalid = prow0 + lane < prm.P;
      const int nacc = iters < NACC ? iters : NACC;          // partial accumulators this tile's chunks have written
      (void)nacc;
      mbar_wait(&acc_full[as], (ti / NBUF) & 1);
      tc_fence_after();
      if constexpr (CPL) {
        // ---- the coupling itself, on the accumulator of the s/t net's out conv (modules_realnvp.py:277-301,
        // 339-361): lane = pixel; columns [0,cio) hold t, [cio,2cio) hold l.  s = (scale*tanh(l)+shift)*(1-m),
        // t *= (1-m); forward x' = x*exp(s)+t, inverse x = (y_unbn - t)*exp(-s); per-sample log-det by shuffles.
        const CplEpilogue& cp = prm.cpl;
        const CplGeom& g = cp.g;
        const int p = prow0 + lane;
        const bool ok = p < prm.P;
        const float keep = (ok && g.ckbd) ? (float)(1 - ((g.cfg + (p & (prm.S - 1)) + ((p >> prm.log2S) & (prm.S - 1))) & 1))
                                          : (ok ? 1.f : 0.f);
        const float scale = *cp.scale, sshift = *cp.sshift;
        constexpr int NH = EW / 4;                       // epilogue warps sharing a TMEM lane quarter split the channels
        const int cper = (ceil_div(g.cio, NH) + 7) & ~7;
        const int c_lo = (ew >> 2) * cper, c_hi = min(g.cio, c_lo + cper);
        float ssum = 0.f;
        for (int cb = c_lo; cb < c_hi; cb += 8) {
          float tv[8], lv[8];
          const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS);
          tmem_ld_32x8(tb + (uint32_t)cb, tv);
          tmem_ld_32x8(tb + (uint32_t)(g.cio + cb), lv);
          if constexpr (X3) {
            for (int a = 1; a <= NACC; ++a) {                       // the other partials, then the lo accumulator
              if (a < NACC && a >= nacc) continue;
              float t2[8], l2[8];
              tmem_ld_32x8(tb + (uint32_t)(a * BN + cb), t2);
              tmem_ld_32x8(tb + (uint32_t)(a * BN + g.cio + cb), l2);
#pragma unroll
              for (int j = 0; j < 8; ++j) { tv[j] += t2[j]; lv[j] += l2[j]; }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = cb + j;
            const bool cv = c < c_hi;                    // uniform across the warp
            float outv = 0.f, s = 0.f;
            if (cv && ok) {
              const float t = (tv[j] + prm.bias[c]) * keep;
              const float l = lv[j] + prm.bias[g.cio + c];
              s = (scale * tanhf(l) + sshift) * keep;
              const float xin = cp.x[(int64_t)p * g.C + g.on_off + c];
              if (cp.mode == 3) {                        // reverse=True: un-BN with the RUNNING statistics, then invert
                const float hl = 0.5f * logf(cp.run_var[c] + 1e-5f);
                const float xt = xin * expf(hl * keep) + cp.run_mean[c] * keep;
                outv = (xt - t) * expf(-s);
                cp.out[(int64_t)p * g.C + g.on_off + c] = outv;
                if (!g.ckbd) cp.out[(int64_t)p * g.C + g.in_off + c] = cp.x[(int64_t)p * g.C + g.in_off + c];
                if (cp.logJ) {                           // the reference returns log_rescale here (:283, :302)
                  cp.logJ[(int64_t)p * g.C + g.on_off + c] = s;
                  if (!g.ckbd) cp.logJ[(int64_t)p * g.C + g.in_off + c] = 0.f;
                }
              } else {
                const float xp = xin * expf(s) + t;
                outv = xp;
                if (cp.mode == 1) {                      // training: x' now, out_bn needs its batch statistics first
                  cp.out[(int64_t)p * g.cio + c] = xp;
                  ssum += s;
                } else {                                 // eval: out_bn with the running statistics right here
                  const float rv = cp.run_var[c];
                  const float hl = 0.5f * logf(rv + 1e-5f);
                  const float yn = (xp - cp.run_mean[c]) * (1.0f / sqrtf(rv + kBnEps));
                  cp.out[(int64_t)p * g.C + g.on_off + c] = keep != 0.f ? yn : xp;
                  if (!g.ckbd) cp.out[(int64_t)p * g.C + g.in_off + c] = cp.x[(int64_t)p * g.C + g.in_off + c];
                  const float lj = s - hl * keep;
                  ssum += lj;
                  if (cp.logJ) {
                    cp.logJ[(int64_t)p * g.C + g.on_off + c] = lj;
                    if (!g.ckbd) cp.logJ[(int64_t)p * g.C + g.in_off + c] = 0.f;
                  }
                }
              }
            }
            if (cp.mode == 1 && cv) {                    // batch statistics of x' for out_bn (all positions)
              const float a1 = warp_sum(outv), a2 = warp_sum(outv * outv);
              if (lane == 0) { atomicAdd(&red_sum[0][c], a1); atomicAdd(&red_sq[0][c], a2); }
            }
          }
        }
        if (cp.mode != 3) {
          // per-sample log-det: the 32 pixels of a warp belong to one sample (hw >= 32) or to 32 / hw samples
          const int hw = prm.S * prm.S;
          const int width = hw < 32 ? hw : 32;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            if (o < width) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
          if ((lane & (width - 1)) == 0 && ok) atomicAdd(&cp.logdet_acc[p >> (2 * prm.log2S)], (double)ssum);
        }
      }
      const bool store_st = !CPL || prm.cpl.store_st;
#pragma unroll
      for (int cl = 0; cl < CH_PER_WARP; ++cl) {
        const int ci = ci_lo + cl;
        const int c0 = ci * 32;
        const int nb = n0 + c0;
        if (!store_st || ci >= chunks_per_tile || nb >= prm.n) break;
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS + c0), v);
        if constexpr (X3) {
          for (int a = 1; a <= NACC; ++a) {                         // the other partials, then the lo accumulator
            if (a < NACC && a >= nacc) continue;
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
              float v2[8];
              tmem_ld_32x8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS + a * BN + c0 + j8), v2);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j8 + j] += v2[j];
            }
          }
        }
        if (prm.bias) {
          if (nb + 32 <= prm.n) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = *reinterpret_cast<const float4*>(prm.bias + nb + j);
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (nb + j < prm.n) ? prm.bias[nb + j] : 0.f;
          }
        }
        float w2[XF ? 1 : 32];                     // second statistic of the fused BN backward
        (void)w2;
        if (prm.has_res) {
          mbar_wait(&res_bar[ew], step & 1);
          if (XF || !bnbwd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r4 = lds128(res_buf + swz_off(lane, j));
              v[4 * j] += r4.x; v[4 * j + 1] += r4.y; v[4 * j + 2] += r4.z; v[4 * j + 3] += r4.w;
            }
          } else if constexpr (!XF) {
            // v = dL/dh (dgrad result); the box holds the raw pre-BN activations x of this layer:
            // gm = v * 1[x*scale+shift > 0]; second statistic sum gm*x (the consumer turns the pair
            // (sum gm, sum gm*x) into sum gm*xhat = rstd*(sum gm*x - mean*sum gm) in double)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r4 = lds128(res_buf + swz_off(lane, j));
              float4 sc4, sh4;
              const int cbase = nb + 4 * j;
              if (coef_in_smem) {                   // n0 == 0: column index == channel index
                sc4 = *reinterpret_cast<const float4*>(&coef[c0 + 4 * j]);
                sh4 = *reinterpret_cast<const float4*>(&coef[BN + c0 + 4 * j]);
              } else if (cbase + 4 <= prm.n) {
                sc4 = __ldg(reinterpret_cast<const float4*>(prm.bn_save + 2 * prm.n + cbase));
                sh4 = __ldg(reinterpret_cast<const float4*>(prm.bn_save + 3 * prm.n + cbase));
              } else {
                sc4 = sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              const float xr[4] = {r4.x, r4.y, r4.z, r4.w};
              const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float gm = fmaf(xr[i], sc[i], sh[i]) > 0.f ? v[4 * j + i] : 0.f;
                v[4 * j + i] = gm;
                w2[4 * j + i] = gm * xr[i];
              }
            }
          }
          // the residual box is consumed: prefetch the one of the next step
          __syncwarp();
          if (lane == 0) {
            // next step of THIS warp: its next chunk of the tile, else its first chunk of its next tile
            const int nci = ci + 1;
            if (cl + 1 < CH_PER_WARP && nci < chunks_per_tile && n0 + nci * 32 < prm.n) {
              prefetch_res(tt, nci);
            } else {
              const int ntt = next_tile_with_work(tt + gridDim.x);
              if (ntt < num_tiles) prefetch_res(ntt, ci_lo);
            }
          }
        }
        if (prm.post_scale) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), sh4 = sc4;
            if (nb + j + 4 <= prm.n) {                      // n % 4 == 0 (checked by the launcher)
              sc4 = __ldg(reinterpret_cast<const float4*>(prm.post_scale + nb + j));
              sh4 = __ldg(reinterpret_cast<const float4*>(prm.post_shift + nb + j));
            }
            v[j] = fmaxf(fmaf(v[j], sc4.x, sh4.x), 0.f);
            v[j + 1] = fmaxf(fmaf(v[j + 1], sc4.y, sh4.y), 0.f);
            v[j + 2] = fmaxf(fmaf(v[j + 2], sc4.z, sh4.z), 0.f);
            v[j + 3] = fmaxf(fmaf(v[j + 3], sc4.w, sh4.w), 0.f);
          }
        }
        if (prm.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
        }
        // stage the 32x32 result box (swizzled) and hand it to TMA
        const uint32_t ob = out_buf + (prm.out_bufs == 2 ? (step & 1) * EPI_BOX_BYTES : 0);
        if (lane == 0) {                                  // the store that last used this buffer has read it
          if (prm.out_bufs == 2) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(ob + swz_off(lane, j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(ob, &tmY, nb, prow0);              // rows >= P and columns >= n are clipped by the map
          tma_store_commit();
        }
        ++step;
        if (keep_stats) {
          // per-channel sums straight from the staged (swizzled) box: lane c adds up column c over the 32
          // rows of this warp -- every step reads one 128-byte row across the warp, conflict free.
          // Rows past P were zero-filled by TMA in the inputs but carry the bias: mask them.
          const int rows_valid = min(32, prm.P - prow0);
          const int jc = lane >> 2, wc = (lane & 3) * 4;
          float s1 = 0.f, s2 = 0.f;
          if (XF || !bnbwd) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              float val = lds32(ob + swz_off(r, jc) + wc);
              val = r < rows_valid ? val : 0.f;
              s1 += val;
              s2 = fmaf(val, val, s2);
            }
          } else if constexpr (!XF) {
            // second statistic of the fused BN backward: sum gm * x, x = raw activations still in res_buf
            // (the next residual prefetch targets the same buffer: it was issued after all lanes had
            // read x into registers, so re-reading here would race -- use the register copy instead)
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              float val = lds32(ob + swz_off(r, jc) + wc);
              val = r < rows_valid ? val : 0.f;
              s1 += val;
            }
            float wsum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) wsum[j] = pvalid ? w2[j] : 0.f;
            s2 = warp_transpose_sum(wsum, lane);
          }
          if (nb + lane >= prm.n) { s1 = 0.f; s2 = 0.f; }
          if (own_ntile) {
            acc_s[cl] += s1;
            acc_q[cl] += s2;
          } else if (nb + lane < prm.n) {
            atomicAdd(&prm.stats[nb + lane], (double)s1);
            atomicAdd(&prm.stats[prm.n + nb + lane], (double)s2);
          }
        }
      }
      // this warp is done with the accumulator: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
    }
    if (lane == 0) tma_store_wait_all();
    if constexpr (CPL) {
      if (prm.cpl.mode == 1) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        const int cio = prm.cpl.g.cio;
        for (int c = threadIdx.x - 64; c < cio; c += 32 * EW) {
          atomicAdd(&prm.cpl.sums[c], (double)red_sum[0][c]);
          atomicAdd(&prm.cpl.sums[cio + c], (double)red_sq[0][c]);
        }
      }
    } else
    if (keep_stats && own_ntile) {
      // every warp publishes a full row of BN partial sums (zero outside its own chunks)
      for (int c = lane; c < BN; c += 32) { red_sum[ew][c] = 0.f; red_sq[ew][c] = 0.f; }
      __syncwarp();
#pragma unroll
      for (int cl = 0; cl < CH_PER_WARP; ++cl) {
        red_sum[ew][(ci_lo + cl) * 32 + lane] = acc_s[cl];
        red_sq[ew][(ci_lo + cl) * 32 + lane] = acc_q[cl];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      const int tt = threadIdx.x - 64;
      for (int c = tt; c < BN; c += 32 * EW) {
        if (c < prm.n) {                                     // n_tiles == 1: column index == channel index
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int w = 0; w < EW; ++w) { s1 += red_sum[w][c]; s2 += red_sq[w][c]; }
          atomicAdd(&prm.stats[c], (double)s1);
          atomicAdd(&prm.stats[prm.n + c], (double)s2);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return (EncodeTiledFn) nullptr;
  }();                                            // thread-safe (magic static): autograd's thread may be first
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
    return RNVP_ERR_CUDA;
  }
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                   box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dim0 %llu dim1 %llu box %u x %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return RNVP_ERR_CUDA;
  }
  return RNVP_OK;
}

// pixel box (bw, bh, bn) covering `npix` consecutive NHWC pixels; false if S does not allow it
static bool pixel_box(int S, int npix, int* bw, int* bh, int* bn) {
  if (S <= 0 || (S & (S - 1)) != 0 || S > npix) return false;     // power of two, at most npix wide
  *bw = S;
  *bh = (npix / S) < S ? (npix / S) : S;
  *bn = npix / (*bw * *bh);
  return (*bw) * (*bh) * (*bn) == npix && *bn <= 256;
}

// activation map over `segs` tensors x_i [B][S][S][ld] fp32 lying seg_stride floats apart, with a (32, bw, bh, bn, 1) box
static int make_act_map(CUtensorMap* m, const float* x, int B, int S, int ld, int bw, int bh, int bn,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int segs = 1, size_t seg_stride = 0) {
  if (segs <= 1) { segs = 1; seg_stride = (size_t)B * S * S * ld; }
  cuuint64_t dims[5] = {(cuuint64_t)ld, (cuuint64_t)S, (cuuint64_t)S, (cuuint64_t)B, (cuuint64_t)segs};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)S * ld * 4, (cuuint64_t)S * S * ld * 4, (cuuint64_t)seg_stride * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, 1};
  return encode_map(m, x, 5, dims, strides, box, swz);
}

bool tf32_supported(int S) {
  int a, b, c;
  return pixel_box(S, 128, &a, &b, &c);
}

// 2-D map over a row-major [P, ld] fp32 tensor exposing `n` real columns, 32x32 boxes
static int make_row_map(CUtensorMap* m, const float* base, int P, int n, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)P};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  return encode_map(m, base, 2, dims, strides, box);
}

static int env_int(const char* name, int lo, int hi, int dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v >= lo && v <= hi ? v : dflt;
}

// per-instantiation launch facts, computed once (thread-safe: C++11 magic static)
struct FwdKernelInfo {
  int max_stages = 0, by_regs = 0, static_smem = 0;
  int status = RNVP_OK;
};
template <int BN, int XF, int CPL, int X3>
static const FwdKernelInfo& fwd_kernel_info() {
  static const FwdKernelInfo info = [] {
    FwdKernelInfo k;
    constexpr int THREADS = TcCfg<BN, XF>::THREADS;
    const char* name = XF ? (BN == 128 ? "RNVP_TC_XSTAGES_128" : (BN == 64 ? "RNVP_TC_XSTAGES_64" : "RNVP_TC_XSTAGES_32"))
                          : (BN == 128 ? "RNVP_TC_STAGES_128" : (BN == 64 ? "RNVP_TC_STAGES_64" : "RNVP_TC_STAGES_32"));
    k.max_stages = env_int(name, 1, TC_MAX_STAGES, TcCfg<BN, XF>::STAGES);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_fwd_tf32_kernel<BN, XF, CPL, X3>) != cudaSuccess) { k.status = RNVP_ERR_CUDA; return k; }
    k.by_regs = 65536 / (pad_to(fa.numRegs * 32, 256) * (THREADS / 32));
    k.static_smem = (int)fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(conv_fwd_tf32_kernel<BN, XF, CPL, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024 - k.static_smem) != cudaSuccess ||
        cudaFuncSetAttribute(conv_fwd_tf32_kernel<BN, XF, CPL, X3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared) != cudaSuccess)
      k.status = RNVP_ERR_CUDA;
    return k;
  }();
  return info;
}

template <int BN, int XF, int CPL = 0, int X3 = 0>
static int launch_fwd(const ConvArgs& a, ConvTcParams prm, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmY, tmR, tmBlo;
  int bw = 0, bh = 0, bn = 0;
  pixel_box(a.S, 128, &bw, &bh, &bn);
  RNVP_TRY(make_act_map(&tmA, a.x, a.B, a.S, a.kpad, bw, bh, bn, CU_TENSOR_MAP_SWIZZLE_128B, a.segs, a.seg_stride));
  const int ktot = a.segs * a.kpad, ldw = a.ldw ? a.ldw : ktot;
  cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)a.npad, (cuuint64_t)a.taps};
  cuuint64_t strides[2] = {(cuuint64_t)ldw * 4, (cuuint64_t)a.npad * ldw * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
  RNVP_TRY(encode_map(&tmB, a.w, 3, dims, strides, box));
  if (X3) RNVP_TRY(encode_map(&tmBlo, a.w + a.w_lo_delta, 3, dims, strides, box));
  else tmBlo = tmB;
  const float* rsrc = a.bn_x ? a.bn_x : a.res;
  RNVP_TRY(make_row_map(&tmY, a.y, prm.P, a.n, a.ldy));
  if (rsrc) RNVP_TRY(make_row_map(&tmR, rsrc, prm.P, a.n, a.ldy));
  else tmR = tmY;
  constexpr int THREADS = TcCfg<BN, XF>::THREADS;
  const FwdKernelInfo& ki = fwd_kernel_info<BN, XF, CPL, X3>();
  RNVP_REQUIRE(ki.status == RNVP_OK, "conv: cudaFuncGetAttributes / cudaFuncSetAttribute failed");
  // shared-memory plan: epilogue staging (a residual box only when the layer has one; one output box per warp
  // for the 32-wide tile, whose warps stage one box per tile) + as many ring stages as keep MIN_CTAS resident
  prm.out_bufs = env_int("RNVP_TC_OUTBUFS", 1, 2, (BN == 32 || X3) ? 1 : 2);
  const int EPI = TcCfg<BN, XF>::EPI_WARPS * (prm.out_bufs + (prm.has_res ? 1 : 0)) * EPI_BOX_BYTES;
  const int stage_bytes = (X3 ? 2 : 1) * (A_TILE_BYTES + BN * 128);          // 3xTF32: hi and lo tiles
  int budget = (228 * 1024) / TcCfg<BN, XF>::MIN_CTAS - 1024 /* driver */ - ki.static_smem - 1024 /* alignment */;
  if ((budget - EPI) / stage_bytes < 2)                                       // the doubled stages want the whole SM
    budget = 228 * 1024 - 1024 - ki.static_smem - 1024;
  int stages = (budget - EPI) / stage_bytes;
  if (stages > ki.max_stages) stages = ki.max_stages;
  if (stages < 2) stages = 2;
  prm.stages = stages;
  const int smem = stages * stage_bytes + EPI + 1024;
  RNVP_REQUIRE(smem + ki.static_smem <= 227 * 1024, "conv: shared-memory plan of %d bytes does not fit", smem);
  // resident CTAs per SM from the kernel's own footprint: shared memory (dynamic + static + 1 KB the driver
  // reserves per CTA) against the 228 KB of an SM, registers against the 64 K file, TMEM columns
  // (cudaOccupancyMaxActiveBlocksPerMultiprocessor under-reports this kernel: it answered 1 where 2 fit)
  int ctas = (228 * 1024) / (smem + ki.static_smem + 1024);
  if (ctas > ki.by_regs) ctas = ki.by_regs;
  constexpr int acc_cols = X3 ? 4 * BN : BN;                 // kernel: [3 partial mains | lo] per buffer in the 3xTF32 tier
  constexpr int tmem_cols = (2 * acc_cols <= 512 ? 2 : 1) * acc_cols < 32 ? 32 : (2 * acc_cols <= 512 ? 2 : 1) * acc_cols;
  if (ctas > 512 / tmem_cols) ctas = 512 / tmem_cols;
  if (ctas < 1) ctas = 1;
  ctas = env_int("RNVP_TC_CTAS", 1, 8, ctas);
  if (getenv("RNVP_DEBUG")) {
    static bool said[2][3] = {};
    bool& s = said[XF][BN == 32 ? 0 : (BN == 64 ? 1 : 2)];
    if (!s) {
      s = true;
      fprintf(stderr, "[rnvp] conv_fwd_tf32<%d,%d>: %d stages, smem %d+%d, regs -> %d, => %d CTAs/SM\n", BN, XF, stages,
              smem, ki.static_smem, ki.by_regs, ctas);
    }
  }
  prm.m_tiles = ceil_div(prm.P, 128);
  prm.n_tiles = ceil_div(a.n, BN);
  int tiles = prm.m_tiles * prm.n_tiles;
  int grid = kNumSMs * ctas;
  if (grid > tiles) grid = tiles;
  RNVP_CUDA(launch_pdl(conv_fwd_tf32_kernel<BN, XF, CPL, X3>, dim3(grid), dim3(THREADS), (size_t)smem, st, tmA, tmB, tmY, tmR, tmBlo, prm));
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

bool conv_tf32_fusable(const ConvArgs& a) {
  int bw, bh, bn;
  return pixel_box(a.S, 128, &bw, &bh, &bn) && a.kpad % 32 == 0 && a.ldy % 4 == 0 && a.n <= 512;
}
// the BN prologue additionally keeps 2 * kpad coefficients in shared memory
bool conv_tf32_prologue_ok(const ConvArgs& a) { return conv_tf32_fusable(a) && a.kpad <= XF_MAX_K; }

int k_conv_fwd_tf32(const ConvArgs& a, cudaStream_t st) {
  const int P = a.B * a.S * a.S;
  if (P == 0) return RNVP_OK;
  int bw, bh, bn;
  if (!pixel_box(a.S, 128, &bw, &bh, &bn) || a.kpad % 32 != 0 || a.ldy % 4 != 0) {
    RNVP_REQUIRE(a.bn_x == nullptr && a.xf == nullptr && a.post_scale == nullptr && a.segs == 1,
                 "fused BN prologue / epilogues / K-concatenated inputs need the tensor-core kernel");
    return k_conv_fwd_fp32(a, st);          // shapes the TMA box cannot express: CUDA-core kernel
  }
  RNVP_REQUIRE(a.bn_x == nullptr || (a.res == nullptr && a.bias == nullptr && a.n % 4 == 0 && a.bn_save && a.stats),
               "fused BN-backward epilogue: needs stats and bn_save, excludes bias / residual, n % 4 == 0");
  RNVP_REQUIRE(a.taps == 1 || a.taps == 9, "conv: taps=%d", a.taps);
  RNVP_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.y & 15) == 0,
               "conv operands must be 16-byte aligned");
  ConvTcParams prm{};
  prm.bias = a.bias; prm.has_res = a.res != nullptr || a.bn_x != nullptr; prm.stats = a.stats;
  prm.bn_save = a.bn_x ? a.bn_save : nullptr;
  prm.round_out = a.round_out;
  RNVP_REQUIRE((a.post_scale == nullptr) == (a.post_shift == nullptr) && (a.post_scale == nullptr || (a.n % 4 == 0 && a.bn_x == nullptr)),
               "conv: the post-affine epilogue needs scale and shift, n % 4 == 0 and no fused BN backward");
  prm.post_scale = a.post_scale; prm.post_shift = a.post_shift;
  RNVP_REQUIRE(a.segs >= 1 && (a.segs == 1 || (a.xf == nullptr && a.seg_stride % 4 == 0 && (a.ldw == 0 || a.ldw % 4 == 0))),
               "conv: K-concatenated input needs 16-byte aligned segments and excludes the BN prologue");
  prm.P = P; prm.n = a.n; prm.taps = a.taps; prm.kchunks = a.segs * (a.kpad / 32);
  prm.segs = a.segs; prm.kps = a.kpad / 32;
  prm.S = a.S;
  prm.log2S = 0;
  while ((1 << prm.log2S) < a.S) ++prm.log2S;
  prm.rev = next_sweep_dir();
  if (a.x3) {
    // fp32-accurate tier: 3xTF32.  Every conv runs the transform-warp kernel (the BN prologue, or a plain split).
    RNVP_REQUIRE(a.bn_x == nullptr && a.round_out == 0 && a.w_lo_delta != 0,
                 "3xTF32: no fused BN-backward epilogue, no operand rounding, and the lo copy of the weights is needed");
    RNVP_REQUIRE(a.segs * a.kpad <= XF_MAX_K || a.xf == nullptr, "3xTF32 BN prologue: too many input channels");
    prm.xf_mode = 3;
    if (a.xf) {
      const BnPrologue& x = *a.xf;
      RNVP_REQUIRE(x.mode == 2 ? x.save != nullptr : (x.gamma && x.beta && x.run_mean && x.run_var), "BN prologue: missing coefficient source");
      RNVP_REQUIRE(x.mode != 1 || (x.sums && x.save), "BN prologue: training mode needs sums and save");
      prm.xf_mode = x.mode; prm.xf_C = x.C; prm.xf_sums = x.sums; prm.xf_count = x.count;
      prm.xf_gamma = x.gamma; prm.xf_beta = x.beta; prm.xf_rm = x.run_mean; prm.xf_rv = x.run_var; prm.xf_save = x.save;
      prm.xf_xg = x.xg;
    }
    if (a.cpl && a.cpl->mode) {
      RNVP_REQUIRE(a.xf && a.n <= 128 && a.bias && !a.res && a.n == 2 * a.cpl->g.cio, "coupling epilogue: bad out conv");
      prm.cpl = *a.cpl;
      if (a.n <= 32) return launch_fwd<32, 1, 1, 1>(a, prm, st);
      if (a.n <= 64) return launch_fwd<64, 1, 1, 1>(a, prm, st);
      return launch_fwd<128, 1, 1, 1>(a, prm, st);
    }
    if (a.n <= 32) return launch_fwd<32, 1, 0, 1>(a, prm, st);
    if (a.n <= 64) return launch_fwd<64, 1, 0, 1>(a, prm, st);
    return launch_fwd<128, 1, 0, 1>(a, prm, st);
  }
  if (a.xf) {
    const BnPrologue& x = *a.xf;
    RNVP_REQUIRE(a.bn_x == nullptr, "BN prologue and BN-backward epilogue are separate kernels");
    RNVP_REQUIRE(a.kpad <= XF_MAX_K && x.C <= a.kpad, "BN prologue: %d input channels unsupported (max %d)", x.C, XF_MAX_K);
    RNVP_REQUIRE(x.mode == 2 ? x.save != nullptr : (x.gamma && x.beta && x.run_mean && x.run_var),
                 "BN prologue: missing coefficient source");
    RNVP_REQUIRE(x.mode != 1 || (x.sums && x.save), "BN prologue: training mode needs sums and save");
    prm.xf_mode = x.mode; prm.xf_C = x.C; prm.xf_sums = x.sums; prm.xf_count = x.count;
    prm.xf_gamma = x.gamma; prm.xf_beta = x.beta; prm.xf_rm = x.run_mean; prm.xf_rv = x.run_var; prm.xf_save = x.save;
    prm.xf_xg = x.xg;
    if (a.cpl && a.cpl->mode) {
      RNVP_REQUIRE(a.n <= 128 && a.bias && !a.res && a.n == 2 * a.cpl->g.cio,
                   "coupling epilogue: the out conv must have a bias, no residual and 2*cio <= 128 outputs");
      prm.cpl = *a.cpl;
      if (a.n <= 32) return launch_fwd<32, 1, 1>(a, prm, st);
      if (a.n <= 64) return launch_fwd<64, 1, 1>(a, prm, st);
      return launch_fwd<128, 1, 1>(a, prm, st);
    }
    if (a.n <= 32) return launch_fwd<32, 1>(a, prm, st);
    if (a.n <= 64) return launch_fwd<64, 1>(a, prm, st);
    return launch_fwd<128, 1>(a, prm, st);
  }
  RNVP_REQUIRE(a.cpl == nullptr || a.cpl->mode == 0, "the coupling epilogue rides on the BN-prologue kernel");
  if (a.n <= 32) return launch_fwd<32, 0>(a, prm, st);
  if (a.n <= 64) return launch_fwd<64, 0>(a, prm, st);
  return launch_fwd<128, 0>(a, prm, st);
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel:  dw[tap][n][k] += sum_p dy[p,n] * x[p+tap,k]   (+ dbias[n] += sum_p dy[p,n])
//
// GEMM view per 64-pixel tile:  D[(tap,k)][n] += X_tap[(tap,k)][pixels] * DY[n][pixels]^T with BOTH
// operands MN-major: the TMA tiles are (32 channels x 64 pixels) boxes whose 128-byte rows are pixels,
// i.e. the GEMM K dimension is the strided one (layout "128B swizzle, 32B atom", mandatory for MN-major
// 32-bit operands).  The M dimension packs (tap, input channel): with <=32 input channels four
// tap-shifted x tiles fill one M=128 operand, so narrow layers do not waste tensor-core rows.
// One CTA owns (k-tile, n-tile, tap range, pixel range): per pixel tile it loads the dy boxes once,
// then one 4-box A stage per M-group, issues 8 x (K = 8 pixels) MMAs per group into that group's
// TMEM columns, and finally flushes the accumulators with coalesced fp32 atomics.  The bias gradient
// (column sums of dy) is taken from the staged dy boxes in shared memory by the epilogue warps, which
// are otherwise idle during the main loop -- it costs no tensor-core work and no extra traffic.
// ---------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 192;               // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue (+ BN transform of the x boxes)
constexpr int WG_BOX_BYTES = 64 * 128;        // 64 pixels x 32 fp32
constexpr int WG_MAX_STAGES = 16;              // ring depths are chosen per launch from the actual stage sizes
constexpr int WG_RING_BYTES = 192 * 1024;     // A ring + B ring
constexpr int WG_SMEM = WG_RING_BYTES + 1024;
constexpr int WG_MAX_GROUPS = 16;

struct WgradTcParams {
  float* dw;
  float* dbias;
  int P, n, npad, kpad, taps, S;            // kpad = segs * seg_k (all input channels of the K-concatenated x)
  int seg_k, lddw;                          // channels per x tensor (tmX's 5th dimension picks the tensor); row stride of dw
  int n_tiles, k_tiles, tap_ranges, taps_per_range;
  int kb;                                   // x boxes (32 channels) per tap inside a k-tile (full tiles)
  int tpm;                                  // taps packed into one M = 128 operand
  int tiles_per_split, num_tiles;           // 64-pixel tiles
  int tmem_cols;
  int a_stages, b_stages, a_stage_bytes, b_stage_bytes;
  int ring_bytes;                           // A ring + MMA slack + B ring
  int a_lbo;                                // byte stride between the 32-row groups of the A operand
  // BN prologue: x in HBM is the raw pre-BN activation of the forward conv; the epilogue warps (idle during the
  // main loop) rewrite every landed x box to tf32(relu(x * scale + shift)) with the coefficients the forward saved
  const float* xf_save;                     // [4C] mean, rstd, scale, shift, or null
  int xf_C, log2S;
  // 3xTF32 (fp32-accurate tier): the epilogue warps split every landed x and dy box into hi (in place: the tensor
  // core truncates the fp32 word itself) and lo = v - trunc(v), written a_half / b_half bytes further into the stage;
  // three MMAs per K step (lo*hi, hi*lo, hi*hi)
  int x3, a_half, b_half;
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// J > 1: a GROUP of J independent weight gradients of one shape in one launch (blockIdx.z = job): the 2R 1x1 and the
// R 3x3 wgrads of a coupling's residual blocks.  At the deep scales a launch has 2-6 us of work under an 8 us launch
// floor, and whatever runs there runs alone (every kernel fills the SMs' shared memory), so launches saved are time saved.
template <int J> struct WgOperands {
  CUtensorMap m[2 * J];                     // (dy, x) maps of job j at [2j], [2j + 1]
  float* dw[J];
  float* dbias[J];
  const float* xf_save[J];
};
template <int J>
__global__ void __launch_bounds__(TC_THREADS) conv_wgrad_tf32_kernel(const __grid_constant__ WgOperands<J> ops,
                                                                     const WgradTcParams prm_in) {
  const int job = J > 1 ? (int)blockIdx.z : 0;
  const CUtensorMap& tmDy = ops.m[2 * job];
  const CUtensorMap& tmX = ops.m[2 * job + 1];
  WgradTcParams prm = prm_in;
  prm.dw = ops.dw[job]; prm.dbias = ops.dbias[job]; prm.xf_save = ops.xf_save[job];
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int WG_A_STAGES = prm.a_stages, WG_B_STAGES = prm.b_stages;
  const int WG_A_STAGE_BYTES = prm.a_stage_bytes, WG_B_STAGE_BYTES = prm.b_stage_bytes;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + prm.ring_bytes - WG_B_STAGES * WG_B_STAGE_BYTES;
  __shared__ __align__(8) uint64_t a_full[WG_MAX_STAGES], a_empty[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t b_full[WG_MAX_STAGES], b_empty[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t a_ready[WG_MAX_STAGES];            // BN prologue / 3xTF32: x boxes of the stage transformed
  __shared__ __align__(8) uint64_t b_ready[WG_MAX_STAGES];            // 3xTF32: dy boxes of the stage split
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float wcoef[2][128];                       // (scale | shift) of this CTA's k-tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool xf = prm.xf_save != nullptr;
  const bool x3 = prm.x3 != 0;
  int bx = blockIdx.x;
  const int kt = bx % prm.k_tiles; bx /= prm.k_tiles;
  const int nt = bx % prm.n_tiles; bx /= prm.n_tiles;
  const int tr = bx;
  const int n0 = nt * 128, k0 = kt * 128;
  const int tap0 = tr * prm.taps_per_range;
  const int ntap = min(prm.taps_per_range, prm.taps - tap0);
  const int kb = min(prm.kb, (prm.kpad - k0) / 32);                     // x boxes per tap in this k-tile
  const int nbx = min(4, ceil_div(pad_to(prm.n, 32) - n0, 32));          // dy boxes (32 channels each)
  const int N = nbx * 32;
  const int groups = ceil_div(ntap, prm.tpm);
  const int t_begin = blockIdx.y * prm.tiles_per_split;
  const int t_end = min(prm.num_tiles, t_begin + prm.tiles_per_split);
  const bool do_bias = prm.dbias != nullptr && kt == 0 && tr == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmX);
    for (int s = 0; s < WG_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&a_ready[s], 4); }
    // a dy stage is released by the MMA commit and, when this CTA owns the bias gradient, by the four
    // epilogue warps that read the boxes for the column sums
    // (3xTF32: the warps hand the stage to the MMA warp through b_ready instead and never touch it afterwards)
    for (int s = 0; s < WG_B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], (do_bias && !x3) ? 5 : 1);
      mbar_init(&b_ready[s], 4);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(&tmem_base_slot, (uint32_t)prm.tmem_cols);
  pdl_wait();                                // the prologue above overlaps the previous kernel's tail
  pdl_trigger();                             // after the wait: at most one successor in flight (see conv kernel)
  if (xf)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      const int c = k0 + (i & 127);
      wcoef[i >> 7][i & 127] = c < prm.xf_C ? prm.xf_save[(2 + (i >> 7)) * prm.xf_C + c] : 0.f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (t_begin < t_end) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        // ring slots, barrier phases and the pixel coordinates advance incrementally (no division per tile)
        const int hw = prm.S * prm.S;
        int p0 = t_begin * 64;
        int img0 = p0 / hw, rem = p0 % hw;
        int as = 0, bs = 0;
        uint32_t aph = 1, bph = 1;                        // parity of the "slot free" waits (first pass: free)
        int xch[4], xseg[4];                              // channel offset / tensor of x box j (K-concatenated x)
#pragma unroll
        for (int j = 0; j < 4; ++j) { xseg[j] = (k0 + 32 * j) / prm.seg_k; xch[j] = (k0 + 32 * j) % prm.seg_k; }
        for (int t = t_begin; t < t_end; ++t) {
          const int row0 = rem / prm.S, col0 = rem % prm.S;
          mbar_wait(&b_empty[bs], bph);
          mbar_expect_tx(&b_full[bs], nbx * WG_BOX_BYTES);
          for (int g = 0; g < nbx; ++g)
            tma_load_5d(b_ring + bs * WG_B_STAGE_BYTES + g * WG_BOX_BYTES, &tmDy, &b_full[bs], n0 + 32 * g, col0, row0, img0, 0);
          if (++bs == WG_B_STAGES) { bs = 0; bph ^= 1; }
          for (int mg = 0; mg < groups; ++mg) {
            const int tl_n = min(prm.tpm, ntap - mg * prm.tpm);       // taps in this M-group
            mbar_wait(&a_empty[as], aph);
            mbar_expect_tx(&a_full[as], tl_n * kb * WG_BOX_BYTES);
            for (int tl = 0; tl < tl_n; ++tl) {
              const int tap = tap0 + mg * prm.tpm + tl;
              int dy = 0, dx = 0;
              if (prm.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < kb)
                  tma_load_5d(a_ring + as * WG_A_STAGE_BYTES + (tl * kb + j) * WG_BOX_BYTES, &tmX, &a_full[as],
                              xch[j], col0 + dx, row0 + dy, img0, xseg[j]);
            }
            if (++as == WG_A_STAGES) { as = 0; aph ^= 1; }
          }
          rem += 64;
          while (rem >= hw) { rem -= hw; ++img0; }       // a 64-pixel tile spans several images when S < 8
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // whole warp in lock step, one elected lane issues, descriptors are warp-uniform values that only change
      // by an address increment per MMA (see the forward kernel: the issue thread was the bound)
      {
        const uint32_t idesc = make_idesc_tf32(128, N, 1, 1);
        const uint32_t leader = elect_one();
        const uint64_t a_desc0 = make_desc(smem_u32(a_ring), (uint32_t)prm.a_lbo, 512, kLayoutSw128Base32);
        const uint64_t b_desc0 = make_desc(smem_u32(b_ring), WG_BOX_BYTES, 512, kLayoutSw128Base32);
        const uint64_t a_step = (uint64_t)(WG_A_STAGE_BYTES >> 4), b_step = (uint64_t)(WG_B_STAGE_BYTES >> 4);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        uint64_t aoff = 0, boff = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(x3 ? &b_ready[bs] : &b_full[bs], bph);
          tc_fence_after();
          const uint64_t bd = b_desc0 + boff;
          const uint32_t acc = (t != t_begin);
          const uint64_t alo = (uint64_t)(prm.a_half >> 4), blo = (uint64_t)(prm.b_half >> 4);
          for (int mg = 0; mg < groups; ++mg) {
            mbar_wait((xf || x3) ? &a_ready[as] : &a_full[as], aph);
            tc_fence_after();
            const uint64_t ad = a_desc0 + aoff;
            const uint32_t d = tmem_base + (uint32_t)(mg * N);
            if (leader) {
              // 8 x (K = 8 pixels = two 512-byte swizzle atoms = 1024 bytes -> +64 in the address field)
              if (x3) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                  umma_tf32(d, ad + alo + 64 * ks, bd + 64 * ks, idesc, acc | (uint32_t)(ks != 0));
                  umma_tf32(d, ad + 64 * ks, bd + blo + 64 * ks, idesc, 1);
                  umma_tf32(d, ad + 64 * ks, bd + 64 * ks, idesc, 1);
                }
              } else {
                umma_tf32(d, ad, bd, idesc, acc);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks) umma_tf32(d, ad + 64 * ks, bd + 64 * ks, idesc, 1);
              }
              umma_commit(&a_empty[as]);
            }
            __syncwarp();
            aoff += a_step;
            if (++as == WG_A_STAGES) { as = 0; aoff = 0; aph ^= 1; }
          }
          if (leader) umma_commit(&b_empty[bs]);
          __syncwarp();
          boff += b_step;
          if (++bs == WG_B_STAGES) { bs = 0; boff = 0; bph ^= 1; }
        }
        if (leader) umma_commit(&acc_bar);
        __syncwarp();
      }
    } else {
      // ===================== epilogue: TMEM -> coalesced fp32 atomics =====================
      const int q = warp & 3;
      const int m = q * 32 + lane;               // accumulator row = (tap_local, k)
      const int kc = kb * 32;
      const int tl = m / kc, k = m % kc;
      if (do_bias || xf || x3) {
        // Main-loop duties of the otherwise idle epilogue warps, tile by tile in the producer's order:
        //  (a) dbias[n] = sum_p dy[p,n]: warp q owns dy box q (32 channels x 64 pixel rows of 128 bytes).  Within a
        //      row the four 32-byte chunks are XOR-permuted by (row & 3) ("128B swizzle, 32B atom" = Swizzle<2,5,2>),
        //      so lane w accumulates word w separately per row phase and un-permutes at the end.
        //  (b) BN prologue: every landed x box -> tf32(relu(x * scale + shift)) in place.  Thread e owns physical
        //      16-byte chunk e & 7 of rows (e >> 3) + 16 i: the row phase (row & 3) is fixed per thread, hence so are
        //      its four logical channels within a box.  Rows whose tap-shifted pixel lies outside the image were
        //      zero-filled by TMA and stay zero (the forward conv pads the ACTIVATED tensor).
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        const int e = threadIdx.x - 64;
        const int c16 = e & 7, er0 = e >> 3;
        const int lch = 8 * ((c16 >> 1) ^ (er0 & 3)) + 4 * (c16 & 1);     // logical channel (within a box) of this chunk
        const int Smask = prm.S - 1;
        const uint32_t a_ring_a = smem_u32(a_ring), b_ring_a = smem_u32(b_ring);
        const uint32_t my_off = (uint32_t)(er0 * 128 + c16 * 16);         // this thread's chunk in rows er0 + 16 i (+ 2048 i)
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        for (int t = t_begin; t < t_end; ++t) {
          if (do_bias || x3) {
            mbar_wait(&b_full[bs], bph);
            if (do_bias && q < nbx) {
              const uint32_t box = b_ring_a + (uint32_t)(bs * WG_B_STAGE_BYTES + q * WG_BOX_BYTES) + 4u * lane;
#pragma unroll
              for (int r0 = 0; r0 < 64; r0 += 16) {
                float v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = lds32(box + 128u * (r0 + r));
#pragma unroll
                for (int r = 0; r < 16; ++r) part[r & 3] += v[r];
              }
            }
            if (x3) {
              // dy boxes stay as they are (hi = what the tensor core reads of them); lo = v - trunc(v) goes next to them
              // (rows past the tensor were zero-filled: lo = 0)
              const uint32_t bst = b_ring_a + (uint32_t)(bs * WG_B_STAGE_BYTES) + my_off;
              for (int g = 0; g < nbx; ++g) {
                float4 w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = lds128(bst + (uint32_t)(g * WG_BOX_BYTES) + 2048u * i);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  sts128(bst + (uint32_t)prm.b_half + (uint32_t)(g * WG_BOX_BYTES) + 2048u * i,
                         make_float4(w[i].x - trunc_tf32(w[i].x), w[i].y - trunc_tf32(w[i].y), w[i].z - trunc_tf32(w[i].z),
                                     w[i].w - trunc_tf32(w[i].w)));
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(&b_ready[bs]);
            } else {
              __syncwarp();
              if (lane == 0) mbar_arrive(&b_empty[bs]);
            }
            if (++bs == WG_B_STAGES) { bs = 0; bph ^= 1; }
          }
          if (xf || x3) {
            for (int mg = 0; mg < groups; ++mg) {
              const int tl_n = min(prm.tpm, ntap - mg * prm.tpm);
              mbar_wait(&a_full[as], aph);
              const uint32_t stage = a_ring_a + (uint32_t)(as * WG_A_STAGE_BYTES) + my_off;
              for (int tli = 0; tli < tl_n; ++tli) {
                const int tap = tap0 + mg * prm.tpm + tli;
                int dy = 0, dx = 0;
                if (prm.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
                bool valid[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int p = t * 64 + er0 + 16 * i;
                  const int x = (p & Smask) + dx, y = ((p >> prm.log2S) & Smask) + dy;
                  valid[i] = p < prm.P && (unsigned)x < (unsigned)prm.S && (unsigned)y < (unsigned)prm.S;
                }
                for (int j = 0; j < kb; ++j) {
                  const uint32_t box = stage + (uint32_t)((tli * kb + j) * WG_BOX_BYTES);
                  const float4 sc = *reinterpret_cast<const float4*>(&wcoef[0][32 * j + lch]);
                  const float4 sh = *reinterpret_cast<const float4*>(&wcoef[1][32 * j + lch]);
                  float4 w[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    if (valid[i]) w[i] = lds128(box + 2048u * i);
                  if (x3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f);        // zero-filled rows stay zero in both halves
                      if (valid[i]) {
                        if (xf) {
                          w[i].x = fmaxf(fmaf(w[i].x, sc.x, sh.x), 0.f); w[i].y = fmaxf(fmaf(w[i].y, sc.y, sh.y), 0.f);
                          w[i].z = fmaxf(fmaf(w[i].z, sc.z, sh.z), 0.f); w[i].w = fmaxf(fmaf(w[i].w, sc.w, sh.w), 0.f);
                        }
                        if (xf) sts128(box + 2048u * i, w[i]);
                        lo = make_float4(w[i].x - trunc_tf32(w[i].x), w[i].y - trunc_tf32(w[i].y), w[i].z - trunc_tf32(w[i].z),
                                         w[i].w - trunc_tf32(w[i].w));
                      }
                      sts128(box + (uint32_t)prm.a_half + 2048u * i, lo);
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      if (valid[i]) {
                        w[i].x = round_tf32(fmaxf(fmaf(w[i].x, sc.x, sh.x), 0.f));
                        w[i].y = round_tf32(fmaxf(fmaf(w[i].y, sc.y, sh.y), 0.f));
                        w[i].z = round_tf32(fmaxf(fmaf(w[i].z, sc.z, sh.z), 0.f));
                        w[i].w = round_tf32(fmaxf(fmaf(w[i].w, sc.w, sh.w), 0.f));
                        sts128(box + 2048u * i, w[i]);
                      }
                    }
                  }
                }
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[as]);
              if (++as == WG_A_STAGES) { as = 0; aph ^= 1; }
            }
          }
        }
        if (do_bias && q < nbx) {
#pragma unroll
          for (int ph = 0; ph < 4; ++ph) {
            const int ch = n0 + q * 32 + ((((lane >> 3) ^ ph) << 3) | (lane & 7));
            if (ch < prm.n) atomicAdd(prm.dbias + ch, part[ph]);
          }
        }
      }
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
      for (int mg = 0; mg < groups; ++mg) {
        const int tap_l = mg * prm.tpm + tl;
        const bool rvalid = tl < prm.tpm && tap_l < ntap;
        float* dst = prm.dw + ((int64_t)(tap0 + tap_l) * prm.npad + n0) * prm.lddw + k0 + k;
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mg * N + c0), v);
          if (rvalid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < prm.n) atomicAdd(dst + (int64_t)(c0 + j) * prm.lddw, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, (uint32_t)prm.tmem_cols);
}

bool wgrad_tf32_prologue_ok(const WgradArgs& a) {
  int bw, bh, bn;
  return pixel_box(a.S, 64, &bw, &bh, &bn) && a.kpad % 32 == 0 && a.lddy % 32 == 0;
}

int k_conv_wgrad_tf32(const WgradArgs& a, cudaStream_t st) {
  const int P = a.B * a.S * a.S;
  if (P == 0) return RNVP_OK;
  int bw, bh, bn;
  if (!pixel_box(a.S, 64, &bw, &bh, &bn) || a.kpad % 32 != 0 || a.lddy % 32 != 0) {
    RNVP_REQUIRE(a.xf_save == nullptr, "wgrad BN prologue needs the tensor-core kernel");
    return k_conv_wgrad_fp32(a, st);
  }
  RNVP_REQUIRE(a.taps == 1 || a.taps == 9, "wgrad: taps=%d", a.taps);
  const int njobs = a.njobs > 1 ? a.njobs : 1;
  RNVP_REQUIRE(njobs <= kMaxWgradJobs && (njobs == 1 || (a.jobs != nullptr && a.segs == 1)),
               "wgrad: a group holds at most %d jobs of one shape, no K-concatenated x", kMaxWgradJobs);
  WgOperands<kMaxWgradJobs> ops{};
  // MN-major fp32 operands: 128B swizzle with 32-byte atoms (the only layout tcgen05 accepts for them)
  for (int j = 0; j < njobs; ++j) {
    const WgradJob one{a.x, a.dy, a.dw, a.dbias, a.xf_save};
    const WgradJob& jb = njobs > 1 ? a.jobs[j] : one;
    RNVP_TRY(make_act_map(&ops.m[2 * j], jb.dy, a.B, a.S, a.lddy, bw, bh, bn, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    RNVP_TRY(make_act_map(&ops.m[2 * j + 1], jb.x, a.B, a.S, a.kpad, bw, bh, bn, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, a.segs,
                          a.seg_stride));
    ops.dw[j] = jb.dw; ops.dbias[j] = jb.dbias; ops.xf_save[j] = jb.xf_save;
  }
  RNVP_REQUIRE(a.segs >= 1 && (a.segs == 1 || (a.xf_save == nullptr && a.seg_stride % 4 == 0)),
               "wgrad: K-concatenated x needs 16-byte aligned segments and excludes the BN prologue");
  const int ktot = a.segs * a.kpad;
  WgradTcParams prm{};
  prm.P = P; prm.n = a.n; prm.npad = a.npad; prm.kpad = ktot; prm.taps = a.taps; prm.S = a.S;
  prm.seg_k = a.kpad; prm.lddw = a.lddw ? a.lddw : ktot;
  prm.x3 = a.x3;
  prm.xf_C = a.xf_C;
  prm.log2S = 0;
  while ((1 << prm.log2S) < a.S) ++prm.log2S;
  prm.n_tiles = ceil_div(a.n, 128);
  prm.k_tiles = ceil_div(ktot, 128);
  prm.kb = (ktot < 128 ? ktot : 128) / 32;
  prm.tpm = prm.kb == 3 ? 1 : 4 / prm.kb;
  if (prm.tpm > a.taps) prm.tpm = a.taps;
  const int N = pad_to(a.n < 128 ? a.n : 128, 32);
  int max_groups = 512 / N;
  if (max_groups > WG_MAX_GROUPS) max_groups = WG_MAX_GROUPS;
  int groups_total = ceil_div(a.taps, prm.tpm);
  prm.tap_ranges = ceil_div(groups_total, max_groups);
  int groups_per_range = ceil_div(groups_total, prm.tap_ranges);
  prm.taps_per_range = groups_per_range * prm.tpm;
  prm.tap_ranges = ceil_div(a.taps, prm.taps_per_range);
  int cols = groups_per_range * N, alloc = 32;
  while (alloc < cols) alloc <<= 1;
  prm.tmem_cols = alloc;
  prm.num_tiles = ceil_div(P, 64);
  const int base = prm.n_tiles * prm.k_tiles * prm.tap_ranges;
  // pixel splits: fill the resident slots in ONE wave (rounding up would leave a second, nearly empty wave)
  int splits = kNumSMs / (base * njobs) < 1 ? 1 : kNumSMs / (base * njobs);
  if (splits > prm.num_tiles) splits = prm.num_tiles;
  prm.tiles_per_split = ceil_div(prm.num_tiles, splits);
  splits = ceil_div(prm.num_tiles, prm.tiles_per_split);
  // smem plan.  The M = 128 A operand always addresses four 32-row groups.  With one real box per
  // stage (<= 32 input channels, one tap) the four groups alias that box (LBO = 0) and a stage is 8 KB;
  // otherwise a stage holds the real boxes and the trailing groups read into the following slots /
  // slack (their accumulator rows are never read back).
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("RNVP_WG_VARIANT");      // debugging / A-B switch: 0 forces the 1-CTA layout
    variant = e ? atoi(e) : 1;
  }
  const int a_boxes = prm.tpm * prm.kb > 4 ? 4 : prm.tpm * prm.kb;
  prm.b_stage_bytes = (N / 32) * WG_BOX_BYTES;
  bool two_ctas = false;
  if (a.x3) {
    // 3xTF32: every stage holds the hi boxes followed by their lo halves: full 4-box A slots (64 KB per stage, two
    // stages) and as many dy stages as fit behind them
    prm.a_half = 4 * WG_BOX_BYTES;
    prm.b_half = prm.b_stage_bytes;
    prm.a_stage_bytes = 2 * prm.a_half;
    prm.b_stage_bytes = 2 * prm.b_half;
    prm.a_lbo = WG_BOX_BYTES;
    prm.a_stages = 2;
    prm.b_stages = (WG_RING_BYTES - prm.a_stages * prm.a_stage_bytes) / prm.b_stage_bytes;
    if (prm.b_stages > 4) prm.b_stages = 4;
    prm.ring_bytes = prm.a_stages * prm.a_stage_bytes + prm.b_stages * prm.b_stage_bytes;
  } else if (variant != 0 && a_boxes == 1) {
    // compact: 8 KB A stages, the four groups alias the one box; small enough for two CTAs per SM
    prm.a_stage_bytes = WG_BOX_BYTES;
    prm.a_lbo = 0;
    prm.a_stages = 4 * groups_per_range;
    if (prm.a_stages > WG_MAX_STAGES) prm.a_stages = WG_MAX_STAGES;
    prm.b_stages = 4;
    prm.ring_bytes = prm.a_stages * prm.a_stage_bytes + prm.b_stages * prm.b_stage_bytes;
    two_ctas = true;
  } else {
    prm.a_stage_bytes = 4 * WG_BOX_BYTES;           // full 4-box slots: trailing groups read idle smem
    prm.a_lbo = WG_BOX_BYTES;
    // BN prologue: the transform is one more pipeline step; RNVP_WG_XF_DEEP=1 trades the second CTA per SM for a
    // four-deep A ring on those launches (A/B switch)
    static const int xf_deep = env_int("RNVP_WG_XF_DEEP", 0, 1, 0);
    if (variant != 0 && !(ops.xf_save[0] && xf_deep) && prm.tmem_cols <= 256 &&
        2 * prm.a_stage_bytes + 2 * prm.b_stage_bytes <= 100 * 1024) {
      prm.a_stages = 2;                             // two CTAs per SM hide the per-tile latency better
      prm.b_stages = (100 * 1024 - 2 * prm.a_stage_bytes) / prm.b_stage_bytes;
      if (prm.b_stages > 4) prm.b_stages = 4;
      two_ctas = true;
    } else {
      prm.a_stages = 4;
      prm.b_stages = (WG_RING_BYTES - prm.a_stages * prm.a_stage_bytes) / prm.b_stage_bytes;
      if (prm.b_stages > 6) prm.b_stages = 6;
    }
    prm.ring_bytes = prm.a_stages * prm.a_stage_bytes + prm.b_stages * prm.b_stage_bytes;
  }
  prm.ring_bytes = (prm.ring_bytes + 1023) / 1024 * 1024;
  static const bool attr_set = [] {              // thread-safe first call (autograd's backward thread may be first)
    return cudaFuncSetAttribute(conv_wgrad_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) == cudaSuccess &&
           cudaFuncSetAttribute(conv_wgrad_tf32_kernel<kMaxWgradJobs>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) ==
               cudaSuccess;
  }();
  RNVP_REQUIRE(attr_set, "wgrad: cudaFuncSetAttribute failed");
  // with the compact layout two CTAs fit per SM: split the pixel range accordingly
  const int dyn_smem = prm.ring_bytes + 1024;
  if (two_ctas && 2 * (dyn_smem + 2048) <= 227 * 1024 && prm.tmem_cols <= 256) {
    int splits2 = 2 * kNumSMs / (base * njobs) < 1 ? 1 : 2 * kNumSMs / (base * njobs);
    if (splits2 > prm.num_tiles) splits2 = prm.num_tiles;
    prm.tiles_per_split = ceil_div(prm.num_tiles, splits2);
    splits = ceil_div(prm.num_tiles, prm.tiles_per_split);
  }
  if (njobs > 1) {
    RNVP_CUDA(launch_pdl(conv_wgrad_tf32_kernel<kMaxWgradJobs>, dim3(base, splits, njobs), dim3(TC_THREADS), (size_t)dyn_smem, st,
                         ops, prm));
  } else {
    WgOperands<1> one{};
    one.m[0] = ops.m[0]; one.m[1] = ops.m[1];
    one.dw[0] = ops.dw[0]; one.dbias[0] = ops.dbias[0]; one.xf_save[0] = ops.xf_save[0];
    RNVP_CUDA(launch_pdl(conv_wgrad_tf32_kernel<1>, dim3(base, splits), dim3(TC_THREADS), (size_t)dyn_smem, st, one, prm));
  }
  RNVP_LAUNCH_CHECK();
  return RNVP_OK;
}

}  // namespace rnvp
