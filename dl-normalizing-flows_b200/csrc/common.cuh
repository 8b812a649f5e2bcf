// Shared helpers for the RealNVP hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include "../../include/rnvp.h"

namespace rnvp {

constexpr float kBnEps = 1e-5f;       // nn.BatchNorm2d default; literal at modules_realnvp.py:289,301
constexpr float kBnMomentum = 0.1f;   // nn.BatchNorm2d default
constexpr int kNumSMs = 148;          // B200

void set_error(const char* fmt, ...);

#define RNVP_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::rnvp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RNVP_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define RNVP_TRY(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != RNVP_OK) return _s;      \
  } while (0)

#define RNVP_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::rnvp::set_error(__VA_ARGS__);                \
      return RNVP_ERR_INVALID;                       \
    }                                                \
  } while (0)

// every kernel launch of the library goes through this macro: it also feeds rnvp_launch_count()
void count_launch();
#define RNVP_LAUNCH_CHECK()          \
  do {                               \
    ::rnvp::count_launch();          \
    RNVP_CUDA(cudaGetLastError());   \
  } while (0)

// optional per-kernel-class CUDA-event timing (rnvp_prof_*): bench.py uses it for the roofline line
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(int kind, int S, int taps, int cin, int cout, cudaStream_t st);
  ~ProfScope();
};
enum { PROF_CONV = 0, PROF_DGRAD = 1, PROF_WGRAD = 2, PROF_BN = 3, PROF_BN_BWD = 4, PROF_CPL = 5 };

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int pad_to(int a, int m) { return (a + m - 1) / m * m; }

bool pdl_enabled();                     // RNVP_PDL=1 switches programmatic dependent launch on
// launch `kernel` with the programmatic-stream-serialization attribute (plain launch when disabled)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks = kNumSMs * 16) {
  int64_t g = ceil_div64(work_items, per_block);
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ---------------------------------------------------------------------------
// geometry of one coupling as the elementwise kernels see it
//   x is NHWC [B,S,S,C].  "in-branch" channels feed the s/t net through in_bn,
//   "transformed" channels receive the affine map.
//   checkerboard (modules_realnvp.py:264-302): both are all C channels; the
//     in-branch sees x*m, the transform acts where (1-m) = 1, m = (cfg+i+j)&1.
//   channelwise (modules_realnvp.py:324-370): cfg=1 -> on = first half,
//     off = second half; cfg=0 the other way round; no spatial mask.
// ---------------------------------------------------------------------------
struct CplGeom {
  int B, S, C;        // x tensor
  int cio;            // channels transformed (= in-branch channels)
  int on_off;         // first transformed channel
  int in_off;         // first in-branch channel
  int ckbd;           // 1 checkerboard, 0 channelwise
  int cfg;            // mask configuration
  int cin_pad;        // padded channel stride of h0 (s/t net input)
  int cst_pad;        // padded channel stride of st (s/t net output, 2*cio real)
  __host__ __device__ int P() const { return B * S * S; }
  __host__ __device__ int cin() const { return ckbd ? 2 * cio + 1 : 2 * cio; }
  // value of the reference's `mask` at pixel p (1 = passes to the s/t net)
  __device__ float mask_in(int p) const {
    if (!ckbd) return 1.f;
    int j = p % S, i = (p / S) % S;
    return (float)((cfg + i + j) & 1);
  }
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Programmatic dependent launch (PDL): a kernel launched with the stream-serialization attribute may start
// while its predecessor drains; it must call pdl_wait() before touching anything the predecessor wrote
// (or writing anything it reads).  pdl_trigger() lets the successor begin launching early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// round-to-nearest fp32 -> tf32 (10-bit mantissa, low 13 bits cleared).  tcgen05 kind::tf32 simply
// ignores the low bits (truncation, biased towards zero); tensors that exist only as conv operands are
// therefore rounded by their producer when the TF32 tier is active.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float maybe_round(float x, int on) { return on ? round_tf32(x) : x; }
// what tcgen05 kind::tf32 reads from an fp32 word; x - trunc_tf32(x) is exact in fp32 (the "lo" operand of the
// 3xTF32 tier)
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }


// ---------------------------------------------------------------------------------------------------------
// One-shot all-reduce of a small double vector over NVLink peer memory, folded INTO the kernel that consumes the
// vector (the batch-norm statistic exchange of data-parallel training, dp.cu).  Every rank owns an inbox
// [world][2 slots][cap] doubles + [world] sequence flags that its peers map through CUDA IPC.  One CTA of the consumer
// (`pusher`) stores the local vector into slot (seq & 1) of every rank's inbox -- remote stores through the NVSwitch --
// and publishes `seq` with system-scope release stores; EVERY CTA then waits until all ranks' numbers have arrived in
// the own inbox, and the pusher replaces the local vector by the rank-ordered (hence bit-identical on all ranks) sum,
// which dp_reduced() reads.  A rank is at most one exchange ahead of the slowest one: finishing exchange k+1 needs every peer's flag
// k+1, which a peer publishes only from the kernel that FOLLOWS the one that read exchange k.  Compared with the
// stand-alone exchange kernel this removes one launch + its stream dependency from the critical path per reduction.
// Failures (a peer that never arrives, diverged call sequences) set the sticky error word: every later entry point of
// the plan then fails with RNVP_ERR_STATE (the numbers of the step in flight are garbage, and known to be).
// ---------------------------------------------------------------------------------------------------------
struct DpXchg {
  void* const* peers = nullptr;      // device array: inbox base of every rank; null = no exchange (single process,
  int rank = 0, world = 1, cap = 0;  //   or the vector was already reduced by the stand-alone kernel / NCCL)
  unsigned long long seq = 0;
  int* err = nullptr;                // sticky error word (mapped host memory)
  __host__ __device__ bool on() const { return peers != nullptr; }
};
// bar_id / nthreads: the named barrier and thread count of the calling group (all of its threads must call);
// tid = index of the caller within the group
__device__ __forceinline__ void dp_exchange(const DpXchg& x, double* local, int n, bool pusher,
                                            int tid, int nthreads, int bar_id) {
  if (!x.on()) return;
  const int slot = (int)(x.seq & 1ull);
  const size_t flag_off = (size_t)x.world * 2 * x.cap * sizeof(double);
  // [world] peer flags, then one LOCAL "exchange seq has arrived" word: only the pusher CTA polls the peers'
  // system-scope flags (hundreds of CTAs hammering them delays the very stores they wait for -- measured); the other
  // CTAs of the grid wait for the pusher's gpu-scope word
  unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x.peers[x.rank]) + flag_off);
  unsigned long long* arrived = my_flags + x.world;
  if (pusher) {
    for (int p = 0; p < x.world; ++p) {
      double* dst = reinterpret_cast<double*>(x.peers[p]) + ((size_t)x.rank * 2 + slot) * x.cap;
      for (int i = tid; i < n; i += nthreads) dst[i] = local[i];
    }
    __threadfence_system();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    if (tid < x.world) {
      unsigned long long* pflag =
          reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x.peers[tid]) + flag_off) + x.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pflag), "l"(x.seq) : "memory");
      const long long t0 = clock64();
      unsigned long long v = 0;
      for (unsigned it = 1;; ++it) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + tid) : "memory");
        if (v >= x.seq) break;
        if ((it & 0xffffu) == 0) {                                         // rarely: the error word lives in host memory
          if (*reinterpret_cast<volatile int*>(x.err) != 0) break;         // an earlier exchange already failed
          if (clock64() - t0 > 240000000000ll) break;                      // ~2 min: the peer is gone
        }
      }
      if (v < x.seq) *reinterpret_cast<volatile int*>(x.err) = 1;
      else if (v > x.seq + 1) *reinterpret_cast<volatile int*>(x.err) = 2;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    // rank-ordered sum, written back over the local vector: the rest of the grid reads plain reduced sums
    const double* inbox = reinterpret_cast<const double*>(x.peers[x.rank]) + (size_t)slot * x.cap;
    for (int i = tid; i < n; i += nthreads) {
      double s = 0.0;
      for (int p = 0; p < x.world; ++p) s += __ldcg(inbox + (size_t)p * 2 * x.cap + i);
      local[i] = s;
    }
    __threadfence();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(arrived), "l"(x.seq) : "memory");
  } else {
    if (tid == 0) {
      unsigned long long v = 0;
      for (unsigned it = 1;; ++it) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(arrived) : "memory");
        if (v >= x.seq) break;
        __nanosleep(100);
        if ((it & 0xfffffu) == 0 && *reinterpret_cast<volatile int*>(x.err) != 0) break;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  }
}
// element i of the (reduced) vector; after dp_exchange the pusher CTA has overwritten `local` with the global sums
__device__ __forceinline__ double dp_reduced(const DpXchg& x, const double* local, int i) {
  return x.on() ? __ldcg(local + i) : local[i];      // L2: written by another CTA of this grid
}

// mean / rstd / scale / shift of one BN channel from (sum, sumsq) over `count` values
struct BnCoef { float mean, rstd, scale, shift, var; };
__device__ __forceinline__ BnCoef bn_coef_from_sums(double s, double ss, double count, float gamma, float beta) {
  double mean = s / count;
  double var = ss / count - mean * mean;
  if (var < 0.0) var = 0.0;
  BnCoef c;
  c.mean = (float)mean;
  c.var = (float)var;
  c.rstd = (float)(1.0 / sqrt(var + (double)kBnEps));
  c.scale = gamma * c.rstd;
  c.shift = beta - c.mean * c.scale;
  return c;
}
__device__ __forceinline__ BnCoef bn_coef_from_running(float rm, float rv, float gamma, float beta) {
  BnCoef c;
  c.mean = rm;
  c.var = rv;
  c.rstd = 1.0f / sqrtf(rv + kBnEps);
  c.scale = gamma * c.rstd;
  c.shift = beta - c.mean * c.scale;
  return c;
}

}  // namespace rnvp
