// Data-parallel plumbing (absent from the reference): an NCCL communicator owned by
// the plan, used for
//   (i)  the gradient all-reduce: the flat gradient buffer is reduced in buckets of whole couplings
//        on a side stream as soon as the backward of those couplings has been enqueued (couplings
//        finish last -> first, and 93 of the 120 M parameters live in the last two groups, so most of
//        the traffic overlaps the rest of the backward pass), and
//   (ii) the batch-norm statistic exchange (per-channel double sums) that makes every BN of the stack
//        a synchronised BN, issued in-stream between the producer and consumer kernels.
// NCCL is resolved with dlopen at first use so that the library loads on a box without it.
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include "kernels.h"

namespace {
// minimal NCCL ABI (nccl.h 2.x): types and enum values are stable across 2.x
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclSum_ = 0, ncclAvg_ = 4 };
enum { ncclFloat32_ = 7, ncclFloat64_ = 8 };
typedef int (*fn_GetUniqueId)(ncclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct Nccl {
  void* lib = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  bool ok = false;
};
Nccl g_nccl;
std::once_flag g_once;

void load_nccl() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return;
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.GetErrorString = (fn_GetErrorString)dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce;
}
int need_nccl() {
  std::call_once(g_once, load_nccl);
  if (!g_nccl.ok) {
    rnvp::set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror());
    return RNVP_ERR_NCCL;
  }
  return RNVP_OK;
}
#define RNVP_NCCL(expr)                                                                       \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != ncclSuccess_) {                                                                 \
      rnvp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                           \
                      g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");      \
      return RNVP_ERR_NCCL;                                                                   \
    }                                                                                         \
  } while (0)
}  // namespace

namespace rnvp {
DpState* plan_dp(rnvp_plan* p);              // runtime.cu (the plan struct is private to it)
int plan_num_couplings(const rnvp_plan* p);

// ---------------------------------------------------------------------------------------------------------
// One-shot all-reduce of a small double vector over NVLink peer memory.
//
// The 840+ batch-norm statistic exchanges of a step are 48 B ... 8 KB each and strictly serialised by data
// dependence, so their cost is pure latency.  Instead of an NCCL call (a kernel launch plus its internal
// handshakes) each exchange is ONE single-CTA kernel: push the local partial sums straight into every
// rank's inbox (remote stores through the NVSwitch), publish a sequence number with a system-scope release
// store, wait until every peer's sequence number has arrived in the own inbox, and add the world's vectors up
// in rank order (so all ranks obtain bit-identical sums).  Inbox slots alternate with the parity of the
// sequence number: a rank can only be one exchange ahead of the slowest one, because finishing exchange k+1
// needs every peer's flag k+1, which a peer publishes only after it has finished reading exchange k.
// A rank may legitimately lag by seconds (checkpoint save, logging, a dataloader stall, first-call module load),
// so the wait is LONG (~2 minutes of spinning); if it still expires, or a peer's sequence number is more than
// one ahead (the ranks issued different numbers of exchanges), the kernel poisons its result with NaN and sets a
// STICKY error word in mapped host memory: every later rnvp_flow_* / rnvp_coupling_* call of the plan fails with
// RNVP_ERR_STATE (make_ctx checks the word), instead of silently training on wrong statistics.
// ---------------------------------------------------------------------------------------------------------
constexpr int kXchgThreads = 256;
__device__ __forceinline__ size_t xchg_flag_off(int world, int cap) { return (size_t)world * 2 * cap * sizeof(double); }

__global__ void __launch_bounds__(kXchgThreads) stats_exchange_kernel(void* const* __restrict__ peers, int rank, int world,
                                                                      int cap, double* __restrict__ buf, int n,
                                                                      unsigned long long seq, int* err) {
  __shared__ int failed;
  if (threadIdx.x == 0) failed = 0;
  const int slot = (int)(seq & 1ull);
  // 1. push: my vector into slot [rank][slot] of every rank's inbox (my own included)
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peers[p]) + ((size_t)rank * 2 + slot) * cap;
    for (int i = threadIdx.x; i < n; i += kXchgThreads) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait for everybody
  if (threadIdx.x < world) {
    const int p = threadIdx.x;
    unsigned long long* pflag = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers[p]) + xchg_flag_off(world, cap)) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pflag), "l"(seq) : "memory");
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(peers[rank]) + xchg_flag_off(world, cap)) + p;
    const long long t0 = clock64();
    unsigned long long v = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > 240000000000ll) break;                           // ~2 min at 2 GHz: the peer is gone
    }
    // a healthy peer is at `seq` or one exchange ahead; anything else is a timeout (1) or a desynchronised
    // call sequence (2): e.g. one rank ran an extra train-mode forward
    if (v < seq) { failed = 1; *reinterpret_cast<volatile int*>(err) = 1; }
    else if (v > seq + 1) { failed = 1; *reinterpret_cast<volatile int*>(err) = 2; }
  }
  __syncthreads();
  if (failed) {
    __threadfence_system();
    for (int i = threadIdx.x; i < n; i += kXchgThreads) buf[i] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  // 4. reduce in rank order
  const double* inbox = reinterpret_cast<const double*>(peers[rank]);
  for (int i = threadIdx.x; i < n; i += kXchgThreads) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += inbox[((size_t)p * 2 + slot) * cap + i];
    buf[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same exchange without the fence and the flag round trip ("LL": data and flag travel in one store).  Every double
// is pushed as ONE 16-byte store {lo32, tag, hi32, tag} with tag = the low 32 bits of the sequence number: whatever the
// fabric does with the 16 bytes, each 8-byte half carries its own tag, so a reader that sees both tags equal to `seq`
// holds the value -- no __threadfence_system() (a remote-write round trip) before a separate flag store (another
// one-way trip).  Readers poll their OWN inbox (local memory) element by element and add the world's vectors up in rank
// order (bit-identical on all ranks).  Slot reuse is safe for the same reason as above: a peer can only start exchange
// k+2 after finishing k+1, which needs my contribution to k+1, which I push after I have read exchange k.
// The LL inbox [world][2][cap] x 16 bytes lies behind the flag words of the fenced protocol's inbox.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t xchg_ll_off(int world, int cap) {
  return (xchg_flag_off(world, cap) + (size_t)(world + 1) * sizeof(unsigned long long) + 15) & ~(size_t)15;
}
__global__ void __launch_bounds__(kXchgThreads) stats_exchange_ll_kernel(void* const* __restrict__ peers, int rank, int world,
                                                                         int cap, double* __restrict__ buf, int n,
                                                                         unsigned long long seq, int* err) {
  __shared__ int failed;
  if (threadIdx.x == 0) failed = 0;
  __syncthreads();
  const int slot = (int)(seq & 1ull);
  const uint32_t tag = (uint32_t)seq;
  const size_t ll = xchg_ll_off(world, cap);
  // 1. push {lo, tag, hi, tag} into slot [rank][slot] of every rank's LL inbox (my own included)
  for (int i = threadIdx.x; i < n; i += kXchgThreads) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[i]);
    const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
    for (int p = 0; p < world; ++p) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(peers[p]) + ll) + ((size_t)rank * 2 + slot) * cap + i;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
    }
  }
  // 2. gather: element i of every rank from my own inbox, in rank order
  const uint4* inbox = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(peers[rank]) + ll) + (size_t)slot * cap;
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < n; i += kXchgThreads) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) {
      const uint4* src = inbox + (size_t)p * 2 * cap + i;
      uint32_t a, ta, b, tb;
      for (unsigned it = 1;; ++it) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(ta), "=r"(b), "=r"(tb) : "l"(src) : "memory");
        if (ta == tag && tb == tag) break;
        // a tag from the future: this rank's peers have moved on by two exchanges -- the call sequences diverged
        if ((int32_t)(ta - tag) > 0 && (int32_t)(tb - tag) > 0) { failed = 2; break; }
        if ((it & 0xffffu) == 0) {
          if (*reinterpret_cast<volatile int*>(&failed) != 0) break;               // another thread of this CTA gave up
          if (*reinterpret_cast<volatile int*>(err) != 0) { failed = 1; break; }   // an earlier exchange already failed
          if (clock64() - t0 > 240000000000ll) { failed = 1; break; }              // ~2 min: the peer is gone
        }
      }
      if (*reinterpret_cast<volatile int*>(&failed) != 0) break;
      s += __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
    }
    buf[i] = s;
  }
  __syncthreads();
  if (failed) {
    if (threadIdx.x == 0) { *reinterpret_cast<volatile int*>(err) = failed; __threadfence_system(); }
    for (int i = threadIdx.x; i < n; i += kXchgThreads) buf[i] = __longlong_as_double(0x7ff8000000000000ll);
  }
}

int dp_allreduce_doubles(DpState* dp, double* buf, size_t n, cudaStream_t st) {
  if (dp->xchg_ready && (int)n <= dp->xchg_cap) {
    const unsigned long long seq = ++dp->xchg_seq;
    static int use_ll = -1;
    if (use_ll < 0) {
      const char* e = getenv("RNVP_XCHG_LL");           // A/B switch: 0 = the fenced data + flag protocol
      use_ll = (e && e[0] == '0') ? 0 : 1;
    }
    if (use_ll) {
      stats_exchange_ll_kernel<<<1, kXchgThreads, 0, st>>>(dp->xchg_peers_dev, dp->rank, dp->world, dp->xchg_cap, buf, (int)n,
                                                           seq, dp->xchg_err);
      RNVP_LAUNCH_CHECK();
      return RNVP_OK;
    }
    // plain launch: programmatic dependent launch bought nothing here (measured) and the exchange is the one
    // kernel whose early start could only add waiting peers
    stats_exchange_kernel<<<1, kXchgThreads, 0, st>>>(dp->xchg_peers_dev, dp->rank, dp->world, dp->xchg_cap, buf, (int)n,
                                                      seq, dp->xchg_err);
    RNVP_LAUNCH_CHECK();
    return RNVP_OK;
  }
  RNVP_REQUIRE(dp->comm != nullptr, "data-parallel communicator not initialised");
  RNVP_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat64_, ncclSum_, (ncclComm_t)dp->comm, st));
  return RNVP_OK;
}

// Synchronised batch norm and the gradient average assume that every rank holds the same local batch (count =
// local pixels * world, ncclAvg).  Each training forward therefore reduces (B, B^2) over the ranks first; equal
// batches <=> world * sum B^2 == (sum B)^2.  A mismatch sets the sticky error word (3).
__global__ void batch_check_kernel(const double* v, int world, int* err) {
  const double s1 = v[0], s2 = v[1];
  if (s1 != s1 || fabs((double)world * s2 - s1 * s1) > 0.5) *reinterpret_cast<volatile int*>(err) = (s1 != s1) ? 1 : 3;
}
int dp_check_equal_batches(DpState* dp, int batch, double* scratch2, cudaStream_t st) {
  if (dp->world <= 1) return RNVP_OK;
  const double h[2] = {(double)batch, (double)batch * (double)batch};
  RNVP_CUDA(cudaMemcpyAsync(scratch2, h, sizeof(h), cudaMemcpyHostToDevice, st));
  RNVP_TRY(dp_allreduce_doubles(dp, scratch2, 2, st));
  if (dp->xchg_err) {
    batch_check_kernel<<<1, 1, 0, st>>>(scratch2, dp->world, dp->xchg_err);
    RNVP_LAUNCH_CHECK();
  }
  return RNVP_OK;
}
int dp_sticky_error(const DpState* dp) {
  return dp->xchg_err_host ? *reinterpret_cast<volatile int*>(dp->xchg_err_host) : 0;
}

int dp_fused_exchange(DpState* dp, double* buf, size_t n, cudaStream_t st, DpXchg* out) {
  *out = DpXchg();
  if (dp->world <= 1) return RNVP_OK;
  // Opt-in (RNVP_DP_FUSED=1).  Bit-exact, but measured SLOWER than the stand-alone exchange kernel at 2 GPUs (82.1 vs
  // 77.0 ms per step, profiles/r02_scaling.md): while the single-CTA kernel waits for the peers the SMs are free for the
  // wgrad side stream, whereas a consumer grid that waits in its prologue holds them.
  static int fused = -1;
  if (fused < 0) {
    const char* e = getenv("RNVP_DP_FUSED");
    fused = (e && e[0] == '1') ? 1 : 0;
  }
  if (fused && dp->xchg_ready && (int)n <= dp->xchg_cap) {
    out->peers = dp->xchg_peers_dev;
    out->rank = dp->rank; out->world = dp->world; out->cap = dp->xchg_cap;
    out->seq = ++dp->xchg_seq;
    out->err = dp->xchg_err;
    return RNVP_OK;
  }
  return dp_allreduce_doubles(dp, buf, n, st);
}

static int launch_bucket(DpState* dp, int64_t begin, int64_t end, cudaStream_t main) {
  if (end <= begin) return RNVP_OK;
  RNVP_CUDA(cudaEventRecord(dp->ev_main, main));
  RNVP_CUDA(cudaStreamWaitEvent(dp->comm_stream, dp->ev_main, 0));
  RNVP_NCCL(g_nccl.AllReduce(dp->flat + begin, dp->flat + begin, (size_t)(end - begin), ncclFloat32_, ncclAvg_,
                             (ncclComm_t)dp->comm, dp->comm_stream));
  dp->launched = true;
  return RNVP_OK;
}

int dp_coupling_done(DpState* dp, int ci, cudaStream_t main) {
  if (dp->world <= 1 || dp->flat == nullptr) return RNVP_OK;
  if (dp->pending_end < 0) dp->pending_end = dp->off[dp->n_cpl];
  const int64_t begin = dp->off[ci];
  if (dp->pending_end - begin >= dp->bucket_elems || ci == 0) {
    RNVP_TRY(launch_bucket(dp, begin, dp->pending_end, main));
    dp->pending_end = begin;
  }
  return RNVP_OK;
}

int dp_join(DpState* dp, cudaStream_t main) {
  if (dp->world <= 1 || dp->flat == nullptr) return RNVP_OK;
  if (dp->launched) {
    RNVP_CUDA(cudaEventRecord(dp->ev_comm, dp->comm_stream));
    RNVP_CUDA(cudaStreamWaitEvent(main, dp->ev_comm, 0));
  }
  dp->pending_end = -1;
  dp->launched = false;
  return RNVP_OK;
}
}  // namespace rnvp

extern "C" {

int rnvp_dp_unique_id(void* id128) {
  RNVP_TRY(need_nccl());
  ncclUniqueId id;
  RNVP_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return RNVP_OK;
}

int rnvp_dp_init(rnvp_plan* plan, const void* id128, int rank, int world) {
  RNVP_REQUIRE(plan && id128 && world >= 1 && rank >= 0 && rank < world, "bad data-parallel arguments");
  RNVP_TRY(need_nccl());
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  RNVP_REQUIRE(dp->comm == nullptr, "data-parallel communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  RNVP_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  dp->comm = comm;
  dp->rank = rank;
  dp->world = world;
  RNVP_CUDA(cudaStreamCreateWithFlags(&dp->comm_stream, cudaStreamNonBlocking));
  RNVP_CUDA(cudaEventCreateWithFlags(&dp->ev_main, cudaEventDisableTiming));
  RNVP_CUDA(cudaEventCreateWithFlags(&dp->ev_comm, cudaEventDisableTiming));
  return RNVP_OK;
}

// ---- NVLink statistic exchange: rank-local inbox, CUDA-IPC handles, peer mappings ------------------------
int rnvp_dp_xchg_alloc(rnvp_plan* plan, int cap_doubles, void* ipc_handle64) {
  RNVP_REQUIRE(plan && ipc_handle64 && cap_doubles > 0, "rnvp_dp_xchg_alloc: bad arguments");
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  RNVP_REQUIRE(dp->world > 1 && dp->world <= 16, "statistic exchange: world size %d unsupported (2..16)", dp->world);
  RNVP_REQUIRE(dp->xchg_local == nullptr, "statistic exchange already allocated");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  // [world][2][cap] doubles, [world] peer sequence flags, 1 local "arrived" word (fused exchange, common.cuh)
  // ... then the LL inbox [world][2][cap] x 16 bytes (stats_exchange_ll_kernel)
  const size_t bytes = ((((size_t)dp->world * 2 * cap_doubles * sizeof(double) + (size_t)(dp->world + 1) * sizeof(unsigned long long)) + 15) &
                        ~(size_t)15) + (size_t)dp->world * 2 * cap_doubles * 16;
  RNVP_CUDA(cudaMalloc(&dp->xchg_local, bytes));
  RNVP_CUDA(cudaMemset(dp->xchg_local, 0, bytes));
  RNVP_CUDA(cudaHostAlloc(&dp->xchg_err_host, sizeof(int), cudaHostAllocMapped));
  *dp->xchg_err_host = 0;
  RNVP_CUDA(cudaHostGetDevicePointer(&dp->xchg_err, dp->xchg_err_host, 0));
  RNVP_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  RNVP_CUDA(cudaIpcGetMemHandle(&h, dp->xchg_local));
  memcpy(ipc_handle64, &h, sizeof(h));
  dp->xchg_cap = cap_doubles;
  return RNVP_OK;
}

int rnvp_dp_xchg_open(rnvp_plan* plan, const void* all_handles) {
  RNVP_REQUIRE(plan && all_handles, "rnvp_dp_xchg_open: bad arguments");
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  RNVP_REQUIRE(dp->xchg_local != nullptr, "call rnvp_dp_xchg_alloc first");
  void* bases[16] = {};
  for (int r = 0; r < dp->world; ++r) {
    if (r == dp->rank) { bases[r] = dp->xchg_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    RNVP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    dp->xchg_peer_host[r] = p;
    bases[r] = p;
  }
  RNVP_CUDA(cudaMalloc(&dp->xchg_peers_dev, 16 * sizeof(void*)));
  RNVP_CUDA(cudaMemcpy(dp->xchg_peers_dev, bases, 16 * sizeof(void*), cudaMemcpyHostToDevice));
  dp->xchg_seq = 0;
  dp->xchg_ready = true;
  return RNVP_OK;
}

// sticky error word of the statistic exchange: 0 healthy, 1 a peer never arrived, 2 the ranks' call sequences
// diverged, 3 the ranks' local batch sizes differ; -1 when the exchange is not in use.  Reading costs nothing (the
// word lives in mapped host memory); it is never cleared: the communicator must be rebuilt.
int rnvp_dp_xchg_errors(rnvp_plan* plan) {
  if (!plan) return -1;
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  if (!dp->xchg_err_host) return -1;
  return *reinterpret_cast<volatile int*>(dp->xchg_err_host);
}

int rnvp_dp_set_grad_layout(rnvp_plan* plan, float* flat, const int64_t* offsets_host, int64_t bucket_elems) {
  RNVP_REQUIRE(plan && flat && offsets_host, "null argument");
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  const int n = rnvp::plan_num_couplings(plan);
  delete[] dp->off;
  dp->off = new int64_t[n + 1];
  for (int i = 0; i <= n; ++i) dp->off[i] = offsets_host[i];
  dp->n_cpl = n;
  dp->flat = flat;
  if (bucket_elems > 0) dp->bucket_elems = bucket_elems;
  dp->pending_end = -1;
  return RNVP_OK;
}

int rnvp_dp_finalize(rnvp_plan* plan) {
  if (!plan) return RNVP_OK;
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  if (dp->comm) {
    cudaStreamSynchronize(dp->comm_stream);
    RNVP_NCCL(g_nccl.CommDestroy((ncclComm_t)dp->comm));
    dp->comm = nullptr;
    cudaStreamDestroy(dp->comm_stream);
    cudaEventDestroy(dp->ev_main);
    cudaEventDestroy(dp->ev_comm);
    dp->comm_stream = nullptr;
  }
  if (dp->xchg_local) {
    cudaDeviceSynchronize();
    dp->xchg_ready = false;
    for (int r = 0; r < 16; ++r)
      if (dp->xchg_peer_host[r]) { cudaIpcCloseMemHandle(dp->xchg_peer_host[r]); dp->xchg_peer_host[r] = nullptr; }
    if (dp->xchg_peers_dev) cudaFree(dp->xchg_peers_dev);
    cudaFree(dp->xchg_local);
    if (dp->xchg_err_host) cudaFreeHost(dp->xchg_err_host);
    dp->xchg_peers_dev = nullptr; dp->xchg_local = nullptr; dp->xchg_err = nullptr; dp->xchg_err_host = nullptr;
  }
  delete[] dp->off;
  dp->off = nullptr;
  dp->flat = nullptr;
  dp->rank = 0;
  dp->world = 1;
  return RNVP_OK;
}

// the reduction the batch-norm statistics go through (NVLink exchange when open and n fits, else NCCL)
int rnvp_dp_allreduce_stats(rnvp_plan* plan, double* buf, size_t n, void* stream) {
  RNVP_REQUIRE(plan && buf, "null argument");
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  if (dp->world <= 1 || n == 0) return RNVP_OK;
  return rnvp::dp_allreduce_doubles(dp, buf, n, (cudaStream_t)stream);
}

int rnvp_dp_allreduce(rnvp_plan* plan, float* buf, size_t n, void* stream) {
  RNVP_REQUIRE(plan, "null plan");
  rnvp::DpState* dp = rnvp::plan_dp(plan);
  if (dp->world <= 1 || n == 0) return RNVP_OK;
  RNVP_REQUIRE(dp->comm != nullptr, "data-parallel communicator not initialised");
  RNVP_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat32_, ncclSum_, (ncclComm_t)dp->comm, (cudaStream_t)stream));
  return RNVP_OK;
}

}  // extern "C"
