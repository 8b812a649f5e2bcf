// Data-parallel plumbing (absent from the reference): an NCCL communicator owned by
// the plan, used for (i) gradient-bucket all-reduce and (ii) the batch-norm statistic exchange that
// turns every BN of the stack into a synchronised BN.  NCCL is resolved with dlopen at first use so
// that the library loads on a box without it (single-GPU use never touches these symbols).
#include <dlfcn.h>
#include <mutex>
#include "kernels.h"

namespace {
// minimal NCCL ABI (nccl.h 2.2x): types and enum values are stable across 2.x
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclSum_ = 0 };
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

// accessors implemented in runtime.cu (the plan struct is private to it)
namespace rnvp {
void** plan_comm_slot(rnvp_plan* p);
void plan_set_ranks(rnvp_plan* p, int rank, int world);
int plan_world(const rnvp_plan* p);

int dp_allreduce_doubles(rnvp_plan* plan, double* buf, size_t n, cudaStream_t st) {
  ncclComm_t comm = (ncclComm_t)*plan_comm_slot(plan);
  RNVP_REQUIRE(comm != nullptr, "data-parallel communicator not initialised");
  RNVP_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat64_, ncclSum_, comm, st));
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
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  RNVP_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  *rnvp::plan_comm_slot(plan) = comm;
  rnvp::plan_set_ranks(plan, rank, world);
  return RNVP_OK;
}

int rnvp_dp_finalize(rnvp_plan* plan) {
  if (!plan) return RNVP_OK;
  void** slot = rnvp::plan_comm_slot(plan);
  if (*slot) {
    RNVP_NCCL(g_nccl.CommDestroy((ncclComm_t)*slot));
    *slot = nullptr;
  }
  rnvp::plan_set_ranks(plan, 0, 1);
  return RNVP_OK;
}

int rnvp_dp_allreduce(rnvp_plan* plan, float* buf, size_t n, void* stream) {
  RNVP_REQUIRE(plan, "null plan");
  if (rnvp::plan_world(plan) <= 1 || n == 0) return RNVP_OK;
  ncclComm_t comm = (ncclComm_t)*rnvp::plan_comm_slot(plan);
  RNVP_REQUIRE(comm != nullptr, "data-parallel communicator not initialised");
  RNVP_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat32_, ncclSum_, comm, (cudaStream_t)stream));
  return RNVP_OK;
}

}  // extern "C"
