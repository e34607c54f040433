// Multi-GPU plumbing: one process per GPU, one NCCL communicator per context.
// The reference is single-process / single-thread and has no communication backend; the only
// collective this library issues is the per-iteration all-reduce of the reduced camera system.
// NCCL is dlopen'ed so that single-GPU use has no dependency on it. The ncclUniqueId is created on
// rank 0 (tslam_nccl_unique_id) and distributed by the host program (torch.distributed in bench.py).
#include <dlfcn.h>
#include <cstring>
#include "ctx.cuh"
#include "solver.cuh"

namespace tsl {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(ncclComm_p);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return TSLAM_OK;
  const char* env = getenv("TSLAM_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return set_error(TSLAM_ERR_NCCL, "cannot dlopen libnccl.so.2 (%s); set TSLAM_NCCL_LIB", dlerror());
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
    return set_error(TSLAM_ERR_NCCL, "libnccl is missing required symbols");
  g_nccl.handle = h;
  return TSLAM_OK;
}

#define TSL_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != 0) return set_error(TSLAM_ERR_NCCL, "%s -> %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
  } while (0)

// ncclDataType_t: ncclFloat64 = 8 ; ncclRedOp_t: ncclSum = 0, ncclMax = 2
static int allreduce(tslam_ctx* ctx, double* buf, size_t n, int op) {
  if (ctx->world <= 1 || n == 0) return TSLAM_OK;
  if (!ctx->nccl_comm) return set_error(TSLAM_ERR_NCCL, "world=%d but tslam_ctx_init_comm was not called", ctx->world);
  TSL_NCCL(g_nccl.AllReduce(buf, buf, n, 8, op, (ncclComm_p)ctx->nccl_comm, ctx->stream));
  return TSLAM_OK;
}
int comm_allreduce_sum(tslam_ctx* ctx, double* buf, size_t n) { return allreduce(ctx, buf, n, 0); }
// ncclInt32 = 2: flag / count tables of the sharded structure analysis (analysis_dev.cu)
int comm_allreduce_sum_i32(tslam_ctx* ctx, int* buf, size_t n) {
  if (ctx->world <= 1 || n == 0) return TSLAM_OK;
  if (!ctx->nccl_comm) return set_error(TSLAM_ERR_NCCL, "world=%d but tslam_ctx_init_comm was not called", ctx->world);
  TSL_NCCL(g_nccl.AllReduce(buf, buf, n, 2, 0, (ncclComm_p)ctx->nccl_comm, ctx->stream));
  return TSLAM_OK;
}
int comm_allreduce_max(tslam_ctx* ctx, double* buf, size_t n) { return allreduce(ctx, buf, n, 2); }

}  // namespace tsl

using namespace tsl;

extern "C" {

int tslam_shard_owner(int landmark_is_free, int landmark_index, int obs_index, int world) {
  if (world < 1) return set_error(TSLAM_ERR_ARG, "world < 1");
  return obs_owner(landmark_is_free != 0, landmark_index, obs_index, world);
}

int tslam_nccl_unique_id(uint8_t id_out[128]) {
  if (!id_out) return set_error(TSLAM_ERR_ARG, "null argument");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t id;
  TSL_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return TSLAM_OK;
}

int tslam_ctx_init_comm(tslam_ctx* ctx, int rank, int world, const uint8_t nccl_unique_id[128]) {
  if (!ctx || !nccl_unique_id) return set_error(TSLAM_ERR_ARG, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return set_error(TSLAM_ERR_ARG, "bad rank/world %d/%d", rank, world);
  TSL_CUDA(cudaSetDevice(ctx->device));
  if (world == 1) { ctx->rank = 0; ctx->world = 1; return TSLAM_OK; }
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t id;
  memcpy(id.internal, nccl_unique_id, 128);
  ncclComm_p comm = nullptr;
  TSL_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm; ctx->rank = rank; ctx->world = world;
  return TSLAM_OK;
}

void tslam_comm_destroy(tslam_ctx* ctx) {
  if (ctx && ctx->nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_p)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
}

}  // extern "C"
