// NCCL plumbing (loaded with dlopen so that the library has no link-time NCCL dependency and shares the
// libnccl.so.2 a host process -- e.g. torch.distributed -- has already loaded).
// Used for (a) the per-evaluation sum of the row-sharded GLM partials and
// (b) the final all-reduce of chain moments.
#include <dlfcn.h>
#include <string.h>
#include "smc_internal.h"

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return SMC_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { smc_set_error("cannot load libnccl.so.2: %s", dlerror()); return SMC_ERR_COMM; }
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) {
    smc_set_error("libnccl is missing expected symbols");
    return SMC_ERR_COMM;
  }
  return SMC_OK;
}
const int kNcclFloat64 = 8, kNcclSum = 0;
}  // namespace

extern "C" int smc_comm_unique_id(void* out_id, int id_bytes) {
  if (!out_id || id_bytes < (int)sizeof(nccl_uid)) { smc_set_error("smc_comm_unique_id: buffer must hold 128 bytes"); return SMC_ERR_INVALID; }
  SMC_TRY(load_nccl());
  nccl_uid id;
  int rc = g_nccl.get_uid(&id);
  if (rc != 0) { smc_set_error("ncclGetUniqueId: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "error"); return SMC_ERR_COMM; }
  memcpy(out_id, &id, sizeof(id));
  return SMC_OK;
}

extern "C" int smc_ctx_comm_init(smc_ctx_t ctx, const void* nccl_unique_id, int id_bytes, int rank, int nranks) {
  if (!ctx || !nccl_unique_id || id_bytes < (int)sizeof(nccl_uid) || rank < 0 || rank >= nranks) {
    smc_set_error("smc_ctx_comm_init: bad argument");
    return SMC_ERR_INVALID;
  }
  SMC_TRY(load_nccl());
  cudaSetDevice(ctx->device);
  nccl_uid id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  void* comm = nullptr;
  int rc = g_nccl.init_rank(&comm, nranks, id, rank);
  if (rc != 0) { smc_set_error("ncclCommInitRank: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "error"); return SMC_ERR_COMM; }
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return SMC_OK;
}

int smc_comm_destroy(smc_ctx_s* c) {
  if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  return SMC_OK;
}

// in-stream sum over ranks of a device buffer of doubles
int smc_comm_allreduce_device(smc_ctx_s* c, double* buf, int64_t n) {
  if (c->nranks <= 1) return SMC_OK;
  if (!c->nccl_comm) { smc_set_error("row-sharded evaluation needs smc_ctx_comm_init"); return SMC_ERR_COMM; }
  int rc = g_nccl.allreduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream);
  if (rc != 0) { smc_set_error("ncclAllReduce: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "error"); return SMC_ERR_COMM; }
  return SMC_OK;
}

extern "C" int smc_ctx_allreduce(smc_ctx_t ctx, double* host_inout, int64_t n) {
  if (!ctx || !host_inout || n < 0) { smc_set_error("smc_ctx_allreduce: bad argument"); return SMC_ERR_INVALID; }
  if (ctx->nranks <= 1 || n == 0) return SMC_OK;
  cudaSetDevice(ctx->device);
  double* d = nullptr;
  SMC_CUDA(smc_dev_alloc(&d, (size_t)n * sizeof(double), ctx->stream));
  SMC_CUDA(cudaMemcpyAsync(d, host_inout, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int rc = smc_comm_allreduce_device(ctx, d, n);
  if (rc == SMC_OK) {
    SMC_CUDA(cudaMemcpyAsync(host_inout, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SMC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  smc_dev_free(d, ctx->stream);
  return rc;
}
