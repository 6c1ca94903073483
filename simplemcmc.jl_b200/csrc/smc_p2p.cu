// Peer-memory sum of the row-sharded GLM statement's (L, G) block over NVLink / NVSwitch: a drop-in for the in-stream NCCL
// all-reduce of smc_comm.cu on one node ("cfg3 rows sharded, few chains": at 4 GPUs and 2 chains a
// leapfrog is 164 us around 44 us of streaming, a few us of which are the NCCL launch).
//
// OPT-IN AND NOT YET RUN ON A DEVICE (written after the round's GPU minutes were spent): a context uses it only after the
// host has called smc_ctx_p2p_export on every rank, exchanged the handles, and called smc_ctx_p2p_import.
//
// Every rank owns one exchange buffer (plain cudaMalloc, exported with cudaIpcGetMemHandle and mapped by the peers):
//   flags[2][16] (unsigned, one per (parity, source rank)) and slots[2][R][kSlotDoubles].
// A call with epoch e (1, 2, ...; the same on all ranks because row-sharded ranks advance identical chains) is two kernels:
//   push: copy the local block into slot[e & 1][my rank] of EVERY rank (remote stores), fence, and the last CTA to finish
//         writes e into flags[e & 1][my rank] of every rank (release, system scope);
//   pull: wait until flags[e & 1][q] >= e for every q (acquire, system scope, bounded), then out[i] = sum over q in rank
//         order of slot[e & 1][q][i]: the same additions in the same order on every rank, so the chains stay bit-identical.
// Two parities suffice: a peer can push epoch e + 2 only after its pull of e + 1, which needs my push of e + 1, which my
// stream orders after my pull of e.
#include <string.h>
#include "smc_internal.h"

namespace {
constexpr int kMaxPeers = 16;
constexpr long long kSlotDoubles = 1ll << 19;            // 4 MB per (parity, rank): (MP + 1) * Cpad up to 524 288
constexpr size_t kFlagBytes = 256;                       // flags[2][16] unsigned = 128 bytes, padded

struct P2PView {
  unsigned char* peer[kMaxPeers];
  int rank, nranks;
  unsigned* count;                                       // local: CTAs of the push kernel that have finished
  int* status;
};
__device__ __forceinline__ unsigned* flag_ptr(unsigned char* base, int parity, int src) {
  return (unsigned*)base + parity * kMaxPeers + src;
}
__device__ __forceinline__ double* slot_ptr(unsigned char* base, int nranks, int parity, int src) {
  return (double*)(base + kFlagBytes) + ((long long)parity * nranks + src) * kSlotDoubles;
}

__global__ void __launch_bounds__(256) k_p2p_push(const double* __restrict__ src, long long n, P2PView v, unsigned epoch) {
  const int parity = (int)(epoch & 1u);
  for (int q = 0; q < v.nranks; q++) {
    double* dst = slot_ptr(v.peer[q], v.nranks, parity, v.rank);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();                                // my stores are visible system-wide before my CTA is counted
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(v.count, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < v.nranks) {
      unsigned* f = flag_ptr(v.peer[threadIdx.x], parity, v.rank);
      asm volatile("st.release.sys.global.u32 [%0], %1;\n" :: "l"(f), "r"(epoch) : "memory");
    }
    if (threadIdx.x == 0) *v.count = 0u;
  }
}

__global__ void __launch_bounds__(256) k_p2p_pull(double* __restrict__ out, long long n, P2PView v, unsigned epoch) {
  const int parity = (int)(epoch & 1u);
  unsigned char* mine = v.peer[v.rank];
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < v.nranks) {
    const unsigned* f = flag_ptr(mine, parity, threadIdx.x);
    unsigned seen = 0;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(f) : "memory");
      if ((int)(seen - epoch) >= 0) break;
      if (clock64() - t0 > 20000000000ll) { ok = 0; break; }          // ~10 s: a peer that never arrives is an error
    }
  }
  __syncthreads();
  if (!ok) { if (threadIdx.x == 0) *v.status = 1; return; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < v.nranks; q++) s += slot_ptr(mine, v.nranks, parity, q)[i];
    out[i] = s;
  }
}
}  // namespace

// local exchange buffer for `nranks` ranks; writes its 64-byte IPC handle
extern "C" int smc_ctx_p2p_export(smc_ctx_t ctx, int nranks, void* handle_out, int handle_bytes) {
  if (!ctx || !handle_out || handle_bytes < (int)sizeof(cudaIpcMemHandle_t) || nranks < 2 || nranks > kMaxPeers) {
    smc_set_error("smc_ctx_p2p_export: bad argument (2..%d ranks, %d-byte handle)", kMaxPeers, (int)sizeof(cudaIpcMemHandle_t));
    return SMC_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  if (!ctx->p2p_local) {
    const size_t bytes = kFlagBytes + (size_t)2 * nranks * kSlotDoubles * sizeof(double);
    SMC_CUDA(cudaMalloc(&ctx->p2p_local, bytes));        // not from the stream-ordered pool: pool memory has no IPC handle
    SMC_CUDA(cudaMemset(ctx->p2p_local, 0, bytes));
    SMC_CUDA(cudaMalloc((void**)&ctx->p2p_count, sizeof(unsigned) + sizeof(int)));
    SMC_CUDA(cudaMemset(ctx->p2p_count, 0, sizeof(unsigned) + sizeof(int)));
    SMC_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  SMC_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
  memcpy(handle_out, &h, sizeof(h));
  return SMC_OK;
}

// handles: nranks consecutive 64-byte handles in rank order (the own one is ignored)
extern "C" int smc_ctx_p2p_import(smc_ctx_t ctx, const void* handles, int nranks, int rank) {
  if (!ctx || !handles || nranks < 2 || nranks > kMaxPeers || rank < 0 || rank >= nranks || !ctx->p2p_local) {
    smc_set_error("smc_ctx_p2p_import: bad argument, or smc_ctx_p2p_export was not called");
    return SMC_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  for (int q = 0; q < nranks; q++) {
    if (q == rank) { ctx->p2p_peer[q] = ctx->p2p_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)q * sizeof(h), sizeof(h));
    SMC_CUDA(cudaIpcOpenMemHandle(&ctx->p2p_peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->p2p_rank = rank;
  ctx->p2p_ranks = nranks;
  ctx->p2p_epoch = 0;
  return SMC_OK;
}

bool smc_p2p_ready(const smc_ctx_s* c, int64_t n) { return c->p2p_ranks > 1 && n <= kSlotDoubles; }

// in-stream sum over ranks of a device buffer of doubles (same contract as smc_comm_allreduce_device)
int smc_p2p_allreduce_device(smc_ctx_s* c, double* buf, int64_t n) {
  P2PView v;
  for (int q = 0; q < kMaxPeers; q++) v.peer[q] = (unsigned char*)(q < c->p2p_ranks ? c->p2p_peer[q] : nullptr);
  v.rank = c->p2p_rank; v.nranks = c->p2p_ranks; v.count = c->p2p_count; v.status = (int*)(c->p2p_count + 1);
  const unsigned epoch = ++c->p2p_epoch;
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 64);
  k_p2p_push<<<blocks, 256, 0, c->stream>>>(buf, (long long)n, v, epoch);
  k_p2p_pull<<<blocks, 256, 0, c->stream>>>(buf, (long long)n, v, epoch);
  c->launches += 2;
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

void smc_p2p_destroy(smc_ctx_s* c) {
  for (int q = 0; q < c->p2p_ranks; q++)
    if (q != c->p2p_rank && c->p2p_peer[q]) cudaIpcCloseMemHandle(c->p2p_peer[q]);
  if (c->p2p_local) cudaFree(c->p2p_local);
  if (c->p2p_count) cudaFree(c->p2p_count);
  c->p2p_local = nullptr; c->p2p_count = nullptr; c->p2p_ranks = 0;
}
