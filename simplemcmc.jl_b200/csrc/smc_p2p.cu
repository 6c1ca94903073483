// Peer-memory sum of the row-sharded GLM statement's (L, G) block over NVLink / NVSwitch (one node): replaces, per
// evaluation, k_glm_reduce + the in-stream NCCL all-reduce of smc_comm.cu + k_glm_publish by ONE kernel
// ("cfg3 rows sharded, few chains": at 4 GPUs and 2 chains a leapfrog was 164 us around 44 us of
// streaming; the row sum is src/distribs.jl:100-102 split over ranks).
//
// Opt-in: a context uses it after the host has called smc_ctx_p2p_export on every rank, exchanged the 64-byte handles, and
// called smc_ctx_p2p_import (simplemcmc.jl_b200/parallel.py::init_p2p).  Every rank owns one exchange buffer (plain
// cudaMalloc, exported with cudaIpcGetMemHandle and mapped by the peers):
//   flags[2][16] (unsigned, one per (parity, source rank)) and slots[2][R][kSlotDoubles].
// k_glm_xchg, epoch e (1, 2, ...; the same on all ranks because row-sharded ranks advance identical chains; it lives in
// DEVICE memory so that the kernel can sit in a captured CUDA graph and still count):
//   1. reduce: the fixed-order sum of this rank's row-split partials, element (j, c) by element;
//   2. push:   that sum goes into slot[e & 1][my rank] of EVERY rank (remote stores over NVLink), fence, and the last CTA
//              to finish writes e into flags[e & 1][my rank] of every rank (st.release.sys);
//   3. wait:   every CTA waits until flags[e & 1][q] >= e for every q (ld.acquire.sys, bounded: a peer that never
//              arrives sets a status word the host checks after the run, SMC_ERR_COMM);
//   4. sum + publish: out = sum over q in rank order of slot[e & 1][q] -- the same additions in the same order on every
//              rank, so the chains stay bit-identical -- written straight into the tape registers (L with the
//              out-of-support flag, G).
// Two parities suffice: a peer can push epoch e + 2 only after it finished epoch e + 1, which needs my push of e + 1, which
// my stream orders after my step 4 of e.
#include <string.h>
#include "smc_internal.h"

namespace {
constexpr int kMaxPeers = 16;
constexpr long long kSlotDoubles = 1ll << 19;            // 4 MB per (parity, rank): (M + 1) * C up to 524 288
constexpr size_t kFlagBytes = 256;                       // flags[2][16] unsigned = 128 bytes, padded
constexpr int kMaxBlocks = 64;                           // all CTAs must be co-resident (they wait on each other's push)

struct P2PView {
  unsigned char* peer[kMaxPeers];
  int rank, nranks;
  unsigned* ctl;                                         // local: [0] push tickets, [1] finish tickets, [2] epoch, [3] status
};
__device__ __forceinline__ unsigned* flag_ptr(unsigned char* base, int parity, int src) {
  return (unsigned*)base + parity * kMaxPeers + src;
}
__device__ __forceinline__ double* slot_ptr(unsigned char* base, int nranks, int parity, int src) {
  return (double*)(base + kFlagBytes) + ((long long)parity * nranks + src) * kSlotDoubles;
}

struct XchgArgs {
  const double* part;      // [RS][MP + 1][Cpad] row-split partials of this rank (NULL: `flat` holds n ready values)
  double* flat;            // generic all-reduce: in/out buffer of n doubles
  int RS, M, MP;
  long long C, Cpad, n;    // n = (M + 1) * C  (GLM)  or the buffer length (flat)
  int* bad;
  double* Ldst;
  double* Gdst;
  int acc;
};

__global__ void __launch_bounds__(256) k_glm_xchg(XchgArgs a, P2PView v) {
  __shared__ int s_flag;
  const unsigned epoch = *(volatile unsigned*)(v.ctl + 2) + 1u;       // updated only after every CTA has passed step 3
  const int parity = (int)(epoch & 1u);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // ---- 1 + 2: reduce and push
  if (a.part && a.RS >= 16) {
    // many row splits (the stream kernels write up to two per SM): a warp per element, the lanes stride the row splits and a
    // fixed shuffle tree combines them -- one thread per element would wait for RS dependent-latency loads in a row
    // (measured: 336 us per leapfrog against 288 us with the NCCL path's k_glm_reduce_warp, 2 GPUs x 2 chains)
    const int lane = threadIdx.x & 31;
    const long long nwarps = stride >> 5;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < a.n; i += nwarps) {
      const long long jj = i / a.C, c = i - jj * a.C;
      const long long j = jj == a.M ? a.MP : jj;
      double t = 0.0;
      for (int r = lane; r < a.RS; r += 32) t += a.part[((long long)r * (a.MP + 1) + j) * a.Cpad + c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (jj == a.M && (a.bad[c] || !(t > -INFINITY))) t = -INFINITY;
      if (lane < v.nranks) slot_ptr(v.peer[lane], v.nranks, parity, v.rank)[i] = t;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
      double t;
      if (a.part) {
        const long long jj = i / a.C, c = i - jj * a.C;
        const long long j = jj == a.M ? a.MP : jj;
        t = 0.0;
        for (int r = 0; r < a.RS; r++) t += a.part[((long long)r * (a.MP + 1) + j) * a.Cpad + c];
        if (jj == a.M && (a.bad[c] || !(t > -INFINITY))) t = -INFINITY;  // the flag survives the cross-rank sum
      } else {
        t = a.flat[i];
      }
      for (int q = 0; q < v.nranks; q++) slot_ptr(v.peer[q], v.nranks, parity, v.rank)[i] = t;
    }
  }
  __threadfence_system();                                // my stores are visible system-wide before my CTA is counted
  __syncthreads();
  if (threadIdx.x == 0) s_flag = (atomicAdd(v.ctl, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (s_flag) {
    __threadfence_system();
    if ((int)threadIdx.x < v.nranks) {
      unsigned* f = flag_ptr(v.peer[threadIdx.x], parity, v.rank);
      asm volatile("st.release.sys.global.u32 [%0], %1;\n" :: "l"(f), "r"(epoch) : "memory");
    }
    if (threadIdx.x == 0) v.ctl[0] = 0u;
  }
  // ---- 3: wait for every rank's flag
  unsigned char* mine = v.peer[v.rank];
  __syncthreads();
  if (threadIdx.x == 0) s_flag = 1;
  __syncthreads();
  if ((int)threadIdx.x < v.nranks) {
    const unsigned* f = flag_ptr(mine, parity, threadIdx.x);
    unsigned seen = 0;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(f) : "memory");
      if ((int)(seen - epoch) >= 0) break;
      if (clock64() - t0 > 20000000000ll) { s_flag = 0; break; }     // ~10 s: a peer that never arrives is an error
    }
  }
  __syncthreads();
  if (!s_flag && threadIdx.x == 0) v.ctl[3] = 1u;
  // ---- 4: rank-ordered sum and publish
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    double s = 0.0;
    for (int q = 0; q < v.nranks; q++) s += __ldcg(slot_ptr(mine, v.nranks, parity, q) + i);
    if (a.part) {
      const long long jj = i / a.C, c = i - jj * a.C;
      if (jj == a.M) {
        if (!(s > -INFINITY)) { a.bad[c] = 1; s = 0.0; }
        a.Ldst[c] = a.acc ? a.Ldst[c] + s : s;
      } else {
        a.Gdst[jj * a.Cpad + c] = s;
      }
    } else {
      a.flat[i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(v.ctl + 1, 1u) == gridDim.x - 1u) {
    v.ctl[1] = 0u;
    v.ctl[2] = epoch;                                     // every CTA has read the old value long ago (it pushed)
    __threadfence();
  }
}

P2PView make_view(const smc_ctx_s* c) {
  P2PView v;
  for (int q = 0; q < kMaxPeers; q++) v.peer[q] = (unsigned char*)(q < c->p2p_ranks ? c->p2p_peer[q] : nullptr);
  v.rank = c->p2p_rank; v.nranks = c->p2p_ranks; v.ctl = c->p2p_count;
  return v;
}
}  // namespace

// local exchange buffer for `nranks` ranks; writes its 64-byte IPC handle
extern "C" int smc_ctx_p2p_export(smc_ctx_t ctx, int nranks, void* handle_out, int handle_bytes) {
  if (!ctx || !handle_out || handle_bytes < (int)sizeof(cudaIpcMemHandle_t) || nranks < 2 || nranks > kMaxPeers) {
    smc_set_error("smc_ctx_p2p_export: bad argument (2..%d ranks, %d-byte handle)", kMaxPeers, (int)sizeof(cudaIpcMemHandle_t));
    return SMC_ERR_INVALID;
  }
  cudaSetDevice(ctx->device);
  if (ctx->p2p_local && ctx->p2p_alloc_ranks != nranks) {
    smc_set_error("smc_ctx_p2p_export: the exchange buffer of this context was sized for %d ranks, not %d", ctx->p2p_alloc_ranks, nranks);
    return SMC_ERR_INVALID;
  }
  if (!ctx->p2p_local) {
    const size_t bytes = kFlagBytes + (size_t)2 * nranks * kSlotDoubles * sizeof(double);
    SMC_CUDA(cudaMalloc(&ctx->p2p_local, bytes));        // not from the stream-ordered pool: pool memory has no IPC handle
    SMC_CUDA(cudaMemset(ctx->p2p_local, 0, bytes));
    SMC_CUDA(cudaMalloc((void**)&ctx->p2p_count, 4 * sizeof(unsigned)));
    SMC_CUDA(cudaMemset(ctx->p2p_count, 0, 4 * sizeof(unsigned)));
    SMC_CUDA(cudaDeviceSynchronize());
    ctx->p2p_alloc_ranks = nranks;
  }
  cudaIpcMemHandle_t h;
  SMC_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
  memcpy(handle_out, &h, sizeof(h));
  return SMC_OK;
}

// handles: nranks consecutive 64-byte handles in rank order (the own one is ignored).  Once per context: the epoch
// counters of the ranks advance together from zero and the flags of an earlier set-up would satisfy later waits.
extern "C" int smc_ctx_p2p_import(smc_ctx_t ctx, const void* handles, int nranks, int rank) {
  if (!ctx || !handles || nranks < 2 || nranks > kMaxPeers || rank < 0 || rank >= nranks || !ctx->p2p_local) {
    smc_set_error("smc_ctx_p2p_import: bad argument, or smc_ctx_p2p_export was not called");
    return SMC_ERR_INVALID;
  }
  if (ctx->p2p_ranks != 0) { smc_set_error("smc_ctx_p2p_import: this context already has a peer exchange (one set-up per context)"); return SMC_ERR_INVALID; }
  if (nranks != ctx->p2p_alloc_ranks) {
    smc_set_error("smc_ctx_p2p_import: %d ranks, but the exchange buffer was exported for %d", nranks, ctx->p2p_alloc_ranks);
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
  return SMC_OK;
}

bool smc_p2p_ready(const smc_ctx_s* c, int64_t n) { return c->p2p_ranks > 1 && n <= kSlotDoubles; }

// the GLM statement's row-split partials -> summed over row splits and ranks -> tape registers, one kernel
int smc_p2p_glm_exchange(smc_ctx_s* c, const double* part, int RS, int M, int MP, int64_t C, int64_t Cpad, int* bad, double* Ldst,
                         double* Gdst, int acc) {
  XchgArgs a;
  a.part = part; a.flat = nullptr; a.RS = RS; a.M = M; a.MP = MP; a.C = C; a.Cpad = Cpad; a.n = (long long)(M + 1) * C;
  a.bad = bad; a.Ldst = Ldst; a.Gdst = Gdst; a.acc = acc;
  const long long threads = RS >= 16 ? a.n * 32 : a.n;      // a warp per element when there are many row splits to sum
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((threads + 255) / 256, kMaxBlocks));
  k_glm_xchg<<<blocks, 256, 0, c->stream>>>(a, make_view(c));
  c->launches += 1;
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

// in-stream sum over ranks of a device buffer of doubles (same contract as smc_comm_allreduce_device)
int smc_p2p_allreduce_device(smc_ctx_s* c, double* buf, int64_t n) {
  XchgArgs a;
  memset(&a, 0, sizeof(a));
  a.flat = buf; a.n = n;
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, kMaxBlocks));
  k_glm_xchg<<<blocks, 256, 0, c->stream>>>(a, make_view(c));
  c->launches += 1;
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

// after a synchronisation point: did any exchange of this context time out?  (clears the word)
int smc_p2p_check(smc_ctx_s* c) {
  if (c->p2p_ranks <= 1 || !c->p2p_count) return SMC_OK;
  unsigned status = 0;
  SMC_CUDA(smc_copy_sync(&status, c->p2p_count + 3, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  if (status) {
    cudaMemsetAsync(c->p2p_count + 3, 0, sizeof(unsigned), c->stream);
    smc_set_error("peer-memory exchange: a rank did not arrive within the time limit (the results of this call are invalid)");
    return SMC_ERR_COMM;
  }
  return SMC_OK;
}

void smc_p2p_destroy(smc_ctx_s* c) {
  for (int q = 0; q < c->p2p_ranks; q++)
    if (q != c->p2p_rank && c->p2p_peer[q]) cudaIpcCloseMemHandle(c->p2p_peer[q]);
  if (c->p2p_local) cudaFree(c->p2p_local);
  if (c->p2p_count) cudaFree(c->p2p_count);
  c->p2p_local = nullptr; c->p2p_count = nullptr; c->p2p_ranks = 0; c->p2p_alloc_ranks = 0;
}
