// libsimplemcmc_cuda.so core: contexts, tape loading, register allocation, the generic statement
// executor (one launch per 3-address statement, every chain in lockstep) and smc_eval.
// Reference path replaced: the generated closure of /root/reference/src/parsing.jl:409-492.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <mutex>
#include "smc_internal.h"
#include "dists.cuh"
#include "exec_common.cuh"

static thread_local char g_err[1024] = "";
void smc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* smc_last_error(void) { return g_err; }
extern "C" int smc_abi_version(void) { return SMC_ABI_VERSION; }

// ------------------------------------------------------------------------------------------ device memory
namespace {
std::mutex g_pool_mu;
std::vector<void*> g_deferred_free;  // frees requested while the stream was capturing
int g_pool_mode = -1;                // -1 not decided, 0 cudaMalloc/cudaFree, 1 stream-ordered pool
bool g_pool_ready[64] = {};

bool pool_enabled() {
  if (g_pool_mode < 0) {
    const char* e = getenv("SMC_NO_POOL");
    g_pool_mode = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return g_pool_mode == 1;
}
void pool_configure() {  // once per device: keep freed blocks in the pool instead of returning them to the driver
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_pool_ready[dev]) return;
  g_pool_ready[dev] = true;
  int supported = 0;
  cudaMemPool_t pool;
  if (cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) != cudaSuccess || !supported ||
      cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) {
    cudaGetLastError();
    g_pool_mode = 0;
    return;
  }
  uint64_t keep = 8ull << 30;
  if (const char* e = getenv("SMC_POOL_KEEP_MB")) keep = (uint64_t)std::max(0ll, atoll(e)) << 20;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}
bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus s = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &s) != cudaSuccess) { cudaGetLastError(); return false; }
  return s != cudaStreamCaptureStatusNone;
}
}  // namespace

cudaError_t smc_dev_alloc_bytes(void** p, size_t bytes, cudaStream_t st) {
  bytes = std::max<size_t>(bytes, 8);
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (pool_enabled()) pool_configure();
  if (!pool_enabled() || stream_capturing(st)) return cudaMalloc(p, bytes);
  if (!g_deferred_free.empty()) {
    for (void* q : g_deferred_free) cudaFree(q);
    g_deferred_free.clear();
  }
  cudaError_t e = cudaMallocAsync(p, bytes, st);
  if (e == cudaErrorMemoryAllocation) {  // give cached blocks back to the driver and retry once
    cudaGetLastError();
    cudaStreamSynchronize(st);
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMallocAsync(p, bytes, st);
  }
  return e;
}

void smc_dev_free(void* p, cudaStream_t st) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!pool_enabled()) { cudaFree(p); return; }
  if (stream_capturing(st)) { g_deferred_free.push_back(p); return; }
  if (cudaFreeAsync(p, st) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
}

cudaError_t smc_copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  if (stream_capturing(st)) return cudaMemcpy(dst, src, bytes, kind);  // a capture cannot be synchronised
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}

// ------------------------------------------------------------------------------------------ context
extern "C" int smc_ctx_create(int device, smc_ctx_t* out) {
  if (!out) { smc_set_error("smc_ctx_create: out is NULL"); return SMC_ERR_INVALID; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    smc_set_error("smc_ctx_create: no CUDA device (%s); libsimplemcmc_cuda has no CPU fallback", cudaGetErrorString(e));
    return SMC_ERR_CUDA;
  }
  if (device < 0 || device >= n) { smc_set_error("smc_ctx_create: device %d out of range (%d devices)", device, n); return SMC_ERR_INVALID; }
  SMC_CUDA(cudaSetDevice(device));
  smc_ctx_s* c = new smc_ctx_s();
  c->device = device;
  cudaDeviceProp p;
  SMC_CUDA(cudaGetDeviceProperties(&p, device));
  c->sm_count = p.multiProcessorCount;
  SMC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = c;
  return SMC_OK;
}

int smc_comm_destroy(smc_ctx_s* c);  // smc_comm.cu
extern "C" int smc_ctx_destroy(smc_ctx_t c) {
  if (!c) return SMC_OK;
  cudaSetDevice(c->device);
  smc_comm_destroy(c);
  smc_p2p_destroy(c);
  if (c->flush_buf) smc_dev_free(c->flush_buf, c->stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SMC_OK;
}

// ------------------------------------------------------------------------------------------ tape loading
static int64_t reg_numel(const smc_tape_reg& r) { return (int64_t)std::max(1u, r.rows) * (int64_t)std::max(1u, r.cols); }

extern "C" int smc_model_compile(smc_ctx_t ctx, const void* tape, size_t nbytes, smc_model_t* out) {
  if (!ctx || !tape || !out) { smc_set_error("smc_model_compile: NULL argument"); return SMC_ERR_INVALID; }
  const uint8_t* p = (const uint8_t*)tape;
  const uint8_t* end = p + nbytes;
  if (nbytes < sizeof(smc_tape_header)) { smc_set_error("tape: truncated header"); return SMC_ERR_INVALID; }
  smc_model_s* m = new smc_model_s();
  m->ctx = ctx;
  memcpy(&m->hdr, p, sizeof(smc_tape_header));
  p += sizeof(smc_tape_header);
  const smc_tape_header& h = m->hdr;
  if (h.magic != SMC_TAPE_MAGIC || h.version != SMC_TAPE_VERSION) {
    smc_set_error("tape: bad magic/version (%08x v%u, expected v%u)", h.magic, h.version, SMC_TAPE_VERSION);
    delete m;
    return SMC_ERR_INVALID;
  }
  size_t need = (size_t)h.n_regs * sizeof(smc_tape_reg) + (size_t)h.n_stmts * sizeof(smc_tape_stmt) +
                (size_t)h.n_data * sizeof(smc_tape_data);
  if ((size_t)(end - p) < need) { smc_set_error("tape: truncated tables"); delete m; return SMC_ERR_INVALID; }
  m->regs.resize(h.n_regs);
  for (uint32_t i = 0; i < h.n_regs; i++) {
    memcpy(&m->regs[i].d, p, sizeof(smc_tape_reg));
    p += sizeof(smc_tape_reg);
    m->regs[i].numel = reg_numel(m->regs[i].d);
    if (m->regs[i].d.kind > SMC_REG_GRAD) { smc_set_error("tape: register %u has unknown kind", i); delete m; return SMC_ERR_INVALID; }
  }
  m->stmts.resize(h.n_stmts);
  for (uint32_t i = 0; i < h.n_stmts; i++) {
    memcpy(&m->stmts[i], p, sizeof(smc_tape_stmt));
    p += sizeof(smc_tape_stmt);
    const smc_tape_stmt& s = m->stmts[i];
    bool ok = s.dst < h.n_regs && s.nsrc <= 4;
    for (uint32_t k = 0; ok && k < s.nsrc; k++) ok = s.src[k] < h.n_regs;
    if (s.op == SMC_OP_GLM) ok = ok && s.aux2 < h.n_regs && s.nsrc == 3;
    if (s.op == SMC_OP_AFFNORM) ok = ok && s.nsrc == 2 && s.aux0 < h.n_idx;
    if (!ok) { smc_set_error("tape: statement %u references a register out of range", i); delete m; return SMC_ERR_INVALID; }
    uint32_t dk = m->regs[s.dst].d.kind;
    if ((s.flags & SMC_F_FWD_ONLY) && (i >= h.reserved[0] || dk != SMC_REG_CHAIN)) {
      smc_set_error("tape: statement %u is flagged forward-only but belongs to the gradient part (or to the header)", i);
      delete m;
      return SMC_ERR_INVALID;
    }
    if (dk == SMC_REG_HDR) m->hdr_stmts.push_back((int)i);
    else if (dk == SMC_REG_CHAIN || dk == SMC_REG_GRAD) m->body_stmts.push_back((int)i);
    else { smc_set_error("tape: statement %u writes a read-only register", i); delete m; return SMC_ERR_INVALID; }
  }
  m->data.resize(h.n_data);
  for (uint32_t i = 0; i < h.n_data; i++) {
    smc_tape_data d;
    memcpy(&d, p, sizeof(d));
    p += sizeof(d);
    d.name[sizeof(d.name) - 1] = 0;
    m->data[i].name = d.name;
    m->data[i].rows = d.rows;
    m->data[i].cols = d.cols;
  }
  for (uint32_t i = 0; i < h.n_regs; i++)
    if (m->regs[i].d.kind == SMC_REG_DATA) {
      int64_t slot = m->regs[i].d.aux;
      if (slot < 0 || slot >= (int64_t)h.n_data) { smc_set_error("tape: data register %u has a bad slot", i); delete m; return SMC_ERR_INVALID; }
      m->data[slot].reg = (int)i;
    }
  cudaSetDevice(ctx->device);
  for (uint32_t i = 0; i < h.n_idx; i++) {
    if ((size_t)(end - p) < 8) { smc_set_error("tape: truncated index table"); delete m; return SMC_ERR_INVALID; }
    uint32_t n;
    memcpy(&n, p, 4);
    p += 8;
    size_t bytes = (size_t)((n + 1) & ~1u) * 4;
    if ((size_t)(end - p) < bytes) { smc_set_error("tape: truncated index table"); delete m; return SMC_ERR_INVALID; }
    int32_t* d = nullptr;
    SMC_CUDA(smc_dev_alloc(&d, std::max<size_t>(bytes, 8), ctx->stream));
    SMC_CUDA(smc_copy_sync(d, p, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    m->idx_host.emplace_back();
    if (n <= 64) { m->idx_host.back().resize(n); memcpy(m->idx_host.back().data(), p, (size_t)n * 4); }
    p += bytes;
    m->idx_dev.push_back(d);
    m->idx_len.push_back(n);
  }
  for (uint32_t i = 0; i < h.reserved[1]; i++) {
    SmcProgram pr;
    if ((size_t)(end - p) < sizeof(smc_prog_header)) { smc_set_error("tape: truncated program"); delete m; return SMC_ERR_INVALID; }
    memcpy(&pr.h, p, sizeof(smc_prog_header));
    p += sizeof(smc_prog_header);
    size_t bytes = (size_t)pr.h.n_instr * sizeof(smc_prog_instr) + (size_t)pr.h.n_red * sizeof(smc_prog_red);
    if ((size_t)(end - p) < bytes || pr.h.n_red > SMC_PROG_MAX_REDUCE || pr.h.n_prologue > pr.h.n_instr || pr.h.n_instr > 1024) {
      smc_set_error("tape: malformed program %u", i); delete m; return SMC_ERR_INVALID;
    }
    pr.instr.resize(pr.h.n_instr);
    memcpy(pr.instr.data(), p, (size_t)pr.h.n_instr * sizeof(smc_prog_instr));
    p += (size_t)pr.h.n_instr * sizeof(smc_prog_instr);
    pr.red.resize(pr.h.n_red);
    memcpy(pr.red.data(), p, (size_t)pr.h.n_red * sizeof(smc_prog_red));
    p += (size_t)pr.h.n_red * sizeof(smc_prog_red);
    for (auto& in : pr.instr) {
      bool ok = in.reg < h.n_regs && in.dst < SMC_PROG_MAX_LOCALS && in.nsrc <= 3;
      for (int k = 0; ok && k < in.nsrc; k++) ok = in.src[k] < SMC_PROG_MAX_LOCALS;
      if (in.op == SMC_P_REDUCE) ok = ok && in.aux0 < pr.h.n_red;
      if (!ok) { smc_set_error("tape: program %u has an operand out of range", i); delete m; return SMC_ERR_INVALID; }
    }
    for (auto& r : pr.red)
      if (r.reg >= h.n_regs) { smc_set_error("tape: program %u reduces into a bad register", i); delete m; return SMC_ERR_INVALID; }
    if (smc_vm_plan(m, pr) != SMC_OK) { delete m; return SMC_ERR_INVALID; }
    m->programs.push_back(pr);
  }
  for (auto& s : m->stmts)
    if (s.op == SMC_OP_FUSED && s.aux0 >= m->programs.size()) { smc_set_error("tape: FUSED statement without a program"); delete m; return SMC_ERR_INVALID; }
  if (h.result_reg >= h.n_regs) { smc_set_error("tape: bad result register"); delete m; return SMC_ERR_INVALID; }
  m->M = h.n_params;
  *out = m;
  return SMC_OK;
}

extern "C" int smc_model_nparams(smc_model_t m, int64_t* out) {
  if (!m || !out) return SMC_ERR_INVALID;
  *out = m->M;
  return SMC_OK;
}

extern "C" int smc_model_destroy(smc_model_t m) {
  if (!m) return SMC_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  smc_graphs_invalidate(m);
  for (auto& r : m->regs) if (r.owned && r.dev) smc_dev_free(r.dev, m->ctx->stream);
  for (auto d : m->idx_dev) smc_dev_free(d, m->ctx->stream);
  for (auto& p : m->scatter) { smc_dev_free(p.ptr, m->ctx->stream); smc_dev_free(p.list, m->ctx->stream); }
  for (auto& p : m->sparse) { if (p.ptr) smc_dev_free(p.ptr, m->ctx->stream); if (p.col) smc_dev_free(p.col, m->ctx->stream); if (p.val) smc_dev_free(p.val, m->ctx->stream); }
  for (auto& g : m->glm) smc_glm_release(g, m->ctx->stream);
  for (auto& a : m->aff) smc_aff_release(a, m->ctx->stream);
  for (auto& pr : m->programs) if (pr.dev_tables) smc_dev_free(pr.dev_tables, m->ctx->stream);
  if (m->beta) smc_dev_free(m->beta, m->ctx->stream);
  if (m->grad) smc_dev_free(m->grad, m->ctx->stream);
  if (m->logp) smc_dev_free(m->logp, m->ctx->stream);
  if (m->bad) smc_dev_free(m->bad, m->ctx->stream);
  if (m->scratch) smc_dev_free(m->scratch, m->ctx->stream);
  if (m->host_stage) cudaFreeHost(m->host_stage);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  delete m;
  return SMC_OK;
}

extern "C" int smc_model_bind_data(smc_model_t m, const char* name, const double* host, const int64_t* dims, int ndims) {
  if (!m || !name || !host || ndims < 0 || ndims > 2) { smc_set_error("smc_model_bind_data: bad argument"); return SMC_ERR_INVALID; }
  cudaSetDevice(m->ctx->device);
  for (auto& d : m->data) {
    if (d.name != name) continue;
    int64_t rows = ndims >= 1 ? dims[0] : 1, cols = ndims >= 2 ? dims[1] : 1;
    if (rows * cols != d.rows * d.cols) {
      smc_set_error("smc_model_bind_data: '%s' has %lld elements, the tape was compiled for %lld", name,
                    (long long)(rows * cols), (long long)(d.rows * d.cols));
      return SMC_ERR_INVALID;
    }
    SmcReg& r = m->regs[d.reg];
    if (!r.dev) { SMC_CUDA(smc_dev_alloc(&r.dev, std::max<int64_t>(r.numel, 1) * sizeof(double), m->ctx->stream)); r.owned = true; }
    SMC_CUDA(cudaMemcpyAsync(r.dev, host, r.numel * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    SMC_CUDA(cudaStreamSynchronize(m->ctx->stream));
    d.bound = true;
    m->finalized = false;
    return SMC_OK;
  }
  smc_set_error("smc_model_bind_data: the tape has no external variable named '%s'", name);
  return SMC_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------ generic executor
struct EwArgs {
  int op, dist, darg, nsrc, acc;
  Opd s[4];
  Dst d;
  long long n, C;
  int* bad;
  int bad_stride;  // 1: per chain, 0: header phase
};

__device__ __forceinline__ double ew_value(const EwArgs& a, long long e, long long c) {
  double x0 = a.nsrc > 0 ? opd_ld(a.s[0], e, c) : 0.0;
  double x1 = a.nsrc > 1 ? opd_ld(a.s[1], e, c) : 0.0;
  double x2 = a.nsrc > 2 ? opd_ld(a.s[2], e, c) : 0.0;
  double v = smc_apply(a.op, a.dist, a.darg, x0, x1, x2);
  if (a.op == SMC_OP_LOGPDF && (v == -INFINITY || isnan(v))) {  // "give up eval", src/distribs.jl:69-70
    a.bad[c * a.bad_stride] = 1;
    v = 0.0;
  }
  return v;
}

__global__ void __launch_bounds__(256) k_elementwise(EwArgs a) {
  long long total = a.n * a.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long e = i / a.C, c = i - e * a.C;
    double v = ew_value(a, e, c);
    double* q = a.d.p + e * a.d.se + c * a.d.sc;
    *q = a.acc ? (*q + v) : v;
  }
}

// chain-parallel reduction: thread = (chain, split); sums its element range left to right
__global__ void __launch_bounds__(256) k_reduce_chains(EwArgs a, double* partial, long long chunk) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  long long s = blockIdx.y, e0 = s * chunk, e1 = min(a.n, e0 + chunk);
  double acc = 0.0;
  for (long long e = e0; e < e1; e++) acc += ew_value(a, e, c);
  partial[s * a.C + c] = acc;
}

// element-parallel reduction: block = (split, chain)
__global__ void __launch_bounds__(256) k_reduce_elems(EwArgs a, double* partial, long long chunk) {
  long long c = blockIdx.y, s = blockIdx.x, e0 = s * chunk, e1 = min(a.n, e0 + chunk);
  double acc = 0.0;
  for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) acc += ew_value(a, e, c);
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    partial[s * a.C + c] = t;
  }
}

// dst[o][c] (=|+=) sum_s partial[s][o][c], fixed order -> deterministic
__global__ void k_reduce_final(const double* partial, long long S, long long outn, long long C, Dst d, int acc) {
  long long total = outn * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long o = i / C, c = i - o * C;
    double t = 0.0;
    for (long long s = 0; s < S; s++) t += partial[(s * outn + o) * C + c];
    double* q = d.p + o * d.se + c * d.sc;
    *q = acc ? (*q + t) : t;
  }
}

struct MmArgs {
  Opd A, B;
  long long ar, br;  // stored leading dimension (rows) of A and B
  int tA, tB;
  long long r, k, n, C;
};
// out(i,j) = sum_l A(i,l) B(l,j), column major, split over l; partial[s][i + j*r][c]
// `direct` (one split only): the sum goes straight to the destination (=|+=) instead of a partial buffer that k_reduce_final
// would copy -- the small products of hierarchical models (k = 1 .. 5) are launch bound, not flop bound
__global__ void __launch_bounds__(256) k_matmul(MmArgs a, double* partial, long long chunk, Dst d, int dacc, int direct) {
  long long total = a.r * a.n * a.C;
  long long s = blockIdx.y, l0 = s * chunk, l1 = min(a.k, l0 + chunk);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long o = idx / a.C, c = idx - o * a.C;
    long long i = o % a.r, j = o / a.r;
    double acc = 0.0;
    for (long long l = l0; l < l1; l++) {
      long long ea = a.tA ? (l + i * a.ar) : (i + l * a.ar);
      long long eb = a.tB ? (j + l * a.br) : (l + j * a.br);
      acc += opd_ld(a.A, ea, c) * opd_ld(a.B, eb, c);
    }
    if (direct) {
      double* q = d.p + o * d.se + c * d.sc;
      *q = dacc ? (*q + acc) : acc;
    } else {
      partial[(s * a.r * a.n + o) * a.C + c] = acc;
    }
  }
}

// MATMUL with a chain-independent left operand (data or header matrix) and a per-chain right operand -- the `X * beta`
// and `transpose(X) * ds` of the generated code (src/diff.jl:52-55) outside the fused GLM pattern.  With the chains as
// the contiguous axis this is the GEMM  out[r x (n*C)] = Aop[r x k] * B[k x (n*C)]: thread = chain (coalesced B and out),
// RI output rows per thread in registers, the A tile staged in shared memory and read as broadcast double2.
struct MmDcArgs {
  const double* A;
  long long ar;   // stored leading dimension of A
  int tA;
  const double* B;
  double* out;
  const double* B2;   // a second product with the same left operand and shapes (upper half of grid.y), or NULL
  double* out2;
  long long Bse, Ose;   // element strides (Cpad)
  long long r, k, n, C, chunk;
  int acc, S;
  double* partial;      // [S][r*n][C] when S > 1
};
template <int RI>
__global__ void __launch_bounds__(256) k_matmul_dc(MmDcArgs a) {
  constexpr int TK = 32;
  extern __shared__ __align__(16) double As[];   // [ty][RI][TK]
  const int tcx = blockDim.x, ty = blockDim.y, tid = threadIdx.y * tcx + threadIdx.x;
  const long long c = (long long)blockIdx.x * tcx + threadIdx.x;
  const long long groups = (a.r + RI - 1) / RI;
  const long long gtiles = (groups + ty - 1) / ty;
  const int second = blockIdx.y >= gtiles * a.n ? 1 : 0;        // (only when B2 is set: the grid is twice as tall)
  const long long by = blockIdx.y - (second ? gtiles * a.n : 0);
  const double* const Bm = second ? a.B2 : a.B;
  double* const outm = second ? a.out2 : a.out;
  const long long j = by / gtiles;                              // output column
  const long long g0 = (by - j * gtiles) * ty;                  // first row group of this block
  const long long i0 = (g0 + threadIdx.y) * RI;
  const long long l0 = (long long)blockIdx.z * a.chunk, l1 = min(a.k, l0 + a.chunk);
  double acc[RI];
#pragma unroll
  for (int ii = 0; ii < RI; ii++) acc[ii] = 0.0;
  const bool cok = c < a.C;
  const double* Bc = Bm + (j * a.k) * a.Bse + (cok ? c : 0);
  for (long long lb = l0; lb < l1; lb += TK) {
    __syncthreads();
    for (int t = tid; t < ty * RI * TK; t += 256) {
      int ll = t % TK, ii = (t / TK) % RI, y = t / (TK * RI);
      long long i = (g0 + y) * RI + ii, l = lb + ll;
      double v = 0.0;
      if (i < a.r && l < l1) v = __ldg(a.A + (a.tA ? (l + i * a.ar) : (i + l * a.ar)));
      As[t] = v;
    }
    __syncthreads();
    const double* as = As + threadIdx.y * RI * TK;
    const int lim = (int)min((long long)TK, l1 - lb);
    if (cok && i0 < a.r) {
      for (int ll = 0; ll < lim; ll += 2) {
        const double b0 = __ldg(Bc + (lb + ll) * a.Bse);
        const double b1 = (ll + 1 < lim) ? __ldg(Bc + (lb + ll + 1) * a.Bse) : 0.0;
#pragma unroll
        for (int ii = 0; ii < RI; ii++) {
          const double2 a2 = *(const double2*)(as + ii * TK + ll);
          acc[ii] = fma(a2.x, b0, acc[ii]);
          acc[ii] = fma(a2.y, b1, acc[ii]);
        }
      }
    }
  }
  if (!cok) return;
#pragma unroll
  for (int ii = 0; ii < RI; ii++) {
    long long i = i0 + ii;
    if (i >= a.r) break;
    long long o = i + j * a.r;
    if (a.S > 1) {
      a.partial[((long long)blockIdx.z * a.r * a.n + o) * a.C + c] = acc[ii];
    } else {
      double* q = outm + o * a.Ose + c;
      *q = a.acc ? (*q + acc[ii]) : acc[ii];
    }
  }
}

// MATMUL whose parameter-independent left operand is sparse (a one-hot design matrix such as sleepstudy's `submatrix`:
// one non-zero per row): CSR of op(A), thread = (output element, chain).  The non-zeros of a row are summed in column
// order, i.e. the dense sum with its exact +0.0 terms dropped.
__global__ void k_spmm_dc(const int32_t* ptr, const int32_t* col, const double* val, const double* B, long long Bse, double* out,
                          long long Ose, long long r, long long n, long long k, long long C, int acc, const double* B2, double* out2) {
  const long long one = r * n * C, total = B2 ? 2 * one : one;     // (B2, out2): a second product with the same matrix
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const bool second = idx >= one;
    const double* const Bm = second ? B2 : B;
    double* const outm = second ? out2 : out;
    const long long id1 = second ? idx - one : idx;
    long long o = id1 / C, c = id1 - o * C;
    long long i = o % r, j = o / r;
    double t = 0.0;
    for (int p = ptr[i]; p < ptr[i + 1]; p++) t = fma(val[p], __ldg(Bm + ((long long)col[p] + j * k) * Bse + c), t);
    double* q = outm + o * Ose + c;
    *q = acc ? (*q + t) : t;
  }
}

__global__ void k_transpose(Opd src, Dst d, long long rows, long long cols, long long C, int acc) {
  long long total = rows * cols * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx / C, c = idx - e * C;
    long long i = e % rows, j = e / rows;  // element (i,j) of src -> (j,i) of dst
    double v = opd_ld(src, e, c);
    double* q = d.p + (j + i * cols) * d.se + c * d.sc;
    *q = acc ? (*q + v) : v;
  }
}

__global__ void k_gather(Opd src, Dst d, const int32_t* idx, long long n, long long C, int acc) {
  long long total = n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long e = i / C, c = i - e * C;
    double v = opd_ld(src, idx[e], c);
    double* q = d.p + e * d.se + c * d.sc;
    *q = acc ? (*q + v) : v;
  }
}

// Scatter through the inverse of the static index table (CSR: for every destination element the source elements that
// map to it, in list order): thread = (destination element, chain), coalesced over chains.  Repeated indices resolve
// exactly as a sequential pass over the list would -- `+=` in list order for the corrected adjoint (SCATTER_ADD), last
// writer wins for the reference's assignment (SCATTER_SET, src/parsing.jl:277-281); untouched elements keep their value.
__global__ void k_scatter_csr(Opd src, Dst d, const int32_t* ptr, const int32_t* list, long long n_dst, long long C, int add) {
  long long total = n_dst * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long e = i / C, c = i - e * C;
    int b = ptr[e], t = ptr[e + 1];
    if (b == t) continue;
    double* q = d.p + e * d.se + c * d.sc;
    if (add) {
      double acc = *q;
      for (int k = b; k < t; k++) acc += opd_ld(src, list[k], c);
      *q = acc;
    } else {
      *q = opd_ld(src, list[t - 1], c);
    }
  }
}

__global__ void k_fill(double* p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_int(int* p, long long n, int v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// logp[c] = bad ? -inf : result ; grad[:,c] = bad ? 0 : grad   (src/parsing.jl:465-467)
__global__ void k_finish_eval(double* result, long long rse, int rsc, int* bad, double* logp, double* grad,
                              long long M, long long C, long long Cpad, int with_grad) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double r = result[c * rsc];
  int b = bad[c] || !(r > -INFINITY);   // a -Inf / NaN total is a zero density as well ("give up eval")
  bad[c] = 0;                           // flags are left clean for the next evaluation
  result[c * rsc] = 0.0;                // ... and so is the accumulator: every statement may add to it (tape version 4)
  logp[c] = b ? -INFINITY : r;
  if (b && with_grad)
    for (long long j = 0; j < M; j++) grad[j * Cpad + c] = 0.0;
}

// [C][M] row major  <->  [M][Cpad]
__global__ void k_rows_to_cols(const double* in, double* out, long long C, long long M, long long Cpad, int shared_init) {
  long long total = C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    out[j * Cpad + c] = shared_init ? in[j] : in[c * M + j];
  }
}
__global__ void k_cols_to_rows(const double* in, double* out, long long C, long long M, long long Cpad) {
  long long total = C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long c = i / M, j = i - c * M;
    out[i] = in[j * Cpad + c];
  }
}


static inline unsigned grid_for(long long total, int block = 256, long long cap = 148LL * 32) {
  long long g = (total + block - 1) / block;
  return (unsigned)std::max<long long>(1, std::min<long long>(g, cap));
}

void smc_graphs_invalidate(smc_model_s* m) {
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  m->graphs_stale = false;
}

int smc_scratch_reserve(smc_model_s* m, int64_t doubles) {
  if (doubles <= m->scratch_cap) return SMC_OK;
  m->graphs_stale = true;  // captured evaluations hold the old pointer
  m->realloc_epoch++;
  if (m->scratch) smc_dev_free(m->scratch, m->ctx->stream);
  m->scratch = nullptr;
  m->scratch_cap = 0;
  SMC_CUDA(smc_dev_alloc(&m->scratch, doubles * sizeof(double), m->ctx->stream));
  m->scratch_cap = doubles;
  return SMC_OK;
}

int smc_stage_reserve(smc_model_s* m, int64_t doubles) {
  if (doubles <= m->host_stage_cap) return SMC_OK;
  if (m->host_stage) cudaFreeHost(m->host_stage);
  m->host_stage = nullptr;
  m->host_stage_cap = 0;
  SMC_CUDA(cudaMallocHost(&m->host_stage, doubles * sizeof(double)));
  m->host_stage_cap = doubles;
  return SMC_OK;
}

Opd smc_make_opd(const smc_model_s* m, uint32_t r) {
  const SmcReg& g = m->regs[r];
  Opd o;
  o.p = nullptr; o.se = 0; o.sc = 0; o.is_imm = 0; o.imm = 0.0;
  switch (g.d.kind) {
    case SMC_REG_IMM: o.is_imm = 1; o.imm = g.d.imm; break;
    case SMC_REG_DATA:
    case SMC_REG_HDR: o.p = g.dev; o.se = g.numel > 1 ? 1 : 0; break;
    case SMC_REG_CHAIN: o.p = g.dev; o.se = g.numel > 1 ? m->Cpad : 0; o.sc = 1; break;
    case SMC_REG_PARAM: o.p = m->beta + g.d.aux * m->Cpad; o.se = g.numel > 1 ? m->Cpad : 0; o.sc = 1; break;
    case SMC_REG_GRAD: o.p = m->grad + g.d.aux * m->Cpad; o.se = g.numel > 1 ? m->Cpad : 0; o.sc = 1; break;
  }
  return o;
}
Dst smc_make_dst(const smc_model_s* m, uint32_t r) {
  Opd o = smc_make_opd(m, r);
  Dst d;
  d.p = const_cast<double*>(o.p);
  d.se = o.se;
  d.sc = o.sc;
  // a scalar destination still needs a valid element stride of 0 and chain stride as computed
  return d;
}

// inverse (CSR) of index table `table` for a destination of n_dst elements, built on first use
static int scatter_plan(smc_model_s* m, uint32_t table, int64_t n_dst, SmcScatterPlan** out) {
  for (auto& p : m->scatter)
    if (p.table == (int)table && p.n_dst == n_dst) { *out = &p; return SMC_OK; }
  const uint32_t n = m->idx_len[table];
  std::vector<int32_t> idx(n);
  SMC_CUDA(smc_copy_sync(idx.data(), m->idx_dev[table], (size_t)n * 4, cudaMemcpyDeviceToHost, m->ctx->stream));
  std::vector<int32_t> ptr(n_dst + 1, 0), list(std::max<uint32_t>(n, 1));
  for (uint32_t e = 0; e < n; e++) {
    if (idx[e] < 0 || idx[e] >= n_dst) { smc_set_error("tape: scatter index out of range"); return SMC_ERR_INVALID; }
    ptr[idx[e] + 1]++;
  }
  for (int64_t k = 0; k < n_dst; k++) ptr[k + 1] += ptr[k];
  std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
  for (uint32_t e = 0; e < n; e++) list[fill[idx[e]]++] = (int32_t)e;   // stable: list order within a destination
  SmcScatterPlan p;
  p.table = (int)table; p.n_dst = n_dst;
  SMC_CUDA(smc_dev_alloc(&p.ptr, ptr.size() * 4, m->ctx->stream));
  SMC_CUDA(smc_dev_alloc(&p.list, list.size() * 4, m->ctx->stream));
  SMC_CUDA(smc_copy_sync(p.ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice, m->ctx->stream));
  SMC_CUDA(smc_copy_sync(p.list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, m->ctx->stream));
  m->scatter.push_back(p);
  *out = &m->scatter.back();
  return SMC_OK;
}

// CSR of op(A) for a data / header matrix operand, built on first use; matrices with more than 15 % non-zeros (or too
// large to inspect cheaply) stay on the dense kernels
static int sparse_plan(smc_model_s* m, uint32_t reg, int tA, long long r, long long k, long long ar, SmcSparse** out) {
  for (auto& p : m->sparse)
    if (p.reg == (int)reg && p.tA == tA) { *out = &p; return SMC_OK; }
  SmcSparse sp;
  sp.reg = (int)reg; sp.tA = tA; sp.r = r; sp.k = k;
  const long long numel = r * k;
  if (numel >= 64 && numel <= (4LL << 20)) {
    std::vector<double> A(numel);
    SMC_CUDA(smc_copy_sync(A.data(), m->regs[reg].dev, (size_t)numel * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    auto at = [&](long long i, long long l) { return A[tA ? (l + i * ar) : (i + l * ar)]; };
    long long nnz = 0;
    for (double v : A) nnz += (v != 0.0);
    bool finite = true;
    for (double v : A) finite = finite && std::isfinite(v);
    if (finite && nnz * 100 <= numel * 15) {
      std::vector<int32_t> ptr(r + 1, 0), col;
      std::vector<double> val;
      col.reserve(nnz); val.reserve(nnz);
      for (long long i = 0; i < r; i++) {
        for (long long l = 0; l < k; l++) {
          double v = at(i, l);
          if (v != 0.0) { col.push_back((int32_t)l); val.push_back(v); }
        }
        ptr[i + 1] = (int32_t)col.size();
      }
      SMC_CUDA(smc_dev_alloc(&sp.ptr, ptr.size() * 4, m->ctx->stream));
      SMC_CUDA(smc_dev_alloc(&sp.col, std::max<size_t>(col.size(), 1) * 4, m->ctx->stream));
      SMC_CUDA(smc_dev_alloc(&sp.val, std::max<size_t>(val.size(), 1) * 8, m->ctx->stream));
      SMC_CUDA(smc_copy_sync(sp.ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice, m->ctx->stream));
      SMC_CUDA(smc_copy_sync(sp.col, col.data(), col.size() * 4, cudaMemcpyHostToDevice, m->ctx->stream));
      SMC_CUDA(smc_copy_sync(sp.val, val.data(), val.size() * 8, cudaMemcpyHostToDevice, m->ctx->stream));
      sp.sparse = true;
    }
  }
  m->sparse.push_back(sp);
  *out = &m->sparse.back();
  return SMC_OK;
}

// `s2` (optional): the next statement of the evaluation when it is a MATMUL with the same left operand, flags and shapes and
// independent of `s` (eval_body checks): where the kernel chosen for `s` can take a second (right operand, destination)
// pair, both products go out in one launch and *paired is set.
static int run_statement(smc_model_s* m, const smc_tape_stmt& s, int64_t C, bool header, const smc_tape_stmt* s2 = nullptr,
                         bool* paired = nullptr) {
  cudaStream_t st = m->ctx->stream;
  const SmcReg& dreg = m->regs[s.dst];
  Dst d = smc_make_dst(m, s.dst);
  int acc = (s.flags & SMC_F_ACC) ? 1 : 0;
  int* bad = header ? (m->bad + m->Cpad) : m->bad;  // slot Cpad = header flag
  switch (s.op) {
    case SMC_OP_MATMUL: {
      MmArgs a;
      a.A = smc_make_opd(m, s.src[0]);
      a.B = smc_make_opd(m, s.src[1]);
      a.tA = (s.flags & SMC_F_TRANS_A) ? 1 : 0;
      a.tB = (s.flags & SMC_F_TRANS_B) ? 1 : 0;
      a.r = s.aux0; a.k = s.aux1; a.n = (long long)s.imm; a.C = C;
      a.ar = a.tA ? a.k : a.r;
      a.br = a.tB ? a.n : a.k;
      long long outn = a.r * a.n;
      const double* B2 = nullptr;      // the partner's right operand and destination, when they fit this statement's layout
      double* out2 = nullptr;
      if (s2 && paired) {
        const Opd b2 = smc_make_opd(m, s2->src[1]);
        const Dst d2 = smc_make_dst(m, s2->dst);
        if (!b2.is_imm && b2.sc == 1 && b2.se == a.B.se && d2.sc == 1 && d2.se == d.se && b2.p && d2.p) { B2 = b2.p; out2 = d2.p; }
      }
      if (!a.A.is_imm && a.A.sc == 0 && !a.B.is_imm && a.B.sc == 1 && !a.tB && d.sc == 1 && !header &&
          m->regs[s.src[0]].numel == a.r * a.k) {
        SmcSparse* sp = nullptr;
        SMC_TRY(sparse_plan(m, s.src[0], a.tA, a.r, a.k, a.ar, &sp));
        if (sp->sparse) {
          k_spmm_dc<<<grid_for((B2 ? 2 : 1) * outn * C), 256, 0, st>>>(sp->ptr, sp->col, sp->val, a.B.p, a.B.se, d.p, d.se, a.r, a.n, a.k,
                                                                      C, acc, B2, out2);
          m->ctx->launches++;
          if (B2) *paired = true;
          break;
        }
      }
      if (!a.A.is_imm && a.A.sc == 0 && !a.B.is_imm && a.B.sc == 1 && !a.tB && C >= 32 && d.sc == 1) {
        // data matrix x per-chain matrix: tiled kernel, thread = chain
        MmDcArgs g;
        g.A = a.A.p; g.ar = a.ar; g.tA = a.tA; g.B = a.B.p; g.out = d.p; g.Bse = a.B.se; g.Ose = d.se;
        g.r = a.r; g.k = a.k; g.n = a.n; g.C = C; g.acc = acc;
        g.B2 = nullptr; g.out2 = nullptr;
        int RI = 16;
        {
          long long best = -1;
          for (int cand : {4, 8, 16}) {
            long long cost = ((a.r + cand - 1) / cand) * (cand + 4);
            if (best < 0 || cost <= best) { best = cost; RI = cand; }
          }
        }
        int tcx = 32;
        while (tcx < 256 && tcx < C) tcx <<= 1;
        int ty = 256 / tcx;
        long long groups = (a.r + RI - 1) / RI;
        while (ty > 1 && ty / 2 >= groups) ty >>= 1, tcx = std::min(256, tcx * 2);
        long long gtiles = (groups + ty - 1) / ty;
        long long ctiles = (C + tcx - 1) / tcx;
        // split the contraction when the output alone gives each SM fewer than ~32 warps of work (e.g. transpose(X) * ds
        // with few output rows: every thread would walk the whole k loop alone, latency bound)
        long long S = 1;
        const long long warps = C * groups * a.n / 32, want = 32LL * m->ctx->sm_count;
        if (warps < want && a.k >= 64) S = std::min<long long>((want + warps - 1) / std::max<long long>(warps, 1), a.k / 32);
        g.chunk = ((a.k + S - 1) / S + 31) / 32 * 32;
        S = (a.k + g.chunk - 1) / g.chunk;
        g.S = (int)S;
        if (S > 1) SMC_TRY(smc_scratch_reserve(m, S * outn * C));
        g.partial = m->scratch;
        if (S == 1 && B2) { g.B2 = B2; g.out2 = out2; *paired = true; }   // both products in one launch (grid.y twice as tall)
        dim3 grid((unsigned)ctiles, (unsigned)(gtiles * a.n * (g.B2 ? 2 : 1)), (unsigned)S), block(tcx, ty);
        size_t smem = (size_t)ty * RI * 32 * sizeof(double);
        if (RI == 4) k_matmul_dc<4><<<grid, block, smem, st>>>(g);
        else if (RI == 8) k_matmul_dc<8><<<grid, block, smem, st>>>(g);
        else k_matmul_dc<16><<<grid, block, smem, st>>>(g);
        m->ctx->launches++;
        if (S > 1) {
          k_reduce_final<<<grid_for(outn * C), 256, 0, st>>>(m->scratch, S, outn, C, d, acc);
          m->ctx->launches++;
        }
        break;
      }
      long long S = 1;
      long long par = outn * C;
      if (par < 148LL * 512 && a.k >= 256) S = std::min<long long>((148LL * 512 + par - 1) / par, (a.k + 63) / 64);
      long long chunk = (a.k + S - 1) / S;
      // (a destination that is also an operand keeps the two-step form: every product is read before anything is written)
      const bool views_may_overlap = dreg.d.kind == SMC_REG_GRAD &&
                                     (m->regs[s.src[0]].d.kind == SMC_REG_GRAD || m->regs[s.src[1]].d.kind == SMC_REG_GRAD);
      const bool direct = S == 1 && s.dst != s.src[0] && s.dst != s.src[1] && d.p != a.A.p && d.p != a.B.p && !views_may_overlap;
      if (!direct) SMC_TRY(smc_scratch_reserve(m, S * outn * C));
      dim3 grid(grid_for(par), (unsigned)S);
      k_matmul<<<grid, 256, 0, st>>>(a, m->scratch, chunk, d, acc, direct ? 1 : 0);
      m->ctx->launches++;
      if (!direct) {
        k_reduce_final<<<grid_for(par), 256, 0, st>>>(m->scratch, S, outn, C, d, acc);
        m->ctx->launches++;
      }
      break;
    }
    case SMC_OP_TRANSPOSE: {
      Opd src = smc_make_opd(m, s.src[0]);
      long long rows = s.aux0, cols = s.aux1;
      k_transpose<<<grid_for(rows * cols * C), 256, 0, st>>>(src, d, rows, cols, C, acc);
      m->ctx->launches++;
      break;
    }
    case SMC_OP_GATHER: {
      Opd src = smc_make_opd(m, s.src[0]);
      long long n = m->idx_len[s.aux0];
      k_gather<<<grid_for(n * C), 256, 0, st>>>(src, d, m->idx_dev[s.aux0], n, C, acc);
      m->ctx->launches++;
      break;
    }
    case SMC_OP_SCATTER_SET:
    case SMC_OP_SCATTER_ADD: {
      Opd src = smc_make_opd(m, s.src[0]);
      SmcScatterPlan* sp = nullptr;
      SMC_TRY(scatter_plan(m, s.aux0, dreg.numel, &sp));
      k_scatter_csr<<<grid_for(dreg.numel * C), 256, 0, st>>>(src, d, sp->ptr, sp->list, dreg.numel, C, s.op == SMC_OP_SCATTER_ADD);
      m->ctx->launches++;
      break;
    }
    case SMC_OP_FUSED:
      return smc_vm_launch(m, m->programs[s.aux0], C, bad);
    case SMC_OP_GLM: {
      for (auto& g : m->glm)
        if (&m->stmts[g.stmt] == &s) return smc_glm_launch(m, g, C);
      smc_set_error("internal: GLM statement without a plan");
      return SMC_ERR_INVALID;
    }
    case SMC_OP_AFFNORM: {
      for (auto& p : m->aff)
        if (&m->stmts[p.stmt] == &s) return smc_aff_launch(m, p, C);
      smc_set_error("internal: AFFNORM statement without a plan");
      return SMC_ERR_INVALID;
    }
    default: {
      EwArgs a;
      a.op = s.op; a.dist = s.aux0; a.darg = s.aux1; a.nsrc = s.nsrc; a.acc = acc;
      long long nsrc_el = 1;
      for (uint32_t k = 0; k < s.nsrc; k++) {
        a.s[k] = smc_make_opd(m, s.src[k]);
        nsrc_el = std::max<long long>(nsrc_el, m->regs[s.src[k]].numel);
      }
      for (uint32_t k = s.nsrc; k < 4; k++) a.s[k] = a.s[0];
      if (s.op == SMC_OP_LOGPDF) nsrc_el = s.aux1;  // length of the first array argument (src/distribs.jl:99)
      long long n = nsrc_el;
      a.d = d; a.C = C; a.bad = bad; a.bad_stride = header ? 0 : 1;
      if (dreg.numel == nsrc_el || nsrc_el == 1) {
        // elementwise; scalar operands (and a scalar result of the operands) broadcast over the destination
        n = dreg.numel;
        a.n = n;
        k_elementwise<<<grid_for(n * C), 256, 0, st>>>(a);
        m->ctx->launches++;
      } else if (dreg.numel == 1) {
        a.n = n;
        long long S;
        if (C >= 4096 || n < 512) {
          S = 1;
          if (C < 148LL * 256) S = std::max<long long>(1, std::min<long long>((148LL * 256) / C, n / 32));
          long long chunk = (n + S - 1) / S;
          S = (n + chunk - 1) / chunk;
          SMC_TRY(smc_scratch_reserve(m, S * C));
          dim3 grid((unsigned)((C + 255) / 256), (unsigned)S);
          k_reduce_chains<<<grid, 256, 0, st>>>(a, m->scratch, chunk);
        } else {
          S = std::max<long long>(1, std::min<long long>((n + 2047) / 2048, std::max<long long>(1, (148LL * 16) / C)));
          long long chunk = (n + S - 1) / S;
          S = (n + chunk - 1) / chunk;
          SMC_TRY(smc_scratch_reserve(m, S * C));
          dim3 grid((unsigned)S, (unsigned)C);
          k_reduce_elems<<<grid, 256, 0, st>>>(a, m->scratch, chunk);
        }
        k_reduce_final<<<grid_for(C), 256, 0, st>>>(m->scratch, S, 1, C, d, acc);
        m->ctx->launches += 2;
      } else {
        smc_set_error("tape: statement with op %u has incompatible shapes (dst %lld, operands %lld)", s.op,
                      (long long)dreg.numel, nsrc_el);
        return SMC_ERR_INVALID;
      }
    }
  }
  return SMC_OK;
}

// The result register (the model's accumulator, src/SimpleMCMC.jl:14) is zero whenever no evaluation is under way: at
// finalize, after every k_finish_eval, after an evaluation that failed half way.  Tapes rely on it since version 4: every
// statement that contributes to the accumulator may be a `+=` (and so may move between the forward and the gradient half).
static void smc_clear_result(smc_model_s* m) {
  if (m->hdr.result_reg >= m->regs.size()) return;
  const SmcReg& rr = m->regs[m->hdr.result_reg];
  if (rr.d.kind == SMC_REG_CHAIN && rr.dev) cudaMemsetAsync(rr.dev, 0, (size_t)rr.numel * m->Cpad * sizeof(double), m->ctx->stream);
}

extern "C" int smc_model_finalize(smc_model_t m, int64_t max_chains, int row_shard) {
  if (!m || max_chains < 1) { smc_set_error("smc_model_finalize: bad argument"); return SMC_ERR_INVALID; }
  cudaSetDevice(m->ctx->device);
  for (auto& d : m->data)
    if (!d.bound) { smc_set_error("external variable '%s' is used by the model but was never bound", d.name.c_str()); return SMC_ERR_UNBOUND; }
  cudaStream_t st = m->ctx->stream;
  SMC_CUDA(cudaStreamSynchronize(st));
  smc_graphs_invalidate(m);
  for (auto& p : m->sparse) { if (p.ptr) smc_dev_free(p.ptr, m->ctx->stream); if (p.col) smc_dev_free(p.col, m->ctx->stream); if (p.val) smc_dev_free(p.val, m->ctx->stream); }
  m->sparse.clear();   // (data may have been bound again)
  {
    const char* ng = getenv("SMC_NO_GRAPHS");
    m->graphs_enabled = (ng && ng[0] == '1') ? 0 : 1;
    const char* gn = getenv("SMC_GRAPH_NCCL");   // opt-in: capture the in-sequence NCCL all-reduce of row-sharded models too
    m->graph_nccl = (gn && gn[0] == '1') ? 1 : 0;
  }
  // chains are the contiguous axis of every per-chain array; pad to 32 only when that does not waste bandwidth
  int64_t Cpad = max_chains >= 32 ? (max_chains + 31) / 32 * 32 : max_chains;
  bool realloc_chain = (Cpad != m->Cpad);
  m->Cmax = max_chains;
  m->row_shard = row_shard;
  if (realloc_chain) {
    for (auto& r : m->regs)
      if (r.d.kind == SMC_REG_CHAIN && r.dev) { smc_dev_free(r.dev, m->ctx->stream); r.dev = nullptr; }
    if (m->beta) smc_dev_free(m->beta, m->ctx->stream);
    if (m->grad) smc_dev_free(m->grad, m->ctx->stream);
    if (m->logp) smc_dev_free(m->logp, m->ctx->stream);
    if (m->bad) smc_dev_free(m->bad, m->ctx->stream);
    m->beta = m->grad = m->logp = nullptr;
    m->bad = nullptr;
    m->Cpad = Cpad;
    size_t free_b = 0, total_b = 0;
    SMC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    double need = 0;
    for (auto& r : m->regs)
      if (r.d.kind == SMC_REG_CHAIN) need += (double)r.numel * Cpad * 8.0;
    if (need > 0.85 * (double)free_b) {
      smc_set_error("the generic executor needs %.1f GB of per-chain registers for %lld chains (%.1f GB free): "
                    "reduce `chains` or use a model the fused kernels cover", need / 1e9, (long long)max_chains, free_b / 1e9);
      return SMC_ERR_MEMORY;
    }
    for (auto& r : m->regs)
      if (r.d.kind == SMC_REG_CHAIN) {
        SMC_CUDA(smc_dev_alloc(&r.dev, (size_t)r.numel * Cpad * sizeof(double), m->ctx->stream));
        r.owned = true;
      }
    SMC_CUDA(smc_dev_alloc(&m->beta, (size_t)std::max<int64_t>(m->M, 1) * Cpad * sizeof(double), m->ctx->stream));
    SMC_CUDA(smc_dev_alloc(&m->grad, (size_t)std::max<int64_t>(m->M, 1) * Cpad * sizeof(double), m->ctx->stream));
    SMC_CUDA(smc_dev_alloc(&m->logp, (size_t)Cpad * sizeof(double), m->ctx->stream));
    SMC_CUDA(smc_dev_alloc(&m->bad, (size_t)(Cpad + 32) * sizeof(int), m->ctx->stream));
    SMC_CUDA(cudaMemsetAsync(m->beta, 0, (size_t)std::max<int64_t>(m->M, 1) * Cpad * sizeof(double), st));
    SMC_CUDA(cudaMemsetAsync(m->grad, 0, (size_t)std::max<int64_t>(m->M, 1) * Cpad * sizeof(double), st));
  }
  for (auto& r : m->regs)
    if (r.d.kind == SMC_REG_HDR && !r.dev) {
      SMC_CUDA(smc_dev_alloc(&r.dev, (size_t)r.numel * sizeof(double), m->ctx->stream));
      r.owned = true;
    }
  smc_clear_result(m);
  if (!m->ev0) { SMC_CUDA(cudaEventCreate(&m->ev0)); SMC_CUDA(cudaEventCreate(&m->ev1)); }
  for (auto& pr : m->programs) SMC_TRY(smc_vm_prepare(m, pr));  // operand tables of the register programs (device pointers)
  // the let-header: parameter independent statements, evaluated once (src/parsing.jl:432-442)
  SMC_CUDA(cudaMemsetAsync(m->bad, 0, (size_t)(Cpad + 32) * sizeof(int), st));
  for (int si : m->hdr_stmts) SMC_TRY(run_statement(m, m->stmts[si], 1, true));
  int hbad = 0;
  SMC_CUDA(cudaMemcpyAsync(&hbad, m->bad + Cpad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SMC_CUDA(cudaStreamSynchronize(st));
  if (hbad) { smc_set_error("Initial values out of model support, try other values (a parameter-independent density is zero)"); return SMC_ERR_SUPPORT; }
  // Row-sharded models: every bound array is this rank's row shard of the caller's data, and only SMC_OP_GLM statements sum
  // over the rows ACROSS the ranks (glm_fused.cu / smc_p2p.cu).  Any other statement that consumes a bound array (or a header
  // value computed from one) would reduce over the local shard only -- every rank a different logp / gradient, chains that
  // drift apart without an error -- so such tapes are refused (Poisson regression, a Normal likelihood with sigma as a
  // parameter, anything lowered with fuse = False).
  if (row_shard) {
    std::vector<char> rowdep(m->regs.size(), 0);
    for (size_t i = 0; i < m->regs.size(); i++)
      if (m->regs[i].d.kind == SMC_REG_DATA && m->regs[i].numel > 1) rowdep[i] = 1;
    auto reads_shard = [&](const smc_tape_stmt& s) -> bool {
      for (uint32_t k = 0; k < s.nsrc && k < 4; k++)
        if (s.src[k] < rowdep.size() && rowdep[s.src[k]]) return true;
      if (s.op == SMC_OP_FUSED && s.aux0 < m->programs.size())
        for (const smc_prog_instr& in : m->programs[s.aux0].instr)
          if ((in.op == SMC_P_LOADV || in.op == SMC_P_LOADS) && in.reg < rowdep.size() && rowdep[in.reg]) return true;
      return false;
    };
    for (int si : m->hdr_stmts) {
      const smc_tape_stmt& s = m->stmts[si];
      if (!reads_shard(s)) continue;
      if (s.dst < rowdep.size()) rowdep[s.dst] = 1;
      if (s.op == SMC_OP_FUSED && s.aux0 < m->programs.size()) {
        for (const smc_prog_instr& in : m->programs[s.aux0].instr)
          if (in.op == SMC_P_STOREV && in.reg < rowdep.size()) rowdep[in.reg] = 1;
        for (const smc_prog_red& rd : m->programs[s.aux0].red)
          if (rd.reg < rowdep.size()) rowdep[rd.reg] = 1;
      }
    }
    for (int si : m->body_stmts) {
      const smc_tape_stmt& s = m->stmts[si];
      if (s.op == SMC_OP_GLM || !reads_shard(s)) continue;
      smc_set_error("row-sharded models: statement %d (opcode %u) reads the row-sharded data outside a fused GLM statement; its sum over "
                    "the rows would cover this rank's shard only.  Supported with shard=\"rows\": Normal (literal sigma), Bernoulli and "
                    "Binomial (literal n) likelihoods of X * beta", si, s.op);
      return SMC_ERR_INVALID;
    }
  }
  // fused GLM statements
  for (auto& g : m->glm) smc_glm_release(g, m->ctx->stream);
  m->glm.clear();
  for (int si : m->body_stmts)
    if (m->stmts[si].op == SMC_OP_GLM) {
      SmcGlmPlan g;
      g.stmt = si;
      SMC_TRY(smc_glm_prepare(m, g));
      m->glm.push_back(g);
    }
  for (auto& a : m->aff) smc_aff_release(a, m->ctx->stream);
  m->aff.clear();
  for (int si : m->body_stmts)
    if (m->stmts[si].op == SMC_OP_AFFNORM) {
      if (row_shard) { smc_set_error("row-sharded models: only GLM statements sum over the rows across ranks (AFFNORM does not)"); return SMC_ERR_INVALID; }
      SmcAffPlan a;
      a.stmt = si;
      SMC_TRY(smc_aff_prepare(m, a));
      m->aff.push_back(a);
    }
  SMC_CUDA(cudaStreamSynchronize(st));
  SMC_CUDA(cudaGetLastError());
  m->finalized = true;
  return SMC_OK;
}

// the launch sequence of one evaluation: every body statement of the tape, all chains in lockstep
static int eval_body(smc_model_s* m, int64_t C, bool with_grad) {
  cudaStream_t st = m->ctx->stream;
  // (the "give up eval" flags m->bad are zero here: cleared at finalize and again by every k_finish_eval)
  const int n_fwd = (int)m->hdr.reserved[0];  // statements past this index only serve the gradient
  static thread_local std::vector<int> run;    // the statements of this evaluation, in order
  run.clear();
  for (int si : m->body_stmts) {
    const smc_tape_stmt& s = m->stmts[si];
    if (!with_grad && (si >= n_fwd || m->regs[s.dst].d.kind == SMC_REG_GRAD)) continue;
    if (with_grad && (s.flags & SMC_F_FWD_ONLY)) continue;   // its work is part of a statement of the gradient half
    run.push_back(si);
  }
  static int pair_mm = -1;                     // SMC_NO_MATMUL_PAIRS=1: one launch per product (A/B switch)
  if (pair_mm < 0) { const char* e = getenv("SMC_NO_MATMUL_PAIRS"); pair_mm = (e && e[0] == '1') ? 0 : 1; }
  for (size_t q = 0; q < run.size(); q++) {
    const smc_tape_stmt& s = m->stmts[run[q]];
    // two consecutive products with one left operand (`submatrix * rand_int`, `submatrix * rand_slope` of
    // examples/sleepstudy.jl and their adjoints; `oneD * mu`, `oneD * sigma` of the hierarchical example): one launch
    const smc_tape_stmt* s2 = nullptr;
    if (pair_mm && s.op == SMC_OP_MATMUL && q + 1 < run.size()) {
      const smc_tape_stmt& t = m->stmts[run[q + 1]];
      if (t.op == SMC_OP_MATMUL && t.src[0] == s.src[0] && t.flags == s.flags && t.aux0 == s.aux0 && t.aux1 == s.aux1 && t.imm == s.imm &&
          t.dst != s.dst && t.src[1] != s.dst && s.src[1] != t.dst && t.src[0] != s.dst && s.src[0] != t.dst &&
          m->regs[t.dst].d.kind == m->regs[s.dst].d.kind && m->regs[t.src[1]].d.kind != SMC_REG_GRAD && m->regs[s.src[1]].d.kind != SMC_REG_GRAD)
        s2 = &t;
    }
    bool paired = false;
    int rc = run_statement(m, s, C, false, s2, &paired);
    if (rc != SMC_OK) {
      cudaMemsetAsync(m->bad, 0, (size_t)m->Cpad * sizeof(int), st);
      smc_clear_result(m);
      return rc;
    }
    if (paired) q++;
  }
  const SmcReg& rr = m->regs[m->hdr.result_reg];
  Opd ro = smc_make_opd(m, m->hdr.result_reg);
  if (ro.is_imm || rr.numel != 1) { smc_set_error("tape: result register must be a computed scalar"); return SMC_ERR_INVALID; }
  // row-sharded models all-reduce inside the GLM statement (its (L, G) are the only row sums), so every
  // rank finishes with identical (logp, grad) here
  k_finish_eval<<<grid_for(C), 256, 0, st>>>(const_cast<double*>(ro.p), ro.se, ro.sc, m->bad, m->logp, m->grad, m->M, C, m->Cpad,
                                             with_grad ? 1 : 0);
  m->ctx->launches++;
  return SMC_OK;
}

// may the launch sequence of an evaluation be captured?  An in-sequence NCCL all-reduce (row-sharded models) keeps it eager
// unless the host opted in (SMC_GRAPH_NCCL=1); the peer-memory exchange is an ordinary kernel with its epoch in device memory
bool smc_comm_graph_ok(const smc_model_s* m) {
  if (!m->row_shard || m->ctx->nranks <= 1) return true;
  return m->graph_nccl || m->ctx->p2p_ranks > 1;
}

// beta[M][Cpad] (device) -> logp[Cpad], grad[M][Cpad] (device).
// The statement sequence of a (chain count, gradient) pair is fixed, so its second evaluation is captured into a
// CUDA graph and every later one is a single graph launch: the sampler loops (one evaluation per leapfrog) stop
// paying one host launch per tape statement.  Row-sharded models (NCCL inside the sequence) and timed evaluations
// stay eager.
int smc_model_eval_device(smc_model_s* m, int64_t C, bool with_grad) {
  if (!m->finalized) { smc_set_error("model is not finalized (call smc_model_finalize after binding data)"); return SMC_ERR_INVALID; }
  if (C < 1 || C > m->Cmax) { smc_set_error("chains %lld outside [1, %lld] given to smc_model_finalize", (long long)C, (long long)m->Cmax); return SMC_ERR_INVALID; }
  cudaStream_t st = m->ctx->stream;
  m->C = C;
  if (!m->graphs_enabled || m->time_glm || !smc_comm_graph_ok(m)) return eval_body(m, C, with_grad);
  if (m->graphs_stale) smc_graphs_invalidate(m);
  int gi = -1;
  for (size_t i = 0; i < m->graphs.size(); i++)
    if (m->graphs[i].C == C && m->graphs[i].with_grad == (with_grad ? 1 : 0)) { gi = (int)i; break; }
  if (gi < 0) {  // first evaluation of this shape: eager, sizes every scratch buffer
    SmcEvalGraph g;
    g.C = C; g.with_grad = with_grad ? 1 : 0; g.state = 1;
    SMC_TRY(eval_body(m, C, with_grad));
    if (!m->graphs_stale) m->graphs.push_back(g);
    return SMC_OK;
  }
  if (m->graphs[gi].state == 1) {
    const int64_t epoch = m->realloc_epoch, l0 = m->ctx->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      m->graphs[gi].state = 3;
      return eval_body(m, C, with_grad);
    }
    int rc = eval_body(m, C, with_grad);
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    bool ok = (rc == SMC_OK && e == cudaSuccess && graph && epoch == m->realloc_epoch);
    cudaGraphExec_t exec = nullptr;
    if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      m->ctx->launches = l0;
      if (m->graphs_stale) smc_graphs_invalidate(m);
      else m->graphs[gi].state = 3;
      return eval_body(m, C, with_grad);
    }
    m->graphs[gi].exec = exec;
    m->graphs[gi].launches = m->ctx->launches - l0;
    m->graphs[gi].state = 2;
    SMC_CUDA(cudaGraphLaunch(exec, st));
    return SMC_OK;
  }
  if (m->graphs[gi].state == 2) {
    SMC_CUDA(cudaGraphLaunch(m->graphs[gi].exec, st));
    m->ctx->launches += m->graphs[gi].launches;
    return SMC_OK;
  }
  return eval_body(m, C, with_grad);
}

int smc_model_eval_inline(smc_model_s* m, int64_t C, bool with_grad) {
  if (!m->finalized || C < 1 || C > m->Cmax) { smc_set_error("smc_model_eval_inline: model not finalized or chains out of range"); return SMC_ERR_INVALID; }
  m->C = C;
  return eval_body(m, C, with_grad);
}

extern "C" int smc_eval(smc_model_t m, const double* beta, int64_t C, double* logp, double* grad) {
  if (!m || !beta || !logp) { smc_set_error("smc_eval: NULL argument"); return SMC_ERR_INVALID; }
  if (!m->finalized) { smc_set_error("model is not finalized (call smc_model_finalize after binding data)"); return SMC_ERR_INVALID; }
  if (C < 1 || C > m->Cmax) { smc_set_error("smc_eval: chains %lld outside [1, %lld]", (long long)C, (long long)m->Cmax); return SMC_ERR_INVALID; }
  cudaSetDevice(m->ctx->device);
  cudaStream_t st = m->ctx->stream;
  int64_t M = m->M;
  SMC_TRY(smc_scratch_reserve(m, std::max<int64_t>(C * M, 1)));
  SMC_CUDA(cudaMemcpyAsync(m->scratch, beta, (size_t)C * M * sizeof(double), cudaMemcpyHostToDevice, st));
  k_rows_to_cols<<<grid_for(C * M), 256, 0, st>>>(m->scratch, m->beta, C, M, m->Cpad, 0);
  m->ctx->launches++;
  SMC_TRY(smc_model_eval_device(m, C, grad != nullptr));
  SMC_CUDA(cudaMemcpyAsync(logp, m->logp, (size_t)C * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (grad) {
    SMC_TRY(smc_scratch_reserve(m, std::max<int64_t>(C * M, 1)));
    k_cols_to_rows<<<grid_for(C * M), 256, 0, st>>>(m->grad, m->scratch, C, M, m->Cpad);
    m->ctx->launches++;
    SMC_CUDA(cudaMemcpyAsync(grad, m->scratch, (size_t)C * M * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  SMC_CUDA(cudaStreamSynchronize(st));
  SMC_CUDA(cudaGetLastError());
  if (m->row_shard) SMC_TRY(smc_p2p_check(m->ctx));
  return SMC_OK;
}

extern "C" int smc_eval_timed(smc_model_t m, int reps, int flush_l2, double* ms_eval, double* ms_glm_kernel) {
  if (!m || reps < 1 || !m->finalized || m->C < 1) { smc_set_error("smc_eval_timed: call smc_eval first"); return SMC_ERR_INVALID; }
  cudaSetDevice(m->ctx->device);
  cudaStream_t st = m->ctx->stream;
  const size_t flush_bytes = 256u << 20;
  if (flush_l2 && !m->ctx->flush_buf) SMC_CUDA(smc_dev_alloc(&m->ctx->flush_buf, flush_bytes, m->ctx->stream));
  double tot = 0.0, glm = 0.0;
  for (int r = 0; r < reps; r++) {
    if (flush_l2) SMC_CUDA(cudaMemsetAsync(m->ctx->flush_buf, r & 0xff, flush_bytes, st));
    m->time_glm = 1;
    for (auto& g : m->glm) g.last_ms = 0.f;
    for (auto& a : m->aff) a.last_ms = 0.f;
    cudaEvent_t e0, e1;
    SMC_CUDA(cudaEventCreate(&e0));
    SMC_CUDA(cudaEventCreate(&e1));
    SMC_CUDA(cudaEventRecord(e0, st));
    int rc = smc_model_eval_device(m, m->C, true);
    SMC_CUDA(cudaEventRecord(e1, st));
    SMC_CUDA(cudaEventSynchronize(e1));
    m->time_glm = 0;
    if (rc != SMC_OK) return rc;
    float ms = 0.f;
    SMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    tot += ms;
    for (auto& g : m->glm) glm += g.last_ms;
    for (auto& a : m->aff) glm += a.last_ms;      // (the fused statement kernels: GLM and AFFNORM)
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (ms_eval) *ms_eval = tot / reps;
  if (ms_glm_kernel) *ms_glm_kernel = glm / reps;
  return SMC_OK;
}
