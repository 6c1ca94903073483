// Internal structures of libsimplemcmc_cuda.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/simplemcmc.h"
#include "../../include/simplemcmc_tape.h"

void smc_set_error(const char* fmt, ...);

#define SMC_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      smc_set_error("CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      return SMC_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define SMC_TRY(call)           \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != SMC_OK) return rc_; \
  } while (0)

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) on the context's stream, with a release
// threshold so that the blocks one sampler call frees are what the next one gets: a simpleHMC call makes ~60
// allocations, and as plain cudaMalloc/cudaFree those took 0.1-1.1 s of host time on a busy box (tools/e2e_breakdown.py).
// SMC_NO_POOL=1 restores cudaMalloc/cudaFree.  Inside a stream capture both fall back to the synchronous calls (the
// free is deferred), so that no allocation ever belongs to a graph.
cudaError_t smc_dev_alloc_bytes(void** p, size_t bytes, cudaStream_t st);
void smc_dev_free(void* p, cudaStream_t st);
template <typename T>
static inline cudaError_t smc_dev_alloc(T** p, size_t bytes, cudaStream_t st) { return smc_dev_alloc_bytes((void**)p, bytes, st); }
// blocking copy ordered on `st` (the pool's allocations are ordered on that stream, not on the legacy default stream)
cudaError_t smc_copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st);

// "once per device" latch for cudaFuncSetAttribute (function attributes belong to a device's context, and one process may
// hold contexts on several devices)
struct SmcOncePerDevice {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

struct smc_ctx_s {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  void* nccl_comm = nullptr;  // ncclComm_t
  int rank = 0, nranks = 1;
  void* flush_buf = nullptr;  // L2 flush scratch (256 MB), allocated on demand
  int64_t launches = 0;       // kernels launched by this library on this context
  // peer-memory exchange of row-sharded models (smc_p2p.cu; opt-in through smc_ctx_p2p_export / _import)
  void* p2p_local = nullptr;
  void* p2p_peer[16] = {};
  unsigned* p2p_count = nullptr;   // device: push tickets, finish tickets, epoch, status
  int p2p_rank = 0, p2p_ranks = 0, p2p_alloc_ranks = 0;
};

struct SmcReg {
  smc_tape_reg d;
  int64_t numel = 1;
  double* dev = nullptr;  // DATA/HDR: [numel]; CHAIN: [numel][Cpad]
  bool owned = false;
};

struct SmcData {
  std::string name;
  int64_t rows = 0, cols = 0;
  int reg = -1;
  bool bound = false;
};

struct SmcGlmPlan {  // one SMC_OP_GLM statement, prepared at finalize
  int stmt = -1;
  int family = 0;
  int sign_neg = 0;
  double par = 0.0;
  int64_t N = 0;       // local rows
  int M = 0, MP = 0;   // columns and columns padded to a multiple of 8
  double* Xrm = nullptr;  // [N][MP] row major, zero padded (DMMA kernel; built lazily)
  const double* Xcm = nullptr;  // [M][N] column major as bound (stream kernel)
  const double* y = nullptr;
  double* ybuf = nullptr;  // response as the kernel consumes it (sign-normalised, broadcast)
  int64_t y_len = 0;
  double y_scalar = 0.0;
  double* part = nullptr;  // partial (L,G) buffers [RS][MP+1][Cpad]
  int64_t part_cap = 0;
  float last_ms = 0.f;
  double lc_sum = 0.0;     // BINOMIAL: sum_i lchoose(n, y_i) over the local rows
  int gram = 0;            // SMC_GLM_F_GRAM: evaluate from Z'Z with Z = [X, ytilde]
  double* ZtZ = nullptr;   // [(M+1)][(M+1)] row major
  // int8 tensor-core path (glm_i8.cu; SMC_GLM_I8=0 switches it off): X as seven balanced 8-bit slices in 128-row tiles
  int8_t* QX8 = nullptr;
  double sx8 = 0.0;        // power-of-two scale of X
  int* i8_status = nullptr;
  int i8_checked = 0;
  int i8_declined = 0;     // X holds NaN / Inf: the statement stays on the FP64 kernel
};

struct SmcAffPlan {  // one SMC_OP_AFFNORM statement (affnorm.cu)
  int stmt = -1;
  std::vector<int32_t> desc;        // host copy of the descriptor table
  unsigned int* counter = nullptr;  // ticket counters of the last-block finish, one per chain group
  float last_ms = 0.f;              // event-timed launch (smc_eval_timed)
};

// Micro-operation of the register-program VM (vm_exec.cu): the tape's LOADV/LOADS/STOREV/REDUCE instructions are
// folded into operand modes / side effects of the computing instruction at plan time.
enum { VM_NONE = 0, VM_LOCAL = 1, VM_SCAL = 2, VM_ARRAY = 3 };
enum { VMF_ACC = 1, VMF_WRITE = 2, VMF_STORE = 4, VMF_REDUCE = 8 };
// A chain micro-op (planner only, never on a tape): r = operand 0, then up to three stages r = stage_k(r, operand k) whose
// intermediate values stay in registers instead of going through a local slot.  dist / darg / code3 hold the stage codes;
// operand 3 is a source (chains never accumulate).  A Normal density stage ends the chain and takes operands k (the
// other one of mean / variate) and k + 1 (sigma, a per-chain scalar).
enum { VM_OP_CHAIN = 254 };
enum { VM_CH_END = 0, VM_CH_ADD, VM_CH_SUB, VM_CH_RSUB, VM_CH_MUL, VM_CH_DIV, VM_CH_RDIV, VM_CH_NEG,
       VM_CH_NLOG,     // Normal log-density of (r - o) / sigma
       VM_CH_ND_RMO,   // (r - o) / sigma^2
       VM_CH_ND_OMR,   // (o - r) / sigma^2
       VM_CH_ND_SIG }; // ((r - o)^2 / sigma^2 - 1) / sigma
struct VmUop {       // 32 bytes
  uint8_t op, dist, darg, flags;
  uint8_t mode[4];   // operands 0..2 and [3] = the accumulate operand (VMF_ACC) / fourth source of a chain
  uint16_t idx[4];   // local slot / scalar slot / array descriptor
  uint16_t dst, store, red, nsrc;
  uint32_t code3;    // third stage of a chain
  uint32_t pad;
};
struct VmDesc {      // 32 bytes; p == NULL: immediate
  const double* p;
  long long se, sc;
  double imm;
};

struct SmcProgram {  // one SMC_OP_FUSED register program
  smc_prog_header h;
  std::vector<smc_prog_instr> instr;
  std::vector<smc_prog_red> red;
  // plan (pointer independent, built at compile time)
  std::vector<VmUop> uops;
  std::vector<uint32_t> desc_regs;  // tape register of every array descriptor (loads and stores)
  std::vector<uint32_t> scal_regs;  // tape register of every per-chain scalar input (staged once per block)
  int n_loc = 0;
  int n_ot = 0;   // operand-table entries (device form)
  // device tables (built at finalize, vm_exec.cu): micro-ops | operand table | scalar input descriptors
  void* dev_tables = nullptr;
  size_t dev_tables_bytes = 0;
};

struct SmcScatterPlan {  // inverse of a static index table for SCATTER statements
  int table = -1;
  int64_t n_dst = 0;
  int32_t* ptr = nullptr;   // [n_dst + 1]
  int32_t* list = nullptr;  // [n]
};

struct SmcSparse {  // a sparse parameter-independent matrix operand of MATMUL (e.g. a one-hot design matrix), as CSR of op(A)
  int reg = -1, tA = 0;
  bool sparse = false;       // false: looked at, dense enough to stay on the dense kernel
  int64_t r = 0, k = 0;      // op(A) is r x k
  int32_t* ptr = nullptr;    // [r + 1]
  int32_t* col = nullptr;    // [nnz]
  double* val = nullptr;     // [nnz]
};

struct SmcEvalGraph {  // one captured evaluation (all statements of the tape for a chain count), replayed with one launch
  int64_t C = 0;
  int with_grad = 0;
  int state = 0;  // 1: ran once eagerly (buffers sized), 2: captured, 3: capture failed -> stays eager
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;  // kernel nodes in the graph
};

struct smc_model_s {
  smc_ctx_s* ctx = nullptr;
  smc_tape_header hdr;
  std::vector<SmcReg> regs;
  std::vector<smc_tape_stmt> stmts;
  std::vector<int> hdr_stmts, body_stmts;
  std::vector<SmcData> data;
  std::vector<int32_t*> idx_dev;
  std::vector<uint32_t> idx_len;
  std::vector<std::vector<int32_t>> idx_host;   // host copies of the short tables (descriptors of AFFNORM statements)
  std::vector<SmcAffPlan> aff;
  std::vector<SmcGlmPlan> glm;
  std::vector<SmcProgram> programs;
  std::vector<SmcScatterPlan> scatter;
  std::vector<SmcSparse> sparse;
  bool finalized = false;
  int row_shard = 0;
  int64_t M = 0;       // n_params
  int64_t Cmax = 0, Cpad = 0;
  int64_t C = 0;       // chains of the current evaluation
  double* beta = nullptr;  // [M][Cpad]  evaluation input (PARAM views)
  double* grad = nullptr;  // [M][Cpad]  evaluation output (GRAD views)
  double* logp = nullptr;  // [Cpad]
  int* bad = nullptr;      // [Cpad] "give up eval" flags (src/distribs.jl:15,17,32,70)
  double* scratch = nullptr;  // partial sums of reductions
  int64_t scratch_cap = 0;
  double* host_stage = nullptr;  // pinned staging
  int64_t host_stage_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int time_glm = 0;
  std::vector<SmcEvalGraph> graphs;
  int graphs_enabled = 1;
  int graph_nccl = 0;
  bool graphs_stale = false;   // a buffer captured graphs point at was reallocated
  int64_t realloc_epoch = 0;
};

// smc_core.cu
int smc_model_eval_device(smc_model_s* m, int64_t C, bool with_grad);  // beta -> logp, grad (device resident)
// the same launch sequence issued kernel by kernel (never as the evaluation's own graph): for callers that capture it into a
// graph of their own where a child graph may not go (the body of a conditional node, samplers.cu)
int smc_model_eval_inline(smc_model_s* m, int64_t C, bool with_grad);
int smc_stage_reserve(smc_model_s* m, int64_t doubles);
void smc_graphs_invalidate(smc_model_s* m);
bool smc_comm_graph_ok(const smc_model_s* m);

// vm_exec.cu
int smc_vm_plan(smc_model_s* m, SmcProgram& pr);      // instruction list -> micro-ops
int smc_vm_prepare(smc_model_s* m, SmcProgram& pr);   // device tables (after registers are allocated)
int smc_vm_launch(smc_model_s* m, SmcProgram& pr, int64_t C, int* bad);

// glm_fused.cu
int smc_glm_prepare(smc_model_s* m, SmcGlmPlan& g);
int smc_glm_launch(smc_model_s* m, SmcGlmPlan& g, int64_t C);
void smc_glm_release(SmcGlmPlan& g, cudaStream_t st);

// affnorm.cu
int smc_aff_prepare(smc_model_s* m, SmcAffPlan& p);
int smc_aff_launch(smc_model_s* m, SmcAffPlan& p, int64_t C);
void smc_aff_release(SmcAffPlan& p, cudaStream_t st);

// smc_p2p.cu
bool smc_p2p_ready(const smc_ctx_s* c, int64_t n);
int smc_p2p_allreduce_device(smc_ctx_s* c, double* buf, int64_t n);
int smc_p2p_glm_exchange(smc_ctx_s* c, const double* part, int RS, int M, int MP, int64_t C, int64_t Cpad, int* bad, double* Ldst,
                         double* Gdst, int acc);
int smc_p2p_check(smc_ctx_s* c);   // after a synchronisation point: SMC_ERR_COMM when an exchange timed out
void smc_p2p_destroy(smc_ctx_s* c);

// glm_i8.cu (Normal GLM statements with 33-64 columns and >= 32 chains; SMC_GLM_I8=0 switches it off)
bool smc_glm_i8_usable(const SmcGlmPlan& g, int64_t C);
int smc_glm_i8_chain_tiles(int64_t C);
int smc_glm_i8_row_tile();
int smc_glm_i8_prepare(smc_model_s* m, SmcGlmPlan& g);   // slices X on the first call; may decline (g.i8_declined)
int smc_glm_i8_launch(smc_model_s* m, SmcGlmPlan& g, const double* beta, int64_t C, int64_t Cpad, int RS, double* part);
void smc_glm_i8_release(SmcGlmPlan& g, cudaStream_t st);
