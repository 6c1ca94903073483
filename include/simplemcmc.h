/* C ABI of libsimplemcmc_cuda.so -- the B200 (sm_100a) executor behind SimpleMCMC.jl's hot path.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   smc_model_compile / bind_data / finalize   the `eval` of the generated closure and its let-header,
 *                                              src/parsing.jl:483-492 (data = globals of Main, :484-485)
 *   smc_eval                                   one call of the closure ll_func(beta) -> (logp, grad),
 *                                              src/parsing.jl:489-492, consumed at src/samplers.jl:46,85,97
 *   smc_hmc_run                                the main loop of simpleHMC, src/samplers.jl:146-167
 *   smc_rwm_run                                the main loop of simpleRWM, src/samplers.jl:89-113
 *   smc_nuts_run                               the main loop of simpleNUTS, src/samplers.jl:199-310
 *   smc_run_output.mean/m2/first/lag1          setRes/addToRes!/calcStats!, src/samplers.jl:330-374: per-chain Welford
 *                                              moments and lag-1 sums on the device (ESS, R-hat without the draws)
 * The only FFI the reference itself has is the per-element ccall into libRmath (src/distribs.jl:65-67);
 * those nine densities are device functions inside the library.
 *
 * Conventions: plain C, no torch/CUDA types in signatures.  Every function returns 0 on success and a
 * negative smc_status on failure; smc_last_error() (thread local) describes the last failure.  All
 * host pointers are caller-owned; the library copies in/out and never frees or retains them.  Calls on
 * one context are stream-ordered and return after completion.  There is no CPU fallback: without a
 * CUDA device smc_ctx_create fails.
 *
 * Layouts: `beta`/`grad`/`init` are [C][M] row-major (chain c's vector is contiguous, Julia Vector per
 * chain).  Draws are [samples][C][M].  Matrices bound as data are column major (Julia).
 */
#ifndef SIMPLEMCMC_H
#define SIMPLEMCMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMC_ABI_VERSION 2

typedef enum {
  SMC_OK = 0,
  SMC_ERR_INVALID = -1, /* bad argument / malformed tape */
  SMC_ERR_CUDA = -2,    /* CUDA runtime failure (message carries the CUDA error string) */
  SMC_ERR_UNBOUND = -3, /* data referenced by the tape has not been bound */
  SMC_ERR_SUPPORT = -4, /* "Initial values out of model support" (src/samplers.jl:86,143,197) */
  SMC_ERR_MEMORY = -5,  /* the generic executor would need more device memory than allowed */
  SMC_ERR_COMM = -6     /* NCCL failure */
} smc_status;

typedef struct smc_ctx_s* smc_ctx_t;
typedef struct smc_model_s* smc_model_t;

int smc_abi_version(void);
const char* smc_last_error(void);

int smc_ctx_create(int device, smc_ctx_t* out);
int smc_ctx_destroy(smc_ctx_t ctx);
/* Row-sharded models: attach an NCCL communicator (id from ncclGetUniqueId, broadcast by the host). */
int smc_ctx_comm_init(smc_ctx_t ctx, const void* nccl_unique_id, int id_bytes, int rank, int nranks);
int smc_comm_unique_id(void* out_id, int id_bytes);
/* Row-sharded models on one node, optional: sum the per-evaluation (L, G) block through peer memory (NVLink) instead of
   NCCL.  Every rank exports the 64-byte IPC handle of its exchange buffer, the host gathers the handles in rank order,
   every rank imports them; from then on the context uses the peer path.  Replaces nothing in the reference (it has no
   multi-process mode); the exchange it serves is the sum over rows of src/distribs.jl:100-102. */
int smc_ctx_p2p_export(smc_ctx_t ctx, int nranks, void* handle_out, int handle_bytes);
int smc_ctx_p2p_import(smc_ctx_t ctx, const void* handles, int nranks, int rank);
/* sum-allreduce a small host vector of doubles over the communicator (chain moments for R-hat) */
int smc_ctx_allreduce(smc_ctx_t ctx, double* host_inout, int64_t n);

int smc_model_compile(smc_ctx_t ctx, const void* tape, size_t nbytes, smc_model_t* out);
int smc_model_destroy(smc_model_t m);
/* dims are Julia dims (column major); ndims 0 (scalar), 1 or 2.  Copies host -> device. */
int smc_model_bind_data(smc_model_t m, const char* name, const double* host, const int64_t* dims, int ndims);
/* Allocate registers for up to max_chains chains, evaluate the parameter-independent header once.
 * row_shard != 0: this rank holds a row shard of the GLM data; smc_eval and the samplers all-reduce
 * (logp, grad) over the communicator after every evaluation; prior terms are counted on rank 0 only. */
int smc_model_finalize(smc_model_t m, int64_t max_chains, int row_shard);
int smc_model_nparams(smc_model_t m, int64_t* out);

/* logp[C], grad[C][M] (grad may be NULL).  A chain outside the support gets logp = -inf and a zero
 * gradient (src/parsing.jl:465-467). */
int smc_eval(smc_model_t m, const double* beta, int64_t C, double* logp, double* grad);
/* Evaluate `reps` times on the device-resident state of the last smc_eval; returns the mean CUDA-event
 * duration of one evaluation and of its dominant (GLM) kernel in milliseconds (0 when no GLM kernel). */
int smc_eval_timed(smc_model_t m, int reps, int flush_l2, double* ms_eval, double* ms_glm_kernel);

typedef struct {
  int64_t steps, burnin; /* src/samplers.jl:124 */
  int64_t isteps;        /* leapfrogs per step (HMC) */
  double stepsize;       /* leapfrog step (HMC) */
  uint64_t seed;         /* Philox4x32-10 key; chain c uses counter stream (c + chain_offset) */
  int64_t chain_offset;  /* global id of local chain 0 (chain sharding across ranks) */
  int32_t store_draws;   /* 0: only moments / accept rate are produced */
  int32_t init_per_chain; /* 0: init is [M] shared by all chains, 1: init is [C][M] */
  /* test hook (first-K-step trajectory parity): when non-NULL the sampler consumes these instead of Philox.
   * normals: [steps][C][M]; uniforms: [steps][C][n_uniform] (HMC n_uniform = 1; RWM 1; NUTS see simplemcmc.h docs) */
  const double* inject_normals;
  const double* inject_uniforms;
  int64_t inject_uniforms_per_step;
} smc_sampler_config;

typedef struct {
  double* draws;   /* [samples][C][M] or NULL */
  double* loglik;  /* [samples][C] or NULL */
  double* accept;  /* [samples][C] (0/1) or NULL */
  double* final_state; /* [C][M] or NULL */
  double* accept_rate; /* [C] or NULL */
  double* mean;    /* [C][M] per-chain mean of post-burnin draws, or NULL */
  double* m2;      /* [C][M] per-chain sum of squared deviations, or NULL */
  double* epsilon; /* NUTS: [steps][C] or NULL (res.misc[:epsilon], src/samplers.jl:307) */
  double* jmax;    /* NUTS: [steps][C] or NULL (res.misc[:jmax], src/samplers.jl:308) */
  double device_ms;        /* CUDA-event time of the sampling loop */
  int64_t n_evals;         /* lockstep evaluations of the tape */
  int64_t n_leapfrogs;     /* sum over chains of leapfrogs actually taken */
  int64_t n_kernel_launches;
  /* ABI 2: lag-1 statistics accumulated on the device, so that calcStats! (src/samplers.jl:352-374: ESS from the lag-1
   * autocorrelation) needs no draws on the host.  With x_t the t-th recorded draw:                                    */
  double* first;   /* [C][M] x_0, or NULL                                                                            */
  double* lag1;    /* [C][M] sum_{t>=1} (x_t - x_0)(x_{t-1} - x_0), or NULL                                          */
} smc_run_output;

int smc_hmc_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out);
int smc_rwm_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out);
int smc_nuts_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out);

/* counter-based RNG exposed for tests: fills out[n] with the normals / uniforms chain `chain` would
 * draw at `step` for purpose stream `stream` */
int smc_philox_normals(uint64_t seed, int64_t chain, int64_t step, int32_t stream, int64_t n, double* out);
int smc_philox_uniforms(uint64_t seed, int64_t chain, int64_t step, int32_t stream, int64_t n, double* out);
/* test hook, host only (no device needed): plan one register program (smc_prog_header + instructions + reductions, as
 * in a tape) and write the planned micro-ops as text, one per line, into out[cap]; returns the number of micro-ops or a
 * negative smc_status.  Lets the planner (operand folding, copy propagation, chains) be checked where there is no GPU. */
int smc_vm_plan_describe(const void* program, size_t nbytes, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
