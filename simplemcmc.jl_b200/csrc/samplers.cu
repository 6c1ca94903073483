// Lockstep samplers: every chain advances through the same kernel sequence; per-chain control flow
// (early trajectory stop, accept/reject) is a mask.  Replaces the main loops of
// /root/reference/src/samplers.jl: simpleHMC :146-167 (+ leapFrog :50-59, update! :47),
// simpleRWM :89-113 (Vihola's robust adaptive Metropolis).  State layout: [M][Cpad], one thread per chain
// or per (component, chain) so that every access is coalesced over chains.
#include <math.h>
#include <algorithm>
#include <vector>
#include "smc_internal.h"
#include "philox.cuh"

namespace {

struct Arena {  // buffers of one sampler call, returned to the context's pool (stream ordered) when the call ends
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  ~Arena() { for (void* p : ptrs) smc_dev_free(p, st); }
  template <typename T>
  int alloc(T** out, size_t n) {
    void* p = nullptr;
    cudaError_t e = smc_dev_alloc(&p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e != cudaSuccess) { smc_set_error("device allocation of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e)); return SMC_ERR_CUDA; }
    ptrs.push_back(p);
    *out = (T*)p;
    return SMC_OK;
  }
};

struct RngSrc {
  unsigned long long seed;
  long long chain_offset;
  const double* inj_n;  // [steps][C][M] or null
  const double* inj_u;  // [steps][C][nu] or null
  long long nu;
};

__device__ __forceinline__ double rng_uniform(const RngSrc& r, long long step, long long c, long long C, int k) {
  if (r.inj_u) return r.inj_u[(step * C + c) * r.nu + k];
  return smc_uniform(r.seed, (unsigned long long)(c + r.chain_offset), (unsigned long long)step, (uint32_t)k);
}

// Thread layout of the per-chain sampler kernels: block = (32 chains) x (LY component lanes).  A warp is 32 consecutive
// chains of one lane, so every [M][Cpad] access is a coalesced line; lane ly owns components q = ly, ly + LY, ... of
// its chain (momentum pairs p = ly, ly + LY, ... when drawing normals).  Per-chain scalars are read by all lanes into
// registers, recomputed identically, and committed by lane 0; sums over components go through shared memory in lane
// order (deterministic).
// LY is a template parameter: 8 lanes (32 chains per block) up to a few ten thousand chains, where parallelism over the
// components matters; 1 lane (256 chains per block) beyond, where each row of the state is then read in 2 KB runs per
// block (measured on cfg1 with 2^20 chains: begin 216 -> 127 us, post 110 -> 51 us).
#define SMC_CXB (256 / LY)

struct Ln {
  long long c;
  int ly;
  bool in;
};
template <int LY>
__device__ __forceinline__ Ln lanes(long long C) {
  Ln L;
  L.c = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (blockDim.x = SMC_CXB, or fewer chains per block: k_nuts_post)
  L.ly = threadIdx.y;
  L.in = L.c < C;
  return L;
}
template <int LY>
__device__ __forceinline__ double ysum(double v, double* sh) {   // sh[LY * SMC_CXB]; call from every thread of the block
  if (LY == 1) return v;                                          // (one lane per chain: nothing to sum, no barrier)
  sh[threadIdx.y * SMC_CXB + threadIdx.x] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int y = 0; y < LY; y++) t += sh[y * SMC_CXB + threadIdx.x];
  __syncthreads();
  return t;
}
template <int LY>
__device__ __forceinline__ void ysum2(double& a, double& b, double* sh) {   // sh[2 * LY * SMC_CXB]
  if (LY == 1) return;
  sh[threadIdx.y * SMC_CXB + threadIdx.x] = a;
  sh[LY * SMC_CXB + threadIdx.y * SMC_CXB + threadIdx.x] = b;
  __syncthreads();
  double ta = 0.0, tb = 0.0;
#pragma unroll
  for (int y = 0; y < LY; y++) { ta += sh[y * SMC_CXB + threadIdx.x]; tb += sh[LY * SMC_CXB + y * SMC_CXB + threadIdx.x]; }
  __syncthreads();
  a = ta; b = tb;
}

// this lane's share of v[j][c] ~ N(0,1), j < M (also stored to v2 when given); returns its part of sum v^2
template <int LY>
__device__ __forceinline__ double draw_normals(const RngSrc& r, long long step, long long c, long long C, long long M,
                                               long long Cpad, double* v, double* v2, int ly) {
  double ss = 0.0;
  if (r.inj_n) {
    const double* src = r.inj_n + (step * C + c) * M;
    for (long long j = ly; j < M; j += LY) {
      double z = src[j];
      v[j * Cpad + c] = z;
      if (v2) v2[j * Cpad + c] = z;
      ss += z * z;
    }
  } else {
    for (long long p = ly; 2 * p < M; p += LY) {
      const long long j = 2 * p;
      double z0, z1;
      smc_normal_pair(r.seed, (unsigned long long)(c + r.chain_offset), (unsigned long long)step, (uint32_t)p, z0, z1);
      v[j * Cpad + c] = z0; ss += z0 * z0;
      if (v2) v2[j * Cpad + c] = z0;
      if (j + 1 < M) {
        v[(j + 1) * Cpad + c] = z1; ss += z1 * z1;
        if (v2) v2[(j + 1) * Cpad + c] = z1;
      }
    }
  }
  return ss;
}

struct Recorder {
  double* draws;   // [samples][M][Cpad]
  double* loglik;  // [samples][Cpad]
  double* accept;  // [samples][Cpad]
  // running statistics of every (chain, component), interleaved so that a warp (32 chains) updates one contiguous
  // 1.25 KB block per component: stats[((c / 32) * M + j) * 5 + k][c % 32], k = mean, m2 (Welford), x_0 (first recorded
  // draw), previous draw, sum_{t>=1} (x_t - x_0)(x_{t-1} - x_0) (lag-1 statistics of calcStats! without the draws).
  // (Measured at 2^20 chains: the layout alone changed little, 795 -> 707 us for the accept kernel; what mattered was
  // issuing all loads of four components before the first store, 707 -> 191 us -- see k_hmc_accept.)
  double* stats;
  double* nacc;    // [Cpad]
  long long burnin;
};

// addToRes! (src/samplers.jl:342-350) + per-chain Welford moments; lane ly records its components of beta
template <int LY>
__device__ __forceinline__ void record(const Recorder& R, long long step, long long c, long long M, long long Cpad,
                                       const double* beta, double ll, bool accepted, int ly) {
  if (step < R.burnin) return;
  long long s = step - R.burnin;
  if (ly == 0) {
    if (R.loglik) R.loglik[s * Cpad + c] = ll;
    if (R.accept) R.accept[s * Cpad + c] = accepted ? 1.0 : 0.0;
    R.nacc[c] += accepted ? 1.0 : 0.0;
  }
  double n = (double)(s + 1);
  for (long long j = ly; j < M; j += LY) {
    double x = beta[j * Cpad + c];
    if (R.draws) R.draws[(s * M + j) * Cpad + c] = x;
    double* st = R.stats + (((c >> 5) * M + j) * 5) * 32 + (c & 31);
    double mu = st[0 * 32], d = x - mu;
    mu += d / n;
    st[0 * 32] = mu;
    st[1 * 32] += d * (x - mu);
    if (s == 0) {
      st[2 * 32] = x;
    } else {
      double x0 = st[2 * 32];
      st[4 * 32] += (x - x0) * (st[3 * 32] - x0);
    }
    st[3 * 32] = x;
  }
}

// ---------------------------------------------------------------------------------------------- HMC
// state0.v = randn ; update!(state0) ; state = state0     (src/samplers.jl:151-153)
// (the iteration number lives in device memory so that a whole iteration can be replayed as one CUDA graph)
template <int LY>
__global__ void __launch_bounds__(256) k_hmc_begin(RngSrc rng, const long long* stepp, long long C, long long M, long long Cpad,
                                                      const double* b0, const double* g0, const double* l0, double* beta,
                                                      double* grad, double* logp, double* v, double* H0, double* H, int* active) {
  __shared__ double sh[LY * SMC_CXB];
  const long long step = *stepp;
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  double ss = 0.0;
  if (L.in) {
    ss = draw_normals<LY>(rng, step, c, C, M, Cpad, v, nullptr, L.ly);
    for (long long j = L.ly; j < M; j += LY) {
      beta[j * Cpad + c] = b0[j * Cpad + c];
      grad[j * Cpad + c] = g0[j * Cpad + c];
    }
  }
  ss = ysum<LY>(ss, sh);
  if (L.in && L.ly == 0) {
    double ll = l0[c];
    logp[c] = ll;
    H0[c] = ll - ss / 2.0;
    H[c] = H0[c];
    active[c] = isfinite(ll) ? 1 : 0;
  }
}

// n.v += n.grad * ve / 2 ; n.beta += ve * n.v          (src/samplers.jl:52-53)
__global__ void k_leapfrog_pre(long long C, long long M, long long Cpad, double eps, const int* active, double* beta,
                               const double* grad, double* v) {
  long long total = C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    if (!active[c]) continue;
    long long o = j * Cpad + c;
    double vv = v[o] + grad[o] * eps / 2.0;
    v[o] = vv;
    beta[o] += eps * vv;
  }
}

// n.v += n.grad * ve / 2 ; update!(n) ; loop condition isfinite(state.llik)   (src/samplers.jl:55-56,156)
template <int LY>
__global__ void __launch_bounds__(256) k_leapfrog_post(long long C, long long M, long long Cpad, double eps, int* active,
                                                          const double* grad, const double* logp, double* v, double* H,
                                                          unsigned long long* nleap) {
  __shared__ double sh[LY * SMC_CXB];
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  const bool act = L.in && active[c];
  double ss = 0.0;
  if (act)
    for (long long j = L.ly; j < M; j += LY) {
      long long o = j * Cpad + c;
      double vv = v[o] + grad[o] * eps / 2.0;
      v[o] = vv;
      ss += vv * vv;
    }
  ss = ysum<LY>(ss, sh);   // (also orders every lane's read of active[c] before lane 0 clears it)
  if (act && L.ly == 0) {
    double ll = logp[c];
    H[c] = ll - ss / 2.0;
    if (!isfinite(ll)) active[c] = 0;
    nleap[c] += 1ull;
  }
}

// if rand() < exp(state.H - state0.H): state0 = state ; addToRes!   (src/samplers.jl:162-166)
// The state copy and the recorder update are done four components at a time with every load issued before the first
// store: with one component per iteration the kernel had ~1 load in flight per thread and ran at 1.5 TB/s.
template <int LY>
__global__ void __launch_bounds__(256) k_hmc_accept(RngSrc rng, const long long* stepp, long long C, long long M, long long Cpad,
                                                     const double* __restrict__ beta, const double* __restrict__ grad,
                                                     const double* __restrict__ logp, const double* __restrict__ H,
                                                     const double* __restrict__ H0, double* b0, double* g0, double* l0, Recorder R) {
  const long long step = *stepp;
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  bool acc = false;
  double ll = 0.0;
  if (L.in) {
    double u = rng_uniform(rng, step, c, C, 0);
    acc = u < exp(H[c] - H0[c]);
    ll = acc ? logp[c] : l0[c];
  }
  __syncthreads();   // every lane has read l0[c]
  if (!L.in) return;
  const bool rec = step >= R.burnin;
  const long long s = step - R.burnin;
  if (L.ly == 0) {
    if (acc) l0[c] = ll;
    if (rec) {
      if (R.loglik) R.loglik[s * Cpad + c] = ll;
      if (R.accept) R.accept[s * Cpad + c] = acc ? 1.0 : 0.0;
      R.nacc[c] += acc ? 1.0 : 0.0;
    }
  }
  if (!acc && !rec) return;
  const double n = (double)(s + 1);
  for (long long j0 = L.ly; j0 < M; j0 += 4 * LY) {
    double x[4], g[4], st0[4], st1[4], st2[4], st3[4], st4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long j = j0 + (long long)k * LY;
      const bool ok = j < M;
      const long long o = (ok ? j : 0) * Cpad + c;
      x[k] = acc ? beta[o] : b0[o];
      g[k] = acc ? grad[o] : 0.0;
      st0[k] = st1[k] = st2[k] = st3[k] = st4[k] = 0.0;
      if (rec && ok) {
        const double* st = R.stats + (((c >> 5) * M + j) * 5) * 32 + (c & 31);
        st0[k] = st[0 * 32]; st1[k] = st[1 * 32];
        if (s > 0) { st2[k] = st[2 * 32]; st3[k] = st[3 * 32]; st4[k] = st[4 * 32]; }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long j = j0 + (long long)k * LY;
      if (j >= M) break;
      const long long o = j * Cpad + c;
      if (acc) { b0[o] = x[k]; g0[o] = g[k]; }
      if (rec) {
        if (R.draws) R.draws[(s * M + j) * Cpad + c] = x[k];
        double* st = R.stats + (((c >> 5) * M + j) * 5) * 32 + (c & 31);
        double mu = st0[k], d = x[k] - mu;
        mu += d / n;
        st[0 * 32] = mu;
        st[1 * 32] = st1[k] + d * (x[k] - mu);
        if (s == 0) st[2 * 32] = x[k];
        else st[4 * 32] = st4[k] + (x[k] - st2[k]) * (st3[k] - st2[k]);
        st[3 * 32] = x[k];
      }
    }
  }
}

__global__ void k_step_inc(long long* stepp) { *stepp += 1; }

// ---------------------------------------------------------------------------------------------- RWM
// jump = randn ; beta += S * jump                      (src/samplers.jl:93-95)
__global__ void k_rwm_propose(RngSrc rng, long long step, long long C, long long M, long long Cpad, const double* b0,
                              const double* S, double* jump, double* beta) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int y = 0; y < 8; y++) draw_normals<8>(rng, step, c, C, M, Cpad, jump, nullptr, y);   // thread per chain here
  for (long long i = 0; i < M; i++) {
    double acc = 0.0;
    for (long long k = 0; k <= i; k++) acc += S[(i + k * M) * Cpad + c] * jump[k * Cpad + c];  // S lower triangular
    beta[i * Cpad + c] = b0[i * Cpad + c] + acc;
  }
}

// accept/reject + robust adaptive Metropolis update of S (src/samplers.jl:99-112).
// chol(S (I + a u u') S')' == chol(S S' + a (S u)(S u)')' : rank-one update (a > 0) / downdate (a < 0) of the lower factor.
__global__ void k_rwm_accept(RngSrc rng, long long step, long long C, long long M, long long Cpad, const double* beta,
                             const double* logp, double* b0, double* l0, double* S, const double* jump, double* w,
                             Recorder R) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double lnew = logp[c], lold = l0[c];
  double alpha = fmin(1.0, exp(lnew - lold));
  double u = rng_uniform(rng, step, c, C, 0);
  bool rejected = u > alpha;
  if (!rejected) {
    for (long long j = 0; j < M; j++) b0[j * Cpad + c] = beta[j * Cpad + c];
    l0[c] = lnew;
  }
  double lcur = l0[c];
  for (int y = 0; y < 8; y++) record<8>(R, step, c, M, Cpad, b0, lcur, lold != lcur, y);
  // adaptation
  double i1 = (double)(step + 1);
  double eta = fmin(1.0, (double)M * pow(i1, -2.0 / 3.0));
  double jj = 0.0;
  for (long long k = 0; k < M; k++) { double z = jump[k * Cpad + c]; jj += z * z; }
  double a = eta * (alpha - 0.234) / jj;  // S (I + a * jump jump') S'
  if (!(a == a) || jj == 0.0) return;
  double sgn = a >= 0.0 ? 1.0 : -1.0, sa = sqrt(fabs(a));
  for (long long i = 0; i < M; i++) {
    double acc = 0.0;
    for (long long k = 0; k <= i; k++) acc += S[(i + k * M) * Cpad + c] * jump[k * Cpad + c];
    w[i * Cpad + c] = sa * acc;
  }
  for (long long k = 0; k < M; k++) {
    double Lkk = S[(k + k * M) * Cpad + c], wk = w[k * Cpad + c];
    double r2 = Lkk * Lkk + sgn * wk * wk;
    double r = sqrt(r2);
    double cc = r / Lkk, ss = wk / Lkk;
    S[(k + k * M) * Cpad + c] = r;
    for (long long i = k + 1; i < M; i++) {
      double Lik = (S[(i + k * M) * Cpad + c] + sgn * ss * w[i * Cpad + c]) / cc;
      S[(i + k * M) * Cpad + c] = Lik;
      w[i * Cpad + c] = cc * w[i * Cpad + c] - ss * Lik;
    }
  }
}

__global__ void k_set_identity(double* S, long long C, long long M, long long Cpad) {
  long long total = C * M * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long e = i / C, c = i - e * C;
    S[e * Cpad + c] = (e % M == e / M) ? 1.0 : 0.0;
  }
}

__global__ void k_copy_state(long long C, long long M, long long Cpad, const double* beta, const double* grad,
                             const double* logp, double* b0, double* g0, double* l0) {
  long long total = C * (M + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    if (j == M) l0[c] = logp[c];
    else {
      b0[j * Cpad + c] = beta[j * Cpad + c];
      if (g0) g0[j * Cpad + c] = grad[j * Cpad + c];
    }
  }
}

__global__ void k_in_rows_to_cols(const double* in, double* out, long long C, long long M, long long Cpad, int shared_init) {
  long long total = C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    out[j * Cpad + c] = shared_init ? in[j] : in[c * M + j];
  }
}
// [rows][Cpad] x nblocks  ->  [nblocks][C][rows]
__global__ void k_out_cols_to_rows(const double* in, double* out, long long nblocks, long long C, long long rows, long long Cpad) {
  long long total = nblocks * C * rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i % rows, c = (i / rows) % C, b = i / (rows * C);
    out[i] = in[(b * rows + j) * Cpad + c];
  }
}
__global__ void k_div(double* p, long long n, double d) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] /= d;
}

inline unsigned gridc(long long n, int block = 128) { return (unsigned)std::max<long long>(1, (n + block - 1) / block); }
// lane-parallel kernels: block = (256 / ly chains) x (ly lanes)
inline int lanes_for(long long C) { return C >= 32768 ? 1 : 8; }
#define LAUNCH_LY(kern, ...)                                                                                   \
  do {                                                                                                         \
    if (ly == 1) kern<1><<<(unsigned)((C + 255) / 256), dim3(256, 1), 0, st>>>(__VA_ARGS__);                  \
    else kern<8><<<(unsigned)((C + 31) / 32), dim3(32, 8), 0, st>>>(__VA_ARGS__);                             \
  } while (0)
// k_nuts_post with `cx` chains per block (LY = 8 lanes each): its merge loop is uniform over the block, so with chains at
// different leaves (asynchronous iterations) a smaller block runs fewer levels and its barriers span fewer warps
#define LAUNCH_POST(cx, ...)                                                                                  \
  do {                                                                                                        \
    if (ly == 1) k_nuts_post<1><<<(unsigned)((C + 255) / 256), dim3(256, 1), 0, st>>>(__VA_ARGS__);           \
    else k_nuts_post<8><<<(unsigned)((C + (cx) - 1) / (cx)), dim3((cx), 8), 0, st>>>(__VA_ARGS__);            \
  } while (0)
inline unsigned gride(long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 32)); }

struct Run {
  smc_model_s* m;
  const smc_sampler_config* cfg;
  smc_run_output* out;
  long long C, M, Cpad, samples;
  Arena A;
  RngSrc rng;
  Recorder R;
  double *b0 = nullptr, *g0 = nullptr, *l0 = nullptr;
  unsigned long long* nleap = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int64_t launches0 = 0;
};

int run_setup(Run& r, const double* init, bool need_grad, long long nu) {
  smc_model_s* m = r.m;
  const smc_sampler_config* cfg = r.cfg;
  if (!m->finalized) { smc_set_error("model is not finalized"); return SMC_ERR_INVALID; }
  if (r.C < 1 || r.C > m->Cmax) { smc_set_error("chains %lld outside [1, %lld]", r.C, (long long)m->Cmax); return SMC_ERR_INVALID; }
  if (cfg->burnin < 0) { smc_set_error("Burnin rounds (%lld) should be >= 0", (long long)cfg->burnin); return SMC_ERR_INVALID; }   // src/samplers.jl:325
  if (cfg->steps <= cfg->burnin) { smc_set_error("Steps (%lld) should be > to burnin (%lld)", (long long)cfg->steps, (long long)cfg->burnin); return SMC_ERR_INVALID; }  // :326
  cudaSetDevice(m->ctx->device);
  cudaStream_t st = m->ctx->stream;
  r.M = m->M; r.Cpad = m->Cpad; r.samples = cfg->steps - cfg->burnin;
  long long C = r.C, M = r.M, Cpad = r.Cpad;
  r.launches0 = m->ctx->launches;
  r.A.st = st;
  SMC_TRY(r.A.alloc(&r.b0, M * Cpad));
  SMC_TRY(r.A.alloc(&r.g0, M * Cpad));
  SMC_TRY(r.A.alloc(&r.l0, Cpad));
  SMC_TRY(r.A.alloc(&r.nleap, Cpad));
  SMC_CUDA(cudaMemsetAsync(r.nleap, 0, Cpad * sizeof(unsigned long long), st));
  r.R.burnin = cfg->burnin;
  r.R.draws = r.R.loglik = r.R.accept = nullptr;
  if (cfg->store_draws && r.out->draws) SMC_TRY(r.A.alloc(&r.R.draws, r.samples * M * Cpad));
  if (r.out->loglik) SMC_TRY(r.A.alloc(&r.R.loglik, r.samples * Cpad));
  if (r.out->accept) SMC_TRY(r.A.alloc(&r.R.accept, r.samples * Cpad));
  const size_t nstats = (size_t)((Cpad + 31) / 32) * M * 5 * 32;
  SMC_TRY(r.A.alloc(&r.R.stats, nstats));
  SMC_CUDA(cudaMemsetAsync(r.R.stats, 0, nstats * sizeof(double), st));
  SMC_TRY(r.A.alloc(&r.R.nacc, Cpad));
  SMC_CUDA(cudaMemsetAsync(r.R.nacc, 0, Cpad * sizeof(double), st));
  r.rng.seed = cfg->seed; r.rng.chain_offset = cfg->chain_offset; r.rng.inj_n = nullptr; r.rng.inj_u = nullptr; r.rng.nu = nu;
  if (cfg->inject_normals) {
    double* d;
    size_t n = (size_t)cfg->steps * C * M;
    SMC_TRY(r.A.alloc(&d, n));
    SMC_CUDA(cudaMemcpyAsync(d, cfg->inject_normals, n * sizeof(double), cudaMemcpyHostToDevice, st));
    r.rng.inj_n = d;
  }
  if (cfg->inject_uniforms) {
    if (cfg->inject_uniforms_per_step < nu) { smc_set_error("inject_uniforms_per_step must be >= %lld for this sampler", nu); return SMC_ERR_INVALID; }
    r.rng.nu = cfg->inject_uniforms_per_step;
    double* d;
    size_t n = (size_t)cfg->steps * C * r.rng.nu;
    SMC_TRY(r.A.alloc(&d, n));
    SMC_CUDA(cudaMemcpyAsync(d, cfg->inject_uniforms, n * sizeof(double), cudaMemcpyHostToDevice, st));
    r.rng.inj_u = d;
  }
  // initial state
  double* stage;
  size_t ninit = cfg->init_per_chain ? (size_t)C * M : (size_t)M;
  SMC_TRY(r.A.alloc(&stage, ninit));
  SMC_CUDA(cudaMemcpyAsync(stage, init, ninit * sizeof(double), cudaMemcpyHostToDevice, st));
  k_in_rows_to_cols<<<gride(C * M), 256, 0, st>>>(stage, m->beta, C, M, Cpad, cfg->init_per_chain ? 0 : 1);
  m->ctx->launches++;
  SMC_TRY(smc_model_eval_device(m, C, need_grad));  // calc!(state0, ll_func), src/samplers.jl:142
  k_copy_state<<<gride(C * (M + 1)), 256, 0, st>>>(C, M, Cpad, m->beta, need_grad ? m->grad : nullptr, m->logp, r.b0,
                                                  need_grad ? r.g0 : nullptr, r.l0);
  m->ctx->launches++;
  // assert(isfinite(state0.llik), "Initial values out of model support, try other values")  src/samplers.jl:143
  std::vector<double> l0h(C);
  SMC_CUDA(cudaMemcpyAsync(l0h.data(), r.l0, C * sizeof(double), cudaMemcpyDeviceToHost, st));
  SMC_CUDA(cudaStreamSynchronize(st));
  for (long long c = 0; c < C; c++)
    if (!isfinite(l0h[c])) {
      smc_set_error("Initial values out of model support, try other values (chain %lld, log-density %g)", c, l0h[c]);
      return SMC_ERR_SUPPORT;
    }
  SMC_CUDA(cudaEventCreate(&r.e0));
  SMC_CUDA(cudaEventCreate(&r.e1));
  SMC_CUDA(cudaEventRecord(r.e0, st));
  return SMC_OK;
}

int copy_out_blocks(Run& r, const double* dev, double* host, long long nblocks, long long rows) {
  // dev [nblocks][rows][Cpad] -> host [nblocks][C][rows], staged in chunks
  if (!host || !dev) return SMC_OK;
  cudaStream_t st = r.m->ctx->stream;
  long long per = r.C * rows;
  long long chunk = std::max<long long>(1, std::min<long long>(nblocks, (64LL << 20) / std::max<long long>(per, 1)));
  double* stage;
  Arena local;
  local.st = st;
  SMC_TRY(local.alloc(&stage, chunk * per));
  for (long long b = 0; b < nblocks; b += chunk) {
    long long nb = std::min(chunk, nblocks - b);
    k_out_cols_to_rows<<<gride(nb * per), 256, 0, st>>>(dev + b * rows * r.Cpad, stage, nb, r.C, rows, r.Cpad);
    r.m->ctx->launches++;
    SMC_CUDA(cudaMemcpyAsync(host + b * per, stage, nb * per * sizeof(double), cudaMemcpyDeviceToHost, st));
    SMC_CUDA(cudaStreamSynchronize(st));
  }
  return SMC_OK;
}

// statistic k of the interleaved recorder block -> host [C][M]
__global__ void k_out_stats(const double* stats, int k, long long C, long long M, double* out) {
  long long total = C * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long c = i / M, j = i - c * M;
    out[i] = stats[((((c >> 5) * M + j) * 5) + k) * 32 + (c & 31)];
  }
}
int copy_out_stats(Run& r, int k, double* host) {
  if (!host) return SMC_OK;
  cudaStream_t st = r.m->ctx->stream;
  double* stage;
  Arena local;
  local.st = st;
  SMC_TRY(local.alloc(&stage, r.C * r.M));
  k_out_stats<<<gride(r.C * r.M), 256, 0, st>>>(r.R.stats, k, r.C, r.M, stage);
  r.m->ctx->launches++;
  SMC_CUDA(cudaMemcpyAsync(host, stage, (size_t)r.C * r.M * sizeof(double), cudaMemcpyDeviceToHost, st));
  SMC_CUDA(cudaStreamSynchronize(st));
  return SMC_OK;
}

int run_finish(Run& r, int64_t n_evals) {
  smc_model_s* m = r.m;
  cudaStream_t st = m->ctx->stream;
  SMC_CUDA(cudaEventRecord(r.e1, st));
  SMC_CUDA(cudaEventSynchronize(r.e1));
  float ms = 0.f;
  SMC_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
  SMC_CUDA(cudaGetLastError());
  r.out->device_ms = ms;
  r.out->n_evals = n_evals;
  cudaEventDestroy(r.e0);
  cudaEventDestroy(r.e1);
  std::vector<unsigned long long> nl(r.C);
  SMC_CUDA(cudaMemcpyAsync(nl.data(), r.nleap, r.C * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  SMC_CUDA(cudaStreamSynchronize(st));
  if (m->row_shard) SMC_TRY(smc_p2p_check(m->ctx));     // a peer that never arrived: the draws are not to be trusted
  unsigned long long tot = 0;
  for (auto x : nl) tot += x;
  r.out->n_leapfrogs = (int64_t)tot;
  SMC_TRY(copy_out_blocks(r, r.R.draws, r.out->draws, r.samples, r.M));
  SMC_TRY(copy_out_blocks(r, r.R.loglik, r.out->loglik, r.samples, 1));
  SMC_TRY(copy_out_blocks(r, r.R.accept, r.out->accept, r.samples, 1));
  SMC_TRY(copy_out_blocks(r, r.b0, r.out->final_state, 1, r.M));
  SMC_TRY(copy_out_stats(r, 0, r.out->mean));
  SMC_TRY(copy_out_stats(r, 1, r.out->m2));
  SMC_TRY(copy_out_stats(r, 2, r.out->first));
  SMC_TRY(copy_out_stats(r, 4, r.out->lag1));
  if (r.out->accept_rate) {
    k_div<<<gride(r.C), 256, 0, st>>>(r.R.nacc, r.C, (double)r.samples);
    m->ctx->launches++;
    SMC_TRY(copy_out_blocks(r, r.R.nacc, r.out->accept_rate, 1, 1));
  }
  r.out->n_kernel_launches = m->ctx->launches - r.launches0;
  return SMC_OK;
}

}  // namespace

extern "C" int smc_hmc_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out) {
  if (!m || !cfg || !init || !out) { smc_set_error("smc_hmc_run: NULL argument"); return SMC_ERR_INVALID; }
  if (cfg->isteps < 1) { smc_set_error("smc_hmc_run: isteps must be >= 1"); return SMC_ERR_INVALID; }
  Run r;
  r.m = m; r.cfg = cfg; r.out = out; r.C = C;
  SMC_TRY(run_setup(r, init, true, 1));
  cudaStream_t st = m->ctx->stream;
  long long M = r.M, Cpad = r.Cpad;
  double *v, *H0, *H;
  int* active;
  SMC_TRY(r.A.alloc(&v, M * Cpad));
  SMC_TRY(r.A.alloc(&H0, Cpad));
  SMC_TRY(r.A.alloc(&H, Cpad));
  SMC_TRY(r.A.alloc(&active, Cpad));
  const int ly = lanes_for(C);
  long long* d_step;
  SMC_TRY(r.A.alloc(&d_step, 1));
  SMC_CUDA(cudaMemsetAsync(d_step, 0, sizeof(long long), st));
  int64_t evals = 1;
  // one iteration = begin, isteps x (half kick + drift, evaluation, half kick), accept
  auto iteration = [&]() -> int {
    LAUNCH_LY(k_hmc_begin, r.rng, d_step, C, M, Cpad, r.b0, r.g0, r.l0, m->beta, m->grad, m->logp, v, H0, H, active);
    for (long long j = 0; j < cfg->isteps; j++) {
      k_leapfrog_pre<<<gride(C * M), 256, 0, st>>>(C, M, Cpad, cfg->stepsize, active, m->beta, m->grad, v);
      SMC_TRY(smc_model_eval_device(m, C, true));
      LAUNCH_LY(k_leapfrog_post, C, M, Cpad, cfg->stepsize, active, m->grad, m->logp, v, H, r.nleap);
    }
    LAUNCH_LY(k_hmc_accept, r.rng, d_step, C, M, Cpad, m->beta, m->grad, m->logp, H, H0, r.b0, r.g0, r.l0, r.R);
    k_step_inc<<<1, 1, 0, st>>>(d_step);
    m->ctx->launches += 3 + 2 * cfg->isteps;
    return SMC_OK;
  };
  // The first two iterations run eagerly (the evaluation sizes its buffers, then captures its own graph); the third is
  // captured whole -- the evaluation graph becomes a child node -- and every later iteration is one graph launch.
  cudaGraphExec_t exec = nullptr;
  int64_t per_iter = 0;
  const bool can_graph = m->graphs_enabled && smc_comm_graph_ok(m);
  int rc = SMC_OK;
  for (long long i = 0; i < cfg->steps && rc == SMC_OK; i++) {
    if (exec) {
      if (cudaGraphLaunch(exec, st) != cudaSuccess) { smc_set_error("cudaGraphLaunch failed in the HMC loop"); rc = SMC_ERR_CUDA; break; }
      m->ctx->launches += per_iter;
      evals += cfg->isteps;
      continue;
    }
    if (i == 2 && can_graph && cfg->steps > 3 && cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      const int64_t l0 = m->ctx->launches, epoch = m->realloc_epoch;
      int brc = iteration();
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      bool ok = (brc == SMC_OK && e == cudaSuccess && graph && epoch == m->realloc_epoch);
      if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { ok = false; exec = nullptr; }
      if (graph) cudaGraphDestroy(graph);
      per_iter = m->ctx->launches - l0;
      m->ctx->launches = l0;
      if (ok) { i--; continue; }        // nothing has run yet: replay the captured iteration for this i
      cudaGetLastError();
      exec = nullptr;
    }
    rc = iteration();
    evals += cfg->isteps;
  }
  if (exec) { cudaStreamSynchronize(st); cudaGraphExecDestroy(exec); }
  if (rc != SMC_OK) return rc;
  return run_finish(r, evals);
}

extern "C" int smc_rwm_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out) {
  if (!m || !cfg || !init || !out) { smc_set_error("smc_rwm_run: NULL argument"); return SMC_ERR_INVALID; }
  Run r;
  r.m = m; r.cfg = cfg; r.out = out; r.C = C;
  SMC_TRY(run_setup(r, init, false, 1));
  cudaStream_t st = m->ctx->stream;
  long long M = r.M, Cpad = r.Cpad;
  double *S, *jump, *w;
  SMC_TRY(r.A.alloc(&S, M * M * Cpad));
  SMC_TRY(r.A.alloc(&jump, M * Cpad));
  SMC_TRY(r.A.alloc(&w, M * Cpad));
  k_set_identity<<<gride(C * M * M), 256, 0, st>>>(S, C, M, Cpad);  // S = eye(nparams), src/samplers.jl:89
  m->ctx->launches++;
  int64_t evals = 1;
  for (long long i = 0; i < cfg->steps; i++) {
    k_rwm_propose<<<gridc(C), 128, 0, st>>>(r.rng, i, C, M, Cpad, r.b0, S, jump, m->beta);
    SMC_TRY(smc_model_eval_device(m, C, false));
    evals++;
    k_rwm_accept<<<gridc(C), 128, 0, st>>>(r.rng, i, C, M, Cpad, m->beta, m->logp, r.b0, r.l0, S, jump, w, r.R);
    m->ctx->launches += 2;
  }
  return run_finish(r, evals);
}

// ---------------------------------------------------------------------------------------------- RNG test hooks
namespace {
__global__ void k_philox_normals(unsigned long long seed, long long chain, long long step, long long n, double* out) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * b >= n) return;
  double z0, z1;
  smc_normal_pair(seed, (unsigned long long)chain, (unsigned long long)step, (uint32_t)b, z0, z1);
  out[2 * b] = z0;
  if (2 * b + 1 < n) out[2 * b + 1] = z1;
}
__global__ void k_philox_uniforms(unsigned long long seed, long long chain, long long step, long long n, double* out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = smc_uniform(seed, (unsigned long long)chain, (unsigned long long)step, (uint32_t)k);
}
}  // namespace

extern "C" int smc_philox_normals(uint64_t seed, int64_t chain, int64_t step, int32_t stream, int64_t n, double* out) {
  (void)stream;
  if (!out || n < 0) return SMC_ERR_INVALID;
  double* d = nullptr;
  SMC_CUDA(cudaMalloc(&d, std::max<int64_t>(n, 1) * sizeof(double)));
  k_philox_normals<<<(unsigned)((n / 2 + 128) / 128), 128>>>(seed, chain, step, n, d);
  cudaError_t e = cudaMemcpy(out, d, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  SMC_CUDA(e);
  return SMC_OK;
}
extern "C" int smc_philox_uniforms(uint64_t seed, int64_t chain, int64_t step, int32_t stream, int64_t n, double* out) {
  (void)stream;
  if (!out || n < 0) return SMC_ERR_INVALID;
  double* d = nullptr;
  SMC_CUDA(cudaMalloc(&d, std::max<int64_t>(n, 1) * sizeof(double)));
  k_philox_uniforms<<<(unsigned)((n + 127) / 128), 128>>>(seed, chain, step, n, d);
  cudaError_t e = cudaMemcpy(out, d, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  SMC_CUDA(e);
  return SMC_OK;
}

// ---------------------------------------------------------------------------------------------- NUTS
// Efficient NUTS (Hoffman & Gelman) as the reference codes it: /root/reference/src/samplers.jl:199-310.
// The recursive buildTree (:224-257) becomes an iteration over the 2^j leaves of a doubling; completed
// sub-trees wait on a per-chain stack indexed by level (a binary counter), merged exactly where the recursion
// would return.  Chains run in lockstep one leapfrog per round; a chain whose tree is finished idles (masked)
// until every chain is done.  Uniform draws are consumed in the reference's order (slice, direction, one per
// merge, one per accepted doubling) from the chain's (step, k) counter stream.
namespace {

constexpr int kMaxDepth = 8;    // `while s && j < 8`, src/samplers.jl:279
constexpr double kDeltaMax = 100.0;

struct Nuts {
  // outer tree
  double *mb, *mv, *mg, *ml, *pb, *pv, *pg, *pl;  // minus / plus edge states
  double *sb, *sg, *sl;                             // selected state of the iteration
  // the trajectory being extended: position, velocity, gradient, log-density of every chain.  The model's evaluation
  // buffers hold only the chains that are still building a tree, packed densely: chain c evaluates in column slot[c].
  double *tb, *tv, *tg, *tl;
  int* slot;
  int* blocksum;
  // sub-tree under construction ("current") and the stack of completed sub-trees
  double *cnb, *cnv, *csb, *csg, *csl;  // current: near edge (beta, v), selected (beta, grad, logp)
  double *knb, *knv, *ksb, *ksg;        // stack[level][M][Cpad]
  double *ksl, *kn, *ka, *kna;          // stack[level][Cpad]
  double *u, *H0, *eps, *hbar, *lebar, *alast, *nalast, *ntot;
  int *j, *leaf, *dir, *active, *newdbl, *ucount, *moved;
  int* n_active;
  unsigned long long* rounds_active;
  // asynchronous iterations (below): the chain's own iteration number, and the round in which it has to end / begin one
  long long* istep;
  int *end_round, *begin_round;
};
// the iteration number a kernel works on: one counter in device memory (lockstep: every chain is in the same iteration, and a
// captured graph replays the kernels while the counter advances) or the chain's own (asynchronous iterations)
struct StepSrc {
  const long long* global;
  const long long* per_chain;
  __device__ __forceinline__ long long of(long long c, bool in) const { return per_chain ? (in ? per_chain[c] : 0) : *global; }
};

// this lane's share of the two U-turn dot products (src/samplers.jl:48):
//   d1 = dot(s2.beta - s1.beta, s1.v),  d2 = dot(s2.beta - s1.beta, s2.v)      (s1 = minus edge, s2 = plus edge)
template <int LY>
__device__ __forceinline__ void uturn_part(const double* minus_b, const double* minus_v, const double* plus_b,
                                           const double* plus_v, long long c, long long M, long long Cpad, int ly,
                                           double& d1, double& d2) {
#pragma unroll 4
  for (long long k = ly; k < M; k += LY) {
    double db = plus_b[k * Cpad + c] - minus_b[k * Cpad + c];
    d1 += db * minus_v[k * Cpad + c];
    d2 += db * plus_v[k * Cpad + c];
  }
}

template <int LY>
__device__ __forceinline__ void copy_vec(double* dst, const double* src, long long c, long long M, long long Cpad, int ly) {
  // four loads in flight before the first store (dst and src may be any of the state vectors: no restrict)
  for (long long k0 = ly; k0 < M; k0 += 4 * LY) {
    double t[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long k = k0 + (long long)u * LY;
      t[u] = k < M ? src[k * Cpad + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long k = k0 + (long long)u * LY;
      if (k < M) dst[k * Cpad + c] = t[u];
    }
  }
}

// dense evaluation slots for the chains that are still building a tree: slot[c] = number of active chains before c
__global__ void k_slot_count(const int* active, long long C, int* blocksum) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int n = __syncthreads_count(c < C && active[c]);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = n;
}
__global__ void k_slot_scan(int* blocksum, int nb) {   // exclusive scan in place, one block
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nb ? blocksum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = sh[threadIdx.x];
    if (i < nb) blocksum[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
}
__global__ void k_slot_assign(const int* active, long long C, const int* blocksum, int* slot) {
  __shared__ int wsum[8];
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int a = (c < C && active[c]) ? 1 : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, a);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = __popc(bal);
  __syncthreads();
  int off = blocksum[blockIdx.x];
  for (int k = 0; k < w; k++) off += wsum[k];
  if (a) slot[c] = off + __popc(bal & ((1u << lane) - 1u));
}

// the three steps above in ONE block (up to 65 536 chains): thread t owns `per` consecutive chains (independent loads, all in
// flight at once), counts its active ones, takes the block-wide exclusive prefix of the counts (shuffle scan per warp, then
// over the 32 warp totals), and numbers its active chains in chain order -- the same slots
__global__ void __launch_bounds__(1024) k_slot_pack(const int* __restrict__ active, long long C, int* __restrict__ slot) {
  __shared__ int wtot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (int)((C + 1023) / 1024);                   // <= 64
  const long long b = (long long)threadIdx.x * per;
  unsigned long long bits = 0ull;                             // this thread's flags
#pragma unroll 8
  for (int u = 0; u < per; u++) {
    const long long c = b + u;
    bits |= (unsigned long long)((c < C && active[c] != 0) ? 1 : 0) << u;
  }
  const int n = __popcll(bits);
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int v = wtot[lane];
    int s2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s2, o);
      if (lane >= o) s2 += t;
    }
    wtot[lane] = s2 - v;                                      // exclusive prefix of the warp totals
  }
  __syncthreads();
  int base = wtot[w] + incl - n;
  for (int u = 0; u < per; u++)
    if ((bits >> u) & 1ull) slot[b + u] = base++;
}

__global__ void k_nuts_init(long long C, Nuts T) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  T.eps[c] = 1.0;       // the initial-epsilon search of :199-210 never runs (state0.H is NaN): epsilon0 = 1
  T.hbar[c] = 0.0;
  T.lebar[c] = 0.0;
}

template <int LY>
__global__ void __launch_bounds__(256) k_nuts_begin(RngSrc rng, StepSrc steps, int round, long long C, long long M, long long Cpad,
                                                       const double* b0, const double* g0, const double* l0, Nuts T) {
  __shared__ double sh[LY * SMC_CXB];
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  const long long step = steps.of(c, L.in);
  // asynchronous iterations: only the chains whose previous iteration ended in the last round begin one now
  const bool go = L.in && (round < 0 || T.begin_round[c] == round);
  double ss = 0.0;
  if (go) {
    ss = draw_normals<LY>(rng, step, c, C, M, Cpad, T.mv, T.pv, L.ly);
    for (long long k = L.ly; k < M; k += LY) {
      long long o = k * Cpad + c;
      double b = b0[o], g = g0[o];
      T.mb[o] = b; T.pb[o] = b; T.sb[o] = b;
      T.mg[o] = g; T.pg[o] = g; T.sg[o] = g;
    }
  }
  ss = ysum<LY>(ss, sh);
  if (!go || L.ly != 0) return;
  double ll = l0[c];
  T.ml[c] = ll; T.pl[c] = ll; T.sl[c] = ll;
  double H0 = ll - ss / 2.0;
  T.H0[c] = H0;
  T.u[c] = log(rng_uniform(rng, step, c, C, 0)) + H0;   // u_slice = log(rand()) + state0.H, :272
  T.ucount[c] = 1;
  T.j[c] = 0;
  T.ntot[c] = 1.0;
  T.active[c] = 1;
  T.newdbl[c] = 1;
  T.moved[c] = 0;
  T.alast[c] = 0.0;
  T.nalast[c] = 1.0;
}

template <int LY>
__global__ void __launch_bounds__(256) k_nuts_pre(RngSrc rng, StepSrc steps, long long C, long long M, long long Cpad,
                                                     double* beta, int use_slots, Nuts T) {
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  const long long step = steps.of(c, L.in);
  const bool act = L.in && T.active[c];
  if (blockIdx.x == 0 && threadIdx.x == 0) *T.n_active = 0;   // counted again by k_nuts_post / k_nuts_end of this round
  int newdbl = 0, dir = 0, k = 0;
  double eps = 0.0, lp_edge = 0.0;
  long long slot = c;
  if (act) {
    newdbl = T.newdbl[c];
    eps = T.eps[c];
    if (use_slots) slot = T.slot[c];
    if (newdbl) {
      k = T.ucount[c];
      double r = rng_uniform(rng, step, c, C, k);
      dir = (r > 0.5) ? 1 : -1;   // dir = (rand() > 0.5) * 2. - 1., :280
      lp_edge = dir < 0 ? T.ml[c] : T.pl[c];
    } else {
      dir = T.dir[c];
    }
  }
  __syncthreads();   // every lane holds the chain's scalars before lane 0 updates them
  if (!act) return;
  if (newdbl) {
    const double* eb = dir < 0 ? T.mb : T.pb;
    const double* ev = dir < 0 ? T.mv : T.pv;
    const double* eg = dir < 0 ? T.mg : T.pg;
    for (long long q = L.ly; q < M; q += LY) {
      long long o = q * Cpad + c;
      T.tb[o] = eb[o]; T.tv[o] = ev[o]; T.tg[o] = eg[o];
    }
    if (L.ly == 0) {
      T.ucount[c] = k + 1;
      T.dir[c] = dir;
      T.tl[c] = lp_edge;
      T.leaf[c] = 0;
      T.newdbl[c] = 0;
    }
  }
  double ve = dir * eps;
  for (long long q0 = L.ly; q0 < M; q0 += 4 * LY) {   // leapFrog, :52-53 (four components' loads before the stores)
    double v4[4], g4[4], b4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long q = q0 + (long long)u * LY, o = (q < M ? q : 0) * Cpad + c;
      v4[u] = T.tv[o]; g4[u] = T.tg[o]; b4[u] = T.tb[o];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long q = q0 + (long long)u * LY;
      if (q >= M) break;
      const long long o = q * Cpad + c;
      double vv = v4[u] + g4[u] * ve / 2.0;
      T.tv[o] = vv;
      double b = b4[u] + ve * vv;
      T.tb[o] = b;
      beta[q * Cpad + slot] = b;   // the evaluation's input column of this chain
    }
  }
}

#ifndef SMC_NUTS_POST_MINB
#define SMC_NUTS_POST_MINB 2     // 233 registers uncapped = one block per SM; capped at 128 (no spills): +13..18 % leapfrogs/s at 16 384 chains (3 and 4 blocks spill and are no better)
#endif
template <int LY>
__global__ void __launch_bounds__(256, SMC_NUTS_POST_MINB) k_nuts_post(RngSrc rng, StepSrc steps, int round, long long C, long long M, long long Cpad,
                                                      const double* egrad, const double* elogp, int use_slots, Nuts T,
                                                      unsigned long long* nleap) {
  __shared__ double sh[2 * LY * SMC_CXB];
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  const long long step = steps.of(c, L.in);
  const int ly = L.ly;
  const bool act = L.in && T.active[c];
  const long long LS = M * Cpad;  // stride between stack levels
  int dir = 1, leaf = 0, j = 0, ucount = 0;
  double eps = 0.0, u = 0.0, H0 = 0.0, ll = 0.0, ntot = 1.0;
  const double* beta = T.tb;    // position and gradient of the trajectory's newest point
  const double* grad = T.tg;
  long long slot = c;
  if (act) {
    if (use_slots) slot = T.slot[c];
    dir = T.dir[c]; eps = T.eps[c]; u = T.u[c]; H0 = T.H0[c]; ll = elogp[slot];
    leaf = T.leaf[c]; j = T.j[c]; ucount = T.ucount[c]; ntot = T.ntot[c];
  }
  const double ve = dir * eps;
  double ss = 0.0;
  if (act)
    for (long long q0 = ly; q0 < M; q0 += 4 * LY) {   // second half kick + update!, :55-56
      double g4[4], v4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long q = q0 + (long long)u * LY, qq = q < M ? q : 0;
        g4[u] = egrad[qq * Cpad + slot];          // the evaluation's output column of this chain
        v4[u] = T.tv[qq * Cpad + c];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long q = q0 + (long long)u * LY;
        if (q >= M) break;
        const long long o = q * Cpad + c;
        T.tg[o] = g4[u];
        double vv = v4[u] + g4[u] * ve / 2.0;
        T.tv[o] = vv;
        ss += vv * vv;
      }
    }
  ss = ysum<LY>(ss, sh);
  const double H = ll - ss / 2.0;
  // ---- leaf (buildTree with j == 0, :230-236)
  double n2 = (u <= H) ? 1.0 : 0.0;
  bool s2 = u < (kDeltaMax + H);
  double a2 = fmin(1.0, exp(H - H0));
  double na2 = 1.0;
  double csl = ll;
  // the "current" sub-tree starts as the leaf itself: near edge (beta, v) and selected state (beta, grad) ARE the newest point
  // of the trajectory, so until a merge replaces them they are read from there (pointers per chain) instead of being copied
  // into cn* / cs* first -- four of the eight to twelve vector copies of a round
  const double *cur_nb = beta, *cur_nv = T.tv, *cur_sb = beta, *cur_sg = grad;
  // ---- merges where the recursion returns (:238-256): levels = trailing one bits of the leaf index.  The loop is
  // uniform over the block (the U-turn test needs every lane); a chain leaves it at `stop_level`.  (With asynchronous
  // iterations the chains of a block are at different leaves and the loop runs to the deepest merge of any of its 32 chains
  // in every round: 62.7 us per launch at 1024 chains against 18.7 in lockstep.  Putting the lanes of a chain into one warp --
  // shuffles instead of shared-memory sums, a loop uniform per warp -- was measured slower in BOTH modes, 3.8e7 against 4.2e7
  // leapfrogs/s in lockstep: the state vectors are then touched in 32-byte pieces.)
  bool merging = act;
  int stop_level = kMaxDepth;
  for (int level = 0; level < kMaxDepth; level++) {
    bool do_merge = false;
    if (merging) {
      if (level >= j) { merging = false; stop_level = level; }
      else if (!((leaf >> level) & 1)) {
        // this sub-tree is a first child here: complete trees wait on the stack; a cut tree is handed up unchanged
        // (`if s1` is false, :239) and may still be merged, as a second child, further up
        if (s2) { merging = false; stop_level = level; }
      } else {
        do_merge = true;
      }
    }
    if (LY == 1) {                   // one lane per chain: no cooperation, so the loop need not be uniform over the block
      if (!merging) break;
      if (!do_merge) continue;
    } else {
      if (!__syncthreads_or(merging ? 1 : 0)) break;
      if (!__syncthreads_or(do_merge ? 1 : 0)) continue;
    }
    double d1 = 0.0, d2 = 0.0;
    bool test_uturn = false;
    if (do_merge) {
      const long long so = (long long)level * LS, sc = (long long)level * Cpad + c;
      double n1 = T.kn[sc];
      double r = rng_uniform(rng, step, c, C, ucount++);
      bool take_second = r <= n2 / (n2 + n1);   // NaN (0/0) compares false, as in the reference
      if (!take_second) {
        copy_vec<LY>(T.csb, T.ksb + so, c, M, Cpad, ly);
        copy_vec<LY>(T.csg, T.ksg + so, c, M, Cpad, ly);
        cur_sb = T.csb; cur_sg = T.csg;
        csl = T.ksl[sc];
      }
      copy_vec<LY>(T.cnb, T.knb + so, c, M, Cpad, ly);   // the merged sub-tree starts where the first one started
      copy_vec<LY>(T.cnv, T.knv + so, c, M, Cpad, ly);
      cur_nb = T.cnb; cur_nv = T.cnv;
      a2 += T.ka[sc];
      na2 += T.kna[sc];
      if (s2) {
        test_uturn = true;
        if (dir > 0) uturn_part<LY>(T.cnb, T.cnv, beta, T.tv, c, M, Cpad, ly, d1, d2);
        else uturn_part<LY>(beta, T.tv, T.cnb, T.cnv, c, M, Cpad, ly, d1, d2);
      }
      n2 += n1;
    }
    ysum2<LY>(d1, d2, sh);
    if (test_uturn) s2 = !(d1 < 0.0 || d2 < 0.0);
  }
  const bool done = act && ((!s2) || (leaf == (1 << j) - 1));
  if (act && !done) {
    const long long so = (long long)stop_level * LS, sc = (long long)stop_level * Cpad + c;
    copy_vec<LY>(T.knb + so, cur_nb, c, M, Cpad, ly);
    copy_vec<LY>(T.knv + so, cur_nv, c, M, Cpad, ly);
    copy_vec<LY>(T.ksb + so, cur_sb, c, M, Cpad, ly);
    copy_vec<LY>(T.ksg + so, cur_sg, c, M, Cpad, ly);
    if (ly == 0) {
      T.tl[c] = ll;
      T.ksl[sc] = csl;
      T.kn[sc] = n2; T.ka[sc] = a2; T.kna[sc] = na2;
      T.leaf[c] = leaf + 1;
      T.ucount[c] = ucount;
      nleap[c] += 1ull;
      atomicAdd(T.n_active, 1);
    }
  }
  // ---- the doubling is complete (or was cut): outer loop body, :281-291
  double d1 = 0.0, d2 = 0.0;
  if (done) {
    double* eb = dir < 0 ? T.mb : T.pb;
    double* ev = dir < 0 ? T.mv : T.pv;
    double* eg = dir < 0 ? T.mg : T.pg;
    for (long long q = ly; q < M; q += LY) {
      long long o = q * Cpad + c;
      eb[o] = beta[o]; ev[o] = T.tv[o]; eg[o] = grad[o];
    }
    uturn_part<LY>(T.mb, T.mv, T.pb, T.pv, c, M, Cpad, ly, d1, d2);
  }
  ysum2<LY>(d1, d2, sh);
  if (!done) return;
  if (s2) {
    double r = rng_uniform(rng, step, c, C, ucount++);
    if (r < n2 / ntot) {   // state = state1
      copy_vec<LY>(T.sb, cur_sb, c, M, Cpad, ly);
      copy_vec<LY>(T.sg, cur_sg, c, M, Cpad, ly);
      if (ly == 0) { T.sl[c] = csl; T.moved[c] = 1; }
    }
  }
  if (ly != 0) return;
  (dir < 0 ? T.ml : T.pl)[c] = ll;
  T.ntot[c] = ntot + n2;
  T.j[c] = j + 1;
  T.alast[c] = a2;
  T.nalast[c] = na2;
  T.ucount[c] = ucount;
  nleap[c] += 1ull;
  bool s = s2 && !(d1 < 0.0 || d2 < 0.0);
  if (s && j + 1 < kMaxDepth) {
    T.newdbl[c] = 1;
    atomicAdd(T.n_active, 1);
  } else {
    T.active[c] = 0;
    if (round >= 0) T.end_round[c] = round;   // asynchronous iterations: this chain ends its iteration in this round
  }
}

// dual averaging of epsilon (:294-302), addToRes! (:305), misc (:307-308), state0 = state (:310)
template <int LY>
__global__ void __launch_bounds__(256) k_nuts_end(StepSrc steps, int round, long long total_steps, long long C, long long M,
                                                     long long Cpad, double* b0, double* g0, double* l0, Nuts T, Recorder R,
                                                     double* eps_out, double* jmax_out) {
  const Ln L = lanes<LY>(C);
  const long long c = L.c;
  if (!L.in) return;
  if (round >= 0 && T.end_round[c] != round) return;     // asynchronous iterations: only the chains whose tree was completed now
  const long long step = steps.of(c, true);
  const int ly = L.ly;
  copy_vec<LY>(b0, T.sb, c, M, Cpad, ly);
  copy_vec<LY>(g0, T.sg, c, M, Cpad, ly);
  const double sl = T.sl[c];
  record<LY>(R, step, c, M, Cpad, b0, sl, T.moved[c] != 0, ly);
  if (ly != 0) return;
  const double delta = 0.7, gam = 0.05, kappa = 0.75, t0 = 10.0;
  const long long nadapt = 1000;
  double i1 = (double)(step + 1);
  double eps;
  if (step + 1 <= nadapt) {
    double mu = log(10.0 * 1.0);
    double hbar = T.hbar[c] * (1.0 - 1.0 / (i1 + t0)) + (delta - T.alast[c] / T.nalast[c]) / (i1 + t0);
    double le = mu - sqrt(i1) / gam * hbar;
    double w = pow(i1, -kappa);
    T.lebar[c] = w * le + (1.0 - w) * T.lebar[c];
    T.hbar[c] = hbar;
    eps = exp(le);
  } else {
    eps = exp(T.lebar[c]);
  }
  T.eps[c] = eps;
  l0[c] = sl;
  if (eps_out) eps_out[step * Cpad + c] = eps;
  if (jmax_out) jmax_out[step * Cpad + c] = (double)T.j[c];
  if (round >= 0) {                                      // on to the chain's next iteration, in the next round
    T.istep[c] = step + 1;
    if (step + 1 < total_steps) {
      T.begin_round[c] = round + 1;
      atomicAdd(T.n_active, 1);
    }
  }
}

__global__ void k_sum_u64(const unsigned long long* v, long long n, unsigned long long* out) {   // one block
  __shared__ unsigned long long sh[256];
  unsigned long long t = 0ull;
  for (long long i = threadIdx.x; i < n; i += 256) t += v[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) *out = sh[0];
}
__global__ void k_nuts_async_init(long long C, long long first_step, Nuts T) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  T.istep[c] = first_step;
  T.begin_round[c] = 0;
  T.end_round[c] = -1;
  T.active[c] = 0;
}

}  // namespace

// ---- device-side round loop (a CUDA graph WHILE node): one graph launch per iteration ---------------------------------------
// ctl[0] = rounds of this iteration, ctl[1] = rounds of the whole run (host statistics)
__global__ void k_nuts_loop_begin(int* n_active, unsigned long long* ctl) {
  *n_active = 0;
  ctl[0] = 0ull;
}
// after k_nuts_post: another round while some chain still builds its tree (src/samplers.jl:263: `while s == 1`, lockstep over
// the chains) and the depth cap has not been reached; clears the counter for that round
__global__ void k_nuts_loop_cond(cudaGraphConditionalHandle h, int* n_active, unsigned long long* ctl, int max_rounds) {
  const unsigned long long rounds = ctl[0] + 1ull;
  ctl[0] = rounds;
  ctl[1] += 1ull;
  const bool more = *n_active > 0 && rounds < (unsigned long long)max_rounds;
  *n_active = 0;
  cudaGraphSetConditional(h, more ? 1u : 0u);
}

extern "C" int smc_nuts_run(smc_model_t m, const smc_sampler_config* cfg, const double* init, int64_t C, smc_run_output* out) {
  if (!m || !cfg || !init || !out) { smc_set_error("smc_nuts_run: NULL argument"); return SMC_ERR_INVALID; }
  Run r;
  r.m = m; r.cfg = cfg; r.out = out; r.C = C;
  // worst case uniforms per step: 1 slice + per doubling (1 direction + merges + 1 accept) = 1 + 8*2 + 247
  const long long max_uniforms = 1 + 2 * kMaxDepth + ((1 << kMaxDepth) - 1 - kMaxDepth);
  SMC_TRY(run_setup(r, init, true, max_uniforms));
  cudaStream_t st = m->ctx->stream;
  long long M = r.M, Cpad = r.Cpad, V = M * Cpad;
  Nuts T;
  double** vecs[] = {&T.mb, &T.mv, &T.mg, &T.pb, &T.pv, &T.pg, &T.sb, &T.sg, &T.tb, &T.tv, &T.tg, &T.cnb, &T.cnv, &T.csb, &T.csg};
  for (auto pp : vecs) SMC_TRY(r.A.alloc(pp, V));
  double** stacks[] = {&T.knb, &T.knv, &T.ksb, &T.ksg};
  for (auto pp : stacks) SMC_TRY(r.A.alloc(pp, V * kMaxDepth));
  double** sstacks[] = {&T.ksl, &T.kn, &T.ka, &T.kna};
  for (auto pp : sstacks) SMC_TRY(r.A.alloc(pp, Cpad * kMaxDepth));
  double** scal[] = {&T.ml, &T.pl, &T.sl, &T.csl, &T.tl, &T.u, &T.H0, &T.eps, &T.hbar, &T.lebar, &T.alast, &T.nalast, &T.ntot};
  for (auto pp : scal) SMC_TRY(r.A.alloc(pp, Cpad));
  int** ints[] = {&T.j, &T.leaf, &T.dir, &T.active, &T.newdbl, &T.ucount, &T.moved, &T.slot};
  for (auto pp : ints) SMC_TRY(r.A.alloc(pp, Cpad));
  const int slot_blocks = (int)((C + 255) / 256);
  const char* sp1 = getenv("SMC_NUTS_SLOT_PACK");   // =0: the three-kernel count / scan / assign at every chain count
  const bool slot_pack1 = C <= 65536 && !(sp1 && sp1[0] == '0');
  SMC_TRY(r.A.alloc(&T.blocksum, slot_blocks + 1));
  SMC_TRY(r.A.alloc(&T.n_active, 4));
  SMC_TRY(r.A.alloc(&T.rounds_active, 4));
  SMC_CUDA(cudaMemsetAsync(T.rounds_active, 0, 4 * sizeof(unsigned long long), st));
  double *eps_dev = nullptr, *jmax_dev = nullptr;
  if (out->epsilon) SMC_TRY(r.A.alloc(&eps_dev, cfg->steps * Cpad));
  if (out->jmax) SMC_TRY(r.A.alloc(&jmax_dev, cfg->steps * Cpad));
  int ly = lanes_for(C);
  if (const char* e = getenv("SMC_NUTS_LY")) ly = (e[0] == '1') ? 1 : 8;      // (A/B measurements)
  int post_cx = 32;
  if (const char* e = getenv("SMC_NUTS_POST_CX")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) post_cx = v; }
  k_nuts_init<<<gridc(C), 128, 0, st>>>(C, T);
  m->ctx->launches++;
  long long* d_step;
  SMC_TRY(r.A.alloc(&d_step, 1));
  SMC_CUDA(cudaMemsetAsync(d_step, 0, sizeof(long long), st));
  unsigned long long* d_ctl;
  SMC_TRY(r.A.alloc(&d_ctl, 2));
  SMC_CUDA(cudaMemsetAsync(d_ctl, 0, 2 * sizeof(unsigned long long), st));
  const StepSrc lock = {d_step, nullptr};
  int64_t evals = 1;
  // Every round ends with "is any chain still building its tree?".  Waiting for that answer before issuing the next
  // round leaves the GPU idle for a host round trip per leapfrog, so when the previous iteration's trees were long the
  // next round is issued before the answer is known (one round in flight ahead): if the answer turns out to be "none",
  // the extra round finds every chain inactive and changes nothing -- at most one wasted evaluation per iteration.
  struct RoundTrip {           // released on every path out of this function
    int* h_active = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~RoundTrip() {
      for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
      if (h_active) cudaFreeHost(h_active);
    }
  } rt;
  SMC_CUDA(cudaMallocHost(&rt.h_active, 2 * sizeof(int)));
  SMC_CUDA(cudaEventCreateWithFlags(&rt.ev[0], cudaEventDisableTiming));
  SMC_CUDA(cudaEventCreateWithFlags(&rt.ev[1], cudaEventDisableTiming));
  int* const h_active = rt.h_active;
  cudaEvent_t* const ev = rt.ev;
  int rc = SMC_OK;
  const int max_rounds = 1 << kMaxDepth;
  int prev_rounds = 0;
  const char* nc = getenv("SMC_NO_COMPACT");
  // (below a few thousand chains an evaluation is latency bound -- its cost does not shrink with the chain count -- and
  // the three slot kernels per round would only add to it; SMC_COMPACT=1 forces packing for tests)
  const char* fc = getenv("SMC_COMPACT");
  const bool compact = !(nc && nc[0] == '1') && !(m->row_shard && m->ctx->nranks > 1) && (C >= 4096 || (fc && fc[0] == '1'));
  long long eval_chains = 0;   // chain evaluations actually performed (sum of the evaluated widths)
  // Opt-in (SMC_NUTS_DEVICE_LOOP=1, below the packing threshold): the round loop on the device.  From the third iteration on an
  // iteration is ONE graph launch -- begin, a WHILE node whose body is a round (pre, evaluation, post, condition) and whose
  // condition a kernel sets from the active-chain counter, end -- so no round waits for the host to read that counter.
  // Bit-identical to the host loop (tests/test_gpu_samplers.py), but not faster on this driver: the WHILE node costs about
  // what the host round trip with its one-round speculation costs (sleepstudy, 1024 chains: 5.8e6 against 6.6e6
  // leapfrogs/s; hierarchical: 2.3e6 against 2.6e6; OU with T = 1e6, one chain: 1.44e4 against 1.33e4 evaluations/s;
  // profiles/nuts_device_loop_r02.json), and a child graph may not sit in a conditional body (tools/cond_graph_probe.cu), so
  // the evaluation is captured kernel by kernel.  With active-chain packing (>= 4096 chains) the evaluated width changes
  // between rounds, which a captured body cannot follow.  Hence the host loop stays the default.
  const char* dl = getenv("SMC_NUTS_DEVICE_LOOP");
  const bool dev_loop = (dl && dl[0] == '1') && !compact && m->graphs_enabled && smc_comm_graph_ok(m) && cfg->steps > 3;
  cudaGraphExec_t loop_exec = nullptr;
  int64_t launches_fixed = 0, launches_round = 0;
  auto build_loop_graph = [&]() -> bool {
    cudaGraph_t G = nullptr;
    if (cudaGraphCreate(&G, 0) != cudaSuccess) return false;
    bool ok = true;
    const int64_t l0 = m->ctx->launches, epoch = m->realloc_epoch;
    // (1) head: clear the counters, begin
    cudaStreamCaptureStatus cs;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    cudaGraph_t gcap = nullptr;
    ok = ok && cudaStreamBeginCaptureToGraph(st, G, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (!ok) { cudaGetLastError(); cudaGraphDestroy(G); return false; }
    k_nuts_loop_begin<<<1, 1, 0, st>>>(T.n_active, d_ctl);
    LAUNCH_LY(k_nuts_begin, r.rng, lock, -1, C, M, Cpad, r.b0, r.g0, r.l0, T);
    ok = ok && cudaStreamGetCaptureInfo(st, &cs, nullptr, &gcap, &deps, &ndeps) == cudaSuccess;
    std::vector<cudaGraphNode_t> head(deps, deps + (ok ? ndeps : 0));
    ok = (cudaStreamEndCapture(st, &gcap) == cudaSuccess) && ok;
    // (2) the WHILE node (the first round always runs: default value 1 at every launch)
    cudaGraphConditionalHandle h;
    cudaGraphNode_t wnode = nullptr;
    cudaGraph_t body = nullptr;
    if (ok) ok = cudaGraphConditionalHandleCreate(&h, G, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) {
      cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
      p.conditional.handle = h;
      p.conditional.type = cudaGraphCondTypeWhile;
      p.conditional.size = 1;
      ok = cudaGraphAddNode(&wnode, G, head.data(), head.size(), &p) == cudaSuccess;
      if (ok) body = p.conditional.phGraph_out[0];
    }
    // (3) body: one round, the evaluation captured inline (its own kernels, not its graph)
    const int64_t l1 = m->ctx->launches;
    if (ok) ok = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      LAUNCH_LY(k_nuts_pre, r.rng, lock, C, M, Cpad, m->beta, 0, T);
      const int e = smc_model_eval_inline(m, C, true);
      LAUNCH_POST(post_cx, r.rng, lock, -1, C, M, Cpad, m->grad, m->logp, 0, T, r.nleap);
      k_nuts_loop_cond<<<1, 1, 0, st>>>(h, T.n_active, d_ctl, max_rounds);
      m->ctx->launches += 3;
      ok = (cudaStreamEndCapture(st, &gcap) == cudaSuccess) && e == SMC_OK;
    }
    launches_round = m->ctx->launches - l1;
    // (4) tail: end of the iteration, next step
    if (ok) ok = cudaStreamBeginCaptureToGraph(st, G, &wnode, nullptr, 1, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      LAUNCH_LY(k_nuts_end, lock, -1, (long long)cfg->steps, C, M, Cpad, r.b0, r.g0, r.l0, T, r.R, eps_dev, jmax_dev);
      k_step_inc<<<1, 1, 0, st>>>(d_step);
      ok = cudaStreamEndCapture(st, &gcap) == cudaSuccess;
    }
    launches_fixed = 4;
    m->ctx->launches = l0;
    ok = ok && epoch == m->realloc_epoch;
    if (ok) ok = cudaGraphInstantiate(&loop_exec, G, 0) == cudaSuccess;
    cudaGraphDestroy(G);
    if (!ok) { cudaGetLastError(); loop_exec = nullptr; }
    return ok;
  };
  // ---- asynchronous iterations ---------------------------------------------------------------------------------------------
  // In lockstep a chain whose tree is complete idles until the longest tree of the iteration is: 45 % (sleepstudy) to 73 %
  // (hierarchical, 2000 iterations) of the lanes at 16 384 chains.  Chains are independent (src/samplers.jl:199-310 is a loop
  // over ONE chain), so nothing requires them to be in the same iteration: here every chain carries its own iteration number
  // (RNG streams, the recorder, the step-size adaptation and the output positions are all keyed by it), a chain whose tree
  // completes in round r ends its iteration in that round (k_nuts_end, masked) and begins the next one in round r + 1
  // (k_nuts_begin, masked), and the rest of the run is one sequence of rounds until no chain has work left.  Every chain
  // executes exactly the arithmetic it would in lockstep: draws, step sizes, tree depths are bit-identical
  // (tests/test_gpu_samplers.py), in roughly half the rounds.  But a round costs 1.5-1.9 times as much -- the chains of a block
  // are at different leaves, so the merge loop of k_nuts_post runs to the deepest merge of any of them, the two masked kernels
  // run every round, and lockstep's late rounds were already cheap through active-chain packing -- so it only pays when
  // lockstep leaves most lanes idle (BASELINE size, 16 384 chains x 2000 iterations, profiles/nuts_baseline_size_r02.jsonl:
  // hierarchical 1.78e7 -> 2.45e7 leapfrogs/s with 0.27 of the lanes active in lockstep; sleepstudy 4.31e7 -> 3.97e7 with
  // 0.54).  Policy: the run starts in lockstep; at iterations 48, 128, 384 and 1100 the active-lane fraction measured since the
  // previous mark decides (< 0.36: the remaining iterations run asynchronously).  Values cannot depend on the decision.
  // SMC_NUTS_ASYNC=1 switches at iteration 0, SMC_NUTS_ASYNC=0 never.
  const char* as = getenv("SMC_NUTS_ASYNC");
  const bool async_ok = C >= 32 && !dev_loop && !(as && as[0] == '0');
  const bool async_forced = async_ok && as && as[0] == '1';
  auto run_async = [&](long long first_step) -> int {
    SMC_TRY(r.A.alloc(&T.istep, Cpad));
    SMC_TRY(r.A.alloc(&T.end_round, Cpad));
    SMC_TRY(r.A.alloc(&T.begin_round, Cpad));
    k_nuts_async_init<<<gridc(C), 128, 0, st>>>(C, first_step, T);
    m->ctx->launches++;
    const StepSrc own = {nullptr, T.istep};
    const long long bucket = std::max<long long>(32, ((C + 15) / 16 + 31) / 32 * 32);
    long long active_ub = C;
    int launched = 0, rounds = 0;
    const long long round_cap = (long long)(cfg->steps - first_step) * max_rounds + 2;
    int arc = SMC_OK;
    auto launch_round = [&]() -> int {
      const int slot = launched & 1, rnd = launched;
      long long Ceval = C;
      if (compact && active_ub * 3 <= C * 2) Ceval = std::min<long long>(C, (std::max<long long>(active_ub, 1) + bucket - 1) / bucket * bucket);
      const int use_slots = Ceval < C ? 1 : 0;
      LAUNCH_LY(k_nuts_begin, r.rng, own, rnd, C, M, Cpad, r.b0, r.g0, r.l0, T);
      if (use_slots) {
        if (slot_pack1) {
          k_slot_pack<<<1, 1024, 0, st>>>(T.active, C, T.slot);
          m->ctx->launches += 1;
        } else {
          k_slot_count<<<slot_blocks, 256, 0, st>>>(T.active, C, T.blocksum);
          k_slot_scan<<<1, 1024, 0, st>>>(T.blocksum, slot_blocks);
          k_slot_assign<<<slot_blocks, 256, 0, st>>>(T.active, C, T.blocksum, T.slot);
          m->ctx->launches += 3;
        }
      }
      LAUNCH_LY(k_nuts_pre, r.rng, own, C, M, Cpad, m->beta, use_slots, T);
      int e = smc_model_eval_device(m, Ceval, true);
      if (e != SMC_OK) return e;
      evals++;
      LAUNCH_POST(post_cx, r.rng, own, rnd, C, M, Cpad, m->grad, m->logp, use_slots, T, r.nleap);
      LAUNCH_LY(k_nuts_end, own, rnd, (long long)cfg->steps, C, M, Cpad, r.b0, r.g0, r.l0, T, r.R, eps_dev, jmax_dev);
      m->ctx->launches += 4;
      cudaMemcpyAsync(h_active + slot, T.n_active, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaEventRecord(ev[slot], st);
      launched++;
      return SMC_OK;
    };
    arc = launch_round();
    while (arc == SMC_OK) {
      const int slot = rounds & 1;
      if (launched == rounds + 1 && launched < round_cap) {     // one round in flight ahead of the host's knowledge
        arc = launch_round();
        if (arc != SMC_OK) break;
      }
      cudaError_t e = cudaEventSynchronize(ev[slot]);
      if (e != cudaSuccess) { smc_set_error("CUDA error '%s' in the NUTS loop", cudaGetErrorString(e)); arc = SMC_ERR_CUDA; break; }
      rounds++;
      if (h_active[slot] == 0 || rounds >= round_cap) break;
      active_ub = h_active[slot];
    }
    cudaStreamSynchronize(st);
    return arc;
  };
  auto finish = [&]() -> int {
    SMC_TRY(run_finish(r, evals));
    SMC_TRY(copy_out_blocks(r, eps_dev, out->epsilon, cfg->steps, 1));
    SMC_TRY(copy_out_blocks(r, jmax_dev, out->jmax, cfg->steps, 1));
    return SMC_OK;
  };
  if (async_forced) {
    SMC_TRY(run_async(0));
    return finish();
  }
  long long probe_rounds = 0, mark_rounds = 0;
  unsigned long long mark_leaps = 0ull;
  bool loop_failed = false;
  for (long long i = 0; i < cfg->steps && rc == SMC_OK; i++) {
    if (dev_loop && !loop_failed && i >= 2) {
      if (!loop_exec && !build_loop_graph()) loop_failed = true;
      if (loop_exec) {
        if (cudaGraphLaunch(loop_exec, st) != cudaSuccess) { smc_set_error("cudaGraphLaunch failed in the NUTS loop"); rc = SMC_ERR_CUDA; break; }
        continue;
      }
    }
    LAUNCH_LY(k_nuts_begin, r.rng, lock, -1, C, M, Cpad, r.b0, r.g0, r.l0, T);
    m->ctx->launches++;
    int launched = 0, rounds = 0;
    // Lockstep evaluation only has to cover the chains that are still building a tree.  Their number never grows within
    // an iteration, so the last count the host has seen is an upper bound: once it drops below ~2/3 of the chains, the
    // active ones are given dense evaluation columns (slot[c]) and the tape is evaluated for that many (rounded up to a
    // bucket so that only a few chain counts, i.e. evaluation graphs, occur).
    long long active_ub = C;
    const long long bucket = std::max<long long>(32, ((C + 15) / 16 + 31) / 32 * 32);
    auto launch_round = [&]() -> int {
      const int slot = launched & 1;
      long long Ceval = C;
      if (compact && active_ub * 3 <= C * 2) Ceval = std::min<long long>(C, (std::max<long long>(active_ub, 1) + bucket - 1) / bucket * bucket);
      const int use_slots = Ceval < C ? 1 : 0;
      if (use_slots) {
        if (slot_pack1) {
          k_slot_pack<<<1, 1024, 0, st>>>(T.active, C, T.slot);
          m->ctx->launches += 1;
        } else {
          k_slot_count<<<slot_blocks, 256, 0, st>>>(T.active, C, T.blocksum);
          k_slot_scan<<<1, 1024, 0, st>>>(T.blocksum, slot_blocks);
          k_slot_assign<<<slot_blocks, 256, 0, st>>>(T.active, C, T.blocksum, T.slot);
          m->ctx->launches += 3;
        }
      }
      LAUNCH_LY(k_nuts_pre, r.rng, lock, C, M, Cpad, m->beta, use_slots, T);
      m->ctx->launches++;
      int e = smc_model_eval_device(m, Ceval, true);
      if (e != SMC_OK) return e;
      evals++;
      eval_chains += Ceval;
      (void)eval_chains;
      LAUNCH_POST(post_cx, r.rng, lock, -1, C, M, Cpad, m->grad, m->logp, use_slots, T, r.nleap);
      m->ctx->launches++;
      cudaMemcpyAsync(h_active + slot, T.n_active, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaEventRecord(ev[slot], st);
      launched++;
      return SMC_OK;
    };
    const bool speculate = prev_rounds >= 16 && !cfg->inject_uniforms;
    rc = launch_round();
    while (rc == SMC_OK) {
      const int slot = rounds & 1;
      if (speculate && launched == rounds + 1 && launched < max_rounds) {
        rc = launch_round();
        if (rc != SMC_OK) break;
      }
      cudaError_t e = cudaEventSynchronize(ev[slot]);
      if (e != cudaSuccess) { smc_set_error("CUDA error '%s' in the NUTS loop", cudaGetErrorString(e)); rc = SMC_ERR_CUDA; break; }
      rounds++;
      if (h_active[slot] == 0 || rounds >= max_rounds) break;
      active_ub = h_active[slot];
      if (launched == rounds) rc = launch_round();
    }
    if (rc != SMC_OK) break;
    prev_rounds = rounds;
    LAUNCH_LY(k_nuts_end, lock, -1, (long long)cfg->steps, C, M, Cpad, r.b0, r.g0, r.l0, T, r.R, eps_dev, jmax_dev);
    k_step_inc<<<1, 1, 0, st>>>(d_step);
    m->ctx->launches += 2;
    probe_rounds += rounds;
    // checkpoints: the active-lane fraction of the window since the previous mark (leapfrogs against rounds x chains,
    // speculative rounds included: they are part of lockstep's cost).  The first iterations (epsilon = 1, short cut trees)
    // say nothing about the run, so the first mark only starts a window.
    bool mark = false;
    for (long long cp : {12LL, 48LL, 128LL, 384LL, 1100LL}) mark = mark || (i + 1 == cp);
    if (async_ok && !cfg->inject_uniforms && mark && cfg->steps - (i + 1) >= 32) {
      unsigned long long nl = 0ull;
      k_sum_u64<<<1, 256, 0, st>>>(r.nleap, C, d_ctl);
      m->ctx->launches++;
      SMC_CUDA(smc_copy_sync(&nl, d_ctl, sizeof(nl), cudaMemcpyDeviceToHost, st));
      const double frac = (double)(nl - mark_leaps) / ((double)std::max<long long>(probe_rounds - mark_rounds, 1) * (double)C);
      const bool first = (i + 1 == 12);
      mark_leaps = nl;
      mark_rounds = probe_rounds;
      if (!first && frac < 0.36) {
        rc = run_async(i + 1);
        if (rc != SMC_OK) return rc;
        return finish();
      }
    }
  }
  cudaStreamSynchronize(st);
  if (loop_exec) {      // rounds run by the device loop: evaluations and launches for the statistics
    unsigned long long ctl[2] = {0ull, 0ull};
    cudaMemcpy(ctl, d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost);
    evals += (int64_t)ctl[1];
    long long graph_iters = 0;
    cudaMemcpy(&graph_iters, d_step, sizeof(long long), cudaMemcpyDeviceToHost);
    graph_iters -= 2;
    m->ctx->launches += launches_fixed * graph_iters + launches_round * (int64_t)ctl[1];
    cudaGraphExecDestroy(loop_exec);
  }
  if (rc != SMC_OK) return rc;
  SMC_TRY(run_finish(r, evals));
  SMC_TRY(copy_out_blocks(r, eps_dev, out->epsilon, cfg->steps, 1));
  SMC_TRY(copy_out_blocks(r, jmax_dev, out->jmax, cfg->steps, 1));
  return SMC_OK;
}
