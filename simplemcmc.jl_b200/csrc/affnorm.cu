// SMC_OP_AFFNORM: an affine residual into a Normal density, forward value and every adjoint sum in ONE launch.
//
//   L = sum_e logpdfNormal(mu, sigma, r_e),   r_e = (((+-t_0) +- t_1) +- ...),   t_j = D_j[e] * s_j  or  s_j
//
// with D_j parameter-independent arrays (or contiguous views of one: x[2:end] and x[1:end-1] are the SAME stream), and
// s_j, mu, sigma per-chain scalars.  This is the shape of examples/ornstein.jl:24-26 (BASELINE configs[4]):
//   resid = x[2:end] - x[1:end-1] * fac - mu * (1. - fac);   resid ~ Normal(0, sigma)
// The reference evaluates it as three vector temporaries forward (src/parsing.jl:123-183 unfolding) and five backward
// (src/diff.jl:29-45 for + - *, :94-99 for Normal); round 1 ran it through the register-program interpreter at 1-6 % of
// the roofline.  Here r_e is computed in registers with the reference's roundings (one per * and per +-, no contraction),
// and what the reverse sweep needs from the element loop are three kinds of sums, all linear in (r - mu):
//   S2 = sum (r - mu)^2,   S_j = sum (r - mu) D_j[e]  (array terms whose scalar is differentiated),   S1 = sum (r - mu)
// from which   L       = -S2 / (2 sigma^2) - n log sigma - n log sqrt(2 pi)                  (src/distribs.jl:23-28)
//              dL/dmu  =  S1 / sigma^2            dL/dsigma = (S2 / sigma^2 - n) / sigma     (src/diff.jl:94-99)
//              dL/ds_j = -sign_j S_j / sigma^2  (array term)   |   -sign_j S1 / sigma^2  (scalar term)
// Two kernels over the same argument block:
//   k_affnorm_stream<CT>  (C <= 31): thread = element, CT chains in registers, loads straight from global memory -- bound
//                          by HBM: 8 n bytes per launch when the arrays are views of one vector;
//   k_affnorm_chains<CT>  (C >= 32): thread = (chain lane, element phase), the arrays go through shared memory as broadcast
//                          reads -- bound by the FP64 pipe: 3 T + 2 (1 + sums) flop per element and chain; its per-block
//                          partial sums are combined by k_affnorm_finish (warp per chain over the blocks).
// The stream kernel's partial sums are combined by its last block to finish (ticket counter).  Both orders are fixed:
// deterministic.
#include <stdlib.h>
#include "exec_common.cuh"

namespace {

constexpr int AFF_MAX_T = 4;
constexpr int AFF_MAX_SUM = 6;       // S2, up to four S_j, S1
constexpr int AFF_STREAM_MAX_BLOCKS = 512;   // per chain group: the last block reads that many partials in one go

struct AffArgs {
  int T, nsum, s1_slot;
  int kind[AFF_MAX_T];               // 0: D * s, 1: s
  int slot[AFF_MAX_T];               // sum slot of an array term whose scalar has an adjoint, else -1
  double sgn[AFF_MAX_T];
  const double* D[AFF_MAX_T];
  Opd s[AFF_MAX_T], mu, sigma;
  long long n, C, cstride;           // cstride: chains covered by the grid (partial sums: [block][slot][cstride])
  double* part;
  unsigned int* counter;             // one ticket counter per chain group
  Dst L;
  int L_acc;
  Dst adj[2 + AFF_MAX_T];            // dL/dmu, dL/dsigma, dL/ds_j
  int has_adj[2 + AFF_MAX_T];
};

__device__ __forceinline__ double opd_scalar(const Opd& o, long long c) { return o.is_imm ? o.imm : o.p[c * o.sc]; }

// Compile-time shape of the residual (the element loop is issue bound: with the term kinds as run-time fields it was
// ~70 instructions per element, 13 of them FP64, 244 us for T = 1e6 x 64 chains; profiles/cfg5_launches_r02.json).
//   bits 0-1: number of terms - 1;  bits 2-9: two bits per term (0: array * scalar, 1: scalar, 2: +array, 3: -array);
//   bits 10-13: the term's sum S_j is needed (array * scalar terms);  bit 14: S1 is needed;  bit 15: mu is the literal 0.
// AFF_GENERIC keeps every decision at run time (any residual the lowering can emit).
constexpr unsigned AFF_GENERIC = 0xFFFFFFFFu;
constexpr unsigned aff_pat(int T, int k0, int k1, int k2, int k3, int need, int s1, int mu0) {
  return (unsigned)(T - 1) | (k0 << 2) | (k1 << 4) | (k2 << 6) | (k3 << 8) | (need << 10) | (s1 << 14) | (mu0 << 15);
}
constexpr unsigned AFF_P_OU = aff_pat(3, 2, 0, 1, 0, 0b0010, 1, 1);      // x[2:end] - x[1:end-1] * s - c      ~ Normal(0, sigma)
constexpr unsigned AFF_P_YXB = aff_pat(2, 2, 0, 0, 0, 0b0010, 0, 1);     // y - x * b                          ~ Normal(0, sigma)
constexpr unsigned AFF_P_XBY = aff_pat(2, 0, 3, 0, 0, 0b0001, 0, 1);     // x * b - y
constexpr unsigned AFF_P_YXBC = aff_pat(3, 2, 0, 0, 0, 0b0110, 0, 1);    // y - x1 * b1 - x2 * b2
template <unsigned PAT>
struct AffPat {
  static constexpr bool GEN = PAT == AFF_GENERIC;
  static constexpr int T = GEN ? AFF_MAX_T : (int)(PAT & 3u) + 1;
  __device__ static __forceinline__ bool is_array(const AffArgs& a, int j) {
    return GEN ? (j < a.T && a.kind[j] == 0) : (j < T && ((PAT >> (2 + 2 * j)) & 3u) != 1u);
  }
  __device__ static __forceinline__ bool live(const AffArgs& a, int j) { return GEN ? j < a.T : j < T; }
  // signed value of term j (sc = sign * scalar)
  __device__ static __forceinline__ double term(const AffArgs& a, int j, double dv, double sc) {
    if (GEN) return a.kind[j] == 0 ? __dmul_rn(dv, sc) : sc;
    const unsigned k = (PAT >> (2 + 2 * j)) & 3u;
    return k == 0u ? __dmul_rn(dv, sc) : (k == 1u ? sc : (k == 2u ? dv : -dv));
  }
  __device__ static __forceinline__ bool need(const AffArgs& a, int j) {
    return GEN ? (j < a.T && a.slot[j] >= 0) : (j < T && ((PAT >> (10 + j)) & 1u) != 0u);
  }
  __device__ static __forceinline__ bool need_s1(const AffArgs& a) { return GEN ? true : ((PAT >> 14) & 1u) != 0u; }
  __device__ static __forceinline__ bool mu_zero() { return GEN ? false : ((PAT >> 15) & 1u) != 0u; }
};
// r - mu of one element and one chain, and its contribution to the sums
template <unsigned PAT>
__device__ __forceinline__ void aff_element(const AffArgs& a, const double (&dv)[AFF_MAX_T], const double (&sc)[AFF_MAX_T], double mu,
                                            double (&acc)[AFF_MAX_SUM]) {
  typedef AffPat<PAT> P;
  double v = P::term(a, 0, dv[0], sc[0]);
#pragma unroll
  for (int j = 1; j < AFF_MAX_T; j++)
    if (P::live(a, j)) v = __dadd_rn(v, P::term(a, j, dv[j], sc[j]));
  const double d = P::mu_zero() ? v : __dsub_rn(v, mu);
  acc[0] = fma(d, d, acc[0]);
#pragma unroll
  for (int j = 0; j < AFF_MAX_T; j++)
    if (P::need(a, j)) acc[1 + j] = fma(d, dv[j], acc[1 + j]);
  if (P::need_s1(a)) acc[AFF_MAX_SUM - 1] += d;
}

// the last block of a chain group finishes: (S2, S_j, S1) -> L and the adjoint scalars of chain c
__device__ __forceinline__ void aff_finish_chain(const AffArgs& a, long long c, const double* S) {
  const double sigma = opd_scalar(a.sigma, c);
  const double inv_s2 = 1.0 / (sigma * sigma), nn = (double)a.n;
  double L = -0.5 * S[0] * inv_s2 - nn * log(sigma) - nn * SMC_LOG_SQRT_2PI;
  if (!(sigma > 0.0)) L = nan("");                                   // src/distribs.jl:25 asserts; here: chain out of support
  double* lp = a.L.p + c * a.L.sc;
  *lp = a.L_acc ? *lp + L : L;
  const double s1 = a.s1_slot >= 0 ? S[a.s1_slot] : 0.0;
  if (a.has_adj[0]) a.adj[0].p[c * a.adj[0].sc] = s1 * inv_s2;
  if (a.has_adj[1]) a.adj[1].p[c * a.adj[1].sc] = (S[0] * inv_s2 - nn) / sigma;
#pragma unroll
  for (int j = 0; j < AFF_MAX_T; j++)
    if (j < a.T && a.has_adj[2 + j]) {
      const double sj = a.kind[j] == 0 ? S[a.slot[j]] : s1;
      a.adj[2 + j].p[c * a.adj[2 + j].sc] = -a.sgn[j] * sj * inv_s2;
    }
}

__device__ __forceinline__ bool aff_take_ticket(unsigned int* counter, unsigned int nblocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == nblocks - 1u);
    if (s_last) *counter = 0u;                                        // ready for the next launch (graph replays)
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ---- few chains: thread = element, CT chains in registers ----------------------------------------------------------------
template <int CT, unsigned PAT>
__global__ void __launch_bounds__(256) k_affnorm_stream(AffArgs a) {
  typedef AffPat<PAT> P;
  constexpr int U = (CT <= 2) ? 8 : 4;                               // elements in flight per thread
  const int tid = threadIdx.x;
  const long long c0 = (long long)blockIdx.y * CT;
  double sc[CT][AFF_MAX_T], mu[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) {
    const long long ch = c0 + c < a.C ? c0 + c : a.C - 1;             // padding lanes repeat the last chain
    mu[c] = opd_scalar(a.mu, ch);
#pragma unroll
    for (int j = 0; j < AFF_MAX_T; j++) sc[c][j] = P::live(a, j) ? a.sgn[j] * opd_scalar(a.s[j], ch) : 0.0;
  }
  double acc[CT][AFF_MAX_SUM];
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int k = 0; k < AFF_MAX_SUM; k++) acc[c][k] = 0.0;

  const long long stride = (long long)gridDim.x * 256 * U;
  for (long long e0 = (long long)blockIdx.x * 256 * U + tid; e0 < a.n; e0 += stride) {
    double dv[U][AFF_MAX_T];
#pragma unroll
    for (int j = 0; j < AFF_MAX_T; j++) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long e = e0 + (long long)u * 256;
        dv[u][j] = (P::is_array(a, j) && e < a.n) ? __ldg(a.D[j] + e) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (e0 + (long long)u * 256 < a.n) {
#pragma unroll
        for (int c = 0; c < CT; c++) aff_element<PAT>(a, dv[u], sc[c], mu[c], acc[c]);
      }
    }
  }
  // block reduction in a fixed order: lanes (xor tree), then warps
  __shared__ double s_red[8][AFF_MAX_SUM * CT];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < AFF_MAX_SUM; k++)
#pragma unroll
    for (int c = 0; c < CT; c++) {
      double v = acc[c][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][k * CT + c] = v;
    }
  __syncthreads();
  // per-block partial, compacted to the slots in use: [block][slot][cstride]
  if (tid < a.nsum * CT) {
    const int slot = tid / CT, c = tid % CT;
    int k = slot;                                                     // slot -> accumulator row
    if (slot > 0) {
      k = AFF_MAX_SUM - 1;
      for (int j = 0; j < AFF_MAX_T; j++)
        if (j < a.T && a.slot[j] == slot) k = 1 + j;
    }
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += s_red[w][k * CT + c];
    a.part[((long long)blockIdx.x * a.nsum + slot) * a.cstride + c0 + c] = v;
  }
  if (!aff_take_ticket(a.counter + blockIdx.y, gridDim.x)) return;
  // last block of this chain group: warp w sums values w, w + 8, ...; lane l takes blocks l, l + 32, ... (at most
  // AFF_STREAM_MAX_BLOCKS / 32 of them, all loads in flight at once), then a fixed xor tree over the lanes
  __shared__ double s_tot[AFF_MAX_SUM * CT];
  for (int v = warp; v < a.nsum * CT; v += 8) {
    const int slot = v / CT, c = v % CT;
    double t[AFF_STREAM_MAX_BLOCKS / 32];
#pragma unroll
    for (int k = 0; k < AFF_STREAM_MAX_BLOCKS / 32; k++) {
      const unsigned int b = lane + 32u * k;
      t[k] = b < gridDim.x ? __ldcg(a.part + ((long long)b * a.nsum + slot) * a.cstride + c0 + c) : 0.0;
    }
    double tt = 0.0;
#pragma unroll
    for (int k = 0; k < AFF_STREAM_MAX_BLOCKS / 32; k++) tt += t[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
    if (lane == 0) s_tot[v] = tt;
  }
  __syncthreads();
  if (tid < CT && c0 + tid < a.C) {
    double S[AFF_MAX_SUM];
#pragma unroll
    for (int k = 0; k < AFF_MAX_SUM; k++) S[k] = k < a.nsum ? s_tot[k * CT + tid] : 0.0;
    aff_finish_chain(a, c0 + tid, S);
  }
}

// ---- many chains: thread = (chain lane, element phase), CT chains per thread, arrays broadcast from shared memory ---------
constexpr int AFF_TE = 1024;         // elements per shared-memory tile
constexpr int AFF_NT = 128;          // threads per block = chain lanes x element phases

template <int CT, unsigned PAT>
__global__ void __launch_bounds__(AFF_NT) k_affnorm_chains(AffArgs a, int nlanes) {
  typedef AffPat<PAT> P;
  extern __shared__ double s_tile[];                                 // [array terms][AFF_TE]
  const int tid = threadIdx.x, lanec = tid % nlanes, phase = tid / nlanes, nphase = AFF_NT / nlanes;
  const long long cb = ((long long)blockIdx.y * nlanes + lanec) * CT;
  double sc[CT][AFF_MAX_T], mu[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) {
    const long long ch = cb + c < a.C ? cb + c : a.C - 1;
    mu[c] = opd_scalar(a.mu, ch);
#pragma unroll
    for (int j = 0; j < AFF_MAX_T; j++) sc[c][j] = P::live(a, j) ? a.sgn[j] * opd_scalar(a.s[j], ch) : 0.0;
  }
  double acc[CT][AFF_MAX_SUM];
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int k = 0; k < AFF_MAX_SUM; k++) acc[c][k] = 0.0;
  int row[AFF_MAX_T];                                                // tile row of an array term
  int nrow = 0;
#pragma unroll
  for (int j = 0; j < AFF_MAX_T; j++) row[j] = P::is_array(a, j) ? nrow++ : 0;

  const long long per = (a.n + gridDim.x - 1) / gridDim.x;
  const long long e_begin = (long long)blockIdx.x * per, e_end = e_begin + per < a.n ? e_begin + per : a.n;
  for (long long te = e_begin; te < e_end; te += AFF_TE) {
    const int len = (int)(e_end - te < AFF_TE ? e_end - te : AFF_TE);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < AFF_MAX_T; j++)
      if (P::is_array(a, j)) {
#pragma unroll
        for (int i0 = 0; i0 < AFF_TE; i0 += AFF_NT) {
          const int i = i0 + tid;
          s_tile[row[j] * AFF_TE + i] = i < len ? __ldg(a.D[j] + te + i) : 0.0;
        }
      }
    __syncthreads();
    // this thread's elements of the tile: phase, phase + nphase, ...; four at a time so that the loads and the four
    // dependent FP64 chains of consecutive elements overlap
    const int cnt = len > phase ? (len - phase + nphase - 1) / nphase : 0;
    const double* tp = s_tile + phase;
    int k = 0;
    for (; k + 4 <= cnt; k += 4) {
      double dv[4][AFF_MAX_T];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int j = 0; j < AFF_MAX_T; j++) dv[u][j] = P::is_array(a, j) ? tp[row[j] * AFF_TE + (k + u) * nphase] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int c = 0; c < CT; c++) aff_element<PAT>(a, dv[u], sc[c], mu[c], acc[c]);
    }
    for (; k < cnt; k++) {
      double dv[AFF_MAX_T];
#pragma unroll
      for (int j = 0; j < AFF_MAX_T; j++) dv[j] = P::is_array(a, j) ? tp[row[j] * AFF_TE + k * nphase] : 0.0;
#pragma unroll
      for (int c = 0; c < CT; c++) aff_element<PAT>(a, dv, sc[c], mu[c], acc[c]);
    }
  }
  // element phases of one chain lane are summed in phase order through shared memory, then one partial per block
  __syncthreads();
  double* s_ph = s_tile;                                              // [phase][AFF_MAX_SUM * CT][nlanes] <= 4 x 12 x 32 doubles
  if (nphase > 1) {
#pragma unroll
    for (int k = 0; k < AFF_MAX_SUM; k++)
#pragma unroll
      for (int c = 0; c < CT; c++) s_ph[((phase * AFF_MAX_SUM + k) * CT + c) * nlanes + lanec] = acc[c][k];
    __syncthreads();
    if (phase == 0) {
#pragma unroll
      for (int k = 0; k < AFF_MAX_SUM; k++)
#pragma unroll
        for (int c = 0; c < CT; c++)
          for (int ph = 1; ph < nphase; ph++) acc[c][k] += s_ph[((ph * AFF_MAX_SUM + k) * CT + c) * nlanes + lanec];
    }
  }
  if (phase == 0) {
#pragma unroll
    for (int c = 0; c < CT; c++) {
      double* p = a.part + (long long)blockIdx.x * a.nsum * a.cstride + cb + c;
      p[0] = acc[c][0];
#pragma unroll
      for (int j = 0; j < AFF_MAX_T; j++)
        if (j < a.T && a.slot[j] >= 0) p[(long long)a.slot[j] * a.cstride] = acc[c][1 + j];
      if (a.s1_slot >= 0) p[(long long)a.s1_slot * a.cstride] = acc[c][AFF_MAX_SUM - 1];
    }
  }
}

// finish of the chains kernel: block = 32 partial lanes x 8 chains; lane l sums blocks l, l + 32, ... in order, then a
// fixed xor tree over the lanes (deterministic), then lane 0 writes L and the adjoint scalars of its chain
__global__ void __launch_bounds__(256) k_affnorm_finish(AffArgs a, int nblocks) {
  const int lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= a.C) return;
  double S[AFF_MAX_SUM];
#pragma unroll
  for (int k = 0; k < AFF_MAX_SUM; k++) {
    double t = 0.0;
    if (k < a.nsum) {
      for (int b = lane; b < nblocks; b += 32) t += __ldcg(a.part + ((long long)b * a.nsum + k) * a.cstride + c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    S[k] = t;
  }
  if (lane == 0) aff_finish_chain(a, c, S);
}

}  // namespace

template <unsigned PAT>
static void launch_stream(const AffArgs& a, int CT, dim3 grid, cudaStream_t st) {
  if (CT == 1) k_affnorm_stream<1, PAT><<<grid, 256, 0, st>>>(a);
  else if (CT == 2) k_affnorm_stream<2, PAT><<<grid, 256, 0, st>>>(a);
  else if (CT == 4) k_affnorm_stream<4, PAT><<<grid, 256, 0, st>>>(a);
  else k_affnorm_stream<8, PAT><<<grid, 256, 0, st>>>(a);
}
template <unsigned PAT>
static void launch_chains(const AffArgs& a, int CT, int nlanes, dim3 grid, size_t smem, cudaStream_t st) {
  if (CT == 1) k_affnorm_chains<1, PAT><<<grid, AFF_NT, smem, st>>>(a, nlanes);
  else k_affnorm_chains<2, PAT><<<grid, AFF_NT, smem, st>>>(a, nlanes);
}
// the compile-time pattern of a statement, or AFF_GENERIC.  A compiled shape may accumulate MORE sums than the statement
// needs (the evaluation without a gradient -- simpleRWM -- needs none of the S_j): which partial sums are written and read is
// decided at run time by the slots, so a statement runs on the first shape with the same terms whose sums cover its own
// (without this the no-gradient tape of examples/ornstein.jl ran on the generic kernel: 0.234 ms against 0.058 ms for the
// evaluation with its gradient, T = 1e6 x 64 chains).
static unsigned aff_pattern_of(const smc_model_s* m, const smc_tape_stmt& s, const std::vector<int32_t>& d, const AffArgs& a) {
  int k[AFF_MAX_T] = {0, 0, 0, 0}, need = 0;
  for (int j = 0; j < a.T; j++) {
    const int32_t* t = &d[3 + 6 * j];
    const SmcReg& sr = m->regs[t[4]];
    if (t[0] == 1) k[j] = 1;
    else if (sr.d.kind == SMC_REG_IMM && sr.d.imm == 1.0 && t[5] < 0) k[j] = t[1] > 0 ? 2 : 3;
    else k[j] = 0;
    if (k[j] == 0 && t[5] >= 0) need |= 1 << j;
  }
  const SmcReg& mr = m->regs[s.src[0]];
  const int mu0 = (mr.d.kind == SMC_REG_IMM && mr.d.imm == 0.0) ? 1 : 0;
  const int s1 = a.s1_slot >= 0 ? 1 : 0;
  for (unsigned known : {AFF_P_OU, AFF_P_YXB, AFF_P_XBY, AFF_P_YXBC}) {
    const unsigned shape = aff_pat(a.T, k[0], k[1], k[2], k[3], 0, 0, mu0);
    const unsigned sums_mask = (0xFu << 10) | (1u << 14);
    if ((known & ~sums_mask) != shape) continue;
    const int kneed = (int)((known >> 10) & 0xFu), ks1 = (int)((known >> 14) & 1u);
    if ((need & ~kneed) == 0 && s1 <= ks1) return known;
  }
  return AFF_GENERIC;
}

// descriptor (an index table of the tape, see lowering.py::serialise):
//   [T, reg(dL/dmu) | -1, reg(dL/dsigma) | -1, then per term: kind, sign, array reg | -1, offset, scalar reg, reg(dL/ds) | -1]
int smc_aff_prepare(smc_model_s* m, SmcAffPlan& p) {
  const smc_tape_stmt& s = m->stmts[p.stmt];
  if (s.aux0 >= m->idx_host.size() || m->idx_host[s.aux0].size() < 3) { smc_set_error("tape: AFFNORM statement without a descriptor"); return SMC_ERR_INVALID; }
  const std::vector<int32_t>& d = m->idx_host[s.aux0];
  const int T = d[0];
  if (T < 1 || T > AFF_MAX_T || (int)d.size() != 3 + 6 * T || s.nsrc != 2) { smc_set_error("tape: malformed AFFNORM descriptor"); return SMC_ERR_INVALID; }
  const int64_t n = s.aux1;
  const int nregs = (int)m->regs.size();
  auto scalar_ok = [&](int r) { return r >= 0 && r < nregs && m->regs[r].numel == 1; };
  auto out_ok = [&](int r) { return r == -1 || (scalar_ok(r) && m->regs[r].d.kind == SMC_REG_CHAIN); };
  bool ok = scalar_ok((int)s.src[0]) && scalar_ok((int)s.src[1]) && out_ok(d[1]) && out_ok(d[2]) && n >= 1;
  for (int j = 0; ok && j < T; j++) {
    const int32_t* t = &d[3 + 6 * j];
    ok = (t[0] == 0 || t[0] == 1) && (t[1] == 1 || t[1] == -1) && scalar_ok(t[4]) && out_ok(t[5]);
    if (ok && t[0] == 0) {
      ok = t[2] >= 0 && t[2] < nregs && t[3] >= 0;
      if (ok) {
        const SmcReg& r = m->regs[t[2]];
        ok = (r.d.kind == SMC_REG_DATA || r.d.kind == SMC_REG_HDR) && (int64_t)t[3] + n <= r.numel;
      }
    }
  }
  if (!ok) { smc_set_error("tape: AFFNORM descriptor references a bad register"); return SMC_ERR_INVALID; }
  p.desc = d;
  if (!p.counter) {
    SMC_CUDA(smc_dev_alloc(&p.counter, 4096 * sizeof(unsigned int), m->ctx->stream));
    SMC_CUDA(cudaMemsetAsync(p.counter, 0, 4096 * sizeof(unsigned int), m->ctx->stream));
  }
  return SMC_OK;
}

void smc_aff_release(SmcAffPlan& p, cudaStream_t st) {
  if (p.counter) smc_dev_free(p.counter, st);
  p.counter = nullptr;
}

int smc_aff_launch(smc_model_s* m, SmcAffPlan& p, int64_t C) {
  const smc_tape_stmt& s = m->stmts[p.stmt];
  const std::vector<int32_t>& d = p.desc;
  cudaStream_t st = m->ctx->stream;
  AffArgs a;
  memset(&a, 0, sizeof(a));
  a.T = d[0];
  a.n = s.aux1;
  a.C = C;
  a.mu = smc_make_opd(m, s.src[0]);
  a.sigma = smc_make_opd(m, s.src[1]);
  a.L = smc_make_dst(m, s.dst);
  a.L_acc = (s.flags & SMC_F_ACC) ? 1 : 0;
  int nsum = 1, narr = 0;
  bool need_s1 = d[1] >= 0;
  for (int k = 0; k < 2; k++)
    if (d[1 + k] >= 0) { a.adj[k] = smc_make_dst(m, (uint32_t)d[1 + k]); a.has_adj[k] = 1; }
  for (int j = 0; j < a.T; j++) {
    const int32_t* t = &d[3 + 6 * j];
    a.kind[j] = t[0];
    a.sgn[j] = (double)t[1];
    a.s[j] = smc_make_opd(m, (uint32_t)t[4]);
    a.slot[j] = -1;
    if (t[0] == 0) {
      a.D[j] = m->regs[t[2]].dev + t[3];
      narr++;
      if (t[5] >= 0) a.slot[j] = nsum++;
    } else if (t[5] >= 0) need_s1 = true;
    if (t[5] >= 0) { a.adj[2 + j] = smc_make_dst(m, (uint32_t)t[5]); a.has_adj[2 + j] = 1; }
  }
  a.s1_slot = need_s1 ? nsum++ : -1;
  a.nsum = nsum;
  a.counter = p.counter;
  const int sms = m->ctx->sm_count;
  static int force_generic = -1;
  if (force_generic < 0) { const char* e = getenv("SMC_AFFNORM_GENERIC"); force_generic = (e && e[0] == '1') ? 1 : 0; }
  const unsigned pat = force_generic ? AFF_GENERIC : aff_pattern_of(m, s, d, a);
  if (m->time_glm) SMC_CUDA(cudaEventRecord(m->ev0, st));
  if (C < 32) {
    const int CT = C <= 1 ? 1 : (C <= 2 ? 2 : (C <= 4 ? 4 : 8));
    const int U = CT <= 2 ? 8 : 4;
    const long long gy = (C + CT - 1) / CT;
    long long gx = std::min<long long>((a.n + 256LL * U - 1) / (256LL * U), std::max<long long>(1, (long long)sms * 4 / gy));
    gx = std::min<long long>(gx, AFF_STREAM_MAX_BLOCKS);
    a.cstride = gy * CT;
    SMC_TRY(smc_scratch_reserve(m, gx * nsum * a.cstride));
    a.part = m->scratch;
    dim3 grid((unsigned)gx, (unsigned)gy);
    switch (pat) {
      case AFF_P_OU: launch_stream<AFF_P_OU>(a, CT, grid, st); break;
      case AFF_P_YXB: launch_stream<AFF_P_YXB>(a, CT, grid, st); break;
      case AFF_P_XBY: launch_stream<AFF_P_XBY>(a, CT, grid, st); break;
      case AFF_P_YXBC: launch_stream<AFF_P_YXBC>(a, CT, grid, st); break;
      default: launch_stream<AFF_GENERIC>(a, CT, grid, st);
    }
  } else {
    // chain lanes per block: 32, 64 or 128 (the other threads are element phases of the same chains); two chains per thread
    // once there are enough chains to fill the device that way
    const int CT = C >= 4096 ? 2 : 1;
    int nlanes = 32;
    while (nlanes < AFF_NT && (long long)nlanes * CT < C) nlanes <<= 1;
    const long long gy = (C + (long long)nlanes * CT - 1) / ((long long)nlanes * CT);
    // blocks of 128 threads at 40 registers: six per SM (three left the FP64 pipe 58 % busy with 18 % of the warp slots in use
    // and fixed-latency waits as the top stall at T = 1e6 x 64 chains, profiles/cfg5_c64_affnorm_ncu_r02.json)
    long long gx = std::max<long long>(1, std::min<long long>((a.n + AFF_TE - 1) / AFF_TE, ((long long)sms * 6 + gy - 1) / gy));
    a.cstride = gy * nlanes * CT;
    SMC_TRY(smc_scratch_reserve(m, gx * nsum * a.cstride));
    a.part = m->scratch;
    dim3 grid((unsigned)gx, (unsigned)gy);
    const size_t smem = (size_t)std::max(std::max(narr, 1) * AFF_TE, (AFF_NT / nlanes) * AFF_MAX_SUM * CT * nlanes) * sizeof(double);
    switch (pat) {
      case AFF_P_OU: launch_chains<AFF_P_OU>(a, CT, nlanes, grid, smem, st); break;
      case AFF_P_YXB: launch_chains<AFF_P_YXB>(a, CT, nlanes, grid, smem, st); break;
      case AFF_P_XBY: launch_chains<AFF_P_XBY>(a, CT, nlanes, grid, smem, st); break;
      case AFF_P_YXBC: launch_chains<AFF_P_YXBC>(a, CT, nlanes, grid, smem, st); break;
      default: launch_chains<AFF_GENERIC>(a, CT, nlanes, grid, smem, st);
    }
    k_affnorm_finish<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(a, (int)gx);
    m->ctx->launches++;
  }
  m->ctx->launches++;
  if (m->time_glm) {
    SMC_CUDA(cudaEventRecord(m->ev1, st));
    SMC_CUDA(cudaEventSynchronize(m->ev1));
    float ms = 0.f;
    SMC_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    p.last_ms += ms;
  }
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}
