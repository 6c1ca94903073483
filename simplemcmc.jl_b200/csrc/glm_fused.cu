// K1: fused GLM log-density + gradient for all chains in lockstep, FP64 tensor-core (DMMA) path.
//
//   eta[N x C] = X[N x M] * B[M x C]                (the generated `tmp = X * beta`)
//   (l, d)     = family(eta, y)   per (row, chain)   (logpdfNormal / logpdfBernoulli + their adjoint chain,
//                                                    /root/reference/src/distribs.jl:12-28, src/diff.jl:94-99,165-166)
//   L[C]       = sum_i l ;  G[M x C] = X' * d        (the generated `dbeta += transpose(X) * dtmp`, diff.jl:55)
//
// One pass over X per lockstep evaluation: a row tile of X is staged once in shared memory (cp.async, 3 stages)
// and used for both products.  Both contractions run on mma.sync.m8n8k4.f64 (DMMA); the accumulator
// fragment of the first product (chain x row) is, element for element, the A fragment of the second
// (chain x row) . (row x param) product, so eta / d never leave registers.  tcgen05 has no f64 kind.
//
// Grid: (chain tiles, row splits).  Each CTA writes a partial (L, G); a second kernel sums the row
// splits in fixed order (deterministic) and publishes into the tape registers.
#include <algorithm>
#include "smc_internal.h"
#include "dists.cuh"

int smc_comm_allreduce_device(smc_ctx_s* c, double* buf, int64_t n);  // smc_comm.cu

namespace {

struct GlmArgs {
  const double* X;     // [N][MP] row major, zero padded
  const double* y;     // [N] or nullptr (then y_scalar)
  const double* beta;  // [M][Cpad]
  double* part;        // [RS][MP+1][Cpad]
  int* bad;
  long long N, rows_per_split, C, Cpad;
  int M;
  double par, y_scalar;
  int sa_neg, sb_neg;  // NORMAL: z = sa*eta + sb*y ; BERNOULLI: u = sa*eta
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int FAM>
__device__ __forceinline__ void family_eval(double eta, double y, const GlmArgs& a, double inv_s, double c0, double& l,
                                            double& d) {
  if (FAM == SMC_GLM_NORMAL) {
    // z = x - mu ; r = z/sigma ; l = -r*r/2 - log(sigma) - log(sqrt(2pi))  (src/distribs.jl:23-28)
    double z = (a.sa_neg ? -eta : eta) + (a.sb_neg ? -y : y);
    double r = z * inv_s;
    l = -r * r * 0.5 + c0;
    double dz = -z * (inv_s * inv_s);  // (mu - x)/(sigma*sigma), src/diff.jl:99
    d = a.sa_neg ? -dz : dz;
  } else {
    // prob = 1/(1+exp(u)) ; l = y ? log(prob) : log(1-prob)   (examples/binomial.jl:15, src/distribs.jl:12-21)
    double u = a.sa_neg ? -eta : eta;
    double e = exp(u), t4 = 1.0 + e, p = 1.0 / t4;
    bool one = (y == 1.0);
    l = one ? log(p) : log(1.0 - p);
    double du = one ? -e * p : p;  // exp(u) * (-1/t4^2) * 1/(p-(1-y)), src/diff.jl:166,73,45
    d = a.sa_neg ? -du : du;
    if (!(y == 0.0 || y == 1.0)) l = nan("");
  }
}

template <int FAM, int MT, int NWC, int NWR, int MP8>
__global__ void __launch_bounds__(NWC* NWR * 32, 1) k_glm(GlmArgs a) {
  constexpr int MP = MP8 * 8, LD = MP + 4, NT = NWC * NWR * 32;
  constexpr int RT = (MP8 >= 4) ? 64 : 128;  // rows per tile
  constexpr int NST = 3;
  constexpr int TC = NWC * MT * 8;  // chains per CTA
  static_assert(RT % (8 * NWR) == 0, "row tile must split evenly over the row warps");
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;                  // [NST][RT][LD]
  double* ys = smem + NST * RT * LD;  // [NST][RT]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wc = warp % NWC, wr = warp / NWC;
  const long long cbase = (long long)blockIdx.x * TC + wc * (MT * 8);
  const long long row_begin = (long long)blockIdx.y * a.rows_per_split;
  const long long row_end = min(a.N, row_begin + a.rows_per_split);
  const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + RT - 1) / RT) : 0;

  // B fragments: bfr[mt][ks] = beta[j = ks*4+q][chain = cbase + mt*8 + g]
  double bfr[MT][MP / 4];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    long long c = cbase + mt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < MP / 4; ks++) {
      int j = ks * 4 + q;
      bfr[mt][ks] = (j < a.M && c < a.C) ? __ldg(a.beta + (long long)j * a.Cpad + c) : 0.0;
    }
  }
  double G[MT][MP8][2];
  double lsum[MT];
  int badmask = 0;
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    lsum[mt] = 0.0;
#pragma unroll
    for (int nt = 0; nt < MP8; nt++) G[mt][nt][0] = G[mt][nt][1] = 0.0;
  }
  double inv_s = 1.0, c0 = 0.0;
  if (FAM == SMC_GLM_NORMAL) {
    inv_s = 1.0 / a.par;
    c0 = -log(a.par) - SMC_LOG_SQRT_2PI;
  }

  auto load_tile = [&](int t) {
    if (t < ntiles) {
      double* xs = Xs + (t % NST) * RT * LD;
      double* yy = ys + (t % NST) * RT;
      long long r0 = row_begin + (long long)t * RT;
      constexpr int CH = MP / 2;  // 16-byte chunks per row
      for (int i = tid; i < RT * CH; i += NT) {
        int r = i / CH, ch = i - r * CH;
        long long row = r0 + r;
        bool ok = row < row_end;
        cp_async16(xs + r * LD + ch * 2, a.X + (ok ? row : 0) * MP + ch * 2, ok ? 16 : 0);
      }
      if (a.y)
        for (int r = tid; r < RT; r += NT) {
          long long row = r0 + r;
          bool ok = row < row_end;
          cp_async8(yy + r, a.y + (ok ? row : 0), ok ? 8 : 0);
        }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < NST - 1; s++) load_tile(s);

  for (int t = 0; t < ntiles; t++) {
    cp_async_wait<NST - 2>();
    __syncthreads();
    load_tile(t + NST - 1);
    const double* xs = Xs + (t % NST) * RT * LD;
    const double* yy = ys + (t % NST) * RT;
    const long long r0 = row_begin + (long long)t * RT;
#pragma unroll 1
    for (int b = wr; b < RT / 8; b += NWR) {
      const int rb = b * 8;
      if (r0 + rb >= row_end) break;
      double acc[MT][2];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
      const double* xrow = xs + (rb + g) * LD + q;
#pragma unroll
      for (int ks = 0; ks < MP / 4; ks++) {
        double xb = xrow[ks * 4];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], bfr[mt][ks], xb);
      }
      // acc[mt][s] = eta(chain cbase+mt*8+g, row rb+2q+s)
      double dd[MT][2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        int r = rb + 2 * q + s;
        bool valid = (r0 + r) < row_end;
        double y = a.y ? yy[r] : a.y_scalar;
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          double l, d;
          family_eval<FAM>(acc[mt][s], y, a, inv_s, c0, l, d);
          if (!valid) { l = 0.0; d = 0.0; }
          if (l == -INFINITY || isnan(l)) { badmask |= (1 << mt); l = 0.0; d = 0.0; }
          lsum[mt] += l;
          dd[mt][s] = d;
        }
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const double* xr2 = xs + (rb + 2 * q + s) * LD + g;
#pragma unroll
        for (int nt = 0; nt < MP8; nt++) {
          double xb = xr2[nt * 8];
#pragma unroll
          for (int mt = 0; mt < MT; mt++) dmma884(G[mt][nt][0], G[mt][nt][1], dd[mt][s], xb);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: reduce over the row warps in shared memory (fixed order), write the partial
  double* Gs = smem;             // [MP+1][TC]  (row MP = L)
  int* bs = (int*)(smem + (MP + 1) * TC);  // [TC]
  for (int i = tid; i < (MP + 1) * TC; i += NT) Gs[i] = 0.0;
  for (int i = tid; i < TC; i += NT) bs[i] = 0;
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    lsum[mt] += __shfl_xor_sync(0xffffffffu, lsum[mt], 1);
    lsum[mt] += __shfl_xor_sync(0xffffffffu, lsum[mt], 2);
  }
  badmask |= __shfl_xor_sync(0xffffffffu, badmask, 1);
  badmask |= __shfl_xor_sync(0xffffffffu, badmask, 2);
  for (int w = 0; w < NWR; w++) {
    if (wr == w) {
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        int cl = wc * (MT * 8) + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < MP8; nt++) {
          Gs[(nt * 8 + 2 * q) * TC + cl] += G[mt][nt][0];
          Gs[(nt * 8 + 2 * q + 1) * TC + cl] += G[mt][nt][1];
        }
        if (q == 0) {
          Gs[MP * TC + cl] += lsum[mt];
          if (badmask & (1 << mt)) bs[cl] = 1;
        }
      }
    }
    __syncthreads();
  }
  double* part = a.part + (long long)blockIdx.y * (MP + 1) * a.Cpad;
  const long long ct0 = (long long)blockIdx.x * TC;
  for (int i = tid; i < (MP + 1) * TC; i += NT) {
    int j = i / TC, cl = i - j * TC;
    long long c = ct0 + cl;
    if (c < a.C) part[(long long)j * a.Cpad + c] = Gs[i];
  }
  for (int cl = tid; cl < TC; cl += NT)
    if (bs[cl] && ct0 + cl < a.C) a.bad[ct0 + cl] = 1;
}

// out[j][c] = sum_rs part[rs][j][c]  (j = MP row is L); a chain flagged bad gets L = -inf so that the flag
// survives a cross-rank sum
__global__ void k_glm_reduce(const double* part, int RS, int MP, long long C, long long Cpad, const int* bad, double* out) {
  long long total = (long long)(MP + 1) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    double t = 0.0;
    for (int r = 0; r < RS; r++) t += part[((long long)r * (MP + 1) + j) * Cpad + c];
    if (j == MP && bad[c]) t = -INFINITY;
    out[j * Cpad + c] = t;
  }
}

// copy (L, G) into the tape registers; L = -inf/NaN marks the chain as out of support
__global__ void k_glm_publish(const double* out, int M, int MP, long long C, long long Cpad, int* bad, double* Ldst,
                              double* Gdst, int acc) {
  long long total = (long long)(M + 1) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    if (j == M) {
      double L = out[(long long)MP * Cpad + c];
      if (!(L > -INFINITY)) { bad[c] = 1; L = 0.0; }
      Ldst[c] = acc ? Ldst[c] + L : L;
    } else {
      Gdst[j * Cpad + c] = out[j * Cpad + c];
    }
  }
}

__global__ void k_to_row_major(const double* Xcm, double* Xrm, long long N, int M, int MP) {
  long long total = N * MP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / MP;
    int j = (int)(i - r * MP);
    Xrm[i] = j < M ? Xcm[(long long)j * N + r] : 0.0;
  }
}

template <int FAM, int MT, int NWC, int NWR, int MP8>
int launch_cfg(const GlmArgs& a, int RS, cudaStream_t st) {
  constexpr int MP = MP8 * 8, LD = MP + 4, RT = (MP8 >= 4) ? 64 : 128, NST = 3, TC = NWC * MT * 8;
  size_t pipe = (size_t)NST * RT * (LD + 1) * sizeof(double);
  size_t epi = (size_t)(MP + 1) * TC * sizeof(double) + TC * sizeof(int);
  size_t smem = std::max(pipe, epi);
  auto kern = k_glm<FAM, MT, NWC, NWR, MP8>;
  static bool attr_done = false;
  if (!attr_done) {
    SMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((unsigned)((a.C + TC - 1) / TC), (unsigned)RS);
  kern<<<grid, NWC * NWR * 32, smem, st>>>(a);
  return SMC_OK;
}

template <int FAM, int MP8>
int launch_mp(const GlmArgs& a, int tc, int RS, cudaStream_t st) {
  switch (tc) {
    case 128: return launch_cfg<FAM, 2, 8, 1, MP8>(a, RS, st);
    case 64: return launch_cfg<FAM, 2, 4, 2, MP8>(a, RS, st);
    case 32: return launch_cfg<FAM, 2, 2, 4, MP8>(a, RS, st);
    case 16: return launch_cfg<FAM, 2, 1, 8, MP8>(a, RS, st);
    default: return launch_cfg<FAM, 1, 1, 8, MP8>(a, RS, st);
  }
}

template <int FAM>
int launch_fam(const GlmArgs& a, int MP, int tc, int RS, cudaStream_t st) {
  switch (MP) {
    case 16: return launch_mp<FAM, 2>(a, tc, RS, st);
    case 32: return launch_mp<FAM, 4>(a, tc, RS, st);
    case 64: return launch_mp<FAM, 8>(a, tc, RS, st);
  }
  smc_set_error("GLM kernel: unsupported padded width %d", MP);
  return SMC_ERR_INVALID;
}

int chain_tile_for(long long C) {
  if (C > 64) return 128;
  if (C > 32) return 64;
  if (C > 16) return 32;
  if (C > 8) return 16;
  return 8;
}

// row splits: fill the SMs with whole waves while keeping >= 4 row tiles per CTA
int row_splits_for(long long chain_tiles, long long N, int RT, int sms) {
  long long max_rs = std::max<long long>(1, N / (4LL * RT));
  long long best = 1;
  double best_eff = 0.0;
  for (long long rs = 1; rs <= std::min<long long>(max_rs, 8LL * sms); rs++) {
    long long ctas = chain_tiles * rs;
    long long waves = (ctas + sms - 1) / sms;
    if (waves > 8) break;
    double eff = (double)ctas / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = rs; }
  }
  return (int)best;
}

}  // namespace

int smc_glm_prepare(smc_model_s* m, SmcGlmPlan& g) {
  const smc_tape_stmt& s = m->stmts[g.stmt];
  const SmcReg& X = m->regs[s.src[0]];
  const SmcReg& B = m->regs[s.src[1]];
  const SmcReg& Y = m->regs[s.src[2]];
  if (X.d.kind != SMC_REG_DATA && X.d.kind != SMC_REG_HDR) { smc_set_error("GLM: X must be parameter independent"); return SMC_ERR_INVALID; }
  g.family = (int)s.aux0;
  g.sign_neg = (int)s.aux1;
  g.par = s.imm;
  g.N = X.d.rows;
  g.M = (int)X.d.cols;
  if (g.M < 1 || g.M > 64 || (int64_t)B.numel != g.M) { smc_set_error("GLM: unsupported width M=%d (1..64)", g.M); return SMC_ERR_INVALID; }
  if (g.family != SMC_GLM_NORMAL && g.family != SMC_GLM_BERNOULLI) { smc_set_error("GLM: family %d is not implemented", g.family); return SMC_ERR_INVALID; }
  g.MP = g.M <= 16 ? 16 : (g.M <= 32 ? 32 : 64);
  cudaStream_t st = m->ctx->stream;
  SMC_CUDA(cudaMalloc(&g.Xrm, (size_t)g.N * g.MP * sizeof(double)));
  k_to_row_major<<<(unsigned)std::min<long long>((g.N * g.MP + 255) / 256, 148LL * 64), 256, 0, st>>>(X.dev, g.Xrm, g.N, g.M, g.MP);
  m->ctx->launches++;
  if (Y.d.kind == SMC_REG_IMM) { g.y = nullptr; g.y_scalar = Y.d.imm; }
  else if (Y.numel == 1) {
    SMC_CUDA(cudaMemcpyAsync(&g.y_scalar, Y.dev, sizeof(double), cudaMemcpyDeviceToHost, st));
    SMC_CUDA(cudaStreamSynchronize(st));
    g.y = nullptr;
  } else {
    if (Y.numel < g.N) { smc_set_error("GLM: response has %lld elements, X has %lld rows", (long long)Y.numel, (long long)g.N); return SMC_ERR_INVALID; }
    g.y = Y.dev;
    g.y_len = Y.numel;
  }
  return SMC_OK;
}

void smc_glm_release(SmcGlmPlan& g) {
  if (g.Xrm) cudaFree(g.Xrm);
  if (g.part) cudaFree(g.part);
  g.Xrm = nullptr;
  g.part = nullptr;
}

int smc_glm_launch(smc_model_s* m, SmcGlmPlan& g, int64_t C) {
  const smc_tape_stmt& s = m->stmts[g.stmt];
  cudaStream_t st = m->ctx->stream;
  int tc = chain_tile_for(C);
  long long chain_tiles = (C + tc - 1) / tc;
  int RT = g.MP >= 32 ? 64 : 128;
  int RS = row_splits_for(chain_tiles, g.N, RT, m->ctx->sm_count);
  long long rows_per_split = ((g.N + RS - 1) / RS + RT - 1) / RT * RT;
  RS = (int)((g.N + rows_per_split - 1) / rows_per_split);
  int64_t need = (int64_t)(RS + 1) * (g.MP + 1) * m->Cpad;
  if (need > g.part_cap) {
    if (g.part) cudaFree(g.part);
    g.part = nullptr;
    SMC_CUDA(cudaMalloc(&g.part, (size_t)need * sizeof(double)));
    g.part_cap = need;
  }
  GlmArgs a;
  a.X = g.Xrm; a.y = g.y; a.beta = m->beta + m->regs[s.src[1]].d.aux * m->Cpad;
  if (m->regs[s.src[1]].d.kind != SMC_REG_PARAM) a.beta = m->regs[s.src[1]].dev;
  a.part = g.part; a.bad = m->bad;
  a.N = g.N; a.rows_per_split = rows_per_split; a.C = C; a.Cpad = m->Cpad; a.M = g.M;
  a.par = g.par; a.y_scalar = g.y_scalar;
  a.sa_neg = g.sign_neg & 1; a.sb_neg = (g.sign_neg >> 1) & 1;
  if (m->time_glm) SMC_CUDA(cudaEventRecord(m->ev0, st));
  int rc = g.family == SMC_GLM_NORMAL ? launch_fam<SMC_GLM_NORMAL>(a, g.MP, tc, RS, st)
                                      : launch_fam<SMC_GLM_BERNOULLI>(a, g.MP, tc, RS, st);
  if (rc != SMC_OK) return rc;
  if (m->time_glm) {
    SMC_CUDA(cudaEventRecord(m->ev1, st));
    SMC_CUDA(cudaEventSynchronize(m->ev1));
    float ms = 0.f;
    SMC_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    g.last_ms += ms;
  }
  SMC_CUDA(cudaGetLastError());
  double* out = g.part + (int64_t)RS * (g.MP + 1) * m->Cpad;
  long long tot = (long long)(g.MP + 1) * C;
  k_glm_reduce<<<(unsigned)std::min<long long>((tot + 255) / 256, 148LL * 16), 256, 0, st>>>(g.part, RS, g.MP, C, m->Cpad, m->bad, out);
  if (m->row_shard && m->ctx->nranks > 1) SMC_TRY(smc_comm_allreduce_device(m->ctx, out, (int64_t)(g.MP + 1) * m->Cpad));
  const SmcReg& Lr = m->regs[s.dst];
  const SmcReg& Gr = m->regs[s.aux2];
  long long tot2 = (long long)(g.M + 1) * C;
  k_glm_publish<<<(unsigned)std::min<long long>((tot2 + 255) / 256, 148LL * 16), 256, 0, st>>>(
      out, g.M, g.MP, C, m->Cpad, m->bad, Lr.dev, Gr.dev, (s.flags & SMC_F_ACC) ? 1 : 0);
  m->ctx->launches += 3;
  return SMC_OK;
}
