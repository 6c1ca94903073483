// K1: fused GLM log-density + gradient for all chains in lockstep, FP64 tensor-core (DMMA) path.
//
//   eta[N x C] = X[N x M] * B[M x C]                (the generated `tmp = X * beta`)
//   (l, d)     = family(eta, y)   per (row, chain)   (logpdfNormal / logpdfBernoulli + their adjoint chain,
//                                                    /root/reference/src/distribs.jl:12-28, src/diff.jl:94-99,165-166)
//   L[C]       = sum_i l ;  G[M x C] = X' * d        (the generated `dbeta += transpose(X) * dtmp`, diff.jl:55)
//
// One pass over X per lockstep evaluation: a row tile of X is staged once in shared memory (cp.async, 4-stage ring)
// and used for both products.  Both contractions run on mma.sync.m8n8k4.f64 (DMMA); the accumulator
// fragment of the first product (chain x row) is, element for element, the A fragment of the second
// (chain x row) . (row x param) product, so eta / d never leave registers.  tcgen05 has no f64 kind.
//
// Grid: (chain tiles, row splits).  Each CTA writes a partial (L, G); a second kernel sums the row
// splits in fixed order (deterministic) and publishes into the tape registers.
// With one to four chains the same statement is an HBM stream instead (k_glm_stream_bulk / k_glm_stream below), and with
// the opt-in sufficient statistics a Gram-matrix evaluation (k_glm_gram).
#include <stdlib.h>
#include <algorithm>
#include "smc_internal.h"
#include "dists.cuh"

int smc_comm_allreduce_device(smc_ctx_s* c, double* buf, int64_t n);  // smc_comm.cu

namespace {

// rows per shared-memory tile.  (128-row tiles with three stages for MP = 64 -- half the CTA-wide barriers -- measured
// 32.72 ms per cfg2 launch against 32.07 ms for 64-row tiles with four stages, so the barrier is not what is left.)
#define SMC_GLM_RT(MP8) ((MP8) >= 4 ? 64 : 128)

struct GlmPub {    // where (L, G) go when a kernel may write the tape registers itself
  double* Ldst;
  double* Gdst;
  int acc, on;
};

struct GlmArgs {
  const double* X;     // [N][MP] row major, zero padded
  const double* y;     // [N] response as consumed by the family (sign-normalised)
  const double* beta;  // [M][Cpad]
  double* part;        // [RS][MP+1][Cpad]
  int* bad;
  long long N, rows_per_split, C, Cpad;
  int M;
  double par;
  int neg_beta;  // BERNOULLI with prob = 1/(1+exp(-eta)): the contraction runs on -beta, so it yields u = -eta directly
  int neg_out;   // ... and G = X' dl/deta = -(X' dl/du)
  double lc_sum;  // BINOMIAL: sum over this rank's rows of lchoose(n, y_i)
  GlmPub pub;    // single row split on a single rank: the epilogue publishes (L, G) itself, no partial buffer / reduce kernel
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
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

// NORMAL.  The reference computes z = x - mu with x/mu = +-eta, -+y; with w = eta - ytilde
// (ytilde = y, or -y when eta and y enter with the same sign: prepared on the host side of the launch) z = +-w, so
//   l = -r*r/2 - log(sigma) - log(sqrt(2pi)),  r = w/sigma            (src/distribs.jl:23-28)
//   dl/deta = -w/sigma^2                                               (src/diff.jl:99 through the +- of diff.jl:36-38)
// The constant -log(sigma) - log(sqrt(2pi)) is added once per row split in the epilogue; zero-filled tail rows
// (X = 0, y = 0) contribute w = 0.
// BERNOULLI.  prob = 1/(1+exp(u)), l = y ? log(prob) : log(1-prob)   (examples/binomial.jl:15, src/distribs.jl:12-21)
//   dl/du = exp(u) * (-1/t4^2) * 1/(prob-(1-y)) = y ? -e*prob : prob   (src/diff.jl:166,73,45)
template <int FAM>
__device__ __forceinline__ void family_eval(double eta, double y, double inv_s, double neg_inv_s2, bool valid, double& lacc,
                                            double& d) {
  if (FAM == SMC_GLM_NORMAL) {
    // Two FP64 instructions per (row, chain): DMMA and DFMA/DMUL share the FP64 pipe (ncu: tensor-FP64 % + pipe_fp64 % =
    // SM throughput), so every multiply here is DMMA time.  The 1/sigma^2 factors of both l and dl/deta are applied
    // once per chain in the epilogues: lacc accumulates sum w^2, the second product runs on w itself.
    double w = eta - y;
    lacc = fma(w, w, lacc);
    d = w;
    (void)inv_s; (void)neg_inv_s2;
  } else if (FAM == SMC_GLM_BERNOULLI) {
    double e = exp(eta), t4 = 1.0 + e, p = 1.0 / t4;
    bool one = (y == 1.0);
    double l = log(one ? p : 1.0 - p);
    d = one ? -e * p : p;
    lacc += valid ? l : 0.0;
  } else {
    // BINOMIAL counts, n trials (passed in inv_s): l = lchoose(n, y) + y log p + (n - y) log(1 - p)   (dbinom, src/distribs.jl:43);
    // lchoose does not depend on the chain: its sum over the rows is added once per chain in the epilogue.
    //   dl/du = exp(u) * (-1/t4^2) * (y/p - (n-y)/(1-p)) = n p - y                        (src/diff.jl:169-170,73,45)
    const double n = inv_s;
    double e = exp(eta), t4 = 1.0 + e, p = 1.0 / t4;
    double l = smc_xlogy(y, p) + smc_xlog1py(n - y, -p);
    d = n * p - y;
    lacc += valid ? l : 0.0;
  }
}

// Row r of an 8-row block sits at MMA column n with r = n < 4 ? n : n ^ 1.  With rows padded to LD = MP + 4 doubles
// (LD mod 16 in {4, 12}) this makes both operand reads bank-conflict free: product 1 reads rows {rowmap(g)} x 4
// consecutive columns per half-warp, product 2 reads rows {rowmap(2q+s)} x 4 consecutive columns.
__device__ __forceinline__ int rowmap(int n) { return n < 4 ? n : (n ^ 1); }

// TRIM: the model has at most MP - 4 columns, so the last k-step of product 1 would multiply zeros and is skipped
// (M = 10 in a 16-wide tile: 3 k-steps instead of 4, 7 DMMAs per row block instead of 8).
// PUB: one row split covers all rows (small N) on a single rank -- the epilogue writes (L, G) into the tape registers
// itself.  (A compile-time switch: as a run-time branch it changed the register allocation of the main loop and cost
// 2.5 % on cfg2, 32.9 vs 32.1 ms per launch.)
// TAIL: the last 8-column tile of product 2 holds only TAIL (1..3) live columns (M = 10 in a 16-wide tile: 2).  A DMMA on it
// spends 16 issue cycles per scheduler on 8 columns of which 8 - TAIL are zeros; DFMA runs at the same flop rate as DMMA on
// this part (34.2 against 37.2 TFLOP/s, profiles/fp64_peak_r01.json), so those columns are accumulated by plain DFMAs
// instead -- thread (g, q) owns chain g and rows (2q, 2q+1) of the block, 2 x TAIL DFMAs of 2 cycles each per m-tile -- and
// the four q lanes of a chain are summed once in the epilogue.  The same columns leave product 1's tensor k-steps too: its
// accumulator fragment (chain g, rows 2q, 2q+1) takes them as beta[8 + p] * x[row][8 + p] DFMAs.  examples/linear.jl
// (M = 10): 16 DMMAs + 32 DFMAs per row block of 32 chains instead of 28 DMMAs (448 -> 320 issue cycles per scheduler).
// (Measured non-improvement on cfg1 with 2^20 chains: the 128-chain tile, MT = 2, at THREE CTAs per SM -- 80 registers, three
// stages, six warps per scheduler instead of four -- 21.0 against 24.4 TFLOP/s for the 256-chain tile at two CTAs per SM.)
template <int FAM, int MT, int NWC, int NWR, int MP8, int TRIM, int PUB, int TAIL = 0>
__global__ void __launch_bounds__(NWC* NWR * 32, (MP8 <= 2 ? 2 : 1)) k_glm(GlmArgs a) {
  constexpr int MP = MP8 * 8, LD = MP + 4, NT = NWC * NWR * 32;
  constexpr int NTD = MP8 - (TAIL > 0 ? 1 : 0);   // 8-column tiles of product 2 on the tensor pipe
  constexpr int TL = TAIL > 0 ? TAIL : 1;
  constexpr int KS = TAIL > 0 ? 2 * (MP8 - 1) : MP / 4 - TRIM;   // k-steps of product 1 on the tensor pipe
  constexpr int RT = SMC_GLM_RT(MP8);        // rows per tile
  constexpr int NST = 4;    // stages; the copy of tile t+1 is in flight while tile t is consumed
  constexpr int TC = NWC * MT * 8;           // chains per CTA
  constexpr int BPW = RT / 8 / NWR;          // 8-row blocks per warp per tile
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
  double bfr[MT][KS];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    long long c = cbase + mt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      int j = ks * 4 + q;
      double bv = (j < a.M && c < a.C) ? __ldg(a.beta + (long long)j * a.Cpad + c) : 0.0;
      bfr[mt][ks] = a.neg_beta ? -bv : bv;
    }
  }
  double bt[MT][TL];                     // TAIL: beta of the columns 8 NTD .. 8 NTD + TAIL - 1 for chain g of every m-tile
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    long long c = cbase + mt * 8 + g;
#pragma unroll
    for (int p = 0; p < TL; p++) {
      int j = NTD * 8 + p;
      double bv = (TAIL > 0 && j < a.M && c < a.C) ? __ldg(a.beta + (long long)j * a.Cpad + c) : 0.0;
      bt[mt][p] = a.neg_beta ? -bv : bv;
    }
  }
  double G[MT][MP8][2];
  double Gt[MT][TL];                     // TAIL: this thread's rows of X' d for the columns 8 NTD .. 8 NTD + TAIL - 1
  double lsum[MT];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    lsum[mt] = 0.0;
#pragma unroll
    for (int nt = 0; nt < MP8; nt++) G[mt][nt][0] = G[mt][nt][1] = 0.0;
#pragma unroll
    for (int p = 0; p < TL; p++) Gt[mt][p] = 0.0;
  }
  double inv_s = 1.0, neg_inv_s2 = 0.0;
  if (FAM == SMC_GLM_NORMAL) {
    inv_s = 1.0 / a.par;
    neg_inv_s2 = -1.0 / (a.par * a.par);
  }
  if (FAM == SMC_GLM_BINOMIAL) inv_s = a.par;   // number of trials

  // Measured alternatives that did NOT help on cfg2 (32.07 ms per launch as written):
  //  * issuing this copy a few chunks per row-block iteration instead of in one burst after the barrier: 33.98 ms
  //    (the extra address arithmetic inside the DMMA loop costs more than the burst);
  //  * full/empty mbarriers (cp.async.mbarrier.arrive) with two tiles in flight instead of the CTA-wide barrier, with
  //    and without starting the upper half of the warps 0.2-3 us late so that the two warps of a scheduler run out of
  //    phase: 32.9 ms at 254 registers, independent of the skew;
  //  * even / odd k-steps of product 1 on separate accumulators (four independent DMMA chains per warp instead of
  //    two): 32.4 ms -- the dependent-DMMA latency is not what leaves the pipe idle.
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
      for (int r = tid; r < RT; r += NT) {
        long long row = r0 + r;
        bool ok = row < row_end;
        cp_async8(yy + r, a.y + (ok ? row : 0), ok ? 8 : 0);
      }
    }
    cp_async_commit();
  };
  // tile t has landed for every thread; the stage of tile t-3 is free again
  auto tile_begin = [&](int t) {
    cp_async_wait<0>();
    __syncthreads();
    load_tile(t + 1);
  };
  // product 1 on the block `gb` of this warp: acc[mt][s] = eta(chain cbase+mt*8+g, row rowmap(2q+s))
  auto gemm1 = [&](int gb, double (&acc)[MT][2]) {
    const int t = gb / BPW, rb = (wr + (gb % BPW) * NWR) * 8;
    const double* xrow = Xs + (t % NST) * RT * LD + (rb + rowmap(g)) * LD + q;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      double xb = xrow[ks * 4];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) dmma884(acc[mt][0], acc[mt][1], bfr[mt][ks], xb);
    }
    if (TAIL > 0) {
      const double* xt = Xs + (t % NST) * RT * LD + rb * LD + NTD * 8;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const double* xr = xt + rowmap(2 * q + s) * LD;
#pragma unroll
        for (int p = 0; p < TL; p++) {
          const double xv = xr[p];
#pragma unroll
          for (int mt = 0; mt < MT; mt++) acc[mt][s] = fma(bt[mt][p], xv, acc[mt][s]);
        }
      }
    }
  };

  // family stage of block gb (registers only): d[mt][s] = dl/deta, lsum += l
  auto family = [&](int gb, const double (&acc)[MT][2], double (&dd)[MT][2]) {
    const int t = gb / BPW, rb = (wr + (gb % BPW) * NWR) * 8;
    const double* yy = ys + (t % NST) * RT;
    const long long r0 = row_begin + (long long)t * RT;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      int r = rb + rowmap(2 * q + s);
      bool valid = (r0 + r) < row_end;
      double y = yy[r];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) family_eval<FAM>(acc[mt][s], y, inv_s, neg_inv_s2, valid, lsum[mt], dd[mt][s]);
    }
  };
  // product 2 on block gb: G[chain][j] += d[chain][row] * X[row][j]
  auto gemm2 = [&](int gb, const double (&dd)[MT][2]) {
    const int t = gb / BPW, rb = (wr + (gb % BPW) * NWR) * 8;
    const double* xs = Xs + (t % NST) * RT * LD;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const double* xr2 = xs + (rb + rowmap(2 * q + s)) * LD + g;
#pragma unroll
      for (int nt = 0; nt < NTD; nt++) {
        double xb = xr2[nt * 8];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) dmma884(G[mt][nt][0], G[mt][nt][1], dd[mt][s], xb);
      }
      if (TAIL > 0) {
        const double* xt = xs + (rb + rowmap(2 * q + s)) * LD + NTD * 8;      // the same address in the eight g lanes: a broadcast
#pragma unroll
        for (int p = 0; p < TL; p++) {
          const double xv = xt[p];
#pragma unroll
          for (int mt = 0; mt < MT; mt++) Gt[mt][p] = fma(dd[mt][s], xv, Gt[mt][p]);
        }
      }
    }
  };

  // Software pipeline over this warp's 8-row blocks, depth 2: iteration gb issues product 2 of block gb-1 and product 1 of
  // block gb+1 (64 DMMAs) and, in their issue shadow, the family stage of block gb, whose result is only consumed one
  // iteration later -- so no FP64-pipe time is lost to the elementwise stage.  Block counts are uniform over the CTA
  // (tail blocks are zero filled); a tile's stage is reused NST - 1 >= 2 tiles later (prefetch distance 1).
  const int total = ntiles * BPW;
  load_tile(0);
  if (total > 0) {
    int tile_ready = -1;
    auto need_tile = [&](int gb) {
      int t = gb / BPW;
      if (t > tile_ready) { tile_begin(t); tile_ready = t; }
    };
    double accA[MT][2], accB[MT][2], ddP[MT][2], ddC[MT][2];
    need_tile(0);
    gemm1(0, accA);
    family(0, accA, ddP);
    {
      const int nxt = min(1, total - 1);
      need_tile(nxt);
      gemm1(nxt, accA);
    }
#pragma unroll 1
    for (int gb = 1; gb < total; gb++) {
      const int nxt = min(gb + 1, total - 1);
      need_tile(nxt);
      gemm2(gb - 1, ddP);
      gemm1(nxt, accB);
      family(gb, accA, ddC);
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        accA[mt][0] = accB[mt][0]; accA[mt][1] = accB[mt][1];
        ddP[mt][0] = ddC[mt][0]; ddP[mt][1] = ddC[mt][1];
      }
    }
    gemm2(total - 1, ddP);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: reduce over the row warps in shared memory (fixed order), write the partial
  double* Gs = smem;             // [MP+1][TC]  (row MP = L)
  int* bs = (int*)(smem + (MP + 1) * TC);  // [TC]
  for (int i = tid; i < (MP + 1) * TC; i += NT) Gs[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < MT; mt++) {
    lsum[mt] += __shfl_xor_sync(0xffffffffu, lsum[mt], 1);
    lsum[mt] += __shfl_xor_sync(0xffffffffu, lsum[mt], 2);
    if (TAIL > 0) {
#pragma unroll
      for (int p = 0; p < TL; p++) {
        Gt[mt][p] += __shfl_xor_sync(0xffffffffu, Gt[mt][p], 1);
        Gt[mt][p] += __shfl_xor_sync(0xffffffffu, Gt[mt][p], 2);
      }
    }
  }
  for (int w = 0; w < NWR; w++) {
    if (wr == w) {
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        int cl = wc * (MT * 8) + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < NTD; nt++) {
          Gs[(nt * 8 + 2 * q) * TC + cl] += G[mt][nt][0];
          Gs[(nt * 8 + 2 * q + 1) * TC + cl] += G[mt][nt][1];
        }
        if (TAIL > 0 && q == 0) {
#pragma unroll
          for (int p = 0; p < TL; p++) Gs[(NTD * 8 + p) * TC + cl] += Gt[mt][p];
        }
        if (q == 0) Gs[MP * TC + cl] += lsum[mt];
      }
    }
    __syncthreads();
  }
  double* part = a.part + (long long)blockIdx.y * (MP + 1) * a.Cpad;
  const long long ct0 = (long long)blockIdx.x * TC;
  const double nrows = (double)(row_end > row_begin ? row_end - row_begin : 0);
  const double lconst = FAM == SMC_GLM_NORMAL ? nrows * (-log(a.par) - SMC_LOG_SQRT_2PI) : 0.0;
  for (int i = tid; i < (MP + 1) * TC; i += NT) {
    int j = i / TC, cl = i - j * TC;
    long long c = ct0 + cl;
    if (c >= a.C) continue;
    double v = Gs[i];
    if (j == MP) {
      if (FAM == SMC_GLM_NORMAL) v = 0.5 * neg_inv_s2 * v + lconst;   // -sum (w/sigma)^2 / 2 - n log(sigma sqrt(2 pi))
      if (FAM == SMC_GLM_BINOMIAL && blockIdx.y == 0) v += a.lc_sum;
      // a -inf / NaN sum means some observation had zero density ("give up eval", src/distribs.jl:15,17)
      if (!(v > -INFINITY)) a.bad[c] = 1;
    } else if (FAM == SMC_GLM_NORMAL) {
      v *= neg_inv_s2;                                                  // X' dl/deta = -(X' w) / sigma^2
    } else if (a.neg_out) {
      v = -v;
    }
    if (PUB) {
      if (j == MP) {
        if (a.bad[c] || !(v > -INFINITY)) { a.bad[c] = 1; v = 0.0; }
        a.pub.Ldst[c] = a.pub.acc ? a.pub.Ldst[c] + v : v;
      } else if (j < a.M) {
        a.pub.Gdst[(long long)j * a.Cpad + c] = v;
      }
    } else {
      part[(long long)j * a.Cpad + c] = v;
    }
  }
  (void)bs;
}

// Sum of the row-split partials in fixed order (deterministic), then either
//   publish: write (L, G) straight into the tape registers (one kernel per GLM statement on a single rank), or
//   stage:   out[j][c] = sum (j = MP row is L) for the NCCL all-reduce of row-sharded models; a chain flagged bad gets
//            L = -inf there so that the flag survives the cross-rank sum.
__device__ __forceinline__ void glm_emit(double t, long long j, long long c, int M, int MP, long long Cpad, int* bad, double* out,
                                         const GlmPub& pub) {
  if (pub.on) {
    if (j == MP) {
      if (bad[c] || !(t > -INFINITY)) { bad[c] = 1; t = 0.0; }
      pub.Ldst[c] = pub.acc ? pub.Ldst[c] + t : t;
    } else if (j < M) {
      pub.Gdst[j * Cpad + c] = t;
    }
  } else {
    if (j == MP && (bad[c] || !(t > -INFINITY))) t = -INFINITY;
    out[j * Cpad + c] = t;
  }
}

__global__ void k_glm_reduce(const double* part, int RS, int M, int MP, long long C, long long Cpad, int* bad, double* out, GlmPub pub) {
  long long total = (long long)(MP + 1) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i / C, c = i - j * C;
    if (pub.on && j >= M && j < MP) continue;
    double t = 0.0;
    for (int r = 0; r < RS; r++) t += part[((long long)r * (MP + 1) + j) * Cpad + c];
    glm_emit(t, j, c, M, MP, Cpad, bad, out, pub);
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

// ytilde[i] = scale * (src ? (bcast ? src[0] : src[i]) : imm)   (bcast carries the immediate when src is null)
__global__ void k_fill_y(const double* src, double imm_or_bcast, double scale, double* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = src ? (imm_or_bcast != 0.0 ? src[0] : src[i]) : imm_or_bcast;
    out[i] = scale * v;
  }
}

// partial sums of lchoose(n, y_i) (one per block, summed in order on the host at prepare time)
__global__ void __launch_bounds__(256) k_binom_const(const double* y, long long N, double n, double* part) {
  __shared__ double sh[256];
  double t = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256)
    t += lgamma(n + 1.0) - lgamma(y[i] + 1.0) - lgamma(n - y[i] + 1.0);
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int h = 128; h >= 1; h >>= 1) {
    if ((int)threadIdx.x < h) sh[threadIdx.x] += sh[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void k_to_row_major(const double* Xcm, double* Xrm, long long N, int M, int MP) {
  long long total = N * MP;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / MP;
    int j = (int)(i - r * MP);
    Xrm[i] = j < M ? Xcm[(long long)j * N + r] : 0.0;
  }
}


// ---- few chains (C <= 4): HBM stream of X ------------------------------------------------------------------------
// With one to four chains the contraction is a matrix-vector product and the DMMA tile (8 chains) would spend 7/8 of
// the family stage (exp / log / divide per row and chain) on padding lanes.  Here a thread owns a row: it reads the
// row's M values from the column-major X the caller bound (consecutive threads = consecutive rows: every load is a
// full coalesced line, no padded columns are read), forms eta for CT chains, evaluates the family and accumulates
// G[j][chain] += d * x[j] in registers.  Algorithmic bytes = bytes moved: 8*N*M + 8*N per evaluation.
struct GlmStreamArgs {
  double lc_sum;       // BINOMIAL: sum of lchoose(n, y_i) over the rows
  const double* X;     // [M][N] column major
  const double* y;     // [N] sign-normalised response
  const double* beta;  // [M][Cpad]
  double* part;        // [gridDim.x][MP+1][Cpad]
  int* bad;
  long long N, C, Cpad;
  int M, MP;
  double par;
  int neg_beta, neg_out;
};

template <int FAM, int MB, int CT>
__global__ void __launch_bounds__(256) k_glm_stream(GlmStreamArgs a) {
  __shared__ double bs[MB * CT];
  __shared__ double red[8][(MB + 1) * CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c0 = (long long)blockIdx.y * CT;
  for (int i = tid; i < MB * CT; i += 256) {
    int j = i / CT, ct = i - j * CT;
    double bv = (j < a.M && c0 + ct < a.C) ? __ldg(a.beta + (long long)j * a.Cpad + c0 + ct) : 0.0;
    bs[i] = a.neg_beta ? -bv : bv;
  }
  __syncthreads();
  double G[MB][CT], lsum[CT];
#pragma unroll
  for (int ct = 0; ct < CT; ct++) {
    lsum[ct] = 0.0;
#pragma unroll
    for (int j = 0; j < MB; j++) G[j][ct] = 0.0;
  }
  double inv_s = 1.0, neg_inv_s2 = 0.0;
  if (FAM == SMC_GLM_NORMAL) {
    inv_s = 1.0 / a.par;
    neg_inv_s2 = -1.0 / (a.par * a.par);
  }
  if (FAM == SMC_GLM_BINOMIAL) inv_s = a.par;
  // software pipeline over this thread's rows: the loads of the next row are in flight while the family stage
  // (exp / divide / log) of the current one runs
  const long long stride = (long long)gridDim.x * 256;
  long long r = (long long)blockIdx.x * 256 + tid;
  double xn[MB], yn = 0.0;
#pragma unroll
  for (int j = 0; j < MB; j++) xn[j] = (j < a.M && r < a.N) ? __ldg(a.X + (long long)j * a.N + r) : 0.0;
  if (r < a.N) yn = __ldg(a.y + r);
  while (r < a.N) {
    double x[MB];
#pragma unroll
    for (int j = 0; j < MB; j++) x[j] = xn[j];
    const double y = yn;
    const long long rn = r + stride;
    if (rn < a.N) {
#pragma unroll
      for (int j = 0; j < MB; j++) xn[j] = (j < a.M) ? __ldg(a.X + (long long)j * a.N + rn) : 0.0;
      yn = __ldg(a.y + rn);
    }
#pragma unroll
    for (int ct = 0; ct < CT; ct++) {
      double eta = 0.0;
#pragma unroll
      for (int j = 0; j < MB; j++) eta = fma(x[j], bs[j * CT + ct], eta);
      double d;
      family_eval<FAM>(eta, y, inv_s, neg_inv_s2, true, lsum[ct], d);
#pragma unroll
      for (int j = 0; j < MB; j++) G[j][ct] = fma(d, x[j], G[j][ct]);
    }
    r = rn;
  }
  // block reduction: shuffle tree inside each warp, then the 8 warps in order
#pragma unroll
  for (int ct = 0; ct < CT; ct++) {
#pragma unroll
    for (int j = 0; j <= MB; j++) {
      double v = j < MB ? G[j < MB ? j : 0][ct] : lsum[ct];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][j * CT + ct] = v;
    }
  }
  __syncthreads();
  double* part = a.part + (long long)blockIdx.x * (a.MP + 1) * a.Cpad;
  for (int i = tid; i < (MB + 1) * CT; i += 256) {
    int j = i / CT, ct = i - j * CT;
    long long c = c0 + ct;
    if (c >= a.C || (j < MB && j >= a.M)) continue;
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += red[w][i];
    if (j == MB) {
      if (FAM == SMC_GLM_NORMAL) {
        v = 0.5 * neg_inv_s2 * v;
        if (blockIdx.x == 0) v += (double)a.N * (-log(a.par) - SMC_LOG_SQRT_2PI);
      }
      if (FAM == SMC_GLM_BINOMIAL && blockIdx.x == 0) v += a.lc_sum;
      if (!(v > -INFINITY)) a.bad[c] = 1;
      part[(long long)a.MP * a.Cpad + c] = v;
    } else {
      part[(long long)j * a.Cpad + c] = FAM == SMC_GLM_NORMAL ? v * neg_inv_s2 : (a.neg_out ? -v : v);
    }
  }
}

// ---- the same stream staged by the bulk-copy engine (TMA, cp.async.bulk) ------------------------------------------------
// Same thread-to-row map, same arithmetic in the same order (results are bit-identical to k_glm_stream); what changes is
// who moves the bytes.  A 256-row tile of X is M contiguous 2 KB column segments plus 2 KB of y: one elected thread
// (a ninth warp) hands them to the copy engine as M + 1 bulk copies that complete on the stage's "full" mbarrier, up to
// NST tiles ahead; each compute warp releases a stage on its "empty" mbarrier when it has read its rows, so the warps
// never wait for one another (with a block barrier per tile the barrier was the top stall: ncu 2.4 of 7.4 cycles).  No registers are tied up by loads in flight (the register-pipelined kernel keeps one row ahead:
// 88 B per thread, 45 KB per SM at two CTAs), so 2-3 stages x 22 KB x 2 CTAs are in flight per SM.
// Needs 16-byte aligned column segments: N even (otherwise the register-pipelined kernel runs).
// Measured non-improvement: three CTAs per SM (72 registers by re-reading the row from the stage instead of holding it,
// 3 stages): 0.177 ms against 0.170 ms on cfg3 with one chain, 0.231 against 0.220 ms for the Binomial counts; two rows
// per thread from two consecutive tiles (two independent exp / divide / log chains to interleave): 0.181 ms; 512-row
// tiles (4 KB segments, two stages): 0.203 ms.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int FAM, int MB, int CT>
__global__ void __launch_bounds__(288, CT == 1 ? 2 : 1) k_glm_stream_bulk(GlmStreamArgs a, int NST) {
  constexpr int TR = 256;
  extern __shared__ __align__(128) double tiles[];            // [NST][M + 1][TR]
  __shared__ __align__(8) unsigned long long full[8], empty[8];
  __shared__ double bs[MB * CT];
  __shared__ double red[8][(MB + 1) * CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;   // warps 0..7 compute, warp 8 feeds the copy engine
  const int rowsz = a.M + 1;
  const long long c0 = (long long)blockIdx.y * CT;
  const long long ntiles = (a.N + TR - 1) / TR;
  if (tid == 0) {
    for (int s2 = 0; s2 < NST; s2++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s2])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s2])), "r"(8) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < MB * CT; i += 288) {
    int j = i / CT, ct = i - j * CT;
    double bv = (j < a.M && c0 + ct < a.C) ? __ldg(a.beta + (long long)j * a.Cpad + c0 + ct) : 0.0;
    bs[i] = a.neg_beta ? -bv : bv;
  }
  __syncthreads();
  auto wait_parity = [](unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  };
  double G[MB][CT], lsum[CT];
#pragma unroll
  for (int ct = 0; ct < CT; ct++) {
    lsum[ct] = 0.0;
#pragma unroll
    for (int j = 0; j < MB; j++) G[j][ct] = 0.0;
  }
  double inv_s = 1.0, neg_inv_s2 = 0.0;
  if (FAM == SMC_GLM_NORMAL) {
    inv_s = 1.0 / a.par;
    neg_inv_s2 = -1.0 / (a.par * a.par);
  }
  if (FAM == SMC_GLM_BINOMIAL) inv_s = a.par;
  if (warp == 8) {
    // producer: tile k of this CTA goes to stage k % NST once the 8 compute warps have released that stage
    if (lane == 0) {
      int s = 0;
      unsigned parity = 1;   // (a fresh barrier passes a wait on the "previous" phase: the first NST tiles do not block)
      for (long long k = 0;; k++) {
        const long long t = (long long)blockIdx.x + k * gridDim.x;
        if (t >= ntiles) break;
        wait_parity(smem_u32(&empty[s]), parity);
        const long long r0 = t * TR;
        const unsigned bytes = (unsigned)(min((long long)TR, a.N - r0) * 8);
        const unsigned bar = smem_u32(&full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (unsigned)rowsz) : "memory");
        double* dst = tiles + (size_t)s * rowsz * TR;
        for (int j = 0; j < a.M; j++)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)j * TR)),
                       "l"(a.X + (long long)j * a.N + r0), "r"(bytes), "r"(bar)
                       : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)a.M * TR)),
                     "l"(a.y + r0), "r"(bytes), "r"(bar)
                     : "memory");
        if (++s == NST) { s = 0; parity ^= 1u; }
      }
    }
  } else {
    int s = 0;
    unsigned parity = 0;
    for (long long k = 0;; k++) {
      const long long t = (long long)blockIdx.x + k * gridDim.x;
      if (t >= ntiles) break;
      wait_parity(smem_u32(&full[s]), parity);
      const long long r = t * TR + tid;
      if (r < a.N) {
        const double* src = tiles + (size_t)s * rowsz * TR + tid;
        double x[MB];
#pragma unroll
        for (int j = 0; j < MB; j++) x[j] = j < a.M ? src[(size_t)j * TR] : 0.0;
        const double y = src[(size_t)a.M * TR];
#pragma unroll
        for (int ct = 0; ct < CT; ct++) {
          double eta = 0.0;
#pragma unroll
          for (int j = 0; j < MB; j++) eta = fma(x[j], bs[j * CT + ct], eta);
          double d;
          family_eval<FAM>(eta, y, inv_s, neg_inv_s2, true, lsum[ct], d);
#pragma unroll
          for (int j = 0; j < MB; j++) G[j][ct] = fma(d, x[j], G[j][ct]);
        }
      }
      __syncwarp();   // the warp's reads of the stage are done (they fed x[] and y above)
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (++s == NST) { s = 0; parity ^= 1u; }
    }
  }
  // block reduction: shuffle tree inside each warp, then the 8 warps in order (as in k_glm_stream)
#pragma unroll
  for (int ct = 0; ct < CT; ct++) {
#pragma unroll
    for (int j = 0; j <= MB; j++) {
      double v = j < MB ? G[j < MB ? j : 0][ct] : lsum[ct];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && warp < 8) red[warp][j * CT + ct] = v;
    }
  }
  __syncthreads();
  double* part = a.part + (long long)blockIdx.x * (a.MP + 1) * a.Cpad;
  for (int i = tid; i < (MB + 1) * CT; i += 288) {
    int j = i / CT, ct = i - j * CT;
    long long c = c0 + ct;
    if (c >= a.C || (j < MB && j >= a.M)) continue;
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += red[w][i];
    if (j == MB) {
      if (FAM == SMC_GLM_NORMAL) {
        v = 0.5 * neg_inv_s2 * v;
        if (blockIdx.x == 0) v += (double)a.N * (-log(a.par) - SMC_LOG_SQRT_2PI);
      }
      if (FAM == SMC_GLM_BINOMIAL && blockIdx.x == 0) v += a.lc_sum;
      if (!(v > -INFINITY)) a.bad[c] = 1;
      part[(long long)a.MP * a.Cpad + c] = v;
    } else {
      part[(long long)j * a.Cpad + c] = FAM == SMC_GLM_NORMAL ? v * neg_inv_s2 : (a.neg_out ? -v : v);
    }
  }
}

// the same with one warp per output (few chains, many row splits): lanes stride the row splits, fixed shuffle tree
__global__ void k_glm_reduce_warp(const double* part, int RS, int M, int MP, long long C, long long Cpad, int* bad, double* out, GlmPub pub) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)(M + 1) * C) return;
  long long jj = w / C, c = w - jj * C;
  long long j = jj == M ? MP : jj;
  double t = 0.0;
  for (int r = lane; r < RS; r += 32) t += part[((long long)r * (MP + 1) + j) * Cpad + c];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) glm_emit(t, j, c, M, MP, Cpad, bad, out, pub);
}

// bulk-copy staging applies when every column segment of a tile is 16-byte aligned; SMC_GLM_NO_BULK=1 keeps the
// register-pipelined kernel (A/B measurements)
static bool stream_bulk_ok(const GlmStreamArgs& a) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("SMC_GLM_NO_BULK"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && a.M <= 16 && (a.N % 2 == 0) && a.N >= 512 && ((uintptr_t)a.X % 16 == 0) && ((uintptr_t)a.y % 16 == 0);
}
template <int FAM, int MB, int CT>
int launch_stream_one(const GlmStreamArgs& a, dim3 grid, cudaStream_t st) {
  if (stream_bulk_ok(a)) {
    const size_t stage = (size_t)(a.M + 1) * 256 * sizeof(double);
    const int NST = (int)std::max<size_t>(2, std::min<size_t>(6, (104 * 1024) / stage));   // two CTAs per SM
    const size_t smem = stage * NST;
    static SmcOncePerDevice once;
    if (once.first()) SMC_CUDA(cudaFuncSetAttribute(k_glm_stream_bulk<FAM, MB, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    k_glm_stream_bulk<FAM, MB, CT><<<grid, 288, smem, st>>>(a, NST);
  } else {
    k_glm_stream<FAM, MB, CT><<<grid, 256, 0, st>>>(a);
  }
  return SMC_OK;
}
template <int FAM, int MB>
int launch_stream_ct(const GlmStreamArgs& a, int CT, dim3 grid, cudaStream_t st) {
  if (CT == 1) return launch_stream_one<FAM, MB, 1>(a, grid, st);
  if (CT == 2) return launch_stream_one<FAM, MB, 2>(a, grid, st);
  return launch_stream_one<FAM, MB, 4>(a, grid, st);
}
template <int FAM>
int launch_stream(const GlmStreamArgs& a, int MB, int CT, dim3 grid, cudaStream_t st) {
  if (MB == 8) return launch_stream_ct<FAM, 8>(a, CT, grid, st);
  if (MB == 16) return launch_stream_ct<FAM, 16>(a, CT, grid, st);
  if (CT > 1) { smc_set_error("GLM stream kernel: unsupported shape"); return SMC_ERR_INVALID; }
  k_glm_stream<FAM, 32, 1><<<grid, 256, 0, st>>>(a);
  return SMC_OK;
}

// the stream path covers M <= 16 with up to 4 chains and M <= 32 with one chain (register budget: x[MB] + G[MB][CT])
bool use_stream(int M, long long C) { return (M <= 16 && C <= 4) || (M <= 32 && C == 1); }


// ---- sufficient statistics (opt-in, SMC_GLM_F_GRAM) -----------------------------------------------------------------
// For the Normal identity-link family  sum_i (x_i.beta - ytilde_i)^2 = beta'(X'X)beta - 2 beta'(X'ytilde) + ytilde'ytilde,
// so with Z = [X, ytilde] the (M+1)x(M+1) matrix Z'Z, computed once when the model is finalized, carries everything an
// evaluation needs: 2*M^2 flops per chain instead of 4*N*M.  This is NOT what the reference's generated code computes
// (it streams X twice per evaluation); it is the algebraic hoisting SURVEY.md 8(d)/8(f) lists as a separately-labelled
// extra, never enabled by default, and it loses accuracy by cancellation when the residual is small against |y|.
struct GramArgs {
  const double* X;   // [M][N] column major
  const double* y;   // [N] ytilde
  long long N, rows_per_block;
  int M;
  double* part;      // [gridDim.x][P*P], P = M + 1
};
__global__ void __launch_bounds__(256) k_gram_partial(GramArgs a) {
  constexpr int TR = 64, MAXPAIRS = 17;   // (64+1)^2 = 4225 pairs <= 17 * 256
  extern __shared__ __align__(16) double zs[];   // [TR][P + 1]
  const int P = a.M + 1, LDZ = P + 1, tid = threadIdx.x;
  const long long r_begin = (long long)blockIdx.x * a.rows_per_block, r_end = min(a.N, r_begin + a.rows_per_block);
  double acc[MAXPAIRS];
#pragma unroll
  for (int t = 0; t < MAXPAIRS; t++) acc[t] = 0.0;
  for (long long r0 = r_begin; r0 < r_end; r0 += TR) {
    __syncthreads();
    for (int i = tid; i < TR * P; i += 256) {
      int col = i / TR, r = i - col * TR;
      long long row = r0 + r;
      double v = 0.0;
      if (row < r_end) v = col < a.M ? __ldg(a.X + (long long)col * a.N + row) : __ldg(a.y + row);
      zs[r * LDZ + col] = v;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MAXPAIRS; t++) {
      int p = tid + t * 256;
      if (p < P * P) {
        int j = p / P, k = p - j * P;
        double s = acc[t];
        for (int r = 0; r < TR; r++) s = fma(zs[r * LDZ + j], zs[r * LDZ + k], s);
        acc[t] = s;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAXPAIRS; t++) {
    int p = tid + t * 256;
    if (p < P * P) a.part[(long long)blockIdx.x * P * P + p] = acc[t];
  }
}
__global__ void k_gram_reduce(const double* part, int nblocks, int PP, double* out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PP) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; b++) t += part[(long long)b * PP + p];   // fixed order
  out[p] = t;
}

struct GramEvalArgs {
  const double* ZtZ;   // [P][P]
  const double* beta;  // [M][Cpad]
  double* Ldst;
  double* Gdst;
  int* bad;
  long long N, C, Cpad;
  int M, acc;
  double par;
};
// block = 32 chains x 8 parameter lanes; A = Z'Z and the block's beta tile in shared memory
__global__ void __launch_bounds__(256) k_glm_gram(GramEvalArgs a) {
  extern __shared__ __align__(16) double gs[];
  const int P = a.M + 1;
  double* A = gs;                  // [P][P]
  double* bs = gs + P * P;         // [M][32]
  double* red = bs + a.M * 32;     // [2][8][32]
  const int cx = threadIdx.x, ly = threadIdx.y, tid = ly * 32 + cx;
  const long long c = (long long)blockIdx.x * 32 + cx;
  for (int i = tid; i < P * P; i += 256) A[i] = __ldg(a.ZtZ + i);
  for (int i = tid; i < a.M * 32; i += 256) {
    int j = i / 32, cl = i - j * 32;
    long long cc = (long long)blockIdx.x * 32 + cl;
    bs[i] = cc < a.C ? __ldg(a.beta + (long long)j * a.Cpad + cc) : 0.0;
  }
  __syncthreads();
  const double inv_s2 = 1.0 / (a.par * a.par);
  double quad = 0.0, lin = 0.0;
  for (int j = ly; j < a.M; j += 8) {
    const double* Aj = A + j * P;
    double q = 0.0;
    for (int k = 0; k < a.M; k++) q = fma(Aj[k], bs[k * 32 + cx], q);
    const double bj = bs[j * 32 + cx], xy = Aj[a.M];
    quad = fma(bj, q, quad);
    lin = fma(bj, xy, lin);
    if (c < a.C) a.Gdst[(long long)j * a.Cpad + c] = -(q - xy) * inv_s2;   // X' dl/deta, dl/deta = -(eta - ytilde)/sigma^2
  }
  red[ly * 32 + cx] = quad;
  red[256 + ly * 32 + cx] = lin;
  __syncthreads();
  if (ly == 0 && c < a.C) {
    double tq = 0.0, tl = 0.0;
    for (int y = 0; y < 8; y++) { tq += red[y * 32 + cx]; tl += red[256 + y * 32 + cx]; }
    const double rss = tq - 2.0 * tl + A[a.M * P + a.M];
    double L = -0.5 * rss * inv_s2 + (double)a.N * (-log(a.par) - SMC_LOG_SQRT_2PI);
    if (!(L > -INFINITY)) { a.bad[c] = 1; L = 0.0; }
    a.Ldst[c] = a.acc ? a.Ldst[c] + L : L;
  }
}

template <int FAM, int MT, int NWC, int NWR, int MP8, int TRIM, int PUB, int TAIL = 0>
int launch_cfg_tp(const GlmArgs& a, int RS, cudaStream_t st) {
  constexpr int MP = MP8 * 8, LD = MP + 4, RT = SMC_GLM_RT(MP8), NST = 4, TC = NWC * MT * 8;
  size_t pipe = (size_t)NST * RT * (LD + 1) * sizeof(double);
  size_t epi = (size_t)(MP + 1) * TC * sizeof(double) + TC * sizeof(int);
  size_t smem = std::max(pipe, epi);
  auto kern = k_glm<FAM, MT, NWC, NWR, MP8, TRIM, PUB, TAIL>;
  static SmcOncePerDevice once;
  if (once.first()) SMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((a.C + TC - 1) / TC), (unsigned)RS);
  kern<<<grid, NWC * NWR * 32, smem, st>>>(a);
  return SMC_OK;
}

template <int FAM, int MT, int NWC, int NWR, int MP8, int TRIM>
int launch_cfg_t(const GlmArgs& a, int RS, cudaStream_t st) {
  if constexpr (MP8 == 2) {   // (the self-publishing variant is only built for narrow models: it matters for small N)
    // 9 or 10 columns (examples/linear.jl, examples/binomial.jl: M = 10) with a full device of chains: the tail columns of
    // product 2 on DFMAs (SMC_GLM_NO_TAIL=1 keeps them on the tensor pipe, for A/B measurements)
    if constexpr (TRIM == 1 && MT >= 2) {
      static int no_tail = -1;
      if (no_tail < 0) { const char* e = getenv("SMC_GLM_NO_TAIL"); no_tail = (e && e[0] == '1') ? 1 : 0; }
      if (!no_tail && a.M >= 9 && a.M <= 10)
        return a.pub.on ? launch_cfg_tp<FAM, MT, NWC, NWR, MP8, TRIM, 1, 2>(a, RS, st)
                        : launch_cfg_tp<FAM, MT, NWC, NWR, MP8, TRIM, 0, 2>(a, RS, st);
    }
    if (a.pub.on) return launch_cfg_tp<FAM, MT, NWC, NWR, MP8, TRIM, 1>(a, RS, st);
  }
  return launch_cfg_tp<FAM, MT, NWC, NWR, MP8, TRIM, 0>(a, RS, st);
}

template <int FAM, int MT, int NWC, int NWR, int MP8>
int launch_cfg(const GlmArgs& a, int RS, cudaStream_t st) {
  return a.M <= MP8 * 8 - 4 ? launch_cfg_t<FAM, MT, NWC, NWR, MP8, 1>(a, RS, st) : launch_cfg_t<FAM, MT, NWC, NWR, MP8, 0>(a, RS, st);
}

template <int FAM, int MP8>
int launch_mp(const GlmArgs& a, int tc, int RS, cudaStream_t st) {
  if constexpr (MP8 == 2) {
    // narrow models with very many chains: four m-tiles (32 chains) per warp, 256 chains per CTA.  A row block is then
    // 28 DMMAs instead of 14 for the same operand loads and loop overhead (G and the beta fragments are small here).
    if (tc == 256) return launch_cfg<FAM, 4, 8, 1, MP8>(a, RS, st);
  }
  switch (tc) {
    case 128: return launch_cfg<FAM, 2, 8, 1, MP8>(a, RS, st);   // (a 16-warp MT=1 variant measured 86.5% vs 88.0%)
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
  if (g.family != SMC_GLM_NORMAL && g.family != SMC_GLM_BERNOULLI && g.family != SMC_GLM_BINOMIAL) { smc_set_error("GLM: family %d is not implemented", g.family); return SMC_ERR_INVALID; }
  g.MP = g.M <= 16 ? 16 : (g.M <= 32 ? 32 : 64);
  cudaStream_t st = m->ctx->stream;
  g.Xcm = X.dev;   // the row-major padded copy for the DMMA kernel is built on its first launch (smc_glm_launch)
  // ytilde[N]: the response as the kernel consumes it (negated when eta and y enter the residual with the same sign,
  // broadcast when the model uses a scalar)
  double scale = 1.0;
  if (g.family == SMC_GLM_NORMAL) {
    int sa_neg = g.sign_neg & 1, sb_neg = (g.sign_neg >> 1) & 1;
    scale = (sa_neg != sb_neg) ? 1.0 : -1.0;   // ytilde = -sa*sb*y
  }
  SMC_CUDA(smc_dev_alloc(&g.ybuf, (size_t)g.N * sizeof(double), st));
  if (Y.d.kind == SMC_REG_IMM) {
    k_fill_y<<<(unsigned)std::min<long long>((g.N + 255) / 256, 148LL * 32), 256, 0, st>>>(nullptr, Y.d.imm, scale, g.ybuf, g.N);
  } else {
    if (Y.numel != 1 && Y.numel < g.N) { smc_set_error("GLM: response has %lld elements, X has %lld rows", (long long)Y.numel, (long long)g.N); return SMC_ERR_INVALID; }
    k_fill_y<<<(unsigned)std::min<long long>((g.N + 255) / 256, 148LL * 32), 256, 0, st>>>(Y.dev, Y.numel == 1 ? 1.0 : 0.0, scale, g.ybuf, g.N);
  }
  m->ctx->launches++;
  g.y = g.ybuf;
  g.y_len = g.N;
  g.lc_sum = 0.0;
  if (g.family == SMC_GLM_BINOMIAL) {
    const int nb = (int)std::max<long long>(1, std::min<long long>((g.N + 255) / 256, 2LL * m->ctx->sm_count));
    double* part = nullptr;
    SMC_CUDA(smc_dev_alloc(&part, (size_t)nb * sizeof(double), st));
    k_binom_const<<<nb, 256, 0, st>>>(g.y, g.N, g.par, part);
    m->ctx->launches++;
    std::vector<double> h(nb);
    SMC_CUDA(cudaMemcpyAsync(h.data(), part, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    SMC_CUDA(cudaStreamSynchronize(st));
    smc_dev_free(part, st);
    for (double v : h) g.lc_sum += v;
  }
  g.gram = (g.family == SMC_GLM_NORMAL && (g.sign_neg & SMC_GLM_F_GRAM) && !m->row_shard) ? 1 : 0;
  if (g.gram) {
    const int P = g.M + 1, PP = P * P;
    const int nblocks = (int)std::max<long long>(1, std::min<long long>((g.N + 63) / 64, 2LL * m->ctx->sm_count));
    GramArgs ga;
    ga.X = g.Xcm; ga.y = g.y; ga.N = g.N; ga.M = g.M;
    ga.rows_per_block = ((g.N + nblocks - 1) / nblocks + 63) / 64 * 64;
    double* part = nullptr;
    SMC_CUDA(smc_dev_alloc(&part, (size_t)nblocks * PP * sizeof(double), st));
    SMC_CUDA(smc_dev_alloc(&g.ZtZ, (size_t)PP * sizeof(double), st));
    ga.part = part;
    const size_t smem = (size_t)64 * (P + 1) * sizeof(double);
    k_gram_partial<<<nblocks, 256, smem, st>>>(ga);
    k_gram_reduce<<<(PP + 255) / 256, 256, 0, st>>>(part, nblocks, PP, g.ZtZ);
    m->ctx->launches += 2;
    SMC_CUDA(cudaStreamSynchronize(st));
    smc_dev_free(part, st);
  }
  return SMC_OK;
}

void smc_glm_release(SmcGlmPlan& g, cudaStream_t st) {
  if (g.Xrm) smc_dev_free(g.Xrm, st);
  if (g.part) smc_dev_free(g.part, st);
  if (g.ybuf) smc_dev_free(g.ybuf, st);
  if (g.ZtZ) smc_dev_free(g.ZtZ, st);
  smc_glm_i8_release(g, st);
  g.ZtZ = nullptr;
  g.ybuf = nullptr;
  g.Xrm = nullptr;
  g.part = nullptr;
}

int smc_glm_launch(smc_model_s* m, SmcGlmPlan& g, int64_t C) {
  const smc_tape_stmt& s = m->stmts[g.stmt];
  cudaStream_t st = m->ctx->stream;
  if (g.gram) {
    GramEvalArgs e;
    e.ZtZ = g.ZtZ;
    e.beta = m->regs[s.src[1]].d.kind == SMC_REG_PARAM ? m->beta + m->regs[s.src[1]].d.aux * m->Cpad : m->regs[s.src[1]].dev;
    e.Ldst = m->regs[s.dst].dev; e.Gdst = m->regs[s.aux2].dev; e.bad = m->bad;
    e.N = g.N; e.C = C; e.Cpad = m->Cpad; e.M = g.M; e.acc = (s.flags & SMC_F_ACC) ? 1 : 0; e.par = g.par;
    const int P = g.M + 1;
    const size_t smem = ((size_t)P * P + (size_t)g.M * 32 + 512) * sizeof(double);
    static SmcOncePerDevice gram_once;
    if (gram_once.first()) SMC_CUDA(cudaFuncSetAttribute(k_glm_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    if (m->time_glm) SMC_CUDA(cudaEventRecord(m->ev0, st));
    k_glm_gram<<<(unsigned)((C + 31) / 32), dim3(32, 8), smem, st>>>(e);
    m->ctx->launches++;
    if (m->time_glm) {
      SMC_CUDA(cudaEventRecord(m->ev1, st));
      SMC_CUDA(cudaEventSynchronize(m->ev1));
      float ms = 0.f;
      SMC_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
      g.last_ms += ms;
    }
    return SMC_OK;
  }
  const bool stream = use_stream(g.M, C);
  if (!stream && !g.Xrm) {
    SMC_CUDA(smc_dev_alloc(&g.Xrm, (size_t)g.N * g.MP * sizeof(double), st));
    k_to_row_major<<<(unsigned)std::min<long long>((g.N * g.MP + 255) / 256, 148LL * 64), 256, 0, st>>>(g.Xcm, g.Xrm, g.N, g.M, g.MP);
    m->ctx->launches++;
  }
  int tc = chain_tile_for(C);
  if (g.MP <= 16 && C >= 2048 && !getenv("SMC_GLM_NO_WIDE")) tc = 256;
  long long chain_tiles = (C + tc - 1) / tc;
  int RT = SMC_GLM_RT(g.MP / 8);
  int RS = row_splits_for(chain_tiles, g.N, RT, m->ctx->sm_count * (g.MP <= 16 ? 2 : 1));  // narrow models run 2 CTAs per SM
  long long rows_per_split = ((g.N + RS - 1) / RS + RT - 1) / RT * RT;
  RS = (int)((g.N + rows_per_split - 1) / rows_per_split);
  if (stream) RS = (int)std::max<long long>(1, std::min<long long>((g.N + 255) / 256, 2LL * m->ctx->sm_count));
  // Normal family, 33-64 columns, >= 32 chains: the sliced-integer kernel on the int8 tensor cores, 32-chain tiles x 128-row
  // tiles (glm_i8.cu; SMC_GLM_I8=0 keeps the FP64 DMMA kernel)
  bool i8 = !stream && smc_glm_i8_usable(g, C);
  if (i8) {
    SMC_TRY(smc_glm_i8_prepare(m, g));
    i8 = !g.i8_declined;
  }
  if (i8) RS = row_splits_for(smc_glm_i8_chain_tiles(C), g.N, smc_glm_i8_row_tile(), m->ctx->sm_count);
  int64_t need = (int64_t)(RS + 1) * (g.MP + 1) * m->Cpad;
  if (need > g.part_cap) {
    m->graphs_stale = true;
    m->realloc_epoch++;
    if (g.part) smc_dev_free(g.part, st);
    g.part = nullptr;
    SMC_CUDA(smc_dev_alloc(&g.part, (size_t)need * sizeof(double), st));
    g.part_cap = need;
  }
  GlmArgs a;
  a.X = g.Xrm; a.y = g.y; a.beta = m->beta + m->regs[s.src[1]].d.aux * m->Cpad;
  if (m->regs[s.src[1]].d.kind != SMC_REG_PARAM) a.beta = m->regs[s.src[1]].dev;
  a.part = g.part; a.bad = m->bad;
  a.N = g.N; a.rows_per_split = rows_per_split; a.C = C; a.Cpad = m->Cpad; a.M = g.M;
  a.par = g.par;
  a.neg_beta = a.neg_out = (g.family != SMC_GLM_NORMAL && (g.sign_neg & 1)) ? 1 : 0;
  a.lc_sum = g.lc_sum;
  const bool sharded = m->row_shard && m->ctx->nranks > 1;
  a.pub.Ldst = m->regs[s.dst].dev; a.pub.Gdst = m->regs[s.aux2].dev; a.pub.acc = (s.flags & SMC_F_ACC) ? 1 : 0;
  a.pub.on = (!stream && RS == 1 && !sharded && g.MP == 16) ? 1 : 0;
  if (m->time_glm) SMC_CUDA(cudaEventRecord(m->ev0, st));
  int rc;
  if (stream) {
    GlmStreamArgs sa;
    sa.X = g.Xcm; sa.y = a.y; sa.beta = a.beta; sa.part = a.part; sa.bad = a.bad;
    sa.N = g.N; sa.C = C; sa.Cpad = m->Cpad; sa.M = g.M; sa.MP = g.MP; sa.par = g.par;
    sa.neg_beta = a.neg_beta; sa.neg_out = a.neg_out; sa.lc_sum = g.lc_sum;
    const int MB = g.M <= 8 ? 8 : (g.M <= 16 ? 16 : 32);
    const int CT = C == 1 ? 1 : (C == 2 ? 2 : 4);
    dim3 grid((unsigned)RS, (unsigned)((C + CT - 1) / CT));
    rc = g.family == SMC_GLM_NORMAL      ? launch_stream<SMC_GLM_NORMAL>(sa, MB, CT, grid, st)
         : g.family == SMC_GLM_BERNOULLI ? launch_stream<SMC_GLM_BERNOULLI>(sa, MB, CT, grid, st)
                                         : launch_stream<SMC_GLM_BINOMIAL>(sa, MB, CT, grid, st);
  } else if (i8) {
    a.pub.on = 0;
    rc = smc_glm_i8_launch(m, g, a.beta, C, m->Cpad, RS, a.part);
  } else {
    rc = g.family == SMC_GLM_NORMAL      ? launch_fam<SMC_GLM_NORMAL>(a, g.MP, tc, RS, st)
         : g.family == SMC_GLM_BERNOULLI ? launch_fam<SMC_GLM_BERNOULLI>(a, g.MP, tc, RS, st)
                                         : launch_fam<SMC_GLM_BINOMIAL>(a, g.MP, tc, RS, st);
  }
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
  const SmcReg& Lr = m->regs[s.dst];
  const SmcReg& Gr = m->regs[s.aux2];
  if (a.pub.on) { m->ctx->launches += 1; return SMC_OK; }   // the kernel published (L, G) itself
  GlmPub pub;
  pub.Ldst = Lr.dev; pub.Gdst = Gr.dev; pub.acc = (s.flags & SMC_F_ACC) ? 1 : 0; pub.on = sharded ? 0 : 1;
  if (sharded && smc_p2p_ready(m->ctx, (int64_t)(g.M + 1) * C)) {
    // rows are split over ranks and the host has set up the peer-memory exchange (smc_ctx_p2p_export / _import): the sum over
    // this rank's row splits, the sum over ranks (in rank order) and the publish step are ONE kernel (smc_p2p.cu)
    return smc_p2p_glm_exchange(m->ctx, g.part, RS, g.M, g.MP, C, m->Cpad, m->bad, Lr.dev, Gr.dev, pub.acc);
  }
  long long tot = (long long)(g.MP + 1) * C;
  if ((long long)(g.M + 1) * C <= 1024 && RS >= 16)
    k_glm_reduce_warp<<<(unsigned)(((long long)(g.M + 1) * C * 32 + 255) / 256), 256, 0, st>>>(g.part, RS, g.M, g.MP, C, m->Cpad, m->bad, out, pub);
  else
    k_glm_reduce<<<(unsigned)std::min<long long>((tot + 255) / 256, 148LL * 16), 256, 0, st>>>(g.part, RS, g.M, g.MP, C, m->Cpad, m->bad, out, pub);
  m->ctx->launches += 2;
  if (sharded) {
    // rows are split over ranks: sum the (L, G) block over the communicator (NCCL, in-stream), then publish
    SMC_TRY(smc_comm_allreduce_device(m->ctx, out, (int64_t)(g.MP + 1) * m->Cpad));
    long long tot2 = (long long)(g.M + 1) * C;
    k_glm_publish<<<(unsigned)std::min<long long>((tot2 + 255) / 256, 148LL * 16), 256, 0, st>>>(
        out, g.M, g.MP, C, m->Cpad, m->bad, Lr.dev, Gr.dev, pub.acc);
    m->ctx->launches++;
  }
  return SMC_OK;
}
