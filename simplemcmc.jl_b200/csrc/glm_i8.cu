// Sliced-integer evaluation of a Normal identity-link GLM statement on the int8 tensor cores (tcgen05.mma kind::i8, int32
// accumulators in TMEM): DESIGN.md section 8.  The path for Normal GLM statements with 33-64 columns and >= 32 chains
// (SMC_GLM_I8=0 keeps them on the FP64 DMMA kernel of glm_fused.cu): measured 20.0 ms per configs[1] evaluation against
// 32.2 ms (profiles/bench_1gpu_r02.json, profiles/glm_i8_v14_ncu_r02.json), logp 2e-16 / gradient 7e-14 against long double at the full size, and the whole
// 1e-12 parity suite runs through it (tests/test_gpu_scale.py, tests/test_zz_gpu_i8.py).
//
// Arithmetic (exact integer products, one FP64 rounding per recombination): every operand is cut into seven balanced
// 8-bit slices against a power-of-two scale -- X once with one global scale, beta per chain and evaluation, the adjoint
// w per chain with a STICKY scale (below) -- slice pairs (i, j) with i + j <= 6 are kept and pairs of equal i + j share an
// int32 accumulator.  One CTA = 32 chains x a range of 128-row tiles; 16 epilogue warps (4 per TMEM lane quadrant) and one
// issuer warp; 480 TMEM columns:
//   product 1: A = X slice i (K-major), B = beta slices j <= 6 - i stacked along N, D at column offset 32 i;
//   product 2: A = X slices (i, i+1) stacked along M, read MN-major from the same shared-memory tile, B = the stacked
//              slices of the adjoint (MN-major, written as seven 8-byte words per thread), D at column offset 256 + 32 i
//              (lanes 64..127 hold the groups shifted by one), ACCUMULATED over tiles and read back only on a flush.
// Output: the partial block [row split][MP + 1][Cpad] of k_glm (rows 0..63 = X' dl/deta, row 64 = log-density), so the
// fixed-order reduction, the publish step and the row-sharded exchange of glm_fused.cu / smc_p2p.cu apply unchanged.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "smc_internal.h"
#include "dists.cuh"

namespace {

constexpr int NC = 32;        // chains per CTA
constexpr int TR = 128;       // rows per tile
constexpr int MP = 64;        // parameter slots
constexpr int NS8 = 7;        // slices
constexpr int NT = 512;       // threads per CTA
constexpr int CQ = NC / (NT / 128);
constexpr uint32_t XT_BYTES = TR * 8 * MP;         // one X tile: 8 slice slots (slot 7 = zeros)
constexpr uint32_t QB_BYTES = 8 * NC * MP;
constexpr unsigned long long H8 = 0x0080808080808080ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, no swizzle: start address, LBO (steps along K), SBO (steps along M/N), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor for kind::i8: D = S32, A = B = signed 8 bit
__device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
#ifndef SMC_I8_WAIT_NS
#define SMC_I8_WAIT_NS 2000u     // suspend-time hint of a failed try_wait
#endif
// bounded wait (a wrong descriptor must end in an error, not in a hung device)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int it = 0; it < (1 << 20); it++) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(SMC_I8_WAIT_NS) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_load_tile(uint8_t* dst, const int8_t* src, uint64_t* bar) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(b), "r"(65536u) : "memory");
#pragma unroll
  for (int k = 0; k < 4; k++)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst + k * 16384)), "l"(src + k * 16384), "r"(16384u), "r"(b) : "memory");
}

// X tile image (shared memory = global): 128-byte core matrices [row group r>>3][slice slot i][parameter chunk p>>4]
// [row r&7][parameter p&15]; product 1 reads slot i K-major (LBO 128, SBO 4096), product 2 reads slots (i, i+1) MN-major
// (SBO 128, LBO 4096)
__host__ __device__ inline uint32_t xoff(int r, int i, int p) {
  return (uint32_t)(r >> 3) * 4096 + (uint32_t)i * 512 + (uint32_t)(p >> 4) * 128 + (uint32_t)(r & 7) * 16 + (uint32_t)(p & 15);
}
// stacked K-major B operands: row n = 32 j + chain, kbytes per row
__device__ inline uint32_t boff(int n, int k, int kbytes) {
  return (uint32_t)(n >> 3) * (uint32_t)(kbytes * 8) + (uint32_t)(k >> 4) * 128 + (uint32_t)(n & 7) * 16 + (uint32_t)(k & 15);
}
// slices of v against scale = 2^54 / f: byte k of the result = slice 6 - k as int8 (slice 0 = most significant)
// (round to nearest: truncation shrinks every element of X towards zero, a bias that adds up over the rows -- at the
// configs[1] size 1.5e-13 of max |G| instead of 1e-15, tools/ozaki_study.py and DESIGN.md section 8)
__device__ __forceinline__ unsigned long long slice7_word(double v, double f) {
  long long t = __double2ll_rn(v * f);
  return ((unsigned long long)t + H8) ^ H8;
}
// power-of-two scale strictly above m, from m's exponent field; factor = 2^54 / scale
__host__ __device__ inline void scale_of8(double m, double* scale, double* factor) {
  unsigned long long bits;
  memcpy(&bits, &m, 8);
  unsigned long long E = (bits >> 52) & 0x7ff;
  if (E < 64 || E > 1900) { *scale = 1.0; *factor = 18014398509481984.0; return; }
  unsigned long long sb = (E + 1) << 52, fb = (2099 - E) << 52;
  memcpy(scale, &sb, 8);
  memcpy(factor, &fb, 8);
}
// eight beta slices (SMC_GLM_I8_BETA8=1; measured: gradient error 3e-15 instead of 7e-14 at 3 % more time): t = round(v 2^62 / scale), byte k = slice 7 - k; the
// resolution of beta is the operand that limits the gradient with seven slices (profiles/ozaki_error_attribution_r01.json)
__device__ __forceinline__ unsigned long long slice8_word(double v, double f62) {
  const unsigned long long H64 = 0x8080808080808080ull;
  long long t = __double2ll_rn(v * f62);
  return ((unsigned long long)t + H64) ^ H64;
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// max |X| over the row-major padded copy (bits of a non-negative double order like the unsigned integer)
__global__ void k_absmax(const double* __restrict__ X, long long n, unsigned long long* out) {
  unsigned long long m = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double a = fabs(X[i]);
    if (!(a <= 1.7976931348623157e308)) a = INFINITY;      // NaN / Inf: reported as Inf
    unsigned long long b = (unsigned long long)__double_as_longlong(a);
    m = b > m ? b : m;
  }
  for (int o = 16; o > 0; o >>= 1) { unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// X (row major [N][64], zero padded columns) -> slice tiles; rows beyond N are zeros
__global__ void k_slice_x8(const double* __restrict__ X, int8_t* __restrict__ QX, long long N, long long prow, double factor) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= prow * MP) return;
  long long row = e / MP;
  int p = (int)(e % MP), r = (int)(row % TR);
  const unsigned long long w = slice7_word(row < N ? X[e] : 0.0, factor);
  int8_t* tile = QX + (row / TR) * (long long)XT_BYTES;
#pragma unroll
  for (int i = 0; i < NS8; i++) tile[xoff(r, i, p)] = (int8_t)(w >> (8 * (6 - i)));
  tile[xoff(r, 7, p)] = 0;
}

struct GlmI8Args {
  const int8_t* QX;     // [tiles][XT_BYTES]
  const double* y;      // [N] response as the Normal family consumes it (w = eta - y)
  const double* beta;   // [M][Cpad]
  double* part;         // [splits][MP + 1][Cpad]
  int* status;          // set when a tensor-core completion never arrived
  long long N, tiles, C, Cpad;
  int M, splits;
  double sx;            // scale of X
  double sigma;
  int m0, m8, m16, m24, m28; // 1, 2^8, 2^16, 2^24, 2^28 (see halves())
  int flush_tiles;      // read accumulator set 2 back at least every so many tiles (<= FLUSH_TILES: the int32 range)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ long long mad_wide(int a, int b, long long c) {
  long long d;
  asm("mad.wide.s32 %0, %1, %2, %3;\n" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
// The group sums as two exact int64 halves: hi = sum_{g<4} a[g] 2^(8 (3 - g)), lo = the remaining groups likewise, each
// on top of `bias`.  The multipliers arrive in registers (kernel arguments), so that every step is one IMAD.WIDE: with
// literal powers of two ptxas expands each step into four or five 32-bit instructions.
struct WideMul { int m0, m8, m16, m24, m28; };
template <int NB>
__device__ __forceinline__ void halves(const int32_t* a, const WideMul& w, long long bias, long long& hi, long long& lo) {
  hi = mad_wide(a[0], w.m24, mad_wide(a[1], w.m16, mad_wide(a[2], w.m8, mad_wide(a[3], w.m0, bias))));
  if (NB == 8) lo = mad_wide(a[4], w.m24, mad_wide(a[5], w.m16, mad_wide(a[6], w.m8, mad_wide(a[7], w.m0, bias))));
  else lo = mad_wide(a[4], w.m16, mad_wide(a[5], w.m8, mad_wide(a[6], w.m0, bias)));
}
// int64 -> double without the conversion unit (I2F.F64.S64 and F2I.S64.F64 issue at a small fraction of the other pipes'
// rate: with them the kernel's top stall was math-pipe throttle on XU, profiles/glm_i8_v10_ncu_r02.json): the bits of
// 1.5 * 2^52 plus an integer |x| < 2^51 ARE the double 1.5 * 2^52 + x, and fma(that, s, -1.5 * 2^52 * s) = x * s exactly.
constexpr long long MAGIC_BITS = 0x4338000000000000LL;
constexpr double MAGIC = 6755399441055744.0;
// adjoint slices in shared memory, MN-major for product 2 (B operand, N = 32 j + chain contiguous, K = rows):
// 128-byte core matrices [row group r>>3][n chunk n>>4][row r&7][n&15]; SBO (along N) = 128, LBO (along K) = QD_LBO
constexpr int QD_N = NS8 * NC;                       // 224 stacked columns
constexpr uint32_t QD_LBO = (QD_N / 16) * 128;       // 1792
constexpr uint32_t QD_BUF = (TR / 8) * QD_LBO;       // 28672 bytes per buffer
constexpr uint32_t SMEM_V10 = 2 * XT_BYTES + QB_BYTES + 2 * QD_BUF;
constexpr int NE = 512, NTI = 544, NWE = NE / 32;    // epilogue threads, threads with the issuer warp, epilogue warps
constexpr int FLUSH_TILES = 64;                      // 7 pairs x 128 rows x 2^14 < 2^24 per tile and group: 64 tiles < 2^30
constexpr unsigned int E_MIN = 64u, E_MAX = 1900u;

// One CTA = 32 chains x a range of 128-row tiles.  Warps 0..15: epilogues (thread = TMEM lane x chain quarter);
// warp 16: bulk copies of X tiles and every tensor instruction.
//   product 1 (t):   eta slices-sums -> TMEM columns [0, 32 NB)
//   epilogue 1 (t):  w = eta - y in registers, sum w^2, exponent of max |w| per chain -> shared memory
//   barrier (A)      (epilogue threads only) ... the issuer may overwrite accumulator set 1 with product 1 (t+1)
//   slicing (t):     w against the chain's STICKY power-of-two scale -> seven int8 slices, stored MN-major as seven 8-byte words
//   product 2 (t):   ACCUMULATES in TMEM columns [256, 480) over consecutive tiles; it is read back (epilogue 2: one FP64
//                    recombination per group sum) only when a chain's scale has to grow, every FLUSH_TILES tiles (int32 range)
//                    and at the end.  A scale grows to one binade above the tile maximum that exceeded it, so growth is rare
//                    (the running maximum of |w| over rows grows like sqrt(log n)); the cost is at most two of the 56 bits.
template <int NB>      // beta slices: 7 or 8
__global__ void __launch_bounds__(NTI, 1) k_glm_i8(GlmI8Args a) {   // 17 warps are allocated as 20: 96 registers
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2[2], barX[2], barA, barB;
  __shared__ uint32_t tmem_slot;
  __shared__ double s_f[NC];
  __shared__ double4 s_sc[NC];                       // (s 2^-36, s 2^-60 | 2^-68) and -MAGIC times each: recombination of eta
  __shared__ unsigned int s_max[4][NC];              // high word of max |w| per chain, buffer = tile mod 4
  __shared__ int s_flag[4];                          // some chain's tile maximum exceeded its sticky scale
  __shared__ int s_badc[NC];
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = (tid >> 7) & 3;
  const bool issuer = warp == NWE;
  const long long c0 = (long long)blockIdx.x * NC;
  const long long t_begin = a.tiles * blockIdx.y / a.splits, t_end = a.tiles * (blockIdx.y + 1) / a.splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&bar1, 1); mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1);
    mbar_init(&barA, NWE); mbar_init(&barB, NWE);
  }
  if (tid < NC) {
    double m = 0.0;
    int bad = 0;
    for (int p = 0; p < a.M; p++) {
      const double b = fabs(a.beta[(long long)p * a.Cpad + c0 + tid]);
      if (!(b <= 1.7976931348623157e308)) bad = 1;         // NaN / Inf parameters: the chain is out of support
      else m = fmax(m, b);
    }
    double sc, f;
    scale_of8(m, &sc, &f);
    const double s = a.sx * sc;                            // power of two
    const double s36 = s * 0x1p-36, s60 = s * (NB == 8 ? 0x1p-68 : 0x1p-60);
    s_sc[tid] = make_double4(s36, s60, -MAGIC * s36, -MAGIC * s60);
    s_f[tid] = NB == 8 ? f * 256.0 : f;
    s_badc[tid] = bad;
#pragma unroll
    for (int b = 0; b < 4; b++) s_max[b][tid] = 0u;
    if (tid < 4) s_flag[tid] = 0;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NTI) {
    const int c = e & (NC - 1), p = e / NC;
    double b = p < a.M ? a.beta[(long long)p * a.Cpad + c0 + c] : 0.0;
    if (s_badc[c]) b = 0.0;
    if (NB == 8) {
      const unsigned long long w = slice8_word(b, s_f[c]);
      for (int j = 0; j < 8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (7 - j)));
    } else {
      const unsigned long long w = slice7_word(b, s_f[c]);
      for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
    }
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  bool ok = true;

  if (issuer) {
    // ================= issuer warp: waits by the whole warp, bulk copies and tensor instructions by lane 0 =================
    auto issue_p1 = [&](uint32_t xs) {
#pragma unroll
      for (int i = 0; i < NS8; i++) {
        const uint32_t idesc = make_idesc(TR, NC * (NB - i), 0, 0);       // beta slices j <= NB - 1 - i
#pragma unroll
        for (int ks = 0; ks < MP / 32; ks++)
          mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
                 make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar1);
    };
    if (t_begin < t_end) {
      if (lane == 0) {
        bulk_load_tile(sX, a.QX + t_begin * (long long)XT_BYTES, &barX[0]);
        if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, a.QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
      }
      ok = mbar_wait(&barX[0], 0);
      fence_after();
      if (lane == 0) issue_p1(xs0);
      __syncwarp();
    }
    int cnt = -1;                                         // tiles accumulated in accumulator set 2 (-1: nothing yet)
    for (long long t = t_begin; t < t_end && ok; t++) {
      const long long k = t - t_begin;
      const int stage = (int)(k & 1);
      if (t + 1 < t_end) {
        // (C) every epilogue warp is through the FP64 half of this tile's slicing (and long through its reads of accumulator
        // set 1): product 1 of tile t+1 now runs under the integer half of the slicing.  FP64 instructions are starved while the
        // tensor core works (tools/fp64_vs_imma_probe.cu: 14 instead of 2 cycles per DFMA), so the tensor instructions are
        // started where the epilogues have integer work only.
        ok = mbar_wait(&barA, (uint32_t)(k & 1)) && mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
        fence_after();
        if (lane == 0) issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
        __syncwarp();
      }
      ok = ok && mbar_wait(&barB, (uint32_t)(k & 1));     // adjoint slices in place (and accumulator set 2 read when flushing)
      fence_after();
      const int flag = *(volatile int*)&s_flag[k & 3];
      const bool fresh = flag != 0 || cnt < 0 || cnt >= a.flush_tiles;
      if (fresh) cnt = 0;
      cnt++;
      if (lane == 0) {
        const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES, qd = qds + (uint32_t)stage * QD_BUF;
#pragma unroll
        for (int i = 0; i < NS8; i += 2) {
          const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 1);
#pragma unroll
          for (int ks = 0; ks < TR / 32; ks++)
            mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                   make_desc(qd + ks * 4 * QD_LBO, QD_LBO, 128), idesc, ((i | ks) != 0 || !fresh) ? 1u : 0u);
        }
        umma_commit(&bar2[stage]);
      }
      __syncwarp();
      if (t + 2 < t_end) {                                // this tile's X buffer is free once product 2 has completed
        ok = ok && mbar_wait(&bar2[stage], (uint32_t)((k >> 1) & 1));
        if (lane == 0) bulk_load_tile(sX + (size_t)stage * XT_BYTES, a.QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
        __syncwarp();
      }
    }
    if (!ok && lane == 0) *a.status = 2;
  }

  // ================= epilogue warps =================
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int ii = rl >> 6;
  const WideMul wm = {a.m0, a.m8, a.m16, a.m24, a.m28};
  double lacc[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) lacc[c] = 0.0;
  const double neg_inv_s2 = -1.0 / (a.sigma * a.sigma);
  double* part = a.part + (long long)blockIdx.y * (MP + 1) * a.Cpad;
  if (!issuer) {
    double d[CQ];
    unsigned int lim[CQ];                                 // (sticky exponent + 1) << 20 of this thread's eight chains: the high word
                                                          // of the power-of-two scale; the same in every thread of the quarter
#pragma unroll
    for (int c = 0; c < CQ; c++) { d[c] = 0.0; lim[c] = (E_MIN + 1u) << 20; }
    // epilogue 1 of tile t (its product 1 is the k-th completion of bar1): w = eta - y, sum w^2, max |w| -> s_max[k & 3]
    auto epilogue1 = [&](long long t, long long k) -> bool {
      const long long row = t * TR + rl;
      const double yr = row < a.N ? a.y[row] : 0.0;
      if (!mbar_wait(&bar1, (uint32_t)(k & 1))) return false;
      fence_after();
      unsigned int hw[CQ];
      bool over = false;
      // integer phase (runs under product 2 of the previous tile): the group sums of all eight chains -> two int64 halves each,
      // as bit patterns of doubles on top of the magic bias
      long long hi[CQ], lo[CQ];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {                      // four chains at a time (registers)
        int32_t acc[NB][4];
#pragma unroll
        for (int cb = 0; cb < NB; cb++) tmem_ld4_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + hf * 4), acc[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NB];
#pragma unroll
          for (int cb = 0; cb < NB; cb++) g[cb] = acc[cb][v];
          // neighbouring groups pair up exactly in int32 (a group sum is below 2^23: at most 8 pairs x 64 terms x 2^14), the
          // pairs in two int64 halves on top of the magic bias: seven instructions instead of fourteen
          const int b0 = g[0] * 256 + g[1], b1 = g[2] * 256 + g[3], b2 = g[4] * 256 + g[5];
          const int b3 = NB == 8 ? g[6] * 256 + g[7] : g[6];
          hi[u] = mad_wide(b0, wm.m16, mad_wide(b1, wm.m0, MAGIC_BITS));
          lo[u] = mad_wide(b2, NB == 8 ? wm.m16 : wm.m8, mad_wide(b3, wm.m0, MAGIC_BITS));
        }
      }
      // (a scheduling fence: every integer result exists before the first FP64 instruction is issued, so that the compiler
      // does not interleave the two phases)
      asm volatile("" : "+l"(hi[0]), "+l"(hi[1]), "+l"(hi[2]), "+l"(hi[3]), "+l"(hi[4]), "+l"(hi[5]), "+l"(hi[6]), "+l"(hi[7]),
                        "+l"(lo[0]), "+l"(lo[1]), "+l"(lo[2]), "+l"(lo[3]), "+l"(lo[4]), "+l"(lo[5]), "+l"(lo[6]), "+l"(lo[7]));
      // FP64 phase (the tensor core is idle): one rounding per recombination, w = eta - y
#pragma unroll
      for (int u = 0; u < CQ; u++) {
        const double4 sc = s_sc[q * CQ + u];
        const double w = (fma(__longlong_as_double(hi[u]), sc.x, sc.z) + fma(__longlong_as_double(lo[u]), sc.y, sc.w)) - yr;
        d[u] = w;
        lacc[u] = fma(w, w, lacc[u]);
        hw[u] = (unsigned int)__double2hiint(w) & 0x7fffffffu;
        over = over || hw[u] >= lim[u];
      }
      // fast path: no |w| of this warp reaches its chain's sticky scale (the usual case after the first tile).  Otherwise the
      // warp publishes its maxima and raises the flag; warps below their limits need not contribute (their maxima are
      // below the limit the new maximum exceeds).
      if (__any_sync(0xffffffffu, over)) {
#pragma unroll
        for (int u = 0; u < CQ; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
        if (lane == 0) {
          const int b = (int)(k & 3);
#pragma unroll
          for (int u = 0; u < CQ; u++) atomicMax(&s_max[b][q * CQ + u], hw[u]);
          s_flag[b] = 1;
        }
      }
      return true;
    };
    // epilogue 2 (a flush of accumulator set 2): thread = ((ii, parameter), chain quarter); scales = the sticky exponents the
    // products were made with.  The recombined sums go straight into this CTA's block of the partial buffer (rows 0..63:
    // X' dl/deta = -(X' w) / sigma^2): the lanes with ii = 0 first, then -- after a barrier among the epilogue threads -- the
    // lanes with ii = 1, which hold the groups shifted by one slice.  Fixed order, so the result is reproducible.
    bool first_flush = true;
    auto epilogue2 = [&]() {
      double val[CQ];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        int32_t acc[NS8][4];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) tmem_ld4_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + hf * 4), acc[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NS8];
#pragma unroll
          for (int cb = 0; cb < NS8; cb++) g[cb] = acc[cb][v];
          long long hi, lo;
          halves<7>(g, wm, 0ll, hi, lo);
          const double sc = a.sx * __hiloint2double((int)lim[u], 0);                // 2^(E + 1 - 1023)
          val[u] = fma((double)hi, ii ? 0x1p-44 : 0x1p-36, (double)lo * (ii ? 0x1p-68 : 0x1p-60)) * (sc * neg_inv_s2);
        }
      }
      double* gp = part + (long long)(rl & 63) * a.Cpad + c0 + q * CQ;
      if (ii == 0) {
#pragma unroll
        for (int u = 0; u < CQ; u++) gp[u] = first_flush ? val[u] : gp[u] + val[u];
      }
      asm volatile("bar.sync 2, 512;\n" ::: "memory");
      if (ii == 1) {
#pragma unroll
        for (int u = 0; u < CQ; u++) gp[u] += val[u];
      }
      first_flush = false;
    };
    if (t_begin < t_end) ok = epilogue1(t_begin, 0);
    int cnt = -1;
    for (long long t = t_begin; t < t_end && ok; t++) {
      const long long k = t - t_begin;
      const int b = (int)(k & 3), stage = (int)(k & 1);
      fence_before();
      asm volatile("bar.sync 1, 512;\n" ::: "memory");      // (A) maxima and flag of tile t complete (epilogue threads only)
      // buffers of tile t+2: last read (by the issuer) before it issued product 1 of tile t+2, which epilogue 1 of tile t+2
      // needs; next written in epilogue 1 of tile t+2, after barrier (A) of tile t+1
      if (tid < NC) s_max[(k + 2) & 3][tid] = 0u;
      if (tid == NC) s_flag[(k + 2) & 3] = 0;
      const bool fresh = s_flag[b] != 0 || cnt < 0 || cnt >= a.flush_tiles;
      if (fresh) {
        if (cnt > 0) {                                      // read the accumulated products back with the scales they were made with
          if (!mbar_wait(&bar2[stage ^ 1], (uint32_t)(((k - 1) >> 1) & 1))) { ok = false; break; }
          fence_after();
          epilogue2();
        }
#pragma unroll
        for (int u = 0; u < CQ; u++) {
          const unsigned int e = min(s_max[b][q * CQ + u] >> 20, E_MAX);
          if (((e + 1u) << 20) > lim[u]) lim[u] = (e + 2u) << 20;       // one binade of headroom: growth stays rare
        }
        cnt = 0;
      }
      cnt++;
      // this tile's slice buffer was last read by product 2 of tile t-2
      if (k >= 2 && !mbar_wait(&bar2[stage], (uint32_t)(((k - 2) >> 1) & 1))) { ok = false; break; }
      // w 2^54 / scale (|.| < 2^54) as an integer without the conversion unit: two magic-number roundings.
      //   h = round(w f 2^-28) sits in the low mantissa word of fma(w, f 2^-28, MAGIC + bias_h); the remainder w f - h 2^28 is
      //   exact (fma) and below 2^27 in magnitude, l = its rounding to an integer sits in the low word of (remainder + MAGIC
      //   + bias_l).  The biases fold the +0x80..80 of the balanced-byte form (and keep the low part non-negative):
      //   h 2^28 + l + H8 = (h + Hh - 1) 2^28 + (l + Hl + 2^28),  H8 = Hh 2^28 + Hl.
      constexpr unsigned long long Hh = H8 >> 28, Hl = H8 & ((1ull << 28) - 1);
      constexpr double MU = MAGIC + (double)(long long)(Hh - 1), MW = MAGIC + (double)(long long)(Hl + (1ull << 28));
      // FP64 half (the tensor core is idle: product 2 of the previous tile completed under the integer phase of epilogue 1)
      int ul[CQ];
      unsigned int wl[CQ];
#pragma unroll
      for (int u = 0; u < CQ; u++) {
        const double f = __hiloint2double((int)((2100u << 20) - lim[u]), 0);        // 2^54 / scale = 2^(1076 - E)
        const double f28 = __hiloint2double((int)((2072u << 20) - lim[u]), 0);      // f 2^-28
        const double uu = fma(d[u], f28, MU);
        const double hs = fma(uu, -0x1p28, MU * 0x1p28);    // -(uu - MU) 2^28, exact: uu - MU is an integer below 2^27
        const double ww = fma(d[u], f, hs) + MW;
        ul[u] = __double2loint(uu);
        wl[u] = (unsigned int)__double2loint(ww);
      }
      asm volatile("" : "+r"(ul[0]), "+r"(ul[1]), "+r"(ul[2]), "+r"(ul[3]), "+r"(ul[4]), "+r"(ul[5]), "+r"(ul[6]), "+r"(ul[7]),
                        "+r"(wl[0]), "+r"(wl[1]), "+r"(wl[2]), "+r"(wl[3]), "+r"(wl[4]), "+r"(wl[5]), "+r"(wl[6]), "+r"(wl[7]));
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&barA);                    // (C) this warp's FP64 work of the tile is done: product 1 of tile t+1 may start
      // integer half (under product 1 of tile t+1)
      uint32_t xl[CQ], xh[CQ];
#pragma unroll
      for (int u = 0; u < CQ; u++) {
        const unsigned long long t = (unsigned long long)mad_wide(ul[u], wm.m28, (long long)(unsigned long long)wl[u]);
        const unsigned long long w = t ^ H8;                // byte k = slice 6 - k
        xl[u] = (uint32_t)w;
        xh[u] = (uint32_t)(w >> 32);
      }
      // 8 x 8 byte transpose: word pair j = slice j of chains 8q .. 8q+7
      uint8_t* dst = sQD + (size_t)stage * QD_BUF + (uint32_t)(rl >> 3) * QD_LBO + (uint32_t)(rl & 7) * 16 +
                     (uint32_t)(q >> 1) * 128 + (uint32_t)(q & 1) * 8;
      {
        uint32_t lo4[4], hi4[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {                       // h = 0: bytes 0..3 (slices 6..3); h = 1: bytes 4..6 (slices 2..0)
          const uint32_t* x = h ? xh : xl;
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const uint32_t x0 = x[half * 4], x1 = x[half * 4 + 1], x2 = x[half * 4 + 2], x3 = x[half * 4 + 3];
            const uint32_t t0 = __byte_perm(x0, x1, 0x5140), t1 = __byte_perm(x2, x3, 0x5140);
            const uint32_t t2 = __byte_perm(x0, x1, 0x7362), t3 = __byte_perm(x2, x3, 0x7362);
            uint32_t* o = half ? hi4 : lo4;
            o[0] = __byte_perm(t0, t1, 0x5410);
            o[1] = __byte_perm(t0, t1, 0x7632);
            o[2] = __byte_perm(t2, t3, 0x5410);
            o[3] = __byte_perm(t2, t3, 0x7632);
          }
#pragma unroll
          for (int kb = 0; kb < (h ? 3 : 4); kb++) {
            const int j = 6 - (h * 4 + kb);
            *(uint2*)(dst + j * 256) = make_uint2(lo4[kb], hi4[kb]);
          }
        }
      }
      fence_async_smem();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&barB);                    // (B) this warp's slices are visible (and its flush reads are done)
      if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1);      // under product 2 of tile t
    }
    if (ok && t_begin < t_end && cnt > 0) {                 // final flush
      const long long k = t_end - 1 - t_begin;
      if (mbar_wait(&bar2[k & 1], (uint32_t)((k >> 1) & 1))) {
        fence_after();
        epilogue2();
      } else ok = false;
    }
    if (!ok && rl == 0) *a.status = 1;
  }
  // ---- partial block of this row split, in k_glm's layout and conventions:
  //      rows 0..63: X' dl/deta = -(X' w) / sigma^2;  row 64: -sum (w / sigma)^2 / 2 - n (log sigma + log sqrt(2 pi))
  fence_before();
  __syncthreads();
  double* sL = (double*)sX;                               // [chain][128]
  if (!issuer) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  }
  __syncthreads();
  const long long r_begin = t_begin * TR, r_end = t_end * TR < a.N ? t_end * TR : a.N;
  const double nrows = (double)(r_end > r_begin ? r_end - r_begin : 0);
  const double lconst = nrows * (-log(a.sigma) - SMC_LOG_SQRT_2PI);
  if (!issuer) {
    for (int c = warp; c < NC; c += NWE) {                  // fixed-order sum over the 128 rows
      double ssum = 0.0;
#pragma unroll
      for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      if (lane == 0 && c0 + c < a.Cpad) {
        double L = 0.5 * neg_inv_s2 * ssum + lconst;
        if (s_badc[c]) L = -INFINITY;                       // the reduction flags the chain (glm_emit)
        part[(long long)MP * a.Cpad + c0 + c] = L;
      }
    }
  }
  if (t_begin >= t_end) {                                 // a row split without tiles still owns a block of the partial buffer
    for (int e = tid; e < NC * MP; e += NTI) part[(long long)(e / NC) * a.Cpad + c0 + (e & (NC - 1))] = 0.0;
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

// the shapes the kernel is written for (SMC_GLM_I8=0 switches the path off: the statement stays on the FP64 DMMA kernel)
bool smc_glm_i8_usable(const SmcGlmPlan& g, int64_t C) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SMC_GLM_I8"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && g.family == SMC_GLM_NORMAL && g.MP == MP && C >= NC && g.N >= 4LL * TR && g.par > 0.0;
}

int smc_glm_i8_chain_tiles(int64_t C) { return (int)((C + NC - 1) / NC); }
int smc_glm_i8_row_tile() { return TR; }

void smc_glm_i8_release(SmcGlmPlan& g, cudaStream_t st) {
  if (g.QX8) smc_dev_free(g.QX8, st);
  if (g.i8_status) smc_dev_free(g.i8_status, st);
  g.QX8 = nullptr;
  g.i8_status = nullptr;
  g.i8_checked = 0;
  g.i8_declined = 0;
}

// Slices X on the first call (always outside a stream capture: the first evaluation of a model is eager).  A design matrix
// with NaN or Inf cannot be sliced: the plan is marked declined and the statement stays on the FP64 kernel, which turns
// such rows into chains out of support like the reference does.
int smc_glm_i8_prepare(smc_model_s* m, SmcGlmPlan& g) {
  if (g.QX8 || g.i8_declined) return SMC_OK;
  cudaStream_t st = m->ctx->stream;
  const long long tiles = (g.N + TR - 1) / TR;
  unsigned long long* dmax = nullptr;
  SMC_CUDA(smc_dev_alloc(&dmax, sizeof(unsigned long long), st));
  SMC_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), st));
  k_absmax<<<(unsigned)std::min<long long>((g.N * g.MP + 255) / 256, 148LL * 16), 256, 0, st>>>(g.Xrm, g.N * g.MP, dmax);
  m->ctx->launches++;
  unsigned long long hbits = 0;
  SMC_CUDA(smc_copy_sync(&hbits, dmax, sizeof(hbits), cudaMemcpyDeviceToHost, st));
  smc_dev_free(dmax, st);
  double xmax;
  memcpy(&xmax, &hbits, 8);
  if (!(xmax <= 1.7976931348623157e308)) { g.i8_declined = 1; return SMC_OK; }
  double fx;
  scale_of8(xmax, &g.sx8, &fx);
  SMC_CUDA(smc_dev_alloc(&g.QX8, (size_t)tiles * XT_BYTES, st));
  SMC_CUDA(smc_dev_alloc(&g.i8_status, sizeof(int), st));
  SMC_CUDA(cudaMemsetAsync(g.i8_status, 0, sizeof(int), st));
  const long long prow = tiles * TR;
  k_slice_x8<<<(unsigned)((prow * MP + 255) / 256), 256, 0, st>>>(g.Xrm, g.QX8, g.N, prow, fx);
  m->ctx->launches++;
  static SmcOncePerDevice once;
  if (once.first()) {
    SMC_CUDA(cudaFuncSetAttribute(k_glm_i8<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_V10));
    SMC_CUDA(cudaFuncSetAttribute(k_glm_i8<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_V10));
  }
  return SMC_OK;
}

// part: [RS][MP + 1][Cpad]; beta: [M][Cpad]
int smc_glm_i8_launch(smc_model_s* m, SmcGlmPlan& g, const double* beta, int64_t C, int64_t Cpad, int RS, double* part) {
  cudaStream_t st = m->ctx->stream;
  if (!g.QX8) { smc_set_error("GLM (int8 path): launch before prepare"); return SMC_ERR_INVALID; }
  const long long tiles = (g.N + TR - 1) / TR;
  GlmI8Args a;
  a.QX = g.QX8; a.y = g.y; a.beta = beta; a.part = part; a.status = g.i8_status;
  a.N = g.N; a.tiles = tiles; a.C = C; a.Cpad = Cpad; a.M = g.M; a.splits = RS; a.sx = g.sx8; a.sigma = g.par;
  a.m0 = 1; a.m8 = 1 << 8; a.m16 = 1 << 16; a.m24 = 1 << 24; a.m28 = 1 << 28;
  static int flush = -1;     // SMC_GLM_I8_FLUSH=n (tests): a shorter read-back period, so that small problems exercise that path
  if (flush < 0) { const char* e = getenv("SMC_GLM_I8_FLUSH"); flush = e ? std::max(1, std::min(FLUSH_TILES, atoi(e))) : FLUSH_TILES; }
  a.flush_tiles = flush;
  dim3 grid((unsigned)smc_glm_i8_chain_tiles(C), (unsigned)RS);
  static int beta8 = -1;
  if (beta8 < 0) { const char* e = getenv("SMC_GLM_I8_BETA8"); beta8 = (e && e[0] == '1') ? 1 : 0; }
  if (beta8) k_glm_i8<8><<<grid, NTI, SMEM_V10, st>>>(a);
  else k_glm_i8<7><<<grid, NTI, SMEM_V10, st>>>(a);
  static int always = -1;    // SMC_GLM_I8_CHECK=1 (tests, outside stream capture): synchronise and check after every launch
  if (always < 0) { const char* e = getenv("SMC_GLM_I8_CHECK"); always = (e && e[0] == '1') ? 1 : 0; }
  if (!g.i8_checked || always) {      // first (eager) launch: a tensor-core completion that never arrived is an error, not a wrong number
    g.i8_checked = 1;
    int status = 0;
    SMC_CUDA(smc_copy_sync(&status, g.i8_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (status != 0) { smc_set_error("GLM (int8 path): tensor-core pipeline timed out (status %d)", status); return SMC_ERR_CUDA; }
  }
  return SMC_OK;
}
