// Prototype of the sliced-integer (Ozaki) GLM evaluation on tcgen05 int8 tensor cores for B200 (sm_100a): the plan of
// DESIGN.md section 8 ("What would lift the FP64 tensor bound itself") as a running kernel.  NOT part of the product path
// (libsimplemcmc_cuda.so does not contain it); it exists to establish, on the device, that
//   logp_c = sum_r -1/2 (y_r - x_r . beta_c)^2      and      G_c = X' (y - X beta_c)          (configs[1]: M = 64)
// come out at FP64 accuracy from int8 slice products accumulated exactly in int32 in TMEM, and to time the un-pipelined
// tile loop.  Numerics follow tools/ozaki_study.py: 8 signed 7-bit slices per operand; X sliced once against one global
// power-of-two scale, beta per chain, the adjoint d per (128-row tile, chain); slice pairs (i, j) with i + j <= 9
// (1-based) are kept; pairs of equal i + j share an int32 accumulator.
//
// One CTA = 32 chains x a range of 128-row tiles, 128 threads, all 512 TMEM columns:
//   product 1  eta = X beta : A = X slice i  (K-major,  128 rows x 64 parameters), B = the beta slices j = 1..9-i stacked
//              along N (32 chains each), D at column offset 32 (i-1): column block cb holds the pairs with i + j = cb + 2.
//              8 x 2 instructions (K = 32) instead of 36 x 2.
//   epilogue 1 thread = row: 8 accumulators -> two int64 -> FP64 eta; d = y - eta to shared memory; per chain: sum d^2,
//              max |d| -> scale; d -> 8 int8 slices (K-major, rows = K) in shared memory.
//   product 2  G = X' d     : A = two X slices (i, i+1) stacked along M, read MN-major from the SAME shared-memory tile
//              (TMEM lane = 64 ii + parameter), B = the d slices j = 1..9-i stacked along N, D at column offset
//              256 + 32 (i-1): lanes 0-63 hold i + j = cb + 2, lanes 64-127 hold i + j = cb + 3.  4 x 4 instructions.
//   epilogue 2 thread = (ii, parameter): 8 accumulators -> FP64 -> G accumulators in registers (32 chains).
//
// This file is a lab notebook: every kernel version that was measured is kept so that profiles/ozaki_glm_proto_v*_r01.jsonl
// can be reproduced.  Variant index (last command-line argument) -> kernel:
//   0  v2, 128 threads                         (= v1)            measured   14.6 TFLOP/s FP64-equivalent
//   1  v2, 512 threads, bulk-copy prefetch                       measured   25.7
//   2  v3, 512 threads (overlap, adjoint in registers)           measured   29.4
//   3  v3 + magic int64->double, bit slicing                     measured   28.8
//   4  v4, 256 threads   5  v4, 512 threads (7 x 8-bit slices)   measured   29.8 / 33.4
//   6  v5, 256 threads   7  v5, 512 threads (both products hidden, straight-line epilogue 1)   measured 28.6 / 36.5 at N = 1e6
//   8  v6 (cheaper scales)   9  v7 (magic-number slicing)   10  v8 (issuer warp)   11  v9 (all three)    NOT YET RUN
// The library form of variant 7 is simplemcmc.jl_b200/csrc/glm_i8.cu.
//
// Build: nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tools/ozaki_glm_proto tools/ozaki_glm_proto.cu
// Run  : tools/ozaki_glm_proto check [rows] [chains] [row_splits] [variant]   (host long-double reference, relative errors)
//        tools/ozaki_glm_proto time  [rows] [chains] [row_splits] [variant]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

namespace {

constexpr int NC = 32;        // chains per CTA
constexpr int TR = 128;       // rows per tile (MMA M of product 1, K of product 2)
constexpr int MP = 64;        // parameters
constexpr int NS = 8;         // slices
constexpr uint32_t XT_BYTES = TR * NS * MP;        // 65536: one X tile, all slices
constexpr uint32_t QB_BYTES = NS * NC * MP;        // 16384: stacked beta slices
constexpr uint32_t QD_BYTES = NS * NC * TR;        // 32768: stacked adjoint slices
constexpr uint32_t SD_BYTES = NC * TR * 8;         // 32768: FP64 adjoint staging
constexpr uint32_t SMEM_BYTES = XT_BYTES + QB_BYTES + QD_BYTES + SD_BYTES;

// ---- PTX wrappers (verified by tools/i8_umma_probe.cu) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
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
// bounded wait.  try_wait suspends the warp (hint: up to 2 us) until the phase completes, so waiting warps do not
// compete for issue slots with the thread that feeds the tensor core; `max_cycles` only bounds the number of retries
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, long long max_cycles) {
  const uint32_t addr = smem_u32(bar);
  for (long long it = 0; it < (max_cycles >> 12); it++) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(2000u) : "memory");
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
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// one X tile (64 KB) by the bulk-copy engine, four 16 KB copies completing on `bar`
__device__ __forceinline__ void bulk_load_tile(uint8_t* dst, const int8_t* src, uint64_t* bar) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(b), "r"(65536u) : "memory");
#pragma unroll
  for (int k = 0; k < 4; k++)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst + k * 16384)), "l"(src + k * 16384), "r"(16384u), "r"(b) : "memory");
}

// ---- layouts ----------------------------------------------------------------------------------------------------------
// X tile (all 8 slices of 128 rows x 64 parameters), one image for both products: 128-byte core matrices
// [row group r>>3][slice i][parameter chunk p>>4][row r&7][parameter p&15].
//   product 1 (K-major A, slice i):            start + 512 i, LBO = 128 (next 16 parameters), SBO = 4096 (next 8 rows)
//   product 2 (MN-major A, slices i and i+1):  start + 512 i, SBO = 128 (next 16 of the 128 "rows" (ii, p)), LBO = 4096
__host__ __device__ inline uint32_t xoff(int r, int i, int p) {
  return (uint32_t)(r >> 3) * 4096 + (uint32_t)i * 512 + (uint32_t)(p >> 4) * 128 + (uint32_t)(r & 7) * 16 + (uint32_t)(p & 15);
}
// stacked K-major B operands: row n = 32 j + chain, K bytes per row = kbytes: [n>>3][k>>4][n&7][k&15]
__host__ __device__ inline uint32_t boff(int n, int k, int kbytes) {
  return (uint32_t)(n >> 3) * (uint32_t)(kbytes * 8) + (uint32_t)(k >> 4) * 128 + (uint32_t)(n & 7) * 16 + (uint32_t)(k & 15);
}

// 8 signed base-128 digits of trunc(v * f), f = 2^56 / scale (|v| < scale): the slices of tools/ozaki_study.py
__host__ __device__ inline void slice8(double v, double f, int8_t* q) {
  long long t = (long long)(v * f);
  unsigned long long mag = (unsigned long long)(t < 0 ? -t : t);
  int sgn = t < 0 ? -1 : 1;
#pragma unroll
  for (int j = 0; j < NS; j++) q[j] = (int8_t)(sgn * (int)((mag >> (49 - 7 * j)) & 127));
}
// power-of-two scale strictly above m (from m's exponent field) and the slicing factor 2^56 / scale
__host__ __device__ inline void scale_of(double m, double* scale, double* factor) {
  unsigned long long bits;
  memcpy(&bits, &m, 8);
  unsigned long long E = (bits >> 52) & 0x7ff;
  if (E < 64 || E > 1900) { *scale = 1.0; *factor = 72057594037927936.0; return; }     // zero / denormal / absurd: 2^0, 2^56
  unsigned long long sb = (E + 1) << 52, fb = (2101 - E) << 52;
  memcpy(scale, &sb, 8);
  memcpy(factor, &fb, 8);
}
// sum_cb a[cb] 2^(-7 (cb + g0)) from eight exact int32 group sums (two int64 halves, one rounding)
__device__ __forceinline__ double recombine(const int32_t* a, double w_hi, double w_lo) {
  long long hi = ((long long)a[0] << 21) + ((long long)a[1] << 14) + ((long long)a[2] << 7) + (long long)a[3];
  long long lo = ((long long)a[4] << 21) + ((long long)a[5] << 14) + ((long long)a[6] << 7) + (long long)a[7];
  return fma((double)hi, w_hi, (double)lo * w_lo);
}

// X (row-major [rows][64] FP64) -> int8 slice tiles, once ("bind time")
__global__ void k_slice_x(const double* __restrict__ X, int8_t* __restrict__ QX, long long rows, double factor) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * MP) return;
  long long row = e / MP;
  int p = (int)(e % MP), r = (int)(row % TR);
  int8_t q[NS];
  slice8(X[e], factor, q);
  int8_t* tile = QX + (row / TR) * (long long)XT_BYTES;
  for (int i = 0; i < NS; i++) tile[xoff(r, i, p)] = q[i];
}

// beta: [64][ldc] (chains contiguous); y: [rows]; outL: [splits][ldc]; outG: [splits][64][ldc]
// NT threads = NT / 128 threads per TMEM lane: thread (rl = tid & 127, q = tid >> 7) owns row / lane rl and the chains
// [q CQ, (q + 1) CQ) in the epilogues.  bulk = 1: the next X tile is prefetched by the bulk-copy engine (cp.async.bulk
// completing on an mbarrier) into a second buffer while this tile is computed.
template <int NT>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
            long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
            double* __restrict__ dbg_eta, int bulk) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_L[NC], s_dscale[NC], s_dfactor[NC];
  uint8_t* sX = smem;                                     // one or two X tiles
  uint8_t* sQB = smem + (bulk ? 2 : 1) * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  double* sD = (double*)(sQD + QD_BYTES);                 // [chain][row]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  // ---- setup: TMEM, barriers, beta slices (per-chain scale) ----
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f;
    scale_of(m, &sc, &f);
    s_sb[tid] = sc;
    s_dfactor[tid] = f;                                   // borrowed until the first tile
    s_L[tid] = 0.0;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    int8_t qs[NS];
    slice8(beta[(size_t)p * ldc + c0 + c], s_dfactor[c], qs);
    for (int j = 0; j < NS; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)qs[j];
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD);
  const int ii = rl >> 6;                                 // product-2 lane block of this thread
  double gacc[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) gacc[c] = 0.0;
  uint32_t phase = 0;
  bool ok = true;
  if (bulk && tid == 0 && t_begin < t_end) bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);

  for (long long t = t_begin; t < t_end && ok; t++) {
    const int stage = bulk ? (int)((t - t_begin) & 1) : 0;
    const uint32_t xs = smem_u32(sX + (size_t)stage * XT_BYTES);
    const double yr = y[t * TR + rl];
    if (!bulk) {
      // X tile: flat 64 KB copy (the global image already has the shared-memory layout)
      const uint4* src = (const uint4*)(QX + t * (long long)XT_BYTES);
      uint4* dst = (uint4*)sX;
#pragma unroll 8
      for (int k = 0; k < (int)(XT_BYTES / 16 / NT); k++) dst[k * NT + tid] = __ldg(src + k * NT + tid);
      fence_async_smem();
      fence_before();
      __syncthreads();
    }
    // ---- product 1 ----
    if (tid == 0) {
      if (bulk) {
        // the other buffer was last read by the previous tile's product 2, which has completed (bar2)
        if (t + 1 < t_end) bulk_load_tile(sX + (size_t)(stage ^ 1) * XT_BYTES, QX + (t + 1) * (long long)XT_BYTES, &barX[stage ^ 1]);
        if (!mbar_wait(&barX[stage], (uint32_t)(((t - t_begin) >> 1) & 1), 4000000000ll)) *status = 2;
      }
      fence_after();
#pragma unroll 1
      for (int i = 0; i < NS; i++) {
        const uint32_t idesc = make_idesc(TR, NC * (NS - i), 0, 0);
        for (int ks = 0; ks < MP / 32; ks++)
          mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
                 make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar1);
    }
    ok = mbar_wait(&bar1, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    // ---- epilogue 1: thread = (row, chain group) ----
#pragma unroll 1
    for (int cc = q * CQ; cc < (q + 1) * CQ; cc += 8) {
      int32_t a[NS][8];
#pragma unroll
      for (int cb = 0; cb < NS; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS];
#pragma unroll
        for (int cb = 0; cb < NS; cb++) g[cb] = a[cb][u];
        double eta = recombine(g, 0x1p-35, 0x1p-63) * (sx * s_sb[cc + u]);
        if (dbg_eta != nullptr && blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + cc + u] = eta;
        sD[(cc + u) * TR + rl] = yr - eta;
      }
    }
    __syncthreads();
    // per chain (one warp each): sum of squares and max |d| -> scale of this tile's adjoint
    for (int c = warp; c < NC; c += NW) {
      double ssum = 0.0, m = 0.0;
#pragma unroll
      for (int k = 0; k < TR / 32; k++) { double d = sD[c * TR + k * 32 + lane]; ssum = fma(d, d, ssum); m = fmax(m, fabs(d)); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { ssum += __shfl_xor_sync(0xffffffffu, ssum, o); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o)); }
      if (lane == 0) {
        s_L[c] += -0.5 * ssum;
        double sc, f;
        scale_of(m, &sc, &f);
        s_dscale[c] = sc;
        s_dfactor[c] = f;
      }
    }
    __syncthreads();
    // d -> int8 slices, K-major with K = rows
#pragma unroll 4
    for (int c = q * CQ; c < (q + 1) * CQ; c++) {
      int8_t qs[NS];
      slice8(sD[c * TR + rl], s_dfactor[c], qs);
#pragma unroll
      for (int j = 0; j < NS; j++) sQD[boff(j * NC + c, rl, TR)] = (uint8_t)qs[j];
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    // ---- product 2 ----
    if (tid == 0) {
      fence_after();
#pragma unroll 1
      for (int i = 0; i < NS; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
    }
    ok = mbar_wait(&bar2, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    // ---- epilogue 2: thread = ((ii, parameter), chain group); lanes 64..127 hold the groups shifted by one ----
    const double w_hi = ii ? 0x1p-42 : 0x1p-35, w_lo = ii ? 0x1p-70 : 0x1p-63;
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS][8];
#pragma unroll
      for (int cb = 0; cb < NS; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS];
#pragma unroll
        for (int cb = 0; cb < NS; cb++) g[cb] = a[cb][u];
        gacc[cc + u] = fma(recombine(g, w_hi, w_lo), sx * s_dscale[q * CQ + cc + u], gacc[cc + u]);
      }
    }
    fence_before();
    __syncthreads();
    phase ^= 1;
  }
  if (!ok && rl == 0) *status = 1;
  // ---- write this row split's partial sums: G = lane block 0 + lane block 1 ----
  __syncthreads();
  double* sG = sD;                                        // [chain][64]
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  if (tid < NC) outL[(size_t)blockIdx.y * ldc + c0 + tid] = s_L[tid];
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v3: overlapped schedule ------------------------------------------------------------------------------------------
// Same arithmetic, but (a) the tensor core runs product 1 of tile t+1 while the threads work on epilogue 2 of tile t
// (issue order: product 2 (t), product 1 (t+1); X tiles double-buffered by the bulk-copy engine), (b) the adjoint stays
// in registers between its computation and its slicing: sum d^2 accumulates per thread, max |d| per (tile, chain) goes
// through redux.sync + one shared-memory atomicMax per warp on the high word of |d| (only the exponent is needed), so the
// FP64 staging array, the statistics pass and one CTA barrier per tile are gone; (c) RFP = 1 recombines the eight group
// sums as a chain of I2F.F64 + DFMA (smallest weight first) instead of two int64 halves.
// exact int64 -> double for |v| < 2^51 with integer adds and one DADD (no I2F.F64.S64): bits of 1.5 * 2^52 plus v
__device__ __forceinline__ double i64_to_double(long long v) {
  return __longlong_as_double(v + 0x4338000000000000ll) - 6755399441055744.0;
}
// the digits of slice8(v, 2^56 / 2^(Emax - 1022)) from the bit pattern of v (no DMUL, no F2I.S64.F64)
__device__ __forceinline__ void slice8_bits(double v, unsigned int Emax, int8_t* q) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned int E = (unsigned int)(bits >> 52) & 0x7ffu;
  const unsigned long long mant = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
  const unsigned int sh = Emax - E;                       // |v| <= the maximum the scale was taken from
  const unsigned long long mag = (E == 0u || sh > 63u) ? 0ull : ((mant << 3) >> sh);
  const int sgn = (long long)bits < 0 ? -1 : 1;
#pragma unroll
  for (int j = 0; j < NS; j++) q[j] = (int8_t)(sgn * (int)((mag >> (49 - 7 * j)) & 127));
}

template <int NT, int RFP>
__device__ __forceinline__ double recombine_v3(const int32_t* a, int ii) {
  if (RFP == 2) {
    long long hi = ((long long)a[0] << 21) + ((long long)a[1] << 14) + ((long long)a[2] << 7) + (long long)a[3];
    long long lo = ((long long)a[4] << 21) + ((long long)a[5] << 14) + ((long long)a[6] << 7) + (long long)a[7];
    return fma(i64_to_double(hi), ii ? 0x1p-42 : 0x1p-35, i64_to_double(lo) * (ii ? 0x1p-70 : 0x1p-63));
  }
  if (RFP == 1) {
    const double w7 = ii ? 0x1p-70 : 0x1p-63;
    double acc = (double)a[7] * w7;
    acc = fma((double)a[6], w7 * 0x1p7, acc);
    acc = fma((double)a[5], w7 * 0x1p14, acc);
    acc = fma((double)a[4], w7 * 0x1p21, acc);
    acc = fma((double)a[3], w7 * 0x1p28, acc);
    acc = fma((double)a[2], w7 * 0x1p35, acc);
    acc = fma((double)a[1], w7 * 0x1p42, acc);
    acc = fma((double)a[0], w7 * 0x1p49, acc);
    return acc;
  }
  return recombine(a, ii ? 0x1p-42 : 0x1p-35, ii ? 0x1p-70 : 0x1p-63);
}

template <int NT, int RFP>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm_v3(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[2][NC];
  uint8_t* sX = smem;                                     // two X tiles
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f;
    scale_of(m, &sc, &f);
    s_sb[tid] = sc;
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    int8_t qs[NS];
    slice8(beta[(size_t)p * ldc + c0 + c], s_f[c], qs);
    for (int j = 0; j < NS; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)qs[j];
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  const int ii = rl >> 6;
  double gacc[CQ], lacc[CQ], etascale[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; etascale[c] = sx * s_sb[q * CQ + c]; }
  bool ok = true;

  auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
    for (int i = 0; i < NS; i++) {
      const uint32_t idesc = make_idesc(TR, NC * (NS - i), 0, 0);
      for (int ks = 0; ks < MP / 32; ks++)
        mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
               make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
    }
    umma_commit(&bar1);
  };
  if (tid == 0 && t_begin < t_end) {
    bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
    if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
    if (!mbar_wait(&barX[0], 0, 4000000000ll)) *status = 2;
    fence_after();
    issue_p1(xs0);
  }

  for (long long t = t_begin; t < t_end; t++) {
    const long long k = t - t_begin;
    const int stage = (int)(k & 1);
    const uint32_t phase = (uint32_t)(k & 1);
    const double yr = y[t * TR + rl];
    ok = mbar_wait(&bar1, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    // ---- epilogue 1: eta, d (registers), sum d^2, max |d| ----
    double d[CQ];
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS][8];
#pragma unroll
      for (int cb = 0; cb < NS; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS];
#pragma unroll
        for (int cb = 0; cb < NS; cb++) g[cb] = a[cb][u];
        const double eta = recombine_v3<NT, RFP>(g, 0) * etascale[cc + u];
        if (dbg_eta != nullptr && blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + cc + u] = eta;
        const double dd = yr - eta;
        d[cc + u] = dd;
        lacc[cc + u] = fma(dd, dd, lacc[cc + u]);
        const unsigned int hw = (unsigned int)__double2hiint(fabs(dd));
        const unsigned int wm = __reduce_max_sync(0xffffffffu, hw);
        if (lane == 0) atomicMax(&s_max[stage][q * CQ + cc + u], wm);
      }
    }
    fence_before();
    __syncthreads();                                      // (A) maxima complete; every thread has read accumulator set 1
    if (tid < NC) s_max[stage ^ 1][tid] = 0u;             // last read before (B) of the previous tile, next written after (B)
    double dscale[CQ];
#pragma unroll
    for (int c = 0; c < CQ; c++) {
      unsigned long long E = (s_max[stage][q * CQ + c] >> 20) & 0x7ffu;
      double sc = 1.0, f = 72057594037927936.0;
      if (E >= 64 && E <= 1900) { sc = __longlong_as_double((long long)((E + 1) << 52)); f = __longlong_as_double((long long)((2101 - E) << 52)); }
      else E = 1022;
      dscale[c] = sx * sc;
      int8_t qs[NS];
      if (RFP == 2) slice8_bits(d[c], (unsigned int)E, qs);
      else slice8(d[c], f, qs);
#pragma unroll
      for (int j = 0; j < NS; j++) sQD[boff(j * NC + q * CQ + c, rl, TR)] = (uint8_t)qs[j];
    }
    fence_async_smem();
    fence_before();
    __syncthreads();                                      // (B) adjoint slices visible; accumulator set 2 free (epilogue 2 of t-1 done)
    if (tid == 0) {
      fence_after();
      const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
      for (int i = 0; i < NS; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
      if (t + 1 < t_end) {                                // product 1 of the next tile runs under epilogue 2 of this one
        if (!mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll)) *status = 2;
        fence_after();
        issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
      }
    }
    ok = mbar_wait(&bar2, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    if (tid == 0 && t + 2 < t_end)                        // this tile's X buffer is free: product 2 has completed
      bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
    // ---- epilogue 2 ----
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS][8];
#pragma unroll
      for (int cb = 0; cb < NS; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS];
#pragma unroll
        for (int cb = 0; cb < NS; cb++) g[cb] = a[cb][u];
        gacc[cc + u] = fma(recombine_v3<NT, RFP>(g, ii), dscale[cc + u], gacc[cc + u]);
      }
    }
  }
  if (!ok && rl == 0) *status = 1;
  // ---- partial sums of this row split ----
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;                              // [chain][64] (16 KB of the 32 KB slice buffer)
  double* sL = (double*)sX;                               // [chain][128]
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
#pragma unroll
  for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  for (int c = warp; c < NC; c += NW) {                   // fixed-order sum over the 128 rows
    double ssum = 0.0;
#pragma unroll
    for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v4: seven balanced 8-bit slices --------------------------------------------------------------------------------------
// int8 holds [-128, 127], so a slice can carry 8 bits when the digits are balanced: t = trunc(v 2^54 / scale),
// u = t + H (H = 0x80 in each of the 7 bytes), slice = byte of u minus 128 = byte of (u xor H) read as int8.  Seven slices
// give the same 56 bits as eight 7-bit ones with 28 instead of 36 slice pairs (7 instead of 8 accumulator groups: 14 + 16
// tensor instructions with N <= 224 per tile, 448 TMEM columns), there is no sign handling, and slicing a value is one
// 64-bit add, one xor and seven byte stores.  Schedule as v3.  The X tile keeps 8 slots per row group (slot 7 = zeros) so
// that the stacked pair (6, 7) of product 2 exists.
constexpr int NS8 = 7;
constexpr unsigned long long H8 = 0x0080808080808080ull;
__host__ __device__ inline unsigned long long slice7_word(double v, double f) {       // f = 2^54 / scale
  long long t = llrint(v * f);      // to nearest: truncating X biases every element towards zero (1.5e-13 of max |G| at N = 1e6)
  return ((unsigned long long)t + H8) ^ H8;               // byte k = slice 6 - k (slice 0 = most significant)
}
__device__ __forceinline__ unsigned long long slice7_bits(double v, unsigned int Emax) {   // scale = 2^(Emax - 1022)
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned int E = (unsigned int)(bits >> 52) & 0x7ffu;
  const unsigned long long mant = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
  const unsigned int sh = Emax - E;
  const unsigned long long mag = (E == 0u || sh > 63u) ? 0ull : ((mant << 1) >> sh);
  const unsigned long long t = (long long)bits < 0 ? (0ull - mag) : mag;
  return (t + H8) ^ H8;
}
__host__ __device__ inline void scale_of8(double m, double* scale, double* factor, unsigned int* Emax) {
  unsigned long long bits;
  memcpy(&bits, &m, 8);
  unsigned long long E = (bits >> 52) & 0x7ff;
  if (E < 64 || E > 1900) { *scale = 1.0; *factor = 18014398509481984.0; *Emax = 1022; return; }      // 2^0, 2^54
  unsigned long long sb = (E + 1) << 52, fb = (2099 - E) << 52;
  memcpy(scale, &sb, 8);
  memcpy(factor, &fb, 8);
  *Emax = (unsigned int)E;
}
[[maybe_unused]] __device__ __forceinline__ double i64_to_double_pre(long long v) {       // exact for |v| < 2^51
  return __longlong_as_double(v + 0x4338000000000000ll) - 6755399441055744.0;
}
// sum_cb a[cb] 2^(-12 - 8 (cb + ii)) from seven exact int32 group sums
__device__ __forceinline__ double recombine7(const int32_t* a, int ii) {
  long long hi = ((long long)a[0] << 24) + ((long long)a[1] << 16) + ((long long)a[2] << 8) + (long long)a[3];
  long long lo = ((long long)a[4] << 16) + ((long long)a[5] << 8) + (long long)a[6];
#ifdef OZAKI_MAGIC_I2F      // A/B switch: integer adds + DADD instead of I2F.F64.S64 (XU pipe)
  return fma(i64_to_double_pre(hi), ii ? 0x1p-44 : 0x1p-36, i64_to_double_pre(lo) * (ii ? 0x1p-68 : 0x1p-60));
#else
  return fma((double)hi, ii ? 0x1p-44 : 0x1p-36, (double)lo * (ii ? 0x1p-68 : 0x1p-60));
#endif
}
__global__ void k_slice_x8(const double* __restrict__ X, int8_t* __restrict__ QX, long long rows, double factor) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * MP) return;
  long long row = e / MP;
  int p = (int)(e % MP), r = (int)(row % TR);
  const unsigned long long w = slice7_word(X[e], factor);
  int8_t* tile = QX + (row / TR) * (long long)XT_BYTES;
  for (int i = 0; i < NS8; i++) tile[xoff(r, i, p)] = (int8_t)(w >> (8 * (6 - i)));
  tile[xoff(r, 7, p)] = 0;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm_v4(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[2][NC];
  uint8_t* sX = smem;                                     // two X tiles
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sc;
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  const int ii = rl >> 6;
  double gacc[CQ], lacc[CQ], etascale[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; etascale[c] = sx * s_sb[q * CQ + c]; }
  bool ok = true;

  auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
    for (int i = 0; i < NS8; i++) {
      const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
      for (int ks = 0; ks < MP / 32; ks++)
        mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
               make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
    }
    umma_commit(&bar1);
  };
  if (tid == 0 && t_begin < t_end) {
    bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
    if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
    if (!mbar_wait(&barX[0], 0, 4000000000ll)) *status = 2;
    fence_after();
    issue_p1(xs0);
  }

  for (long long t = t_begin; t < t_end; t++) {
    const long long k = t - t_begin;
    const int stage = (int)(k & 1);
    const uint32_t phase = (uint32_t)(k & 1);
    const double yr = y[t * TR + rl];
    ok = mbar_wait(&bar1, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    // ---- epilogue 1: eta, d (registers), sum d^2, max |d| ----
    double d[CQ];
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        const double eta = recombine7(g, 0) * etascale[cc + u];
        if (dbg_eta != nullptr && blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + cc + u] = eta;
        const double dd = yr - eta;
        d[cc + u] = dd;
        lacc[cc + u] = fma(dd, dd, lacc[cc + u]);
        const unsigned int hw = (unsigned int)__double2hiint(fabs(dd));
        const unsigned int wm = __reduce_max_sync(0xffffffffu, hw);
        if (lane == 0) atomicMax(&s_max[stage][q * CQ + cc + u], wm);
      }
    }
    fence_before();
    __syncthreads();                                      // (A) maxima complete; every thread has read accumulator set 1
    if (tid < NC) s_max[stage ^ 1][tid] = 0u;
    double dscale[CQ];
#pragma unroll
    for (int c = 0; c < CQ; c++) {
      unsigned int E = (s_max[stage][q * CQ + c] >> 20) & 0x7ffu;
      double sc = 1.0;
      if (E >= 64u && E <= 1900u) sc = __longlong_as_double((long long)((unsigned long long)(E + 1u) << 52));
      else E = 1022u;
      dscale[c] = sx * sc;
      const unsigned long long w = slice7_bits(d[c], E);
      const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
      uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);      // slice j at + j * NC rows = j * 4096 bytes
      dst[0 * NC * TR] = (uint8_t)(wh >> 16);
      dst[1 * NC * TR] = (uint8_t)(wh >> 8);
      dst[2 * NC * TR] = (uint8_t)wh;
      dst[3 * NC * TR] = (uint8_t)(wl >> 24);
      dst[4 * NC * TR] = (uint8_t)(wl >> 16);
      dst[5 * NC * TR] = (uint8_t)(wl >> 8);
      dst[6 * NC * TR] = (uint8_t)wl;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();                                      // (B) adjoint slices visible; accumulator set 2 free
    if (tid == 0) {
      fence_after();
      const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
      for (int i = 0; i < NS8; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
      if (t + 1 < t_end) {
        if (!mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll)) *status = 2;
        fence_after();
        issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
      }
    }
    ok = mbar_wait(&bar2, phase, 4000000000ll);
    fence_after();
    if (!ok) break;
    if (tid == 0 && t + 2 < t_end)
      bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
    // ---- epilogue 2 ----
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        gacc[cc + u] = fma(recombine7(g, ii), dscale[cc + u], gacc[cc + u]);
      }
    }
  }
  if (!ok && rl == 0) *status = 1;
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;                              // [chain][64]
  double* sL = (double*)sX;                               // [chain][128]
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
#pragma unroll
  for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  for (int c = warp; c < NC; c += NW) {
    double ssum = 0.0;
#pragma unroll
    for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v5: v4 with both tensor products hidden ----------------------------------------------------------------------------
// Issue order changed so that the CTA never idles on product 2: product 1 of tile t+1 is issued at barrier (A) of tile t
// (its accumulators are free once epilogue 1 has read them) and runs under the slicing of tile t; product 2 of tile t is
// issued at barrier (B) and runs under epilogue 1 of tile t+1, whose results (the adjoint of the next tile) wait in
// registers; epilogue 2 of tile t follows.  Scales are re-derived from shared memory where that saves registers.
template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm_v5(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[2][NC];
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sx * sc;                                  // scale of eta
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  const int ii = rl >> 6;
  double gacc[CQ], lacc[CQ], d[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; d[c] = 0.0; }
  bool ok = true;

  auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
    for (int i = 0; i < NS8; i++) {
      const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
      for (int ks = 0; ks < MP / 32; ks++)
        mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
               make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
    }
    umma_commit(&bar1);
  };
  // epilogue 1 of tile t (its product 1 is the k-th completion of bar1): eta, d -> registers, sum d^2, max |d|
  auto epilogue1 = [&](long long t, long long k) -> bool {
    const double yr = y[t * TR + rl];
    if (!mbar_wait(&bar1, (uint32_t)(k & 1), 4000000000ll)) return false;
    fence_after();
    const int st = (int)(k & 1);
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
      // straight-line code for the eight values (no side effects in between, so the compiler can interleave the chains)
      unsigned int hw[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        const double eta = recombine7(g, 0) * s_sb[q * CQ + cc + u];
        if (DBG) { if (blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + cc + u] = eta; }
        const double dd = yr - eta;
        d[cc + u] = dd;
        lacc[cc + u] = fma(dd, dd, lacc[cc + u]);
        hw[u] = (unsigned int)__double2hiint(fabs(dd));
      }
#pragma unroll
      for (int u = 0; u < 8; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicMax(&s_max[st][q * CQ + cc + u], hw[u]);
      }
    }
    return true;
  };
  if (tid == 0 && t_begin < t_end) {
    bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
    if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
    if (!mbar_wait(&barX[0], 0, 4000000000ll)) *status = 2;
    fence_after();
    issue_p1(xs0);
  }
  if (t_begin < t_end) ok = epilogue1(t_begin, 0);

  for (long long t = t_begin; t < t_end && ok; t++) {
    const long long k = t - t_begin;
    const int stage = (int)(k & 1);
    fence_before();
    __syncthreads();                                      // (A) maxima of tile t complete; accumulator set 1 has been read
    if (tid == 0 && t + 1 < t_end) {                      // product 1 of tile t+1 runs under the slicing of tile t
      if (!mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll)) *status = 2;
      fence_after();
      issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
    }
    if (tid >= NT - NC) s_max[stage ^ 1][tid - (NT - NC)] = 0u;
#pragma unroll
    for (int c = 0; c < CQ; c++) {
      unsigned int E = (s_max[stage][q * CQ + c] >> 20) & 0x7ffu;
      if (E < 64u || E > 1900u) E = 1022u;
      const unsigned long long w = slice7_bits(d[c], E);
      const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
      uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);
      dst[0 * NC * TR] = (uint8_t)(wh >> 16);
      dst[1 * NC * TR] = (uint8_t)(wh >> 8);
      dst[2 * NC * TR] = (uint8_t)wh;
      dst[3 * NC * TR] = (uint8_t)(wl >> 24);
      dst[4 * NC * TR] = (uint8_t)(wl >> 16);
      dst[5 * NC * TR] = (uint8_t)(wl >> 8);
      dst[6 * NC * TR] = (uint8_t)wl;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();                                      // (B) adjoint slices visible; accumulator set 2 free
    if (tid == 0) {
      fence_after();
      const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
      for (int i = 0; i < NS8; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
    }
    if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1);      // under product 2 of tile t
    if (!mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll)) ok = false;
    fence_after();
    if (!ok) break;
    if (tid == 0 && t + 2 < t_end)
      bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
    // ---- epilogue 2 of tile t ----
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        unsigned int E = (s_max[stage][q * CQ + cc + u] >> 20) & 0x7ffu;   // zeroed only after barrier (A) of tile t+1
        double sc = 1.0;
        if (E >= 64u && E <= 1900u) sc = __longlong_as_double((long long)((unsigned long long)(E + 1u) << 52));
        gacc[cc + u] = fma(recombine7(g, ii), sx * sc, gacc[cc + u]);
      }
    }
  }
  if (!ok && rl == 0) *status = 1;
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;
  double* sL = (double*)sX;
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
#pragma unroll
  for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  for (int c = warp; c < NC; c += NW) {
    double ssum = 0.0;
#pragma unroll
    for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v6 (NOT YET RUN: written after the round's GPU minutes were spent) = v5 with cheaper scale handling ------------------
// The per-(tile, chain) maximum is reduced on the exponent field itself, the range checks are dropped (a tile whose adjoint
// is all zeros / denormals slices to zeros with any scale; exponents near the top of the range are out of this model's
// reach) and the scale of epilogue 2, a product of two powers of two, is assembled from the exponents with integer adds
// instead of a DMUL: about 11 instructions less per value and tile in the slicing and in epilogue 2 (of about 590 per
// thread and tile), one more in epilogue 1.
// ---- (text of v5:) issue order changed so that the CTA never idles on product 2: product 1 of tile t+1 is issued at barrier (A) of tile t
// (its accumulators are free once epilogue 1 has read them) and runs under the slicing of tile t; product 2 of tile t is
// issued at barrier (B) and runs under epilogue 1 of tile t+1, whose results (the adjoint of the next tile) wait in
// registers; epilogue 2 of tile t follows.  Scales are re-derived from shared memory where that saves registers.
template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm_v6(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[2][NC];
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sx * sc;                                  // scale of eta
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  const int ii = rl >> 6;
  // exponent field of sx * 2^(E - 1022) = E + 1 + (field(sx) - 1023)
  const unsigned int sx_exp = (((unsigned int)__double2hiint(sx) >> 20) & 0x7ffu) + 1u - 1023u;
  double gacc[CQ], lacc[CQ], d[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; d[c] = 0.0; }
  bool ok = true;

  auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
    for (int i = 0; i < NS8; i++) {
      const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
      for (int ks = 0; ks < MP / 32; ks++)
        mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
               make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
    }
    umma_commit(&bar1);
  };
  // epilogue 1 of tile t (its product 1 is the k-th completion of bar1): eta, d -> registers, sum d^2, max |d|
  auto epilogue1 = [&](long long t, long long k) -> bool {
    const double yr = y[t * TR + rl];
    if (!mbar_wait(&bar1, (uint32_t)(k & 1), 4000000000ll)) return false;
    fence_after();
    const int st = (int)(k & 1);
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
      // straight-line code for the eight values (no side effects in between, so the compiler can interleave the chains)
      unsigned int hw[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        const double eta = recombine7(g, 0) * s_sb[q * CQ + cc + u];
        if (DBG) { if (blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + cc + u] = eta; }
        const double dd = yr - eta;
        d[cc + u] = dd;
        lacc[cc + u] = fma(dd, dd, lacc[cc + u]);
        hw[u] = ((unsigned int)__double2hiint(dd) >> 20) & 0x7ffu;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicMax(&s_max[st][q * CQ + cc + u], hw[u]);
      }
    }
    return true;
  };
  if (tid == 0 && t_begin < t_end) {
    bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
    if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
    if (!mbar_wait(&barX[0], 0, 4000000000ll)) *status = 2;
    fence_after();
    issue_p1(xs0);
  }
  if (t_begin < t_end) ok = epilogue1(t_begin, 0);

  for (long long t = t_begin; t < t_end && ok; t++) {
    const long long k = t - t_begin;
    const int stage = (int)(k & 1);
    fence_before();
    __syncthreads();                                      // (A) maxima of tile t complete; accumulator set 1 has been read
    if (tid == 0 && t + 1 < t_end) {                      // product 1 of tile t+1 runs under the slicing of tile t
      if (!mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll)) *status = 2;
      fence_after();
      issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
    }
    if (tid >= NT - NC) s_max[stage ^ 1][tid - (NT - NC)] = 0u;
#pragma unroll
    for (int c = 0; c < CQ; c++) {
      const unsigned long long w = slice7_bits(d[c], s_max[stage][q * CQ + c]);
      const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
      uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);
      dst[0 * NC * TR] = (uint8_t)(wh >> 16);
      dst[1 * NC * TR] = (uint8_t)(wh >> 8);
      dst[2 * NC * TR] = (uint8_t)wh;
      dst[3 * NC * TR] = (uint8_t)(wl >> 24);
      dst[4 * NC * TR] = (uint8_t)(wl >> 16);
      dst[5 * NC * TR] = (uint8_t)(wl >> 8);
      dst[6 * NC * TR] = (uint8_t)wl;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();                                      // (B) adjoint slices visible; accumulator set 2 free
    if (tid == 0) {
      fence_after();
      const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
      for (int i = 0; i < NS8; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
    }
    if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1);      // under product 2 of tile t
    if (!mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll)) ok = false;
    fence_after();
    if (!ok) break;
    if (tid == 0 && t + 2 < t_end)
      bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
    // ---- epilogue 2 of tile t ----
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        // sx * 2^(E - 1022): both powers of two, so the exponent fields add (zeroed only after barrier (A) of tile t+1)
        const double sc = __hiloint2double((int)((s_max[stage][q * CQ + cc + u] + sx_exp) << 20), 0);
        gacc[cc + u] = fma(recombine7(g, ii), sc, gacc[cc + u]);
      }
    }
  }
  if (!ok && rl == 0) *status = 1;
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;
  double* sL = (double*)sX;
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
#pragma unroll
  for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  for (int c = warp; c < NC; c += NW) {
    double ssum = 0.0;
#pragma unroll
    for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v7 (NOT YET RUN) = v6 with the adjoint sliced by two magic-number roundings instead of bit manipulation ------------
// t = rint(v 2^54 / scale) is assembled from h = rint(v f 2^-27) and l = rint(v f - h 2^27), each read from the low word of
// (x + 1.5 2^52): five FP64 instructions (the FP64 pipe is 7 % busy in v5) and about ten integer ones replace the roughly
// thirty of slice7_bits (exponent extraction, 64-bit funnel shifts, clamps, conditional negation).  The algorithm is
// checked on the CPU (tests/test_ozaki_numerics.py::test_magic_number_slicing).  The exponent of a tile's maximum is
// clamped to >= 64 so that the factors stay finite for an all-zero tile.
__device__ __forceinline__ unsigned long long slice7_magic(double v, unsigned int Emax) {   // scale = 2^(Emax - 1022)
  const double f = __hiloint2double((int)((2099u - Emax) << 20), 0);          // 2^54 / scale
  const double f27 = __hiloint2double((int)((2072u - Emax) << 20), 0);        // 2^27 / scale
  const double a = fma(v, f27, 6755399441055744.0);
  const int h = __double2loint(a);
  const double hd = a - 6755399441055744.0;
  const double r = fma(hd, -134217728.0, v * f);                              // exact
  const int l = __double2loint(r + 6755399441055744.0);
  const long long t = (long long)h * 134217728ll + (long long)l;
  return ((unsigned long long)t + H8) ^ H8;
}
template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
k_ozaki_glm_v7(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int Q = NT / 128, CQ = NC / Q, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[2][NC];
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = tid >> 7;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1); }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sx * sc;                                  // scale of eta
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  const int ii = rl >> 6;
  // exponent field of sx * 2^(E - 1022) = E + 1 + (field(sx) - 1023)
  const unsigned int sx_exp = (((unsigned int)__double2hiint(sx) >> 20) & 0x7ffu) + 1u - 1023u;
  double gacc[CQ], lacc[CQ], d[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; d[c] = 0.0; }
  bool ok = true;

  auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
    for (int i = 0; i < NS8; i++) {
      const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
      for (int ks = 0; ks < MP / 32; ks++)
        mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
               make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
    }
    umma_commit(&bar1);
  };
  // epilogue 1 of tile t (its product 1 is the k-th completion of bar1): eta, d -> registers, sum d^2, max |d|
  auto epilogue1 = [&](long long t, long long k) -> bool {
    const double yr = y[t * TR + rl];
    if (!mbar_wait(&bar1, (uint32_t)(k & 1), 4000000000ll)) return false;
    fence_after();
    const int st = (int)(k & 1);
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
      // straight-line code for the eight values (no side effects in between, so the compiler can interleave the chains)
      unsigned int hw[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        const double eta = recombine7(g, 0) * s_sb[q * CQ + cc + u];
        if (DBG) { if (blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + cc + u] = eta; }
        const double dd = yr - eta;
        d[cc + u] = dd;
        lacc[cc + u] = fma(dd, dd, lacc[cc + u]);
        hw[u] = ((unsigned int)__double2hiint(dd) >> 20) & 0x7ffu;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicMax(&s_max[st][q * CQ + cc + u], hw[u]);
      }
    }
    return true;
  };
  if (tid == 0 && t_begin < t_end) {
    bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
    if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
    if (!mbar_wait(&barX[0], 0, 4000000000ll)) *status = 2;
    fence_after();
    issue_p1(xs0);
  }
  if (t_begin < t_end) ok = epilogue1(t_begin, 0);

  for (long long t = t_begin; t < t_end && ok; t++) {
    const long long k = t - t_begin;
    const int stage = (int)(k & 1);
    fence_before();
    __syncthreads();                                      // (A) maxima of tile t complete; accumulator set 1 has been read
    if (tid == 0 && t + 1 < t_end) {                      // product 1 of tile t+1 runs under the slicing of tile t
      if (!mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll)) *status = 2;
      fence_after();
      issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
    }
    if (tid >= NT - NC) s_max[stage ^ 1][tid - (NT - NC)] = 0u;
#pragma unroll
    for (int c = 0; c < CQ; c++) {
      const unsigned long long w = slice7_magic(d[c], max(s_max[stage][q * CQ + c], 64u));
      const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
      uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);
      dst[0 * NC * TR] = (uint8_t)(wh >> 16);
      dst[1 * NC * TR] = (uint8_t)(wh >> 8);
      dst[2 * NC * TR] = (uint8_t)wh;
      dst[3 * NC * TR] = (uint8_t)(wl >> 24);
      dst[4 * NC * TR] = (uint8_t)(wl >> 16);
      dst[5 * NC * TR] = (uint8_t)(wl >> 8);
      dst[6 * NC * TR] = (uint8_t)wl;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();                                      // (B) adjoint slices visible; accumulator set 2 free
    if (tid == 0) {
      fence_after();
      const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
      for (int i = 0; i < NS8; i += 2) {
        const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
        for (int ks = 0; ks < TR / 32; ks++)
          mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                 make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar2);
    }
    if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1);      // under product 2 of tile t
    if (!mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll)) ok = false;
    fence_after();
    if (!ok) break;
    if (tid == 0 && t + 2 < t_end)
      bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
    // ---- epilogue 2 of tile t ----
#pragma unroll
    for (int cc = 0; cc < CQ; cc += 8) {
      int32_t a[NS8][8];
#pragma unroll
      for (int cb = 0; cb < NS8; cb++) tmem_ld8_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + cc), a[cb]);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int32_t g[NS8];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][u];
        // sx * 2^(E - 1022): both powers of two, so the exponent fields add (zeroed only after barrier (A) of tile t+1)
        const double sc = __hiloint2double((int)((max(s_max[stage][q * CQ + cc + u], 64u) + sx_exp) << 20), 0);
        gacc[cc + u] = fma(recombine7(g, ii), sc, gacc[cc + u]);
      }
    }
  }
  if (!ok && rl == 0) *status = 1;
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;
  double* sL = (double*)sX;
  if (ii == 1) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
  }
#pragma unroll
  for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  __syncthreads();
  if (ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  for (int c = warp; c < NC; c += NW) {
    double ssum = 0.0;
#pragma unroll
    for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v8 (NOT YET RUN) = v5 with a dedicated issuer warp -------------------------------------------------------------------
// In v1-v7 thread 0 feeds the tensor core (about 13 instructions per instruction issued, 30 per tile) AND is one of the
// epilogue threads: its warp reaches every CTA barrier late by the issue work, and 15 warps wait for it twice per tile.
// Here a 17th warp issues the tensor instructions and the bulk copies and does nothing else; the 16 epilogue warps hand
// over through mbarriers (one elected arrive per warp): barA = "accumulator set 1 has been read" (product 1 of the next
// tile may go), barB = "this tile's adjoint slices are in shared memory and accumulator set 2 has been read" (product 2
// may go).  Barrier (A) among the epilogue threads (maxima complete before slicing) becomes a named barrier over 512
// threads, barrier (B) disappears.  544 threads: 120 registers per thread.
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
template <bool DBG>
__global__ void __launch_bounds__(544, 1)
k_ozaki_glm_v8(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int NT = 544, NE = 512, CQ = 8, NW = 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2], barA, barB;
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[3][NC];   // three buffers: without barrier (B) a fast warp reaches epilogue 1 of tile t+1
                                          // while a slow one is still clearing a buffer (see the loop)
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = (tid >> 7) & 3;
  const bool issuer = warp == NW;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1);
    mbar_init(&barA, NW); mbar_init(&barB, NW);
  }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sx * sc;
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
    s_max[2][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  bool ok = true;

  if (issuer) {
    // ================= issuer warp: bulk copies and tensor instructions (lane 0), waits by the whole warp =================
    auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
      for (int i = 0; i < NS8; i++) {
        const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
        for (int ks = 0; ks < MP / 32; ks++)
          mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
                 make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar1);
    };
    if (t_begin < t_end) {
      if (lane == 0) {
        bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
        if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
      }
      ok = mbar_wait(&barX[0], 0, 4000000000ll);
      fence_after();
      if (lane == 0) issue_p1(xs0);
      __syncwarp();
    }
    for (long long t = t_begin; t < t_end && ok; t++) {
      const long long k = t - t_begin;
      const int stage = (int)(k & 1);
      if (t + 1 < t_end) {                                // accumulator set 1 read by every epilogue warp: product 1 (t+1)
        ok = mbar_wait(&barA, (uint32_t)(k & 1), 4000000000ll) &&
             mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll);
        fence_after();
        if (lane == 0) issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
        __syncwarp();
      }
      ok = ok && mbar_wait(&barB, (uint32_t)(k & 1), 4000000000ll);    // adjoint slices in place, accumulator set 2 read
      fence_after();
      if (lane == 0) {
        const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
        for (int i = 0; i < NS8; i += 2) {
          const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
          for (int ks = 0; ks < TR / 32; ks++)
            mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                   make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
        }
        umma_commit(&bar2);
      }
      __syncwarp();
      if (t + 2 < t_end) {                                // this tile's X buffer is free once product 2 has completed
        ok = ok && mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll);
        if (lane == 0) bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
        __syncwarp();
      }
    }
    if (!ok && lane == 0) *status = 2;
  }

  // ================= epilogue warps =================
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int ii = rl >> 6;
  double gacc[CQ], lacc[CQ], d[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; d[c] = 0.0; }
  if (!issuer) {
    auto epilogue1 = [&](long long t, long long k, int st) -> bool {
      const double yr = y[t * TR + rl];
      if (!mbar_wait(&bar1, (uint32_t)(k & 1), 4000000000ll)) return false;
      fence_after();
      unsigned int hw[8];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {                      // four chains at a time: 28 instead of 56 registers of accumulators
        int32_t a[NS8][4];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) tmem_ld4_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + hf * 4), a[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NS8];
#pragma unroll
          for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][v];
          const double eta = recombine7(g, 0) * s_sb[q * CQ + u];
          if (DBG) { if (blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + u] = eta; }
          const double dd = yr - eta;
          d[u] = dd;
          lacc[u] = fma(dd, dd, lacc[u]);
          hw[u] = (unsigned int)__double2hiint(fabs(dd));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicMax(&s_max[st][q * CQ + u], hw[u]);
      }
      return true;
    };
    if (t_begin < t_end) ok = epilogue1(t_begin, 0, 0);
    int cur = 0;                                          // s_max buffer of tile t: (t - t_begin) mod 3
    for (long long t = t_begin; t < t_end && ok; t++, cur = cur == 2 ? 0 : cur + 1) {
      const long long k = t - t_begin;
      const int nxt = cur == 2 ? 0 : cur + 1, nn = nxt == 2 ? 0 : nxt + 1;
      fence_before();
      asm volatile("bar.sync 1, 512;\n" ::: "memory");     // (A) maxima of tile t complete (epilogue threads only)
      if (lane == 0) mbar_arrive(&barA);                    // accumulator set 1 has been read by this warp
      // buffer of tile t+2: last read in epilogue 2 of tile t-1 (before this barrier), next written in epilogue 1 of tile
      // t+2, which every warp runs after barrier (A) of tile t+1
      if (tid >= NE - NC) s_max[nn][tid - (NE - NC)] = 0u;
#pragma unroll
      for (int c = 0; c < CQ; c++) {
        unsigned int E = (s_max[cur][q * CQ + c] >> 20) & 0x7ffu;
        if (E < 64u || E > 1900u) E = 1022u;
        const unsigned long long w = slice7_bits(d[c], E);
        const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
        uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);
        dst[0 * NC * TR] = (uint8_t)(wh >> 16);
        dst[1 * NC * TR] = (uint8_t)(wh >> 8);
        dst[2 * NC * TR] = (uint8_t)wh;
        dst[3 * NC * TR] = (uint8_t)(wl >> 24);
        dst[4 * NC * TR] = (uint8_t)(wl >> 16);
        dst[5 * NC * TR] = (uint8_t)(wl >> 8);
        dst[6 * NC * TR] = (uint8_t)wl;
      }
      fence_async_smem();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&barB);                    // (B) this warp's slices are visible, its epilogue 2 of t-1 is done
      if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1, nxt); // under product 2 of tile t
      if (!mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll)) ok = false;
      fence_after();
      if (!ok) break;
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        int32_t a[NS8][4];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) tmem_ld4_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + hf * 4), a[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NS8];
#pragma unroll
          for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][v];
          unsigned int E = (s_max[cur][q * CQ + u] >> 20) & 0x7ffu;     // cleared only after barrier (A) of tile t+1
          double sc = 1.0;
          if (E >= 64u && E <= 1900u) sc = __longlong_as_double((long long)((unsigned long long)(E + 1u) << 52));
          gacc[u] = fma(recombine7(g, ii), sx * sc, gacc[u]);
        }
      }
    }
    if (!ok && rl == 0) *status = 1;
  }
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;
  double* sL = (double*)sX;
  if (!issuer) {
    if (ii == 1) {
#pragma unroll
      for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
    }
#pragma unroll
    for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  }
  __syncthreads();
  if (!issuer && ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  if (!issuer) {
    for (int c = warp; c < NC; c += NW) {
      double ssum = 0.0;
#pragma unroll
      for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
    }
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- v9 (NOT YET RUN) = v8 (issuer warp) + v6 (exponent-only maxima, integer-built scales) + v7 (magic-number slicing) -----
template <bool DBG>
__global__ void __launch_bounds__(544, 1)
k_ozaki_glm_v9(const int8_t* __restrict__ QX, const double* __restrict__ y, const double* __restrict__ beta, int ldc,
               long long tiles, int splits, double sx, double* __restrict__ outL, double* __restrict__ outG, int* status,
               double* __restrict__ dbg_eta) {
  constexpr int NT = 544, NE = 512, CQ = 8, NW = 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar1, bar2, barX[2], barA, barB;
  __shared__ uint32_t tmem_slot;
  __shared__ double s_sb[NC], s_f[NC];
  __shared__ unsigned int s_max[3][NC];   // three buffers: without barrier (B) a fast warp reaches epilogue 1 of tile t+1
                                          // while a slow one is still clearing a buffer (see the loop)
  uint8_t* sX = smem;
  uint8_t* sQB = smem + 2 * XT_BYTES;
  uint8_t* sQD = sQB + QB_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rl = tid & 127, q = (tid >> 7) & 3;
  const bool issuer = warp == NW;
  const int c0 = blockIdx.x * NC;
  const long long t_begin = tiles * blockIdx.y / splits, t_end = tiles * (blockIdx.y + 1) / splits;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&bar1, 1); mbar_init(&bar2, 1); mbar_init(&barX[0], 1); mbar_init(&barX[1], 1);
    mbar_init(&barA, NW); mbar_init(&barB, NW);
  }
  if (tid < NC) {
    double m = 0.0;
    for (int p = 0; p < MP; p++) m = fmax(m, fabs(beta[(size_t)p * ldc + c0 + tid]));
    double sc, f; unsigned int em;
    scale_of8(m, &sc, &f, &em);
    s_sb[tid] = sx * sc;
    s_f[tid] = f;
    s_max[0][tid] = 0u;
    s_max[1][tid] = 0u;
    s_max[2][tid] = 0u;
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    const unsigned long long w = slice7_word(beta[(size_t)p * ldc + c0 + c], s_f[c]);
    for (int j = 0; j < NS8; j++) sQB[boff(j * NC + c, p, MP)] = (uint8_t)(w >> (8 * (6 - j)));
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t qbs = smem_u32(sQB), qds = smem_u32(sQD), xs0 = smem_u32(sX);
  bool ok = true;

  if (issuer) {
    // ================= issuer warp: bulk copies and tensor instructions (lane 0), waits by the whole warp =================
    auto issue_p1 = [&](uint32_t xs) {
#pragma unroll 1
      for (int i = 0; i < NS8; i++) {
        const uint32_t idesc = make_idesc(TR, NC * (NS8 - i), 0, 0);
        for (int ks = 0; ks < MP / 32; ks++)
          mma_i8(tmem + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 256, 128, 4096),
                 make_desc(qbs + ks * 256, 128, MP * 8), idesc, (i | ks) != 0);
      }
      umma_commit(&bar1);
    };
    if (t_begin < t_end) {
      if (lane == 0) {
        bulk_load_tile(sX, QX + t_begin * (long long)XT_BYTES, &barX[0]);
        if (t_begin + 1 < t_end) bulk_load_tile(sX + XT_BYTES, QX + (t_begin + 1) * (long long)XT_BYTES, &barX[1]);
      }
      ok = mbar_wait(&barX[0], 0, 4000000000ll);
      fence_after();
      if (lane == 0) issue_p1(xs0);
      __syncwarp();
    }
    for (long long t = t_begin; t < t_end && ok; t++) {
      const long long k = t - t_begin;
      const int stage = (int)(k & 1);
      if (t + 1 < t_end) {                                // accumulator set 1 read by every epilogue warp: product 1 (t+1)
        ok = mbar_wait(&barA, (uint32_t)(k & 1), 4000000000ll) &&
             mbar_wait(&barX[stage ^ 1], (uint32_t)(((k + 1) >> 1) & 1), 4000000000ll);
        fence_after();
        if (lane == 0) issue_p1(xs0 + (uint32_t)(stage ^ 1) * XT_BYTES);
        __syncwarp();
      }
      ok = ok && mbar_wait(&barB, (uint32_t)(k & 1), 4000000000ll);    // adjoint slices in place, accumulator set 2 read
      fence_after();
      if (lane == 0) {
        const uint32_t xs = xs0 + (uint32_t)stage * XT_BYTES;
#pragma unroll 1
        for (int i = 0; i < NS8; i += 2) {
          const uint32_t idesc = make_idesc(128, NC * (NS8 - i), 1, 0);
          for (int ks = 0; ks < TR / 32; ks++)
            mma_i8(tmem + 256u + (uint32_t)(i * NC), make_desc(xs + i * 512 + ks * 16384, 4096, 128),
                   make_desc(qds + ks * 256, 128, TR * 8), idesc, (i | ks) != 0);
        }
        umma_commit(&bar2);
      }
      __syncwarp();
      if (t + 2 < t_end) {                                // this tile's X buffer is free once product 2 has completed
        ok = ok && mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll);
        if (lane == 0) bulk_load_tile(sX + (size_t)stage * XT_BYTES, QX + (t + 2) * (long long)XT_BYTES, &barX[stage]);
        __syncwarp();
      }
    }
    if (!ok && lane == 0) *status = 2;
  }

  // ================= epilogue warps =================
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int ii = rl >> 6;
  const unsigned int sx_exp = (((unsigned int)__double2hiint(sx) >> 20) & 0x7ffu) + 1u - 1023u;
  double gacc[CQ], lacc[CQ], d[CQ];
#pragma unroll
  for (int c = 0; c < CQ; c++) { gacc[c] = 0.0; lacc[c] = 0.0; d[c] = 0.0; }
  if (!issuer) {
    auto epilogue1 = [&](long long t, long long k, int st) -> bool {
      const double yr = y[t * TR + rl];
      if (!mbar_wait(&bar1, (uint32_t)(k & 1), 4000000000ll)) return false;
      fence_after();
      unsigned int hw[8];
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {                      // four chains at a time: 28 instead of 56 registers of accumulators
        int32_t a[NS8][4];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) tmem_ld4_nowait(tmem + lane_base + (uint32_t)(cb * NC + q * CQ + hf * 4), a[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NS8];
#pragma unroll
          for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][v];
          const double eta = recombine7(g, 0) * s_sb[q * CQ + u];
          if (DBG) { if (blockIdx.y == 0 && t == t_begin) dbg_eta[(size_t)rl * ldc + c0 + q * CQ + u] = eta; }
          const double dd = yr - eta;
          d[u] = dd;
          lacc[u] = fma(dd, dd, lacc[u]);
          hw[u] = ((unsigned int)__double2hiint(dd) >> 20) & 0x7ffu;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) hw[u] = __reduce_max_sync(0xffffffffu, hw[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicMax(&s_max[st][q * CQ + u], hw[u]);
      }
      return true;
    };
    if (t_begin < t_end) ok = epilogue1(t_begin, 0, 0);
    int cur = 0;                                          // s_max buffer of tile t: (t - t_begin) mod 3
    for (long long t = t_begin; t < t_end && ok; t++, cur = cur == 2 ? 0 : cur + 1) {
      const long long k = t - t_begin;
      const int nxt = cur == 2 ? 0 : cur + 1, nn = nxt == 2 ? 0 : nxt + 1;
      fence_before();
      asm volatile("bar.sync 1, 512;\n" ::: "memory");     // (A) maxima of tile t complete (epilogue threads only)
      if (lane == 0) mbar_arrive(&barA);                    // accumulator set 1 has been read by this warp
      // buffer of tile t+2: last read in epilogue 2 of tile t-1 (before this barrier), next written in epilogue 1 of tile
      // t+2, which every warp runs after barrier (A) of tile t+1
      if (tid >= NE - NC) s_max[nn][tid - (NE - NC)] = 0u;
#pragma unroll
      for (int c = 0; c < CQ; c++) {
        const unsigned long long w = slice7_magic(d[c], max(s_max[cur][q * CQ + c], 64u));
        const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
        uint8_t* dst = sQD + boff(q * CQ + c, rl, TR);
        dst[0 * NC * TR] = (uint8_t)(wh >> 16);
        dst[1 * NC * TR] = (uint8_t)(wh >> 8);
        dst[2 * NC * TR] = (uint8_t)wh;
        dst[3 * NC * TR] = (uint8_t)(wl >> 24);
        dst[4 * NC * TR] = (uint8_t)(wl >> 16);
        dst[5 * NC * TR] = (uint8_t)(wl >> 8);
        dst[6 * NC * TR] = (uint8_t)wl;
      }
      fence_async_smem();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&barB);                    // (B) this warp's slices are visible, its epilogue 2 of t-1 is done
      if (t + 1 < t_end) ok = epilogue1(t + 1, k + 1, nxt); // under product 2 of tile t
      if (!mbar_wait(&bar2, (uint32_t)(k & 1), 4000000000ll)) ok = false;
      fence_after();
      if (!ok) break;
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        int32_t a[NS8][4];
#pragma unroll
        for (int cb = 0; cb < NS8; cb++) tmem_ld4_nowait(tmem + lane_base + 256u + (uint32_t)(cb * NC + q * CQ + hf * 4), a[cb]);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const int u = hf * 4 + v;
          int32_t g[NS8];
#pragma unroll
          for (int cb = 0; cb < NS8; cb++) g[cb] = a[cb][v];
          // sx * 2^(E - 1022) from the exponents (the buffer is cleared only after barrier (A) of tile t+1)
          const double sc = __hiloint2double((int)((max(s_max[cur][q * CQ + u], 64u) + sx_exp) << 20), 0);
          gacc[u] = fma(recombine7(g, ii), sc, gacc[u]);
        }
      }
    }
    if (!ok && rl == 0) *status = 1;
  }
  fence_before();
  __syncthreads();
  double* sG = (double*)sQD;
  double* sL = (double*)sX;
  if (!issuer) {
    if (ii == 1) {
#pragma unroll
      for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + (rl & 63)] = gacc[c];
    }
#pragma unroll
    for (int c = 0; c < CQ; c++) sL[(q * CQ + c) * TR + rl] = lacc[c];
  }
  __syncthreads();
  if (!issuer && ii == 0) {
#pragma unroll
    for (int c = 0; c < CQ; c++) sG[(q * CQ + c) * MP + rl] += gacc[c];
  }
  if (!issuer) {
    for (int c = warp; c < NC; c += NW) {
      double ssum = 0.0;
#pragma unroll
      for (int kk = 0; kk < TR / 32; kk++) ssum += sL[c * TR + kk * 32 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      if (lane == 0) outL[(size_t)blockIdx.y * ldc + c0 + c] = -0.5 * ssum;
    }
  }
  __syncthreads();
  for (int e = tid; e < NC * MP; e += NT) {
    int c = e & (NC - 1), p = e / NC;
    outG[((size_t)blockIdx.y * MP + p) * ldc + c0 + c] = sG[c * MP + p];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------------------------------------------------
struct Rng {
  uint64_t s;
  double uni() { s = s * 6364136223846793005ull + 1442695040888963407ull; return ((s >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double normal() { double u = uni(), v = uni(); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }
};

struct Buffers {
  const int8_t* QX8; double sx8;
  const int8_t* QX; const double *Y, *B; double *L, *G, *eta; int* status;
};

template <int NT>
int launch_variant(const Buffers& b, int bulk, dim3 grid, int ldc, long long tiles, int splits, double sx, int reps, float* best_ms) {
  const size_t shm = SMEM_BYTES + (bulk ? XT_BYTES : 0);
  CK(cudaFuncSetAttribute(k_ozaki_glm<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    k_ozaki_glm<NT><<<grid, NT, shm>>>(b.QX, b.Y, b.B, ldc, tiles, splits, sx, b.L, b.G, b.status, b.eta, bulk);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

template <int NT, int RFP>
int launch_v3(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, double sx, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v3<NT, RFP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    k_ozaki_glm_v3<NT, RFP><<<grid, NT, shm>>>(b.QX, b.Y, b.B, ldc, tiles, splits, sx, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

template <int NT>
int launch_v4(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v4<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    k_ozaki_glm_v4<NT><<<grid, NT, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

template <int NT>
int launch_v5(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v5<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  CK(cudaFuncSetAttribute(k_ozaki_glm_v5<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    if (b.eta != nullptr) k_ozaki_glm_v5<NT, true><<<grid, NT, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    else k_ozaki_glm_v5<NT, false><<<grid, NT, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

int launch_v9(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v9<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  CK(cudaFuncSetAttribute(k_ozaki_glm_v9<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    if (b.eta != nullptr) k_ozaki_glm_v9<true><<<grid, 544, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    else k_ozaki_glm_v9<false><<<grid, 544, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

int launch_v8(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  CK(cudaFuncSetAttribute(k_ozaki_glm_v8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    if (b.eta != nullptr) k_ozaki_glm_v8<true><<<grid, 544, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    else k_ozaki_glm_v8<false><<<grid, 544, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

int launch_v7(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v7<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  CK(cudaFuncSetAttribute(k_ozaki_glm_v7<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    if (b.eta != nullptr) k_ozaki_glm_v7<512, true><<<grid, 512, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    else k_ozaki_glm_v7<512, false><<<grid, 512, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

int launch_v6(const Buffers& b, dim3 grid, int ldc, long long tiles, int splits, int reps, float* best_ms) {
  const size_t shm = 2 * XT_BYTES + QB_BYTES + QD_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki_glm_v6<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  CK(cudaFuncSetAttribute(k_ozaki_glm_v6<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < reps; rep++) {
    CK(cudaEventRecord(e0));
    if (b.eta != nullptr) k_ozaki_glm_v6<512, true><<<grid, 512, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    else k_ozaki_glm_v6<512, false><<<grid, 512, shm>>>(b.QX8, b.Y, b.B, ldc, tiles, splits, b.sx8, b.L, b.G, b.status, b.eta);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (reps == 1 || rep > 0) best = fminf(best, ms);
  }
  *best_ms = best;
  return 0;
}

int run(bool check, long long rows, int chains, int splits, int only) {
  const long long tiles = (rows + TR - 1) / TR, prow = tiles * TR;
  const int ldc = (chains + NC - 1) / NC * NC;
  Rng rng{12345};
  std::vector<double> X((size_t)prow * MP, 0.0), Y((size_t)prow, 0.0), B((size_t)MP * ldc, 0.0), b0(MP);
  for (int p = 0; p < MP; p++) b0[p] = rng.normal();
  double xmax = 1.0;
  for (long long r = 0; r < rows; r++) {
    double eta = 0.0;
    for (int p = 0; p < MP; p++) {
      double v = p == 0 ? 1.0 : rng.normal();
      X[(size_t)r * MP + p] = v;
      eta += v * b0[p];
      xmax = fmax(xmax, fabs(v));
    }
    Y[(size_t)r] = eta + rng.normal();
  }
  for (int p = 0; p < MP; p++)
    for (int c = 0; c < chains; c++) B[(size_t)p * ldc + c] = b0[p] + 1e-3 * rng.normal();      // chains near the mode
  double sx, fx;
  scale_of(xmax, &sx, &fx);

  double *dX, *dY, *dB, *dL, *dG, *dEta = nullptr; int8_t* dQX; int* dS;
  CK(cudaMalloc(&dX, X.size() * 8)); CK(cudaMalloc(&dY, Y.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8));
  CK(cudaMalloc(&dQX, (size_t)tiles * XT_BYTES)); CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMalloc(&dL, (size_t)splits * ldc * 8)); CK(cudaMalloc(&dG, (size_t)splits * MP * ldc * 8));
  if (check) CK(cudaMalloc(&dEta, (size_t)TR * ldc * 8));
  CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dY, Y.data(), Y.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  k_slice_x<<<(unsigned)((prow * MP + 255) / 256), 256>>>(dX, dQX, prow, fx);
  CK(cudaGetLastError());
  int8_t* dQX8; double sx8, fx8; unsigned int emx;
  scale_of8(xmax, &sx8, &fx8, &emx);
  CK(cudaMalloc(&dQX8, (size_t)tiles * XT_BYTES));
  k_slice_x8<<<(unsigned)((prow * MP + 255) / 256), 256>>>(dX, dQX8, prow, fx8);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  dim3 grid(ldc / NC, splits);
  Buffers bufs{dQX8, sx8, dQX, dY, dB, dL, dG, dEta, dS};

  // reference in long double (check only); errors relative to max |.| per chain (the parity metric of tests/test_gpu_parity.py)
  std::vector<long double> Lref, Gref, eta_ref;
  double err_G64 = 0, err_L64 = 0;
  if (check) {
    Lref.assign(chains, 0.0L); Gref.assign((size_t)MP * chains, 0.0L); eta_ref.assign((size_t)TR * chains, 0.0L);
    for (int c = 0; c < chains; c++) {
      std::vector<double> g64(MP, 0.0); double L64 = 0.0;
      for (long long r = 0; r < rows; r++) {
        long double eta = 0.0L; double e64 = 0.0;
        for (int p = 0; p < MP; p++) {
          eta += (long double)X[(size_t)r * MP + p] * (long double)B[(size_t)p * ldc + c];
          e64 = fma(X[(size_t)r * MP + p], B[(size_t)p * ldc + c], e64);
        }
        long double d = (long double)Y[(size_t)r] - eta; double d64 = Y[(size_t)r] - e64;
        Lref[c] += -0.5L * d * d; L64 += -0.5 * d64 * d64;
        for (int p = 0; p < MP; p++) { Gref[(size_t)p * chains + c] += (long double)X[(size_t)r * MP + p] * d; g64[p] = fma(X[(size_t)r * MP + p], d64, g64[p]); }
        if (r < TR) eta_ref[(size_t)r * chains + c] = eta;
      }
      long double gmax = 0.0L, gerr64 = 0.0L;
      for (int p = 0; p < MP; p++) { gmax = fmaxl(gmax, fabsl(Gref[(size_t)p * chains + c])); gerr64 = fmaxl(gerr64, fabsl((long double)g64[p] - Gref[(size_t)p * chains + c])); }
      err_G64 = fmax(err_G64, (double)(gerr64 / gmax));
      err_L64 = fmax(err_L64, (double)fabsl(((long double)L64 - Lref[c]) / Lref[c]));
    }
  }

  const double flop = 4.0 * (double)rows * MP * chains;
  // {threads, bulk prefetch (v2) or FP64 recombination (v3), kernel version}
  const int variants[12][3] = {{128, 0, 2}, {512, 1, 2}, {512, 0, 3}, {512, 2, 3}, {256, 0, 4}, {512, 0, 4}, {256, 0, 5}, {512, 0, 5},
                               {512, 0, 6}, {512, 0, 7}, {544, 0, 8}, {544, 0, 9}};   // 8..11: v6, v7, v8, v9 -- not yet run on a device
  for (int v = 0; v < 12; v++) {
    if (only >= 0 && v != only) continue;
    const int NT = variants[v][0], bulk = variants[v][1], ver = variants[v][2];
    CK(cudaMemset(dS, 0, sizeof(int)));
    CK(cudaMemset(dL, 0xff, (size_t)splits * ldc * 8)); CK(cudaMemset(dG, 0xff, (size_t)splits * MP * ldc * 8));
    if (check) CK(cudaMemset(dEta, 0, (size_t)TR * ldc * 8));
    float best = 0;
    const int reps = check ? 1 : 3;
    int rc;
    if (ver == 2) rc = NT == 128 ? launch_variant<128>(bufs, bulk, grid, ldc, tiles, splits, sx, reps, &best)
                                 : launch_variant<512>(bufs, bulk, grid, ldc, tiles, splits, sx, reps, &best);
    else if (ver == 9) rc = launch_v9(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (ver == 8) rc = launch_v8(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (ver == 7) rc = launch_v7(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (ver == 6) rc = launch_v6(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (ver == 5) rc = NT == 256 ? launch_v5<256>(bufs, grid, ldc, tiles, splits, reps, &best) : launch_v5<512>(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (ver == 4) rc = NT == 256 ? launch_v4<256>(bufs, grid, ldc, tiles, splits, reps, &best) : launch_v4<512>(bufs, grid, ldc, tiles, splits, reps, &best);
    else if (NT == 256) rc = launch_v3<256, 0>(bufs, grid, ldc, tiles, splits, sx, reps, &best);
    else rc = bulk == 2 ? launch_v3<512, 2>(bufs, grid, ldc, tiles, splits, sx, reps, &best)
            : bulk == 1 ? launch_v3<512, 1>(bufs, grid, ldc, tiles, splits, sx, reps, &best)
                        : launch_v3<512, 0>(bufs, grid, ldc, tiles, splits, sx, reps, &best);
    if (rc) return rc;
    int st = 0; CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<double> L((size_t)splits * ldc), G((size_t)splits * MP * ldc);
    CK(cudaMemcpy(L.data(), dL, L.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G.data(), dG, G.size() * 8, cudaMemcpyDeviceToHost));
    if (!check) {
      double l0 = 0; for (int s2 = 0; s2 < splits; s2++) l0 += L[(size_t)s2 * ldc];
      printf("{\"probe\": \"ozaki_glm_time\", \"version\": %d, \"threads\": %d, \"flag\": %d, \"rows\": %lld, \"chains\": %d, \"row_splits\": %d, \"ctas\": %d, "
             "\"ms\": %.4f, \"timeout\": %d, \"fp64_equiv_TFLOPs\": %.3f, \"us_per_tile\": %.3f, \"logp0\": %.17g}\n", ver, NT, bulk, rows, chains, splits,
             (int)(grid.x * grid.y), best, st, flop / (best * 1e-3) * 1e-12,
             best * 1e3 / ((double)tiles * grid.x / 148.0), l0);
      continue;
    }
    std::vector<double> eta_dev((size_t)TR * ldc);
    CK(cudaMemcpy(eta_dev.data(), dEta, eta_dev.size() * 8, cudaMemcpyDeviceToHost));
    double err_eta = 0, err_L = 0, err_G = 0;
    for (int c = 0; c < chains; c++) {
      long double eta_max = 0.0L, eta_err = 0.0L;
      for (int r = 0; r < TR && r < rows; r++) {
        eta_max = fmaxl(eta_max, fabsl(eta_ref[(size_t)r * chains + c]));
        eta_err = fmaxl(eta_err, fabsl((long double)eta_dev[(size_t)r * ldc + c] - eta_ref[(size_t)r * chains + c]));
      }
      long double Ld = 0.0L, gmax = 0.0L, gerr = 0.0L;
      for (int s2 = 0; s2 < splits; s2++) Ld += L[(size_t)s2 * ldc + c];
      for (int p = 0; p < MP; p++) {
        long double gd = 0.0L;
        for (int s2 = 0; s2 < splits; s2++) gd += G[((size_t)s2 * MP + p) * ldc + c];
        gmax = fmaxl(gmax, fabsl(Gref[(size_t)p * chains + c])); gerr = fmaxl(gerr, fabsl(gd - Gref[(size_t)p * chains + c]));
      }
      double e1 = (double)(eta_err / eta_max), e2 = (double)fabsl((Ld - Lref[c]) / Lref[c]), e3 = (double)(gerr / gmax);
      err_eta = e1 > err_eta || e1 != e1 ? e1 : err_eta; err_L = e2 > err_L || e2 != e2 ? e2 : err_L; err_G = e3 > err_G || e3 != e3 ? e3 : err_G;
    }
    printf("{\"probe\": \"ozaki_glm_check\", \"version\": %d, \"threads\": %d, \"flag\": %d, \"rows\": %lld, \"chains\": %d, \"row_splits\": %d, \"timeout\": %d, "
           "\"ms\": %.4f, \"relerr_eta_tile0\": %.3e, \"relerr_logp\": %.3e, \"relerr_grad\": %.3e, \"fp64_host_relerr_logp\": %.3e, "
           "\"fp64_host_relerr_grad\": %.3e, \"meets_1e-12\": %s}\n", ver, NT, bulk, rows, chains, splits, st, best, err_eta, err_L, err_G, err_L64, err_G64,
           (st == 0 && err_L <= 1e-12 && err_G <= 1e-12) ? "true" : "false");
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "check"))
    return run(true, argc > 2 ? atoll(argv[2]) : 16384, argc > 3 ? atoi(argv[3]) : 64, argc > 4 ? atoi(argv[4]) : 3, argc > 5 ? atoi(argv[5]) : -1);
  if (argc >= 2 && !strcmp(argv[1], "time"))
    return run(false, argc > 2 ? atoll(argv[2]) : 1000000, argc > 3 ? atoi(argv[3]) : 4096, argc > 4 ? atoi(argv[4]) : 37, argc > 5 ? atoi(argv[5]) : -1);
  printf("usage: ozaki_glm_proto check [rows] [chains] [row_splits] [variant] | time [rows] [chains] [row_splits] [variant]\n");
  return 2;
}
