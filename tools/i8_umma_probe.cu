// tcgen05 (UMMA) kind::i8 probe for B200 (sm_100a): the first brick of the sliced-integer contraction that DESIGN.md
// plans for lifting the FP64 tensor bound of k_glm (int8 slice products accumulated exactly in int32 in TMEM).
//
//   check  : one CTA, D[M x N] (int32, TMEM) = A[M x K] * B[N x K]^T from int8 operands staged in shared memory in
//            the no-swizzle canonical core-matrix layout, in the three operand arrangements the GLM kernel needs:
//              kk : A K-major, B K-major    (product 1: eta = X * beta;   X row tile, K = parameters)
//              mk : A MN-major, B K-major   (product 2: G = X' * D from the SAME X tile; K = rows)
//              km : A K-major, B MN-major   (product 2 transposed: chains on the MMA-M side)
//            compared bit for bit with a host integer product.
//   rate   : every SM issues back-to-back 128 x N x 32 MMAs on resident operands -> int8 op/s of the instruction itself
//            (the roofline denominator of a tcgen05 int8 kernel; cuBLASLt int8 GEMM on this pool: 2.9 Pop/s).
//
// Every mbarrier wait is bounded (clock64), so a wrong descriptor yields a reported timeout, not a hung GPU.
// Build: nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tools/i8_umma_probe tools/i8_umma_probe.cu
// Run  : tools/i8_umma_probe check kk|mk|km [N] [swap] |     tools/i8_umma_probe rate [N] [mmas]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

namespace {

constexpr int kM = 128;      // MMA M (rows of D = TMEM lanes)
constexpr int kK = 64;       // contraction length of the check (two K = 32 instructions)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in
// 16-byte units, version 1 (Blackwell), layout type in bits 61..63 (0 = no swizzle).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: D = S32, A = B = signed 8 bit.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
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
// bounded wait: false on time-out
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, long long max_cycles) {
  long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return true;
    if (clock64() - t0 > max_cycles) return false;
  }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// Byte offset of element (mn, k) of an int8 operand tile in the no-swizzle canonical layout.  A core matrix is 128
// contiguous bytes: K-major: 8 mn-rows x 16 k-bytes;  MN-major: 8 k-rows x 16 mn-bytes.  SBO steps along MN, LBO along K.
__host__ __device__ inline uint32_t off_kmajor(int mn, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(mn >> 3) * sbo + (uint32_t)(k >> 4) * lbo + (uint32_t)(mn & 7) * 16 + (uint32_t)(k & 15);
}
__host__ __device__ inline uint32_t off_mnmajor(int mn, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(mn >> 4) * sbo + (uint32_t)(k >> 3) * lbo + (uint32_t)(k & 7) * 16 + (uint32_t)(mn & 15);
}

struct Geometry {           // strides of one operand tile (rows = MN extent, kK deep)
  uint32_t lbo, sbo, kstep; // kstep = descriptor start-address advance for the next K = 32 instruction
};
__host__ __device__ inline Geometry geom(int mn_major, int rows) {
  Geometry g;
  if (!mn_major) { g.lbo = 128; g.sbo = (kK / 16) * 128; g.kstep = 2 * g.lbo; }    // [mn group][k chunk] core matrices
  else { g.sbo = 128; g.lbo = (uint32_t)(rows / 16) * 128; g.kstep = 4 * g.lbo; }  // [k group][mn chunk] core matrices
  return g;
}

// status: 0 ok, 1 = MMA completion never arrived
__global__ void __launch_bounds__(128) k_check(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                               int N, int a_mn, int b_mn, int swap, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;                         // kM x kK bytes
  uint8_t* sB = smem + kM * kK;               // N x kK bytes
  const int tid = threadIdx.x, warp = tid >> 5;
  const Geometry ga = geom(a_mn, kM), gb = geom(b_mn, N);
  for (int e = tid; e < kM * kK; e += blockDim.x) {          // A is [kM][kK] row-major in global memory
    int m = e / kK, k = e % kK;
    sA[a_mn ? off_mnmajor(m, k, ga.lbo, ga.sbo) : off_kmajor(m, k, ga.lbo, ga.sbo)] = (uint8_t)A[e];
  }
  for (int e = tid; e < N * kK; e += blockDim.x) {           // B is [N][kK] row-major
    int n = e / kK, k = e % kK;
    sB[b_mn ? off_mnmajor(n, k, gb.lbo, gb.sbo) : off_kmajor(n, k, gb.lbo, gb.sbo)] = (uint8_t)B[e];
  }
  const uint32_t cols = N < 32 ? 32 : N;      // power of two >= 32 (N in {32, 64, 128, 256})
  if (warp == 0) tmem_alloc(&tmem_slot, cols);
  if (tid == 32) mbar_init(&bar, 1);
  fence_async_smem();                          // generic-proxy smem writes -> visible to the tensor core's async proxy
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(kM, N, a_mn, b_mn);
    for (int ks = 0; ks < kK / 32; ks++) {
      // swap = 1 exchanges the two offsets in the descriptors only (a cross-check of which field steps along K)
      uint64_t da = make_desc(smem_u32(sA) + ks * ga.kstep, swap ? ga.sbo : ga.lbo, swap ? ga.lbo : ga.sbo);
      uint64_t db = make_desc(smem_u32(sB) + ks * gb.kstep, swap ? gb.sbo : gb.lbo, swap ? gb.lbo : gb.sbo);
      mma_i8(tmem, da, db, idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0, 2000000000ll);
  fence_after();
  if (!ok) { if (tid == 0) *status = 1; }
  else {
    const int row = warp * 32 + (tid & 31);    // warp w reads TMEM lanes 32w .. 32w+31
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 8; j++) D[row * N + c0 + j] = (int32_t)v[j];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, cols);
}

// issue-rate kernel: one CTA per SM, `mmas` back-to-back 128 x N x 32 instructions on resident operands,
// alternating between two accumulators so that consecutive instructions are independent
__global__ void __launch_bounds__(128) k_rate(int N, int mmas, int* status, int32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (kM + 256) * 32; e += blockDim.x) smem[e] = (uint8_t)((e * 7 + blockIdx.x) & 3);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) mbar_init(&bar, 1);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(kM, N, 0, 0);
    const uint64_t da = make_desc(smem_u32(smem), 128, 256);                 // K = 32: two k-chunks per 8-row group
    const uint64_t db = make_desc(smem_u32(smem + kM * 32), 128, 256);
    for (int i = 0; i < mmas; i++) mma_i8(tmem + (uint32_t)(i & 1) * 256, da, db, idesc, i > 1);
    umma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0, 8000000000ll);
  fence_after();
  if (!ok && tid == 0) *status = 1;
  if (ok && sink != nullptr) {
    uint32_t v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16), v);
    if (v[0] == 0x7fffffffu) sink[0] = (int32_t)v[1];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int run_check(const char* mode, int N, int swap) {
  const int a_mn = mode[0] == 'm', b_mn = mode[1] == 'm';
  std::vector<int8_t> A(kM * kK), B((size_t)N * kK);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)((int)((s >> 16) % 127) - 63); };   // 7-bit signed slices
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  int8_t *dA, *dB; int32_t* dD; int* dS;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, sizeof(int32_t) * kM * N)); CK(cudaMalloc(&dS, sizeof(int)));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, sizeof(int32_t) * kM * N)); CK(cudaMemset(dS, 0, sizeof(int)));
  size_t shm = (size_t)(kM + N) * kK;
  CK(cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  k_check<<<1, 128, shm>>>(dA, dB, dD, N, a_mn, b_mn, swap, dS);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D((size_t)kM * N); int st = 0;
  CK(cudaMemcpy(D.data(), dD, sizeof(int32_t) * D.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
  long bad = 0; int first = -1;
  for (int m = 0; m < kM; m++)
    for (int n = 0; n < N; n++) {
      int32_t ref = 0;
      for (int k = 0; k < kK; k++) ref += (int32_t)A[m * kK + k] * (int32_t)B[n * kK + k];
      if (ref != D[(size_t)m * N + n]) { if (first < 0) first = m * N + n; bad++; }
    }
  printf("{\"probe\": \"check\", \"mode\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"swap\": %d, \"timeout\": %d, \"mismatches\": %ld, \"first_bad\": %d, "
         "\"d0\": %d, \"bit_exact\": %s}\n", mode, kM, N, kK, swap, st, bad, first, D[0], (st == 0 && bad == 0) ? "true" : "false");
  return 0;
}

int run_rate(int N, int mmas) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int* dS; int32_t* dSink;
  CK(cudaMalloc(&dS, sizeof(int))); CK(cudaMalloc(&dSink, 64)); CK(cudaMemset(dS, 0, sizeof(int)));
  size_t shm = (size_t)(kM + 256) * 32;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    k_rate<<<sms, 128, shm>>>(N, mmas, dS, dSink);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  int st = 0; CK(cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost));
  double ops = 2.0 * kM * N * 32 * (double)mmas * sms;
  printf("{\"probe\": \"rate\", \"M\": %d, \"N\": %d, \"K\": 32, \"mmas_per_sm\": %d, \"sms\": %d, \"ms\": %.4f, \"timeout\": %d, "
         "\"int8_Pops\": %.4f, \"ns_per_mma\": %.2f}\n", kM, N, mmas, sms, best, st, ops / (best * 1e-3) * 1e-15, best * 1e6 / mmas);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "check")) return run_check(argv[2], argc > 3 ? atoi(argv[3]) : 64, argc > 4 ? atoi(argv[4]) : 0);
  if (argc >= 2 && !strcmp(argv[1], "rate")) return run_rate(argc > 2 ? atoi(argv[2]) : 256, argc > 3 ? atoi(argv[3]) : 20000);
  printf("usage: i8_umma_probe check kk|mk|km [N] [swap] | rate [N] [mmas]\n");
  return 2;
}
