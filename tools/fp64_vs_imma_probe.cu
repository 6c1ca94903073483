// Do FP64 instructions issue while the 5th-generation tensor core runs tcgen05.mma kind::i8 on the same SM?  (ncu counts the
// two under one "shared" pipe, and in k_glm_i8 the first FP64 instruction of every phase collects most of the stall samples.)
// One CTA per SM: warp 0 (lane 0) issues back-to-back 128 x 256 x 32 int8 MMAs on resident shared-memory operands (or idles),
// warps 1..W run a fixed number of dependent-chain iterations of DFMA (ILP 8) or of IMAD (ILP 8) and clock themselves.
// Prints JSON lines: cycles per instruction and scheduler, with the tensor core idle and busy.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return true;
    if (clock64() - t0 > 2000000000ll) return false;
  }
}

// mode 0: DFMA, mode 1: IMAD (32-bit multiply-add), mode 2: IADD3/LOP3 mix (ALU pipe)
template <int MODE>
__global__ void __launch_bounds__(288) k_probe(int mma_on, int iters, long long* cyc, double* sink, int* status, unsigned* mma_count) {
  extern __shared__ __align__(1024) uint8_t smem[];      // A: 128 x 32 bytes, B: 256 x 32 bytes (contents irrelevant)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < (128 + 256) * 32; e += blockDim.x) smem[e] = (uint8_t)(e * 7 + 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) { mbar_init(&bar, 1); stop = 0; }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (mma_on && lane == 0) {
      const uint32_t idesc = make_idesc(128, 256);
      const uint64_t da = make_desc(smem_u32(smem), 128, 256), db = make_desc(smem_u32(smem + 128 * 32), 128, 256);
      unsigned n = 0, parity = 0;
      while (!stop) {
        for (int k = 0; k < 16; k++) mma_i8(tmem, da, db, idesc, 1u);
        umma_commit(&bar);
        if (!mbar_wait(&bar, parity)) { *status = 1; break; }
        parity ^= 1u;
        n += 16;
      }
      if (blockIdx.x == 0) *mma_count = n;
    }
  } else {
    double a[8];
    int b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = (double)(tid + k); b[k] = tid * 3 + k; }
    const double x = 1.0000001, y = 1e-9;
    const int im = 3 + (int)(cyc[0] & 1);
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (MODE == 0) a[k] = fma(a[k], x, y);
        if (MODE == 1) b[k] = b[k] * im + k;
        if (MODE == 2) b[k] = ((b[k] + im) ^ (b[k] >> 3)) + k;
      }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + (double)b[k];
    sink[blockIdx.x * blockDim.x + tid] = s;
    if (lane == 0) cyc[1 + blockIdx.x * 8 + (warp - 1)] = t1 - t0;
    // the compute warps are done: stop the issuer
    asm volatile("bar.sync 1, 256;\n" ::: "memory");    // warps 1..8
    if (tid == 32) stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(256u) : "memory");
}

template <int MODE>
static void run(const char* name, int mma_on, long long* cyc, double* sink, int* status, unsigned* cnt) {
  const int sms = 148, iters = 4000;
  cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaMemset(status, 0, 4); cudaMemset(cnt, 0, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_probe<MODE><<<sms, 288, 16384>>>(mma_on, iters, cyc, sink, status, cnt);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_probe<MODE><<<sms, 288, 16384>>>(mma_on, iters, cyc, sink, status, cnt);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[1 + 148 * 8];
  int hs; unsigned hc;
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&hs, status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148 * 8; i++) avg += (double)h[1 + i] / (148 * 8);
  // 8 compute warps = 2 per scheduler, 8 instructions per iteration and warp
  printf("{\"probe\": \"fp64_vs_imma\", \"instruction\": \"%s\", \"tensor_core\": \"%s\", \"cycles_per_instruction_per_scheduler\": %.3f, "
         "\"kernel_ms\": %.3f, \"mmas_issued_by_sm0\": %u, \"tensor_busy_fraction_estimate\": %.2f, \"status\": %d, \"err\": \"%s\"}\n",
         name, mma_on ? "busy (back-to-back 128x256x32 int8 MMAs)" : "idle", avg / (iters * 8.0 * 2.0), ms, hc,
         hc * (128.0 * 256 * 32 / 8192.0) / (avg > 0 ? avg : 1), hs, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* cyc; double* sink; int* status; unsigned* cnt;
  cudaMalloc(&cyc, (1 + 148 * 8) * sizeof(long long)); cudaMemset(cyc, 0, (1 + 148 * 8) * sizeof(long long));
  cudaMalloc(&sink, 148 * 288 * sizeof(double)); cudaMalloc(&status, 4); cudaMalloc(&cnt, 4);
  for (int on = 0; on < 2; on++) {
    run<0>("DFMA", on, cyc, sink, status, cnt);
    run<1>("IMAD", on, cyc, sink, status, cnt);
    run<2>("IADD3/LOP3/SHF", on, cyc, sink, status, cnt);
  }
  return 0;
}
