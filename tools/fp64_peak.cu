// FP64 peak microbenchmarks for B200 (sm_100a): DFMA and DMMA (mma.sync f64) issue rates,
// plus a read-only HBM stream.  MEASURED_PEAKS.json holds no FP64 figure, so
// the roofline denominator for the FP64-bound GLM kernel is measured here.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* d, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_mma884(double* out, int iters) {
  double d[ILP][2];
  double a = threadIdx.x * 1e-4, b = 1.0 + threadIdx.x * 1e-5;
#pragma unroll
  for (int i = 0; i < ILP; i++) { d[i][0] = i; d[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) mma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}
template <int ILP, int KK>
__global__ void k_mma16(double* out, int iters) {
  double d[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-4 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-5 * i;
#pragma unroll
  for (int i = 0; i < ILP; i++) { d[i][0] = i; d[i][1] = -i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (KK == 4) mma1684(d[i], a, b[0]);
      if (KK == 8) mma1688(d[i], a, b);
      if (KK == 16) mma16816(d[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}

__global__ void k_read(const double2* __restrict__ p, size_t n2, double* out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 v = __ldg(p + i);
    s += v.x + v.y;
  }
  if (s == 123.456) out[0] = s;
}

template <typename F>
float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  if (cudaGetLastError() != cudaSuccess) return -1.0f;      // e.g. 1024 threads x the kernel's registers exceed the register file
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      int grid = sms * bps;
      float ms = timeit([&] { k_dfma<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
      double tf = 2.0 * 8 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
      printf(",\n \"dfma_t%d_b%d_tflops\": %.3f", threads, bps, tf);
    }
  }
  for (int threads : {128, 256, 512, 1024}) {
    int grid = sms;
    int warps = threads / 32;
    {
      float ms = timeit([&] { k_mma884<8><<<grid, threads>>>(out, iters); });
      double tf = 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
      printf(",\n \"dmma_m8n8k4_ilp8_t%d_tflops\": %.3f", threads, tf);
    }
    {
      float ms = timeit([&] { k_mma884<2><<<grid, threads>>>(out, iters); });
      double tf = 2.0 * 8 * 8 * 4 * 2 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
      printf(",\n \"dmma_m8n8k4_ilp2_t%d_tflops\": %.3f", threads, tf);
    }
    {
      float ms = timeit([&] { k_mma16<4, 4><<<grid, threads>>>(out, iters); });
      double tf = 2.0 * 16 * 8 * 4 * 4 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
      printf(",\n \"dmma_m16n8k4_ilp4_t%d_tflops\": %.3f", threads, tf);
    }
    {
      float ms = timeit([&] { k_mma16<4, 8><<<grid, threads>>>(out, iters); });
      double tf = 2.0 * 16 * 8 * 8 * 4 * iters * (double)grid * warps / (ms * 1e-3) / 1e12;
      printf(",\n \"dmma_m16n8k8_ilp4_t%d_tflops\": %.3f", threads, tf);
    }
    {
      float ms = timeit([&] { k_mma16<4, 16><<<grid, threads>>>(out, iters / 2); });
      double tf = 2.0 * 16 * 8 * 16 * 4 * (iters / 2) * (double)grid * warps / (ms * 1e-3) / 1e12;
      if (ms > 0) printf(",\n \"dmma_m16n8k16_ilp4_t%d_tflops\": %.3f", threads, tf);      // (a configuration that cannot launch is left out)
    }
  }
  {
    size_t bytes = (size_t)4 << 30;
    double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    for (int bps : {4, 8, 16}) {
      float ms = timeit([&] { k_read<<<sms * bps, 512>>>(buf, bytes / 16, out); });
      printf(",\n \"hbm_read_b%d_gbs\": %.1f", bps, bytes / (ms * 1e-3) / 1e9);
    }
    CK(cudaFree(buf));
  }
  printf("\n}\n");
  return 0;
}
