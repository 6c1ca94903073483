// FP64 issue behaviour on sm_100a with few warps per scheduler (the situation of k_glm_i8's epilogues: 4 warps per SMSP):
// cycles per DFMA warp-instruction and scheduler as a function of independent chains per thread (ILP), warps per scheduler
// and integer instructions mixed in.  One CTA per SM.  Prints JSON lines.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MIX>
__global__ void k_probe(double* out, long long* cyc, int iters, double x, double y, int im) {
  double a[K];
  int b[4] = {(int)threadIdx.x, 3, 5, 7};
#pragma unroll
  for (int k = 0; k < K; k++) a[k] = (double)(threadIdx.x + k);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      a[k] = fma(a[k], x, y);
#pragma unroll
      for (int m = 0; m < MIX; m++) b[(k + m) & 3] = b[(k + m) & 3] * im + (k + m);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(b[0] + b[1] + b[2] + b[3]);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int MIX>
static void run(int threads, double* out, long long* cyc) {
  const int iters = 2000, sms = 148;
  k_probe<K, MIX><<<sms, threads>>>(out, cyc, iters, 1.0000001, 1e-9, 3);
  cudaDeviceSynchronize();
  k_probe<K, MIX><<<sms, threads>>>(out, cyc, iters, 1.0000001, 1e-9, 3);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += (double)h[i] / sms;
  const int wps = threads / 128;
  printf("{\"probe\": \"fp64_issue\", \"ilp\": %d, \"int_per_dfma\": %d, \"warps_per_smsp\": %d, \"cycles_per_dfma_per_smsp\": %.3f, "
         "\"cycles_per_iteration_per_warp\": %.2f, \"err\": \"%s\"}\n",
         K, MIX, wps, avg / ((double)iters * K * wps), avg / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int threads : {128, 256, 512, 1024}) {
    run<1, 0>(threads, out, cyc); run<2, 0>(threads, out, cyc); run<4, 0>(threads, out, cyc); run<8, 0>(threads, out, cyc);
    run<4, 2>(threads, out, cyc); run<4, 4>(threads, out, cyc); run<8, 4>(threads, out, cyc); run<1, 4>(threads, out, cyc);
  }
  return 0;
}
