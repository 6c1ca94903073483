// CUDA graph conditional WHILE node probe (the device-side round loop of simpleNUTS): does a kernel-set condition drive a
// captured loop on this driver, and may the loop body hold a child graph?   ./cond_graph_probe [child]
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int* n) { *n += 1; }
__global__ void k_cond(cudaGraphConditionalHandle h, int* n, int limit) { cudaGraphSetConditional(h, *n < limit ? 1u : 0u); }
int main(int argc, char** argv) {
  const bool use_child = argc > 1;
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaStream_t s, s2; cudaStreamCreate(&s); cudaStreamCreate(&s2);
  // a child graph (the way an evaluation graph enters a captured sampler iteration)
  cudaGraph_t cg; cudaGraphExec_t cx;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
  k_body<<<1, 1, 0, s>>>(d);
  cudaStreamEndCapture(s, &cg);
  cudaGraphInstantiate(&cx, cg, 0);
  cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
  k_body<<<1, 1, 0, s>>>(d);
  cudaStreamCaptureStatus st; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
  cudaStreamGetCaptureInfo(s, &st, nullptr, &g, &deps, &nd);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  cudaError_t e = cudaGraphAddNode(&node, g, deps, nd, &p);
  printf("add node: %s\n", cudaGetErrorString(e));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
  e = cudaStreamBeginCaptureToGraph(s2, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  printf("begin body: %s\n", cudaGetErrorString(e));
  k_body<<<1, 1, 0, s2>>>(d);
  if (use_child) { e = cudaGraphLaunch(cx, s2); printf("child graph launch inside the body capture: %s\n", cudaGetErrorString(e)); }
  k_cond<<<1, 1, 0, s2>>>(h, d, 10);
  e = cudaStreamEndCapture(s2, nullptr);
  printf("end body: %s\n", cudaGetErrorString(e));
  k_body<<<1, 1, 0, s>>>(d);
  e = cudaStreamEndCapture(s, &g);
  printf("end: %s\n", cudaGetErrorString(e));
  cudaGraphExec_t x; e = cudaGraphInstantiate(&x, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
  for (int r = 0; r < 3; r++) { cudaMemset(d, 0, 4); cudaGraphLaunch(x, s); cudaStreamSynchronize(s); int hn; cudaMemcpy(&hn, d, 4, cudaMemcpyDeviceToHost); printf("n = %d (expect 11 without the child graph: head 1, the body runs while n < 10, tail 1)\n", hn); }
  return 0;
}
