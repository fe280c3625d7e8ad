// cost of a dependent kernel boundary: plain launches vs CUDA graph (scratch)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tiny(double *p, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0;
}
__global__ void wide(double *p, int n) {   // 435 CTAs x 128 threads touching memory
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = p[i] * 1.0000001 + 1.0;
}
int main() {
  double *p; cudaMalloc(&p, 1 << 20); cudaMemset(p, 0, 1 << 20);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int NK = 40;
  for (int mode = 0; mode < 2; mode++) {
    for (int use_graph = 0; use_graph < 2; use_graph++) {
      cudaGraphExec_t exec = nullptr;
      if (use_graph) {
        cudaGraph_t g;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < NK; i++) { if (mode) wide<<<435, 128, 0, s>>>(p, 435 * 128); else tiny<<<1, 32, 0, s>>>(p, 1); }
        cudaStreamEndCapture(s, &g);
        cudaGraphInstantiate(&exec, g, 0);
      }
      float best = 1e9;
      for (int rep = 0; rep < 10; rep++) {
        cudaEventRecord(e0, s);
        if (use_graph) cudaGraphLaunch(exec, s);
        else for (int i = 0; i < NK; i++) { if (mode) wide<<<435, 128, 0, s>>>(p, 435 * 128); else tiny<<<1, 32, 0, s>>>(p, 1); }
        cudaEventRecord(e1, s); cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("%s kernels, %s: %.2f us per kernel\n", mode ? "wide (435x128)" : "tiny (1x32)", use_graph ? "graph" : "stream", best * 1e3 / NK);
    }
  }
  return 0;
}
