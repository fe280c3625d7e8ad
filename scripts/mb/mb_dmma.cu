// FP64 peak of one B200: mma.sync.m8n8k4.f64 (SASS DMMA) and plain DFMA, from
// registers only (no memory traffic), as a function of resident warps per SM.
// Prints TFLOP/s; the best DMMA figure is the `peak` of the flop-bound roofline
// (bench.py --workload c4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_dmma mb_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ACC>
__global__ void k_dmma(int iters, double *out) {
  double c[ACC][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < ACC; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int ACC>
__global__ void k_dfma(int iters, double *out) {
  double c[ACC];
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
#pragma unroll
  for (int i = 0; i < ACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double *out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  double best_mma = 0, best_fma = 0;
  for (int wps : {4, 8, 16, 32, 64}) {
    const int threads = wps >= 32 ? 1024 : wps * 32, ctas = nsm * (wps >= 32 ? wps / 32 : 1);
    float ms;
    k_dmma<8><<<ctas, threads>>>(100, out);
    cudaEventRecord(e0); k_dmma<8><<<ctas, threads>>>(iters, out); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)ctas * threads / 32;
    const double tf = warps * iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;  // m8n8k4: 256 FMA
    k_dfma<8><<<ctas, threads>>>(100, out);
    cudaEventRecord(e0); k_dfma<8><<<ctas, threads>>>(iters, out); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms2; cudaEventElapsedTime(&ms2, e0, e1);
    const double tf2 = warps * 32 * iters * 8.0 * 2.0 / (ms2 * 1e-3) / 1e12;
    printf("warps/SM %2d: DMMA %.2f TFLOP/s (%.3f ms)   DFMA %.2f TFLOP/s (%.3f ms)   %s\n", wps, tf, ms,
           tf2, ms2, cudaGetErrorString(cudaGetLastError()));
    if (tf > best_mma) best_mma = tf;
    if (tf2 > best_fma) best_fma = tf2;
  }
  // one warp alone on an SM: issue interval of dependent / independent DMMAs
  {
    float ms;
    cudaEventRecord(e0); k_dmma<1><<<1, 32>>>(iters, out); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("one warp, 1 dependent chain : %.1f ns per DMMA\n", ms * 1e6 / iters);
    cudaEventRecord(e0); k_dmma<8><<<1, 32>>>(iters, out); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("one warp, 8 independent     : %.1f ns per DMMA\n", ms * 1e6 / iters / 8);
    cudaEventRecord(e0); k_dmma<8><<<1, 128>>>(iters, out); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("four warps (4 sub-partitions), 8 independent each: %.1f ns per DMMA per warp\n", ms * 1e6 / iters / 8);
  }
  printf("{\"fp64_dmma_tflops\": %.2f, \"fp64_dfma_tflops\": %.2f, \"sms\": %d}\n", best_mma, best_fma, nsm);
  return 0;
}
