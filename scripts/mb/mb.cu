#include <cstdio>
#include <cuda_runtime.h>
#include "../../hqp_b200/csrc/lq_device.cuh"

__global__ void k_ldl(double *A, long long *out) {
  __shared__ double S[30 * 30];
  for (int i = threadIdx.x; i < 900; i += blockDim.x) S[i] = A[i];
  __syncthreads();
  long long t0 = clock64();
  int st = 0;
  if (threadIdx.x < 32) st = warp_ldlt_reg<10>(S + 20 * 30 + 20, 30);
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = st; }
  // dependent DFMA chain
  double x = A[threadIdx.x], y = A[threadIdx.x + 1];
  __syncthreads();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) x = fma(x, y, 1.0);
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  A[threadIdx.x] = x;
  // dependent shuffles of a double
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) x = __shfl_sync(0xffffffffu, x, (i * 7) & 31) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  A[threadIdx.x + 32] = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) x = fast_rcp(x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  A[threadIdx.x + 64] = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) x = 1.0 / (x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
  A[threadIdx.x + 96] = x;
  // dependent DMMA chain
  double c0 = 0, c1 = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) dmma_m8n8k4(c0, c1, x, y);
  t1 = clock64();
  if (threadIdx.x == 0) out[6] = t1 - t0;
  A[threadIdx.x + 128] = c0 + c1;
  // LDS dependent chain
  int idx = threadIdx.x;
  __shared__ int chain[128];
  chain[threadIdx.x] = (threadIdx.x * 5 + 3) & 127;
  __syncthreads();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) idx = chain[idx];
  t1 = clock64();
  if (threadIdx.x == 0) out[7] = t1 - t0;
  A[threadIdx.x + 160] = idx;
  // syncthreads cost
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) out[8] = t1 - t0;
}

int main() {
  double *A; long long *out;
  cudaMalloc(&A, 4096 * 8); cudaMalloc(&out, 16 * 8);
  double h[4096];
  for (int i = 0; i < 4096; i++) h[i] = 0.001 * ((i * 37) % 101);
  for (int i = 0; i < 30; i++) h[i * 30 + i] += 10.0;
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) {
    cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
    k_ldl<<<1, 128>>>(A, out);
    cudaDeviceSynchronize();
  }
  long long o[16];
  cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
  printf("ldl10 %lld cycles (st %lld)\n64 dep DFMA %lld (%.1f each)\n64 dep shfl+dadd %lld (%.1f each)\n16 fast_rcp %lld (%.1f each)\n16 div %lld (%.1f each)\n32 dep DMMA %lld (%.1f each)\n32 dep LDS %lld (%.1f each)\n16 syncthreads %lld (%.1f each)\n",
         o[0], o[1], o[2], o[2] / 64.0, o[3], o[3] / 64.0, o[4], o[4] / 16.0, o[5], o[5] / 16.0, o[6], o[6] / 32.0, o[7], o[7] / 32.0, o[8], o[8] / 16.0);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
