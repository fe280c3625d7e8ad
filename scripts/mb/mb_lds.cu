// issue rate of independent shared-memory loads from ONE warp (scratch)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, long long *cyc, int iters) {
  __shared__ __align__(16) double S[32 * 48];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 48; i += blockDim.x) S[i] = i * 1e-3;
  __syncthreads();
  double acc[20];
  for (int j = 0; j < 20; j++) acc[j] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double *p = S + (it & 7) * 32;
#pragma unroll
    for (int j = 0; j < 20; j++) {
      if (MODE == 0) acc[j] += p[j * 32 + lane];                 // LDS.64, lanes consecutive
      if (MODE == 1) acc[j] += p[j * 2];                          // LDS.64 broadcast
      if (MODE == 2) { double2 v = *reinterpret_cast<const double2 *>(p + j * 32 + 2 * (lane & 15)); acc[j] += v.x + v.y; }  // LDS.128
      if (MODE == 3) acc[j] = fma(acc[j], 1.0000001, 0.5);       // no loads: 20 independent DFMA
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < 20; j++) s += acc[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
  const int iters = 200;
  for (int rep = 0; rep < 2; rep++) {
    k<0><<<1, 32>>>(out, cyc, iters); k<1><<<1, 32>>>(out, cyc, iters);
    k<2><<<1, 32>>>(out, cyc, iters); k<3><<<1, 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
  }
  long long c[8]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  printf("1 warp : LDS.64 %.1f | LDS.64 bcast %.1f | LDS.128 %.1f | DFMA only %.1f  cycles per (load + DADD)\n",
         c[0] / (20.0 * iters), c[1] / (20.0 * iters), c[2] / (20.0 * iters), c[3] / (20.0 * iters));
  for (int rep = 0; rep < 2; rep++) {
    k<0><<<1, 128>>>(out, cyc, iters); k<1><<<1, 128>>>(out, cyc, iters);
    k<2><<<1, 128>>>(out, cyc, iters); k<3><<<1, 128>>>(out, cyc, iters);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  printf("4 warps: LDS.64 %.1f | LDS.64 bcast %.1f | LDS.128 %.1f | DFMA only %.1f\n",
         c[0] / (20.0 * iters), c[1] / (20.0 * iters), c[2] / (20.0 * iters), c[3] / (20.0 * iters));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
