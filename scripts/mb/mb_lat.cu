// latency probes for the warp-level Gauss-Jordan (scratch)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../hqp_b200/csrc/lq_device.cuh"

__global__ void k(double *A, long long *out) {
  __shared__ __align__(16) double sb[64];
  const int lane = threadIdx.x & 31;
  if (warp_id_uniform() != 0) return;
  // T1: redux + ballot + ffs chain
  unsigned key = (unsigned)(A[lane] * 1e6) + lane;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const unsigned best = __reduce_max_sync(0xffffffffu, key);
    const unsigned ball = __ballot_sync(0xffffffffu, key == best);
    const int rl = __ffs(ball) - 1;
    key = key ^ (unsigned)(rl + i + lane);
  }
  long long t1 = clock64();
  if (lane == 0) out[0] = t1 - t0;
  A[64 + lane] = key;
  // T2: 20 independent DFMAs per round, rounds dependent through m
  double a[20], m = A[lane + 1];
#pragma unroll
  for (int j = 0; j < 20; j++) a[j] = A[lane + j];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) {
#pragma unroll
    for (int j = 0; j < 20; j++) a[j] = fma(-m, 1.0001 + j, a[j]);
    m = a[i];
  }
  t1 = clock64();
  if (lane == 0) out[1] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int j = 0; j < 20; j++) s += a[j];
  A[128 + lane] = s;
  // T3: STS -> syncwarp -> LDS chain
  double v = A[lane];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) {
    if (lane == (i & 31)) sb[i & 1] = v;
    __syncwarp();
    v += sb[i & 1];
  }
  t1 = clock64();
  if (lane == 0) out[2] = t1 - t0;
  A[160 + lane] = v;
  // T4: fabs/hi -> redux -> ballot -> select me -> STS -> LDS -> DMUL (pivot header)
  double x = A[lane] + 1.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const unsigned k2 = (unsigned)__double2hiint(fabs(x));
    const double myinv = fast_rcp(x);
    const unsigned best = __reduce_max_sync(0xffffffffu, k2);
    const unsigned ball = __ballot_sync(0xffffffffu, k2 == best);
    const int rl = __ffs(ball) - 1;
    if (lane == rl) sb[2 + (i & 1)] = myinv;
    __syncwarp();
    x = x * sb[2 + (i & 1)] + 1.0 + lane;
  }
  t1 = clock64();
  if (lane == 0) out[3] = t1 - t0;
  A[192 + lane] = x;
  // T5: 20 independent DFMA, no dependence between rounds except through a[] themselves
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) {
#pragma unroll
    for (int j = 0; j < 20; j++) a[j] = fma(-m, 1.0001 + j, a[j]);
  }
  t1 = clock64();
  if (lane == 0) out[4] = t1 - t0;
  s = 0;
#pragma unroll
  for (int j = 0; j < 20; j++) s += a[j];
  A[224 + lane] = s;
}
int main() {
  double *A; long long *out;
  cudaMalloc(&A, 4096 * 8); cudaMalloc(&out, 16 * 8);
  double h[4096];
  for (int i = 0; i < 4096; i++) h[i] = 0.001 * ((i * 37) % 101) + 0.5;
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) { k<<<1, 128>>>(A, out); cudaDeviceSynchronize(); }
  long long o[16];
  cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
  printf("redux+ballot+ffs chain: %.1f cyc each\n20 indep DFMA + dep round: %.1f cyc per round (%.1f per DFMA)\nSTS->syncwarp->LDS: %.1f each\npivot header: %.1f each\n20 DFMA rounds (8-deep chains): %.1f per round\n",
         o[0] / 32.0, o[1] / 16.0, o[1] / 320.0, o[2] / 32.0, o[3] / 32.0, o[4] / 16.0);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
