// what does one dependent chain step cost? (scratch) -- single warp, data in smem
#include <cstdio>
#include <cuda_runtime.h>
#define N 20
template <int VAR>
__global__ void k(const double *Mg, double *out, long long *cyc, int steps) {
  __shared__ __align__(16) double Ms[N * N], as[N], tv[2 * N];
  const int lane = threadIdx.x;
  for (int i = lane; i < N * N; i += 32) Ms[i] = Mg[i];
  if (lane < N) as[lane] = Mg[lane];
  __syncwarp();
  double t = 0.001 * lane;
  int par = 0;
  const bool act = lane < N;
  const int rs = act ? lane : 0;
  long long t0 = clock64();
  for (int s = 0; s < steps; s++, par ^= 1) {
    double *tvp = tv + par * N;
    if (VAR == 3) {            // all lanes store (no divergence), clamp index
      tvp[rs] = t;
    } else if (act) tvp[lane] = t;
    double m[N];
#pragma unroll
    for (int l = 0; l < N; l++) m[l] = Ms[l * N + rs];
    __syncwarp();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int l = 0; l < N; l += 4) {
      const double2 ta = *reinterpret_cast<const double2 *>(tvp + l);
      const double2 tb = *reinterpret_cast<const double2 *>(tvp + l + 2);
      a0 = fma(m[l], ta.x, a0); a1 = fma(m[l + 1], ta.y, a1);
      a2 = fma(m[l + 2], tb.x, a2); a3 = fma(m[l + 3], tb.y, a3);
    }
    if (VAR == 3) {
      t = as[rs] + ((a0 + a1) + (a2 + a3));
      if (act) out[(size_t)s * N + lane] = t;
    } else if (act) {
      t = as[lane] + ((a0 + a1) + (a2 + a3));
      if (VAR != 1) out[(size_t)s * N + lane] = t;   // VAR 1: no global store
    }
    if (VAR == 2) __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) cyc[VAR] = t1 - t0;
  if (act) out[lane] = t;
}
int main() {
  double *M, *out; long long *cyc;
  cudaMalloc(&M, N * N * 8); cudaMalloc(&out, 1000 * N * 8); cudaMalloc(&cyc, 64);
  cudaMemset(M, 0, N * N * 8);
  const int steps = 500;
  k<0><<<1, 32>>>(M, out, cyc, steps); k<1><<<1, 32>>>(M, out, cyc, steps);
  k<2><<<1, 32>>>(M, out, cyc, steps); k<3><<<1, 32>>>(M, out, cyc, steps);
  cudaDeviceSynchronize();
  k<0><<<1, 32>>>(M, out, cyc, steps); k<1><<<1, 32>>>(M, out, cyc, steps);
  k<2><<<1, 32>>>(M, out, cyc, steps); k<3><<<1, 32>>>(M, out, cyc, steps);
  cudaDeviceSynchronize();
  long long c[8]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  printf("base %.1f | no STG %.1f | extra syncwarp %.1f | no divergence %.1f cycles/step  %s\n", c[0] / (double)steps,
         c[1] / (double)steps, c[2] / (double)steps, c[3] / (double)steps, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
