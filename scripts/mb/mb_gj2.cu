// microbenchmark + correctness: warp_gj_inverse (round 1) vs the look-ahead
// variants warp_gj_inverse_la<N,false> (shared-row broadcast) and <N,true>
// (shuffle broadcast).  One warp, clock64 around 8 back-to-back inversions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_gj2 mb_gj2.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../hqp_b200/csrc/lq_device.cuh"

template <int N, int V>
__global__ void __launch_bounds__(256) k_gj(const double *A, double *out, long long *cyc) {
  __shared__ __align__(16) double M0[N * (N + 1)], Mi[N * (N + 1)];
  __shared__ int rowsel[65];
  __shared__ __align__(16) double rowbuf[2 * (N + 2)];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) M0[(i / N) * (N + 1) + i % N] = A[i];
  __syncthreads();
  int fl = 0;
  long long t0 = 0, t1 = 0;
  if (warp_id_uniform() == 0) {
    t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 8; rep++) {
      if (V == 0) fl |= warp_gj_inverse<N>(M0, N + 1, Mi, N + 1, rowbuf, rowsel);
      if (V == 1) fl |= warp_gj_inverse_la<N, false>(M0, N + 1, Mi, N + 1, rowbuf, rowsel);
      if (V == 2) fl |= warp_gj_inverse_la<N, true>(M0, N + 1, Mi, N + 1, rowbuf, rowsel);
      __syncwarp();
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) { cyc[0] = (t1 - t0) / 8; cyc[1] = fl; }
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) out[i] = Mi[(i / N) * (N + 1) + i % N];
}

template <int N, int V>
void run(unsigned seed, const char *name) {
  double *A, *out; long long *cyc;
  cudaMalloc(&A, N * N * 8); cudaMalloc(&out, N * N * 8); cudaMalloc(&cyc, 64);
  double h[N * N], o[N * N];
  srand(seed);
  for (int i = 0; i < N * N; i++) h[i] = (rand() / (double)RAND_MAX - 0.5);
  for (int i = 0; i < N; i++) h[i * N + (i * 7 + 3) % N] += 3.0;  // needs pivoting
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  long long c[2];
  for (int rep = 0; rep < 3; rep++) { k_gj<N, V><<<1, 256>>>(A, out, cyc); cudaDeviceSynchronize(); }
  cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
  cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  double e = 0;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += h[i * N + k] * o[k * N + j];
      e = fmax(e, fabs(s - (i == j)));
    }
  printf("N=%d %-22s %6lld cycles (%.0f per pivot)  |MX-I| %.2e flags %lld  %s\n", N, name, c[0],
         (double)c[0] / N, e, c[1], cudaGetErrorString(cudaGetLastError()));
  cudaFree(A); cudaFree(out); cudaFree(cyc);
}
int main() {
  run<12, 0>(1, "round-1"); run<12, 1>(1, "look-ahead smem"); run<12, 2>(1, "look-ahead shfl");
  run<20, 0>(2, "round-1"); run<20, 1>(2, "look-ahead smem"); run<20, 2>(2, "look-ahead shfl");
  run<32, 0>(3, "round-1"); run<32, 1>(3, "look-ahead smem"); run<32, 2>(3, "look-ahead shfl");
  return 0;
}
