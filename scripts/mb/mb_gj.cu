// microbenchmark + correctness of warp_gj_inverse vs cta_gauss_jordan (scratch)
#define LQ_GJ_STAMPS
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../hqp_b200/csrc/lq_device.cuh"

template <int N>
__global__ void __launch_bounds__(256) k_gj(const double *A, double *out, long long *cyc) {
  extern __shared__ double dyn[];
  double *M = dyn, *X = M + N * 3 * N, *Mi = X + N * 2 * N, *M0 = Mi + N * N;
  __shared__ int piv_s[64], st_s, rowsel[65];
  __shared__ __align__(16) double rowbuf[2 * (N + 2)];
  __shared__ double inv_s[2];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int r = i / N, c = i % N;
    M0[i] = A[i];
    M[r * 3 * N + c] = A[i];
    M[r * 3 * N + N + c] = r == c ? 1.0 : 0.0;
    M[r * 3 * N + 2 * N + c] = 0.0;
  }
  if (threadIdx.x == 0) st_s = 0;
  __syncthreads();
  long long t0 = clock64();
  cta_gauss_jordan<N>(M, 3 * N, N, 3 * N, X, piv_s, inv_s, &st_s);
  long long t1 = clock64();
  int fl = 0;
  if (warp_id_uniform() == 0) fl = warp_gj_inverse<N>(M0, N, Mi, N, rowbuf, rowsel);
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = fl | (st_s << 8); }
  // out: [inverse from CTA GJ (X[:, :N], ld 2N) | inverse from warp GJ]
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int r = i / N, c = i % N;
    out[i] = X[r * 2 * N + c];
    out[N * N + i] = Mi[i];
  }
}

template <int N>
void run(unsigned seed) {
  double *A, *out; long long *cyc;
  cudaMalloc(&A, N * N * 8); cudaMalloc(&out, 2 * N * N * 8); cudaMalloc(&cyc, 64);
  double h[N * N], o[2 * N * N];
  srand(seed);
  for (int i = 0; i < N * N; i++) h[i] = (rand() / (double)RAND_MAX - 0.5);
  for (int i = 0; i < N; i++) h[i * N + (i * 7 + 3) % N] += 3.0;  // needs pivoting
  cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice);
  long long c[3];
  cudaFuncSetAttribute(k_gj<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * N * N * 8);
  for (int rep = 0; rep < 3; rep++) { k_gj<N><<<1, 256, 7 * N * N * 8>>>(A, out, cyc); cudaDeviceSynchronize(); }
  cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
  cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  double e[2] = {0, 0};
  for (int v = 0; v < 2; v++)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        double s = 0;
        for (int k = 0; k < N; k++) s += h[i * N + k] * o[v * N * N + k * N + j];
        e[v] = fmax(e[v], fabs(s - (i == j)));
      }
  { long long st[80]; cudaMemcpyFromSymbol(st, g_gj_stamps, sizeof st); printf("  per-pivot:"); for (int p = 0; p < N; p++) printf(" %lld", st[p + 1] - st[p]); printf("\n"); }
  printf("N=%d cta_gj %lld cyc (|MX-I| %.2e)  warp_gj %lld cyc (|MX-I| %.2e) flags %llx  %s\n", N, c[0], e[0],
         c[1], e[1], c[2], cudaGetErrorString(cudaGetLastError()));
}
int main() { run<12>(1); run<20>(2); run<40>(3); return 0; }
