// microbenchmark of chain_run_warp (scratch): cycles per step, 1 CTA and full grid
#include <cstdio>
#include <cuda_runtime.h>
#include "../../hqp_b200/csrc/lq_solve.cuh"

template <bool TRANS, int N>
__global__ void k_chain(const double *M, const double *a, const double *b, double *out, int cnt,
                        int CH, long long *cyc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem cs = chain_smem_init(N, smem_raw, CH);
  constexpr int NR = (N + 31) / 32;
  double t[NR];
  for (int r = 0; r < NR; r++) t[r] = 0.001 * threadIdx.x;
  const size_t off = (size_t)blockIdx.x * cnt;
  long long t0 = clock64();
  chain_run_warp<TRANS, N>(M + off * N * N, a + off * N, TRANS ? b + off * N : nullptr, nullptr,
                           out + off * N, 0, TRANS ? cnt - 1 : 0, TRANS ? -1 : 1, cnt, t, cs.ring,
                           cs.bars, cs.tvec, CH);
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <bool TRANS, int N>
void run(int grid, int cnt, int CH) {
  size_t nst = (size_t)grid * cnt;
  double *M, *a, *b, *out; long long *cyc;
  cudaMalloc(&M, nst * N * N * 8); cudaMalloc(&a, nst * N * 8); cudaMalloc(&b, nst * N * 8);
  cudaMalloc(&out, nst * N * 8); cudaMalloc(&cyc, 64);
  cudaMemset(M, 0, nst * N * N * 8); cudaMemset(a, 0, nst * N * 8); cudaMemset(b, 0, nst * N * 8);
  size_t smem = ((size_t)2 * CH * (N * N + 2 * N) + 3 * N + 8) * 8 + 64;
  cudaFuncSetAttribute(k_chain<TRANS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_chain<TRANS, N><<<grid, 32, smem>>>(M, a, b, out, cnt, CH, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("N=%d %s grid=%d cnt=%d CH=%d: %.1f cyc/step (CTA 0), kernel %.1f us  %s\n", N, TRANS ? "back" : "fwd ",
         grid, cnt, CH, (double)c / cnt, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(M); cudaFree(a); cudaFree(b); cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int CH : {4, 8}) {
    run<true, 20>(1, 23, CH); run<false, 20>(1, 23, CH);
    run<true, 20>(435, 23, CH); run<false, 20>(435, 23, CH);
  }
  run<true, 20>(1, 230, 23); run<false, 20>(1, 230, 23);
  run<true, 12>(4096, 50, 4); run<false, 12>(4096, 50, 4);
  run<true, 40>(148, 676, 4); run<false, 40>(148, 676, 4);
  return 0;
}
