// Efficiency of the large-block CTA GEMM (cta_mm_big, lq_device.cuh) in isolation:
// every CTA multiplies its own L2-resident operands, as the segment kernels do at
// nx = 200.  Prints TFLOP/s over all SMs per operand orientation and shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -I hqp_b200/csrc -o scratch_bin/mb_biggemm scripts/mb/mb_biggemm.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lq_device.cuh"

// mode bit 0: A transposed access (ar = 1), bit 1: B "rows contiguous" (bc = 1)
__global__ void __launch_bounds__(LQ_BIG_NT) k_gemm(double *ws, size_t stride, int M, int N, int Kd,
                                                    int ld, int mode, int reps, int sub) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *stg = reinterpret_cast<double *>(smem_raw);
  double *A = ws + blockIdx.x * stride, *B = A + (size_t)ld * ld, *C = B + (size_t)ld * ld;
  const int ar = (mode & 1) ? 1 : ld, ac = (mode & 1) ? ld : 1;
  const int br = (mode & 2) ? ld : 1, bc = (mode & 2) ? 1 : ld;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = 0; r < reps; r++) {
    if (sub) {  // all warps but warp 0 (the K3 situation while warp 0 factors Guu)
      if (warp > 0) cta_mm_big(stg, C, ld, nullptr, 0, 0.0, 1.0, A, ar, ac, B, br, bc, M, N, Kd, 0, warp - 1, nw - 1);
    } else {
      if (mode & 4) cta_mm_big(stg, C, ld, r == 0 ? nullptr : C, ld, 1.0, 0.5, A, ar, ac, B, br, bc, M, N, Kd, 0, warp, nw);
      else cta_mm_big(stg, C, ld, nullptr, 0, 0.0, 1.0, A, ar, ac, B, br, bc, M, N, Kd, 0, warp, nw);
    }
    __syncthreads();
  }
}

__global__ void k_check(const double *ws, size_t stride, int M, int N, int Kd, int ld, int mode, double *err, double scale) {
  const double *A = ws + blockIdx.x * stride, *B = A + (size_t)ld * ld, *C = B + (size_t)ld * ld;
  const int ar = (mode & 1) ? 1 : ld, ac = (mode & 1) ? ld : 1;
  const int br = (mode & 2) ? ld : 1, bc = (mode & 2) ? 1 : ld;
  double e = 0;
  for (int idx = threadIdx.x; idx < M * N; idx += blockDim.x) {
    const int i = idx / N, j = idx % N;
    double s = 0;
    for (int l = 0; l < Kd; l++) s += A[i * ar + l * ac] * B[l * br + j * bc];
    e = fmax(e, fabs(scale * s - C[(size_t)i * ld + j]));
  }
  atomicMax(reinterpret_cast<unsigned long long *>(err), (unsigned long long)__double_as_longlong(e));
}

int main(int argc, char **argv) {
  const int only_shape = argc > 1 ? atoi(argv[1]) : -1, only_mode = argc > 2 ? atoi(argv[2]) : -1;
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int ld = argc > 3 ? atoi(argv[3]) : 250;
  const size_t stride = (size_t)3 * ld * ld + 8;
  double *ws, *err;
  cudaMalloc(&ws, nsm * stride * 8);
  cudaMalloc(&err, 8);
  std::vector<double> h(nsm * stride);
  for (size_t i = 0; i < h.size(); i++) h[i] = ((i * 2654435761u) % 2001) / 1000.0 - 1.0;
  cudaMemcpy(ws, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)LQ_BIG_STG * 8;
  cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Shape { int M, N, K; } shapes[] = {{200, 200, 200}, {200, 250, 200}, {200, 200, 50}, {50, 200, 200},
                                           {50, 50, 200}, {200, 50, 200}, {250, 250, 200}, {130, 130, 130}, {50, 250, 200}, {70, 90, 70}, {20, 70, 70}};
  int si = -1;
  for (auto sh : shapes) {
    si++;
    if (only_shape >= 0 && si != only_shape) continue;
    for (int mode = 0; mode < 8; mode++) {
      if (only_mode >= 0 && mode != only_mode) continue;
      for (int sub = 0; sub < 2; sub++) {
        if (sub) continue;
        const int reps = 20;
        k_gemm<<<nsm, LQ_BIG_NT, smem>>>(ws, stride, sh.M, sh.N, sh.K, ld, mode, 2, sub);
        cudaEventRecord(e0);
        k_gemm<<<nsm, LQ_BIG_NT, smem>>>(ws, stride, sh.M, sh.N, sh.K, ld, mode, reps, sub);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemset(err, 0, 8);
        k_check<<<nsm, 256>>>(ws, stride, sh.M, sh.N, sh.K, ld, mode, err, (mode & 4) ? 0.5 * reps : 1.0);
        double he;
        cudaMemcpy(&he, err, 8, cudaMemcpyDeviceToHost);
        const double fl = 2.0 * sh.M * sh.N * sh.K * reps * nsm;
        printf("M %3d N %3d K %3d mode %d%s: %7.3f ms  %6.2f TFLOP/s  %8.0f cycles/product  maxerr %.2e  %s\n", sh.M,
               sh.N, sh.K, mode, sub ? " (15 warps)" : "", ms, fl / (ms * 1e-3) / 1e12,
               ms * 1e-3 * 1.965e9 / reps, he, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
