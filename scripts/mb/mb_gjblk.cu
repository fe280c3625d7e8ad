// Blocked Gauss-Jordan solve of the large-block combine (cta_gj_solve_blocked) against the
// unblocked in-place inverse + GEMM, per CTA on its own L2-resident [M0 | R].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I hqp_b200/csrc \
//        -o scratch_bin/mb_gjblk scripts/mb/mb_gjblk.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lq_device.cuh"

__global__ void __launch_bounds__(LQ_BIG_NT) k_solve(double *ws, size_t stride, int n, int nc, int blocked) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *stg = reinterpret_cast<double *>(smem_raw);
  double *W = ws + blockIdx.x * stride, *X = W + (size_t)n * nc, *scr = X + (size_t)n * (nc - n);
  __shared__ int piv[LQ_GJ_MAXN];
  __shared__ int st;
  if (threadIdx.x == 0) st = 0;
  __syncthreads();
  cta_inverse_apply_big(stg, blocked ? scr : nullptr, W, nc, n, nc, X, nc - n, piv, &st);
  if (threadIdx.x == 0 && st) printf("block %d status %d\n", blockIdx.x, st);
}

__global__ void k_resid(const double *orig, const double *ws, size_t stride, int n, int nc, double *err) {
  const double *W0 = orig + blockIdx.x * stride, *X = ws + blockIdx.x * stride + (size_t)n * nc;
  const int nr = nc - n;
  double e = 0;
  for (int idx = threadIdx.x; idx < n * nr; idx += blockDim.x) {
    const int i = idx / nr, j = idx % nr;
    double s = 0;
    for (int l = 0; l < n; l++) s += W0[(size_t)i * nc + l] * X[(size_t)l * nr + j];
    e = fmax(e, fabs(s - W0[(size_t)i * nc + n + j]));
  }
  atomicMax(reinterpret_cast<unsigned long long *>(err), (unsigned long long)__double_as_longlong(e));
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 200, mult = argc > 2 ? atoi(argv[2]) : 3;
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int nc = mult * n;
  const size_t stride = ((size_t)n * nc + (size_t)n * (nc - n) + big_gj_scratch_doubles(n) + 16) & ~(size_t)1;
  double *ws, *orig, *err;
  cudaMalloc(&ws, nsm * stride * 8);
  cudaMalloc(&orig, nsm * stride * 8);
  cudaMalloc(&err, 8);
  std::vector<double> h(nsm * stride, 0.0);
  unsigned long long sd = 12345;
  for (int b = 0; b < nsm; b++)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < nc; j++) {
        sd = sd * 6364136223846793005ULL + 1442695040888963407ULL;
        const double u = ((sd >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
        h[b * stride + (size_t)i * nc + j] = (j < n ? 0.3 * u / 4 + (i == (j * 7 + 3) % n ? 1.0 : 0.0) : u);
      }
  cudaMemcpy(orig, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)LQ_BIG_STG * 8;
  cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int blocked = 1; blocked >= 0; blocked--) {
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
      cudaMemcpy(ws, orig, nsm * stride * 8, cudaMemcpyDeviceToDevice);
      cudaEventRecord(e0);
      k_solve<<<nsm, LQ_BIG_NT, smem>>>(ws, stride, n, nc, blocked);
      cudaEventRecord(e1);
      cudaError_t er = cudaEventSynchronize(e1);
      if (er != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(er)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    cudaMemset(err, 0, 8);
    k_resid<<<nsm, 256>>>(orig, ws, stride, n, nc, err);
    double he;
    cudaMemcpy(&he, err, 8, cudaMemcpyDeviceToHost);
#ifdef LQ_GJ_STAMPS
    if (blocked) {
      long long cyc[8];
      cudaMemcpyFromSymbol(cyc, g_gj_cyc, sizeof cyc);
      printf("   cycles over 3 runs: panel load %lld, panel elimination %lld, Y %lld, U %lld, update GEMM %lld\n", cyc[0], cyc[1],
             cyc[2], cyc[3], cyc[4]);
    }
#endif
    printf("n %d nc %d %s: %.3f ms per solve (all SMs busy), max |M0 X - R| = %.2e  %s\n", n, nc,
           blocked ? "blocked  " : "unblocked", best, he, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
