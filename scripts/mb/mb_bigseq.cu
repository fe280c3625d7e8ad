// Sequence test for cta_mm_big: different products back to back inside one kernel, operands
// written in-kernel, as the large-block stage kernels do.  argv[1] = bitmask of steps to run.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lq_device.cuh"

__global__ void __launch_bounds__(LQ_BIG_NT) k_seq(double *ws, size_t stride, int nx, int nu, int steps, int reps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *stg = reinterpret_cast<double *>(smem_raw);
  const int nm = nx + nu, n2 = nx * nx;
  double *V = ws + blockIdx.x * stride, *fx = V + n2, *fu = fx + n2, *T = fu + nx * nu, *G = T + nx * nm,
         *Rux = G + nm * nm, *Phi = Rux + nu * nx, *P0 = Phi + n2, *P1 = P0 + n2;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = 0; r < reps; r++) {
    if (steps & 1) {  // T = V [fx fu]
      cta_mm_big(stg, T, nm, nullptr, 0, 0.0, 1.0, V, 1, nx, fx, nx, 1, nx, nx, nx, 0, warp, nw);
      cta_mm_big(stg, T + nx, nm, nullptr, 0, 0.0, 1.0, V, 1, nx, fu, nu, 1, nx, nu, nx, 0, warp, nw);
      __syncthreads();
    }
    if (steps & 2) {  // G += [fx fu]' T
      cta_mm_big(stg, G, nm, G, nm, 1.0, 1e-3, fx, 1, nx, T, nm, 1, nx, nx, nx, 0, warp, nw);
      cta_mm_big(stg, G + nx * nm, nm, G + nx * nm, nm, 1.0, 1e-3, fu, 1, nu, T, nm, 1, nu, nm, nx, 0, warp, nw);
      __syncthreads();
    }
    if (steps & 4) {  // generic writes + symmetrize
      for (int i = threadIdx.x; i < nu * nx; i += blockDim.x) Rux[i] = 1e-3 * G[nx * nm + (i / nx) * nm + i % nx];
      cta_symmetrize_big(stg, V, nx, nx);
      __syncthreads();
    }
    if (steps & 8) {  // V = G - Gux' Rux ; Phi = fx - fu Rux
      cta_mm_big(stg, V, nx, G, nm, 1.0, -1e-3, G + nx * nm, 1, nm, Rux, nx, 1, nx, nx, nu, 0, warp, nw);
      cta_mm_big(stg, Phi, nx, fx, nx, 1.0, -1.0, fu, nu, 1, Rux, nx, 1, nx, nx, nu, 0, warp, nw);
      __syncthreads();
    }
    if (steps & 16) {  // Psi
      cta_mm_big(stg, P1, nx, nullptr, 0, 0.0, 1e-2, Phi, 1, nx, P0, nx, 1, nx, nx, nx, 0, warp, nw);
      __syncthreads();
    }
    if (steps & 32) {  // row-major product (tree)
      cta_mm_big(stg, P0, nx, nullptr, 0, 0.0, 1e-2, P1, nx, 1, Phi, nx, 1, nx, nx, nx, 0, warp, nw);
      __syncthreads();
    }
  }
}

int main(int argc, char **argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 63, nx = argc > 2 ? atoi(argv[2]) : 200, nu = argc > 3 ? atoi(argv[3]) : 50;
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  setvbuf(stdout, nullptr, _IONBF, 0);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int nm = nx + nu;
  const size_t stride = ((size_t)6 * nx * nx + 3 * nx * nu + nx * nm + nm * nm + 16) & ~(size_t)1;
  double *ws;
  cudaMalloc(&ws, nsm * stride * 8);
  std::vector<double> h(nsm * stride);
  for (size_t i = 0; i < h.size(); i++) h[i] = (((i * 2654435761u) % 2001) / 1000.0 - 1.0) * 0.05;
  cudaMemcpy(ws, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)LQ_BIG_STG * 8;
  cudaFuncSetAttribute(k_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_seq<<<nsm, LQ_BIG_NT, smem>>>(ws, stride, nx, nu, steps, reps);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("steps %d nx %d nu %d reps %d: %.3f ms  %s\n", steps, nx, nu, reps, ms, cudaGetErrorString(err));
  return 0;
}
