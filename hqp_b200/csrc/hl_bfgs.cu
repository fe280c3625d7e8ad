// Block-diagonal BFGS update of the Lagrangian Hessian on the GPU.
//
// What the reference does once per SQP iteration, one diagonal block of Q after the
// other on one CPU core (Hqp_HL_BFGS::update, hqp/Hqp_HL_BFGS.C:216-243, block update
// update_b_Q :149-213):
//     sv = s'u,  Qs = Q s,  sQ = s'Q,  sQs = sQ s
//     Powell's damping (:176-184): if sv < gamma sQs: v = theta u + (1 - theta) Qs
//     Q <- Q - Qs sQ' / sQs + v v' / sv                       (:194-202)
//     eigenvalue control (:204-212): lambda_min of the new block (symmeig, a Householder
//     + implicit QL solve of the whole (nx+nu)^2 block), shifted up to theta = eps^2
// The blocks are independent: here ONE WARP owns a block; the block and a working copy
// live in shared memory, the smallest eigenvalue comes from a cyclic Jacobi iteration in
// round-robin ordering (n/2 disjoint rotations at a time, rows then columns), which
// converges quadratically to eigenvalues accurate to a few ulp of ||Q|| -- the accuracy
// of the reference's QL.  Arithmetic order of the rank-2 update follows the reference
// ((a*b)/c, subtraction before addition), no FMA contraction in that part.
//
// C ABI: include/hqp_hlcuda.h.  No CPU fallback: without a device every entry point
// returns HQPHL_E_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "hqp_hlcuda.h"

namespace {

thread_local std::string g_err;

struct BfgsArgs {
  int nblocks;
  const int *bsize;          // [nblocks]
  const long long *qoff;     // [nblocks] offset of the block in Q (doubles)
  const int *voff;           // [nblocks] offset of the block in s, u
  double *Q;
  const double *s, *u;
  double alpha, gamma, eps;
  int eigen_control;
  int max_n;                 // largest block
  int *info;                 // [0]: blocks whose diagonal was shifted, [1]: blocks skipped (sv or sQs zero),
                             // [2]: blocks whose Jacobi iteration hit the sweep limit
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// smallest eigenvalue of the symmetric n x n matrix W (ld = n, destroyed) by one warp.
// rot: 3 * (n/2 + 1) doubles of scratch (c, s and the index pair per rotation).  Returns the sweep count in *sweeps.
__device__ double warp_jacobi_min_eig(double *W, int n, double *rot, int *sweeps) {
  const int lane = threadIdx.x & 31;
  const int ne = (n + 1) & ~1;  // players of the tournament (one dummy for odd n)
  const int np = ne >> 1;
  double *cs = rot, *sn = rot + np;
  int *pq = reinterpret_cast<int *>(rot + 2 * np);  // [np][2]
  int sw = 0;
  for (; sw < 40; sw++) {
    // off-diagonal and total Frobenius norms
    double off = 0.0, tot = 0.0;
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e - i * n;
      const double a = W[e];
      tot = fma(a, a, tot);
      if (i != j) off = fma(a, a, off);
    }
    off = warp_sum(off);
    tot = warp_sum(tot);
    if (!(off > 1e-30 * tot) || !(tot > 0.0)) break;  // (relative 1e-15 on the norms; also NaN)
    for (int r = 0; r < ne - 1; r++) {
      // pairs of round r: player ne-1 stays, the others rotate; (p, q, c, s) per pair in
      // shared memory, so that the two passes below carry no index arithmetic
      for (int k = lane; k < np; k += 32) {
        int p = k == 0 ? ne - 1 : (r + k) % (ne - 1);
        int q = (r + ne - 1 - k) % (ne - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = W[p * n + q];
          if (apq != 0.0) {
            const double tau = (W[q * n + q] - W[p * n + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = p;  // the dummy player of an odd n: identity on row p
        }
        cs[k] = c;
        sn[k] = s;
        pq[2 * k] = p;
        pq[2 * k + 1] = q;
      }
      __syncwarp();
      // rows: (row p, row q) <- (c row p - s row q, s row p + c row q); lanes along the row
      for (int k = 0; k < np; k++) {
        const int p = pq[2 * k], q = pq[2 * k + 1];
        const double c = cs[k], s = sn[k];
        if (s != 0.0) {
          for (int j = lane; j < n; j += 32) {
            const double x = W[p * n + j], y = W[q * n + j];
            W[p * n + j] = c * x - s * y;
            W[q * n + j] = s * x + c * y;
          }
        }
      }
      __syncwarp();
      // columns; lanes along the column
      for (int k = 0; k < np; k++) {
        const int p = pq[2 * k], q = pq[2 * k + 1];
        const double c = cs[k], s = sn[k];
        if (s != 0.0) {
          for (int i = lane; i < n; i += 32) {
            const double x = W[i * n + p], y = W[i * n + q];
            W[i * n + p] = c * x - s * y;
            W[i * n + q] = s * x + c * y;
          }
        }
      }
      __syncwarp();
    }
  }
  double mn = INFINITY;
  for (int i = lane; i < n; i += 32) mn = fmin(mn, W[i * n + i]);
  *sweeps = sw;
  return warp_min(mn);
}

// true iff W - theta I is positive definite: LDL' without interchanges, all pivots > 0
// (W destroyed).  lambda_min(W) > theta then needs no eigenvalue at all -- the usual case
// for a damped BFGS matrix; only blocks that fail pay for the Jacobi iteration.
__device__ bool warp_pd_certificate(double *W, int n, double theta) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < n; i += 32) W[i * n + i] -= theta;
  __syncwarp();
  for (int p = 0; p < n; p++) {
    const double d = W[p * n + p];
    if (!(d > 0.0)) return false;  // (the same value on every lane)
    const double inv = 1.0 / d;
    const int r = n - p - 1;
    for (int e = lane; e < r * r; e += 32) {
      const int ii = e / r, jj = e - ii * r;
      if (jj <= ii) {
        const int i = p + 1 + ii, j = p + 1 + jj;
        W[i * n + j] = fma(-W[i * n + p] * inv, W[j * n + p], W[i * n + j]);
      }
    }
    __syncwarp();
  }
  return true;
}

// one warp per block; dynamic shared memory: per warp 2 max_n^2 + 6 max_n + 8 doubles
__global__ void hl_bfgs_kernel(BfgsArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t per = (size_t)2 * a.max_n * a.max_n + 6 * a.max_n + 8;
  double *Qb = smem + warp * per;
  double *W = Qb + (size_t)a.max_n * a.max_n;
  double *sv_ = W + (size_t)a.max_n * a.max_n;   // s
  double *vv = sv_ + a.max_n;                     // u, then v
  double *Qs = vv + a.max_n, *sQ = Qs + a.max_n, *rot = sQ + a.max_n;
  for (int b = blockIdx.x * nw + warp; b < a.nblocks; b += gridDim.x * nw) {
    const int n = a.bsize[b];
    double *Qg = a.Q + a.qoff[b];
    const double *s = a.s + a.voff[b], *u = a.u + a.voff[b];
    for (int e = lane; e < n * n; e += 32) Qb[e] = Qg[e];
    for (int i = lane; i < n; i += 32) { sv_[i] = s[i]; vv[i] = u[i]; }
    __syncwarp();
    // Qs = Q s, sQ = s'Q (the block may carry different lower and upper parts when
    // eigenvalue control is off: vm_mlt and mv_mlt of the reference, :158-160)
    double part = 0.0;
    for (int i = lane; i < n; i += 32) {
      double r = 0.0, c = 0.0;
      for (int j = 0; j < n; j++) {
        r += Qb[i * n + j] * sv_[j];
        c += Qb[j * n + i] * sv_[j];
      }
      Qs[i] = r;
      sQ[i] = c;
      part += sv_[i] * vv[i];
    }
    double sv = warp_sum(part);  // s'u
    __syncwarp();
    part = 0.0;
    for (int i = lane; i < n; i += 32) part += sQ[i] * sv_[i];
    const double sQs = warp_sum(part);
    double gamma = a.gamma;
    if (gamma < 0.0) {  // damping adapted to the step length (:168-172)
      gamma = -gamma;
      gamma = gamma + (1.0 - gamma) * (1.0 - a.alpha);
    }
    if (sv < gamma * sQs) {  // Powell's modification (:176-181)
      const double theta = (1.0 - gamma) * sQs / (sQs - sv);
      part = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double v = theta * vv[i] + (1.0 - theta) * Qs[i];
        vv[i] = v;
        part += sv_[i] * v;
      }
      sv = warp_sum(part);
    }
    __syncwarp();
    if (!(sv != 0.0) || !(sQs != 0.0)) {  // (:186-187) block left as it is
      if (lane == 0) atomicAdd(&a.info[1], 1);
      continue;
    }
    // rank-2 update of the upper triangle, mirrored when eigenvalue control is on (:194-202)
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e - i * n;
      if (j >= i) {
        double q = Qb[e];
        q = __dsub_rn(q, __ddiv_rn(__dmul_rn(Qs[i], sQ[j]), sQs));
        q = __dadd_rn(q, __ddiv_rn(__dmul_rn(vv[i], vv[j]), sv));
        Qb[e] = q;
        if (a.eigen_control) Qb[j * n + i] = q;
      }
    }
    __syncwarp();
    double shift = 0.0;
    if (a.eigen_control) {
      double theta = a.eps * a.eps;
      if (sQs < theta && sQs >= 0.0) theta = sQs;
      for (int e = lane; e < n * n; e += 32) W[e] = Qb[e];
      __syncwarp();
      if (!warp_pd_certificate(W, n, theta)) {
        __syncwarp();
        for (int e = lane; e < n * n; e += 32) W[e] = Qb[e];
        __syncwarp();
        int sweeps;
        const double lmin = warp_jacobi_min_eig(W, n, rot, &sweeps) - theta;
        if (lmin < 0.0) shift = -lmin;
        if (lane == 0) {
          if (shift != 0.0) atomicAdd(&a.info[0], 1);
          if (sweeps >= 40) atomicAdd(&a.info[2], 1);
        }
      }
    }
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e - i * n;
      Qg[e] = Qb[e] + (i == j ? shift : 0.0);
    }
    __syncwarp();
  }
}

int fail(const std::string &m, int code) {
  g_err = m;
  return code;
}

#define CU(call)                                                         \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess)                                               \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_), HQPHL_E_CUDA); \
  } while (0)

int launch(const BfgsArgs &a, cudaStream_t st) {
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t per = ((size_t)2 * a.max_n * a.max_n + 6 * a.max_n + 8) * sizeof(double);
  int warps = (int)std::min<size_t>(8, (200 * 1024) / per);
  if (warps < 1) return fail("hqphl: block too large for the shared-memory kernel", HQPHL_E_UNSUPPORTED);
  CU(cudaFuncSetAttribute(hl_bfgs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per * warps)));
  const int ctas = std::max(1, std::min((a.nblocks + warps - 1) / warps, sms * 2));
  hl_bfgs_kernel<<<ctas, 32 * warps, per * warps, st>>>(a);
  CU(cudaGetLastError());
  return HQPHL_OK;
}

}  // namespace

extern "C" {

const char *hqphl_last_error(void) { return g_err.c_str(); }

int hqphl_bfgs_update_dev(void *cuda_stream, int nblocks, int max_bsize, const int *d_bsize,
                          const long long *d_qoff, const int *d_voff, double *d_Q, const double *d_s,
                          const double *d_u, double alpha, double gamma, double eps, int eigen_control,
                          int *d_info3) {
  if (nblocks < 0 || max_bsize < 1 || !d_bsize || !d_qoff || !d_voff || !d_Q || !d_s || !d_u || !d_info3)
    return fail("hqphl_bfgs_update_dev: bad argument", HQPHL_E_ARG);
  if (nblocks == 0) return HQPHL_OK;
  BfgsArgs a{nblocks, d_bsize, d_qoff, d_voff, d_Q, d_s, d_u, alpha, gamma, eps, eigen_control, max_bsize, d_info3};
  return launch(a, static_cast<cudaStream_t>(cuda_stream));
}

int hqphl_bfgs_update(int device, int nblocks, const int *bsize, double *Q, const double *s, const double *u,
                      double alpha, double gamma, double eps, int eigen_control, int *info3) {
  if (nblocks < 0 || (nblocks > 0 && (!bsize || !Q || !s || !u)))
    return fail("hqphl_bfgs_update: bad argument", HQPHL_E_ARG);
  if (info3) info3[0] = info3[1] = info3[2] = 0;
  if (nblocks == 0) return HQPHL_OK;
  CU(cudaSetDevice(device));
  std::vector<long long> qoff(nblocks);
  std::vector<int> voff(nblocks);
  long long nq = 0;
  int nv = 0, mx = 0;
  for (int b = 0; b < nblocks; b++) {
    if (bsize[b] < 1) return fail("hqphl_bfgs_update: block size < 1", HQPHL_E_ARG);
    qoff[b] = nq;
    voff[b] = nv;
    nq += (long long)bsize[b] * bsize[b];
    nv += bsize[b];
    mx = std::max(mx, bsize[b]);
  }
  int *d_bs = nullptr, *d_vo = nullptr, *d_info = nullptr;
  long long *d_qo = nullptr;
  double *d_Q = nullptr, *d_s = nullptr, *d_u = nullptr;
  int rc = HQPHL_OK;
  auto cleanup = [&]() {
    cudaFree(d_bs); cudaFree(d_vo); cudaFree(d_info); cudaFree(d_qo); cudaFree(d_Q); cudaFree(d_s); cudaFree(d_u);
  };
#define CUC(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      cleanup();                                                                         \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_), HQPHL_E_CUDA);     \
    }                                                                                    \
  } while (0)
  CUC(cudaMalloc(&d_bs, nblocks * sizeof(int)));
  CUC(cudaMalloc(&d_vo, nblocks * sizeof(int)));
  CUC(cudaMalloc(&d_qo, nblocks * sizeof(long long)));
  CUC(cudaMalloc(&d_info, 3 * sizeof(int)));
  CUC(cudaMalloc(&d_Q, nq * sizeof(double)));
  CUC(cudaMalloc(&d_s, nv * sizeof(double)));
  CUC(cudaMalloc(&d_u, nv * sizeof(double)));
  CUC(cudaMemcpy(d_bs, bsize, nblocks * sizeof(int), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(d_vo, voff.data(), nblocks * sizeof(int), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(d_qo, qoff.data(), nblocks * sizeof(long long), cudaMemcpyHostToDevice));
  CUC(cudaMemset(d_info, 0, 3 * sizeof(int)));
  CUC(cudaMemcpy(d_Q, Q, nq * sizeof(double), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(d_s, s, nv * sizeof(double), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(d_u, u, nv * sizeof(double), cudaMemcpyHostToDevice));
  rc = hqphl_bfgs_update_dev(nullptr, nblocks, mx, d_bs, d_qo, d_vo, d_Q, d_s, d_u, alpha, gamma, eps,
                             eigen_control, d_info);
  if (rc == HQPHL_OK) {
    CUC(cudaDeviceSynchronize());
    CUC(cudaMemcpy(Q, d_Q, nq * sizeof(double), cudaMemcpyDeviceToHost));
    if (info3) CUC(cudaMemcpy(info3, d_info, 3 * sizeof(int), cudaMemcpyDeviceToHost));
  }
  cleanup();
  return rc;
#undef CUC
}

}  // extern "C"
