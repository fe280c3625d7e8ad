// General stage equality rows  E dx = r2|eq  (terminal constraints, mixed
// state/control equalities; Hqp_IpLQDOCP::Formax hqp/Hqp_IpLQDOCP.C:1280-1321).
//
// The reference eliminates them inside the Riccati sweep with a rank-revealing
// null-space step per stage (GE_QP, meschach/addon_hqp.c:399-475), which is
// sequential in k.  Here they are handled by block elimination on top of the
// parallel-in-time factor, which stays untouched:
//     factor:  for every row j solve the LQ system with right-hand side
//              r1 = -E_j' (n_eq extra solves, all through the same kernels);
//              S = E [dx^1 .. dx^n_eq];  Sinv = S^{-1}   (n_eq x n_eq, one CTA)
//     step:    base solve, then  y = Sinv (r2|eq - E dx0),
//              (dx,dy,dz,dw) += sum_j y_j (dx,dy,dz,dw)^j,   dy|eq = y.
// Exact (not a penalty): the same KKT solution as the reference whenever the
// LQ problem without the equality rows is itself well posed (Guu > 0).
#pragma once

#include "lq_device.cuh"

struct LqEq {
  int n_eq, nnz;
  const int *stage, *ptr, *lcol;  // CSR over the stage-local columns
  const double *val;              // [nnz]
  double *DX, *DY, *DZ, *DW;      // [n_eq][N], [n_eq][me], [n_eq][m], [n_eq][m]
  double *S, *Sinv;               // [n_eq*n_eq]
  double *y;                      // [n_eq]
  double *ety;                    // [N]  E' dy|eq  (residuum)
};

// r1 (already zeroed, length N) <- -E_j' ; one CTA
__global__ void eq_unit_rhs_kernel(LqDev d, LqEq q, int j, double *r1) {
  const int k = q.stage[j];
  for (int e = q.ptr[j] + threadIdx.x; e < q.ptr[j + 1]; e += blockDim.x)
    r1[(size_t)k * d.nm + q.lcol[e]] = -q.val[e];
}

// S[i][j] = E_i . dx^j ; one thread per entry
__global__ void eq_schur_kernel(LqDev d, LqEq q) {
  const int n = q.n_eq;
  for (int t = threadIdx.x + blockIdx.x * blockDim.x; t < n * n; t += gridDim.x * blockDim.x) {
    const int i = t / n, j = t - i * n;
    const double *dxj = q.DX + (size_t)j * d.N + (size_t)q.stage[i] * d.nm;
    double s = 0.0;
    for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++) s = fma(q.val[e], dxj[q.lcol[e]], s);
    q.S[t] = s;
  }
}

// Sinv = S^{-1} by Gauss-Jordan with partial pivoting in shared memory (one CTA)
__global__ void eq_invert_kernel(LqDev d, LqEq q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = q.n_eq;
  double *M = reinterpret_cast<double *>(smem_raw);  // n x 2n
  double *X = M + (size_t)n * 2 * n;                  // n x n
  __shared__ int st_s, piv_s[64];
  __shared__ double inv_s[2];
  if (threadIdx.x == 0) st_s = 0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t - i * n;
    M[i * 2 * n + j] = q.S[t];
    M[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
  }
  cta_gauss_jordan<0>(M, 2 * n, n, 2 * n, X, piv_s, inv_s, &st_s);
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) q.Sinv[t] = X[t];
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, LQ_FLAG_SING);
}

// y = Sinv (r2|eq - E dx0) ; one CTA
__global__ void eq_multiplier_kernel(LqDev d, LqEq q, const double *__restrict__ r2,
                                     const double *__restrict__ dx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);  // n_eq
  const int n = q.n_eq;
  const int eq0 = d.me - n;  // equality rows are the last rows of r2 / dy
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double *dxk = dx + (size_t)q.stage[i] * d.nm;
    double s = r2[eq0 + i];
    for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++) s = fma(-q.val[e], dxk[q.lcol[e]], s);
    t[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s = fma(q.Sinv[i * n + j], t[j], s);
    q.y[i] = s;
  }
}

// out += sum_j y_j out^j for the four solution vectors; dy|eq = y
__global__ void eq_combine_kernel(LqDev d, LqEq q, double *dx, double *dy, double *dz,
                                  double *dw) {
  const int n = q.n_eq;
  const int eq0 = d.me - n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)d.N; i += stride) {
    double s = dx[i];
    for (int j = 0; j < n; j++) s = fma(q.y[j], q.DX[(size_t)j * d.N + i], s);
    dx[i] = s;
  }
  for (size_t i = t0; i < (size_t)d.me; i += stride) {
    if (i >= (size_t)eq0) {
      dy[i] = q.y[i - eq0];
    } else {
      double s = dy[i];
      for (int j = 0; j < n; j++) s = fma(q.y[j], q.DY[(size_t)j * d.me + i], s);
      dy[i] = s;
    }
  }
  for (size_t i = t0; i < (size_t)d.m; i += stride) {
    double sz = dz[i], sw = dw[i];
    for (int j = 0; j < n; j++) {
      sz = fma(q.y[j], q.DZ[(size_t)j * d.m + i], sz);
      sw = fma(q.y[j], q.DW[(size_t)j * d.m + i], sw);
    }
    dz[i] = sz;
    dw[i] = sw;
  }
}

// residuum pieces of the equality rows: ety = E' dy|eq (dense, length N) and
// max |r2|eq - E dx| into *res.  One CTA; n_eq is small.
__global__ void eq_residuum_kernel(LqDev d, LqEq q, const double *__restrict__ r2,
                                   const double *__restrict__ dx, const double *__restrict__ dy,
                                   double *t2, double *res) {
  const int n = q.n_eq;
  const int eq0 = d.me - n;
  for (int i = threadIdx.x; i < d.N; i += blockDim.x) q.ety[i] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {  // rows may share columns: serial scatter keeps it deterministic
    for (int i = 0; i < n; i++) {
      const double yi = dy[eq0 + i];
      for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++)
        q.ety[(size_t)q.stage[i] * d.nm + q.lcol[e]] += q.val[e] * yi;
    }
  }
  double mx = 0.0;
  bool bad = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double *dxk = dx + (size_t)q.stage[i] * d.nm;
    double s = r2[eq0 + i];
    for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++) s = fma(-q.val[e], dxk[q.lcol[e]], s);
    if (t2) t2[eq0 + i] = s;
    mx = fmax(mx, fabs(s));
    bad |= (s != s);
  }
  if (bad) mx = __longlong_as_double(0x7ff8000000000000LL);
  if (mx != 0.0) atomic_max_nonneg(res, mx);
}
