// Solve kernels: one KKT solve with the current factor
// (Hqp_IpLQDOCP::step hqp/Hqp_IpLQDOCP.C:869-976 + ExRiccatiSolveSc :2007-2182).
//
// The reference's two dependent sweeps carry, per stage, a BKP solve and four
// mat-vecs on the critical path.  Here everything that does not depend on the
// sweep is hoisted into stage-parallel passes, so that the sequential chains
// carry ONE nx x nx mat-vec per stage, and the chains themselves are cut into
// the same P segments as the factor (segment transition Psi_s from K3):
//
//   pre   (stage-parallel)  g = -r1 + C'((z r3 + r4)/w);  wv = gx - Rux' gu;
//                           q = Vxx[k+1] f_k;  v[K] = gx_K
//   back1 (segment chains)  v = wv + Phi'(v+ + q) from v_b = 0    -> segv0
//   back2 (P-step chain)    segvb[s] = segv0[s+1] + Psi[s+1]' segvb[s+1]
//   back3 (segment chains)  same chain from the true v_b, stores v[k]
//   mid   (stage-parallel)  Ru = Guu^{-1}(gu + fu'(v+ + q));  c = f - fu Ru
//   fwd1  (segment chains)  x+ = Phi x + c from x_a = 0            -> segx0
//   fwd2  (P-step chain)    segxa[s+1] = Psi[s] segxa[s] + segx0[s]
//   fwd3  (segment chains)  same chain from the true x_a, stores x[k]
//   post  (stage-parallel)  u = -(Rux x + Ru); p = Vxx+ x+ + v+; dx,dy,dw,dz
#pragma once

#include "lq_device.cuh"

// y(n) = [y0 +] A' t  or  A t for an n x n row-major matrix in GLOBAL memory,
// 4 lanes per output row, result valid in every lane of the quad.
template <bool TRANS>
__device__ __forceinline__ double quad_matvec(const double *__restrict__ A, const double *t,
                                              int n, int i, int part) {
  double s = 0.0;
  if (i < n) {
    if (TRANS) {
      for (int l = part; l < n; l += 4) s = fma(A[l * n + i], t[l], s);
    } else {
      for (int l = part; l < n; l += 4) s = fma(A[i * n + l], t[l], s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}

// ---- pre ------------------------------------------------------------------
// grid (K+1, batch), block >= nm threads
__global__ void solve_pre_kernel(LqDev d, const double *__restrict__ r1,
                                 const double *__restrict__ r2, const double *__restrict__ r3,
                                 const double *__restrict__ r4) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *gk = reinterpret_cast<double *>(smem_raw);  // nm
  double *fk = gk + d.nm;                              // nx
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int k = blockIdx.x, b = blockIdx.y;
  const int dk = (k < d.K) ? nm : nx;
  const double *z = d.z + (size_t)b * d.m, *w = d.w + (size_t)b * d.m;
  const double *cv = d.cval + (size_t)b * d.nnz;
  const double *r3b = r3 + (size_t)b * d.m, *r4b = r4 + (size_t)b * d.m;
  const size_t xo = (size_t)b * d.N + (size_t)k * nm;
  for (int i = threadIdx.x; i < dk; i += blockDim.x) {
    double s = -r1[xo + i];
    const int gv = k * nm + i;
    for (int e = d.vcol_ptr[gv]; e < d.vcol_ptr[gv + 1]; e++) {
      const int r = d.vcol_row[e];
      s = fma(cv[d.vcol_nz[e]], (z[r] * r3b[r] + r4b[r]) / w[r], s);
    }
    gk[i] = s;
    d.g[xo + i] = s;
  }
  if (k < d.K)
    for (int i = threadIdx.x; i < nx; i += blockDim.x)
      fk[i] = r2[(size_t)b * d.me + (size_t)k * nx + i];
  __syncthreads();
  if (k == d.K) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x)
      d.v[((size_t)b * (d.K + 1) + k) * nx + i] = gk[i];
    return;
  }
  const size_t ks = (size_t)b * d.K + k;
  const double *Rux = d.Rux + ks * nu * nx;
  const double *Vp = d.V + ((size_t)b * (d.K + 1) + k + 1) * nx * nx;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) {
    double s = gk[i];
    for (int l = 0; l < nu; l++) s = fma(-Rux[l * nx + i], gk[nx + l], s);
    d.wv[ks * nx + i] = s;
    double t = 0.0;
    for (int l = 0; l < nx; l++) t = fma(Vp[l * nx + i], fk[l], t);  // Vxx symmetric
    d.q[ks * nx + i] = t;
  }
}

// ---- backward chains -------------------------------------------------------
// grid (P, batch), block = 4 * ceil32(nx) threads.
// mode 0: start from 0, write segv0[s]; mode 1: start from segvb[s], store v[k].
__global__ void solve_back_kernel(LqDev d, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);  // nx
  const int nx = d.nx;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  const size_t so = ((size_t)b * d.P + s) * nx;
  double vi = 0.0;
  if (mode == 1 && i < nx) vi = d.segvb[so + i];
  for (int k = kb - 1; k >= ka; k--) {
    const size_t ks = (size_t)b * d.K + k;
    if (i < nx && part == 0) t[i] = vi + d.q[ks * nx + i];
    __syncthreads();
    const double a = quad_matvec<true>(d.Phi + ks * nx * nx, t, nx, i, part);
    if (i < nx) {
      vi = d.wv[ks * nx + i] + a;
      if (mode == 1 && part == 0) d.v[((size_t)b * (d.K + 1) + k) * nx + i] = vi;
    }
    __syncthreads();
  }
  if (mode == 0 && i < nx && part == 0) d.segv0[so + i] = vi;
}

// grid (batch): segvb[P-1] = v[K]; segvb[s] = segv0[s+1] + Psi[s+1]' segvb[s+1]
__global__ void solve_back_scan_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);
  const int nx = d.nx, b = blockIdx.x;
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  double vi = 0.0;
  if (i < nx) vi = d.v[((size_t)b * (d.K + 1) + d.K) * nx + i];
  for (int s = d.P - 1; s >= 0; s--) {
    const size_t so = ((size_t)b * d.P + s) * nx;
    if (i < nx && part == 0) {
      d.segvb[so + i] = vi;
      t[i] = vi;
    }
    __syncthreads();
    const double a = quad_matvec<true>(d.segPsi + so * nx, t, nx, i, part);
    if (i < nx) vi = d.segv0[so + i] + a;
    __syncthreads();
  }
}

// ---- mid -------------------------------------------------------------------
// grid (K, batch), block >= max(nx,nu) threads
__global__ void solve_mid_kernel(LqDev d, const double *__restrict__ r2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  double *t = reinterpret_cast<double *>(smem_raw);  // nx
  double *Gu = t + nx;                               // nu
  const int k = blockIdx.x, b = blockIdx.y;
  const size_t ks = (size_t)b * d.K + k;
  const double *vp = d.v + ((size_t)b * (d.K + 1) + k + 1) * nx;
  const double *fu = d.fu + ks * nx * nu;
  const double *g = d.g + (size_t)b * d.N + (size_t)k * nm;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) t[i] = vp[i] + d.q[ks * nx + i];
  __syncthreads();
  for (int j = threadIdx.x; j < nu; j += blockDim.x) {
    double s = g[nx + j];
    for (int l = 0; l < nx; l++) s = fma(fu[l * nu + j], t[l], s);
    Gu[j] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) thread_ldlt_solve(d.LD + ks * nu * nu, nu, nu, Gu, 1);
  __syncthreads();
  for (int j = threadIdx.x; j < nu; j += blockDim.x) d.Ru[ks * nu + j] = Gu[j];
  for (int i = threadIdx.x; i < nx; i += blockDim.x) {
    double s = r2[(size_t)b * d.me + (size_t)k * nx + i];
    for (int l = 0; l < nu; l++) s = fma(-fu[i * nu + l], Gu[l], s);
    d.c[ks * nx + i] = s;
  }
}

// ---- forward chains ---------------------------------------------------------
// mode 0: from x_a = 0, write segx0[s] (= x at segment end);
// mode 1: from segxa[s], store x[k], k = a..b
__global__ void solve_fwd_kernel(LqDev d, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);
  const int nx = d.nx;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  const size_t so = ((size_t)b * d.P + s) * nx;
  double xi = 0.0;
  if (mode == 1 && i < nx) {
    xi = d.segxa[so + i];
    if (part == 0) d.x[((size_t)b * (d.K + 1) + ka) * nx + i] = xi;
  }
  for (int k = ka; k < kb; k++) {
    const size_t ks = (size_t)b * d.K + k;
    if (i < nx && part == 0) t[i] = xi;
    __syncthreads();
    const double a = quad_matvec<false>(d.Phi + ks * nx * nx, t, nx, i, part);
    if (i < nx) {
      xi = d.c[ks * nx + i] + a;
      if (mode == 1 && part == 0) d.x[((size_t)b * (d.K + 1) + k + 1) * nx + i] = xi;
    }
    __syncthreads();
  }
  if (mode == 0 && i < nx && part == 0) d.segx0[so + i] = xi;
}

// grid (batch): x_0 (fixed: -a_0, hqp/Hqp_IpLQDOCP.C:2099-2100; free:
// -Vxx[0]^{-1} v[0], :2111-2117), then segxa[s+1] = Psi[s] segxa[s] + segx0[s]
__global__ void solve_fwd_scan_kernel(LqDev d, const double *__restrict__ r2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);
  const int nx = d.nx, b = blockIdx.x;
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  double xi = 0.0;
  if (d.fixed_x0) {
    if (i < nx) xi = -r2[(size_t)b * d.me + (size_t)d.K * nx + i];
  } else {
    if (threadIdx.x < nx) t[threadIdx.x] = -d.v[(size_t)b * (d.K + 1) * nx + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) thread_ldlt_solve(d.V0f + (size_t)b * nx * nx, nx, nx, t, 1);
    __syncthreads();
    if (i < nx) xi = t[i];
    __syncthreads();
  }
  for (int s = 0; s < d.P; s++) {
    const size_t so = ((size_t)b * d.P + s) * nx;
    if (i < nx && part == 0) {
      d.segxa[so + i] = xi;
      t[i] = xi;
    }
    __syncthreads();
    const double a = quad_matvec<false>(d.segPsi + so * nx, t, nx, i, part);
    if (i < nx) xi = d.segx0[so + i] + a;
    __syncthreads();
  }
}

// ---- post -------------------------------------------------------------------
// grid (K+1, batch), block >= nm threads
__global__ void solve_post_kernel(LqDev d, const double *__restrict__ r3,
                                  const double *__restrict__ r4, double *__restrict__ dx,
                                  double *__restrict__ dy, double *__restrict__ dz,
                                  double *__restrict__ dw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  double *xs = reinterpret_cast<double *>(smem_raw);  // nm : dx of this stage
  double *xn = xs + nm;                                // nx : x[k+1]
  const int k = blockIdx.x, b = blockIdx.y;
  const double *xk = d.x + ((size_t)b * (d.K + 1) + k) * nx;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = xk[i];
  if (k < d.K)
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xn[i] = xk[nx + i];
  __syncthreads();
  const size_t ks = (size_t)b * d.K + k;
  if (k < d.K) {
    const double *Rux = d.Rux + ks * nu * nx;
    for (int j = threadIdx.x; j < nu; j += blockDim.x) {
      double s = d.Ru[ks * nu + j];
      for (int l = 0; l < nx; l++) s = fma(Rux[j * nx + l], xs[l], s);
      xs[nx + j] = -s;  // u_k
    }
    // p_k = Vxx[k+1] x[k+1] + v[k+1]   (:2169-2171)
    const double *Vp = d.V + ((size_t)b * (d.K + 1) + k + 1) * nx * nx;
    const double *vp = d.v + ((size_t)b * (d.K + 1) + k + 1) * nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = vp[i];
      for (int l = 0; l < nx; l++) s = fma(Vp[l * nx + i], xn[l], s);
      dy[(size_t)b * d.me + (size_t)k * nx + i] = s;
    }
  }
  if (k == 0 && d.fixed_x0) {  // y_0 = -(Vx[0] + Vxx[0] x_0)   (:2153-2159)
    const double *V0 = d.V + (size_t)b * (d.K + 1) * nx * nx;
    const double *v0 = d.v + (size_t)b * (d.K + 1) * nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = v0[i];
      for (int l = 0; l < nx; l++) s = fma(V0[l * nx + i], xs[l], s);
      dy[(size_t)b * d.me + (size_t)d.K * nx + i] = -s;
    }
  }
  __syncthreads();
  const int dk = (k < d.K) ? nm : nx;
  // dx = -[x;u]   (:952)
  for (int i = threadIdx.x; i < dk; i += blockDim.x) {
    xs[i] = -xs[i];
    dx[(size_t)b * d.N + (size_t)k * nm + i] = xs[i];
  }
  __syncthreads();
  // dw = C dx - r3 ; dz = (r4 - z dw)/w   (:955-960)
  const double *cv = d.cval + (size_t)b * d.nnz;
  for (int rr = d.srow_ptr[k] + threadIdx.x; rr < d.srow_ptr[k + 1]; rr += blockDim.x) {
    const int r = d.srow[rr];
    double s = 0.0;
    for (int e = d.ineq_ptr[r]; e < d.ineq_ptr[r + 1]; e++)
      s = fma(cv[e], xs[d.ineq_lcol[e]], s);
    const size_t ro = (size_t)b * d.m + r;
    const double dwr = s - r3[ro];
    dw[ro] = dwr;
    dz[ro] = (r4[ro] - d.z[ro] * dwr) / d.w[ro];
  }
}

// ---- residuum ----------------------------------------------------------------
// Hqp_IpMatrix::residuum (hqp/Hqp_IpMatrix.C:131-178); grid (K+1, batch).
// Writes the four residual vectors to t1..t4 (may be NULL) and atomically
// maxes their inf-norm into *res (must be zeroed before the launch).
__global__ void residuum_kernel(LqDev d, const double *__restrict__ r1,
                                const double *__restrict__ r2, const double *__restrict__ r3,
                                const double *__restrict__ r4, const double *__restrict__ dx,
                                const double *__restrict__ dy, const double *__restrict__ dz,
                                const double *__restrict__ dw, double *t1, double *t2,
                                double *t3, double *t4, double *res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  double *xs = reinterpret_cast<double *>(smem_raw);  // nm: dx stage k
  double *yk = xs + nm;                                // nx: dy dynamics rows k
  __shared__ double red[32];
  const int k = blockIdx.x, b = blockIdx.y;
  const int dk = (k < d.K) ? nm : nx;
  const size_t xo = (size_t)b * d.N + (size_t)k * nm;
  const size_t yo = (size_t)b * d.me;
  for (int i = threadIdx.x; i < dk; i += blockDim.x) xs[i] = dx[xo + i];
  if (k < d.K)
    for (int i = threadIdx.x; i < nx; i += blockDim.x) yk[i] = dy[yo + (size_t)k * nx + i];
  __syncthreads();
  double mx = 0.0;
  bool bad = false;  // NaN seen
  const double *Qk = d.Q + ((size_t)b * (d.K + 1) + k) * nm * nm;
  const double *cv = d.cval + (size_t)b * d.nnz;
  const size_t ks = (size_t)b * d.K + k;
  // t1 = r1 + Q dx - A' dy - C' dz, rows of stage k
  for (int i = threadIdx.x; i < dk; i += blockDim.x) {
    double s = r1[xo + i];
    for (int l = 0; l < dk; l++) s = fma(Qk[l * nm + i], xs[l], s);  // Q symmetric
    if (k < d.K) {
      if (i < nx) {
        const double *fx = d.fx + ks * nx * nx;
        for (int l = 0; l < nx; l++) s = fma(-fx[l * nx + i], yk[l], s);
      } else {
        const double *fu = d.fu + ks * nx * nu;
        for (int l = 0; l < nx; l++) s = fma(-fu[l * nu + (i - nx)], yk[l], s);
      }
    }
    if (i < nx) {
      if (k > 0) s += dy[yo + (size_t)(k - 1) * nx + i];  // -(-I)' dy_{k-1}
      else if (d.fixed_x0) s -= dy[yo + (size_t)d.K * nx + i];
    }
    const int gv = k * nm + i;
    for (int e = d.vcol_ptr[gv]; e < d.vcol_ptr[gv + 1]; e++)
      s = fma(-cv[d.vcol_nz[e]], dz[(size_t)b * d.m + d.vcol_row[e]], s);
    if (t1) t1[xo + i] = s;
    mx = fmax(mx, fabs(s));
    bad |= (s != s);
  }
  // t2 = r2 - A dx, dynamics rows of stage k (+ x0 rows with stage 0)
  if (k < d.K) {
    const double *fx = d.fx + ks * nx * nx, *fu = d.fu + ks * nx * nu;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = -dx[xo + nm + i];
      for (int l = 0; l < nx; l++) s = fma(fx[i * nx + l], xs[l], s);
      for (int l = 0; l < nu; l++) s = fma(fu[i * nu + l], xs[nx + l], s);
      const double t = r2[yo + (size_t)k * nx + i] - s;
      if (t2) t2[yo + (size_t)k * nx + i] = t;
      mx = fmax(mx, fabs(t));
      bad |= (t != t);
    }
  }
  if (k == 0 && d.fixed_x0)
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const double t = r2[yo + (size_t)d.K * nx + i] - xs[i];
      if (t2) t2[yo + (size_t)d.K * nx + i] = t;
      mx = fmax(mx, fabs(t));
      bad |= (t != t);
    }
  // t3 = r3 - (C dx - dw) ; t4 = r4 - (z dw + w dz), rows of stage k
  for (int rr = d.srow_ptr[k] + threadIdx.x; rr < d.srow_ptr[k + 1]; rr += blockDim.x) {
    const int r = d.srow[rr];
    double s = 0.0;
    for (int e = d.ineq_ptr[r]; e < d.ineq_ptr[r + 1]; e++)
      s = fma(cv[e], xs[d.ineq_lcol[e]], s);
    const size_t ro = (size_t)b * d.m + r;
    const double a3 = r3[ro] - (s - dw[ro]);
    const double a4 = r4[ro] - (d.z[ro] * dw[ro] + d.w[ro] * dz[ro]);
    if (t3) t3[ro] = a3;
    if (t4) t4[ro] = a4;
    mx = fmax(mx, fmax(fabs(a3), fabs(a4)));
    bad |= (a3 != a3) | (a4 != a4);
  }
  if (bad) mx = __longlong_as_double(0x7ff8000000000000LL);
  // block max; NaN is propagated explicitly (fmax would drop it)
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (mx != mx) ? mx : ((other != other) ? other : fmax(mx, other));
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 1; i < nw; i++) {
      const double other = red[i];
      mx = (mx != mx) ? mx : ((other != other) ? other : fmax(mx, other));
    }
    atomic_max_nonneg(res, mx);
  }
}

// y_i += alpha * x_i on the four solution vectors (refinement update,
// hqp/Hqp_IpMatrix.C:103-106)
__global__ void axpy4_kernel(double alpha, const double *__restrict__ x1, double *y1, size_t n1,
                             const double *__restrict__ x2, double *y2, size_t n2,
                             const double *__restrict__ x3, double *y3, size_t n3,
                             const double *__restrict__ x4, double *y4, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n1; i += stride) y1[i] = fma(alpha, x1[i], y1[i]);
  for (size_t i = t0; i < n2; i += stride) y2[i] = fma(alpha, x2[i], y2[i]);
  for (size_t i = t0; i < n3; i += stride) y3[i] = fma(alpha, x3[i], y3[i]);
  for (size_t i = t0; i < n4; i += stride) y4[i] = fma(alpha, x4[i], y4[i]);
}
