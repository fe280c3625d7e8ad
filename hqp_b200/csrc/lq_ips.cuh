// Interior-point vector kernels: the per-iteration residual, complementarity
// and step-length passes of Hqp_IpsMehrotra::step (hqp/Hqp_IpsMehrotra.C:355-693)
// on device-resident vectors, fused so that every vector is read once per pass.
//
// The reference spends these passes in Meschach primitives over the whole
// vectors (v_star, v_mltadd, in_prod, sp_mv_mlt ...).  Here the
// sparse mat-vecs use the stage structure (Q block diagonal, A block
// bidiagonal, C stage-local), so one CTA per stage produces the stage's part of
// r1..r4 plus its partial reductions; reductions are finished in a FIXED order
// by one CTA (bit-reproducible from run to run: iteration counts depend on mu
// and the gap).
#pragma once

#include "lq_device.cuh"
#include "lq_eq.cuh"

#define IPS_NQ 8  // reduced quantities per pass

struct IpsVec {
  const double *c, *b, *dv;  // QP vectors: linear cost, A x + b = 0, C x + d >= 0
  double *x, *y, *z, *w;
};

// block reduction of `v` (sum or max over the CTA); result valid in thread 0
__device__ __forceinline__ double block_reduce(double v, bool is_max, double *red) {
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, other) : v + other;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 1; i < nw; i++) v = is_max ? fmax(v, red[i]) : v + red[i];
  }
  return v;
}

// ---------------------------------------------------------------------------
// KKT residuals (hqp/Hqp_IpsMehrotra.C:425-445), stage k per CTA:
//   r1 = Q x + c - A' y - C' z     r2 = -(A x + b)
//   r3 = -(C x + d - w)            r4 = -z w
// partial[blk][0..7] = x'Qx, x'c, y'b, z'd, z'w, |r1|inf, |r2|inf, |r3|inf
// ety = E' y|eq (general equality rows; NULL if none).
// grid ceil((K+1) / IPS_RES_WPB), block 32 * IPS_RES_WPB, smem IPS_RES_WPB (nm + 2 nx) doubles
// ---------------------------------------------------------------------------
#define IPS_RES_WPB 4  // stages (warps) per CTA of the residual pass
__global__ void __launch_bounds__(32 * IPS_RES_WPB)
ips_residual_kernel(LqDev d, IpsVec v, double *r1, double *r2, double *r3, double *r4,
                    const double *__restrict__ ety, double *partial) {
  // One WARP per stage (round 1: one 64-thread CTA per stage with eight block
  // reductions, 16 CTA barriers -- 82 us at C2, launch- and barrier-bound): no CTA
  // barrier at all, the eight partial sums / maxima of the stage by warp shuffles.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *xs = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (nm + 2 * nx);  // nm
  double *xn = xs + nm;                                // nx : x_{k+1}
  double *yk = xn + nx;                                // nx : y dynamics rows k
  const int k = blockIdx.x * IPS_RES_WPB + warp;
  __shared__ double red[IPS_RES_WPB][IPS_NQ];
  double xQx = 0, xc = 0, yb = 0, zd = 0, zw = 0, n1 = 0, n2 = 0, n3 = 0;
  const int dk = (k < d.K) ? nm : nx;
  const size_t xo = (size_t)k * nm;
  if (k == d.K && d.has_next) {
    // horizon split: the trailing block duplicates the next range's x_0; its rows
    // (and its share of every reduction) belong to that rank
    for (int i = lane; i < nx; i += 32) r1[xo + i] = 0.0;
  } else if (k <= d.K) {
  const double *xnext = (d.has_next && d.halo && k == d.K - 1)
                            ? d.halo + (size_t)(d.rank + 1) * 2 * nx : nullptr;
  const double *yprev = (d.has_prev && d.halo && k == 0)
                            ? d.halo + (size_t)(d.rank - 1) * 2 * nx + nx : nullptr;
  for (int i = lane; i < dk; i += 32) xs[i] = v.x[xo + i];
  if (k < d.K)
    for (int i = lane; i < nx; i += 32) {
      xn[i] = xnext ? xnext[i] : v.x[xo + nm + i];
      yk[i] = v.y[(size_t)k * nx + i];
    }
  __syncwarp();
  const double *Qk = d.Q + (size_t)k * nm * nm;
  const double *cv = d.cval;
  for (int i = lane; i < dk; i += 32) {
    double qx = 0.0;
#pragma unroll 10
    for (int l = 0; l < dk; l++) qx = fma(Qk[l * nm + i], xs[l], qx);
    xQx = fma(xs[i], qx, xQx);
    xc = fma(xs[i], v.c[xo + i], xc);
    double s = qx + v.c[xo + i];
    if (k < d.K) {
      if (i < nx) {
        const double *fx = d.fx + (size_t)k * nx * nx;
#pragma unroll 10
        for (int l = 0; l < nx; l++) s = fma(-fx[l * nx + i], yk[l], s);
      } else {
        const double *fu = d.fu + (size_t)k * nx * nu;
#pragma unroll 10
        for (int l = 0; l < nx; l++) s = fma(-fu[l * nu + (i - nx)], yk[l], s);
      }
    }
    if (i < nx) {
      if (k > 0) s += v.y[(size_t)(k - 1) * nx + i];
      else if (d.fixed_x0) s -= v.y[(size_t)d.K * nx + i];
      else if (yprev) s += yprev[i];
    }
    const int gv = k * nm + i;
    for (int e = d.vcol_ptr[gv]; e < d.vcol_ptr[gv + 1]; e++)
      s = fma(-cv[d.vcol_nz[e]], v.z[d.vcol_row[e]], s);
    if (ety) s -= ety[xo + i];
    r1[xo + i] = s;
    n1 = fmax(n1, fabs(s));
  }
  if (k < d.K) {
    const double *fx = d.fx + (size_t)k * nx * nx, *fu = d.fu + (size_t)k * nx * nu;
    for (int i = lane; i < nx; i += 32) {
      double s = -xn[i];
#pragma unroll 10
      for (int l = 0; l < nx; l++) s = fma(fx[i * nx + l], xs[l], s);
#pragma unroll 10
      for (int l = 0; l < nu; l++) s = fma(fu[i * nu + l], xs[nx + l], s);
      const size_t ro = (size_t)k * nx + i;
      const double t = -(s + v.b[ro]);
      r2[ro] = t;
      yb = fma(yk[i], v.b[ro], yb);
      n2 = fmax(n2, fabs(t));
    }
  }
  if (k == 0 && d.fixed_x0)
    for (int i = lane; i < nx; i += 32) {
      const size_t ro = (size_t)d.K * nx + i;
      const double t = -(xs[i] + v.b[ro]);
      r2[ro] = t;
      yb = fma(v.y[ro], v.b[ro], yb);
      n2 = fmax(n2, fabs(t));
    }
  for (int rr = d.srow_ptr[k] + lane; rr < d.srow_ptr[k + 1]; rr += 32) {
    const int r = d.srow[rr];
    double s = 0.0;
    for (int e = d.ineq_ptr[r]; e < d.ineq_ptr[r + 1]; e++)
      s = fma(cv[e], xs[d.ineq_lcol[e]], s);
    const double zr = v.z[r], wr = v.w[r];
    const double t = -(s + v.dv[r] - wr);
    r3[r] = t;
    r4[r] = -zr * wr;
    zd = fma(zr, v.dv[r], zd);
    zw = fma(zr, wr, zw);
    n3 = fmax(n3, fabs(t));
  }
  for (int o = 16; o > 0; o >>= 1) {
    xQx += __shfl_xor_sync(0xffffffffu, xQx, o);
    xc += __shfl_xor_sync(0xffffffffu, xc, o);
    yb += __shfl_xor_sync(0xffffffffu, yb, o);
    zd += __shfl_xor_sync(0xffffffffu, zd, o);
    zw += __shfl_xor_sync(0xffffffffu, zw, o);
    n1 = fmax(n1, __shfl_xor_sync(0xffffffffu, n1, o));
    n2 = fmax(n2, __shfl_xor_sync(0xffffffffu, n2, o));
    n3 = fmax(n3, __shfl_xor_sync(0xffffffffu, n3, o));
  }
  }  // (stage of this warp)
  // one partial row per CTA: the stages of its warps combined in warp order
  if (lane == 0) {
    red[warp][0] = xQx; red[warp][1] = xc; red[warp][2] = yb; red[warp][3] = zd;
    red[warp][4] = zw;  red[warp][5] = n1; red[warp][6] = n2; red[warp][7] = n3;
  }
  __syncthreads();
  if (threadIdx.x < IPS_NQ) {
    const int j = threadIdx.x;
    double a = red[0][j];
    for (int w2 = 1; w2 < IPS_RES_WPB; w2++) a = (j >= 5) ? fmax(a, red[w2][j]) : a + red[w2][j];
    partial[(size_t)blockIdx.x * IPS_NQ + j] = a;
  }
}

// general equality rows of the residual pass; one CTA.  Writes r2|eq, ety = E'y|eq
// and its partial row (index nblk of `partial`): y'b and |r2|inf contributions.
__global__ void ips_eq_residual_kernel(LqDev d, LqEq q, IpsVec v, double *r2, double *partial_row) {
  __shared__ double red[32];
  const int n = q.n_eq, eq0 = d.me - n;
  for (int i = threadIdx.x; i < d.N; i += blockDim.x) q.ety[i] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < n; i++) {
      const double yi = v.y[eq0 + i];
      for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++)
        q.ety[(size_t)q.stage[i] * d.nm + q.lcol[e]] += q.val[e] * yi;
    }
  double yb = 0.0, n2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double *xk = v.x + (size_t)q.stage[i] * d.nm;
    double s = v.b[eq0 + i];
    for (int e = q.ptr[i]; e < q.ptr[i + 1]; e++) s = fma(q.val[e], xk[q.lcol[e]], s);
    r2[eq0 + i] = -s;
    yb = fma(v.y[eq0 + i], v.b[eq0 + i], yb);
    n2 = fmax(n2, fabs(s));
  }
  double t;
  t = block_reduce(yb, false, red);
  if (threadIdx.x == 0) {
    for (int j = 0; j < IPS_NQ; j++) partial_row[j] = 0.0;
    partial_row[2] = t;
  }
  t = block_reduce(n2, true, red);
  if (threadIdx.x == 0) partial_row[6] = t;
}

// Finish the reductions in a fixed order: out[j] = sum (j < nsum) or max over
// the nblk partial rows (nq <= 8 quantities per row).  One CTA, one pass: every
// thread folds all quantities of its rows, then warp shuffles and one shared
// round (the first version looped over the quantities with a 9-barrier tree each:
// 27 us per call at 10^4 rows, 37 calls per QP solve).
#define IPS_FIN_THREADS 1024
__global__ void __launch_bounds__(IPS_FIN_THREADS)
ips_finalize_kernel(const double *__restrict__ partial, int nblk, int nq, int nsum, double *out) {
  __shared__ double sh[IPS_FIN_THREADS / 32][8];
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.0;
  for (int i = threadIdx.x; i < nblk; i += IPS_FIN_THREADS) {
    const double *row = partial + (size_t)i * nq;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < nq) {
        const double pv = row[j];
        acc[j] = (j >= nsum) ? fmax(acc[j], pv) : acc[j] + pv;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (j < nq) {
      double a = acc[j];
      for (int o = 16; o > 0; o >>= 1) {
        const double b2 = __shfl_xor_sync(0xffffffffu, a, o);
        a = (j >= nsum) ? fmax(a, b2) : a + b2;
      }
      if (lane == 0) sh[warp][j] = a;
    }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < nq) {
        double a = sh[lane][j];
        for (int o = 16; o > 0; o >>= 1) {
          const double b2 = __shfl_xor_sync(0xffffffffu, a, o);
          a = (j >= nsum) ? fmax(a, b2) : a + b2;
        }
        if (lane == 0) out[j] = a;
      }
    }
  }
}

__device__ __forceinline__ void atomic_min_pos(double *addr, double val) {
  atomicMin(reinterpret_cast<unsigned long long *>(addr),
            static_cast<unsigned long long>(__double_as_longlong(val)));
}

// ---------------------------------------------------------------------------
// Ratio tests (hqp/Hqp_IpsMehrotra.C:566-574, 586-591, 604-612, 627-646):
//   out[0] = min over dz_i<0 of -z_i/dz_i      (zmin, +inf if none)
//   out[1] = min over dw_i<0 of -w_i/dw_i      (wmin)
//   out[2] = max over dz_i dw_i > 0 of dz_i dw_i / z_i / w_i   (Terlaky t)
// out must be initialised to {+inf, +inf, 0}.  min/max are exact and order
// independent, so atomics on the bit patterns keep the result reproducible.
// ---------------------------------------------------------------------------
__global__ void ips_ratio_kernel(int m, const double *__restrict__ z, const double *__restrict__ w,
                                 const double *__restrict__ dz, const double *__restrict__ dw,
                                 double *out) {
  double zmin = __longlong_as_double(0x7ff0000000000000LL), wmin = zmin, t = 0.0;
  auto one = [&](double zi, double wi, double a, double b) {
    if (a < 0.0) zmin = fmin(zmin, -zi / a);
    if (b < 0.0) wmin = fmin(wmin, -wi / b);
    if (a * b > 0.0) t = fmax(t, a * b / zi / wi);
  };
  // 128-bit loads over the even part of the vectors (cudaMalloc'ed: 256-byte aligned)
  const size_t m2 = (size_t)m >> 1;
  const double2 *z2 = reinterpret_cast<const double2 *>(z), *w2 = reinterpret_cast<const double2 *>(w);
  const double2 *a2 = reinterpret_cast<const double2 *>(dz), *b2 = reinterpret_cast<const double2 *>(dw);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m2;
       i += (size_t)gridDim.x * blockDim.x) {
    const double2 zi = z2[i], wi = w2[i], a = a2[i], b = b2[i];
    one(zi.x, wi.x, a.x, b.x);
    one(zi.y, wi.y, a.y, b.y);
  }
  if ((m & 1) && blockIdx.x == 0 && threadIdx.x == 0) one(z[m - 1], w[m - 1], dz[m - 1], dw[m - 1]);
  for (int o = 16; o > 0; o >>= 1) {
    zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
    t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  // one set of atomics per CTA, not per warp (three hot addresses)
  __shared__ double red[3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { red[0][warp] = zmin; red[1][warp] = wmin; red[2][warp] = t; }
  __syncthreads();
  if (warp == 0) {
    zmin = lane < nw ? red[0][lane] : __longlong_as_double(0x7ff0000000000000LL);
    wmin = lane < nw ? red[1][lane] : __longlong_as_double(0x7ff0000000000000LL);
    t = lane < nw ? red[2][lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
      wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
      t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    if (lane == 0) {
      atomic_min_pos(out + 0, zmin);
      atomic_min_pos(out + 1, wmin);
      atomic_max_nonneg(out + 2, t);
    }
  }
}

// first index attaining the minima found by ips_ratio_kernel (:636-645 keep the
// first strict minimum); idx initialised to {INT_MAX, INT_MAX}
__global__ void ips_argmin_kernel(int m, const double *__restrict__ z, const double *__restrict__ w,
                                  const double *__restrict__ dz, const double *__restrict__ dw,
                                  const double *__restrict__ mins, int *idx) {
  const double zmin = mins[0], wmin = mins[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x) {
    if (dz[i] < 0.0 && -z[i] / dz[i] == zmin) atomicMin(idx + 0, (int)i);
    if (dw[i] < 0.0 && -w[i] / dw[i] == wmin) atomicMin(idx + 1, (int)i);
  }
}

// out[0..3] = (z, dz, w, dw) at idx[1] (the blocking slack), out[4..7] at idx[0]
// (the blocking multiplier), zeros where no index was found; out[8], out[9] =
// idx[0], idx[1]: everything Mehrotra's step-length rule needs (:650-663) in one
// device-to-host copy
__global__ void ips_gather_blocking_kernel(int m, const int *__restrict__ idx,
                                           const double *__restrict__ z,
                                           const double *__restrict__ dz,
                                           const double *__restrict__ w,
                                           const double *__restrict__ dw, double *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int iz = idx[0], iw = idx[1];
  const bool vw = iw < m, vz = iz < m;
  out[0] = vw ? z[iw] : 0.0;
  out[1] = vw ? dz[iw] : 0.0;
  out[2] = vw ? w[iw] : 0.0;
  out[3] = vw ? dw[iw] : 0.0;
  out[4] = vz ? z[iz] : 0.0;
  out[5] = vz ? dz[iz] : 0.0;
  out[6] = vz ? w[iz] : 0.0;
  out[7] = vz ? dw[iz] : 0.0;
  out[8] = (double)iz;
  out[9] = (double)iw;
}

// horizon split: local first-minimum indices -> indices on the whole horizon
// (rows of the ranges are numbered range after range); out[0..1] as doubles
// (exact below 2^53), 1e300 where this range has no candidate
__global__ void ips_global_index_kernel(int m, long long row0, const int *__restrict__ idx,
                                        double *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out[0] = idx[0] < m ? (double)(row0 + idx[0]) : 1.0e300;
  out[1] = idx[1] < m ? (double)(row0 + idx[1]) : 1.0e300;
}

// as ips_gather_blocking_kernel with the winners given by their horizon-wide index
// gidx[0..1] (after the min-all-reduce): only the owning range writes the entries,
// so that a sum-all-reduce delivers them everywhere; out[8], out[9] = the indices
__global__ void ips_gather_blocking_dist_kernel(int m, long long row0,
                                                const double *__restrict__ gidx,
                                                const double *__restrict__ z,
                                                const double *__restrict__ dz,
                                                const double *__restrict__ w,
                                                const double *__restrict__ dw, double *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double gz = gidx[0], gw = gidx[1];
  const long long iz = gz < 1.0e299 ? (long long)gz - row0 : -1;
  const long long iw = gw < 1.0e299 ? (long long)gw - row0 : -1;
  const bool vw = iw >= 0 && iw < m, vz = iz >= 0 && iz < m;
  out[0] = vw ? z[iw] : 0.0;
  out[1] = vw ? dz[iw] : 0.0;
  out[2] = vw ? w[iw] : 0.0;
  out[3] = vw ? dw[iw] : 0.0;
  out[4] = vz ? z[iz] : 0.0;
  out[5] = vz ? dz[iz] : 0.0;
  out[6] = vz ? w[iz] : 0.0;
  out[7] = vz ? dw[iz] : 0.0;
  out[8] = gz;  // (not all-reduced: identical on every rank already)
  out[9] = gw;
}

// corrector right-hand side r4 = -(z w + dza dwa - smm)  (:597-600, :616-619)
__global__ void ips_corrector_rhs_kernel(int m, const double *__restrict__ z,
                                         const double *__restrict__ w,
                                         const double *__restrict__ dza,
                                         const double *__restrict__ dwa, double smm,
                                         double *r4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x)
    r4[i] = -(z[i] * w[i] + dza[i] * dwa[i] - smm);
}

// partial[blk][0] = sum (z + a dz)(w + a dw)   (mu_pl, :651-653)
// alpha = min(mins[0], mins[1]) is read from the device scalars the ratio test left.
__global__ void ips_mupl_kernel(int m, const double *__restrict__ mins, const double *__restrict__ z,
                                const double *__restrict__ w, const double *__restrict__ dz,
                                const double *__restrict__ dw, double *partial) {
  __shared__ double red[32];
  const double alpha = fmin(mins[0], mins[1]);
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x)
    s = fma(fma(alpha, dz[i], z[i]), fma(alpha, dw[i], w[i]), s);
  s = block_reduce(s, false, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// iterate update (:676-681): x,y,z,w += alpha d; partial[blk] = {z'w, |x|inf}
__global__ void ips_update_kernel(int N, int me, int m, double alpha, double *x,
                                  const double *__restrict__ dx, double *y,
                                  const double *__restrict__ dy, double *z,
                                  const double *__restrict__ dz, double *w,
                                  const double *__restrict__ dw, double *partial) {
  __shared__ double red[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double zw = 0.0, nx_ = 0.0;
  bool bad = false;
  for (size_t i = t0; i < (size_t)N; i += stride) {
    const double v = fma(alpha, dx[i], x[i]);
    x[i] = v;
    nx_ = fmax(nx_, fabs(v));
    bad |= !(fabs(v) <= 1.7976931348623157e308);
  }
  for (size_t i = t0; i < (size_t)me; i += stride) y[i] = fma(alpha, dy[i], y[i]);
  for (size_t i = t0; i < (size_t)m; i += stride) {
    const double zi = fma(alpha, dz[i], z[i]), wi = fma(alpha, dw[i], w[i]);
    z[i] = zi;
    w[i] = wi;
    zw = fma(zi, wi, zw);
  }
  if (bad) nx_ = __longlong_as_double(0x7ff0000000000000LL);
  zw = block_reduce(zw, false, red);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 2] = zw;
  nx_ = block_reduce(nx_, true, red);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 2 + 1] = nx_;
}

// cold start shift (hqp/Hqp_IpsMehrotra.C:299-315), pass 1:
//   partial[blk] = {sum dz, sum dw, |dz|inf, |dw|inf, max(-dz), max(-dw)}
__global__ void ips_cold_stats_kernel(int m, const double *__restrict__ dz,
                                      const double *__restrict__ dw, double *partial) {
  __shared__ double red[32];
  double sz = 0, sw = 0, az = 0, aw = 0, nz = 0, nw = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x) {
    sz += dz[i];
    sw += dw[i];
    az = fmax(az, fabs(dz[i]));
    aw = fmax(aw, fabs(dw[i]));
    nz = fmax(nz, -dz[i]);
    nw = fmax(nw, -dw[i]);
  }
  double *out = partial + (size_t)blockIdx.x * 6;
  double q;
  q = block_reduce(sz, false, red); if (threadIdx.x == 0) out[0] = q;
  q = block_reduce(sw, false, red); if (threadIdx.x == 0) out[1] = q;
  q = block_reduce(az, true, red);  if (threadIdx.x == 0) out[2] = q;
  q = block_reduce(aw, true, red);  if (threadIdx.x == 0) out[3] = q;
  q = block_reduce(nz, true, red);  if (threadIdx.x == 0) out[4] = q;
  q = block_reduce(nw, true, red);  if (threadIdx.x == 0) out[5] = q;
}

// z = dz + delz, w = dw + delw; optional partial[blk] = sum z w
__global__ void ips_shift_kernel(int m, double delz, double delw, const double *__restrict__ dz,
                                 const double *__restrict__ dw, double *z, double *w,
                                 double *partial) {
  __shared__ double red[32];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x) {
    const double zi = dz[i] + delz, wi = dw[i] + delw;
    if (z) { z[i] = zi; w[i] = wi; }
    s = fma(zi, wi, s);
  }
  s = block_reduce(s, false, red);
  if (partial && threadIdx.x == 0) partial[blockIdx.x] = s;
}

// y = a*x (+ fill) helpers for the cold-start right-hand sides
__global__ void ips_scale_copy_kernel(size_t n, double a, const double *__restrict__ x, double *y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = a * x[i];
}
__global__ void ips_fill_kernel(size_t n, double a, double *y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = a;
}

// norm_data pieces (hqp/Hqp_IpsMehrotra.C:459-461): row-sum norms of Q (stored
// upper triangle), A, C as Meschach's sp_norm_inf (meschach/addon2_hqp.c:724-741)
// sees them.  grid (K+1), block 64; partial[blk] = {nQ, nA, nC}
__global__ void ips_norm_data_kernel(LqDev d, double *partial) {
  __shared__ double red[32];
  const int nx = d.nx, nu = d.nu, nm = d.nm, k = blockIdx.x;
  const int dk = (k < d.K) ? nm : nx;
  double nq = 0, na = 0, nc = 0;
  const double *Qk = d.Q + (size_t)k * nm * nm;
  for (int i = threadIdx.x; i < dk; i += blockDim.x) {
    double s = 0.0;
    for (int j = i; j < dk; j++) s += fabs(Qk[i * nm + j]);
    nq = fmax(nq, s);
  }
  if (k < d.K) {
    const double *fx = d.fx + (size_t)k * nx * nx, *fu = d.fu + (size_t)k * nx * nu;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = 1.0;
      for (int l = 0; l < nx; l++) s += fabs(fx[i * nx + l]);
      for (int l = 0; l < nu; l++) s += fabs(fu[i * nu + l]);
      na = fmax(na, s);
    }
  }
  if (k == 0 && d.fixed_x0) na = fmax(na, 1.0);
  for (int rr = d.srow_ptr[k] + threadIdx.x; rr < d.srow_ptr[k + 1]; rr += blockDim.x) {
    const int r = d.srow[rr];
    double s = 0.0;
    for (int e = d.ineq_ptr[r]; e < d.ineq_ptr[r + 1]; e++) s += fabs(d.cval[e]);
    nc = fmax(nc, s);
  }
  double *out = partial + (size_t)blockIdx.x * 3;
  double q;
  q = block_reduce(nq, true, red); if (threadIdx.x == 0) out[0] = q;
  q = block_reduce(na, true, red); if (threadIdx.x == 0) out[1] = q;
  q = block_reduce(nc, true, red); if (threadIdx.x == 0) out[2] = q;
}

// |v|inf of a plain vector: partial[blk] = max
__global__ void ips_absmax_kernel(size_t n, const double *__restrict__ x, double *partial) {
  __shared__ double red[32];
  double a = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    a = fmax(a, fabs(x[i]));
  a = block_reduce(a, true, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = a;
}

// ---- Franke (hqp/Hqp_IpsFranke.C) ---------------------------------------------------
// right-hand sides of one iteration (:293-299): r_i = -zeta a_i, r4 = z w - mu
__global__ void ips_franke_rhs_kernel(int N, int me, int m, double zeta, double mu,
                                      const double *__restrict__ a1, const double *__restrict__ a2,
                                      const double *__restrict__ a3, const double *__restrict__ z,
                                      const double *__restrict__ w, double *r1, double *r2,
                                      double *r3, double *r4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)N; i += stride) r1[i] = -zeta * a1[i];
  for (size_t i = t0; i < (size_t)me; i += stride) r2[i] = -zeta * a2[i];
  for (size_t i = t0; i < (size_t)m; i += stride) {
    r3[i] = -zeta * a3[i];
    r4[i] = z[i] * w[i] - mu;
  }
}

// maximal feasible step (:316-334): out[0] = min over dz_i > 0 of z_i / dz_i and over
// dw_i > 0 of w_i / dw_i (the step is x -= alpha dx); out[0] initialised to +inf.
// (the reference's running test  z_i < val * dz_i  picks the same minimum)
__global__ void ips_franke_ratio_kernel(int m, const double *__restrict__ z,
                                        const double *__restrict__ w,
                                        const double *__restrict__ dz,
                                        const double *__restrict__ dw, double *out) {
  double v = __longlong_as_double(0x7ff0000000000000LL);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x) {
    const double a = dz[i], b = dw[i];
    if (a > 0.0) v = fmin(v, z[i] / a);
    if (b > 0.0) v = fmin(v, w[i] / b);
  }
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v >= 0.0) atomic_min_pos(out, v);
}

// cold start slacks (:184-188): w = (Ltilde + d) + 1e-10
__global__ void ips_franke_w0_kernel(int m, double Ltilde, const double *__restrict__ dvec, double *w) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)m;
       i += (size_t)gridDim.x * blockDim.x) {
    const double t = Ltilde + dvec[i];
    w[i] = t + 1e-10;
  }
}

__global__ void ips_add_scalar_kernel(size_t n, double a, double *y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] += a;
}
