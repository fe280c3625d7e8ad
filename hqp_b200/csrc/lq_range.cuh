// Horizon split across GPUs: each rank owns a contiguous stage
// range, condenses it to the top element of its trees and exchanges only
//   factor:  (A, C, J) of the range + the local terminal block   4 nx^2 doubles
//            Psi of the range (after K3)                           nx^2 doubles
//   solve:   (v0, v_term) backward, (x0, x_start) forward          2 nx doubles each
// with one all-gather each (done by the caller over NCCL).  Every rank then
// applies, redundantly, the elements of the ranks behind (before) it -- at most
// world-1 sequential steps -- and continues with its local down-sweeps.
#pragma once

#include "lq_device.cuh"
#include "lq_factor.cuh"

// xf[0..3][nx*nx] <- A, C, J of the single top element of the factor tree and the
// local terminal block Q_K + C_K'(z/w)C_K (zero when the caller passed none).
// One CTA.
// slot: where the element of the whole range lives (top of the binary tree, or
// suffix 0 of the suffix scan).
__global__ void range_export_factor_kernel(LqDev d, double *xf, int slot) {
  const int nx = d.nx, nm = d.nm, n2 = nx * nx;
  const size_t o = (size_t)slot * n2;  // batch == 1
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    xf[i] = d.segA[o + i];
    xf[n2 + i] = d.segC[o + i];
    xf[2 * n2 + i] = d.segJ[o + i];
    const int r = i / nx, c = i - r * nx;
    xf[3 * n2 + i] = d.Q[(size_t)d.K * nm * nm + r * nm + c] +
                     (r == c ? d.hdiag[(size_t)d.K * nm + r] : 0.0);
  }
  // general inequality rows of the terminal stage
  for (int rr = d.grow_ptr[d.K]; rr < d.grow_ptr[d.K + 1]; rr++) {
    __syncthreads();
    const int r = d.grow[rr];
    const int e0 = d.ineq_ptr[r], ne = d.ineq_ptr[r + 1] - e0;
    const double wz = d.z[r] / d.w[r];
    for (int e = threadIdx.x; e < ne * ne; e += blockDim.x) {
      const int ea = e / ne, eb = e - ea * ne;
      xf[3 * n2 + d.ineq_lcol[e0 + ea] * nx + d.ineq_lcol[e0 + eb]] +=
          wz * d.cval[e0 + ea] * d.cval[e0 + eb];
    }
  }
}

// Vext <- value Hessian at the end of this range: start from the terminal block
// of the last rank and apply the elements of the ranks world-1 .. rank+1
// (V_start = J + A'(I + S C)^{-1} S A).  gathered: [world][4][nx*nx].  One CTA.
template <int NX>
__global__ void __launch_bounds__(NX == 0 ? LQ_BIG_NT : LQ_NT2) range_scan_factor_kernel(LqDev d, const double *gathered,
                                                                int rank, int world) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, n2 = nx * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int ldm = NX > 0 ? 2 * nx + 1 : 2 * nx;  // odd row stride for the warp inverse
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  double *S = sm.take(n2), *A = sm.take(n2), *Cg = sm.take(n2);
  double *M = sm.take(nx * ldm), *X = sm.take(n2);
  double *inv_scr = NX > 0 ? sm.take(nx * (nx + 1) + 2 * (nx + 2)) : nullptr;
  __shared__ int st_s, piv_small[65];
  __shared__ double inv_s[2];
  int *piv_s = nx <= 64 ? piv_small : reinterpret_cast<int *>(sm.take((nx + 2) / 2));
  if (threadIdx.x == 0) st_s = 0;
  const double *last = gathered + (size_t)(world - 1) * 4 * n2;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) S[i] = last[3 * n2 + i];
  __syncthreads();
  for (int r = world - 1; r > rank; r--) {
    const double *E = gathered + (size_t)r * 4 * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      A[i] = E[i];
      Cg[i] = E[n2 + i];
    }
    __syncthreads();
    cta_mmx<TC, LQ_NT2 / 32>(stg, M, ldm, nullptr, 0, 0.0, 1.0, S, nx, 1, Cg, nx, 1, nx, nx, nx);
    cta_mmx<TC, LQ_NT2 / 32>(stg, M + nx, ldm, nullptr, 0, 0.0, 1.0, S, nx, 1, A, nx, 1, nx, nx, nx);
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) M[i * ldm + i] += 1.0;
    if constexpr (NX > 0)
      cta_inverse_apply<NX, LQ_NT2 / 32>(M, ldm, 2 * nx, X, inv_scr, piv_s, &st_s);
    else if (stg)
      cta_inverse_apply_big(stg, big_gj_scratch(d), M, ldm, nx, 2 * nx, X, nx, piv_s, &st_s);
    else
      cta_gauss_jordan<NX>(M, 2 * nx, nx, 2 * nx, X, piv_s, inv_s, &st_s);
    cta_mmx<TC, LQ_NT2 / 32>(stg, S, nx, E + 2 * n2, nx, 1.0, 1.0, A, 1, nx, X, nx, 1, nx, nx, nx);
    __syncthreads();
    cta_symmetrize(S, nx, nx);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n2; i += blockDim.x) d.Vext[i] = (rank == world - 1) ? 0.0 : S[i];
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, st_s);
}

// xv[0][nx] <- zero-boundary solution of the whole range (top element of the solve
// tree), xv[1][nx] <- this rank's boundary value: backward v[K] (the terminal
// gradient), forward the initial state x_0 (rank 0 only).
template <bool BACK>
__global__ void range_export_vec_kernel(LqDev d, const double *__restrict__ r2, double *xv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);
  const int nx = d.nx;
  const size_t top = (size_t)d.st.off[d.st.nlev - 1] * nx;
  const double *in0 = BACK ? d.segv0 : d.segx0;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xv[i] = in0[top + i];
  if (BACK) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xv[nx + i] = d.v[(size_t)d.K * nx + i];
  } else if (d.has_prev) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xv[nx + i] = 0.0;
  } else if (d.fixed_x0) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xv[nx + i] = -r2[(size_t)d.K * nx + i];
  } else {
    if (threadIdx.x < nx) t[threadIdx.x] = -d.v[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) thread_ldlt_solve(d.V0f, nx, nx, t, 1);
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xv[nx + i] = t[i];
  }
}

// Boundary vector of this range from the gathered pieces.
//   BACK: t = v_term(last rank); for r = world-1 .. rank+1: t = v0_r + Psi_r' t;
//         v[K] += t   (rank < world-1)
//   FWD : t = x_start(rank 0);   for r = 0 .. rank-1:       t = x0_r + Psi_r t;
//         xstart = t  (rank > 0)
// gv: [world][2][nx], gpsi: [world][nx*nx].  One CTA, nx <= blockDim.x.
// psi_stride: doubles between the transitions of consecutive ranks in gpsi.
template <bool BACK>
__global__ void range_scan_vec_kernel(LqDev d, const double *__restrict__ gv,
                                      const double *__restrict__ gpsi, int rank, int world,
                                      int psi_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *t = reinterpret_cast<double *>(smem_raw);  // nx
  double *u = t + d.nx;
  const int nx = d.nx, n2 = nx * nx;
  const int i = threadIdx.x;
  const int src = BACK ? world - 1 : 0;
  if (i < nx) t[i] = gv[((size_t)src * 2 + 1) * nx + i];
  __syncthreads();
  if (BACK) {
    for (int r = world - 1; r > rank; r--) {
      const double *P = gpsi + (size_t)r * psi_stride;
      if (i < nx) {
        double s = gv[(size_t)r * 2 * nx + i];
        for (int l = 0; l < nx; l++) s = fma(P[l * nx + i], t[l], s);
        u[i] = s;
      }
      __syncthreads();
      if (i < nx) t[i] = u[i];
      __syncthreads();
    }
    if (rank < world - 1 && i < nx) d.v[(size_t)d.K * nx + i] += t[i];
  } else {
    for (int r = 0; r < rank; r++) {
      const double *P = gpsi + (size_t)r * psi_stride;
      if (i < nx) {
        double s = gv[(size_t)r * 2 * nx + i];
        for (int l = 0; l < nx; l++) s = fma(P[i * nx + l], t[l], s);
        u[i] = s;
      }
      __syncthreads();
      if (i < nx) t[i] = u[i];
      __syncthreads();
    }
    if (rank > 0 && i < nx) d.xstart[i] = t[i];
  }
}

// Status word of this range appended to its transition block (slot n2 of xpsi), and
// the OR of all ranges' words back into d.status after the all-gather: every rank
// then takes the same branch on a singular / non-PD block anywhere on the horizon.
__global__ void range_status_pack_kernel(LqDev d, double *xpsi) {
  if (threadIdx.x == 0) xpsi[d.nx * d.nx] = (double)*d.status;
}
__global__ void range_status_merge_kernel(LqDev d, const double *gpsi, int world, int psi_stride) {
  if (threadIdx.x == 0) {
    int st = 0;
    for (int r = 0; r < world; r++) st |= (int)gpsi[(size_t)r * psi_stride + d.nx * d.nx];
    *d.status = st;
  }
}
