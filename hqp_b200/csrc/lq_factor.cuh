// Factor kernels: parallel-in-time Riccati for the stage-structured KKT system.
//
// The reference sweeps the horizon sequentially (ExRiccatiFactorSc,
// hqp/Hqp_IpLQDOCP.C:1811-1969: for k = K-1..0, each stage needs Vxx[k+1]).
// Here the horizon of every instance is cut into P segments of L stages:
//
//   K1 seg_element_kernel   (P x batch CTAs)  each segment is condensed, with a
//        ZERO terminal cost, into its boundary element (A, C, J):
//            J = Riccati value Hessian at the segment start,
//            A = closed-loop transition across the segment,
//            C = closed-loop controllability Gramian weighted by Guu^{-1},
//        i.e. the segment's Schur complement onto (x_start, costate_end).
//   K2 seg_scan_kernel      (batch CTAs)  back-substitutes across segments:
//            Vb_s = V at the end of segment s,
//            V at its start = J + A' (I + Vb C)^{-1} Vb A      (exact identity)
//   K3 seg_riccati_kernel   (P x batch CTAs)  the ordinary Riccati recursion
//        inside every segment from its now-known terminal Vb, storing
//        Vxx[k], Rux[k], LDL'(Guu[k]), Phi[k] = fx - fu Rux and the segment
//        transition Psi_s = Phi[b-1] ... Phi[a].
//
// With P = 1 only K2 (terminal block) and K3 run: the reference's sequential
// sweep, one CTA per instance (the batched-MPC configuration).
#pragma once

#include "lq_device.cuh"

// shared-memory carve-up helper
struct SmemCarver {
  double *p;
  __device__ explicit SmemCarver(void *base) : p(reinterpret_cast<double *>(base)) {}
  __device__ double *take(int n) {
    double *r = p;
    p += (n + 1) & ~1;  // keep 16-byte alignment
    return r;
  }
};

// G(nm x nm) <- Q_k + C_k' diag(z/w) C_k for stage k of instance b
// (factor prologue hqp/Hqp_IpLQDOCP.C:805-832 + CTDC :68-103), and
// F(nx x nm) <- [fx_k fu_k] when k < K.  Ends with __syncthreads().
__device__ __forceinline__ void load_stage(const LqDev &d, int b, int k, double *G, double *F) {
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const double *Qk = d.Q + ((size_t)b * (d.K + 1) + k) * nm * nm;
  for (int i = threadIdx.x; i < nm * nm; i += blockDim.x) G[i] = Qk[i];
  if (k < d.K && F) {
    const double *fx = d.fx + ((size_t)b * d.K + k) * nx * nx;
    const double *fu = d.fu + ((size_t)b * d.K + k) * nx * nu;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
      const int r = i / nx, c = i - r * nx;
      F[r * nm + c] = fx[i];
    }
    for (int i = threadIdx.x; i < nx * nu; i += blockDim.x) {
      const int r = i / nu, c = i - r * nu;
      F[r * nm + nx + c] = fu[i];
    }
  }
  __syncthreads();
  const int r0 = d.srow_ptr[k], r1 = d.srow_ptr[k + 1];
  if (r1 > r0) {
    const double *cv = d.cval + (size_t)b * d.nnz;
    const double *z = d.z + (size_t)b * d.m, *w = d.w + (size_t)b * d.m;
    for (int rr = r0; rr < r1; rr++) {
      const int r = d.srow[rr];
      const int e0 = d.ineq_ptr[r], ne = d.ineq_ptr[r + 1] - e0;
      const double wz = z[r] / w[r];
      if (ne == 1) {  // simple bound: diagonal update
        if (threadIdx.x == 0) {
          const int c = d.ineq_lcol[e0];
          const double a = cv[e0];
          G[c * nm + c] += wz * a * a;
        }
      } else {
        for (int e = threadIdx.x; e < ne * ne; e += blockDim.x) {
          const int ea = e / ne, eb = e - ea * ne;
          G[d.ineq_lcol[e0 + ea] * nm + d.ineq_lcol[e0 + eb]] +=
              wz * cv[e0 + ea] * cv[e0 + eb];
        }
      }
      // rows of one stage may hit the same entries: serialise them
      __syncthreads();
    }
  }
}

// One Riccati stage on shared-memory blocks (FormGxx hqp/Hqp_IpLQDOCP.C:1077-1111
// + unconstrained-u branch :1854-1882 + Vxx :1940-1961).
//   in : V (nx x nx) = Vxx[k+1] (ignored when zero_V), F = [fx fu], G = H_k
//   out: G = [Gxx Gxu; Gux LDL'(Guu)], Rux (nu x nx), V <- Vxx[k] (symmetric),
//        Phi (nx x nx) = fx - fu Rux
// T is scratch (nx x nm).  st_s: shared status word.
__device__ __forceinline__ void riccati_stage(int nx, int nu, bool zero_V, double *V,
                                              const double *F, double *G, double *T,
                                              double *Rux, double *Phi, int *st_s) {
  const int nm = nx + nu;
  if (!zero_V) {
    // T = V F ; G += F' T
    cta_mm(T, nm, nullptr, 0, 0.0, 1.0, V, nx, 1, F, nm, 1, nx, nm, nx);
    __syncthreads();
    cta_mm(G, nm, G, nm, 1.0, 1.0, F, 1, nm, T, nm, 1, nm, nm, nx);
    __syncthreads();
    cta_symmetrize(G, nm, nm);
    __syncthreads();
  }
  double *Guu = G + nx * nm + nx;
  if (threadIdx.x < 32) {
    const int st = warp_ldlt(Guu, nm, nu);
    if (st && threadIdx.x == 0) atomicOr(st_s, st);
  }
  __syncthreads();
  // Rux = Guu^{-1} Gux : one right-hand side (column of Gux) per thread
  for (int j = threadIdx.x; j < nx; j += blockDim.x) {
    for (int i = 0; i < nu; i++) Rux[i * nx + j] = G[(nx + i) * nm + j];
    thread_ldlt_solve(Guu, nm, nu, Rux + j, nx);
  }
  __syncthreads();
  // V = Gxx - Gxu Rux ; Phi = fx - fu Rux
  cta_mm(V, nx, G, nm, 1.0, -1.0, G + nx, nm, 1, Rux, nx, 1, nx, nx, nu);
  cta_mm(Phi, nx, F, nm, 1.0, -1.0, F + nx, nm, 1, Rux, nx, 1, nx, nx, nu);
  __syncthreads();
  cta_symmetrize(V, nx, nx);
  __syncthreads();
}

// ---------------------------------------------------------------------------
// K1: condense segment s of instance b with zero terminal cost.
// ---------------------------------------------------------------------------
__global__ void seg_element_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  SmemCarver sm(smem_raw);
  double *J = sm.take(nx * nx), *A0 = sm.take(nx * nx), *A1 = sm.take(nx * nx);
  double *Cg = sm.take(nx * nx), *F = sm.take(nx * nm), *T = sm.take(nx * nm);
  double *G = sm.take(nm * nm), *Rux = sm.take(nu * nx), *Phi = sm.take(nx * nx);
  double *W = sm.take(nx * nu), *Y = sm.take(nx * nu);
  __shared__ int st_s;
  if (threadIdx.x == 0) st_s = 0;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    J[i] = 0.0;
    Cg[i] = 0.0;
    A0[i] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  double *A = A0, *An = A1;
  for (int k = kb - 1; k >= ka; k--) {
    load_stage(d, b, k, G, F);
    riccati_stage(nx, nu, k == kb - 1, J, F, G, T, Rux, Phi, &st_s);
    // W = A fu ; Y = W Guu^{-1} ; C += Y W' ; A <- A Phi
    cta_mm(W, nu, nullptr, 0, 0.0, 1.0, A, nx, 1, F + nx, nm, 1, nx, nu, nx);
    cta_mm(An, nx, nullptr, 0, 0.0, 1.0, A, nx, 1, Phi, nx, 1, nx, nx, nx);
    __syncthreads();
    const double *LD = G + nx * nm + nx;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      for (int j = 0; j < nu; j++) Y[i * nu + j] = W[i * nu + j];
      thread_ldlt_solve(LD, nm, nu, Y + i * nu, 1);
    }
    __syncthreads();
    cta_mm(Cg, nx, Cg, nx, 1.0, 1.0, Y, nu, 1, W, 1, nu, nx, nx, nu);
    double *t = A; A = An; An = t;
    __syncthreads();
  }
  cta_symmetrize(Cg, nx, nx);
  __syncthreads();
  const size_t o = ((size_t)b * d.P + s) * nx * nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
    d.segA[o + i] = A[i];
    d.segC[o + i] = Cg[i];
    d.segJ[o + i] = J[i];
  }
  // a non-positive pivot here comes from the artificial zero terminal cost
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, LQ_FLAG_NOTPD);
}

// ---------------------------------------------------------------------------
// K2: terminal block + back-substitution across the P segment elements.
// ---------------------------------------------------------------------------
__global__ void seg_scan_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nm = d.nm;
  const int b = blockIdx.x;
  SmemCarver sm(smem_raw);
  double *S = sm.take(nx * nx), *A = sm.take(nx * nx), *Cg = sm.take(nx * nx);
  double *M = sm.take(nx * 2 * nx);
  double *G = sm.take(nm * nm);
  __shared__ int st_s, piv_s[2];
  if (threadIdx.x == 0) st_s = 0;
  // terminal stage: Vxx[K] = H_K (hqp/Hqp_IpLQDOCP.C:1800-1804)
  load_stage(d, b, d.K, G, nullptr);
  double *VK = d.V + ((size_t)b * (d.K + 1) + d.K) * nx * nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    S[i] = G[r * nm + c];
    VK[i] = S[i];
  }
  __syncthreads();
  for (int s = d.P - 1; s >= 0; s--) {
    const size_t o = ((size_t)b * d.P + s) * nx * nx;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) d.segVb[o + i] = S[i];
    if (d.P == 1) break;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
      A[i] = d.segA[o + i];
      Cg[i] = d.segC[o + i];
    }
    __syncthreads();
    // M = [I + S C | S A]
    cta_mm(M, 2 * nx, nullptr, 0, 0.0, 1.0, S, nx, 1, Cg, nx, 1, nx, nx, nx);
    cta_mm(M + nx, 2 * nx, nullptr, 0, 0.0, 1.0, S, nx, 1, A, nx, 1, nx, nx, nx);
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) M[i * 2 * nx + i] += 1.0;
    cta_gauss_jordan(M, 2 * nx, nx, 2 * nx, piv_s, &st_s);
    // S <- J + A' X, symmetrised
    cta_mm(S, nx, d.segJ + o, nx, 1.0, 1.0, A, 1, nx, M + nx, 2 * nx, 1, nx, nx, nx);
    __syncthreads();
    cta_symmetrize(S, nx, nx);
    __syncthreads();
    if (s > 0) {  // V at the boundary a_s: the value segment s-1 builds on
      double *Va = d.V + ((size_t)b * (d.K + 1) + (size_t)s * d.L) * nx * nx;
      for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) Va[i] = S[i];
    }
  }
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, st_s);
}

// ---------------------------------------------------------------------------
// K3: Riccati recursion inside segment s from its terminal Vb.
// ---------------------------------------------------------------------------
__global__ void seg_riccati_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  SmemCarver sm(smem_raw);
  double *V = sm.take(nx * nx), *F = sm.take(nx * nm), *T = sm.take(nx * nm);
  double *G = sm.take(nm * nm), *Rux = sm.take(nu * nx), *Phi = sm.take(nx * nx);
  double *P0 = sm.take(nx * nx), *P1 = sm.take(nx * nx);
  __shared__ int st_s;
  if (threadIdx.x == 0) st_s = 0;
  const size_t so = ((size_t)b * d.P + s) * nx * nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    V[i] = d.segVb[so + i];
    P0[i] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  double *Psi = P0, *Psin = P1;
  for (int k = kb - 1; k >= ka; k--) {
    load_stage(d, b, k, G, F);
    riccati_stage(nx, nu, false, V, F, G, T, Rux, Phi, &st_s);
    const size_t ks = (size_t)b * d.K + k;
    // boundary values Vxx[a_s], s > 0, were fixed by K2
    if (k > ka || s == 0) {
      double *Vk = d.V + ((size_t)b * (d.K + 1) + k) * nx * nx;
      for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) Vk[i] = V[i];
    }
    double *Rk = d.Rux + ks * nu * nx, *Lk = d.LD + ks * nu * nu, *Pk = d.Phi + ks * nx * nx;
    for (int i = threadIdx.x; i < nu * nx; i += blockDim.x) Rk[i] = Rux[i];
    for (int i = threadIdx.x; i < nu * nu; i += blockDim.x) {
      const int r = i / nu, c = i - r * nu;
      Lk[i] = G[(nx + r) * nm + nx + c];
    }
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) Pk[i] = Phi[i];
    // Psi <- Psi Phi
    cta_mm(Psin, nx, nullptr, 0, 0.0, 1.0, Psi, nx, 1, Phi, nx, 1, nx, nx, nx);
    double *t = Psi; Psi = Psin; Psin = t;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) d.segPsi[so + i] = Psi[i];
  // an indefinite (but non-singular) Guu is accepted like the reference's BKP
  if (threadIdx.x == 0 && (st_s & LQ_FLAG_SING)) atomicOr(d.status, LQ_FLAG_SING);
}

// LDL^T of Vxx[0] for a free initial state (hqp/Hqp_IpLQDOCP.C:1971-1996 with
// an empty cbx[0]); one CTA per instance.
__global__ void x0_factor_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, b = blockIdx.x;
  double *A = reinterpret_cast<double *>(smem_raw);
  const double *V0 = d.V + (size_t)b * (d.K + 1) * nx * nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) A[i] = V0[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    const int st = warp_ldlt(A, nx, nx);  // indefinite is allowed, singular is not
    if ((st & LQ_FLAG_SING) && threadIdx.x == 0) atomicOr(d.status, LQ_FLAG_SING);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) d.V0f[(size_t)b * nx * nx + i] = A[i];
}
