// Factor kernels: parallel-in-time Riccati for the stage-structured KKT system.
//
// The reference sweeps the horizon sequentially (ExRiccatiFactorSc,
// hqp/Hqp_IpLQDOCP.C:1811-1969: for k = K-1..0, each stage needs Vxx[k+1]).
// Here the horizon of every instance is cut into P segments of L stages, and
// the segments are grouped R at a time into a small hierarchy (level 0 =
// segments, level l+1 = groups of R level-l elements):
//
//   hdiag_kernel        (stage-parallel) diagonal of C'(z/w)C from bound rows
//   K1 seg_element_kernel (P x batch CTAs)  each segment is condensed, with a
//        ZERO terminal cost, into its boundary element (A, C, J):
//            J = Riccati value Hessian at the segment start,
//            A = closed-loop transition across the segment,
//            C = closed-loop controllability Gramian weighted by Guu^{-1},
//        i.e. the segment's Schur complement onto (x_start, costate_end).
//   K2a elem_compose_kernel (per level, groups in parallel) composes R
//        consecutive elements into one:  with M = (I + C_i J_j)^{-1},
//            A = A_j M A_i,  C = A_j M C_i A_j' + C_j,  J = A_i' J_j M A_i + J_i
//   K2b elem_scan_kernel   top level (one CTA per instance) and, going back
//        down, every group in parallel: from the value Hessian S at an
//        element's end,  V at its start = J + A' (I + S C)^{-1} S A  (exact).
//   K3 seg_riccati_kernel (P x batch CTAs)  the ordinary Riccati recursion
//        inside every segment from its now-known terminal Vb, storing
//        Vxx[k], Rux[k], LDL'(Guu[k]), Phi[k] = fx - fu Rux and the segment
//        transition Psi_s = Phi[b-1] ... Phi[a].
//   K4 psi_compose_kernel (per level) Psi of a group = product of its children.
//
// Stage inputs (Q_k, fx_k, fu_k, hdiag_k) are staged into shared memory with
// TMA bulk copies one stage ahead of the arithmetic (double buffer + mbarrier).
// With P = 1 only the terminal block and K3 run: the reference's sequential
// sweep, one CTA per instance (the batched-MPC configuration).
#pragma once

#include "lq_device.cuh"

#ifdef LQ_TIMING
#define LQ_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) d.dbg[i] = clock64(); } while (0)
#else
#define LQ_STAMP(i) do { } while (0)
#endif

#ifndef LQ_K13_MINB
#define LQ_K13_MINB 3  // resident CTAs per SM the segment kernels are compiled for
#endif

#ifndef LQ_USE_DMMA
#define LQ_USE_DMMA 1  // FP64 tensor-core block products in the templated kernels
#endif

// ---------------------------------------------------------------------------
// hdiag[b][j] = sum over single-entry rows r of C in column j of (z_r/w_r) c_r^2
// (the bound rows of Hqp_Docp, hqp/Hqp_Docp.C:658-666).  grid-stride over
// batch*N variables.
// ---------------------------------------------------------------------------
__global__ void hdiag_kernel(LqDev d) {
  pdl_enter();
  const size_t total = (size_t)d.batch * d.N;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / d.N), j = (int)(t - (size_t)b * d.N);
    const double *z = d.z + (size_t)b * d.m, *w = d.w + (size_t)b * d.m;
    const double *cv = d.cval + (size_t)b * d.nnz;
    double s = 0.0;
    for (int e = d.vcol_ptr[j]; e < d.vcol_ptr[j + 1]; e++) {
      const int r = d.vcol_row[e];
      if (d.ineq_ptr[r + 1] - d.ineq_ptr[r] == 1) {
        const double a = cv[d.vcol_nz[e]];
        s = fma(z[r] / w[r], a * a, s);
      }
    }
    d.hdiag[t] = s;
  }
}

// ---------------------------------------------------------------------------
// Double-buffered stage loader.
// ---------------------------------------------------------------------------
struct StagePipe {
  double *base;   // two consecutive slots of [G | fx | fu | hd]
  uint64_t *bar;  // [2]
  int slot, o_fx, o_fu, o_hd;
  __device__ __forceinline__ double *G(int buf) const { return base + buf * slot; }
  __device__ __forceinline__ double *fx(int buf) const { return base + buf * slot + o_fx; }
  __device__ __forceinline__ double *fu(int buf) const { return base + buf * slot + o_fu; }
  __device__ __forceinline__ double *hd(int buf) const { return base + buf * slot + o_hd; }
};

// issued by ONE thread: start the bulk copies of stage k into buffer `buf`
__device__ __forceinline__ void stage_issue(const LqDev &d, int nx, int nu,
                                            const StagePipe &sp, int b, int k, int buf) {
  const int nm = nx + nu;
  const uint32_t bq = nm * nm * 8, bx = nx * nx * 8, bu = nx * nu * 8, bh = nm * 8;
  const double *Qk = d.Q + ((size_t)b * (d.K + 1) + k) * nm * nm;
  const double *hk = d.hdiag + (size_t)b * d.N + (size_t)k * nm;
  if (k < d.K) {
    mbar_expect_tx(&sp.bar[buf], bq + bx + bu + bh);
    tma_load_1d(sp.G(buf), Qk, bq, &sp.bar[buf]);
    tma_load_1d(sp.fx(buf), d.fx + ((size_t)b * d.K + k) * nx * nx, bx, &sp.bar[buf]);
    tma_load_1d(sp.fu(buf), d.fu + ((size_t)b * d.K + k) * nx * nu, bu, &sp.bar[buf]);
    tma_load_1d(sp.hd(buf), hk, bh, &sp.bar[buf]);
  } else {
    // terminal stage: only nx diagonal entries exist (nx*8 is a 16-byte multiple
    // whenever use_tma holds)
    mbar_expect_tx(&sp.bar[buf], bq + nx * 8);
    tma_load_1d(sp.G(buf), Qk, bq, &sp.bar[buf]);
    tma_load_1d(sp.hd(buf), hk, nx * 8, &sp.bar[buf]);
  }
}

// all threads: make stage k available in buffer `buf` as
//   G = Q_k + C_k' diag(z/w) C_k   (factor prologue hqp/Hqp_IpLQDOCP.C:805-832,
//   CTDC :68-103),  fx, fu.  `parity` = phase of the buffer's mbarrier.
// Ends with __syncthreads().
// fup != nullptr: fu is also re-packed there with row stride LU.
__device__ __forceinline__ void stage_acquire(const LqDev &d, int nx, int nu,
                                              const StagePipe &sp, int b, int k, int buf,
                                              uint32_t parity, double *fup = nullptr,
                                              int LU = 0) {
  const int nm = nx + nu;
  const int dk = k < d.K ? nm : nx;
  double *G = sp.G(buf);
  if (d.use_tma) {
    mbar_wait(&sp.bar[buf], parity);
  } else {
    const double *Qk = d.Q + ((size_t)b * (d.K + 1) + k) * nm * nm;
    if (d.gws) cta_copy_big(G, Qk, nm * nm, threadIdx.x, blockDim.x);
    else for (int i = threadIdx.x; i < nm * nm; i += blockDim.x) G[i] = Qk[i];
    if (k < d.K && !d.gws) {  // (large blocks: fx, fu are read where they lie, stage_fx / stage_fu)
      const double *fx = d.fx + ((size_t)b * d.K + k) * nx * nx;
      const double *fu = d.fu + ((size_t)b * d.K + k) * nx * nu;
      for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) sp.fx(buf)[i] = fx[i];
      for (int i = threadIdx.x; i < nx * nu; i += blockDim.x) sp.fu(buf)[i] = fu[i];
    }
    const double *hk = d.hdiag + (size_t)b * d.N + (size_t)k * nm;
    for (int i = threadIdx.x; i < dk; i += blockDim.x) sp.hd(buf)[i] = hk[i];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < dk; i += blockDim.x) G[i * nm + i] += sp.hd(buf)[i];
  if (fup && k < d.K) {
    const double *fu = sp.fu(buf);
    for (int i = threadIdx.x; i < nx * nu; i += blockDim.x) {
      const int r = i / nu, c = i - r * nu;
      fup[r * LU + c] = fu[i];
    }
  }
  // general rows (more than one nonzero): dense rank-1 updates, one row at a
  // time so that the summation order is fixed
  const int r0 = d.grow_ptr[k], r1 = d.grow_ptr[k + 1];
  if (r1 > r0) {
    const double *cv = d.cval + (size_t)b * d.nnz;
    const double *z = d.z + (size_t)b * d.m, *w = d.w + (size_t)b * d.m;
    for (int rr = r0; rr < r1; rr++) {
      __syncthreads();
      const int r = d.grow[rr];
      const int e0 = d.ineq_ptr[r], ne = d.ineq_ptr[r + 1] - e0;
      const double wz = z[r] / w[r];
      for (int e = threadIdx.x; e < ne * ne; e += blockDim.x) {
        const int ea = e / ne, eb = e - ea * ne;
        G[d.ineq_lcol[e0 + ea] * nm + d.ineq_lcol[e0 + eb]] += wz * cv[e0 + ea] * cv[e0 + eb];
      }
    }
  }
  __syncthreads();
}

// fx / fu of stage k as the stage kernels read them: the pipe's buffer, or (large
// blocks, whose products stage their operands themselves) the HBM slab directly
__device__ __forceinline__ const double *stage_fx(const LqDev &d, const StagePipe &sp, int nx, int b, int k, int buf) {
  return d.gws ? d.fx + ((size_t)b * d.K + k) * nx * nx : sp.fx(buf);
}
__device__ __forceinline__ const double *stage_fu(const LqDev &d, const StagePipe &sp, int nx, int nu, int b, int k,
                                                  int buf) {
  return d.gws ? d.fu + ((size_t)b * d.K + k) * nx * nu : sp.fu(buf);
}

// set up pipe storage and barriers; returns after a CTA barrier
// (bars: two mbarriers in SHARED memory, used by the bulk-copy path only)
__device__ __forceinline__ void stage_pipe_init(int nx, int nu, SmemCarver &sm, StagePipe &sp,
                                                uint64_t *bars, bool use_tma) {
  const int nm = nx + nu;
  sp.o_fx = (nm * nm + 1) & ~1;
  sp.o_fu = sp.o_fx + ((nx * nx + 1) & ~1);
  sp.o_hd = sp.o_fu + ((nx * nu + 1) & ~1);
  sp.slot = sp.o_hd + ((nm + 1) & ~1);
  sp.base = sm.take(2 * sp.slot);
  sp.bar = bars;
  if (use_tma && threadIdx.x == 0) {
    mbar_init(&sp.bar[0], 1);
    mbar_init(&sp.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// One Riccati stage on shared-memory blocks (FormGxx hqp/Hqp_IpLQDOCP.C:1077-1111
// + unconstrained-u branch :1854-1882 + Vxx :1940-1961).
//   in : V (nx x nx) = Vxx[k+1] (ignored when zero_V), fx, fu, G = H_k
//   out: G = [Gxx . ; Gux LDL'(Guu)] (lower blocks), Rux (nu x nx),
//        V <- Gxx - Gux' Rux (NOT yet symmetrised), Phi = fx - fu Rux
// T is scratch (nx x nm).  When `el` (K1): also W = A fu, Y = W Guu^{-1},
// Cg += Y W' are folded into the same barrier phases.
// Ends with __syncthreads().
// ---------------------------------------------------------------------------
// Row strides of the CTA-internal blocks.  A DMMA fragment load is a 64-bit
// shared load that the hardware serves per half-warp (lanes g = 0..3, t = 0..3),
// for both operand orientations -- element (row g, col t) at g*ld + t or at
// t*ld + g -- conflict-free exactly when ld = 4 (mod 8) doubles (measured: stride
// 24 doubled the wavefronts of these loads, 20 is ideal).  The tensor-core
// kernels therefore pad the internal row strides to the next value = 4 (mod 8):
// 20 -> 20, 30 -> 36, 10 -> 12, 12 -> 12, 16 -> 20, 40 -> 44, 50 -> 52, and the
// K1 accumulators are kept transposed so that no operand needs another layout.
// G and fx arrive by TMA in their dense global layout; fu (small) is re-packed.
__host__ __device__ constexpr int lq_pad4(int n) {
  return n <= 4 ? 4 : ((n - 4 + 7) / 8) * 8 + 4;
}

// K1 accumulators: At = A' (nx x nx), Wt = (A fu)' and Yt = (W Guu^{-1})'
// (nu x nx), Cg (nx x nx); all with row stride LV
struct ElemAcc {
  double *At, *Wt, *Yt, *Cg;
};

// V, Phi, Rux (nu x nx), the ElemAcc blocks: stride LV; T (nx x nm): stride LT
// fu: nx x nu with row stride LU
// `idle(w, nw)`: work for the nw warps that have nothing to do while warp 0
// factors Guu and solves for Rux (K3: the copy-outs and the Psi update of the
// previous stage); called with (0, 1) on a one-warp CTA.
struct NoIdle {
  __device__ __forceinline__ void operator()(int, int) const {}
};
template <int NU, bool TC, int NW, class Idle = NoIdle>
__device__ __forceinline__ void riccati_stage(double *stg, int nx, int nu, int LV, int LT, int LU,
                                              bool zero_V, double *V, const double *fx,
                                              const double *fu, double *G, double *T, double *Rux,
                                              double *Phi, int *st_s, const ElemAcc *el,
                                              Idle idle = Idle(), int *kind_s = nullptr) {
  const int nm = nx + nu;
  const int tid = threadIdx.x, nthr = blockDim.x;
  LQ_STAMP2(1);
  if (!zero_V) {
    // T = V [fx fu]   (V symmetric: read as V')
    cta_mmx<TC, NW>(stg, T, LT, nullptr, 0, 0.0, 1.0, V, 1, LV, fx, nx, 1, nx, nx, nx);
    cta_mmx<TC, NW>(stg, T + nx, LT, nullptr, 0, 0.0, 1.0, V, 1, LV, fu, LU, 1, nx, nu, nx, 3);
  }
  if (el)  // Wt = fu' At  (W = A fu)
    cta_mmx<TC, NW>(stg, el->Wt, LV, nullptr, 0, 0.0, 1.0, fu, 1, LU, el->At, LV, 1, nu, nx, nx, 2);
  if (!zero_V || el) __syncthreads();
  LQ_STAMP2(2);
  if (!zero_V) {
    // Gxx += fx' Tx ; Gux += fu' Tx ; Guu += fu' Tu  (lower blocks only)
    // (Gxx and V are symmetric: tiles on and below the diagonal only, mirrored later)
    cta_mmx<TC, NW>(stg, G, nm, G, nm, 1.0, 1.0, fx, 1, nx, T, LT, 1, nx, nx, nx, 0, -1, TC);
    if (!TC && stg) {  // (large blocks: [Gux Guu] += fu' [Tx Tu] as one product)
      cta_mmx<TC, NW>(stg, G + nx * nm, nm, G + nx * nm, nm, 1.0, 1.0, fu, 1, LU, T, LT, 1, nu, nm, nx);
    } else {
      cta_mmx<TC, NW>(stg, G + nx * nm, nm, G + nx * nm, nm, 1.0, 1.0, fu, 1, LU, T, LT, 1, nu, nx, nx);
      cta_mmx<TC, NW>(stg, G + nx * nm + nx, nm, G + nx * nm + nx, nm, 1.0, 1.0, fu, 1, LU, T + nx, LT, 1,
                  nu, nu, nx, 3);
    }
    __syncthreads();
  }
  double *Guu = G + nx * nm + nx;
  LQ_STAMP2(3);
  // Large blocks: G lives in global memory, where a dependent substitution step
  // costs an L2 round trip.  The factor of Guu is kept in the shared-memory area
  // `lds` behind the GEMM staging ring, and the ring itself holds the right-hand
  // sides of the substitutions, one column per thread, `yc` columns at a time
  // (NULL when it does not fit: the slow path).
  double *lds = nullptr, *ybuf = nullptr;
  int yc = 0;
  if constexpr (!TC) {
    if (stg && big_ldlt_fits(nx, nu, nthr)) {
      lds = stg + big_seg_union_doubles(nx, nu, nthr);
      ybuf = stg;
      yc = big_ycols(nx, nu, nthr);
    }
  }
  const int ldl = nu + 1;
  if (lds) {
    // the deferred work of the previous stage first (K3: copy-outs and the Psi product
    // on the tensor cores, by the whole CTA), then LDL^T of Guu in shared memory
    idle((int)(tid >> 5), NW);
    __shared__ int ldlt_flag;
    for (int e = tid; e < nu * nu; e += nthr) {
      const int i = e / nu, j = e - i * nu;
      if (j <= i) lds[i * ldl + j] = Guu[i * nm + j];
    }
    const int st = cta_ldlt(lds, ldl, nu, &ldlt_flag);  // (barriers on entry and exit)
    if (st && tid == 0) atomicOr(st_s, st);
    for (int e = tid; e < nu * nu; e += nthr) {
      const int i = e / nu, j = e - i * nu;
      if (j <= i) Guu[i * nm + j] = lds[i * ldl + j];
    }
    // Rux = Guu^{-1} Gux and (K1) Yt = Guu^{-1} Wt: one right-hand side per thread
    if (tid < yc) {
      double *y = ybuf + tid;
      const int ncols = el ? 2 * nx : nx;
      for (int j = tid; j < ncols; j += yc) {
        const double *src = j < nx ? G + nx * nm + j : el->Wt + (j - nx);
        double *dst = j < nx ? Rux + j : el->Yt + (j - nx);
        const int sld = j < nx ? nm : LV;
        for (int i = 0; i < nu; i++) y[i * yc] = src[(size_t)i * sld];
        thread_ldlt_solve(lds, ldl, nu, y, yc);
        for (int i = 0; i < nu; i++) dst[(size_t)i * LV] = y[i * yc];
      }
    }
  } else if (el) {
    // K1: LDL' by warp 0, then the 2 nx triangular solves spread over the CTA
    if (warp_id_uniform() == 0) {  // uniform branch: no WARPSYNC around the shuffles inside
      const int st = warp_ldlt_any<NU>(Guu, nm, nu);
      if (st && tid == 0) atomicOr(st_s, st);
    }
    __syncthreads();
    LQ_STAMP2(4);
    // Rux = Guu^{-1} Gux : one right-hand side (column of Gux) per thread;
    // columns of Yt = Guu^{-1} Wt on the next nx threads
    for (int j = tid; j < 2 * nx; j += nthr) {
      if (j < nx) {
        for (int i = 0; i < nu; i++) Rux[i * LV + j] = G[(nx + i) * nm + j];
        ldlt_solve_any<NU>(Guu, nm, nu, Rux + j, LV);
      } else {
        const int i = j - nx;
        for (int l = 0; l < nu; l++) el->Yt[l * LV + i] = el->Wt[l * LV + i];
        ldlt_solve_any<NU>(Guu, nm, nu, el->Yt + i, LV);
      }
    }
  } else {
    // K3: warp 0 factors and solves back to back (no CTA barrier in between);
    // the other warps run the deferred work of the previous stage meanwhile
    const int wu = warp_id_uniform();
    if (wu == 0) {
      // (a16) positive definite blocks: LDL^T without interchanges; a block with a
      // non-positive pivot is left untouched and inverted with scaling and pivoting
      constexpr bool PIV = NU > 0 && NU <= 32;
      int st = warp_ldlt_any<NU>(Guu, nm, nu, PIV);
      bool inv = false;
      if constexpr (PIV) {
        if (st) {  // (the same on every lane: pivots are broadcast by shuffles)
          constexpr int ME = (NU + 1) & ~1;
          st = warp_pivoted_inverse<NU>(Guu, nm, T,
                                        reinterpret_cast<int *>(T + ME * (ME + 1) + 2 * (ME + 2) + ME + 2));
          inv = true;
        }
      }
      if (st && tid == 0) atomicOr(st_s, st);
      if (kind_s && tid == 0) *kind_s = inv ? 1 : 0;
      __syncwarp();
      LQ_STAMP2(4);
      for (int j = tid; j < nx; j += 32) {
        for (int i = 0; i < nu; i++) Rux[i * LV + j] = G[(nx + i) * nm + j];
        if constexpr (PIV) {
          if (inv) thread_inv_apply<NU>(Guu, nm, Rux + j, LV);
          else ldlt_solve_any<NU>(Guu, nm, nu, Rux + j, LV);
        } else {
          ldlt_solve_any<NU>(Guu, nm, nu, Rux + j, LV);
        }
      }
      if (NW == 1) {
        __syncwarp();
        idle(0, 1);
      }
    } else {
      idle(wu - 1, NW - 1);
    }
  }
  __syncthreads();
  LQ_STAMP2(5);
  // V = Gxx - Gux' Rux ; Phi = fx - fu Rux ; K1: Cg += Y W' = Yt' Wt
  cta_mmx<TC, NW>(stg, V, LV, G, nm, 1.0, -1.0, G + nx * nm, 1, nm, Rux, LV, 1, nx, nx, nu, 0, -1, TC);
  cta_mmx<TC, NW>(stg, Phi, LV, fx, nx, 1.0, -1.0, fu, LU, 1, Rux, LV, 1, nx, nx, nu);
  if (el)
    cta_mmx<TC, NW>(stg, el->Cg, LV, el->Cg, LV, 1.0, 1.0, el->Yt, 1, LV, el->Wt, LV, 1, nx, nx, nu, 2,
                    -1, TC);
  __syncthreads();
}

// ---------------------------------------------------------------------------
// K1: condense segment s of instance b with zero terminal cost.
// ---------------------------------------------------------------------------
// NW warps per CTA: 4 = the CTA-cooperative version (every product split over the
// warps, barriers between the phases of a stage); 1 = one warp per segment (no
// barrier wait, all DMMAs of a stage issued by the same warp).
template <int NX, int NU, int NW>
__global__ void __launch_bounds__(32 * NW, (NX == 20 && NW == 4 ? LQ_K13_MINB : 0))
seg_element_kernel(LqDev d) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, nu = NX > 0 ? NU : d.nu, nm = nx + nu;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const int LV = TC ? lq_pad4(nx) : nx, LT = TC ? lq_pad4(nm) : nm, LU = TC ? lq_pad4(nu) : nu;
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  // large blocks: shared memory is the GEMM staging area (cta_mm_big)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  StagePipe sp;
  __shared__ __align__(8) uint64_t pipe_bars[2];
  stage_pipe_init(nx, nu, sm, sp, pipe_bars, d.use_tma);
  double *fup = TC ? sm.take(nx * LU) : nullptr;
  double *J = sm.take(nx * LV), *A0 = sm.take(nx * LV), *A1 = sm.take(nx * LV);
  double *Cg = sm.take(nx * LV), *T = sm.take(nx * LT);
  double *Rux = sm.take(nu * LV), *Phi = sm.take(nx * LV);
  double *Wt = sm.take(nu * LV), *Yt = sm.take(nu * LV);
  __shared__ int st_s;
  if (threadIdx.x == 0) {
    st_s = 0;
    if (d.use_tma) stage_issue(d, nx, nu, sp, b, kb - 1, 0);
  }
  for (int i = threadIdx.x; i < nx * LV; i += blockDim.x) {
    const int r = i / LV, c = i - r * LV;
    J[i] = 0.0;
    Cg[i] = 0.0;
    A0[i] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  double *At = A0, *Atn = A1;  // At = A' (transposed accumulator)
  int it = 0;
  for (int k = kb - 1; k >= ka; k--, it++) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && d.use_tma && k > ka) {
      fence_proxy_async();
      stage_issue(d, nx, nu, sp, b, k - 1, buf ^ 1);
    }
    stage_acquire(d, nx, nu, sp, b, k, buf, (it >> 1) & 1, fup, LU);
    ElemAcc el{At, Wt, Yt, Cg};
    riccati_stage<NU, TC, NW>(stg, nx, nu, LV, LT, LU, k == kb - 1, J, stage_fx(d, sp, nx, b, k, buf),
                          TC ? fup : stage_fu(d, sp, nx, nu, b, k, buf), sp.G(buf), T, Rux, Phi, &st_s, &el);
    // J symmetrised ; A <- A Phi, i.e. At <- Phi' At
    if constexpr (TC) cta_symmetrize_tc<NW>(J, LV, nx, true);
    else if (stg) cta_symmetrize_big(stg, J, LV, nx);
    else cta_symmetrize(J, LV, nx);
    cta_mmx<TC, NW>(stg, Atn, LV, nullptr, 0, 0.0, 1.0, Phi, 1, LV, At, LV, 1, nx, nx, nx);
    double *t = At; At = Atn; Atn = t;
    __syncthreads();
  }
  if constexpr (TC) cta_symmetrize_tc<NW>(Cg, LV, nx, true);
  else if (stg) cta_symmetrize_big(stg, Cg, LV, nx);
  else cta_symmetrize(Cg, LV, nx);
  __syncthreads();
  const size_t o = ((size_t)b * d.ft.nel + s) * nx * nx;
  for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    d.segA[o + i] = At[c * LV + r];
    d.segC[o + i] = Cg[r * LV + c];
    d.segJ[o + i] = J[r * LV + c];
  }
  // a non-positive pivot here comes from the artificial zero terminal cost
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, LQ_FLAG_NOTPD);
}

// ---------------------------------------------------------------------------
// K2a: compose the children [g*R, min((g+1)*R, cnt)) of level `lev` into
// element g of level lev+1.  grid (cnt_{lev+1}, batch).
// ---------------------------------------------------------------------------

template <int NX>
__global__ void __launch_bounds__(NX == 0 ? LQ_BIG_NT : LQ_NT2) elem_compose_kernel(LqDev d, int lev) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LQ_STAMP(0);
  const int nx = NX > 0 ? NX : d.nx, n2 = nx * nx, n3 = 3 * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int ldm = NX > 0 ? n3 + 1 : n3;  // odd row stride for the warp inverse
  const int g = blockIdx.x, b = blockIdx.y;
  const int c0 = g * d.ft.R, c1 = min(d.ft.cnt[lev], c0 + d.ft.R);
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  // large blocks: shared memory is the GEMM staging area (cta_mm_big)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  const int ldx = NX > 0 ? 2 * nx + 4 : 2 * nx;  // = 4 (mod 8): conflict-free fragment reads
  double *Aj = sm.take(n2), *Cj = sm.take(n2), *Jj = sm.take(n2);
  double *Ai = sm.take(n2), *Ji = sm.take(n2);
  double *M = sm.take(nx * ldm), *T1 = sm.take(n2), *T2 = sm.take(n2), *T3 = sm.take(n2);
  double *X = sm.take(nx * ldx);  // [X_A | X_C]
  double *inv_scr = NX > 0 ? sm.take(nx * (nx + 1) + 2 * (nx + 2)) : nullptr;
  __shared__ int st_s, piv_small[65];
  __shared__ double inv_s[2];
  int *piv_s = nx <= 64 ? piv_small : reinterpret_cast<int *>(sm.take((nx + 2) / 2));
  constexpr int NWC = LQ_NT2 / 32;
  if (threadIdx.x == 0) st_s = 0;
  __syncthreads();  // (before any load: costs nothing on the critical path)
  const size_t base = ((size_t)b * d.ft.nel + d.ft.off[lev]) * n2;
  {
    // (no barrier here: these loads and the first child's are in flight together)
    const size_t o = base + (size_t)(c1 - 1) * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      Aj[i] = d.segA[o + i];
      Cj[i] = d.segC[o + i];
      Jj[i] = d.segJ[o + i];
    }
  }
  LQ_STAMP(1);
  for (int c = c1 - 2; c >= c0; c--) {
    const size_t o = base + (size_t)c * n2;
    // M = [I + C_i J_j | A_i | C_i]
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const int r = i / nx, cc = i - r * nx;
      const double ai = d.segA[o + i], ci = d.segC[o + i];
      Ai[i] = ai;
      Ji[i] = d.segJ[o + i];
      M[r * ldm + nx + cc] = ai;
      M[r * ldm + 2 * nx + cc] = ci;
      T1[i] = ci;
    }
    __syncthreads();
    LQ_STAMP(2);
    cta_mmx<TC, NWC>(stg, M, ldm, nullptr, 0, 0.0, 1.0, T1, nx, 1, Jj, nx, 1, nx, nx, nx);
    __syncthreads();
    LQ_STAMP(3);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) M[i * ldm + i] += 1.0;
    if constexpr (NX > 0)
      cta_inverse_apply<NX, NWC>(M, ldm, n3, X, inv_scr, piv_s, &st_s, ldx);
    else if (stg)
      cta_inverse_apply_big(stg, big_gj_scratch(d), M, ldm, nx, n3, X, ldx, piv_s, &st_s);
    else
      cta_gauss_jordan<NX>(M, n3, nx, n3, X, piv_s, inv_s, &st_s);
    LQ_STAMP(4);
    // T1 = A_j X_C ; T2 = J_j X_A ; T3 = A_j X_A (the new A)
    cta_mmx<TC, NWC>(stg, T1, nx, nullptr, 0, 0.0, 1.0, Aj, nx, 1, X + nx, ldx, 1, nx, nx, nx);
    cta_mmx<TC, NWC>(stg, T2, nx, nullptr, 0, 0.0, 1.0, Jj, nx, 1, X, ldx, 1, nx, nx, nx, 4);
    cta_mmx<TC, NWC>(stg, T3, nx, nullptr, 0, 0.0, 1.0, Aj, nx, 1, X, ldx, 1, nx, nx, nx, 2);
    __syncthreads();
    // C = T1 A_j' + C_j ; J = A_i' T2 + J_i
    cta_mmx<TC, NWC>(stg, Cj, nx, Cj, nx, 1.0, 1.0, T1, nx, 1, Aj, 1, nx, nx, nx, nx);
    cta_mmx<TC, NWC>(stg, Jj, nx, Ji, nx, 1.0, 1.0, Ai, 1, nx, T2, nx, 1, nx, nx, nx, 4);
    __syncthreads();
    if constexpr (TC) {
      cta_symmetrize_tc<NWC>(Cj, nx, nx);
      cta_symmetrize_tc<NWC>(Jj, nx, nx);
    } else if (stg) {
      cta_symmetrize_big(stg, Cj, nx, nx);
      cta_symmetrize_big(stg, Jj, nx, nx);
    } else {
      cta_symmetrize(Cj, nx, nx);
      cta_symmetrize(Jj, nx, nx);
    }
    { double *t = Aj; Aj = T3; T3 = t; }
    __syncthreads();
    LQ_STAMP(5);
  }
  const size_t o = ((size_t)b * d.ft.nel + d.ft.off[lev + 1] + g) * n2;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    d.segA[o + i] = Aj[i];
    d.segC[o + i] = Cj[i];
    d.segJ[o + i] = Jj[i];
  }
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, st_s);
  LQ_STAMP(6);
}

// The virtual terminal element of the suffix scan, in slot P of both regions:
// J = Vxx[K] = Q_K + C_K'(z/w)C_K (hqp/Hqp_IpLQDOCP.C:1800-1804), A = C = 0; also the
// end value of the last segment for K3.  grid (batch), any block size.
__global__ void elem_terminal_kernel(LqDev d) {
  pdl_enter();
  const int nx = d.nx, nm = d.nm, n2 = nx * nx, b = blockIdx.x;
  const size_t eb = (size_t)b * d.ft.nel;
  double *J0 = d.segJ + (eb + d.P) * n2, *J1 = d.segJ + (eb + 2 * d.P + 1) * n2;
  const double *QK = d.Q + ((size_t)b * (d.K + 1) + d.K) * nm * nm;
  const double *hk = d.hdiag + (size_t)b * d.N + (size_t)d.K * nm;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    // (horizon split: plus the value of everything behind this range)
    J0[i] = QK[r * nm + c] + (r == c ? hk[r] : 0.0) + (d.has_next ? d.Vext[i] : 0.0);
    d.segA[(eb + d.P) * n2 + i] = 0.0;
    d.segC[(eb + d.P) * n2 + i] = 0.0;
    d.segA[(eb + 2 * d.P + 1) * n2 + i] = 0.0;
    d.segC[(eb + 2 * d.P + 1) * n2 + i] = 0.0;
  }
  const double *cv = d.cval + (size_t)b * d.nnz;
  for (int rr = d.grow_ptr[d.K]; rr < d.grow_ptr[d.K + 1]; rr++) {
    __syncthreads();
    const int r = d.grow[rr];
    const int e0 = d.ineq_ptr[r], ne = d.ineq_ptr[r + 1] - e0;
    const double wz = d.z[(size_t)b * d.m + r] / d.w[(size_t)b * d.m + r];
    for (int e = threadIdx.x; e < ne * ne; e += blockDim.x) {
      const int ea = e / ne, ebb = e - ea * ne;
      J0[d.ineq_lcol[e0 + ea] * nx + d.ineq_lcol[e0 + ebb]] += wz * cv[e0 + ea] * cv[e0 + ebb];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const double v = J0[i];
    J1[i] = v;
    d.segVb[(eb + d.P - 1) * n2 + i] = v;
  }
}

// ---------------------------------------------------------------------------
// K2 as ONE sweep (suffix scan over the segment elements).  The terminal value
// Vxx[K] is a virtual element P = (A 0, C 0, J Vxx[K]) (elem_terminal_kernel);
// level l replaces element s by the composition of s and s + 2^l (beyond the end:
// unchanged).  Composing with a "complete" element (A = C = 0) is exactly the
// Riccati map of element s applied to that value, so after ceil(log2 (P+1)) levels
// the J of element s is the value Hessian at the START of segment s, i.e. the end
// value of segment s-1 -- every segment boundary with log2 P dependent combines
// instead of the 2 log2 P of the up- and down-sweep of a binary tree.  All P
// combines of a level run concurrently (one wave); the extra arithmetic is free,
// the machine idles during the tree anyway.
// Elements ping-pong between two regions of P+1 slots (src, dst: slot offsets).
// last != 0: also segVb[s-1] <- J.     grid (P, batch)
// ---------------------------------------------------------------------------
// jmax: last slot that takes part (P with the terminal element in the scan, P-1
// without); jfix >= 0: every element is combined with slot jfix instead of
// s + stride (a stage range of a split horizon learns its terminal value only
// after the exchange: the scan runs without it, then ONE more level applies the
// terminal element in slot P to all suffixes at once).
template <int NX>
__global__ void __launch_bounds__(NX == 0 ? LQ_BIG_NT : LQ_NT2) elem_hs_kernel(LqDev d, int stride, int src, int dst,
                                                      int last, int jmax, int jfix) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, n2 = nx * nx, n3 = 3 * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int ldm = NX > 0 ? n3 + 1 : n3;  // odd row stride for the warp inverse
  const int s = blockIdx.x, b = blockIdx.y;
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  const int ldx = NX > 0 ? 2 * nx + 4 : 2 * nx;
  double *Aj = sm.take(n2), *Cj = sm.take(n2), *Jj = sm.take(n2);
  double *Ai = sm.take(n2), *Ji = sm.take(n2);
  double *M = sm.take(nx * ldm), *T1 = sm.take(n2), *T2 = sm.take(n2), *T3 = sm.take(n2);
  double *X = sm.take(nx * ldx);  // [X_A | X_C]
  double *inv_scr = NX > 0 ? sm.take(nx * (nx + 1) + 2 * (nx + 2)) : nullptr;
  __shared__ int st_s, piv_small[65];
  __shared__ double inv_s[2];
  int *piv_s = nx <= 64 ? piv_small : reinterpret_cast<int *>(sm.take((nx + 2) / 2));
  constexpr int NWC = LQ_NT2 / 32;
  if (threadIdx.x == 0) st_s = 0;
  const size_t eb = (size_t)b * d.ft.nel;
  const size_t oi = (eb + src + s) * n2, oo = (eb + dst + s) * n2;
  const int j = jfix >= 0 ? jfix : s + stride;
  if (j > jmax) {
    // already the composition up to the end of the horizon
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const double jv = d.segJ[oi + i];
      d.segA[oo + i] = d.segA[oi + i];
      d.segC[oo + i] = d.segC[oi + i];
      d.segJ[oo + i] = jv;
      if (last && s > 0) d.segVb[(eb + s - 1) * n2 + i] = jv;
    }
    return;
  }
  const size_t oj = (eb + src + j) * n2;
  // M = [I + C_i J_j | A_i | C_i]
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int r = i / nx, cc = i - r * nx;
    Aj[i] = d.segA[oj + i];
    Cj[i] = d.segC[oj + i];
    Jj[i] = d.segJ[oj + i];
    const double ai = d.segA[oi + i], ci = d.segC[oi + i];
    Ai[i] = ai;
    Ji[i] = d.segJ[oi + i];
    M[r * ldm + nx + cc] = ai;
    M[r * ldm + 2 * nx + cc] = ci;
    T1[i] = ci;
  }
  __syncthreads();
  cta_mmx<TC, NWC>(stg, M, ldm, nullptr, 0, 0.0, 1.0, T1, nx, 1, Jj, nx, 1, nx, nx, nx);
  __syncthreads();
  for (int i = threadIdx.x; i < nx; i += blockDim.x) M[i * ldm + i] += 1.0;
  if constexpr (NX > 0)
    cta_inverse_apply<NX, NWC>(M, ldm, n3, X, inv_scr, piv_s, &st_s, ldx);
  else if (stg)
    cta_inverse_apply_big(stg, big_gj_scratch(d), M, ldm, nx, n3, X, ldx, piv_s, &st_s);
  else
    cta_gauss_jordan<NX>(M, n3, nx, n3, X, piv_s, inv_s, &st_s);
  // T1 = A_j X_C ; T2 = J_j X_A ; T3 = A_j X_A (the new A)
  cta_mmx<TC, NWC>(stg, T1, nx, nullptr, 0, 0.0, 1.0, Aj, nx, 1, X + nx, ldx, 1, nx, nx, nx);
  cta_mmx<TC, NWC>(stg, T2, nx, nullptr, 0, 0.0, 1.0, Jj, nx, 1, X, ldx, 1, nx, nx, nx, 4);
  cta_mmx<TC, NWC>(stg, T3, nx, nullptr, 0, 0.0, 1.0, Aj, nx, 1, X, ldx, 1, nx, nx, nx, 2);
  __syncthreads();
  // C = T1 A_j' + C_j ; J = A_i' T2 + J_i
  cta_mmx<TC, NWC>(stg, Cj, nx, Cj, nx, 1.0, 1.0, T1, nx, 1, Aj, 1, nx, nx, nx, nx);
  cta_mmx<TC, NWC>(stg, Jj, nx, Ji, nx, 1.0, 1.0, Ai, 1, nx, T2, nx, 1, nx, nx, nx, 4);
  __syncthreads();
  if constexpr (TC) {
    cta_symmetrize_tc<NWC>(Cj, nx, nx);
    cta_symmetrize_tc<NWC>(Jj, nx, nx);
  } else if (stg) {
    cta_symmetrize_big(stg, Cj, nx, nx);
    cta_symmetrize_big(stg, Jj, nx, nx);
  } else {
    cta_symmetrize(Cj, nx, nx);
    cta_symmetrize(Jj, nx, nx);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const double jv = Jj[i];
    d.segA[oo + i] = T3[i];
    d.segC[oo + i] = Cj[i];
    d.segJ[oo + i] = jv;
    if (last && s > 0) d.segVb[(eb + s - 1) * n2 + i] = jv;
  }
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, st_s);
}

// ---------------------------------------------------------------------------
// K2b: back-substitution of the value Hessian through the elements.
//   top = 1: level `lev` is the top level; one CTA per instance starts from the
//            terminal stage Vxx[K] = H_K (hqp/Hqp_IpLQDOCP.C:1800-1804) and walks
//            all elements of the level.          grid (1, batch)
//   top = 0: CTA g walks the children of element g of level lev+1 starting from
//            that element's segVb.               grid (cnt_{lev+1}, batch)
// For every element visited: segVb[e] = S (Vxx at its end), then
//   S <- J + A' (I + S C)^{-1} S A.
// ---------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(NX == 0 ? LQ_BIG_NT : LQ_NT2) elem_scan_kernel(LqDev d, int lev, int top) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, nm = d.nm, n2 = nx * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int g = blockIdx.x, b = blockIdx.y;
  const int ldm = NX > 0 ? 2 * nx + 1 : 2 * nx;  // odd row stride for the warp inverse
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  // large blocks: shared memory is the GEMM staging area (cta_mm_big)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  double *S = sm.take(n2), *A = sm.take(n2), *Cg = sm.take(n2);
  double *M = sm.take(nx * ldm);
  double *X = sm.take(n2);
  double *inv_scr = NX > 0 ? sm.take(nx * (nx + 1) + 2 * (nx + 2)) : nullptr;
  __shared__ int st_s, piv_small[65];
  __shared__ double inv_s[2];
  int *piv_s = nx <= 64 ? piv_small : reinterpret_cast<int *>(sm.take((nx + 2) / 2));
  if (threadIdx.x == 0) st_s = 0;
  int c0, c1;
  if (top) {
    c0 = 0;
    c1 = d.ft.cnt[lev];
    // terminal stage: Q_K + hdiag_K + general rows of stage K
    const double *QK = d.Q + ((size_t)b * (d.K + 1) + d.K) * nm * nm;
    const double *hk = d.hdiag + (size_t)b * d.N + (size_t)d.K * nm;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const int r = i / nx, c = i - r * nx;
      S[i] = QK[r * nm + c] + (r == c ? hk[r] : 0.0);
    }
    const double *cv = d.cval + (size_t)b * d.nnz;
    for (int rr = d.grow_ptr[d.K]; rr < d.grow_ptr[d.K + 1]; rr++) {
      __syncthreads();
      const int r = d.grow[rr];
      const int e0 = d.ineq_ptr[r], ne = d.ineq_ptr[r + 1] - e0;
      const double wz = d.z[(size_t)b * d.m + r] / d.w[(size_t)b * d.m + r];
      for (int e = threadIdx.x; e < ne * ne; e += blockDim.x) {
        const int ea = e / ne, eb = e - ea * ne;
        S[d.ineq_lcol[e0 + ea] * nx + d.ineq_lcol[e0 + eb]] += wz * cv[e0 + ea] * cv[e0 + eb];
      }
    }
    __syncthreads();
    if (d.has_next)  // horizon split: value of everything behind this range
      for (int i = threadIdx.x; i < n2; i += blockDim.x) S[i] += d.Vext[i];
    __syncthreads();
    double *VK = d.V + ((size_t)b * (d.K + 1) + d.K) * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) VK[i] = S[i];
  } else {
    c0 = g * d.ft.R;
    c1 = min(d.ft.cnt[lev], c0 + d.ft.R);
    const size_t o = ((size_t)b * d.ft.nel + d.ft.off[lev + 1] + g) * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) S[i] = d.segVb[o + i];
  }
  __syncthreads();
  const size_t base = ((size_t)b * d.ft.nel + d.ft.off[lev]) * n2;
  for (int c = c1 - 1; c >= c0; c--) {
    const size_t o = base + (size_t)c * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) d.segVb[o + i] = S[i];
    // the value at the start of the first child is the end value of the
    // previous group (known to the parent level) or, at the very top, Vxx[0],
    // which K3 produces itself: skip the last update
    if (c == c0) break;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      A[i] = d.segA[o + i];
      Cg[i] = d.segC[o + i];
    }
    __syncthreads();
    // M = [I + S C | S A]
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
    // S <- J + A' X, symmetrised
    cta_mmx<TC, LQ_NT2 / 32>(stg, S, nx, d.segJ + o, nx, 1.0, 1.0, A, 1, nx, X, nx, 1, nx, nx, nx);
    __syncthreads();
    if constexpr (TC) cta_symmetrize_tc<LQ_NT2 / 32>(S, nx, nx);
    else if (stg) cta_symmetrize_big(stg, S, nx, nx);
    else cta_symmetrize(S, nx, nx);
    __syncthreads();
  }
  if (threadIdx.x == 0 && st_s) atomicOr(d.status, st_s);
}

// ---------------------------------------------------------------------------
// K3: Riccati recursion inside segment s from its terminal Vb.
// ---------------------------------------------------------------------------
template <int NX, int NU, int NW>
__global__ void __launch_bounds__(32 * NW, (NX == 20 && NW == 4 ? LQ_K13_MINB : 0))
seg_riccati_kernel(LqDev d) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, nu = NX > 0 ? NU : d.nu, nm = nx + nu, n2 = nx * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const int LV = TC ? lq_pad4(nx) : nx, LT = TC ? lq_pad4(nm) : nm, LU = TC ? lq_pad4(nu) : nu;
  SmemCarver sm(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // (compiled sizes: provably shared memory)
  // large blocks: shared memory is the GEMM staging area (cta_mm_big)
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  StagePipe sp;
  __shared__ __align__(8) uint64_t pipe_bars[2];
  stage_pipe_init(nx, nu, sm, sp, pipe_bars, d.use_tma);
  double *fup = TC ? sm.take(nx * LU) : nullptr;
  double *V = sm.take(nx * LV), *T = sm.take(nx * LT);
  double *RuxA = sm.take(nu * LV), *RuxB = sm.take(nu * LV), *Phi = sm.take(nx * LV);
  double *P0 = sm.take(nx * LV), *P1 = sm.take(nx * LV);
  __shared__ int st_s, kind_s;
  if (threadIdx.x == 0) {
    st_s = 0;
    kind_s = 0;
    if (d.use_tma) stage_issue(d, nx, nu, sp, b, kb - 1, 0);
  }
  const size_t so = ((size_t)b * d.ft.nel + s) * n2;   // factor tree (segVb)
  const size_t po = ((size_t)b * d.st.nel + s) * n2;   // solve tree (segPsi)
  double *Vend = d.V + ((size_t)b * (d.K + 1) + kb) * n2;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int r = i / nx, c = i - r * nx;
    const double v = d.segVb[so + i];
    V[r * LV + c] = v;
    Vend[i] = v;  // Vxx at the segment end (for the last segment: Vxx[K])
    P0[r * LV + c] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  double *Pt = P0, *Ptn = P1;  // Pt = Psi' (transposed accumulator)
  double *Rux = RuxA, *Rprev = RuxB;
  const bool need_psi = d.P > 1 || d.has_prev || d.has_next;
  // Results of stage kp (Rux in `Rprev`, Phi, symmetrised V) leave for HBM and
  // enter Psi while stage kp-1 factors Guu: written by `nw` warps, `w` = index
  // of the calling warp among them.  V and Phi are not overwritten before the
  // barrier that ends that phase, Rux is double-buffered.
  auto flush = [&](int kp, const double *Rp, const double *Ptc, double *Ptd, int w, int nw) {
    const size_t ks = (size_t)b * d.K + kp;
    double *Rk = d.Rux + ks * nu * nx, *Pk = d.Phi + ks * n2;
    const int t0 = w * 32 + (int)(threadIdx.x & 31), nt = nw * 32;
    // Psi <- Psi Phi, i.e. Pt <- Phi' Pt (only the solve hierarchy reads Psi: not
    // needed when the segment is the whole horizon)
    if (need_psi) {
      if (nw == NW)
        cta_mmx<TC, NW>(stg, Ptd, LV, nullptr, 0, 0.0, 1.0, Phi, 1, LV, Ptc, LV, 1, nx, nx, nx);
      else
        cta_mmx<TC, (NW > 1 ? NW - 1 : 1)>(stg, Ptd, LV, nullptr, 0, 0.0, 1.0, Phi, 1, LV, Ptc, LV, 1,
                                          nx, nx, nx, 0, w);
    }
    if (stg) {  // (large blocks: LV == nx, plain copies out of the workspace)
      cta_copy_big(Rk, Rp, nu * nx, t0, nt);
      cta_copy_big(Pk, Phi, n2, t0, nt);
      if (kp > ka || s == 0) cta_copy_big(d.V + ((size_t)b * (d.K + 1) + kp) * n2, V, n2, t0, nt);
      return;
    }
    for (int i = t0; i < nu * nx; i += nt) {
      const int r = i / nx, c = i - r * nx;
      Rk[i] = Rp[r * LV + c];
    }
    for (int i = t0; i < n2; i += nt) {
      const int r = i / nx, c = i - r * nx;
      Pk[i] = Phi[r * LV + c];
    }
    // interior value Hessians; Vxx[a_s], s > 0, is the end value of segment
    // s-1 and is written there
    if (kp > ka || s == 0) {
      double *Vk = d.V + ((size_t)b * (d.K + 1) + kp) * n2;
      for (int i = t0; i < n2; i += nt) {
        const int r = i / nx, c = i - r * nx;
        Vk[i] = V[r * LV + c];
      }
    }
  };
  int it = 0;
  for (int k = kb - 1; k >= ka; k--, it++) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && d.use_tma && k > ka) {
      fence_proxy_async();
      stage_issue(d, nx, nu, sp, b, k - 1, buf ^ 1);
    }
    LQ_STAMP2(0);
    stage_acquire(d, nx, nu, sp, b, k, buf, (it >> 1) & 1, fup, LU);
    double *G = sp.G(buf);
    const bool have_prev = it > 0;
    const double *Rp = Rprev, *Ptc = Pt;
    double *Ptd = Ptn;
    riccati_stage<NU, TC, NW>(stg, nx, nu, LV, LT, LU, false, V, stage_fx(d, sp, nx, b, k, buf),
                              TC ? fup : stage_fu(d, sp, nx, nu, b, k, buf), G, T,
                              Rux, Phi, &st_s, nullptr, [&](int w, int nw) {
                                if (have_prev) flush(k + 1, Rp, Ptc, Ptd, w, nw);
                              }, &kind_s);
    if (threadIdx.x == 0) d.ldkind[(size_t)b * d.K + k] = kind_s;
    if (have_prev) { double *t = Pt; Pt = Ptn; Ptn = t; }
    LQ_STAMP2(6);
    const size_t ks = (size_t)b * d.K + k;
    double *Lk = d.LD + ks * nu * nu;
    if constexpr (TC) cta_symmetrize_tc<NW>(V, LV, nx, true);
    else if (stg) cta_symmetrize_big(stg, V, LV, nx);
    else cta_symmetrize(V, LV, nx);
    for (int i = threadIdx.x; i < nu * nu; i += blockDim.x) {
      const int r = i / nu, c = i - r * nu;
      Lk[i] = G[(nx + r) * nm + nx + c];
    }
    { double *t = Rux; Rux = Rprev; Rprev = t; }
    __syncthreads();
    LQ_STAMP2(7);
  }
  // the last stage's results, by the whole CTA
  flush(ka, Rprev, Pt, Ptn, (int)(threadIdx.x >> 5), NW);
  __syncthreads();
  if (need_psi)
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const int r = i / nx, c = i - r * nx;
      d.segPsi[po + i] = Ptn[c * LV + r];
    }
  // an indefinite (but non-singular) Guu is accepted like the reference's BKP
  if (threadIdx.x == 0 && (st_s & LQ_FLAG_SING)) atomicOr(d.status, LQ_FLAG_SING);
}

// K4: Psi of element g of level lev+1 = Psi[c1-1] ... Psi[c0] of its children.
// The children are loaded `chunk` at a time and multiplied as a binary tree
// inside the CTA (log2 rounds of independent products on the tensor cores)
// instead of R-1 dependent products; the product so far re-enters the next
// chunk as its first (rightmost) factor.  grid (cnt_{lev+1}, batch), LQ_NT2
// threads, smem: (chunk + ceil(chunk/2)) * nx*nx doubles.
template <int NX>
__global__ void __launch_bounds__(NX == 0 ? LQ_BIG_NT : LQ_NT2) psi_compose_kernel(LqDev d, int lev, int chunk) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx, n2 = nx * nx;
  constexpr bool TC = LQ_USE_DMMA && NX > 0;
  const int g = blockIdx.x, b = blockIdx.y;
  const int c0 = g * d.st.R, c1 = min(d.st.cnt[lev], c0 + d.st.R);
  double *bufA = reinterpret_cast<double *>(NX > 0 ? smem_raw : cta_workspace(d, smem_raw));  // chunk blocks
  double *const stg = (NX == 0 && d.gws) ? reinterpret_cast<double *>(smem_raw) : nullptr;
  double *bufB = bufA + (size_t)chunk * n2;             // ceil(chunk/2) blocks
  const size_t base = ((size_t)b * d.st.nel + d.st.off[lev]) * n2;
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int TJ = (nx + 7) >> 3, ntiles = TJ * TJ;
  double *src = bufA;
  int have = 0;
  __shared__ uint64_t tma_bar;
  uint32_t tma_phase = 0;
  if (threadIdx.x == 0) {
    mbar_init(&tma_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  for (int done = c0; done < c1;) {
    const int take = min(chunk - have, c1 - done);
    if (d.use_tma) {
      // the children are contiguous: one bulk copy (a per-thread copy loop pays
      // the memory latency once per unrolled group of loads)
      if (threadIdx.x == 0) {
        if (done > c0) fence_proxy_async();  // bufA was read through the generic proxy
        const uint32_t bytes = (uint32_t)take * n2 * 8;
        mbar_expect_tx(&tma_bar, bytes);
        tma_load_1d(bufA + (size_t)have * n2, d.segPsi + base + (size_t)done * n2, bytes, &tma_bar);
      }
      mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1;
    } else {
      for (int i = threadIdx.x; i < take * n2; i += blockDim.x)
        bufA[(size_t)have * n2 + i] = d.segPsi[base + (size_t)done * n2 + i];
    }
    done += take;
    int cnt = have + take;
    __syncthreads();
    double *dst = bufB;
    src = bufA;
    while (cnt > 1) {
      const int half = cnt >> 1, odd = cnt & 1;
      // pair (2i+1, 2i) -> src[2i+1] * src[2i]; an unpaired last block is copied
      if constexpr (TC) {
        const int wpp = max(1, nwarp / half);  // warps per product
        const int stride = max(1, nwarp / wpp);
        for (int pr = warp / wpp; pr < half; pr += stride) {
          const int wq = warp % wpp;
          const double *A = src + (size_t)(2 * pr + 1) * n2, *B = src + (size_t)(2 * pr) * n2;
          double *C = dst + (size_t)pr * n2;
          for (int ta = wq; ta < ntiles; ta += 2 * wpp) {
            const int tb = ta + wpp;
            mm_tc_tile_pair(C, nx, nullptr, 0, 0.0, 1.0, A, nx, 1, B, nx, 1, nx, nx, nx, TJ, ta,
                            tb < ntiles ? tb : -1, gq, t);
          }
        }
      } else {
        for (int pr = 0; pr < half; pr++)
          cta_mmx<false, LQ_NT2 / 32>(stg, dst + (size_t)pr * n2, nx, nullptr, 0, 0.0, 1.0,
                                      src + (size_t)(2 * pr + 1) * n2, nx, 1,
                                      src + (size_t)(2 * pr) * n2, nx, 1, nx, nx, nx);
      }
      if (odd)
        for (int i = threadIdx.x; i < n2; i += blockDim.x)
          dst[(size_t)half * n2 + i] = src[(size_t)(cnt - 1) * n2 + i];
      __syncthreads();
      cnt = half + odd;
      double *tmp = src; src = dst; dst = tmp;
    }
    if (done < c1 && src != bufA) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) bufA[i] = src[i];
      // the loads of the next chunk touch other blocks; the barrier after them orders this copy
    }
    have = 1;
  }
  const size_t o = ((size_t)b * d.st.nel + d.st.off[lev + 1] + g) * n2;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) d.segPsi[o + i] = src[i];
}

// LDL^T of Vxx[0] for a free initial state (hqp/Hqp_IpLQDOCP.C:1971-1996 with
// an empty cbx[0]); one CTA per instance.
__global__ void x0_factor_kernel(LqDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, b = blockIdx.x;
  double *A = reinterpret_cast<double *>(cta_workspace(d, smem_raw));
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
