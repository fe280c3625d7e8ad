// Solve kernels: one KKT solve with the current factor
// (Hqp_IpLQDOCP::step hqp/Hqp_IpLQDOCP.C:869-976 + ExRiccatiSolveSc :2007-2182).
//
// The reference's two dependent sweeps carry, per stage, a BKP solve and four
// mat-vecs on the critical path.  Here everything that does not depend on the
// sweep is hoisted into stage-parallel passes, so that the sequential chains
// carry ONE nx x nx mat-vec per stage; the chains are cut into the factor's P
// segments, and the P segment boundaries are resolved through the same R-ary
// hierarchy (segment / group transitions Psi from K3 / K4):
//
//   pre    (stage-parallel)  g = -r1 + C'((z r3 + r4)/w);  wv = gx - Rux' gu;
//                            q = Vxx[k+1] f_k;  v[K] = gx_K
//   back   fine chains       v = wv + Phi'(v+ + q) from v_b = 0         -> segv0
//          up / top / down   t <- segv0[e] + Psi[e]' t over the hierarchy -> segvb
//          fine chains again from the true v_b, storing v[k]
//   mid    (stage-parallel)  Ru = Guu^{-1}(gu + fu'(v+ + q));  c = f - fu Ru
//   fwd    fine chains       x+ = Phi x + c from x_a = 0                -> segx0
//          up / top / down   t <- segx0[e] + Psi[e] t                    -> segxa
//          fine chains again from the true x_a, storing x[k]
//   post   (stage-parallel)  u = -(Rux x + Ru); p = Vxx+ x+ + v+; dx,dy,dw,dz
//
// Every chain is one instance of chain_run(): matrices and vectors of the next
// steps are staged into a shared-memory ring by TMA bulk copies ahead of use.
#pragma once

#include "lq_device.cuh"

#define LQ_RING 4  // default stages per staged chunk of a segment chain (two chunk buffers)

// One sequential affine chain executed by a whole CTA (4 lanes per row):
//     for j = 0..cnt-1, e = first + j*dir:
//         if pre:  pre[e] = t
//         t <- a[e] + op(M[e]) (t + b[e])          op = transpose if TRANS
//         if post: post[e + post_shift] = t
// M: n x n row-major blocks (stride n*n), a/b/pre/post: n-vectors (stride n).
// t lives in registers (valid for part == 0 lanes and replicated in the quad).
// The operands of CH consecutive steps are contiguous in memory, so each chunk
// is staged with THREE bulk copies (M, a, b) on one mbarrier -- issuing one
// small copy per step from a single thread costs more than the step itself --
// and the next chunk is in flight while the current one is consumed (two chunk
// buffers).  buf: 2 * CH * (n*n + 2n) doubles, bars: 2, tvec: 2n doubles.
template <bool TRANS>
__device__ __forceinline__ double chain_run(int n, bool use_tma, const double *M,
                                            const double *a, const double *b, double *pre,
                                            double *post, int post_shift, int first, int dir,
                                            int cnt, double t, double *buf, uint64_t *bars,
                                            double *tvec, int CH) {
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int n2 = n * n;
  const size_t chunk_sz = (size_t)CH * (n2 + 2 * n);
  const int nchunk = (cnt + CH - 1) / CH;
  // chunk c covers steps [c*CH, c*CH + len): elements e_lo .. e_lo + len - 1
  auto issue = [&](int c) {
    const int j0 = c * CH, len = min(CH, cnt - j0);
    const int e_first = first + j0 * dir;
    const int e_lo = dir > 0 ? e_first : e_first - (len - 1);
    double *cb = buf + (size_t)(c & 1) * chunk_sz;
    uint64_t *bar = &bars[c & 1];
    const uint32_t bm = (uint32_t)len * n2 * 8, bv = (uint32_t)len * n * 8;
    mbar_expect_tx(bar, bm + bv + (b ? bv : 0));
    tma_load_1d(cb, M + (size_t)e_lo * n2, bm, bar);
    tma_load_1d(cb + (size_t)CH * n2, a + (size_t)e_lo * n, bv, bar);
    if (b) tma_load_1d(cb + (size_t)CH * (n2 + n), b + (size_t)e_lo * n, bv, bar);
  };
  if (use_tma && threadIdx.x == 0) {
    issue(0);
    if (nchunk > 1) issue(1);
  }
  for (int c = 0; c < nchunk; c++) {
    const int j0 = c * CH, len = min(CH, cnt - j0);
    const int e_first = first + j0 * dir;
    const int e_lo = dir > 0 ? e_first : e_first - (len - 1);
    const double *cb = buf + (size_t)(c & 1) * chunk_sz;
    if (use_tma) mbar_wait(&bars[c & 1], (c >> 1) & 1);
    for (int jj = 0; jj < len; jj++) {
      const int j = j0 + jj;
      const int e = e_first + jj * dir;
      const double *Ms, *as, *bs;
      if (use_tma) {
        const int o = e - e_lo;
        Ms = cb + (size_t)o * n2;
        as = cb + (size_t)CH * n2 + (size_t)o * n;
        bs = b ? cb + (size_t)CH * (n2 + n) + (size_t)o * n : nullptr;
      } else {
        Ms = M + (size_t)e * n2;
        as = a + (size_t)e * n;
        bs = b ? b + (size_t)e * n : nullptr;
      }
      double *tv = tvec + (j & 1) * n;  // double buffer: one barrier per step
      if (i < n && part == 0) {
        if (pre) pre[(size_t)e * n + i] = t;
        tv[i] = bs ? t + bs[i] : t;
      }
      __syncthreads();
      double s = 0.0;
      if (i < n) {
        if (TRANS) {
          for (int l = part; l < n; l += 4) s = fma(Ms[l * n + i], tv[l], s);
        } else {
          for (int l = part; l < n; l += 4) s = fma(Ms[i * n + l], tv[l], s);
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (i < n) {
        t = as[i] + s;
        if (post && part == 0) post[(size_t)(e + post_shift) * n + i] = t;
      }
    }
    if (use_tma && c + 2 < nchunk) {
      // this chunk buffer is refilled only after every thread has read it
      __syncthreads();
      if (threadIdx.x == 0) {
        fence_proxy_async();
        issue(c + 2);
      }
    }
  }
  return t;
}

struct ChainSmem {
  double *ring, *tvec;
  uint64_t *bars;
};

// (the ring only exists on the bulk-copy path; without it the chain reads its
//  matrices from global memory and CH is just the loop blocking)
__device__ __forceinline__ ChainSmem chain_smem_init(int n, void *raw, int CH, bool ring = true) {
  SmemCarver sm(raw);
  ChainSmem cs;
  cs.ring = ring ? sm.take(2 * CH * (n * n + 2 * n)) : nullptr;
  cs.tvec = sm.take(3 * n);
  cs.bars = sm.take_bars(2);
  if (threadIdx.x == 0) {
    mbar_init(&cs.bars[0], 1);
    mbar_init(&cs.bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  return cs;
}

// shared-memory loads as PTX: `asm volatile` keeps nvcc from moving or merging
// them (the prefetch and the dependent-path loads stay where they are written),
// the plain ld.shared leaves the final schedule to ptxas
__device__ __forceinline__ double lds_ordered(const double *p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];"
               : "=d"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ double2 lds_ordered2(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

// The same chain run by ONE WARP (lane i owns rows i, i + 32, ...): no CTA
// barrier on the dependent path.  Per step the vector goes through a
// double-buffered shared row (one store per lane, __syncwarp, broadcast reads:
// ~40 cycles, against ~5 per shuffled element) and each lane accumulates its own
// rows in four independent chains.
//   TRANS : lane i reads column i of M, M[l][i]: consecutive lanes, no conflict
//   !TRANS: lane i reads row i in the skewed order l' = (l + i) mod N, so that
//           the 32 simultaneous reads M[i][l'] fall into distinct banks (row
//           stride N = 20, 12, 40 doubles would otherwise collide 2- to 8-way)
// The matrices are staged by the chunked TMA ring; t[] is in/out.  tvec: 2 N.
template <bool TRANS, int N>
__device__ __forceinline__ void chain_run_warp(const double *M, const double *a, const double *b,
                                               double *pre, double *post, int post_shift,
                                               int first, int dir, int cnt,
                                               double (&t)[(N + 31) / 32], double *buf,
                                               uint64_t *bars, double *tvec, int CH) {
  static_assert(N % 4 == 0, "chain_run_warp: four accumulator chains, double2 reads");
  constexpr int NR = (N + 31) / 32, n2 = N * N;
  const int lane = threadIdx.x & 31;
  const size_t chunk_sz = (size_t)CH * (n2 + 2 * N);
  const int nchunk = (cnt + CH - 1) / CH;
  auto issue = [&](int c) {
    const int j0 = c * CH, len = min(CH, cnt - j0);
    const int e_first = first + j0 * dir;
    const int e_lo = dir > 0 ? e_first : e_first - (len - 1);
    double *cb = buf + (size_t)(c & 1) * chunk_sz;
    uint64_t *bar = &bars[c & 1];
    const uint32_t bm = (uint32_t)len * n2 * 8, bv = (uint32_t)len * N * 8;
    mbar_expect_tx(bar, bm + bv + (b ? bv : 0));
    tma_load_1d(cb, M + (size_t)e_lo * n2, bm, bar);
    tma_load_1d(cb + (size_t)CH * n2, a + (size_t)e_lo * N, bv, bar);
    if (b) tma_load_1d(cb + (size_t)CH * (n2 + N), b + (size_t)e_lo * N, bv, bar);
  };
  if (lane == 0) {
    issue(0);
    if (nchunk > 1) issue(1);
  }
  int par = 0;
  // matrix entries of one lane for one step (row `rs`): column rs of M (TRANS) or
  // row rs in skewed order
  auto load_m = [&](double (&m)[N], const double *Ms, int rs) {
#pragma unroll
    for (int l = 0; l < N; l++) {
      if (TRANS) {
        m[l] = lds_ordered(Ms + l * N + rs);
      } else {
        const int cidx = (l + rs < N) ? l + rs : l + rs - N;  // (l + row) mod N
        m[l] = lds_ordered(Ms + rs * N + cidx);
      }
    }
  };
  // the vector as this lane needs it: all of it (TRANS) or in the skewed order
  auto load_t = [&](double (&tl)[N], const double *tv, int rs) {
    if (TRANS) {
#pragma unroll
      for (int l = 0; l < N; l += 2) {
        const double2 ta = lds_ordered2(tv + l);
        tl[l] = ta.x;
        tl[l + 1] = ta.y;
      }
    } else {
#pragma unroll
      for (int l = 0; l < N; l++) {
        const int cidx = (l + rs < N) ? l + rs : l + rs - N;
        tl[l] = lds_ordered(tv + cidx);
      }
    }
  };
  auto dot = [&](const double (&m)[N], const double (&tl)[N]) -> double {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
    for (int l = 0; l < N; l += 4) {
      acc0 = fma(m[l], tl[l], acc0);
      acc1 = fma(m[l + 1], tl[l + 1], acc1);
      acc2 = fma(m[l + 2], tl[l + 2], acc2);
      acc3 = fma(m[l + 3], tl[l + 3], acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
  };
  for (int c = 0; c < nchunk; c++) {
    const int j0 = c * CH, len = min(CH, cnt - j0);
    const int e_first = first + j0 * dir;
    const int e_lo = dir > 0 ? e_first : e_first - (len - 1);
    const double *cb = buf + (size_t)(c & 1) * chunk_sz;
    mbar_wait(&bars[c & 1], (c >> 1) & 1);
    // one step; the matrix entries of the NEXT step (same chunk) are requested
    // before this step's multiply-adds: they do not depend on the chain, and the
    // ~30 shared-memory loads of a step otherwise sit on its dependent path
    auto step = [&](int jj, double (&mcur)[N], double (&mnext)[N]) {
      const int e = e_first + jj * dir, o = e - e_lo;
      const double *Ms = cb + (size_t)o * n2;
      const double *as = cb + (size_t)CH * n2 + (size_t)o * N;
      const double *bs = b ? cb + (size_t)CH * (n2 + N) + (size_t)o * N : nullptr;
      double *tv = tvec + par * N;
      par ^= 1;
#pragma unroll
      for (int r = 0; r < NR; r++) {
        const int row = lane + 32 * r;
        if (row < N) {
          if (pre) pre[(size_t)e * N + row] = t[r];
          tv[row] = bs ? t[r] + bs[row] : t[r];
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < NR; r++) {
        const int row = lane + 32 * r;
        const bool act = row < N;
        const int rs = act ? row : 0;
        double s, tl[N];
        load_t(tl, tv, rs);  // first in the queue: these are on the dependent path
        if (NR == 1) {
          if (jj + 1 < len) load_m(mnext, Ms + dir * n2, rs);
          s = dot(mcur, tl);
        } else {
          double m[N];
          load_m(m, Ms, rs);
          s = dot(m, tl);
        }
        if (act) {
          t[r] = as[row] + s;
          if (post) post[(size_t)(e + post_shift) * N + row] = t[r];
        }
      }
    };
    double mA[N], mB[N];
    if (NR == 1) load_m(mA, cb + (size_t)(e_first - e_lo) * n2, lane < N ? lane : 0);
    int jj = 0;
    for (; jj + 1 < len; jj += 2) {
      step(jj, mA, mB);
      step(jj + 1, mB, mA);
    }
    if (jj < len) step(jj, mA, mB);
    if (c + 2 < nchunk) {
      // this chunk buffer is refilled only after every lane has read it
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();
        issue(c + 2);
      }
    }
  }
}

// Stage-parallel passes run one WARP per stage (no CTA barriers, coalesced
// column-wise reads of the stage blocks) and several stages per warp, so that a
// pass is a few hundred fat CTAs instead of K tiny ones (CTA launch rate, not
// bandwidth, limited the one-CTA-per-stage version).
#ifndef LQ_WPB
#define LQ_WPB 4                       // warps per CTA
#endif


// ---- pre ------------------------------------------------------------------
// grid (ceil((K+1)/(spw*LQ_WPB)), batch), block 32*LQ_WPB; smem: LQ_WPB * (nm + nx) doubles
// W lanes per stage (32: one stage per warp at a time; 16: two -- for nm <= 16 a
// full warp per stage leaves more than half of the lanes idle).
template <int SPW, int W>
__global__ void __launch_bounds__(32 * LQ_WPB) solve_pre_kernel(
    LqDev d, const double *__restrict__ r1, const double *__restrict__ r2,
    const double *__restrict__ r3, const double *__restrict__ r4) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = 32 / W;
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / W, sl = lane % W;
  double *gk = reinterpret_cast<double *>(smem_raw) + (warp * G + sub) * (nm + nx);  // nm
  double *fk = gk + nm;                                                              // nx
  const int b = blockIdx.y;
  const double *z = d.z + (size_t)b * d.m, *w = d.w + (size_t)b * d.m;
  const double *cv = d.cval + (size_t)b * d.nnz;
  const double *r3b = r3 + (size_t)b * d.m, *r4b = r4 + (size_t)b * d.m;
  for (int s = 0; s < SPW; s++) {
    const int k0 = (blockIdx.x * LQ_WPB + warp) * G * SPW + s;  // stage of lane group 0
    if (k0 > d.K) break;                                        // (uniform in the warp)
    const int k = k0 + sub * SPW;
    const bool on = k <= d.K;
    const int dk = (k < d.K) ? nm : nx;
    const size_t xo = (size_t)b * d.N + (size_t)k * nm;
    if (on) {
      for (int i = sl; i < dk; i += W) {
        double a = -r1[xo + i];
        const int gv = k * nm + i;
        for (int e = d.vcol_ptr[gv]; e < d.vcol_ptr[gv + 1]; e++) {
          const int r = d.vcol_row[e];
          a = fma(cv[d.vcol_nz[e]], (z[r] * r3b[r] + r4b[r]) / w[r], a);
        }
        gk[i] = a;
        d.g[xo + i] = a;
      }
      if (k < d.K)
        for (int i = sl; i < nx; i += W) fk[i] = r2[(size_t)b * d.me + (size_t)k * nx + i];
    }
    __syncwarp();
    if (on) {
      if (k == d.K) {
        for (int i = sl; i < nx; i += W) d.v[((size_t)b * (d.K + 1) + k) * nx + i] = gk[i];
      } else {
        const size_t ks = (size_t)b * d.K + k;
        const double *Rux = d.Rux + ks * nu * nx;
        const double *Vp = d.V + ((size_t)b * (d.K + 1) + k + 1) * nx * nx;
        for (int i = sl; i < nx; i += W) {
          double a0 = gk[i], a1 = 0.0;
#pragma unroll 10
          for (int l = 0; l < nu; l++) a0 = fma(-Rux[l * nx + i], gk[nx + l], a0);
#pragma unroll 10
          for (int l = 0; l < nx; l++) a1 = fma(Vp[l * nx + i], fk[l], a1);  // Vxx symmetric
          d.wv[ks * nx + i] = a0;
          d.q[ks * nx + i] = a1;
        }
      }
    }
    __syncwarp();
  }
}

// ---- fine chains -------------------------------------------------------------
// grid (P, batch).  NX > 0: one warp per chain (chain_run_warp), 32 threads;
// NX == 0: any nx, block = 4 * ceil32(nx) threads (chain_run).
// back: mode 0 from 0, writes segv0[s]; mode 1 from segvb[s], stores v[k]
template <int NX>
__global__ void solve_back_kernel(LqDev d, int mode, int ring_n) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx;
  ChainSmem cs = chain_smem_init(nx, smem_raw, ring_n, NX > 0 || d.use_tma);
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const size_t so = ((size_t)b * d.st.nel + s) * nx;
  const size_t ks0 = (size_t)b * d.K;
  double *vout = mode == 1 ? d.v + (size_t)b * (d.K + 1) * nx : nullptr;
  if constexpr (NX > 0) {
    constexpr int NR = (NX + 31) / 32;
    const int lane = threadIdx.x;
    double t[NR];
#pragma unroll
    for (int r = 0; r < NR; r++)
      t[r] = (mode == 1 && lane + 32 * r < NX) ? d.segvb[so + lane + 32 * r] : 0.0;
    chain_run_warp<true, NX>(d.Phi + ks0 * nx * nx, d.wv + ks0 * nx, d.q + ks0 * nx, nullptr, vout,
                             0, kb - 1, -1, kb - ka, t, cs.ring, cs.bars, cs.tvec, ring_n);
    if (mode == 0) {
#pragma unroll
      for (int r = 0; r < NR; r++)
        if (lane + 32 * r < NX) d.segv0[so + lane + 32 * r] = t[r];
    }
  } else {
    const int i = threadIdx.x >> 2;
    double t = 0.0;
    if (mode == 1 && i < nx) t = d.segvb[so + i];
    t = chain_run<true>(nx, d.use_tma, d.Phi + ks0 * nx * nx, d.wv + ks0 * nx, d.q + ks0 * nx,
                        nullptr, vout, 0, kb - 1, -1, kb - ka, t, cs.ring, cs.bars, cs.tvec,
                        ring_n);
    if (mode == 0 && i < nx && (threadIdx.x & 3) == 0) d.segv0[so + i] = t;
  }
}

// fwd: mode 0 from x_a = 0, writes segx0[s] (x at the segment end);
//      mode 1 from segxa[s], stores x[k], k = a..b
template <int NX>
__global__ void solve_fwd_kernel(LqDev d, int mode, int ring_n) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx;
  ChainSmem cs = chain_smem_init(nx, smem_raw, ring_n, NX > 0 || d.use_tma);
  const int s = blockIdx.x, b = blockIdx.y;
  const int ka = s * d.L, kb = min(d.K, ka + d.L);
  const size_t so = ((size_t)b * d.st.nel + s) * nx;
  const size_t ks0 = (size_t)b * d.K;
  double *xb = d.x + (size_t)b * (d.K + 1) * nx;
  if constexpr (NX > 0) {
    constexpr int NR = (NX + 31) / 32;
    const int lane = threadIdx.x;
    double t[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int row = lane + 32 * r;
      t[r] = 0.0;
      if (mode == 1 && row < NX) {
        t[r] = d.segxa[so + row];
        xb[(size_t)ka * nx + row] = t[r];
      }
    }
    chain_run_warp<false, NX>(d.Phi + ks0 * nx * nx, d.c + ks0 * nx, nullptr, nullptr,
                              mode == 1 ? xb : nullptr, 1, ka, +1, kb - ka, t, cs.ring, cs.bars,
                              cs.tvec, ring_n);
    if (mode == 0) {
#pragma unroll
      for (int r = 0; r < NR; r++)
        if (lane + 32 * r < NX) d.segx0[so + lane + 32 * r] = t[r];
    }
  } else {
    const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
    double t = 0.0;
    if (mode == 1 && i < nx) {
      t = d.segxa[so + i];
      if (part == 0) xb[(size_t)ka * nx + i] = t;
    }
    t = chain_run<false>(nx, d.use_tma, d.Phi + ks0 * nx * nx, d.c + ks0 * nx, nullptr, nullptr,
                         mode == 1 ? xb : nullptr, 1, ka, +1, kb - ka, t, cs.ring, cs.bars,
                         cs.tvec, ring_n);
    if (mode == 0 && i < nx && part == 0) d.segx0[so + i] = t;
  }
}

// ---- hierarchy scans -----------------------------------------------------------
// phase 0 (up)  : CTA g composes the children of element g of level lev+1:
//                 from t = 0, t <- v0[e] + Psi[e]^(T) t; result -> v0 of (lev+1, g)
// phase 1 (top) : one CTA per instance walks all elements of level lev from the
//                 boundary value (v[K] backward / x_0 forward), recording the
//                 value at every element's far side -> vb / xa
// phase 2 (down): CTA g walks the children of (lev+1, g) from that element's
//                 vb / xa, recording the children's values
// backward (BACK = true): children visited last -> first with Psi'.
template <bool BACK, int NX>
__global__ void solve_scan_kernel(LqDev d, int lev, int phase, const double *__restrict__ r2,
                                  int ring_n) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = NX > 0 ? NX : d.nx;
  ChainSmem cs = chain_smem_init(nx, smem_raw, ring_n, NX > 0 || d.use_tma);
  const int g = blockIdx.x, b = blockIdx.y;
  const size_t eb = (size_t)b * d.st.nel + d.st.off[lev];           // first element of level
  const size_t pb = (size_t)b * d.st.nel + (phase == 1 ? 0 : d.st.off[lev + 1]);
  double *in0 = BACK ? d.segv0 : d.segx0;   // zero-boundary solutions
  double *bnd = BACK ? d.segvb : d.segxa;   // boundary values
  // start value of row `row` (phase 1: the boundary condition of the horizon)
  double *scr = cs.tvec;
  if (phase == 1 && !BACK && !d.has_prev && !d.fixed_x0) {
    // x_0 = -Vxx[0]^{-1} v[0] (hqp/Hqp_IpLQDOCP.C:2111-2117)
    for (int r = threadIdx.x; r < nx; r += blockDim.x)
      scr[r] = -d.v[(size_t)b * (d.K + 1) * nx + r];
    __syncthreads();
    // runtime dims on purpose: a serial one-off solve, not worth unrolling
    if (threadIdx.x == 0) thread_ldlt_solve(d.V0f + (size_t)b * d.nx * d.nx, d.nx, d.nx, scr, 1);
    __syncthreads();
  }
  auto start = [&](int row) -> double {
    if (row >= nx) return 0.0;
    if (phase == 0) return 0.0;
    if (phase == 2) return bnd[(pb + g) * nx + row];
    if (BACK) return d.v[((size_t)b * (d.K + 1) + d.K) * nx + row];
    if (d.has_prev) return d.xstart[row];  // horizon split: state from the ranks before
    if (d.fixed_x0)                         // x_0 = -a_0 (hqp/Hqp_IpLQDOCP.C:2099-2100)
      return -r2[(size_t)b * d.me + (size_t)d.K * nx + row];
    return scr[row];
  };
  int c0 = 0, c1 = d.st.cnt[lev];
  if (phase != 1) {
    c0 = g * d.st.R;
    c1 = min(d.st.cnt[lev], c0 + d.st.R);
  }
  const int first = BACK ? c1 - 1 : c0, dir = BACK ? -1 : +1;
  if constexpr (NX > 0) {
    constexpr int NR = (NX + 31) / 32;
    const int lane = threadIdx.x;
    double t[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) t[r] = start(lane + 32 * r);
    chain_run_warp<BACK, NX>(d.segPsi + eb * nx * nx, in0 + eb * nx, nullptr,
                             phase == 0 ? nullptr : bnd + eb * nx, nullptr, 0, first, dir,
                             c1 - c0, t, cs.ring, cs.bars, cs.tvec, ring_n);
    if (phase == 0) {
#pragma unroll
      for (int r = 0; r < NR; r++)
        if (lane + 32 * r < NX) in0[(pb + g) * nx + lane + 32 * r] = t[r];
    }
  } else {
    const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
    double t = start(i);
    __syncthreads();  // scr (inside tvec) is reused by the chain
    t = chain_run<BACK>(nx, d.use_tma, d.segPsi + eb * nx * nx, in0 + eb * nx, nullptr,
                        phase == 0 ? nullptr : bnd + eb * nx, nullptr, 0, first, dir, c1 - c0, t,
                        cs.ring, cs.bars, cs.tvec, ring_n);
    if (phase == 0 && i < nx && part == 0) in0[(pb + g) * nx + i] = t;
  }
}

// warp-cooperative solve of (L D L') y = b, y in shared memory (m <= 32 per pass
// of 32 lanes; larger m loops).  LD row-major in GLOBAL memory (ld = m): the
// column sweeps read it once, coalesced per column pair.
// `sl`, W: lane index inside, and width of, the lane group that owns this system
// (W = 32: the whole warp; W = 16: two systems per warp).  `on` = false: the group
// has no system but takes part in the warp barriers.
__device__ __forceinline__ void warp_ldlt_solve_g(const double *__restrict__ LD, int m,
                                                  double *y, int sl, int W = 32,
                                                  bool on = true) {
  // forward: for each column l, rows i > l subtract L[i][l] y[l]
  for (int l = 0; l < m; l++) {
    if (on) {
      const double yl = y[l];
      for (int i = l + 1 + sl; i < m; i += W) y[i] = fma(-LD[i * m + l], yl, y[i]);
    }
    __syncwarp();
  }
  if (on)
    for (int i = sl; i < m; i += W) y[i] *= LD[i * m + i];
  __syncwarp();
  // backward: for each row l (descending), entries i < l subtract L[l][i] y[l]
  for (int l = m - 1; l > 0; l--) {
    if (on) {
      const double yl = y[l];
      for (int i = sl; i < l; i += W) y[i] = fma(-LD[l * m + i], yl, y[i]);
    }
    __syncwarp();
  }
}

// ---- mid -------------------------------------------------------------------
// grid (ceil(K/(spw*LQ_WPB)), batch), block 32*LQ_WPB; smem: LQ_WPB * (nx + nu) doubles
template <int SPW, int W>
__global__ void __launch_bounds__(32 * LQ_WPB) solve_mid_kernel(LqDev d, const double *__restrict__ r2) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = 32 / W;
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / W, sl = lane % W;
  double *t = reinterpret_cast<double *>(smem_raw) + (warp * G + sub) * (nx + nu + nu * nu);  // nx
  double *Gu = t + nx;                                                                        // nu
  double *LDs = Gu + nu;                                                                      // nu x nu
  const int b = blockIdx.y;
  for (int s = 0; s < SPW; s++) {
    const int k0 = (blockIdx.x * LQ_WPB + warp) * G * SPW + s;
    if (k0 >= d.K) break;  // (uniform in the warp)
    const int k = k0 + sub * SPW;
    const bool on = k < d.K;
    const size_t ks = (size_t)b * d.K + (on ? k : 0);
    const double *vp = d.v + ((size_t)b * (d.K + 1) + (on ? k : 0) + 1) * nx;
    const double *fu = d.fu + ks * nx * nu;
    const double *g = d.g + (size_t)b * d.N + (size_t)(on ? k : 0) * nm;
    if (on) {
      // the factor of Guu goes to shared memory up front: the substitution below
      // would otherwise pay a global-memory latency on each of its 2 nu dependent steps
      const double *LDg = d.LD + ks * nu * nu;
#pragma unroll 4
      for (int i = sl; i < nu * nu; i += W) LDs[i] = LDg[i];
      for (int i = sl; i < nx; i += W) t[i] = vp[i] + d.q[ks * nx + i];
    }
    __syncwarp();
    if (on)
      for (int j = sl; j < nu; j += W) {
        double a = g[nx + j];
#pragma unroll 10
        for (int l = 0; l < nx; l++) a = fma(fu[l * nu + j], t[l], a);
        Gu[j] = a;
      }
    __syncwarp();
    // LD holds the LDL^T factor of Guu or, for an indefinite block handled by the
    // pivoted fallback of the factor, its explicit inverse (d.ldkind)
    const bool expl = on && d.ldkind[ks] != 0;
    warp_ldlt_solve_g(LDs, nu, Gu, sl, W, on && !expl);
    if (__any_sync(0xffffffffu, expl)) {  // (rare: an indefinite block in this warp's stages)
      // the pivoted fallback exists for nu <= W only (compiled sizes, nu <= 32 / 16):
      // one lane per row, all of Gu read before any of it is overwritten
      double a = 0.0;
      if (expl && sl < nu)
        for (int l = 0; l < nu; l++) a = fma(LDs[sl * nu + l], Gu[l], a);
      __syncwarp();
      if (expl && sl < nu) Gu[sl] = a;
      __syncwarp();
    }
    if (on) {
      for (int j = sl; j < nu; j += W) d.Ru[ks * nu + j] = Gu[j];
      for (int i = sl; i < nx; i += W) {
        double a = r2[(size_t)b * d.me + (size_t)k * nx + i];
#pragma unroll 10
        for (int l = 0; l < nu; l++) a = fma(-fu[i * nu + l], Gu[l], a);
        d.c[ks * nx + i] = a;
      }
    }
    __syncwarp();
  }
}

// ---- post -------------------------------------------------------------------
// grid (ceil((K+1)/(spw*LQ_WPB)), batch), block 32*LQ_WPB; smem: LQ_WPB * (nm + nx) doubles
template <int SPW, int W>
__global__ void __launch_bounds__(32 * LQ_WPB) solve_post_kernel(
    LqDev d, const double *__restrict__ r3, const double *__restrict__ r4,
    double *__restrict__ dx, double *__restrict__ dy, double *__restrict__ dz,
    double *__restrict__ dw) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  constexpr int G = 32 / W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / W, sl = lane % W;
  double *xs = reinterpret_cast<double *>(smem_raw) + (warp * G + sub) * (nm + nx);  // nm: [x_k ; u_k] -> dx
  double *xn = xs + nm;                                                   // nx: x_{k+1}
  const int b = blockIdx.y;
  const double *cv = d.cval + (size_t)b * d.nnz;
  for (int s = 0; s < SPW; s++) {
    const int k0 = (blockIdx.x * LQ_WPB + warp) * G * SPW + s;
    if (k0 > d.K) break;  // (uniform in the warp)
    const int k = k0 + sub * SPW;
    const bool on = k <= d.K;
    const double *xk = d.x + ((size_t)b * (d.K + 1) + (on ? k : 0)) * nx;
    const size_t ks = (size_t)b * d.K + k;
    if (on)
      for (int i = sl; i < nx; i += W) {
        xs[i] = xk[i];
        if (k < d.K) xn[i] = xk[nx + i];
      }
    __syncwarp();
    if (on && k < d.K) {
      const double *Rux = d.Rux + ks * nu * nx;
      for (int j = sl; j < nu; j += W) {
        double a = d.Ru[ks * nu + j];
#pragma unroll 10
        for (int l = 0; l < nx; l++) a = fma(Rux[j * nx + l], xs[l], a);
        xs[nx + j] = -a;  // u_k
      }
      // p_k = Vxx[k+1] x[k+1] + v[k+1]   (:2169-2171)
      const double *Vp = d.V + ((size_t)b * (d.K + 1) + k + 1) * nx * nx;
      const double *vp = d.v + ((size_t)b * (d.K + 1) + k + 1) * nx;
      for (int i = sl; i < nx; i += W) {
        double a = vp[i];
#pragma unroll 10
        for (int l = 0; l < nx; l++) a = fma(Vp[l * nx + i], xn[l], a);
        dy[(size_t)b * d.me + (size_t)k * nx + i] = a;
      }
    }
    if (on && k == 0 && d.fixed_x0) {  // y_0 = -(Vx[0] + Vxx[0] x_0)   (:2153-2159)
      const double *V0 = d.V + (size_t)b * (d.K + 1) * nx * nx;
      const double *v0 = d.v + (size_t)b * (d.K + 1) * nx;
      for (int i = sl; i < nx; i += W) {
        double a = v0[i];
#pragma unroll 10
        for (int l = 0; l < nx; l++) a = fma(V0[l * nx + i], xs[l], a);
        dy[(size_t)b * d.me + (size_t)d.K * nx + i] = -a;
      }
    }
    __syncwarp();
    const int dk = (k < d.K) ? nm : nx;
    // dx = -[x;u]   (:952)
    if (on)
      for (int i = sl; i < dk; i += W) {
        const double a = -xs[i];
        xs[i] = a;
        dx[(size_t)b * d.N + (size_t)k * nm + i] = a;
      }
    __syncwarp();
    // dw = C dx - r3 ; dz = (r4 - z dw)/w   (:955-960)
    if (on)
      for (int rr = d.srow_ptr[k] + sl; rr < d.srow_ptr[k + 1]; rr += W) {
        const int r = d.srow[rr];
        double a = 0.0;
        for (int e = d.ineq_ptr[r]; e < d.ineq_ptr[r + 1]; e++)
          a = fma(cv[e], xs[d.ineq_lcol[e]], a);
        const size_t ro = (size_t)b * d.m + r;
        const double dwr = a - r3[ro];
        dw[ro] = dwr;
        dz[ro] = (r4[ro] - d.z[ro] * dwr) / d.w[ro];
      }
    __syncwarp();
  }
}

// ---- residuum ----------------------------------------------------------------
// Hqp_IpMatrix::residuum (hqp/Hqp_IpMatrix.C:131-178);
// grid (ceil((K+1)/LQ_RES_WPB), batch), block 32 * LQ_RES_WPB, smem LQ_RES_WPB (nm + nx) doubles.
// Writes the four residual vectors to t1..t4 (may be NULL) and atomically
// maxes their inf-norm into *res (must be zeroed before the launch).
#define LQ_RES_WPB 4  // stages (warps) per CTA
__global__ void __launch_bounds__(32 * LQ_RES_WPB)
residuum_kernel(LqDev d, const double *__restrict__ r1, const double *__restrict__ r2,
                const double *__restrict__ r3, const double *__restrict__ r4,
                const double *__restrict__ dx, const double *__restrict__ dy,
                const double *__restrict__ dz, const double *__restrict__ dw, double *t1,
                double *t2, double *t3, double *t4, double *res, const double *__restrict__ ety) {
  // one WARP per stage (round 1: one small CTA per stage -- CTA turnover, not
  // bandwidth, set its 55 us at C2); one atomic max per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = d.nx, nu = d.nu, nm = d.nm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *xs = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (nm + nx);  // nm: dx stage k
  double *yk = xs + nm;                                                        // nx: dy dynamics rows k
  __shared__ double red[LQ_RES_WPB];
  const int k = blockIdx.x * LQ_RES_WPB + warp, b = blockIdx.y;
  double mx = 0.0;
  bool bad = false;  // NaN seen
  if (k <= d.K) {
    const int dk = (k < d.K) ? nm : nx;
    const size_t xo = (size_t)b * d.N + (size_t)k * nm;
    const size_t yo = (size_t)b * d.me;
    if (k == d.K && d.has_next) {
      // horizon split: the trailing state block is the next range's x_0 -- its rows
      // belong to that rank
      if (t1)
        for (int i = lane; i < nx; i += 32) t1[xo + i] = 0.0;
    } else {
      // state after the last local stage / multiplier before the first: from the
      // neighbouring ranges when the horizon is split
      const double *xnext = (d.has_next && d.halo && k == d.K - 1)
                                ? d.halo + (size_t)(d.rank + 1) * 2 * nx : nullptr;
      const double *yprev = (d.has_prev && d.halo && k == 0)
                                ? d.halo + (size_t)(d.rank - 1) * 2 * nx + nx : nullptr;
      for (int i = lane; i < dk; i += 32) xs[i] = dx[xo + i];
      if (k < d.K)
        for (int i = lane; i < nx; i += 32) yk[i] = dy[yo + (size_t)k * nx + i];
      __syncwarp();
      const double *Qk = d.Q + ((size_t)b * (d.K + 1) + k) * nm * nm;
      const double *cv = d.cval + (size_t)b * d.nnz;
      const size_t ks = (size_t)b * d.K + k;
      // t1 = r1 + Q dx - A' dy - C' dz, rows of stage k
      for (int i = lane; i < dk; i += 32) {
        double s = r1[xo + i];
#pragma unroll 10
        for (int l = 0; l < dk; l++) s = fma(Qk[l * nm + i], xs[l], s);  // Q symmetric
        if (k < d.K) {
          if (i < nx) {
            const double *fx = d.fx + ks * nx * nx;
#pragma unroll 10
            for (int l = 0; l < nx; l++) s = fma(-fx[l * nx + i], yk[l], s);
          } else {
            const double *fu = d.fu + ks * nx * nu;
#pragma unroll 10
            for (int l = 0; l < nx; l++) s = fma(-fu[l * nu + (i - nx)], yk[l], s);
          }
        }
        if (i < nx) {
          if (k > 0) s += dy[yo + (size_t)(k - 1) * nx + i];  // -(-I)' dy_{k-1}
          else if (d.fixed_x0) s -= dy[yo + (size_t)d.K * nx + i];
          else if (yprev) s += yprev[i];
        }
        const int gv = k * nm + i;
        for (int e = d.vcol_ptr[gv]; e < d.vcol_ptr[gv + 1]; e++)
          s = fma(-cv[d.vcol_nz[e]], dz[(size_t)b * d.m + d.vcol_row[e]], s);
        if (ety) s -= ety[xo + i];  // general equality rows (batch == 1)
        if (t1) t1[xo + i] = s;
        mx = fmax(mx, fabs(s));
        bad |= (s != s);
      }
      // t2 = r2 - A dx, dynamics rows of stage k (+ x0 rows with stage 0)
      if (k < d.K) {
        const double *fx = d.fx + ks * nx * nx, *fu = d.fu + ks * nx * nu;
        for (int i = lane; i < nx; i += 32) {
          double s = xnext ? -xnext[i] : -dx[xo + nm + i];
#pragma unroll 10
          for (int l = 0; l < nx; l++) s = fma(fx[i * nx + l], xs[l], s);
#pragma unroll 10
          for (int l = 0; l < nu; l++) s = fma(fu[i * nu + l], xs[nx + l], s);
          const double t = r2[yo + (size_t)k * nx + i] - s;
          if (t2) t2[yo + (size_t)k * nx + i] = t;
          mx = fmax(mx, fabs(t));
          bad |= (t != t);
        }
      }
      if (k == 0 && d.fixed_x0)
        for (int i = lane; i < nx; i += 32) {
          const double t = r2[yo + (size_t)d.K * nx + i] - xs[i];
          if (t2) t2[yo + (size_t)d.K * nx + i] = t;
          mx = fmax(mx, fabs(t));
          bad |= (t != t);
        }
      // t3 = r3 - (C dx - dw) ; t4 = r4 - (z dw + w dz), rows of stage k
      for (int rr = d.srow_ptr[k] + lane; rr < d.srow_ptr[k + 1]; rr += 32) {
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
    }
  }
  if (bad) mx = __longlong_as_double(0x7ff8000000000000LL);
  // warp, then CTA max; NaN is propagated explicitly (fmax would drop it)
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (mx != mx) ? mx : ((other != other) ? other : fmax(mx, other));
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < LQ_RES_WPB; i++) {
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
