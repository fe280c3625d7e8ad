// Device-side data layout and CTA-level dense FP64 primitives for the
// stage-structured KKT factor/solve (sm_100a).
//
// Everything a kernel needs is passed by value in one LqDev struct.  All stage
// blocks live in contiguous HBM slabs (the reference keeps one heap block per
// stage, hqp/t_mesch.h:43-140); index = (instance * stages + stage) * block.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define LQ_MAXLEV 16
#define LQ_NT 128  // threads per CTA of the segment kernels K1/K3 (compile-time)
#define LQ_NT2 256 // threads per CTA of the tree kernels (compose / scan): quads x columns

// R-ary hierarchy over the P level-0 segments of one instance
struct LqTree {
  int R;                // group size
  int nlev;             // number of levels (level 0 = segments)
  int cnt[LQ_MAXLEV];   // elements per level
  int off[LQ_MAXLEV];   // element offset of the level inside the per-instance arrays
  int nel;              // total elements per instance
};

struct LqDev {
  int K, nx, nu, nm, batch;
  int P, L;           // level-0 segments per instance, stages per segment
  LqTree ft;          // factor hierarchy (segA/segC/segJ/segVb), binary: every
                      // level is one fully parallel combine deep
  LqTree st;          // solve hierarchy (segPsi and the boundary vectors), wide:
                      // its per-element work is one mat-vec, launches dominate
  int fixed_x0;
  int m, nnz;         // inequality rows / nonzeros per instance
  int N, me;          // per-instance vector lengths
  int use_tma;        // per-stage slabs are 16-byte multiples: bulk-copy path
  // horizon split across ranks (lq_range.cuh): this handle owns a contiguous
  // stage range of a longer horizon
  int hs;             // factor tree as a one-sweep suffix scan (elem_hs_kernel): the last
                      // segment carries the terminal value, log2(P) levels of P combines
  int spw;            // stages per warp of the stage-parallel solve passes
  int lgw;            // lanes per stage there (32, or 16: two stages side by side)
  int has_prev;       // a rank before this one supplies the state at stage 0
  int has_next;       // a rank behind this one supplies the terminal value
  double *Vext;       // [nx*nx] value Hessian handed over from the ranks behind
  double *xstart;     // [nx]    state at stage 0 handed over from the ranks before
  // Large stage blocks (nx > 64): the CTA-internal blocks of the factor kernels do
  // not fit the 227 KB of shared memory; they then live in a per-CTA slice of this
  // global workspace (L2-resident) and shared memory only stages GEMM panels
  double *gws;          // NULL: blocks in shared memory
  size_t gws_stride;    // doubles per CTA
  // halo of the residual passes: [world][2 nx] = (first state, last dynamics
  // multiplier) of every range, all-gathered before the pass (NULL: no split)
  const double *halo;
  int rank, world;
  // inequality structure (shared by all instances)
  const int *ineq_stage, *ineq_ptr, *ineq_lcol;
  const int *srow_ptr;  // [K+2] rows sorted by stage
  const int *srow;      // [m]
  const int *grow_ptr;  // [K+2] rows with more than one nonzero, sorted by stage
  const int *grow;      // [..]
  const int *vcol_ptr;  // [N+1] per variable: entries of C in that column
  const int *vcol_row;  // [nnz] row of the entry
  const int *vcol_nz;   // [nnz] index of the entry in the CSR value array
  // values (update)
  const double *Q, *fx, *fu, *cval;
  // captured at factor
  const double *z, *w;
  double *hdiag;   // [batch][N] diagonal of C'(z/w)C contributed by single-entry rows
  // factor state
  double *V;       // [batch][K+1][nx*nx]   value-function Hessians Vxx
  double *Rux;     // [batch][K][nu*nx]     feedback gains
  double *LD;      // [batch][K][nu*nu]     LDL^T of Guu: unit L below, 1/D on the diagonal
  int *ldkind;     // [batch][K] 0: LD holds the LDL^T factor; 1: the explicit inverse of an
                   //            indefinite Guu from the scaled, pivoted fallback
  double *Phi;     // [batch][K][nx*nx]     closed loop fx - fu Rux
  // factor hierarchy: element e of level l at [(b*ft.nel + ft.off[l] + e) * nx*nx]
  double *segA, *segC, *segJ;  // boundary elements (zero-terminal-cost condensation)
  double *segVb;   // Vxx at the element's end
  // solve hierarchy: element e of level l at [(b*st.nel + st.off[l] + e) * ...]
  double *segPsi;  // closed-loop transition across the element
  double *V0f;     // [batch][nx*nx]        LDL^T of Vxx[0] (free x0)
  int *status;     // device status word (0 ok)
  long long *dbg;  // [16] clock64 stamps (LQ_TIMING builds only)
  // solve scratch
  double *g;       // [batch][N]     reduced gradient (gx,gu)
  double *wv;      // [batch][K][nx] gx - Rux' gu
  double *q;       // [batch][K][nx] Vxx[k+1] f_k
  double *v;       // [batch][K+1][nx]
  double *Ru;      // [batch][K][nu]
  double *c;       // [batch][K][nx] f_k - fu Ru
  double *x;       // [batch][K+1][nx]
  double *segv0, *segvb, *segx0, *segxa;  // [batch][st.nel][nx]
};

#ifdef LQ_TIMING
__device__ long long g_dbg[32];  // clock64 stamps of one CTA (timing builds only)
#define LQ_STAMP2(i) do { if (threadIdx.x == 0 && blockIdx.x == 7) g_dbg[i] = clock64(); } while (0)
#else
#define LQ_STAMP2(i) do { } while (0)
#endif

// Programmatic dependent launch: let the next kernel of the stream be scheduled
// now (it blocks in its own pdl_enter()), then wait until the previous kernel
// has completed and its writes are visible.  No-ops without a programmatic
// dependency.  Must run before the first access to global memory.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// warp index as a value the compiler knows to be warp-uniform (a shuffle
// result): branches on it are uniform, so the *_sync collectives inside
// warp-specialised code are not wrapped in WARPSYNC / ENDCOLLECTIVE pairs
__device__ __forceinline__ int warp_id_uniform() {
  return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
}

// status word bits (device) -> HQPCU_E_SING / HQPCU_E_NOTPD (host)
#define LQ_FLAG_SING 1
#define LQ_FLAG_NOTPD 2

// ---------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers: cp.async.bulk global -> shared,
// completion signalled on an mbarrier (SASS: UBLKCP / SYNCS).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy accesses to a buffer must be ordered before the async proxy
// (TMA) overwrites it
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------
// C(MxN, ldc) = beta*C0 + alpha * A * B with arbitrary element strides:
//   A(i,l) = A[i*ar + l*ac],  B(l,j) = B[l*br + j*bc].
// Each thread owns 2x2 output tiles (register blocking: 4 loads per 4 FMAs);
// tiles are dealt round-robin over [tid, tid+nthr, ...).  All sizes and strides
// are plain ints: inside the <NX,NU>-templated kernels they are compile-time
// constants after inlining, so the l-loop unrolls into immediate-offset LDS +
// DFMA with no address arithmetic; the <0,0> instantiation keeps runtime loops.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cta_mm(double *C, int ldc, const double *C0, int ldc0,
                                       double beta, double alpha, const double *A, int ar,
                                       int ac, const double *B, int br, int bc, int M, int N,
                                       int Kd, int tid, int nthr) {
  const int TI = (M + 1) >> 1, TJ = (N + 1) >> 1;
  for (int t = tid; t < TI * TJ; t += nthr) {
    const int ti = t / TJ, tj = t - ti * TJ;
    const int i0 = 2 * ti, j0 = 2 * tj;
    const bool i1 = i0 + 1 < M, j1 = j0 + 1 < N;
    const double *a0 = A + i0 * ar;
    const double *a1 = i1 ? a0 + ar : a0;
    const double *b0 = B + j0 * bc;
    const double *b1 = j1 ? b0 + bc : b0;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
    for (int l = 0; l < Kd; l++) {
      const double x0 = a0[l * ac], x1 = a1[l * ac];
      const double y0 = b0[l * br], y1 = b1[l * br];
      c00 = fma(x0, y0, c00);
      c01 = fma(x0, y1, c01);
      c10 = fma(x1, y0, c10);
      c11 = fma(x1, y1, c11);
    }
    c00 *= alpha; c01 *= alpha; c10 *= alpha; c11 *= alpha;
    if (C0) {
      c00 = fma(beta, C0[i0 * ldc0 + j0], c00);
      if (j1) c01 = fma(beta, C0[i0 * ldc0 + j0 + 1], c01);
      if (i1) c10 = fma(beta, C0[(i0 + 1) * ldc0 + j0], c10);
      if (i1 && j1) c11 = fma(beta, C0[(i0 + 1) * ldc0 + j0 + 1], c11);
    }
    C[i0 * ldc + j0] = c00;
    if (j1) C[i0 * ldc + j0 + 1] = c01;
    if (i1) C[(i0 + 1) * ldc + j0] = c10;
    if (i1 && j1) C[(i0 + 1) * ldc + j0 + 1] = c11;
  }
}
__device__ __forceinline__ void cta_mm(double *C, int ldc, const double *C0, int ldc0,
                                       double beta, double alpha, const double *A, int ar,
                                       int ac, const double *B, int br, int bc, int M, int N,
                                       int Kd) {
  cta_mm(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, threadIdx.x,
         blockDim.x);
}

// ---------------------------------------------------------------------------
// Same product on the FP64 tensor cores: mma.sync.m8n8k4.f64 (SASS: DMMA).
// One warp owns 8x8 output tiles; per k-step of 4 every lane loads ONE element
// of A and ONE of B (fragment layout of the PTX ISA: A row = lane/4, col =
// lane%4; B row = lane%4, col = lane/4; C row = lane/4, cols = 2*(lane%4)+{0,1}),
// i.e. 2 shared-memory loads per 256 multiply-adds instead of 8 loads per 128
// with the 2x2 FMA tiles.  ncu showed the FMA version bound by shared-memory
// wavefronts (74 % LSU) with the FP64 pipe at 14 %, which is the case where the
// tensor-core path pays even for 20 x 20 blocks (padded to 24): see
// profiles/r01_ncu_full_factor_v2.md.  Used by the <NX,NU>-templated kernels;
// sizes/strides are compile-time constants there and the k-loop fully unrolls.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// one 8x8 tile pair (tiles ta, tb; tb < 0: none) of the product, all indices
// compile-time after unrolling: interior tiles carry no bounds predicates
__device__ __forceinline__ void mm_tc_tile_pair(double *C, int ldc, const double *C0, int ldc0,
                                                double beta, double alpha, const double *A,
                                                int ar, int ac, const double *B, int br, int bc,
                                                int M, int N, int Kd, int TJ, int ta, int tb,
                                                int g, int t) {
  const bool has2 = tb >= 0;
  const int i0 = (ta / TJ) << 3, j0 = (ta % TJ) << 3;
  const int i1 = has2 ? (tb / TJ) << 3 : 0, j1 = has2 ? (tb % TJ) << 3 : 0;
  // bounds checks only where the tile crosses the matrix edge (compile-time test)
  const bool va0 = (i0 + 8 <= M) || (i0 + g < M), vb0 = (j0 + 8 <= N) || (j0 + g < N);
  const bool va1 = has2 && ((i1 + 8 <= M) || (i1 + g < M));
  const bool vb1 = has2 && ((j1 + 8 <= N) || (j1 + g < N));
  const double *a0p = A + (i0 + g) * ar + t * ac, *b0p = B + (j0 + g) * bc + t * br;
  const double *a1p = A + (i1 + g) * ar + t * ac, *b1p = B + (j1 + g) * bc + t * br;
  double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < Kd; k0 += 4) {
    const bool vk = (k0 + 4 <= Kd) || (k0 + t < Kd);
    const double a0 = (va0 && vk) ? a0p[k0 * ac] : 0.0;
    const double b0 = (vb0 && vk) ? b0p[k0 * br] : 0.0;
    dmma_m8n8k4(c00, c01, a0, b0);
    if (has2) {
      const double a1 = (va1 && vk) ? a1p[k0 * ac] : 0.0;
      const double b1 = (vb1 && vk) ? b1p[k0 * br] : 0.0;
      dmma_m8n8k4(c10, c11, a1, b1);
    }
  }
  {
    const int ic = i0 + g, jc = j0 + 2 * t;
    const bool vr = (i0 + 8 <= M) || (ic < M);
    const bool v0 = vr && ((j0 + 8 <= N) || (jc < N));
    const bool v1 = vr && ((j0 + 8 <= N) || (jc + 1 < N));
    if (v0) {
      double r = alpha * c00;
      if (C0) r = fma(beta, C0[ic * ldc0 + jc], r);
      C[ic * ldc + jc] = r;
    }
    if (v1) {
      double r = alpha * c01;
      if (C0) r = fma(beta, C0[ic * ldc0 + jc + 1], r);
      C[ic * ldc + jc + 1] = r;
    }
  }
  if (has2) {
    const int ic = i1 + g, jc = j1 + 2 * t;
    const bool vr = (i1 + 8 <= M) || (ic < M);
    const bool v0 = vr && ((j1 + 8 <= N) || (jc < N));
    const bool v1 = vr && ((j1 + 8 <= N) || (jc + 1 < N));
    if (v0) {
      double r = alpha * c10;
      if (C0) r = fma(beta, C0[ic * ldc0 + jc], r);
      C[ic * ldc + jc] = r;
    }
    if (v1) {
      double r = alpha * c11;
      if (C0) r = fma(beta, C0[ic * ldc0 + jc + 1], r);
      C[ic * ldc + jc + 1] = r;
    }
  }
}

// R x S block of 8x8 tiles (tile rows ti0.., tile columns tj0..) by one warp:
// per k-step R fragments of A and S of B feed R*S DMMAs, so a 1x2 block loads 3
// fragments for 2 DMMAs and a 3x1 block 4 for 3 (a lone tile: 2 for 1) -- the
// segment kernels are bound by shared-memory wavefronts, most of them these
// fragment loads (profiles/r01_ncu_full_k1k3_v4.md).  The C tile is read and
// written as double2 when the row strides are even (conflict-free for strides
// = 8 mod 16 doubles, half the instructions otherwise).
template <int R, int S>
__device__ __forceinline__ void mm_tc_block(double *C, int ldc, const double *C0, int ldc0,
                                            double beta, double alpha, const double *A, int ar,
                                            int ac, const double *B, int br, int bc, int M,
                                            int N, int Kd, int ti0, int tj0, int nr, int g,
                                            int t, int ns = S) {
  double acc[R][S][2];
  bool va[R], vb[S];
  const double *ap[R], *bp[S];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i0 = (ti0 + r) << 3;
    va[r] = r < nr && ((i0 + 8 <= M) || (i0 + g < M));
    ap[r] = A + (i0 + g) * ar + t * ac;
#pragma unroll
    for (int s = 0; s < S; s++) acc[r][s][0] = acc[r][s][1] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int j0 = (tj0 + s) << 3;
    vb[s] = s < ns && ((j0 + 8 <= N) || (j0 + g < N));
    bp[s] = B + (j0 + g) * bc + t * br;
  }
#pragma unroll
  for (int k0 = 0; k0 < Kd; k0 += 4) {
    const bool vk = (k0 + 4 <= Kd) || (k0 + t < Kd);
    double af[R], bf[S];
#pragma unroll
    for (int r = 0; r < R; r++) af[r] = (va[r] && vk) ? ap[r][k0 * ac] : 0.0;
#pragma unroll
    for (int s = 0; s < S; s++) bf[s] = (vb[s] && vk) ? bp[s][k0 * br] : 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (r < nr) {
#pragma unroll
        for (int s = 0; s < S; s++)
          if (s < ns) dmma_m8n8k4(acc[r][s][0], acc[r][s][1], af[r], bf[s]);
      }
    }
  }
  const bool vec = (ldc % 2 == 0) && (N % 2 == 0) && (C0 == nullptr || ldc0 % 2 == 0);
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (r >= nr) continue;
    const int i0 = (ti0 + r) << 3, ic = i0 + g;
    const bool vr = (i0 + 8 <= M) || (ic < M);
#pragma unroll
    for (int s = 0; s < S; s++) {
      if (s >= ns) continue;
      const int j0 = (tj0 + s) << 3, jc = j0 + 2 * t;
      const bool v0 = vr && ((j0 + 8 <= N) || (jc < N));
      const bool v1 = vr && ((j0 + 8 <= N) || (jc + 1 < N));
      double r0 = alpha * acc[r][s][0], r1 = alpha * acc[r][s][1];
      if (vec) {
        if (v0) {
          if (C0) {
            const double2 c0 = *reinterpret_cast<const double2 *>(C0 + ic * ldc0 + jc);
            r0 = fma(beta, c0.x, r0);
            r1 = fma(beta, c0.y, r1);
          }
          *reinterpret_cast<double2 *>(C + ic * ldc + jc) = make_double2(r0, r1);
        }
      } else {
        if (v0) {
          if (C0) r0 = fma(beta, C0[ic * ldc0 + jc], r0);
          C[ic * ldc + jc] = r0;
        }
        if (v1) {
          if (C0) r1 = fma(beta, C0[ic * ldc0 + jc + 1], r1);
          C[ic * ldc + jc + 1] = r1;
        }
      }
    }
  }
}

// Whole-CTA product on the tensor cores.  Work units: per tile row, pairs of
// adjacent tiles (1x2 blocks, shared A fragment); if the tile-column count is
// odd, the last column in vertical strips of up to three tiles (shared B
// fragment).  Units are dealt round-robin to the NW warps starting at warp
// `wofs` (lets back-to-back products without a barrier between them spread over
// different warps).  Tile indices are compile-time after unrolling; the warp
// test is uniform.
template <int NW>
__device__ __forceinline__ void cta_mm_tc(double *C, int ldc, const double *C0, int ldc0,
                                          double beta, double alpha, const double *A, int ar,
                                          int ac, const double *B, int br, int bc, int M, int N,
                                          int Kd, int wofs = 0, int warp_id = -1,
                                          bool lower = false) {
  // warp_id >= 0: the product is shared by a subset of NW warps numbered 0..NW-1
  const int lane = threadIdx.x & 31, warp = warp_id >= 0 ? warp_id : (int)(threadIdx.x >> 5);
  const int g = lane >> 2, t = lane & 3;
  const int TI = (M + 7) >> 3, TJ = (N + 7) >> 3;
  if (lower) {
    // symmetric result (M == N): only the tiles on and below the diagonal; the
    // caller mirrors them (cta_symmetrize_tc, mirror mode).  Per tile row:
    // 1x2 blocks and, for an odd count, the diagonal tile alone.
    int u = 0;
#pragma unroll
    for (int ti = 0; ti < TI; ti++) {
#pragma unroll
      for (int tj = 0; tj <= ti; tj += 2, u++) {
        if (NW == 1 || warp == (u + wofs) % NW)
          mm_tc_block<1, 2>(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, ti, tj,
                            1, g, t, tj + 1 <= ti ? 2 : 1);
      }
    }
    return;
  }
  if constexpr (NW == 1) {
    // one warp owns the whole product: 3x3 blocks, 6 fragment loads per 9 DMMAs
#pragma unroll
    for (int bi = 0; bi < TI; bi += 3) {
#pragma unroll
      for (int bj = 0; bj < TJ; bj += 3)
        mm_tc_block<3, 3>(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, bi, bj,
                          TI - bi < 3 ? TI - bi : 3, g, t, TJ - bj < 3 ? TJ - bj : 3);
    }
    return;
  }
  const int npr = TJ >> 1;                       // 1x2 blocks per tile row
  const int nrow_units = TI * npr;
  const int ncol_units = (TJ & 1) ? (TI + 2) / 3 : 0;
#pragma unroll
  for (int u = 0; u < nrow_units; u++) {
    if (warp == (u + wofs) % NW)
      mm_tc_block<1, 2>(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, u / npr,
                        2 * (u % npr), 1, g, t);
  }
#pragma unroll
  for (int c = 0; c < ncol_units; c++) {
    if (warp == (nrow_units + c + wofs) % NW) {
      const int nr = TI - 3 * c < 3 ? TI - 3 * c : 3;
      mm_tc_block<3, 1>(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, 3 * c,
                        TJ - 1, nr, g, t);
    }
  }
}

// ---------------------------------------------------------------------------
// Large-block product (stage blocks that live in global memory, nx > 64):
//   C (M x N) = beta * C0 + alpha * A * B,  same strided operands as cta_mm.
// CTA-cooperative tensor-core GEMM.  The output is cut into CTA tiles of
// (wmb * 40) x (wnb * 24) elements (5 x 3 = 15 warp blocks of 5 x 3 DMMA tiles each:
// 200 x 72; 2 x 7 blocks for products with few rows); the k-direction is streamed
// in chunks of LQ_BIG_KC through a ring of LQ_BIG_NS shared-memory stages that ALL
// warps read: the A panel of a chunk feeds every warp column and the B panel every
// warp row, so a byte staged from L2 is used by 3 .. 7 warps (round 2's first
// version staged private panels per warp and spent 97 % of its instructions in the
// staging loop; profiles/r02_ncu_full_c4.md).
// Staging: one bulk async copy (cp.async.bulk, SASS UBLKCP) per panel row -- 128 B
// along k when k is the operand's contiguous direction, one copy per k otherwise --
// completing on the stage's mbarrier, issued LQ_BIG_NS chunks ahead; one named
// barrier per chunk recycles the stage.  Operands whose strides or base break the
// 16-byte rules of the bulk copy take 8-byte cp.async copies (zero-filled) with
// the same ring, two stages deep.
// Panel layouts (conflict-free 64-bit fragment reads, see lq_pad4):
//   [row][k], row stride LQ_BIG_KC + 4 = 20      (k contiguous in memory)
//   [k][row], row stride rows + 4 = 4|12 (mod 16) (rows contiguous in memory)
// stg: LQ_BIG_STG doubles of shared memory, 16-byte aligned.  Called by `nwarps`
// warps numbered warp_log = 0 .. nwarps-1 (the whole CTA or a subset, e.g. all but
// the warp that factors Guu); they synchronise on named barrier 1.
// ---------------------------------------------------------------------------
#define LQ_BIG_KC 16
#define LQ_BIG_NS 4
#define LQ_BIG_LDS (LQ_BIG_KC + 4)
#define LQ_BIG_RT 5                                    // DMMA tiles per warp block: rows
#define LQ_BIG_CT 3                                    //                            columns
#define LQ_BIG_ROWS (5 * 8 * LQ_BIG_RT + 3 * 8 * LQ_BIG_CT)  // panel rows of a stage: 200 + 72
#define LQ_BIG_STAGE (LQ_BIG_ROWS * LQ_BIG_LDS)
#define LQ_BIG_STG (LQ_BIG_NS * LQ_BIG_STAGE)          // doubles of staging per CTA (174 KB)
#define LQ_BIG_NT 512                                  // threads per CTA of the large-block kernels
#define LQ_BIG_SMEM_DOUBLES 27000                      // what a large-block kernel may use (216 KB)

// Large-block segment kernels: shared memory = [U | LDL^T factor of Guu], U = the
// GEMM staging ring, which between the products doubles as the right-hand sides of
// the substitutions (one column per thread, `big_ycols` columns at a time).
__host__ __device__ inline int big_ycols(int nx, int nu, int nthr) {
  const long avail = ((long)LQ_BIG_SMEM_DOUBLES - (long)nu * (nu + 1)) / (nu > 0 ? nu : 1);
  long yc = nthr < 2 * nx ? nthr : 2 * nx;
  if (yc > avail) yc = avail;
  return (int)(yc & ~1L);
}
__host__ __device__ inline size_t big_seg_union_doubles(int nx, int nu, int nthr) {
  const size_t y = (size_t)nu * big_ycols(nx, nu, nthr);
  return y > (size_t)LQ_BIG_STG ? y : (size_t)LQ_BIG_STG;
}
__host__ __device__ inline size_t big_seg_smem_doubles(int nx, int nu, int nthr) {
  return big_seg_union_doubles(nx, nu, nthr) + (size_t)nu * (nu + 1) + 2;
}
__host__ __device__ inline bool big_ldlt_fits(int nx, int nu, int nthr) {
  return (size_t)nu * (nu + 1) + LQ_BIG_STG + 2 <= (size_t)LQ_BIG_SMEM_DOUBLES &&
         big_ycols(nx, nu, nthr) >= 32;
}

__device__ __forceinline__ void cp_async_f64(double *dst_smem, const double *src, bool valid) {
  const int sz = valid ? 8 : 0;  // src-size 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src),
               "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void big_bar(int nt) { asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory"); }

// One operand panel of the chunk k0 .. of X(i, l) = X[i * xr + l * xc], rows r0 ..,
// by the nt threads lt = 0 .. nt-1 (operands without 16-byte alignment).
//   TR = false: panel[r][k] (stride LQ_BIG_LDS);  TR = true: panel[k][r] (stride ldt)
// generic: 8-byte cp.async, zero-filled outside M x Kd; rows = panel rows to fill
template <bool TR>
__device__ __forceinline__ void big_fill_gen(double *panel, int ldt, const double *X, int xr,
                                             int xc, int r0, int M, int rows, int k0, int Kd,
                                             int lt, int nt) {
  constexpr int KC = LQ_BIG_KC;
  if constexpr (!TR) {
    for (int e = lt; e < rows * KC; e += nt) {
      const int r = e / KC, k = e % KC;
      const bool ok = r0 + r < M && k0 + k < Kd;
      cp_async_f64(panel + r * LQ_BIG_LDS + k, ok ? X + (size_t)(r0 + r) * xr + (size_t)(k0 + k) * xc : X, ok);
    }
  } else {
    for (int e = lt; e < rows * KC; e += nt) {
      const int k = e / rows, r = e - k * rows;
      const bool ok = r0 + r < M && k0 + k < Kd;
      cp_async_f64(panel + k * ldt + r, ok ? X + (size_t)(r0 + r) * xr + (size_t)(k0 + k) * xc : X, ok);
    }
  }
}

__device__ __forceinline__ void mbar_add_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// (the core is not inlined: ~40 call sites in the large-block kernels; four copies -- one per
//  pair of operand orientations, so that the fragment addressing keeps constant strides --
//  hold the build of the library at a third of the time of full inlining, and a call costs
//  nothing next to a 10^5-cycle product)
template <bool TA, bool TB>
__device__ __noinline__ void cta_mm_big_core(double *stg, double *C, int crs, int ccs, const double *C0,
                                             int c0rs, int c0cs, double beta, double alpha,
                                             const double *A, int ar, int ac, const double *B, int br,
                                             int bc, int M, int N, int Kd, int warp_log, int nwarps) {
  constexpr int KC = LQ_BIG_KC, NS = LQ_BIG_NS, LDS = LQ_BIG_LDS, RT = LQ_BIG_RT, CT = LQ_BIG_CT;
  constexpr int WR = 8 * RT, WC = 8 * CT;  // rows / columns of a warp block
  __shared__ __align__(8) uint64_t bars[2 * NS];  // full[NS] (bytes landed), empty[NS] (warps done)
  uint64_t *full = bars, *empty = bars + NS;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int lt = warp_log * 32 + lane, nt = nwarps * 32;
  // warp grid of a CTA tile
  int wnb = nwarps >= 15 ? 3 : (nwarps >= 8 ? 2 : 1);
  int wmb = nwarps / wnb < 5 ? nwarps / wnb : 5;
  if (M <= 2 * WR && nwarps >= 14) { wmb = 2; wnb = 7; }
  const int BM = wmb * WR, BN = wnb * WC, ncw = wmb * wnb;
  const int wr = warp_log / wnb, wc = warp_log - wr * wnb;
  const bool cw = warp_log < ncw;  // this warp computes
  // operand orientation -> panel layout; bulk copies where the 16-byte rules hold
  constexpr bool ta = TA, tb = TB;
  // (a subset of the CTA's warps takes the plain path as well)
#ifndef LQ_BIG_SUBFAST
#define LQ_BIG_SUBFAST 0
#endif
  const bool fast =
      (LQ_BIG_SUBFAST || nt == (int)blockDim.x) &&
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && (Kd & 1) == 0 &&
      (ta ? ((ac & 1) == 0 && (M & 1) == 0) : (ac == 1 && (ar & 1) == 0)) &&
      (tb ? ((br & 1) == 0 && (N & 1) == 0) : (br == 1 && (bc & 1) == 0));
  const bool solo = fast && ta && tb && ncw < nwarps;  // one spare warp stages everything
  const int ldtA = BM + 4, ldtB = BN + 4;
  const int nchunk = (Kd + KC - 1) / KC;
  const int ntj_t = (N + BN - 1) / BN;
  const int total = ((M + BM - 1) / BM) * ntj_t * nchunk;
  // fragment addressing: element (row, k) of a panel at row * rs + k * ks
  const int a_rs = ta ? 1 : LDS, a_ks = ta ? ldtA : 1;
  const int b_rs = tb ? 1 : LDS, b_ks = tb ? ldtB : 1;
  const int a_off = (wr * WR + g) * a_rs + t * a_ks, b_off = BM * LDS + (wc * WC + g) * b_rs + t * b_ks;
  // The operands were written by this CTA a moment ago with ordinary stores (generic
  // proxy); the bulk copies below read them through the async proxy, and overwrite a
  // staging area that held right-hand sides / scratch.  Every thread orders its own
  // writes -- global and shared -- before the barrier that precedes the first copy.
#ifndef LQ_BIG_FENCE
#define LQ_BIG_FENCE 2
#endif
#if LQ_BIG_FENCE == 0
  fence_proxy_async();
#elif LQ_BIG_FENCE == 1
  __threadfence();
  fence_proxy_async();
#elif LQ_BIG_FENCE == 2
  __threadfence();
  asm volatile("fence.proxy.async.global;" ::: "memory");
  fence_proxy_async();
#else
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
  if (fast && lt == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
      mbar_init(&full[s], solo ? 1 : nwarps);
      mbar_init(&empty[s], ncw);
    }
    mbar_fence_init();
  }
  big_bar(nt);
  if (fast) {
    // ---- bulk-copy ring.  The (tile, chunk) pairs form ONE stream q = 0 .. total-1
    // (slot q % NS, phase (q / NS) & 1).  Every warp stages its share of chunk q + D
    // (D = NS - 2) at the top of its iteration q: it waits until the ncw computing
    // warps have released that slot (`empty`; they did so two chunks ago, so the wait
    // is normally over already), issues at most one bulk copy per lane and arrives on
    // the slot's `full` barrier with the byte count of its copies.  The computing
    // warps then only wait for data, not for each other: no CTA barrier inside the
    // product, and the first chunks of the next CTA tile are in flight while the last
    // ones of the current tile are consumed.  (One staging warp alone cannot issue
    // the ~270 row copies of a k-contiguous operand fast enough: 47 cycles per copy.)
    // When both panels arrive as a few long copies (rows contiguous in memory: 2 x 16
    // copies per chunk) and the warp grid leaves a warp without a block, that warp alone
    // stages, a full ring ahead, and the computing warps do not touch the copies at all.
    constexpr int D = NS - 2;
    auto issue = [&](int q, int lt, int nt) {
      if (q >= total) return;
      const int s = q % NS;
      if (q >= NS) {
        mbar_wait(&empty[s], ((q / NS) - 1) & 1);
        __syncwarp();
      }
      const int tile = q / nchunk, c = q - tile * nchunk;
      const int i0 = (tile / ntj_t) * BM, j0 = (tile % ntj_t) * BN;
      const int mt = M - i0 < BM ? M - i0 : BM, nc = N - j0 < BN ? N - j0 : BN;
      const int k0 = c * KC, kc = Kd - k0 < KC ? Kd - k0 : KC;
      double *pa = stg + (size_t)s * LQ_BIG_STAGE, *pb = pa + BM * LDS;
      // copies lt, lt + nt, .. of the chunk (A copies first, then B copies): normally
      // at most one per thread
      const int na = ta ? kc : mt, nb = tb ? kc : nc;
      auto copy_of = [&](int e, const double *&src, double *&dst) -> uint32_t {
        if (e < na) {
          if (ta) { src = A + (size_t)(k0 + e) * ac + i0; dst = pa + e * ldtA; return (uint32_t)mt * 8; }
          src = A + (size_t)(i0 + e) * ar + k0; dst = pa + e * LDS; return (uint32_t)kc * 8;
        }
        e -= na;
        if (tb) { src = B + (size_t)(k0 + e) * br + j0; dst = pb + e * ldtB; return (uint32_t)nc * 8; }
        src = B + (size_t)(j0 + e) * bc + k0; dst = pb + e * LDS; return (uint32_t)kc * 8;
      };
      if (kc & 3) {  // zero-filled k-tail (plain stores, published by the arrive below)
        const int kc4 = (kc + 3) & ~3;
        for (int e = lt; e < (kc4 - kc) * mt; e += nt) {
          const int k = kc + e / mt, r = e % mt;
          pa[ta ? k * ldtA + r : r * LDS + k] = 0.0;
        }
        for (int e = lt; e < (kc4 - kc) * nc; e += nt) {
          const int k = kc + e / nc, r = e % nc;
          pb[tb ? k * ldtB + r : r * LDS + k] = 0.0;
        }
        __syncwarp();
      }
      // every thread announces the bytes of its own copies, then lane 0 arrives for the
      // warp.  (Uniform trip count, predicated body, and the warp reconverges before it
      // goes on: what follows -- mma.sync, the next wait -- needs all 32 lanes together.)
      for (int e0 = 0; e0 < na + nb; e0 += nt) {
        const int e = e0 + lt;
        if (e < na + nb) {
          const double *src;
          double *dst;
          const uint32_t b = copy_of(e, src, dst);
          mbar_add_tx(&full[s], b);
          tma_load_1d(dst, src, b, &full[s]);
        }
        __syncwarp();
      }
      if (lane == 0) mbar_arrive(&full[s]);
      __syncwarp();
    };
    if (solo) {
      if (warp_log == ncw)
        for (int q = 0; q < total; q++) issue(q, lane, 32);
    } else {
      for (int q = 0; q < D; q++) issue(q, lt, nt);
      if (!cw)
        for (int q = 0; q < total; q++) issue(q + D, lt, nt);
    }
    if (cw) {
      int q = 0;
      for (int i0 = 0; i0 < M; i0 += BM) {
        for (int j0 = 0; j0 < N; j0 += BN) {
          const int mt = M - i0 < BM ? M - i0 : BM, nc = N - j0 < BN ? N - j0 : BN;
          int nti = (mt - wr * WR + 7) >> 3, ntj = (nc - wc * WC + 7) >> 3;
          nti = nti < 0 ? 0 : (nti > RT ? RT : nti);
          ntj = ntj < 0 ? 0 : (ntj > CT ? CT : ntj);
          if (nti == 0 || ntj == 0) nti = ntj = 0;
          double acc[RT][CT][2];
#pragma unroll
          for (int r = 0; r < RT; r++)
#pragma unroll
            for (int qq = 0; qq < CT; qq++) acc[r][qq][0] = acc[r][qq][1] = 0.0;
          for (int c = 0; c < nchunk; c++, q++) {
            if (!solo) issue(q + D, lt, nt);
            const int s = q % NS;
            mbar_wait(&full[s], (q / NS) & 1);
            __syncwarp();
            const double *ap = stg + (size_t)s * LQ_BIG_STAGE + a_off;
            const double *bp = stg + (size_t)s * LQ_BIG_STAGE + b_off;
            const int kc = Kd - c * KC < KC ? Kd - c * KC : KC;
            const int kc4 = (kc + 3) & ~3;
            if (nti == RT && ntj == CT) {
#pragma unroll 4
              for (int kk = 0; kk < kc4; kk += 4) {
                double af[RT], bf[CT];
#pragma unroll
                for (int r = 0; r < RT; r++) af[r] = ap[r * 8 * a_rs + kk * a_ks];
#pragma unroll
                for (int qq = 0; qq < CT; qq++) bf[qq] = bp[qq * 8 * b_rs + kk * b_ks];
#pragma unroll
                for (int r = 0; r < RT; r++)
#pragma unroll
                  for (int qq = 0; qq < CT; qq++) dmma_m8n8k4(acc[r][qq][0], acc[r][qq][1], af[r], bf[qq]);
              }
            } else if (nti > 0) {
              for (int kk = 0; kk < kc4; kk += 4) {
                double af[RT], bf[CT];
#pragma unroll
                for (int r = 0; r < RT; r++) af[r] = ap[r * 8 * a_rs + kk * a_ks];
#pragma unroll
                for (int qq = 0; qq < CT; qq++) bf[qq] = bp[qq * 8 * b_rs + kk * b_ks];
#pragma unroll
                for (int r = 0; r < RT; r++) {
                  if (r < nti) {
#pragma unroll
                    for (int qq = 0; qq < CT; qq++)
                      if (qq < ntj) dmma_m8n8k4(acc[r][qq][0], acc[r][qq][1], af[r], bf[qq]);
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with slot s
          }
          // (C0 may alias C: all loads of the tile first, then all stores -- interleaved,
          //  every load would wait for the store before it, one L2 round trip each; the
          //  two adjacent columns of a lane move as one 16-byte access where they can)
          const bool vec = ccs == 1 && (crs & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                           (N & 1) == 0 &&
                           (!C0 || (c0cs == 1 && (c0rs & 1) == 0 && (reinterpret_cast<uintptr_t>(C0) & 15) == 0));
          if (C0) {
            // (no branches around the loads: every load is issued from a valid address,
            //  so that several of them are in flight at once -- with a branch per element
            //  each one waits out its own L2 round trip, 15 in a row per tile)
#pragma unroll
            for (int r = 0; r < RT; r++) {
              const int ic = i0 + wr * WR + r * 8 + g;
              double c0a[CT], c0b[CT];
#pragma unroll
              for (int qq = 0; qq < CT; qq++) {
                const int jc = j0 + wc * WC + qq * 8 + 2 * t;
                const bool ok = r < nti && qq < ntj && ic < M && jc < N;
                if (vec) {
                  const double2 c2 = *reinterpret_cast<const double2 *>(ok ? C0 + (size_t)ic * c0rs + jc : C0);
                  c0a[qq] = c2.x;
                  c0b[qq] = c2.y;
                } else {
                  const bool ok1 = ok && jc + 1 < N;
                  c0a[qq] = *(ok ? C0 + (size_t)ic * c0rs + (size_t)jc * c0cs : C0);
                  c0b[qq] = *(ok1 ? C0 + (size_t)ic * c0rs + (size_t)(jc + 1) * c0cs : C0);
                }
              }
#pragma unroll
              for (int qq = 0; qq < CT; qq++) {
                acc[r][qq][0] = fma(beta, c0a[qq], alpha * acc[r][qq][0]);
                acc[r][qq][1] = fma(beta, c0b[qq], alpha * acc[r][qq][1]);
              }
            }
          }
          const double asc = C0 ? 1.0 : alpha;
#pragma unroll
          for (int r = 0; r < RT; r++) {
            const int ic = i0 + wr * WR + r * 8 + g;
            if (r >= nti || ic >= M) continue;
#pragma unroll
            for (int qq = 0; qq < CT; qq++) {
              const int jc = j0 + wc * WC + qq * 8 + 2 * t;
              if (qq >= ntj) continue;
              if (vec) {
                if (jc < N)
                  *reinterpret_cast<double2 *>(C + (size_t)ic * crs + jc) =
                      make_double2(asc * acc[r][qq][0], asc * acc[r][qq][1]);
              } else {
                if (jc < N) C[(size_t)ic * crs + (size_t)jc * ccs] = asc * acc[r][qq][0];
                if (jc + 1 < N) C[(size_t)ic * crs + (size_t)(jc + 1) * ccs] = asc * acc[r][qq][1];
              }
            }
          }
        }
      }
    }
    big_bar(nt);  // the ring and its barriers are free again
    // (the next product arms the barriers with other counts: re-initialising a live
    //  mbarrier is undefined, it has to be invalidated first)
    if (lt == 0) {
#pragma unroll
      for (int s = 0; s < 2 * NS; s++) mbar_inval(&bars[s]);
    }
    return;
  }
  // ---- operands that break the 16-byte rules: 8-byte cp.async copies by all threads,
  // two slots, one CTA-wide barrier per chunk
  for (int i0 = 0; i0 < M; i0 += BM) {
    for (int j0 = 0; j0 < N; j0 += BN) {
      const int mt = M - i0 < BM ? M - i0 : BM, nc = N - j0 < BN ? N - j0 : BN;
      int nti = 0, ntj = 0;
      if (cw) {
        nti = (mt - wr * WR + 7) >> 3;
        nti = nti < 0 ? 0 : (nti > RT ? RT : nti);
        ntj = (nc - wc * WC + 7) >> 3;
        ntj = ntj < 0 ? 0 : (ntj > CT ? CT : ntj);
        if (nti == 0 || ntj == 0) nti = ntj = 0;
      }
      double acc[RT][CT][2];
#pragma unroll
      for (int r = 0; r < RT; r++)
#pragma unroll
        for (int qq = 0; qq < CT; qq++) acc[r][qq][0] = acc[r][qq][1] = 0.0;
      auto fill = [&](int c, int s) {
        double *pa = stg + (size_t)s * LQ_BIG_STAGE, *pb = pa + BM * LDS;
        const int ra = (mt + 7) & ~7, rb = (nc + 7) & ~7;
        if (ta) big_fill_gen<true>(pa, ldtA, A, ar, ac, i0, M, ra, c * KC, Kd, lt, nt);
        else big_fill_gen<false>(pa, ldtA, A, ar, ac, i0, M, ra, c * KC, Kd, lt, nt);
        if (tb) big_fill_gen<true>(pb, ldtB, B, bc, br, j0, N, rb, c * KC, Kd, lt, nt);
        else big_fill_gen<false>(pb, ldtB, B, bc, br, j0, N, rb, c * KC, Kd, lt, nt);
        cp_async_commit();
      };
      fill(0, 0);
      for (int c = 0; c < nchunk; c++) {
        cp_async_wait_all();
        big_bar(nt);  // chunk c has landed for every thread; chunk c-1 is consumed
        if (c + 1 < nchunk) fill(c + 1, (c + 1) & 1);
        const double *ap = stg + (size_t)(c & 1) * LQ_BIG_STAGE + a_off;
        const double *bp = stg + (size_t)(c & 1) * LQ_BIG_STAGE + b_off;
        if (nti > 0) {
          for (int kk = 0; kk < KC; kk += 4) {
            double af[RT], bf[CT];
#pragma unroll
            for (int r = 0; r < RT; r++) af[r] = ap[r * 8 * a_rs + kk * a_ks];
#pragma unroll
            for (int qq = 0; qq < CT; qq++) bf[qq] = bp[qq * 8 * b_rs + kk * b_ks];
#pragma unroll
            for (int r = 0; r < RT; r++) {
              if (r < nti) {
#pragma unroll
                for (int qq = 0; qq < CT; qq++)
                  if (qq < ntj) dmma_m8n8k4(acc[r][qq][0], acc[r][qq][1], af[r], bf[qq]);
              }
            }
          }
        }
      }
      big_bar(nt);  // before the next tile refills slot 0
#pragma unroll
      for (int r = 0; r < RT; r++) {
        const int ic = i0 + wr * WR + r * 8 + g;
        if (r >= nti || ic >= M) continue;
#pragma unroll
        for (int qq = 0; qq < CT; qq++) {
          const int jc = j0 + wc * WC + qq * 8 + 2 * t;
          if (qq >= ntj) continue;
          double r0 = alpha * acc[r][qq][0], r1 = alpha * acc[r][qq][1];
          if (jc < N) {
            if (C0) r0 = fma(beta, C0[(size_t)ic * c0rs + (size_t)jc * c0cs], r0);
            C[(size_t)ic * crs + (size_t)jc * ccs] = r0;
          }
          if (jc + 1 < N) {
            if (C0) r1 = fma(beta, C0[(size_t)ic * c0rs + (size_t)(jc + 1) * c0cs], r1);
            C[(size_t)ic * crs + (size_t)(jc + 1) * ccs] = r1;
          }
        }
      }
    }
  }
}

// C (M x N, ldc) = beta * C0 + alpha * A * B: orientation dispatch in front of the core
__device__ __forceinline__ void cta_mm_big(double *stg, double *C, int ldc, const double *C0,
                                           int ldc0, double beta, double alpha, const double *A,
                                           int ar, int ac, const double *B, int br, int bc, int M,
                                           int N, int Kd, int wofs, int warp_log, int nwarps) {
  (void)wofs;
  // a product with few rows is computed as C' = B' A' (the CTA tile is tall): the
  // operands swap roles and the result is stored through transposed strides
  int crs = ldc, ccs = 1, c0rs = ldc0, c0cs = 1;
  if (M <= 2 * 8 * LQ_BIG_RT && N > M) {
    const double *tp = A; A = B; B = tp;
    int ti = ar; ar = bc; bc = ti;
    ti = ac; ac = br; br = ti;
    ti = M; M = N; N = ti;
    crs = 1; ccs = ldc; c0rs = 1; c0cs = ldc0;
  }
  const bool ta = ar == 1 && ac != 1, tb = bc == 1 && br != 1;
  if (ta && tb)
    cta_mm_big_core<true, true>(stg, C, crs, ccs, C0, c0rs, c0cs, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, warp_log, nwarps);
  else if (ta)
    cta_mm_big_core<true, false>(stg, C, crs, ccs, C0, c0rs, c0cs, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, warp_log, nwarps);
  else if (tb)
    cta_mm_big_core<false, true>(stg, C, crs, ccs, C0, c0rs, c0cs, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, warp_log, nwarps);
  else
    cta_mm_big_core<false, false>(stg, C, crs, ccs, C0, c0rs, c0cs, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, warp_log, nwarps);
}

// compile-time choice between the tensor-core and the FMA product
// stg != nullptr (large blocks in global memory, generic kernels only): cta_mm_big.
template <bool TC, int NW = LQ_NT / 32>
__device__ __forceinline__ void cta_mmx(double *stg, double *C, int ldc, const double *C0,
                                        int ldc0, double beta, double alpha, const double *A,
                                        int ar, int ac, const double *B, int br, int bc, int M,
                                        int N, int Kd, int wofs = 0, int warp_id = -1,
                                        bool lower = false) {
  if constexpr (!TC) {
    if (stg) {
      cta_mm_big(stg, C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, wofs,
                 warp_id >= 0 ? warp_id : (int)(threadIdx.x >> 5), warp_id >= 0 ? NW : (int)(blockDim.x >> 5));
      return;
    }
  }
  if constexpr (TC) {
    cta_mm_tc<NW>(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd, wofs, warp_id,
                  lower);
  } else {
    if (warp_id >= 0)  // subset of NW warps
      cta_mm(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd,
             warp_id * 32 + (int)(threadIdx.x & 31), NW * 32);
    else
      cta_mm(C, ldc, C0, ldc0, beta, alpha, A, ar, ac, B, br, bc, M, N, Kd);
  }
}

// dst[0..n) = src[0..n) by the whole CTA (global -> global, large blocks): 16-byte
// accesses when both ends allow it, four independent loads in flight per thread
__device__ __forceinline__ void cta_copy_big(double *dst, const double *src, int n, int tid, int nthr) {
  if ((((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) && (n & 1) == 0) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    const int n2 = n >> 1;
    int i = tid;
    for (; i + 3 * nthr < n2; i += 4 * nthr) {
      const double2 a = s2[i], b = s2[i + nthr], c = s2[i + 2 * nthr], e = s2[i + 3 * nthr];
      d2[i] = a; d2[i + nthr] = b; d2[i + 2 * nthr] = c; d2[i + 3 * nthr] = e;
    }
    for (; i < n2; i += nthr) d2[i] = s2[i];
  } else {
    int i = tid;
    for (; i + 3 * nthr < n; i += 4 * nthr) {
      const double a = src[i], b = src[i + nthr], c = src[i + 2 * nthr], e = src[i + 3 * nthr];
      dst[i] = a; dst[i + nthr] = b; dst[i + 2 * nthr] = c; dst[i + 3 * nthr] = e;
    }
    for (; i < n; i += nthr) dst[i] = src[i];
  }
}

// A <- 0.5 (A + A') for a block in global memory (large blocks): 32 x 32 tile pairs
// (I, J), I <= J, one pair per warp at a time; both tiles are read along their rows
// and exchanged through a padded shared-memory tile (scr: 33 * 32 doubles per warp),
// where cta_symmetrize reads one of every two elements down a column.
__device__ __forceinline__ void cta_symmetrize_big(double *scr, double *A, int lda, int n) {
  // (at most 15 warps: 15 padded tiles fit the GEMM staging ring, LQ_BIG_STG doubles)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (int)(blockDim.x >> 5) < 15 ? (int)(blockDim.x >> 5) : 15;
  if (warp >= nw) return;
  double *tile = scr + (size_t)warp * (33 * 32);
  const int nt = (n + 31) >> 5;
  for (int pr = warp; pr < nt * (nt + 1) / 2; pr += nw) {
    // pair index -> (I, J), I <= J
    int I = 0, rem = pr;
    while (rem >= nt - I) { rem -= nt - I; I++; }
    const int J = I + rem;
    const int i0 = I << 5, j0 = J << 5;
    // tile (J, I) into shared memory, transposed access later
    for (int r = 0; r < 32; r++) {
      const int gi = j0 + r, gj = i0 + lane;
      tile[r * 33 + lane] = (gi < n && gj < n) ? A[(size_t)gi * lda + gj] : 0.0;
    }
    __syncwarp();
    for (int r = 0; r < 32; r++) {
      const int gi = i0 + r, gj = j0 + lane;
      if (gi < n && gj < n) {
        const double v = 0.5 * (A[(size_t)gi * lda + gj] + tile[lane * 33 + r]);
        A[(size_t)gi * lda + gj] = v;
        tile[lane * 33 + r] = v;
      }
    }
    __syncwarp();
    if (I != J) {
      for (int r = 0; r < 32; r++) {
        const int gi = j0 + r, gj = i0 + lane;
        if (gi < n && gj < n) A[(size_t)gi * lda + gj] = tile[r * 33 + lane];
      }
    }
    __syncwarp();
  }
}

// A <- 0.5 (A + A')  (n x n, lda); one thread per (i<j) pair
__device__ __forceinline__ void cta_symmetrize(double *A, int lda, int n) {
  const int total = n * n;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    if (i < j) {
      const double v = 0.5 * (A[i * lda + j] + A[j * lda + i]);
      A[i * lda + j] = v;
      A[j * lda + i] = v;
    }
  }
}

// A <- 0.5 (A + A') through registers, 8x8 tile pairs (ti <= tj) dealt to the NW
// warps: row-wise double2 reads of both tiles, the transposed partner of every
// element fetched with shuffles, row-wise double2 writes -- no column-strided
// shared-memory access (the element-wise version above spends ~1.5 k cycles per
// call on 4-way bank conflicts at lda = 20).  lda and n even, A 16-byte aligned.
// mirror = true: the tiles above the diagonal are stale (product computed with
// `lower`): they are overwritten with the transposed lower tiles, only the
// diagonal tiles are averaged.
template <int NW>
__device__ __forceinline__ void cta_symmetrize_tc(double *A, int lda, int n, bool mirror = false) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int TN = (n + 7) >> 3;
  int u = 0;
#pragma unroll
  for (int ti = 0; ti < TN; ti++) {
#pragma unroll
    for (int tj = ti; tj < TN; tj++, u++) {
      if (warp != u % NW) continue;
      const int i0 = ti << 3, j0 = tj << 3;
      const bool va = (i0 + g < n) && (j0 + 2 * t < n);  // a: rows of tile (ti, tj)
      const bool vb = (j0 + g < n) && (i0 + 2 * t < n);  // b: rows of tile (tj, ti)
      double2 a = make_double2(0.0, 0.0), b2 = make_double2(0.0, 0.0);
      if (va) a = *reinterpret_cast<const double2 *>(A + (i0 + g) * lda + j0 + 2 * t);
      if (vb) b2 = *reinterpret_cast<const double2 *>(A + (j0 + g) * lda + i0 + 2 * t);
      // partner of a.{x,y} = element (j0 + 2t + e, i0 + g): array b of lane
      // (g' = 2t + e, t' = g / 2), component g % 2; and symmetrically for b
      const int s0 = (2 * t) * 4 + (g >> 1), s1 = (2 * t + 1) * 4 + (g >> 1);
      const bool odd = g & 1;
      const double bx0 = __shfl_sync(0xffffffffu, b2.x, s0), by0 = __shfl_sync(0xffffffffu, b2.y, s0);
      const double bx1 = __shfl_sync(0xffffffffu, b2.x, s1), by1 = __shfl_sync(0xffffffffu, b2.y, s1);
      const double ax0 = __shfl_sync(0xffffffffu, a.x, s0), ay0 = __shfl_sync(0xffffffffu, a.y, s0);
      const double ax1 = __shfl_sync(0xffffffffu, a.x, s1), ay1 = __shfl_sync(0xffffffffu, a.y, s1);
      const bool copy = mirror && ti != tj;
      const double2 an = copy ? make_double2(odd ? by0 : bx0, odd ? by1 : bx1)
                              : make_double2(0.5 * (a.x + (odd ? by0 : bx0)),
                                             0.5 * (a.y + (odd ? by1 : bx1)));
      const double2 bn = make_double2(0.5 * (b2.x + (odd ? ay0 : ax0)), 0.5 * (b2.y + (odd ? ay1 : ax1)));
      if (va) *reinterpret_cast<double2 *>(A + (i0 + g) * lda + j0 + 2 * t) = an;
      if (ti != tj && !copy && vb) *reinterpret_cast<double2 *>(A + (j0 + g) * lda + i0 + 2 * t) = bn;
    }
  }
}

// ---------------------------------------------------------------------------
// In-place LDL^T (no pivoting) of the symmetric m x m block A (lda) by warp 0;
// only the lower triangle is read.  On exit: strictly lower part = unit L,
// diagonal = 1/D.  Returns (to all lanes of warp 0) a bit set: LQ_FLAG_SING for
// a zero / non-finite pivot, LQ_FLAG_NOTPD for a negative one (the
// factorisation continues).
// The reference uses a scaled Bunch-Kaufman here (hqp/Hqp_IpLQDOCP.C:1860-1879);
// on the convex QPs of an IP iteration Guu is positive definite and LDL^T
// without interchanges is backward stable.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_ldlt(double *A, int lda, int m) {
  const int lane = threadIdx.x & 31;
  int st = 0;
  for (int p = 0; p < m; p++) {
    const double d = A[p * lda + p];
    if (!(d > 0.0)) st |= (d == 0.0 || d != d) ? LQ_FLAG_SING : LQ_FLAG_NOTPD;
    const double inv = 1.0 / d;
    const int r = m - p - 1;  // trailing size
    // trailing update A[i][j] -= A[i][p] A[j][p] / d, p < j <= i
    for (int e = lane; e < r * r; e += 32) {
      const int ii = e / r, jj = e - ii * r;
      if (jj <= ii) {
        const int i = p + 1 + ii, j = p + 1 + jj;
        A[i * lda + j] = fma(-A[i * lda + p] * inv, A[j * lda + p], A[i * lda + j]);
      }
    }
    __syncwarp();
    for (int i = p + 1 + lane; i < m; i += 32) A[i * lda + p] *= inv;
    if (lane == 0) A[p * lda + p] = inv;
    __syncwarp();
  }
  return st;
}

// The same factorisation by the whole CTA (large blocks, m up to a few hundred, A in
// shared memory): per pivot the trailing update is spread over all threads.  Returns
// the flag word to every thread; `flag_s`: one int of shared memory.  Ends with a barrier.
__device__ __forceinline__ int cta_ldlt(double *A, int lda, int m, int *flag_s) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0) *flag_s = 0;
  __syncthreads();
  for (int p = 0; p < m; p++) {
    const double d = A[p * lda + p];
    const double inv = 1.0 / d;
    const int r = m - p - 1;  // trailing size
    // trailing update A[i][j] -= A[i][p] A[j][p] / d, p < j <= i  (column p is still unscaled)
    for (int e = tid; e < r * r; e += nthr) {
      const int ii = e / r, jj = e - ii * r;
      if (jj <= ii) {
        const int i = p + 1 + ii, j = p + 1 + jj;
        A[i * lda + j] = fma(-A[i * lda + p] * inv, A[j * lda + p], A[i * lda + j]);
      }
    }
    __syncthreads();
    for (int i = p + 1 + tid; i < m; i += nthr) A[i * lda + p] *= inv;
    if (tid == 0) {
      A[p * lda + p] = inv;
      if (!(d > 0.0)) *flag_s |= (d == 0.0 || d != d) ? LQ_FLAG_SING : LQ_FLAG_NOTPD;
    }
    __syncthreads();
  }
  return *flag_s;
}

// fast FP64 reciprocal: hardware seed (MUFU.RCP64H) + two Newton steps; full
// double precision for normal, finite, non-zero x (pivots are checked by the
// callers), without the slow path of a generic IEEE division
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// Register-resident LDL^T for compile-time M <= 32: lane i keeps row i of the
// block; per pivot the pivot column is broadcast with warp shuffles (no shared
// memory traffic, no integer division, no barriers inside the sweep).
// keep_on_fail: a block with a non-positive / zero pivot is left untouched (the
// caller switches to the pivoted inverse, warp_pivoted_inverse).
template <int M>
__device__ __forceinline__ int warp_ldlt_reg(double *A, int lda, bool keep_on_fail = false) {
  const int lane = threadIdx.x & 31;
  const int row = lane < M ? lane : M - 1;
  double a[M];
#pragma unroll
  for (int j = 0; j < M; j++) a[j] = A[row * lda + j];
  int st = 0;
#pragma unroll
  for (int p = 0; p < M; p++) {
    const double dp = __shfl_sync(0xffffffffu, a[p], p);
    if (!(dp > 0.0)) st |= (dp == 0.0 || dp != dp) ? LQ_FLAG_SING : LQ_FLAG_NOTPD;
    const double inv = fast_rcp(dp);
    const double li = a[p] * inv;  // multiplier L[row][p] (meaningful for row > p)
#pragma unroll
    for (int j = p + 1; j < M; j++) {
      const double ajp = __shfl_sync(0xffffffffu, a[p], j);  // A[j][p] before scaling
      a[j] = fma(-li, ajp, a[j]);
    }
    a[p] = (lane == p) ? inv : li;
  }
  __syncwarp();  // lanes >= M read row M-1 above: order them before its write-back
  if (lane < M && !(keep_on_fail && st)) {
#pragma unroll
    for (int j = 0; j < M; j++)
      if (j <= lane) A[lane * lda + j] = a[j];
  }
  __syncwarp();
  return st;
}

template <int M>
__device__ __forceinline__ int warp_ldlt_any(double *A, int lda, int m, bool keep_on_fail = false) {
  if constexpr (M > 0 && M <= 32)
    return warp_ldlt_reg<M>(A, lda, keep_on_fail);
  else
    return warp_ldlt(A, lda, m);
}

// Solve (L D L') y = b in place for one right-hand side held at y[i*ys],
// i < m, by ONE thread.  LD as produced by warp_ldlt.
__device__ __forceinline__ void thread_ldlt_solve(const double *LD, int lda, int m, double *y,
                                                  int ys) {
#pragma unroll
  for (int i = 1; i < m; i++) {
    double s = y[i * ys];
#pragma unroll
    for (int l = 0; l < i; l++) s = fma(-LD[i * lda + l], y[l * ys], s);
    y[i * ys] = s;
  }
#pragma unroll
  for (int i = 0; i < m; i++) y[i * ys] *= LD[i * lda + i];
#pragma unroll
  for (int i = m - 2; i >= 0; i--) {
    double s = y[i * ys];
#pragma unroll
    for (int l = i + 1; l < m; l++) s = fma(-LD[l * lda + i], y[l * ys], s);
    y[i * ys] = s;
  }
}

// Same solve with the right-hand side held in registers (compile-time M): the
// substitution chains stay in the register file instead of round-tripping
// through shared memory on every step.
template <int M>
__device__ __forceinline__ void thread_ldlt_solve_reg(const double *LD, int lda, double *y,
                                                      int ys) {
  double r[M];
#pragma unroll
  for (int i = 0; i < M; i++) r[i] = y[i * ys];
#pragma unroll
  for (int i = 1; i < M; i++) {
#pragma unroll
    for (int l = 0; l < i; l++) r[i] = fma(-LD[i * lda + l], r[l], r[i]);
  }
#pragma unroll
  for (int i = 0; i < M; i++) r[i] *= LD[i * lda + i];
#pragma unroll
  for (int i = M - 2; i >= 0; i--) {
#pragma unroll
    for (int l = i + 1; l < M; l++) r[i] = fma(-LD[l * lda + i], r[l], r[i]);
  }
#pragma unroll
  for (int i = 0; i < M; i++) y[i * ys] = r[i];
}

template <int M>
__device__ __forceinline__ void ldlt_solve_any(const double *LD, int lda, int m, double *y,
                                               int ys) {
  if constexpr (M > 0 && M <= 16)
    thread_ldlt_solve_reg<M>(LD, lda, y, ys);
  else
    thread_ldlt_solve(LD, lda, m, y, ys);
}

// ---------------------------------------------------------------------------
// Inverse of the N x N matrix M (shared memory, ldm) by ONE warp, register
// resident: lane i owns rows i (and i + 32 for N > 32) -- in-place Gauss-Jordan
// with implicit partial pivoting.  Per pivot the dependent path is a redux over
// 32-bit magnitude keys (the high words of |a_ip|: enough to pick a pivot
// within 2^-20 of the column maximum), the pivot lane's row and the reciprocal
// it computed meanwhile published through a double-buffered shared row, and N
// independent multiply-adds per lane; no CTA barrier (the CTA-wide version
// spent ~900 cycles per pivot, measured on B200).
// Minv (shared, ldi) <- M^{-1}; scratch in shared memory: rowbuf 2 (N + 2)
// doubles (16-byte aligned), rowsel 65 ints.
// Returns LQ_FLAG_SING on a zero pivot column (the sweep still completes).
// ---------------------------------------------------------------------------
#ifdef LQ_GJ_STAMPS
__device__ long long g_gj_stamps[80];
#endif
#ifndef LQ_GJ_MODE
#define LQ_GJ_MODE 0  // warp inverse of the tree combines: 0 plain, 1 look-ahead, 2 look-ahead + shuffles
#endif
template <int N>
__device__ __forceinline__ int warp_gj_inverse(const double *M, int ldm, double *Minv, int ldi,
                                               double *rowbuf, int *rowsel) {
  constexpr int NR = (N + 31) / 32;
  static_assert(NR <= 2 && N % 2 == 0, "warp_gj_inverse: even N <= 64");
  const int lane = threadIdx.x & 31;
  double a[NR][N];
  bool used[NR];
  int mystep[NR];
  double dinv[NR];  // reciprocal pivot of the own row(s): rows are scaled at the end
#pragma unroll
  for (int rr = 0; rr < NR; rr++) {
    const int row = lane + 32 * rr;
    used[rr] = row >= N;  // rows past the matrix never become pivots
    mystep[rr] = 0;
    dinv[rr] = 1.0;
#pragma unroll
    for (int j = 0; j < N; j++) a[rr][j] = row < N ? M[row * ldm + j] : 0.0;
  }
  int flag = 0;
#pragma unroll
  for (int p = 0; p < N; p++) {
    // local candidate, its key and its reciprocal
#ifdef LQ_GJ_STAMPS
    if (lane == 0) g_gj_stamps[p] = clock64();
#endif
    unsigned key = 0u;
    int krr = 0;
    bool cand = false;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
      if (!used[rr]) {
        const unsigned k = (unsigned)__double2hiint(fabs(a[rr][p]));
        if (!cand || k > key) { key = k; krr = rr; }
        cand = true;
      }
    }
    const double mine = (NR > 1 && krr) ? a[NR - 1][p] : a[0][p];
    const double myinv = fast_rcp(mine);
    const unsigned best = __reduce_max_sync(0xffffffffu, cand ? key : 0u);
    const unsigned ball = __ballot_sync(0xffffffffu, cand && key == best);
    if (best == 0u) flag |= LQ_FLAG_SING;
    const int rl = __ffs(ball) - 1;
    const bool me = lane == rl;
    // the pivot lane publishes its (unscaled) row and 1/pivot through shared
    // memory: a broadcast read costs one wavefront, N shuffles cost ~10 cycles each
    double *rb = rowbuf + (p & 1) * (N + 2);
    if (me) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        double2 v;
        v.x = (NR > 1 && krr) ? a[NR - 1][j] : a[0][j];
        v.y = (NR > 1 && krr) ? a[NR - 1][j + 1] : a[0][j + 1];
        *reinterpret_cast<double2 *>(rb + j) = v;
      }
      rb[N] = myinv;
      if (NR > 1) rb[N + 1] = (double)krr;  // (double-buffered with the row)
      rowsel[p] = lane + 32 * krr;
    }
    __syncwarp();
    const double inv = rb[N];
    const int rsel = NR > 1 ? (int)rb[N + 1] : 0;
    // row_i -= (a_ip / piv) row_r for i != r; the pivot row stays unscaled, the
    // unit column of the right-hand identity takes the place of column p
    double m[NR];
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
      const bool isr = me && rr == rsel;
      m[rr] = isr ? 0.0 : a[rr][p] * inv;
      a[rr][p] = isr ? 1.0 : -m[rr];
      if (isr) { used[rr] = true; mystep[rr] = p; dinv[rr] = inv; }
    }
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const double2 prj = *reinterpret_cast<const double2 *>(rb + j);
#pragma unroll
      for (int rr = 0; rr < NR; rr++) {
        if (j != p) a[rr][j] = fma(-m[rr], prj.x, a[rr][j]);
        if (j + 1 != p) a[rr][j + 1] = fma(-m[rr], prj.y, a[rr][j + 1]);
      }
    }
  }
  __syncwarp();
#ifdef LQ_GJ_STAMPS
  if (lane == 0) g_gj_stamps[N] = clock64();
#endif
  // register a[rr][q] / piv of the row chosen at step p is Minv[p][rowsel[q]]
#pragma unroll
  for (int rr = 0; rr < NR; rr++) {
    if (lane + 32 * rr < N) {
#pragma unroll
      for (int q = 0; q < N; q++) Minv[mystep[rr] * ldi + rowsel[q]] = a[rr][q] * dinv[rr];
    }
  }
  return flag;
}

// Same inverse for N <= 32 with the pivot search taken off the dependent path
// ("look-ahead"): within step p the column p+1 is updated first, its magnitude
// keys go into the redux / ballot that pick pivot p+1 and every candidate lane
// starts the reciprocal of its own entry -- all of which then overlaps the N-2
// remaining column updates of step p instead of following them.  The dependent
// chain of a step shrinks to: publish the pivot row, read 1/pivot, one multiply,
// one multiply-add.  BCAST_SHFL: the pivot row travels by shuffles from the
// (runtime) pivot lane instead of through the shared row.
template <int N, bool BCAST_SHFL = false>
__device__ __forceinline__ int warp_gj_inverse_la(const double *M, int ldm, double *Minv,
                                                  int ldi, double *rowbuf, int *rowsel) {
  static_assert(N <= 32 && N % 2 == 0, "warp_gj_inverse_la: even N <= 32");
  const int lane = threadIdx.x & 31;
  double a[N];
  bool used = lane >= N;  // rows past the matrix never become pivots
  int mystep = 0;
  double dinv = 1.0;
#pragma unroll
  for (int j = 0; j < N; j++) a[j] = lane < N ? M[lane * ldm + j] : 0.0;
  int flag = 0;
  unsigned key = used ? 0u : (unsigned)__double2hiint(fabs(a[0]));
  double myinv = fast_rcp(a[0]);
  unsigned best = __reduce_max_sync(0xffffffffu, key);
  unsigned ball = __ballot_sync(0xffffffffu, !used && key == best);
#pragma unroll
  for (int p = 0; p < N; p++) {
    if (best == 0u) flag |= LQ_FLAG_SING;
    const int rl = __ffs(ball) - 1;
    const bool me = lane == rl;
    double *rb = rowbuf + (p & 1) * (N + 2);
    double inv;
    if constexpr (BCAST_SHFL) {
      inv = __shfl_sync(0xffffffffu, myinv, rl);
      if (me) rowsel[p] = lane;
    } else {
      if (me) {
#pragma unroll
        for (int j = 0; j < N; j += 2)
          *reinterpret_cast<double2 *>(rb + j) = make_double2(a[j], a[j + 1]);
        rb[N] = myinv;
        rowsel[p] = lane;
      }
      __syncwarp();
      inv = rb[N];
    }
    const double m = me ? 0.0 : a[p] * inv;
    if (me) { used = true; mystep = p; dinv = inv; }
    // column p+1 first, then the selection of pivot p+1
    if (p + 1 < N) {
      const double prn = BCAST_SHFL ? __shfl_sync(0xffffffffu, a[p + 1], rl) : rb[p + 1];
      a[p + 1] = fma(-m, prn, a[p + 1]);
      key = used ? 0u : (unsigned)__double2hiint(fabs(a[p + 1]));
      myinv = fast_rcp(a[p + 1]);
      best = __reduce_max_sync(0xffffffffu, key);
      ball = __ballot_sync(0xffffffffu, !used && key == best);
    }
    // the unit column of the right-hand identity takes the place of column p
    if constexpr (BCAST_SHFL) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (j == p || j == p + 1) continue;
        const double prj = __shfl_sync(0xffffffffu, a[j], rl);
        a[j] = fma(-m, prj, a[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        const double2 prj = *reinterpret_cast<const double2 *>(rb + j);
        if (j != p && j != p + 1) a[j] = fma(-m, prj.x, a[j]);
        if (j + 1 != p && j + 1 != p + 1) a[j + 1] = fma(-m, prj.y, a[j + 1]);
      }
    }
    a[p] = me ? 1.0 : -m;
  }
  __syncwarp();
  // register a[q] / piv of the row chosen at step p is Minv[p][rowsel[q]]
  if (lane < N) {
#pragma unroll
    for (int q = 0; q < N; q++) Minv[mystep * ldi + rowsel[q]] = a[q] * dinv;
  }
  return flag;
}

// ---------------------------------------------------------------------------
// Indefinite stage block (a16): the reference factors Guu with a diagonally
// scaled Bunch-Kaufman (sc_i = 1/sqrt(g_ii) where g_ii > 1, BKPfactor;
// hqp/Hqp_IpLQDOCP.C:1860-1879, meschach/bkpfacto.c:102-226).  Here LDL^T without
// interchanges is tried first (positive definite blocks: the IP iterations of a
// convex QP); a block with a non-positive pivot is redone by ONE warp as
//     Guu^{-1} = S (S Guu S)^{-1} S,   S = diag(sc),
// with the register-resident Gauss-Jordan whose pivot search is a warp redux over
// magnitude keys (warp_gj_inverse: partial pivoting, stable for symmetric
// indefinite blocks).  A (M x M, lda, full symmetric block) is replaced by the
// explicit inverse; scr: M (M+1) + 2 (M+2) + M doubles, rowsel: 65 ints.
// Returns LQ_FLAG_SING for a singular block.
// ---------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ int warp_pivoted_inverse(double *A, int lda, double *scr, int *rowsel) {
  constexpr int ME = (M + 1) & ~1;  // warp_gj_inverse wants an even order: pad with an identity row
  const int lane = threadIdx.x & 31;
  double *Ms = scr, *rowbuf = scr + ME * (ME + 1), *sc = rowbuf + 2 * (ME + 2);
  if (lane < M) {
    const double g = A[lane * lda + lane];
    sc[lane] = g > 1.0 ? rsqrt(g) : 1.0;
  }
  __syncwarp();
  for (int e = lane; e < ME * ME; e += 32) {
    const int i = e / ME, j = e - i * ME;
    // (symmetrise from the lower triangle: the upper one may be stale)
    double v = (i == j) ? 1.0 : 0.0;
    if (i < M && j < M) v = sc[i] * sc[j] * (j <= i ? A[i * lda + j] : A[j * lda + i]);
    Ms[i * (ME + 1) + j] = v;
  }
  __syncwarp();
  const int fl = warp_gj_inverse<ME>(Ms, ME + 1, Ms, ME + 1, rowbuf, rowsel);
  __syncwarp();
  for (int e = lane; e < M * M; e += 32) {
    const int i = e / M, j = e - i * M;
    A[i * lda + j] = sc[i] * sc[j] * Ms[i * (ME + 1) + j];
  }
  __syncwarp();
  return fl;
}

// y <- Ainv y for one right-hand side held at y[i*ys] by ONE thread (Ainv: the
// explicit inverse left by warp_pivoted_inverse)
template <int M>
__device__ __forceinline__ void thread_inv_apply(const double *Ainv, int lda, double *y, int ys) {
  double r[M], o[M];
#pragma unroll
  for (int i = 0; i < M; i++) r[i] = y[i * ys];
#pragma unroll
  for (int i = 0; i < M; i++) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < M; l++) s = fma(Ainv[i * lda + l], r[l], s);
    o[i] = s;
  }
#pragma unroll
  for (int i = 0; i < M; i++) y[i * ys] = o[i];
}

// ---------------------------------------------------------------------------
// Gauss-Jordan with (implicit) partial pivoting on the augmented n x nc matrix
// M (ldm), nc > n.  Rows are never swapped: column p is eliminated with the
// not-yet-used row of largest modulus, the row permutation is kept in piv_s and
// undone when the solution A^{-1} B is copied to X (n x (nc-n), ldx = nc-n).
// Whole CTA cooperates; ONE barrier per pivot.  A quad of lanes owns a column
// j > p (lane q of the quad the rows i = q mod 4), so the per-thread work of a
// pivot is n/4 FMAs; the quad that owns column p+1 also picks the next pivot
// row (quad shuffle reduction) and its reciprocal while it updates that column.
// piv_s: >= n ints, inv_s: 2 doubles (shared).  Status is OR-ed into *st_s.
// NX > 0: n is a compile-time constant (unrolled row loops).
// ---------------------------------------------------------------------------
template <int NX>
__device__ __forceinline__ void cta_gauss_jordan(double *M, int ldm, int n, int nc, double *X,
                                                 int *piv_s, double *inv_s, int *st_s) {
  __syncthreads();
  if (threadIdx.x < 32) {  // pivot row of column 0
    double best = -1.0;
    int bi = 0;
    for (int i = (threadIdx.x & 31); i < n; i += 32) {
      const double a = fabs(M[i * ldm]);
      if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (threadIdx.x == 0) {
      piv_s[0] = bi;
      inv_s[0] = 1.0 / M[bi * ldm];
      if (!(best > 0.0)) *st_s |= LQ_FLAG_SING;
    }
  }
  __syncthreads();
  // rows already consumed as pivot rows (n <= 256: four 64-bit words)
  unsigned long long used0 = 0ull, used1 = 0ull, used2 = 0ull, used3 = 0ull;
  auto is_used = [&](int i) -> bool {
    const unsigned long long wsel = i < 128 ? (i < 64 ? used0 : used1) : (i < 192 ? used2 : used3);
    return (wsel >> (i & 63)) & 1ull;
  };
  const int q = threadIdx.x & 3, jl = threadIdx.x >> 2, ncol = blockDim.x >> 2;
  for (int p = 0; p < n; p++) {
    const int r = piv_s[p];
    const double inv = inv_s[p & 1];
    {
      const unsigned long long bit = 1ull << (r & 63);
      if (r < 64) used0 |= bit; else if (r < 128) used1 |= bit;
      else if (r < 192) used2 |= bit; else used3 |= bit;
    }
    for (int j0 = p + 1; j0 < nc; j0 += ncol) {  // uniform trip count: shuffles below
      const int j = j0 + jl;
      const bool act = j < nc;
      const bool next = (j == p + 1) && (p + 1 < n);
      const double prj = act ? M[r * ldm + j] * inv : 0.0;
      double best = -1.0, bv = 0.0;
      int bi = 0;
      if (act) {
        constexpr int NR = NX > 0 ? (NX + 3) / 4 : 1;
        if constexpr (NX > 0) {
          double f[NR], c[NR];
#pragma unroll
          for (int k = 0; k < NR; k++) {
            const int i = q + 4 * k;
            f[k] = i < NX ? M[i * ldm + p] : 0.0;
            c[k] = i < NX ? M[i * ldm + j] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < NR; k++) {
            const int i = q + 4 * k;
            c[k] = fma(-f[k], prj, c[k]);
            if (i < NX && i != r) M[i * ldm + j] = c[k];
            if (i < NX && !is_used(i) && fabs(c[k]) > best) {
              best = fabs(c[k]); bv = c[k]; bi = i;
            }
          }
        } else {
          for (int i = q; i < n; i += 4) {
            const double v = fma(-M[i * ldm + p], prj, M[i * ldm + j]);
            if (i != r) M[i * ldm + j] = v;
            if (!is_used(i) && fabs(v) > best) { best = fabs(v); bv = v; bi = i; }
          }
        }
      }
      __syncwarp();  // every lane of the quad has read M[r][j] before it is overwritten
      if (act && q == (r & 3)) M[r * ldm + j] = prj;  // pivot row: scaled, not eliminated
      // next pivot: reduce (best, bi, bv) over the quad that owns column p+1;
      // ties go to the smaller row index
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bv = ov; bi = oi; }
      }
      if (next && q == 0) {
        piv_s[p + 1] = bi;
        inv_s[(p + 1) & 1] = fast_rcp(bv);
        if (!(best > 0.0)) *st_s |= LQ_FLAG_SING;
      }
    }
    __syncthreads();
  }
  // X[p][:] = row piv_s[p] of the right-hand part
  const int nr = nc - n;
  for (int e = threadIdx.x; e < n * nr; e += blockDim.x) {
    const int p = e / nr, j = e - p * nr;
    X[e] = M[piv_s[p] * ldm + n + j];
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// In-place inverse of the n x n block M (ldm; global memory, large blocks) by the
// whole CTA: Gauss-Jordan with partial pivoting and explicit row interchanges, the
// pivot row and the multiplier column staged in shared memory (scr: 2 n doubles),
// one rank-1 update of the whole block per pivot (coalesced along the rows), the
// column permutation undone at the end.  piv_s: n ints.  Barriers on entry and exit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cta_gj_inverse_big(double *M, int ldm, int n, int *piv_s,
                                                   double *scr, int *st_s) {
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ double piv_inv;
  double *rowp = scr, *colp = scr + n;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  for (int p = 0; p < n; p++) {
    // pivot search in column p, rows p .. n-1 (ties: smaller row)
    double best = -1.0;
    int bi = p;
    for (int i = p + tid; i < n; i += nthr) {
      const double a = fabs(M[(size_t)i * ldm + p]);
      if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      const int nw = (nthr + 31) >> 5;
      for (int w2 = 1; w2 < nw; w2++)
        if (red_v[w2] > best || (red_v[w2] == best && red_i[w2] < bi)) { best = red_v[w2]; bi = red_i[w2]; }
      piv_s[p] = bi;
      if (!(best > 0.0)) *st_s |= LQ_FLAG_SING;
      piv_inv = 1.0 / M[(size_t)bi * ldm + p];
    }
    __syncthreads();
    const int r = piv_s[p];
    const double inv = piv_inv;
    // interchange rows p and r; the pivot row is scaled and staged
    for (int j = tid; j < n; j += nthr) {
      const double a = M[(size_t)r * ldm + j];
      if (r != p) M[(size_t)r * ldm + j] = M[(size_t)p * ldm + j];
      const double v = (j == p) ? inv : a * inv;
      M[(size_t)p * ldm + j] = v;
      rowp[j] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) colp[i] = (i == p) ? 0.0 : M[(size_t)i * ldm + p];
    __syncthreads();
    // rank-1 update of every other row; column p receives -f / pivot.  A thread
    // owns a column (coalesced along the rows, no index arithmetic) and keeps 16
    // independent loads in flight: the block lives in L2, not in shared memory
    for (int j = tid; j < n; j += nthr) {
      const double rj = rowp[j];
      const bool isp = j == p;
      double *col = M + j;
#pragma unroll 1
      for (int i0 = 0; i0 < n; i0 += 16) {
        double cur[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const int i = i0 + q;
          cur[q] = (i < n && !isp) ? col[(size_t)i * ldm] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const int i = i0 + q;
          if (i < n && i != p) col[(size_t)i * ldm] = fma(-colp[i], rj, cur[q]);
        }
      }
    }
    __syncthreads();
  }
  // (P A)^{-1} = A^{-1} P': undo the interchanges on the columns, last first
  for (int p = n - 1; p >= 0; p--) {
    const int r = piv_s[p];
    if (r == p) continue;  // (uniform)
    for (int i = tid; i < n; i += nthr) {
      const double a = M[(size_t)i * ldm + p];
      M[(size_t)i * ldm + p] = M[(size_t)i * ldm + r];
      M[(size_t)i * ldm + r] = a;
    }
    __syncthreads();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// X = M0^{-1} R for the augmented W = [M0 | R] (n x nc, ldm; global memory, large
// blocks) by the whole CTA: BLOCKED Gauss-Jordan elimination with partial pivoting.
// Round 2's first version inverted M0 with one rank-1 update of the whole block per
// pivot -- 200 dependent passes over L2, 3.4 of the 4.5 ms of a tree level at nx =
// 200.  Here LQ_GJ_W pivots are taken at a time:
//   1. the panel W[:, k0 .. k0+w) is eliminated in shared memory, pivot by pivot
//      (search over the rows not used yet, implicit row pivoting: no interchanges in
//      memory); its columns keep the multipliers F;
//   2. the w "effective" pivot rows Y of all columns to the right follow from a small
//      triangular recurrence (one column per thread);
//   3. every other entry to the right takes ONE rank-w update  W += U Y  on the
//      tensor cores (cta_mm_big; U = -F, with the unit / zero pattern of the pivot
//      rows), instead of w rank-1 passes.
// At the end row piv[c] of the right-hand part is row c of X.
// scr: (LQ_GJ_W) * (n + nc + 4) doubles of GLOBAL scratch (operands of the update);
// stg: the GEMM staging ring (panel and flags live there between the products).
// piv_s: n ints.  Barriers on entry and exit.
// ---------------------------------------------------------------------------
#define LQ_GJ_W 32
#define LQ_GJ_MAXN 1024
#ifdef LQ_GJ_STAMPS
__device__ long long g_gj_cyc[8];  // cycles per phase, thread 0 of CTA 0 (microbenchmark builds)
#define GJ_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_gj_cyc[i] += t_ - gj_t0; gj_t0 = t_; } } while (0)
#else
#define GJ_STAMP(i) do { } while (0)
#endif
__host__ __device__ inline size_t big_gj_scratch_doubles(int nx) {
  return (size_t)LQ_GJ_W * (4 * (size_t)nx + 8);
}
__device__ __forceinline__ void cta_gj_solve_blocked(double *stg, double *scr, double *W, int ldm, int n,
                                                     int nc, double *X, int ldx, int *piv_s, int *st_s) {
  constexpr int PW = LQ_GJ_W + 1;  // panel row stride in shared memory
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ int blk_row[LQ_GJ_W];
  __shared__ double prow[LQ_GJ_W];
  __shared__ unsigned char used[LQ_GJ_MAXN];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  double *Pn = stg;                              // [n][PW] panel, then multipliers
  const int ldu = (n + 1) & ~1, ldy = (nc + 1) & ~1;
  double *Ut = scr, *Yg = scr + (size_t)LQ_GJ_W * ldu;  // [w][ldu], [w][ldy]
  for (int i = tid; i < n; i += nthr) used[i] = 0;
  // panel elimination: nsplit threads per row (no index arithmetic inside the pivot loop)
  const int nsplit = nthr / n > 0 ? nthr / n : 1;
  const int my_row = nthr >= n ? tid / nsplit : n, my_part = tid - (tid / nsplit) * nsplit;
  __syncthreads();
#ifdef LQ_GJ_STAMPS
  long long gj_t0 = clock64();
#endif
  for (int k0 = 0; k0 < n; k0 += LQ_GJ_W) {
    const int w = n - k0 < LQ_GJ_W ? n - k0 : LQ_GJ_W;
    // ---- 1. panel into shared memory and eliminated there
    for (int e = tid; e < n * w; e += nthr) {
      const int i = e / w, q = e - i * w;
      Pn[i * PW + q] = W[(size_t)i * ldm + k0 + q];
    }
    __syncthreads();
    GJ_STAMP(0);
    for (int p = 0; p < w; p++) {
      // pivot of column p among the rows not used yet (ties: smaller row)
      double best = -1.0;
      int bi = n;
      for (int i = tid; i < n; i += nthr) {
        if (!used[i]) {
          const double a = fabs(Pn[i * PW + p]);
          if (a > best || (a == best && i < bi) || bi == n) { best = a; bi = i; }
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi < n && (bi == n || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
      }
      if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        // second round over the per-warp results, then the copy of the (unscaled) pivot row
        const int nw = (nthr + 31) >> 5;
        best = lane < nw ? red_v[lane] : -1.0;
        bi = lane < nw ? red_i[lane] : n;
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi < n && (bi == n || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
        }
        if (bi >= n) bi = 0;  // (cannot happen: n - k0 - p rows are unused)
        if (lane == 0) {
          if (!(best > 0.0)) *st_s |= LQ_FLAG_SING;
          piv_s[k0 + p] = bi;
          blk_row[p] = bi;
          used[bi] = 1;
        }
        for (int q = lane; q < w; q += 32) prow[q] = Pn[bi * PW + q];
      }
      __syncthreads();
      const int r = blk_row[p];
      const double inv = 1.0 / prow[p];
      // eliminate column p from every other row of the panel (columns q > p; column p
      // keeps the multiplier); the pivot row is scaled in the same pass (the others read
      // its copy `prow`).  `nsplit` threads share a row.
      if (my_row < n) {
        double *row = Pn + my_row * PW;
        if (my_row != r) {
          const double f = -row[p] * inv;
          for (int q = p + 1 + my_part; q < w; q += nsplit) row[q] = fma(f, prow[q], row[q]);
        } else {
          for (int q = p + 1 + my_part; q < w; q += nsplit) row[q] = prow[q] * inv;
          if (my_part == 0) row[p] = inv;
        }
      }
      __syncthreads();
    }
    GJ_STAMP(1);
    const int j0 = k0 + w, no = nc - j0;  // columns to the right
    if (no <= 0) break;
    // ---- 2. effective pivot rows  y_p = inv_p (W[r_p] - sum_{p' < p} F[r_p][p'] y_p'),
    //         one column per thread; the pivot rows themselves are cleared (they are
    //         rebuilt by the update below)
    for (int j = tid; j < no; j += nthr) {
      double y[LQ_GJ_W];
#pragma unroll
      for (int p = 0; p < LQ_GJ_W; p++)
        y[p] = p < w ? W[(size_t)blk_row[p] * ldm + j0 + j] : 0.0;
#pragma unroll
      for (int p = 0; p < LQ_GJ_W; p++) {
        if (p < w) {
          const int r = blk_row[p];
          y[p] *= Pn[r * PW + p];  // (1 / pivot)
          Yg[(size_t)p * ldy + j] = y[p];
          W[(size_t)r * ldm + j0 + j] = 0.0;
          // right-looking: the later pivot rows lose their multiple of y_p now (independent updates)
#pragma unroll
          for (int pp = 0; pp < LQ_GJ_W; pp++)
            if (pp > p && pp < w) y[pp] = fma(-Pn[blk_row[pp] * PW + p], y[p], y[pp]);
        }
      }
    }
    GJ_STAMP(2);
    // U' (w x n): -F for the other rows; pivot row r_p0: unit at p0, -F right of it, 0 left
    for (int e = tid; e < n * w; e += nthr) {
      const int p = e / n, i = e - p * n;
      int p0 = -1;
#pragma unroll
      for (int pp = 0; pp < LQ_GJ_W; pp++)
        if (pp < w && blk_row[pp] == i) p0 = pp;
      double u = -Pn[i * PW + p];
      if (p0 >= 0) u = p == p0 ? 1.0 : (p > p0 ? u : 0.0);
      Ut[(size_t)p * ldu + i] = u;
    }
    __syncthreads();
    GJ_STAMP(3);
    // ---- 3. W[:, j0 ..] += U Y
    cta_mm_big(stg, W + j0, ldm, W + j0, ldm, 1.0, 1.0, Ut, 1, ldu, Yg, ldy, 1, n, no, w, 0,
               (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5));
    __syncthreads();
    GJ_STAMP(4);
  }
  // X[c][:] = row piv[c] of the right-hand part
  const int nr = nc - n;
  for (int e = tid; e < n * nr; e += nthr) {
    const int c = e / nr, j = e - c * nr;
    X[(size_t)c * ldx + j] = W[(size_t)piv_s[c] * ldm + n + j];
  }
  __syncthreads();
}

// X = M0^{-1} R for the augmented M = [M0 | R] (n x nc, ldm), large blocks.  gjs: the
// CTA's global scratch for the blocked elimination (NULL or n too large: M0 is inverted
// in place, one rank-1 update per pivot, then applied to R on the tensor cores).
__device__ __forceinline__ void cta_inverse_apply_big(double *stg, double *gjs, double *M, int ldm, int n,
                                                      int nc, double *X, int ldx, int *piv_s,
                                                      int *st_s) {
  if (gjs && n <= LQ_GJ_MAXN && n <= (int)blockDim.x && (size_t)n * (LQ_GJ_W + 1) <= (size_t)LQ_BIG_STG) {
    cta_gj_solve_blocked(stg, gjs, M, ldm, n, nc, X, ldx, piv_s, st_s);
    return;
  }
  cta_gj_inverse_big(M, ldm, n, piv_s, stg, st_s);
  cta_mm_big(stg, X, ldx, nullptr, 0, 0.0, 1.0, M, ldm, 1, M + n, ldm, 1, n, nc - n, n, 0,
             (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5));
  __syncthreads();
}

// X = M0^{-1} R for the augmented M = [M0 | R] (n x nc, ldm) of the compiled
// sizes: warp 0 inverts M0 in registers (warp_gj_inverse), then the whole CTA
// applies the inverse to R on the tensor cores.  Same contract as
// cta_gauss_jordan (X: n x (nc - n), ld = ldx or nc - n; barriers on entry and exit);
// ldm should be odd (conflict-free row-per-lane reads).
// scratch: NX (NX + 1) + 2 (NX + 2) doubles, rowsel: 65 ints.
template <int NX, int NW>
__device__ __forceinline__ void cta_inverse_apply(double *M, int ldm, int nc, double *X,
                                                  double *scratch, int *rowsel, int *st_s,
                                                  int ldx = 0) {
  if (ldx == 0) ldx = nc - NX;
  static_assert(NX > 0, "compiled sizes only");
  __syncthreads();
  double *Minv = scratch, *rowbuf = scratch + NX * (NX + 1);
  if (warp_id_uniform() == 0) {
    int fl;
    if constexpr (NX <= 32 && LQ_GJ_MODE == 1)
      fl = warp_gj_inverse_la<NX, false>(M, ldm, Minv, NX + 1, rowbuf, rowsel);
    else if constexpr (NX <= 32 && LQ_GJ_MODE == 2)
      fl = warp_gj_inverse_la<NX, true>(M, ldm, Minv, NX + 1, rowbuf, rowsel);
    else
      fl = warp_gj_inverse<NX>(M, ldm, Minv, NX + 1, rowbuf, rowsel);
    if (fl && (threadIdx.x & 31) == 0) *st_s |= fl;
  }
  __syncthreads();
  cta_mm_tc<NW>(X, ldx, nullptr, 0, 0.0, 1.0, Minv, NX + 1, 1, M + NX, ldm, 1, NX, nc - NX,
                NX);
  __syncthreads();
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double val) {
  // for val >= 0 the IEEE bit pattern is monotone as an unsigned integer;
  // NaN (all-ones exponent, non-zero mantissa) compares above +inf and sticks.
  atomicMax(reinterpret_cast<unsigned long long *>(addr),
            static_cast<unsigned long long>(__double_as_longlong(fabs(val))));
}

// base of the CTA's block storage: dynamic shared memory, or its slice of the
// global workspace when the stage blocks are too large for it
__device__ __forceinline__ unsigned char *cta_workspace(const LqDev &d, unsigned char *smem_raw) {
  if (!d.gws) return smem_raw;
  const size_t cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
  return reinterpret_cast<unsigned char *>(d.gws + cta * d.gws_stride);
}

// the CTA's scratch for the blocked elimination: the tail of its workspace slice
__device__ __forceinline__ double *big_gj_scratch(const LqDev &d) {
  if (!d.gws) return nullptr;
  const size_t cta = blockIdx.x + (size_t)gridDim.x * blockIdx.y;
  return d.gws + (cta + 1) * d.gws_stride - big_gj_scratch_doubles(d.nx);
}

// shared-memory carve-up helper (16-byte granularity)
struct SmemCarver {
  unsigned char *p;
  __device__ explicit SmemCarver(void *base) : p(reinterpret_cast<unsigned char *>(base)) {}
  __device__ double *take(int n) {
    double *r = reinterpret_cast<double *>(p);
    p += (size_t)((n + 1) & ~1) * sizeof(double);
    return r;
  }
  __device__ uint64_t *take_bars(int n) {
    uint64_t *r = reinterpret_cast<uint64_t *>(p);
    p += (size_t)((n + 1) & ~1) * sizeof(uint64_t);
    return r;
  }
};
