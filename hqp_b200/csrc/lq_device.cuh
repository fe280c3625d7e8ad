// Device-side data layout and CTA-level dense FP64 primitives for the
// stage-structured KKT factor/solve (sm_100a).
//
// Everything a kernel needs is passed by value in one LqDev struct.  All stage
// blocks live in contiguous HBM slabs (the reference keeps one heap block per
// stage, hqp/t_mesch.h:43-140); index = (instance * stages + stage) * block.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct LqDev {
  int K, nx, nu, nm, batch;
  int P, L;           // segments per instance, stages per segment
  int fixed_x0;
  int m, nnz;         // inequality rows / nonzeros per instance
  int N, me;          // per-instance vector lengths
  // inequality structure (shared by all instances)
  const int *ineq_stage, *ineq_ptr, *ineq_lcol;
  const int *srow_ptr;  // [K+2] rows sorted by stage
  const int *srow;      // [m]
  const int *vcol_ptr;  // [N+1] per variable: entries of C in that column
  const int *vcol_row;  // [nnz] row of the entry
  const int *vcol_nz;   // [nnz] index of the entry in the CSR value array
  // values (update)
  const double *Q, *fx, *fu, *cval;
  // captured at factor
  const double *z, *w;
  // factor state
  double *V;       // [batch][K+1][nx*nx]   value-function Hessians Vxx
  double *Rux;     // [batch][K][nu*nx]     feedback gains
  double *LD;      // [batch][K][nu*nu]     LDL^T of Guu: unit L below, 1/D on the diagonal
  double *Phi;     // [batch][K][nx*nx]     closed loop fx - fu Rux
  double *segA, *segC, *segJ;  // [batch][P][nx*nx] segment elements
  double *segPsi;  // [batch][P][nx*nx]     closed-loop transition over the segment
  double *segVb;   // [batch][P][nx*nx]     Vxx at the segment end
  double *V0f;     // [batch][nx*nx]        LDL^T of Vxx[0] (free x0)
  int *status;     // device status word (0 ok)
  // solve scratch
  double *g;       // [batch][N]     reduced gradient (gx,gu)
  double *wv;      // [batch][K][nx] gx - Rux' gu
  double *q;       // [batch][K][nx] Vxx[k+1] f_k
  double *v;       // [batch][K+1][nx]
  double *Ru;      // [batch][K][nu]
  double *c;       // [batch][K][nx] f_k - fu Ru
  double *x;       // [batch][K+1][nx]
  double *segv0, *segvb, *segx0, *segxa;  // [batch][P][nx]
};

// status word bits (device) -> HQPCU_E_SING / HQPCU_E_NOTPD (host)
#define LQ_FLAG_SING 1
#define LQ_FLAG_NOTPD 2

// ---------------------------------------------------------------------------
// C(MxN, ldc) = beta*C0 + alpha * A * B with arbitrary element strides:
//   A(i,l) = A[i*ar + l*ac],  B(l,j) = B[l*br + j*bc].
// One output element per thread per pass; consecutive threads walk j, so B
// rows are read conflict-free and A is a broadcast.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cta_mm(double *C, int ldc, const double *C0, int ldc0,
                                       double beta, double alpha, const double *A, int ar,
                                       int ac, const double *B, int br, int bc, int M, int N,
                                       int Kd) {
  const int total = M * N;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / N, j = idx - i * N;
    const double *a = A + i * ar;
    const double *b = B + j * bc;
    double s0 = 0.0, s1 = 0.0;
    int l = 0;
    for (; l + 1 < Kd; l += 2) {
      s0 = fma(a[l * ac], b[l * br], s0);
      s1 = fma(a[(l + 1) * ac], b[(l + 1) * br], s1);
    }
    if (l < Kd) s0 = fma(a[l * ac], b[l * br], s0);
    double r = alpha * (s0 + s1);
    if (C0) r += beta * C0[i * ldc0 + j];
    C[i * ldc + j] = r;
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

// ---------------------------------------------------------------------------
// In-place LDL^T (no pivoting) of the symmetric m x m block A (lda) by warp 0.
// On exit: strictly lower part = unit L, diagonal = 1/D.  Returns (to all
// lanes of warp 0) a bit set: LQ_FLAG_SING for a zero / non-finite pivot,
// LQ_FLAG_NOTPD for a negative one (the factorisation continues).
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

// Solve (L D L') y = b in place for one right-hand side held at y[i*ys],
// i < m, by ONE thread.  LD as produced by warp_ldlt.
__device__ __forceinline__ void thread_ldlt_solve(const double *LD, int lda, int m, double *y,
                                                  int ys) {
  for (int i = 1; i < m; i++) {
    double s = y[i * ys];
    for (int l = 0; l < i; l++) s = fma(-LD[i * lda + l], y[l * ys], s);
    y[i * ys] = s;
  }
  for (int i = 0; i < m; i++) y[i * ys] *= LD[i * lda + i];
  for (int i = m - 2; i >= 0; i--) {
    double s = y[i * ys];
    for (int l = i + 1; l < m; l++) s = fma(-LD[l * lda + i], y[l * ys], s);
    y[i * ys] = s;
  }
}

// ---------------------------------------------------------------------------
// Gauss-Jordan with partial pivoting on the augmented n x nc matrix M (ldm),
// nc >= n: on exit columns n..nc-1 hold A^{-1} B.  Whole CTA cooperates.
// piv_s: shared scratch (>= 2 ints).  Returns status through *st_s (shared).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cta_gauss_jordan(double *M, int ldm, int n, int nc, int *piv_s,
                                                 int *st_s) {
  for (int p = 0; p < n; p++) {
    __syncthreads();
    if (threadIdx.x < 32) {  // pivot search in column p, rows p..n-1
      double best = -1.0;
      int bi = p;
      for (int i = p + (threadIdx.x & 31); i < n; i += 32) {
        const double a = fabs(M[i * ldm + p]);
        if (a > best) { best = a; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) {
        piv_s[0] = bi;
        if (!(best > 0.0)) *st_s |= LQ_FLAG_SING;
      }
    }
    __syncthreads();
    const int r = piv_s[0];
    if (r != p)
      for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        const double t = M[p * ldm + j];
        M[p * ldm + j] = M[r * ldm + j];
        M[r * ldm + j] = t;
      }
    __syncthreads();
    const double inv = 1.0 / M[p * ldm + p];
    __syncthreads();
    for (int j = threadIdx.x; j < nc; j += blockDim.x) M[p * ldm + j] *= inv;
    __syncthreads();
    // eliminate column p from every other row; only columns > p matter
    const int ncols = nc - p - 1;
    for (int e = threadIdx.x; e < n * ncols; e += blockDim.x) {
      const int i = e / ncols, j = p + 1 + (e - i * ncols);
      if (i != p) M[i * ldm + j] = fma(-M[i * ldm + p], M[p * ldm + j], M[i * ldm + j]);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double val) {
  // for val >= 0 the IEEE bit pattern is monotone as an unsigned integer;
  // NaN (all-ones exponent, non-zero mantissa) compares above +inf and sticks.
  atomicMax(reinterpret_cast<unsigned long long *>(addr),
            static_cast<unsigned long long>(__double_as_longlong(fabs(val))));
}
