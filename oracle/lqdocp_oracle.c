/*
 * Plain-C restatement of the reference's LQ-DOCP KKT path.  See
 * lqdocp_oracle.h for scope.  TEST INFRASTRUCTURE ONLY.
 *
 * All matrices are dense row-major.  Sign conventions are the reference's:
 *   [-Q A' C' 0; A 0 0 0; C 0 0 -I; 0 0 W Z] [dx dy dz dw] = [r1 r2 r3 r4]
 * (hqp/Hqp_IpsMehrotra.C:27-31), the Riccati sweep works on xi = -dx
 * (hqp/Hqp_IpLQDOCP.C:952).
 */
#include "lqdocp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct lqo_fact {
  lqo_problem p;
  int nm, N, me;
  double *H;    /* (K+1) nm nm : Q + C' (z/w) C                              */
  double *Vxx;  /* (K+1) nx nx                                               */
  double *Rux;  /* K nu nx                                                    */
  double *Gxu;  /* K nx nu                                                    */
  double *Guf;  /* K nu nu : BKP-factored scaled Guu                          */
  int *piv, *blk; /* K nu each                                               */
  double *sc;   /* K nu                                                       */
  double *V0f;  /* nx nx BKP factor of scaled Vxx[0] (free x0)                */
  int *piv0, *blk0;
  double *sc0;
  double *z, *w;
  /* step work space */
  double *gx, *gu, *Vx, *Ru, *xs, *us, *ps;
  /* solve work space */
  double *t1, *t2, *t3, *t4, *ex, *ey, *ez, *ew;
};

/* ------------------------------------------------------------------------
 * Bunch-Kaufman-Parlett  P'AP = M D M'  with 1x1 / 2x2 diagonal blocks.
 * Follows the pivoting rule of meschach/bkpfacto.c:102-226 (alpha =
 * (1+sqrt(17))/8, :42): 1x1 without interchange if |a_ii| >= alpha*lambda
 * (:139) or |a_ii|*sigma >= alpha*lambda^2 (:165), 1x1 after swapping i<->r
 * if |a_rr| >= alpha*sigma (:167), else a 2x2 block on (i, r->i+1) (:174).
 * Storage here: full symmetric matrix kept symmetric during elimination;
 * multipliers overwrite the strictly lower part column by column.
 * blk[i] == 1: 1x1 pivot at i;  blk[i] == 2: 2x2 pivot at (i,i+1), blk[i+1]=0.
 * ---------------------------------------------------------------------- */
static void sym_swap(double *A, int n, int a, int b) {
  if (a == b) return;
  for (int k = 0; k < n; k++) {
    double t = A[a * n + k]; A[a * n + k] = A[b * n + k]; A[b * n + k] = t;
  }
  for (int k = 0; k < n; k++) {
    double t = A[k * n + a]; A[k * n + a] = A[k * n + b]; A[k * n + b] = t;
  }
}

void lqo_bkp_factor(double *A, int n, int *piv, int *blk) {
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  for (int i = 0; i < n; i++) { piv[i] = i; blk[i] = 1; }
  int i = 0;
  while (i < n) {
    int one = 1;
    double aii = fabs(A[i * n + i]);
    double lambda = 0.0;
    int r = (i + 1 < n) ? i + 1 : i;
    for (int k = i + 1; k < n; k++) {
      double t = fabs(A[i * n + k]);
      if (t >= lambda) { lambda = t; r = k; }
    }
    if (!(aii >= alpha * lambda)) {
      double sigma = 0.0;
      for (int k = i; k < n; k++) {
        if (k == r) continue;
        double t = fabs(A[r * n + k]);
        if (t > sigma) sigma = t;
      }
      if (aii * sigma >= alpha * lambda * lambda || i + 1 == n) {
        one = 1;
      } else if (fabs(A[r * n + r]) >= alpha * sigma) {
        sym_swap(A, n, i, r);
        int t = piv[i]; piv[i] = piv[r]; piv[r] = t;
        one = 1;
      } else {
        sym_swap(A, n, i + 1, r);
        int t = piv[i + 1]; piv[i + 1] = piv[r]; piv[r] = t;
        one = 0;
      }
    }
    if (one) {
      blk[i] = 1;
      double d = A[i * n + i];
      if (d != 0.0) { /* zero pivot is reported by the solve, bkpfacto.c:186,275 */
        for (int j = i + 1; j < n; j++) {
          double l = A[j * n + i] / d;
          for (int k = i + 1; k < n; k++) A[j * n + k] -= l * A[i * n + k];
          A[j * n + i] = l;
        }
        /* keep the trailing block exactly symmetric */
        for (int j = i + 1; j < n; j++)
          for (int k = j + 1; k < n; k++) A[k * n + j] = A[j * n + k];
        for (int j = i + 1; j < n; j++) A[i * n + j] = A[j * n + i];
      }
      i += 1;
    } else {
      blk[i] = 2; blk[i + 1] = 0;
      double a = A[i * n + i], b = A[i * n + i + 1], c = A[(i + 1) * n + i + 1];
      double det = a * c - b * b;
      for (int j = i + 2; j < n; j++) {
        double aj = A[j * n + i], bj = A[j * n + i + 1];
        double s = (c * aj - b * bj) / det;
        double t = (a * bj - b * aj) / det;
        for (int k = i + 2; k < n; k++)
          A[j * n + k] -= s * A[i * n + k] + t * A[(i + 1) * n + k];
        A[j * n + i] = s;
        A[j * n + i + 1] = t;
      }
      for (int j = i + 2; j < n; j++)
        for (int k = j + 1; k < n; k++) A[k * n + j] = A[j * n + k];
      i += 2;
    }
  }
}

/* meschach/bkpfacto.c:230-311; E_SING on a zero 1x1 pivot / zero 2x2 det
 * (:273-286). */
int lqo_bkp_solve(const double *A, int n, const int *piv, const int *blk,
                  const double *b, double *x) {
  double *t = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) t[i] = b[piv[i]];
  /* M t = t : unit lower with 2x2 blocks skipped */
  for (int i = 0; i < n; i++) {
    int first = (blk[i] == 0) ? i - 1 : i; /* second row of a 2x2 block */
    double s = t[i];
    for (int j = 0; j < first; j++) s -= A[i * n + j] * t[j];
    t[i] = s;
  }
  for (int i = 0; i < n;) {
    if (blk[i] == 1) {
      double d = A[i * n + i];
      if (d == 0.0) { free(t); return LQO_E_SING; }
      t[i] /= d;
      i += 1;
    } else {
      double a = A[i * n + i], bb = A[i * n + i + 1], c = A[(i + 1) * n + i + 1];
      double det = a * c - bb * bb;
      if (det == 0.0) { free(t); return LQO_E_SING; }
      double b1 = t[i], b2 = t[i + 1];
      t[i] = (c * b1 - bb * b2) / det;
      t[i + 1] = (a * b2 - bb * b1) / det;
      i += 2;
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    int start = (blk[i] == 2) ? i + 2 : i + 1;
    double s = t[i];
    for (int j = start; j < n; j++) s -= A[j * n + i] * t[j];
    t[i] = s;
  }
  for (int i = 0; i < n; i++) x[piv[i]] = t[i];
  free(t);
  return LQO_OK;
}

/* ---------------------------------------------------------------------- */
static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }
static int *ialloc(size_t n) { return (int *)calloc(n ? n : 1, sizeof(int)); }

lqo_fact *lqo_alloc(const lqo_problem *p) {
  lqo_fact *f = (lqo_fact *)calloc(1, sizeof *f);
  f->p = *p;
  int nx = p->nx, nu = p->nu, K = p->K, nm = nx + nu;
  f->nm = nm;
  f->N = K * nm + nx;
  f->me = K * nx + (p->fixed_x0 ? nx : 0);
  f->H = dalloc((size_t)(K + 1) * nm * nm);
  f->Vxx = dalloc((size_t)(K + 1) * nx * nx);
  f->Rux = dalloc((size_t)K * nu * nx);
  f->Gxu = dalloc((size_t)K * nx * nu);
  f->Guf = dalloc((size_t)K * nu * nu);
  f->piv = ialloc((size_t)K * nu);
  f->blk = ialloc((size_t)K * nu);
  f->sc = dalloc((size_t)K * nu);
  f->V0f = dalloc((size_t)nx * nx);
  f->piv0 = ialloc(nx);
  f->blk0 = ialloc(nx);
  f->sc0 = dalloc(nx);
  f->z = dalloc(p->m);
  f->w = dalloc(p->m);
  f->gx = dalloc((size_t)(K + 1) * nx);
  f->gu = dalloc((size_t)K * nu);
  f->Vx = dalloc((size_t)(K + 1) * nx);
  f->Ru = dalloc((size_t)K * nu);
  f->xs = dalloc((size_t)(K + 1) * nx);
  f->us = dalloc((size_t)K * nu);
  f->ps = dalloc((size_t)K * nx);
  f->t1 = dalloc(f->N); f->t2 = dalloc(f->me); f->t3 = dalloc(p->m); f->t4 = dalloc(p->m);
  f->ex = dalloc(f->N); f->ey = dalloc(f->me); f->ez = dalloc(p->m); f->ew = dalloc(p->m);
  return f;
}

void lqo_free(lqo_fact *f) {
  if (!f) return;
  free(f->H); free(f->Vxx); free(f->Rux); free(f->Gxu); free(f->Guf);
  free(f->piv); free(f->blk); free(f->sc); free(f->V0f); free(f->piv0);
  free(f->blk0); free(f->sc0); free(f->z); free(f->w); free(f->gx); free(f->gu);
  free(f->Vx); free(f->Ru); free(f->xs); free(f->us); free(f->ps);
  free(f->t1); free(f->t2); free(f->t3); free(f->t4);
  free(f->ex); free(f->ey); free(f->ez); free(f->ew);
  free(f);
}

const double *lqo_Vxx(const lqo_fact *f) { return f->Vxx; }
const double *lqo_Rux(const lqo_fact *f) { return f->Rux; }

/* out(nx x nx) = 0.5 (in + in')  -- m_symm, meschach/addon_hqp.c:174-192 */
static void symmetrize(double *A, int n) {
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      double v = 0.5 * (A[i * n + j] + A[j * n + i]);
      A[i * n + j] = A[j * n + i] = v;
    }
}

/* scaled BKP of a symmetric block: sc_i = 1/sqrt(a_ii) if a_ii > 1 else 1
 * (hqp/Hqp_IpLQDOCP.C:1860-1872, 1982-1995) */
static void scaled_bkp(const double *G, int n, double *F, double *sc, int *piv, int *blk) {
  for (int i = 0; i < n; i++) {
    double d = G[i * n + i];
    sc[i] = d > 1.0 ? 1.0 / sqrt(d) : 1.0;
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) F[i * n + j] = sc[i] * G[i * n + j] * sc[j];
  lqo_bkp_factor(F, n, piv, blk);
}

/* x = sc .* F^{-1} (sc .* b)  (hqp/Hqp_IpLQDOCP.C:2045-2047) */
static int scaled_bkp_solve(const double *F, int n, const double *sc, const int *piv,
                            const int *blk, const double *b, double *x) {
  double tmp[n > 0 ? n : 1];
  for (int i = 0; i < n; i++) tmp[i] = sc[i] * b[i];
  int rc = lqo_bkp_solve(F, n, piv, blk, tmp, x);
  if (rc) return rc;
  for (int i = 0; i < n; i++) x[i] *= sc[i];
  return LQO_OK;
}

/* ------------------------------------------------------------------------
 * Hqp_IpLQDOCP::factor (hqp/Hqp_IpLQDOCP.C:796-862): wz = z/w (:810),
 * CTC = C' diag(wz) C on the stage blocks (CTDC :68-103), then
 * ExRiccatiFactorSc (:1794-1999), unconstrained-u branch.
 * ---------------------------------------------------------------------- */
int lqo_factor(lqo_fact *f, const double *z, const double *w) {
  const lqo_problem *p = &f->p;
  const int nx = p->nx, nu = p->nu, K = p->K, nm = f->nm;
  memcpy(f->z, z, sizeof(double) * p->m);
  memcpy(f->w, w, sizeof(double) * p->m);
  memcpy(f->H, p->Q, sizeof(double) * (size_t)(K + 1) * nm * nm);
  for (int r = 0; r < p->m; r++) {
    double wz = z[r] / w[r];
    double *Hk = f->H + (size_t)p->ineq_stage[r] * nm * nm;
    for (int a = p->ineq_ptr[r]; a < p->ineq_ptr[r + 1]; a++)
      for (int b = p->ineq_ptr[r]; b < p->ineq_ptr[r + 1]; b++)
        Hk[p->ineq_lcol[a] * nm + p->ineq_lcol[b]] += wz * p->ineq_val[a] * p->ineq_val[b];
  }
  /* terminal stage (:1800-1804) */
  {
    const double *HK = f->H + (size_t)K * nm * nm;
    double *VK = f->Vxx + (size_t)K * nx * nx;
    for (int i = 0; i < nx; i++)
      for (int j = 0; j < nx; j++) VK[i * nx + j] = HK[i * nm + j];
  }
  double *T = dalloc((size_t)nx * nm);   /* V+ [fx fu] */
  double *Gxx = dalloc((size_t)nx * nx);
  double *Guu = dalloc((size_t)nu * nu);
  double *row = dalloc(nu), *col = dalloc(nu);
  int rc = LQO_OK;
  for (int k = K - 1; k >= 0; k--) {
    const double *Hk = f->H + (size_t)k * nm * nm;
    const double *fx = p->fx + (size_t)k * nx * nx;
    const double *fu = p->fu + (size_t)k * nx * nu;
    const double *Vp = f->Vxx + (size_t)(k + 1) * nx * nx;
    double *Gxu = f->Gxu + (size_t)k * nx * nu;
    /* FormGxx (:1077-1111) */
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) {
        double s = 0.0;
        for (int l = 0; l < nx; l++) s += Vp[i * nx + l] * fx[l * nx + j];
        T[i * nm + j] = s;
      }
      for (int j = 0; j < nu; j++) {
        double s = 0.0;
        for (int l = 0; l < nx; l++) s += Vp[i * nx + l] * fu[l * nu + j];
        T[i * nm + nx + j] = s;
      }
    }
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) {
        double s = 0.0;
        for (int l = 0; l < nx; l++) s += fx[l * nx + i] * T[l * nm + j];
        Gxx[i * nx + j] = Hk[i * nm + j] + s;
      }
      for (int j = 0; j < nu; j++) {
        double s = 0.0;
        for (int l = 0; l < nx; l++) s += fx[l * nx + i] * T[l * nm + nx + j];
        Gxu[i * nu + j] = Hk[i * nm + nx + j] + s;
      }
    }
    for (int i = 0; i < nu; i++)
      for (int j = 0; j < nu; j++) {
        double s = 0.0;
        for (int l = 0; l < nx; l++) s += fu[l * nu + i] * T[l * nm + nx + j];
        Guu[i * nu + j] = Hk[(nx + i) * nm + nx + j] + s;
      }
    symmetrize(Gxx, nx);
    symmetrize(Guu, nu);
    /* unconstrained u (:1854-1882) */
    double *F = f->Guf + (size_t)k * nu * nu;
    double *sc = f->sc + (size_t)k * nu;
    int *piv = f->piv + (size_t)k * nu, *blk = f->blk + (size_t)k * nu;
    scaled_bkp(Guu, nu, F, sc, piv, blk);
    double *Rux = f->Rux + (size_t)k * nu * nx;
    for (int i = 0; i < nx; i++) { /* BKPsolveMT, addon_hqp.c:311-335 */
      for (int j = 0; j < nu; j++) row[j] = Gxu[i * nu + j];
      int e = scaled_bkp_solve(F, nu, sc, piv, blk, row, col);
      if (e) rc = e;
      for (int j = 0; j < nu; j++) Rux[j * nx + i] = col[j];
    }
    /* Vxx = Gxx - Gxu Rux, symmetrised (:1940-1961) */
    double *V = f->Vxx + (size_t)k * nx * nx;
    for (int i = 0; i < nx; i++)
      for (int j = 0; j < nx; j++) {
        double s = 0.0;
        for (int l = 0; l < nu; l++) s += Gxu[i * nu + l] * Rux[l * nx + j];
        V[i * nx + j] = Gxx[i * nx + j] - s;
      }
    symmetrize(V, nx);
  }
  /* initial state (:1971-1996) */
  if (!p->fixed_x0) scaled_bkp(f->Vxx, nx, f->V0f, f->sc0, f->piv0, f->blk0);
  free(T); free(Gxx); free(Guu); free(row); free(col);
  return rc;
}

/* ------------------------------------------------------------------------
 * Hqp_IpLQDOCP::step (hqp/Hqp_IpLQDOCP.C:869-976) + ExRiccatiSolveSc
 * (:2007-2182), unconstrained-u branch.
 * ---------------------------------------------------------------------- */
int lqo_step(lqo_fact *f, const double *r1, const double *r2, const double *r3,
             const double *r4, double *dx, double *dy, double *dz, double *dw) {
  const lqo_problem *p = &f->p;
  const int nx = p->nx, nu = p->nu, K = p->K, nm = f->nm;
  int rc = LQO_OK;
  /* gx,gu = -r1 (+ C'v), v = (z r3 + r4)/w  (:885-921) */
  for (int k = 0; k <= K; k++) {
    for (int i = 0; i < nx; i++) f->gx[k * nx + i] = -r1[k * nm + i];
    if (k < K) for (int i = 0; i < nu; i++) f->gu[k * nu + i] = -r1[k * nm + nx + i];
  }
  for (int r = 0; r < p->m; r++) {
    double v = (f->z[r] * r3[r] + r4[r]) / f->w[r];
    int k = p->ineq_stage[r];
    for (int a = p->ineq_ptr[r]; a < p->ineq_ptr[r + 1]; a++) {
      int lc = p->ineq_lcol[a];
      if (lc < nx) f->gx[k * nx + lc] += p->ineq_val[a] * v;
      else f->gu[k * nu + lc - nx] += p->ineq_val[a] * v;
    }
  }
  const double *fdyn = r2;          /* f[k] = r2 | dynamics rows (:896-897) */
  const double *a0 = r2 + K * nx;   /* a[0] = r2 | x0 rows (:891)           */
  double t[nx > 0 ? nx : 1], Gx[nx > 0 ? nx : 1], Gu[nu > 0 ? nu : 1];
  /* backward (:2016-2094) */
  memcpy(f->Vx + (size_t)K * nx, f->gx + (size_t)K * nx, sizeof(double) * nx);
  for (int k = K - 1; k >= 0; k--) {
    const double *fx = p->fx + (size_t)k * nx * nx;
    const double *fu = p->fu + (size_t)k * nx * nu;
    const double *Vp = f->Vxx + (size_t)(k + 1) * nx * nx;
    const double *Vxp = f->Vx + (size_t)(k + 1) * nx;
    const double *fk = fdyn + (size_t)k * nx;
    for (int i = 0; i < nx; i++) { /* FormGx (:1227-1238) */
      double s = Vxp[i];
      for (int j = 0; j < nx; j++) s += Vp[i * nx + j] * fk[j];
      t[i] = s;
    }
    for (int i = 0; i < nx; i++) {
      double s = f->gx[k * nx + i];
      for (int l = 0; l < nx; l++) s += fx[l * nx + i] * t[l];
      Gx[i] = s;
    }
    for (int i = 0; i < nu; i++) {
      double s = f->gu[k * nu + i];
      for (int l = 0; l < nx; l++) s += fu[l * nu + i] * t[l];
      Gu[i] = s;
    }
    double *Ru = f->Ru + (size_t)k * nu;
    int e = scaled_bkp_solve(f->Guf + (size_t)k * nu * nu, nu, f->sc + (size_t)k * nu,
                             f->piv + (size_t)k * nu, f->blk + (size_t)k * nu, Gu, Ru);
    if (e) rc = e;
    const double *Gxu = f->Gxu + (size_t)k * nx * nu;
    for (int i = 0; i < nx; i++) { /* Vx = Gx - Gxu Ru (:2088-2089) */
      double s = Gx[i];
      for (int l = 0; l < nu; l++) s -= Gxu[i * nu + l] * Ru[l];
      f->Vx[k * nx + i] = s;
    }
  }
  /* initial state (:2096-2119) */
  if (p->fixed_x0) {
    for (int i = 0; i < nx; i++) f->xs[i] = -a0[i];
  } else {
    double rhs[nx > 0 ? nx : 1];
    for (int i = 0; i < nx; i++) rhs[i] = -f->Vx[i];
    int e = scaled_bkp_solve(f->V0f, nx, f->sc0, f->piv0, f->blk0, rhs, f->xs);
    if (e) rc = e;
  }
  /* forward (:2123-2177) */
  for (int k = 0; k < K; k++) {
    const double *fx = p->fx + (size_t)k * nx * nx;
    const double *fu = p->fu + (size_t)k * nx * nu;
    const double *Rux = f->Rux + (size_t)k * nu * nx;
    const double *x = f->xs + (size_t)k * nx;
    double *u = f->us + (size_t)k * nu;
    double *xn = f->xs + (size_t)(k + 1) * nx;
    for (int i = 0; i < nu; i++) {
      double s = f->Ru[k * nu + i];
      for (int j = 0; j < nx; j++) s += Rux[i * nx + j] * x[j];
      u[i] = -s;
    }
    for (int i = 0; i < nx; i++) {
      double s = fdyn[k * nx + i];
      for (int j = 0; j < nx; j++) s += fx[i * nx + j] * x[j];
      for (int j = 0; j < nu; j++) s += fu[i * nu + j] * u[j];
      xn[i] = s;
    }
    const double *Vp = f->Vxx + (size_t)(k + 1) * nx * nx;
    for (int i = 0; i < nx; i++) { /* p = Vxx x + Vx (:2169-2171) */
      double s = f->Vx[(k + 1) * nx + i];
      for (int j = 0; j < nx; j++) s += Vp[i * nx + j] * xn[j];
      f->ps[k * nx + i] = s;
    }
  }
  /* dx = -[x;u], dy = [p; y0] (:938-952); y0 = -(Vx0 + Vxx0 x0) (:2153-2159) */
  for (int k = 0; k <= K; k++) {
    for (int i = 0; i < nx; i++) dx[k * nm + i] = -f->xs[k * nx + i];
    if (k < K) {
      for (int i = 0; i < nu; i++) dx[k * nm + nx + i] = -f->us[k * nu + i];
      for (int i = 0; i < nx; i++) dy[k * nx + i] = f->ps[k * nx + i];
    }
  }
  if (p->fixed_x0)
    for (int i = 0; i < nx; i++) {
      double s = f->Vx[i];
      for (int j = 0; j < nx; j++) s += f->Vxx[i * nx + j] * f->xs[j];
      dy[K * nx + i] = -s;
    }
  /* dw = C dx - r3 ; dz = (r4 - z dw)/w (:955-960) */
  for (int r = 0; r < p->m; r++) {
    int k = p->ineq_stage[r];
    double s = 0.0;
    for (int a = p->ineq_ptr[r]; a < p->ineq_ptr[r + 1]; a++)
      s += p->ineq_val[a] * dx[k * nm + p->ineq_lcol[a]];
    dw[r] = s - r3[r];
    dz[r] = (r4[r] - f->z[r] * dw[r]) / f->w[r];
  }
  return rc;
}

/* ------------------------------------------------------------------------
 * Hqp_IpMatrix::residuum (hqp/Hqp_IpMatrix.C:131-178): inf-norm of
 *   r1 + Q dx - A' dy - C' dz,  r2 - A dx,  r3 - (C dx - dw),
 *   r4 - (z dw + w dz)
 * The four residual vectors are left in t1..t4.
 * ---------------------------------------------------------------------- */
static double residual_vectors(lqo_fact *f, const double *r1, const double *r2,
                               const double *r3, const double *r4, const double *dx,
                               const double *dy, const double *dz, const double *dw) {
  const lqo_problem *p = &f->p;
  const int nx = p->nx, nu = p->nu, K = p->K, nm = f->nm;
  double res = 0.0;
  for (int k = 0; k <= K; k++) {
    int dk = k < K ? nm : nx;
    const double *Qk = p->Q + (size_t)k * nm * nm;
    for (int i = 0; i < dk; i++) {
      double s = r1[k * nm + i];
      for (int j = 0; j < dk; j++) s += Qk[i * nm + j] * dx[k * nm + j];
      f->t1[k * nm + i] = s;
    }
  }
  for (int k = 0; k < K; k++) {
    const double *fx = p->fx + (size_t)k * nx * nx;
    const double *fu = p->fu + (size_t)k * nx * nu;
    for (int i = 0; i < nx; i++) {
      double y = dy[k * nx + i];
      double s = -dx[(k + 1) * nm + i];
      for (int j = 0; j < nx; j++) {
        s += fx[i * nx + j] * dx[k * nm + j];
        f->t1[k * nm + j] -= fx[i * nx + j] * y;
      }
      for (int j = 0; j < nu; j++) {
        s += fu[i * nu + j] * dx[k * nm + nx + j];
        f->t1[k * nm + nx + j] -= fu[i * nu + j] * y;
      }
      f->t1[(k + 1) * nm + i] += y;
      f->t2[k * nx + i] = r2[k * nx + i] - s;
    }
  }
  if (p->fixed_x0)
    for (int i = 0; i < nx; i++) {
      f->t1[i] -= dy[K * nx + i];
      f->t2[K * nx + i] = r2[K * nx + i] - dx[i];
    }
  for (int r = 0; r < p->m; r++) {
    int k = p->ineq_stage[r];
    double s = 0.0;
    for (int a = p->ineq_ptr[r]; a < p->ineq_ptr[r + 1]; a++) {
      s += p->ineq_val[a] * dx[k * nm + p->ineq_lcol[a]];
      f->t1[k * nm + p->ineq_lcol[a]] -= p->ineq_val[a] * dz[r];
    }
    f->t3[r] = r3[r] - (s - dw[r]);
    f->t4[r] = r4[r] - (f->z[r] * dw[r] + f->w[r] * dz[r]);
  }
  for (int i = 0; i < f->N; i++) if (fabs(f->t1[i]) > res) res = fabs(f->t1[i]);
  for (int i = 0; i < f->me; i++) if (fabs(f->t2[i]) > res) res = fabs(f->t2[i]);
  for (int i = 0; i < p->m; i++) if (fabs(f->t3[i]) > res) res = fabs(f->t3[i]);
  for (int i = 0; i < p->m; i++) if (fabs(f->t4[i]) > res) res = fabs(f->t4[i]);
  return res;
}

double lqo_residuum(lqo_fact *f, const double *r1, const double *r2, const double *r3,
                    const double *r4, const double *dx, const double *dy,
                    const double *dz, const double *dw) {
  return residual_vectors(f, r1, r2, r3, r4, dx, dy, dz, dw);
}

/* Hqp_IpMatrix::solve (hqp/Hqp_IpMatrix.C:65-128) */
int lqo_solve(lqo_fact *f, double eps, const double *r1, const double *r2,
              const double *r3, const double *r4, double *dx, double *dy, double *dz,
              double *dw, double *res_out, int *nsteps) {
  const int N = f->N, me = f->me, m = f->p.m;
  int steps = 1;
  int rc = lqo_step(f, r1, r2, r3, r4, dx, dy, dz, dw);
  double res = residual_vectors(f, r1, r2, r3, r4, dx, dy, dz, dw);
  double *s1 = dalloc(N), *s2 = dalloc(me), *s3 = dalloc(m), *s4 = dalloc(m);
  for (int it = 0; it < 5 && res > eps && !rc; it++) {
    double res_last = res;
    memcpy(s1, f->t1, sizeof(double) * N);
    memcpy(s2, f->t2, sizeof(double) * me);
    memcpy(s3, f->t3, sizeof(double) * m);
    memcpy(s4, f->t4, sizeof(double) * m);
    rc = lqo_step(f, s1, s2, s3, s4, f->ex, f->ey, f->ez, f->ew);
    steps++;
    double alpha = 1.0;
    do {
      for (int i = 0; i < N; i++) dx[i] += alpha * f->ex[i];
      for (int i = 0; i < me; i++) dy[i] += alpha * f->ey[i];
      for (int i = 0; i < m; i++) { dz[i] += alpha * f->ez[i]; dw[i] += alpha * f->ew[i]; }
      res = residual_vectors(f, r1, r2, r3, r4, dx, dy, dz, dw);
      if (res > res_last) {
        for (int i = 0; i < N; i++) dx[i] -= alpha * f->ex[i];
        for (int i = 0; i < me; i++) dy[i] -= alpha * f->ey[i];
        for (int i = 0; i < m; i++) { dz[i] -= alpha * f->ez[i]; dw[i] -= alpha * f->ew[i]; }
        alpha -= 0.3;
      }
    } while (res > res_last && alpha > 0.0);
    if (alpha <= 0.0) break;
  }
  free(s1); free(s2); free(s3); free(s4);
  if (res_out) *res_out = res;
  if (nsteps) *nsteps = steps;
  return rc;
}
