/*
 * Plain-C restatement of the reference's LQ-DOCP KKT path (CPU oracle).
 *
 * TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this.  The product (hqp_b200/) never
 * links or calls it.
 *
 * What is restated (reference file:line in lqdocp_oracle.c next to each
 * function): Hqp_IpLQDOCP::factor/step with the default ExRiccatiFactorSc /
 * ExRiccatiSolveSc sweeps for stages WITHOUT stage equality constraints (the
 * "S.n == 0" branch, hqp/Hqp_IpLQDOCP.C:1854-1882, 2040-2048), fixed or free
 * initial state, Meschach's Bunch-Kaufman-Parlett factor/solve
 * (meschach/bkpfacto.c:102-311), and Hqp_IpMatrix::residuum / ::solve
 * (hqp/Hqp_IpMatrix.C:65-178).
 * Not restated: the GE_QP null-space branch for stage equality constraints
 * (hqp/Hqp_IpLQDOCP.C:1883-1938) -- parity for those inputs is checked against
 * the compiled reference itself (oracle/_ref) and the fixtures generated from
 * it (tests/golden/).
 *
 * Parity pin: tests/test_oracle.py compares every entry point with the
 * compiled reference (oracle/_ref/libhqpharness.so) when present and with the
 * committed fixtures in tests/golden/ always.
 *
 * Layout = the C ABI's (include/hqp_ipcuda.h): variables [x0,u0,...,xK];
 * equality rows: K*nx dynamics rows, then nx x0-fixing rows when fixed_x0.
 */
#ifndef LQDOCP_ORACLE_H
#define LQDOCP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int nx, nu, K;
  int fixed_x0;
  int m;                  /* inequality rows of C */
  const int *ineq_stage;  /* [m]   stage of row i */
  const int *ineq_ptr;    /* [m+1] CSR row pointers */
  const int *ineq_lcol;   /* local column in [0, nx+nu) of the row's stage */
  const double *ineq_val;
  const double *Q;        /* (K+1) x (nx+nu)^2 full symmetric blocks */
  const double *fx;       /* K x nx x nx */
  const double *fu;       /* K x nx x nu */
} lqo_problem;

typedef struct lqo_fact lqo_fact;

#define LQO_OK 0
#define LQO_E_SING 4 /* Meschach E_SING, meschach/err.h:88 */

lqo_fact *lqo_alloc(const lqo_problem *p);
void lqo_free(lqo_fact *f);

/* Hqp_IpLQDOCP::factor */
int lqo_factor(lqo_fact *f, const double *z, const double *w);
/* Hqp_IpLQDOCP::step */
int lqo_step(lqo_fact *f, const double *r1, const double *r2, const double *r3,
             const double *r4, double *dx, double *dy, double *dz, double *dw);
/* Hqp_IpMatrix::residuum */
double lqo_residuum(lqo_fact *f, const double *r1, const double *r2,
                    const double *r3, const double *r4, const double *dx,
                    const double *dy, const double *dz, const double *dw);
/* Hqp_IpMatrix::solve: step + <=5 damped refinement steps until res <= eps */
int lqo_solve(lqo_fact *f, double eps, const double *r1, const double *r2,
              const double *r3, const double *r4, double *dx, double *dy,
              double *dz, double *dw, double *res, int *nsteps);

/* read-only views for tests: Vxx [(K+1) nx nx], Rux [K nu nx] */
const double *lqo_Vxx(const lqo_fact *f);
const double *lqo_Rux(const lqo_fact *f);

/* dense Bunch-Kaufman-Parlett exposed for unit tests (n x n row-major, in
 * place; piv/blk length n) */
void lqo_bkp_factor(double *A, int n, int *piv, int *blk);
int lqo_bkp_solve(const double *A, int n, const int *piv, const int *blk,
                  const double *b, double *x);

#ifdef __cplusplus
}
#endif
#endif
