/*
 * hqp_docpcuda.h -- C ABI of libhqpdocp.so: the stage loop of Hqp_Docp::update /
 * Hqp_Docp::update_fbd on an NVIDIA B200.
 *
 * What it replaces.  Once per SQP iteration (update) and once per line-search trial
 * (update_fbd) the reference walks all stages k = 0..K on the host
 *   Hqp_Docp::update_fbd     hqp/Hqp_Docp.C:831-891   values: f_k - x_{k+1} -> qp->b,
 *                                                     sum f0_k -> f, c_k -> qp->b / qp->d
 *   Hqp_Docp::update         hqp/Hqp_Docp.C:944-1075  values + the stage derivatives
 *                                                     fx fu f0x f0u cx cu -> qp->A / qp->C / qp->c
 *   Hqp_Docp::update_grds    hqp/Hqp_Docp.C:1097-1180 default derivatives: forward differences,
 *                                                     dv = 1e-4 |v| + 1e-6, nx+nu+1 model calls a stage
 *   Hqp_Docp::update_bounds  hqp/Hqp_Docp.C:893-940   x - x_fix -> qp->b; x - lb, ub - x -> qp->d
 * calling the model's update_vals() per stage (an "omp parallel for" over ncpu threads).
 * Here the model is a DEVICE function (hqp_b200/csrc/docp_models.cuh; a model is one struct
 * with a vals<T>() template, written once) and one call evaluates all stages: one thread per
 * stage for the values, one thread per (stage, variable) for the derivative columns.
 *
 * Derivatives: HQPDOCP_GRAD_FD repeats Hqp_Docp::update_grds' forward differences operation by
 * operation (same increments, same quotients); HQPDOCP_GRAD_AD evaluates the same vals<T>() on
 * forward-mode dual numbers -- exact first derivatives at the cost of the differences, the role
 * ADOL-C plays for hqp/Hqp_DocpAdol.C (not buildable here: ADOL-C is absent).  The Lagrangian
 * Hessian is NOT touched: Hqp_Docp::update_hela's default is empty (hqp/Hqp_Docp.C:1183-1189)
 * and the SQP solver's own approximation fills Q (row f2, include/hqp_hlcuda.h), so the
 * multipliers y, z of Hqp_Docp::update(y, z) have no role here.
 *
 * Layout.  Uniform stages: nx states at k = 0..K, nu controls and nc constraints at
 * k = 0..K-1, ncK constraints at k = K (Hqp_Docp asserts nus[K] == 0, hqp/Hqp_Docp.C:543).
 *   x   [N = K (nx+nu) + nx]             [x0 u0 x1 u1 ... xK]          (hqp/Hqp_Docp.C:506-531)
 *   b   [me = K nx + n_xu_eq + n_cns_eq] dynamics rows, then x/u fixings, then constraint
 *                                        equalities                     (:573-576, :590-600)
 *   d   [m = n_xu_lb + n_xu_ub + n_cns_lb + n_cns_ub]                   (:578-582)
 *   g   [N]  objective gradient f0x, f0u per stage (qp->c)
 *   fx  [K][nx*nx], fu [K][nx*nu]        row-major, the layout hqpcu_update_dev /
 *                                        hqpcu_update_stages_dev take (include/hqp_ipcuda.h)
 *   cx  [K*nc + ncK][nx], cu [K*nc][nu]  rows of the stage constraints (global constraint index)
 * The six association tables are Hqp_DocpAssoc's idxs / vals as parse_constr builds them
 * (hqp/Hqp_Docp.C:370-397): x/u tables index x, constraint tables index the global constraint
 * number k nc + i.  Periodical states (:936-940) are not supported.
 *
 * Plain C, FP64, every call returns a status; host pointers unless the name ends in _dev.
 * No CPU fallback: without a CUDA device the calls return HQPDOCP_E_CUDA.
 */
#ifndef HQP_DOCPCUDA_H
#define HQP_DOCPCUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define HQPDOCP_OK 0
#define HQPDOCP_E_ARG 1
#define HQPDOCP_E_UNSUPPORTED 2 /* unknown model, or dimensions outside the model's / kernel's range */
#define HQPDOCP_E_CUDA 100      /* CUDA runtime error; see hqpdocp_last_error() */

/* device models (hqp_b200/csrc/docp_models.cuh) */
#define HQPDOCP_MODEL_DID 0     /* hqp_docp/Prg_DID.C:78-98: double integrator, nx 2, nu 1, nc 0|1;
                                   par = { dt }                                                    */
#define HQPDOCP_MODEL_SYNTHNL 1 /* the synthetic SQP-driven workload (config 5's model): dense linear
                                   part + rational state nonlinearity, tracking objective with an x-u
                                   cross term, one quadratic path constraint;
                                   par = { eps, A[nx*nx], B[nx*nu], qw[nx], rw[nu] }, spar[k] = r_k[nx] */

#define HQPDOCP_GRAD_FD 0 /* Hqp_Docp::update_grds, hqp/Hqp_Docp.C:1097-1180 */
#define HQPDOCP_GRAD_AD 1 /* forward-mode dual numbers on the same model code */

typedef struct hqpdocp_handle hqpdocp_handle;

typedef struct hqpdocp_assoc { /* Hqp_DocpAssoc (hqp/Hqp_Docp.C:36-58): dim entries */
  int dim;
  const int *idxs;    /* index into x (x/u tables) or global constraint index (constraint tables) */
  const double *vals; /* the fixed value / bound */
} hqpdocp_assoc;

typedef struct hqpdocp_dims {
  int K;              /* stages with controls; states x_0 .. x_K                */
  int nx, nu;         /* uniform stage dimensions                                */
  int nc, ncK;        /* constraints per stage k < K, and at k = K               */
  int model;          /* HQPDOCP_MODEL_*                                         */
  int npar;           /* global model parameters                                 */
  const double *par;  /* [npar]                                                  */
  int nspar;          /* per-stage model parameters                              */
  const double *spar; /* [(K+1)][nspar], may be NULL if nspar == 0               */
  hqpdocp_assoc xu_eq, xu_lb, xu_ub;    /* _xu_eq / _xu_lb / _xu_ub              */
  hqpdocp_assoc cns_eq, cns_lb, cns_ub; /* _cns_eq / _cns_lb / _cns_ub           */
  int device;         /* CUDA device ordinal                                     */
  /* A handle may own a contiguous RANGE of the stages of a longer horizon (one handle per GPU,
   * the same split as hqpcu_comm_init's, include/hqp_ipcuda.h): K_total > 0 is the horizon's K
   * and k_first the global index of this handle's first stage; the handle then evaluates the
   * global stages k_first .. k_first+K-1 and, only if k_first+K == K_total, the final stage.
   * x still holds K (nx+nu) + nx values: x_{k_first+K} is read for the last dynamics rows (a halo
   * of nx values when the stage belongs to the next range).  spar, the tables and all outputs
   * are those of the range, with LOCAL indices.  No exchange between ranges is needed: the
   * objective is the sum of the ranges' f.  K_total = 0: the whole horizon (k_first = 0).      */
  int k_first, K_total;
} hqpdocp_dims;

const char *hqpdocp_last_error(void);
int hqpdocp_create(const hqpdocp_dims *dims, hqpdocp_handle **out);
int hqpdocp_destroy(hqpdocp_handle *h);
int hqpdocp_set_stream(hqpdocp_handle *h, void *cuda_stream);
/* N, me, m as laid out above */
int hqpdocp_sizes(const hqpdocp_handle *h, long long *N, long long *me, long long *m);
/* kernels launched by this handle since creation */
long long hqpdocp_launch_count(const hqpdocp_handle *h);

/* Hqp_Docp::update_fbd (:831-891): x [N] in; f (the sum of the stage objectives, added in
 * stage order), b [me], d [m] out. */
int hqpdocp_update_fbd(hqpdocp_handle *h, const double *x, double *f, double *b, double *d);
/* Hqp_Docp::update (:944-1075): additionally g [N], fx, fu, cx, cu (cx / cu may be NULL when
 * nc == ncK == 0). */
int hqpdocp_update(hqpdocp_handle *h, int grad_mode, const double *x, double *f, double *b,
                   double *d, double *g, double *fx, double *fu, double *cx, double *cu);

/* The same on DEVICE pointers, asynchronous on the handle's stream; f: one double on the
 * device.  fx / fu can be the arrays handed to hqpcu_update_dev next: nothing crosses PCIe. */
int hqpdocp_update_fbd_dev(hqpdocp_handle *h, const double *x, double *f, double *b, double *d);
int hqpdocp_update_dev(hqpdocp_handle *h, int grad_mode, const double *x, double *f, double *b,
                       double *d, double *g, double *fx, double *fu, double *cx, double *cu);

#ifdef __cplusplus
}
#endif
#endif
