/*
 * hqp_ipcuda.h -- C ABI of libhqpcuda.so, the B200 (sm_100a) KKT engine behind
 * HQP's Hqp_IpMatrix plugin interface.
 *
 * Drop-in boundary: the reference's interior-point solvers call a
 * matrix module through  init / update / factor / step / solve / residuum
 * (hqp/Hqp_IpMatrix.h:63-88).  The host module hqp_b200/host/Hqp_IpCuda.C
 * implements that C++ interface and forwards every call to the functions below,
 * exactly as hqp/Hqp_IpPARDISO.C:161-166 forwards to hqp/pardiso_wrapper.h:32-48.
 *
 * Conventions
 *  - plain C: pointers, ints, doubles; no C++ types, no exceptions.
 *  - every function returns an int status (HQPCU_OK = 0).  HQPCU_E_SING = 4 is
 *    Meschach's E_SING (meschach/err.h:88): the host module turns it into
 *    m_error(E_SING, ...) AFTER the call returned, so no longjmp crosses a CUDA
 *    call in flight ("Failure detection").
 *  - all arithmetic is FP64.  Vectors use the reference's QP layout
 *    (hqp/Hqp_Docp.C:465-755):
 *        x  (N  = K*(nx+nu)+nx)      [x0,u0,x1,u1,...,xK]
 *        y  (me = K*nx [+nx] +n_eq)  dynamics rows, then the nx rows fixing x0
 *                                    (if fixed_x0), then general equality rows
 *        z,w (m = n_ineq)            rows of C in the caller's order
 *    With batch > 1 every vector / slab is the concatenation of `batch`
 *    instances (instance-major).
 *  - The linear system solved (hqp/Hqp_IpsMehrotra.C:27-31):
 *        -Q dx + A' dy + C' dz      = r1
 *         A dx                      = r2
 *         C dx              - dw    = r3
 *                 W dz    + Z dw    = r4
 *  - "_dev" entry points take DEVICE pointers and enqueue on the handle's
 *    stream without synchronising; the others take HOST pointers, copy in/out
 *    and return after the results are on the host.
 */
#ifndef HQP_IPCUDA_H
#define HQP_IPCUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define HQPCU_OK 0
#define HQPCU_E_SIZES 1   /* inconsistent dimensions (Meschach E_SIZES)          */
#define HQPCU_E_SING 4    /* zero pivot in a stage LDL^T / LU (Meschach E_SING)  */
#define HQPCU_E_NULL 8    /* NULL argument (Meschach E_NULL)                     */
#define HQPCU_E_CUDA 100  /* CUDA runtime error; see hqpcu_last_error()          */
#define HQPCU_E_UNSUPPORTED 101 /* structure outside the supported scope         */
#define HQPCU_E_NOTPD 102 /* parallel-in-time factor met a non-positive pivot at
                             a segment end; caller may retry with nseg = 1       */

typedef struct hqpcu_handle hqpcu_handle;

/* Stage structure of the QP: what Hqp_IpLQDOCP::init derives from the sparsity
 * of A and C (Get_Dim hqp/Hqp_IpLQDOCP.C:201-287, Get_Constr_Dim :368-407,
 * Check_Structure :298-354). */
typedef struct hqpcu_dims {
  int K;        /* stages with controls (_kmax); states x_0..x_K               */
  int nx, nu;   /* uniform stage dimensions (_nk[k], _mk[k])                    */
  int batch;    /* independent instances sharing this structure (>= 1)          */
  int fixed_x0; /* 1: x_0 is fixed by nx identity rows (_fixed_x0, :344-351)    */
  int n_ineq;   /* rows of C per instance                                       */
  const int *ineq_stage; /* [n_ineq]   stage of each row (_rcki, :391-406)      */
  const int *ineq_ptr;   /* [n_ineq+1] CSR row pointers                         */
  const int *ineq_lcol;  /* [nnz]      column inside the stage block, 0..nx+nu-1*/
  int n_eq;     /* general stage equality rows per instance (_raki, :381-388)   */
  const int *eq_stage;   /* [n_eq]                                              */
  const int *eq_ptr;     /* [n_eq+1]                                            */
  const int *eq_lcol;    /* [nnz_eq]                                            */
  int device;   /* CUDA device ordinal                                          */
  int nseg;     /* horizon segments per instance for the parallel-in-time
                   factor/solve; 0 = choose automatically, 1 = sequential sweep */
  int ngpu;     /* 0 / 1: one GPU.  N > 1: the handle is a dispatcher over the N
                   devices device .. device+N-1 of THIS process (Hqp_IpCuda's
                   mat_ngpu): the horizon is split into N contiguous stage ranges,
                   one worker thread and one NCCL rank per GPU; the host-pointer
                   entry points (update, factor, step, solve, residuum,
                   mehrotra_*, franke_solve) take and return FULL-length vectors.
                   batch == 1, no general equality rows, dense update only.     */
} hqpcu_dims;

/* --- life cycle (Hqp_IpCuda ctor/dtor + init; If_Module deletes and re-creates
 *     the module object on every "qp_mat_solver" write, iftcl/If_Module.h:66-90) */
int hqpcu_create(const hqpcu_dims *dims, hqpcu_handle **out);
int hqpcu_destroy(hqpcu_handle *h);
/* cudaStream_t to enqueue on (NULL = default stream) */
int hqpcu_set_stream(hqpcu_handle *h, void *cuda_stream);
const char *hqpcu_last_error(void);
/* kernels launched by this handle since creation (bench.py "gpu_launches") */
long long hqpcu_launch_count(const hqpcu_handle *h);
/* refined solves (hqpcu_solve*, and the two per IP iteration inside
 * hqpcu_mehrotra_*) since creation and the steps they took: steps / solves =
 * mean number of KKT solves per Hqp_IpMatrix::solve (1 = no refinement needed,
 * hqp/Hqp_IpMatrix.C:65-128)                                                   */
int hqpcu_solve_stats(const hqpcu_handle *h, long long *solves, long long *steps);
/* segments actually used per instance */
int hqpcu_nseg(const hqpcu_handle *h);

/* --- update: once per SQP iteration (Hqp_IpLQDOCP::update, :722-787) -------
 * Q   [batch][(K+1)][(nx+nu)^2] full symmetric stage blocks (block K uses its
 *                               leading nx x nx part)
 * fx  [batch][K][nx*nx], fu [batch][K][nx*nu]  row-major dynamics Jacobians
 * ineq_val [batch][nnz], eq_val [batch][nnz_eq] values on the CSR patterns     */
int hqpcu_update(hqpcu_handle *h, const double *Q, const double *fx,
                 const double *fu, const double *ineq_val, const double *eq_val);
int hqpcu_update_dev(hqpcu_handle *h, const double *Q, const double *fx,
                     const double *fu, const double *ineq_val,
                     const double *eq_val);

/* update of a stage window / of the inequality values from DEVICE memory (a host
 * that builds its matrices on the GPU, or a 10^6-stage horizon produced chunk by
 * chunk): nq blocks of Q for stages k0 .. k0+nq-1 (<= K), nf blocks of fx and fu
 * for stages k0 .. k0+nf-1 (< K).  batch == 1.                                  */
int hqpcu_update_stages_dev(hqpcu_handle *h, int k0, int nq, const double *Q, int nf,
                            const double *fx, const double *fu);
int hqpcu_update_ineq_dev(hqpcu_handle *h, const double *ineq_val);

/* --- update from sparse values -----------------------------
 * Replaces the host-side walk that writes every SPMAT entry into a dense
 * stage block (sp_extract_mat_iter in Hqp_IpLQDOCP::update,
 * hqp/Hqp_IpLQDOCP.C:747-755; meschach/addon2_hqp.c:649-709) by a device-side
 * scatter: the caller registers ONCE (the sparsity pattern is fixed between
 * init() calls) where each stored entry of Q and of the dynamics rows of A goes,
 * then uploads only the value array per SQP iteration.
 *   dst[i]  position of value i in the virtual slab [ Q | fx | fu ] of one
 *           instance (offsets 0, (K+1) nm^2, (K+1) nm^2 + K nx^2)
 *   dst2[i] second position written with the same value (the mirrored entry of
 *           the symmetric Q block), or -1
 * hqpcu_update_values: vals [n] (host); the slabs are zeroed, then filled by
 * one kernel.  batch == 1.                                                      */
int hqpcu_set_value_map(hqpcu_handle *h, long long n, const long long *dst,
                        const long long *dst2);
int hqpcu_update_values(hqpcu_handle *h, const double *vals,
                        const double *ineq_val, const double *eq_val);

/* --- factor: once per IP iteration (Hqp_IpLQDOCP::factor, :796-862) --------- */
int hqpcu_factor(hqpcu_handle *h, const double *z, const double *w);
int hqpcu_factor_dev(hqpcu_handle *h, const double *z, const double *w);

/* status of everything enqueued so far by the _dev entry points; synchronises
 * the stream.  HQPCU_E_NOTPD after hqpcu_factor_dev: call hqpcu_set_nseg(h, 1)
 * and factor again (hqpcu_factor does this by itself).                         */
int hqpcu_sync_status(hqpcu_handle *h);
/* change the number of horizon segments (<= the count chosen at create) */
int hqpcu_set_nseg(hqpcu_handle *h, int nseg);

/* --- step: one KKT solve with the current factor (Hqp_IpLQDOCP::step,
 *     :869-976).  z,w are the ones given to the last factor call.              */
int hqpcu_step(hqpcu_handle *h, const double *r1, const double *r2,
               const double *r3, const double *r4, double *dx, double *dy,
               double *dz, double *dw);
int hqpcu_step_dev(hqpcu_handle *h, const double *r1, const double *r2,
                   const double *r3, const double *r4, double *dx, double *dy,
                   double *dz, double *dw);

/* --- residuum: inf-norm of the four block residuals
 *     (Hqp_IpMatrix::residuum, hqp/Hqp_IpMatrix.C:131-178)                      */
int hqpcu_residuum(hqpcu_handle *h, const double *r1, const double *r2,
                   const double *r3, const double *r4, const double *dx,
                   const double *dy, const double *dz, const double *dw,
                   double *res);
int hqpcu_residuum_dev(hqpcu_handle *h, const double *r1, const double *r2,
                       const double *r3, const double *r4, const double *dx,
                       const double *dy, const double *dz, const double *dw,
                       double *res /* host */);

/* --- solve: step + up to 5 damped refinement steps until res <= eps
 *     (Hqp_IpMatrix::solve, hqp/Hqp_IpMatrix.C:65-128).  nsteps (may be NULL)
 *     receives the number of step() calls performed.                           */
int hqpcu_solve(hqpcu_handle *h, double eps, const double *r1, const double *r2,
                const double *r3, const double *r4, double *dx, double *dy,
                double *dz, double *dw, double *res, int *nsteps);
int hqpcu_solve_dev(hqpcu_handle *h, double eps, const double *r1,
                    const double *r2, const double *r3, const double *r4,
                    double *dx, double *dy, double *dz, double *dw,
                    double *res /* host */, int *nsteps);

/* --- horizon split across GPUs.  The handle owns a contiguous
 *     stage range of a longer horizon (created with the LOCAL K; a range that is
 *     followed by another passes a zero terminal block and no rows at stage K;
 *     a range that is preceded by another has fixed_x0 = 0 and no x0 rows).
 *     All pointers are DEVICE pointers; the caller performs the all-gathers
 *     (NCCL) between the phases:
 *       factor_begin  -> xf [4*nx*nx]  : (A, C, J) of the range, terminal block
 *         all-gather xf  -> gathered [world][4*nx*nx]
 *       factor_finish -> xpsi [nx*nx]  : transition of the range
 *         all-gather xpsi -> gpsi [world][nx*nx]          (once per factor)
 *       step_begin    -> xv [2*nx]     : (v0, v_term)
 *         all-gather xv  -> gv [world][2*nx]
 *       step_mid      -> xx [2*nx]     : (x0, x_start)
 *         all-gather xx  -> gx [world][2*nx]
 *       step_finish   -> dx, dy, dz, dw of the range                          */
int hqpcu_range_config(hqpcu_handle *h, int has_prev, int has_next);
int hqpcu_range_factor_begin(hqpcu_handle *h, const double *z, const double *w, double *xf);
int hqpcu_range_factor_finish(hqpcu_handle *h, const double *gathered, int rank, int world,
                              double *xpsi);
int hqpcu_range_step_begin(hqpcu_handle *h, const double *r1, const double *r2,
                           const double *r3, const double *r4, double *xv);
int hqpcu_range_step_mid(hqpcu_handle *h, const double *gv, const double *gpsi, int rank,
                         int world, double *xx);
int hqpcu_range_step_finish(hqpcu_handle *h, const double *gx, const double *gpsi, int rank,
                            int world, double *dx, double *dy, double *dz, double *dw);

/* --- horizon split with the exchanges INSIDE the library.
 *     hqpcu_comm_init turns the handle into rank `rank` of `world` contiguous stage
 *     ranges (same local-K / local-row conventions as hqpcu_range_config) and
 *     creates an NCCL communicator from a 128-byte unique id that the caller
 *     distributes (rank 0: hqpcu_comm_unique_id, then MPI / torch.distributed /
 *     a file).  NCCL is bound at run time (dlopen of libnccl.so.2; an NCCL already
 *     in the process is shared); without it these two calls return
 *     HQPCU_E_UNSUPPORTED and everything else keeps working on one GPU.
 *     Afterwards EVERY entry point above (update, factor, step, solve, residuum,
 *     mehrotra_solve, host or _dev flavour) works on the range's local slices and
 *     must be called by all ranks together: the all-gathers of the boundary
 *     elements / vectors and the all-reduces of the residual norm and of the IP
 *     scalars (sum: mu, gap, dots; min: step lengths; max: norms --
 *     hqp/Hqp_IpsMehrotra.C:425-465, 566-574, 627-681; hqp/Hqp_IpMatrix.C:131-178)
 *     are issued on the handle's stream as nodes of the same CUDA graphs as the
 *     kernels.  A singular / non-PD status is merged over the ranks, so all ranks
 *     take the same branch.  world == 1 is allowed (no NCCL needed).
 *     batch == 1, no general equality rows; a non-positive pivot at a segment end
 *     returns HQPCU_E_NOTPD (no sequential fallback across ranks).              */
int hqpcu_comm_unique_id(unsigned char *id128);
int hqpcu_comm_init(hqpcu_handle *h, const unsigned char *id128, int rank, int world);
/* rank, world and the number of inequality rows on the whole horizon (any may be NULL) */
int hqpcu_comm_info(const hqpcu_handle *h, int *rank, int *world, long long *m_global);

/* --- device-resident interior-point solve: Hqp_IpsMehrotra::cold_start + ::solve
 *     (hqp/Hqp_IpsMehrotra.C:209-327, 355-733; predictor-corrector with Terlaky's
 *     safeguard, qp_init_method 0) of
 *         min 1/2 x'Qx + c'x   s.t.  A x + b = 0,  C x + d >= 0
 *     for the Q, A, C given to hqpcu_update.  The residual / complementarity /
 *     step-length vector passes run as fused kernels on vectors that stay in
 *     HBM; only scalars return to the host between iterations.  batch == 1.
 *     result: Hqp_Result (hqp/Hqp_impl.h:37-43) 0 optimal, 3 suboptimal,
 *     4 degenerate.  max_iters <= 0: the reference default (200).  w, gap may
 *     be NULL.                                                                 */
int hqpcu_mehrotra_solve(hqpcu_handle *h, const double *c, const double *b,
                         const double *d, double eps, int max_iters, double *x,
                         double *y, double *z, double *w, int *iters, int *result,
                         double *gap);
/* Hqp_IpsMehrotra::hot_start + solve (hqp/Hqp_IpsMehrotra.C:330-352, 696-733;
 * called by Hqp_SqpSolver.C:288-293 for every QP after the first): x, y are
 * IN/OUT (the previous solution), z and w restart from the iterate this handle
 * remembered during its previous solve (:475-478).  A hot start that does not
 * contract (test > test1 / 1.2^(iter-1)), exceeds max_warm_iters (0: 25) or
 * ends without "optimal" is restarted cold inside the call; the iterations it
 * cost are included in *iters like the reference's _fail_iters.  Without a
 * previous solve this is a cold-started solve.                                 */
int hqpcu_mehrotra_hot_solve(hqpcu_handle *h, const double *c, const double *b,
                             const double *d, double eps, int max_iters,
                             int max_warm_iters, double *x, double *y, double *z,
                             double *w, int *iters, int *result, double *gap);

/* --- device-resident Franke interior-point solve: Hqp_IpsFranke::cold_start /
 *     hot_start + ::solve (hqp/Hqp_IpsFranke.C:157-417), the default QP solver of
 *     the shipped docp example (hqp/Hqp_SqpSolver.C:67).  Same problem and vector
 *     conventions as hqpcu_mehrotra_solve.  hot = 0: cold start from x = y = 0;
 *     hot != 0: x, y, z, w are IN/OUT (the previous iterate, :226-268), with the
 *     reference's cold restarts when the gap grows.  beta <= 0: 0.995 (qp_beta),
 *     mu0: qp_mu0 (0 = Wright's Ltilde), max_warm_iters <= 0: 15.  result:
 *     Hqp_Result 0 optimal, 1 feasible, 2 infeasible, 3 suboptimal, 4 degenerate. */
int hqpcu_franke_solve(hqpcu_handle *h, const double *c, const double *b, const double *d,
                       double eps, int max_iters, int hot, int max_warm_iters, double beta,
                       double mu0, double *x, double *y, double *z, double *w, int *iters,
                       int *result, double *gap);

/* --- per-kernel timing with CUDA events on the launching stream (bench.py's
 *     roofline section).  hqpcu_profile_read synchronises and writes a JSON
 *     object {"kernel": {"ms": total, "n": launches}, ...} for everything
 *     launched since the previous read.                                         */
/* ---- SQP-level vector operations on the device-resident QP --
 * What Hqp_SqpSolver / Hqp_SqpPowell compute around every QP solve with sparse mat-vecs
 * over the whole horizon (hqp/Hqp_SqpSolver.C:155-174, 225-226, 258-259, 299-301, 430-445;
 * hqp/Hqp_SqpPowell.C:189-244), from the Q / A / C values the last hqpcu_update left in
 * HBM.  Vectors in the ABI's layout (the one hqpcu_step uses), host pointers.
 *   grd_L = c - A'y - C'z
 *   merit: out8 = { phi, phi1, s'Qs, c's, ||b||inf, ||min(d,0)||inf, sum re|As+b|,
 *                   -sum r min(0, Cs+d) }  with  phi  = f + sum re|b| - sum r min(0,d),
 *                                                 phi1 = f + c's + out8[6] + out8[7]
 *   quad:  x'Qx
 * One instance, one GPU (no batch, no split horizon). */
int hqpcu_sqp_grd_L(hqpcu_handle *h, const double *c, const double *y, const double *z,
                    double *grd_L);
int hqpcu_sqp_merit(hqpcu_handle *h, double f, const double *c, const double *s, const double *b,
                    const double *d, const double *re, const double *r, double *out8);
int hqpcu_sqp_quad(hqpcu_handle *h, const double *x, double *xQx);

int hqpcu_profile(hqpcu_handle *h, int on);
int hqpcu_profile_read(hqpcu_handle *h, char *buf, int len);

/* --- read-back of factor state for tests: Vxx [batch][(K+1)][nx*nx],
 *     Rux [batch][K][nu*nx] (either may be NULL)                               */
int hqpcu_get_factor(hqpcu_handle *h, double *Vxx, double *Rux);

#ifdef __cplusplus
}
#endif
#endif
