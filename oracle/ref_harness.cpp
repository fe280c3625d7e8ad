// ctypes-callable harness around the UNMODIFIED reference (libhqpref.so).
//
// TEST INFRASTRUCTURE ONLY: used by tests/, by __graft_entry__.smoke() and by
// bench.py's cpu_baseline / --impl reference legs as the checker / the CPU
// baseline.  Never on the product path.
//
// Everything here drives the reference through its own public C++ API:
//   Hqp_Program            (hqp/Hqp_Program.h:33-65)
//   Hqp_IpMatrix plugins   (hqp/Hqp_IpMatrix.h:63-88) created by class id
//                          through If_ClassList (iftcl/If_Class.h:93-108)
//   Hqp_IpsMehrotra/Franke (hqp/Hqp_Solver.h:64-71)
//   the hqp_docp example   (hqp_docp/Docp_Main.C:21-110) with the outer loop of
//                          hqp/hqp_solve.tcl:83-265 restated in C++ because no
//                          Tcl interpreter exists in this image.

// std headers first: hqp/Meschach.h:42 defines min/max as macros
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <string>

#include <If.h>
#include <If_Class.h>
#include <If_Element.h>
#include <If_Procedure.h>

#include <Hqp.h>
#include <Hqp_HL_BFGS.h>
#include <Hqp_IpMatrix.h>
#include <Hqp_IpsFranke.h>
#include <Hqp_IpsMehrotra.h>
#include <Hqp_Program.h>
#include <Hqp_Solver.h>
#include <Prg_DID.h>

extern "C" {
#include <meschach/addon2_hqp.h>
#include <meschach/sparse.h>
}

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static bool g_inited = false;
static std::string g_last_error;

//--------------------------------------------------------------------------
// hqp_solve.tcl:83-265 restated: outer SQP loop on the same command names.
// Returns the Tcl result string ("optimal") or throws the error word.
//--------------------------------------------------------------------------
static int geti(const char *n) { int v = 0; If_GetInt(n, &v); return v; }
static double getr(const char *n) { double v = 0; If_GetReal(n, &v); return v; }
static std::string gets_(const char *n) {
  const char *s = NULL;
  If_GetString(n, &s);
  return s ? s : "";
}

static int g_qp_iters_total = 0;
static int g_verbose = 0;

static const char *hqp_solve_loop() {
  int qp_iters = 0, nullsteps = 0, hela_restart = 0;
  if (g_verbose)
    printf("%3s %12s %10s %10s [%3s %3.3s] %10s %10s %8s\n", "it", "obj",
           "||inf||", "||grdL||", "qp", "result", "||s||", "s'Qs", "stepsize");
  while (1) {
    if (qp_iters == 0) {
      If_Eval("sqp_qp_update");                       // hqp_solve.tcl:110-113
      if (g_verbose)
        printf("%3d %12.6g %10.4g %10.4g ", geti("sqp_iter"), getr("prg_f"),
               getr("sqp_norm_inf"), getr("sqp_norm_grd_L"));
      double chk = getr("prg_f") * getr("sqp_norm_inf");
      if (chk != chk) return "evaluation";            // :118-121
    } else {
      if (g_verbose)
        printf("%3d %12.6g %10.4g ", geti("sqp_iter"), getr("prg_f"),
               getr("sqp_norm_inf"));
      double chk = getr("prg_f") * getr("sqp_norm_inf");
      if (chk != chk) return "evaluation";            // :127-130
      // the extra break test :136-145 is hot-start only (hot = 0 here)
      If_Eval("sqp_qp_update");                       // :147
      if (g_verbose) printf("%10.4g ", getr("sqp_norm_grd_L"));
    }
    if (getr("sqp_xQx") < 0.0) {                      // :168-173
      If_Eval("sqp_hela_restart");
      hela_restart = 1;
    } else
      hela_restart = 0;
    if (geti("sqp_iter") > 0) {                       // :176-180
      if (getr("sqp_norm_inf") < getr("sqp_eps") &&
          getr("sqp_norm_grd_L") < getr("sqp_eps"))
        break;
    }
    If_Eval("sqp_qp_solve");                          // :182-183
    qp_iters += geti("qp_iter");
    std::string qres = gets_("qp_result");
    if (g_verbose) printf("[%3d %3.3s] ", geti("qp_iter"), qres.c_str());
    if (geti("qp_iter") == 0) {                       // :189-192
      g_last_error = qres;
      g_qp_iters_total = qp_iters;
      return g_last_error.c_str();
    }
    if (g_verbose) printf("%10.4g %10.4g ", getr("sqp_norm_s"), getr("sqp_sQs"));
    if (getr("sqp_sQs") < 0.0) If_Eval("sqp_hela_restart");   // :198-200
    if (geti("sqp_iter") > 0 && getr("sqp_sQs") >= 0.0 && !hela_restart) {  // :202-212
      double eps = getr("sqp_eps");
      if (getr("sqp_norm_inf") < eps && qres == "optimal") {
        if (getr("sqp_sQs") < eps * eps) break;
        if (geti("sqp_iter") > 2) {
          if (getr("sqp_norm_s") < eps * getr("sqp_norm_x") &&
              getr("sqp_norm_df") < eps * fabs(getr("prg_f")) &&
              getr("sqp_sQs") < eps)
            break;
        }
      }
    }
    If_Eval("sqp_step");                              // :214
    if (g_verbose) printf("%8.3g\n", getr("sqp_alpha"));
    g_qp_iters_total = qp_iters;
    if (geti("qp_iter") >= geti("qp_max_iters") && qres != "feasible")
      return "subiters";                              // :218-220
    if (geti("sqp_iter") >= geti("sqp_max_iters")) return "iters";      // :221-223
    if (geti("sqp_inf_iters") >= geti("sqp_max_inf_iters"))             // :224-230
      return qres == "suboptimal" ? "infeasible" : "degenerate";
    if (getr("sqp_alpha") < 1e-8 &&
        getr("sqp_norm_df") < getr("sqp_eps") * fabs(getr("prg_f")))    // :232-239
      nullsteps++;
    else
      nullsteps = 0;
    if (nullsteps > 5) return "stall";
  }
  if (g_verbose) printf("\n%43d qp-it\n", qp_iters);
  g_qp_iters_total = qp_iters;
  return "optimal";
}

static const char *g_solve_result = "";
static void hqp_solve_cmd() { g_solve_result = hqp_solve_loop(); }

//--------------------------------------------------------------------------
static VEC *vec_from(const double *p, int n) {
  VEC *v = v_get(n > 0 ? n : 1);
  v = v_resize(v, n);
  if (n > 0 && p) memcpy(v->ve, p, sizeof(double) * n);
  return v;
}
static void vec_to(const VEC *v, double *p) {
  if (p && v->dim > 0) memcpy(p, v->ve, sizeof(double) * v->dim);
}

struct RefMat {
  Hqp_IpMatrix *mat;
};

extern "C" {

int ref_init(void) {
  if (g_inited) return 0;
  static char arg0[] = "ref_harness";
  char *argv[] = {arg0, NULL};
  if (If_CreateInterp(1, argv) != IF_OK) return 1;
  Hqp_Init(If_Interp());
  static If_List procs;
  procs.append(new If_Procedure("hqp_solve", &hqp_solve_cmd));
  g_inited = true;
  return 0;
}

void ref_set_verbose(int v) { g_verbose = v; }

// dlopen a plugin shared object (e.g. the Hqp_IpCuda module) so that its
// IF_CLASS_DEFINE static constructor registers the class id with
// If_ClassList<Hqp_IpMatrix> (iftcl/If_Class.h:54-55).
int ref_load_plugin(const char *path) {
  ref_init();
  void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    g_last_error = dlerror();
    fprintf(stderr, "ref_load_plugin: %s\n", g_last_error.c_str());
    return 1;
  }
  return 0;
}

const char *ref_last_error(void) { return g_last_error.c_str(); }

//------------------------------------------------------------------ QP --
// CSR inputs; Q holds the UPPER triangle only (symsp convention,
// hqp/Hqp_Program.h:47).
void *ref_qp_create(int n, int me, int m, const int *Qp, const int *Qj,
                    const double *Qv, const double *c, const int *Ap,
                    const int *Aj, const double *Av, const double *b,
                    const int *Cp, const int *Cj, const double *Cv,
                    const double *d) {
  ref_init();
  Hqp_Program *qp = new Hqp_Program;
  int elq = 1, ela = 1, elc = 1;
  for (int i = 0; i < n; i++) if (Qp[i + 1] - Qp[i] > elq) elq = Qp[i + 1] - Qp[i];
  for (int i = 0; i < me; i++) if (Ap[i + 1] - Ap[i] > ela) ela = Ap[i + 1] - Ap[i];
  for (int i = 0; i < m; i++) if (Cp[i + 1] - Cp[i] > elc) elc = Cp[i + 1] - Cp[i];
  qp->resize(n, me, m, elq, ela, elc);
  for (int i = 0; i < n; i++)
    for (int e = Qp[i]; e < Qp[i + 1]; e++) sp_set_val(qp->Q, i, Qj[e], Qv[e]);
  for (int i = 0; i < me; i++)
    for (int e = Ap[i]; e < Ap[i + 1]; e++) sp_set_val(qp->A, i, Aj[e], Av[e]);
  for (int i = 0; i < m; i++)
    for (int e = Cp[i]; e < Cp[i + 1]; e++) sp_set_val(qp->C, i, Cj[e], Cv[e]);
  if (n) memcpy(qp->c->ve, c, sizeof(double) * n);
  if (me) memcpy(qp->b->ve, b, sizeof(double) * me);
  if (m) memcpy(qp->d->ve, d, sizeof(double) * m);
  v_zero(qp->x);
  return qp;
}

void ref_qp_free(void *qp) { delete (Hqp_Program *)qp; }

// overwrite the numerical values of Q/A/C in place (pattern unchanged):
// what Hqp_Docp::update does once per SQP iteration.
int ref_qp_set_values(void *qp_, const int *Qp, const int *Qj, const double *Qv,
                      const int *Ap, const int *Aj, const double *Av,
                      const int *Cp, const int *Cj, const double *Cv) {
  Hqp_Program *qp = (Hqp_Program *)qp_;
  for (int i = 0; i < qp->Q->m; i++)
    for (int e = Qp[i]; e < Qp[i + 1]; e++) sp_set_val(qp->Q, i, Qj[e], Qv[e]);
  for (int i = 0; i < qp->A->m; i++)
    for (int e = Ap[i]; e < Ap[i + 1]; e++) sp_set_val(qp->A, i, Aj[e], Av[e]);
  for (int i = 0; i < qp->C->m; i++)
    for (int e = Cp[i]; e < Cp[i + 1]; e++) sp_set_val(qp->C, i, Cj[e], Cv[e]);
  return 0;
}

//-------------------------------------------------------------- plugin --
void *ref_mat_create(const char *name) {
  ref_init();
  If_ClassList<Hqp_IpMatrix> *list = If_ClassList_Hqp_IpMatrix();
  if (!list) return NULL;
  Hqp_IpMatrix *m = list->createObject(name);
  if (!m) return NULL;
  RefMat *r = new RefMat;
  r->mat = m;
  return r;
}

void ref_mat_free(void *h) {
  RefMat *r = (RefMat *)h;
  delete r->mat;
  delete r;
}

int ref_set_real(const char *name, double v) { return If_SetReal(name, v); }
int ref_set_int(const char *name, int v) { return If_SetInt(name, v); }
int ref_set_string(const char *name, const char *v) { return If_SetString(name, v); }
double ref_get_real(const char *name) { return getr(name); }
int ref_get_int(const char *name) { return geti(name); }

// each returns 0, or the Meschach error number raised (E_SING = 4)
int ref_mat_init(void *h, void *qp) {
  int err = 0;
  m_catchall(((RefMat *)h)->mat->init((Hqp_Program *)qp), err = _err_num);
  return err;
}
int ref_mat_update(void *h, void *qp) {
  int err = 0;
  m_catchall(((RefMat *)h)->mat->update((Hqp_Program *)qp), err = _err_num);
  return err;
}
int ref_mat_factor(void *h, void *qp, int m, const double *z, const double *w) {
  VEC *zv = vec_from(z, m), *wv = vec_from(w, m);
  int err = 0;
  m_catchall(((RefMat *)h)->mat->factor((Hqp_Program *)qp, zv, wv),
             err = _err_num);
  v_free(zv);
  v_free(wv);
  return err;
}

// mode 0: step()   1: solve() (step + iterative refinement, Hqp_IpMatrix.C:65-128)
// mode 2: residuum() of the given dx..dw (inputs, not overwritten)
int ref_mat_apply(void *h, void *qp_, int mode, const double *z,
                  const double *w, const double *r1, const double *r2,
                  const double *r3, const double *r4, double *dx, double *dy,
                  double *dz, double *dw, double *res) {
  Hqp_Program *qp = (Hqp_Program *)qp_;
  Hqp_IpMatrix *mat = ((RefMat *)h)->mat;
  int n = qp->Q->n, me = qp->A->m, m = qp->C->m;
  VEC *zv = vec_from(z, m), *wv = vec_from(w, m);
  VEC *v1 = vec_from(r1, n), *v2 = vec_from(r2, me), *v3 = vec_from(r3, m),
      *v4 = vec_from(r4, m);
  VEC *x = vec_from(mode == 2 ? dx : NULL, n), *y = vec_from(mode == 2 ? dy : NULL, me),
      *zz = vec_from(mode == 2 ? dz : NULL, m), *ww = vec_from(mode == 2 ? dw : NULL, m);
  int err = 0;
  double r = 0.0;
  m_catchall(
      if (mode == 0) mat->step(qp, zv, wv, v1, v2, v3, v4, x, y, zz, ww);
      else if (mode == 1) r = mat->solve(qp, zv, wv, v1, v2, v3, v4, x, y, zz, ww);
      else r = mat->residuum(qp, zv, wv, v1, v2, v3, v4, x, y, zz, ww),
      err = _err_num);
  if (mode != 2) {
    vec_to(x, dx); vec_to(y, dy); vec_to(zz, dz); vec_to(ww, dw);
  }
  if (res) *res = r;
  v_free(zv); v_free(wv); v_free(v1); v_free(v2); v_free(v3); v_free(v4);
  v_free(x); v_free(y); v_free(zz); v_free(ww);
  return err;
}

// time `reps` x (1 factor + nstep step) on one thread; returns best-of times.
int ref_mat_time(void *h, void *qp_, const double *z, const double *w,
                 const double *r1, const double *r2, const double *r3,
                 const double *r4, int reps, int nstep, double *t_factor,
                 double *t_step) {
  Hqp_Program *qp = (Hqp_Program *)qp_;
  Hqp_IpMatrix *mat = ((RefMat *)h)->mat;
  int n = qp->Q->n, me = qp->A->m, m = qp->C->m;
  VEC *zv = vec_from(z, m), *wv = vec_from(w, m);
  VEC *v1 = vec_from(r1, n), *v2 = vec_from(r2, me), *v3 = vec_from(r3, m),
      *v4 = vec_from(r4, m);
  VEC *x = vec_from(NULL, n), *y = vec_from(NULL, me), *zz = vec_from(NULL, m),
      *ww = vec_from(NULL, m);
  double bf = 1e300, bs = 1e300;
  int err = 0;
  m_catchall(
      for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        mat->factor(qp, zv, wv);
        double t1 = now_s();
        for (int s = 0; s < nstep; s++)
          mat->step(qp, zv, wv, v1, v2, v3, v4, x, y, zz, ww);
        double t2 = now_s();
        if (t1 - t0 < bf) bf = t1 - t0;
        if (nstep > 0 && (t2 - t1) / nstep < bs) bs = (t2 - t1) / nstep;
      },
      err = _err_num);
  *t_factor = bf;
  *t_step = bs;
  v_free(zv); v_free(wv); v_free(v1); v_free(v2); v_free(v3); v_free(v4);
  v_free(x); v_free(y); v_free(zz); v_free(ww);
  return err;
}

//----------------------------------------------------------- IP solver --
// Stand-alone IP solve of one QP: solver = "Mehrotra" |
// "Franke", mat = "LQDOCP" | "RedSpBKP" | "SpBKP" | plugin id.
// result: 0 optimal (Hqp_Optimal) ... as hqp/Hqp.h Hqp_Result; returns
// Meschach error number if one escaped.
int ref_ips_solve(void *qp_, const char *solver, const char *mat, double eps,
                  int max_iters, double *x, double *y, double *z, int *iters,
                  int *result, double *seconds) {
  ref_init();
  Hqp_Program *qp = (Hqp_Program *)qp_;
  If_ClassList<Hqp_Solver> *list = If_ClassList_Hqp_Solver();
  Hqp_Solver *s = list ? list->createObject(solver) : NULL;
  if (!s) return -1;
  int err = 0;
  // (a solver module without an exchangeable matrix solver -- "CudaMehrotra",
  //  "CudaFranke" -- is selected with an empty mat)
  if (mat && *mat && If_SetString("qp_mat_solver", mat) != IF_OK) {
    delete s;
    return -2;
  }
  if (getenv("HQP_MAT_NGPU")) If_SetInt("mat_ngpu", atoi(getenv("HQP_MAT_NGPU")));
  s->qp(qp);
  s->eps(eps);
  if (max_iters > 0) s->max_iters(max_iters);
  double t0 = 0, t1 = 0;
  m_catchall(s->init(); s->update(); s->cold_start(); t0 = now_s(); s->solve();
             t1 = now_s(), err = _err_num);
  if (!err) {
    vec_to(qp->x, x);
    vec_to(s->y(), y);
    vec_to(s->z(), z);
    *iters = s->iter();
    *result = (int)s->result();
  }
  if (seconds) *seconds = t1 - t0;
  delete s;
  return err;
}

// A sequence of IP solves on one Hqp_Solver object: the first cold-started, the
// following hot-started (Hqp_SqpSolver.C:284-294) after the linear terms c, b, d
// of the program were replaced by row k of cs / bs / ds.  Outputs per solve.
int ref_ips_solve_seq(void *qp_, const char *solver, const char *mat, double eps,
                      int max_iters, int nsolve, const double *cs, const double *bs,
                      const double *ds, double *xs, double *ys, double *zs, int *iters,
                      int *results) {
  ref_init();
  Hqp_Program *qp = (Hqp_Program *)qp_;
  If_ClassList<Hqp_Solver> *list = If_ClassList_Hqp_Solver();
  Hqp_Solver *s = list ? list->createObject(solver) : NULL;
  if (!s) return -1;
  if (If_SetString("qp_mat_solver", mat) != IF_OK) {
    delete s;
    return -2;
  }
  int err = 0;
  s->qp(qp);
  s->eps(eps);
  if (max_iters > 0) s->max_iters(max_iters);
  const int n = qp->c->dim, me = qp->b->dim, m = qp->d->dim;
  for (int k = 0; k < nsolve && !err; k++) {
    for (int i = 0; i < n; i++) qp->c->ve[i] = cs[(size_t)k * n + i];
    for (int i = 0; i < me; i++) qp->b->ve[i] = bs[(size_t)k * me + i];
    for (int i = 0; i < m; i++) qp->d->ve[i] = ds[(size_t)k * m + i];
    if (k == 0) {
      m_catchall(s->init(); s->update(); s->cold_start(); s->solve(), err = _err_num);
    } else {
      m_catchall(s->update(); s->hot_start(); s->solve(), err = _err_num);
    }
    if (err) break;
    vec_to(qp->x, xs + (size_t)k * n);
    vec_to(s->y(), ys + (size_t)k * me);
    if (m) vec_to(s->z(), zs + (size_t)k * m);
    iters[k] = s->iter();
    results[k] = (int)s->result();
  }
  delete s;
  return err;
}

//------------------------------------------------------- docp example --
// hqp_docp/Docp_Main.C restated with selectable solvers.  One call per
// process is the supported use (global theSqpSolver state).
int ref_docp_did(int kmax, const char *qp_solver, const char *mat_solver,
                 double sqp_eps, int with_cns, double *objective, int *sqp_iters,
                 int *qp_iters, int *line_steps, char *result, int result_len) {
  ref_init();
  Prg_DID *prg = new Prg_DID();
  If_SetReal("sqp_eps", sqp_eps);
  if (qp_solver && *qp_solver)
    if (If_SetString("sqp_qp_solver", qp_solver) != IF_OK) return -1;
  // (a solver module without an exchangeable matrix solver -- "CudaMehrotra" --
  // is selected with an empty mat_solver)
  if (mat_solver && *mat_solver && If_SetString("qp_mat_solver", mat_solver) != IF_OK) {
    fprintf(stderr, "qp_mat_solver %s: %s\n", mat_solver, If_ResultString());
    return -2;
  }
  // qp_eps of a freshly selected solver module is Hqp_Solver's 1e-10; the as-shipped
  // Franke instance runs with the 1e-9 of Hqp_SqpSolver's constructor -- tests that
  // compare a new module with the as-shipped run set it explicitly
  if (getenv("HQP_QP_EPS")) If_SetReal("qp_eps", atof(getenv("HQP_QP_EPS")));
  // option of a plugged-in matrix module (Hqp_IpCuda: horizon split over GPUs)
  if (getenv("HQP_MAT_NGPU")) If_SetInt("mat_ngpu", atoi(getenv("HQP_MAT_NGPU")));
  if (kmax > 0) If_SetInt("prg_kmax", kmax);
  If_SetInt("prg_with_cns", with_cns);
  int rc = 0;
  if (If_Eval("prg_setup") != IF_OK) rc = -3;
  if (!rc && If_Eval("prg_simulate") != IF_OK) rc = -4;
  if (!rc && If_Eval("sqp_init") != IF_OK) rc = -5;
  if (rc) {
    fprintf(stderr, "ref_docp_did setup: %s\n", If_ResultString());
    return rc;
  }
  g_solve_result = "";
  if (If_Eval("hqp_solve") != IF_OK) {
    snprintf(result, result_len, "error: %s", If_ResultString());
  } else
    snprintf(result, result_len, "%s", g_solve_result);
  *objective = getr("prg_f");
  *sqp_iters = geti("sqp_iter");
  *line_steps = geti("prg_fbd_evals");
  *qp_iters = g_qp_iters_total;
  delete prg;
  return 0;
}


// ---- row f4: hqp_solve for the program created last (ref_docp_create: the constructor of an
// Hqp_SqpProgram hands itself to theSqpSolver, hqp/Hqp_SqpProgram.C:63-64); the calls of
// hqp_docp/Docp_Main.C:68-76 after prg_setup
int ref_prg_solve(double sqp_eps, int simulate, const char *qp_solver, const char *mat_solver,
                  double *objective, int *sqp_iters, int *qp_iters, char *result, int result_len) {
  If_SetReal("sqp_eps", sqp_eps);
  if (qp_solver && *qp_solver && If_SetString("sqp_qp_solver", qp_solver) != IF_OK) return -1;
  if (mat_solver && *mat_solver && If_SetString("qp_mat_solver", mat_solver) != IF_OK) return -2;
  if (simulate && If_Eval("prg_simulate") != IF_OK) return -4;
  if (If_Eval("sqp_init") != IF_OK) {
    fprintf(stderr, "ref_prg_solve sqp_init: %s\n", If_ResultString());
    return -5;
  }
  g_solve_result = "";
  g_qp_iters_total = 0;
  if (If_Eval("hqp_solve") != IF_OK)
    snprintf(result, result_len, "error: %s", If_ResultString());
  else
    snprintf(result, result_len, "%s", g_solve_result);
  *objective = getr("prg_f");
  *sqp_iters = geti("sqp_iter");
  *qp_iters = g_qp_iters_total;
  return 0;
}

// ---- row f2: the reference's own block update Hqp_HL_BFGS::update_b_Q
// (hqp/Hqp_HL_BFGS.C:149-213; protected, reached through a subclass), on ONE dense
// block Q (n x n, row-major, both triangles), in place
class RefBfgsBlock : public Hqp_HL_BFGS {
 public:
  void block(const VEC *s, const VEC *u, Real alpha, MAT *Q, Real gamma, bool ec, Real eps) {
    _gamma = gamma;
    _eigen_control = ec;
    _eps = eps;
    _logging = false;
    update_b_Q(s, u, alpha, Q);
  }
};

int ref_hl_bfgs_block(int n, double *Q, const double *s, const double *u, double alpha,
                      double gamma, double eps, int eigen_control) {
  if (ref_init()) return -1;
  static RefBfgsBlock *hl = NULL;
  if (!hl) hl = new RefBfgsBlock();
  MAT *M = m_get(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) M->me[i][j] = Q[(size_t)i * n + j];
  VEC *vs = vec_from(s, n), *vu = vec_from(u, n);
  hl->block(vs, vu, alpha, M, gamma, eigen_control != 0, eps);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Q[(size_t)i * n + j] = M->me[i][j];
  m_free(M);
  v_free(vs);
  v_free(vu);
  return 0;
}

// ---- row f2 through the module interface: an Hqp_HL module by name ("BFGS" = the
// reference, "CudaBFGS" = hqp_b200/host/Hqp_HL_CudaBFGS.C from the plugin library) is set
// up on the QP's Hessian and asked for one update (Hqp_HL::setup / ::update)
#include <Hqp_SqpProgram.h>
class HarnessSqp : public Hqp_SqpProgram {
 public:
  explicit HarnessSqp(Hqp_Program *q) { _qp = q; }
  ~HarnessSqp() { _qp = NULL; }  // (the QP belongs to the caller)
  void setup() {}
  void init_x() {}
  void update_fbd() {}
  void update(const VECP, const VECP) {}
  const char *name() { return "harness"; }
};

int ref_hl_update(const char *hela, void *qp_, const double *s, const double *u, double alpha,
                  double gamma, double eps, int eigen_control) {
  if (ref_init()) return -1;
  Hqp_Program *qp = (Hqp_Program *)qp_;
  If_ClassList<Hqp_HL> *list = If_ClassList_Hqp_HL();
  Hqp_HL *hl = list ? list->createObject(hela) : NULL;
  if (!hl) return -2;
  If_SetReal("sqp_hela_gamma", gamma);
  If_SetReal("sqp_hela_eps", eps);
  If_SetInt("sqp_hela_eigen_control", eigen_control != 0);  // (If_Bool accepts 0 / 1)
  int rc = 0;
  {
    HarnessSqp prg(qp);
    VEC *vs = vec_from(s, qp->Q->n), *vu = vec_from(u, qp->Q->n);
    int code = 0;
    m_catchall(hl->setup(&prg); hl->update(vs, vu, alpha, &prg), code = 1);
    rc = code ? -3 : 0;
    v_free(vs);
    v_free(vu);
  }
  delete hl;
  return rc;
}

// dense copy (both triangles) of the diagonal block offs .. offs+size-1 of the QP's Hessian
int ref_qp_get_Q_block(void *qp_, int offs, int size, double *out) {
  Hqp_Program *qp = (Hqp_Program *)qp_;
  MAT *M = m_get(size, size);
  symsp_extract_mat(qp->Q, offs, M);
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++) out[(size_t)i * size + j] = M->me[i][j];
  m_free(M);
  return 0;
}

// ---- row f3: the reference's own SQP-level vector operations -- Hqp_SqpSolver::grd_L
// (hqp/Hqp_SqpSolver.C:430-445), ::norm_inf (:155-174), x'Qx / s'Qs (:225-226, 299-301) and
// Hqp_SqpPowell::phi / ::phi1 (hqp/Hqp_SqpPowell.C:189-244); protected members, reached
// through a subclass working on the caller's Hqp_Program
#include <Hqp_SqpPowell.h>
class HarnessPowell : public Hqp_SqpPowell {
 public:
  // out8 = { phi, phi1, s'Qs, c's, ||b||inf via norm_inf part 1, max(0,-min d), 0, 0 }
  void eval(HarnessSqp *prg, double f, const VEC *y, const VEC *z, const VEC *re, const VEC *r,
            double *grd, double *out8) {
    _prg = prg;
    prg->set_f(f);
    _y = v_copy(y, _y);
    _z = v_copy(z, _z);
    _re = v_copy(re, _re);
    _r = v_copy(r, _r);
    Hqp_Program *qp = prg->qp();
    VEC *g = grd_L(qp, VNULL);
    if (grd) memcpy(grd, g->ve, sizeof(double) * g->dim);
    v_free(g);
    out8[0] = phi();
    out8[1] = phi1();
    VEC *t = sp_mv_symmlt(qp->Q, qp->x, VNULL);
    out8[2] = in_prod(t, qp->x);
    v_free(t);
    out8[3] = in_prod(qp->c, qp->x);
    out8[4] = norm_inf(qp);
    out8[5] = out8[6] = out8[7] = 0.0;
    _prg = NULL;
  }
};

// qp->x must hold the step s (ref_qp_set_x); y, z multipliers; re, r penalty weights
int ref_sqp_eval(void *qp_, double f, const double *s, const double *y, const double *z,
                 const double *re, const double *r, double *grd_L, double *out8) {
  if (ref_init()) return -1;
  Hqp_Program *qp = (Hqp_Program *)qp_;
  static HarnessPowell *sp = NULL;
  if (!sp) sp = new HarnessPowell();
  const int n = qp->Q->m, me = qp->A->m, m = qp->C->m;
  memcpy(qp->x->ve, s, sizeof(double) * n);
  VEC *vy = vec_from(y, me), *vz = vec_from(z, m), *vre = vec_from(re, me), *vr = vec_from(r, m);
  {
    HarnessSqp prg(qp);
    sp->eval(&prg, f, vy, vz, vre, vr, grd_L, out8);
  }
  v_free(vy); v_free(vz); v_free(vre); v_free(vr);
  return 0;
}
}  // extern "C"
