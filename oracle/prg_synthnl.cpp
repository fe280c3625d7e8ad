// prg_synthnl.cpp -- TEST INFRASTRUCTURE ONLY (oracle/README.md).  Row f4's pin against the
// unmodified reference: the synthetic stage model of hqp_b200/csrc/docp_models.cuh
// (ModelSynthNL) restated as an Hqp_Docp subclass, so that the reference's OWN
// Hqp_Docp::setup / ::update / ::update_fbd / ::update_grds / ::update_bounds
// (hqp/Hqp_Docp.C:397-1180) produce qp->A, C, b, c, d and f for it; and the reference's own
// example Prg_DID (hqp_docp/Prg_DID.C) driven through the same calls.
//
// Only update_vals() and setup_vars() below are model code; they evaluate the same expressions
// in the same order as ModelSynthNL::vals (operation order is part of the model's contract).
#include <cstring>
#include <vector>

#include <Hqp_Docp.h>
#include <Hqp_Program.h>
#include <If_Class.h>
#include <If_Element.h>
#include <Prg_DID.h>

#include "../hqp_b200/host/Hqp_DocpCuda.h"  // (the product's host mix-in under test)

extern "C" int ref_init(void);

namespace {

class Prg_SynthNL : public Hqp_Docp {
 public:
  int K, nx, nu, nc, ncK;
  std::vector<double> par, spar, xinit;

  const char *name() { return "SynthNL"; }

  void setup_horizon(int &k0, int &kf) {
    k0 = 0;
    kf = K;
  }

  // bounds chosen to populate all six association tables (parse_constr, hqp/Hqp_Docp.C:370-397)
  void setup_vars(int k, VECP x, VECP x_min, VECP x_max, IVECP x_int, VECP u, VECP u_min, VECP u_max,
                  IVECP u_int, VECP c, VECP c_min, VECP c_max) {
    const int nd = nx + nu;
    alloc_vars(x, x_min, x_max, x_int, nx);
    for (int i = 0; i < nx; i++) x[i] = xinit[(size_t)k * nd + i];
    if (k == 0)
      for (int i = 0; i < nx; i++) x_min[i] = x_max[i] = x[i];  // fixed initial state
    else if (k < K) {
      x_max[nx > 1 ? 1 : 0] = 2.0;   // a state bound from above
      x_min[0] = -3.0;               // and one from below
    }
    if (k < K) {
      alloc_vars(u, u_min, u_max, u_int, nu);
      for (int j = 0; j < nu; j++) {
        u[j] = xinit[(size_t)k * nd + nx + j];
        u_min[j] = -1.0;
        u_max[j] = 1.0;
      }
      if (nc > 0) {
        alloc_vars(c, c_min, c_max, IVNULL, nc);
        for (int i = 0; i < nc; i++) {
          if (i % 3 == 0) c_max[i] = 2.0;
          else if (i % 3 == 1) { c_min[i] = -1.0; c_max[i] = 1.0; }
          else c_min[i] = c_max[i] = 0.1;
        }
      }
    } else if (ncK > 0) {
      alloc_vars(c, c_min, c_max, IVNULL, ncK);
      for (int i = 0; i < ncK; i++) c_min[i] = -1.0;
    }
  }

  void update_vals(int k, const VECP x, const VECP u, VECP f, Real &f0, VECP c) {
    const double eps = par[0];
    const double *A = &par[1], *B = A + nx * nx, *qw = B + nx * nu, *rw = qw + nx;
    const double *r = &spar[(size_t)k * nx];
    double s = 0.0;
    for (int i = 0; i < nx; i++) {
      double e = x[i] - r[i];
      s = s + qw[i] * e * e;
    }
    if (k < K) {
      for (int i = 0; i < nx; i++) {
        double acc = 0.0;
        for (int j = 0; j < nx; j++) acc = acc + A[i * nx + j] * x[j];
        for (int j = 0; j < nu; j++) acc = acc + B[i * nu + j] * u[j];
        f[i] = acc + eps * (x[i] / (1.0 + x[i] * x[i]));
      }
      for (int j = 0; j < nu; j++) s = s + rw[j] * u[j] * u[j];
      s = 0.5 * s;
      const int nm = nx < nu ? nx : nu;
      for (int j = 0; j < nm; j++) s = s + eps * x[j] * u[j];
      f0 = s;
      if (nc > 0) {
        double q = 0.0;
        for (int i = 0; i < nx; i++) q = q + x[i] * x[i];
        c[0] = q / (double)nx + eps * u[0] * x[0];
        for (int i = 1; i < nc; i++) c[i] = x[i] * u[i % nu];
      }
    } else {
      f0 = 0.5 * s;
      for (int i = 0; i < ncK; i++) c[i] = x[i] * x[i];
    }
  }
};

// the same program with its stage loop on the GPU (the mix-in under test)
class Prg_SynthNLCuda : public Hqp_DocpCuda<Prg_SynthNL> {
 public:
  const char *name() { return "SynthNLCuda"; }
  int cuda_model() { return HQPDOCP_MODEL_SYNTHNL; }
  void cuda_params(int, std::vector<double> &p, int &nspar, std::vector<double> &sp) {
    p = par;
    nspar = nx;
    sp = spar;
  }
};

struct DocpRef {
  Hqp_SqpProgram *prg;
};

}  // namespace

extern "C" {

// model 0: the reference's Prg_DID (par[0] unused: dt = 1/K inside Prg_DID; nc = prg_with_cns);
// model 1: Prg_SynthNL; models 2 / 3: the same two programs with the stage loop on the GPU
// ("DIDCuda" from the plugin library, ref_load_plugin first; Prg_SynthNLCuda above).  xinit [N] = the iterate the program is set up at (model 1 only; Prg_DID
// has its own initial values, hqp_docp/Prg_DID.C:43-49).
void *ref_docp_create(int model, int K, int nx, int nu, int nc, int ncK, const double *par, int npar,
                      const double *spar, int nspar, const double *xinit) {
  if (ref_init()) return NULL;
  Hqp_SqpProgram *prg = NULL;
  if (model == 0 || model == 2) {
    if (model == 0)
      prg = new Prg_DID();
    else {
      If_ClassList<Hqp_SqpProgram> *list = If_ClassList_Hqp_SqpProgram();
      prg = list ? list->createObject("DIDCuda") : NULL;
      if (!prg) return NULL;
    }
    If_SetInt("prg_kmax", K);
    If_SetInt("prg_with_cns", nc != 0);
  } else if (model == 1 || model == 3) {
    Prg_SynthNL *p = model == 1 ? new Prg_SynthNL() : new Prg_SynthNLCuda();
    p->K = K; p->nx = nx; p->nu = nu; p->nc = nc; p->ncK = ncK;
    p->par.assign(par, par + npar);
    p->spar.assign(spar, spar + (size_t)(K + 1) * nspar);
    p->xinit.assign(xinit, xinit + (size_t)K * (nx + nu) + nx);
    prg = p;
  } else
    return NULL;
  int code = 0;
  m_catchall(prg->setup(), code = 1);
  if (code) {
    delete prg;
    return NULL;
  }
  DocpRef *h = new DocpRef();
  h->prg = prg;
  return h;
}

void ref_docp_free(void *h_) {
  DocpRef *h = (DocpRef *)h_;
  if (!h) return;
  delete h->prg;
  delete h;
}

int ref_docp_sizes(void *h_, int *N, int *me, int *m) {
  Hqp_Program *qp = ((DocpRef *)h_)->prg->qp();
  *N = qp->c->dim;
  *me = qp->b->dim;
  *m = qp->d->dim;
  return 0;
}

int ref_docp_get_x(void *h_, double *x) {
  Hqp_SqpProgram *prg = ((DocpRef *)h_)->prg;
  memcpy(x, prg->x()->ve, sizeof(double) * prg->x()->dim);
  return 0;
}

// x [N] (NULL: keep the current iterate); fbd_only: Hqp_Docp::update_fbd, else ::update(y=0, z=0).
// Out: f, b [me], d [m], c [N] (qp->c), A [me*N], C [m*N] dense row-major (may be NULL).
int ref_docp_update(void *h_, const double *x, int fbd_only, double *f, double *b, double *d, double *c,
                    double *A, double *C) {
  Hqp_SqpProgram *prg = ((DocpRef *)h_)->prg;
  Hqp_Program *qp = prg->qp();
  const int N = qp->c->dim, me = qp->b->dim, m = qp->d->dim;
  if (x) {
    VEC *vx = v_get(N);
    memcpy(vx->ve, x, sizeof(double) * N);
    prg->set_x(vx);
    v_free(vx);
  }
  VEC *y = v_get(me), *z = v_get(m);
  v_zero(y);
  v_zero(z);
  int code = 0;
  if (fbd_only) {
    m_catchall(prg->update_fbd(), code = 1);
  } else {
    m_catchall(prg->update(y, z), code = 1);
  }
  v_free(y);
  v_free(z);
  if (code) return -1;
  *f = prg->f();
  memcpy(b, qp->b->ve, sizeof(double) * me);
  if (m) memcpy(d, qp->d->ve, sizeof(double) * m);
  if (c) memcpy(c, qp->c->ve, sizeof(double) * N);
  if (A)
    for (int i = 0; i < me; i++)
      for (int j = 0; j < N; j++) A[(size_t)i * N + j] = sp_get_val(qp->A, i, j);
  if (C)
    for (int i = 0; i < m; i++)
      for (int j = 0; j < N; j++) C[(size_t)i * N + j] = sp_get_val(qp->C, i, j);
  return 0;
}

}  // extern "C"
