/*
 * Hqp_IpsCuda.C -- see Hqp_IpsCuda.h
 */
// (standard headers first: Meschach.h defines min / max macros)
#include <assert.h>
#include <stdio.h>
#include <vector>

#include <If_Int.h>
#include <If_Real.h>
#include <If_Method.h>

#include "Hqp_Program.h"
#include "Hqp_IpsCuda.h"

#include "hqp_ipcuda.h"

typedef If_Method<Hqp_IpsCuda> If_Cmd;

IF_CLASS_DEFINE("CudaMehrotra", Hqp_IpsCuda, Hqp_Solver);
IF_CLASS_DEFINE("CudaFranke", Hqp_IpsCudaFranke, Hqp_Solver);

//--------------------------------------------------------------------------
Hqp_IpsCuda::Hqp_IpsCuda(int franke)
{
  _n = _me = _m = 0;
  _w = VNULL;
  _hot = 0;
  _franke = franke;
  _max_warm_iters = franke ? 15 : 25;  // hqp/Hqp_IpsFranke.C:82, hqp/Hqp_IpsMehrotra.C:111
  _beta = 0.995;                       // hqp/Hqp_IpsFranke.C:79
  _mu0 = 0.0;                          // :78
  _logging = 0;
  _gap = 0.0;
  if (franke) {
    _ifList.append(new If_Real("qp_beta", &_beta));
    _ifList.append(new If_Real("qp_mu0", &_mu0));
  }

  // the option names of Hqp_IpsMehrotra (:117-130) that apply here
  _ifList.append(new If_Int("qp_iter", &_iter));
  _ifList.append(new If_Int("qp_max_iters", &_max_iters));
  _ifList.append(new If_Real("qp_eps", &_eps));
  _ifList.append(new If_Real("qp_gap", &_gap));
  _ifList.append(new If_Int("qp_max_warm_iters", &_max_warm_iters));
  _ifList.append(new If_Int("qp_logging", &_logging));
  _ifList.append(new If_Cmd("qp_init", &Hqp_IpsCuda::init, this));
  _ifList.append(new If_Cmd("qp_update", &Hqp_IpsCuda::update, this));
  _ifList.append(new If_Cmd("qp_cold_start", &Hqp_IpsCuda::cold_start, this));
  _ifList.append(new If_Cmd("qp_hot_start", &Hqp_IpsCuda::hot_start, this));
  _ifList.append(new If_Cmd("qp_solve", &Hqp_IpsCuda::solve, this));
}

//--------------------------------------------------------------------------
Hqp_IpsCuda::~Hqp_IpsCuda()
{
  v_free(_w);
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::init()
{
  assert(_qp != NULL);
  _n = _qp->Q->n;
  _me = _qp->A->m;
  _m = _qp->C->m;
  _y = v_resize(_y, _me);
  _z = v_resize(_z, _m);
  _w = v_resize(_w, _m);
  _mat.init(_qp);  // structure detection + device engine (Hqp_IpCuda::init)
  _bp.assign(_mat.me_abi() > 0 ? _mat.me_abi() : 1, 0.0);
  _yp.assign(_mat.me_abi() > 0 ? _mat.me_abi() : 1, 0.0);
  _cp.assign(_mat.n_abi() > 0 ? _mat.n_abi() : 1, 0.0);
  _xp.assign(_mat.n_abi() > 0 ? _mat.n_abi() : 1, 0.0);
  _hot = 0;
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::update()
{
  _mat.update(_qp);
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::cold_start()
{
  _iter = 0;
  _hot = 0;
  _result = Hqp_Infeasible;
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::hot_start()
{
  _iter = 0;
  _hot = 1;
  _result = Hqp_Infeasible;
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::step()
{
  // single IP iterations are not exposed by the device-resident loop
  m_error(E_INTERN, "Hqp_IpsCuda::step: use solve()");
}

//--------------------------------------------------------------------------
void Hqp_IpsCuda::solve()
{
  static double none[1];
  const std::vector<int> &rowmap = _mat.rowmap();
  const bool ident = _mat.identity_rows();
  int i, iters = 0, res = (int)Hqp_Infeasible;
  double gap = 0.0;

  // equality rows in the engine's order (dynamics, x0, general rows); variables in
  // the engine's (possibly padded) stage layout
  const double *b = _qp->b->ve, *c = _qp->c->ve;
  double *y = _y->ve, *x = _qp->x->ve;
  if (!ident) {
    _mat.to_abi_y(_qp->b->ve, &_bp[0]);
    _mat.to_abi_y(_y->ve, &_yp[0]);
    b = &_bp[0];
    y = &_yp[0];
  }
  if (_mat.padded()) {
    _mat.to_abi_x(_qp->c->ve, &_cp[0]);
    _mat.to_abi_x(_qp->x->ve, &_xp[0]);
    c = &_cp[0];
    x = &_xp[0];
  }
  (void)rowmap;
  int rc;
  if (_franke)
    rc = hqpcu_franke_solve(_mat.handle(), c, b, _m ? _qp->d->ve : none, _eps, _max_iters,
                            _hot, _max_warm_iters, _beta, _mu0, x, y, _m ? _z->ve : none,
                            _m ? _w->ve : none, &iters, &res, &gap);
  else if (_hot)
    rc = hqpcu_mehrotra_hot_solve(_mat.handle(), c, b, _m ? _qp->d->ve : none, _eps,
                                  _max_iters, _max_warm_iters, x, y,
                                  _m ? _z->ve : none, _m ? _w->ve : none, &iters, &res, &gap);
  else
    rc = hqpcu_mehrotra_solve(_mat.handle(), c, b, _m ? _qp->d->ve : none, _eps,
                              _max_iters, x, y, _m ? _z->ve : none,
                              _m ? _w->ve : none, &iters, &res, &gap);
  if (rc != HQPCU_OK) {
    fprintf(stderr, "Hqp_IpsCuda::solve: %s\n", hqpcu_last_error());
    m_error(rc == HQPCU_E_SING ? E_SING : E_INTERN, "Hqp_IpsCuda::solve");
  }
  if (!ident) _mat.from_abi_y(&_yp[0], _y->ve);
  if (_mat.padded()) _mat.from_abi_x(&_xp[0], _qp->x->ve);
  _iter = iters;
  _gap = gap;
  _result = (Hqp_Result)res;
  if (_logging)
    printf("Hqp_IpsCuda::solve: %s start, %d iterations, result %d, gap %g\n",
           _hot ? "hot" : "cold", iters, res, gap);
}
