/*
 * Hqp_DocpCuda.h -- host side of SURVEY.md section 8, row f4: an Hqp_Docp program whose
 * stage loop (Hqp_Docp::update / ::update_fbd, hqp/Hqp_Docp.C:831-891, 944-1075) runs on
 * the GPU through the C ABI include/hqp_docpcuda.h.
 *
 * Hqp_DocpCuda<Base> is a mix-in over any Hqp_Docp subclass Base (a program written for the
 * reference: setup_horizon / setup_vars / setup_struct / update_vals stay as they are and
 * still serve setup() and prg_simulate).  The subclass names the device model that evaluates
 * the same stage functions (hqp_b200/csrc/docp_models.cuh) and its parameters:
 *
 *     class Prg_DIDCuda : public Hqp_DocpCuda<Prg_DID> {
 *       int  cuda_model() { return HQPDOCP_MODEL_DID; }
 *       void cuda_params(int K, std::vector<double> &par, int &nspar, std::vector<double> &spar)
 *         { par.assign(1, 1.0 / K); nspar = 0; }
 *     };
 *
 * What is overridden, and what each override replaces:
 *   setup()       Base::setup(), then the bounds of every stage are asked for once more
 *                 (setup_vars) and parsed into the six association tables the way
 *                 Hqp_Docp::setup_x / parse_constr do (hqp/Hqp_Docp.C:370-397, 465-541 -- the
 *                 reference keeps its own tables private), and the device handle is created
 *   update_fbd()  hqpdocp_update_fbd: f, qp->b, qp->d for the current x
 *   update(y, z)  hqpdocp_update: additionally qp->c and the stage Jacobians, which are then
 *                 written into the existing entries of qp->A / qp->C (what sp_update_mrow and
 *                 Hqp_DocpAssoc::sp_update_mat do, hqp/Hqp_Docp.C:1045-1064).  The Hessian is
 *                 left to the SQP solver's approximation, as with the default update_hela.
 * Interface element: prg_cuda_grad (0: forward differences as Hqp_Docp::update_grds, 1: dual
 * numbers; default 1), prg_cuda_device.
 * Restrictions (m_error E_FORMAT from setup): uniform stage dimensions, no Periodical states.
 * With a device-resident QP solver (Hqp_IpsCuda) the Jacobians need not come back at all:
 * hqpdocp_update_dev writes fx / fu where hqpcu_update_dev reads them.
 */
#ifndef Hqp_DocpCuda_H
#define Hqp_DocpCuda_H

#include <cmath>
#include <vector>

#include <Hqp_Docp.h>
#include <Hqp_Program.h>
#include <If_Int.h>

#include "hqp_docpcuda.h"

template <class Base>
class Hqp_DocpCuda : public Base {
 public:
  Hqp_DocpCuda() : _h(NULL), _grad(HQPDOCP_GRAD_AD), _device(0), _K(0), _nx(0), _nu(0), _nc(0), _ncK(0) {
    this->_ifList.append(new If_Int("prg_cuda_grad", &_grad));
    this->_ifList.append(new If_Int("prg_cuda_device", &_device));
  }
  ~Hqp_DocpCuda() { hqpdocp_destroy(_h); }

  virtual int cuda_model() = 0;
  virtual void cuda_params(int K, std::vector<double> &par, int &nspar, std::vector<double> &spar) = 0;

  void setup() {
    Base::setup();
    hqpdocp_destroy(_h);
    _h = NULL;
    const IVECP nxs = this->nxs(), nus = this->nus();
    _K = (int)nxs->dim - 1;
    _nx = nxs[0];
    _nu = nus[0];
    for (int k = 0; k <= _K; k++)
      if (nxs[k] != _nx || nus[k] != (k < _K ? _nu : 0))
        m_error(E_FORMAT, "Hqp_DocpCuda::setup: stage dimensions must be uniform");
    // the bounds of every stage once more, parsed like Hqp_Docp::setup_x does
    struct Tab { std::vector<int> idx; std::vector<double> val; } t[6];  // xu_eq lb ub, cns_eq lb ub
    VECP x = v_get(1), x_min = v_get(1), x_max = v_get(1), u = v_get(1), u_min = v_get(1), u_max = v_get(1);
    VECP c = v_get(1), c_min = v_get(1), c_max = v_get(1);
    IVECP x_int = iv_get(1), u_int = iv_get(1);
    int k0 = 0, kf = 0;
    this->setup_horizon(k0, kf);
    int ncns = 0;
    for (int k = 0; k <= _K; k++) {
      v_resize(x, 0); v_resize(x_min, 0); v_resize(x_max, 0); iv_resize(x_int, 0);
      v_resize(u, 0); v_resize(u_min, 0); v_resize(u_max, 0); iv_resize(u_int, 0);
      v_resize(c, 0); v_resize(c_min, 0); v_resize(c_max, 0);
      this->setup_vars(k0 + k, x, x_min, x_max, x_int, u, u_min, u_max, u_int, c, c_min, c_max);
      const int nck = (int)c_min->dim;
      if (k == 0) _nc = nck;
      if (k < _K ? nck != _nc : false) m_error(E_FORMAT, "Hqp_DocpCuda::setup: constraints per stage must be uniform");
      if (k == _K) _ncK = nck;
      const int kxu = k * (_nx + _nu);
      parse(x_min, x_max, kxu, t[0], t[1], t[2]);
      parse(u_min, u_max, kxu + _nx, t[0], t[1], t[2]);
      parse(c_min, c_max, ncns, t[3], t[4], t[5]);
      ncns += nck;
    }
    v_free(x); v_free(x_min); v_free(x_max); v_free(u); v_free(u_min); v_free(u_max);
    v_free(c); v_free(c_min); v_free(c_max); iv_free(x_int); iv_free(u_int);
    for (int i = 0; i < 6; i++) _n[i] = (int)t[i].idx.size();
    _cns_idx[0] = t[3].idx; _cns_idx[1] = t[4].idx; _cns_idx[2] = t[5].idx;
    Hqp_Program *qp = this->_qp;
    if ((int)qp->b->dim != _K * _nx + _n[0] + _n[3] || (int)qp->d->dim != _n[1] + _n[2] + _n[4] + _n[5])
      m_error(E_FORMAT, "Hqp_DocpCuda::setup: constraint layout differs from Hqp_Docp's (Periodical states?)");
    std::vector<double> par, spar;
    int nspar = 0;
    cuda_params(_K, par, nspar, spar);
    hqpdocp_dims D = hqpdocp_dims();
    D.K = _K; D.nx = _nx; D.nu = _nu; D.nc = _nc; D.ncK = _ncK;
    D.model = cuda_model();
    D.npar = (int)par.size(); D.par = par.empty() ? NULL : &par[0];
    D.nspar = nspar; D.spar = spar.empty() ? NULL : &spar[0];
    hqpdocp_assoc *a[6] = {&D.xu_eq, &D.xu_lb, &D.xu_ub, &D.cns_eq, &D.cns_lb, &D.cns_ub};
    for (int i = 0; i < 6; i++) {
      a[i]->dim = _n[i];
      a[i]->idxs = _n[i] ? &t[i].idx[0] : NULL;
      a[i]->vals = _n[i] ? &t[i].val[0] : NULL;
    }
    D.device = _device;
    if (hqpdocp_create(&D, &_h) != HQPDOCP_OK) {
      _h = NULL;
      m_error(E_FORMAT, hqpdocp_last_error());
    }
    _fx.resize((size_t)_K * _nx * _nx);
    _fu.resize((size_t)_K * _nx * _nu + 1);
    _cx.resize((size_t)ncns * _nx + 1);
    _cu.resize((size_t)_K * _nc * _nu + 1);
  }

  void update_fbd() {
    need_handle();
    Hqp_Program *qp = this->_qp;
    double f = 0.0;
    if (hqpdocp_update_fbd(_h, this->_x->ve, &f, qp->b->ve, qp->d->ve) != HQPDOCP_OK)
      m_error(E_INTERN, hqpdocp_last_error());
    this->_f = f;
  }

  void update(const VECP y, const VECP z) {
    need_handle();
    Hqp_Program *qp = this->_qp;
    if ((const VEC *)y == NULL || (const VEC *)z == NULL) m_error(E_NULL, "Hqp_DocpCuda::update");
    if (y->dim != qp->b->dim || z->dim != qp->d->dim) m_error(E_SIZES, "Hqp_DocpCuda::update");
    double f = 0.0;
    if (hqpdocp_update(_h, _grad ? HQPDOCP_GRAD_AD : HQPDOCP_GRAD_FD, this->_x->ve, &f, qp->b->ve, qp->d->ve,
                       qp->c->ve, &_fx[0], &_fu[0], &_cx[0], &_cu[0]) != HQPDOCP_OK)
      m_error(E_INTERN, hqpdocp_last_error());
    this->_f = f;
    // the Jacobians into the existing entries of A and C (hqp/Hqp_Docp.C:1045-1064)
    const int nd = _nx + _nu;
    for (int k = 0; k < _K; k++)
      for (int i = 0; i < _nx; i++) {
        SPROW *row = qp->A->row + k * _nx + i;
        put(row, k * nd, _nx, &_fx[((size_t)k * _nx + i) * _nx], 1.0);
        put(row, k * nd + _nx, _nu, &_fu[((size_t)k * _nx + i) * _nu], 1.0);
      }
    cns_rows(qp->A, _K * _nx + _n[0], _cns_idx[0], 1.0);
    cns_rows(qp->C, _n[1] + _n[2], _cns_idx[1], 1.0);
    cns_rows(qp->C, _n[1] + _n[2] + _n[4], _cns_idx[2], -1.0);
  }

 private:
  hqpdocp_handle *_h;
  int _grad, _device;
  int _K, _nx, _nu, _nc, _ncK;
  int _n[6];
  std::vector<int> _cns_idx[3];
  std::vector<double> _fx, _fu, _cx, _cu;

  void need_handle() {
    if (!_h) m_error(E_NULL, "Hqp_DocpCuda: setup() has not run");
  }

  template <class Tab>
  static void parse(const VECP cmin, const VECP cmax, int idx, Tab &eq, Tab &lb, Tab &ub) {
    for (int i = 0; i < (int)cmin->dim; i++, idx++) {  // Hqp_Docp::parse_constr, hqp/Hqp_Docp.C:370-397
      if (cmin[i] == cmax[i]) {
        if (!is_finite(cmin[i])) m_error(E_FORMAT, "Hqp_DocpCuda::setup: Periodical states are not supported");
        eq.idx.push_back(idx); eq.val.push_back(cmin[i]);
      } else {
        if (cmin[i] > -Inf) { lb.idx.push_back(idx); lb.val.push_back(cmin[i]); }
        if (cmax[i] < Inf) { ub.idx.push_back(idx); ub.val.push_back(cmax[i]); }
      }
    }
  }

  // existing entries of a sparse row with columns j_offs .. j_offs+n-1 <- sign * src
  static void put(SPROW *row, int j_offs, int n, const double *src, double sign) {
    int j_idx = sprow_idx(row, j_offs);
    if (j_idx < 0) {
      if (j_idx == -1) return;
      j_idx = -(j_idx + 2);
    }
    for (; j_idx < row->len; j_idx++) {
      const int j = row->elt[j_idx].col - j_offs;
      if (j >= n) break;
      row->elt[j_idx].val = sign * src[j];
    }
  }

  void cns_rows(SPMAT *M, int row0, const std::vector<int> &idx, double sign) {
    const int nd = _nx + _nu;
    for (size_t i = 0; i < idx.size(); i++) {
      const int ci = idx[i];
      const int k = (_nc > 0 && ci < _K * _nc) ? ci / _nc : _K;
      SPROW *row = M->row + row0 + (int)i;
      put(row, k * nd, _nx, &_cx[(size_t)ci * _nx], sign);
      if (k < _K) put(row, k * nd + _nx, _nu, &_cu[(size_t)ci * _nu], sign);
    }
  }
};

#endif
