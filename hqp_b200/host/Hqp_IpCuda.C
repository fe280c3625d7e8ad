/*
 * Hqp_IpCuda.C -- class definition (see Hqp_IpCuda.h)
 *
 * Host side of the drop-in boundary.  What the reference's own stage-structured
 * module does on the CPU (hqp/Hqp_IpLQDOCP.C) is split here into
 *   - structure detection from the sparsity of A, C, Q (init),
 *   - packing of the SPMAT values into contiguous stage slabs (update),
 *   - forwarding of factor / step / solve / residuum to libhqpcuda.so.
 * Numerical failure is reported the reference's way: the C ABI returns
 * HQPCU_E_SING and this class raises m_error(E_SING, ...) after the call has
 * returned, so that Hqp_IpsMehrotra / Hqp_IpsFranke catch it with m_catch and
 * set Hqp_Degenerate (hqp/Hqp_IpsMehrotra.C:524-536).
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include <If_Int.h>

#include "Hqp_Program.h"
#include "Hqp_IpCuda.h"

#include "hqp_ipcuda.h"

IF_CLASS_DEFINE("Cuda", Hqp_IpCuda, Hqp_IpMatrix);

//--------------------------------------------------------------------------
Hqp_IpCuda::Hqp_IpCuda()
{
  _h = NULL;
  _nseg = 0;
  _device = 0;
  _ngpu = 1;
  _dev_solve = 1;
  _sparse_update = 1;
  _K = _nx = _nu = _n = _me = _m = 0;
  _fixed_x0 = 0;
  _n_eq = 0;
  _identity_rows = true;

  _ifList.append(new If_Int("mat_nseg", &_nseg));
  _ifList.append(new If_Int("mat_device", &_device));
  _ifList.append(new If_Int("mat_ngpu", &_ngpu));
  _ifList.append(new If_Int("mat_dev_solve", &_dev_solve));
  _ifList.append(new If_Int("mat_sparse_update", &_sparse_update));
}

//--------------------------------------------------------------------------
Hqp_IpCuda::~Hqp_IpCuda()
{
  free_handle();
}

void Hqp_IpCuda::free_handle()
{
  if (_h) {
    hqpcu_destroy(_h);
    _h = NULL;
  }
}

//--------------------------------------------------------------------------
void Hqp_IpCuda::check(int status, const char *where)
{
  if (status == HQPCU_OK)
    return;
  if (status == HQPCU_E_SING)
    m_error(E_SING, where);
  fprintf(stderr, "Hqp_IpCuda: %s failed with status %d: %s\n", where, status,
          hqpcu_last_error());
  if (status == HQPCU_E_SIZES || status == HQPCU_E_UNSUPPORTED)
    m_error(E_SIZES, where);
  if (status == HQPCU_E_NULL)
    m_error(E_NULL, where);
  m_error(E_INTERN, where);
}

//--------------------------------------------------------------------------
//   Derive the stage structure from the sparsity of qp->A, qp->C, qp->Q.
//   Layout expected (hqp/Hqp_Docp.C:465-755): variables [x0,u0,x1,...,xK];
//   the first rows of A are the dynamics [fx fu -I], recognisable by their
//   trailing -1.0 on consecutive columns.
//--------------------------------------------------------------------------
void Hqp_IpCuda::init(const Hqp_Program *qp)
{
  const SPMAT *A = qp->A, *C = qp->C, *Q = qp->Q;
  const int N = A->n;
  int i, k;

  assert(A->m > 0 && N == (int)Q->n);

  // --- dynamics rows and stage offsets
  std::vector<int> xoff(1, 0);   // first column of x_k
  std::vector<int> nxk;          // states of stage k+1 = rows of stage k
  int prev = -1, rows_in_stage = 0, ndyn = 0;
  bool closed = false;
  for (i = 0; i < (int)A->m; i++) {
    const SPROW *r = A->row + i;
    if (r->len < 2 || r->elt[r->len - 1].val != -1.0)
      break;
    const int c = r->elt[r->len - 1].col;
    if (c <= prev)
      break;
    if (prev < 0 || c - prev > 1) {           // first row of a new stage
      if (prev >= 0) nxk.push_back(rows_in_stage);
      xoff.push_back(c);
      rows_in_stage = 1;
    } else
      rows_in_stage++;
    prev = c;
    ndyn = i + 1;
    if (c == N - 1) {
      nxk.push_back(rows_in_stage);
      closed = true;
      break;
    }
  }
  if (!closed || nxk.empty())
    m_error(E_FORMAT, "Hqp_IpCuda::init: no DOCP structure in A");
  _K = (int)nxk.size();
  _nx = nxk[0];
  _nu = xoff[1] - _nx;
  for (k = 0; k < _K; k++) {
    const int nuk = (k + 1 < (int)xoff.size() - 0 ? xoff[k + 1] - xoff[k] : 0) - _nx;
    if (nxk[k] != _nx || nuk != _nu)
      m_error(E_SIZES, "Hqp_IpCuda::init: non-uniform stage dimensions");
  }
  if (_nu < 1 || xoff[_K] + _nx != N)
    m_error(E_SIZES, "Hqp_IpCuda::init: unsupported stage dimensions");
  const int nm = _nx + _nu;
  _n = N;
  _me = A->m;
  _m = C->m;

  // --- every dynamics row must live in [x_k u_k | x_{k+1}]
  for (i = 0; i < ndyn; i++) {
    const SPROW *r = A->row + i;
    k = i / _nx;
    if (r->elt[0].col < k * nm || r->elt[r->len - 2].col >= (k + 1) * nm ||
        r->elt[r->len - 1].col != (k + 1) * nm + i % _nx)
      m_error(E_FORMAT, "Hqp_IpCuda::init: dynamics row leaves its stage");
  }

  // --- remaining equality rows: x0 fixing rows, then general stage rows
  std::vector<int> rest_stage;
  for (i = ndyn; i < (int)A->m; i++) {
    const SPROW *r = A->row + i;
    if (r->len < 1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: empty equality row");
    const int k0 = r->elt[0].col / nm < _K ? r->elt[0].col / nm : _K;
    const int k1 = r->elt[r->len - 1].col / nm < _K ? r->elt[r->len - 1].col / nm : _K;
    if (k0 != k1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: equality row couples stages");
    rest_stage.push_back(k0);
  }
  // fixed initial state: exactly nx stage-0 rows, the j-th one being +1 at x0_j
  std::vector<int> rows0;
  for (i = 0; i < (int)rest_stage.size(); i++)
    if (rest_stage[i] == 0) rows0.push_back(ndyn + i);
  _fixed_x0 = ((int)rows0.size() == _nx);
  for (i = 0; _fixed_x0 && i < _nx; i++) {
    const SPROW *r = A->row + rows0[i];
    _fixed_x0 = (r->len == 1 && r->elt[0].col == i && r->elt[0].val == 1.0);
  }
  _rowmap.clear();
  for (i = 0; i < ndyn; i++) _rowmap.push_back(i);
  if (_fixed_x0)
    for (i = 0; i < _nx; i++) _rowmap.push_back(rows0[i]);
  std::vector<int> eq_stage, eq_lcol;
  _eq_ptr.assign(1, 0);
  for (i = 0; i < (int)rest_stage.size(); i++) {
    const int row = ndyn + i;
    bool is_x0 = false;
    if (_fixed_x0)
      for (k = 0; k < _nx; k++) is_x0 = is_x0 || rows0[k] == row;
    if (is_x0) continue;
    const SPROW *r = A->row + row;
    _rowmap.push_back(row);
    eq_stage.push_back(rest_stage[i]);
    for (k = 0; k < r->len; k++) eq_lcol.push_back(r->elt[k].col - rest_stage[i] * nm);
    _eq_ptr.push_back((int)eq_lcol.size());
  }
  _n_eq = (int)eq_stage.size();
  _identity_rows = true;
  for (i = 0; i < _me; i++) _identity_rows = _identity_rows && _rowmap[i] == i;

  // --- inequality rows
  std::vector<int> ineq_stage(_m), ineq_lcol;
  _ineq_ptr.assign(1, 0);
  for (i = 0; i < _m; i++) {
    const SPROW *r = C->row + i;
    if (r->len < 1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: empty inequality row");
    const int k0 = r->elt[0].col / nm < _K ? r->elt[0].col / nm : _K;
    const int k1 = r->elt[r->len - 1].col / nm < _K ? r->elt[r->len - 1].col / nm : _K;
    if (k0 != k1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: inequality row couples stages");
    ineq_stage[i] = k0;
    for (k = 0; k < r->len; k++) ineq_lcol.push_back(r->elt[k].col - k0 * nm);
    _ineq_ptr.push_back((int)ineq_lcol.size());
  }

  // --- Q must be block diagonal over the stages (upper triangle stored)
  for (i = 0; i < N; i++) {
    const SPROW *r = Q->row + i;
    if (r->len == 0) continue;
    const int k0 = i / nm < _K ? i / nm : _K;
    const int kl = r->elt[r->len - 1].col / nm < _K ? r->elt[r->len - 1].col / nm : _K;
    const int kf = r->elt[0].col / nm < _K ? r->elt[0].col / nm : _K;
    if (kf != k0 || kl != k0)
      m_error(E_FORMAT, "Hqp_IpCuda::init: Q couples stages");
  }

  // --- (re)create the device engine
  free_handle();
  hqpcu_dims dims;
  memset(&dims, 0, sizeof dims);
  dims.K = _K; dims.nx = _nx; dims.nu = _nu; dims.batch = 1;
  dims.fixed_x0 = _fixed_x0;
  dims.n_ineq = _m;
  dims.ineq_stage = _m ? &ineq_stage[0] : NULL;
  dims.ineq_ptr = &_ineq_ptr[0];
  dims.ineq_lcol = ineq_lcol.empty() ? NULL : &ineq_lcol[0];
  static const int zero = 0;
  if (!_m) { dims.ineq_stage = &zero; dims.ineq_lcol = &zero; }
  dims.n_eq = _n_eq;
  dims.eq_stage = _n_eq ? &eq_stage[0] : NULL;
  dims.eq_ptr = &_eq_ptr[0];
  dims.eq_lcol = _n_eq ? &eq_lcol[0] : NULL;
  dims.device = _device;
  dims.nseg = _nseg;
  dims.ngpu = _ngpu > 1 ? _ngpu : 0;
  check(hqpcu_create(&dims, &_h), "Hqp_IpCuda::init");

  _Q.assign((size_t)(_K + 1) * nm * nm, 0.0);
  _fx.assign((size_t)_K * _nx * _nx, 0.0);
  _fu.assign((size_t)_K * _nx * _nu, 0.0);
  _cval.assign(_ineq_ptr.back() > 0 ? _ineq_ptr.back() : 1, 0.0);
  _eval.assign(_eq_ptr.back() > 0 ? _eq_ptr.back() : 1, 0.0);
  _r2p.assign(_me > 0 ? _me : 1, 0.0);
  _dyp.assign(_me > 0 ? _me : 1, 0.0);

  if (_ngpu > 1) _sparse_update = 0;  // (the dispatcher takes dense stage slabs)
  if (_sparse_update) build_value_map(qp);
  update(qp);
}

//--------------------------------------------------------------------------
//   where every stored entry of Q (upper triangle) and of the
//   dynamics rows of A lives in the device slabs.  Registered once per init();
//   update() then uploads the values only and the device scatters them
//   (replaces the sp_extract_mat walk of hqp/Hqp_IpLQDOCP.C:747-755).
void Hqp_IpCuda::build_value_map(const Hqp_Program *qp)
{
  const SPMAT *A = qp->A, *Q = qp->Q;
  const int nm = _nx + _nu;
  const long long szQ = (long long)(_K + 1) * nm * nm, szX = (long long)_K * _nx * _nx;
  std::vector<long long> dst, dst2;
  int i, j, k;
  for (i = 0; i < _n; i++) {
    const SPROW *r = Q->row + i;
    k = i / nm < _K ? i / nm : _K;
    const int li = i - k * nm;
    for (j = 0; j < r->len; j++) {
      const int lj = r->elt[j].col - k * nm;
      if (lj < 0 || lj >= nm)
        m_error(E_FORMAT, "Hqp_IpCuda::init: Q couples stages");
      dst.push_back((long long)k * nm * nm + (long long)li * nm + lj);
      dst2.push_back(li != lj ? (long long)k * nm * nm + (long long)lj * nm + li : -1);
    }
  }
  for (i = 0; i < _K * _nx; i++) {
    const SPROW *r = A->row + i;
    k = i / _nx;
    const int li = i - k * _nx;
    for (j = 0; j < r->len - 1; j++) {
      const int lc = r->elt[j].col - k * nm;
      if (lc < 0 || lc >= nm)
        m_error(E_FORMAT, "Hqp_IpCuda::init: dynamics row leaves its stage");
      dst.push_back(lc < _nx ? szQ + ((long long)k * _nx + li) * _nx + lc
                             : szQ + szX + ((long long)k * _nx + li) * _nu + lc - _nx);
      dst2.push_back(-1);
    }
  }
  _vals.assign(dst.size() > 0 ? dst.size() : 1, 0.0);
  check(hqpcu_set_value_map(_h, (long long)dst.size(), dst.empty() ? NULL : &dst[0],
                            dst2.empty() ? NULL : &dst2[0]), "Hqp_IpCuda::init");
}

//--------------------------------------------------------------------------
//   Once per SQP iteration: values of Q, A, C -> contiguous stage slabs
//   (the counterpart of Hqp_IpLQDOCP::update, hqp/Hqp_IpLQDOCP.C:722-787).
//--------------------------------------------------------------------------
void Hqp_IpCuda::update(const Hqp_Program *qp)
{
  const SPMAT *A = qp->A, *C = qp->C, *Q = qp->Q;
  const int nm = _nx + _nu;
  int i, j, k;

  assert(_h != NULL && (int)A->m == _me && (int)C->m == _m && (int)Q->n == _n);

  if (_sparse_update) {
    // values in the order of build_value_map(); the device scatters them
    size_t p = 0;
    for (i = 0; i < _n; i++) {
      const SPROW *r = Q->row + i;
      for (j = 0; j < r->len; j++, p++) {
        if (p >= _vals.size()) m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of Q changed");
        _vals[p] = r->elt[j].val;
      }
    }
    for (i = 0; i < _K * _nx; i++) {
      const SPROW *r = A->row + i;
      for (j = 0; j < r->len - 1; j++, p++) {
        if (p >= _vals.size()) m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of A changed");
        _vals[p] = r->elt[j].val;
      }
      if (r->len < 1 || r->elt[r->len - 1].val != -1.0)
        m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row lost its -1");
    }
    if (p != _vals.size() && !(p == 0 && _vals.size() == 1))
      m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of Q or A changed");
    for (i = 0; i < _n_eq; i++) {
      const SPROW *r = A->row + _rowmap[_me - _n_eq + i];
      if (r->len != _eq_ptr[i + 1] - _eq_ptr[i])
        m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of A changed");
      for (j = 0; j < r->len; j++) _eval[_eq_ptr[i] + j] = r->elt[j].val;
    }
    for (i = 0; i < _m; i++) {
      const SPROW *r = C->row + i;
      if (r->len != _ineq_ptr[i + 1] - _ineq_ptr[i])
        m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of C changed");
      for (j = 0; j < r->len; j++) _cval[_ineq_ptr[i] + j] = r->elt[j].val;
    }
    check(hqpcu_update_values(_h, &_vals[0], &_cval[0], &_eval[0]), "Hqp_IpCuda::update");
    return;
  }

  // Q: upper triangle -> full symmetric stage blocks
  std::fill(_Q.begin(), _Q.end(), 0.0);
  for (i = 0; i < _n; i++) {
    const SPROW *r = Q->row + i;
    k = i / nm < _K ? i / nm : _K;
    double *blk = &_Q[(size_t)k * nm * nm];
    const int li = i - k * nm;
    for (j = 0; j < r->len; j++) {
      const int lj = r->elt[j].col - k * nm;
      if (lj < 0 || lj >= nm)
        m_error(E_FORMAT, "Hqp_IpCuda::update: Q couples stages");
      blk[li * nm + lj] = r->elt[j].val;
      blk[lj * nm + li] = r->elt[j].val;
    }
  }
  // A: dynamics rows -> fx, fu
  std::fill(_fx.begin(), _fx.end(), 0.0);
  std::fill(_fu.begin(), _fu.end(), 0.0);
  for (i = 0; i < _K * _nx; i++) {
    const SPROW *r = A->row + i;
    k = i / _nx;
    const int li = i - k * _nx;
    for (j = 0; j < r->len - 1; j++) {
      const int lc = r->elt[j].col - k * nm;
      if (lc < 0 || lc >= nm)
        m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row leaves its stage");
      if (lc < _nx)
        _fx[((size_t)k * _nx + li) * _nx + lc] = r->elt[j].val;
      else
        _fu[((size_t)k * _nx + li) * _nu + lc - _nx] = r->elt[j].val;
    }
    if (r->elt[r->len - 1].val != -1.0)
      m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row lost its -1");
  }
  // general equality rows
  for (i = 0; i < _n_eq; i++) {
    const SPROW *r = A->row + _rowmap[_me - _n_eq + i];
    if (r->len != _eq_ptr[i + 1] - _eq_ptr[i])
      m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of A changed");
    for (j = 0; j < r->len; j++) _eval[_eq_ptr[i] + j] = r->elt[j].val;
  }
  // C
  for (i = 0; i < _m; i++) {
    const SPROW *r = C->row + i;
    if (r->len != _ineq_ptr[i + 1] - _ineq_ptr[i])
      m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of C changed");
    for (j = 0; j < r->len; j++) _cval[_ineq_ptr[i] + j] = r->elt[j].val;
  }
  check(hqpcu_update(_h, &_Q[0], &_fx[0], &_fu[0], &_cval[0], &_eval[0]),
        "Hqp_IpCuda::update");
}

//--------------------------------------------------------------------------
void Hqp_IpCuda::factor(const Hqp_Program *, const VEC *z, const VEC *w)
{
  assert((int)z->dim == _m && (int)w->dim == _m);
  static double none = 0.0;
  check(hqpcu_factor(_h, _m ? z->ve : &none, _m ? w->ve : &none), "Hqp_IpCuda::factor");
}

//--------------------------------------------------------------------------
const double *Hqp_IpCuda::pack_r2(const VEC *r2)
{
  if (_identity_rows)
    return r2->ve;
  for (int i = 0; i < _me; i++) _r2p[i] = r2->ve[_rowmap[i]];
  return &_r2p[0];
}

void Hqp_IpCuda::unpack_dy(VEC *dy)
{
  if (_identity_rows)
    return;
  for (int i = 0; i < _me; i++) dy->ve[_rowmap[i]] = _dyp[i];
}

//--------------------------------------------------------------------------
void Hqp_IpCuda::step(const Hqp_Program *, const VEC *, const VEC *,
                      const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
                      VEC *dx, VEC *dy, VEC *dz, VEC *dw)
{
  assert((int)r1->dim == _n && (int)dx->dim == _n);
  assert((int)r2->dim == _me && (int)dy->dim == _me);
  assert((int)r3->dim == _m && (int)dz->dim == _m);
  assert((int)r4->dim == _m && (int)dw->dim == _m);
  static double none[1];
  check(hqpcu_step(_h, r1->ve, pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                   dx->ve, _identity_rows ? dy->ve : &_dyp[0], _m ? dz->ve : none,
                   _m ? dw->ve : none),
        "Hqp_IpCuda::step");
  unpack_dy(dy);
}

//--------------------------------------------------------------------------
//   step + iterative refinement (Hqp_IpMatrix::solve, hqp/Hqp_IpMatrix.C:65-128)
//   with all intermediate vectors resident on the device
//--------------------------------------------------------------------------
Real Hqp_IpCuda::solve(const Hqp_Program *qp, const VEC *z, const VEC *w,
                       const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
                       VEC *dx, VEC *dy, VEC *dz, VEC *dw)
{
  if (!_dev_solve)
    return Hqp_IpMatrix::solve(qp, z, w, r1, r2, r3, r4, dx, dy, dz, dw);
  static double none[1];
  double res = 0.0;
  check(hqpcu_solve(_h, _eps, r1->ve, pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                    dx->ve, _identity_rows ? dy->ve : &_dyp[0], _m ? dz->ve : none,
                    _m ? dw->ve : none, &res, NULL),
        "Hqp_IpCuda::solve");
  unpack_dy(dy);
  return res;
}

//--------------------------------------------------------------------------
Real Hqp_IpCuda::residuum(const Hqp_Program *, const VEC *, const VEC *,
                          const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
                          VEC *dx, VEC *dy, VEC *dz, VEC *dw)
{
  static double none[1];
  double res = 0.0;
  const double *dyv = dy->ve;
  if (!_identity_rows) {
    for (int i = 0; i < _me; i++) _dyp[i] = dy->ve[_rowmap[i]];
    dyv = &_dyp[0];
  }
  check(hqpcu_residuum(_h, r1->ve, pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                       dx->ve, dyv, _m ? dz->ve : none, _m ? dw->ve : none, &res),
        "Hqp_IpCuda::residuum");
  return res;
}
