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
  _padded = false;
  _n_abi = _me_abi = 0;

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
    const int c = r->elt[r->len - 1].col, c1 = r->elt[r->len - 2].col;
    if (c <= prev)
      break;
    // first row of a new stage: a gap in the -1 columns (the controls in between) or,
    // for a stage without controls, a row whose other entries already lie in the block
    // the -1 columns of the running stage point to (Get_Dim, hqp/Hqp_IpLQDOCP.C:223)
    if (prev < 0 || c - prev > 1 || c - c1 < rows_in_stage) {
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
  // per-stage dimensions: nxs[k] states, nus[k] controls (nus[K] = 0); the number of
  // states of stage 0 is not visible in A: like Get_Dim (hqp/Hqp_IpLQDOCP.C:266) it is
  // taken as min(nx_1, variables of stage 0)
  std::vector<int> nxs(_K + 1), nus(_K + 1, 0), dyn0(_K + 1, 0);
  for (k = 0; k < _K; k++) nxs[k + 1] = nxk[k];
  nxs[0] = nxs[1] < xoff[1] - xoff[0] ? nxs[1] : xoff[1] - xoff[0];
  for (k = 0; k < _K; k++) {
    nus[k] = xoff[k + 1] - xoff[k] - nxs[k];
    dyn0[k + 1] = dyn0[k] + nxs[k + 1];
    if (nus[k] < 0)
      m_error(E_SIZES, "Hqp_IpCuda::init: unsupported stage dimensions");
  }
  if (xoff[_K] + nxs[_K] != N)
    m_error(E_SIZES, "Hqp_IpCuda::init: unsupported stage dimensions");
  _nx = 0;
  _nu = 1;
  _padded = false;
  for (k = 0; k <= _K; k++) {
    if (nxs[k] > _nx) _nx = nxs[k];
    if (nus[k] > _nu) _nu = nus[k];
  }
  for (k = 0; k <= _K; k++)
    _padded = _padded || nxs[k] != _nx || (k < _K && nus[k] != _nu);
  if (_padded && _ngpu > 1)
    m_error(E_SIZES, "Hqp_IpCuda::init: mat_ngpu with non-uniform stage dimensions");
  const int nm = _nx + _nu;
  _n = N;
  _me = A->m;
  _m = C->m;
  _n_abi = _K * nm + _nx;
  // variable -> slot of the padded layout; stage of a variable
  _xmap.assign(N, 0);
  std::vector<int> &vstage = _vstage;
  vstage.assign(N, 0);
  for (k = 0; k <= _K; k++) {
    const int cnt = (k < _K ? xoff[k + 1] : N) - xoff[k];
    for (i = 0; i < cnt; i++) {
      _xmap[xoff[k] + i] = k * nm + (i < nxs[k] ? i : _nx + (i - nxs[k]));
      vstage[xoff[k] + i] = k;
    }
  }
  // local column of a variable inside its (padded) stage block
#define LCOL(col) (_xmap[col] - _vstage[col] * nm)

  // --- every dynamics row must live in [x_k u_k | x_{k+1}]
  for (k = 0; k < _K; k++)
    for (i = 0; i < nxs[k + 1]; i++) {
      const SPROW *r = A->row + dyn0[k] + i;
      if (r->elt[0].col < xoff[k] || r->elt[r->len - 2].col >= xoff[k + 1] ||
          r->elt[r->len - 1].col != xoff[k + 1] + i)
        m_error(E_FORMAT, "Hqp_IpCuda::init: dynamics row leaves its stage");
    }

  // --- remaining equality rows: x0 fixing rows, then general stage rows
  std::vector<int> rest_stage;
  for (i = ndyn; i < (int)A->m; i++) {
    const SPROW *r = A->row + i;
    if (r->len < 1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: empty equality row");
    const int k0 = vstage[r->elt[0].col], k1 = vstage[r->elt[r->len - 1].col];
    if (k0 != k1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: equality row couples stages");
    rest_stage.push_back(k0);
  }
  // fixed initial state: exactly nx_0 stage-0 rows, the j-th one being +1 at x0_j
  std::vector<int> rows0;
  for (i = 0; i < (int)rest_stage.size(); i++)
    if (rest_stage[i] == 0) rows0.push_back(ndyn + i);
  _fixed_x0 = ((int)rows0.size() == nxs[0]);
  for (i = 0; _fixed_x0 && i < nxs[0]; i++) {
    const SPROW *r = A->row + rows0[i];
    _fixed_x0 = (r->len == 1 && r->elt[0].col == i && r->elt[0].val == 1.0);
  }
  // rows of the device layout -> rows of qp->A (-1: a padded row, right-hand side 0)
  _rowmap.clear();
  for (k = 0; k < _K; k++)
    for (i = 0; i < _nx; i++) _rowmap.push_back(i < nxs[k + 1] ? dyn0[k] + i : -1);
  if (_fixed_x0)
    for (i = 0; i < _nx; i++) _rowmap.push_back(i < nxs[0] ? rows0[i] : -1);
  std::vector<int> eq_stage, eq_lcol;
  _eq_ptr.assign(1, 0);
  for (i = 0; i < (int)rest_stage.size(); i++) {
    const int row = ndyn + i;
    bool is_x0 = false;
    if (_fixed_x0)
      for (k = 0; k < nxs[0]; k++) is_x0 = is_x0 || rows0[k] == row;
    if (is_x0) continue;
    const SPROW *r = A->row + row;
    _rowmap.push_back(row);
    eq_stage.push_back(rest_stage[i]);
    for (k = 0; k < r->len; k++) eq_lcol.push_back(LCOL(r->elt[k].col));
    _eq_ptr.push_back((int)eq_lcol.size());
  }
  _n_eq = (int)eq_stage.size();
  _me_abi = (int)_rowmap.size();
  _identity_rows = !_padded && _me_abi == _me;
  for (i = 0; _identity_rows && i < _me; i++) _identity_rows = _rowmap[i] == i;

  // --- inequality rows
  std::vector<int> ineq_stage(_m), ineq_lcol;
  _ineq_ptr.assign(1, 0);
  for (i = 0; i < _m; i++) {
    const SPROW *r = C->row + i;
    if (r->len < 1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: empty inequality row");
    const int k0 = vstage[r->elt[0].col], k1 = vstage[r->elt[r->len - 1].col];
    if (k0 != k1)
      m_error(E_FORMAT, "Hqp_IpCuda::init: inequality row couples stages");
    ineq_stage[i] = k0;
    for (k = 0; k < r->len; k++) ineq_lcol.push_back(LCOL(r->elt[k].col));
    _ineq_ptr.push_back((int)ineq_lcol.size());
  }

  // --- Q must be block diagonal over the stages (upper triangle stored)
  for (i = 0; i < N; i++) {
    const SPROW *r = Q->row + i;
    if (r->len == 0) continue;
    if (vstage[r->elt[0].col] != vstage[i] || vstage[r->elt[r->len - 1].col] != vstage[i])
      m_error(E_FORMAT, "Hqp_IpCuda::init: Q couples stages");
  }
  // padded variables: unit diagonal of Q (decoupled, positive definite)
  _pad_diag.clear();
  if (_padded) {
    std::vector<char> used((size_t)_n_abi, 0);
    for (i = 0; i < N; i++) used[_xmap[i]] = 1;
    for (i = 0; i < _n_abi; i++)
      if (!used[i]) {
        k = i / nm < _K ? i / nm : _K;
        const int l = i - k * nm;
        _pad_diag.push_back((long long)k * nm * nm + (long long)l * nm + l);
      }
    // (the u block of stage K does not exist in the device layout either)
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
  _r2p.assign(_me_abi > 0 ? _me_abi : 1, 0.0);
  _dyp.assign(_me_abi > 0 ? _me_abi : 1, 0.0);
  _r1p.assign(_n_abi, 0.0);
  _dxp.assign(_n_abi, 0.0);
  // stage geometry kept for update()
  _xoff = xoff;
  _nxs = nxs;
  _dyn0 = dyn0;

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
    k = _vstage[i];
    const int li = LCOL(i);
    for (j = 0; j < r->len; j++) {
      if (_vstage[r->elt[j].col] != k)
        m_error(E_FORMAT, "Hqp_IpCuda::init: Q couples stages");
      const int lj = LCOL(r->elt[j].col);
      dst.push_back((long long)k * nm * nm + (long long)li * nm + lj);
      dst2.push_back(li != lj ? (long long)k * nm * nm + (long long)lj * nm + li : -1);
    }
  }
  for (k = 0; k < _K; k++)
    for (int li = 0; li < _nxs[k + 1]; li++) {
      const SPROW *r = A->row + _dyn0[k] + li;
      for (j = 0; j < r->len - 1; j++) {
        if (_vstage[r->elt[j].col] != k)
          m_error(E_FORMAT, "Hqp_IpCuda::init: dynamics row leaves its stage");
        const int lc = LCOL(r->elt[j].col);
        dst.push_back(lc < _nx ? szQ + ((long long)k * _nx + li) * _nx + lc
                               : szQ + szX + ((long long)k * _nx + li) * _nu + lc - _nx);
        dst2.push_back(-1);
      }
    }
  // unit diagonal of the padded variables (constant values at the end of the list)
  for (size_t q = 0; q < _pad_diag.size(); q++) {
    dst.push_back(_pad_diag[q]);
    dst2.push_back(-1);
  }
  _vals.assign(dst.size() > 0 ? dst.size() : 1, 0.0);
  for (size_t q = 0; q < _pad_diag.size(); q++) _vals[dst.size() - _pad_diag.size() + q] = 1.0;
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

  // general equality rows and C (both paths)
  for (i = 0; i < _n_eq; i++) {
    const SPROW *r = A->row + _rowmap[_me_abi - _n_eq + i];
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

  if (_sparse_update) {
    // values in the order of build_value_map(); the device scatters them
    const size_t nvar = _vals.size() - _pad_diag.size();
    size_t p = 0;
    for (i = 0; i < _n; i++) {
      const SPROW *r = Q->row + i;
      for (j = 0; j < r->len; j++, p++) {
        if (p >= nvar) m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of Q changed");
        _vals[p] = r->elt[j].val;
      }
    }
    for (k = 0; k < _K; k++)
      for (int li = 0; li < _nxs[k + 1]; li++) {
        const SPROW *r = A->row + _dyn0[k] + li;
        for (j = 0; j < r->len - 1; j++, p++) {
          if (p >= nvar) m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of A changed");
          _vals[p] = r->elt[j].val;
        }
        if (r->len < 1 || r->elt[r->len - 1].val != -1.0)
          m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row lost its -1");
      }
    if (p != nvar && !(p == 0 && _vals.size() == 1))
      m_error(E_FORMAT, "Hqp_IpCuda::update: pattern of Q or A changed");
    check(hqpcu_update_values(_h, &_vals[0], &_cval[0], &_eval[0]), "Hqp_IpCuda::update");
    return;
  }

  // Q: upper triangle -> full symmetric stage blocks
  std::fill(_Q.begin(), _Q.end(), 0.0);
  for (i = 0; i < _n; i++) {
    const SPROW *r = Q->row + i;
    k = _vstage[i];
    double *blk = &_Q[(size_t)k * nm * nm];
    const int li = LCOL(i);
    for (j = 0; j < r->len; j++) {
      if (_vstage[r->elt[j].col] != k)
        m_error(E_FORMAT, "Hqp_IpCuda::update: Q couples stages");
      const int lj = LCOL(r->elt[j].col);
      blk[li * nm + lj] = r->elt[j].val;
      blk[lj * nm + li] = r->elt[j].val;
    }
  }
  for (size_t q = 0; q < _pad_diag.size(); q++) _Q[(size_t)_pad_diag[q]] = 1.0;
  // A: dynamics rows -> fx, fu
  std::fill(_fx.begin(), _fx.end(), 0.0);
  std::fill(_fu.begin(), _fu.end(), 0.0);
  for (k = 0; k < _K; k++)
    for (int li = 0; li < _nxs[k + 1]; li++) {
      const SPROW *r = A->row + _dyn0[k] + li;
      for (j = 0; j < r->len - 1; j++) {
        if (_vstage[r->elt[j].col] != k)
          m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row leaves its stage");
        const int lc = LCOL(r->elt[j].col);
        if (lc < _nx)
          _fx[((size_t)k * _nx + li) * _nx + lc] = r->elt[j].val;
        else
          _fu[((size_t)k * _nx + li) * _nu + lc - _nx] = r->elt[j].val;
      }
      if (r->elt[r->len - 1].val != -1.0)
        m_error(E_FORMAT, "Hqp_IpCuda::update: dynamics row lost its -1");
    }
  check(hqpcu_update(_h, &_Q[0], &_fx[0], &_fu[0], &_cval[0], &_eval[0]),
        "Hqp_IpCuda::update");
}
#undef LCOL

//--------------------------------------------------------------------------
void Hqp_IpCuda::factor(const Hqp_Program *, const VEC *z, const VEC *w)
{
  assert((int)z->dim == _m && (int)w->dim == _m);
  static double none = 0.0;
  check(hqpcu_factor(_h, _m ? z->ve : &none, _m ? w->ve : &none), "Hqp_IpCuda::factor");
}

//--------------------------------------------------------------------------
void Hqp_IpCuda::to_abi_y(const double *user, double *abi) const
{
  for (int i = 0; i < _me_abi; i++) abi[i] = _rowmap[i] >= 0 ? user[_rowmap[i]] : 0.0;
}

void Hqp_IpCuda::from_abi_y(const double *abi, double *user) const
{
  for (int i = 0; i < _me_abi; i++)
    if (_rowmap[i] >= 0) user[_rowmap[i]] = abi[i];
}

void Hqp_IpCuda::to_abi_x(const double *user, double *abi) const
{
  if (_padded)
    for (int i = 0; i < _n_abi; i++) abi[i] = 0.0;
  for (int i = 0; i < _n; i++) abi[_xmap[i]] = user[i];
}

void Hqp_IpCuda::from_abi_x(const double *abi, double *user) const
{
  for (int i = 0; i < _n; i++) user[i] = abi[_xmap[i]];
}

const double *Hqp_IpCuda::pack_r2(const VEC *r2)
{
  if (_identity_rows)
    return r2->ve;
  to_abi_y(r2->ve, &_r2p[0]);
  return &_r2p[0];
}

void Hqp_IpCuda::unpack_dy(VEC *dy)
{
  if (_identity_rows)
    return;
  from_abi_y(&_dyp[0], dy->ve);
}

const double *Hqp_IpCuda::pack_r1(const VEC *r1)
{
  if (!_padded)
    return r1->ve;
  to_abi_x(r1->ve, &_r1p[0]);
  return &_r1p[0];
}

void Hqp_IpCuda::unpack_dx(VEC *dx)
{
  if (_padded)
    from_abi_x(&_dxp[0], dx->ve);
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
  check(hqpcu_step(_h, pack_r1(r1), pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                   _padded ? &_dxp[0] : dx->ve, _identity_rows ? dy->ve : &_dyp[0],
                   _m ? dz->ve : none, _m ? dw->ve : none),
        "Hqp_IpCuda::step");
  unpack_dx(dx);
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
  check(hqpcu_solve(_h, _eps, pack_r1(r1), pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                    _padded ? &_dxp[0] : dx->ve, _identity_rows ? dy->ve : &_dyp[0],
                    _m ? dz->ve : none, _m ? dw->ve : none, &res, NULL),
        "Hqp_IpCuda::solve");
  unpack_dx(dx);
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
  const double *dyv = dy->ve, *dxv = dx->ve;
  if (!_identity_rows) {
    to_abi_y(dy->ve, &_dyp[0]);
    dyv = &_dyp[0];
  }
  if (_padded) {
    to_abi_x(dx->ve, &_dxp[0]);
    dxv = &_dxp[0];
  }
  check(hqpcu_residuum(_h, pack_r1(r1), pack_r2(r2), _m ? r3->ve : none, _m ? r4->ve : none,
                       dxv, dyv, _m ? dz->ve : none, _m ? dw->ve : none, &res),
        "Hqp_IpCuda::residuum");
  return res;
}
