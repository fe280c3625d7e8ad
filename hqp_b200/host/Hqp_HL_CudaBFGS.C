/*
 * Hqp_HL_CudaBFGS.C -- see Hqp_HL_CudaBFGS.h.  New code against the reference's
 * public headers; the numerics live behind the C ABI of include/hqp_hlcuda.h.
 */
#include "Hqp_HL_CudaBFGS.h"

#include <string.h>

extern "C" {
#include <meschach/addon2_hqp.h>
#include <meschach/sparse.h>
}

#include <If_Int.h>

#include "Hqp_Program.h"
#include "Hqp_SqpProgram.h"
#include "hqp_hlcuda.h"

IF_CLASS_DEFINE("CudaBFGS", Hqp_HL_CudaBFGS, Hqp_HL);

Hqp_HL_CudaBFGS::Hqp_HL_CudaBFGS() {
  _device = 0;
  _ifList.append(new If_Int("sqp_hela_device", &_device));
}

void Hqp_HL_CudaBFGS::update(const VEC *s, const VEC *u, Real alpha, Hqp_SqpProgram *prg) {
  SPMAT *Q = prg->qp()->Q;
  int offs, size;

  // gather: the blocks next_block() finds (hqp/Hqp_HL_BFGS.C:249-293), dense with both
  // triangles as update_b_Q receives them (symsp_extract_mat, :245), one after the other
  _offs.clear();
  _bs.clear();
  _Qp.clear();
  _sp.clear();
  _up.clear();
  _b_begin = 0;
  while (next_block(Q, &offs, &size)) {
    if (size != (int)_b_Q->n) _b_Q = m_resize(_b_Q, size, size);
    symsp_extract_mat(Q, offs, _b_Q);
    for (int i = 0; i < size; i++) _Qp.insert(_Qp.end(), _b_Q->me[i], _b_Q->me[i] + size);
    _sp.insert(_sp.end(), s->ve + offs, s->ve + offs + size);
    _up.insert(_up.end(), u->ve + offs, u->ve + offs + size);
    _offs.push_back(offs);
    _bs.push_back(size);
  }
  if (_bs.empty()) return;

  int info[3];
  const int rc = hqphl_bfgs_update(_device, (int)_bs.size(), &_bs[0], &_Qp[0], &_sp[0], &_up[0], alpha,
                                   _gamma, _eps, _eigen_control ? 1 : 0, info);
  if (rc != HQPHL_OK) {
    fprintf(stderr, "Hqp_HL_CudaBFGS: %s\n", hqphl_last_error());
    m_error(E_INTERN, "Hqp_HL_CudaBFGS::update");  // no CPU fallback
  }

  // scatter: upper triangles back into the SPMAT (symsp_insert_symmat, :247)
  const double *q = &_Qp[0];
  for (size_t b = 0; b < _bs.size(); b++) {
    const int n = _bs[b];
    if (n != (int)_b_Q->n) _b_Q = m_resize(_b_Q, n, n);
    for (int i = 0; i < n; i++) memcpy(_b_Q->me[i], q + (size_t)i * n, n * sizeof(double));
    symsp_insert_symmat(Q, _offs[b], _b_Q);
    q += (size_t)n * n;
  }
}
