/*
 * Hqp_HL_CudaBFGS.h --
 *   - Hessian-of-the-Lagrangian module for HQP's SQP solvers that performs the
 *     block-diagonal BFGS update (Powell damping, eigenvalue control) of ALL diagonal
 *     blocks of Q in one call on an NVIDIA B200 (include/hqp_hlcuda.h, libhqphl.so)
 *   - derives from the reference's Hqp_HL_BFGS (hqp/Hqp_HL_BFGS.h:34-60): setup, init,
 *     posdef and the options sqp_hela_gamma / sqp_hela_eigen_control / sqp_hela_eps /
 *     sqp_hela_bsize are the reference's own; only update() is replaced
 *   - selected like every other module:  sqp_hela CudaBFGS
 *
 * This file is new code; it only includes the reference's public headers.
 */
#ifndef Hqp_HL_CudaBFGS_H
#define Hqp_HL_CudaBFGS_H

#include <vector>

#include "Hqp_HL_BFGS.h"

class Hqp_HL_CudaBFGS : public Hqp_HL_BFGS {
 protected:
  int _device;  ///< sqp_hela_device: CUDA device ordinal
  std::vector<int> _offs, _bs;
  std::vector<double> _Qp, _sp, _up;

 public:
  Hqp_HL_CudaBFGS();

  /** replaces the per-block loop of Hqp_HL_BFGS::update (hqp/Hqp_HL_BFGS.C:216-243) */
  void update(const VEC *s, const VEC *u, Real alpha, Hqp_SqpProgram *);

  const char *name() { return "CudaBFGS"; }
};

#endif
