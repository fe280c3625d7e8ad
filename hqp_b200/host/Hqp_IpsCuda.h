/*
 * Hqp_IpsCuda.h --
 *   - QP solver module for HQP: Mehrotra's predictor-corrector interior-point
 *     method with the whole iteration (KKT factor/solve AND the vector updates
 *     of Hqp_IpsMehrotra::step, hqp/Hqp_IpsMehrotra.C:355-693) resident on an
 *     NVIDIA B200; per solve only c, b, d go to the device and x, y, z come back
 *   - implements the solver interface Hqp_Solver (hqp/Hqp_Solver.h:40-96) and
 *     is selected like every other solver:  sqp_qp_solver CudaMehrotra
 *   - cold_start / hot_start / solve follow Hqp_IpsMehrotra (:209-352, 696-733);
 *     the numerics live behind hqpcu_mehrotra_solve / hqpcu_mehrotra_hot_solve
 *     (include/hqp_ipcuda.h)
 *
 * This file is new code; it only includes the reference's public headers.
 */
#ifndef Hqp_IpsCuda_H
#define Hqp_IpsCuda_H

#include <vector>

#include "Hqp_Solver.h"
#include "Hqp_IpCuda.h"

class Hqp_IpsCuda : public Hqp_Solver {
 protected:
  Hqp_IpCuda _mat;   ///< owns the device engine (structure detection, update)
  int _n, _me, _m;
  VEC *_w;           ///< slacks
  int _hot;          ///< next solve() is hot started
  int _max_warm_iters;
  int _logging;
  Real _gap;
  int _franke;       ///< 0: Mehrotra predictor-corrector, 1: Franke (Hqp_IpsFranke)
  Real _beta, _mu0;  ///< Franke: qp_beta, qp_mu0 (hqp/Hqp_IpsFranke.C:78-79)
  std::vector<double> _bp, _yp;  ///< b / y in the engine's equality row order
  std::vector<double> _cp, _xp;  ///< c / x in the engine's padded stage layout (non-uniform stages)

 public:
  Hqp_IpsCuda(int franke = 0);
  ~Hqp_IpsCuda();

  void init();
  void update();
  void cold_start();
  void hot_start();
  void step();
  void solve();

  Real gap() { return _gap; }
  const char *name() { return _franke ? "CudaFranke" : "CudaMehrotra"; }
};

/** Hqp_IpsFranke (hqp/Hqp_IpsFranke.C, the solver docp ships with) with the whole
 *  iteration on the device: sqp_qp_solver CudaFranke (hqpcu_franke_solve). */
class Hqp_IpsCudaFranke : public Hqp_IpsCuda {
 public:
  Hqp_IpsCudaFranke() : Hqp_IpsCuda(1) {}
};

#endif
