/*
 * Hqp_IpCuda.h --
 *   - matrix module for HQP's interior-point QP solvers that factors and
 *     solves the stage-structured KKT system on an NVIDIA B200 through the
 *     C ABI of libhqpcuda.so (include/hqp_ipcuda.h)
 *   - implements the plugin interface Hqp_IpMatrix (hqp/Hqp_IpMatrix.h:42-89)
 *     and is selected like every other module:  qp_mat_solver Cuda
 *   - same layering as Hqp_IpPARDISO + pardiso_wrapper
 *     (hqp/Hqp_IpPARDISO.C:161-166, hqp/pardiso_wrapper.h:32-48): a thin C++
 *     class that owns no numerics
 *
 * This file is new code; it only includes the reference's public headers.
 */
#ifndef Hqp_IpCuda_H
#define Hqp_IpCuda_H

#include <vector>

#include "Hqp_IpMatrix.h"

struct hqpcu_handle;

class Hqp_IpCuda : public Hqp_IpMatrix {
 protected:
  hqpcu_handle *_h;
  // interface options
  int _nseg;      ///< mat_nseg: horizon segments (0 = automatic, 1 = sequential)
  int _device;    ///< mat_device: CUDA device ordinal (the first one with mat_ngpu > 1)
  int _ngpu;      ///< mat_ngpu: split the horizon over this many GPUs (hqpcu_dims::ngpu)
  int _dev_solve; ///< mat_dev_solve: run the refinement loop of solve() on the device
  int _sparse_update; ///< mat_sparse_update: upload SPMAT values + device-side scatter (row f1)

  // stage structure derived from the sparsity of A and C (cf. Hqp_IpLQDOCP
  // Get_Dim / Get_Constr_Dim / Check_Structure)
  int _K, _nx, _nu, _n, _me, _m;
  int _fixed_x0;
  int _n_eq;
  std::vector<int> _rowmap;   ///< ABI equality row -> row of qp->A
  std::vector<int> _ineq_ptr; ///< CSR pattern of qp->C as passed to the ABI
  std::vector<int> _eq_ptr;
  // packed host slabs (reused between updates)
  std::vector<double> _Q, _fx, _fu, _cval, _eval;
  std::vector<double> _vals;  ///< values of Q / dynamics rows of A in map order (sparse update)
  void build_value_map(const Hqp_Program *qp);
  std::vector<double> _r2p, _dyp; ///< permuted copies when _rowmap is not the identity
  bool _identity_rows;
  // non-uniform stage dimensions (Hqp_IpLQDOCP::Get_Dim handles _nk[k], _mk[k] per
  // stage, hqp/Hqp_IpLQDOCP.C:201-287): every stage is padded to the largest nx / nu
  // with decoupled unit blocks (identity Hessian, zero dynamics), which leaves the
  // solution of the original variables unchanged; _xmap: variable -> padded slot
  bool _padded;
  int _n_abi, _me_abi;            ///< vector lengths on the device side of the ABI
  std::vector<int> _xmap;
  std::vector<double> _r1p, _dxp; ///< padded copies of x-like vectors
  std::vector<long long> _pad_diag; ///< slab positions of the padded diagonal entries of Q
  std::vector<int> _xoff, _nxs, _dyn0, _vstage; ///< stage geometry of the QP's own layout

  void free_handle();
  void check(int status, const char *where);
  const double *pack_r2(const VEC *r2);
  void unpack_dy(VEC *dy);
  const double *pack_r1(const VEC *r1);
  void unpack_dx(VEC *dx);

 public:
  Hqp_IpCuda();
  ~Hqp_IpCuda();

  void init(const Hqp_Program *);
  void update(const Hqp_Program *);
  void factor(const Hqp_Program *, const VEC *z, const VEC *w);
  void step(const Hqp_Program *, const VEC *z, const VEC *w,
            const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
            VEC *dx, VEC *dy, VEC *dz, VEC *dw);
  Real solve(const Hqp_Program *, const VEC *z, const VEC *w,
             const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
             VEC *dx, VEC *dy, VEC *dz, VEC *dw);
  Real residuum(const Hqp_Program *, const VEC *z, const VEC *w,
                const VEC *r1, const VEC *r2, const VEC *r3, const VEC *r4,
                VEC *dx, VEC *dy, VEC *dz, VEC *dw);

  const char *name() { return "Cuda"; }

  // used by Hqp_IpsCuda (the device-resident IP solver built on this engine)
  hqpcu_handle *handle() { return _h; }
  bool identity_rows() const { return _identity_rows; }
  const std::vector<int> &rowmap() const { return _rowmap; }
  // vectors of the QP <-> vectors of the (possibly padded / permuted) device layout
  bool padded() const { return _padded; }
  int n_abi() const { return _n_abi; }
  int me_abi() const { return _me_abi; }
  void to_abi_x(const double *user, double *abi) const;
  void from_abi_x(const double *abi, double *user) const;
  void to_abi_y(const double *user, double *abi) const;
  void from_abi_y(const double *abi, double *user) const;
};

#endif
