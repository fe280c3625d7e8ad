// docp_models.cuh -- device-resident stage models for libhqpdocp.so (SURVEY.md section 8,
// row f4).  A model is the device counterpart of an Hqp_Docp subclass' update_vals()
// (hqp/Hqp_Docp.h:75-76): ONE function template over the scalar type,
//
//   template <class T, class In, class Out>
//   static __device__ void vals(const ModelArgs &m, int k, const In &x, const In &u,
//                               Out &f, T &f0, Out &c);
//
// instantiated with T = double (values, forward differences) and T = Dual (forward-mode
// derivatives).  x[i] / u[i] yield a T; f[i] = ... / c[i] = ... accept a T and are WRITE-ONLY
// (the kernels hand in views: inputs come from a stage vector shared by all column threads
// with the thread's own column perturbed or seeded, outputs go straight to a shared-memory
// tile or to HBM -- no per-thread arrays, nothing in local memory).  Operation order is part of the contract: the CPU restatements
// (oracle/docp_oracle.py, oracle/prg_synthnl.cpp) evaluate the same expressions in the same
// order, the library is compiled with -fmad=false, so values agree to the last bit and
// differences quotients (which amplify rounding by 1e4..1e6) stay comparable.
#pragma once

struct ModelArgs {
  int K;               // stages with controls of the WHOLE horizon; vals() gets the global stage k
  int nx, nu, nc, ncK;
  const double *par;   // [npar]   global parameters (device)
  const double *spar;  // [(K+1)][nspar] per-stage parameters (device) or nullptr
  int nspar;
  int k0;              // global index of the handle's first stage (spar row 0)
};

// forward-mode dual number: value and ONE directional derivative
struct Dual {
  double v, d;
  Dual() = default;
  __device__ Dual(double v_) : v(v_), d(0.0) {}
  __device__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ inline Dual operator/(Dual a, Dual b) {
  const double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ inline Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
__device__ inline Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
__device__ inline Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
__device__ inline Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
__device__ inline Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
__device__ inline Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
__device__ inline Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
__device__ inline Dual operator/(double a, Dual b) {
  const double q = a / b.v;
  return Dual(q, (-q * b.d) / b.v);
}

// ---- hqp_docp/Prg_DID.C:78-98: double integrator.  nx = 2, nu = 1, nc = 0 | 1
// (prg_with_cns), ncK = 0.  par = { dt } with dt = 1 / kmax.
struct ModelDID {
  static bool dims_ok(int nx, int nu, int nc, int ncK, int npar, int nspar) {
    return nx == 2 && nu == 1 && (nc == 0 || nc == 1) && ncK == 0 && npar == 1 && nspar == 0;
  }
  template <class T, class In, class Out>
  static __device__ void vals(const ModelArgs &m, int k, const In &x, const In &u, Out &f, T &f0, Out &c) {
    const double dt = m.par[0];
    if (k < m.K) {
      f[0] = x[0] + u[0] * dt;
      f[1] = x[0] * dt + x[1] + u[0] * 0.5 * dt * dt;
      f0 = u[0] * u[0] * dt;
      if (m.nc) c[0] = x[1] + 0.5 * dt * x[0];
    } else {
      f0 = T(0.0);
    }
  }
};

// ---- the synthetic SQP-driven workload ("Hqp_Docp subclass
// with mildly nonlinear dynamics"):
//   f_i  = sum_j A_ij x_j + sum_j B_ij u_j + eps x_i / (1 + x_i^2)                 k < K
//   f0   = 1/2 [ sum_i qw_i (x_i - r_ki)^2 + sum_j rw_j u_j^2 ] + eps sum_{j < min(nx,nu)} x_j u_j
//          (k = K: the state term only)
//   c_0  = (sum_i x_i^2) / nx + eps u_0 x_0,   c_i = x_i u_{i mod nu}  (0 < i < nc)   k < K
//   c_i  = x_i^2  (i < ncK)                                                          k = K
// par = { eps, A[nx*nx], B[nx*nu], qw[nx], rw[nu] } row-major; spar[k] = r_k[nx].
struct ModelSynthNL {
  static bool dims_ok(int nx, int nu, int nc, int ncK, int npar, int nspar) {
    return nx >= 1 && nu >= 1 && nc <= nx && ncK <= nx &&
           npar == 1 + nx * nx + nx * nu + nx + nu && nspar == nx;
  }
  template <class T, class In, class Out>
  static __device__ void vals(const ModelArgs &m, int k, const In &x, const In &u, Out &f, T &f0, Out &c) {
    const int nx = m.nx, nu = m.nu;
    const double eps = m.par[0];
    const double *A = m.par + 1, *B = A + nx * nx, *qw = B + nx * nu, *rw = qw + nx;
    const double *r = m.spar + (size_t)(k - m.k0) * m.nspar;
    T s(0.0);
    for (int i = 0; i < nx; i++) {
      T e = x[i] - r[i];
      s = s + qw[i] * e * e;
    }
    if (k < m.K) {
      for (int i = 0; i < nx; i++) {
        T acc(0.0);
        for (int j = 0; j < nx; j++) acc = acc + A[i * nx + j] * x[j];
        for (int j = 0; j < nu; j++) acc = acc + B[i * nu + j] * u[j];
        f[i] = acc + eps * (x[i] / (1.0 + x[i] * x[i]));
      }
      for (int j = 0; j < nu; j++) s = s + rw[j] * u[j] * u[j];
      s = 0.5 * s;
      const int nm = nx < nu ? nx : nu;
      for (int j = 0; j < nm; j++) s = s + eps * x[j] * u[j];
      f0 = s;
      if (m.nc > 0) {
        T q(0.0);
        for (int i = 0; i < nx; i++) q = q + x[i] * x[i];
        c[0] = q / (double)nx + eps * u[0] * x[0];
        for (int i = 1; i < m.nc; i++) c[i] = x[i] * u[i % nu];
      }
    } else {
      f0 = 0.5 * s;
      for (int i = 0; i < m.ncK; i++) c[i] = x[i] * x[i];
    }
  }
};
