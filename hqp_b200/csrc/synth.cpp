// Seeded synthetic LQ-DOCP generator (host C++, no CUDA).
//
// Produces the stage slabs the C ABI consumes (include/hqp_ipcuda.h) for the
// workloads named in BASELINE.json.  The oracle harness (oracle/ref_harness.cpp)
// builds the reference's Hqp_Program from the SAME slabs, so both arms of every
// parity test and of bench.py see bit-identical inputs.
//
// Draw order follows SURVEY.md App. B.7 (libstdc++ std::mt19937_64 +
// uniform_real_distribution<double>(-1,1)), which is the layout Hqp_Docp
// produces (hqp/Hqp_Docp.C:585-755): variables [x0,u0,x1,u1,...,xK]; equality
// rows = K*nx dynamics rows [fx fu -I] followed by nx rows fixing x0;
// inequality rows 2*(k*nu+j), 2*(k*nu+j)+1 = +u_j + 1 >= 0, -u_j + 1 >= 0.

#include <cmath>
#include <cstddef>
#include <random>
#include <vector>

extern "C" {

// Q   : (K+1) blocks of (nx+nu)^2, row-major, full symmetric; block K only
//       uses its leading nx x nx part (rest zero)
// c   : K*(nx+nu)+nx
// fx  : K blocks nx*nx row-major; fu: K blocks nx*nu row-major
// b   : K*nx + nx   (dynamics offsets, then -x0)
// returns 0
int hqp_synth_lqdocp(int nx, int nu, int K, unsigned long long seed,
                     double *Q, double *c, double *fx, double *fu, double *b) {
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const int nm = nx + nu;
  std::vector<double> M((size_t)nm * nm);
  const double rs = 1.0 / std::sqrt((double)nx);
  for (int k = 0; k <= K; k++) {
    const int d = (k < K) ? nm : nx;
    for (int i = 0; i < nm * nm; i++) M[i] = U(g);
    double *Qk = Q + (size_t)k * nm * nm;
    for (int i = 0; i < nm * nm; i++) Qk[i] = 0.0;
    for (int i = 0; i < d; i++)
      for (int j = i; j < d; j++) {
        double s = 0.0;
        for (int l = 0; l < nm; l++) s += M[l * nm + i] * M[l * nm + j] / nm;
        if (i == j) s += 0.1;
        Qk[i * nm + j] = s;
        Qk[j * nm + i] = s;
      }
    double *ck = c + (size_t)k * nm;
    for (int i = 0; i < d; i++) ck[i] = U(g);
    if (k < K) {
      double *fxk = fx + (size_t)k * nx * nx;
      double *fuk = fu + (size_t)k * nx * nu;
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < nx; j++)
          fxk[i * nx + j] = (i == j ? 1.0 : 0.0) + 0.1 * U(g) * rs;
        for (int j = 0; j < nu; j++) fuk[i * nu + j] = U(g);
        b[(size_t)k * nx + i] = 0.01 * U(g);
      }
    }
  }
  for (int i = 0; i < nx; i++) b[(size_t)K * nx + i] = -U(g);
  return 0;
}

// Direct-plugin RHS of SURVEY.md App. B.7 ("phase.cpp"): z,w in [0.5,1.5]
// interleaved per row with r3,r4, then r1, then r2.
int hqp_synth_rhs(int N, int me, int m, unsigned long long seed, double *z,
                  double *w, double *r1, double *r2, double *r3, double *r4) {
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  for (int i = 0; i < m; i++) {
    z[i] = 0.5 + 0.5 * (U(g) + 1.0);
    w[i] = 0.5 + 0.5 * (U(g) + 1.0);
    r3[i] = U(g);
    r4[i] = U(g);
  }
  for (int i = 0; i < N; i++) r1[i] = U(g);
  for (int i = 0; i < me; i++) r2[i] = U(g);
  return 0;
}

}  // extern "C"
