// docp_update.cu -- libhqpdocp.so: the stage loop of Hqp_Docp::update / ::update_fbd
// (hqp/Hqp_Docp.C:831-891, 944-1075), the default difference quotients of
// Hqp_Docp::update_grds (:1097-1180) and Hqp_Docp::update_bounds (:893-940) for
// device-resident stage models (docp_models.cuh).  C ABI: include/hqp_docpcuda.h.
// SURVEY.md section 8, row f4.  sm_100a; compiled with -fmad=false (see docp_models.cuh).
//
// Kernels (all streaming; nothing is re-read from HBM except the iterate x, which the
// nx+nu column threads of a stage share through L1/L2):
//   docp_vals_kernel   (update_fbd) one thread per stage: f_k, f0_k, c_k; writes b's dynamics
//                      rows (f_k - x_{k+1}), the stage objective and the constraint values
//   docp_stage_kernel  (update) one thread per (stage, variable) plus one for the plain values:
//                      one perturbed evaluation (FD) or one dual-number evaluation (AD) each,
//                      results into a shared-memory tile per stage, then coalesced row stores of
//                      fx fu cx cu g b -- inputs and outputs through views, no local memory
//   docp_assoc_kernel  the association tables -> b, d (update_bounds + the c_k rows)
//   docp_sum_kernel    f = sum_k f0_k, fixed order (one CTA)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "docp_models.cuh"
#include "hqp_docpcuda.h"

namespace {

constexpr int MAXX = 64, MAXU = 32, MAXC = 8;
thread_local std::string g_err;

int fail(const std::string &m, int code) {
  g_err = m;
  return code;
}

#define CU(call)                                                                           \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_), HQPDOCP_E_CUDA);     \
  } while (0)

struct Assoc {
  int dim = 0;
  int *idxs = nullptr;
  double *vals = nullptr;
};

struct UpdArgs {
  ModelArgs m;
  const double *x;  // [N]
  double *fbase;    // [K][nx]   f_k before x_{k+1} is subtracted (scratch)
  double *f0k;      // [K+1]     stage objectives (scratch)
  double *cval;     // [K nc + ncK] constraint values (scratch)
  double *b;        // [me]
  double *g, *fx, *fu, *cx, *cu;
  int npar;
  int KL;    // stages with controls of THIS handle (local indices 0 .. KL)
  int halo;  // 1: local stage KL belongs to the next stage range -- only its x is read
};

// ---- views handed to Model::vals (docp_models.cuh)
// input, plain: element i at p[i * stride] -- a stage vector in shared memory (stride 1), a
// column of the transposed stage block (values kernel) or a thread's own perturbed copy
struct ColView {
  const double *p;
  int stride;
  __device__ double operator[](int i) const { return p[i * stride]; }
};
// input, perturbed in place: the stage vector shared by all column threads, element `col`
// replaced by the thread's own perturbed value (when the per-column copies do not fit)
struct PertView {
  const double *base;
  int col;
  const double *slot;  // the thread's perturbed value, in shared memory like base
  __device__ double operator[](int i) const { return *(i == col ? slot : base + i); }
};
// input, dual numbers: the shared stage vector, the thread's own column carries the seed
struct SeedView {
  const double *base;
  int col;
  __device__ Dual operator[](int i) const { return Dual(base[i], i == col ? 1.0 : 0.0); }
};
// output: element i goes to p[i * stride]; a Dual leaves its derivative
struct OutRef {
  double *p;
  __device__ void operator=(double v) { *p = v; }
  __device__ void operator=(Dual v) { *p = v.d; }
};
struct OutView {
  double *p;
  int stride;
  __device__ OutRef operator[](int i) const { return OutRef{p + i * stride}; }
};

constexpr int VT = 128;        // stages per CTA of the values kernel
constexpr int VLD = VT + 1;    // padded leading dimension of its transposed tiles

// values only (Hqp_Docp::update_fbd): a CTA takes VT consecutive stages, one thread each.  The
// stages' x_k u_k arrive with coalesced loads and are transposed into shared memory (element i
// of stage s at [i][s]: conflict-free for the evaluation), results leave the same way.
// Shared memory: [ par copy | xT nd x VLD | outT (nx + ncm) x VLD ].
template <class Model, bool PARSH>
__global__ void __launch_bounds__(VT) docp_vals_kernel(UpdArgs a, int ncm) {
  extern __shared__ double sh[];
  ModelArgs m = a.m;
  const int nx = m.nx, nd = m.nx + m.nu;
  const int npar_sh = PARSH ? a.npar : 0;
  double *xT = sh + npar_sh, *outT = xT + nd * VLD;
  if (PARSH) {
    for (int i = threadIdx.x; i < npar_sh; i += VT) sh[i] = m.par[i];
    m.par = sh;
  }
  const long long k0 = (long long)blockIdx.x * VT;
  const int ns = (int)(k0 + VT <= a.KL + 1 ? VT : a.KL + 1 - k0);
  const long long lo = k0 * nd, n_all = (long long)a.KL * nd + nx;
  for (int t = threadIdx.x; t < ns * nd; t += VT) {
    const int s = t / nd, i = t - s * nd;
    xT[i * VLD + s] = lo + t < n_all ? a.x[lo + t] : 0.0;
  }
  for (int i = 0; i < nx + ncm; i++) outT[i * VLD + threadIdx.x] = 0.0;  // v_zero(fk), hqp/Hqp_Docp.C:856
  __syncthreads();
  if ((int)threadIdx.x < ns) {
    const long long k = k0 + threadIdx.x;
    ColView x{xT + threadIdx.x, VLD}, u{xT + nx * VLD + threadIdx.x, VLD};
    OutView f{outT + threadIdx.x, VLD}, c{outT + nx * VLD + threadIdx.x, VLD};
    double f0 = 0.0;
    if (!(k == a.KL && a.halo)) Model::template vals<double>(m, (int)k + m.k0, x, u, f, f0, c);
    a.f0k[k] = f0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ns * nx; t += VT) {
    const int s = t / nx, i = t - s * nx;
    const long long k = k0 + s;
    if (k < a.KL) {
      const double v = outT[i * VLD + s];
      a.fbase[k * nx + i] = v;
      a.b[k * nx + i] = v - a.x[(k + 1) * nd + i];  // v_sub(fk, x_{k+1}), :861
    }
  }
  for (int t = threadIdx.x; t < ns * ncm; t += VT) {
    const int s = t / ncm, i = t - s * ncm;
    const long long k = k0 + s;
    if (i < (k < a.KL ? m.nc : (a.halo ? 0 : m.ncK))) a.cval[k * m.nc + i] = outT[(nx + i) * VLD + s];
  }
}

// values + derivatives (Hqp_Docp::update): a CTA takes S consecutive stages; thread (s, j) of a
// stage evaluates the model once -- j < nd: column j (perturbed: Hqp_Docp::update_grds,
// hqp/Hqp_Docp.C:1127-1171; or seeded: dual numbers), j = nd: the plain values -- and leaves its
// nx + nc + 1 results in column j of the stage's shared-memory tile.  After the barrier the CTA
// forms the quotients (FD) and streams rows of fx, fu, cx, cu, g, b out with coalesced stores.
// XCOPY (FD): every column thread reads its own perturbed copy of the stage vector (element i of
// column j at [i][j]: one conflict-free LDS per access instead of compare + select + LDS).
// Shared memory: [ par copy | per stage: x_k u_k (nd) | tile (nx + ncm + 1) x (nd + 1) |
//                  XCOPY: copies nd x (nd + 1) ] | one perturbed value per thread ].
template <class Model, int MODE, bool PARSH, bool XCOPY>
__global__ void __launch_bounds__(256) docp_stage_kernel(UpdArgs a, int S, int ncm) {
  extern __shared__ double sh[];
  ModelArgs m = a.m;
  const int nx = m.nx, nd = m.nx + m.nu, ld = nd + 1, rows = nx + ncm + 1;
  const int npar_sh = PARSH ? a.npar : 0;
  double *stage_sh = sh + npar_sh;
  const int per_stage = nd + rows * ld + (XCOPY ? nd * ld : 0);
  if (PARSH) {
    for (int i = threadIdx.x; i < npar_sh; i += blockDim.x) sh[i] = m.par[i];
    m.par = sh;
  }
  const long long k0 = (long long)blockIdx.x * S;
  const int ns = (int)(k0 + S <= a.KL + 1 ? S : a.KL + 1 - k0);
  // stage vectors: ns * nd contiguous doubles of x (the final stage has nx only)
  {
    const long long lo = k0 * nd, n_all = (long long)a.KL * nd + nx;
    for (int t = threadIdx.x; t < ns * nd; t += blockDim.x) {
      const int s = t / nd, i = t - s * nd;
      stage_sh[s * per_stage + i] = lo + t < n_all ? a.x[lo + t] : 0.0;
    }
  }
  __syncthreads();
  if (XCOPY) {
    for (int t = threadIdx.x; t < ns * nd * nd; t += blockDim.x) {
      const int s = t / (nd * nd), r = t - s * nd * nd, i = r / nd, jj = r - i * nd;
      const double *xs = stage_sh + s * per_stage;
      const double v = xs[i];
      // x->ve[j] += dvj with dvj = 1e-4 |v| + 1e-6, hqp/Hqp_Docp.C:1129-1131
      stage_sh[s * per_stage + nd + rows * ld + i * ld + jj] = i == jj ? v + (1e-4 * fabs(v) + 1e-6) : v;
    }
    __syncthreads();
  }
  const int s = threadIdx.x / ld, j = threadIdx.x - s * ld;
  if (s < ns) {
    const long long k = k0 + s;
    const bool last = k == a.KL;
    const int ncols = last ? nx : nd;
    if (!(last && a.halo) && (j < ncols || j == nd)) {
      const double *xs = stage_sh + s * per_stage;
      double *tile = stage_sh + s * per_stage + nd;
      const int nf = last ? 0 : nx, nc = last ? m.ncK : m.nc;
      OutView f{tile + j, ld}, c{tile + nx * ld + j, ld};
      for (int i = 0; i < nf; i++) tile[i * ld + j] = 0.0;  // v_zero(df) / v_zero(dc), :1136-1138
      for (int i = 0; i < nc; i++) tile[(nx + i) * ld + j] = 0.0;
      double *f0_out = tile + (nx + ncm) * ld + j;
      if (j == nd || MODE == HQPDOCP_GRAD_FD) {
        double f0 = 0.0;
        if (XCOPY) {
          const double *xc = tile + rows * ld + j;
          ColView x{j == nd ? xs : xc, j == nd ? 1 : ld}, u{j == nd ? xs + nx : xc + nx * ld, j == nd ? 1 : ld};
          Model::template vals<double>(m, (int)k + m.k0, x, u, f, f0, c);
        } else {
          int col = -1;
          double *slot = sh + npar_sh + S * per_stage + threadIdx.x;
          if (j < nd) {
            const double v = xs[j];
            *slot = v + (1e-4 * fabs(v) + 1e-6);  // x->ve[j] += dvj, :1129-1131
            col = j;
          }
          PertView x{xs, col, slot}, u{xs + nx, col - nx, slot};
          Model::template vals<double>(m, (int)k + m.k0, x, u, f, f0, c);
        }
        *f0_out = f0;
      } else {
        SeedView x{xs, j}, u{xs + nx, j - nx};
        Dual f0(0.0);
        Model::template vals<Dual>(m, (int)k + m.k0, x, u, f, f0, c);
        *f0_out = f0.d;
      }
    }
  }
  __syncthreads();
  // write-out: rows of the tiles, consecutive threads on consecutive columns
  for (int s2 = 0; s2 < ns; s2++) {
    const long long k = k0 + s2;
    const bool last = k == a.KL;
    if (last && a.halo) {  // the next range's stage: nothing of it is ours
      if (threadIdx.x == 0) a.f0k[k] = 0.0;
      continue;
    }
    const double *xs = stage_sh + s2 * per_stage;
    const double *tile = xs + nd;
    const int nf = last ? 0 : nx, nc = last ? m.ncK : m.nc, ncols = last ? nx : nd;
    const int nrow = nf + nc + 1;  // f rows, c rows, the objective row
    for (int t = threadIdx.x; t < nrow * ld; t += blockDim.x) {
      int r = t / ld;
      const int jj = t - r * ld;
      if (jj >= ncols && jj != nd) continue;
      // tile row: f rows 0..nf-1, c rows nx.., objective row nx + ncm
      const int trow = r < nf ? r : (r < nf + nc ? nx + (r - nf) : nx + ncm);
      const double v = tile[trow * ld + jj];
      if (jj == nd) {  // the plain values
        if (r < nf) {
          a.fbase[k * nx + r] = v;
          a.b[k * nx + r] = v - a.x[(k + 1) * nd + r];  // v_sub(fk, x_{k+1}), :1018
        } else if (r < nf + nc) {
          a.cval[k * m.nc + (r - nf)] = v;
        } else {
          a.f0k[k] = v;
        }
        continue;
      }
      double q = v;
      if (MODE == HQPDOCP_GRAD_FD) {
        const double dvj = 1e-4 * fabs(xs[jj]) + 1e-6;
        q = (v - tile[trow * ld + nd]) / dvj;  // (df - f) / dvj, :1140-1148
      }
      if (r < nf) {
        if (jj < nx) a.fx[(k * nx + r) * nx + jj] = q;
        else a.fu[(k * nx + r) * m.nu + (jj - nx)] = q;
      } else if (r < nf + nc) {
        const long long ci = k * m.nc + (r - nf);
        if (jj < nx) a.cx[ci * nx + jj] = q;
        else a.cu[ci * m.nu + (jj - nx)] = q;
      } else {
        a.g[k * nd + jj] = q;  // f0x / f0u written in place into qp->c (:965-966)
      }
    }
  }
}

struct AssocArgs {
  const double *x, *cval;
  double *b, *d;
  long long b_off;  // K nx
  Assoc xu_eq, xu_lb, xu_ub, cns_eq, cns_lb, cns_ub;
};

__global__ void __launch_bounds__(256) docp_assoc_kernel(AssocArgs a, long long total) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long i = t;
    // b: [dynamics | xu_eq | cns_eq]   (hqp/Hqp_Docp.C:573-576)
    if (i < a.xu_eq.dim) { a.b[a.b_off + i] = a.x[a.xu_eq.idxs[i]] - a.xu_eq.vals[i]; continue; }  // :905-911
    i -= a.xu_eq.dim;
    if (i < a.cns_eq.dim) {  // :867-871
      a.b[a.b_off + a.xu_eq.dim + i] = a.cval[a.cns_eq.idxs[i]] - a.cns_eq.vals[i];
      continue;
    }
    i -= a.cns_eq.dim;
    // d: [xu_lb | xu_ub | cns_lb | cns_ub]   (:578-582)
    double *d = a.d;
    if (i < a.xu_lb.dim) { d[i] = a.x[a.xu_lb.idxs[i]] - a.xu_lb.vals[i]; continue; }  // :913-919
    i -= a.xu_lb.dim;
    d += a.xu_lb.dim;
    if (i < a.xu_ub.dim) { d[i] = -a.x[a.xu_ub.idxs[i]] + a.xu_ub.vals[i]; continue; }  // :921-927
    i -= a.xu_ub.dim;
    d += a.xu_ub.dim;
    if (i < a.cns_lb.dim) { d[i] = a.cval[a.cns_lb.idxs[i]] - a.cns_lb.vals[i]; continue; }  // :873-876
    i -= a.cns_lb.dim;
    d += a.cns_lb.dim;
    d[i] = a.cns_ub.vals[i] - a.cval[a.cns_ub.idxs[i]];  // :878-881
  }
}

// f = sum_k f0_k in a fixed order: thread t adds its contiguous slice in stage order, the
// 1024 partial sums are combined by a fixed binary tree (the reference adds stage by stage,
// hqp/Hqp_Docp.C:860, 885 -- equal up to the rounding of a re-associated sum)
__global__ void __launch_bounds__(1024) docp_sum_kernel(const double *f0k, long long n, double *f) {
  __shared__ double part[1024];
  const long long per = (n + 1023) / 1024;
  const long long lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  double s = 0.0;
  for (long long i = lo; i < hi; i++) s += f0k[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int w = 1; w < 1024; w <<= 1) {
    if ((threadIdx.x & (2 * w - 1)) == 0) part[threadIdx.x] += part[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *f = part[0];
}

}  // namespace

struct hqpdocp_handle {
  hqpdocp_dims dims{};
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr;
  long long N = 0, me = 0, m = 0, ncns = 0;
  long long launches = 0;
  bool staged = false;
  int xcopy_mode = -1;  // HQPDOCP_XCOPY: -1 choose, 0 never, 1 always (when it fits)
  double *d_par = nullptr, *d_spar = nullptr;
  Assoc t[6];  // xu_eq xu_lb xu_ub cns_eq cns_lb cns_ub
  double *fbase = nullptr, *f0k = nullptr, *cval = nullptr;
  // staging for the host-pointer entry points
  double *s_x = nullptr, *s_f = nullptr, *s_b = nullptr, *s_d = nullptr, *s_g = nullptr, *s_fx = nullptr,
         *s_fu = nullptr, *s_cx = nullptr, *s_cu = nullptr;
};

namespace {

int upload_assoc(const hqpdocp_assoc &src, Assoc &dst, long long idx_limit, const char *name) {
  dst.dim = src.dim;
  if (src.dim < 0 || (src.dim > 0 && (!src.idxs || !src.vals)))
    return fail(std::string("hqpdocp_create: bad association table ") + name, HQPDOCP_E_ARG);
  if (src.dim == 0) return HQPDOCP_OK;
  for (int i = 0; i < src.dim; i++)
    if (src.idxs[i] < 0 || src.idxs[i] >= idx_limit)
      return fail(std::string("hqpdocp_create: index out of range in ") + name, HQPDOCP_E_ARG);
  CU(cudaMalloc(&dst.idxs, sizeof(int) * src.dim));
  CU(cudaMalloc(&dst.vals, sizeof(double) * src.dim));
  CU(cudaMemcpy(dst.idxs, src.idxs, sizeof(int) * src.dim, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dst.vals, src.vals, sizeof(double) * src.dim, cudaMemcpyHostToDevice));
  return HQPDOCP_OK;
}

// the S (stages per CTA, <= 256 threads) that keeps the most evaluation threads resident on an
// SM (227 KB of shared memory, 2048 threads, 32 CTAs); 0 if not even one stage fits
int pick_stages(int ld, size_t fixed, size_t per_stage, long long *resident) {
  int S = 0;
  long long best = -1;
  for (int c = 1; c <= std::max(1, 256 / ld); c++) {
    const size_t sm = (fixed + c * per_stage + 256) * 8 + 1024;
    if (sm > 200 * 1024) break;
    const int thr_c = ((c * ld + 31) / 32) * 32;
    const long long ctas = std::min<long long>({(long long)(227 * 1024 / sm), 2048 / thr_c, 32});
    if (ctas * c * ld > best) {
      best = ctas * c * ld;
      S = c;
    }
  }
  if (resident) *resident = best;
  return S;
}

template <class Model, bool PARSH>
int launch_vals(hqpdocp_handle *h, const UpdArgs &a, int ncm, size_t smem) {
  static thread_local int optin_dev = -1;  // (the attribute is per function and device)
  if (optin_dev != h->device) {
    CU(cudaFuncSetAttribute(docp_vals_kernel<Model, PARSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    optin_dev = h->device;
  }
  const long long grid = (h->dims.K + 1 + VT - 1) / VT;
  docp_vals_kernel<Model, PARSH><<<(unsigned)grid, VT, smem, h->stream>>>(a, ncm);
  return HQPDOCP_OK;
}

template <class Model, int MODE, bool PARSH, bool XCOPY>
int launch_stage(hqpdocp_handle *h, const UpdArgs &a, int S, int ncm, size_t smem) {
  const int ld = h->dims.nx + h->dims.nu + 1;
  static thread_local int optin_dev = -1;
  if (optin_dev != h->device) {
    CU(cudaFuncSetAttribute(docp_stage_kernel<Model, MODE, PARSH, XCOPY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            200 * 1024));
    optin_dev = h->device;
  }
  const long long grid = ((long long)h->dims.K + 1 + S - 1) / S;
  docp_stage_kernel<Model, MODE, PARSH, XCOPY><<<(unsigned)grid, ((S * ld + 31) / 32) * 32, smem, h->stream>>>(a, S, ncm);
  return HQPDOCP_OK;
}

template <class Model>
int launch_model(hqpdocp_handle *h, bool grads, int mode, const UpdArgs &a) {
  const hqpdocp_dims &D = h->dims;
  const int nd = D.nx + D.nu, ld = nd + 1;
  const int ncm = std::max(D.nc, D.ncK);
  if (((long long)D.K + 1) > 0x7fffffffLL) return fail("hqpdocp: horizon too long for one launch", HQPDOCP_E_UNSUPPORTED);
  int rc;
  if (!grads) {
    const size_t tiles = (size_t)(nd + D.nx + ncm) * VLD;
    const bool parsh = (D.npar + tiles) * 8 <= 200 * 1024;
    const size_t smem = ((parsh ? D.npar : 0) + tiles) * 8;
    if (smem > 200 * 1024) return fail("hqpdocp: stage too large for the values kernel", HQPDOCP_E_UNSUPPORTED);
    rc = parsh ? launch_vals<Model, true>(h, a, ncm, smem) : launch_vals<Model, false>(h, a, ncm, smem);
  } else {
    const size_t tile = (size_t)nd + (size_t)(D.nx + ncm + 1) * ld, copies = (size_t)nd * ld;
    const bool fd = mode == HQPDOCP_GRAD_FD;
    // parameter copy in shared memory unless it crowds out the stages; per-column copies of the
    // stage vector (FD) only when they cost no resident threads (measured at C2 / C5: with half
    // the threads resident the copies lose, 0.229 vs 0.157 ms and 10.8 vs 5.5 ms)
    bool parsh = (size_t)D.npar * 8 <= 64 * 1024;
    long long r_plain = 0, r_copy = 0;
    int S = pick_stages(ld, parsh ? D.npar : 0, tile, &r_plain);
    if (S == 0 && parsh) {
      parsh = false;
      S = pick_stages(ld, 0, tile, &r_plain);
    }
    if (S == 0) return fail("hqpdocp: stage too large for the update kernel", HQPDOCP_E_UNSUPPORTED);
    bool xcopy = false;
    if (fd && h->xcopy_mode != 0) {
      const int Sc = pick_stages(ld, parsh ? D.npar : 0, tile + copies, &r_copy);
      if (Sc > 0 && (h->xcopy_mode == 1 || r_copy >= r_plain)) {
        xcopy = true;
        S = Sc;
      }
    }
    const size_t smem = ((parsh ? D.npar : 0) + S * (tile + (xcopy ? copies : 0)) + 256) * sizeof(double);
    if (fd) {
      if (xcopy) rc = parsh ? launch_stage<Model, HQPDOCP_GRAD_FD, true, true>(h, a, S, ncm, smem)
                            : launch_stage<Model, HQPDOCP_GRAD_FD, false, true>(h, a, S, ncm, smem);
      else rc = parsh ? launch_stage<Model, HQPDOCP_GRAD_FD, true, false>(h, a, S, ncm, smem)
                      : launch_stage<Model, HQPDOCP_GRAD_FD, false, false>(h, a, S, ncm, smem);
    } else {
      rc = parsh ? launch_stage<Model, HQPDOCP_GRAD_AD, true, false>(h, a, S, ncm, smem)
                 : launch_stage<Model, HQPDOCP_GRAD_AD, false, false>(h, a, S, ncm, smem);
    }
  }
  if (rc) return rc;
  h->launches++;
  CU(cudaGetLastError());
  return HQPDOCP_OK;
}

int run(hqpdocp_handle *h, bool grads, int mode, const double *x, double *f, double *b, double *d, double *g,
        double *fx, double *fu, double *cx, double *cu) {
  if (!h) return fail("hqpdocp: NULL handle", HQPDOCP_E_ARG);
  const hqpdocp_dims &D = h->dims;
  if (!x || !f || !b || (h->m > 0 && !d)) return fail("hqpdocp_update: NULL argument", HQPDOCP_E_ARG);
  if (grads) {
    if (mode != HQPDOCP_GRAD_FD && mode != HQPDOCP_GRAD_AD)
      return fail("hqpdocp_update: unknown grad_mode", HQPDOCP_E_ARG);
    if (!g || !fx || !fu || (h->ncns > 0 && !cx) || (D.nc > 0 && !cu))
      return fail("hqpdocp_update: NULL derivative array", HQPDOCP_E_ARG);
  }
  CU(cudaSetDevice(h->device));
  UpdArgs a{};
  a.m = ModelArgs{D.K_total, D.nx, D.nu, D.nc, D.ncK, h->d_par, h->d_spar, D.nspar, D.k_first};
  a.KL = D.K;
  a.halo = D.k_first + D.K < D.K_total;
  a.x = x;
  a.fbase = h->fbase;
  a.f0k = h->f0k;
  a.cval = h->cval;
  a.b = b;
  a.g = g; a.fx = fx; a.fu = fu; a.cx = cx; a.cu = cu;
  a.npar = D.npar;
  int rc;
  switch (D.model) {
    case HQPDOCP_MODEL_DID: rc = launch_model<ModelDID>(h, grads, mode, a); break;
    case HQPDOCP_MODEL_SYNTHNL: rc = launch_model<ModelSynthNL>(h, grads, mode, a); break;
    default: return fail("hqpdocp: unknown model", HQPDOCP_E_UNSUPPORTED);
  }
  if (rc) return rc;
  const long long na = (long long)h->t[0].dim + h->t[1].dim + h->t[2].dim + h->t[3].dim + h->t[4].dim + h->t[5].dim;
  if (na > 0) {
    AssocArgs aa{x, h->cval, b, d, (long long)D.K * D.nx, h->t[0], h->t[1], h->t[2], h->t[3], h->t[4], h->t[5]};
    const int grid = (int)std::max<long long>(1, std::min<long long>((na + 255) / 256, (long long)h->sms * 8));
    docp_assoc_kernel<<<grid, 256, 0, h->stream>>>(aa, na);
    h->launches++;
  }
  docp_sum_kernel<<<1, 1024, 0, h->stream>>>(h->f0k, (long long)D.K + 1, f);
  h->launches++;
  CU(cudaGetLastError());
  return HQPDOCP_OK;
}

// device staging of the host-pointer entry points, allocated on first use; all or nothing
int stage_alloc(hqpdocp_handle *h) {
  if (h->staged) return HQPDOCP_OK;
  const hqpdocp_dims &D = h->dims;
  const size_t K = D.K;
  struct { double **p; size_t n; } want[] = {
      {&h->s_x, (size_t)h->N}, {&h->s_f, 1}, {&h->s_b, (size_t)std::max<long long>(1, h->me)},
      {&h->s_d, (size_t)std::max<long long>(1, h->m)}, {&h->s_g, (size_t)h->N},
      {&h->s_fx, std::max<size_t>(1, K * D.nx * D.nx)}, {&h->s_fu, std::max<size_t>(1, K * D.nx * D.nu)},
      {&h->s_cx, std::max<size_t>(1, (size_t)h->ncns * D.nx)}, {&h->s_cu, std::max<size_t>(1, K * D.nc * D.nu)}};
  for (auto &w : want) {
    cudaError_t e = cudaMalloc(w.p, sizeof(double) * w.n);
    if (e != cudaSuccess) {
      for (auto &v : want) {
        cudaFree(*v.p);
        *v.p = nullptr;
      }
      return fail(std::string("hqpdocp: staging buffers: ") + cudaGetErrorString(e), HQPDOCP_E_CUDA);
    }
  }
  h->staged = true;
  return HQPDOCP_OK;
}

}  // namespace

extern "C" {

const char *hqpdocp_last_error(void) { return g_err.c_str(); }

int hqpdocp_destroy(hqpdocp_handle *h) {
  if (!h) return HQPDOCP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream); else cudaDeviceSynchronize();
  cudaFree(h->d_par); cudaFree(h->d_spar);
  for (auto &t : h->t) { cudaFree(t.idxs); cudaFree(t.vals); }
  cudaFree(h->fbase); cudaFree(h->f0k); cudaFree(h->cval);
  cudaFree(h->s_x); cudaFree(h->s_f); cudaFree(h->s_b); cudaFree(h->s_d); cudaFree(h->s_g);
  cudaFree(h->s_fx); cudaFree(h->s_fu); cudaFree(h->s_cx); cudaFree(h->s_cu);
  delete h;
  return HQPDOCP_OK;
}

int hqpdocp_create(const hqpdocp_dims *dims, hqpdocp_handle **out) {
  if (!dims || !out) return fail("hqpdocp_create: NULL argument", HQPDOCP_E_ARG);
  *out = nullptr;
  const hqpdocp_dims &D = *dims;
  if (D.K < 1 || D.nx < 1 || D.nu < 0 || D.nc < 0 || D.ncK < 0 || D.npar < 0 || D.nspar < 0 ||
      (D.npar > 0 && !D.par) || (D.nspar > 0 && !D.spar))
    return fail("hqpdocp_create: bad dimensions", HQPDOCP_E_ARG);
  if (D.K_total < 0 || D.k_first < 0 || (D.K_total > 0 && D.k_first + D.K > D.K_total) ||
      (D.K_total == 0 && D.k_first != 0))
    return fail("hqpdocp_create: bad stage range", HQPDOCP_E_ARG);
  if (D.nx > MAXX || D.nu > MAXU || D.nc > MAXC || D.ncK > MAXC)
    return fail("hqpdocp_create: nx <= 64, nu <= 32, nc <= 8 per stage", HQPDOCP_E_UNSUPPORTED);
  bool ok;
  switch (D.model) {
    case HQPDOCP_MODEL_DID: ok = ModelDID::dims_ok(D.nx, D.nu, D.nc, D.ncK, D.npar, D.nspar); break;
    case HQPDOCP_MODEL_SYNTHNL: ok = ModelSynthNL::dims_ok(D.nx, D.nu, D.nc, D.ncK, D.npar, D.nspar); break;
    default: return fail("hqpdocp_create: unknown model", HQPDOCP_E_UNSUPPORTED);
  }
  if (!ok) return fail("hqpdocp_create: dimensions / parameter counts do not fit the model", HQPDOCP_E_UNSUPPORTED);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (D.device < 0 || D.device >= ndev) return fail("hqpdocp_create: no such CUDA device", HQPDOCP_E_CUDA);
  CU(cudaSetDevice(D.device));
  hqpdocp_handle *h = new hqpdocp_handle();
  h->dims = D;
  if (h->dims.K_total == 0) {  // the whole horizon
    h->dims.K_total = D.K;
    h->dims.k_first = 0;
  }
  const bool halo = h->dims.k_first + D.K < h->dims.K_total;
  h->device = D.device;
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, D.device);
  if (const char *e = getenv("HQPDOCP_XCOPY")) h->xcopy_mode = atoi(e);
  const long long K = D.K;
  h->N = K * (D.nx + D.nu) + D.nx;
  h->ncns = K * D.nc + (halo ? 0 : D.ncK);
  h->me = K * D.nx + D.xu_eq.dim + D.cns_eq.dim;
  h->m = (long long)D.xu_lb.dim + D.xu_ub.dim + D.cns_lb.dim + D.cns_ub.dim;
  int rc = HQPDOCP_OK;
  auto bail = [&](int code) {
    std::string keep = g_err;
    hqpdocp_destroy(h);
    g_err = keep;
    return code;
  };
#define CUH(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return bail(fail(std::string(#call) + ": " + cudaGetErrorString(e_), HQPDOCP_E_CUDA)); \
  } while (0)
  if (D.npar) {
    CUH(cudaMalloc(&h->d_par, sizeof(double) * D.npar));
    CUH(cudaMemcpy(h->d_par, D.par, sizeof(double) * D.npar, cudaMemcpyHostToDevice));
  }
  if (D.nspar) {
    CUH(cudaMalloc(&h->d_spar, sizeof(double) * (K + 1) * D.nspar));
    CUH(cudaMemcpy(h->d_spar, D.spar, sizeof(double) * (K + 1) * D.nspar, cudaMemcpyHostToDevice));
  }
  const hqpdocp_assoc *src[6] = {&D.xu_eq, &D.xu_lb, &D.xu_ub, &D.cns_eq, &D.cns_lb, &D.cns_ub};
  const char *nm[6] = {"xu_eq", "xu_lb", "xu_ub", "cns_eq", "cns_lb", "cns_ub"};
  for (int i = 0; i < 6; i++)
    if ((rc = upload_assoc(*src[i], h->t[i], i < 3 ? h->N : h->ncns, nm[i])) != HQPDOCP_OK) return bail(rc);
  CUH(cudaMalloc(&h->fbase, sizeof(double) * K * D.nx));
  CUH(cudaMalloc(&h->f0k, sizeof(double) * (K + 1)));
  CUH(cudaMalloc(&h->cval, sizeof(double) * std::max<long long>(1, h->ncns)));
#undef CUH
  // the handle's own pointers in dims would dangle: keep sizes only
  h->dims.par = nullptr;
  h->dims.spar = nullptr;
  for (auto *p : {&h->dims.xu_eq, &h->dims.xu_lb, &h->dims.xu_ub, &h->dims.cns_eq, &h->dims.cns_lb, &h->dims.cns_ub}) {
    p->idxs = nullptr;
    p->vals = nullptr;
  }
  *out = h;
  return HQPDOCP_OK;
}

int hqpdocp_set_stream(hqpdocp_handle *h, void *cuda_stream) {
  if (!h) return fail("hqpdocp_set_stream: NULL handle", HQPDOCP_E_ARG);
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  return HQPDOCP_OK;
}

int hqpdocp_sizes(const hqpdocp_handle *h, long long *N, long long *me, long long *m) {
  if (!h) return fail("hqpdocp_sizes: NULL handle", HQPDOCP_E_ARG);
  if (N) *N = h->N;
  if (me) *me = h->me;
  if (m) *m = h->m;
  return HQPDOCP_OK;
}

long long hqpdocp_launch_count(const hqpdocp_handle *h) { return h ? h->launches : 0; }

int hqpdocp_update_fbd_dev(hqpdocp_handle *h, const double *x, double *f, double *b, double *d) {
  return run(h, false, 0, x, f, b, d, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int hqpdocp_update_dev(hqpdocp_handle *h, int grad_mode, const double *x, double *f, double *b, double *d,
                       double *g, double *fx, double *fu, double *cx, double *cu) {
  return run(h, true, grad_mode, x, f, b, d, g, fx, fu, cx, cu);
}

int hqpdocp_update_fbd(hqpdocp_handle *h, const double *x, double *f, double *b, double *d) {
  if (!h || !x || !f || !b || (h->m > 0 && !d)) return fail("hqpdocp_update_fbd: NULL argument", HQPDOCP_E_ARG);
  CU(cudaSetDevice(h->device));
  int rc = stage_alloc(h);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->s_x, x, sizeof(double) * h->N, cudaMemcpyHostToDevice, h->stream));
  if ((rc = run(h, false, 0, h->s_x, h->s_f, h->s_b, h->s_d, nullptr, nullptr, nullptr, nullptr, nullptr)))
    return rc;
  CU(cudaMemcpyAsync(f, h->s_f, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(b, h->s_b, sizeof(double) * h->me, cudaMemcpyDeviceToHost, h->stream));
  if (h->m > 0) CU(cudaMemcpyAsync(d, h->s_d, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return HQPDOCP_OK;
}

int hqpdocp_update(hqpdocp_handle *h, int grad_mode, const double *x, double *f, double *b, double *d, double *g,
                   double *fx, double *fu, double *cx, double *cu) {
  if (!h || !x || !f || !b || (h->m > 0 && !d) || !g || !fx || !fu)
    return fail("hqpdocp_update: NULL argument", HQPDOCP_E_ARG);
  const hqpdocp_dims &D = h->dims;
  if ((h->ncns > 0 && !cx) || (D.nc > 0 && !cu)) return fail("hqpdocp_update: NULL constraint Jacobian", HQPDOCP_E_ARG);
  CU(cudaSetDevice(h->device));
  int rc = stage_alloc(h);
  if (rc) return rc;
  const size_t K = D.K;
  CU(cudaMemcpyAsync(h->s_x, x, sizeof(double) * h->N, cudaMemcpyHostToDevice, h->stream));
  if ((rc = run(h, true, grad_mode, h->s_x, h->s_f, h->s_b, h->s_d, h->s_g, h->s_fx, h->s_fu, h->s_cx, h->s_cu)))
    return rc;
  CU(cudaMemcpyAsync(f, h->s_f, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(b, h->s_b, sizeof(double) * h->me, cudaMemcpyDeviceToHost, h->stream));
  if (h->m > 0) CU(cudaMemcpyAsync(d, h->s_d, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(g, h->s_g, sizeof(double) * h->N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(fx, h->s_fx, sizeof(double) * K * D.nx * D.nx, cudaMemcpyDeviceToHost, h->stream));
  if (D.nu > 0) CU(cudaMemcpyAsync(fu, h->s_fu, sizeof(double) * K * D.nx * D.nu, cudaMemcpyDeviceToHost, h->stream));
  if (h->ncns > 0)
    CU(cudaMemcpyAsync(cx, h->s_cx, sizeof(double) * h->ncns * D.nx, cudaMemcpyDeviceToHost, h->stream));
  if (D.nc > 0 && D.nu > 0)
    CU(cudaMemcpyAsync(cu, h->s_cu, sizeof(double) * K * D.nc * D.nu, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return HQPDOCP_OK;
}

}  // extern "C"
