// docp_update.cu -- libhqpdocp.so: the stage loop of Hqp_Docp::update / ::update_fbd
// (hqp/Hqp_Docp.C:831-891, 944-1075), the default difference quotients of
// Hqp_Docp::update_grds (:1097-1180) and Hqp_Docp::update_bounds (:893-940) for
// device-resident stage models (docp_models.cuh).  C ABI: include/hqp_docpcuda.h.
// SURVEY.md section 8, row f4.  sm_100a; compiled with -fmad=false (see docp_models.cuh).
//
// Kernels (all streaming; nothing is re-read from HBM except the iterate x, which the
// nx+nu column threads of a stage share through L1/L2):
//   docp_vals_kernel   one thread per stage: f_k, f0_k, c_k; writes b's dynamics rows
//                      (f_k - x_{k+1}), the stage objective and the constraint values
//   docp_grds_kernel   one thread per (stage, variable): one perturbed evaluation (FD) or one
//                      dual-number evaluation (AD); lanes of a warp hold neighbouring columns
//                      of the same stage, so the row-major fx/fu/cx/cu stores coalesce
//   docp_assoc_kernel  the association tables -> b, d (update_bounds + the c_k rows)
//   docp_sum_kernel    f = sum_k f0_k, fixed order (one CTA)
#include <cuda_runtime.h>

#include <algorithm>
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
};

template <class Model>
__global__ void __launch_bounds__(128) docp_vals_kernel(UpdArgs a) {
  const ModelArgs &m = a.m;
  const int nd = m.nx + m.nu;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k <= m.K;
       k += (long long)gridDim.x * blockDim.x) {
    double x[MAXX], u[MAXU], f[MAXX], c[MAXC], f0 = 0.0;
    const double *xs = a.x + k * nd;
    for (int i = 0; i < m.nx; i++) x[i] = xs[i];
    const int nu = k < m.K ? m.nu : 0, nc = k < m.K ? m.nc : m.ncK;
    for (int j = 0; j < nu; j++) u[j] = xs[m.nx + j];
    for (int i = 0; i < m.nx; i++) f[i] = 0.0;  // v_zero(fk), hqp/Hqp_Docp.C:856
    for (int i = 0; i < nc; i++) c[i] = 0.0;
    Model::template vals<double>(m, (int)k, x, u, f, f0, c);
    a.f0k[k] = f0;
    if (k < m.K) {
      const double *xn = xs + nd;
      for (int i = 0; i < m.nx; i++) {
        a.fbase[k * m.nx + i] = f[i];
        a.b[k * m.nx + i] = f[i] - xn[i];  // v_sub(fk, x_{k+1}), :861
      }
    }
    for (int i = 0; i < nc; i++) a.cval[k * m.nc + i] = c[i];
  }
}

// one thread per (stage k, variable j of [x_k u_k]); stage K has its nx states only
template <class Model, int MODE>
__global__ void __launch_bounds__(128) docp_grds_kernel(UpdArgs a, long long total) {
  const ModelArgs &m = a.m;
  const int nd = m.nx + m.nu;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    long long k = w / nd;
    if (k > m.K) k = m.K;
    const int j = (int)(w - k * nd);
    const bool last = k == m.K;
    const int nu = last ? 0 : m.nu, nc = last ? m.ncK : m.nc, nf = last ? 0 : m.nx;
    const double *xs = a.x + k * nd;
    const double *fb = a.fbase + k * m.nx;
    const double *cb = a.cval + k * m.nc;
    double gj;
    if constexpr (MODE == HQPDOCP_GRAD_FD) {
      // Hqp_Docp::update_grds, hqp/Hqp_Docp.C:1127-1171
      double x[MAXX], u[MAXU], f[MAXX], c[MAXC], f0 = 0.0;
      for (int i = 0; i < m.nx; i++) x[i] = xs[i];
      for (int i = 0; i < nu; i++) u[i] = xs[m.nx + i];
      double *v = j < m.nx ? &x[j] : &u[j - m.nx];
      const double dvj = 1e-4 * fabs(*v) + 1e-6;
      *v += dvj;
      for (int i = 0; i < nf; i++) f[i] = 0.0;
      for (int i = 0; i < nc; i++) c[i] = 0.0;
      Model::template vals<double>(m, (int)k, x, u, f, f0, c);
      if (j < m.nx) {
        for (int i = 0; i < nf; i++) a.fx[(k * m.nx + i) * m.nx + j] = (f[i] - fb[i]) / dvj;
        for (int i = 0; i < nc; i++) a.cx[(k * m.nc + i) * m.nx + j] = (c[i] - cb[i]) / dvj;
      } else {
        for (int i = 0; i < nf; i++) a.fu[(k * m.nx + i) * m.nu + (j - m.nx)] = (f[i] - fb[i]) / dvj;
        for (int i = 0; i < nc; i++) a.cu[(k * m.nc + i) * m.nu + (j - m.nx)] = (c[i] - cb[i]) / dvj;
      }
      gj = (f0 - a.f0k[k]) / dvj;
    } else {
      Dual x[MAXX], u[MAXU], f[MAXX], c[MAXC], f0;
      for (int i = 0; i < m.nx; i++) x[i] = Dual(xs[i], i == j ? 1.0 : 0.0);
      for (int i = 0; i < nu; i++) u[i] = Dual(xs[m.nx + i], m.nx + i == j ? 1.0 : 0.0);
      Model::template vals<Dual>(m, (int)k, x, u, f, f0, c);
      if (j < m.nx) {
        for (int i = 0; i < nf; i++) a.fx[(k * m.nx + i) * m.nx + j] = f[i].d;
        for (int i = 0; i < nc; i++) a.cx[(k * m.nc + i) * m.nx + j] = c[i].d;
      } else {
        for (int i = 0; i < nf; i++) a.fu[(k * m.nx + i) * m.nu + (j - m.nx)] = f[i].d;
        for (int i = 0; i < nc; i++) a.cu[(k * m.nc + i) * m.nu + (j - m.nx)] = c[i].d;
      }
      gj = f0.d;
    }
    a.g[k * nd + j] = gj;  // f0x / f0u written in place into qp->c (:965-966)
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

template <class Model>
int launch_model(hqpdocp_handle *h, bool grads, int mode, const UpdArgs &a) {
  const long long K = h->dims.K;
  const int nd = h->dims.nx + h->dims.nu;
  {
    const int thr = 128;
    const long long want = (K + 1 + thr - 1) / thr;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->sms * 16));
    docp_vals_kernel<Model><<<grid, thr, 0, h->stream>>>(a);
    h->launches++;
  }
  if (grads) {
    const long long total = K * nd + h->dims.nx;
    const int thr = 128;
    const long long want = (total + thr - 1) / thr;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->sms * 16));
    if (mode == HQPDOCP_GRAD_FD)
      docp_grds_kernel<Model, HQPDOCP_GRAD_FD><<<grid, thr, 0, h->stream>>>(a, total);
    else
      docp_grds_kernel<Model, HQPDOCP_GRAD_AD><<<grid, thr, 0, h->stream>>>(a, total);
    h->launches++;
  }
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
  a.m = ModelArgs{D.K, D.nx, D.nu, D.nc, D.ncK, h->d_par, h->d_spar, D.nspar};
  a.x = x;
  a.fbase = h->fbase;
  a.f0k = h->f0k;
  a.cval = h->cval;
  a.b = b;
  a.g = g; a.fx = fx; a.fu = fu; a.cx = cx; a.cu = cu;
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

int stage_alloc(hqpdocp_handle *h) {
  if (h->s_x) return HQPDOCP_OK;
  const hqpdocp_dims &D = h->dims;
  const size_t K = D.K;
  CU(cudaMalloc(&h->s_x, sizeof(double) * h->N));
  CU(cudaMalloc(&h->s_f, sizeof(double)));
  CU(cudaMalloc(&h->s_b, sizeof(double) * std::max<long long>(1, h->me)));
  CU(cudaMalloc(&h->s_d, sizeof(double) * std::max<long long>(1, h->m)));
  CU(cudaMalloc(&h->s_g, sizeof(double) * h->N));
  CU(cudaMalloc(&h->s_fx, sizeof(double) * std::max<size_t>(1, K * D.nx * D.nx)));
  CU(cudaMalloc(&h->s_fu, sizeof(double) * std::max<size_t>(1, K * D.nx * D.nu)));
  CU(cudaMalloc(&h->s_cx, sizeof(double) * std::max<size_t>(1, (size_t)h->ncns * D.nx)));
  CU(cudaMalloc(&h->s_cu, sizeof(double) * std::max<size_t>(1, K * D.nc * D.nu)));
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
  h->device = D.device;
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, D.device);
  const long long K = D.K;
  h->N = K * (D.nx + D.nu) + D.nx;
  h->ncns = K * D.nc + D.ncK;
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
