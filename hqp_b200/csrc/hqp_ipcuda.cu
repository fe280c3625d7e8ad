// libhqpcuda.so -- host side of the C ABI declared in include/hqp_ipcuda.h.
//
// Owns every device allocation of one matrix-module instance, builds the
// stage-sorted / column-sorted index maps the kernels need from the caller's
// CSR structure, launches the factor / solve kernels (lq_factor.cuh,
// lq_solve.cuh) on the handle's stream and mirrors the reference's control
// flow that sits directly above them (Hqp_IpMatrix::solve refinement loop,
// hqp/Hqp_IpMatrix.C:65-128).  There is no CPU fallback: every entry point
// either runs the CUDA path or returns an error status.

#include "../../include/hqp_ipcuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound with dlopen (hqp_dist_host.inc)

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "lq_device.cuh"
#include "lq_factor.cuh"
#include "lq_solve.cuh"
#include "lq_eq.cuh"
#include "lq_ips.cuh"
#include "lq_range.cuh"

static thread_local std::string g_err;

// Compile-time specialisations of the factor kernels for the stage shapes named
// in BASELINE.json; any other shape runs the <0,0> (runtime-dimension) build of
// the same code.
#ifndef LQ_SCAN_R
#define LQ_SCAN_R 32  // largest radix of the solve hierarchy (buffers are sized for it)
#endif
#define LQ_DISPATCH_NXNU(nx_, nu_, CALL)                                       \
  do {                                                                        \
    if ((nx_) == 20 && (nu_) == 10) { CALL(20, 10); }                         \
    else if ((nx_) == 12 && (nu_) == 4) { CALL(12, 4); }                      \
    else if ((nx_) == 40 && (nu_) == 10) { CALL(40, 10); }                    \
    else { CALL(0, 0); }                                                      \
  } while (0)
#define LQ_DISPATCH_NX(nx_, nu_, CALL)                                         \
  do {                                                                        \
    if ((nx_) == 20 && (nu_) == 10) { CALL(20); }                             \
    else if ((nx_) == 12 && (nu_) == 4) { CALL(12); }                         \
    else if ((nx_) == 40 && (nu_) == 10) { CALL(40); }                        \
    else { CALL(0); }                                                         \
  } while (0)

#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return HQPCU_E_CUDA;                                                    \
    }                                                                         \
  } while (0)

// end of a launch sequence: the first launch error of the sequence (LAUNCHP
// records cudaLaunchKernelEx's status) or whatever the runtime still holds
#define CUL(h)                                                                \
  do {                                                                        \
    cudaError_t e_ = (h)->launch_err;                                         \
    (h)->launch_err = cudaSuccess;                                            \
    const cudaError_t g_ = cudaGetLastError();                                \
    if (e_ == cudaSuccess) e_ = g_;                                           \
    if (e_ != cudaSuccess) {                                                  \
      g_err = std::string("kernel launch: ") + cudaGetErrorString(e_);        \
      return HQPCU_E_CUDA;                                                    \
    }                                                                         \
  } while (0)

struct IpsState;
struct DistState;
struct MgState;

struct hqpcu_handle {
  LqDev d;
  IpsState *ips = nullptr;           // device-resident IP solver state (lazy)
  DistState *dist = nullptr;         // horizon split over NCCL (hqp_dist_host.inc)
  MgState *mg = nullptr;             // dispatcher over several GPUs of this process (hqp_mg_host.inc)
  std::vector<double> eq_rowsum;     // row sums |E| of the general equality rows
  std::vector<int> dims_eq_ptr;      // host copy of the equality CSR pointers
  hqpcu_dims dims;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  cudaError_t launch_err = cudaSuccess;  // first failed cudaLaunchKernelEx since the last check
  long long n_solves = 0, n_solve_steps = 0;  // hqpcu_solve_stats
  int device = 0;
  int nseg_req = 0;       // segment count asked for (0 = automatic); hqpcu_set_nseg updates it
  bool big = false;       // stage blocks exceed shared memory: global-workspace kernels (LqDev::gws)
  int nt2() const { return big ? LQ_BIG_NT : LQ_NT2; }  // threads per CTA of the tree kernels
  bool demoted = false;   // E_NOTPD fallback to the sequential sweep is in force until the next update
  std::vector<void *> allocs;
  // owned device copies
  double *Q = nullptr, *fx = nullptr, *fu = nullptr, *cval = nullptr, *z = nullptr, *w = nullptr;
  // staging for the host-pointer API and the refinement loop
  double *u_r1 = nullptr, *u_r2 = nullptr, *u_r3 = nullptr, *u_r4 = nullptr;
  double *u_dx = nullptr, *u_dy = nullptr, *u_dz = nullptr, *u_dw = nullptr;
  double *t1 = nullptr, *t2 = nullptr, *t3 = nullptr, *t4 = nullptr;
  double *e1 = nullptr, *e2 = nullptr, *e3 = nullptr, *e4 = nullptr;
  double *res_dev = nullptr;
  double *sqp_part = nullptr, *sqp_host = nullptr;  // reduction scratch of hqpcu_sqp_* (row f3)
  double *res_host = nullptr;  // pinned
  int *status_host = nullptr;  // pinned
  int status_seen = 0;         // status word as of the last residual read-back
  bool factored = false;
  // general stage equality rows (block elimination on top of the factor)
  LqEq q;
  double *eqval = nullptr, *eq_r1 = nullptr;
  // optional per-kernel CUDA-event timing (bench.py roofline section)
  bool profiling = false;
  struct Span { const char *name; cudaEvent_t e0, e1; };
  std::vector<Span> spans;
  size_t smem_k1 = 0, smem_k2 = 0, smem_k3 = 0, smem_cmp = 0, smem_psi = 0, smem_chain = 0;
  int thr_factor = 128, thr_chain = 128, thr_stage = 64;
  int max_el = 0;  // elements per instance the seg* arrays were sized for
  int n_sm = 148, k1_ctas_per_sm = 3;
  int seg_warps = 4;      // warps per segment CTA of K1/K3 (1: one warp per segment)
  int seg_warps_req = 0;  // HQPCU_SEG_WARPS override (0: automatic)
  int ring_scan = 8, psi_chunk = 2, ring_chain = LQ_RING;
  size_t smem_scan = 0;
  // CUDA graphs of the fixed launch sequences (factor; step per pointer set):
  // replayed with one cudaGraphLaunch instead of 13-21 dependent launches
  struct GraphEntry { std::vector<const void *> key; cudaGraphExec_t exec; long long launches; };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;
  bool use_hs = true;    // factor tree as a one-sweep suffix scan (HQPCU_HS=0: up/down tree)
  int hs_final = 0;      // slot offset of the region that holds the finished suffix elements
  // slot of the element that condenses the whole stage range (exported by a split horizon)
  int range_slot() const { return d.hs ? hs_final : d.ft.off[d.ft.nlev - 1]; }
  // programmatic dependent launch inside the hot sequences: measured SLOWER at C2
  // (0.845 vs 0.794 ms per unit: early-scheduled dependents hold SM resources), so
  // off unless HQPCU_PDL=1
  bool use_pdl = false;
  cudaStream_t cap_stream = nullptr;  // capture happens here (the user stream may be stream 0)
  // row f1: registered scatter map (values -> stage slabs)
  long long vm_n = 0;
  long long *vm_dst = nullptr, *vm_dst2 = nullptr;
  double *vm_vals = nullptr;
  // horizon split: right-hand sides remembered between the three step phases
  const double *rg_r1 = nullptr, *rg_r2 = nullptr, *rg_r3 = nullptr, *rg_r4 = nullptr;
  bool ranged() const { return d.has_prev || d.has_next; }
  // level scanned sequentially by the "top" kernels: the single root when the
  // range is part of a longer horizon, else the level below it
  int ftop() const { return (ranged() || d.ft.nlev < 2) ? d.ft.nlev - 1 : d.ft.nlev - 2; }
  int stop() const { return (ranged() || d.st.nlev < 2) ? d.st.nlev - 1 : d.st.nlev - 2; }
};

// launch wrapper: counts the launch and, when profiling, brackets it with
// CUDA events on the launching stream
#define LAUNCH(h, kname, ...)                                                 \
  do {                                                                        \
    hqpcu_handle::Span sp_{#kname, nullptr, nullptr};                         \
    if ((h)->profiling) {                                                     \
      cudaEventCreate(&sp_.e0);                                               \
      cudaEventCreate(&sp_.e1);                                               \
      cudaEventRecord(sp_.e0, (h)->stream);                                   \
    }                                                                         \
    kname __VA_ARGS__;                                                        \
    (h)->launches++;                                                          \
    if ((h)->profiling) {                                                     \
      cudaEventRecord(sp_.e1, (h)->stream);                                   \
      (h)->spans.push_back(sp_);                                              \
    }                                                                         \
  } while (0)

// Launch with programmatic stream serialisation (PDL): the kernel may be
// scheduled while its predecessor drains; it calls pdl_enter() before touching
// global memory, which waits for the predecessor's completion and flush.
template <class... KArgs, class... Args>
static cudaError_t launch_ex(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                             size_t smem, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define LAUNCHP(h, kname, grid_, block_, smem_, strm_, ...)                    \
  do {                                                                        \
    hqpcu_handle::Span sp_{#kname, nullptr, nullptr};                         \
    if ((h)->profiling) {                                                     \
      cudaEventCreate(&sp_.e0);                                               \
      cudaEventCreate(&sp_.e1);                                               \
      cudaEventRecord(sp_.e0, (h)->stream);                                   \
    }                                                                         \
    {                                                                         \
      const cudaError_t le_ = launch_ex((h)->use_pdl && !(h)->profiling, kname, grid_, block_,    \
                                        smem_, strm_, __VA_ARGS__);                               \
      if (le_ != cudaSuccess && (h)->launch_err == cudaSuccess) (h)->launch_err = le_;            \
    }                                                                         \
    (h)->launches++;                                                          \
    if ((h)->profiling) {                                                     \
      cudaEventRecord(sp_.e1, (h)->stream);                                   \
      (h)->spans.push_back(sp_);                                              \
    }                                                                         \
  } while (0)

static void drop_graphs(hqpcu_handle *h) {
  for (auto &g : h->graphs) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

// Run `body` (a fixed sequence of launches on h->stream that depends only on
// `key`) through a cached CUDA graph: captured on first use, replayed afterwards.
template <class F>
static int run_graphed(hqpcu_handle *h, std::vector<const void *> key, F &&body) {
  if (!h->use_graphs || h->profiling || !h->cap_stream) return body();
  for (auto &g : h->graphs)
    if (g.key == key) {
      CU(cudaGraphLaunch(g.exec, h->stream));
      h->launches += g.launches;
      return HQPCU_OK;
    }
  cudaStream_t user = h->stream;
  const long long l0 = h->launches;
  h->stream = h->cap_stream;
  cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {  // capture unavailable: plain launches
    cudaGetLastError();
    h->stream = user;
    h->use_graphs = false;
    return body();
  }
  int rc = body();
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(h->cap_stream, &graph);
  h->stream = user;
  const long long nl = h->launches - l0;
  h->launches = l0;
  if (rc || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc) return rc;
    h->use_graphs = false;
    return body();
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    cudaGetLastError();
    h->use_graphs = false;
    return body();
  }
  if (h->graphs.size() >= 16) {  // bounded cache: drop the oldest
    cudaGraphExecDestroy(h->graphs.front().exec);
    h->graphs.erase(h->graphs.begin());
  }
  h->graphs.push_back({std::move(key), exec, nl});
  CU(cudaGraphLaunch(exec, h->stream));
  h->launches += nl;
  return HQPCU_OK;
}

template <typename T>
static int dev_alloc(hqpcu_handle *h, T **p, size_t count) {
  void *q = nullptr;
  CU(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  CU(cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = static_cast<T *>(q);
  return HQPCU_OK;
}

template <typename T>
static int dev_upload(hqpcu_handle *h, const T **p, const std::vector<T> &v) {
  T *q = nullptr;
  int rc = dev_alloc(h, &q, v.size());
  if (rc) return rc;
  if (!v.empty()) CU(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *p = q;
  return HQPCU_OK;
}

static size_t pad2(size_t n) { return (n + 1) & ~size_t(1); }

static void build_tree(LqTree &t, int P, int R) {
  t.R = R;
  t.nlev = 0;
  int cnt = P, off = 0;
  for (;;) {
    t.cnt[t.nlev] = cnt;
    t.off[t.nlev] = off;
    off += cnt;
    t.nlev++;
    if (cnt == 1 || t.nlev == LQ_MAXLEV) break;
    cnt = (cnt + R - 1) / R;
  }
  t.nel = off;
}

// segments per instance, stages per segment and the hierarchies above them
// One warp per segment (no CTA barriers, 3x3 register-blocked products) wins
// when there are many small independent recursions (measured at 4096 x nx=12:
// factor 0.55 vs 0.67 ms); a single horizon of few, fat segments needs the
// four cooperating warps.
static void choose_seg_warps(hqpcu_handle *h) {
  if (h->seg_warps_req)
    h->seg_warps = h->seg_warps_req;
  else if (h->smem_k3 && 8 * h->smem_k3 <= 227 * 1024 &&
           (long long)h->d.P * h->dims.batch >= 8LL * h->n_sm)
    h->seg_warps = 1;
  else
    // 5x5 tiles per product at nx = 40: eight warps per segment (measured at the
    // C5 slice: factor 13.4 vs 14.7 ms)
    h->seg_warps = h->dims.nx >= 32 ? 8 : 4;
}

static void choose_segments(hqpcu_handle *h, int nseg) {
  drop_graphs(h);  // the launch geometry is about to change
  const int K = h->dims.K;
  LqDev &d = h->d;
  int P = nseg;
  if (P <= 0) {
    // enough independent instances already fill the machine: sequential sweep
    if (h->dims.batch >= 64 || K < 32)
      P = 1;
    else {
      // one wave of segment CTAs: (SMs x resident CTAs per SM of the K1 kernel),
      // but never fewer than 8 stages per segment
      // (one resident CTA per SM at nx >= 32: two waves, the shorter solve chains
      //  pay for it -- measured at the C5 slice)
      const int wave = std::max(1, h->n_sm * std::max(1, h->k1_ctas_per_sm)) *
                       (h->dims.nx >= 32 && !h->big ? 2 : 1);
      const int per_inst = std::max(1, wave / std::max(1, h->dims.batch));
      P = std::max(2, std::min(per_inst, K / 8));
    }
  }
  P = std::max(1, std::min(P, std::max(1, K / 2)));
  P = std::min(P, 16384);
  int L = (K + P - 1) / P;
  if (L < 1) L = 1;
  P = std::max(1, (K + L - 1) / L);
  d.P = P;
  d.L = L;
  // stage-parallel solve passes: one stage per warp for a single horizon, two
  // when many small instances make CTA turnover the cost (C3: step 735 vs 790 us)
  {
    const char *env = getenv("HQPCU_SPW");
    d.spw = env ? std::max(1, std::min(2, atoi(env))) : (h->dims.batch >= 64 ? 2 : 1);
    // lanes per stage: half a warp when a stage is at most 16 wide and there are
    // enough of them (a full warp would leave most lanes idle); implies spw = 2
    const char *lg = getenv("HQPCU_LANE_GROUP");
    d.lgw = lg ? (atoi(lg) == 16 ? 16 : 32)
               : ((h->dims.nx + h->dims.nu <= 16 && h->dims.batch >= 64) ? 16 : 32);
    if (d.lgw == 16) d.spw = 2;
  }
  choose_seg_warps(h);
  build_tree(d.ft, P, 2);
  d.ft.nel = std::max(d.ft.nel, 2 * P + 2);  // (two ping-pong regions of P+1 slots for the suffix scan)
  d.hs = (h->use_hs && (P > 1 || h->ranged())) ? 1 : 0;
  // solve hierarchy: up (R steps) + top (P/R) + down (R) sequential chain steps,
  // shortest for R ~ sqrt(P) (measured at P = 435: R = 21 beats 32 by 5 % per step)
  int R = (int)std::ceil(std::sqrt((double)P));
  R = std::max(8, std::min(R, LQ_SCAN_R));
  build_tree(d.st, P, R);
}

// The attribute is global per function and device: handles with different block
// sizes share it, so it is only ever raised (a smaller, later handle must not
// lower the limit under an earlier one).
static int set_smem(const void *fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> cur;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t &have = cur[{dev, fn}];
  if (bytes > 48 * 1024 && bytes > have) {
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  // all of the L1/shared array as shared memory: these kernels live in shared
  // memory and their occupancy is bounded by it (ncu: occupancy_limit_shared_mem)
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                          cudaSharedmemCarveoutMaxShared));
  return HQPCU_OK;
}

static void ips_free(hqpcu_handle *h);
// horizon split over NCCL (hqp_dist_host.inc): after hqpcu_comm_init the handle is
// one rank's stage range and every entry point works on that range's slices
extern "C" {  // (defined inside this file's extern "C" block)
static int mg_create(const hqpcu_dims *dims, hqpcu_handle **out);
static void mg_free(hqpcu_handle *h);
static int mg_update(hqpcu_handle *h, const double *Q, const double *fx, const double *fu,
                     const double *cv);
static int mg_factor(hqpcu_handle *h, const double *z, const double *w);
static int mg_apply(hqpcu_handle *h, int mode, double eps, const double *r1, const double *r2,
                    const double *r3, const double *r4, double *dx, double *dy, double *dz,
                    double *dw, double *res, int *nsteps);
static void dist_free(hqpcu_handle *h);
static bool dist_on(const hqpcu_handle *h);
static int dist_launch_factor(hqpcu_handle *h);
static int dist_launch_step(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                            const double *r4, double *dx, double *dy, double *dz, double *dw);
}

extern "C" {

const char *hqpcu_last_error(void) { return g_err.c_str(); }

int hqpcu_create(const hqpcu_dims *dims, hqpcu_handle **out) {
  if (!dims || !out) return HQPCU_E_NULL;
  *out = nullptr;
  if (dims->K < 1 || dims->nx < 1 || dims->nu < 0 || dims->batch < 1 || dims->n_ineq < 0 ||
      dims->n_eq < 0) {
    g_err = "hqpcu_create: bad dimensions";
    return HQPCU_E_SIZES;
  }
  if (dims->nu < 1) {
    g_err = "hqpcu_create: nu = 0 is not supported";
    return HQPCU_E_UNSUPPORTED;
  }
  if (dims->ngpu > 1) return mg_create(dims, out);  // dispatcher over several GPUs
  if (dims->nx > 256 || dims->nu > 256) {
    g_err = "hqpcu_create: stage blocks larger than 256 are not supported";
    return HQPCU_E_UNSUPPORTED;
  }
  if (dims->n_eq > 0 && dims->batch != 1) {
    g_err = "hqpcu_create: general stage equality rows need batch == 1";
    return HQPCU_E_UNSUPPORTED;
  }
  if (dims->n_eq > 64) {
    g_err = "hqpcu_create: more than 64 general stage equality rows";
    return HQPCU_E_UNSUPPORTED;
  }
  if (dims->n_eq > 0 && (!dims->eq_stage || !dims->eq_ptr || !dims->eq_lcol)) return HQPCU_E_NULL;
  if (dims->n_ineq > 0 && (!dims->ineq_stage || !dims->ineq_ptr || !dims->ineq_lcol))
    return HQPCU_E_NULL;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (dims->device < 0 || dims->device >= ndev) {
    g_err = "hqpcu_create: no such CUDA device";
    return HQPCU_E_CUDA;
  }
  CU(cudaSetDevice(dims->device));

  hqpcu_handle *h = new hqpcu_handle;
  h->dims = *dims;
  h->device = dims->device;
  {
    const char *sw = getenv("HQPCU_SEG_WARPS");
    if (sw) h->seg_warps_req = atoi(sw) == 1 ? 1 : (atoi(sw) == 8 ? 8 : 4);
    const char *pe = getenv("HQPCU_PDL");
    h->use_pdl = pe && pe[0] == '1';
    const char *hs = getenv("HQPCU_HS");
    h->use_hs = !(hs && hs[0] == '0');
    const char *env = getenv("HQPCU_GRAPHS");  // "0": plain launches (debugging, ncu per-kernel lists)
    h->use_graphs = !(env && env[0] == '0');
    if (h->use_graphs &&
        cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      h->cap_stream = nullptr;
    }
  }
  h->nseg_req = dims->nseg;
  LqDev &d = h->d;
  memset(&d, 0, sizeof d);
  const int K = dims->K, nx = dims->nx, nu = dims->nu, nm = nx + nu, B = dims->batch;
  const int m = dims->n_ineq;
  d.K = K; d.nx = nx; d.nu = nu; d.nm = nm; d.batch = B;
  d.fixed_x0 = dims->fixed_x0 ? 1 : 0;
  d.m = m;
  d.nnz = m ? dims->ineq_ptr[m] : 0;
  d.N = K * nm + nx;
  d.me = K * nx + (d.fixed_x0 ? nx : 0) + dims->n_eq;
  memset(&h->q, 0, sizeof h->q);
  h->q.n_eq = dims->n_eq;
  h->q.nnz = dims->n_eq ? dims->eq_ptr[dims->n_eq] : 0;
  {
    // resident K1 CTAs per SM (shared-memory bound) -> one wave of segments
    const size_t p2 = 2 * (pad2((size_t)nm * nm) + pad2((size_t)nx * nx) + pad2((size_t)nx * nu) +
                           pad2((size_t)nm)) * sizeof(double);
    const size_t k1 = p2 + (5 * pad2((size_t)nx * nx) + pad2((size_t)nx * nm) +
                            pad2((size_t)nu * nx) + 2 * pad2((size_t)nx * nu)) * sizeof(double);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dims->device);
    // 3 was the measured optimum at nx=20 (4 fit by size but run in two waves)
    h->k1_ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (227 * 1024) / (k1 + 1024)));
    h->big = k1 + 1024 > 227 * 1024;
  }
  choose_segments(h, dims->nseg);

  // ---- index maps -------------------------------------------------------
  std::vector<int> stage(dims->ineq_stage, dims->ineq_stage + m);
  std::vector<int> ptr(m + 1, 0), lcol(d.nnz);
  if (m) {
    std::copy(dims->ineq_ptr, dims->ineq_ptr + m + 1, ptr.begin());
    std::copy(dims->ineq_lcol, dims->ineq_lcol + d.nnz, lcol.begin());
  }
  for (int r = 0; r < m; r++) {
    const int k = stage[r];
    if (k < 0 || k > K || ptr[r + 1] < ptr[r]) {
      delete h;
      g_err = "hqpcu_create: inequality row outside the horizon";
      return HQPCU_E_SIZES;
    }
    const int dk = k < K ? nm : nx;
    for (int e = ptr[r]; e < ptr[r + 1]; e++)
      if (lcol[e] < 0 || lcol[e] >= dk) {
        delete h;
        g_err = "hqpcu_create: inequality column outside its stage block";
        return HQPCU_E_SIZES;
      }
  }
  std::vector<int> srow_ptr(K + 2, 0), srow(m);
  for (int r = 0; r < m; r++) srow_ptr[stage[r] + 1]++;
  for (int k = 0; k <= K; k++) srow_ptr[k + 1] += srow_ptr[k];
  {
    std::vector<int> fill(srow_ptr.begin(), srow_ptr.end() - 1);
    for (int r = 0; r < m; r++) srow[fill[stage[r]]++] = r;
  }
  std::vector<int> grow_ptr(K + 2, 0), grow;
  for (int k = 0; k <= K; k++) {
    for (int rr = srow_ptr[k]; rr < srow_ptr[k + 1]; rr++)
      if (ptr[srow[rr] + 1] - ptr[srow[rr]] > 1) grow.push_back(srow[rr]);
    grow_ptr[k + 1] = (int)grow.size();
  }
  std::vector<int> vcol_ptr(d.N + 1, 0), vcol_row(d.nnz), vcol_nz(d.nnz);
  for (int r = 0; r < m; r++)
    for (int e = ptr[r]; e < ptr[r + 1]; e++) vcol_ptr[stage[r] * nm + lcol[e] + 1]++;
  for (int j = 0; j < d.N; j++) vcol_ptr[j + 1] += vcol_ptr[j];
  {
    std::vector<int> fill(vcol_ptr.begin(), vcol_ptr.end() - 1);
    for (int r = 0; r < m; r++)
      for (int e = ptr[r]; e < ptr[r + 1]; e++) {
        const int pos = fill[stage[r] * nm + lcol[e]]++;
        vcol_row[pos] = r;
        vcol_nz[pos] = e;
      }
  }
  int rc = HQPCU_OK;
#define TRY(x) do { if ((rc = (x)) != HQPCU_OK) { hqpcu_destroy(h); return rc; } } while (0)
  TRY(dev_upload(h, &d.ineq_stage, stage));
  TRY(dev_upload(h, &d.ineq_ptr, ptr));
  TRY(dev_upload(h, &d.ineq_lcol, lcol));
  TRY(dev_upload(h, &d.srow_ptr, srow_ptr));
  TRY(dev_upload(h, &d.srow, srow));
  TRY(dev_upload(h, &d.grow_ptr, grow_ptr));
  TRY(dev_upload(h, &d.grow, grow));
  TRY(dev_upload(h, &d.vcol_ptr, vcol_ptr));
  TRY(dev_upload(h, &d.vcol_row, vcol_row));
  TRY(dev_upload(h, &d.vcol_nz, vcol_nz));

  if (dims->n_eq) {
    const int ne = dims->n_eq;
    std::vector<int> es(dims->eq_stage, dims->eq_stage + ne), ep(dims->eq_ptr, dims->eq_ptr + ne + 1),
        ec(dims->eq_lcol, dims->eq_lcol + h->q.nnz);
    for (int r = 0; r < ne; r++) {
      const int k = es[r];
      bool ok = k >= 0 && k <= K && ep[r + 1] >= ep[r];
      for (int e = ep[r]; ok && e < ep[r + 1]; e++) ok = ec[e] >= 0 && ec[e] < (k < K ? nm : nx);
      if (!ok) {
        g_err = "hqpcu_create: equality row outside its stage block";
        hqpcu_destroy(h);
        return HQPCU_E_SIZES;
      }
    }
    h->dims_eq_ptr = ep;
    TRY(dev_upload(h, &h->q.stage, es));
    TRY(dev_upload(h, &h->q.ptr, ep));
    TRY(dev_upload(h, &h->q.lcol, ec));
    TRY(dev_alloc(h, &h->eqval, (size_t)h->q.nnz));
    h->q.val = h->eqval;
    TRY(dev_alloc(h, &h->q.DX, (size_t)ne * d.N));
    TRY(dev_alloc(h, &h->q.DY, (size_t)ne * d.me));
    TRY(dev_alloc(h, &h->q.DZ, (size_t)ne * std::max(m, 1)));
    TRY(dev_alloc(h, &h->q.DW, (size_t)ne * std::max(m, 1)));
    TRY(dev_alloc(h, &h->q.S, (size_t)ne * ne));
    TRY(dev_alloc(h, &h->q.Sinv, (size_t)ne * ne));
    TRY(dev_alloc(h, &h->q.y, (size_t)ne));
    TRY(dev_alloc(h, &h->q.ety, (size_t)d.N));
    TRY(dev_alloc(h, &h->eq_r1, (size_t)d.N));
  }

  // ---- slabs --------------------------------------------------------------
  const size_t SB = (size_t)B;
  TRY(dev_alloc(h, &h->Q, SB * (K + 1) * nm * nm));
  TRY(dev_alloc(h, &h->fx, SB * K * nx * nx));
  TRY(dev_alloc(h, &h->fu, SB * K * nx * nu));
  TRY(dev_alloc(h, &h->cval, SB * d.nnz));
  TRY(dev_alloc(h, &h->z, SB * m));
  TRY(dev_alloc(h, &h->w, SB * m));
  d.Q = h->Q; d.fx = h->fx; d.fu = h->fu; d.cval = h->cval; d.z = h->z; d.w = h->w;
  TRY(dev_alloc(h, &d.V, SB * (K + 1) * nx * nx));
  TRY(dev_alloc(h, &d.Rux, SB * K * nu * nx));
  TRY(dev_alloc(h, &d.LD, SB * K * nu * nu));
  TRY(dev_alloc(h, &d.ldkind, SB * K));
  TRY(dev_alloc(h, &d.Phi, SB * K * nx * nx));
  // bulk-copy (TMA) path needs every per-stage slab to be a 16-byte multiple
  d.use_tma = (nx % 2 == 0 && nu % 2 == 0 && !h->big) ? 1 : 0;
  TRY(dev_alloc(h, &d.hdiag, SB * d.N));
  // element arrays are sized for the hierarchy chosen here; hqpcu_set_nseg may
  // only shrink it
  // (the solve hierarchy's radix follows P: any radix >= 2 needs < 2 P elements)
  const size_t PM = (size_t)2 * std::max(d.P, 1) + 64;
  const size_t PF = (size_t)std::max(d.ft.nel, 1);
  h->max_el = d.P;
  h->dims.nseg = d.P;
  TRY(dev_alloc(h, &d.segA, SB * PF * nx * nx));
  TRY(dev_alloc(h, &d.segC, SB * PF * nx * nx));
  TRY(dev_alloc(h, &d.segJ, SB * PF * nx * nx));
  TRY(dev_alloc(h, &d.segPsi, SB * PM * nx * nx));
  TRY(dev_alloc(h, &d.segVb, SB * PF * nx * nx));
  TRY(dev_alloc(h, &d.V0f, SB * nx * nx));
  TRY(dev_alloc(h, &d.Vext, (size_t)nx * nx));
  TRY(dev_alloc(h, &d.xstart, (size_t)nx));
  TRY(dev_alloc(h, &d.status, 1));
  TRY(dev_alloc(h, &d.dbg, 16));
  TRY(dev_alloc(h, &d.g, SB * d.N));
  TRY(dev_alloc(h, &d.wv, SB * K * nx));
  TRY(dev_alloc(h, &d.q, SB * K * nx));
  TRY(dev_alloc(h, &d.v, SB * (K + 1) * nx));
  TRY(dev_alloc(h, &d.Ru, SB * K * nu));
  TRY(dev_alloc(h, &d.c, SB * K * nx));
  TRY(dev_alloc(h, &d.x, SB * (K + 1) * nx));
  TRY(dev_alloc(h, &d.segv0, SB * PM * nx));
  TRY(dev_alloc(h, &d.segvb, SB * PM * nx));
  TRY(dev_alloc(h, &d.segx0, SB * PM * nx));
  TRY(dev_alloc(h, &d.segxa, SB * PM * nx));
  TRY(dev_alloc(h, &h->u_r1, SB * d.N)); TRY(dev_alloc(h, &h->u_dx, SB * d.N));
  TRY(dev_alloc(h, &h->t1, SB * d.N));   TRY(dev_alloc(h, &h->e1, SB * d.N));
  TRY(dev_alloc(h, &h->u_r2, SB * d.me)); TRY(dev_alloc(h, &h->u_dy, SB * d.me));
  TRY(dev_alloc(h, &h->t2, SB * d.me));   TRY(dev_alloc(h, &h->e2, SB * d.me));
  TRY(dev_alloc(h, &h->u_r3, SB * m)); TRY(dev_alloc(h, &h->u_dz, SB * m));
  TRY(dev_alloc(h, &h->t3, SB * m));   TRY(dev_alloc(h, &h->e3, SB * m));
  TRY(dev_alloc(h, &h->u_r4, SB * m)); TRY(dev_alloc(h, &h->u_dw, SB * m));
  TRY(dev_alloc(h, &h->t4, SB * m));   TRY(dev_alloc(h, &h->e4, SB * m));
  TRY(dev_alloc(h, &h->res_dev, 1));
  if (cudaMallocHost(&h->res_host, sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&h->status_host, sizeof(int)) != cudaSuccess) {
    g_err = "cudaMallocHost failed";
    hqpcu_destroy(h);
    return HQPCU_E_CUDA;
  }

  // ---- launch geometry ------------------------------------------------------
  h->thr_factor = nm <= 32 ? 128 : 256;
  h->thr_chain = 4 * (((nx + 31) / 32) * 32);
  h->thr_stage = std::max(32, ((nm + 31) / 32) * 32);
  const size_t nn = pad2((size_t)nx * nx), nf = pad2((size_t)nx * nm), gg = pad2((size_t)nm * nm),
               ru = pad2((size_t)nu * nx), xu = pad2((size_t)nx * nu), hd = pad2((size_t)nm);
  const size_t pipe = 2 * (gg + nn + xu + hd) * sizeof(double) + 2 * sizeof(uint64_t);
  {
    // CTA-internal blocks of K1/K3 with the padded strides of the tensor-core
    // instantiations (lq_pad4); the generic kernels use less
    bool compiled = false;
#define IS_(NX_) if ((NX_) > 0) compiled = true;
    LQ_DISPATCH_NX(nx, nu, IS_);
#undef IS_
    const size_t LV = compiled ? lq_pad4(nx) : nx, LT = compiled ? lq_pad4(nm) : nm;
    const size_t bv = pad2((size_t)nx * LV), bt = pad2((size_t)nx * LT), bu = pad2((size_t)nu * LV);
    const size_t bf = compiled ? pad2((size_t)nx * lq_pad4(nu)) : 0;  // re-packed fu
    h->smem_k1 = pipe + (bf + 4 * bv + bt + bu + bv + 2 * bu) * sizeof(double);
    h->smem_k3 = pipe + (bf + bv + bt + 2 * bu + 3 * bv) * sizeof(double);
  }
  choose_seg_warps(h);
  // (+ odd-stride augmented matrix and the scratch of the warp inverse)
  const size_t invs = pad2((size_t)nx * (nx + 1) + 2 * (nx + 2));
  h->smem_k2 = (4 * nn + pad2((size_t)nx * (2 * nx + 1)) + invs) * sizeof(double);
  h->smem_cmp = (5 * nn + pad2((size_t)nx * (3 * nx + 1)) + 3 * nn + pad2((size_t)nx * (2 * nx + 4)) + invs) *
                sizeof(double);
  // children of one group multiplied as a tree in shared memory, in chunks that fit
  h->psi_chunk = std::max(2, std::min(LQ_SCAN_R, (int)((192 * 1024) / (nn * sizeof(double)) * 2 / 3)));
  h->smem_psi = (size_t)(h->psi_chunk + (h->psi_chunk + 1) / 2) * nn * sizeof(double);
  {
    const char *env = getenv("HQPCU_CHAIN_CHUNK");  // tuning knob: stages per TMA chunk
    h->ring_chain = env ? std::max(1, atoi(env)) : LQ_RING;
  }
  h->smem_chain = (pad2((size_t)2 * h->ring_chain * (nx * nx + 2 * nx)) + pad2((size_t)3 * nx)) * sizeof(double) +
                  2 * sizeof(uint64_t) + 16;
  // the hierarchy scans run few CTAs: stage every matrix of a group up front
  h->ring_scan = std::max(1, std::min(LQ_SCAN_R, (int)((100 * 1024) / ((nx * nx + 2 * nx) * sizeof(double)))));
  h->smem_scan = (pad2((size_t)2 * h->ring_scan * (nx * nx + 2 * nx)) + pad2((size_t)3 * nx)) * sizeof(double) +
                 2 * sizeof(uint64_t) + 16;
  const size_t smem_max = 227 * 1024;
  if (h->smem_k1 > smem_max || h->smem_k2 > smem_max || h->smem_k3 > smem_max ||
      h->smem_cmp > smem_max || h->smem_psi > smem_max)
    h->big = true;
  if (h->big) {
    // Large stage blocks: the CTA-internal blocks live in a per-CTA slice of a
    // global workspace (LqDev::gws); shared memory holds barriers and flags only.
    d.use_tma = 0;
    const size_t need = std::max({h->smem_k1, h->smem_k2, h->smem_k3, h->smem_cmp, h->smem_psi,
                                  nn * sizeof(double)}) + (size_t)(nx + 16) * sizeof(double) +
                        big_gj_scratch_doubles(nx) * sizeof(double);  // (tail: scratch of the blocked elimination)
    d.gws_stride = pad2(need / sizeof(double) + 2);
    TRY(dev_alloc(h, &d.gws, (size_t)std::max(d.P, 1) * B * d.gws_stride));
    // shared memory = the GEMM staging ring of the CTA (cta_mm_big), for K1/K3 followed by
    // the LDL^T factor of Guu; LQ_BIG_NT threads per CTA
    h->smem_k1 = h->smem_k3 = (big_ldlt_fits(nx, nu, LQ_BIG_NT) ? big_seg_smem_doubles(nx, nu, LQ_BIG_NT)
                                                                : (size_t)LQ_BIG_STG) * sizeof(double);
    h->smem_k2 = h->smem_cmp = h->smem_psi = (size_t)LQ_BIG_STG * sizeof(double);
    // (the chains read their matrices from global memory directly: no ring)
    h->smem_chain = h->smem_scan = pad2((size_t)3 * nx) * sizeof(double) + 2 * sizeof(uint64_t) + 16;
  }
  // the middle pass keeps one LDL^T factor of Guu per warp in shared memory
  {
    const size_t sv = (size_t)LQ_WPB * 2 * (nx + nu + (size_t)nu * nu) * sizeof(double);
    if (sv > smem_max) {
      g_err = "hqpcu_create: nu too large for the stage-parallel solve pass";
      hqpcu_destroy(h);
      return HQPCU_E_UNSUPPORTED;
    }
    TRY(set_smem((const void *)(solve_mid_kernel<1, 32>), sv));
    TRY(set_smem((const void *)(solve_mid_kernel<2, 32>), sv));
    TRY(set_smem((const void *)(solve_mid_kernel<2, 16>), sv));
  }
#define SET_A(NX_, NU_)                                                        \
  TRY(set_smem((const void *)seg_element_kernel<NX_, NU_, 4>, h->smem_k1));   \
  TRY(set_smem((const void *)seg_riccati_kernel<NX_, NU_, 4>, h->smem_k3));   \
  TRY(set_smem((const void *)seg_element_kernel<NX_, NU_, ((NX_) >= 32 ? 8 : 4)>, h->smem_k1)); \
  TRY(set_smem((const void *)seg_riccati_kernel<NX_, NU_, ((NX_) >= 32 ? 8 : 4)>, h->smem_k3)); \
  TRY(set_smem((const void *)seg_element_kernel<NX_, NU_, ((NX_) > 0 ? 1 : 4)>, h->smem_k1)); \
  TRY(set_smem((const void *)seg_riccati_kernel<NX_, NU_, ((NX_) > 0 ? 1 : 4)>, h->smem_k3)); \
  TRY(set_smem((const void *)seg_element_kernel<NX_, NU_, ((NX_) == 0 ? LQ_BIG_NT / 32 : 4)>, h->smem_k1)); \
  TRY(set_smem((const void *)seg_riccati_kernel<NX_, NU_, ((NX_) == 0 ? LQ_BIG_NT / 32 : 4)>, h->smem_k3));
#define SET_B(NX_)                                                             \
  TRY(set_smem((const void *)elem_scan_kernel<NX_>, h->smem_k2));             \
  TRY(set_smem((const void *)range_scan_factor_kernel<NX_>, h->smem_k2));     \
  TRY(set_smem((const void *)elem_compose_kernel<NX_>, h->smem_cmp));         \
  TRY(set_smem((const void *)elem_hs_kernel<NX_>, h->smem_cmp));              \
  TRY(set_smem((const void *)psi_compose_kernel<NX_>, h->smem_psi));
  LQ_DISPATCH_NXNU(nx, nu, SET_A);
  LQ_DISPATCH_NX(nx, nu, SET_B);
#undef SET_A
#undef SET_B
  if (!h->big) TRY(set_smem((const void *)x0_factor_kernel, nn * sizeof(double)));
  if (h->smem_chain > smem_max) {
    g_err = "hqpcu_create: chain ring exceeds shared memory";
    hqpcu_destroy(h);
    return HQPCU_E_UNSUPPORTED;
  }
#define SET_C(NX_)                                                             \
  TRY(set_smem((const void *)solve_back_kernel<NX_>, h->smem_chain));         \
  TRY(set_smem((const void *)solve_fwd_kernel<NX_>, h->smem_chain));          \
  TRY(set_smem((const void *)(solve_scan_kernel<true, NX_>), h->smem_scan));  \
  TRY(set_smem((const void *)(solve_scan_kernel<false, NX_>), h->smem_scan));
  LQ_DISPATCH_NX(nx, nu, SET_C);
#undef SET_C
  if (dims->n_eq)
    TRY(set_smem((const void *)eq_invert_kernel, (size_t)3 * dims->n_eq * dims->n_eq * sizeof(double)));
#undef TRY
  *out = h;
  return HQPCU_OK;
}

int hqpcu_destroy(hqpcu_handle *h) {
  if (!h) return HQPCU_OK;
  if (h->mg) {
    mg_free(h);
    delete h;
    return HQPCU_OK;
  }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  drop_graphs(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->vm_dst) { cudaFree(h->vm_dst); cudaFree(h->vm_dst2); cudaFree(h->vm_vals); }
  for (void *p : h->allocs) cudaFree(p);
  ips_free(h);
  dist_free(h);
  if (h->res_host) cudaFreeHost(h->res_host);
  if (h->sqp_host) cudaFreeHost(h->sqp_host);
  if (h->status_host) cudaFreeHost(h->status_host);
  delete h;
  return HQPCU_OK;
}

int hqpcu_set_stream(hqpcu_handle *h, void *s) {
  if (!h) return HQPCU_E_NULL;
  h->stream = static_cast<cudaStream_t>(s);
  return HQPCU_OK;
}

long long hqpcu_launch_count(const hqpcu_handle *h) { return h ? h->launches : 0; }
int hqpcu_solve_stats(const hqpcu_handle *h, long long *solves, long long *steps) {
  if (!h || !solves || !steps) return HQPCU_E_NULL;
  *solves = h->n_solves;
  *steps = h->n_solve_steps;
  return HQPCU_OK;
}
int hqpcu_nseg(const hqpcu_handle *h);

// ------------------------------------------------------------------ update --
// new matrix values: a sequential-sweep fallback taken for the previous values
// (E_NOTPD at a segment end) ends here
static void undo_demotion(hqpcu_handle *h) {
  if (!h->demoted) return;
  h->demoted = false;
  choose_segments(h, h->nseg_req);
}

static int update_impl(hqpcu_handle *h, const double *Q, const double *fx, const double *fu,
                       const double *cv, const double *ev, cudaMemcpyKind kind) {
  if (!h || !Q || !fx || !fu || (h->d.nnz && !cv) || (h->q.nnz && !ev)) return HQPCU_E_NULL;
  undo_demotion(h);
  const LqDev &d = h->d;
  const size_t B = d.batch;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->Q, Q, B * (d.K + 1) * d.nm * d.nm * sizeof(double), kind, h->stream));
  CU(cudaMemcpyAsync(h->fx, fx, B * d.K * d.nx * d.nx * sizeof(double), kind, h->stream));
  CU(cudaMemcpyAsync(h->fu, fu, B * d.K * d.nx * d.nu * sizeof(double), kind, h->stream));
  if (d.nnz)
    CU(cudaMemcpyAsync(h->cval, cv, B * d.nnz * sizeof(double), kind, h->stream));
  if (h->q.nnz)
    CU(cudaMemcpyAsync(h->eqval, ev, (size_t)h->q.nnz * sizeof(double), kind, h->stream));
  h->factored = false;
  return HQPCU_OK;
}

int hqpcu_update(hqpcu_handle *h, const double *Q, const double *fx, const double *fu,
                 const double *ineq_val, const double *eq_val) {
  if (h && h->mg) {
    if (!Q || !fx || !fu || (h->d.m && !ineq_val)) return HQPCU_E_NULL;
    return mg_update(h, Q, fx, fu, ineq_val);
  }
  int rc = update_impl(h, Q, fx, fu, ineq_val, eq_val, cudaMemcpyHostToDevice);
  if (rc) return rc;
  h->eq_rowsum.assign(h->q.n_eq, 0.0);
  for (int i = 0; i < h->q.n_eq; i++)
    for (int e = h->dims_eq_ptr[i]; e < h->dims_eq_ptr[i + 1]; e++) h->eq_rowsum[i] += fabs(eq_val[e]);
  CU(cudaStreamSynchronize(h->stream));
  return HQPCU_OK;
}

// ---- update from sparse values (row f1) ---------------------------------------
__global__ void scatter_values_kernel(long long n, const double *__restrict__ vals,
                                      const long long *__restrict__ dst,
                                      const long long *__restrict__ dst2, double *Q, double *fx,
                                      double *fu, long long szQ, long long szX) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = vals[i];
    long long p = dst[i];
    for (int rep = 0; rep < 2; rep++) {
      if (p >= 0) {
        if (p < szQ) Q[p] = v;
        else if (p < szQ + szX) fx[p - szQ] = v;
        else fu[p - szQ - szX] = v;
      }
      p = dst2[i];
    }
  }
}

int hqpcu_set_value_map(hqpcu_handle *h, long long n, const long long *dst,
                        const long long *dst2) {
  if (!h || n < 0 || (n && (!dst || !dst2))) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  if (d.batch != 1) {
    g_err = "hqpcu_set_value_map: batch == 1";
    return HQPCU_E_UNSUPPORTED;
  }
  const long long szQ = (long long)(d.K + 1) * d.nm * d.nm, szX = (long long)d.K * d.nx * d.nx,
                  szU = (long long)d.K * d.nx * d.nu;
  for (long long i = 0; i < n; i++)
    if (dst[i] < 0 || dst[i] >= szQ + szX + szU || dst2[i] < -1 || dst2[i] >= szQ + szX + szU) {
      g_err = "hqpcu_set_value_map: target outside the stage slabs";
      return HQPCU_E_SIZES;
    }
  CU(cudaSetDevice(h->device));
  if (h->vm_dst) { cudaFree(h->vm_dst); cudaFree(h->vm_dst2); cudaFree(h->vm_vals); }
  h->vm_dst = h->vm_dst2 = nullptr;
  h->vm_vals = nullptr;
  h->vm_n = n;
  const size_t nn = (size_t)std::max<long long>(n, 1);
  CU(cudaMalloc(&h->vm_dst, nn * sizeof(long long)));
  CU(cudaMalloc(&h->vm_dst2, nn * sizeof(long long)));
  CU(cudaMalloc(&h->vm_vals, nn * sizeof(double)));
  CU(cudaMemcpy(h->vm_dst, dst, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->vm_dst2, dst2, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice));
  return HQPCU_OK;
}

int hqpcu_update_values(hqpcu_handle *h, const double *vals, const double *ineq_val,
                        const double *eq_val) {
  if (!h || !h->vm_dst || (h->vm_n && !vals) || (h->d.nnz && !ineq_val) || (h->q.nnz && !eq_val))
    return HQPCU_E_NULL;
  undo_demotion(h);
  const LqDev &d = h->d;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t szQ = (size_t)(d.K + 1) * d.nm * d.nm, szX = (size_t)d.K * d.nx * d.nx,
               szU = (size_t)d.K * d.nx * d.nu;
  CU(cudaMemcpyAsync(h->vm_vals, vals, (size_t)h->vm_n * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->Q, 0, szQ * sizeof(double), s));
  CU(cudaMemsetAsync(h->fx, 0, szX * sizeof(double), s));
  CU(cudaMemsetAsync(h->fu, 0, szU * sizeof(double), s));
  if (h->vm_n) {
    const int blocks = (int)std::min<long long>((h->vm_n + 255) / 256, 148 * 16);
    LAUNCH(h, scatter_values_kernel, <<<blocks, 256, 0, s>>>(h->vm_n, h->vm_vals, h->vm_dst, h->vm_dst2,
                                                            h->Q, h->fx, h->fu, (long long)szQ,
                                                            (long long)szX));
    CUL(h);
  }
  if (d.nnz)
    CU(cudaMemcpyAsync(h->cval, ineq_val, d.nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  if (h->q.nnz)
    CU(cudaMemcpyAsync(h->eqval, eq_val, (size_t)h->q.nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  h->eq_rowsum.assign(h->q.n_eq, 0.0);
  for (int i = 0; i < h->q.n_eq; i++)
    for (int e = h->dims_eq_ptr[i]; e < h->dims_eq_ptr[i + 1]; e++) h->eq_rowsum[i] += fabs(eq_val[e]);
  h->factored = false;
  CU(cudaStreamSynchronize(s));
  return HQPCU_OK;
}

int hqpcu_update_dev(hqpcu_handle *h, const double *Q, const double *fx, const double *fu,
                     const double *ineq_val, const double *eq_val) {
  return update_impl(h, Q, fx, fu, ineq_val, eq_val, cudaMemcpyDeviceToDevice);
}

// Update of a stage window from device memory: a host that produces its matrices
// on the device (or a long horizon generated chunk by chunk) never holds a second
// full copy.  nq blocks of Q for stages k0.., nf blocks of fx / fu for stages k0..
int hqpcu_update_stages_dev(hqpcu_handle *h, int k0, int nq, const double *Q, int nf,
                            const double *fx, const double *fu) {
  if (!h || (nq && !Q) || (nf && (!fx || !fu))) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  if (d.batch != 1 || k0 < 0 || nq < 0 || nf < 0 || k0 + nq > d.K + 1 || k0 + nf > d.K) {
    g_err = "hqpcu_update_stages_dev: stage window outside the horizon (batch == 1)";
    return HQPCU_E_SIZES;
  }
  undo_demotion(h);
  CU(cudaSetDevice(h->device));
  const size_t qq = (size_t)d.nm * d.nm, xx = (size_t)d.nx * d.nx, xu = (size_t)d.nx * d.nu;
  if (nq)
    CU(cudaMemcpyAsync(h->Q + (size_t)k0 * qq, Q, (size_t)nq * qq * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
  if (nf) {
    CU(cudaMemcpyAsync(h->fx + (size_t)k0 * xx, fx, (size_t)nf * xx * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->fu + (size_t)k0 * xu, fu, (size_t)nf * xu * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
  }
  h->factored = false;
  return HQPCU_OK;
}

int hqpcu_update_ineq_dev(hqpcu_handle *h, const double *ineq_val) {
  if (!h || (h->d.nnz && !ineq_val)) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  if (h->d.nnz)
    CU(cudaMemcpyAsync(h->cval, ineq_val, (size_t)h->d.batch * h->d.nnz * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->stream));
  h->factored = false;
  return HQPCU_OK;
}

// ------------------------------------------------------------------ factor --
static int launch_step_base(hqpcu_handle *h, const double *r1, const double *r2,
                            const double *r3, const double *r4, double *dx, double *dy,
                            double *dz, double *dw);

// extra solves + Schur complement of the general equality rows (lq_eq.cuh)
static int launch_eq_factor(hqpcu_handle *h) {
  const LqDev &d = h->d;
  const LqEq &q = h->q;
  cudaStream_t s = h->stream;
  const int ne = q.n_eq;
  // zero right-hand sides: t2..t4 double as scratch (overwritten by every residuum)
  CU(cudaMemsetAsync(h->t2, 0, (size_t)d.me * sizeof(double), s));
  if (d.m) {
    CU(cudaMemsetAsync(h->t3, 0, (size_t)d.m * sizeof(double), s));
    CU(cudaMemsetAsync(h->t4, 0, (size_t)d.m * sizeof(double), s));
  }
  for (int j = 0; j < ne; j++) {
    CU(cudaMemsetAsync(h->eq_r1, 0, (size_t)d.N * sizeof(double), s));
    LAUNCH(h, eq_unit_rhs_kernel, <<<1, 64, 0, s>>>(d, q, j, h->eq_r1));
    int rc = launch_step_base(h, h->eq_r1, h->t2, h->t3, h->t4, q.DX + (size_t)j * d.N,
                              q.DY + (size_t)j * d.me, q.DZ + (size_t)j * std::max(d.m, 1),
                              q.DW + (size_t)j * std::max(d.m, 1));
    if (rc) return rc;
  }
  LAUNCH(h, eq_schur_kernel, <<<(ne * ne + 127) / 128, 128, 0, s>>>(d, q));
  LAUNCH(h, eq_invert_kernel, <<<1, 128, (size_t)3 * ne * ne * sizeof(double), s>>>(d, q));
  CUL(h);
  return HQPCU_OK;
}

#define L_K1(NX_, NU_)                                                                         \
  do {                                                                                         \
    if ((NX_) == 0 && h->big)                                                                  \
      LAUNCHP(h, (seg_element_kernel<NX_, NU_, ((NX_) == 0 ? LQ_BIG_NT / 32 : 4)>), gseg, LQ_BIG_NT, h->smem_k1, s, d); \
    else if ((NX_) >= 32 && h->seg_warps == 8)                                                 \
      LAUNCHP(h, (seg_element_kernel<NX_, NU_, ((NX_) >= 32 ? 8 : 4)>), gseg, 256, h->smem_k1, s, d); \
    else if ((NX_) > 0 && h->seg_warps == 1)                                                   \
      LAUNCHP(h, (seg_element_kernel<NX_, NU_, ((NX_) > 0 ? 1 : 4)>), gseg, 32, h->smem_k1, s, d); \
    else                                                                                       \
      LAUNCHP(h, (seg_element_kernel<NX_, NU_, 4>), gseg, 128, h->smem_k1, s, d);        \
  } while (0)
#define L_K3(NX_, NU_)                                                                         \
  do {                                                                                         \
    if ((NX_) == 0 && h->big)                                                                  \
      LAUNCHP(h, (seg_riccati_kernel<NX_, NU_, ((NX_) == 0 ? LQ_BIG_NT / 32 : 4)>), gseg, LQ_BIG_NT, h->smem_k3, s, d); \
    else if ((NX_) >= 32 && h->seg_warps == 8)                                                 \
      LAUNCHP(h, (seg_riccati_kernel<NX_, NU_, ((NX_) >= 32 ? 8 : 4)>), gseg, 256, h->smem_k3, s, d); \
    else if ((NX_) > 0 && h->seg_warps == 1)                                                   \
      LAUNCHP(h, (seg_riccati_kernel<NX_, NU_, ((NX_) > 0 ? 1 : 4)>), gseg, 32, h->smem_k3, s, d); \
    else                                                                                       \
      LAUNCHP(h, (seg_riccati_kernel<NX_, NU_, 4>), gseg, 128, h->smem_k3, s, d);        \
  } while (0)
#define L_CMP(NX_) LAUNCHP(h, elem_compose_kernel<NX_>, gl, h->nt2(), h->smem_cmp, s, d, l)
#define L_HS(NX_) LAUNCHP(h, elem_hs_kernel<NX_>, gseg, h->nt2(), h->smem_cmp, s, d, stride, src, dst, last, jmax, jfix)
#define L_TOP(NX_) LAUNCHP(h, elem_scan_kernel<NX_>, dim3(1, d.batch), h->nt2(), h->smem_k2, s, d, h->ftop(), 1)
#define L_DWN(NX_) LAUNCHP(h, elem_scan_kernel<NX_>, gl, h->nt2(), h->smem_k2, s, d, l, 0)
#define L_PSI(NX_) LAUNCHP(h, psi_compose_kernel<NX_>, gl, h->nt2(), h->smem_psi, s, d, l, h->psi_chunk)

// factor, part 1: bound diagonal, segment elements, tree up-sweep
static int launch_factor_up(hqpcu_handle *h) {
  const LqDev &d = h->d;
  cudaStream_t s = h->stream;
  CU(cudaMemsetAsync(d.status, 0, sizeof(int), s));
  const dim3 gseg(d.P, d.batch);
  if (d.m) {
    const size_t tot = (size_t)d.batch * d.N;
    const int blocks = (int)std::min<size_t>((tot + 255) / 256, 148 * 8);
    LAUNCHP(h, hdiag_kernel, blocks, 256, 0, s, d);
  }
  if (d.P > 1 || h->ranged()) {
    LQ_DISPATCH_NXNU(d.nx, d.nu, L_K1);
    if (d.hs && !h->ranged()) {
      // one sweep: ceil(log2 (P+1)) levels of P concurrent combines (elem_hs_kernel)
      LAUNCHP(h, elem_terminal_kernel, d.batch, 128, 0, s, d);
      const int jmax = d.P, jfix = -1;
      int lev = 0;
      for (int stride = 1; stride <= d.P; stride <<= 1, lev++) {
        const int src = (lev & 1) * (d.P + 1), dst = ((lev + 1) & 1) * (d.P + 1);
        const int last = (stride << 1) > d.P ? 1 : 0;
        LQ_DISPATCH_NX(d.nx, d.nu, L_HS);
      }
    } else if (d.hs) {
      // stage range of a split horizon: the scan without the terminal element (it is
      // only known after the exchange); suffix 0 = the element of the whole range
      const int jmax = d.P - 1, jfix = -1, last = 0;
      int lev = 0;
      for (int stride = 1; stride < d.P; stride <<= 1, lev++) {
        const int src = (lev & 1) * (d.P + 1), dst = ((lev + 1) & 1) * (d.P + 1);
        LQ_DISPATCH_NX(d.nx, d.nu, L_HS);
      }
      h->hs_final = (lev & 1) * (d.P + 1);
    } else {
      for (int l = 0; l < h->ftop(); l++) {
        const dim3 gl(d.ft.cnt[l + 1], d.batch);
        LQ_DISPATCH_NX(d.nx, d.nu, L_CMP);
      }
    }
  }
  CUL(h);
  return HQPCU_OK;
}

// factor, part 2: value Hessians back down the tree, Riccati inside segments
static int launch_factor_down(hqpcu_handle *h) {
  const LqDev &d = h->d;
  cudaStream_t s = h->stream;
  const dim3 gseg(d.P, d.batch);
  if (d.hs && h->ranged()) {
    // the value behind this range is known now: terminal element into slot P, then
    // ONE level applies it to every suffix -> all segment end values
    LAUNCHP(h, elem_terminal_kernel, d.batch, 128, 0, s, d);
    const int stride = 0, src = h->hs_final, dst = h->hs_final, last = 1, jmax = d.P, jfix = d.P;
    LQ_DISPATCH_NX(d.nx, d.nu, L_HS);
  }
  if (!d.hs) {  // (suffix-scan mode: every segment's end value is known already)
    LQ_DISPATCH_NX(d.nx, d.nu, L_TOP);
    for (int l = h->ftop() - 1; l >= 0; l--) {
      const dim3 gl(d.ft.cnt[l + 1], d.batch);
      LQ_DISPATCH_NX(d.nx, d.nu, L_DWN);
    }
  }
  LQ_DISPATCH_NXNU(d.nx, d.nu, L_K3);
  for (int l = 0; l < h->stop(); l++) {
    const dim3 gl(d.st.cnt[l + 1], d.batch);
    LQ_DISPATCH_NX(d.nx, d.nu, L_PSI);
  }
  if (!d.fixed_x0 && !d.has_prev)
    LAUNCH(h, x0_factor_kernel,
           <<<d.batch, 32, h->big ? 0 : pad2((size_t)d.nx * d.nx) * sizeof(double), s>>>(d));
  CUL(h);
  return HQPCU_OK;
}
#undef L_K1
#undef L_K3
#undef L_CMP
#undef L_HS
#undef L_TOP
#undef L_DWN
#undef L_PSI

static int launch_factor(hqpcu_handle *h) {
  if (dist_on(h)) return dist_launch_factor(h);
  if (h->ranged()) {
    g_err = "this handle is a stage range of a split horizon: use hqpcu_range_*";
    return HQPCU_E_UNSUPPORTED;
  }
  int rc = run_graphed(h, {(const void *)(uintptr_t)1}, [&]() {
    int r = launch_factor_up(h);
    if (!r) r = launch_factor_down(h);
    if (!r && h->q.n_eq) {
      h->factored = true;  // the extra solves below go through the step launchers
      r = launch_eq_factor(h);
    }
    return r;
  });
  if (!rc) h->factored = true;
  return rc;
}

static int read_status(hqpcu_handle *h) {
  CU(cudaMemcpyAsync(h->status_host, h->d.status, sizeof(int), cudaMemcpyDeviceToHost,
                     h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const int st = *h->status_host;
  if (st & LQ_FLAG_SING) return HQPCU_E_SING;
  if (st & LQ_FLAG_NOTPD) return HQPCU_E_NOTPD;
  return HQPCU_OK;
}

static int factor_impl(hqpcu_handle *h, const double *z, const double *w, cudaMemcpyKind kind) {
  if (!h) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  if (d.m && (!z || !w)) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  if (d.m) {
    const size_t bytes = (size_t)d.batch * d.m * sizeof(double);
    CU(cudaMemcpyAsync(h->z, z, bytes, kind, h->stream));
    CU(cudaMemcpyAsync(h->w, w, bytes, kind, h->stream));
  }
  return launch_factor(h);
}

int hqpcu_factor_dev(hqpcu_handle *h, const double *z, const double *w) {
  return factor_impl(h, z, w, cudaMemcpyDeviceToDevice);
}

// Synchronise on a factor that was just enqueued and act on its status.  The
// zero-terminal-cost segment condensation needs Huu > 0 at segment ends: on
// E_NOTPD the factor is repeated as the sequential sweep ON THE GPU (still no CPU
// path).  The demotion holds for the matrix values that caused it -- the next
// hqpcu_update* restores the configured segment count (undo_demotion).
static int finish_factor(hqpcu_handle *h) {
  int rc = read_status(h);
  if (rc == HQPCU_E_NOTPD && h->d.P > 1 && !h->ranged()) {
    choose_segments(h, 1);
    h->demoted = true;
    rc = launch_factor(h);
    if (rc) return rc;
    rc = read_status(h);
  }
  // indefinite but non-singular: accepted by the sequential sweep like the
  // reference's BKP; a stage range of a split horizon has no sequential fallback
  if (rc == HQPCU_E_NOTPD && !h->ranged()) rc = HQPCU_OK;
  return rc;
}

int hqpcu_factor(hqpcu_handle *h, const double *z, const double *w) {
  if (h && h->mg) return mg_factor(h, z, w);
  int rc = factor_impl(h, z, w, cudaMemcpyHostToDevice);
  if (rc) return rc;
  return finish_factor(h);
}

// status of everything enqueued so far by the _dev entry points (synchronises)
int hqpcu_sync_status(hqpcu_handle *h) {
  if (!h) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  return read_status(h);
}

int hqpcu_set_nseg(hqpcu_handle *h, int nseg) {
  if (!h) return HQPCU_E_NULL;
  const int oldP = h->d.P;
  choose_segments(h, nseg);
  if (h->d.P > h->max_el) {
    choose_segments(h, oldP);
    g_err = "hqpcu_set_nseg: cannot exceed the segment count of hqpcu_create";
    return HQPCU_E_SIZES;
  }
  h->nseg_req = nseg;
  h->demoted = false;
  h->factored = false;
  return HQPCU_OK;
}

// chain kernels: one warp per chain for the compiled sizes, else the CTA version
static void launch_back(hqpcu_handle *h, int mode) {
  const LqDev &d = h->d;
  const dim3 gseg(d.P, d.batch);
  cudaStream_t s = h->stream;
#define L_(NX_) LAUNCHP(h, solve_back_kernel<NX_>, gseg, (NX_) ? 32 : h->thr_chain, h->smem_chain, s, d, mode, h->ring_chain)
  LQ_DISPATCH_NX(d.nx, d.nu, L_);
#undef L_
}
static void launch_fwd(hqpcu_handle *h, int mode) {
  const LqDev &d = h->d;
  const dim3 gseg(d.P, d.batch);
  cudaStream_t s = h->stream;
#define L_(NX_) LAUNCHP(h, solve_fwd_kernel<NX_>, gseg, (NX_) ? 32 : h->thr_chain, h->smem_chain, s, d, mode, h->ring_chain)
  LQ_DISPATCH_NX(d.nx, d.nu, L_);
#undef L_
}
static void launch_scan(hqpcu_handle *h, bool back, int groups, int lev, int phase,
                        const double *r2) {
  const LqDev &d = h->d;
  const dim3 g(groups, d.batch);
  cudaStream_t s = h->stream;
#define L_(NX_)                                                                               \
  do {                                                                                        \
    if (back)                                                                                 \
      LAUNCHP(h, (solve_scan_kernel<true, NX_>), g, (NX_) ? 32 : h->thr_chain, h->smem_scan, s, d, lev, phase, r2, h->ring_scan); \
    else                                                                                      \
      LAUNCHP(h, (solve_scan_kernel<false, NX_>), g, (NX_) ? 32 : h->thr_chain, h->smem_scan, s, d, lev, phase, r2, h->ring_scan); \
  } while (0)
  LQ_DISPATCH_NX(d.nx, d.nu, L_);
#undef L_
}

// -------------------------------------------------------------------- step --
// solve, part 1: stage-parallel prologue, zero-boundary backward chains, up-sweep
static int launch_step_a(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                         const double *r4) {
  const LqDev &d = h->d;
  if (!h->factored) {
    g_err = "step before factor";
    return HQPCU_E_NULL;
  }
  const int spb = d.spw * LQ_WPB * (32 / d.lgw);  // stages per CTA
  const dim3 gall((d.K + 1 + spb - 1) / spb, d.batch), gseg(d.P, d.batch);
  const size_t sv = (size_t)LQ_WPB * (32 / d.lgw) * (d.nm + d.nx) * sizeof(double);
  cudaStream_t s = h->stream;
  if (d.lgw == 16)
    LAUNCHP(h, (solve_pre_kernel<2, 16>), gall, 32 * LQ_WPB, sv, s, d, r1, r2, r3, r4);
  else if (d.spw == 1)
    LAUNCHP(h, (solve_pre_kernel<1, 32>), gall, 32 * LQ_WPB, sv, s, d, r1, r2, r3, r4);
  else
    LAUNCHP(h, (solve_pre_kernel<2, 32>), gall, 32 * LQ_WPB, sv, s, d, r1, r2, r3, r4);
  // a single segment that is the whole horizon starts from known boundary
  // values: no zero-boundary pass
  if (d.P > 1 || h->ranged()) launch_back(h, 0);
  for (int l = 0; l < h->stop(); l++)
    launch_scan(h, true, d.st.cnt[l + 1], l, 0, r2);
  CUL(h);
  return HQPCU_OK;
}

// solve, part 2: backward boundary values down the tree, true backward chains,
// stage-parallel middle pass, zero-boundary forward chains, up-sweep
static int launch_step_b(hqpcu_handle *h, const double *r2) {
  const LqDev &d = h->d;
  const int spb = d.spw * LQ_WPB * (32 / d.lgw);
  const dim3 gk((d.K + spb - 1) / spb, d.batch), gseg(d.P, d.batch);
  const size_t sv = (size_t)LQ_WPB * (32 / d.lgw) * (d.nx + d.nu + d.nu * d.nu) * sizeof(double);
  cudaStream_t s = h->stream;
  launch_scan(h, true, 1, h->stop(), 1, r2);
  for (int l = h->stop() - 1; l >= 0; l--)
    launch_scan(h, true, d.st.cnt[l + 1], l, 2, r2);
  launch_back(h, 1);
  if (d.lgw == 16)
    LAUNCHP(h, (solve_mid_kernel<2, 16>), gk, 32 * LQ_WPB, sv, s, d, r2);
  else if (d.spw == 1)
    LAUNCHP(h, (solve_mid_kernel<1, 32>), gk, 32 * LQ_WPB, sv, s, d, r2);
  else
    LAUNCHP(h, (solve_mid_kernel<2, 32>), gk, 32 * LQ_WPB, sv, s, d, r2);
  if (d.P > 1 || h->ranged()) launch_fwd(h, 0);
  for (int l = 0; l < h->stop(); l++)
    launch_scan(h, false, d.st.cnt[l + 1], l, 0, r2);
  CUL(h);
  return HQPCU_OK;
}

// solve, part 3: forward boundary values down the tree, true forward chains,
// stage-parallel epilogue
static int launch_step_c(hqpcu_handle *h, const double *r2, const double *r3, const double *r4,
                         double *dx, double *dy, double *dz, double *dw) {
  const LqDev &d = h->d;
  const int spb = d.spw * LQ_WPB * (32 / d.lgw);  // stages per CTA
  const dim3 gall((d.K + 1 + spb - 1) / spb, d.batch), gseg(d.P, d.batch);
  const size_t sv = (size_t)LQ_WPB * (32 / d.lgw) * (d.nm + d.nx) * sizeof(double);
  cudaStream_t s = h->stream;
  launch_scan(h, false, 1, h->stop(), 1, r2);
  for (int l = h->stop() - 1; l >= 0; l--)
    launch_scan(h, false, d.st.cnt[l + 1], l, 2, r2);
  launch_fwd(h, 1);
  if (d.lgw == 16)
    LAUNCHP(h, (solve_post_kernel<2, 16>), gall, 32 * LQ_WPB, sv, s, d, r3, r4, dx, dy, dz, dw);
  else if (d.spw == 1)
    LAUNCHP(h, (solve_post_kernel<1, 32>), gall, 32 * LQ_WPB, sv, s, d, r3, r4, dx, dy, dz, dw);
  else
    LAUNCHP(h, (solve_post_kernel<2, 32>), gall, 32 * LQ_WPB, sv, s, d, r3, r4, dx, dy, dz, dw);
  CUL(h);
  return HQPCU_OK;
}

static int launch_step_base(hqpcu_handle *h, const double *r1, const double *r2,
                            const double *r3, const double *r4, double *dx, double *dy,
                            double *dz, double *dw) {
  if (h->ranged()) {
    g_err = "this handle is a stage range of a split horizon: use hqpcu_range_*";
    return HQPCU_E_UNSUPPORTED;
  }
  int rc = launch_step_a(h, r1, r2, r3, r4);
  if (!rc) rc = launch_step_b(h, r2);
  if (!rc) rc = launch_step_c(h, r2, r3, r4, dx, dy, dz, dw);
  return rc;
}

static int launch_step_plain(hqpcu_handle *h, const double *r1, const double *r2,
                             const double *r3, const double *r4, double *dx, double *dy,
                             double *dz, double *dw) {
  int rc = launch_step_base(h, r1, r2, r3, r4, dx, dy, dz, dw);
  if (rc || !h->q.n_eq) return rc;
  const LqDev &d = h->d;
  cudaStream_t s = h->stream;
  LAUNCH(h, eq_multiplier_kernel, <<<1, 64, (size_t)(h->q.n_eq + 2) * sizeof(double), s>>>(d, h->q, r2, dx));
  const int blocks = (int)std::min<size_t>(((size_t)d.N + 255) / 256, 148 * 8);
  LAUNCH(h, eq_combine_kernel, <<<blocks, 256, 0, s>>>(d, h->q, dx, dy, dz, dw));
  CUL(h);
  return HQPCU_OK;
}

static int launch_step(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                       const double *r4, double *dx, double *dy, double *dz, double *dw) {
  if (!h->factored) {
    g_err = "step before factor";
    return HQPCU_E_NULL;
  }
  if (dist_on(h)) return dist_launch_step(h, r1, r2, r3, r4, dx, dy, dz, dw);
  return run_graphed(h, {(const void *)(uintptr_t)2, r1, r2, r3, r4, dx, dy, dz, dw}, [&]() {
    return launch_step_plain(h, r1, r2, r3, r4, dx, dy, dz, dw);
  });
}

int hqpcu_step_dev(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                   const double *r4, double *dx, double *dy, double *dz, double *dw) {
  if (!h || !r1 || !r2 || !dx || !dy) return HQPCU_E_NULL;
  if (h->d.m && (!r3 || !r4 || !dz || !dw)) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  return launch_step(h, r1, r2, r3, r4, dx, dy, dz, dw);
}

static int h2d_rhs(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                   const double *r4) {
  const LqDev &d = h->d;
  const size_t B = d.batch;
  CU(cudaMemcpyAsync(h->u_r1, r1, B * d.N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->u_r2, r2, B * d.me * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (d.m) {
    CU(cudaMemcpyAsync(h->u_r3, r3, B * d.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->u_r4, r4, B * d.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  return HQPCU_OK;
}

static int d2h_sol(hqpcu_handle *h, double *dx, double *dy, double *dz, double *dw) {
  const LqDev &d = h->d;
  const size_t B = d.batch;
  CU(cudaMemcpyAsync(dx, h->u_dx, B * d.N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(dy, h->u_dy, B * d.me * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (d.m) {
    CU(cudaMemcpyAsync(dz, h->u_dz, B * d.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(dw, h->u_dw, B * d.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return HQPCU_OK;
}

int hqpcu_step(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
               const double *r4, double *dx, double *dy, double *dz, double *dw) {
  if (!h || !r1 || !r2 || !dx || !dy) return HQPCU_E_NULL;
  if (h->d.m && (!r3 || !r4 || !dz || !dw)) return HQPCU_E_NULL;
  if (h->mg) return mg_apply(h, 0, 0.0, r1, r2, r3, r4, dx, dy, dz, dw, nullptr, nullptr);
  CU(cudaSetDevice(h->device));
  int rc = h2d_rhs(h, r1, r2, r3, r4);
  if (rc) return rc;
  rc = launch_step(h, h->u_r1, h->u_r2, h->u_r3, h->u_r4, h->u_dx, h->u_dy, h->u_dz, h->u_dw);
  if (rc) return rc;
  return d2h_sol(h, dx, dy, dz, dw);
}

#include "hqp_dist_host.inc"

// ---------------------------------------------------------------- residuum --
static int launch_residuum(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                           const double *r4, const double *dx, const double *dy,
                           const double *dz, const double *dw, bool keep, double *res) {
  const LqDev &d = h->d;
  if (dist_on(h)) {  // boundary blocks of the neighbouring ranges
    const int rc = dist_halo(h, dx, dy);
    if (rc) return rc;
  }
  CU(cudaMemsetAsync(h->res_dev, 0, sizeof(double), h->stream));
  const dim3 gall((d.K + LQ_RES_WPB) / LQ_RES_WPB, d.batch);
  const size_t sv = (size_t)LQ_RES_WPB * (d.nm + d.nx) * sizeof(double);
  if (h->q.n_eq)
    LAUNCH(h, eq_residuum_kernel, <<<1, 256, 0, h->stream>>>(d, h->q, r2, dx, dy,
                                                              keep ? h->t2 : nullptr, h->res_dev));
  LAUNCH(h, residuum_kernel, <<<gall, 32 * LQ_RES_WPB, sv, h->stream>>>(
      d, r1, r2, r3, r4, dx, dy, dz, dw, keep ? h->t1 : nullptr, keep ? h->t2 : nullptr,
      keep ? h->t3 : nullptr, keep ? h->t4 : nullptr, h->res_dev,
      h->q.n_eq ? h->q.ety : nullptr));
  CUL(h);
  if (dist_on(h)) {  // max over the ranges (on the bit pattern: a NaN stays on top)
    const int rc = dist_allreduce_bits_max(h, h->res_dev, 1);
    if (rc) return rc;
  }
  CU(cudaMemcpyAsync(h->res_host, h->res_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  // (the status word of the factor rides along: callers that enqueue factor + solve
  //  back to back read both with this one synchronisation, factor_solve)
  CU(cudaMemcpyAsync(h->status_host, h->d.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *res = *h->res_host;
  h->status_seen = *h->status_host;
  return HQPCU_OK;
}

int hqpcu_residuum_dev(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                       const double *r4, const double *dx, const double *dy, const double *dz,
                       const double *dw, double *res) {
  if (!h || !res) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  return launch_residuum(h, r1, r2, r3, r4, dx, dy, dz, dw, false, res);
}

int hqpcu_residuum(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                   const double *r4, const double *dx, const double *dy, const double *dz,
                   const double *dw, double *res) {
  if (!h || !res) return HQPCU_E_NULL;
  if (h->mg)
    return mg_apply(h, 2, 0.0, r1, r2, r3, r4, const_cast<double *>(dx), const_cast<double *>(dy),
                    const_cast<double *>(dz), const_cast<double *>(dw), res, nullptr);
  CU(cudaSetDevice(h->device));
  const LqDev &d = h->d;
  const size_t B = d.batch;
  int rc = h2d_rhs(h, r1, r2, r3, r4);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->u_dx, dx, B * d.N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->u_dy, dy, B * d.me * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (d.m) {
    CU(cudaMemcpyAsync(h->u_dz, dz, B * d.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->u_dw, dw, B * d.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  return launch_residuum(h, h->u_r1, h->u_r2, h->u_r3, h->u_r4, h->u_dx, h->u_dy, h->u_dz,
                         h->u_dw, false, res);
}

// one KKT solve with the current factor: the single-GPU launch sequence or, for a
// stage range of a split horizon, the one with the NCCL exchanges in it
static int step_any(hqpcu_handle *h, const double *r1, const double *r2, const double *r3,
                    const double *r4, double *dx, double *dy, double *dz, double *dw) {
  return launch_step(h, r1, r2, r3, r4, dx, dy, dz, dw);
}

// ------------------------------------------------------------------- solve --
// Hqp_IpMatrix::solve (hqp/Hqp_IpMatrix.C:65-128) on device-resident vectors.
static int solve_core(hqpcu_handle *h, double eps, const double *r1, const double *r2,
                      const double *r3, const double *r4, double *dx, double *dy, double *dz,
                      double *dw, double *res_out, int *nsteps) {
  const LqDev &d = h->d;
  const size_t B = d.batch;
  const size_t n1 = B * d.N, n2 = B * d.me, n3 = B * d.m;
  int steps = 1;
  int rc = step_any(h, r1, r2, r3, r4, dx, dy, dz, dw);
  if (rc) return rc;
  double res = 0.0;
  rc = launch_residuum(h, r1, r2, r3, r4, dx, dy, dz, dw, true, &res);
  if (rc) return rc;
  const int ablocks = (int)std::min<size_t>((std::max(n1, n2) + 255) / 256, 148 * 8);
  for (int it = 0; it < 5 && res > eps; it++) {
    const double res_last = res;
    rc = step_any(h, h->t1, h->t2, h->t3, h->t4, h->e1, h->e2, h->e3, h->e4);
    if (rc) return rc;
    steps++;
    double alpha = 1.0;
    do {
      LAUNCH(h, axpy4_kernel, <<<ablocks, 256, 0, h->stream>>>(alpha, h->e1, dx, n1, h->e2, dy, n2,
                                                              h->e3, dz, n3, h->e4, dw, n3));
      rc = launch_residuum(h, r1, r2, r3, r4, dx, dy, dz, dw, true, &res);
      if (rc) return rc;
      if (res > res_last) {
        LAUNCH(h, axpy4_kernel, <<<ablocks, 256, 0, h->stream>>>(-alpha, h->e1, dx, n1, h->e2, dy,
                                                                n2, h->e3, dz, n3, h->e4, dw, n3));
        alpha -= 0.3;
      }
    } while (res > res_last && alpha > 0.0);
    if (alpha <= 0.0) break;
  }
  if (res_out) *res_out = res;
  if (nsteps) *nsteps = steps;
  h->n_solves++;
  h->n_solve_steps += steps;
  return HQPCU_OK;
}

int hqpcu_solve_dev(hqpcu_handle *h, double eps, const double *r1, const double *r2,
                    const double *r3, const double *r4, double *dx, double *dy, double *dz,
                    double *dw, double *res, int *nsteps) {
  if (!h || !r1 || !r2 || !dx || !dy) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  return solve_core(h, eps, r1, r2, r3, r4, dx, dy, dz, dw, res, nsteps);
}

int hqpcu_solve(hqpcu_handle *h, double eps, const double *r1, const double *r2,
                const double *r3, const double *r4, double *dx, double *dy, double *dz,
                double *dw, double *res, int *nsteps) {
  if (!h || !r1 || !r2 || !dx || !dy) return HQPCU_E_NULL;
  if (h->d.m && (!r3 || !r4 || !dz || !dw)) return HQPCU_E_NULL;
  if (h->mg) return mg_apply(h, 1, eps, r1, r2, r3, r4, dx, dy, dz, dw, res, nsteps);
  CU(cudaSetDevice(h->device));
  int rc = h2d_rhs(h, r1, r2, r3, r4);
  if (rc) return rc;
  rc = solve_core(h, eps, h->u_r1, h->u_r2, h->u_r3, h->u_r4, h->u_dx, h->u_dy, h->u_dz,
                  h->u_dw, res, nsteps);
  if (rc) return rc;
  return d2h_sol(h, dx, dy, dz, dw);
}

// Factor with (z, w) already on the device, then one refined solve; the factor's
// status is read with the solve's first residual -- one synchronisation for both
// (the IP loops call this once per iteration).  Rare paths: a singular block
// returns HQPCU_E_SING (the solve ran on garbage, harmlessly); a non-positive pivot
// at a segment end repeats factor and solve as the sequential sweep.
static int factor_solve(hqpcu_handle *h, const double *z, const double *w, double eps,
                        const double *r1, const double *r2, const double *r3, const double *r4,
                        double *dx, double *dy, double *dz, double *dw, double *res) {
  int rc = factor_impl(h, z, w, cudaMemcpyDeviceToDevice);
  if (rc) return rc;
  rc = solve_core(h, eps, r1, r2, r3, r4, dx, dy, dz, dw, res, nullptr);
  if (rc) return rc;
  const int st = h->status_seen;
  if (st & LQ_FLAG_SING) return HQPCU_E_SING;
  if (st & LQ_FLAG_NOTPD) {
    if (h->ranged()) return HQPCU_E_NOTPD;
    if (h->d.P > 1) {
      choose_segments(h, 1);
      h->demoted = true;
      if ((rc = launch_factor(h))) return rc;
      rc = read_status(h);
      if (rc == HQPCU_E_NOTPD) rc = HQPCU_OK;
      if (rc) return rc;
      return solve_core(h, eps, r1, r2, r3, r4, dx, dy, dz, dw, res, nullptr);
    }
  }
  return HQPCU_OK;
}

// ------------------------------------------------------ horizon split (8e) --
int hqpcu_range_config(hqpcu_handle *h, int has_prev, int has_next) {
  if (!h) return HQPCU_E_NULL;
  if (h->d.batch != 1 || h->q.n_eq) {
    g_err = "hqpcu_range_config: batch == 1 and no general equality rows";
    return HQPCU_E_UNSUPPORTED;
  }
  if (has_prev && h->d.fixed_x0) {
    g_err = "hqpcu_range_config: only the first range may fix x0";
    return HQPCU_E_SIZES;
  }
  drop_graphs(h);  // the launch sequences depend on the range flags
  h->d.has_prev = has_prev ? 1 : 0;
  h->d.has_next = has_next ? 1 : 0;
  h->d.hs = (h->use_hs && (h->d.P > 1 || h->ranged())) ? 1 : 0;
  h->factored = false;
  return HQPCU_OK;
}

int hqpcu_range_factor_begin(hqpcu_handle *h, const double *z, const double *w, double *xf) {
  if (!h || !xf) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  CU(cudaSetDevice(h->device));
  if (d.m) {
    const size_t bytes = (size_t)d.m * sizeof(double);
    CU(cudaMemcpyAsync(h->z, z, bytes, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->w, w, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  return run_graphed(h, {(const void *)(uintptr_t)10, xf}, [&]() {
    int rc = launch_factor_up(h);
    if (rc) return rc;
    LAUNCH(h, range_export_factor_kernel, <<<1, 128, 0, h->stream>>>(h->d, xf, h->range_slot()));
    CUL(h);
    return (int)HQPCU_OK;
  });
}

int hqpcu_range_factor_finish(hqpcu_handle *h, const double *gathered, int rank, int world,
                              double *xpsi) {
  if (!h || !gathered || !xpsi) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  CU(cudaSetDevice(h->device));
  const int rc = run_graphed(
      h, {(const void *)(uintptr_t)11, gathered, xpsi, (const void *)(uintptr_t)rank,
          (const void *)(uintptr_t)world},
      [&]() {
        cudaStream_t s = h->stream;
#define L_RS(NX_) LAUNCH(h, range_scan_factor_kernel<NX_>, <<<1, h->nt2(), h->smem_k2, s>>>(d, gathered, rank, world))
        LQ_DISPATCH_NX(d.nx, d.nu, L_RS);
#undef L_RS
        int rc = launch_factor_down(h);
        if (rc) return rc;
        const size_t n2 = (size_t)d.nx * d.nx;
        CU(cudaMemcpyAsync(xpsi, d.segPsi + (size_t)d.st.off[d.st.nlev - 1] * n2,
                           n2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return (int)HQPCU_OK;
      });
  if (!rc) h->factored = true;  // (not inside the body: a graph replay does not run it)
  return rc;
}

int hqpcu_range_step_begin(hqpcu_handle *h, const double *r1, const double *r2,
                           const double *r3, const double *r4, double *xv) {
  if (!h || !r1 || !r2 || !xv) return HQPCU_E_NULL;
  if (!h->factored) {
    g_err = "step before factor";
    return HQPCU_E_NULL;
  }
  CU(cudaSetDevice(h->device));
  h->rg_r1 = r1; h->rg_r2 = r2; h->rg_r3 = r3; h->rg_r4 = r4;
  return run_graphed(h, {(const void *)(uintptr_t)12, r1, r2, r3, r4, xv}, [&]() {
    int rc = launch_step_a(h, r1, r2, r3, r4);
    if (rc) return rc;
    LAUNCH(h, range_export_vec_kernel<true>,
           <<<1, 64, (size_t)(h->d.nx + 2) * sizeof(double), h->stream>>>(h->d, r2, xv));
    CUL(h);
    return (int)HQPCU_OK;
  });
}

int hqpcu_range_step_mid(hqpcu_handle *h, const double *gv, const double *gpsi, int rank,
                         int world, double *xx) {
  if (!h || !gv || !gpsi || !xx) return HQPCU_E_NULL;
  if (!h->factored) {
    g_err = "step before factor";
    return HQPCU_E_NULL;
  }
  CU(cudaSetDevice(h->device));
  const int thr = std::max(64, ((h->d.nx + 31) / 32) * 32);
  const size_t sm = (size_t)(2 * h->d.nx + 2) * sizeof(double);
  return run_graphed(
      h, {(const void *)(uintptr_t)13, gv, gpsi, xx, h->rg_r2, (const void *)(uintptr_t)rank,
          (const void *)(uintptr_t)world},
      [&]() {
        LAUNCH(h, range_scan_vec_kernel<true>, <<<1, thr, sm, h->stream>>>(h->d, gv, gpsi, rank, world, h->d.nx * h->d.nx));
        int rc = launch_step_b(h, h->rg_r2);
        if (rc) return rc;
        LAUNCH(h, range_export_vec_kernel<false>,
               <<<1, 64, (size_t)(h->d.nx + 2) * sizeof(double), h->stream>>>(h->d, h->rg_r2, xx));
        CUL(h);
        return (int)HQPCU_OK;
      });
}

int hqpcu_range_step_finish(hqpcu_handle *h, const double *gx, const double *gpsi, int rank,
                            int world, double *dx, double *dy, double *dz, double *dw) {
  if (!h || !gx || !gpsi || !dx || !dy) return HQPCU_E_NULL;
  if (!h->factored) {
    g_err = "step before factor";
    return HQPCU_E_NULL;
  }
  CU(cudaSetDevice(h->device));
  const int thr = std::max(64, ((h->d.nx + 31) / 32) * 32);
  const size_t sm = (size_t)(2 * h->d.nx + 2) * sizeof(double);
  return run_graphed(
      h, {(const void *)(uintptr_t)14, gx, gpsi, dx, dy, dz, dw, h->rg_r2, h->rg_r3, h->rg_r4,
          (const void *)(uintptr_t)rank, (const void *)(uintptr_t)world},
      [&]() {
        LAUNCH(h, range_scan_vec_kernel<false>, <<<1, thr, sm, h->stream>>>(h->d, gx, gpsi, rank, world, h->d.nx * h->d.nx));
        return launch_step_c(h, h->rg_r2, h->rg_r3, h->rg_r4, dx, dy, dz, dw);
      });
}

int hqpcu_profile(hqpcu_handle *h, int on) {
  if (!h) return HQPCU_E_NULL;
  h->profiling = on != 0;
  return HQPCU_OK;
}

// Synchronises, sums the CUDA-event time of every launch recorded since the
// last call and writes {"kernel": {"ms": total, "n": launches}, ...} to buf.
int hqpcu_profile_read(hqpcu_handle *h, char *buf, int len) {
  if (!h || !buf || len < 3) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int> cnt;
  for (auto &sp : h->spans) {
    float t = 0.f;
    cudaEventElapsedTime(&t, sp.e0, sp.e1);
    cudaEventDestroy(sp.e0);
    cudaEventDestroy(sp.e1);
    size_t i = 0;
    for (; i < names.size(); i++)
      if (names[i] == sp.name) break;
    if (i == names.size()) { names.push_back(sp.name); ms.push_back(0); cnt.push_back(0); }
    ms[i] += t;
    cnt[i]++;
  }
  h->spans.clear();
  std::string out = "{";
  for (size_t i = 0; i < names.size(); i++) {
    char tmp[160];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"n\": %d}", i ? ", " : "",
             names[i].c_str(), ms[i], cnt[i]);
    out += tmp;
  }
  out += "}";
  if ((int)out.size() + 1 > len) return HQPCU_E_SIZES;
  memcpy(buf, out.c_str(), out.size() + 1);
  return HQPCU_OK;
}

int hqpcu_debug_stamps(hqpcu_handle *h, long long *out16) {
  if (!h || !out16) return HQPCU_E_NULL;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(out16, h->d.dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
#ifdef LQ_TIMING
  CU(cudaMemcpyFromSymbol(out16 + 16, g_dbg, 16 * sizeof(long long)));  // caller passes 32
#endif
  return HQPCU_OK;
}

int hqpcu_get_factor(hqpcu_handle *h, double *Vxx, double *Rux) {
  if (!h) return HQPCU_E_NULL;
  const LqDev &d = h->d;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  if (Vxx)
    CU(cudaMemcpy(Vxx, d.V, (size_t)d.batch * (d.K + 1) * d.nx * d.nx * sizeof(double),
                  cudaMemcpyDeviceToHost));
  if (Rux)
    CU(cudaMemcpy(Rux, d.Rux, (size_t)d.batch * d.K * d.nu * d.nx * sizeof(double),
                  cudaMemcpyDeviceToHost));
  return HQPCU_OK;
}

}  // extern "C"

#include "hqp_ips_host.inc"
#include "hqp_franke_host.inc"
#include "hqp_sqp_host.inc"
extern "C" {
#include "hqp_mg_host.inc"
int hqpcu_nseg(const hqpcu_handle *h) {
  if (h && h->mg) return h->mg->w[0]->part ? h->mg->w[0]->part->d.P : 0;
  return h ? h->d.P : 0;
}
}
