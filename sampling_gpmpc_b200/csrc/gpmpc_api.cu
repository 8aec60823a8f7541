// libgpmpc_b200.so -- host side of the C ABI declared in include/gpmpc_b200.h.
// Plain CUDA runtime, no torch types: the Python host (sampling_gpmpc_b200/engine.py) binds this with ctypes
// and passes torch tensors' data_ptr() / the current torch stream handle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "gpmpc_assemble.cuh"
#include "gpmpc_block.cuh"
#include "gpmpc_post.cuh"
#include "gpmpc_step.cuh"
#include "gpmpc_block_mma.cuh"
#include "gpmpc_k0.cuh"
#include "gpmpc_horizon.cuh"
#include "gpmpc_eig.cuh"
#include "gpmpc_rng.cuh"

// static shared memory of k_sample_eig (rotation tables + coefficients), rounded up
#define EIG_STATIC_SMEM (44 * 1024)

namespace {
std::string g_create_error;

// cudaFuncSetAttribute is per DEVICE: remember (kernel, device) pairs, not one process-wide flag per kernel
std::set<std::pair<const void*, int>> g_func_configured;

// Every entry point runs on the handle's device whatever the caller's current device is (restored on return).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
}  // namespace
#define ON_HANDLE_DEVICE(h) DeviceGuard dev_guard_((h) ? (h)->device : -1)

struct gpmpc_handle {
  DevState st{};
  gpmpc_dims dims{};
  int device = 0;
  bool have_hypers = false, have_real = false;
  bool condition = true;
  std::string err;
  // host mirrors / bookkeeping
  std::vector<double> h_ls, h_os, h_noise;
  long long factor_version = 0;  // bumped whenever the factor changes
  long long cache_version = -1;  // factor_version the posterior cache was built against
  int cache_H = 0;
  int ws_n = 0, ws_q = 0, ws_H = 0;  // workspace capacity
  // Sigma* / mean w.r.t. the REAL data only, produced by the same pass as the full ones when the caller is about to reset the
  // hallucinated set (gpmpc_linearise with reset_first): the append after the reset then needs no second posterior pass
  double *S2buf = nullptr, *mu2buf = nullptr;
  bool want_real_only = false, real_cache_valid = false;
  // chol(Sigma_app + noise) factorised beside the draw for the append that follows (k_pm_finish's second CTA row -> st.Lpre):
  // requested by gpmpc_linearise; pre_kind 1 = against the current factor (valid while factor_version == pre_version),
  // 2 = against the real data alone (valid for the append right after a reset), 0 = none
  uint32_t* status_pinned = nullptr;  // pinned staging word of gpmpc_status
  bool prefactor_next = false;
  int pre_kind = 0, pre_H = 0;
  long long pre_version = -1;
  int real_cache_H = 0;
  long long launches = 0;
  double last_bytes = 0.0, last_flops = 0.0;
  // rollout scratch
  double *r_xu = nullptr, *r_xstar = nullptr, *r_y = nullptr;
  int r_ns = 0;
  unsigned char* d_active = nullptr;
  int d_active_cap = 0;
  // scratch of the consumers (gpmpc_traj_stats / gpmpc_stage_hulls), grown on demand
  void* c_scratch = nullptr;
  size_t c_scratch_bytes = 0;
  int max_dyn_smem = 0, num_sms = 148;
  // fused-horizon rollout (k_horizon, one launch per rollout).  It LOSES to the step-wise path at the bench shape (254.7 ms vs
  // 220.4 ms per rollout on the same box, profiles/r2_horizon_probe.txt) and WINS where the step-wise rollout is
  // launch-latency bound -- the reference's own sample counts: 1.0 vs 2.5 ms at 20 samples, 1.2 vs 2.4 ms at 200, 4.9 vs 5.8 ms
  // at 2000 (profiles/r2_small_batch_rollout.txt).  gpmpc_set_option("rollout_fused", v) / GPMPC_ROLLOUT_FUSED=v: 0 never
  // (default of the C ABI: a failed jitter ladder is then redone in-stream through the eigen root), 1 wherever the shape allows,
  // 2 automatic = only while every warp of the fused kernel gets at most HZ_AUTO_MAX_SERIAL samples.  GPMPC_HZ_GROUPS caps the
  // sample groups per CTA (L2 footprint = #SMs x groups x g_ny factors), GPMPC_HZ_STAGGER_NS spreads the groups' starts
  int fused_rollout = 0;
  int hz_groups_cap = 0;
  int step_grid_cap = 0;             // > 0: CTAs (= SMs) the step kernel may take; the rest stay free for a concurrent stream
  int sr_grid_cap = 0;               // > 0: CTAs (= SMs) the shared-rows GEMM may take (experiment: GEMM of one half beside the step of the other)
  int step_warps_cap = 0;            // > 0: warps per CTA of the step kernel (two such CTAs of two streams share an SM)
  long long hz_stagger_ns = -1;      // < 0: automatic (one sample-horizon of the previous fused rollout)
  double hz_last_ms = 0.0;           // device time of the previous fused launch (for the automatic stagger)
  int hz_last_samples_per_group = 0;
  cudaEvent_t hz_ev[2] = {nullptr, nullptr};
  bool last_rollout_fused = false;
  std::vector<cudaEvent_t> slice_ev;
  // grouped rollout (gpmpc_set_grouping): consecutive blocks of grp_size samples are one reference Agent each; the min-distance
  // filter of update_hallucinated_Dyn_dataset (src/agent.py:164-202) reduces its flags per group
  int grp_size = 0;
  double grp_min_dist = -1.0;
  unsigned char *grp_flags = nullptr, *grp_decision = nullptr;
  // scratch of gpmpc_linearise: [x, u] on the device and the gathered GP inputs
  double *lin_xu = nullptr, *lin_xg = nullptr;
  size_t lin_xu_count = 0, lin_xg_count = 0;
  int eig_epoch = 0;  // draw launches so far (gpmpc_eig.cuh: a failing element publishes the epoch of its launch)
  // large-m path: shared rows of all elements by one batched GEMM (k_shared_rows) whenever inv(L_oo) does not fit in
  // shared memory beside the warps (m > ~130) and m >= wo_min_m; GPMPC_WO_MIN_M overrides the threshold (tests use it
  // to reach the one-element-per-pass path through L2, which measured 28 % slower at m = 180, 7.8x slower at m = 1000)
  int wo_min_m = 1;
  bool force_big = false;            // tests: k_step_big (+ k-slab GEMM) for any m
  int big_slab_cap = 0;              // tests: cap on the rows of K per GEMM pass (several slabs at small m)
  bool force_block_fallback = false;  // tests: gpmpc_step through posterior + append even where k_step_big applies
  bool wo_slab_nb3 = true;  // k_step<WO> path, 1200 < m <= 3600: GEMM with 3 column blocks in k-slabs instead of 2 / 1 in one pass
  bool force_wo = false;  // experiment: the batched shared-rows GEMM also for small m (gpmpc_set_option "force_wo")
  int wo_max_nb = 3;  // GPMPC_WO_MAX_NB: cap on the column blocks per tile (tests reach the NB = 2 / 1 instantiations with it)
  // SQP-mode model call: tensor-core kernel k_posterior_mma (default) or the scalar substitution kernel k_posterior
  // (gpmpc_set_block_kernels / GPMPC_BLOCK_SCALAR=1: the independent reference semantics the parity tests compare with)
  bool block_mma = true;
  // a hallucinated point with only SOME of its T scalars in the factor exists (gpmpc_append_masked; agent.py:402): the
  // kernels that find a point's rows through hrow0 (K1, K2m) are bypassed for the scalar kernels until the next reset
  bool has_partial = false;
  int first_partial_point = -1;    // index of the first partially observed point (-1: none)
  std::vector<int> c_before_point; // factor rows in use before hallucinated point p (for gpmpc_truncate_hallucinated)
  size_t wo_count = 0;
  // optional per-launch timing of the fused step kernel inside gpmpc_rollout (CUDA events on its stream)
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  int ev_used = 0;
};

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                           \
      return GPMPC_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

static int fail(gpmpc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

// opt a kernel in to `bytes` of dynamic shared memory on the handle's device (once per kernel and device)
template <typename F>
static cudaError_t opt_in_smem(gpmpc_handle* h, F func, int bytes, bool prefer_shared = false) {
  const auto key = std::make_pair((const void*)func, h->device);
  if (g_func_configured.count(key)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && prefer_shared)
    e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) g_func_configured.insert(key);
  return e;
}

template <typename Tp>
static cudaError_t dev_alloc(Tp** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  return cudaMalloc((void**)p, count * sizeof(Tp));
}

static void free_factor_state(gpmpc_handle* h) {
  DevState& st = h->st;
  cudaFree(st.Xh); cudaFree(st.Yh); cudaFree(st.hobs_pt); cudaFree(st.hobs_task); cudaFree(st.hrow0);
  cudaFree(st.Lh); cudaFree(st.beta_h); cudaFree(st.pstate);
  st.Xh = st.Yh = st.Lh = st.beta_h = nullptr;
  st.pstate = nullptr;
  st.hobs_pt = st.hobs_task = st.hrow0 = nullptr;
}

// (re)allocates the per-element state for `cap_points`, keeping what is already stored
static int alloc_factor_state(gpmpc_handle* h, int cap_points, cudaStream_t stream) {
  DevState old = h->st;
  DevState& st = h->st;
  // the factor exists only while conditioning is on; a record-only handle keeps just the data set
  const int c_cap = h->condition ? cap_points * st.T : 0;
  // sub-panel layout (gpmpc_state.cuh): an element's rows are one contiguous stream whose prefix does not
  // depend on the capacity, so growing is a strided copy of the used prefixes
  const long long stride = c_cap ? (long long)subpanel_off((c_cap + 7) / 8, st.mo) : 0;
  double *Xh, *Yh, *Lh, *beta_h;
  int *hp, *ht, *hr;
  const size_t B = (size_t)st.B;
  const size_t lh_count = c_cap ? B * (size_t)stride : 1;
  CUDA_TRY(h, dev_alloc(&Xh, B * cap_points * st.d));
  CUDA_TRY(h, dev_alloc(&Yh, B * cap_points * st.T));
  CUDA_TRY(h, dev_alloc(&Lh, lh_count));
  CUDA_TRY(h, dev_alloc(&beta_h, B * c_cap));
  CUDA_TRY(h, dev_alloc(&hp, (size_t)c_cap));
  CUDA_TRY(h, dev_alloc(&ht, (size_t)c_cap));
  CUDA_TRY(h, dev_alloc(&hr, (size_t)cap_points));
  unsigned char* ps = nullptr;
  if (h->grp_size > 0) {
    CUDA_TRY(h, dev_alloc(&ps, B * (size_t)std::max(cap_points, 1)));
    CUDA_TRY(h, cudaMemsetAsync(ps, 0, B * (size_t)std::max(cap_points, 1), stream));
    if (old.pstate && old.np > 0)
      CUDA_TRY(h, cudaMemcpy2DAsync(ps, (size_t)cap_points, old.pstate, (size_t)old.cap_points, (size_t)old.np, B,
                                    cudaMemcpyDeviceToDevice, stream));
  }
  // padding columns [m, mo) must read as 0; rows not yet appended are never used but kept finite
  CUDA_TRY(h, cudaMemsetAsync(Lh, 0, lh_count * sizeof(double), stream));
  CUDA_TRY(h, cudaMemsetAsync(beta_h, 0, std::max<size_t>(1, B * c_cap) * sizeof(double), stream));
  if (old.Xh && old.np > 0) {
    CUDA_TRY(h, cudaMemcpy2DAsync(Xh, (size_t)cap_points * st.d * 8, old.Xh, (size_t)old.cap_points * st.d * 8,
                                  (size_t)old.np * st.d * 8, B, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(Yh, (size_t)cap_points * st.T * 8, old.Yh, (size_t)old.cap_points * st.T * 8,
                                  (size_t)old.np * st.T * 8, B, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(h, cudaMemcpyAsync(hr, old.hrow0, (size_t)old.np * 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (old.Lh && old.c > 0 && c_cap >= old.c) {
    const size_t used = subpanel_off((old.c + 7) / 8, st.mo) * 8;  // bytes of the sub-panels in use
    CUDA_TRY(h, cudaMemcpy2DAsync(Lh, (size_t)stride * 8, old.Lh, (size_t)old.elem_stride * 8, used, B,
                                  cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(h, cudaMemcpy2DAsync(beta_h, (size_t)c_cap * 8, old.beta_h, (size_t)old.c_cap * 8, (size_t)old.c * 8, B,
                                  cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(h, cudaMemcpyAsync(hp, old.hobs_pt, (size_t)old.c * 4, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(h, cudaMemcpyAsync(ht, old.hobs_task, (size_t)old.c * 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (old.Xh) {
    CUDA_TRY(h, cudaStreamSynchronize(stream));
    free_factor_state(h);
  }
  st.Xh = Xh; st.Yh = Yh; st.Lh = Lh; st.beta_h = beta_h; st.hobs_pt = hp; st.hobs_task = ht; st.hrow0 = hr;
  st.pstate = ps;
  st.cap_points = cap_points; st.c_cap = c_cap; st.elem_stride = stride;
  h->dims.cap_points = cap_points;
  return GPMPC_OK;
}

static int ensure_workspace(gpmpc_handle* h, int H) {
  DevState& st = h->st;
  const int n = st.m + st.c, q = H * st.T;
  if (n <= h->ws_n && q <= h->ws_q && H <= h->ws_H && st.W) return GPMPC_OK;
  // sized for the reserved capacity where one is known (gpmpc_reserve / Agent: H * max_sqp_iter), so that the SQP loop never
  // re-allocates (cudaFree / cudaMalloc synchronise the device: 10-20 ms spikes otherwise)
  const int new_n = std::max(std::max(n + n / 2, st.m + st.c_cap), h->ws_n), new_q = std::max(q, h->ws_q), new_H = std::max(H, h->ws_H);
  cudaFree(st.W); cudaFree(st.S); cudaFree(st.C); cudaFree(st.mu); cudaFree(st.xc); cudaFree(st.E); cudaFree(st.Lpre);
  cudaFree(h->S2buf); cudaFree(h->mu2buf);
  st.W = st.S = st.C = st.mu = st.xc = st.E = st.Lpre = nullptr;
  h->pre_kind = 0;
  h->S2buf = h->mu2buf = nullptr;
  h->real_cache_valid = false;
  const size_t B = (size_t)st.B;
  CUDA_TRY(h, dev_alloc(&st.W, B * new_n * (size_t)new_q));
  CUDA_TRY(h, dev_alloc(&st.S, B * (size_t)new_q * new_q));
  CUDA_TRY(h, dev_alloc(&st.C, B * (size_t)new_q * new_q));
  CUDA_TRY(h, dev_alloc(&st.Lpre, B * ((size_t)new_q * (new_q + 1) / 2 + 1)));
  CUDA_TRY(h, dev_alloc(&st.mu, B * (size_t)new_q));
  CUDA_TRY(h, dev_alloc(&h->S2buf, B * (size_t)new_q * new_q));
  CUDA_TRY(h, dev_alloc(&h->mu2buf, B * (size_t)new_q));
  CUDA_TRY(h, dev_alloc(&st.xc, B * (size_t)new_H * st.d));
  // iteration matrix of the eigen-root fallback, only where it does not fit in shared memory
  const bool e_global = (size_t)new_q * new_q * 8 > (size_t)(h->max_dyn_smem - EIG_STATIC_SMEM);
  CUDA_TRY(h, dev_alloc(&st.E, e_global ? B * (size_t)new_q * new_q : 1));
  h->ws_n = new_n; h->ws_q = new_q; h->ws_H = new_H;
  h->cache_version = -1;
  return GPMPC_OK;
}

// algorithmic work of one conditioning step per batch element (DESIGN.md "roofline" section)
static void count_work(gpmpc_handle* h, int H, bool append) {
  const DevState& st = h->st;
  const double m = st.m, c = st.c, q = (double)H * st.T, B = st.B, d = st.d, n = m + c;
  double bytes = 8.0 * (c * m + c * (c + 1) / 2.0)            // own factor rows, read once
                 + 8.0 * (H * d + 3.0 * q)                      // x in; mean, var, y out
                 + 8.0 * c / st.T * d;                          // hallucinated inputs for the kernel vector
  if (append) bytes += 8.0 * (q * n + q * (q + 1) / 2.0 + q);   // new rows + beta
  double flops = q * m * m + 2.0 * q * c * m + q * c * c + q * q * n + q * q * q / 3.0 + 2.0 * q * n +
                 q * (n + q) * (3.0 * d + 12.0);
  h->last_bytes = bytes * B;
  h->last_flops = flops * B;
}

// dynamic shared memory for the packed q x q Cholesky of the draw / append (0: does not fit, global-memory path)
static size_t tri_bytes(gpmpc_handle* h, int q) {
  const size_t b = (size_t)q * (q + 1) / 2 * sizeof(double);
  return b + 1024 <= (size_t)h->max_dyn_smem ? b : 0;
}

static int configure_block_smem(gpmpc_handle* h) {
  CUDA_TRY(h, opt_in_smem(h, k_sample, h->max_dyn_smem - 8192));
  CUDA_TRY(h, opt_in_smem(h, k_append, h->max_dyn_smem - 8192));
  return GPMPC_OK;
}

// The eigen-root redo of a draw, queued right behind it; returns at once on the device unless an element of THIS draw
// (st.eig_epoch) failed its jitter ladder.  `st` must be the DevState the draw kernel was launched with.
static int launch_sample_eig(gpmpc_handle* h, const DevState& st, int H, const double* eps, const gpmpc_sample_opts& o,
                             double* y, int* jl, cudaStream_t stream) {
  if (!eps || (o.flags & GPMPC_OPT_NO_EIG_FALLBACK) || H * st.T == 1) return GPMPC_OK;
  const int q = H * st.T;
  if (q > 2 * EIG_MAX_PAIRS) return GPMPC_OK;  // beyond the fallback's tables: the NotPD status stands
  const int dyn_max = h->max_dyn_smem - EIG_STATIC_SMEM;
  CUDA_TRY(h, opt_in_smem(h, k_sample_eig, dyn_max));
  const size_t a_bytes = (size_t)q * q * sizeof(double);
  const int in_smem = a_bytes <= (size_t)dyn_max ? 1 : 0;
  k_sample_eig<<<st.B, EIG_THREADS, in_smem ? a_bytes : 0, stream>>>(st, H, eps, o, y, jl, in_smem);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

template <int D, int T>
static int launch_posterior_mma(gpmpc_handle* h, const DevState& st, const double* x, int H, double* mean, double* var,
                                const double* eps, const gpmpc_sample_opts& o, double* y, int* jl, cudaStream_t stream) {
  auto solve = k_pm_solve<D, T>;
  auto gram = k_pm_gram<T>;
  const int q = H * st.T, QB = (q + 7) / 8;
  const size_t slab = (size_t)2 * PM_NSP * PM_SLABC * 8 * sizeof(double);
  CUDA_TRY(h, opt_in_smem(h, solve, (int)slab));
  CUDA_TRY(h, opt_in_smem(h, k_pm_finish, h->max_dyn_smem - 8192));
  // the column blocks of one element over ceil(QB / 4) CTAs; the Gram tiles over enough CTAs to fill the GPU
  dim3 gs(st.B, (QB + PM_WARPS - 1) / PM_WARPS);
  solve<<<gs, PM_WARPS * 32, slab, stream>>>(st, x, H);
  const int tiles = QB * (QB + 1) / 2 + QB;
  const int want = std::max(1, std::min((tiles + PM_WARPS - 1) / PM_WARPS, (4 * h->num_sms + st.B - 1) / st.B));
  dim3 gg(st.B, want);
  gram<<<gg, PM_WARPS * 32, 0, stream>>>(st, x, H);
  h->launches += 1;
  if (!mean && !var && !eps) return GPMPC_OK;  // W, Sigma*, mean stay in the workspace (recompute before an append)
  size_t tri = eps ? tri_bytes(h, q) : 0;
  if (tri + 8192 > (size_t)h->max_dyn_smem) tri = 0;
  // the append's own Cholesky beside the draw's (second CTA row), when the caller announced the append (gpmpc_linearise)
  const int pre = (h->prefactor_next && tri && q >= 2 && st.Lpre) ? (st.S2 ? 2 : 1) : 0;
  h->prefactor_next = false;
  h->pre_kind = pre;
  h->pre_H = H;
  h->pre_version = h->factor_version;
  k_pm_finish<<<dim3(st.B, pre ? 2 : 1), BLK_THREADS, tri, stream>>>(st, H, mean, var, eps, o, y, jl, tri ? 1 : 0, pre);
  h->launches += 1;
  return launch_sample_eig(h, st, H, eps, o, y, jl, stream);
}

static int dispatch_posterior_mma(gpmpc_handle* h, const DevState& st, const double* x, int H, double* mean,
                                  double* var, const double* eps, const gpmpc_sample_opts& o, double* y, int* jl,
                                  cudaStream_t stream) {
#define PM_CASE(D_)                                                                                       \
  case D_:                                                                                                \
    return st.T == 1 ? launch_posterior_mma<D_, 1>(h, st, x, H, mean, var, eps, o, y, jl, stream)         \
                     : launch_posterior_mma<D_, D_ + 1>(h, st, x, H, mean, var, eps, o, y, jl, stream);
  switch (st.d) {
    PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) PM_CASE(5) PM_CASE(6)
  }
#undef PM_CASE
  return fail(h, GPMPC_ERR_ARG, "unsupported d");
}

extern "C" {

const char* gpmpc_version(void) { return "gpmpc_b200 0.4 (sm_100a)"; }

int64_t gpmpc_base_samples(uint8_t* rng_state, int64_t state_bytes, int64_t slots, int64_t n, double beta, double* out) {
  if (!rng_state || !out || state_bytes != (int64_t)gpmpc_rng::STATE_BYTES || slots < 0 || n < 1 || !(beta > 0.0))
    return GPMPC_ERR_ARG;
  return gpmpc_rng::truncated_candidates(rng_state, slots, n, beta, out);
}

const char* gpmpc_last_error(const gpmpc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int gpmpc_create(const gpmpc_dims* dims, gpmpc_handle** out) {
  if (!dims || !out) return fail(nullptr, GPMPC_ERR_ARG, "null argument");
  *out = nullptr;
  if (dims->ns < 1 || dims->g_ny < 1 || dims->d < 1 || dims->d > GPMPC_MAX_D || dims->n_real < 1 ||
      dims->cap_points < 0 || !(dims->T == 1 || dims->T == dims->d + 1))
    return fail(nullptr, GPMPC_ERR_ARG, "bad dims (need 1<=d<=6, T in {1,d+1}, ns,g_ny,n_real>=1)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, GPMPC_ERR_CUDA, "no CUDA device: gpmpc_b200 has no CPU fallback");
  gpmpc_handle* h = new gpmpc_handle();
  h->dims = *dims;
  cudaGetDevice(&h->device);
  cudaDeviceGetAttribute(&h->max_dyn_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
  if (const char* e = getenv("GPMPC_WO_MIN_M")) h->wo_min_m = atoi(e);
  if (const char* e = getenv("GPMPC_BLOCK_SCALAR")) h->block_mma = atoi(e) == 0;
  if (const char* e = getenv("GPMPC_WO_MAX_NB")) h->wo_max_nb = std::min(3, std::max(1, atoi(e)));
  if (const char* e = getenv("GPMPC_WO_SLAB_NB3")) h->wo_slab_nb3 = atoi(e) != 0;
  if (const char* e = getenv("GPMPC_ROLLOUT_FUSED")) h->fused_rollout = std::max(0, std::min(2, atoi(e)));
  if (const char* e = getenv("GPMPC_HZ_GROUPS")) h->hz_groups_cap = atoi(e);
  if (const char* e = getenv("GPMPC_HZ_STAGGER_NS")) h->hz_stagger_ns = atoll(e);
  DevState& st = h->st;
  st.ns = dims->ns; st.g_ny = dims->g_ny; st.d = dims->d; st.T = dims->T; st.n_real = dims->n_real;
  st.B = dims->ns * dims->g_ny;
  st.cap_points = dims->cap_points;
  st.jitter = 1e-6;
  unsigned* status;
  if (dev_alloc(&status, 1) != cudaSuccess || cudaMemset(status, 0, 4) != cudaSuccess) {
    delete h;
    return fail(nullptr, GPMPC_ERR_CUDA, "cudaMalloc(status) failed");
  }
  st.status = status;
  if (dev_alloc(&st.eig_flag, 1) != cudaSuccess || cudaMemset(st.eig_flag, 0, 4) != cudaSuccess) {
    cudaFree(status);
    delete h;
    return fail(nullptr, GPMPC_ERR_CUDA, "cudaMalloc(eig_flag) failed");
  }
  if (dev_alloc(&st.fin, (size_t)st.B * (st.T + st.T * (st.T + 1) / 2)) != cudaSuccess) {
    cudaFree(status);
    delete h;
    return fail(nullptr, GPMPC_ERR_CUDA, "cudaMalloc(fin) failed");
  }
  *out = h;
  return GPMPC_OK;
}

int gpmpc_destroy(gpmpc_handle* h) {
  ON_HANDLE_DEVICE(h);
  if (!h) return GPMPC_OK;
  DevState& st = h->st;
  free_factor_state(h);
  cudaFree((void*)st.Xr); cudaFree((void*)st.obs_pt); cudaFree((void*)st.obs_task); cudaFree((void*)st.y_obs);
  cudaFree((void*)st.ls); cudaFree((void*)st.os); cudaFree((void*)st.noise);
  cudaFree(st.Loo); cudaFree(st.LooP); cudaFree(st.beta_o); cudaFree(st.status); cudaFree(st.fin); cudaFree(st.Wo);
  if (h->status_pinned) cudaFreeHost(h->status_pinned);
  cudaFree(st.W); cudaFree(st.S); cudaFree(st.C); cudaFree(st.mu); cudaFree(st.xc); cudaFree(st.E); cudaFree(st.Lpre); cudaFree(st.eig_flag);
  cudaFree(h->S2buf); cudaFree(h->mu2buf);
  cudaFree(h->r_xu); cudaFree(h->r_xstar); cudaFree(h->r_y); cudaFree(h->d_active); cudaFree(h->c_scratch);
  cudaFree((void*)st.Yr); cudaFree((void*)st.real_full);
  cudaFree(h->grp_flags); cudaFree(h->grp_decision);
  cudaFree(h->lin_xu); cudaFree(h->lin_xg);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->slice_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : h->hz_ev) if (e) cudaEventDestroy(e);
  delete h;
  return GPMPC_OK;
}

int gpmpc_set_hypers(gpmpc_handle* h, const double* lengthscale, const double* outputscale,
                     const double* noise, double jitter) {
  ON_HANDLE_DEVICE(h);
  if (!h || !lengthscale || !outputscale || !noise) return fail(h, GPMPC_ERR_ARG, "null argument");
  DevState& st = h->st;
  for (int i = 0; i < st.g_ny * st.d; ++i)
    if (!(lengthscale[i] > 0.0)) return fail(h, GPMPC_ERR_ARG, "lengthscale must be > 0");
  for (int i = 0; i < st.g_ny; ++i)
    if (!(outputscale[i] > 0.0)) return fail(h, GPMPC_ERR_ARG, "outputscale must be > 0");
  h->h_ls.assign(lengthscale, lengthscale + st.g_ny * st.d);
  h->h_os.assign(outputscale, outputscale + st.g_ny);
  h->h_noise.assign(noise, noise + st.g_ny * st.T);
  double *ls, *os, *nz;
  if (!st.ls) {
    CUDA_TRY(h, dev_alloc(&ls, (size_t)st.g_ny * st.d));
    CUDA_TRY(h, dev_alloc(&os, (size_t)st.g_ny));
    CUDA_TRY(h, dev_alloc(&nz, (size_t)st.g_ny * st.T));
    st.ls = ls; st.os = os; st.noise = nz;
  }
  CUDA_TRY(h, cudaMemcpy((void*)st.ls, lengthscale, sizeof(double) * st.g_ny * st.d, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy((void*)st.os, outputscale, sizeof(double) * st.g_ny, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy((void*)st.noise, noise, sizeof(double) * st.g_ny * st.T, cudaMemcpyHostToDevice));
  st.jitter = jitter;
  h->have_hypers = true;
  h->have_real = false;  // factor must be rebuilt
  return GPMPC_OK;
}

int gpmpc_set_real_data(gpmpc_handle* h, const double* X, const double* Y, void* stream_) {
  ON_HANDLE_DEVICE(h);
  if (!h || !X || !Y) return fail(h, GPMPC_ERR_ARG, "null argument");
  if (!h->have_hypers) return fail(h, GPMPC_ERR_STATE, "call gpmpc_set_hypers first");
  cudaStream_t stream = (cudaStream_t)stream_;
  DevState& st = h->st;
  const int n = st.n_real, T = st.T, g_ny = st.g_ny, d = st.d;
  std::vector<double> hy((size_t)g_ny * n * T);
  CUDA_TRY(h, cudaMemcpyAsync(hy.data(), Y, hy.size() * 8, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(h, cudaStreamSynchronize(stream));
  // a slot is observed only if it is non-NaN for every batch element (here: every output)
  std::vector<int> pt, task;
  for (int p = 0; p < n; ++p)
    for (int t = 0; t < T; ++t) {
      bool ok = true;
      for (int j = 0; j < g_ny; ++j) ok = ok && !std::isnan(hy[((size_t)j * n + p) * T + t]);
      if (ok) { pt.push_back(p); task.push_back(t); }
    }
  const int m = (int)pt.size();
  if (m == 0) return fail(h, GPMPC_ERR_ARG, "no observed real data");
  std::vector<double> yobs((size_t)g_ny * m);
  for (int j = 0; j < g_ny; ++j)
    for (int i = 0; i < m; ++i) yobs[(size_t)j * m + i] = hy[((size_t)j * n + pt[i]) * T + task[i]];

  h->have_real = false;  // until every allocation and the factorisation below have succeeded
  h->real_cache_valid = false;
  cudaFree((void*)st.Xr); cudaFree((void*)st.obs_pt); cudaFree((void*)st.obs_task); cudaFree((void*)st.y_obs);
  cudaFree(st.Loo); cudaFree(st.LooP); cudaFree(st.beta_o);
  cudaFree((void*)st.Yr); cudaFree((void*)st.real_full);
  st.Xr = st.y_obs = st.Yr = nullptr;
  st.obs_pt = st.obs_task = st.real_full = nullptr;
  st.Loo = st.LooP = st.beta_o = nullptr;
  // per output: is every task of real point p observed?  (sample_gp's y_train_isnan, src/agent.py:674-679)
  std::vector<int> full((size_t)g_ny * n);
  for (int j = 0; j < g_ny; ++j)
    for (int p = 0; p < n; ++p) {
      bool ok = true;
      for (int t = 0; t < T; ++t) ok = ok && !std::isnan(hy[((size_t)j * n + p) * T + t]);
      full[(size_t)j * n + p] = ok ? 1 : 0;
    }
  double* Yr; int* rf;
  CUDA_TRY(h, dev_alloc(&Yr, hy.size()));
  CUDA_TRY(h, dev_alloc(&rf, full.size()));
  CUDA_TRY(h, cudaMemcpyAsync(Yr, Y, hy.size() * 8, cudaMemcpyDeviceToDevice, stream));
  CUDA_TRY(h, cudaMemcpyAsync(rf, full.data(), full.size() * 4, cudaMemcpyHostToDevice, stream));
  st.Yr = Yr; st.real_full = rf;
  double *Xr, *yo; int *op, *ot;
  CUDA_TRY(h, dev_alloc(&Xr, (size_t)n * d));
  CUDA_TRY(h, dev_alloc(&op, (size_t)m));
  CUDA_TRY(h, dev_alloc(&ot, (size_t)m));
  CUDA_TRY(h, dev_alloc(&yo, (size_t)g_ny * m));
  CUDA_TRY(h, dev_alloc(&st.Loo, (size_t)g_ny * m * m));
  CUDA_TRY(h, dev_alloc(&st.LooP, (size_t)g_ny * subpanel_off((m + 7) / 8, 0)));
  CUDA_TRY(h, dev_alloc(&st.beta_o, (size_t)g_ny * m));
  CUDA_TRY(h, cudaMemcpyAsync(Xr, X, (size_t)n * d * 8, cudaMemcpyDeviceToDevice, stream));
  CUDA_TRY(h, cudaMemcpyAsync(op, pt.data(), (size_t)m * 4, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(h, cudaMemcpyAsync(ot, task.data(), (size_t)m * 4, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(h, cudaMemcpyAsync(yo, yobs.data(), yobs.size() * 8, cudaMemcpyHostToDevice, stream));
  st.Xr = Xr; st.obs_pt = op; st.obs_task = ot; st.y_obs = yo;
  const bool m_changed = (m != st.m);
  st.m = m;
  st.mo = (m + 7) & ~7;
  st.c = 0; st.np = 0;
  h->has_partial = false;
  h->first_partial_point = -1;
  h->c_before_point.clear();
  if (m_changed || !st.Lh) {
    // the slab height depends on m: (re)allocate the per-element state from scratch
    free_factor_state(h);
    int rc = alloc_factor_state(h, st.cap_points, stream);
    if (rc) return rc;
  }
  {
    int coop_ok = 0;
    cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, h->device);
    // m in the thousands: blocked factorisation + strip-wise inverse on the tensor cores (gpmpc_k0.cuh; 2.7 s -> ~0.1 s at
    // m = 10^4); below GPMPC_K0_BLOCKED_MIN_M the per-pivot kernels, whose results the reference-sized tests are pinned to
    int blocked_min_m = 768;
    if (const char* e = getenv("GPMPC_K0_BLOCKED_MIN_M")) blocked_min_m = atoi(e);
    if (coop_ok && m >= blocked_min_m) {
      const size_t smem = (size_t)3 * K0P * K0P_LD * sizeof(double);
      CUDA_TRY(h, opt_in_smem(h, k_factor_real_blocked, (int)smem));
      void* args[] = {(void*)&st};
      CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)k_factor_real_blocked, dim3(h->num_sms), dim3(K0F_THREADS), args, smem, stream));
      const int Pm = (m + 7) / 8;
      k_invert_real_mma<<<dim3((Pm + 1) / 2, g_ny), K0I_WARPS * 32, 0, stream>>>(st);
      k_beta_from_inverse<<<dim3((Pm + 3) / 4, g_ny), 128, 0, stream>>>(st);
    } else {
      CUDA_TRY(h, opt_in_smem(h, k_factor_real, h->max_dyn_smem));
      CUDA_TRY(h, opt_in_smem(h, k_invert_real, h->max_dyn_smem));
      int k0b_warps = K0B_WARPS;  // one column of inv(L_oo) per warp, m doubles of shared memory each
      while (k0b_warps > 1 && (size_t)m * 8 * k0b_warps > (size_t)h->max_dyn_smem) k0b_warps >>= 1;
      if ((size_t)m * 8 * k0b_warps > (size_t)h->max_dyn_smem)
        return fail(h, GPMPC_ERR_ARG, "more observed real scalars than K0 supports (m <= ~29000)");
      // m in the thousands: the factorisation cooperatively over all SMs (one grid barrier per pivot); GPMPC_K0_COOP_MIN_M
      int coop_min_m = 768;
      if (const char* e = getenv("GPMPC_K0_COOP_MIN_M")) coop_min_m = atoi(e);
      if (coop_ok && m >= coop_min_m && (size_t)m * 16 + 1024 <= (size_t)h->max_dyn_smem) {
        CUDA_TRY(h, opt_in_smem(h, k_factor_real_coop, h->max_dyn_smem));
        void* args[] = {(void*)&st};
        CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)k_factor_real_coop, dim3(h->num_sms), dim3(K0_THREADS), args,
                                                (size_t)m * 16, stream));
      } else {
        k_factor_real<<<g_ny, K0_THREADS, (size_t)m * 8, stream>>>(st);
      }
      k_invert_real<<<dim3((m + k0b_warps - 1) / k0b_warps, g_ny), k0b_warps * 32, (size_t)m * 8 * k0b_warps, stream>>>(st);
    }
    h->launches++;
  }
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(stream));  // host staging vectors go out of scope
  h->have_real = true;
  h->factor_version++;
  h->ws_n = 0;  // force workspace re-evaluation
  return GPMPC_OK;
}

int gpmpc_reset_hallucinated(gpmpc_handle* h) {
  if (!h) return GPMPC_ERR_ARG;
  // (st.pstate needs no clearing: the state of point p is written by the step that records it, entries at p >= np are never read)
  h->st.c = 0;
  h->st.np = 0;
  h->has_partial = false;
  h->first_partial_point = -1;
  h->c_before_point.clear();
  h->factor_version++;
  return GPMPC_OK;
}

int gpmpc_truncate_hallucinated(gpmpc_handle* h, int32_t n_points) {
  if (!h || n_points < 0 || n_points > h->st.np) return fail(h, GPMPC_ERR_ARG, "bad point count");
  if (h->grp_size > 0) return fail(h, GPMPC_ERR_STATE, "grouped handle");
  if (n_points == h->st.np) return GPMPC_OK;
  if ((int)h->c_before_point.size() < h->st.np) return fail(h, GPMPC_ERR_STATE, "point bookkeeping incomplete");
  // rows / records beyond the cut are simply not used any more (appending overwrites them)
  h->st.c = h->condition ? h->c_before_point[n_points] : 0;
  h->st.np = n_points;
  h->c_before_point.resize(n_points);
  if (h->first_partial_point >= n_points) { h->first_partial_point = -1; h->has_partial = false; }
  h->factor_version++;
  return GPMPC_OK;
}

int gpmpc_reserve(gpmpc_handle* h, int32_t cap_points, void* stream) {
  ON_HANDLE_DEVICE(h);
  if (!h || cap_points < 0) return fail(h, GPMPC_ERR_ARG, "bad capacity");
  if (!h->have_real) { h->st.cap_points = std::max(h->st.cap_points, cap_points); return GPMPC_OK; }
  if (cap_points <= h->st.cap_points) return GPMPC_OK;
  return alloc_factor_state(h, cap_points, (cudaStream_t)stream);
}

int gpmpc_set_condition_on_hallucinated(gpmpc_handle* h, int32_t on) {
  ON_HANDLE_DEVICE(h);
  if (!h) return GPMPC_ERR_ARG;
  const bool was = h->condition;
  h->condition = on != 0;
  if (h->condition && !was && h->have_real) {
    if (h->st.c != 0 || h->st.np != 0)
      return fail(h, GPMPC_ERR_STATE, "switch conditioning on only with an empty hallucinated set");
    if (h->st.c_cap < h->st.cap_points * h->st.T) return alloc_factor_state(h, h->st.cap_points, 0);
  }
  return GPMPC_OK;
}

static int check_ready(gpmpc_handle* h) {
  if (!h) return GPMPC_ERR_ARG;
  if (!h->have_real) return fail(h, GPMPC_ERR_STATE, "call gpmpc_set_hypers and gpmpc_set_real_data first");
  return GPMPC_OK;
}

int gpmpc_posterior(gpmpc_handle* h, const double* x, int32_t H, double* mean, double* var,
                    const double* eps, const gpmpc_sample_opts* opts, double* y, int32_t* jitter_level,
                    void* stream) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || H < 1) return fail(h, GPMPC_ERR_ARG, "bad x / H");
  if (eps && (!opts || !y)) return fail(h, GPMPC_ERR_ARG, "eps given without opts / y");
  if (h->grp_size > 0 && h->st.np > 0)
    return fail(h, GPMPC_ERR_STATE, "grouped handle (gpmpc_set_grouping): only gpmpc_step / gpmpc_rollout see the per-Agent point masks");
  rc = ensure_workspace(h, H);
  if (rc) return rc;
  if (eps) h->st.eig_epoch = ++h->eig_epoch;
  DevState st = h->st;
  st.W_stride = (long long)h->ws_n * (H * st.T);
  gpmpc_sample_opts o = opts ? *opts : gpmpc_sample_opts{-1.0, -1.0, 0, 0};
  const bool with_real = h->want_real_only && h->block_mma && !h->has_partial && H * st.T <= PM_MAX_Q && st.c > 0;
  h->want_real_only = false;
  h->real_cache_valid = with_real;  // any other model call overwrites W: the real-only cache goes with it
  h->real_cache_H = H;
  if (with_real) { st.S2 = h->S2buf; st.mu2 = h->mu2buf; }
  h->pre_kind = 0;  // (set again by launch_posterior_mma when this call prefactors)
  if (!(h->block_mma && !h->has_partial && H * st.T <= PM_MAX_Q)) h->prefactor_next = false;
  if (h->block_mma && !h->has_partial && H * st.T <= PM_MAX_Q) {
    rc = dispatch_posterior_mma(h, st, x, H, mean, var, eps, o, y, jitter_level, (cudaStream_t)stream);
    if (rc) return rc;
  } else {
    k_posterior<<<st.B, BLK_THREADS, 0, (cudaStream_t)stream>>>(st, x, H, mean, var, eps, o, y, jitter_level);
    rc = launch_sample_eig(h, st, H, eps, o, y, jitter_level, (cudaStream_t)stream);
    if (rc) return rc;
  }
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  h->cache_version = h->factor_version;
  h->cache_H = H;
  count_work(h, H, false);
  return GPMPC_OK;
}

int gpmpc_sample(gpmpc_handle* h, const double* eps, const gpmpc_sample_opts* opts, double* y,
                 int32_t* jitter_level, void* stream) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!eps || !opts || !y) return fail(h, GPMPC_ERR_ARG, "null argument");
  if (h->cache_version != h->factor_version || h->cache_H < 1)
    return fail(h, GPMPC_ERR_STATE, "gpmpc_sample needs a preceding gpmpc_posterior on the current factor");
  h->st.eig_epoch = ++h->eig_epoch;
  DevState st = h->st;
  st.W_stride = (long long)h->ws_n * (h->cache_H * st.T);
  rc = configure_block_smem(h);
  if (rc) return rc;
  size_t tri = tri_bytes(h, h->cache_H * st.T);
  if (tri + 8192 > (size_t)h->max_dyn_smem) tri = 0;
  k_sample<<<st.B, BLK_THREADS, tri, (cudaStream_t)stream>>>(st, h->cache_H, eps, *opts, y, jitter_level, tri ? 1 : 0);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return launch_sample_eig(h, st, h->cache_H, eps, *opts, y, jitter_level, (cudaStream_t)stream);
}

int gpmpc_append(gpmpc_handle* h, const double* x, const double* y, const uint8_t* point_active, int32_t H,
                 void* stream_) {
  if (!h || H < 1 || !point_active) return gpmpc_append_masked(h, x, y, nullptr, H, stream_);
  std::vector<uint8_t> sc((size_t)H * h->st.T);
  for (int i = 0; i < H; ++i)
    for (int t = 0; t < h->st.T; ++t) sc[(size_t)i * h->st.T + t] = point_active[i] ? 1 : 0;
  return gpmpc_append_masked(h, x, y, sc.data(), H, stream_);
}

int gpmpc_append_masked(gpmpc_handle* h, const double* x, const double* y, const uint8_t* scalar_active, int32_t H,
                        void* stream_) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || !y || H < 1) return fail(h, GPMPC_ERR_ARG, "bad x / y / H");
  if (h->grp_size > 0) return fail(h, GPMPC_ERR_STATE, "grouped handle (gpmpc_set_grouping): append through gpmpc_step / gpmpc_rollout");
  cudaStream_t stream = (cudaStream_t)stream_;
  DevState& hst = h->st;
  const int T = hst.T;
  if (H * T > 512) return fail(h, GPMPC_ERR_ARG, "gpmpc_append supports at most 512 new scalars per call");
  int n_active = 0;  // scalars entering the factor
  bool partial = false;
  for (int i = 0; i < H; ++i) {
    int mine = 0;
    for (int t = 0; t < T; ++t) mine += (!scalar_active || scalar_active[(size_t)i * T + t]) ? 1 : 0;
    n_active += mine;
    partial = partial || (mine != 0 && mine != T);
  }
  if (hst.np + H > hst.cap_points) {
    rc = alloc_factor_state(h, std::max(hst.np + H, hst.cap_points * 2), stream);
    if (rc) return rc;
  }
  const bool grow = h->condition && n_active > 0;
  unsigned char* d_act = nullptr;
  if (grow) {
    rc = ensure_workspace(h, H);
    if (rc) return rc;
    if (scalar_active && n_active < H * T) {
      if (h->d_active_cap < H * T) {
        cudaFree(h->d_active);
        CUDA_TRY(h, dev_alloc(&h->d_active, (size_t)H * T));
        h->d_active_cap = H * T;
      }
      CUDA_TRY(h, cudaMemcpyAsync(h->d_active, scalar_active, (size_t)H * T, cudaMemcpyHostToDevice, stream));
      d_act = h->d_active;
    }
  }
  DevState st = h->st;
  st.W_stride = (long long)h->ws_n * (H * st.T);
  int reuse = (h->cache_version == h->factor_version && h->cache_H == H) ? 1 : 0;
  int prefactored = (reuse && h->pre_kind == 1 && h->pre_version == h->factor_version && h->pre_H == H && !d_act) ? 1 : 0;
  if (grow && !reuse && hst.c == 0 && h->real_cache_valid && h->real_cache_H == H) {
    prefactored = (h->pre_kind == 2 && h->pre_H == H && !d_act) ? 1 : 0;
    // the hallucinated set was reset after the model call: W's real rows, Sigma* and the mean w.r.t. the real data alone
    // came out of that same pass (k_pm_gram's second accumulators); k_append checks on the device that x is the same
    st.S = h->S2buf;
    st.mu = h->mu2buf;
    reuse = 1;
  }
  h->real_cache_valid = false;
  if (grow && !reuse && h->block_mma && !h->has_partial && H * st.T <= PM_MAX_Q) {
    // the factor changed since the last model call (e.g. the reset of agent.py:261-272 between the model call and the
    // conditioning): W, Sigma*, mean for these points against the CURRENT factor, on the tensor-core kernels
    rc = dispatch_posterior_mma(h, st, x, H, nullptr, nullptr, nullptr, gpmpc_sample_opts{-1.0, -1.0, 0, 0}, nullptr, nullptr, stream);
    if (rc) return rc;
    h->launches++;
    h->cache_version = h->factor_version;
    h->cache_H = H;
    reuse = 1;
  }
  rc = configure_block_smem(h);
  if (rc) return rc;
  size_t tri = grow ? tri_bytes(h, H * st.T) : 0;
  if (tri + 8192 > (size_t)h->max_dyn_smem) tri = 0;  // k_append also holds 2 KB of static shared memory
  if (!grow || !tri || !reuse) prefactored = 0;
  h->pre_kind = 0;
  k_append<<<st.B, BLK_THREADS, tri, stream>>>(st, x, y, d_act, H, st.np, reuse, grow ? 1 : 0, tri ? 1 : 0, prefactored);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  if (d_act) CUDA_TRY(h, cudaStreamSynchronize(stream));  // caller may reuse point_active's host memory
  count_work(h, H, grow);
  {
    int cc = hst.c;
    for (int i = 0; i < H; ++i) {
      h->c_before_point.push_back(cc);
      int mine = 0;
      for (int t = 0; t < T; ++t) mine += (!scalar_active || scalar_active[(size_t)i * T + t]) ? 1 : 0;
      if (grow) cc += mine;
      if (grow && mine != 0 && mine != T && h->first_partial_point < 0) h->first_partial_point = hst.np + i;
    }
  }
  hst.np += H;
  if (grow) {
    hst.c += n_active;
    h->factor_version++;
    h->has_partial = h->has_partial || partial;
  }
  return GPMPC_OK;
}

}  // extern "C"

template <int T>
static int launch_step_finish(gpmpc_handle* h, const DevState& st, const double* x, const double* eps,
                              const gpmpc_sample_opts& o, double* mean, double* var, double* y, int* jl, int grow,
                              cudaStream_t stream) {
  const int fin_blocks = (st.B + FIN_THREADS - 1) / FIN_THREADS;
  if (h->grp_size > 0 && eps && grow) {
    // grouped step: draw, per-element "too close" flags, per-Agent all / any reduction, then the append (null rows for a
    // masked / dropped point)
    k_step_finish<T><<<fin_blocks, FIN_THREADS, 0, stream>>>(st, x, eps, o, mean, var, y, jl, grow, FIN_DRAW);
    const int warps = 4;
    k_filter_new_points<<<(st.B + warps - 1) / warps, warps * 32, 0, stream>>>(st, x, 1, h->grp_min_dist, 1, nullptr, nullptr,
                                                                              h->grp_flags);
    const int n_groups = (st.ns + h->grp_size - 1) / h->grp_size;
    k_group_decide<<<(n_groups + warps - 1) / warps, warps * 32, 0, stream>>>(st.ns, st.g_ny, h->grp_size, h->grp_flags,
                                                                              h->grp_decision);
    k_step_finish<T><<<fin_blocks, FIN_THREADS, 0, stream>>>(st, x, eps, o, nullptr, nullptr, y, nullptr, grow, FIN_APPEND,
                                                             GroupArgs{h->grp_flags, h->grp_decision, h->grp_size * st.g_ny});
    h->launches += 4;
    CUDA_TRY(h, cudaGetLastError());
    return GPMPC_OK;
  }
  // the plain fused step: the diagonal-block part of the append specialised on c mod 8 (uniform over the launch)
#define FIN_CASE(IF_) case IF_: k_step_finish<T, IF_><<<fin_blocks, FIN_THREADS, 0, stream>>>(st, x, eps, o, mean, var, y, jl, grow, FIN_ALL); break;
  switch ((eps && grow && T > 1) ? (st.c & 7) : -1) {
    FIN_CASE(0) FIN_CASE(1) FIN_CASE(2) FIN_CASE(3) FIN_CASE(4) FIN_CASE(5) FIN_CASE(6) FIN_CASE(7)
    default: k_step_finish<T><<<fin_blocks, FIN_THREADS, 0, stream>>>(st, x, eps, o, mean, var, y, jl, grow, FIN_ALL);
  }
#undef FIN_CASE
  h->launches++;
  if (eps && T > 1 && !(o.flags & GPMPC_OPT_NO_EIG_FALLBACK)) {
    // GPyTorch's batch-wide eigen-root fallback: exits on the device unless an element of this step failed its ladder
    k_step_finish<T><<<fin_blocks, FIN_THREADS, 0, stream>>>(st, x, eps, o, mean, var, y, jl, grow, FIN_EIG_REDO);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

// Regime A (large m): W_o = inv(L_oo) K_o for every element, NB column blocks of 8 per tile
// slab = rows of K per pass (multiple of 8); slab >= mo: the one-pass form
template <int D, int T, int NB>
static int launch_shared_rows(gpmpc_handle* h, const DevState& st, const double* x, cudaStream_t stream, int slab = 1 << 30) {
  constexpr int E = 8 * NB / T;
  auto kern = k_shared_rows<D, T, NB>;
  CUDA_TRY(h, opt_in_smem(h, kern, h->max_dyn_smem));
  slab = std::min(slab, st.mo);
  const size_t smem = ((size_t)slab * 8 * NB + ((E * D + 1) & ~1)) * 8;
  const int n_tiles = (st.ns + E - 1) / E;
  // few row panels (m of a few hundred): 8 warps per CTA balance them better than 16 and two CTAs share an SM
  const int Pm = st.mo / 8;
  const int threads = Pm <= 32 ? 256 : SR_THREADS;
  const int per_sm = threads == 256 && 2 * (smem + 1024) <= (size_t)h->max_dyn_smem ? 2 : 1;
  int sms = h->num_sms;
  if (h->sr_grid_cap > 0) sms = std::min(sms, h->sr_grid_cap);
  dim3 grid(std::min(n_tiles, std::max(1, sms * per_sm / st.g_ny)), st.g_ny);
  for (int k_lo = 0; k_lo < st.mo; k_lo += slab) {
    kern<<<grid, threads, smem, stream>>>(st, x, k_lo, std::min(st.mo, k_lo + slab));
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

// The step for training sets beyond k_step's per-warp w array (m in the thousands): batched shared-rows GEMM in k-slabs,
// then k_step_big (one CTA per element, w_o in global memory), then the usual finishing kernel.
template <int D, int T>
static int launch_step_big(gpmpc_handle* h, const DevState& st, const double* x, const double* eps, const gpmpc_sample_opts& o,
                           double* mean, double* var, double* y, int* jl, int grow, cudaStream_t stream, bool* handled) {
  *handled = false;
  const size_t budget = (size_t)h->max_dyn_smem;
  const int P8 = (st.c + 7) / 8;
  const int FS = T + T * (T + 1) / 2;
  const size_t big_smem = ((size_t)std::max(8 * P8, 1) * T + (size_t)(BIG_THREADS / 32) * std::max(8 * T, FS)) * 8;
  if (big_smem + 1024 > budget) return GPMPC_OK;  // own rows beyond shared memory too: the general block kernels
  const size_t need = (size_t)st.B * st.mo * T;
  if (h->wo_count < need) {
    cudaFree(h->st.Wo);
    h->st.Wo = nullptr;
    h->wo_count = 0;
    CUDA_TRY(h, dev_alloc(&h->st.Wo, need));
    h->wo_count = need;
  }
  DevState stw = st;
  stw.Wo = h->st.Wo;
  // column blocks per tile: as many as leave a slab of >= 512 rows of K in shared memory
  int nb = std::min(3, h->wo_max_nb), slab = 0;
  for (; nb >= 1; --nb) {
    slab = (int)((budget - 2048) / ((size_t)64 * nb)) & ~7;
    if (slab >= 512 || nb == 1) break;
  }
  if (h->big_slab_cap >= 8) slab = std::min(slab, h->big_slab_cap);
  if (slab < 8) return GPMPC_OK;
  int rc = nb == 3 ? launch_shared_rows<D, T, 3>(h, stw, x, stream, slab)
           : nb == 2 ? launch_shared_rows<D, T, 2>(h, stw, x, stream, slab)
                     : launch_shared_rows<D, T, 1>(h, stw, x, stream, slab);
  if (rc) return rc;
  CUDA_TRY(h, opt_in_smem(h, k_step_big, h->max_dyn_smem));
  k_step_big<<<st.B, BIG_THREADS, big_smem, stream>>>(stw, x, grow);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  *handled = true;
  return launch_step_finish<T>(h, stw, x, eps, o, mean, var, y, jl, grow, stream);
}

template <int D, int T, bool LOO_SMEM, bool WO = false>
static int launch_step_impl(gpmpc_handle* h, const DevState& st, const double* x, const double* eps,
                            const gpmpc_sample_opts& o, double* mean, double* var, double* y, int* jl, int grow,
                            int warps, size_t smem, cudaStream_t stream) {
  auto kern = k_step<D, T, LOO_SMEM, WO>;
  CUDA_TRY(h, opt_in_smem(h, kern, h->max_dyn_smem, true));
  // persistent CTAs, one per SM, split over the g_ny outputs; each warp loops over samples
  const int want = (st.ns + warps - 1) / warps;
  int resident = std::max(1, h->num_sms / st.g_ny);
  if (h->step_grid_cap > 0) resident = std::max(1, std::min(resident, h->step_grid_cap / st.g_ny));
  dim3 grid(std::min(want, resident), st.g_ny);
  kern<<<grid, warps * 32, smem, stream>>>(st, x, grow);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return launch_step_finish<T>(h, st, x, eps, o, mean, var, y, jl, grow, stream);
}

// c == 0 and nothing appended: the shared-factor kernel (8 / T elements per tensor-core tile)
template <int D, int T>
static int launch_step_shared(gpmpc_handle* h, const DevState& st, const double* x, const double* eps,
                              const gpmpc_sample_opts& o, double* mean, double* var, double* y, int* jl,
                              cudaStream_t stream, bool* handled) {
  constexpr int G = 8 / T;
  const int Pm = (st.m + 7) / 8;
  const size_t loop_sz = subpanel_off(Pm, 0);
  const size_t nr_even = (st.n_real + 1) & ~1;
  const size_t fixed = (loop_sz + nr_even * D + st.mo) * 8 + (size_t)((st.n_real * T + 1) & ~1) * 4 + 128;
  const size_t per_warp = ((size_t)st.mo * 8 + 8 * D) * 8;
  const size_t budget = (size_t)h->max_dyn_smem;
  *handled = fixed + 4 * per_warp <= budget;
  if (!*handled) return GPMPC_OK;  // inv(L_oo) too large for shared memory: the general kernel takes over
  const int groups = (st.ns + G - 1) / G;
  int warps = (int)std::min<size_t>(STEP_MAX_WARPS, (budget - fixed) / per_warp);
  const int per_cta_need = (groups * st.g_ny + h->num_sms - 1) / h->num_sms;
  warps = std::max(1, std::min(warps, per_cta_need));
  auto kern = k_step_shared<D, T>;
  CUDA_TRY(h, opt_in_smem(h, kern, h->max_dyn_smem));
  const int want = (groups + warps - 1) / warps;
  const int resident = std::max(1, h->num_sms / st.g_ny);
  dim3 grid(std::min(want, resident), st.g_ny);
  kern<<<grid, warps * 32, fixed + (size_t)warps * per_warp, stream>>>(st, x);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return launch_step_finish<T>(h, st, x, eps, o, mean, var, y, jl, 0, stream);
}

template <int D, int T>
static int launch_step(gpmpc_handle* h, const DevState& st, const double* x, const double* eps,
                       const gpmpc_sample_opts& o, double* mean, double* var, double* y, int* jl, int grow,
                       cudaStream_t stream, bool* handled) {
  if (st.c == 0 && !grow) {
    int rc = launch_step_shared<D, T>(h, st, x, eps, o, mean, var, y, jl, stream, handled);
    if (rc || *handled) return rc;
  }
  if (h->force_big && h->grp_size == 0) return launch_step_big<D, T>(h, st, x, eps, o, mean, var, y, jl, grow, stream, handled);
  // shared-memory budget, mirroring the carve-up at the top of k_step
  const int m = st.m, Pm = (m + 7) / 8, P8 = (st.c + 7) / 8;
  const size_t loop_sz = subpanel_off(Pm, 0);
  const size_t m_even = (m + 1) & ~1;
  const size_t nr_even = (st.n_real + 1) & ~1;
  const size_t shared_tab = (nr_even * D + m_even) * 8 + (size_t)(((st.n_real * T + 1) & ~1) + ((st.np + 1) & ~1)) * 4 + 128;
  const size_t wv_rows = st.mo + 8 * P8;
  const size_t wv_sz = (wv_rows * T + 8 + 15) & ~(size_t)15, wb_sz = (wv_rows + 15) & ~(size_t)15;
  const size_t per_warp = (wv_sz + wb_sz + (size_t)STEP_NST * STEP_SEG * 8) * 8 + STEP_NST * 8;
  const size_t budget = (size_t)h->max_dyn_smem;
  // L_oo lives in shared memory if at least 8 warps still fit beside it; otherwise it is read through L1/L2
  bool loo_smem = loop_sz * 8 + shared_tab + 8 * per_warp <= budget;
  // ... or, for large m, not by this kernel at all: batched GEMM first (k_shared_rows), NB = column blocks per tile
  int nb = 0;
  if ((!loo_smem || h->force_wo) && m >= h->wo_min_m)
    for (nb = h->wo_max_nb; nb >= 1; --nb)
      if ((size_t)st.mo * 8 * nb * 8 + 1024 <= budget) break;
  if (nb >= 1) {
    const size_t fixed_wo = m_even * 8 + (size_t)((st.np + 1) & ~1) * 4 + 128;
    if (fixed_wo + per_warp <= budget) {
      const size_t need = (size_t)st.B * st.mo * T;
      if (h->wo_count < need) {
        cudaFree(h->st.Wo);
        h->st.Wo = nullptr;
        h->wo_count = 0;
        CUDA_TRY(h, dev_alloc(&h->st.Wo, need));
        h->wo_count = need;
      }
      DevState stw = st;
      stw.Wo = h->st.Wo;
      // m beyond the one-pass limit of three column blocks (~1200): rather than fewer column blocks per tile -- inv(L_oo) is then
      // re-streamed from L2 once per 5.3 / 2.7 samples and the GEMM turns L2-bound (0.28-0.35 of the DMMA peak at m = 2000-3000) --
      // keep three and walk K in slabs of as many rows as fit, accumulating into Wo ("wo_slab_nb3" 0 switches this off)
      int slab = 1 << 30;
      if (nb < 3 && h->wo_max_nb >= 3 && h->wo_slab_nb3) {
        nb = 3;
        slab = (int)((budget - 2048) / ((size_t)64 * 3)) & ~7;
      }
      int rc = nb == 3 ? launch_shared_rows<D, T, 3>(h, stw, x, stream, slab)
               : nb == 2 ? launch_shared_rows<D, T, 2>(h, stw, x, stream)
                         : launch_shared_rows<D, T, 1>(h, stw, x, stream);
      if (rc) return rc;
      int warps = (int)std::min<size_t>(STEP_MAX_WARPS, (budget - fixed_wo) / per_warp);
      const int per_cta_need = (st.ns * st.g_ny + h->num_sms - 1) / h->num_sms;
      warps = std::max(1, std::min(warps, std::max(per_cta_need, 1)));
      *handled = true;
      return launch_step_impl<D, T, false, true>(h, stw, x, eps, o, mean, var, y, jl, grow, warps,
                                                 fixed_wo + (size_t)warps * per_warp, stream);
    }
  }
  const size_t fixed = shared_tab + (loo_smem ? loop_sz * 8 : 0);
  *handled = fixed + per_warp <= budget;
  if (!*handled && h->grp_size == 0 && !h->force_block_fallback)  // m beyond the per-warp w array: GEMM in k-slabs + k_step_big
    return launch_step_big<D, T>(h, st, x, eps, o, mean, var, y, jl, grow, stream, handled);
  if (!*handled) return GPMPC_OK;  // general block kernels take over
  int warps = (int)std::min<size_t>(STEP_MAX_WARPS, (budget - fixed) / per_warp);
  if (h->step_warps_cap > 0) warps = std::min(warps, h->step_warps_cap);
  // small launches: spread the samples over the SMs rather than filling few CTAs
  const int per_cta_need = (st.ns * st.g_ny + h->num_sms - 1) / h->num_sms;
  warps = std::max(1, std::min(warps, std::max(per_cta_need, 1)));
  const size_t smem = fixed + (size_t)warps * per_warp;
  return loo_smem ? launch_step_impl<D, T, true>(h, st, x, eps, o, mean, var, y, jl, grow, warps, smem, stream)
                  : launch_step_impl<D, T, false>(h, st, x, eps, o, mean, var, y, jl, grow, warps, smem, stream);
}

static int dispatch_step(gpmpc_handle* h, const DevState& st, const double* x, const double* eps,
                         const gpmpc_sample_opts& o, double* mean, double* var, double* y, int* jl, int grow,
                         cudaStream_t stream, bool* handled) {
#define STEP_CASE(D_)                                                                                  \
  case D_:                                                                                             \
    return st.T == 1 ? launch_step<D_, 1>(h, st, x, eps, o, mean, var, y, jl, grow, stream, handled) \
                     : launch_step<D_, D_ + 1>(h, st, x, eps, o, mean, var, y, jl, grow, stream, handled);
  switch (st.d) {
    STEP_CASE(1) STEP_CASE(2) STEP_CASE(3) STEP_CASE(4) STEP_CASE(5) STEP_CASE(6)
  }
#undef STEP_CASE
  return fail(h, GPMPC_ERR_ARG, "unsupported d");
}

extern "C" {

int gpmpc_step(gpmpc_handle* h, const double* x, const double* eps, const gpmpc_sample_opts* opts,
               double* mean, double* var, double* y, int32_t* jitter_level, void* stream_) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x) return fail(h, GPMPC_ERR_ARG, "null x");
  if (eps && (!opts || !y)) return fail(h, GPMPC_ERR_ARG, "eps given without opts / y");
  cudaStream_t stream = (cudaStream_t)stream_;
  DevState& hst = h->st;
  if (eps && hst.np + 1 > hst.cap_points) {
    rc = alloc_factor_state(h, std::max(hst.np + 1, hst.cap_points * 2), stream);
    if (rc) return rc;
  }
  const int grow = (eps && h->condition) ? 1 : 0;
  gpmpc_sample_opts o = opts ? *opts : gpmpc_sample_opts{-1.0, -1.0, 0, 0};
  if (eps) hst.eig_epoch = ++h->eig_epoch;
  count_work(h, 1, grow);
  bool handled = false;
  if (!h->has_partial) {  // the fused kernel finds a point's factor rows through hrow0: whole points only
    rc = dispatch_step(h, h->st, x, eps, o, mean, var, y, jitter_level, grow, stream, &handled);
    if (rc) return rc;
  }
  if (!handled && h->grp_size > 0)
    return fail(h, GPMPC_ERR_CAPACITY, "grouped step: the factor does not fit the fused step kernel");
  if (!handled) {
    // factor too tall for the register-resident sweep: same recursion through the general block kernels
    const double w_bytes = h->last_bytes, w_flops = h->last_flops;
    rc = gpmpc_posterior(h, x, 1, mean, var, eps, opts, y, jitter_level, stream_);
    if (rc) return rc;
    if (eps) rc = gpmpc_append(h, x, y, nullptr, 1, stream_);
    h->last_bytes = w_bytes;
    h->last_flops = w_flops;
    return rc;
  }
  if (eps) {
    h->c_before_point.push_back(hst.c);
    hst.np += 1;
    if (grow) {
      hst.c += hst.T;
      h->factor_version++;
    }
  }
  return GPMPC_OK;
}

int gpmpc_assemble(gpmpc_handle* h, const gpmpc_env* env, const double* xu, const double* y_gp, int32_t H,
                   double* out, void* stream) {
  ON_HANDLE_DEVICE(h);
  if (!h || !env || !xu || !y_gp || !out || H < 1) return fail(h, GPMPC_ERR_ARG, "null argument");
  if (env->nx > GPMPC_MAX_NX || env->nx + env->nu > 2 * GPMPC_MAX_NX || env->g_ny != h->st.g_ny)
    return fail(h, GPMPC_ERR_ARG, "bad env dims");
  const long long total = (long long)h->st.ns * env->nx * H;
  const int threads = 128;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
  k_assemble<<<blocks, threads, 0, (cudaStream_t)stream>>>(*env, h->st.ns, H, h->st.T, xu, y_gp, out);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

}  // extern "C"

// ---- fused-horizon rollout (gpmpc_horizon.cuh) ---------------------------------------------------------------------------
// sample groups per CTA that fit (0: the step-wise path serves this shape)
static int horizon_groups(const gpmpc_handle* h, int n_steps) {
  const DevState& st = h->st;
  if (!h->fused_rollout || !h->condition || h->has_partial || h->grp_size > 0 || st.T != st.d + 1 || st.c != 0 || st.np != 0)
    return 0;
  int groups = std::min(HZ_MAX_WARPS / st.g_ny, 15);
  if (h->hz_groups_cap > 0) groups = std::min(groups, h->hz_groups_cap);
  for (; groups >= 1; --groups) {
    const HzLayout L = hz_layout(st.g_ny, st.n_real, st.m, st.mo, st.d, st.T, n_steps, groups);
    if ((size_t)L.total * 8 + 256 <= (size_t)h->max_dyn_smem) break;
  }
  if (groups < 1) return 0;
  // automatic: only while the step-wise rollout would be launch-latency bound (few samples per warp of the fused kernel)
  if (h->fused_rollout == 2 && (st.ns + h->num_sms * groups - 1) / (h->num_sms * groups) > HZ_AUTO_MAX_SERIAL) return 0;
  // small batches: one sample group per CTA over as many SMs as there are samples, rather than few full CTAs
  if (h->hz_groups_cap <= 0) groups = std::min(groups, std::max(1, (st.ns + h->num_sms - 1) / h->num_sms));
  return groups;
}

template <int D, int T>
static int launch_horizon_impl(gpmpc_handle* h, const gpmpc_env& env, const HorizonArgs& a, cudaStream_t stream) {
  auto kern = k_horizon<D, T>;
  CUDA_TRY(h, opt_in_smem(h, kern, h->max_dyn_smem, true));
  const DevState& st = h->st;
  const HzLayout L = hz_layout(st.g_ny, st.n_real, st.m, st.mo, D, T, a.n_steps, a.groups);
  const int n = a.s_end - a.s_begin;
  const int grid = std::max(1, std::min(h->num_sms, (n + a.groups - 1) / a.groups));
  kern<<<grid, a.groups * st.g_ny * 32, (size_t)L.total * 8, stream>>>(st, env, a);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

static int launch_horizon(gpmpc_handle* h, const gpmpc_env& env, const HorizonArgs& a, cudaStream_t stream) {
#define HZ_CASE(D_) case D_: return launch_horizon_impl<D_, D_ + 1>(h, env, a, stream);
  switch (h->st.d) { HZ_CASE(1) HZ_CASE(2) HZ_CASE(3) HZ_CASE(4) HZ_CASE(5) HZ_CASE(6) }
#undef HZ_CASE
  return fail(h, GPMPC_ERR_ARG, "unsupported d");
}

// bookkeeping after the fused launches: every element now holds n_steps fully observed points
static int horizon_commit(gpmpc_handle* h, int n_steps, cudaStream_t stream) {
  DevState& hst = h->st;
  k_fill_row_tables<<<(n_steps * hst.T + 127) / 128, 128, 0, stream>>>(hst, n_steps);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  double bytes = 0.0, flops = 0.0;
  for (int t = 0; t < n_steps; ++t) {
    hst.c = hst.T * t;
    count_work(h, 1, true);
    bytes += h->last_bytes;
    flops += h->last_flops;
  }
  h->last_bytes = bytes;
  h->last_flops = flops;
  h->c_before_point.clear();
  for (int t = 0; t < n_steps; ++t) h->c_before_point.push_back(hst.T * t);
  hst.np = n_steps;
  hst.c = hst.T * n_steps;
  h->factor_version++;
  h->last_rollout_fused = true;
  return GPMPC_OK;
}

static unsigned long long horizon_stagger(const gpmpc_handle* h, int samples_per_group) {
  if (h->hz_stagger_ns >= 0) return (unsigned long long)h->hz_stagger_ns;
  // automatic: the duration of one sample's horizon in the previous fused launch (0 on the very first launch)
  if (h->hz_last_ms <= 0.0 || h->hz_last_samples_per_group <= 0 || samples_per_group < 8) return 0ull;
  return (unsigned long long)(h->hz_last_ms * 1e6 / h->hz_last_samples_per_group);
}

extern "C" {

int gpmpc_linearise(gpmpc_handle* h, const gpmpc_env* env, const double* xu, int32_t xu_on_host, int32_t H, const double* eps,
                    const gpmpc_sample_opts* opts, int32_t reset_first, double* mean, double* var, double* y,
                    int32_t* jitter_level, double* out, double* out_host, void* stream_) {
  int rc = check_ready(h);
  if (rc) return rc;
  ON_HANDLE_DEVICE(h);
  if (!env || !xu || !eps || !opts || !mean || !var || !y || !out || H < 1) return fail(h, GPMPC_ERR_ARG, "null argument");
  if (env->g_ny != h->st.g_ny || env->d != h->st.d) return fail(h, GPMPC_ERR_ARG, "env does not match handle");
  if (h->grp_size > 0) return fail(h, GPMPC_ERR_STATE, "grouped handle");
  cudaStream_t stream = (cudaStream_t)stream_;
  const DevState& st = h->st;
  const int nz = env->nx + env->nu;
  const size_t n_xu = (size_t)st.ns * env->nx * H * nz, n_xg = (size_t)st.B * H * st.d;
  if (h->lin_xu_count < n_xu) {
    cudaFree(h->lin_xu);
    h->lin_xu_count = 0;
    CUDA_TRY(h, dev_alloc(&h->lin_xu, n_xu));
    h->lin_xu_count = n_xu;
  }
  if (h->lin_xg_count < n_xg) {
    cudaFree(h->lin_xg);
    h->lin_xg_count = 0;
    CUDA_TRY(h, dev_alloc(&h->lin_xg, n_xg));
    h->lin_xg_count = n_xg;
  }
  const double* xu_dev = xu;
  if (xu_on_host) {
    CUDA_TRY(h, cudaMemcpyAsync(h->lin_xu, xu, n_xu * 8, cudaMemcpyHostToDevice, stream));
    xu_dev = h->lin_xu;
  }
  {
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n_xg + threads - 1) / threads, (size_t)h->num_sms * 8);
    k_gather_gp_inputs<<<blocks, threads, 0, stream>>>(*env, st.ns, H, xu_dev, h->lin_xg);
    h->launches++;
  }
  h->want_real_only = reset_first != 0;
  h->prefactor_next = h->condition;  // the append below factorises Sigma* + noise: do it beside the draw
  rc = gpmpc_posterior(h, h->lin_xg, H, mean, var, eps, opts, y, jitter_level, stream_);
  if (rc) return rc;
  if (reset_first) {
    rc = gpmpc_reset_hallucinated(h);
    if (rc) return rc;
  }
  rc = gpmpc_append(h, h->lin_xg, y, nullptr, H, stream_);
  if (rc) return rc;
  rc = gpmpc_assemble(h, env, xu_dev, y, H, out, stream_);
  if (rc) return rc;
  if (out_host)
    CUDA_TRY(h, cudaMemcpyAsync(out_host, out, (size_t)st.ns * env->nx * H * (1 + nz) * 8, cudaMemcpyDeviceToHost, stream));
  return GPMPC_OK;
}

int gpmpc_rollout(gpmpc_handle* h, const gpmpc_env* env, const double* x0, const double* u_ff,
                  const double* eps, const gpmpc_sample_opts* opts, int32_t n_steps, double* traj,
                  void* stream_) {
  return gpmpc_rollout_gated(h, env, x0, u_ff, eps, opts, n_steps, traj, nullptr, stream_);
}

int gpmpc_rollout_gated(gpmpc_handle* h, const gpmpc_env* env, const double* x0, const double* u_ff,
                        const double* eps, const gpmpc_sample_opts* opts, int32_t n_steps, double* traj,
                        void* const* eps_ready, void* stream_) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!env || !x0 || !u_ff || !eps || !opts || !traj || n_steps < 1) return fail(h, GPMPC_ERR_ARG, "null argument");
  if (env->g_ny != h->st.g_ny || env->d != h->st.d) return fail(h, GPMPC_ERR_ARG, "env does not match handle");
  cudaStream_t stream = (cudaStream_t)stream_;
  DevState& hst = h->st;
  const int ns = hst.ns, nz = env->nx + env->nu;
  if (h->r_ns < ns) {
    cudaFree(h->r_xu); cudaFree(h->r_xstar); cudaFree(h->r_y);
    CUDA_TRY(h, dev_alloc(&h->r_xu, (size_t)ns * 2 * GPMPC_MAX_NX));
    CUDA_TRY(h, dev_alloc(&h->r_xstar, (size_t)hst.B * GPMPC_MAX_D));
    CUDA_TRY(h, dev_alloc(&h->r_y, (size_t)hst.B * GPMPC_MAX_T));
    h->r_ns = ns;
  }
  (void)nz;
  if (h->condition && hst.np + n_steps > hst.cap_points) {
    rc = alloc_factor_state(h, hst.np + n_steps, stream);
    if (rc) return rc;
  } else if (!h->condition && hst.np + n_steps > hst.cap_points) {
    rc = alloc_factor_state(h, hst.np + n_steps, stream);
    if (rc) return rc;
  }
  h->last_rollout_fused = false;
  const int hz_groups = horizon_groups(h, n_steps);
  if (hz_groups > 0) {
    // ONE launch for the whole horizon (gpmpc_horizon.cuh); base samples still on their way from the host are awaited first
    if (eps_ready)
      for (int t = 0; t < n_steps; ++t)
        if (eps_ready[t]) CUDA_TRY(h, cudaStreamWaitEvent(stream, (cudaEvent_t)eps_ready[t], 0));
    if (!h->hz_ev[0]) {
      CUDA_TRY(h, cudaEventCreate(&h->hz_ev[0]));
      CUDA_TRY(h, cudaEventCreate(&h->hz_ev[1]));
    } else if (h->hz_last_samples_per_group < 0) {
      // the previous fused launch was timed: its duration feeds the automatic stagger (the events are long complete when
      // a caller has read the trajectories; otherwise this waits for that launch)
      float ms = 0.f;
      if (cudaEventSynchronize(h->hz_ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, h->hz_ev[0], h->hz_ev[1]) == cudaSuccess)
        h->hz_last_ms = ms;
      h->hz_last_samples_per_group = -h->hz_last_samples_per_group;
    }
    HorizonArgs a{x0, u_ff, eps, traj, n_steps, 0, ns, hz_groups, 0ull, *opts};
    const int grid = std::max(1, std::min(h->num_sms, (ns + hz_groups - 1) / hz_groups));
    const int per_group = (ns + grid * hz_groups - 1) / (grid * hz_groups);
    a.stagger_ns = horizon_stagger(h, per_group);
    CUDA_TRY(h, cudaEventRecord(h->hz_ev[0], stream));
    rc = launch_horizon(h, *env, a, stream);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->hz_ev[1], stream));
    h->hz_last_samples_per_group = -per_group;  // negative: duration not read yet
    h->ev_used = 0;
    return horizon_commit(h, n_steps, stream);
  }
  const int threads = 128, blocks = (ns + threads - 1) / threads;
  const size_t eps_stride = (size_t)hst.B * hst.T;
  double bytes = 0.0, flops = 0.0;
  k_rollout_state<<<blocks, threads, 0, stream>>>(*env, ns, hst.T, 0, n_steps, 0, x0, u_ff, h->r_y, h->r_xu,
                                                   h->r_xstar, traj);
  h->launches++;
  if (h->timing) {
    while ((int)h->ev.size() < 2 * n_steps) {
      cudaEvent_t e;
      CUDA_TRY(h, cudaEventCreate(&e));
      h->ev.push_back(e);
    }
    h->ev_used = 0;
  }
  for (int t = 0; t < n_steps; ++t) {
    // base samples of step t may still be on their way from the host (copy stream): wait for their event here
    if (eps_ready && eps_ready[t]) CUDA_TRY(h, cudaStreamWaitEvent(stream, (cudaEvent_t)eps_ready[t], 0));
    if (h->timing) CUDA_TRY(h, cudaEventRecord(h->ev[2 * t], stream));
    rc = gpmpc_step(h, h->r_xstar, eps + (size_t)t * eps_stride, opts, nullptr, nullptr, h->r_y, nullptr, stream);
    if (rc) return rc;
    if (h->timing) {
      CUDA_TRY(h, cudaEventRecord(h->ev[2 * t + 1], stream));
      h->ev_used = 2 * (t + 1);
    }
    bytes += h->last_bytes;
    flops += h->last_flops;
    k_rollout_state<<<blocks, threads, 0, stream>>>(*env, ns, hst.T, t + 1, n_steps, 1, x0, u_ff, h->r_y,
                                                     h->r_xu, h->r_xstar, traj);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  h->last_bytes = bytes;
  h->last_flops = flops;
  return GPMPC_OK;
}

static int ensure_scratch(gpmpc_handle* h, size_t bytes) {
  if (bytes <= h->c_scratch_bytes) return GPMPC_OK;
  cudaFree(h->c_scratch);
  h->c_scratch = nullptr;
  h->c_scratch_bytes = 0;
  CUDA_TRY(h, cudaMalloc(&h->c_scratch, bytes));
  h->c_scratch_bytes = bytes;
  return GPMPC_OK;
}

int gpmpc_fs_advance(gpmpc_handle* h, const gpmpc_env* env, const double* xu, const double* y, const double* x_target,
                     double c_i, const double* u_next, int32_t* samples_left, double* x_next, double* xu_next, void* stream) {
  if (!h || !env || !xu || !y || !x_target || !samples_left || !x_next) return fail(h, GPMPC_ERR_ARG, "null argument");
  ON_HANDLE_DEVICE(h);
  if (env->g_ny != h->st.g_ny || env->nx > GPMPC_MAX_NX) return fail(h, GPMPC_ERR_ARG, "env does not match handle");
  if ((u_next == nullptr) != (xu_next == nullptr)) return fail(h, GPMPC_ERR_ARG, "u_next and xu_next go together");
  const int threads = 128;
  k_fs_advance<<<(h->st.ns + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(*env, h->st.ns, h->st.T, xu, y, x_target,
                                                                                       c_i, u_next, samples_left, x_next, xu_next);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

int gpmpc_min_dist_overwrite(gpmpc_handle* h, const double* x, int32_t H, const double* mean, const double* var,
                             double min_dist, double beta, double* y, void* stream) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || !y || H < 1 || (beta >= 0.0 && (!mean || !var))) return fail(h, GPMPC_ERR_ARG, "bad x / y / H / moments");
  const long long pairs = (long long)h->st.B * H;
  const int warps = 4;
  k_min_dist_overwrite<<<(unsigned)((pairs + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
      h->st, x, H, mean, var, min_dist, beta, y);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

int gpmpc_filter_new_points(gpmpc_handle* h, const double* x, int32_t H, double min_dist, int32_t use_hallucinated,
                            double* y, int32_t* counts, void* stream) {
  ON_HANDLE_DEVICE(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!x || !y || !counts || H < 1) return fail(h, GPMPC_ERR_ARG, "bad x / y / counts / H");
  CUDA_TRY(h, cudaMemsetAsync(counts, 0, (size_t)h->st.g_ny * H * 4, (cudaStream_t)stream));
  const long long pairs = (long long)h->st.B * H;
  const int warps = 4;
  k_filter_new_points<<<(unsigned)((pairs + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
      h->st, x, H, min_dist, use_hallucinated, y, counts);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

int gpmpc_pack_plin(gpmpc_handle* h, const gpmpc_env* env, const double* lin, const double* x_h, const double* tail,
                    int32_t n_tail, int32_t H, int32_t use_feedback_K, double* out, void* stream) {
  ON_HANDLE_DEVICE(h);
  if (!h || !env || !lin || !x_h || !out || H < 1 || n_tail < 0 || (n_tail > 0 && !tail))
    return fail(h, GPMPC_ERR_ARG, "null argument");
  if (env->nx < 1 || env->nx > GPMPC_MAX_NX || env->nu < 1 || env->nu > GPMPC_MAX_NX)
    return fail(h, GPMPC_ERR_ARG, "bad env dims");
  const int ns = h->st.ns, nx = env->nx, nu = env->nu;
  const long long total = ((long long)ns * (nx * nx + nx * nu + 2 * nx) + n_tail) * H;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)h->num_sms * 8);
  k_pack_plin<<<blocks, threads, 0, (cudaStream_t)stream>>>(ns, nx, nu, H, n_tail, use_feedback_K ? 1 : 0, *env, lin,
                                                             x_h, tail, out);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

int gpmpc_traj_stats(gpmpc_handle* h, const double* traj, int32_t ns, int32_t nx, int32_t H1, const double* ref,
                     double* box_min, double* box_max, double* max_dev, void* stream_) {
  ON_HANDLE_DEVICE(h);
  if (!h || !traj || ns < 1 || nx < 1 || H1 < 1) return fail(h, GPMPC_ERR_ARG, "bad traj / dims");
  if (max_dev && !ref) return fail(h, GPMPC_ERR_ARG, "max_dev needs a reference trajectory");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int cols = nx * H1, threads = 128, bx = (cols + threads - 1) / threads;
  // enough sample ranges to fill the machine a few times over; each reads its rows coalesced along (i, t)
  const int parts = std::max(1, std::min(ns, (h->num_sms * 8 + bx - 1) / bx));
  int rc = ensure_scratch(h, (size_t)parts * 3 * cols * 8);
  if (rc) return rc;
  double* part = (double*)h->c_scratch;
  k_traj_stats_partial<<<dim3(bx, parts), threads, 0, stream>>>(ns, cols, traj, ref, part);
  k_traj_stats_final<<<bx, threads, 0, stream>>>(parts, cols, part, box_min, box_max, max_dev);
  h->launches += 2;
  CUDA_TRY(h, cudaGetLastError());
  return GPMPC_OK;
}

// exact 2-D hull of a handful of points (Andrew's monotone chain, strict turns: collinear points are not vertices,
// like qhull's default).  pts sorted by (x, y, idx); returns positions into pts, counter-clockwise from the
// lexicographically smallest point.
namespace {
struct HullPt { double x, y; int idx; };
inline long double orient(const HullPt& a, const HullPt& b, const HullPt& c) {
  return ((long double)b.x - a.x) * ((long double)c.y - a.y) - ((long double)b.y - a.y) * ((long double)c.x - a.x);
}
std::vector<int> monotone_chain(std::vector<HullPt>& p) {
  std::sort(p.begin(), p.end(), [](const HullPt& a, const HullPt& b) {
    return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.idx < b.idx);
  });
  // identical coordinates: keep the lowest sample index
  size_t w = 0;
  for (size_t i = 0; i < p.size(); ++i)
    if (w == 0 || p[i].x != p[w - 1].x || p[i].y != p[w - 1].y) p[w++] = p[i];
  p.resize(w);
  const int n = (int)p.size();
  std::vector<int> hull;
  if (n <= 2) {
    for (int i = 0; i < n; ++i) hull.push_back(i);
    return hull;
  }
  hull.resize(2 * (size_t)n);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && orient(p[hull[k - 2]], p[hull[k - 1]], p[i]) <= 0) --k;
    hull[k++] = i;
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && orient(p[hull[k - 2]], p[hull[k - 1]], p[i]) <= 0) --k;
    hull[k++] = i;
  }
  hull.resize(k - 1);
  return hull;
}
}  // namespace

int gpmpc_hull2d(const double* xy, int32_t n, int32_t* hull_pos, int32_t* hull_n) {
  if (!xy || !hull_pos || !hull_n || n < 0) return GPMPC_ERR_ARG;
  std::vector<HullPt> p((size_t)n);
  for (int i = 0; i < n; ++i) p[i] = HullPt{xy[2 * i], xy[2 * i + 1], i};
  const std::vector<int> hv = monotone_chain(p);
  for (size_t k = 0; k < hv.size(); ++k) hull_pos[k] = p[hv[k]].idx;
  *hull_n = (int32_t)hv.size();
  return GPMPC_OK;
}

int gpmpc_stage_hulls(gpmpc_handle* h, const double* traj, int32_t ns, int32_t nx, int32_t H1, int32_t i0, int32_t i1,
                      int32_t max_vertices, int32_t* hull_idx, int32_t* hull_n, void* stream_) {
  ON_HANDLE_DEVICE(h);
  if (!h || !traj || !hull_idx || !hull_n || ns < 1 || nx < 1 || H1 < 1 || max_vertices < 1 || i0 < 0 || i1 < 0 ||
      i0 >= nx || i1 >= nx || i0 == i1)
    return fail(h, GPMPC_ERR_ARG, "bad traj / dims / coordinate pair");
  cudaStream_t stream = (cudaStream_t)stream_;
  HullDirs dirs;
  for (int k = 0; k < HULL_DIRS; ++k) {  // counter-clockwise
    dirs.cx[k] = std::cos(2.0 * M_PI * k / HULL_DIRS);
    dirs.cy[k] = std::sin(2.0 * M_PI * k / HULL_DIRS);
  }
  const int bx = (H1 + 31) / 32, ty = 8;
  const int parts = std::max(1, std::min((ns + ty - 1) / ty, (h->num_sms * 4 + bx - 1) / bx));
  // survivors per stage: the polygon of HULL_DIRS extremes leaves a thin rim; start generous, grow if it overflows
  int cap = std::max(4096, std::min(ns, ns / 8 + 1024));
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t n_ext = (size_t)H1 * HULL_DIRS;
    const size_t off_pval = 0, off_poly = off_pval + (size_t)parts * n_ext * 8, off_cxy = off_poly + n_ext * 16,
                 off_pidx = off_cxy + (size_t)H1 * cap * 16, off_ext = off_pidx + (size_t)parts * n_ext * 4,
                 off_cand = off_ext + n_ext * 4, off_cnt = off_cand + (size_t)H1 * cap * 4,
                 total = off_cnt + (size_t)H1 * 4;
    int rc = ensure_scratch(h, total);
    if (rc) return rc;
    char* base = (char*)h->c_scratch;
    double* pval = (double*)(base + off_pval);
    double* poly = (double*)(base + off_poly);
    double* cxy = (double*)(base + off_cxy);
    int* pidx = (int*)(base + off_pidx);
    int* ext = (int*)(base + off_ext);
    int* cand = (int*)(base + off_cand);
    int* cnt = (int*)(base + off_cnt);
    CUDA_TRY(h, cudaMemsetAsync(cnt, 0, (size_t)H1 * 4, stream));
    k_hull_extremes_partial<<<dim3(bx, parts), dim3(32, ty), 0, stream>>>(ns, nx, H1, i0, i1, dirs, traj, pval, pidx);
    k_hull_extremes_final<<<(unsigned)((n_ext + 127) / 128), 128, 0, stream>>>(parts, nx, H1, i0, i1, pval, pidx, traj,
                                                                              poly, ext);
    k_hull_filter<<<dim3(bx, parts), dim3(32, ty), 0, stream>>>(ns, nx, H1, i0, i1, cap, traj, poly, cand, cxy, cnt);
    h->launches += 3;
    CUDA_TRY(h, cudaGetLastError());
    std::vector<int> hc((size_t)H1), hext(n_ext);
    std::vector<double> hpoly(n_ext * 2);
    CUDA_TRY(h, cudaMemcpyAsync(hc.data(), cnt, (size_t)H1 * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(h, cudaMemcpyAsync(hext.data(), ext, n_ext * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(h, cudaMemcpyAsync(hpoly.data(), poly, n_ext * 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(h, cudaStreamSynchronize(stream));
    const int worst = *std::max_element(hc.begin(), hc.end());
    if (worst > cap) {  // degenerate cloud (e.g. collinear): every point survived the filter
      if (attempt == 1 || cap >= ns) return fail(h, GPMPC_ERR_CAPACITY, "hull candidate list overflow");
      cap = ns;
      continue;
    }
    // host staging sized by the longest list actually produced (the capacity is ~ns / 8 per stage: 100+ MB of vectors per call
    // at 10^6 samples when sized by it)
    const size_t wl = (size_t)std::max(worst, 1);
    std::vector<int> hcand((size_t)H1 * wl);
    std::vector<double> hxy((size_t)H1 * wl * 2);
    if (worst > 0) {
      CUDA_TRY(h, cudaMemcpy2DAsync(hcand.data(), wl * 4, cand, (size_t)cap * 4, (size_t)worst * 4, H1,
                                    cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(h, cudaMemcpy2DAsync(hxy.data(), wl * 16, cxy, (size_t)cap * 16, (size_t)worst * 16, H1,
                                    cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(h, cudaStreamSynchronize(stream));
    }
    for (int t = 0; t < H1; ++t) {
      std::vector<HullPt> p;
      p.reserve((size_t)hc[t] + HULL_DIRS);
      for (int k = 0; k < hc[t]; ++k)
        p.push_back(HullPt{hxy[((size_t)t * wl + k) * 2], hxy[((size_t)t * wl + k) * 2 + 1], hcand[(size_t)t * wl + k]});
      for (int k = 0; k < HULL_DIRS; ++k)
        p.push_back(HullPt{hpoly[((size_t)t * HULL_DIRS + k) * 2], hpoly[((size_t)t * HULL_DIRS + k) * 2 + 1],
                           hext[(size_t)t * HULL_DIRS + k]});
      const std::vector<int> hv = monotone_chain(p);
      if ((int)hv.size() > max_vertices) return fail(h, GPMPC_ERR_CAPACITY, "more hull vertices than max_vertices");
      hull_n[t] = (int32_t)hv.size();
      for (size_t k = 0; k < hv.size(); ++k) hull_idx[(size_t)t * max_vertices + k] = p[hv[k]].idx;
    }
    return GPMPC_OK;
  }
  return fail(h, GPMPC_ERR_CAPACITY, "hull candidate list overflow");
}

int32_t gpmpc_num_hallucinated(const gpmpc_handle* h) { return h ? h->st.np : -1; }
int32_t gpmpc_num_factor_rows(const gpmpc_handle* h) { return h ? h->st.c : -1; }
int32_t gpmpc_num_real_observed(const gpmpc_handle* h) { return h ? h->st.m : -1; }

int gpmpc_export_hallucinated(const gpmpc_handle* h_, double* X, double* Y, void* stream) {
  ON_HANDLE_DEVICE(h_);
  gpmpc_handle* h = const_cast<gpmpc_handle*>(h_);
  if (!h || !X || !Y) return fail(h, GPMPC_ERR_ARG, "null argument");
  const DevState& st = h->st;
  if (st.np == 0) return GPMPC_OK;
  CUDA_TRY(h, cudaMemcpy2DAsync(X, (size_t)st.np * st.d * 8, st.Xh, (size_t)st.cap_points * st.d * 8,
                                (size_t)st.np * st.d * 8, st.B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  CUDA_TRY(h, cudaMemcpy2DAsync(Y, (size_t)st.np * st.T * 8, st.Yh, (size_t)st.cap_points * st.T * 8,
                                (size_t)st.np * st.T * 8, st.B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GPMPC_OK;
}

int gpmpc_export_point_states(const gpmpc_handle* h_, uint8_t* out, void* stream) {
  gpmpc_handle* h = const_cast<gpmpc_handle*>(h_);
  if (!h || !out) return fail(h, GPMPC_ERR_ARG, "null argument");
  ON_HANDLE_DEVICE(h);
  const DevState& st = h->st;
  if (st.np == 0) return GPMPC_OK;
  if (!st.pstate) {
    CUDA_TRY(h, cudaMemsetAsync(out, 0, (size_t)st.B * st.np, (cudaStream_t)stream));
    return GPMPC_OK;
  }
  CUDA_TRY(h, cudaMemcpy2DAsync(out, (size_t)st.np, st.pstate, (size_t)st.cap_points, (size_t)st.np, st.B,
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GPMPC_OK;
}

int gpmpc_status(gpmpc_handle* h, uint32_t* status, int32_t clear, void* stream) {
  ON_HANDLE_DEVICE(h);
  if (!h || !status) return fail(h, GPMPC_ERR_ARG, "null argument");
  // through a pinned staging word: a device->pageable copy goes through the driver's bounce buffer and costs ~10 us more per
  // SQP linearisation (the status is read once per linearisation, behind the outputs' own copy)
  if (!h->status_pinned) CUDA_TRY(h, cudaHostAlloc((void**)&h->status_pinned, sizeof(uint32_t), cudaHostAllocDefault));
  CUDA_TRY(h, cudaMemcpyAsync(h->status_pinned, h->st.status, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (clear) CUDA_TRY(h, cudaMemsetAsync(h->st.status, 0, 4, (cudaStream_t)stream));
  CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream));
  *status = *h->status_pinned;
  return GPMPC_OK;
}

int64_t gpmpc_state_bytes(const gpmpc_handle* h) {
  if (!h) return -1;
  const DevState& st = h->st;
  const int64_t B = st.B;
  return 8 * (B * (int64_t)st.elem_stride + B * st.c_cap + B * st.cap_points * (int64_t)(st.d + st.T)) +
         8 * (int64_t)st.g_ny * ((int64_t)st.m * st.m + (int64_t)subpanel_off((st.m + 7) / 8, 0) + st.m);
}

int gpmpc_last_launch_work(const gpmpc_handle* h, double* bytes, double* flops) {
  if (!h) return GPMPC_ERR_ARG;
  if (bytes) *bytes = h->last_bytes;
  if (flops) *flops = h->last_flops;
  return GPMPC_OK;
}

int64_t gpmpc_launch_count(const gpmpc_handle* h) { return h ? h->launches : -1; }

int gpmpc_set_block_kernels(gpmpc_handle* h, int32_t mma) {
  if (!h) return fail(h, GPMPC_ERR_ARG, "null handle");
  h->block_mma = mma != 0;
  return GPMPC_OK;
}

int gpmpc_set_grouping(gpmpc_handle* h, int32_t group_size, double min_dist) {
  if (!h || group_size < 0) return fail(h, GPMPC_ERR_ARG, "bad group size");
  ON_HANDLE_DEVICE(h);
  DevState& st = h->st;
  if (st.np != 0) return fail(h, GPMPC_ERR_STATE, "set the grouping on an empty hallucinated set (gpmpc_reset_hallucinated first)");
  if (group_size > 0 && st.T == 1 && !h->condition) return fail(h, GPMPC_ERR_ARG, "grouping needs conditioning");
  cudaFree(h->grp_flags); cudaFree(h->grp_decision); cudaFree(st.pstate);
  h->grp_flags = h->grp_decision = nullptr;
  st.pstate = nullptr;
  h->grp_size = group_size;
  h->grp_min_dist = min_dist;
  if (group_size == 0) return GPMPC_OK;
  CUDA_TRY(h, dev_alloc(&h->grp_flags, (size_t)st.B));
  CUDA_TRY(h, dev_alloc(&h->grp_decision, (size_t)(st.ns + group_size - 1) / group_size));
  if (st.Lh || st.Xh) {  // the per-element state exists already: add the point states
    CUDA_TRY(h, dev_alloc(&st.pstate, (size_t)st.B * std::max(st.cap_points, 1)));
    CUDA_TRY(h, cudaMemset(st.pstate, 0, (size_t)st.B * std::max(st.cap_points, 1)));
  }
  return GPMPC_OK;
}

int gpmpc_set_option(gpmpc_handle* h, const char* name, int64_t value) {
  if (!h || !name) return fail(h, GPMPC_ERR_ARG, "null argument");
  const std::string n(name);
  if (n == "rollout_fused") h->fused_rollout = (int)std::max<int64_t>(0, std::min<int64_t>(2, value));
  else if (n == "hz_groups") h->hz_groups_cap = (int)value;
  else if (n == "step_grid_cap") h->step_grid_cap = (int)value;
  else if (n == "step_warps_cap") h->step_warps_cap = (int)value;
  else if (n == "sr_grid_cap") h->sr_grid_cap = (int)value;
  else if (n == "prefactor_next") h->prefactor_next = value != 0 && h->condition;
  else if (n == "hz_stagger_ns") h->hz_stagger_ns = value;
  else if (n == "force_wo") h->force_wo = value != 0;
  else if (n == "wo_slab_nb3") h->wo_slab_nb3 = value != 0;
  else if (n == "force_block_fallback") h->force_block_fallback = value != 0;
  else if (n == "force_big") h->force_big = value != 0;
  else if (n == "big_slab_cap") h->big_slab_cap = (int)value & ~7;
  else return fail(h, GPMPC_ERR_ARG, "unknown option " + n);
  return GPMPC_OK;
}

int gpmpc_get_option(gpmpc_handle* h, const char* name, int64_t* value) {
  if (!h || !name || !value) return fail(h, GPMPC_ERR_ARG, "null argument");
  const std::string n(name);
  if (n == "rollout_fused") *value = h->fused_rollout;
  else if (n == "last_rollout_fused") *value = h->last_rollout_fused ? 1 : 0;
  else if (n == "hz_groups") *value = h->hz_groups_cap;
  else if (n == "step_grid_cap") *value = h->step_grid_cap;
  else if (n == "step_warps_cap") *value = h->step_warps_cap;
  else if (n == "hz_stagger_ns") *value = h->hz_stagger_ns;
  else if (n == "force_wo") *value = h->force_wo;
  else if (n == "force_block_fallback") *value = h->force_block_fallback;
  else if (n == "force_big") *value = h->force_big;
  else if (n == "big_slab_cap") *value = h->big_slab_cap;
  else return fail(h, GPMPC_ERR_ARG, "unknown option " + n);
  return GPMPC_OK;
}

int gpmpc_set_timing(gpmpc_handle* h, int32_t on) {
  if (!h) return GPMPC_ERR_ARG;
  h->timing = on != 0;
  return GPMPC_OK;
}

int gpmpc_rollout_kernel_ms(gpmpc_handle* h, double* total_ms, int32_t* launches) {
  ON_HANDLE_DEVICE(h);
  if (!h || !total_ms || !launches) return fail(h, GPMPC_ERR_ARG, "null argument");
  *total_ms = 0.0;
  *launches = h->ev_used / 2;
  if (h->last_rollout_fused && h->hz_ev[1]) {  // the fused-horizon rollout is one launch
    float ms = 0.f;
    CUDA_TRY(h, cudaEventSynchronize(h->hz_ev[1]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->hz_ev[0], h->hz_ev[1]));
    *total_ms = ms;
    *launches = 1;
    h->hz_last_ms = ms;
    if (h->hz_last_samples_per_group < 0) h->hz_last_samples_per_group = -h->hz_last_samples_per_group;
    return GPMPC_OK;
  }
  if (h->ev_used == 0) return GPMPC_OK;
  CUDA_TRY(h, cudaEventSynchronize(h->ev[h->ev_used - 1]));
  for (int i = 0; i < h->ev_used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    *total_ms += ms;
  }
  return GPMPC_OK;
}

}  // extern "C"
