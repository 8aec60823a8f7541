// K1: fused rollout step (H = 1).  One WARP per batch element (sample s, output j); the warps of a CTA
// share output j, so the shared real-data factor L_oo (packed, column-major), the observed real inputs and
// beta_o are staged once per CTA in shared memory.  Per element and step, in ONE pass:
//
//   A  kernel vector k(x*, X) against real + hallucinated scalars (one exp per lane-owned scalar)
//   B  forward substitution against the shared factor (column sweep in shared memory)
//   C  forward substitution against the element's own bordered rows, streamed ONCE from HBM with
//      coalesced 16-byte loads, 4 rows in flight per warp, warp-shuffle reduction of the 4*T dot products
//   D  Sigma* = K** - w^T w and mean = w^T beta accumulated on the fly (registers, redundantly per lane)
//   E  T x T Cholesky with GPyTorch's jitter ladder, y = mean + L eps, zero-variance / truncation
//   F  rank-T append: new rows [w^T | chol(Sigma* + noise)] written back coalesced, beta_h, data set
//
// HBM traffic per element-step = its factor rows (read once) + T new rows (written once) + O(T) I/O:
// this kernel is HBM-bound (roofline section); algorithmic bytes are counted by the host.
#pragma once
#include "gpmpc_state.cuh"

#define STEP_WARPS 4
#define STEP_RB 4  // own rows in flight per warp

template <int T>
struct TriT {  // lower-triangular T x T in registers
  double v[T * (T + 1) / 2];
  __device__ __forceinline__ double& at(int r, int s) { return v[r * (r + 1) / 2 + s]; }
};

// in-register Cholesky (every lane does the same arithmetic); returns false on a non-positive / NaN pivot
template <int T>
__device__ __forceinline__ bool chol_T(const TriT<T>& S, double add, TriT<T>& L) {
  bool ok = true;
#pragma unroll
  for (int r = 0; r < T; ++r) {
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      double v = S.v[r * (r + 1) / 2 + s];
      if (s == r) v += add;
#pragma unroll
      for (int k = 0; k < s; ++k) v -= L.v[r * (r + 1) / 2 + k] * L.v[s * (s + 1) / 2 + k];
      if (s == r) {
        if (!(v > 0.0)) ok = false;
        L.v[r * (r + 1) / 2 + s] = sqrt(v);
      } else {
        L.v[r * (r + 1) / 2 + s] = v / L.v[s * (s + 1) / 2 + s];
      }
    }
  }
  return ok;
}

template <int D, int T>
__global__ void __launch_bounds__(STEP_WARPS * 32)
k_step(DevState st, const double* __restrict__ x, const double* __restrict__ eps, gpmpc_sample_opts opts,
       double* __restrict__ mean, double* __restrict__ var, double* __restrict__ y,
       int* __restrict__ jitter_level, int grow_factor, int n_pad, int loo_in_smem) {
  extern __shared__ double smem[];
  const int j = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s_idx = blockIdx.x * STEP_WARPS + warp;
  const int m = st.m, c = st.c, n = m + c;
  const size_t tri = (size_t)m * (m + 1) / 2;

  // ---- CTA-shared tables ---------------------------------------------------------------------
  double* sLT = smem;                               // [tri] (only if loo_in_smem)
  double* sXo = sLT + (loo_in_smem ? tri : 0);      // [m][D] input of observed real scalar i
  double* sBo = sXo + (size_t)m * D;                // [m]
  int* sTo = (int*)(sBo + m);                       // [m] task of observed real scalar i
  double* wbase = (double*)(sTo + ((m + 1) & ~1));  // per-warp w: [T][n_pad]
  const double* gLT = st.LooT + (size_t)j * tri;
  if (loo_in_smem)
    for (size_t i = threadIdx.x; i < tri; i += blockDim.x) sLT[i] = gLT[i];
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double* xp = st.Xr + (size_t)st.obs_pt[i] * D;
#pragma unroll
    for (int a = 0; a < D; ++a) sXo[i * D + a] = xp[a];
    sBo[i] = st.beta_o[(size_t)j * m + i];
    sTo[i] = st.obs_task[i];
  }
  __syncthreads();
  if (s_idx >= st.ns) return;  // no block-level sync below this line
  const double* LT = loo_in_smem ? sLT : gLT;
  const int b = s_idx * st.g_ny + j;
  double* w = wbase + (size_t)warp * T * n_pad;

  double ls[D], xs[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    ls[a] = st.ls[j * D + a];
    xs[a] = x[(size_t)b * D + a];
  }
  const double os = st.os[j];

  // ---- A: kernel vector ------------------------------------------------------------------------
  for (int i = lane; i < n; i += 32) {
    const double* xa;
    int ta;
    if (i < m) {
      xa = sXo + i * D;
      ta = sTo[i];
    } else {
      int k = i - m;
      xa = st.Xh + ((size_t)b * st.cap_points + st.hobs_pt[k]) * D;
      ta = st.hobs_task[k];
    }
    double g[D], sq = 0.0, ga = 0.0, la2 = 1.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double r = xa[a] - xs[a];
      double t = r / ls[a];
      sq += t * t;
      g[a] = t / ls[a];  // r_a / l_a^2
      if (a == ta - 1) { ga = g[a]; la2 = ls[a] * ls[a]; }
    }
    double k0 = os * exp(-0.5 * sq);
    if (ta == 0) {
      w[i] = k0;
#pragma unroll
      for (int tb = 1; tb < T; ++tb) w[tb * n_pad + i] = k0 * g[tb - 1];
    } else {
      w[i] = -k0 * ga;
#pragma unroll
      for (int tb = 1; tb < T; ++tb) {
        double h = -ga * g[tb - 1];
        if (tb == ta) h += 1.0 / la2;
        w[tb * n_pad + i] = k0 * h;
      }
    }
  }

  TriT<T> Sacc;
  double macc[T];
#pragma unroll
  for (int i = 0; i < T * (T + 1) / 2; ++i) Sacc.v[i] = 0.0;
#pragma unroll
  for (int r = 0; r < T; ++r) macc[r] = 0.0;

  // ---- B: shared block, column sweep -------------------------------------------------------------
  for (int jj = 0; jj < m; ++jj) {
    __syncwarp();
    const double* col = LT + packed_col(jj, m);
    const double dj = col[0];
    double wj[T];
#pragma unroll
    for (int r = 0; r < T; ++r) wj[r] = w[r * n_pad + jj] / dj;
    const double bo = sBo[jj];
#pragma unroll
    for (int r = 0; r < T; ++r) {
      macc[r] += wj[r] * bo;
#pragma unroll
      for (int s = 0; s <= r; ++s) Sacc.at(r, s) += wj[r] * wj[s];
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < T; ++r) w[r * n_pad + jj] = wj[r];
    }
    for (int i = jj + 1 + lane; i < m; i += 32) {
      const double l = col[i - jj];
#pragma unroll
      for (int r = 0; r < T; ++r) w[r * n_pad + i] -= l * wj[r];
    }
  }
  __syncwarp();

  // ---- C: own bordered rows, streamed from HBM ---------------------------------------------------
  const double* Lb = st.Lh + (size_t)b * st.c_cap * st.ldL;
  const double* bh = st.beta_h + (size_t)b * st.c_cap;
  for (int i0 = 0; i0 < c; i0 += STEP_RB) {
    const int nrow = min(STEP_RB, c - i0);
    const int len = m + i0;  // prefix whose w is final
    double acc[STEP_RB][T];
#pragma unroll
    for (int a = 0; a < STEP_RB; ++a)
#pragma unroll
      for (int r = 0; r < T; ++r) acc[a][r] = 0.0;
    for (int k = 2 * lane; k < len; k += 64) {
      const bool two = (k + 1 < len);
      double2 l2[STEP_RB];
#pragma unroll
      for (int a = 0; a < STEP_RB; ++a)
        if (a < nrow) l2[a] = *reinterpret_cast<const double2*>(Lb + (size_t)(i0 + a) * st.ldL + k);
        else l2[a] = make_double2(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < T; ++r) {
        const double w0 = w[r * n_pad + k];
        const double w1 = two ? w[r * n_pad + k + 1] : 0.0;
#pragma unroll
        for (int a = 0; a < STEP_RB; ++a) acc[a][r] += l2[a].x * w0 + l2[a].y * w1;
      }
    }
#pragma unroll
    for (int a = 0; a < STEP_RB; ++a)
#pragma unroll
      for (int r = 0; r < T; ++r) acc[a][r] = warp_sum(acc[a][r]);
    // resolve the nrow x nrow triangular corner (uniform over lanes)
    double wn[STEP_RB][T];
#pragma unroll
    for (int a = 0; a < STEP_RB; ++a) {
      if (a < nrow) {
        const double* row = Lb + (size_t)(i0 + a) * st.ldL + len;
        double v[T];
#pragma unroll
        for (int r = 0; r < T; ++r) v[r] = w[r * n_pad + len + a] - acc[a][r];
#pragma unroll
        for (int bb = 0; bb < a; ++bb) {
          const double l = row[bb];
#pragma unroll
          for (int r = 0; r < T; ++r) v[r] -= l * wn[bb][r];
        }
        const double dg = row[a];
        const double be = bh[i0 + a];
#pragma unroll
        for (int r = 0; r < T; ++r) {
          wn[a][r] = v[r] / dg;
          macc[r] += wn[a][r] * be;
        }
#pragma unroll
        for (int r = 0; r < T; ++r)
#pragma unroll
          for (int s = 0; s <= r; ++s) Sacc.at(r, s) += wn[a][r] * wn[a][s];
      }
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < STEP_RB; ++a)
        if (a < nrow) {
#pragma unroll
          for (int r = 0; r < T; ++r) w[r * n_pad + len + a] = wn[a][r];
        }
    }
    __syncwarp();
  }

  // ---- D: posterior moments ----------------------------------------------------------------------
  TriT<T> S;
#pragma unroll
  for (int r = 0; r < T; ++r)
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      double kss = 0.0;
      if (r == s) kss = (r == 0) ? os : os / (ls[r > 0 ? r - 1 : 0] * ls[r > 0 ? r - 1 : 0]);
      S.at(r, s) = kss - Sacc.at(r, s);
    }
  double vr[T];
#pragma unroll
  for (int r = 0; r < T; ++r) vr[r] = fmax(S.at(r, r), GP_MIN_VARIANCE);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) {
      if (mean) mean[(size_t)b * T + r] = macc[r];
      if (var) var[(size_t)b * T + r] = vr[r];
    }
  }
  if (!eps) return;

  // ---- E: draw ------------------------------------------------------------------------------------
  TriT<T> Lc;
  int level = 0;
  if (T == 1) {
    Lc.v[0] = opts.unclamped_sqrt_1x1 ? sqrt(S.v[0]) : sqrt(fmax(S.v[0], 0.0));
  } else {
    bool ok = chol_T<T>(S, 0.0, Lc);
    double jit = st.jitter;
    while (!ok && level < GP_MAX_TRIES) {
      ++level;
      ok = chol_T<T>(S, jit, Lc);
      jit *= 10.0;
    }
    if (!ok) level = 4;
  }
  double yv[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    double acc = macc[r];
#pragma unroll
    for (int s = 0; s <= r; ++s) acc += Lc.at(r, s) * eps[(size_t)b * T + s];
    yv[r] = level < 4 ? acc : nan("");
  }
  bool zero = opts.variance_is_zero >= 0.0;
#pragma unroll
  for (int r = 0; r < T; ++r) zero = zero && (vr[r] <= opts.variance_is_zero);
#pragma unroll
  for (int r = 0; r < T; ++r) {
    if (zero) yv[r] = macc[r];
    if (opts.beta >= 0.0) {
      const double sd = sqrt(vr[r]);
      yv[r] = fmin(fmax(yv[r], macc[r] - opts.beta * sd), macc[r] + opts.beta * sd);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) y[(size_t)b * T + r] = yv[r];
    if (jitter_level) jitter_level[b] = level;
    if (level == 4) atomicOr(st.status, GPMPC_ST_SAMPLE_NOT_PD);
  }

  // ---- F: condition on (x*, y) --------------------------------------------------------------------
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < D; ++a) st.Xh[((size_t)b * st.cap_points + st.np) * D + a] = xs[a];
#pragma unroll
    for (int r = 0; r < T; ++r) st.Yh[((size_t)b * st.cap_points + st.np) * T + r] = yv[r];
  }
  if (!grow_factor) return;
  TriT<T> Sn = S, Ln;
#pragma unroll
  for (int r = 0; r < T; ++r) Sn.at(r, r) += st.noise[j * T + r];
  if (!chol_T<T>(Sn, 0.0, Ln)) {
    if (lane == 0) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
  }
  double* Lnew = st.Lh + ((size_t)b * st.c_cap + c) * st.ldL;
  double bn[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    for (int k = lane; k < n; k += 32) Lnew[(size_t)r * st.ldL + k] = w[r * n_pad + k];
    double t = yv[r] - macc[r];
#pragma unroll
    for (int s = 0; s < r; ++s) t -= Ln.at(r, s) * bn[s];
    bn[r] = t / Ln.at(r, r);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) {
#pragma unroll
      for (int s = 0; s <= r; ++s) Lnew[(size_t)r * st.ldL + n + s] = Ln.at(r, s);
      st.beta_h[(size_t)b * st.c_cap + c + r] = bn[r];
    }
    if (b == 0) {
#pragma unroll
      for (int r = 0; r < T; ++r) {
        st.hobs_pt[c + r] = st.np;
        st.hobs_task[c + r] = r;
      }
    }
  }
}
