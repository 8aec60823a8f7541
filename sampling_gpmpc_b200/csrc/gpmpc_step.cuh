// K1: fused rollout step (H = 1).  One WARP per batch element (sample s, output j); the warps of a CTA
// share output j, so the shared real-data factor L_oo (packed, column-major), the observed real inputs and
// beta_o are staged once per CTA in shared memory.  Per element and step, in ONE pass:
//
//   A  kernel vector k(x*, X) against real + hallucinated scalars (one exp per lane-owned scalar)
//   B  forward substitution against the shared factor (column sweep in shared memory)
//   C  forward substitution against the element's own bordered rows, streamed ONCE from HBM with
//      coalesced 16-byte loads, RB rows in flight per warp; the RB*T partial dot products are transposed
//      through shared memory (lane v finishes dot product v), lanes r < T resolve the RB x RB corner
//   D  Sigma* = K** - w^T w and mean = w^T beta: one pass over w + warp-shuffle all-reduce
//   E  T x T Cholesky with GPyTorch's jitter ladder, y = mean + L eps, zero-variance / truncation
//   F  rank-T append: new rows [w^T | chol(Sigma* + noise)] written back coalesced, beta_h, data set
//
// HBM traffic per element-step = its factor rows (read once) + T new rows (written once) + O(T) I/O:
// this kernel is HBM-bound (roofline section); algorithmic bytes are counted by the host.
#pragma once
#include "gpmpc_state.cuh"

#define STEP_WARPS 4

// own rows streamed per block (all in flight at once); RB*T partial dot products must fit one per lane
template <int T> struct StepRB { static constexpr int value = T == 1 ? 8 : T == 2 ? 8 : T == 3 ? 6 : T == 4 ? 4 : T == 5 ? 5 : T == 6 ? 5 : 4; };
#define STEP_RED_LD 33  // row stride of the per-warp reduction scratch (conflict-free column reads)

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
  constexpr int RB = StepRB<T>::value;
  constexpr int NV = RB * T;
  extern __shared__ __align__(16) double smem[];
  const int j = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s_idx = blockIdx.x * STEP_WARPS + warp;
  const int m = st.m, c = st.c, n = m + c;
  const size_t tri = (size_t)m * (m + 1) / 2;
  const size_t tri_pad = (tri + 1) & ~(size_t)1;
  const int m_pad = (m + 1) & ~1;

  // ---- CTA-shared tables (every table starts 16-byte aligned) ------------------------------------
  double* sLT = smem;                                  // [tri_pad] (only if loo_in_smem); diagonal = 1/L_jj
  double* sXo = sLT + (loo_in_smem ? tri_pad : 0);     // [m_pad*D] input of observed real scalar i
  double* sBo = sXo + (size_t)m_pad * D;               // [m_pad]
  int* sTo = (int*)(sBo + m_pad);                      // [2*m_pad] ints: task of observed real scalar i
  double* wbase = (double*)(sTo + 2 * m_pad);          // per-warp: w [T][n_pad], red [NV][33], tot [32]
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
  const int per_warp = T * n_pad + NV * STEP_RED_LD + 32 + ((NV * STEP_RED_LD) & 1);
  double* w = wbase + (size_t)warp * per_warp;
  double* red = w + T * n_pad;
  double* tot = red + NV * STEP_RED_LD + ((NV * STEP_RED_LD) & 1);

  double ls[D], il[D], xs[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    ls[a] = st.ls[j * D + a];
    il[a] = 1.0 / ls[a];
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
    double g[D], sq = 0.0, ga = 0.0, il2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double r = xa[a] - xs[a];
      double t = r * il[a];
      sq = fma(t, t, sq);
      g[a] = t * il[a];  // r_a / l_a^2
      if (a == ta - 1) { ga = g[a]; il2 = il[a] * il[a]; }
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
        if (tb == ta) h += il2;
        w[tb * n_pad + i] = k0 * h;
      }
    }
  }
  __syncwarp();

  // ---- B: shared block, column sweep; entry jj is finalised (scaled by 1/L_jj) by the lane that
  //         applies column jj-1 to it, so each column costs one warp sync --------------------------------
  if (lane < T) w[lane * n_pad] *= LT[0];
  for (int jj = 0; jj + 1 < m; ++jj) {
    __syncwarp();
    const double* col = LT + packed_col(jj, m);
    double wj[T];
#pragma unroll
    for (int r = 0; r < T; ++r) wj[r] = w[r * n_pad + jj];
    const double rd_next = col[m - jj];  // = LT[packed_col(jj + 1, m)]: reciprocal diagonal of column jj+1
    for (int i = jj + 1 + lane; i < m; i += 32) {
      const double l = col[i - jj];
      const double sc = (i == jj + 1) ? rd_next : 1.0;
#pragma unroll
      for (int r = 0; r < T; ++r) w[r * n_pad + i] = (w[r * n_pad + i] - l * wj[r]) * sc;
    }
  }
  __syncwarp();

  // ---- C: own bordered rows, streamed from HBM ---------------------------------------------------
  const double* Lb = st.Lh + (size_t)b * st.c_cap * st.ldL;
  const size_t ldL = st.ldL;
  int i0 = 0;
  for (; i0 + RB <= c; i0 += RB) {
    const int len = m + i0;  // prefix whose w is final
    const double* rows = Lb + (size_t)i0 * ldL;
    double acc[RB][T];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
      for (int r = 0; r < T; ++r) acc[a][r] = 0.0;
    for (int k = 2 * lane; k < len; k += 64) {
      double2 wv[T];
#pragma unroll
      for (int r = 0; r < T; ++r) {
        wv[r] = *reinterpret_cast<const double2*>(w + r * n_pad + k);
        if (k + 1 >= len) wv[r].y = 0.0;  // entry `len` is not final yet
      }
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const double2 l2 = *reinterpret_cast<const double2*>(rows + a * ldL + k);
#pragma unroll
        for (int r = 0; r < T; ++r) acc[a][r] = fma(l2.x, wv[r].x, fma(l2.y, wv[r].y, acc[a][r]));
      }
    }
    // transpose-reduce through shared memory: lane v sums partial dot product v over the 32 lanes
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
      for (int r = 0; r < T; ++r) red[(a * T + r) * STEP_RED_LD + lane] = acc[a][r];
    __syncwarp();
    if (lane < NV) {
      const double* rr = red + lane * STEP_RED_LD;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        s0 += rr[i];
        s1 += rr[i + 1];
        s2 += rr[i + 2];
        s3 += rr[i + 3];
      }
      tot[lane] = (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
    if (lane < T) {  // lane r resolves the RB x RB triangular corner for right-hand side r
      double wn[RB];
      double* wr = w + lane * n_pad + len;
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const double* row = rows + a * ldL + len;
        double v = wr[a] - tot[a * T + lane];
#pragma unroll
        for (int bb = 0; bb < a; ++bb) v -= row[bb] * wn[bb];
        wn[a] = v * row[a];  // diagonal slot holds 1/L_kk
        wr[a] = wn[a];
      }
    }
    __syncwarp();
  }
  for (; i0 < c; ++i0) {  // tail rows (c not a multiple of RB)
    const int len = m + i0;
    const double* row = Lb + (size_t)i0 * ldL;
    double acc[T];
#pragma unroll
    for (int r = 0; r < T; ++r) acc[r] = 0.0;
    for (int k = lane; k < len; k += 32) {
      const double l = row[k];
#pragma unroll
      for (int r = 0; r < T; ++r) acc[r] = fma(l, w[r * n_pad + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < T; ++r) acc[r] = warp_sum(acc[r]);
    const double rd = row[len];
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < T; ++r) w[r * n_pad + len] = (w[r * n_pad + len] - acc[r]) * rd;
    }
    __syncwarp();
  }

  // ---- D: posterior moments: S = K** - sum_i w_i w_i^T, mean = sum_i w_i beta_i -------------------------
  TriT<T> Sacc;
  double macc[T];
#pragma unroll
  for (int i = 0; i < T * (T + 1) / 2; ++i) Sacc.v[i] = 0.0;
#pragma unroll
  for (int r = 0; r < T; ++r) macc[r] = 0.0;
  const double* bh = st.beta_h + (size_t)b * st.c_cap;
  for (int i = lane; i < n; i += 32) {
    const double be = i < m ? sBo[i] : bh[i - m];
    double wi[T];
#pragma unroll
    for (int r = 0; r < T; ++r) wi[r] = w[r * n_pad + i];
#pragma unroll
    for (int r = 0; r < T; ++r) {
      macc[r] = fma(wi[r], be, macc[r]);
#pragma unroll
      for (int s = 0; s <= r; ++s) Sacc.at(r, s) = fma(wi[r], wi[s], Sacc.at(r, s));
    }
  }
#pragma unroll
  for (int i = 0; i < T * (T + 1) / 2; ++i) Sacc.v[i] = warp_sum(Sacc.v[i]);
#pragma unroll
  for (int r = 0; r < T; ++r) macc[r] = warp_sum(macc[r]);
  TriT<T> S;
#pragma unroll
  for (int r = 0; r < T; ++r)
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      double kss = 0.0;
      if (r == s) kss = (r == 0) ? os : os * (il[r > 0 ? r - 1 : 0] * il[r > 0 ? r - 1 : 0]);
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
  double* Lnew = st.Lh + ((size_t)b * st.c_cap + c) * ldL;
  double bn[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    for (int k = lane; k < n; k += 32) Lnew[(size_t)r * ldL + k] = w[r * n_pad + k];
    double t = yv[r] - macc[r];
#pragma unroll
    for (int s = 0; s < r; ++s) t -= Ln.at(r, s) * bn[s];
    bn[r] = t / Ln.at(r, r);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) {
#pragma unroll
      for (int s = 0; s < r; ++s) Lnew[(size_t)r * ldL + n + s] = Ln.at(r, s);
      Lnew[(size_t)r * ldL + n + r] = 1.0 / Ln.at(r, r);  // reciprocal diagonal (see gpmpc_state.cuh)
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
