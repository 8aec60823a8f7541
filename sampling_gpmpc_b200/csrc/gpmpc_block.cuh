// General (any H, any n) kernels: one CTA per batch element, state and workspace in global memory
// (L2-resident at the SQP sizes: B <= a few hundred elements).  These serve the SQP linearisation
// (q = H*T = 51..150 joint test scalars) and are the reference semantics for the fused rollout kernel.
//
//   k_factor_real : K0  chol(K_oo + Sigma) per GP output, once per real-data change
//   k_posterior   : K_{o*}, W = L^{-1} K_{o*}, Sigma* = K** - W^T W, mean = W^T beta  (+ optional draw)
//   k_sample      : chol(Sigma*) with GPyTorch's jitter ladder, y = mean + L eps, post-processing
//   k_append      : chol(Sigma* + noise) -> new bordered rows [W^T | L_nn], beta_h
#pragma once
#include "gpmpc_state.cuh"

#define BLK_THREADS 256

// ------------------------------------------------------------------------------------------------
// In-place lower Cholesky of the n x n matrix A (row-major, leading dim ld), cooperative over the CTA.
// LAPACK potrf semantics as used by torch.linalg.cholesky_ex: returns k+1 if pivot k is <= 0 or NaN.
// ------------------------------------------------------------------------------------------------
__device__ int block_cholesky(double* A, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    double akk = A[(size_t)k * ld + k];
    if (!(akk > 0.0)) return k + 1;  // uniform: every thread reads the same value
    double lkk = sqrt(akk);
    __syncthreads();
    if (tid == 0) A[(size_t)k * ld + k] = lkk;
    for (int i = k + 1 + tid; i < n; i += nt) A[(size_t)i * ld + k] /= lkk;
    __syncthreads();
    const int rem = n - k - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      int i = k + 1 + idx / rem, cc = k + 1 + idx % rem;
      if (cc <= i) A[(size_t)i * ld + cc] -= A[(size_t)i * ld + k] * A[(size_t)cc * ld + k];
    }
  }
  __syncthreads();
  return 0;
}

// The same factorisation for the m x m real-data block (K0; m up to a few thousand, A in global memory / L2): the scaled
// pivot column is cached in shared memory (scol, n doubles) so that the rank-1 update reads A[i][cc] once, coalesced over
// cc, and nothing else from global memory; rows over the warps, columns over the lanes, no index divisions.
__device__ int block_cholesky_wide(double* A, int n, int ld, double* scol) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = A[(size_t)k * ld + k];
    if (!(akk > 0.0)) return k + 1;  // uniform: every thread reads the same value
    const double lkk = sqrt(akk);
    __syncthreads();
    if (tid == 0) A[(size_t)k * ld + k] = lkk;
    for (int i = k + 1 + tid; i < n; i += nt) {
      const double v = A[(size_t)i * ld + k] / lkk;
      A[(size_t)i * ld + k] = v;
      scol[i] = v;
    }
    __syncthreads();
    for (int i = k + 1 + wid; i < n; i += nw) {
      const double lik = scol[i];
      double* row = A + (size_t)i * ld;
      for (int cc = k + 1 + lane; cc <= i; cc += 32) row[cc] -= lik * scol[cc];
    }
  }
  __syncthreads();
  return 0;
}

// The same factorisation on a PACKED lower triangle in shared memory (P[r(r+1)/2 + s], s <= r): the q x q matrices of
// the SQP-mode draw / append are latency bound in global memory (three L2 round trips per pivot); in shared memory a
// pivot step costs a few hundred cycles.  Same pivot test, same return value.
__device__ int block_cholesky_packed_pivotwise(double* P, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4, ny = nt >> 4;  // 16 columns x (nt / 16) rows of the trailing block per pass
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = P[k * (k + 1) / 2 + k];
    if (!(akk > 0.0)) return k + 1;
    const double lkk = sqrt(akk);
    __syncthreads();
    if (tid == 0) P[k * (k + 1) / 2 + k] = lkk;
    for (int i = k + 1 + tid; i < n; i += nt) P[i * (i + 1) / 2 + k] /= lkk;
    __syncthreads();
    for (int i = k + 1 + ty; i < n; i += ny) {
      const int ro = i * (i + 1) / 2;
      const double lik = P[ro + k];
      for (int cc = k + 1 + tx; cc <= i; cc += 16) P[ro + cc] -= lik * P[cc * (cc + 1) / 2 + k];
    }
  }
  __syncthreads();
  return 0;
}

// The same arithmetic, entry by entry and in the same order (every entry takes its updates  a -= L[i][k] L[c][k]  for ascending
// k, each one fused multiply-add, then the same division / square root: BIT-IDENTICAL factors), organised in panels of 8
// columns: inside a panel a thread owns one row and keeps its 8 entries in registers (left-looking: column c takes the updates
// of the panel's earlier columns when its turn comes; one CTA barrier per pivot, the pivot row handed on through shared
// memory), then the trailing triangle takes all 8 updates of an entry in one pass (one load and one store per entry instead of
// 8, two barriers per panel instead of 16).  The pivot-wise form above costs ~1100 clk per pivot at q = 51, nearly all of it
// barriers and shared-memory round trips of the trailing update; here the chain per pivot is sqrt + division + one barrier.
// n > blockDim.x: the pivot-wise form (a thread holds one row of the panel).  GPMPC_CHOL_PIVOTWISE=1 at build time selects
// the pivot-wise form everywhere (tools/chol_ab.py: bit-equality and timing of the two).
__device__ int block_cholesky_packed(double* P, int n) {
#ifdef GPMPC_CHOL_PIVOTWISE
  return block_cholesky_packed_pivotwise(P, n);
#else
  if (n > (int)blockDim.x) return block_cholesky_packed_pivotwise(P, n);
  __shared__ double s_row[2][8];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4, ny = nt >> 4;
  int par = 0;
  __syncthreads();  // the caller's fill is complete
  for (int b0 = 0; b0 < n; b0 += 8) {
    const int nbk = min(8, n - b0);
    const int i = b0 + tid;  // this thread's row of the panel
    const bool mine = i < n;
    const int ro = i * (i + 1) / 2 + b0;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (mine && c < nbk && b0 + c <= i) ? P[ro + c] : 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nbk) {  // (uniform)
        if (tid == c) {  // the pivot row b0 + c: complete its diagonal entry, hand the row on
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < c) a[c] = fma(-a[k], a[k], a[c]);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k <= c) s_row[par][k] = a[k];
        }
        __syncthreads();
        const double akk = s_row[par][c];
        if (!(akk > 0.0)) return b0 + c + 1;  // uniform: every thread reads the same value
        const double lkk = sqrt(akk);
        if (tid == c) {
          a[c] = lkk;
        } else if (mine && tid > c) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < c) a[c] = fma(-a[k], s_row[par][k], a[c]);
          a[c] = a[c] / lkk;
        }
        par ^= 1;  // (the buffer is rewritten two pivots later: every thread has passed the next pivot's barrier by then)
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (mine && c < nbk && b0 + c <= i) P[ro + c] = a[c];
    __syncthreads();
    const int t0 = b0 + nbk;  // first trailing column
    for (int i2 = t0 + ty; i2 < n; i2 += ny) {
      const int r2 = i2 * (i2 + 1) / 2;
      double li[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) li[k] = k < nbk ? P[r2 + b0 + k] : 0.0;
      for (int cc = t0 + tx; cc <= i2; cc += 16) {
        const int rc = cc * (cc + 1) / 2 + b0;
        double v = P[r2 + cc];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nbk) v = fma(-li[k], P[rc + k], v);
        P[r2 + cc] = v;
      }
    }
    __syncthreads();
  }
  return 0;
#endif
}

// ------------------------------------------------------------------------------------------------
// K0: shared real-data block.  grid = g_ny, block = BLK_THREADS.
// ------------------------------------------------------------------------------------------------
#define K0_THREADS 1024
__global__ void __launch_bounds__(K0_THREADS) k_factor_real(DevState st) {
  extern __shared__ __align__(16) double k0_col[];  // [m] pivot column of the factorisation
  const int j = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int m = st.m, d = st.d;
  double* A = st.Loo + (size_t)j * m * m;
  const double* ls = st.ls + j * d;
  const double os = st.os[j];
  const double* noise = st.noise + j * st.T;
  int level = 0;
  for (;;) {
    // psd_safe_cholesky: try 0 without jitter, then total jitter*10^(level-1) on the diagonal
    double add = level == 0 ? 0.0 : st.jitter * pow(10.0, (double)(level - 1));
    for (int idx = tid; idx < m * m; idx += nt) {
      int i = idx / m, cc = idx % m;
      double v = 0.0;
      if (cc <= i) {
        v = cov_scalar(st.Xr + (size_t)st.obs_pt[i] * d, st.obs_task[i], st.Xr + (size_t)st.obs_pt[cc] * d,
                       st.obs_task[cc], ls, os, d);
        if (cc == i) v += noise[st.obs_task[i]] + add;
      }
      A[idx] = v;
    }
    int info = block_cholesky_wide(A, m, m, k0_col);
    if (info == 0) break;
    if (level == GP_MAX_TRIES) {
      if (tid == 0) atomicOr(st.status, GPMPC_ST_TRAIN_NOT_PD);
      break;
    }
    ++level;
    __syncthreads();
  }
  if (level > 0 && tid == 0) atomicOr(st.status, GPMPC_ST_TRAIN_JITTER | ((unsigned)level << 8));
  // inv(L_oo) is filled in by k_invert_real (next launch); padding rows / the upper part stay 0
  const int Pm = (m + 7) >> 3;
  double* LP = st.LooP + (size_t)j * subpanel_off(Pm, 0);
  for (int idx = tid; idx < (int)subpanel_off(Pm, 0); idx += nt) LP[idx] = 0.0;
  // beta_o = L^{-1} y_o : forward substitution, one warp, lanes over the row's dot product
  if (tid < 32) {
    const double* y = st.y_obs + (size_t)j * m;
    double* beta = st.beta_o + (size_t)j * m;
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int k = tid; k < i; k += 32) acc += A[(size_t)i * m + k] * beta[k];
      acc = warp_sum(acc);
      if (tid == 0) beta[i] = (y[i] - acc) / A[(size_t)i * m + i];
      __syncwarp();
    }
  }
}

// K0c: the same factorisation for m in the thousands, COOPERATIVE over the whole GPU (cudaLaunchCooperativeKernel, one CTA
// per SM, one grid barrier per pivot): every CTA keeps its own copy of the scaled pivot column in shared memory (double
// buffered), the rows of the trailing update are spread over all warps of the grid, and the scaled column of pivot k is
// written back by the rows' owners after the barrier that opens pivot k + 1 (nobody reads column k of A after that
// barrier).  Same pivot test and jitter ladder as K0 (every CTA sees the same pivot value, so the decisions are uniform).
#include <cooperative_groups.h>
__global__ void __launch_bounds__(K0_THREADS) k_factor_real_coop(DevState st) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double k0c_col[];  // [2][m]
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int gtid = blockIdx.x * nt + tid, gnt = gridDim.x * nt;
  const int gw = gtid >> 5, gnw = gnt >> 5;
  const int m = st.m, d = st.d;
  for (int j = 0; j < st.g_ny; ++j) {
    double* A = st.Loo + (size_t)j * m * m;
    const double* ls = st.ls + j * d;
    const double os = st.os[j];
    const double* noise = st.noise + j * st.T;
    int level = 0;
    for (;;) {
      const double add = level == 0 ? 0.0 : st.jitter * pow(10.0, (double)(level - 1));
      for (int i = gw; i < m; i += gnw) {  // row i of the lower triangle by one warp
        const double* xi = st.Xr + (size_t)st.obs_pt[i] * d;
        const int ti = st.obs_task[i];
        for (int cc = lane; cc < m; cc += 32) {
          double v = 0.0;
          if (cc <= i) {
            v = cov_scalar(xi, ti, st.Xr + (size_t)st.obs_pt[cc] * d, st.obs_task[cc], ls, os, d);
            if (cc == i) v += noise[ti] + add;
          }
          A[(size_t)i * m + cc] = v;
        }
      }
      int info = 0;
      double lkk_prev = 0.0;
      for (int k = 0; k < m; ++k) {
        grid.sync();  // the trailing update of pivot k - 1 is complete everywhere
        double* scol = k0c_col + (size_t)(k & 1) * m;
        const double* prev = k0c_col + (size_t)((k + 1) & 1) * m;
        const double akk = A[(size_t)k * m + k];
        if (!(akk > 0.0)) { info = k + 1; break; }  // uniform over the grid: nobody writes a(k,k) during pivot k
        const double lkk = sqrt(akk);
        for (int i = k + 1 + tid; i < m; i += nt) scol[i] = A[(size_t)i * m + k] / lkk;
        __syncthreads();
        for (int i = k + 1 + gw; i < m; i += gnw) {
          double* row = A + (size_t)i * m;
          const double lik = scol[i];
          for (int cc = k + 1 + lane; cc <= i; cc += 32) row[cc] -= lik * scol[cc];
        }
        // column k - 1 (scaled, from the other buffer) and its diagonal go back to A now: every CTA has left pivot k - 1,
        // and pivot k reads column k and touches columns > k only
        if (k > 0) {
          for (int i = k + gtid; i < m; i += gnt) A[(size_t)i * m + (k - 1)] = prev[i];
          if (gtid == 0) A[(size_t)(k - 1) * m + (k - 1)] = lkk_prev;
        }
        lkk_prev = lkk;
      }
      grid.sync();
      if (info == 0) {
        if (gtid == 0) A[(size_t)(m - 1) * m + (m - 1)] = lkk_prev;  // the last pivot has no column below it
        break;
      }
      if (level == GP_MAX_TRIES) {
        if (gtid == 0) atomicOr(st.status, GPMPC_ST_TRAIN_NOT_PD);
        break;
      }
      ++level;
    }
    if (level > 0 && gtid == 0) atomicOr(st.status, GPMPC_ST_TRAIN_JITTER | ((unsigned)level << 8));
    const int Pm = (m + 7) >> 3;
    double* LP = st.LooP + (size_t)j * subpanel_off(Pm, 0);
    for (size_t idx = gtid; idx < subpanel_off(Pm, 0); idx += gnt) LP[idx] = 0.0;
    grid.sync();
    if (blockIdx.x == 0 && tid < 32) {  // beta_o = L^{-1} y_o
      const double* y = st.y_obs + (size_t)j * m;
      double* beta = st.beta_o + (size_t)j * m;
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int k = tid; k < i; k += 32) acc += A[(size_t)i * m + k] * beta[k];
        acc = warp_sum(acc);
        if (tid == 0) beta[i] = (y[i] - acc) / A[(size_t)i * m + i];
        __syncwarp();
      }
    }
  }
}

// K0b: explicit inverse of L_oo in sub-panel layout (gpmpc_state.cuh) for the fused / tensor-core kernels: w_o = inv(L_oo) k_o
// is then a plain product.  cond(L_oo) ~ 1e3 at the reference's configurations: the inverse costs ~1e-14 relative accuracy
// in the posterior variance, 5 orders below the parity tolerance.  The columns of the inverse are independent:
// one WARP per column jj solves L x = e_jj by forward substitution, lanes over the row's dot product, the solution kept
// in shared memory; grid = (ceil(m / warps), g_ny), so the whole GPU works on it (one thread per column in one CTA took
// more than half of K0's 9 s at m = 3000).
#define K0B_WARPS 4
__global__ void __launch_bounds__(K0B_WARPS * 32) k_invert_real(DevState st) {
  extern __shared__ __align__(16) double k0b_x[];  // [K0B_WARPS][m]
  const int j = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int m = st.m, jj = blockIdx.x * (blockDim.x >> 5) + wid;  // (the launch picks 4, 2 or 1 warps per CTA by m)
  if (jj >= m) return;
  const double* A = st.Loo + (size_t)j * m * m;
  double* LP = st.LooP + (size_t)j * subpanel_off((m + 7) >> 3, 0);
  double* xs = k0b_x + (size_t)wid * m;
  for (int i = jj; i < m; ++i) {
    const double* row = A + (size_t)i * m;
    double acc = 0.0;
    for (int k = jj + lane; k < i; k += 32) acc += row[k] * xs[k];
    acc = warp_sum(acc);
    if (lane == 0) xs[i] = ((i == jj ? 1.0 : 0.0) - acc) / row[i];
    __syncwarp();
  }
  for (int i = jj + lane; i < m; i += 32) LP[subpanel_off(i >> 3, 0) + sp_idx(jj, i & 7)] = xs[i];
}

// ------------------------------------------------------------------------------------------------
// Posterior pieces (device functions shared by k_posterior and k_append's recompute path)
// ------------------------------------------------------------------------------------------------
#define FS_ROWS 16

// W[i][r] = cov(train scalar i, test scalar r); then W <- L^{-1} W; S = K** - W^T W; mu = W^T beta.
__device__ void block_posterior(const DevState& st, int b, const double* __restrict__ x, int H) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int j = b % st.g_ny, d = st.d, T = st.T;
  const int n = st.m + st.c, q = H * T;
  double* W = st.W + (size_t)b * st.W_stride;
  double* S = st.S + (size_t)b * q * q;
  double* mu = st.mu + (size_t)b * q;
  double* xc = st.xc + (size_t)b * H * d;
  const double* ls = st.ls + j * d;
  const double os = st.os[j];
  const double* xb = x + (size_t)b * H * d;

  for (int idx = tid; idx < H * d; idx += nt) xc[idx] = xb[idx];
  for (int idx = tid; idx < n * q; idx += nt) {
    int i = idx / q, r = idx % q, ta;
    const double* xa = train_scalar(st, b, i, ta);
    W[idx] = cov_scalar(xa, ta, xb + (size_t)(r / T) * d, r % T, ls, os, d);
  }
  __syncthreads();

  // blocked forward substitution against the bordered factor
  for (int i0 = 0; i0 < n; i0 += FS_ROWS) {
    const int rb = min(FS_ROWS, n - i0);
    for (int idx = tid; idx < rb * q; idx += nt) {
      int a = idx / q, r = idx % q;
      double acc = 0.0;
      for (int k = 0; k < i0; ++k) acc += factor_entry(st, b, j, i0 + a, k) * W[(size_t)k * q + r];
      W[(size_t)(i0 + a) * q + r] -= acc;
    }
    __syncthreads();
    for (int r = tid; r < q; r += nt) {  // a thread owns column r: no sync needed inside the block
      for (int a = 0; a < rb; ++a) {
        const int i = i0 + a;
        double v = W[(size_t)i * q + r];
        for (int bb = 0; bb < a; ++bb) v -= factor_entry(st, b, j, i, i0 + bb) * W[(size_t)(i0 + bb) * q + r];
        // own rows keep 1/L_kk in their diagonal slot, the shared block keeps L_kk on its diagonal
        W[(size_t)i * q + r] = (i >= st.m) ? v * *own_entry(st, b, i - st.m, i)
                                           : v / st.Loo[((size_t)j * st.m + i) * st.m + i];
      }
    }
    __syncthreads();
  }

  // Sigma* (lower triangle) and mean
  const double* beta_o = st.beta_o + (size_t)j * st.m;
  const double* beta_h = st.beta_h + (size_t)b * st.c_cap;
  for (int idx = tid; idx < q * q; idx += nt) {
    int r = idx / q, s = idx % q;
    if (s > r) continue;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += W[(size_t)i * q + r] * W[(size_t)i * q + s];
    double kss = cov_scalar(xb + (size_t)(r / T) * d, r % T, xb + (size_t)(s / T) * d, s % T, ls, os, d);
    S[idx] = kss - acc;
  }
  for (int r = tid; r < q; r += nt) {
    double acc = 0.0;
    for (int i = 0; i < st.m; ++i) acc += W[(size_t)i * q + r] * beta_o[i];
    for (int i = st.m; i < n; ++i) acc += W[(size_t)i * q + r] * beta_h[i - st.m];
    mu[r] = acc;
  }
  __syncthreads();
}

__device__ void block_postprocess(const DevState& st, int b, int H, gpmpc_sample_opts opts, double* __restrict__ yb);

// chol(Sigma*) with the psd_safe_cholesky ladder, y = mu + L eps, then sample_gp's post-processing
// (zero-variance -> mean, truncation to mean +- beta sqrt(var); src/agent.py:646-663,701-708).
// tri != NULL: shared-memory scratch of at least q(q+1)/2 doubles for the factorisation (packed lower triangle).
__device__ void block_sample(const DevState& st, int b, int H, const double* __restrict__ eps,
                             gpmpc_sample_opts opts, double* __restrict__ y, int* __restrict__ jitter_level,
                             double* tri = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int T = st.T, q = H * T;
  const double* S = st.S + (size_t)b * q * q;
  double* C = st.C + (size_t)b * q * q;
  const double* mu = st.mu + (size_t)b * q;
  const double* e = eps + (size_t)b * q;
  double* yb = y + (size_t)b * q;
  int level = 0;
  if (q == 1) {
    // LinearOperator._cholesky of a 1x1: clamp_min(0).sqrt(); zero_mean_mvn_samples: plain sqrt
    if (tid == 0) C[0] = opts.unclamped_sqrt_1x1 ? sqrt(S[0]) : sqrt(fmax(S[0], 0.0));
    __syncthreads();
  } else {
    for (;;) {
      double add = level == 0 ? 0.0 : st.jitter * pow(10.0, (double)(level - 1));
      int info;
      if (tri) {
        __syncthreads();
        for (int idx = tid; idx < q * q; idx += nt) {  // coalesced over the rows of S
          const int r = idx / q, s = idx - r * q;
          if (s <= r) tri[r * (r + 1) / 2 + s] = S[idx] + (s == r ? add : 0.0);
        }
        info = block_cholesky_packed(tri, q);
      } else {
        for (int idx = tid; idx < q * q; idx += nt) {
          int r = idx / q, s = idx % q;
          double v = s <= r ? S[idx] : 0.0;
          if (s == r) v += add;
          C[idx] = v;
        }
        info = block_cholesky(C, q, q);
      }
      if (info == 0) break;
      // NaN anywhere makes GPyTorch raise NanError instead of climbing the ladder
      if (level == GP_MAX_TRIES) {
        level = 4;
        // a NaN in the matrix is GPyTorch's NanError (and eigh of a NaN matrix raises too): no eigen fallback for that
        int nan_here = 0;
        for (int idx = tid; idx < q * q; idx += nt) nan_here |= (idx % q <= idx / q && isnan(S[idx])) ? 1 : 0;
        const int has_nan = __syncthreads_or(nan_here);
        if (tid == 0) {
          atomicOr(st.status, GPMPC_ST_SAMPLE_NOT_PD | (has_nan ? GPMPC_ST_NAN_INPUT : 0u));
          // GPyTorch: the whole batch falls back to the eigen root (gpmpc_eig.cuh) -- tell the kernel queued behind this one
          if (!has_nan && !(opts.flags & GPMPC_OPT_NO_EIG_FALLBACK)) atomicMax(st.eig_flag, st.eig_epoch);
        }
        break;
      }
      ++level;
      __syncthreads();
    }
  }
  if (tid == 0 && jitter_level) jitter_level[b] = level;
  for (int r = tid; r < q; r += nt) {
    double acc = mu[r];
    if (level < 4 && tri && q > 1)
      for (int s = 0; s <= r; ++s) acc += tri[r * (r + 1) / 2 + s] * e[s];
    else if (level < 4)
      for (int s = 0; s <= r; ++s) acc += C[(size_t)r * q + s] * e[s];
    else
      acc = nan("");
    yb[r] = acc;
  }
  __syncthreads();
  block_postprocess(st, b, H, opts, yb);
}

// sample_gp's post-processing of a draw yb [q] (zero-variance -> mean, truncation; src/agent.py:646-663,701-708)
__device__ void block_postprocess(const DevState& st, int b, int H, gpmpc_sample_opts opts, double* __restrict__ yb) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int T = st.T, q = H * T;
  const double* S = st.S + (size_t)b * q * q;
  const double* mu = st.mu + (size_t)b * q;
  for (int h = tid; h < H; h += nt) {
    bool zero = opts.variance_is_zero >= 0.0;
    if (zero)
      for (int t = 0; t < T; ++t) {
        double v = fmax(S[(size_t)(h * T + t) * q + h * T + t], GP_MIN_VARIANCE);
        zero = zero && (v <= opts.variance_is_zero);
      }
    for (int t = 0; t < T; ++t) {
      int r = h * T + t;
      double v = fmax(S[(size_t)r * q + r], GP_MIN_VARIANCE), mr = mu[r], yy = yb[r];
      if (zero) yy = mr;
      if (opts.beta >= 0.0) {
        double sd = sqrt(v);
        yy = fmin(fmax(yy, mr - opts.beta * sd), mr + opts.beta * sd);
      }
      yb[r] = yy;
    }
  }
}

__global__ void __launch_bounds__(BLK_THREADS)
k_posterior(DevState st, const double* __restrict__ x, int H, double* __restrict__ mean,
            double* __restrict__ var, const double* __restrict__ eps, gpmpc_sample_opts opts,
            double* __restrict__ y, int* __restrict__ jitter_level) {
  const int b = blockIdx.x, q = H * st.T;
  block_posterior(st, b, x, H);
  const double* S = st.S + (size_t)b * q * q;
  const double* mu = st.mu + (size_t)b * q;
  for (int r = threadIdx.x; r < q; r += blockDim.x) {
    if (mean) mean[(size_t)b * q + r] = mu[r];
    if (var) var[(size_t)b * q + r] = fmax(S[(size_t)r * q + r], GP_MIN_VARIANCE);
  }
  if (eps) block_sample(st, b, H, eps, opts, y, jitter_level);
}

// tri_ok: the launch carries q(q+1)/2 doubles of dynamic shared memory for the packed Cholesky
__global__ void __launch_bounds__(BLK_THREADS)
k_sample(DevState st, int H, const double* __restrict__ eps, gpmpc_sample_opts opts, double* __restrict__ y,
         int* __restrict__ jitter_level, int tri_ok) {
  extern __shared__ __align__(16) double dyn_tri[];
  block_sample(st, blockIdx.x, H, eps, opts, y, jitter_level, tri_ok ? dyn_tri : nullptr);
}

// ------------------------------------------------------------------------------------------------
// Conditioning: append the active points' T scalars each to element b's factor.
//   reuse != 0 : the workspace may hold W, S, mu for exactly these x (checked per element on device)
//   active     : DEVICE uint8[H*T] (per test scalar) or NULL; pt_base = index of the first new point in Xh/Yh
// New rows k = c + r':  L[m+k][0..n) = W[:, act(r')],  L[m+k][n + s'] = chol(S_act + noise)[r'][s'].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK_THREADS)
k_append(DevState st, const double* __restrict__ x, const double* __restrict__ ylab,
         const unsigned char* __restrict__ active, int H, int pt_base, int reuse, int grow_factor, int tri_ok,
         int prefactored) {
  extern __shared__ __align__(16) double dyn_tri[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int j = b % st.g_ny, d = st.d, T = st.T, q = H * T;
  const int n = st.m + st.c;
  __shared__ int sh_same;
  __shared__ int sh_act[512];  // active test scalars (q' <= 512 checked by the host)
  __shared__ int sh_qa;

  // record the points (labels keep their NaNs); `active` flags SCALARS (point h, task t) -> active[h * T + t].  A fully
  // active point's T factor rows are consecutive from hrow0; a point with no or only some active scalars gets
  // hrow0 = -1 (the host then routes model calls to the kernels that walk hobs_pt / hobs_task row by row)
  if (b == 0)
    for (int h = tid; h < H; h += nt) {
      int row = -1;
      if (grow_factor) {
        int mine = 0, before = 0;
        for (int t = 0; t < T; ++t) mine += (!active || active[h * T + t]) ? 1 : 0;
        for (int s2 = 0; s2 < h * T; ++s2) before += (!active || active[s2]) ? 1 : 0;
        if (mine == T) row = st.c + before;
      }
      st.hrow0[pt_base + h] = row;
    }
  for (int idx = tid; idx < H * d; idx += nt)
    st.Xh[((size_t)b * st.cap_points + pt_base) * d + idx] = x[(size_t)b * H * d + idx];
  for (int idx = tid; idx < H * T; idx += nt)
    st.Yh[((size_t)b * st.cap_points + pt_base) * T + idx] = ylab[(size_t)b * q + idx];
  if (!grow_factor) return;

  if (tid == 0) {
    sh_same = reuse;
    int qa = 0;
    for (int s2 = 0; s2 < H * T; ++s2)
      if (!active || active[s2]) sh_act[qa++] = s2;
    sh_qa = qa;
  }
  __syncthreads();
  if (reuse) {
    const double* xc = st.xc + (size_t)b * H * d;
    for (int idx = tid; idx < H * d; idx += nt)
      if (xc[idx] != x[(size_t)b * H * d + idx]) sh_same = 0;  // benign race: all writers store 0
    __syncthreads();
  }
  if (!sh_same) block_posterior(st, b, x, H);
  const int qa = sh_qa;
  if (qa == 0) return;

  const double* W = st.W + (size_t)b * st.W_stride;
  const double* S = st.S + (size_t)b * q * q;
  double* C = st.C + (size_t)b * q * q;  // used as qa x qa, leading dim qa
  const double* mu = st.mu + (size_t)b * q;
  const double* noise = st.noise + j * T;
  int info;
  if (tri_ok) {
    const int ntri = qa * (qa + 1) / 2;
    if (prefactored && sh_same) {
      // chol(Sigma_app + noise) was factorised beside the draw (k_pm_finish, second CTA row): pick it up
      const double* pre = st.Lpre + (size_t)b * (ntri + 1);
      for (int idx = tid; idx < ntri; idx += nt) dyn_tri[idx] = pre[idx];
      info = (int)pre[ntri];
      __syncthreads();
    } else {
      for (int idx = tid; idx < qa * qa; idx += nt) {
        const int rr = idx / qa, ss = idx - rr * qa;
        if (ss <= rr)
          dyn_tri[rr * (rr + 1) / 2 + ss] = S[(size_t)sh_act[rr] * q + sh_act[ss]] + (ss == rr ? noise[sh_act[rr] % T] : 0.0);
      }
      info = block_cholesky_packed(dyn_tri, qa);  // the rest of the kernel reads the factor straight from shared memory
    }
  } else {
    for (int idx = tid; idx < qa * qa; idx += nt) {
      int rr = idx / qa, ss = idx % qa;
      double v = 0.0;
      if (ss <= rr) {
        v = S[(size_t)sh_act[rr] * q + sh_act[ss]];
        if (ss == rr) v += noise[sh_act[rr] % T];
      }
      C[idx] = v;
    }
    info = block_cholesky(C, qa, qa);
  }
  if (info != 0) {
    if (tid == 0) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
    return;
  }
  __shared__ double sh_beta[512];
  if (b == 0)
    for (int rr = tid; rr < qa; rr += nt) {
      st.hobs_pt[st.c + rr] = pt_base + sh_act[rr] / T;
      st.hobs_task[st.c + rr] = sh_act[rr] % T;
    }
  if (tri_ok) {
    // Factor in shared memory: warp 0 solves for beta_new (a serial chain of qa rows) WHILE the other warps write the new rows
    // and complete the touched diagonal blocks' transposed inverses -- none of the three needs the others' results.
    if (tid < 32) {
      double* beta = st.beta_h + (size_t)b * st.c_cap + st.c;
      const double* yb = ylab + (size_t)b * q;
      for (int rr = tid; rr < qa; rr += 32) sh_beta[rr] = yb[sh_act[rr]] - mu[sh_act[rr]];
      __syncwarp();
      for (int rr = 0; rr < qa; ++rr) {
        double acc = 0.0;
        for (int ss = tid; ss < rr; ss += 32) acc += dyn_tri[rr * (rr + 1) / 2 + ss] * sh_beta[ss];
        acc = warp_sum(acc);
        if (tid == 0) sh_beta[rr] = (sh_beta[rr] - acc) / dyn_tri[rr * (rr + 1) / 2 + rr];
        __syncwarp();
      }
      for (int rr = tid; rr < qa; rr += 32) beta[rr] = sh_beta[rr];
      return;
    }
    const int wt = tid - 32, wnt = nt - 32;  // the writers
    // new rows k = c + rr, written column by column (rr fastest: contiguous within a column)
    // The part left of the new block, one k-block (4 storage columns of one row = one 32-byte sector, gpmpc_state.cuh) per thread
    // and two per pass: eight loads of W in flight (one entry per pass was a chain of dependent L2 round trips), the sector
    // written whole by two 16-byte stores.  Lanes = consecutive new rows: their sectors are contiguous inside a k-block, so a
    // warp's store covers whole 256-byte k-blocks.  Storage column t: t < m shared (W row t), [m, mo) padding (zero), t >= mo own
    // (W row m + t - mo); the columns of a k-block that straddles the new block's first column go entry by entry below.
    {
      const int c0 = st.c, mo = st.mo, m = st.m;
      const int nkb_full = (mo + c0) >> 2;
      double* Le = st.Lh + (size_t)b * st.elem_stride;
      for (int idx0 = wt; idx0 < qa * nkb_full; idx0 += 2 * wnt) {
        double v[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = idx0 + u * wnt;
          if (idx < qa * nkb_full) {
            const int kb = idx / qa, rr = idx - kb * qa;
            const double* wc = W + sh_act[rr];
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
              const int t = 4 * kb + jq;
              const int r = t < m ? t : (t >= mo ? t - mo + m : -1);
              v[u][jq] = r >= 0 ? __ldcg(wc + (size_t)r * q) : 0.0;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = idx0 + u * wnt;
          if (idx < qa * nkb_full) {
            const int kb = idx / qa, rr = idx - kb * qa, row = c0 + rr;
            double2* dst = reinterpret_cast<double2*>(Le + subpanel_off(row >> 3, mo) + (size_t)kb * 32 + (row & 7) * 4);
            dst[0] = make_double2(v[u][0], v[u][1]);
            dst[1] = make_double2(v[u][2], v[u][3]);
          }
        }
      }
      const int t_rem = 4 * nkb_full, n_rem = mo + c0 - t_rem;  // 0 .. 3 own columns left of the new block
      for (int idx = wt; idx < qa * n_rem; idx += wnt) {
        const int tq = idx / qa, rr = idx - tq * qa, k = m + (t_rem + tq - mo);
        *own_entry(st, b, c0 + rr, k) = W[(size_t)k * q + sh_act[rr]];
      }
    }
    for (int idx = wt; idx < qa * qa; idx += wnt) {
      int ss = idx / qa, rr = idx % qa;
      if (ss < rr) *own_entry(st, b, st.c + rr, n + ss) = dyn_tri[rr * (rr + 1) / 2 + ss];
      else if (ss == rr) *own_entry(st, b, st.c + rr, n + rr) = 1.0 / dyn_tri[rr * (rr + 1) / 2 + rr];
    }
    // transposed inverses of the 8 x 8 diagonal blocks the new rows touch (gpmpc_state.cuh; same arithmetic, same order as
    // warp_update_dinv): one warp per block, LANE j < 8 owns column j of the block's inverse in registers,
    //   inv[i][j] = -(1 / L_ii) * sum_{t = j}^{i - 1} L[i][t] inv[t][j],   inv[j][j] = 1 / L_jj,
    // fed from the factor in shared memory (new rows and columns), from W (new rows under old columns of a block that was
    // partly filled before) and from the old rows' slots in global memory -- nothing this kernel itself writes is read back.
    const int lane = tid & 31, wi = (tid >> 5) - 1, wn = (nt >> 5) - 1, c0 = st.c;
    for (int kb = (c0 & ~7) + 8 * wi; kb < c0 + qa; kb += 8 * wn) {
      if (lane >= 8) continue;
      const int jc = lane;
      double invc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kb + i;  // own row
        invc[i] = 0.0;
        if (k >= c0 + qa) continue;
        if (k < c0) {
          // an old row of a partly filled block: its slots are complete
          if (i == jc) invc[i] = __ldcg(own_entry(st, b, k, st.m + k));
          else if (i > jc) invc[i] = __ldcg(own_entry(st, b, kb + jc, st.m + k));
          continue;
        }
        const int rr = k - c0;
        const double rd = 1.0 / dyn_tri[rr * (rr + 1) / 2 + rr];
        if (i == jc) invc[i] = rd;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t < i) {
            const int oc = kb + t;  // own column
            const double lit = oc < c0 ? W[(size_t)(st.m + oc) * q + sh_act[rr]] : dyn_tri[rr * (rr + 1) / 2 + (oc - c0)];
            if (t >= jc) acc = fma(lit, invc[t], acc);
          }
        }
        if (jc < i) {
          invc[i] = -rd * acc;
          __stcg(own_entry(st, b, kb + jc, st.m + k), invc[i]);
        }
      }
    }
    return;
  }
  // new rows k = c + rr, written column by column (rr fastest: contiguous within a column)
  for (int idx = tid; idx < qa * n; idx += nt) {
    int k = idx / qa, rr = idx % qa;
    *own_entry(st, b, st.c + rr, k) = W[(size_t)k * q + sh_act[rr]];
  }
  for (int idx = tid; idx < qa * qa; idx += nt) {
    int ss = idx / qa, rr = idx % qa;
    if (ss < rr) *own_entry(st, b, st.c + rr, n + ss) = C[(size_t)rr * qa + ss];
    else if (ss == rr) *own_entry(st, b, st.c + rr, n + rr) = 1.0 / C[(size_t)rr * qa + rr];
  }
  // beta_new = L_nn^{-1} (y - mu): forward substitution by one warp, the partial solution kept in shared memory
  // (sh_beta aliases nothing: 512 doubles) so that a row costs a shared-memory round trip, not an L2 one
  if (tid < 32) {
    double* beta = st.beta_h + (size_t)b * st.c_cap + st.c;
    const double* yb = ylab + (size_t)b * q;
    for (int rr = tid; rr < qa; rr += 32) sh_beta[rr] = yb[sh_act[rr]] - mu[sh_act[rr]];
    __syncwarp();
    for (int rr = 0; rr < qa; ++rr) {
      double acc = 0.0;
      for (int ss = tid; ss < rr; ss += 32) acc += C[(size_t)rr * qa + ss] * sh_beta[ss];
      acc = warp_sum(acc);
      if (tid == 0) sh_beta[rr] = (sh_beta[rr] - acc) / C[(size_t)rr * qa + rr];
      __syncwarp();
    }
    for (int rr = tid; rr < qa; rr += 32) beta[rr] = sh_beta[rr];
  }
  __syncthreads();
  // transposed inverses of the 8 x 8 diagonal blocks the new rows touch: the blocks are independent, one warp each
  for (int kb = (st.c & ~7) + 8 * (tid >> 5); kb < st.c + qa; kb += 8 * (nt >> 5))
    warp_update_dinv(st, b, max(kb, st.c), min(kb + 8, st.c + qa), tid & 31);
}
