// Device-visible state of one gpmpc handle and the small device helpers every kernel shares.
// fp64 throughout: the reference runs with torch.set_default_dtype(torch.float64) (src/agent.py:15).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gpmpc_b200.h"

#define GP_MIN_VARIANCE 1e-10  // gpytorch.settings.min_variance for double
#define GP_MAX_TRIES 3         // gpytorch.settings.cholesky_max_tries

// ---- sub-panel layout of a lower-triangular factor -------------------------------------------------------
// Rows are grouped in SUB-PANELS of 8.  Sub-panel p holds the 8 rows L[8p .. 8p+7][t] of every storage column
// t < off + 8p + 8 (64 bytes per column, "column group"); sub-panels follow one another in row order, so an
// element's factor is ONE contiguous stream that the fused step kernel pulls through shared memory with TMA bulk
// copies in consumption order.  Inside a sub-panel the columns are taken four at a time ("k-block", 256 bytes = the
// A operand of one mma.m8n8k4) and a k-block is stored ROW-major, [row 0..7][column 0..3] (sp_idx):
//   * the A fragment of lane = 4*row + column is double number `lane` of the k-block: a warp reads one contiguous
//     256-byte run, conflict-free (the plain [column][8 rows] order put a half-warp on banks 0-7 and 16-23 only:
//     2-way conflict on every A load, measured 289 conflict wavefronts per element-step at c = 30);
//   * four consecutive columns of ONE row are one aligned 32-byte DRAM sector, so appending a row writes whole
//     sectors.  (Round 1 stored [row half][column][row of the half]: equally conflict-free, but a new row then
//     put 8 bytes into every 32-byte sector it touched -- L2 had to fetch the other 24 from DRAM before writing
//     the sector back: ~4x write amplification plus the fill reads on the appended rows, 15 % of the DRAM traffic
//     of the car rollout.)
//   off = 0 for the shared real-data factor L_oo (storage column = column);
//   off = mo = roundup8(m) for an element's own rows: storage columns [0, m) are the shared columns, [m, mo)
//   are zero padding (so that every column range the kernel iterates over in steps of 4 is aligned), and
//   storage column mo + k is own column k.
// Inside an 8 x 8 diagonal block D the strictly lower slots hold L, the diagonal slots hold 1/L_kk and the strictly
// UPPER slot (row j, column i), j < i, holds inv(D)[i][j] -- the transposed inverse of the block, maintained by
// whoever appends rows (warp_update_dinv) -- so that the fused kernel can apply a whole block solve as two FP64
// tensor-core MMAs (w_blk = inv(D) rhs) instead of an 8-step substitution chain.  Substitution-based readers
// (gpmpc_block.cuh) use the lower part only.
__host__ __device__ __forceinline__ size_t subpanel_off(int p, int off) {
  // doubles before sub-panel p:  8 * sum_{q<p} (off + 8q + 8)
  return (size_t)8 * ((size_t)p * (off + 8) + (size_t)4 * p * (p - 1));
}

// index (doubles) inside a sub-panel of the entry (storage column t, row r of the sub-panel, 0 <= r < 8)
__host__ __device__ __forceinline__ size_t sp_idx(int t, int r) {
  return (size_t)(t >> 2) * 32 + (size_t)(r & 7) * 4 + (t & 3);
}

struct DevState {
  int ns, g_ny, d, T, n_real, B;
  int m;           // observed real scalars (shared by every batch element)
  int mo;          // storage-column offset of the own columns = roundup8(m)
  int c;           // hallucinated scalars currently in the factor (own rows in use)
  int np;          // hallucinated points recorded per batch element
  int cap_points;  // capacity in points
  int c_cap;       // capacity in factor rows = cap_points * T
  long long elem_stride;  // doubles between consecutive elements' factors = subpanel_off(ceil(c_cap/8), mo)
  double jitter;
  // shared (per GP output) ---------------------------------------------------------------
  const double* Xr;     // [n_real][d]
  const int* obs_pt;    // [m] real point of observed scalar i
  const int* obs_task;  // [m] its task
  const double* y_obs;  // [g_ny][m]
  const double* Yr;     // [g_ny][n_real][T] real targets as given (NaN = unobserved), for the min-distance overwrite
  const int* real_full; // [g_ny][n_real] 1 if every task of the real point is observed for that output
  const double* ls;     // [g_ny][d]
  const double* os;     // [g_ny]
  const double* noise;  // [g_ny][T]
  double* Loo;          // [g_ny][m][m] row-major lower Cholesky factor of K_oo + Sigma
  double* LooP;         // [g_ny][subpanel_off(ceil(m/8), 0)] inv(L_oo) in sub-panel layout (zeros above the diagonal)
  double* beta_o;       // [g_ny][m]  L_oo^{-1} y_o
  // per batch element --------------------------------------------------------------------
  double* Xh;           // [B][cap_points][d]
  double* Yh;           // [B][cap_points][T]  labels as appended (NaN kept, for export)
  int* hobs_pt;         // [c_cap] hallucinated point of factor row k   (uniform over b)
  int* hobs_task;       // [c_cap] its task
  int* hrow0;           // [cap_points] first factor row of hallucinated point p (its T tasks are consecutive), -1 if not in the factor
  double* Lh;           // [B][elem_stride] own rows L[m+k][0 .. m+k] in sub-panel layout (off = mo)
  double* beta_h;       // [B][c_cap]
  unsigned char* pstate; // [B][cap_points] grouped rollouts only (else NULL): 0 = point in the factor, 1 = recorded but MASKED
                        //   (a label of its Agent was NaN: observation_nan_policy("mask") drops the slot for the whole batch),
                        //   2 = DROPPED (filtered for all samples of an output, src/agent.py:186-191: never recorded)
  double* Wo;           // [B][mo][T]  shared rows inv(L_oo) k_o of the current step (k_shared_rows -> k_step<WO>), large m only
  double* fin;          // [B][T + T(T+1)/2]  W^T beta and lower(W^T W) of the last fused step (k_step -> k_step_finish)
  unsigned* status;     // device status word (GPMPC_ST_*)
  int* eig_flag;        // epoch of the last draw in which some element's jitter ladder failed (gpmpc_eig.cuh)
  int eig_epoch;        // epoch of the current draw launch
  // workspace of the block kernels ---------------------------------------------------------
  double* W;            // [B][n_ws][q]   L^{-1} K_{o*}
  long long W_stride;   // n_ws * q
  double* S;            // [B][q][q]  Sigma* (lower triangle valid)
  double* C;            // [B][q][q]  scratch for Cholesky
  double* mu;           // [B][q]
  double* S2;           // [B][q][q]  Sigma* w.r.t. the REAL data only (the model of a call whose hallucinated set is about to be
  double* mu2;          // [B][q]     reset, agent.py:261-272: k_append then needs no second posterior pass); NULL: not wanted
  double* xc;           // [B][H][d]  test points the cache was built for
  double* E;            // [B][q][q]  iteration matrix of the eigen-root fallback when it does not fit in shared memory
  double* Lpre;         // [B][q(q+1)/2 + 1]  packed chol(Sigma_app + noise) + its info word, factorised beside the draw by
                        //   k_pm_finish's second CTA row for the k_append that follows (gpmpc_linearise); NULL: not allocated
};

// cov( task ta of f at xa , task tb of f at xb ) for the scaled SE kernel with derivative tasks
// (gpytorch RBFKernelGrad / RBFKernel under ScaleKernel).  r = xa - xb.
__device__ inline double cov_scalar(const double* __restrict__ xa, int ta, const double* __restrict__ xb,
                                    int tb, const double* __restrict__ ls, double os, int d) {
  double s = 0.0, ra = 0.0, rb = 0.0, la = 1.0, lb = 1.0;
  for (int a = 0; a < d; ++a) {
    double r = xa[a] - xb[a];
    double l = ls[a];
    double t = r / l;
    s += t * t;
    if (a == ta - 1) { ra = r; la = l; }
    if (a == tb - 1) { rb = r; lb = l; }
  }
  double k = os * exp(-0.5 * s);
  if (ta == 0 && tb == 0) return k;
  if (ta == 0) return k * ((rb / lb) / lb);
  if (tb == 0) return -k * ((ra / la) / la);
  double h = -((ra / la) / la) * ((rb / lb) / lb);
  if (ta == tb) h += 1.0 / (la * la);
  return k * h;
}

// training scalar i of batch element b -> (pointer to its input point, task)
__device__ __forceinline__ const double* train_scalar(const DevState& st, int b, int i, int& task) {
  if (i < st.m) {
    task = st.obs_task[i];
    return st.Xr + (size_t)st.obs_pt[i] * st.d;
  }
  int k = i - st.m;
  task = st.hobs_task[k];
  return st.Xh + ((size_t)b * st.cap_points + st.hobs_pt[k]) * st.d;
}

// address of L[m+k][col] (logical column col <= m+k) of batch element b's own rows; col == m+k is the
// diagonal slot, which holds 1 / L[m+k][m+k]
__device__ __forceinline__ double* own_entry(const DevState& st, int b, int k, int col) {
  const int t = col < st.m ? col : col - st.m + st.mo;
  return st.Lh + (size_t)b * st.elem_stride + subpanel_off(k >> 3, st.mo) + sp_idx(t, k & 7);
}

// strictly-lower entry L[i][col] (col < i) of the full bordered factor [[L_oo, 0], [own rows]] (output j)
__device__ __forceinline__ double factor_entry(const DevState& st, int b, int j, int i, int col) {
  if (i < st.m) return st.Loo[((size_t)j * st.m + i) * st.m + col];
  return *own_entry(st, b, i - st.m, col);
}

// Completes the diagonal blocks of own rows [k0, k1) of element b (already written: L entries and 1/L_kk) with
// the transposed inverse in the strictly upper slots.  Called by ONE full warp; rows in increasing order:
//   inv(D)[i][j] = -(1/L_ii) * sum_{t=j}^{i-1} D[i][t] inv(D)[t][j],   j < i (indices within the block)
__device__ __forceinline__ void warp_update_dinv(const DevState& st, int b, int k0, int k1, int lane) {
  for (int k = k0; k < k1; ++k) {
    const int i = k & 7, kb = k - i;
    if (lane < i) {
      const int j = lane;
      const double rd = __ldcg(own_entry(st, b, k, st.m + k));
      double acc = 0.0;
      for (int t = j; t < i; ++t)
        acc = fma(__ldcg(own_entry(st, b, k, st.m + kb + t)), __ldcg(own_entry(st, b, kb + j, st.m + kb + t)), acc);
      __stcg(own_entry(st, b, kb + j, st.m + k), -rd * acc);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
