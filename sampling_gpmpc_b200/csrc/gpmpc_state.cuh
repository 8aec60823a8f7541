// Device-visible state of one gpmpc handle and the small device helpers every kernel shares.
// fp64 throughout: the reference runs with torch.set_default_dtype(torch.float64) (src/agent.py:15).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gpmpc_b200.h"

#define GP_MIN_VARIANCE 1e-10  // gpytorch.settings.min_variance for double
#define GP_MAX_TRIES 3         // gpytorch.settings.cholesky_max_tries

struct DevState {
  int ns, g_ny, d, T, n_real, B;
  int m;           // observed real scalars (shared by every batch element)
  int c;           // hallucinated scalars currently in the factor (rows of Lh in use)
  int np;          // hallucinated points recorded per batch element
  int cap_points;  // capacity in points
  int c_cap;       // capacity in factor rows = cap_points * T
  int ldC;         // column stride of LhT (doubles), multiple of 4, >= c_cap
  double jitter;
  // shared (per GP output) ---------------------------------------------------------------
  const double* Xr;     // [n_real][d]
  const int* obs_pt;    // [m] real point of observed scalar i
  const int* obs_task;  // [m] its task
  const double* y_obs;  // [g_ny][m]
  const double* ls;     // [g_ny][d]
  const double* os;     // [g_ny]
  const double* noise;  // [g_ny][T]
  double* Loo;          // [g_ny][m][m] row-major lower Cholesky factor of K_oo + Sigma
  double* LooT;         // [g_ny][m(m+1)/2] same factor, packed column-major (column j contiguous), diagonal = 1/L_jj
  double* beta_o;       // [g_ny][m]  L_oo^{-1} y_o
  // per batch element --------------------------------------------------------------------
  double* Xh;           // [B][cap_points][d]
  double* Yh;           // [B][cap_points][T]  labels as appended (NaN kept, for export)
  int* hobs_pt;         // [c_cap] hallucinated point of factor row k   (uniform over b)
  int* hobs_task;       // [c_cap] its task
  // The element's bordered rows  L[m+k][0..m+k]  are stored COLUMN-major: LhT[b][j][k] = L[m+k][j] for
  // j < m+k (strictly below the diagonal); everything else in the (m+c_cap) x ldC slab stays 0 from the
  // allocation memset.  Forward substitution then sweeps columns: a lane owns rows, reads of one column are
  // contiguous over rows, and no cross-lane reduction is needed.  Diagonals live in rdiag as 1/L_kk.
  double* LhT;          // [B][m + c_cap][ldC]
  double* rdiag;        // [B][c_cap]  1 / L[m+k][m+k]
  double* beta_h;       // [B][c_cap]
  unsigned* status;     // device status word (GPMPC_ST_*)
  // workspace of the block kernels ---------------------------------------------------------
  double* W;            // [B][n_ws][q]   L^{-1} K_{o*}
  long long W_stride;   // n_ws * q
  double* S;            // [B][q][q]  Sigma* (lower triangle valid)
  double* C;            // [B][q][q]  scratch for Cholesky
  double* mu;           // [B][q]
  double* xc;           // [B][H][d]  test points the cache was built for
};

__device__ __forceinline__ size_t packed_col(int j, int m) {
  // start of column j in the packed column-major lower triangle (element (i,j), i>=j, at +i-j)
  return (size_t)j * m - ((size_t)j * (j - 1)) / 2;
}

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

// address of L[m+k][col] (col < m+k) of batch element b's own rows
__device__ __forceinline__ double* own_entry(const DevState& st, int b, int k, int col) {
  return st.LhT + ((size_t)b * (st.m + st.c_cap) + col) * st.ldC + k;
}

// strictly-lower entry L[i][col] (col < i) of the full bordered factor [[L_oo, 0], [own rows]] (output j)
__device__ __forceinline__ double factor_entry(const DevState& st, int b, int j, int i, int col) {
  if (i < st.m) return st.Loo[((size_t)j * st.m + i) * st.m + col];
  return *own_entry(st, b, i - st.m, col);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
