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
  int ldL;         // row stride of Lh (doubles), multiple of 4, >= m + c_cap
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
  double* Lh;           // [B][c_cap][ldL]  bordered rows: row k has m+k+1 entries, the last one is 1/L_kk
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

// row i of the full bordered factor L = [[L_oo, 0], [Lh rows]] for batch element b (output j)
__device__ __forceinline__ const double* factor_row(const DevState& st, int b, int j, int i) {
  if (i < st.m) return st.Loo + ((size_t)j * st.m + i) * st.m;
  return st.Lh + ((size_t)b * st.c_cap + (i - st.m)) * st.ldL;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
