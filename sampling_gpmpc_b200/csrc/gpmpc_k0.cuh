// K0 for training sets in the thousands (BASELINE configs[4] reaches m = 10^4): BLOCKED factorisation and inverse of the shared
// real-data block on the FP64 tensor cores.  The per-pivot cooperative kernel (k_factor_real_coop, gpmpc_block.cuh) re-reads and
// re-writes the whole trailing matrix for every pivot (m^3 / 3 * 16 bytes of DRAM traffic: 1.3 s at m = 10^4), and k_invert_real
// re-reads L once per COLUMN of the inverse (m^3 / 6 * 8 bytes: 1.4 s).  Here
//   k_factor_real_blocked : right-looking Cholesky in panels of K0P = 64 columns, cooperative over the GPU, two grid barriers per
//                           PANEL: every CTA factorises the 64 x 64 diagonal block in its own shared memory (redundantly: no
//                           barrier, no broadcast), the rows below are solved against it one row per thread, and the trailing
//                           matrix takes A22 -= L21 L21^T tile by tile (64 x 64 x 64, DMMA m8n8k4, operands in shared memory)
//   k_invert_real_mma     : inv(L_oo) by 16-column strips, one CTA per strip, straight into the sub-panel layout: per 8-row
//                           sub-panel the dot products against the strip's own earlier rows are DMMA chains split over the
//                           CTA's warps (A = rows of L from global memory, B = the strip's rows already written), then the
//                           8 x 8 diagonal block is solved by substitution
// Same pivot test and jitter ladder as K0 (every CTA takes the same decisions: identical arithmetic on identical data).
// The summation ORDER differs from the per-pivot kernels (rounding-level differences in L_oo): used from m >= 768 on only
// (GPMPC_K0_BLOCKED_MIN_M), i.e. never at the reference's own configurations (m <= 180), whose results stay bit for bit.
#pragma once
#include <cooperative_groups.h>
#include "gpmpc_block.cuh"
#include "gpmpc_step.cuh"

#define K0P 64            // panel width
#define K0P_LD 68         // leading dimension of the shared-memory tiles (68 = 4 mod 16: conflict-free 64-bit fragment loads)
#define K0F_THREADS 256   // 8 warps: the panel solve keeps a 64-entry row in registers

// lower Cholesky of the nb x nb block in shared memory (leading dimension K0P_LD), every thread of the CTA; k + 1 on failure
__device__ __forceinline__ int k0_chol_tile(double* s, int nb) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4, ny = nt >> 4;
  for (int k = 0; k < nb; ++k) {
    __syncthreads();
    const double akk = s[k * K0P_LD + k];
    if (!(akk > 0.0)) return k + 1;
    const double lkk = sqrt(akk);
    for (int i = k + 1 + tid; i < nb; i += nt) s[i * K0P_LD + k] /= lkk;
    __syncthreads();
    if (tid == 0) s[k * K0P_LD + k] = lkk;
    for (int i = k + 1 + ty; i < nb; i += ny) {
      const double lik = s[i * K0P_LD + k];
      for (int cc = k + 1 + tx; cc <= i; cc += 16) s[i * K0P_LD + cc] -= lik * s[cc * K0P_LD + k];
    }
  }
  __syncthreads();
  return 0;
}

__global__ void __launch_bounds__(K0F_THREADS) k_factor_real_blocked(DevState st) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double k0s[];  // s11 [64][68] | sI [64][68] | sJ [64][68]
  double* s11 = k0s;
  double* sI = s11 + K0P * K0P_LD;
  double* sJ = sI + K0P * K0P_LD;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
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
      for (int i = gw; i < m; i += gnw) {  // row i of the lower triangle by one warp (zeros above the diagonal)
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
      for (int kp = 0; kp < m; kp += K0P) {
        grid.sync();  // the trailing update of the previous panel (first panel: the fill) is complete everywhere
        const int nb = min(K0P, m - kp);
        // ---- a: the diagonal block, factorised by every CTA for itself ----
        for (int idx = tid; idx < nb * nb; idx += nt) {
          const int r = idx / nb, cc = idx - r * nb;
          s11[r * K0P_LD + cc] = cc <= r ? A[(size_t)(kp + r) * m + kp + cc] : 0.0;
        }
        const int bad = k0_chol_tile(s11, nb);  // (begins with a CTA barrier)
        if (bad) { info = kp + bad; break; }    // uniform over the grid
        // ---- b: rows below the block: L21[i][:] = A[i][panel] inv(L11)^T, one row per thread, the row in registers ----
        for (int i = kp + nb + gtid; i < m; i += gnt) {
          double* arow = A + (size_t)i * m + kp;
          double x[K0P];
#pragma unroll
          for (int jj = 0; jj < K0P; ++jj) {
            if (jj < nb) {
              double v = arow[jj];
#pragma unroll
              for (int t = 0; t < K0P; ++t)
                if (t < jj) v = fma(-x[t], s11[jj * K0P_LD + t], v);
              x[jj] = v / s11[jj * K0P_LD + jj];
            }
          }
#pragma unroll
          for (int jj = 0; jj < K0P; ++jj)
            if (jj < nb) arow[jj] = x[jj];
        }
        grid.sync();  // L21 is complete; every CTA has read the unfactorised diagonal block
        if (blockIdx.x == 0)
          for (int idx = tid; idx < nb * nb; idx += nt) {
            const int r = idx / nb, cc = idx - r * nb;
            if (cc <= r) A[(size_t)(kp + r) * m + kp + cc] = s11[r * K0P_LD + cc];
          }
        // ---- c: trailing update A22 -= L21 L21^T, lower 64 x 64 tiles over the CTAs ----
        const int r0 = kp + nb;                      // first trailing row / column
        const int nblk = (m - r0 + K0P - 1) / K0P;   // trailing block rows
        const int n_tiles = nblk * (nblk + 1) / 2;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          int bi = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
          while ((bi + 1) * (bi + 2) / 2 <= tile) ++bi;
          while (bi * (bi + 1) / 2 > tile) --bi;
          const int bj = tile - bi * (bi + 1) / 2;
          const int ri = r0 + bi * K0P, rj = r0 + bj * K0P;
          __syncthreads();  // the previous tile's reads of sI / sJ are complete
          {
            // thread (tr, tc): column tc of rows tr, tr + 4, ... of both operand tiles; every load is issued before the first
            // store (32 independent loads per thread in flight)
            const int tr = tid >> 6, tc = tid & 63;
            double vi[16], vj[16];
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int r = tr + 4 * it;
              vi[it] = (ri + r < m && tc < nb) ? __ldcg(A + (size_t)(ri + r) * m + kp + tc) : 0.0;
              vj[it] = (rj + r < m && tc < nb) ? __ldcg(A + (size_t)(rj + r) * m + kp + tc) : 0.0;
            }
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int r = tr + 4 * it;
              sI[r * K0P_LD + tc] = vi[it];
              sJ[r * K0P_LD + tc] = vj[it];
            }
          }
          __syncthreads();
          // warp w: rows 8 w .. 8 w + 7 of the tile, all 64 columns (8 accumulator tiles)
          double acc[8][2];
#pragma unroll
          for (int n8 = 0; n8 < 8; ++n8) acc[n8][0] = acc[n8][1] = 0.0;
          const double* ap = sI + (8 * warp + gid) * K0P_LD + tig;
          const double* bp = sJ + gid * K0P_LD + tig;
#pragma unroll 4
          for (int k4 = 0; k4 < K0P / 4; ++k4) {
            const double a = ap[4 * k4];
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8) dmma(acc[n8][0], acc[n8][1], a, bp[n8 * 8 * K0P_LD + 4 * k4]);
          }
          const int row = ri + 8 * warp + gid;
          if (row < m) {
            double* crow = A + (size_t)row * m + rj + 2 * tig;
            double cv[8][2];
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8) {  // every load before the first store
              const int col = rj + 8 * n8 + 2 * tig;
              cv[n8][0] = col <= row ? __ldcg(crow + 8 * n8) : 0.0;
              cv[n8][1] = col + 1 <= row ? __ldcg(crow + 8 * n8 + 1) : 0.0;
            }
#pragma unroll
            for (int n8 = 0; n8 < 8; ++n8) {
              const int col = rj + 8 * n8 + 2 * tig;
              if (col <= row) crow[8 * n8] = cv[n8][0] - acc[n8][0];
              if (col + 1 <= row) crow[8 * n8 + 1] = cv[n8][1] - acc[n8][1];
            }
          }
        }
      }
      grid.sync();
      if (info == 0) break;
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
    // (beta_o = L^{-1} y_o follows the inverse: k_beta_from_inverse -- the one-warp substitution of the per-pivot kernels is a
    // serial pass over all of L, 35 ms at m = 10^4)
  }
}

// beta_o = inv(L_oo) y_o from the explicit inverse: one warp per 8-row sub-panel, 256-byte coalesced k-blocks
// ([row 0..7][column 0..3]: lane = 4 row + column), the four column lanes of a row reduced by two shuffles.
__global__ void __launch_bounds__(128) k_beta_from_inverse(DevState st) {
  const int j = blockIdx.y, lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int m = st.m, Pm = (m + 7) >> 3;
  if (p >= Pm) return;
  const double* sp = st.LooP + (size_t)j * subpanel_off(Pm, 0) + subpanel_off(p, 0);
  const double* y = st.y_obs + (size_t)j * m;
  double acc = 0.0;
  for (int kb = 0; kb < 2 * p + 2; ++kb) {
    const int col = 4 * kb + (lane & 3);
    if (col < m) acc = fma(sp[(size_t)kb * 32 + lane], y[col], acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  const int row = 8 * p + (lane >> 2);
  if ((lane & 3) == 0 && row < m) st.beta_o[(size_t)j * m + row] = acc;
}

// inv(L_oo), 16 columns per CTA (strip s: columns 16 s .. 16 s + 15), grid = (ceil(Pm / 2), g_ny).  The strip's rows are produced
// sub-panel by sub-panel (8 rows): X[P][strip] = inv(L_PP) (E_P - sum_{k < 8 P} L[P][k] X[k][strip]); the sum runs over the strip's
// own earlier rows only (X is lower triangular), split over the CTA's warps by k-blocks of 4.
#define K0I_WARPS 16
__global__ void __launch_bounds__(K0I_WARPS * 32) k_invert_real_mma(DevState st) {
  __shared__ double s_part[K0I_WARPS][2][64];  // the warps' partial 8 x 8 sums for the strip's two column blocks: [row][col]
  __shared__ double s_lpp[64];                 // the sub-panel's diagonal block of L
  const int j = blockIdx.y, strip = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, Pm = (m + 7) >> 3;
  const double* A = st.Loo + (size_t)j * m * m;
  double* LP = st.LooP + (size_t)j * subpanel_off(Pm, 0);
  const int p_first = 2 * strip;  // first sub-panel with a non-zero entry in the strip
  // this lane's B-fragment element X[4 k4 + tig][16 strip + 8 nb + gid] sits, inside sub-panel (k4 >> 1), at
  //   ((16 strip + 8 nb + gid) >> 2) * 32 + (4 (k4 & 1) + tig) * 4 + (gid & 3)
  const size_t b_lane0 = (size_t)((16 * strip + gid) >> 2) * 32 + (size_t)tig * 4 + (gid & 3);
  for (int p = p_first; p < Pm; ++p) {
    const int row = 8 * p + gid;
    const double* arow = A + (size_t)min(row, m - 1) * m + tig;
    const bool rowok = row < m;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    // k-blocks [4 strip, 2 p): columns 16 strip .. 8 p - 1 of L's rows 8 p .. 8 p + 7
    for (int k4 = 4 * strip + warp; k4 < 2 * p; k4 += 4 * K0I_WARPS) {
      double a[4], b0[4], b1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = k4 + u * K0I_WARPS;
        const bool ok = kk < 2 * p;
        a[u] = (ok && rowok) ? __ldg(arow + 4 * kk) : 0.0;
        const double* bp = LP + subpanel_off(kk >> 1, 0) + b_lane0 + (size_t)(kk & 1) * 16;
        b0[u] = ok ? __ldcg(bp) : 0.0;
        // column block + 8: two k-blocks further; sub-panel 2 strip holds no such columns (they lie above the diagonal)
        b1[u] = (ok && (kk >> 1) > 2 * strip) ? __ldcg(bp + 64) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dmma(c00, c01, a[u], b0[u]);
        dmma(c10, c11, a[u], b1[u]);
      }
    }
    s_part[warp][0][gid * 8 + 2 * tig] = c00;
    s_part[warp][0][gid * 8 + 2 * tig + 1] = c01;
    s_part[warp][1][gid * 8 + 2 * tig] = c10;
    s_part[warp][1][gid * 8 + 2 * tig + 1] = c11;
    if (tid < 64) {
      const int r = tid >> 3, cc = tid & 7;
      const int gr = 8 * p + r, gc = 8 * p + cc;
      s_lpp[tid] = (gr < m && gc < m && cc <= r) ? A[(size_t)gr * m + gc] : (r == cc ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid < 16) {
      const int nb = tid >> 3, cc = tid & 7, col = 16 * strip + tid;
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double sum = 0.0;
        for (int w = 0; w < K0I_WARPS; ++w) sum += s_part[w][nb][i * 8 + cc];
        double rhs = ((8 * p + i == col) ? 1.0 : 0.0) - sum;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (t < i) rhs = fma(-s_lpp[i * 8 + t], x[t], rhs);
        x[i] = rhs / s_lpp[i * 8 + i];
      }
      if (col < m && col < 8 * (p + 1)) {  // (columns right of the sub-panel's diagonal block do not exist in the layout)
        double* out = LP + subpanel_off(p, 0) + (size_t)(col >> 2) * 32 + (col & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * p + i < m) __stcg(out + i * 4, x[i]);
      }
    }
    __syncthreads();  // the sub-panel's rows of X are visible to every warp of the CTA
  }
}
