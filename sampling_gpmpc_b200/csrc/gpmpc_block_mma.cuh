// K2m: the SQP-mode model call (any H) on the FP64 tensor cores.
//
// The joint posterior of H test points (q = H*T scalars) given n = m + c training scalars is dominated by the
// triangular solve W = L^{-1} K_{o*} with q right-hand sides (q n^2 / 2 MACs) and by Sigma* = K** - W^T W (q^2 n / 2):
// contraction bound once the hallucinated set has grown over a few SQP iterations (car-residual config: q = 150, n up
// to several thousand), where the scalar kernel k_posterior (gpmpc_block.cuh, kept as the substitution-based reference
// semantics) reaches ~1 % of the FP64 peak.  One CTA per batch element, 8 warps, each warp owns 8-column blocks of the
// right-hand sides (columns never interact, so no CTA-wide synchronisation is needed for W itself):
//   0  K_{o*}: one exp per (training point, test point) pair, the T x T derivative block from it
//   1  shared rows: W_o = inv(L_oo) K_o, tile-rows last to first, in place (as K1 phase B)
//   2  own rows, left-looking over the 8-row sub-panels of the element's factor stream: the sub-panel's k-blocks are
//      staged in shared memory once (all column blocks need the same A operand), every warp runs the DMMA chain
//      dot = L[rows][cols < n_off] W over its column blocks with W read back through L2, rhs = K - dot, and the 8 x 8
//      diagonal block is applied as inv(D) rhs (two more DMMAs; inverse kept in the block's upper triangle)
//   3  Sigma* = K** - W^T W by 8 x 8 output tiles (A and B fragments are the same access pattern on W), mean = W^T beta
// then the draw / post-processing of gpmpc_block.cuh (block_sample) in the same launch.
// W [n][q], S [q][q], mu [q], xc keep the layouts of the scalar kernel: k_sample / k_append consume either.
#pragma once
#include "gpmpc_block.cuh"
#include "gpmpc_step.cuh"

#define PM_THREADS 256
#define PM_MAXOWN 4     // column blocks per warp (q <= 8 * 8 * 4 = 256 test scalars per call)
#define PM_SLAB 1024    // storage columns of a sub-panel staged per pass (64 KB of shared memory)

template <int D, int T>
__global__ void __launch_bounds__(PM_THREADS, 2)
k_posterior_mma(DevState st, const double* __restrict__ x, int H, double* __restrict__ mean,
                double* __restrict__ var, const double* __restrict__ eps, gpmpc_sample_opts opts,
                double* __restrict__ y, int* __restrict__ jitter_level) {
  extern __shared__ __align__(128) double sA[];  // [PM_SLAB * 8] one slab of a sub-panel (k-block layout)
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int j = b % st.g_ny;
  const int m = st.m, mo = st.mo, c = st.c, n = m + c, q = H * T;
  const int QB = (q + 7) >> 3;
  double* W = st.W + (size_t)b * st.W_stride;
  double* S = st.S + (size_t)b * q * q;
  double* mu = st.mu + (size_t)b * q;
  double* xc = st.xc + (size_t)b * H * D;
  const double* xb = x + (size_t)b * H * D;
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j * D + a];
  const double os = st.os[j];

  // ---- 0: kernel matrix ---------------------------------------------------------------------------------------
  for (int idx = tid; idx < H * D; idx += nt) xc[idx] = xb[idx];
  for (int idx = tid; idx < m * H; idx += nt) {
    const int i = idx / H, h = idx - i * H;
    double xs[D], out[T];
#pragma unroll
    for (int a = 0; a < D; ++a) xs[a] = xb[h * D + a];
    kernel_row<D, T>(st.Xr + (size_t)st.obs_pt[i] * D, st.obs_task[i], xs, il, os, out);
#pragma unroll
    for (int tb = 0; tb < T; ++tb) W[(size_t)i * q + h * T + tb] = out[tb];
  }
  for (int idx = tid; idx < st.np * H; idx += nt) {
    const int p = idx / H, h = idx - p * H;
    const int r0 = st.hrow0[p];
    if (r0 < 0) continue;  // recorded but masked point: no factor rows
    double xa[D], xs[D], kb[T][T];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      xa[a] = st.Xh[((size_t)b * st.cap_points + p) * D + a];
      xs[a] = xb[h * D + a];
    }
    kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
    for (int ta = 0; ta < T; ++ta)
#pragma unroll
      for (int tb = 0; tb < T; ++tb) W[(size_t)(m + r0 + ta) * q + h * T + tb] = kb[ta][tb];
  }
  __syncthreads();

  // W row of storage column t (the factor's column order): t < m shared, [m, mo) padding (none), t >= mo own
  const uint32_t a_lane = a_lane_off(gid, tig);
  int own[PM_MAXOWN];  // this warp's column blocks
  int nown = 0;
  for (int cb = warp; cb < QB && nown < PM_MAXOWN; cb += nw) own[nown++] = cb;

  // ---- 1: shared rows ------------------------------------------------------------------------------------------
  {
    const int Pm = (m + 7) >> 3;
    const double* gL = st.LooP + (size_t)j * subpanel_off(Pm, 0);
    for (int o = 0; o < nown; ++o) {
      const int colB = own[o] * 8 + gid;  // B-fragment column of this lane
      for (int p8 = Pm - 1; p8 >= 0; --p8) {
        const double* ap = gL + subpanel_off(p8, 0) + a_lane / 8;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < 2 * p8 + 2; k += 2) {
          const double a0 = __ldcg(ap + k * 32), a1 = __ldcg(ap + k * 32 + 32);
          const int r0 = 4 * k + tig, r1 = r0 + 4;
          const double b0 = (r0 < m && colB < q) ? __ldcg(W + (size_t)r0 * q + colB) : 0.0;
          const double b1 = (r1 < m && colB < q) ? __ldcg(W + (size_t)r1 * q + colB) : 0.0;
          dmma(acc[0], acc[1], a0, b0);
          dmma(acc[2], acc[3], a1, b1);
        }
        // mma.sync: every lane's reads of this tile-row's inputs have completed
        const int row = 8 * p8 + gid, col = own[o] * 8 + 2 * tig;
        if (row < m) {
          if (col < q) __stcg(W + (size_t)row * q + col, acc[0] + acc[2]);
          if (col + 1 < q) __stcg(W + (size_t)row * q + col + 1, acc[1] + acc[3]);
        }
        __syncwarp();
      }
    }
  }

  // ---- 2: own rows ----------------------------------------------------------------------------------------------
  const int P8 = (c + 7) >> 3;
  const double* Le = st.Lh + (size_t)b * st.elem_stride;
  const uint32_t sA_s = smem_u32(sA);
  for (int p = 0; p < P8; ++p) {
    const int n_off = mo + 8 * p;           // off-diagonal storage columns of this sub-panel
    const int ncol = n_off + 8;             // + its diagonal block
    const double* gp = Le + subpanel_off(p, mo);
    double acc[PM_MAXOWN][4];
#pragma unroll
    for (int o = 0; o < PM_MAXOWN; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.0;
    for (int s0 = 0; s0 < ncol; s0 += PM_SLAB) {
      const int s1 = min(ncol, s0 + PM_SLAB);
      __syncthreads();  // the previous slab is no longer read
      {
        const double2* src = (const double2*)(gp + (size_t)s0 * 8);
        double2* dst = (double2*)sA;
        for (int idx = tid; idx < (s1 - s0) * 4; idx += nt) dst[idx] = __ldcg(src + idx);
      }
      __syncthreads();
      const int k_end = (min(s1, n_off) - s0) >> 2;  // k-steps of off-diagonal columns in this slab
#pragma unroll
      for (int o = 0; o < PM_MAXOWN; ++o) {
        if (o < nown) {
          const int colB = own[o] * 8 + gid;
          const bool cok = colB < q;
          for (int k = 0; k < k_end; k += 2) {
            const double a0 = lds(sA_s + k * 256 + a_lane), a1 = lds(sA_s + k * 256 + 256 + a_lane);
            const int t0 = s0 + 4 * k + tig, t1 = t0 + 4;
            // storage column -> W row
            const int r0 = t0 < m ? t0 : (t0 >= mo ? t0 - mo + m : -1);
            const int r1 = t1 < m ? t1 : (t1 >= mo ? t1 - mo + m : -1);
            const double b0 = (cok && r0 >= 0) ? __ldcg(W + (size_t)r0 * q + colB) : 0.0;
            const double b1 = (cok && r1 >= 0) ? __ldcg(W + (size_t)r1 * q + colB) : 0.0;
            dmma(acc[o][0], acc[o][1], a0, b0);
            dmma(acc[o][2], acc[o][3], a1, b1);
          }
        }
      }
      if (s1 == ncol) {
        // the slab ends with the diagonal block: rhs = K - dot, w_blk = inv(D) rhs
        const uint32_t dblk = sA_s + (uint32_t)(n_off - s0) * 64;
        const int nvalid = min(8, c - 8 * p);
        double a0 = 0.0, a1 = 0.0;
        if (tig <= gid) a0 = lds(dblk + (uint32_t)sp_idx(gid, tig) * 8);
        if (tig + 4 <= gid) a1 = lds(dblk + (uint32_t)sp_idx(gid, tig + 4) * 8);
#pragma unroll
        for (int o = 0; o < PM_MAXOWN; ++o) {
          if (o < nown) {
            const int col = own[o] * 8 + 2 * tig;
            double* wrow = W + (size_t)(m + 8 * p + gid) * q;
            const bool live = gid < nvalid;
            if (live) {
              if (col < q) __stcg(wrow + col, __ldcg(wrow + col) - (acc[o][0] + acc[o][2]));
              if (col + 1 < q) __stcg(wrow + col + 1, __ldcg(wrow + col + 1) - (acc[o][1] + acc[o][3]));
            }
            __syncwarp();
            const int colB = own[o] * 8 + gid;
            const double b0 = (colB < q && tig < nvalid) ? __ldcg(W + (size_t)(m + 8 * p + tig) * q + colB) : 0.0;
            const double b1 = (colB < q && tig + 4 < nvalid) ? __ldcg(W + (size_t)(m + 8 * p + tig + 4) * q + colB) : 0.0;
            double d0 = 0.0, d1 = 0.0;
            dmma(d0, d1, a0, b0);
            dmma(d0, d1, a1, b1);  // mma.sync: every lane's rhs loads have completed
            if (live) {
              if (col < q) __stcg(wrow + col, d0);
              if (col + 1 < q) __stcg(wrow + col + 1, d1);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- 3: Sigma* (lower triangle) and mean ---------------------------------------------------------------------
  {
    const int n4 = (n + 3) >> 2;
    const int n_tiles = QB * (QB + 1) / 2;
    for (int tile = warp; tile < n_tiles; tile += nw) {
      int rb = 0;
      while ((rb + 1) * (rb + 2) / 2 <= tile) ++rb;
      const int sb = tile - rb * (rb + 1) / 2;
      const int ca = rb * 8 + gid, cbb = sb * 8 + gid;
      const bool oka = ca < q, okb = cbb < q;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < n4; k += 2) {
        const int r0 = 4 * k + tig, r1 = r0 + 4;
        const double a0 = (oka && r0 < n) ? __ldcg(W + (size_t)r0 * q + ca) : 0.0;
        const double a1 = (oka && r1 < n) ? __ldcg(W + (size_t)r1 * q + ca) : 0.0;
        const double b0 = (okb && r0 < n) ? __ldcg(W + (size_t)r0 * q + cbb) : 0.0;
        const double b1 = (okb && r1 < n) ? __ldcg(W + (size_t)r1 * q + cbb) : 0.0;
        dmma(acc[0], acc[1], a0, b0);
        dmma(acc[2], acc[3], a1, b1);
      }
      const int r = rb * 8 + gid;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int s = sb * 8 + 2 * tig + hh;
        if (r < q && s <= r) {
          const double kss = cov_scalar(xb + (size_t)(r / T) * D, r % T, xb + (size_t)(s / T) * D, s % T,
                                        st.ls + j * D, os, D);
          S[(size_t)r * q + s] = kss - (acc[hh] + acc[2 + hh]);
        }
      }
    }
    const double* beta_o = st.beta_o + (size_t)j * m;
    const double* beta_h = st.beta_h + (size_t)b * st.c_cap;
    for (int r = tid; r < q; r += nt) {
      double a = 0.0;
      for (int i = 0; i < m; ++i) a += __ldcg(W + (size_t)i * q + r) * beta_o[i];
      for (int i = m; i < n; ++i) a += __ldcg(W + (size_t)i * q + r) * beta_h[i - m];
      mu[r] = a;
    }
  }
  __syncthreads();

  for (int r = tid; r < q; r += nt) {
    if (mean) mean[(size_t)b * q + r] = mu[r];
    if (var) var[(size_t)b * q + r] = fmax(S[(size_t)r * q + r], GP_MIN_VARIANCE);
  }
  if (eps) block_sample(st, b, H, eps, opts, y, jitter_level);
}
