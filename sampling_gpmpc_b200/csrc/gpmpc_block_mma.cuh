// K2m: the SQP-mode model call (any H) on the FP64 tensor cores.
//
// The joint posterior of H test points (q = H*T scalars) given n = m + c training scalars is dominated by the
// triangular solve W = L^{-1} K_{o*} with q right-hand sides (q n^2 / 2 MACs) and by Sigma* = K** - W^T W (q^2 n / 2):
// contraction bound once the hallucinated set has grown over a few SQP iterations (car-residual config: q = 150, n up
// to several thousand), where the scalar kernel k_posterior (gpmpc_block.cuh, kept as the substitution-based reference
// semantics) reaches ~1 % of the FP64 peak.  The right-hand sides are cut into 8-column blocks; columns never interact
// in the solve, so a warp owns ONE block for the whole solve and the blocks of an element are spread over several
// CTAs (grid = elements x splits: SQP mode has only tens to hundreds of elements, fewer than SMs):
//  k_pm_solve (grid B x ceil(QB / 4), 4 warps):
//   0  K_{o*}: one exp per (training point, test point) pair, the T x T derivative block from it
//   1  shared rows: W_o = inv(L_oo) K_o, tile-rows last to first, in place (as K1 phase B)
//   2  own rows, left-looking over ROW BLOCKS of 8 sub-panels (64 rows) of the element's factor stream: the stream is
//      staged in shared memory by TMA bulk copies (all column blocks need the same A operand); for the columns left of
//      the block every k-step reads one W fragment back through L2 and feeds 8 DMMA chains (one per sub-panel), then the
//      block's triangle is solved sub-panel by sub-panel: rhs = K - dot, and the 8 x 8 diagonal block is applied as
//      inv(D) rhs (two more DMMAs; inverse kept in the block's upper triangle)
//  k_pm_gram (grid B x splits, 4 warps):
//   3  Sigma* = K** - W^T W by 8 x 8 output tiles (A and B fragments are the same access pattern on W), mean = W^T beta
//  k_pm_finish (grid B): mean / variance out, then the draw / post-processing of gpmpc_block.cuh (block_sample, packed
//  Cholesky in shared memory).
// W [n][q], S [q][q], mu [q], xc keep the layouts of the scalar kernel: k_sample / k_append consume either.
#pragma once
#include "gpmpc_block.cuh"
#include "gpmpc_step.cuh"

#ifndef PM_WARPS
#define PM_WARPS 4       // warps = column blocks per CTA of k_pm_solve
#endif
#define PM_NSP 8         // sub-panels per row block of the solve (64 rows share every W fragment)
#define PM_SLABC 64      // storage columns per sub-panel per stage: a buffer is PM_NSP x PM_SLABC x 64 B = 32 KB, two resident
#define PM_MAX_Q 2048    // test scalars per call served by this path (QB <= 256 column blocks)

template <int D, int T>
__global__ void __launch_bounds__(PM_WARPS * 32, PM_WARPS <= 4 ? 3 : (PM_WARPS <= 8 ? 2 : 1))
k_pm_solve(DevState st, const double* __restrict__ x, int H) {
  extern __shared__ __align__(128) double sA[];  // [2][PM_NSP][PM_SLABC * 8] two stages of the factor stream (k-block layout)
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int j = b % st.g_ny;
  const int m = st.m, mo = st.mo, c = st.c, q = H * T;
  const int QB = (q + 7) >> 3;
  double* W = st.W + (size_t)b * st.W_stride;
  const double* xb = x + (size_t)b * H * D;
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j * D + a];
  const double os = st.os[j];

  // ---- 0: kernel matrix, this CTA's columns [c_lo, c_hi) -----------------------------------------------------
  const int c_lo = blockIdx.y * PM_WARPS * 8, c_hi = min(q, c_lo + PM_WARPS * 8);
  const int h_lo = c_lo / T, h_n = (c_hi - 1) / T - h_lo + 1;  // test points touching those columns
  if (blockIdx.y == 0) {
    double* xc = st.xc + (size_t)b * H * D;
    for (int idx = tid; idx < H * D; idx += nt) xc[idx] = xb[idx];
  }
  for (int idx = tid; idx < m * h_n; idx += nt) {
    const int i = idx / h_n, h = h_lo + idx - i * h_n;
    double xs[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xs[a] = xb[h * D + a];
    const int pt = st.obs_pt[i], ta = st.obs_task[i];
    if (real_point_full<T>(st, i, pt, ta, m)) {  // one exp for the point's T x T block (thread of its task-0 row)
      if (ta != 0) continue;
      double xa[D], kb[T][T];
#pragma unroll
      for (int a = 0; a < D; ++a) xa[a] = st.Xr[(size_t)pt * D + a];
      kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
      for (int t2 = 0; t2 < T; ++t2)
#pragma unroll
        for (int tb = 0; tb < T; ++tb) {
          const int col = h * T + tb;
          if (col >= c_lo && col < c_hi) W[(size_t)(i + t2) * q + col] = kb[t2][tb];
        }
    } else {
      double out[T];
      kernel_row<D, T>(st.Xr + (size_t)pt * D, ta, xs, il, os, out);
#pragma unroll
      for (int tb = 0; tb < T; ++tb) {
        const int col = h * T + tb;
        if (col >= c_lo && col < c_hi) W[(size_t)i * q + col] = out[tb];
      }
    }
  }
  for (int idx = tid; idx < st.np * h_n; idx += nt) {
    const int p = idx / h_n, h = h_lo + idx - p * h_n;
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
      for (int tb = 0; tb < T; ++tb) {
        const int col = h * T + tb;
        if (col >= c_lo && col < c_hi) W[(size_t)(m + r0 + ta) * q + col] = kb[ta][tb];
      }
  }
  __syncthreads();

  const uint32_t a_lane = a_lane_off(gid, tig);
  const int cb = blockIdx.y * PM_WARPS + warp;   // this warp's column block
  const bool has = cb < QB;                      // (idle warps still take part in the CTA barriers below)
  const int colB = cb * 8 + gid;                 // B-fragment column of this lane
  const bool cok = has && colB < q;
  const int colC = cb * 8 + 2 * tig;             // C-fragment columns colC, colC + 1

  // ---- 1: shared rows ------------------------------------------------------------------------------------------
  if (has) {
    const int Pm = (m + 7) >> 3;
    const double* gL = st.LooP + (size_t)j * subpanel_off(Pm, 0);
    for (int p8 = Pm - 1; p8 >= 0; --p8) {
      const double* ap = gL + subpanel_off(p8, 0) + a_lane / 8;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < 2 * p8 + 2; k += 2) {
        const double a0 = __ldcg(ap + k * 32), a1 = __ldcg(ap + k * 32 + 32);
        const int r0 = 4 * k + tig, r1 = r0 + 4;
        const double b0 = (r0 < m && cok) ? __ldcg(W + (size_t)r0 * q + colB) : 0.0;
        const double b1 = (r1 < m && cok) ? __ldcg(W + (size_t)r1 * q + colB) : 0.0;
        dmma(acc[0], acc[1], a0, b0);
        dmma(acc[2], acc[3], a1, b1);
      }
      // mma.sync: every lane's reads of this tile-row's inputs have completed
      const int row = 8 * p8 + gid;
      if (row < m) {
        if (colC < q) __stcg(W + (size_t)row * q + colC, acc[0] + acc[2]);
        if (colC + 1 < q) __stcg(W + (size_t)row * q + colC + 1, acc[1] + acc[3]);
      }
      __syncwarp();
    }
  }

  // ---- 2: own rows ----------------------------------------------------------------------------------------------
  // Left-looking over ROW BLOCKS of PM_NSP sub-panels (64 rows): for the columns left of the block every k-step loads ONE
  // B fragment of W (an L2 round trip) and feeds PM_NSP tensor-core products, one per sub-panel of the block -- W is
  // re-read once per 64 rows instead of once per 8 (the 8-row version was bound by exactly these re-reads:
  // profiles/r1_sqp_growth.txt).  The block's own triangle is then solved sub-panel by sub-panel.
  // The factor stream arrives through two 32 KB buffers: a "stage" is either the same <= PM_SLABC columns of all the
  // block's sub-panels (PM_NSP TMA bulk copies on one mbarrier) or the block's triangle; stage i + 1 is in flight while
  // the warps work on stage i.  A warp owns its column block for the whole solve, so W needs no CTA-wide barrier.
  const int P8 = (c + 7) >> 3;
  const int NBLK = (P8 + PM_NSP - 1) / PM_NSP;
  const double* Le = st.Lh + (size_t)b * st.elem_stride;
  const uint32_t sA_s = smem_u32(sA);
  // W row of storage column t (the factor's column order): t < m shared, [m, mo) padding (none), t >= mo own
  auto load_b = [&](double (&bv)[8], int kk0, int kk_end) {  // k-steps kk0 .. kk0+7 (global, 4 storage columns each)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = 4 * (kk0 + u) + tig;
      const int r = t < m ? t : (t >= mo ? t - mo + m : -1);
      bv[u] = (cok && r >= 0 && kk0 + u < kk_end) ? __ldcg(W + (size_t)r * q + colB) : 0.0;
    }
  };
  __shared__ uint64_t bars[2];
  uint64_t l2_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_stream));
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  auto bulk = [&](double* dst, const double* src, uint32_t bytes, uint64_t* bar) {
    // L2 evict-first: the factor passes through once per CTA and must not push W out of L2
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_stream) : "memory");
  };
  auto n_part1 = [&](int blk) { return (mo + 8 * PM_NSP * blk + PM_SLABC - 1) / PM_SLABC; };  // stages before the triangle
  auto issue = [&](int blk, int stg, int buf) {  // thread 0 only
    const int p0 = blk * PM_NSP, nsp = min(PM_NSP, P8 - p0), base = mo + 8 * p0;
    double* dst = sA + (size_t)buf * PM_NSP * PM_SLABC * 8;
    if (stg < n_part1(blk)) {
      const int s0 = stg * PM_SLABC, s1 = min(base, s0 + PM_SLABC);
      const uint32_t each = (uint32_t)(s1 - s0) * 64u;
      mbar_expect_tx(&bars[buf], each * nsp);
      for (int jj = 0; jj < nsp; ++jj)
        bulk(dst + (size_t)jj * PM_SLABC * 8, Le + subpanel_off(p0 + jj, mo) + (size_t)s0 * 8, each, &bars[buf]);
    } else {  // triangle: sub-panel jj contributes its columns [base, base + 8 jj + 8), packed one after the other
      mbar_expect_tx(&bars[buf], 256u * nsp * (nsp + 1));
      for (int jj = 0; jj < nsp; ++jj)
        bulk(dst + 32 * jj * (jj + 1), Le + subpanel_off(p0 + jj, mo) + (size_t)base * 8, (uint32_t)(8 * jj + 8) * 64u, &bars[buf]);
    }
  };
  if (tid == 0 && NBLK > 0) issue(0, 0, 0);
  int it = 0;  // stage counter: buffer it & 1, barrier phase (it >> 1) & 1
  uint32_t buf_s = sA_s;
  auto next_stage = [&](int blk, int stg) {  // all threads: hand over to stage (blk, stg), start the copy of its successor
    __syncthreads();  // every warp is done with the previous stage's buffer
    if (tid == 0) {
      if (stg < n_part1(blk)) issue(blk, stg + 1, (it + 1) & 1);
      else if (blk + 1 < NBLK) issue(blk + 1, 0, (it + 1) & 1);
    }
    mbar_wait(&bars[it & 1], (it >> 1) & 1);
    buf_s = sA_s + (uint32_t)(it & 1) * PM_NSP * PM_SLABC * 64;
    ++it;
  };
  for (int blk = 0; blk < NBLK; ++blk) {
    const int p0 = blk * PM_NSP, nsp = min(PM_NSP, P8 - p0), base = mo + 8 * p0;
    const int kk_end = base >> 2, rounds = (kk_end + 7) >> 3, n1 = n_part1(blk);
    double acc[PM_NSP][4];
#pragma unroll
    for (int jj = 0; jj < PM_NSP; ++jj) acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.0;
    // -- columns left of the block: one W fragment per k-step, PM_NSP products
    double bn[8];
    load_b(bn, 0, kk_end);
    for (int r = 0; r < rounds; ++r) {
      if ((r & (PM_SLABC / 32 - 1)) == 0) next_stage(blk, r / (PM_SLABC / 32));
      double bv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) bv[u] = bn[u];
      if (r + 1 < rounds) load_b(bn, 8 * (r + 1), kk_end);
      const uint32_t ab = buf_s + (uint32_t)((8 * r) & (PM_SLABC / 4 - 1)) * 256 + a_lane;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (8 * r + u < kk_end) {
#pragma unroll
          for (int jj = 0; jj < PM_NSP; ++jj)
            if (jj < nsp) dmma(acc[jj][2 * (u & 1)], acc[jj][2 * (u & 1) + 1], lds(ab + jj * PM_SLABC * 64 + u * 256), bv[u]);
        }
      }
    }
    // -- the block's triangle, sub-panel by sub-panel
    next_stage(blk, n1);
    if (has) {
      // The block's finished sub-panels stay in registers (C-fragment layout: lane (gid, tig) holds rows gid, columns 2 tig and
      // 2 tig + 1 of its column block) and reach the later sub-panels' products as B fragments through warp shuffles -- lane
      // (gid, tig) needs row 4 kk + tig, column gid, i.e. lane (4 kk + tig, gid >> 1)'s element gid & 1 -- instead of a store
      // to W and a dependent L2 load per k-step; the same for rhs -> inv(D) rhs.  Same values, same products, same order.
      double wres[PM_NSP][2];
      const int src_t = gid >> 1;
      const bool odd = gid & 1;
#pragma unroll
      for (int jj = 0; jj < PM_NSP; ++jj) {
        if (jj < nsp) {
          const int p = p0 + jj;
          const uint32_t tri = buf_s + (uint32_t)(32 * jj * (jj + 1)) * 8;
          const int nvalid = min(8, c - 8 * p);
          const bool live = gid < nvalid;
          double* wrow = W + (size_t)(m + 8 * p + gid) * q;
          // the kernel entries of this sub-panel's rows (written by phase 0): in flight during the in-block products
          const double k0 = (live && colC < q) ? __ldcg(wrow + colC) : 0.0;
          const double k1 = (live && colC + 1 < q) ? __ldcg(wrow + colC + 1) : 0.0;
#pragma unroll
          for (int kk = 0; kk < 2 * PM_NSP; ++kk) {  // in-block columns base .. base + 8 jj: the block's earlier sub-panels
            if (kk < 2 * jj) {
              const int src = 4 * (4 * (kk & 1) + tig) + src_t;
              const double v0 = __shfl_sync(0xffffffffu, wres[kk >> 1][0], src);
              const double v1 = __shfl_sync(0xffffffffu, wres[kk >> 1][1], src);
              const double bq = cok ? (odd ? v1 : v0) : 0.0;
              dmma(acc[jj][2 * (kk & 1)], acc[jj][2 * (kk & 1) + 1], lds(tri + kk * 256 + a_lane), bq);
            }
          }
          // rhs = K - dot, w_blk = inv(D) rhs (diagonal block = the last 8 columns of this sub-panel's part)
          const uint32_t dblk = tri + (uint32_t)(8 * jj) * 64;
          double a0 = 0.0, a1 = 0.0;
          if (tig <= gid) a0 = lds(dblk + (uint32_t)sp_idx(gid, tig) * 8);
          if (tig + 4 <= gid) a1 = lds(dblk + (uint32_t)sp_idx(gid, tig + 4) * 8);
          const double r0 = (live && colC < q) ? k0 - (acc[jj][0] + acc[jj][2]) : 0.0;
          const double r1 = (live && colC + 1 < q) ? k1 - (acc[jj][1] + acc[jj][3]) : 0.0;
          const double u0 = __shfl_sync(0xffffffffu, r0, 4 * tig + src_t), u1 = __shfl_sync(0xffffffffu, r1, 4 * tig + src_t);
          const double w0 = __shfl_sync(0xffffffffu, r0, 4 * (tig + 4) + src_t), w1 = __shfl_sync(0xffffffffu, r1, 4 * (tig + 4) + src_t);
          const double b0 = (cok && tig < nvalid) ? (odd ? u1 : u0) : 0.0;
          const double b1 = (cok && tig + 4 < nvalid) ? (odd ? w1 : w0) : 0.0;
          double d0 = 0.0, d1 = 0.0;
          dmma(d0, d1, a0, b0);
          dmma(d0, d1, a1, b1);
          if (live) {
            if (colC < q) __stcg(wrow + colC, d0);
            if (colC + 1 < q) __stcg(wrow + colC + 1, d1);
          }
          wres[jj][0] = d0;
          wres[jj][1] = d1;
        }
      }
      __syncwarp();  // this block's rows of W are complete before the next block's left part reads them back
    }
  }
}

template <int T>
__global__ void __launch_bounds__(PM_WARPS * 32)
k_pm_gram(DevState st, const double* __restrict__ x, int H) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int j = b % st.g_ny, d = st.d;
  const int m = st.m, n = m + st.c, q = H * T;
  const int QB = (q + 7) >> 3;
  const double* W = st.W + (size_t)b * st.W_stride;
  double* S = st.S + (size_t)b * q * q;
  double* mu = st.mu + (size_t)b * q;
  const double* xb = x + (size_t)b * H * d;
  const double os = st.os[j];
  const int n4 = (n + 3) >> 2;
  const int n_tiles = QB * (QB + 1) / 2;
  const int gw = blockIdx.y * nw + warp, gstride = gridDim.y * nw;
  for (int tile = gw; tile < n_tiles + QB; tile += gstride) {
    if (tile >= n_tiles) {
      // mean of column block cb: lanes (gid, tig) take rows 4k + tig, reduced over tig
      const int col = (tile - n_tiles) * 8 + gid;
      const double* beta_o = st.beta_o + (size_t)j * m;
      const double* beta_h = st.beta_h + (size_t)b * st.c_cap;
      double a = 0.0, a_real = 0.0;
      if (col < q)
        for (int r = tig; r < n; r += 4) {
          a = fma(__ldcg(W + (size_t)r * q + col), r < m ? beta_o[r] : beta_h[r - m], a);
          if (r + 4 >= m && r < m) a_real = a;  // the partial sum after this lane's last real row: same order as a pass over m rows
        }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (tig == 0 && col < q) mu[col] = a;
      if (st.mu2) {
        a_real += __shfl_xor_sync(0xffffffffu, a_real, 1);
        a_real += __shfl_xor_sync(0xffffffffu, a_real, 2);
        if (tig == 0 && col < q) st.mu2[(size_t)b * q + col] = a_real;
      }
      continue;
    }
    int rb = 0;
    while ((rb + 1) * (rb + 2) / 2 <= tile) ++rb;
    const int sb = tile - rb * (rb + 1) / 2;
    const int ca = rb * 8 + gid, cbb = sb * 8 + gid;
    const bool oka = ca < q, okb = cbb < q;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double acr[4] = {0.0, 0.0, 0.0, 0.0};  // the same sums over the REAL rows only (st.S2): rows >= m enter as zeros
    const bool want_real = st.S2 != nullptr;
    for (int k = 0; k < n4; k += 8) {
      double av[8], bv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = 4 * (k + u) + tig;
        av[u] = (oka && r < n) ? __ldcg(W + (size_t)r * q + ca) : 0.0;
        bv[u] = (okb && r < n) ? __ldcg(W + (size_t)r * q + cbb) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) dmma(acc[2 * (u & 1)], acc[2 * (u & 1) + 1], av[u], bv[u]);
      if (want_real && 4 * k < m) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool real = 4 * (k + u) + tig < m;
          dmma(acr[2 * (u & 1)], acr[2 * (u & 1) + 1], real ? av[u] : 0.0, real ? bv[u] : 0.0);
        }
      }
    }
    const int r = rb * 8 + gid;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int s = sb * 8 + 2 * tig + hh;
      if (r < q && s <= r) {
        const double kss = cov_scalar(xb + (size_t)(r / T) * d, r % T, xb + (size_t)(s / T) * d, s % T,
                                      st.ls + j * d, os, d);
        S[(size_t)r * q + s] = kss - (acc[hh] + acc[2 + hh]);
        if (want_real) st.S2[(size_t)b * q * q + (size_t)r * q + s] = kss - (acr[hh] + acr[2 + hh]);
      }
    }
  }
}

// mean / variance out, then the draw (tri_ok: q(q+1)/2 doubles of dynamic shared memory for the packed Cholesky).
// pre_kind != 0 (grid.y = 2): the CTAs with blockIdx.y = 1 factorise, at the same time, the matrix the append that follows
// will need -- chol(Sigma_app + noise), Sigma_app = st.S (1) or st.S2 (2: the hallucinated set is about to be reset), all
// scalars active -- into st.Lpre: the second q x q Cholesky of an SQP linearisation leaves the critical path
// (same fill expression and the same block_cholesky_packed as k_append: bit-identical factor).
__global__ void __launch_bounds__(BLK_THREADS)
k_pm_finish(DevState st, int H, double* __restrict__ mean, double* __restrict__ var, const double* __restrict__ eps,
            gpmpc_sample_opts opts, double* __restrict__ y, int* __restrict__ jitter_level, int tri_ok, int pre_kind) {
  extern __shared__ __align__(16) double dyn_tri[];
  const int b = blockIdx.x, q = H * st.T;
  if (blockIdx.y == 1) {
    const int tid = threadIdx.x, nt = blockDim.x, T = st.T;
    const double* Sa = (pre_kind == 2 ? st.S2 : st.S) + (size_t)b * q * q;
    const double* noise = st.noise + (b % st.g_ny) * T;
    const int ntri = q * (q + 1) / 2;
    for (int idx = tid; idx < q * q; idx += nt) {
      const int r = idx / q, s2 = idx - r * q;
      if (s2 <= r) dyn_tri[r * (r + 1) / 2 + s2] = Sa[idx] + (s2 == r ? noise[r % T] : 0.0);
    }
    const int info = block_cholesky_packed(dyn_tri, q);
    double* out = st.Lpre + (size_t)b * (ntri + 1);
    for (int idx = tid; idx < ntri; idx += nt) out[idx] = dyn_tri[idx];
    if (tid == 0) out[ntri] = (double)info;
    return;
  }
  const double* S = st.S + (size_t)b * q * q;
  const double* mu = st.mu + (size_t)b * q;
  for (int r = threadIdx.x; r < q; r += blockDim.x) {
    if (mean) mean[(size_t)b * q + r] = mu[r];
    if (var) var[(size_t)b * q + r] = fmax(S[(size_t)r * q + r], GP_MIN_VARIANCE);
  }
  if (eps) block_sample(st, b, H, eps, opts, y, jitter_level, tri_ok ? dyn_tri : nullptr);
}
