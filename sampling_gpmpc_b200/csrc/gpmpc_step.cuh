// K1: fused rollout step (H = 1): posterior + draw + post-processing + rank-T append in one launch.
//
// One WARP owns one batch element (sample s, output j) at a time and loops over samples (persistent CTAs:
// blockIdx.y = output j, so the warps of a CTA share L_oo, the observed real inputs and beta_o in shared
// memory).  The step is a forward substitution  w = L^{-1} k(X, x*)  against the bordered factor
// [[L_oo, 0], [V, L_hh]], done LEFT-LOOKING over SUB-PANELS of 8 rows (layout: gpmpc_state.cuh) on the FP64
// tensor cores: one mma.sync.m8n8k4.f64 multiplies an 8-row x 4-column tile of L with the 4 x T tile of w
// (T <= 7 right-hand sides = the N dimension; unused columns are don't-care), so the inner loop is two shared
// loads and one DMMA per 4 factor columns and needs no cross-lane reduction.
//
//   A  kernel vector k(X, x*): lanes over training scalars, T right-hand sides each, written to the per-warp
//      shared array wv[storage column][T]  (in place: wv holds k first, w = L^{-1} k afterwards)
//   B  the m shared rows against L_oo (shared memory, sub-panel layout)
//   C  the element's own c rows, streamed from HBM exactly once:  the element's factor is one contiguous
//      stream of sub-panels, cut into chunks of STEP_SEG column groups (2 KB) that one elected lane pulls into
//      a per-warp ring of STEP_NST shared-memory slots with TMA bulk copies (cp.async.bulk + mbarrier
//      complete_tx).  The ring runs STEP_NST-1 chunks ahead of the consumer and ACROSS elements: while an
//      element's epilogue runs, the first chunks of the warp's next element are already in flight.
//      Per sub-panel: dot = L[rows][cols < n_off] w  (DMMA chain, two accumulator sets), rhs = k - dot, then the
//      8 x 8 diagonal block is applied as w_blk = inv(D) rhs with two more DMMAs (inv(D) is kept transposed in
//      the block's upper triangle) -- no substitution chain anywhere.
//   D  W^T [W | beta] with the same DMMA loop: Sigma* = K** - W^T W, mean = W^T beta
//   E  T x T Cholesky with GPyTorch's jitter ladder, y = mean + L eps, zero-variance / truncation
//   F  rank-T append: w goes to the T new rows' column groups, then the touched diagonal blocks' inverses
//
// HBM traffic per element-step = its own factor read once (8-row granularity) + T new rows written once + O(c)
// inputs: HBM-bound by design (DESIGN.md "Roofline"); the host counts the algorithmic bytes per launch.
#pragma once
#include "gpmpc_state.cuh"

#ifndef STEP_WARPS
#define STEP_WARPS 4
#endif
#define FULL_MASK 0xffffffffu
#define STEP_SEG 32  // 8-row column groups per TMA chunk / ring slot (32 * 64 B = 2 KB); multiple of 8
#define STEP_NST 4   // ring slots per warp (power of 2): one being consumed, three in flight

// ---- TMA bulk copy + mbarrier (one ring per warp; the warp is its own producer and consumer) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

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

// cov( task ta at xa , tasks 0..T-1 at xs ), r = xa - xs  
template <int D, int T>
__device__ __forceinline__ void kernel_row(const double* __restrict__ xa, int ta, const double (&xs)[D],
                                           const double (&il)[D], double os, double (&out)[T]) {
  double g[D], sq = 0.0, ga = 0.0, il2 = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const double r = xa[a] - xs[a];
    const double t = r * il[a];
    sq = fma(t, t, sq);
    g[a] = t * il[a];  // r_a / l_a^2
    if (a == ta - 1) { ga = g[a]; il2 = il[a] * il[a]; }
  }
  const double k0 = os * exp(-0.5 * sq);
  if (ta == 0) {
    out[0] = k0;
#pragma unroll
    for (int tb = 1; tb < T; ++tb) out[tb] = k0 * g[tb - 1];
  } else {
    out[0] = -k0 * ga;
#pragma unroll
    for (int tb = 1; tb < T; ++tb) {
      double h = -ga * g[tb - 1];
      if (tb == ta) h += il2;
      out[tb] = k0 * h;
    }
  }
}

// T right-hand-side values of one storage column from wv (row stride TP doubles, 16-byte aligned when TP is even)
template <int T>
__device__ __forceinline__ void load_w(const double* __restrict__ p, double (&w)[T]) {
  if constexpr (T == 1) {
    w[0] = p[0];
  } else {
#pragma unroll
    for (int r = 0; r + 1 < T; r += 2) {
      const double2 v = *reinterpret_cast<const double2*>(p + r);
      w[r] = v.x;
      w[r + 1] = v.y;
    }
    if constexpr (T & 1) w[T - 1] = p[T - 1];
  }
}

// D(8x8) += A(8x4) B(4x8) in fp64 on the tensor cores.  Fragments (lane = 4*gid + tig): a = A[gid][tig],
// b = B[tig][gid], c0/c1 = C[gid][2*tig], C[gid][2*tig+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// c += L[8 rows][4*n4 columns] * w[4*n4 columns][.]: column groups contiguous at `lp` (lane pointer already at
// + tig*8 + gid), w rows at `wp` (lane pointer already at + tig*TP + gid).  n4 is even; two accumulator sets.
template <int TP>
__device__ __forceinline__ void mma_accumulate(double (&c)[4], const double* __restrict__ lp,
                                               const double* __restrict__ wp, int n4) {
#pragma unroll 2
  for (int it = 0; it < n4; it += 2) {
    const double a0 = lp[it * 32], b0 = wp[it * 4 * TP];
    const double a1 = lp[it * 32 + 32], b1 = wp[it * 4 * TP + 4 * TP];
    dmma(c[0], c[1], a0, b0);
    dmma(c[2], c[3], a1, b1);
  }
}

// Finishes one sub-panel: rows n_off .. n_off+7 of wv (wv_blk) hold the kernel entries, c the off-diagonal dot
// products (C layout).  rhs = k - dot goes back to wv, w_blk = inv(D) rhs comes out of two DMMAs and replaces it.
// Rows >= nvalid (padding / not yet appended) are forced to 0.
template <int T, int TP>
__device__ __forceinline__ void subpanel_finish(const double (&c)[4], const double* __restrict__ dblk,
                                                double* __restrict__ wv_blk, int nvalid, int lane) {
  const int gid = lane >> 2, tig = lane & 3;
  const bool v0 = 2 * tig < T, v1 = 2 * tig + 1 < T, live = gid < nvalid;
  double* mine = wv_blk + gid * TP + 2 * tig;
  if (v0) mine[0] = live ? mine[0] - (c[0] + c[2]) : 0.0;
  if (v1) mine[1] = live ? mine[1] - (c[1] + c[3]) : 0.0;
  // inv(D)[gid][k], k = tig and tig + 4: slot (row k, column gid) of the block, zero above the diagonal
  const double a0 = tig <= gid ? dblk[gid * 8 + tig] : 0.0;
  const double a1 = tig + 4 <= gid ? dblk[gid * 8 + tig + 4] : 0.0;
  __syncwarp();
  const double b0 = wv_blk[tig * TP + gid], b1 = wv_blk[(tig + 4) * TP + gid];
  double d0 = 0.0, d1 = 0.0;
  dmma(d0, d1, a0, b0);
  dmma(d0, d1, a1, b1);
  __syncwarp();
  if (v0) mine[0] = live ? d0 : 0.0;
  if (v1) mine[1] = live ? d1 : 0.0;
  __syncwarp();
}

template <int D, int T, bool LOO_SMEM>
__global__ void __launch_bounds__(STEP_WARPS * 32)
k_step(DevState st, const double* __restrict__ x, const double* __restrict__ eps, gpmpc_sample_opts opts,
       double* __restrict__ mean, double* __restrict__ var, double* __restrict__ y,
       int* __restrict__ jitter_level, int grow_factor) {
  constexpr int TP = T == 1 ? 1 : ((T + 1) & ~1);
  extern __shared__ __align__(128) double smem[];
  const int j_out = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, mo = st.mo, c = st.c;
  const int Pm = (m + 7) >> 3, P8 = (c + 7) >> 3;
  const int loop_sz = (int)subpanel_off(Pm, 0);
  const int m_even = (m + 1) & ~1;

  // ---- shared-memory carve-up (sizes mirrored by launch_step in gpmpc_api.cu) -----------------------------
  double* sL = smem;                                         // [loop_sz] L_oo sub-panels (only if LOO_SMEM)
  double* sXo = sL + (LOO_SMEM ? loop_sz : 0);               // [m_even*D] input of observed real scalar i
  double* sBo = sXo + (size_t)m_even * D;                    // [m_even]
  int* sTo = (int*)(sBo + m_even);                           // [2*m_even] ints: task of observed real scalar i
  const int wv_rows = mo + 8 * P8;
  const int wv_sz = (wv_rows * TP + 8 + 15) & ~15;           // per warp, doubles (+8: don't-care reads of idle lanes)
  const int wb_sz = (wv_rows + 15) & ~15;
  const int per_warp = wv_sz + wb_sz + 64 + STEP_NST * STEP_SEG * 8;
  double* warp_base = (double*)(sTo + 2 * m_even);
  warp_base = (double*)(((uintptr_t)warp_base + 127) & ~(uintptr_t)127);
  double* wv = warp_base + (size_t)warp * per_warp;          // [wv_rows][TP]  k, then w
  double* wb = wv + wv_sz;                                   // [wv_rows]      beta by storage column
  double* sc = wb + wb_sz;                                   // [8][8]         W^T [W | beta]
  double* ring = sc + 64;                                    // [STEP_NST][STEP_SEG*8]
  uint64_t* bars = (uint64_t*)(warp_base + (size_t)STEP_WARPS * per_warp) + warp * STEP_NST;

  if (threadIdx.x < STEP_WARPS * STEP_NST) mbar_init(bars - warp * STEP_NST + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const double* gL = st.LooP + (size_t)j_out * loop_sz;
  if (LOO_SMEM)
    for (int idx = threadIdx.x; idx < loop_sz; idx += blockDim.x) sL[idx] = gL[idx];
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
    const double* xp = st.Xr + (size_t)st.obs_pt[idx] * D;
#pragma unroll
    for (int a = 0; a < D; ++a) sXo[idx * D + a] = xp[a];
    sBo[idx] = st.beta_o[(size_t)j_out * m + idx];
    sTo[idx] = st.obs_task[idx];
  }
  // padding rows [m, mo) and rows >= c of wv / wb are zero for the whole launch
  for (int idx = lane; idx < wv_sz + wb_sz; idx += 32) wv[idx] = 0.0;
  __syncthreads();
  for (int idx = lane; idx < m; idx += 32) wb[idx] = sBo[idx];
  __syncwarp();
  // no block-level synchronisation below this line: every warp runs its own element loop

  const double* Lp = LOO_SMEM ? sL : gL;
  const int nwarps_total = gridDim.x * STEP_WARPS;
  const int s_first = blockIdx.x * STEP_WARPS + warp;
  double il[D];
#pragma unroll
  for (int a = 0; a < D; ++a) il[a] = 1.0 / st.ls[j_out * D + a];
  const double os = st.os[j_out];

  // ---- producer (warp-uniform state; lane 0 issues): the own factors of this warp's elements, in
  //      consumption order, as one sequence of chunks that never cross a sub-panel boundary --------------------
  int prod_s = P8 > 0 ? s_first : st.ns;
  int prod_p8 = 0, prod_left = mo + 8;
  const double* prod_src = st.Lh + (size_t)(prod_s < st.ns ? prod_s * st.g_ny + j_out : 0) * st.elem_stride;
  unsigned issued = 0, consumed = 0;
  auto produce = [&]() {  // keeps STEP_NST chunks outstanding while anything is left
    while (issued < consumed + STEP_NST && prod_s < st.ns) {
      const int ng = min(STEP_SEG, prod_left);
      if (lane == 0) {
        const unsigned slot = issued & (STEP_NST - 1);
        mbar_expect_tx(bars + slot, (uint32_t)ng * 64u);
        tma_bulk_g2s(ring + (size_t)slot * STEP_SEG * 8, prod_src, (uint32_t)ng * 64u, bars + slot);
      }
      ++issued;
      prod_src += ng * 8;
      prod_left -= ng;
      if (prod_left == 0) {
        if (++prod_p8 == P8) {
          prod_p8 = 0;
          prod_s += nwarps_total;
          if (prod_s < st.ns) prod_src = st.Lh + (size_t)(prod_s * st.g_ny + j_out) * st.elem_stride;
        }
        prod_left = mo + 8 * prod_p8 + 8;
      }
    }
  };
  produce();

  for (int s_idx = s_first; s_idx < st.ns; s_idx += nwarps_total) {
    const int b = s_idx * st.g_ny + j_out;
    __syncwarp();  // wv is about to be rewritten: every lane is done with the previous element
    double xs[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xs[a] = x[(size_t)b * D + a];

    // ---- A: kernel vector (and the element's beta) ----------------------------------------------------------
    for (int i = lane; i < m; i += 32) {
      double kv[T];
      kernel_row<D, T>(sXo + i * D, sTo[i], xs, il, os, kv);
#pragma unroll
      for (int r = 0; r < T; ++r) wv[i * TP + r] = kv[r];
    }
    const double* bh = st.beta_h + (size_t)b * st.c_cap;
    for (int k = lane; k < c; k += 32) {
      double kv[T];
      const double* xa = st.Xh + ((size_t)b * st.cap_points + st.hobs_pt[k]) * D;
      const double be = bh[k];
      kernel_row<D, T>(xa, st.hobs_task[k], xs, il, os, kv);
#pragma unroll
      for (int r = 0; r < T; ++r) wv[(mo + k) * TP + r] = kv[r];
      wb[mo + k] = be;
    }
    __syncwarp();

    // ---- B: shared rows against L_oo --------------------------------------------------------------------------
    for (int p8 = 0; p8 < Pm; ++p8) {
      const int n_off = 8 * p8;
      const double* base = Lp + subpanel_off(p8, 0);
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      mma_accumulate<TP>(acc, base + tig * 8 + gid, wv + tig * TP + gid, n_off >> 2);
      subpanel_finish<T, TP>(acc, base + n_off * 8, wv + n_off * TP, min(8, m - n_off), lane);
    }

    // ---- C: own rows, streamed through the TMA ring -------------------------------------------------------------
    for (int p8 = 0; p8 < P8; ++p8) {
      const int n_off = mo + 8 * p8;
      const int nch = (n_off + 8 + STEP_SEG - 1) / STEP_SEG;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* slot_base = ring;
      for (int ch = 0; ch < nch; ++ch) {
        const unsigned slot = consumed & (STEP_NST - 1);
        mbar_wait(bars + slot, (consumed / STEP_NST) & 1);
        slot_base = ring + (size_t)slot * STEP_SEG * 8;
        const int t0 = ch * STEP_SEG;
        const int t1 = min(t0 + STEP_SEG, n_off);
        if (t1 > t0) mma_accumulate<TP>(acc, slot_base + tig * 8 + gid, wv + (t0 + tig) * TP + gid, (t1 - t0) >> 2);
        if (ch < nch - 1) {  // fully consumed (the diagonal block lives in the last chunk): refill the slot
          __syncwarp();
          ++consumed;
          produce();
        }
      }
      subpanel_finish<T, TP>(acc, slot_base + (size_t)(n_off - (nch - 1) * STEP_SEG) * 8, wv + n_off * TP,
                             min(8, c - 8 * p8), lane);
      ++consumed;
      produce();
    }

    // ---- D: posterior moments: C[r][s] = sum_t w[t][r] w[t][s],  C[r][7] = sum_t w[t][r] beta[t] ------------------
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* wp = wv + tig * TP + gid;
      const double* bp = wb + tig;
      const bool is_beta = gid == 7;
#pragma unroll 2
      for (int t = 0; t < wv_rows; t += 8) {
        const double a0 = wp[t * TP], a1 = wp[(t + 4) * TP];
        const double e0 = bp[t], e1 = bp[t + 4];
        dmma(acc[0], acc[1], a0, is_beta ? e0 : a0);
        dmma(acc[2], acc[3], a1, is_beta ? e1 : a1);
      }
      *reinterpret_cast<double2*>(sc + gid * 8 + 2 * tig) = make_double2(acc[0] + acc[2], acc[1] + acc[3]);
      __syncwarp();
    }
    TriT<T> S;
    double macc[T];
#pragma unroll
    for (int r = 0; r < T; ++r) {
      macc[r] = sc[r * 8 + 7];
#pragma unroll
      for (int s = 0; s <= r; ++s) {
        double kss = 0.0;
        if (r == s) kss = (r == 0) ? os : os * (il[r > 0 ? r - 1 : 0] * il[r > 0 ? r - 1 : 0]);
        S.at(r, s) = kss - sc[r * 8 + s];
      }
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
    if (!eps) continue;

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
    if (!grow_factor) continue;
    TriT<T> Sn = S, Ln;
#pragma unroll
    for (int r = 0; r < T; ++r) Sn.at(r, r) += st.noise[j_out * T + r];
    if (!chol_T<T>(Sn, 0.0, Ln)) {
      if (lane == 0) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
    }
    // new rows c .. c+T-1: entry of storage column t is w[t][r]; row c+r lives in sub-panel (c+r)/8
    double* Le = st.Lh + (size_t)b * st.elem_stride;
    double* rowp[T];
#pragma unroll
    for (int r = 0; r < T; ++r) rowp[r] = Le + subpanel_off((c + r) >> 3, mo) + ((c + r) & 7);
    for (int t = lane; t < mo + c; t += 32) {
      if (t >= m && t < mo) continue;  // padding columns stay 0
      double w[T];
      load_w<T>(wv + t * TP, w);
#pragma unroll
      for (int r = 0; r < T; ++r) rowp[r][(size_t)t * 8] = w[r];
    }
    if (lane == 0) {
      double bn[T];
#pragma unroll
      for (int r = 0; r < T; ++r) {
        double t = yv[r] - macc[r];
#pragma unroll
        for (int s = 0; s < r; ++s) {
          t -= Ln.at(r, s) * bn[s];
          rowp[r][(size_t)(mo + c + s) * 8] = Ln.at(r, s);
        }
        bn[r] = t / Ln.at(r, r);
        rowp[r][(size_t)(mo + c + r) * 8] = 1.0 / Ln.at(r, r);
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
    __syncwarp();
    warp_update_dinv(st, b, c, c + T, lane);
  }
}
