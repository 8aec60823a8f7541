// K1: fused rollout step (H = 1).  One WARP per batch element (sample s, output j); the warps of a CTA
// share output j, so the shared real-data factor L_oo (packed column-major, diagonal stored as 1/L_jj),
// the observed real inputs and beta_o are staged once per CTA in shared memory.
//
// The whole step is a COLUMN SWEEP of the bordered factor with lanes owning rows:
//   lane l holds, in registers, the T right-hand sides of rows l, l+32, ... (RSR slots for the m shared
//   rows, RSO slots for the element's own c rows).  Processing column j means: the owner lane finalises
//   w_j = k_j / L_jj and broadcasts it (T warp shuffles); every lane updates its rows i > j with
//   k_i -= L[i][j] w_j.  Reads of one column are contiguous over rows (coalesced), there is no cross-lane
//   reduction and no shared-memory scratch, and every load address is known up front, so the loads of
//   later columns are issued while the shuffle/FMA chain of earlier columns is still running.
//
//   A  kernel vector k(x*, X): each lane evaluates the entries of the rows it owns (one exp per row)
//   B  shared columns j < m against L_oo (shared memory), final w_j also parked in shared memory
//   V  the element's own rows against the shared columns   } the element's factor columns stream from HBM
//   T  the element's own triangular block (c shuffles)      } through a per-warp ring of TMA bulk copies
//      (cp.async.bulk + mbarrier complete_tx, STEP_P stages of STEP_G columns): one elected lane issues
//      one copy per column, STEP_P-1 stages ahead of the stage being consumed
//   D  Sigma* = K** - sum_i w_i w_i^T and mean = sum_i w_i beta_i from registers + warp-shuffle all-reduce
//   E  T x T Cholesky with GPyTorch's jitter ladder, y = mean + L eps, zero-variance / truncation
//   F  rank-T append: w_i goes to column i of the T new rows (T contiguous doubles per column)
//
// HBM traffic per element-step = its own factor entries read once + T new rows written once + O(T) I/O:
// HBM-bound by design (DESIGN.md "Roofline"); the host counts the algorithmic bytes per launch.
#pragma once
#include "gpmpc_state.cuh"

#define STEP_WARPS 4
#define FULL_MASK 0xffffffffu
#define STEP_G 2  // factor columns per TMA stage
#define STEP_P 4  // stages in the per-warp ring (STEP_P - 1 stages in flight while one is consumed)

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

template <int D, int T, int RSR, int RSO>
__global__ void __launch_bounds__(STEP_WARPS * 32)
k_step(DevState st, const double* __restrict__ x, const double* __restrict__ eps, gpmpc_sample_opts opts,
       double* __restrict__ mean, double* __restrict__ var, double* __restrict__ y,
       int* __restrict__ jitter_level, int grow_factor, int loo_in_smem, int cs) {
  extern __shared__ __align__(16) double smem[];
  const int j_out = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s_idx = blockIdx.x * STEP_WARPS + warp;
  const int m = st.m, c = st.c;
  const size_t tri = (size_t)m * (m + 1) / 2;
  const size_t tri_pad = (tri + 1) & ~(size_t)1;
  const int m_pad = (m + 1) & ~1;

  // ---- CTA-shared tables ---------------------------------------------------------------------------
  double* sLT = smem;                                  // [tri_pad] (only if loo_in_smem); diagonal = 1/L_jj
  double* sXo = sLT + (loo_in_smem ? tri_pad : 0);     // [m_pad*D] input of observed real scalar i
  double* sBo = sXo + (size_t)m_pad * D;               // [m_pad]
  int* sTo = (int*)(sBo + m_pad);                      // [2*m_pad] ints: task of observed real scalar i
  double* wo_base = (double*)(sTo + 2 * m_pad);        // per warp: final w of the shared rows, [T][m_pad]
  double* ring_base = wo_base + (size_t)STEP_WARPS * T * m_pad;  // per warp: [STEP_P][STEP_G][cs]
  uint64_t* bar_base = (uint64_t*)(ring_base + (size_t)STEP_WARPS * STEP_P * STEP_G * cs);  // [warps][STEP_P]
  if (threadIdx.x < STEP_WARPS * STEP_P) mbar_init(bar_base + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const double* gLT = st.LooT + (size_t)j_out * tri;
  if (loo_in_smem)
    for (size_t i = threadIdx.x; i < tri; i += blockDim.x) sLT[i] = gLT[i];
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double* xp = st.Xr + (size_t)st.obs_pt[i] * D;
#pragma unroll
    for (int a = 0; a < D; ++a) sXo[i * D + a] = xp[a];
    sBo[i] = st.beta_o[(size_t)j_out * m + i];
    sTo[i] = st.obs_task[i];
  }
  __syncthreads();
  if (s_idx >= st.ns) return;  // no block-level sync below this line
  const double* LT = loo_in_smem ? sLT : gLT;
  const int b = s_idx * st.g_ny + j_out;
  double* wo = wo_base + (size_t)warp * T * m_pad;
  double* ring = ring_base + (size_t)warp * STEP_P * STEP_G * cs;
  uint64_t* bars = bar_base + warp * STEP_P;
  const size_t ldC = st.ldC;
  double* LhTb = st.LhT + (size_t)b * (m + st.c_cap) * ldC;

  double il[D], xs[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    il[a] = 1.0 / st.ls[j_out * D + a];
    xs[a] = x[(size_t)b * D + a];
  }
  const double os = st.os[j_out];

  // ---- A: kernel vector, each lane for the rows it owns ------------------------------------------------
  double kr[RSR][T], ko[RSO][T], rdl[RSO];
#pragma unroll
  for (int rs = 0; rs < RSR; ++rs) {
    const int i = lane + 32 * rs;
    if (i < m) {
      kernel_row<D, T>(sXo + i * D, sTo[i], xs, il, os, kr[rs]);
    } else {
#pragma unroll
      for (int r = 0; r < T; ++r) kr[rs][r] = 0.0;
    }
  }
  const double* rdg = st.rdiag + (size_t)b * st.c_cap;
#pragma unroll
  for (int rs = 0; rs < RSO; ++rs) {
    const int i = lane + 32 * rs;
    if (i < c) {
      const double* xa = st.Xh + ((size_t)b * st.cap_points + st.hobs_pt[i]) * D;
      kernel_row<D, T>(xa, st.hobs_task[i], xs, il, os, ko[rs]);
      rdl[rs] = rdg[i];
    } else {
#pragma unroll
      for (int r = 0; r < T; ++r) ko[rs][r] = 0.0;
      rdl[rs] = 0.0;
    }
  }

  // ---- B: shared columns against L_oo ------------------------------------------------------------------
#pragma unroll
  for (int jb = 0; jb < RSR; ++jb) {
    const int jend = min(32, m - 32 * jb);
    for (int jl = 0; jl < jend; ++jl) {
      const int j = 32 * jb + jl;
      const double* col = LT + packed_col(j, m);
      const double rd = col[0];
      double wj[T];
#pragma unroll
      for (int r = 0; r < T; ++r) wj[r] = __shfl_sync(FULL_MASK, kr[jb][r] * rd, jl);
      if (lane == jl) {
#pragma unroll
        for (int r = 0; r < T; ++r) {
          kr[jb][r] = wj[r];
          wo[r * m_pad + j] = wj[r];
        }
      }
#pragma unroll
      for (int rs = jb; rs < RSR; ++rs) {
        const int i = lane + 32 * rs;
        if (i > j && i < m) {
          const double l = col[i - j];
#pragma unroll
          for (int r = 0; r < T; ++r) kr[rs][r] = fma(-l, wj[r], kr[rs][r]);
        }
      }
    }
  }
  __syncwarp();

  // ---- V + T: the element's own factor columns, streamed through the TMA ring ------------------------------
  // group g = STEP_G consecutive columns; groups 0..nV-1 cover the shared columns 0..m-1 (all c own rows),
  // groups nV.. cover the triangular columns k = 0..c-2 (rows k+1..c-1).  Column copies start at an even row
  // and have even length so that source, destination and size are 16-byte aligned.
  const int c_even = (c + 1) & ~1;
  const int nV = c > 0 ? (m + STEP_G - 1) / STEP_G : 0;
  const int nT = c > 1 ? (c - 1 + STEP_G - 1) / STEP_G : 0;
  const int ngroups = nV + nT;
  auto issue_group = [&](int g) {  // lane 0 only
    const int sidx = g % STEP_P;
    double* dst0 = ring + (size_t)sidx * STEP_G * cs;
    uint32_t total = 0;
    int col0, ncol, r0[STEP_G];
    if (g < nV) {
      col0 = g * STEP_G;
      ncol = min(STEP_G, m - col0);
#pragma unroll
      for (int u = 0; u < STEP_G; ++u) r0[u] = 0;
    } else {
      const int k0 = (g - nV) * STEP_G;
      col0 = m + k0;
      ncol = min(STEP_G, c - 1 - k0);
#pragma unroll
      for (int u = 0; u < STEP_G; ++u) r0[u] = (k0 + u + 1) & ~1;
    }
#pragma unroll
    for (int u = 0; u < STEP_G; ++u)
      if (u < ncol) total += (uint32_t)(c_even - r0[u]) * 8u;
    mbar_expect_tx(bars + sidx, total);
#pragma unroll
    for (int u = 0; u < STEP_G; ++u)
      if (u < ncol)
        tma_bulk_g2s(dst0 + (size_t)u * cs + r0[u], LhTb + (size_t)(col0 + u) * ldC + r0[u],
                     (uint32_t)(c_even - r0[u]) * 8u, bars + sidx);
  };
  if (lane == 0)
    for (int g = 0; g < min(STEP_P - 1, ngroups); ++g) issue_group(g);

  for (int g = 0; g < nV; ++g) {
    // the stage freed by group g-1 is refilled before this group is consumed
    __syncwarp();
    if (lane == 0 && g + STEP_P - 1 < ngroups) issue_group(g + STEP_P - 1);
    mbar_wait(bars + g % STEP_P, (g / STEP_P) & 1);
    const double* stg = ring + (size_t)(g % STEP_P) * STEP_G * cs;
#pragma unroll
    for (int u = 0; u < STEP_G; ++u) {
      const int j = g * STEP_G + u;
      if (j < m) {
        double wj[T];
#pragma unroll
        for (int r = 0; r < T; ++r) wj[r] = wo[r * m_pad + j];
#pragma unroll
        for (int rs = 0; rs < RSO; ++rs) {
          const int i = lane + 32 * rs;
          if (i < c) {
            const double l = stg[u * cs + i];
#pragma unroll
            for (int r = 0; r < T; ++r) ko[rs][r] = fma(-l, wj[r], ko[rs][r]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int jb = 0; jb < RSO; ++jb) {
    for (int gi = 0; gi < 32 / STEP_G; ++gi) {
      const int k0 = 32 * jb + STEP_G * gi;
      if (k0 >= c) break;
      const int g = nV + k0 / STEP_G;
      const bool has_data = k0 < c - 1;
      if (has_data) {
        __syncwarp();
        if (lane == 0 && g + STEP_P - 1 < ngroups) issue_group(g + STEP_P - 1);
        mbar_wait(bars + g % STEP_P, (g / STEP_P) & 1);
      }
      const double* stg = ring + (size_t)(g % STEP_P) * STEP_G * cs;
#pragma unroll
      for (int u = 0; u < STEP_G; ++u) {
        const int k = k0 + u;
        if (k < c) {
          const int jl = STEP_G * gi + u;
          double wj[T];
#pragma unroll
          for (int r = 0; r < T; ++r) wj[r] = __shfl_sync(FULL_MASK, ko[jb][r] * rdl[jb], jl);
          if (lane == jl) {
#pragma unroll
            for (int r = 0; r < T; ++r) ko[jb][r] = wj[r];
          }
#pragma unroll
          for (int rs = jb; rs < RSO; ++rs) {
            const int i = lane + 32 * rs;
            if (i > k && i < c) {
              const double l = stg[u * cs + i];
#pragma unroll
              for (int r = 0; r < T; ++r) ko[rs][r] = fma(-l, wj[r], ko[rs][r]);
            }
          }
        }
      }
    }
  }

  // ---- D: posterior moments --------------------------------------------------------------------------------
  TriT<T> Sacc;
  double macc[T];
#pragma unroll
  for (int i = 0; i < T * (T + 1) / 2; ++i) Sacc.v[i] = 0.0;
#pragma unroll
  for (int r = 0; r < T; ++r) macc[r] = 0.0;
  const double* bh = st.beta_h + (size_t)b * st.c_cap;
#pragma unroll
  for (int rs = 0; rs < RSR; ++rs) {
    const int i = lane + 32 * rs;
    const double be = i < m ? sBo[i] : 0.0;
#pragma unroll
    for (int r = 0; r < T; ++r) {
      macc[r] = fma(kr[rs][r], be, macc[r]);
#pragma unroll
      for (int s = 0; s <= r; ++s) Sacc.at(r, s) = fma(kr[rs][r], kr[rs][s], Sacc.at(r, s));
    }
  }
#pragma unroll
  for (int rs = 0; rs < RSO; ++rs) {
    const int i = lane + 32 * rs;
    const double be = i < c ? bh[i] : 0.0;
#pragma unroll
    for (int r = 0; r < T; ++r) {
      macc[r] = fma(ko[rs][r], be, macc[r]);
#pragma unroll
      for (int s = 0; s <= r; ++s) Sacc.at(r, s) = fma(ko[rs][r], ko[rs][s], Sacc.at(r, s));
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
  for (int r = 0; r < T; ++r) Sn.at(r, r) += st.noise[j_out * T + r];
  if (!chol_T<T>(Sn, 0.0, Ln)) {
    if (lane == 0) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
  }
  // new rows c .. c+T-1: entry of column i is w_i (T contiguous doubles per column)
#pragma unroll
  for (int rs = 0; rs < RSR; ++rs) {
    const int i = lane + 32 * rs;
    if (i < m) {
      double* dst = LhTb + (size_t)i * ldC + c;
#pragma unroll
      for (int r = 0; r < T; ++r) dst[r] = kr[rs][r];
    }
  }
#pragma unroll
  for (int rs = 0; rs < RSO; ++rs) {
    const int i = lane + 32 * rs;
    if (i < c) {
      double* dst = LhTb + (size_t)(m + i) * ldC + c;
#pragma unroll
      for (int r = 0; r < T; ++r) dst[r] = ko[rs][r];
    }
  }
  if (lane == 0) {
    double bn[T];
#pragma unroll
    for (int r = 0; r < T; ++r) {
      double t = yv[r] - macc[r];
#pragma unroll
      for (int s = 0; s < r; ++s) {
        t -= Ln.at(r, s) * bn[s];
        LhTb[(size_t)(m + c + s) * ldC + c + r] = Ln.at(r, s);
      }
      bn[r] = t / Ln.at(r, r);
      st.rdiag[(size_t)b * st.c_cap + c + r] = 1.0 / Ln.at(r, r);
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
