// K1h: the WHOLE conditioned rollout in one launch (gpmpc_rollout): a warp owns one batch element (sample s, output j)
// for the entire horizon, the g_ny warps of a sample sit in the same CTA and exchange the sampled values through shared
// memory (one named barrier per step), the next state is computed in-kernel.  Replaces, per horizon step, the launches
// k_step + k_step_finish (+ the eigen-redo launch) + k_rollout_state of the step-wise path (gpmpc_step.cuh), and with them
//   * 150-200 launches per rollout and the st.fin round trip between k_step and k_step_finish,
//   * every global read of phase A (the element's hallucinated inputs and beta stay in the warp's shared memory),
//   * and -- the point -- the HBM traffic of the factor: between two steps an element's factor is touched by nobody else,
//     and the persistent grid holds only  #SMs x warps  elements at a time (148 x 12 = 1776: ~100 MB of factor on average
//     when the groups are staggered over the horizon), so the stream the TMA ring pulls in at step t+1 is what the warp
//     wrote / read at step t and still sits in the 126 MB L2.  The step-wise path re-reads all 58 GB of factor state from
//     HBM at every step (1.1 TB per rollout); here HBM sees each factor row once, when it is evicted.
// The arithmetic is the step-wise path's, phase by phase and in the same order (the device functions of gpmpc_step.cuh;
// rollout_next_state / rollout_inputs below are shared with k_rollout_state), and the global state it leaves behind
// (factor rows in sub-panel layout, beta_h, Xh / Yh, row tables) is identical, so trajectories are BIT-IDENTICAL to the
// step-wise rollout (tests/test_gpu_horizon.py) and posterior / step calls can continue on the handle afterwards.
//
// A failed jitter ladder cannot take GPyTorch's batch-wide eigen-root fallback here (other elements are already steps
// ahead): the element's draw is NaN, GPMPC_ST_SAMPLE_NOT_PD is raised without GPMPC_ST_SAMPLE_EIG and the host re-runs the
// rollout on the step-wise path, which has the in-stream redo (ForwardRollout.check).
#pragma once
#include "gpmpc_assemble.cuh"
#include "gpmpc_step.cuh"

#ifndef HZ_MAX_WARPS
#define HZ_MAX_WARPS 12  // 384 threads: up to 168 registers per thread
#endif
// "rollout_fused" = 2 (automatic): samples one warp group may take one after the other (measured: ahead at 4, behind at 34)
#define HZ_AUTO_MAX_SERIAL 4
#ifndef HZ_SEG
#define HZ_SEG 48        // 8-row column groups per TMA chunk / ring slot (3 KB); multiple of 8
#endif
#ifndef HZ_NST
#define HZ_NST 2         // ring slots per warp
#endif
#define HZ_SLOT_BYTES (HZ_SEG * 64)

struct HorizonArgs {
  const double* x0;    // [ns][nx]
  const double* u_ff;  // [n_steps][nu]
  const double* eps;   // [n_steps][B*T]
  double* traj;        // [ns][nx][n_steps+1]
  int n_steps;
  int s_begin, s_end;  // samples [s_begin, s_end) of the handle are rolled out by this launch
  int groups;          // sample groups (of g_ny warps) per CTA
  unsigned long long stagger_ns;  // the groups' starts are spread uniformly over this many ns (0: all start together)
  gpmpc_sample_opts opts;
};

// shared-memory layout in doubles, computed identically by the launcher and the kernel
struct HzLayout {
  int loop_sz, m_even, off_sXr, off_sBo, off_sRrow, off_env, off_u, off_sY, off_warp, wv_sz, wb_sz, xh_sz, per_warp, off_bars, total;
};
#define HZ_ENV_DOUBLES ((int)((sizeof(gpmpc_env) + 7) / 8))
__host__ __device__ inline HzLayout hz_layout(int g_ny, int n_real, int m, int mo, int D, int T, int n_steps, int groups) {
  HzLayout L;
  L.loop_sz = (int)subpanel_off((m + 7) >> 3, 0);
  L.m_even = (m + 1) & ~1;
  const int nr_even = (n_real + 1) & ~1;
  L.off_sXr = g_ny * L.loop_sz;
  L.off_sBo = L.off_sXr + nr_even * D;
  L.off_sRrow = L.off_sBo + g_ny * L.m_even;
  L.off_env = L.off_sRrow + ((n_real * T + 1) & ~1) / 2;
  L.off_u = L.off_env + HZ_ENV_DOUBLES;
  L.off_sY = L.off_u + n_steps * GPMPC_MAX_NX;
  L.off_warp = (L.off_sY + groups * 2 * g_ny * T + 15) & ~15;
  const int wv_rows = mo + 8 * ((T * n_steps + 7) >> 3);
  L.wv_sz = (wv_rows * T + 8 + 15) & ~15;
  L.wb_sz = (wv_rows + 15) & ~15;
  L.xh_sz = (n_steps * D + 15) & ~15;
  L.per_warp = L.wv_sz + L.wb_sz + L.xh_sz + 64 + 16 + HZ_NST * HZ_SEG * 8;
  L.off_bars = L.off_warp + groups * g_ny * L.per_warp;
  L.total = L.off_bars + groups * g_ny * HZ_NST;
  return L;
}

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int D, int T>
__global__ void __launch_bounds__(HZ_MAX_WARPS * 32, 1)
k_horizon(DevState st, const __grid_constant__ gpmpc_env env, HorizonArgs a) {
  extern __shared__ __align__(128) double smem[];
  const int g_ny = st.g_ny;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int m = st.m, mo = st.mo;
  const int Pm = (m + 7) >> 3;
  const int nx = env.nx, nu = env.nu;
  const HzLayout L = hz_layout(g_ny, st.n_real, m, mo, D, T, a.n_steps, a.groups);
  const int loop_sz = L.loop_sz, m_even = L.m_even;

  double* sL_all = smem;                          // [g_ny][loop_sz]  inv(L_oo), sub-panel layout
  double* sXr = smem + L.off_sXr;                 // [n_real][D]
  double* sBo_all = smem + L.off_sBo;             // [g_ny][m_even]   beta_o
  int* sRrow = (int*)(smem + L.off_sRrow);        // [n_real * T]     factor row of (real point, task), -1 = unobserved
  double* sY_all = smem + L.off_sY;               // [groups][2][g_ny][T]
  gpmpc_env* envs_w = (gpmpc_env*)(smem + L.off_env);  // the env hooks in shared memory: dynamically indexed tables without
  const gpmpc_env& envs = *envs_w;                     // constant-bank round trips
  double* sU = smem + L.off_u;                    // [n_steps][nu]
  const int nw_used = a.groups * g_ny;
  for (int idx = threadIdx.x; idx < (int)(sizeof(gpmpc_env) / 4); idx += blockDim.x)
    ((int*)envs_w)[idx] = ((const int*)&env)[idx];
  for (int idx = threadIdx.x; idx < a.n_steps * nu; idx += blockDim.x) sU[idx] = a.u_ff[idx];

  for (int idx = threadIdx.x; idx < g_ny * loop_sz; idx += blockDim.x) sL_all[idx] = st.LooP[idx];
  for (int idx = threadIdx.x; idx < st.n_real * D; idx += blockDim.x) sXr[idx] = st.Xr[idx];
  for (int idx = threadIdx.x; idx < st.n_real * T; idx += blockDim.x) sRrow[idx] = -1;
  for (int idx = threadIdx.x; idx < g_ny * m; idx += blockDim.x) sBo_all[(idx / m) * m_even + idx % m] = st.beta_o[idx];
  __syncthreads();
  for (int idx = threadIdx.x; idx < m; idx += blockDim.x) sRrow[st.obs_pt[idx] * T + st.obs_task[idx]] = idx;
  __syncthreads();
  if (warp >= nw_used) return;  // no block-level synchronisation below this line (named barriers per group only)

  const int grp = warp / g_ny, j_out = warp - grp * g_ny;
  double* wv = smem + L.off_warp + (size_t)warp * L.per_warp;  // [wv_rows][T]  k, then w
  double* wb = wv + L.wv_sz;                                    // [wv_rows]     beta by storage column
  double* sXh = wb + L.wb_sz;                                   // [n_steps][D]  this element's hallucinated inputs
  double* db = sXh + L.xh_sz;                                   // [8][8]  the partially filled diagonal block, db[col t * 8 + row i]
  double* zs = db + 64;                                         // [nx + nu]  [x, u] of the current step
  double* ring = zs + 16;                                     // [HZ_NST][HZ_SEG * 8]
  uint64_t* bars = (uint64_t*)(smem + L.off_bars) + warp * HZ_NST;
  const double* sL = sL_all + (size_t)j_out * loop_sz;
  const double* sBo = sBo_all + (size_t)j_out * m_even;
  double* sY = sY_all + (size_t)grp * 2 * g_ny * T;

  if (lane < HZ_NST) mbar_init(bars + lane, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const uint32_t wv_s = smem_u32(wv), wb_s = smem_u32(wb), ring_s = smem_u32(ring), bars_s = smem_u32(bars);
  const uint32_t sL_s = smem_u32(sL);
  const uint32_t a_lane = a_lane_off(gid, tig);
  const uint32_t b_lane = (tig * T + gid) * 8;
  double il[D];
#pragma unroll
  for (int q = 0; q < D; ++q) il[q] = 1.0 / st.ls[j_out * D + q];
  const double os = st.os[j_out];
  double noise[T], kdiag[T];
#pragma unroll
  for (int r = 0; r < T; ++r) {
    noise[r] = st.noise[j_out * T + r];
    kdiag[r] = r == 0 ? os : __dmul_rn(os, __dmul_rn(il[r - 1], il[r - 1]));
  }
  // (k_step_finish computes the prior variance of task r as os * ((1/l) * (1/l)) with 1/l = 1.0 / ls: the same expression)

  // ---- TMA producer: one factor stream per horizon step (lane 0 issues; state is warp-uniform) --------------------------
  const char* prod_base = nullptr;
  unsigned prod_bytes = 0, prod_off = 0, prod_slot = 0;
  auto produce_one = [&]() {
    if (prod_off >= prod_bytes) return;
    const unsigned bytes = min((unsigned)HZ_SLOT_BYTES, prod_bytes - prod_off);
    const uint32_t bar = bars_s + prod_slot * 8;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.u32 p, %4, 0;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];\n\t}"
        ::"r"(bar), "r"(bytes), "r"(ring_s + prod_slot * HZ_SLOT_BYTES), "l"(prod_base + prod_off), "r"(lane) : "memory");
    prod_slot = prod_slot + 1 == HZ_NST ? 0 : prod_slot + 1;
    prod_off += bytes;
  };
  unsigned cons_slot = 0, cons_parity = 0;

  const int groups_total = gridDim.x * a.groups;
  const int g_index = blockIdx.x * a.groups + grp;
  if (a.stagger_ns) {
    // spread the groups' positions in the horizon (and with them the factor bytes resident in L2) uniformly
    const unsigned long long t0 = global_timer_ns();
    const unsigned h = (unsigned)g_index * 2654435761u;  // co-resident groups get unrelated phases
    const unsigned long long delay = (unsigned long long)((double)(h >> 8) / 16777216.0 * (double)a.stagger_ns);
    while (global_timer_ns() - t0 < delay) __nanosleep(2000);
  }

  const int barrier_id = 1 + grp, barrier_threads = 32 * g_ny;
  const size_t epsB = (size_t)st.B * T;

  for (int s_idx = a.s_begin + g_index; s_idx < a.s_end; s_idx += groups_total) {
    const int b = s_idx * g_ny + j_out;
    double* Le = st.Lh + (size_t)b * st.elem_stride;
    double* Xb = st.Xh + (size_t)b * st.cap_points * D;
    double* Yb = st.Yh + (size_t)b * st.cap_points * T;
    double* bh = st.beta_h + (size_t)b * st.c_cap;
    // fresh element: w / beta arrays zero (rows >= c and the padding rows [m, mo) must read as 0 throughout), beta_o in place
    __syncwarp();
    for (int idx = lane; idx < L.wv_sz + L.wb_sz; idx += 32) wv[idx] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < m; idx += 32) wb[idx] = sBo[idx];
    // [x, u] of step 0: one row per lane (x: lanes < nx, then u: lanes < nu), through the warp's zs array
    if (lane < nx) {
      const double x0i = a.x0[(size_t)s_idx * nx + lane];
      zs[lane] = x0i;
      if (j_out == 0) a.traj[((size_t)s_idx * nx + lane) * (a.n_steps + 1)] = x0i;
    }
    __syncwarp();
    if (lane < nu) zs[nx + lane] = rollout_input_row(envs, zs, sU, lane);
    prod_bytes = prod_off = 0;
    __syncwarp();

    for (int t = 0; t < a.n_steps; ++t) {
      const int c = T * t, np = t;
      const int P8 = (c + 7) >> 3;
      const int wv_rows = mo + 8 * P8;
      double xs[D];
#pragma unroll
      for (int q = 0; q < D; ++q) xs[q] = zs[envs.g_idx_inputs[q]];
      double epsv[T];
#pragma unroll
      for (int r = 0; r < T; ++r) epsv[r] = a.eps[(size_t)t * epsB + (size_t)b * T + r];

      // ---- A: kernel vector (hallucinated inputs from shared memory; their beta is already in wb) -----------------------
      for (int p0 = 0; p0 < np; p0 += 32) {
        const int p = p0 + lane;
        if (p < np) {
          double xa[D], kb[T][T];
#pragma unroll
          for (int q = 0; q < D; ++q) xa[q] = sXh[p * D + q];
          kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
          for (int ta = 0; ta < T; ++ta)
#pragma unroll
            for (int tb = 0; tb < T; ++tb) wv[(mo + p * T + ta) * T + tb] = kb[ta][tb];
        }
      }
      for (int p = lane; p < st.n_real; p += 32) {
        double xa[D], kb[T][T];
#pragma unroll
        for (int q = 0; q < D; ++q) xa[q] = sXr[p * D + q];
        kernel_block<D, T>(xa, xs, il, os, kb);
#pragma unroll
        for (int ta = 0; ta < T; ++ta) {
          const int row = sRrow[p * T + ta];
          if (row >= 0) {
#pragma unroll
            for (int tb = 0; tb < T; ++tb) wv[row * T + tb] = kb[ta][tb];
          }
        }
      }
      __syncwarp();

      // ---- B: shared rows w_o = inv(L_oo) k_o, tile-rows last to first (in place) ----------------------------------------
      for (int p8 = Pm - 1; p8 >= 0; --p8) {
        const uint32_t boff = (uint32_t)subpanel_off(p8, 0) * 8;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        mma_accumulate<T>(acc, sL_s + boff + a_lane, wv_s + b_lane, 2 * p8 + 2);
        const uint32_t mine = wv_s + ((8 * p8 + gid) * T + 2 * tig) * 8;
        if (2 * tig < T) sts(mine, acc[0] + acc[2]);
        if (2 * tig + 1 < T) sts(mine + 8, acc[1] + acc[3]);
      }
      __syncwarp();

      // ---- C: own rows, streamed through the TMA ring (from L2: written / read by this warp one step ago) -----------------
      //      (a per-pair consumer loop with a power-of-two ring was tried and lost: 305 ms against 275 ms per rollout --
      //      the 4-fold unrolled mma_accumulate over long pieces issues fewer instructions per k-block)
      if (P8 > 0) {
        const unsigned elem_bytes = (unsigned)step_groups_per_element(c, mo) * 64u;
        mbar_wait_s(bars_s + cons_slot * 8, cons_parity);
        unsigned left_in_elem = elem_bytes / 64;
        int cpos = 0;
        auto next_chunk = [&]() {
          __syncwarp();
          produce_one();
          cpos = 0;
          if (++cons_slot == HZ_NST) { cons_slot = 0; cons_parity ^= 1; }
          if (left_in_elem > 0) mbar_wait_s(bars_s + cons_slot * 8, cons_parity);
        };
        for (int p8 = 0; p8 < P8; ++p8) {
          const int n_off = mo + 8 * p8;
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          uint32_t wa = wv_s + b_lane;
          int rem = n_off;
          while (rem > 0) {
            const int piece = min(rem, HZ_SEG - cpos);
            mma_accumulate<T>(acc, ring_s + cons_slot * HZ_SLOT_BYTES + cpos * 64 + a_lane, wa, piece >> 2);
            wa += piece * T * 8;
            rem -= piece;
            cpos += piece;
            left_in_elem -= piece;
            if (cpos == HZ_SEG) next_chunk();
          }
          const uint32_t dblk = ring_s + cons_slot * HZ_SLOT_BYTES + cpos * 64;
          subpanel_finish<T, true>(acc, dblk, nullptr, wv_s + n_off * T * 8, min(8, c - 8 * p8), gid, tig);
          cpos += 8;
          left_in_elem -= 8;
          if (cpos == HZ_SEG || left_in_elem == 0) next_chunk();
        }
      }

      // ---- D: posterior moments: C[r][s] = sum_t w[t][r] w[t][s],  C[r][7] = sum_t w[t][r] beta[t] ------------------------
      double c0, c1;
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t wa = wv_s + b_lane, ba = wb_s + tig * 8;
        const bool is_beta = gid == 7;
        for (int tt = 0; tt < wv_rows; tt += 8) {
          const double a0 = lds<0>(wa), a1 = lds<4 * T * 8>(wa);
          const double e0 = lds<0>(ba), e1 = lds<32>(ba);
          dmma(acc[0], acc[1], a0, is_beta ? e0 : a0);
          dmma(acc[2], acc[3], a1, is_beta ? e1 : a1);
          wa += 8 * T * 8;
          ba += 64;
        }
        c0 = acc[0] + acc[2];
        c1 = acc[1] + acc[3];
      }
      // lane (gid, tig) holds C[gid][2 tig] (c0), C[gid][2 tig + 1] (c1): every lane gathers mean and lower(W^T W)
      double macc[T];
      TriT<T> S;
#pragma unroll
      for (int r = 0; r < T; ++r) {
        macc[r] = __shfl_sync(0xffffffffu, c1, 4 * r + 3);  // C[r][7]
#pragma unroll
        for (int s2 = 0; s2 <= r; ++s2) {
          const double v0 = __shfl_sync(0xffffffffu, c0, 4 * r + (s2 >> 1));
          const double v1 = __shfl_sync(0xffffffffu, c1, 4 * r + (s2 >> 1));
          const double wtw = (s2 & 1) ? v1 : v0;
          S.at(r, s2) = (r == s2 ? kdiag[r] : 0.0) - wtw;
        }
      }

      // ---- E: draw + post-processing (k_step_finish, same arithmetic; every lane computes the same values) ---------------
      double vr[T];
#pragma unroll
      for (int r = 0; r < T; ++r) vr[r] = fmax(S.at(r, r), GP_MIN_VARIANCE);
      TriT<T> Lc;
      int level = 0;
      if (T == 1) {
        Lc.v[0] = a.opts.unclamped_sqrt_1x1 ? sqrt(S.v[0]) : sqrt(fmax(S.v[0], 0.0));
      } else {
        bool ok = chol_T<T>(S, 0.0, Lc);
        double jit = st.jitter;
        while (!ok && level < GP_MAX_TRIES) {
          ++level;
          ok = chol_T<T>(S, jit, Lc);
          jit *= 10.0;
        }
        if (!ok) {
          level = 4;
          bool has_nan = false;
#pragma unroll
          for (int i = 0; i < T * (T + 1) / 2; ++i) has_nan = has_nan || isnan(S.v[i]);
          if (lane == 0) atomicOr(st.status, GPMPC_ST_SAMPLE_NOT_PD | (has_nan ? GPMPC_ST_NAN_INPUT : 0u));
        }
      }
      double yv[T];
#pragma unroll
      for (int r = 0; r < T; ++r) {
        double acc = macc[r];
#pragma unroll
        for (int s2 = 0; s2 <= r; ++s2) acc += Lc.at(r, s2) * epsv[s2];
        yv[r] = level < 4 ? acc : nan("");
      }
      bool zero = a.opts.variance_is_zero >= 0.0;
#pragma unroll
      for (int r = 0; r < T; ++r) zero = zero && (vr[r] <= a.opts.variance_is_zero);
#pragma unroll
      for (int r = 0; r < T; ++r) {
        if (zero) yv[r] = macc[r];
        if (a.opts.beta >= 0.0) {
          const double sd = sqrt(vr[r]);
          yv[r] = fmin(fmax(yv[r], macc[r] - a.opts.beta * sd), macc[r] + a.opts.beta * sd);
        }
      }
      // the sampled values to the sample's group (buffer t & 1), the record of the point to global memory
      if (lane < T) {
        double mine = yv[0];
#pragma unroll
        for (int r = 1; r < T; ++r) mine = lane == r ? yv[r] : mine;
        sY[((t & 1) * g_ny + j_out) * T + lane] = mine;
        Yb[(size_t)np * T + lane] = mine;
      }
      if (lane < D) {
        double mine = xs[0];
#pragma unroll
        for (int q = 1; q < D; ++q) mine = lane == q ? xs[q] : mine;
        Xb[(size_t)np * D + lane] = mine;
        sXh[np * D + lane] = mine;
      }

      // ---- F: condition on (x*, y): T new factor rows -----------------------------------------------------------------
      TriT<T> Sn = S, Ln;
#pragma unroll
      for (int r = 0; r < T; ++r) Sn.at(r, r) += noise[r];
      if (!chol_T<T>(Sn, 0.0, Ln) && lane == 0) atomicOr(st.status, GPMPC_ST_APPEND_NOT_PD);
      double bn[T], rdn[T];
#pragma unroll
      for (int r = 0; r < T; ++r) {
        double tmp = yv[r] - macc[r];
#pragma unroll
        for (int s2 = 0; s2 < r; ++s2) tmp -= Ln.at(r, s2) * bn[s2];
        bn[r] = tmp / Ln.at(r, r);
        rdn[r] = 1.0 / Ln.at(r, r);
      }
      if (lane < T) {
        double mine = bn[0];
#pragma unroll
        for (int r = 1; r < T; ++r) mine = lane == r ? bn[r] : mine;
        bh[c + lane] = mine;
        wb[mo + c + lane] = mine;
      }
      // entries left of the diagonal block are w itself (storage columns [0, m) and [mo, mo + c))
      {
        double* rowp[T];
#pragma unroll
        for (int r = 0; r < T; ++r) rowp[r] = Le + subpanel_off((c + r) >> 3, mo) + sp_idx(0, (c + r) & 7);
        for (int tt = lane; tt < mo + c; tt += 32) {
          if (tt >= m && tt < mo) continue;
          const size_t to = sp_idx(tt, 0);
#pragma unroll
          for (int r = 0; r < T; ++r) rowp[r][to] = wv[tt * T + r];
        }
      }
      // the diagonal-block part, row by row: db mirrors the (partially filled) 8 x 8 block the new rows fall into -- lower:
      // L, diagonal: 1/L_kk, upper slot (row jc, column i): inv(D)[i][jc] -- and persists in shared memory between steps
#pragma unroll
      for (int r = 0; r < T; ++r) {
        const int k = c + r, i = k & 7, kb = k - i;    // block row i of block kb
        const int i_old = max(0, c - kb);              // rows of this block that were there before this step
        double* gblk = Le + subpanel_off(kb >> 3, mo) + (size_t)(mo + kb) * 8;
        __syncwarp();
        if (lane < i) {
          // L[k][kb + lane]: under an old column of the block it is w (already in global memory by the loop above),
          // under a new column it is Ln[r][.]
          double e;
          if (lane < i_old) {
            e = wv[(mo + kb + lane) * T + r];
          } else {
            const int s2 = kb + lane - c;  // new row index of that column
            e = 0.0;
#pragma unroll
            for (int q = 0; q < T; ++q)
              if (q < r && q == s2) e = Ln.at(r, q);
            gblk[sp_idx(lane, i)] = e;
          }
          db[lane * 8 + i] = e;
        } else if (lane == i) {
          db[i * 8 + i] = rdn[r];
          gblk[sp_idx(i, i)] = rdn[r];
        }
        // new columns that belong to an EARLIER block (the T new rows straddle a block boundary)
        if (lane < r && c + lane < kb) {
          double e = 0.0;
#pragma unroll
          for (int q = 0; q < T; ++q)
            if (q < r && q == lane) e = Ln.at(r, q);
          Le[subpanel_off(kb >> 3, mo) + sp_idx(mo + c + lane, i)] = e;
        }
        __syncwarp();
        if (lane < i) {
          const int jc = lane;
          double acc = 0.0;
          for (int tt = jc; tt < i; ++tt) acc = fma(db[tt * 8 + i], db[tt * 8 + jc], acc);
          const double inv = -rdn[r] * acc;
          db[i * 8 + jc] = inv;  // slot (row jc, column i)
          gblk[sp_idx(i, jc)] = inv;
        }
      }
      __syncwarp();

      // ---- the factor stream of the next step: make this step's global writes visible to the async proxy, prefetch -------
      if (t + 1 < a.n_steps) {
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        prod_base = (const char*)Le;
        prod_bytes = (unsigned)step_groups_per_element(c + T, mo) * 64u;
        prod_off = 0;
#pragma unroll
        for (int i = 0; i < HZ_NST; ++i) produce_one();
      }

      // ---- next state: the g_ny outputs of the sample meet; every warp advances its own copy of [x, u], one row per lane ----
      named_barrier(barrier_id, barrier_threads);
      {
        const double* yb = sY + (size_t)(t & 1) * g_ny * T;
        double xn = 0.0;
        if (lane < nx) xn = rollout_next_state_row(envs, T, zs, [&](int j) { return yb + j * T; }, lane);
        __syncwarp();  // every lane has read the old [x, u]
        if (lane < nx) {
          zs[lane] = xn;
          if (j_out == 0) a.traj[((size_t)s_idx * nx + lane) * (a.n_steps + 1) + t + 1] = xn;
        }
        __syncwarp();
        if (t + 1 < a.n_steps && lane < nu) zs[nx + lane] = rollout_input_row(envs, zs, sU + (size_t)(t + 1) * nu, lane);
        __syncwarp();
      }
    }
  }
}

// row tables of a handle whose every element holds n_points fully observed hallucinated points (what n_points fused
// steps leave behind): hrow0[p] = p T, hobs_pt[k] = k / T, hobs_task[k] = k % T
__global__ void k_fill_row_tables(DevState st, int n_points) {
  const int T = st.T;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points * T; p += gridDim.x * blockDim.x) {
    if (p < n_points) st.hrow0[p] = p * T;
    st.hobs_pt[p] = p / T;
    st.hobs_task[p] = p % T;
  }
}
