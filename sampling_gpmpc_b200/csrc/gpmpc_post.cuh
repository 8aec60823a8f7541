// K4: the data-set rules around the draw (Dyn_gp_min_data_dist) and the formats either side of the hot path:
// the acados stage-parameter vector the SQP driver consumes (src/solver.py:98-131) and the per-stage
// reductions the trajectory consumers need (tightening Delta_k, bounding boxes, convex-hull candidates).
// All of it is HBM-bound index / byte work: one pass over the data, coalesced along the fastest axis.
#pragma once
#include "gpmpc_state.cuh"

// Euclidean distance exactly as torch.linalg.vector_norm(a - b, dim=-1) defines it (sqrt of the sum of squares)
__device__ __forceinline__ double point_dist(const double* __restrict__ a, const double* __restrict__ b, int d) {
  double s = 0.0;
  for (int k = 0; k < d; ++k) {
    const double r = a[k] - b[k];
    s = __dadd_rn(s, __dmul_rn(r, r));  // no FMA contraction: the comparison against min_dist must not depend on it
  }
  return sqrt(s);
}

// sample_gp's min-distance overwrite followed by the truncation (src/agent.py:666-708).  One warp per (b, h):
// lanes scan the model's training points (real, then recorded hallucinated); points with a NaN target count as
// infinitely far (agent.py:674-679); the nearest point's targets replace the draw when it is within min_dist
// (first index wins a tie, like torch.min on CPU); then y is clipped to mean +- beta sqrt(var).
__global__ void k_min_dist_overwrite(DevState st, const double* __restrict__ x, int H, const double* __restrict__ mean,
                                     const double* __restrict__ var, double min_dist, double beta,
                                     double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= (long long)st.B * H) return;
  const int b = (int)(pair / H), j = b % st.g_ny, d = st.d, T = st.T;
  const double* xs = x + pair * d;
  const int n = st.n_real + st.np;
  double best = INFINITY;
  int best_i = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    const double* xt;
    bool full;
    if (i < st.n_real) {
      xt = st.Xr + (size_t)i * d;
      full = st.real_full[(size_t)j * st.n_real + i] != 0;
    } else {
      const size_t p = (size_t)b * st.cap_points + (i - st.n_real);
      xt = st.Xh + p * d;
      full = true;
      for (int t = 0; t < T; ++t) full = full && !isnan(st.Yh[p * T + t]);
    }
    const double dist = full ? point_dist(xs, xt, d) : INFINITY;
    if (dist < best) { best = dist; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane < T) {
    double v = y[pair * T + lane];
    if (best <= min_dist) {
      v = best_i < st.n_real ? st.Yr[((size_t)j * st.n_real + best_i) * T + lane]
                             : st.Yh[((size_t)b * st.cap_points + (best_i - st.n_real)) * T + lane];
    }
    if (beta >= 0.0) {
      const double mu = mean[pair * T + lane], sd = sqrt(var[pair * T + lane]);
      const double hw = __dmul_rn(beta, sd);  // torch: mean -+ beta * sqrt(var) as separate roundings
      v = fmin(fmax(v, __dsub_rn(mu, hw)), __dadd_rn(mu, hw));
    }
    y[pair * T + lane] = v;
  }
}

// update_hallucinated_Dyn_dataset's filter (src/agent.py:166-181): a new point closer than min_dist to ANY input
// already in the element's data set (real or hallucinated, observed or not) gets NaN labels; counts[j][h] is the
// number of this handle's samples for which point h of output j was filtered (the host turns it into the
// reference's all-over-samples / any-over-batch flags, after an all-reduce when the samples are sharded).
__global__ void k_filter_new_points(DevState st, const double* __restrict__ x, int H, double min_dist,
                                    int use_hallucinated, double* __restrict__ y, int* __restrict__ counts,
                                    unsigned char* __restrict__ flags = nullptr) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= (long long)st.B * H) return;
  const int b = (int)(pair / H), h = (int)(pair % H), j = b % st.g_ny, d = st.d, T = st.T;
  const double* xs = x + pair * d;
  const int n = st.n_real + (use_hallucinated ? st.np : 0);
  bool hit = false;
  for (int i = lane; i < n && !hit; i += 32) {
    if (i >= st.n_real && st.pstate && st.pstate[(size_t)b * st.cap_points + (i - st.n_real)] == 2) continue;  // dropped: never stored
    const double* xt = i < st.n_real ? st.Xr + (size_t)i * d
                                     : st.Xh + ((size_t)b * st.cap_points + (i - st.n_real)) * d;
    hit = point_dist(xs, xt, d) <= min_dist;
  }
  hit = __any_sync(0xffffffffu, hit);
  if (flags && lane == 0) flags[pair] = hit ? 1 : 0;
  if (!hit) return;
  if (y && lane < T) y[pair * T + lane] = nan("");
  if (counts && lane == 0) atomicAdd(counts + (size_t)j * H + h, 1);
}

// Group-wise reduction of the flags of ONE new point per element (H = 1), groups = consecutive blocks of group_size samples
// = one reference Agent each (simulate_true_reachable_set.py: a new Agent of num_dyn_samples samples per repeat):
//   dropped  (2) iff for some output the point is filtered for ALL samples of the group      (src/agent.py:186-191)
//   masked   (1) iff filtered for ANY batch element of the group (GPyTorch's any-over-batch NaN mask)
//   appended (0) otherwise.                                              One warp per group.
__global__ void k_group_decide(int ns, int g_ny, int group_size, const unsigned char* __restrict__ flags,
                               unsigned char* __restrict__ decision) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_groups = (ns + group_size - 1) / group_size;
  if (g >= n_groups) return;
  const int s0 = g * group_size, s1 = min(ns, s0 + group_size);
  bool any = false, drop = false;
  for (int j = 0; j < g_ny; ++j) {
    bool all_j = true, any_j = false;
    for (int s = s0 + lane; s < s1; s += 32) {
      const bool f = flags[(size_t)s * g_ny + j] != 0;
      all_j = all_j && f;
      any_j = any_j || f;
    }
    all_j = __all_sync(0xffffffffu, all_j);
    any_j = __any_sync(0xffffffffu, any_j);
    drop = drop || all_j;
    any = any || any_j;
  }
  if (lane == 0) decision[g] = drop ? 2 : (any ? 1 : 0);
}

// The acados stage parameter p_lin (src/solver.py:98-131; consumed by src/utils/model.py:34-41), all stages at
// once: out[stage] = [ for every sample i: A_i (nx*nx row-major) | B_i (nx*nu) | x_lin_i (nx) | f_i (nx) ] ++ tail[stage]
// with A = y_grad (+ u_grad K with feedback, solver.py:90), B = u_grad, f = gp_val taken from the assembled
// linearisation lin [ns][nx][H][1+nx+nu], x_lin = x_h[stage][i*nx .. ].  One thread per output scalar.
__global__ void k_pack_plin(int ns, int nx, int nu, int H, int n_tail, int use_K, gpmpc_env env,
                            const double* __restrict__ lin, const double* __restrict__ x_h,
                            const double* __restrict__ tail, double* __restrict__ out) {
  const int per = nx * nx + nx * nu + 2 * nx, w = 1 + nx + nu;
  const long long P = (long long)ns * per + n_tail, total = P * H;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int stage = (int)(idx / P);
    const long long e = idx % P;
    double v;
    if (e >= (long long)ns * per) {
      v = tail[(size_t)stage * n_tail + (e - (long long)ns * per)];
    } else {
      const long long i = e / per;
      int r = (int)(e % per);
      const double* li = lin + ((size_t)i * nx * H + stage) * w;  // row a of sample i: + a*H*w
      if (r < nx * nx) {
        const int a = r / nx, c = r % nx;
        const double* row = li + (size_t)a * H * w;
        v = row[1 + c];
        if (use_K) {
          double acc = __dmul_rn(row[1 + nx], env.K_fb[c]);  // u_grad @ K, summed in k order, then added to y_grad
          for (int k = 1; k < nu; ++k) acc = __dadd_rn(acc, __dmul_rn(row[1 + nx + k], env.K_fb[k * nx + c]));
          v = __dadd_rn(v, acc);
        }
      } else if ((r -= nx * nx) < nx * nu) {
        v = li[(size_t)(r / nu) * H * w + 1 + nx + r % nu];
      } else if ((r -= nx * nu) < nx) {
        v = x_h[(size_t)stage * ns * nx + i * nx + r];
      } else {
        v = li[(size_t)(r - nx) * H * w];
      }
    }
    out[idx] = v;
  }
}

// ---- trajectory consumers ----------------------------------------------------------------------------------
// traj [ns][nx][H1] (H1 fastest).  Stage-wise reductions over the samples: bounding box and the tightening
// Delta[i][t] = max_n |x^n_t[i] - ref[i][t]|  (extra/approx_sampling_mpc/README.md:19-27).  Thread e owns column
// e = i*H1 + t (coalesced along e), blocks own sample ranges and write partials; k_traj_stats_final reduces them.
// max / min are exact, so the result does not depend on the split.
__global__ void k_traj_stats_partial(int ns, int cols, const double* __restrict__ traj, const double* __restrict__ ref,
                                     double* __restrict__ part /* [gridDim.y][3][cols] */) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cols) return;
  const int per = (ns + gridDim.y - 1) / gridDim.y;
  const int s0 = blockIdx.y * per, s1 = min(ns, s0 + per);
  const double r = ref ? ref[e] : 0.0;
  double lo = INFINITY, hi = -INFINITY, dev = 0.0;
  for (int s = s0; s < s1; ++s) {
    const double v = traj[(size_t)s * cols + e];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
    dev = fmax(dev, fabs(v - r));
  }
  double* p = part + (size_t)blockIdx.y * 3 * cols;
  p[e] = lo; p[cols + e] = hi; p[2 * cols + e] = dev;
}

__global__ void k_traj_stats_final(int nparts, int cols, const double* __restrict__ part, double* __restrict__ box_min,
                                   double* __restrict__ box_max, double* __restrict__ max_dev) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cols) return;
  double lo = INFINITY, hi = -INFINITY, dev = 0.0;
  for (int k = 0; k < nparts; ++k) {
    const double* p = part + (size_t)k * 3 * cols;
    lo = fmin(lo, p[e]); hi = fmax(hi, p[cols + e]); dev = fmax(dev, p[2 * cols + e]);
  }
  if (box_min) box_min[e] = lo;
  if (box_max) box_max[e] = hi;
  if (max_dev) max_dev[e] = dev;
}

// Convex-hull candidates of the per-stage point clouds (traj[:, i0, t], traj[:, i1, t]) over the samples
// (benchmarking/generate_convex_hull.py:88-100).  A hull vertex maximises SOME direction, so:
//   pass 1: per stage, the extreme sample in each of HULL_DIRS fixed directions (argmax of a dot product);
//   pass 2: every point strictly inside the polygon of those extremes cannot be a vertex and is dropped;
//           the few survivors' sample indices are compacted per stage.
// The exact hull of the survivors (a few hundred points) is a host job (gpmpc_api.cu, monotone chain).
#define HULL_DIRS 16

struct HullDirs {
  double cx[HULL_DIRS], cy[HULL_DIRS];
};

// block = (32 stage lanes) x (blockDim.y sample lanes); partial [gridDim.y][H1][HULL_DIRS] of (value, index)
__global__ void k_hull_extremes_partial(int ns, int nx, int H1, int i0, int i1, HullDirs dirs,
                                        const double* __restrict__ traj, double* __restrict__ pval,
                                        int* __restrict__ pidx) {
  const int t = blockIdx.x * 32 + threadIdx.x;
  const int per = (ns + gridDim.y - 1) / gridDim.y;
  const int s0 = blockIdx.y * per, s1 = min(ns, s0 + per);
  double best[HULL_DIRS];
  int bi[HULL_DIRS];
#pragma unroll
  for (int k = 0; k < HULL_DIRS; ++k) { best[k] = -INFINITY; bi[k] = -1; }
  if (t < H1) {
    for (int s = s0 + threadIdx.y; s < s1; s += blockDim.y) {
      const double px = traj[((size_t)s * nx + i0) * H1 + t], py = traj[((size_t)s * nx + i1) * H1 + t];
#pragma unroll
      for (int k = 0; k < HULL_DIRS; ++k) {
        const double v = dirs.cx[k] * px + dirs.cy[k] * py;
        if (v > best[k]) { best[k] = v; bi[k] = s; }
      }
    }
  }
  __shared__ double sv[8][32];
  __shared__ int si[8][32];
#pragma unroll
  for (int k = 0; k < HULL_DIRS; ++k) {
    sv[threadIdx.y][threadIdx.x] = best[k];
    si[threadIdx.y][threadIdx.x] = bi[k];
    __syncthreads();
    if (threadIdx.y == 0 && t < H1) {
      double v = best[k];
      int ix = bi[k];
      for (int r = 1; r < blockDim.y; ++r) {
        const double ov = sv[r][threadIdx.x];
        const int oi = si[r][threadIdx.x];
        if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) { v = ov; ix = oi; }
      }
      pval[((size_t)blockIdx.y * H1 + t) * HULL_DIRS + k] = v;
      pidx[((size_t)blockIdx.y * H1 + t) * HULL_DIRS + k] = ix;
    }
    __syncthreads();
  }
}

// reduces the partials and stores the polygon of stage t: poly [H1][HULL_DIRS][2] coordinates of the extremes,
// ext_idx [H1][HULL_DIRS] their sample indices (always hull candidates, also when the cloud is a single point)
__global__ void k_hull_extremes_final(int nparts, int nx, int H1, int i0, int i1, const double* __restrict__ pval,
                                      const int* __restrict__ pidx, const double* __restrict__ traj,
                                      double* __restrict__ poly, int* __restrict__ ext_idx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H1 * HULL_DIRS) return;
  const int t = e / HULL_DIRS;
  double v = -INFINITY;
  int ix = -1;
  for (int p = 0; p < nparts; ++p) {
    const double ov = pval[(size_t)p * H1 * HULL_DIRS + e];
    const int oi = pidx[(size_t)p * H1 * HULL_DIRS + e];
    if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) { v = ov; ix = oi; }
  }
  ext_idx[e] = ix;
  poly[(size_t)e * 2 + 0] = traj[((size_t)ix * nx + i0) * H1 + t];
  poly[(size_t)e * 2 + 1] = traj[((size_t)ix * nx + i1) * H1 + t];
}

// pass 2: keeps sample s at stage t unless it lies strictly inside the extremes' polygon (directions are in
// counter-clockwise order, so consecutive distinct extremes form a convex CCW polygon).  Survivors are appended to
// cand [H1][cap] (+ their coordinates cand_xy [H1][cap][2]) through a per-stage counter (order fixed up on the host by sorting).
__global__ void k_hull_filter(int ns, int nx, int H1, int i0, int i1, int cap, const double* __restrict__ traj,
                              const double* __restrict__ poly, int* __restrict__ cand, double* __restrict__ cand_xy,
                              int* __restrict__ count) {
  const int t = blockIdx.x * 32 + threadIdx.x;
  if (t >= H1) return;
  double vx[HULL_DIRS], vy[HULL_DIRS];
#pragma unroll
  for (int k = 0; k < HULL_DIRS; ++k) {
    vx[k] = poly[((size_t)t * HULL_DIRS + k) * 2];
    vy[k] = poly[((size_t)t * HULL_DIRS + k) * 2 + 1];
  }
  const int per = (ns + gridDim.y - 1) / gridDim.y;
  const int s0 = blockIdx.y * per, s1 = min(ns, s0 + per);
  for (int s = s0 + threadIdx.y; s < s1; s += blockDim.y) {
    const double px = traj[((size_t)s * nx + i0) * H1 + t], py = traj[((size_t)s * nx + i1) * H1 + t];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < HULL_DIRS; ++k) {
      const int k2 = (k + 1) % HULL_DIRS;
      const double ex = vx[k2] - vx[k], ey = vy[k2] - vy[k];
      // degenerate edge (same extreme in both directions) constrains nothing
      const double cr = ex * (py - vy[k]) - ey * (px - vx[k]);
      const double scale = fabs(ex) * fabs(py - vy[k]) + fabs(ey) * fabs(px - vx[k]);
      if ((ex != 0.0 || ey != 0.0) && !(cr > 1e-12 * scale)) inside = false;  // on / outside / too close to call: keep
    }
    if (!inside) {
      const int slot = atomicAdd(count + t, 1);
      if (slot < cap) {
        cand[(size_t)t * cap + slot] = s;
        cand_xy[((size_t)t * cap + slot) * 2] = px;
        cand_xy[((size_t)t * cap + slot) * 2 + 1] = py;
      }
    }
  }
}
