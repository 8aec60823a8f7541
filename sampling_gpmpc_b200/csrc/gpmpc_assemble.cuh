// K3: what Agent.dyn_fg_jacobians does around the GP call, as one elementwise kernel, and the state
// bookkeeping of a forward rollout (so a whole horizon runs on the stream without host round trips).
#pragma once
#include "gpmpc_state.cuh"

// transform_sensitivity (identity: pendulum1D.py:240-241; residual car: car_model_residual.py:211-224)
// followed by the scatter into pad_g (src/agent.py:545-550).  Returns entry p of the padded row.
template <typename ENV>
__device__ __forceinline__ double transformed_entry(const ENV& env, const double* __restrict__ yg, int T,
                                                    double v, int p) {
  if (env.transform == 1) {
    // [v g, v dg/dphi, g, v dg/ddelta]; with T == 1 torch broadcasting copies g into every slot
    if (p == 2) return yg[0];
    int src = p == 0 ? 0 : (p == 1 ? 1 : 2);
    if (src > T - 1) src = T - 1;
    return v * yg[src];
  }
  return yg[p < T ? p : T - 1];
}

// out[s][i][h][:] = [f_i, df_i/dx, df_i/du] + sum_j B_d[i][j] * pad(transform(y_gp[s][j][h][:]))
__global__ void k_assemble(gpmpc_env env, int ns, int H, int T, const double* __restrict__ xu,
                           const double* __restrict__ y_gp, double* __restrict__ out) {
  const int nx = env.nx, nu = env.nu, nz = nx + nu, w = 1 + nz;
  const long long total = (long long)ns * nx * H;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(idx % H);
    const int i = (int)((idx / H) % nx);
    const long long s = idx / ((long long)H * nx);
    const double* z = xu + ((s * nx + 0) * H + h) * nz;  // row 0 of the nx tiled copies
    double* o = out + ((s * nx + i) * H + h) * w;
    double f = 0.0;
    for (int k = 0; k < nz; ++k) {
      const double Fik = env.F_known[i * nz + k];
      f += Fik * z[k];
      o[1 + k] = Fik;
    }
    o[0] = f;
    const double v = nz > 3 ? z[3] : 0.0;
    for (int j = 0; j < env.g_ny; ++j) {
      const double bij = env.B_d[i * env.g_ny + j];
      if (bij == 0.0) continue;
      const double* yg = y_gp + ((s * env.g_ny + j) * H + h) * T;
      for (int p = 0; p < env.n_pad; ++p) o[env.pad_g[p]] += bij * transformed_entry(env, yg, T, v, p);
    }
  }
}

// Row i of the next state of one sample from z = [x, u] of this step and the g_ny sampled value rows y_of(j)[T]:
// x+_i = F_known[i] z + sum_j B_d[i][j] pad(transform(y_j))[0]   (k_rollout_state evaluates every row in one thread, the
// fused-horizon kernel one row per lane: same expression order, bit-identical trajectories)
template <typename ENV, typename YOF>
__device__ __forceinline__ double rollout_next_state_row(const ENV& env, int T, const double* z, YOF y_of, int i) {
  const int nz = env.nx + env.nu;
  const double v = nz > 3 ? z[3] : 0.0;
  double f = 0.0;
  for (int k = 0; k < nz; ++k) f += env.F_known[i * nz + k] * z[k];
  for (int j = 0; j < env.g_ny; ++j) {
    const double bij = env.B_d[i * env.g_ny + j];
    if (bij != 0.0) {
      const double* yg = y_of(j);
      for (int p = 0; p < env.n_pad; ++p)
        if (env.pad_g[p] == 0) f += bij * transformed_entry(env, yg, T, v, p);
    }
  }
  return f;
}
template <typename ENV, typename YOF>
__device__ __forceinline__ void rollout_next_state(const ENV& env, int T, const double* z, YOF y_of, double* xc) {
  for (int i = 0; i < env.nx; ++i) xc[i] = rollout_next_state_row(env, T, z, y_of, i);
}

// input k of z = [x, u_t (+ feedback)]  (simulate_forward_sampling_car.py:122: u = u_ff - K (x_equi - x))
template <typename ENV>
__device__ __forceinline__ double rollout_input_row(const ENV& env, const double* xc, const double* u_t, int k) {
  double u = u_t[k];
  if (env.use_feedback)
    for (int i = 0; i < env.nx; ++i) u -= env.K_fb[k * env.nx + i] * (env.x_equi[i] - xc[i]);
  return u;
}
template <typename ENV>
__device__ __forceinline__ void rollout_inputs(const ENV& env, const double* xc, const double* u_t, double* z) {
  for (int i = 0; i < env.nx; ++i) z[i] = xc[i];
  for (int k = 0; k < env.nu; ++k) z[env.nx + k] = rollout_input_row(env, xc, u_t, k);
}

// Rollout bookkeeping, one thread per sample.  phase 0: x_cur = x0.  phase 1: x_cur = F xu + B_d pad(y)[0].
// Then (if t < n_steps) builds xu_t = [x_cur, u_t (+feedback)], records traj[:, :, t] and gathers the GP input.
__global__ void k_rollout_state(gpmpc_env env, int ns, int T, int t, int n_steps, int phase,
                                const double* __restrict__ x0, const double* __restrict__ u_ff,
                                const double* __restrict__ y_gp, double* __restrict__ xu,
                                double* __restrict__ xstar, double* __restrict__ traj) {
  const int nx = env.nx, nu = env.nu, nz = nx + nu;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  double xc[GPMPC_MAX_NX];
  double* z = xu + (size_t)s * nz;
  if (phase == 0) {
    for (int i = 0; i < nx; ++i) xc[i] = x0[(size_t)s * nx + i];
  } else {
    double zl[2 * GPMPC_MAX_NX];
    for (int k = 0; k < nz; ++k) zl[k] = z[k];
    rollout_next_state(env, T, zl, [&](int j) { return y_gp + ((size_t)s * env.g_ny + j) * T; }, xc);
  }
  for (int i = 0; i < nx; ++i) traj[((size_t)s * nx + i) * (n_steps + 1) + t] = xc[i];
  if (t >= n_steps) return;
  double zn[2 * GPMPC_MAX_NX];
  rollout_inputs(env, xc, u_ff + (size_t)t * nu, zn);
  for (int k = 0; k < nz; ++k) z[k] = zn[k];
  for (int j = 0; j < env.g_ny; ++j)
    for (int a = 0; a < env.d; ++a) xstar[((size_t)s * env.g_ny + j) * env.d + a] = zn[env.g_idx_inputs[a]];
}

// One step of the rejection rollout of Agent.prepare_dynamics_set (src/agent.py:365-415), one thread per sample:
//   x_next   = known_dyn(xu) + B_d g_val                        (:381-383; g_val = the sampled VALUE task of every output)
//   survive  = prod_i ( |x_target_i - x_next_i| - c_i < 0 )     (:384-389, the sample-survival test)
//   samples_left *= survive
//   xu_next  = [x_next, u_next] tiled over the nx rows          (:407-415), when u_next is given
// xu [ns][nx][1][nx+nu] (row 0 of the nx tiled copies is read), y [ns][g_ny][1][T], x_target [ns][nx].
__global__ void k_fs_advance(gpmpc_env env, int ns, int T, const double* __restrict__ xu, const double* __restrict__ y,
                             const double* __restrict__ x_target, double c_i, const double* __restrict__ u_next,
                             int* __restrict__ samples_left, double* __restrict__ x_next, double* __restrict__ xu_next) {
  const int nx = env.nx, nu = env.nu, nz = nx + nu;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  double z[2 * GPMPC_MAX_NX], xn[GPMPC_MAX_NX];
  for (int k = 0; k < nz; ++k) z[k] = xu[(size_t)s * nx * nz + k];
  rollout_next_state(env, T, z, [&](int j) { return y + ((size_t)s * env.g_ny + j) * T; }, xn);
  int ok = 1;
  for (int i = 0; i < nx; ++i) {
    ok = ok && (fabs(x_target[(size_t)s * nx + i] - xn[i]) - c_i < 0.0);
    x_next[(size_t)s * nx + i] = xn[i];
  }
  samples_left[s] *= ok;
  if (!xu_next) return;
  for (int r = 0; r < nx; ++r) {
    double* o = xu_next + ((size_t)s * nx + r) * nz;
    for (int i = 0; i < nx; ++i) o[i] = xn[i];
    for (int k = 0; k < nu; ++k) o[nx + k] = u_next[k];
  }
}

// get_g_xu_hat (src/environments/*.py: xu_hat[:, 0:g_ny, :, g_idx_inputs]) as a kernel: xg[b][h][a] = xu[s][0][h][g_idx[a]]
// (the nx rows of xu are tiled copies; the reference asserts that, pendulum1D.py:165-170)
__global__ void k_gather_gp_inputs(gpmpc_env env, int ns, int H, const double* __restrict__ xu, double* __restrict__ xg) {
  const int nz = env.nx + env.nu, d = env.d, g_ny = env.g_ny;
  const long long total = (long long)ns * g_ny * H * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(idx % d);
    const int hh = (int)((idx / d) % H);
    const long long s = idx / ((long long)d * H * g_ny);
    xg[idx] = xu[((s * env.nx + 0) * H + hh) * nz + env.g_idx_inputs[a]];
  }
}
