/*
 * gpmpc_b200.h -- C ABI of libgpmpc_b200.so: batched posterior sampling of GP dynamics on B200 (sm_100a).
 *
 * The reference (manish-pra/sampling-gpmpc) has no FFI for this path: its seam is Python duck typing
 * (SURVEY.md section 8b).  Each entry point below names the reference code it replaces; the Python
 * binding a maintainer would add is a ctypes stub (sampling_gpmpc_b200/engine.py).
 *
 * Conventions
 *   - All numeric arrays are float64, contiguous, row-major, in exactly the reference's layouts:
 *     a "batch element" is b = s*g_ny + j (sample s, GP output j)  <-> torch batch_shape (ns, g_ny);
 *     test inputs  x  [B][H][d];  posterior quantities  [B][H][T]  with T fastest (GPyTorch's
 *     interleaved multitask order: scalar index = point*T + task, task 0 = value, task a = d/dx_a).
 *   - Pointers marked DEVICE are CUDA device pointers owned by the caller (e.g. torch tensors); the
 *     library borrows them for the duration of the call.  Pointers marked HOST are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every call is
 *     asynchronous w.r.t. the host and ordered on that stream; nothing here synchronises unless
 *     its comment says so.
 *   - Return value: 0 ok; <0 invalid argument / capacity / CUDA error (gpmpc_last_error() has text).
 *     Numerical status never crosses as a return code: it is written per batch element into the
 *     caller's `jitter_level` array and the handle's device status word (gpmpc_status()).
 *   - One handle per Agent per GPU; not thread-safe; no internal threads.
 *   - There is no CPU fallback: every entry point fails with GPMPC_ERR_CUDA without a device.
 */
#ifndef GPMPC_B200_H
#define GPMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPMPC_MAX_D 6   /* GP input dimension (BASELINE.json: "input dim <= 6") */
#define GPMPC_MAX_T 7   /* tasks per point: 1 (value only) or d+1 (value + gradient) */
#define GPMPC_MAX_NX 8  /* state / input dims of the dynamics for the assembly kernel */

#define GPMPC_OK 0
#define GPMPC_ERR_ARG (-1)
#define GPMPC_ERR_CAPACITY (-2)
#define GPMPC_ERR_CUDA (-3)
#define GPMPC_ERR_STATE (-4)

/* bits of the device status word (gpmpc_status) */
#define GPMPC_ST_TRAIN_JITTER 1u   /* real-data block needed jitter (psd_safe_cholesky ladder, level in bits 8..9) */
#define GPMPC_ST_TRAIN_NOT_PD 2u   /* real-data block not PD after the ladder (GPyTorch: NotPSDError) */
#define GPMPC_ST_SAMPLE_NOT_PD 4u  /* some posterior covariance not PD after the ladder (jitter_level[b] == 4) */
#define GPMPC_ST_APPEND_NOT_PD 8u  /* a bordered block (Sigma* + noise) was not PD when conditioning */
#define GPMPC_ST_NAN_INPUT 16u     /* NaN reached a Cholesky (GPyTorch: NanError) */
#define GPMPC_ST_SAMPLE_EIG 32u    /* a draw fell back to the eigen root for the whole batch (GPyTorch: root_decomposition
                                      catches the NotPSDError and uses symeig; jitter_level[b] == 4 for every b) */

typedef struct gpmpc_handle gpmpc_handle;

typedef struct {
  int32_t ns;          /* dynamics samples held by this handle (this rank's shard) */
  int32_t g_ny;        /* GP outputs per sample */
  int32_t d;           /* GP input dimension g_nx + g_nu */
  int32_t T;           /* tasks per point: 1 or d+1 */
  int32_t n_real;      /* real (measured) training points, shared by all samples */
  int32_t cap_points;  /* capacity: hallucinated points per batch element (grown by gpmpc_reserve) */
} gpmpc_dims;

/* sample_gp's post-processing switches (src/agent.py:646-708); a negative threshold disables, like the yaml */
typedef struct {
  double beta;              /* Dyn_gp_beta: truncate to mean +- beta*sqrt(variance); <0 = no truncation */
  double variance_is_zero;  /* Dyn_gp_variance_is_zero */
  int32_t unclamped_sqrt_1x1; /* 1 = GPyTorch's no-base-samples 1x1 path (sqrt without clamp), else 0 */
  int32_t flags;            /* GPMPC_OPT_* */
} gpmpc_sample_opts;
/* Without this flag a joint Cholesky that still fails after the jitter ladder makes the WHOLE batch draw through the
 * eigen root evecs*sqrt(clamp(evals,0)), like GPyTorch's root_decomposition (src/agent.py:641); with it the failing
 * elements return NaN and GPMPC_ST_SAMPLE_NOT_PD is the caller's to raise (psd_safe_cholesky's NotPSDError). */
#define GPMPC_OPT_NO_EIG_FALLBACK 1

/* env hooks of Agent.dyn_fg_jacobians as plain data (src/agent.py:532-564, src/environments/*.py) */
typedef struct {
  int32_t nx, nu, g_ny, d;
  int32_t g_idx_inputs[GPMPC_MAX_D];        /* columns of [x,u] that feed the GP (get_g_xu_hat) */
  int32_t pad_g[GPMPC_MAX_NX];              /* where transformed [g, dg] land in [f, df/dx, df/du] */
  int32_t n_pad;                            /* entries of pad_g in use */
  int32_t transform;                        /* 0 identity, 1 car-residual: [v g, v dg/dphi, g, v dg/ddelta] */
  double B_d[GPMPC_MAX_NX * GPMPC_MAX_NX];  /* (nx, g_ny) row-major */
  double F_known[GPMPC_MAX_NX * 2 * GPMPC_MAX_NX]; /* (nx, nx+nu) row-major: f(x,u) = F [x;u] */
  int32_t use_feedback;                     /* rollouts: u = u_ff - K (x_equi - x)  (simulate_forward_sampling_car.py:122) */
  int32_t reserved;
  double K_fb[GPMPC_MAX_NX * GPMPC_MAX_NX]; /* (nu, nx) row-major */
  double x_equi[GPMPC_MAX_NX];
} gpmpc_env;

/* ---- one SQP GP linearisation in one call ----------------------------------------------------------------- */

/* What src/solver.py:84-94 does per SQP iteration around the GP -- get_g_xu_hat, model_i(x), .sample(base_samples), sample_gp's
 * post-processing, update_hallucinated_Dyn_dataset, the assembly of dyn_fg_jacobians and its device->host copy -- queued on
 * `stream` by ONE call:  xu [ns][nx][H][nx+nu] (HOST, pinned, when xu_on_host, else DEVICE) = get_batch_x_hat's tensor;
 * eps DEVICE [B][H][T];  reset_first: the deferred reset of src/agent.py:261-272 (the model of this call still holds the
 * previous MPC step's points, the data set it appends to is empty);  mean, var, y DEVICE [B][H][T], jitter_level [B];
 * out DEVICE [ns][nx][H][1+nx+nu] = [f, df/dx, df/du], also copied to out_host (HOST, pinned) when not NULL.
 * Nothing synchronises: wait for `stream` before reading out_host, then gpmpc_status. */
int gpmpc_linearise(gpmpc_handle* h, const gpmpc_env* env, const double* xu, int32_t xu_on_host, int32_t H, const double* eps,
                    const gpmpc_sample_opts* opts, int32_t reset_first, double* mean, double* var, double* y,
                    int32_t* jitter_level, double* out, double* out_host, void* stream);

/* ---- rejection rollout (Agent.prepare_dynamics_set, src/agent.py:331-443) ------------------------------------ */

/* Forgets every hallucinated point from index n_points on (their factor rows are simply no longer used): the end of
 * prepare_dynamics_set's forward-sampling data set (FS_*_train_batch, src/agent.py:344-345,399-405), which the reference
 * drops by re-fitting on [real || hallucinated] (:438-441). */
int gpmpc_truncate_hallucinated(gpmpc_handle* h, int32_t n_points);

/* One step of the rejection rollout, all DEVICE: xu [ns][nx][1][nx+nu] current [x, u] (tiled over nx as get_batch_x_hat lays
 * it out), y [B][1][T] the draw at it, x_target [ns][nx] = X_soln[i+1], c_i = ci_list[i];  x_next [ns][nx] = known_dyn(xu) +
 * B_d g (src/agent.py:381-383), samples_left [ns] int32 *= prod_i(|x_target - x_next| - c_i < 0) (:384-389); with u_next
 * [nu] also xu_next [ns][nx][1][nx+nu] = [x_next, u_next] (:407-415).  u_next = xu_next = NULL on the last step. */
int gpmpc_fs_advance(gpmpc_handle* h, const gpmpc_env* env, const double* xu, const double* y, const double* x_target,
                     double c_i, const double* u_next, int32_t* samples_left, double* x_next, double* xu_next, void* stream);

/* ---- several reference Agents in one handle ------------------------------------------------------------- */

/* benchmarking/simulate_true_reachable_set.py:167-259 builds a NEW Agent of num_dyn_samples samples for each of its 10^4
 * repeats; the B200 path rolls all repeats out as one batch.  With Dyn_gp_min_data_dist >= 0 the Agent's
 * update_hallucinated_Dyn_dataset (src/agent.py:164-202) couples the samples of ONE Agent: a new point whose label was NaN'd
 * (closer than min_dist to the element's own data set) for ANY element of the Agent is masked out of the Agent's model by
 * observation_nan_policy("mask"), and it is not stored at all if it was filtered for ALL samples of some output.
 * group_size > 0 declares consecutive blocks of group_size samples to be one Agent each: every gpmpc_step / gpmpc_rollout
 * step then runs that filter and those two reductions PER GROUP (no host round trip) and a masked / dropped point enters the
 * group's factors as null rows.  group_size = 0 switches it off.  Call on an empty hallucinated set.  While it is on the
 * block entry points (gpmpc_posterior with hallucinated points, gpmpc_append) refuse the handle. */
int gpmpc_set_grouping(gpmpc_handle* h, int32_t group_size, double min_dist);
/* DEVICE out[B * num_hallucinated] uint8: 0 = point in the element's factor, 1 = recorded but masked, 2 = dropped (not part of
 * the Agent's data set: skip it when reading gpmpc_export_hallucinated).  All 0 without grouping. */
int gpmpc_export_point_states(const gpmpc_handle* h, uint8_t* out, void* stream);

/* ---- tuning switches ------------------------------------------------------------------------------- */

/* name = "rollout_fused"  (0 default: one gpmpc_step per horizon step; 1: gpmpc_rollout runs the whole horizon in ONE launch,
 *                          csrc/gpmpc_horizon.cuh, where the shape allows; 2: automatic = one launch only while the step-wise
 *                          rollout would be launch-latency bound, i.e. for the reference's own sample counts (20 ... ~2000),
 *                          where it is 1.2 - 2.5x faster; at the bench shape it is slower.  The one-launch kernel cannot take
 *                          the batch-wide eigen-root fallback of a failed jitter ladder: it flags GPMPC_ST_SAMPLE_NOT_PD and
 *                          the caller repeats the rollout with "rollout_fused" 0 -- rollout.ForwardRollout does),
 *        "hz_groups"      (cap on the samples one CTA of the fused kernel holds at a time; 0 = as many as fit),
 *        "hz_stagger_ns"  (spread of the sample groups' start times in the fused kernel; -1 = automatic, 0 = none),
 *        "prefactor_next" (1: the next gpmpc_posterior call that draws also factorises Sigma* + noise, beside the draw's own
 *                          Cholesky, for the gpmpc_append of the same points that follows -- what gpmpc_linearise does by itself;
 *                          one-shot, ignored where it does not apply),
 *        "step_warps_cap" / "step_grid_cap" / "sr_grid_cap" (experiments: warps per CTA / CTAs of the fused step kernel, 0 = no cap; two handles
 *                          on two streams can then share the SMs -- measured: no gain).
 * The results do not depend on any of them (bit-identical trajectories). */
int gpmpc_set_option(gpmpc_handle* h, const char* name, int64_t value);
/* reads an option back; also "last_rollout_fused" (1 if the last gpmpc_rollout took the one-launch kernel) */
int gpmpc_get_option(gpmpc_handle* h, const char* name, int64_t* value);

/* ---- base samples (host only) -------------------------------------------------------------------- */

/* The truncated standard-normal base samples of Agent.random_vector_within_bounds (src/agent.py:76-104), drawn from
 * torch's CPU generator stream-identically and without the Python loop: `slots` candidates of `n` float64 scalars each
 * (slots = n_mpc * n_sqp * ns, n = g_ny * H * T), every candidate one emulated torch.normal(0, 1, size=(1, g_ny, H, T))
 * call, redrawn until all |w| <= beta.  rng_state: HOST, the bytes of torch.get_rng_state() (CPUGeneratorImplState,
 * state_bytes must be 5056); advanced in place -- write it back with torch.set_rng_state().  out: HOST [slots * n].
 * Returns the number of candidates drawn (>= slots), < 0 on a bad argument.  No CUDA involved. */
int64_t gpmpc_base_samples(uint8_t* rng_state, int64_t state_bytes, int64_t slots, int64_t n, double beta, double* out);

/* ---- lifetime ---------------------------------------------------------------------------------- */

/* Allocates the persistent per-sample state on the current CUDA device.
 * Replaces: Agent.__init__'s empty Hallcinated_*_train + real_data_batch (src/agent.py:52-67,204-214). */
int gpmpc_create(const gpmpc_dims* dims, gpmpc_handle** out);
int gpmpc_destroy(gpmpc_handle* h);
const char* gpmpc_last_error(const gpmpc_handle* h); /* h may be NULL: error of the last failed create */

/* ---- model definition -------------------------------------------------------------------------- */

/* Hyper-parameters per GP output (HOST arrays): lengthscale[g_ny*d], outputscale[g_ny],
 * noise[g_ny*T] = task_noises[t] + noise (the diagonal MultitaskGaussianLikelihood(rank=0) adds to the
 * training block), jitter = Dyn_gp_jitter.
 * Replaces: BatchMultitaskGPModelWithDerivatives_fromParams.__init__ (src/GP_model.py:121-143) and
 * gpytorch.settings.cholesky_jitter (src/agent.py:634-638).  The reference tiles these over samples;
 * they are identical for every sample, so the library keeps one copy per output. */
int gpmpc_set_hypers(gpmpc_handle* h, const double* lengthscale, const double* outputscale,
                     const double* noise, double jitter);

/* Real training data (DEVICE): X[n_real*d], Y[g_ny*n_real*T] with NaN = unobserved slot.  Factorises the
 * shared block once per output (kernel K0; from 768 observed scalars on the blocked tensor-core form of csrc/gpmpc_k0.cuh,
 * environment GPMPC_K0_BLOCKED_MIN_M moves that switch) and clears the hallucinated set.  SYNCHRONISES once to read the
 * observed-slot count.  Replaces: the real-data part of every ExactGP re-fit (src/agent.py:223-248 ->
 * gpytorch ExactGP / DefaultPredictionStrategy, psd_safe_cholesky of K_oo + Sigma). */
int gpmpc_set_real_data(gpmpc_handle* h, const double* X, const double* Y, void* stream);

/* Drops all hallucinated points (src/agent.py:261-272: the reset when sqp_iter == 0). */
int gpmpc_reset_hallucinated(gpmpc_handle* h);
/* Ensures room for `cap_points` hallucinated points per batch element (re-lays the factor out). */
int gpmpc_reserve(gpmpc_handle* h, int32_t cap_points, void* stream);
/* 1 (default): appended points extend the factor.  0: they are only recorded (the reference's
 * use_model_without_derivatives "real data only" model, src/agent.py:221-226). */
int gpmpc_set_condition_on_hallucinated(gpmpc_handle* h, int32_t on);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* Posterior of every batch element at its own H test points, joint over the q = H*T scalars.
 * x DEVICE [B*H*d]; outputs DEVICE mean[B*H*T], var[B*H*T] (diag clamped at 1e-10 like .variance).
 * If eps != NULL also draws y = mean + chol(Sigma*) eps with GPyTorch's jitter ladder and applies
 * sample_gp's post-processing (opts); y DEVICE [B*H*T], jitter_level DEVICE int32 [B] (0..3, 4 = not PD).
 * The joint covariance stays cached in the handle for gpmpc_sample / gpmpc_append.
 * Replaces: self.model_i(x_input) and .sample(base_samples) (src/agent.py:640-641) and the
 * post-processing at src/agent.py:646-663,701-708. */
int gpmpc_posterior(gpmpc_handle* h, const double* x, int32_t H, double* mean, double* var,
                    const double* eps, const gpmpc_sample_opts* opts, double* y, int32_t* jitter_level,
                    void* stream);

/* Draw from the posterior cached by the last gpmpc_posterior (MultitaskMultivariateNormal.sample). */
int gpmpc_sample(gpmpc_handle* h, const double* eps, const gpmpc_sample_opts* opts, double* y,
                 int32_t* jitter_level, void* stream);

/* Condition every batch element on its H new points: x DEVICE [B*H*d], y DEVICE [B*H*T] (labels as they
 * stand after post-processing).  point_active HOST uint8[H] or NULL (= all): points with 0 are recorded in
 * the data set but do not enter the factor (GPyTorch's any-over-batch NaN mask).
 * Appends T rows per point to each element's bordered Cholesky factor: [W^T | chol(Sigma* + noise)].
 * Replaces: Agent.update_hallucinated_Dyn_dataset + the next train_hallucinated_dynGP re-fit
 * (src/agent.py:164-202,216-248). */
int gpmpc_append(gpmpc_handle* h, const double* x, const double* y, const uint8_t* point_active,
                 int32_t H, void* stream);

/* The same with one flag per new SCALAR: scalar_active HOST uint8[H*T] (point h, task t -> [h*T + t]; NULL = all).
 * observation_nan_policy("mask") drops single label slots, e.g. the derivative slots prepare_dynamics_set NaNs
 * (src/agent.py:402): a point may then enter the factor with some of its T rows only.  While such a point is in the
 * factor, model calls are served by the row-by-row kernels (k_posterior) instead of the fused / tensor-core ones, which
 * assume T consecutive rows per point; gpmpc_reset_hallucinated clears the condition. */
int gpmpc_append_masked(gpmpc_handle* h, const double* x, const double* y, const uint8_t* scalar_active,
                        int32_t H, void* stream);

/* Fused rollout step (H = 1): posterior + draw + post-processing + append in ONE launch, one warp per
 * batch element, the element's factor rows streamed once from HBM.  Same outputs as gpmpc_posterior.
 * Replaces one iteration of the loops at benchmarking/simulate_true_reachable_set.py:179-259 and
 * benchmarking/simulate_forward_sampling_car.py:117-138 (model call .. update_hallucinated_Dyn_dataset). */
int gpmpc_step(gpmpc_handle* h, const double* x, const double* eps, const gpmpc_sample_opts* opts,
               double* mean, double* var, double* y, int32_t* jitter_level, void* stream);

/* dyn_fg_jacobians' assembly: out DEVICE [ns][nx][H][1+nx+nu] = df_known + B_d pad(transform(y_gp)).
 * xu DEVICE [ns][nx][H][nx+nu] (the reference's tiled layout; row 0 of the nx copies is read),
 * y_gp DEVICE [ns][g_ny][H][T].  Replaces src/agent.py:537-554 and the env hooks in SURVEY.md 8(a) a14. */
int gpmpc_assemble(gpmpc_handle* h, const gpmpc_env* env, const double* xu, const double* y_gp, int32_t H,
                   double* out, void* stream);

/* Whole forward rollout on the stream, no host round trips: for t < n_steps
 *   xu_t = [x_t, u_t (+ feedback)], traj[:, :, t] = x_t, y = step(xu_t[g_idx]), x_{t+1} = F xu_t + B_d pad(y)[0].
 * x0 DEVICE [ns*nx]; u_ff DEVICE [n_steps*nu]; eps DEVICE [n_steps][B*T]; traj DEVICE [ns][nx][n_steps+1].
 * Replaces the loop at benchmarking/simulate_forward_sampling_car.py:117-138. */
int gpmpc_rollout(gpmpc_handle* h, const gpmpc_env* env, const double* x0, const double* u_ff,
                  const double* eps, const gpmpc_sample_opts* opts, int32_t n_steps, double* traj,
                  void* stream);

/* The same rollout with the base samples still in flight from the host: eps_ready HOST array of n_steps CUDA event
 * handles (cudaEvent_t, NULL entries allowed, the array itself may be NULL); before step t is queued the stream waits
 * for eps_ready[t], recorded by the caller on its copy stream after the host-to-device copy of eps[t] was queued.
 * The upload of epistimic_random_vector (simulate_forward_sampling_car.py:84-90, 126) then overlaps the horizon
 * instead of preceding it. */
int gpmpc_rollout_gated(gpmpc_handle* h, const gpmpc_env* env, const double* x0, const double* u_ff,
                        const double* eps, const gpmpc_sample_opts* opts, int32_t n_steps, double* traj,
                        void* const* eps_ready, void* stream);

/* ---- the data-set rules around the draw (Dyn_gp_min_data_dist >= 0) ------------------------------ */

/* sample_gp's min-distance overwrite + truncation (src/agent.py:666-708), applied to y DEVICE [B*H*T] in place:
 * where the nearest FULLY OBSERVED training point of the model (real, then recorded hallucinated; a point with a
 * NaN target counts as infinitely far, agent.py:674-679) is within min_dist of test point (b, h), its targets
 * replace the draw; then y is clipped to mean +- beta*sqrt(var) (beta < 0: no clipping; mean/var may be NULL). */
int gpmpc_min_dist_overwrite(gpmpc_handle* h, const double* x, int32_t H, const double* mean, const double* var,
                             double min_dist, double beta, double* y, void* stream);

/* update_hallucinated_Dyn_dataset's filter (src/agent.py:166-181): labels y DEVICE [B*H*T] of every new point that
 * is within min_dist of ANY input already in its element's data set (real; plus the recorded hallucinated ones
 * if use_hallucinated) become NaN.  counts DEVICE int32 [g_ny*H] = number of this handle's samples filtered per
 * (output, point): the caller sums it over ranks and forms the reference's all-over-samples (drop the point,
 * agent.py:186-191) and any-over-batch (GPyTorch's NaN mask, SURVEY.md A.4 -> gpmpc_append's point_active). */
int gpmpc_filter_new_points(gpmpc_handle* h, const double* x, int32_t H, double min_dist, int32_t use_hallucinated,
                            double* y, int32_t* counts, void* stream);

/* ---- formats either side of the path ---------------------------------------------- */

/* The acados stage parameter vector of every stage in one launch (src/solver.py:98-131, layout fixed by
 * src/utils/model.py:34-41):  out DEVICE [H][P],  P = ns*(nx*nx + nx*nu + 2*nx) + n_tail,
 *   out[stage] = [ per sample i:  y_grad_i (nx*nx, row-major) | u_grad_i (nx*nu) | x_h[stage][i*nx..] | gp_val_i (nx) ]
 *                ++ tail[stage]           (tail = hstack(u_h, xg, w, tilde_eps) rows, DEVICE [H][n_tail])
 * lin DEVICE [ns][nx][H][1+nx+nu] is gpmpc_assemble's output, x_h DEVICE [H][ns*nx] the SQP iterate.
 * use_feedback_K: y_grad += u_grad @ env->K_fb (solver.py:90). */
int gpmpc_pack_plin(gpmpc_handle* h, const gpmpc_env* env, const double* lin, const double* x_h, const double* tail,
                    int32_t n_tail, int32_t H, int32_t use_feedback_K, double* out, void* stream);

/* Stage-wise reductions over the trajectories traj DEVICE [ns][nx][H1]: bounding box and the tightening
 * Delta[i][t] = max_n |traj[n][i][t] - ref[i][t]| (extra/approx_sampling_mpc/README.md:19-27).  Outputs DEVICE
 * [nx][H1], any may be NULL; ref DEVICE [nx][H1] (required for max_dev).  Sharded samples: reduce the three
 * small arrays with MIN / MAX over ranks instead of gathering the trajectories. */
int gpmpc_traj_stats(gpmpc_handle* h, const double* traj, int32_t ns, int32_t nx, int32_t H1, const double* ref,
                     double* box_min, double* box_max, double* max_dev, void* stream);

/* Convex hull of the per-stage point clouds (traj[:, i0, t], traj[:, i1, t]) over the samples
 * (benchmarking/generate_convex_hull.py:88-100, scipy ConvexHull(...).vertices).  The GPU discards everything
 * strictly inside the polygon of 16 directional extremes; the exact hull of the survivors is taken on the host.
 * hull_idx HOST int32 [H1][max_vertices]: sample indices of stage t's vertices, counter-clockwise from the
 * lexicographically smallest point; hull_n HOST int32 [H1].  SYNCHRONISES the stream. */
int gpmpc_stage_hulls(gpmpc_handle* h, const double* traj, int32_t ns, int32_t nx, int32_t H1, int32_t i0, int32_t i1,
                      int32_t max_vertices, int32_t* hull_idx, int32_t* hull_n, void* stream);
/* Host helper (no device work): exact hull of n HOST points xy[n][2], e.g. to merge the per-rank hull vertices
 * of sharded samples.  hull_pos HOST int32 [n] receives positions into xy, counter-clockwise. */
int gpmpc_hull2d(const double* xy, int32_t n, int32_t* hull_pos, int32_t* hull_n);

/* ---- introspection ----------------------------------------------------------------------------- */

int32_t gpmpc_num_hallucinated(const gpmpc_handle* h);  /* points recorded per batch element */
int32_t gpmpc_num_factor_rows(const gpmpc_handle* h);   /* observed hallucinated scalars in the factor */
int32_t gpmpc_num_real_observed(const gpmpc_handle* h); /* m: observed real scalars */
/* Copies the recorded hallucinated set to DEVICE X[B][np][d], Y[B][np][T]
 * (serves model.train_inputs[0] / train_targets, src/visu.py:483-484). */
int gpmpc_export_hallucinated(const gpmpc_handle* h, double* X, double* Y, void* stream);
/* Reads (and optionally clears) the device status word.  SYNCHRONISES the stream. */
int gpmpc_status(gpmpc_handle* h, uint32_t* status, int32_t clear, void* stream);
/* Bytes of persistent device state, and algorithmic HBM bytes / flops of the last hot-path launch
 * (the figures DESIGN.md's roofline uses). */
int64_t gpmpc_state_bytes(const gpmpc_handle* h);
int gpmpc_last_launch_work(const gpmpc_handle* h, double* bytes, double* flops);
/* Number of kernels this library launched since create (bench.py's gpu_launches). */
int64_t gpmpc_launch_count(const gpmpc_handle* h);
/* Per-launch timing of the fused step kernel inside gpmpc_rollout: when on, a CUDA event pair brackets every
 * step-kernel launch on the rollout's stream; gpmpc_rollout_kernel_ms waits for the last one and returns the
 * summed device time and the number of launches of the last rollout (bench.py's roofline.achieved). */
/* Which kernel serves gpmpc_posterior: mma != 0 (default) the FP64 tensor-core kernel (k_posterior_mma: shared rows by
 * the explicit inverse of L_oo, own rows by 8-row sub-panels), mma == 0 the scalar forward-substitution kernel -- the
 * same posterior by independent arithmetic, kept as the reference semantics for the parity tests. */
int gpmpc_set_block_kernels(gpmpc_handle* h, int32_t mma);
int gpmpc_set_timing(gpmpc_handle* h, int32_t on);
int gpmpc_rollout_kernel_ms(gpmpc_handle* h, double* total_ms, int32_t* launches);
const char* gpmpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPMPC_B200_H */
