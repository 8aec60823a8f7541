"""Host-side mirror of the reference ``Agent``'s hot-path interface (src/agent.py), over the CUDA engine.

Same method names, argument meaning and return layouts as the reference, so ``src/solver.py:84-94`` and the
rollout scripts can drive it unchanged:

    train_hallucinated_dynGP(sqp_iter, use_model_without_derivatives=False)      agent.py:216-272
    get_batch_x_hat(x_h, u_h) / get_batch_x_hat_u_diff(x_h, u_h)                  agent.py:480-527
    dyn_fg_jacobians(xu_hat, sqp_iter) -> gp_val, y_grad, u_grad (numpy)           agent.py:532-564
    get_batch_gp_sensitivities(xu_hat, sqp_iter)                                  agent.py:566-627
    sample_gp(x_input, base_samples)                                              agent.py:629-730
    update_hallucinated_Dyn_dataset(newX, newY)                                   agent.py:164-202
    mpc_iteration(i), random_vector_within_bounds(), epistimic_random_vector      agent.py:76-104,529-530
    Hallcinated_X_train / Hallcinated_Y_train, model_i.train_inputs / train_targets, model_i_call.mean/.variance

What differs is only *how*: the reference throws the GPyTorch model away and re-fits it on
[real || hallucinated] data on every call; here ``train_hallucinated_dynGP`` is bookkeeping, because the
engine keeps each sample's bordered Cholesky factor on the GPU and ``update_hallucinated_Dyn_dataset``
appends rows to it.  All arithmetic happens in libgpmpc_b200.so; torch holds memory and moves bytes.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .engine import GPEngine, make_env_struct
from .envs import EnvSpec, env_spec_from_env_model, make_env_spec

F64 = torch.float64


def gp_hypers_from_params(params: dict, g_ny: int, d: int, use_grad: bool):
    """The yaml -> hyper-parameter broadcast of GP_model.py:121-143, per GP output (the reference tiles the
    same values over every sample).  Returns lengthscale (g_ny,d), outputscale (g_ny,), noise (g_ny,T)."""
    ag = params["agent"]
    ls = np.asarray(ag["Dyn_gp_lengthscale"]["both"], dtype=np.float64).reshape(-1, d)
    ls = np.broadcast_to(ls, (g_ny, d)).copy()
    os_ = np.broadcast_to(np.asarray(ag["Dyn_gp_outputscale"]["both"], dtype=np.float64).reshape(-1), (g_ny,)).copy()
    val = ag["Dyn_gp_task_noises"]["val"]
    val = np.asarray(val if use_grad else [val[0]], dtype=np.float64) * ag["Dyn_gp_task_noises"]["multiplier"]
    noise = np.broadcast_to(val + ag["Dyn_gp_noise"], (g_ny, val.shape[0])).copy()
    return ls, os_, noise


def reachable_set_ball(params: dict, V_k: np.ndarray):
    """The epsilon-ball constraint tightenings of src/utils/reachable_set.py:3-39 (host scalar math; prepare_dynamics_set
    reads ci_list[i] as the acceptance radius of stage i).  Returns (tilde_eps_list, ci_list)."""
    opt, tight = params["optimizer"], params["agent"]["tight"]
    H = opt["H"]
    P = np.array(opt["terminal_tightening"]["P"])
    L = tight["Lipschitz"]
    var_eps = tight["dyn_eps"] + tight["w_bound"]
    B_d_norm = np.sum(np.sqrt(np.diag(P[:3][:3]))) * V_k
    P_inv = np.linalg.inv(P)
    K = np.array(opt["terminal_tightening"]["K"])
    x_t, u_t = np.sqrt(np.diag(P_inv)), np.sqrt(np.diag(K @ P_inv @ K.T))
    tilde_eps_list = [np.concatenate([(x_t * 0).tolist(), (u_t * 0).tolist(), [0]])]
    ci_list = []
    for stage in range(1, H + 1):
        B_eps_k = var_eps * B_d_norm[stage - 1] * np.sum(np.power(L, np.arange(0, stage)))
        tilde_eps_list.append(np.concatenate([(x_t * B_eps_k).tolist(), (u_t * B_eps_k).tolist(), [B_eps_k]]))
        ci_list.append(B_eps_k)
    return tilde_eps_list, ci_list


class _ModelView:
    """What callers read off ``agent.model_i`` (src/visu.py:483-484, src/agent.py:642-643)."""

    def __init__(self, agent: "Agent", version: int, with_hallucinated: bool):
        self._agent, self._version, self._with_h = agent, version, with_hallucinated
        self.batch_shape = agent.batch_shape

    def _data(self):
        ag = self._agent
        if ag._data_version != self._version:
            raise RuntimeError("model_i's training-data snapshot is gone: the factor was updated after "
                               "train_hallucinated_dynGP (call it again before reading train_inputs)")
        X, Y = ag.Dyn_gp_X_train_batch, ag.Dyn_gp_Y_train_batch
        if self._with_h:
            Xh, Yh = ag.engine.export_hallucinated()
            X, Y = torch.cat([X, Xh], 2), torch.cat([Y, Yh], 2)
        return X, Y

    @property
    def train_inputs(self):
        return (self._data()[0],)

    @property
    def train_targets(self):
        return self._data()[1]

    def eval(self):
        return self


class _PosteriorView:
    """``agent.model_i_call``: .mean / .variance / .stddev like the MultitaskMultivariateNormal the reference keeps."""

    def __init__(self, mean, variance, jitter_level=None):
        self.mean, self.variance, self.jitter_level = mean, variance, jitter_level

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev.mul(2)
        return self.mean.sub(s2), self.mean.add(s2)


class Agent:
    def __init__(self, params: dict, env_model=None, *, spec: Optional[EnvSpec] = None,
                 X_real: Optional[torch.Tensor] = None, Y_real: Optional[torch.Tensor] = None,
                 epistimic_random_vector: Optional[torch.Tensor] = None, generate_base_samples: bool = True,
                 rank: int = 0, world_size: int = 1, device: Optional[torch.device] = None):
        self.params = params
        ag = params["agent"]
        self.g_nx, self.g_nu, self.g_ny = ag["g_dim"]["nx"], ag["g_dim"]["nu"], ag["g_dim"]["ny"]
        self.nx, self.nu = ag["dim"]["nx"], ag["dim"]["nu"]
        self.ns_global = ag["num_dyn_samples"]
        # contiguous shard of the sample index; global indices are kept for eps slicing
        self.rank, self.world_size = rank, world_size
        per = -(-self.ns_global // world_size)
        self.s_lo, self.s_hi = min(rank * per, self.ns_global), min((rank + 1) * per, self.ns_global)
        self.ns = self.s_hi - self.s_lo
        self.in_dim_x = self.g_nx + self.g_nu
        self.in_dim_y = 1 if params["env"]["use_model_without_derivatives"] else 1 + self.in_dim_x
        self.batch_shape = torch.Size([self.ns, self.g_ny])
        self.torch_device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.use_cuda = True

        if spec is None:
            spec = env_spec_from_env_model(env_model, params) if env_model is not None else make_env_spec(params)
        self.spec = spec
        self.env_model = env_model
        if X_real is None:
            X_real, Y_real = (env_model.initial_training_data() if env_model is not None
                              else spec.initial_training_data(params))
        self.Dyn_gp_X_train = X_real.to(self.torch_device, F64)
        self.Dyn_gp_Y_train = Y_real.to(self.torch_device, F64)
        if self.in_dim_y == 1:
            self.Dyn_gp_Y_train = self.Dyn_gp_Y_train[:, :, [0]]
        n_real = self.Dyn_gp_X_train.shape[0]

        self.engine = GPEngine(self.ns, self.g_ny, self.in_dim_x, self.in_dim_y, n_real, cap_points=0,
                               device=self.torch_device)
        self.engine.set_condition_on_hallucinated(self.in_dim_y > 1)  # value-only model never conditions (agent.py:221-226)
        ls, os_, noise = gp_hypers_from_params(params, self.g_ny, self.in_dim_x, use_grad=self.in_dim_y > 1)
        self.engine.set_hypers(ls, os_, noise, ag["Dyn_gp_jitter"])
        self.engine.set_real_data(self.Dyn_gp_X_train, self.Dyn_gp_Y_train.contiguous())
        self.engine.reserve(params["optimizer"]["H"] * max(1, params["optimizer"]["SEMPC"]["max_sqp_iter"]))
        self.env_struct = make_env_struct(spec)
        self.outputscale = os_

        self._data_version = 0      # bumps whenever the engine's data set changes
        self._pending_reset = False  # sqp_iter == 0: data set is logically empty, factor still holds the old one
        self._appended_since_train = False
        self.model_i = None
        self.model_i_call = None
        self.mpc_iter = 0
        self.tilde_eps_list = self.ci_list = None
        if "P" in params["optimizer"].get("terminal_tightening", {}) and "Lipschitz" in ag.get("tight", {}):
            self.tilde_eps_list, self.ci_list = reachable_set_ball(params, np.ones(params["optimizer"]["H"] + 1))  # agent.py:71-73
        if epistimic_random_vector is not None:
            self.epistimic_random_vector = epistimic_random_vector[:, :, self.s_lo:self.s_hi].to(self.torch_device, F64)
        elif generate_base_samples:
            self.epistimic_random_vector = self.random_vector_within_bounds()[:, :, self.s_lo:self.s_hi]
        else:
            self.epistimic_random_vector = None

    # ---- views the reference exposes as attributes ---------------------------------------------
    @property
    def Dyn_gp_X_train_batch(self):  # real_data_batch, agent.py:204-214 (a view: nothing is tiled in memory)
        return self.Dyn_gp_X_train.expand(self.ns, self.g_ny, *self.Dyn_gp_X_train.shape)

    @property
    def Dyn_gp_Y_train_batch(self):
        return self.Dyn_gp_Y_train.expand(self.ns, *self.Dyn_gp_Y_train.shape)

    def _hallucinated(self):
        if self._pending_reset:
            return (torch.empty(self.ns, self.g_ny, 0, self.in_dim_x, dtype=F64, device=self.torch_device),
                    torch.empty(self.ns, self.g_ny, 0, self.in_dim_y, dtype=F64, device=self.torch_device))
        return self.engine.export_hallucinated()

    @property
    def Hallcinated_X_train(self):
        return self._hallucinated()[0]

    @property
    def Hallcinated_Y_train(self):
        return self._hallucinated()[1]

    # ---- a2 / f2: agent.py:76-104, the same generator stream without the Python loop (base_samples.py) ------------
    def random_vector_within_bounds(self) -> torch.Tensor:
        from .base_samples import truncated_normal_base_samples
        p = self.params
        dev = self.torch_device if p["common"]["use_cuda"] else torch.device("cpu")  # the generator the reference uses
        out = truncated_normal_base_samples(p["common"]["num_MPC_itrs"], p["optimizer"]["SEMPC"]["max_sqp_iter"],
                                            self.ns_global, self.g_ny, p["optimizer"]["H"], self.in_dim_y,
                                            p["agent"]["Dyn_gp_beta"], dev)
        return out.to(self.torch_device)

    def mpc_iteration(self, i):
        self.mpc_iter = i

    # ---- a4 ------------------------------------------------------------------------------------
    def train_hallucinated_dynGP(self, sqp_iter, use_model_without_derivatives=False):
        """Reference: build a new GPyTorch model on [real || hallucinated] (agent.py:216-258), then, if
        sqp_iter == 0, empty the hallucinated set (agent.py:261-272) -- so this call's model still contains the
        previous MPC step's points.  Here the factor already *is* that model; the reset is deferred to the next
        append so the coming posterior call sees the same training set the reference's model_i would."""
        if use_model_without_derivatives != (self.in_dim_y == 1):
            raise NotImplementedError("use_model_without_derivatives must agree with params['env'] (as in the reference scripts)")
        if self._pending_reset:  # two trains without an append in between: the old model is simply dropped
            self.engine.reset_hallucinated()
            self._data_version += 1
            self._pending_reset = False
        self.engine.set_condition_on_hallucinated(not use_model_without_derivatives)
        self.model_i = _ModelView(self, self._data_version, with_hallucinated=not use_model_without_derivatives)
        self._appended_since_train = False
        if sqp_iter == 0:
            self._pending_reset = True

    # ---- a13: agent.py:480-527.  Built in PINNED HOST memory with numpy views (a few microseconds at the SQP sizes); the
    #      one H2D copy happens inside the C call that consumes it (gpmpc_linearise) -- engine methods that are handed the
    #      tensor directly move it themselves -------------------------------------------------------------------------
    def _xu_buffer(self, H):
        key = (self.ns, self.nx, H, self.nx + self.nu)
        if getattr(self, "_xu_key", None) != key:
            self._xu_host = torch.empty(key, dtype=F64, pin_memory=True)
            self._xu_np = self._xu_host.numpy()
            self._xu_key = key
        return self._xu_host, self._xu_np

    def get_batch_x_hat(self, x_h, u_h):
        H = self.params["optimizer"]["H"]
        host, out = self._xu_buffer(H)
        x_h = np.asarray(x_h, dtype=np.float64).reshape(H, self.ns_global, self.nx)[:, self.s_lo:self.s_hi]
        u_h = np.asarray(u_h, dtype=np.float64).reshape(H, -1)
        out[:, :, :, :self.nx] = x_h.transpose(1, 0, 2)[:, None, :, :]
        out[:, :, :, self.nx:] = u_h[None, None, :, :]
        return host

    def get_batch_x_hat_u_diff(self, x_h, u_h):
        H = self.params["optimizer"]["H"]
        host, out = self._xu_buffer(H)
        x_h = np.asarray(x_h, dtype=np.float64).reshape(H, self.ns_global, self.nx)[:, self.s_lo:self.s_hi]
        u_h = np.asarray(u_h, dtype=np.float64).reshape(H, self.ns_global, self.nu)[:, self.s_lo:self.s_hi]
        out[:, :, :, :self.nx] = x_h.transpose(1, 0, 2)[:, None, :, :]
        out[:, :, :, self.nx:] = u_h.transpose(1, 0, 2)[:, None, :, :]
        return host

    def get_g_xu_hat(self, xu_hat):
        return xu_hat.to(self.torch_device)[:, 0:self.g_ny, :, list(self.spec.g_idx_inputs)].contiguous()

    # ---- a9 ------------------------------------------------------------------------------------
    def sample_gp(self, x_input, base_samples=None):
        ag = self.params["agent"]
        if self._appended_since_train:
            raise NotImplementedError("sample_gp after update_hallucinated_Dyn_dataset without a new "
                                      "train_hallucinated_dynGP: the reference would still use the old model")
        if ag["Dyn_gp_min_data_dist"] >= 0.0:
            return self._sample_gp_min_dist(x_input, base_samples)
        if base_samples is None:  # GPyTorch draws torch.randn(*batch, q, 1) inside .sample()
            base_samples = torch.randn(*x_input.shape[:-1], self.in_dim_y, dtype=F64, device=self.torch_device)
            opts = self.engine.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"], unclamped_sqrt_1x1=True)
        else:
            opts = self.engine.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"])
        mean, var, y, jl = self.engine.posterior(x_input, base_samples, opts)
        self.model_i_call = _PosteriorView(mean, var, jl)
        self.model_i_samples = y
        return y

    def _sample_gp_min_dist(self, x_input, base_samples):
        """Dyn_gp_min_data_dist >= 0 (agent.py:666-698): the nearest fully observed training point's targets
        replace the draw when it is closer than the threshold.  Draw (with the zero-variance rule) without
        truncation, then ONE kernel does the nearest-point overwrite and the truncation (agent.py:701-708)."""
        ag = self.params["agent"]
        if base_samples is None:
            base_samples = torch.randn(*x_input.shape[:-1], self.in_dim_y, dtype=F64, device=self.torch_device)
        opts = self.engine.opts(-1.0, ag["Dyn_gp_variance_is_zero"])
        mean, var, y, jl = self.engine.posterior(x_input, base_samples, opts)
        self.engine.min_dist_overwrite(x_input, mean, var, y, ag["Dyn_gp_min_data_dist"], ag["Dyn_gp_beta"])
        self.model_i_call = _PosteriorView(mean, var, jl)
        self.model_i_samples = y
        return y

    # ---- a11 -----------------------------------------------------------------------------------
    def update_hallucinated_Dyn_dataset(self, newX, newY):
        min_distance = self.params["agent"]["Dyn_gp_min_data_dist"]
        active = None
        if min_distance >= 0.0:
            # agent.py:166-191: NaN the labels of near-duplicates per sample (one kernel), drop a point only if it is
            # filtered for ALL samples of some output; GPyTorch then masks a slot that is NaN for ANY batch
            # element (A.4).  The all / any run over every sample, i.e. over all ranks: one tiny all-reduce.
            newY = newY.contiguous().clone()
            counts = self.engine.filter_new_points(newX, newY, min_distance, use_hallucinated=not self._pending_reset)
            from .rollout import reduce_filter_counts
            flags = reduce_filter_counts(counts, self.ns_global, self.world_size)
            keep = ~flags[0]
            if not keep.all():
                newX, newY = newX[:, :, keep, :], newY[:, :, keep, :]
            active = (~flags[1][keep]).astype(np.uint8)
        if self._pending_reset:
            self.engine.reset_hallucinated()
            self._pending_reset = False
        if newX.shape[2] > 0:
            self.engine.append(newX.contiguous(), newY.contiguous(), active)
        self._data_version += 1
        self._appended_since_train = True

    # ---- a15: the rejection rollout (agent.py:283-443; common.dynamics_rejection) ---------------------------------
    def train_forward_sampling_dynGP(self):
        """Reference: a new GPyTorch model on [real || forward-sampling || hallucinated] data (agent.py:283-329).  Here the
        factor already holds the hallucinated AND the forward-sampling points (prepare_dynamics_set appends them as it
        goes; the order of training points does not change a posterior), so this is bookkeeping like
        train_hallucinated_dynGP."""
        self.model_i = _ModelView(self, self._data_version, with_hallucinated=True)
        self._appended_since_train = False

    def prepare_dynamics_set(self, X_soln, U_soln, X_kp1, base_samples=None):
        """agent.py:331-443: roll every sampled dynamics forward from the MEASURED next state x(k+1) along the solver's input
        sequence, conditioning each sample on its own draws (value task only: the derivative labels are NaN'd, :402), and
        keep the samples whose trajectories stay within c_i of the solver's own prediction (:351-394); the hallucinated data
        of a rejected sample is then replaced by that of a random surviving one (:418-436, numpy's global generator, two
        independent choices for inputs and labels exactly like the reference) and the model is restored (:438-441).
        X_soln (H+1, ns*nx), U_soln (H, nu), X_kp1 (nx, 1).  base_samples: optional list of (ns, g_ny, 1, T) draws, one per
        rollout step, instead of the library-internal torch.randn of `.sample()` (tests).  Returns samples_left (ns,)."""
        if self._pending_reset or self.in_dim_y == 1:
            raise NotImplementedError("prepare_dynamics_set follows a completed SQP solve of the derivative model")
        ag, opt = self.params["agent"], self.params["optimizer"]
        ns, dev, eng = self.ns, self.torch_device, self.engine
        var_eps = (ag["tight"]["dyn_eps"] + ag["tight"]["w_bound"]) * np.sqrt(opt["terminal_tightening"]["P"][1][1])
        # (sharded: X_soln holds every rank's samples; this rank rolls out its own block and the survivor resampling at the end
        # is the one exchange of the population, rollout.resample_rejected)
        X_soln = torch.as_tensor(np.asarray(X_soln), dtype=F64).reshape(-1, self.ns_global, self.nx)[:, self.s_lo:self.s_hi].contiguous().to(dev)
        X_kp1 = torch.as_tensor(np.asarray(X_kp1), dtype=F64).reshape(self.nx, -1).transpose(0, 1).to(dev)  # (1, nx)
        U_soln = torch.as_tensor(np.asarray(U_soln), dtype=F64).to(dev)
        n_stage = X_soln.shape[0]
        diff = X_soln[1] - X_kp1
        samples_left = torch.prod((torch.abs(diff) - var_eps < 0).to(torch.int32), dim=1).to(torch.int32).contiguous()
        xu_hat = torch.cat([X_kp1, U_soln[[1]]], dim=-1).expand(ns, self.nx, 1, self.nx + self.nu).contiguous()
        n_h0 = eng.num_hallucinated
        value_only = np.zeros((1, self.in_dim_y), dtype=np.uint8)
        value_only[0, 0] = 1
        opts = eng.opts()  # the script-level draw: no truncation, no zero-variance rule (agent.py:375-377)
        for i in range(1, n_stage - 1):
            g_xu_hat = self.get_g_xu_hat(xu_hat)
            if base_samples is not None:
                eps = base_samples[i - 1]
                eps = (eps[self.s_lo:self.s_hi] if eps.shape[0] == self.ns_global and self.world_size > 1 else eps).to(dev, F64).contiguous()
            else:
                eps = torch.randn(ns, self.g_ny, self.in_dim_y, 1, dtype=F64, device=dev).reshape(ns, self.g_ny, 1, self.in_dim_y)
            mean, var, y, jl = eng.posterior(g_xu_hat, eps, opts)
            self.model_i_call = _PosteriorView(mean, var, jl)
            last = i == n_stage - 2
            _, xu_next = eng.fs_advance(self.env_struct, xu_hat, y, X_soln[i + 1], float(self.ci_list[i]),
                                        None if last else U_soln[i + 1], samples_left)
            if last:
                break
            eng.append_masked(g_xu_hat, y, value_only)  # FS_*_train_batch: the point with its value label only
            self._data_version += 1
            self.train_forward_sampling_dynGP()
            xu_hat = xu_next
        eng.truncate_hallucinated(n_h0)  # the forward-sampling set is dropped again
        self._data_version += 1
        from .rollout import resample_rejected
        Xh, Yh = eng.export_hallucinated()
        changed, Xh, Yh, active = resample_rejected(samples_left, Xh, Yh, self.ns_global, self.rank, self.world_size)
        if changed:
            # the replaced samples' factors: rebuilt by conditioning on the new data set (the reference re-fits, :438-441)
            eng.reset_hallucinated()
            step = max(1, 512 // self.in_dim_y)
            active = active.reshape(-1, self.in_dim_y)  # (points, T): GPyTorch's any-over-batch slot mask
            for p0 in range(0, Xh.shape[2], step):
                eng.append_masked(Xh[:, :, p0:p0 + step].contiguous(), Yh[:, :, p0:p0 + step].contiguous(), active[p0:p0 + step])
            self._data_version += 1
        self.train_hallucinated_dynGP(sqp_iter=opt["SEMPC"]["max_sqp_iter"])
        eng.raise_on_status()
        return samples_left

    # ---- a10 -----------------------------------------------------------------------------------
    def get_batch_gp_sensitivities(self, xu_hat, sqp_iter):
        ag = self.params["agent"]
        g_xu_hat = self.get_g_xu_hat(xu_hat)
        H = self.params["optimizer"]["H"]
        update = True
        if (ag["true_dyn_as_sample"] or ag["mean_as_dyn_sample"]) and self.ns_global == 1:
            y_sample = torch.zeros(1, self.g_ny, H, self.in_dim_y, dtype=F64, device=self.torch_device)
            update = False
        elif (ag["true_dyn_as_sample"] and ag["mean_as_dyn_sample"]) and self.ns_global == 2:
            y_sample = torch.zeros(2, self.g_ny, H, self.in_dim_y, dtype=F64, device=self.torch_device)[self.s_lo:self.s_hi]
            update = False
        else:
            y_sample = self.sample_gp(g_xu_hat, base_samples=self.epistimic_random_vector[self.mpc_iter][sqp_iter])
        if not update:
            mean, var = self.engine.posterior(g_xu_hat)
            self.model_i_call = _PosteriorView(mean, var)
        idx = 0  # global sample index of the next overwrite (agent.py:607-623)
        if ag["true_dyn_as_sample"]:
            if self.s_lo <= idx < self.s_hi:
                true_dyn = self.spec.prior_data(g_xu_hat[idx - self.s_lo, 0].cpu()).to(self.torch_device)
                if self.in_dim_y == 1:
                    true_dyn = true_dyn[:, :, [0]]
                y_sample[idx - self.s_lo] = true_dyn
            idx += 1
        if ag["mean_as_dyn_sample"]:
            if self.s_lo <= idx < self.s_hi:
                y_sample[idx - self.s_lo] = self.model_i_call.mean[idx - self.s_lo]
            idx += 1
        if update:
            self.update_hallucinated_Dyn_dataset(g_xu_hat, y_sample)
        return y_sample

    # ---- a12 -----------------------------------------------------------------------------------
    def dyn_fg_jacobians_device(self, xu_hat, sqp_iter) -> torch.Tensor:
        """(ns, nx, H, 1+nx+nu) on the device: [f, df/dx, df/du] of every sampled dynamics."""
        y_gp = self.get_batch_gp_sensitivities(xu_hat, sqp_iter)
        return self.engine.assemble(self.env_struct, xu_hat, y_gp)

    def _fused_linearisation_applies(self, xu_hat) -> bool:
        ag = self.params["agent"]
        return (self.in_dim_y > 1 and not ag["true_dyn_as_sample"] and not ag["mean_as_dyn_sample"]
                and ag["Dyn_gp_min_data_dist"] < 0.0 and not self._appended_since_train
                and self.epistimic_random_vector is not None and torch.is_tensor(xu_hat)
                and (xu_hat.is_cuda or xu_hat.is_pinned()) and xu_hat.is_contiguous())

    def dyn_fg_jacobians(self, xu_hat, sqp_iter):
        if self._fused_linearisation_applies(xu_hat):
            # everything of solver.py:84-94 behind the model build in ONE C call: gather of the GP inputs, posterior, draw,
            # post-processing, (deferred reset,) append, assembly, device->host copy
            ag = self.params["agent"]
            opts = self.engine.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"])
            if not hasattr(self, "_lin_bufs"):
                self._lin_bufs = {}
            mean, var, y_gp, jl, _, host = self.engine.linearise(
                self.env_struct, xu_hat, self.epistimic_random_vector[self.mpc_iter][sqp_iter], opts, self._pending_reset,
                self._lin_bufs)
            self._pending_reset = False
            self.model_i_call = _PosteriorView(mean, var, jl)  # views of buffers the next linearisation overwrites, like the
            self.model_i_samples = y_gp                         # reference's model_i_call / model_i_samples are replaced
            self._data_version += 1
            self._appended_since_train = True
            self.engine.raise_on_status()  # waits for the stream: `host` is complete
            h = host.numpy().copy()  # the pinned staging buffer is reused by the next call
            return h[:, :, :, 0:1], h[:, :, :, 1:1 + self.nx], h[:, :, :, 1 + self.nx:1 + self.nx + self.nu]  # views of the copy
        y = self.dyn_fg_jacobians_device(xu_hat, sqp_iter)
        host = torch.empty(y.shape, dtype=F64, pin_memory=True)
        host.copy_(y, non_blocking=True)  # ONE device->host copy (the reference does three, agent.py:555-557)
        # once per SQP iteration, behind the same stream sync: NotPSDError where psd_safe_cholesky would raise it (a failed
        # conditioning block leaves the factor without the new rows -- never carry on silently)
        self.engine.raise_on_status()
        h = host.numpy()
        return h[:, :, :, [0]], h[:, :, :, 1:1 + self.nx], h[:, :, :, 1 + self.nx:1 + self.nx + self.nu]


    def linearise_p_lin(self, x_h, u_h, sqp_iter, tail, use_feedback_K: bool = False, u_per_sample: bool = False) -> np.ndarray:
        """solver.py:84-131 between `ocp_solver.get` and `ocp_solver.set(stage, "p", p_lin)` as ONE device pass and ONE
        device->host copy: the linearisation of every sampled dynamics at the iterate (dyn_fg_jacobians) packed straight into
        the acados stage parameters.  x_h (H, ns*nx), u_h (H, nu) [or (H, ns*nu) with u_per_sample], tail (H, n_tail) =
        hstack(u_h, xg, w, tilde_eps) per stage.  Returns (H, P); row `stage` is that stage's p_lin."""
        xu = self.get_batch_x_hat_u_diff(x_h, u_h) if u_per_sample else self.get_batch_x_hat(x_h, u_h)
        if self._fused_linearisation_applies(xu):
            ag = self.params["agent"]
            opts = self.engine.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"])
            if not hasattr(self, "_lin_bufs"):
                self._lin_bufs = {}
            mean, var, y_gp, jl, lin, _ = self.engine.linearise(
                self.env_struct, xu, self.epistimic_random_vector[self.mpc_iter][sqp_iter], opts, self._pending_reset,
                self._lin_bufs, copy_to_host=False)
            self._pending_reset = False
            self.model_i_call = _PosteriorView(mean, var, jl)
            self.model_i_samples = y_gp
            self._data_version += 1
            self._appended_since_train = True
        else:
            lin = self.dyn_fg_jacobians_device(xu, sqp_iter)
        out = self.pack_p_lin(lin, x_h, tail, use_feedback_K)
        self.engine.raise_on_status()
        return out

    # ---- f1: the acados stage parameters (solver.py:98-131) -------------------------------------
    def pack_p_lin(self, lin: torch.Tensor, x_h, tail, use_feedback_K: bool = False) -> np.ndarray:
        """lin = dyn_fg_jacobians_device(...) (ns,nx,H,1+nx+nu); x_h (H, ns*nx) the SQP iterate; tail (H, n_tail) =
        hstack(u_h, xg, w, tilde_eps) per stage.  Returns the (H, P) array whose row `stage` is exactly the p_lin
        the reference concatenates sample by sample: ONE kernel and ONE device->host copy instead of the
        O(H ns^2) numpy concatenate loop."""
        env = self.env_struct
        if use_feedback_K:
            from .engine import make_env_struct
            K = np.asarray(self.params["optimizer"]["terminal_tightening"]["K"], dtype=np.float64)
            env = make_env_struct(self.spec, K, np.zeros(self.nx))
        x_h = torch.as_tensor(np.asarray(x_h), dtype=F64)
        tail = None if tail is None else torch.as_tensor(np.asarray(tail), dtype=F64)
        out = self.engine.pack_plin(env, lin, x_h, tail, use_feedback_K)
        host = torch.empty(out.shape, dtype=F64, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
