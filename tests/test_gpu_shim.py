"""GPU: the module-level drop-in (sampling_gpmpc_b200.gpytorch_shim, SURVEY.md 8b level B1).

The reference keeps `import gpytorch` and builds a NEW model on [real || hallucinated] data at every SQP iteration
(src/agent.py:216-258, src/GP_model.py:50-143).  `ShimRefAgent` below does exactly that sequence of gpytorch-API calls
against the shim -- likelihood, ExactGP subclass with mean/kernel modules, hyper-parameters set through the property
setters, `model(x)` under the reference's settings contexts, `.sample(base_samples=)` -- while everything above the
GPyTorch line is the oracle's restatement of the Agent (oracle/agent_ref.py, pinned to the fixtures at 1e-13 by
tests/test_oracle_golden.py).  The replay is compared with the golden fixtures the UNMODIFIED reference Agent produced
(tests/golden/*.npz).  Tolerance: |a-b| <= 1e-9 * max(|b|, s), s = outputscale (variances) / sqrt(outputscale) (rest).
"""
import sys

import numpy as np
import pytest
import torch

from oracle.agent_ref import RefAgent
from tests.replay import CASES, load_case, outputscales, replay, scaled_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9
WELL_CONDITIONED = [c for c in CASES if c not in ("pendulum2D_sqp", "car_sqp", "car_residual_sqp")]  # see tests/test_gpu_parity.py


def _model_class(G):
    class Model(G.models.ExactGP):  # the composition of GP_model.py:50-91, written against the gpytorch API
        def __init__(self, tx, ty, lik, batch_shape, d, use_grad):
            super().__init__(tx, ty, lik)
            self.mean_module = (G.means.ConstantMeanGrad if use_grad else G.means.ConstantMean)(batch_shape=batch_shape)
            self.base_kernel = (G.kernels.RBFKernelGrad if use_grad else G.kernels.RBFKernel)(
                ard_num_dims=d, batch_shape=batch_shape)
            self.covar_module = G.kernels.ScaleKernel(self.base_kernel, batch_shape=batch_shape)

        def forward(self, x):
            return G.distributions.MultitaskMultivariateNormal(self.mean_module(x), self.covar_module(x))
    return Model


def _build(G, params, X, Y, batch_shape, use_grad):
    ag = params["agent"]
    ns, g_ny = batch_shape
    lik = G.likelihoods.MultitaskGaussianLikelihood(num_tasks=Y.shape[-1], rank=0,
                                                    noise_constraint=G.constraints.GreaterThan(0.0),
                                                    batch_shape=batch_shape)
    model = _model_class(G)(X, Y, lik, batch_shape, X.shape[-1], use_grad)
    # GP_model.py:121-143
    model.likelihood.noise = torch.tile(torch.tensor([ag["Dyn_gp_noise"]], dtype=torch.float64), dims=(ns, g_ny, 1))
    val = ag["Dyn_gp_task_noises"]["val"] if use_grad else ag["Dyn_gp_task_noises"]["val"][0]
    model.likelihood.task_noises = torch.tile(
        torch.tensor(val, dtype=torch.float64) * ag["Dyn_gp_task_noises"]["multiplier"], dims=(ns, g_ny, 1))
    model.covar_module.base_kernel.lengthscale = torch.tile(
        torch.tensor(ag["Dyn_gp_lengthscale"]["both"], dtype=torch.float64), dims=(ns, 1, 1, 1))
    model.covar_module.outputscale = torch.tile(
        torch.tensor(ag["Dyn_gp_outputscale"]["both"], dtype=torch.float64), dims=(ns, 1))
    model.eval()
    lik.eval()
    return model.cuda()


class _CpuPosterior:
    def __init__(self, post):
        self._p = post
        self.mean, self.variance = post.mean.cpu(), post.variance.cpu()
        self.jitter_level = None

    def sample(self, base_samples=None):
        y = self._p.sample(base_samples=base_samples)
        self.jitter_level = self._p.jitter_level.cpu()
        return y.cpu()


class _CpuModel:
    """Hands the CPU-resident RefAgent CPU tensors; the settings contexts are the reference's (agent.py:630-638)."""

    def __init__(self, G, model, jitter):
        self.G, self.model, self.jitter = G, model, jitter
        self.train_x, self.train_y = model.train_inputs[0], model.train_targets

    def __call__(self, x):
        S = self.G.settings
        with torch.no_grad(), S.observation_nan_policy("mask"), \
                S.fast_computations(covar_root_decomposition=False, log_prob=False, solves=False), \
                S.cholesky_jitter(float_value=self.jitter, double_value=self.jitter, half_value=self.jitter):
            return _CpuPosterior(self.model(x.cuda()))


class ShimRefAgent(RefAgent):
    def __init__(self, G, *a, **k):
        super().__init__(*a, **k)
        self.G = G

    def train_hallucinated_dynGP(self, sqp_iter, use_model_without_derivatives=False):
        if use_model_without_derivatives:
            data_X, data_Y = self.Dyn_gp_X_train_batch, self.Dyn_gp_Y_train_batch[:, :, :, [0]]
        else:
            data_X, data_Y = self.concatenate_real_hallucinated_data()
        model = _build(self.G, self.params, data_X, data_Y, self.batch_shape, not use_model_without_derivatives)
        self.model_i = _CpuModel(self.G, model, self.params["agent"]["Dyn_gp_jitter"])
        if sqp_iter == 0:  # agent.py:261-272
            self.Hallcinated_X_train = torch.empty(self.ns, self.g_ny, 0, self.in_dim_x, dtype=torch.float64)
            self.Hallcinated_Y_train = torch.empty(self.ns, self.g_ny, 0, self.in_dim_y, dtype=torch.float64)


@pytest.mark.parametrize("case", WELL_CONDITIONED)
def test_shim_replay_matches_golden(case):
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    from sampling_gpmpc_b200.envs import make_env_spec
    shim.reset_backends()
    z, params = load_case(case)
    agent = ShimRefAgent(shim.namespace(), params, make_env_spec(params), torch.tensor(z["X_real"]),
                         torch.tensor(z["Y_real"]), epistimic_random_vector=torch.tensor(z["eps"]))
    os_ = outputscales(params)
    s_val = float(np.sqrt(os_.max()))
    worst = {"mean": 0.0, "variance": 0.0, "y_sample": 0.0, "lin": 0.0}

    def check(k, ag, res):
        xscale = max(1.0, float(np.abs(z[f"x_h_{k}"]).max()))
        mean, var = ag.model_i_call.mean.numpy(), ag.model_i_call.variance.numpy()
        for j in range(mean.shape[1]):
            worst["mean"] = max(worst["mean"], scaled_close(mean[:, j], z[f"mean_{k}"][:, j], np.sqrt(os_[j]), RTOL))
            worst["variance"] = max(worst["variance"], scaled_close(var[:, j], z[f"variance_{k}"][:, j], os_[j], RTOL))
        if f"jitter_level_{k}" in z.files:
            assert np.array_equal(ag.model_i_call.jitter_level.numpy(), z[f"jitter_level_{k}"])
            ys = ag.model_i_samples.numpy()
            for j in range(ys.shape[1]):
                worst["y_sample"] = max(worst["y_sample"],
                                        scaled_close(ys[:, j], z[f"y_sample_{k}"][:, j], np.sqrt(os_[j]), RTOL))
        for got, name in zip(res, ("gp_val", "y_grad", "u_grad")):
            worst["lin"] = max(worst["lin"], scaled_close(got, z[f"{name}_{k}"], s_val * xscale, RTOL))
        assert ag.Hallcinated_X_train.shape[2] == int(z[f"n_halluc_{k}"])

    replay(agent, z, params, on_call=check)
    for q, v in worst.items():
        assert v <= 1.0, f"{case}: {q} off by {v:.3g} x tolerance through the gpytorch shim"
    # the engine behind the shim grew incrementally: one handle, never more rows than the data set has scalars
    (be,) = shim._BACKENDS.values()
    assert be.eng.num_real_observed == int(np.sum(~np.isnan(z["Y_real"][0][:, : agent.in_dim_y])))
    shim.reset_backends()


def test_shim_incremental_append_reset_and_nan_mask():
    """Model rebuilt on a growing data set -> only the new points are appended; a point NaN'd for SOME samples is masked
    for all; a shrunken data set resets; results equal a fresh engine fed the same data in one go."""
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    from sampling_gpmpc_b200 import configs
    shim.reset_backends()
    G = shim.namespace()
    params = configs.pendulum1D_sqp(num_dyn_samples=6, n_mpc=1)
    from sampling_gpmpc_b200.envs import make_env_spec
    X, Y = make_env_spec(params).initial_training_data(params)
    ns, g_ny, d, T, H = 6, 1, 2, 3, 4
    bs = torch.Size([ns, g_ny])
    g = torch.Generator().manual_seed(5)
    Xr, Yr = torch.tile(X, (ns, g_ny, 1, 1)), torch.tile(Y, (ns, 1, 1, 1))
    xs = [2.2 + torch.rand(ns, g_ny, H, d, generator=g, dtype=torch.float64) for _ in range(4)]
    ys = [0.01 * torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64) for _ in range(3)]
    ys[1][2, 0, 1, :] = float("nan")  # near-duplicate filtered for one sample only
    eps = torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)

    def model_on(k):
        Xa = torch.cat([Xr] + xs[:k], 2)
        Ya = torch.cat([Yr] + ys[:k], 2)
        return _build(G, params, Xa, Ya, bs, True)

    outs, rows = [], []
    for k in range(4):  # 0, 4, 8, 12 hallucinated points
        m = model_on(k)
        with G.settings.cholesky_jitter(double_value=1e-6):
            p = m(xs[3].cuda())
            y = p.sample(base_samples=eps.cuda())
        lo, hi = p.confidence_region()
        assert torch.equal(hi, p.mean + 2 * p.stddev) and torch.equal(lo, p.mean - 2 * p.stddev)
        outs.append((p.mean.clone(), p.variance.clone(), y.clone()))
        (be,) = shim._BACKENDS.values()
        rows.append(be.eng.num_factor_rows)
    assert rows == [0, 4 * T, 7 * T, 11 * T]  # the NaN'd point holds no factor rows, for any sample
    assert m.train_inputs[0].shape[2] == X.shape[0] + 12 and m.train_targets.shape[2] == X.shape[0] + 12
    eng_id = id(be.eng)
    # shrink (new MPC step): reset, then identical to the first model
    p0 = model_on(0)(xs[3].cuda())
    assert be.eng.num_factor_rows == 0 and id(next(iter(shim._BACKENDS.values())).eng) == eng_id
    assert torch.equal(p0.mean, outs[0][0]) and torch.equal(p0.variance, outs[0][1])
    # one-go build of the largest set == the incremental one (same kernels, block sizes differ: tolerance)
    shim.reset_backends()
    p3 = model_on(3)(xs[3].cuda())
    y3 = p3.sample(base_samples=eps.cuda())
    os_ = float(params["agent"]["Dyn_gp_outputscale"]["both"][0])
    assert scaled_close(p3.mean.cpu(), outs[3][0].cpu(), np.sqrt(os_), RTOL) <= 1.0
    assert scaled_close(p3.variance.cpu(), outs[3][1].cpu(), os_, RTOL) <= 1.0
    # the 12 x 12 joint covariance of 4 nearby test points is numerically singular (lambda_min ~ 1e-12 os): its Cholesky
    # root, hence the draw, amplifies the rounding-level difference between the two factorisations ~1e5 x (the same
    # effect that excludes the ill-conditioned fixtures from free-running replay, tests/test_gpu_parity.py)
    assert scaled_close(y3.cpu(), outs[3][2].cpu(), np.sqrt(os_), 1e-5) <= 1.0
    # a stale posterior object (another model call happened in between) still samples its own distribution
    pa = model_on(3)
    post_a = pa(xs[3].cuda())
    _ = pa(xs[0].cuda())
    ya = post_a.sample(base_samples=eps.cuda())
    assert scaled_close(ya.cpu(), y3.cpu(), np.sqrt(os_), RTOL) <= 1.0
    shim.reset_backends()


def test_shim_partially_observed_points_match_the_oracle():
    """observation_nan_policy('mask') per label SLOT: hallucinated points whose derivative slots are NaN (what
    prepare_dynamics_set does, src/agent.py:402) keep their value row only; a slot NaN for one sample is masked for all.
    The shim (gpmpc_append_masked; model calls then go through the row-by-row kernels) against the CPU oracle on the same
    data, through three model rebuilds (partial points first, whole points after them, then a reset)."""
    from oracle import gp_ref
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    from sampling_gpmpc_b200.envs import make_env_spec
    shim.reset_backends()
    G = shim.namespace()
    params = configs.pendulum1D_sqp(num_dyn_samples=5, n_mpc=1)
    X, Y = make_env_spec(params).initial_training_data(params)
    ns, g_ny, d, T, H = 5, 1, 2, 3, 6
    bs = torch.Size([ns, g_ny])
    g = torch.Generator().manual_seed(9)
    Xr, Yr = torch.tile(X, (ns, g_ny, 1, 1)), torch.tile(Y, (ns, 1, 1, 1))
    xs = [2.2 + torch.rand(ns, g_ny, H, d, generator=g, dtype=torch.float64) for _ in range(3)]
    ys = [0.01 * torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64) for _ in range(2)]
    ys[0][:, :, 1, 1:] = float("nan")     # point 1: value only (both derivative slots dropped)
    ys[0][3, 0, 4, 2] = float("nan")      # point 4: one derivative slot NaN for ONE sample -> dropped for all
    ys[0][:, :, 5, :] = float("nan")      # point 5: not in the factor at all
    os_ = float(params["agent"]["Dyn_gp_outputscale"]["both"][0])
    xq = xs[2]
    for k, want_rows in ((1, 6 * T - 2 - 1 - 3), (2, 6 * T - 6 + 6 * T), (0, 0)):
        Xa, Ya = torch.cat([Xr] + xs[:k], 2), torch.cat([Yr] + ys[:k], 2)
        with G.settings.cholesky_jitter(double_value=1e-6):
            post = _build(G, params, Xa, Ya, bs, True)(xq.cuda())
        ref = gp_ref.make_gp_from_params(params, Xa, Ya, bs, use_grad=True)(xq)
        (be,) = shim._BACKENDS.values()
        assert be.eng.num_factor_rows == want_rows
        assert scaled_close(post.mean.cpu(), ref.mean, np.sqrt(os_), RTOL) <= 1.0
        assert scaled_close(post.variance.cpu(), ref.variance, os_, RTOL) <= 1.0
    shim.reset_backends()


def test_shim_installs_as_gpytorch():
    """`import gpytorch` resolves to the shim after install(); the census of SURVEY.md 8(b) is complete."""
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    saved = {k: v for k, v in sys.modules.items() if k == "gpytorch" or k.startswith("gpytorch.")}
    try:
        shim.install()
        import gpytorch
        from gpytorch.kernels import RBFKernel, ScaleKernel  # agent.py:8-11
        assert gpytorch.models.ExactGP is shim.ExactGP and RBFKernel is shim.RBFKernel and ScaleKernel is shim.ScaleKernel
        for path in ("means.ConstantMean", "means.ConstantMeanGrad", "kernels.RBFKernelGrad",
                     "likelihoods.MultitaskGaussianLikelihood", "constraints.GreaterThan",
                     "distributions.MultitaskMultivariateNormal", "settings.observation_nan_policy",
                     "settings.fast_computations", "settings.cholesky_jitter", "settings.fast_pred_var"):
            mod, name = path.split(".")
            assert hasattr(getattr(gpytorch, mod), name), path
    finally:
        for k in [k for k in sys.modules if k == "gpytorch" or k.startswith("gpytorch.")]:
            del sys.modules[k]
        sys.modules.update(saved)
