"""GPU: error behaviour of the C ABI as INTEGRATION.md states it (return code < 0 + gpmpc_last_error -> GPEngineError; numerical
failure -> per-element jitter_level 4, NaN draw, status flag -> NotPSDError, the counterpart of GPyTorch's NotPSDError after
three jitter escalations)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(ns=3, g_ny=2, d=2, T=3, n_real=12, jitter=1e-6):
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n_real, d, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((g_ny, n_real, T), float("nan"), dtype=torch.float64)
    Y[:, :, 0] = torch.sin(X).sum(1)
    eng = GPEngine(ns, g_ny, d, T, n_real)
    eng.set_hypers(np.ones((g_ny, d)), np.ones(g_ny), np.full((g_ny, T), 1e-6), jitter)
    eng.set_real_data(X, Y)
    return eng


def test_argument_and_state_errors():
    from sampling_gpmpc_b200.engine import GPEngine, GPEngineError
    with pytest.raises(GPEngineError, match="bad dims"):
        GPEngine(4, 1, 2, 2, 10)  # T must be 1 or d + 1
    with pytest.raises(GPEngineError, match="bad dims"):
        GPEngine(4, 1, 7, 8, 10)  # d <= 6
    e = GPEngine(2, 1, 2, 3, 5)
    with pytest.raises(GPEngineError, match="set_hypers"):
        e.set_real_data(torch.zeros(5, 2, dtype=torch.float64), torch.zeros(1, 5, 3, dtype=torch.float64))
    e.set_hypers(np.ones((1, 2)), np.ones(1), np.full((1, 3), 1e-6), 1e-6)
    with pytest.raises(GPEngineError):
        e.posterior(torch.zeros(2, 1, 1, 2, dtype=torch.float64))  # no real data yet
    with pytest.raises(GPEngineError, match="no observed real data"):
        e.set_real_data(torch.zeros(5, 2, dtype=torch.float64), torch.full((1, 5, 3), float("nan"), dtype=torch.float64))

    eng = _engine()
    x = torch.rand(3, 2, 4, 2, dtype=torch.float64) - 0.5
    eps = torch.zeros(3, 2, 4, 3, dtype=torch.float64)
    with pytest.raises(GPEngineError, match="preceding gpmpc_posterior"):
        eng.sample(eps, 4)
    m, v, y, jl = eng.posterior(x, eps, eng.opts())
    eng.append(x, y)
    with pytest.raises(GPEngineError, match="preceding gpmpc_posterior"):
        eng.sample(eps, 4)  # the factor changed: the cached posterior is stale
    big = torch.rand(3, 2, 171, 2, dtype=torch.float64)
    with pytest.raises(GPEngineError, match="at most 512"):
        eng.append(big, torch.zeros(3, 2, 171, 3, dtype=torch.float64))
    assert eng.status() == 0 and eng.num_factor_rows == 12


@pytest.mark.parametrize("mma", [True, False])
def test_not_positive_definite_is_reported_like_gpytorch(mma):
    """A covariance that stays non-PD through the jitter ladder (here: NaN test inputs of one sample): that element gets
    jitter_level 4 and NaN draws, the others are unaffected, the status word carries the flags and raise_on_status raises
    the counterpart of GPyTorch's NanError (a NaN matrix never takes the eigen-root fallback: eigh raises on it too)."""
    from sampling_gpmpc_b200.engine import NanError, ST_NAN_INPUT, ST_SAMPLE_EIG, ST_SAMPLE_NOT_PD
    eng = _engine()
    eng.set_block_kernels(mma)
    x = torch.rand(3, 2, 4, 2, dtype=torch.float64) - 0.5
    x[1] = float("nan")
    eps = torch.randn(3, 2, 4, 3, dtype=torch.float64).clamp(-2, 2)
    m, v, y, jl = eng.posterior(x, eps, eng.opts(beta=3.0))
    jl = jl.cpu().numpy()
    assert (jl[1] == 4).all() and (jl[[0, 2]] < 4).all()
    assert torch.isnan(y[1]).all() and torch.isfinite(y[[0, 2]]).all()
    assert eng.status() & ST_SAMPLE_NOT_PD and eng.status() & ST_NAN_INPUT and not eng.status() & ST_SAMPLE_EIG
    with pytest.raises(NanError):
        eng.raise_on_status()
    assert eng.status() == 0  # cleared by raise_on_status


def test_shim_sample_raises_on_nan_like_gpytorch():
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    from sampling_gpmpc_b200.engine import NanError
    shim.reset_backends()
    G = shim.namespace()
    ns, g_ny, n, d, T = 2, 1, 9, 2, 3
    g = torch.Generator().manual_seed(1)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.full((g_ny, n, T), float("nan"), dtype=torch.float64)
    Y[:, :, 0] = X.sum(1)
    bs = torch.Size([ns, g_ny])

    class Model(G.models.ExactGP):
        def __init__(self, tx, ty, lik):
            super().__init__(tx, ty, lik)
            self.mean_module = G.means.ConstantMeanGrad(batch_shape=bs)
            self.base_kernel = G.kernels.RBFKernelGrad(ard_num_dims=d, batch_shape=bs)
            self.covar_module = G.kernels.ScaleKernel(self.base_kernel, batch_shape=bs)

    lik = G.likelihoods.MultitaskGaussianLikelihood(num_tasks=T, rank=0, noise_constraint=G.constraints.GreaterThan(0.0),
                                                    batch_shape=bs)
    model = Model(torch.tile(X, (ns, g_ny, 1, 1)), torch.tile(Y, (ns, 1, 1, 1)), lik)
    model.likelihood.noise = torch.full((ns, g_ny, 1), 1e-6, dtype=torch.float64)
    model.likelihood.task_noises = torch.full((ns, g_ny, T), 1e-6, dtype=torch.float64)
    model.covar_module.base_kernel.lengthscale = torch.ones(ns, g_ny, 1, d, dtype=torch.float64)
    model.covar_module.outputscale = torch.ones(ns, g_ny, dtype=torch.float64)
    xq = torch.rand(ns, g_ny, 3, d, generator=g, dtype=torch.float64)
    xq[0] = float("nan")
    post = model(xq.cuda())
    with pytest.raises(NanError):
        post.sample(base_samples=torch.zeros(ns, g_ny, 3, T, dtype=torch.float64))
    shim.reset_backends()


@pytest.mark.parametrize("n_real", [40, 800])
def test_real_data_factorisation_failure_is_flagged(n_real):
    """K0's jitter ladder when the real-data block cannot be factorised (a NaN input): all four attempts fail, the status
    word carries TRAIN_NOT_PD and nothing hangs -- on the one-CTA kernel (n_real = 40) and on the cooperative multi-CTA one
    (n_real = 800: every CTA must take the same decision at every grid barrier)."""
    from sampling_gpmpc_b200.engine import GPEngine, ST_TRAIN_NOT_PD
    g = torch.Generator().manual_seed(3)
    X = torch.rand(n_real, 2, generator=g, dtype=torch.float64)
    X[n_real // 2, 0] = float("nan")
    Y = torch.zeros(2, n_real, 1, dtype=torch.float64)
    eng = GPEngine(2, 2, 2, 1, n_real)
    eng.set_hypers(np.ones((2, 2)), np.ones(2), np.full((2, 1), 1e-6), 1e-6)
    eng.set_real_data(X, Y)
    assert eng.status() & ST_TRAIN_NOT_PD
    # and a clean factorisation right after it on the same handle
    X[n_real // 2, 0] = 0.5
    eng.status(clear=True)
    eng.set_real_data(X, Y)
    assert eng.status() & ST_TRAIN_NOT_PD == 0
    m, v = eng.posterior(torch.rand(2, 2, 3, 2, dtype=torch.float64))
    assert torch.isfinite(m).all() and (v > 0).all()
