"""Oracle-backed stand-in for the `gpytorch` symbols the reference imports (TEST INFRASTRUCTURE ONLY).

gpytorch 1.13 is not installable in this environment, so the reference's own ``src/agent.py`` /
``src/GP_model.py`` cannot be imported as they are.  ``install()`` registers a minimal module tree
named ``gpytorch`` in ``sys.modules`` exposing exactly the symbols listed in SURVEY.md §8(b)
(census: agent.py:4,8-11,235-248,306-317,365-376,595-605,630-641; GP_model.py:18-24,54-91,121-143),
whose arithmetic is ``oracle/gp_ref.py``.  With it, the UNMODIFIED reference Agent runs on CPU; the
generating script ``tests/golden/make_golden.py`` uses that to freeze fixtures, so the Agent-level
logic in the fixtures (post-processing, dataset bookkeeping, Jacobian assembly) is the reference's
own code.  The GP arithmetic underneath is still the restatement: parity unpinned for GPyTorch itself.

Never imported by the product package.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch

from . import gp_ref


class _Module:
    """Just enough of torch.nn.Module for the reference's call pattern (.eval()/.cuda()/.train())."""

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def cuda(self, *a, **k):
        return self

    def cpu(self):
        return self

    def to(self, *a, **k):
        return self


class GreaterThan:
    def __init__(self, lower_bound, *a, **k):
        self.lower_bound = lower_bound


class ConstantMean(_Module):
    def __init__(self, batch_shape=torch.Size([]), **k):
        self.batch_shape = batch_shape

    def __call__(self, x):
        return torch.zeros(*x.shape[:-1], dtype=x.dtype)


class ConstantMeanGrad(_Module):
    def __init__(self, batch_shape=torch.Size([]), **k):
        self.batch_shape = batch_shape

    def __call__(self, x):
        return torch.zeros(*x.shape[:-1], x.shape[-1] + 1, dtype=x.dtype)


class RBFKernel(_Module):
    use_grad = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), **k):
        self.ard_num_dims = ard_num_dims
        self.batch_shape = batch_shape
        d = ard_num_dims or 1
        self.lengthscale = torch.full((*batch_shape, 1, d), 0.6931471805599453, dtype=torch.float64)


class RBFKernelGrad(RBFKernel):
    use_grad = True


class ScaleKernel(_Module):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **k):
        self.base_kernel = base_kernel
        self.batch_shape = batch_shape
        self.outputscale = torch.full(tuple(batch_shape), 0.6931471805599453, dtype=torch.float64)

    def __call__(self, x):  # the prior is never used on its own by the reference
        return (self, x)


class MultitaskGaussianLikelihood(_Module):
    def __init__(self, num_tasks, rank=0, noise_constraint=None, batch_shape=torch.Size([]), **k):
        assert rank == 0, "the reference only uses rank=0 (agent.py:237)"
        self.num_tasks = num_tasks
        self.batch_shape = batch_shape
        self.noise = torch.full((*batch_shape, 1), 0.6931471805599453, dtype=torch.float64)
        self.task_noises = torch.full((*batch_shape, num_tasks), 0.6931471805599453, dtype=torch.float64)


class MultitaskMultivariateNormal:
    """Only ever constructed inside the reference's ``forward`` (GP_model.py:91); the posterior the
    reference reads is the RefPosterior returned by ExactGP.__call__."""

    def __init__(self, mean, covar):
        self.mean_prior = mean
        self.covar_prior = covar


_SETTINGS = {"jitter": 1e-8}  # gpytorch.settings.cholesky_jitter default for float64


class ExactGP(_Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def __call__(self, x):
        base = self.covar_module.base_kernel
        gp = gp_ref.RefExactGP(
            self.train_inputs[0].detach().cpu(), self.train_targets.detach().cpu(),
            torch.as_tensor(base.lengthscale).detach().cpu(),
            torch.as_tensor(self.covar_module.outputscale).detach().cpu(),
            torch.as_tensor(self.likelihood.noise).detach().cpu(),
            torch.as_tensor(self.likelihood.task_noises).detach().cpu(),
            use_grad=base.use_grad, jitter=_SETTINGS["jitter"])
        return gp(x.detach().cpu())


@contextlib.contextmanager
def _noop(*a, **k):
    yield


@contextlib.contextmanager
def _cholesky_jitter(float_value=None, double_value=None, half_value=None, *a, **k):
    prev = _SETTINGS["jitter"]
    if double_value is not None:
        _SETTINGS["jitter"] = float(double_value)
    try:
        yield
    finally:
        _SETTINGS["jitter"] = prev


def install(stub_matplotlib: bool = True) -> None:
    """Register the stand-in as ``gpytorch`` (and a no-op ``matplotlib.pyplot``, which agent.py:13
    imports but the hot path never calls)."""
    g = types.ModuleType("gpytorch")
    sub = {}
    for name in ("models", "means", "kernels", "likelihoods", "constraints", "distributions", "settings"):
        sub[name] = types.ModuleType(f"gpytorch.{name}")
        setattr(g, name, sub[name])
        sys.modules[f"gpytorch.{name}"] = sub[name]
    sub["models"].ExactGP = ExactGP
    sub["means"].ConstantMean = ConstantMean
    sub["means"].ConstantMeanGrad = ConstantMeanGrad
    sub["kernels"].RBFKernel = RBFKernel
    sub["kernels"].RBFKernelGrad = RBFKernelGrad
    sub["kernels"].ScaleKernel = ScaleKernel
    sub["likelihoods"].MultitaskGaussianLikelihood = MultitaskGaussianLikelihood
    sub["constraints"].GreaterThan = GreaterThan
    sub["distributions"].MultitaskMultivariateNormal = MultitaskMultivariateNormal
    sub["settings"].observation_nan_policy = _noop
    sub["settings"].fast_computations = _noop
    sub["settings"].fast_pred_var = _noop
    sub["settings"].cholesky_jitter = _cholesky_jitter
    sys.modules["gpytorch"] = g
    if stub_matplotlib and "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            p = types.ModuleType("matplotlib.pyplot")
            p.rcParams = {}
            m.pyplot = p
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = p
