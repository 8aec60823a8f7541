"""Module-level drop-in (level B1): the `gpytorch` symbols the reference imports, backed by
libgpmpc_b200.so, so that the UNMODIFIED `src/agent.py` / `src/GP_model.py` / benchmarking scripts of
sampling-gpmpc evaluate their GP on the B200 through the C ABI.

    import sampling_gpmpc_b200.gpytorch_shim as shim; shim.install()      # before `import src.agent`
    # or, without touching any file:  PYTHONPATH=<repo>/sampling_gpmpc_b200/shim python main.py -i 1 ...

Census of what the reference touches (agent.py:4,8-11,235-248,306-317,365-376,595-605,630-641; GP_model.py:18-24,
54-91,121-143; solver.py:245-257; visu.py:483-484; simulate_true_reachable_set.py:199-209):
`models.ExactGP`, `means.ConstantMean[Grad]`, `kernels.RBFKernel[Grad] / ScaleKernel`,
`likelihoods.MultitaskGaussianLikelihood(rank=0)`, `constraints.GreaterThan`,
`distributions.MultitaskMultivariateNormal`, `settings.{observation_nan_policy, fast_computations, fast_pred_var,
cholesky_jitter}`.  Everything else of GPyTorch is absent on purpose (training, priors, other kernels).

How the reference's call pattern maps onto the persistent factor: the reference builds a NEW model on
[real || hallucinated] data, tiled over (ns, g_ny), at every SQP iteration (agent.py:216-258).  `ExactGP.__call__`
here finds the block of training points that is identical for every sample (the real data: one shared factor, K0),
compares the rest with what the engine already holds and appends only the new points (k_append) -- or resets when
the hallucinated set shrank (agent.py:261-272).  The posterior object returned mirrors what the reference reads from
a MultitaskMultivariateNormal: `.mean`, `.variance`, `.stddev`, `.confidence_region()`, `.sample(base_samples=)`.
There is no CPU path: without a CUDA device / the built library the first model call raises.
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .engine import GPEngine, NotPSDError

F64 = torch.float64
_SOFTPLUS_0 = 0.6931471805599453  # value of an untouched GPyTorch raw parameter (softplus(0))
_SETTINGS = {"jitter": 1e-8, "nan_policy": "ignore"}  # gpytorch.settings.cholesky_jitter default for float64


class _Module:
    """The slice of torch.nn.Module the reference uses on these objects."""

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

    def __call__(self, x):  # prior mean 0: the constant is never set or trained by the reference
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ConstantMeanGrad(_Module):
    def __init__(self, batch_shape=torch.Size([]), **k):
        self.batch_shape = batch_shape

    def __call__(self, x):
        return torch.zeros(*x.shape[:-1], x.shape[-1] + 1, dtype=x.dtype, device=x.device)


class RBFKernel(_Module):
    use_grad = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), **k):
        self.ard_num_dims = ard_num_dims
        self.batch_shape = batch_shape
        self.lengthscale = torch.full((*batch_shape, 1, ard_num_dims or 1), _SOFTPLUS_0, dtype=F64)


class RBFKernelGrad(RBFKernel):
    use_grad = True


class ScaleKernel(_Module):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **k):
        self.base_kernel = base_kernel
        self.batch_shape = batch_shape
        self.outputscale = torch.full(tuple(batch_shape), _SOFTPLUS_0, dtype=F64)

    def __call__(self, x):  # only reached through the reference's forward(), which this shim never needs
        return (self, x)


class MultitaskGaussianLikelihood(_Module):
    def __init__(self, num_tasks, rank=0, noise_constraint=None, batch_shape=torch.Size([]), **k):
        if rank != 0:
            raise NotImplementedError("only rank=0 task noise (agent.py:237) is implemented")
        self.num_tasks = num_tasks
        self.batch_shape = batch_shape
        self.noise = torch.full((*batch_shape, 1), _SOFTPLUS_0, dtype=F64)
        self.task_noises = torch.full((*batch_shape, num_tasks), _SOFTPLUS_0, dtype=F64)

    def __call__(self, *a, **k):
        raise NotImplementedError("likelihood(model(x)) is not on the hot path: the reference samples model(x) "
                                  "(agent.py:375,605,640)")


class MultitaskMultivariateNormal:
    """As constructed inside the reference's forward() (GP_model.py:91): a prior placeholder.  The object the
    reference reads its posterior from is the `Posterior` returned by ExactGP.__call__."""

    def __init__(self, mean, covar):
        self.mean_prior = mean
        self.covar_prior = covar


# ---- the engine behind the models of one process -------------------------------------------------------------
class _Backend:
    """One GPEngine per (ns, g_ny, d, T, n_real, device) plus what it currently holds."""

    def __init__(self, eng: GPEngine, Xs: torch.Tensor, Ys: torch.Tensor, hypers):
        self.eng, self.Xs, self.Ys, self.hypers = eng, Xs, Ys, hypers
        self.Xh: Optional[torch.Tensor] = None  # (ns, g_ny, nh, d) hallucinated points as appended
        self.Yh: Optional[torch.Tensor] = None
        self.version = 0       # bumped whenever the engine's training set changes
        self.post_token = 0    # identifies the posterior cache currently inside the engine

    @property
    def nh(self) -> int:
        return 0 if self.Xh is None else self.Xh.shape[2]


_BACKENDS: Dict[Tuple, _Backend] = {}


def reset_backends() -> None:
    """Drop every engine (frees the GPU state)."""
    _BACKENDS.clear()


def _nan_equal(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return (a == b) | (a.isnan() & b.isnan())


def _uniform_over_samples(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.shape[0] > 1 and not bool((t == t[:1]).all()):
        raise NotImplementedError(f"{what} differs between dynamics samples; the reference tiles one value "
                                  "(GP_model.py:121-143) and the shared real-data factor relies on it")
    return t[0]


class Posterior:
    """What the reference reads off `model(x)`: agent.py:640-706, simulate_true_reachable_set.py:208-236."""

    def __init__(self, model: "ExactGP", backend: _Backend, x: torch.Tensor, mean, var, token: int, out_device=None):
        self._model, self._be, self._x, self._token = model, backend, x, token
        # results live where the caller's test inputs live (GPyTorch computes on the inputs' device): with
        # common.use_cuda False the reference's own post-processing (agent.py:646-708) then keeps working on CPU tensors
        self._out = out_device if out_device is not None else mean.device
        self.mean, self._var = mean.to(self._out), var.to(self._out)
        self.jitter_level: Optional[torch.Tensor] = None

    @property
    def variance(self) -> torch.Tensor:  # clamped at gpytorch.settings.min_variance inside the kernel (A.5)
        return self._var

    @property
    def stddev(self) -> torch.Tensor:
        return self._var.sqrt()

    def confidence_region(self):
        s2 = self.stddev.mul(2)
        return self.mean.sub(s2), self.mean.add(s2)

    def sample(self, sample_shape=torch.Size(), base_samples: Optional[torch.Tensor] = None) -> torch.Tensor:
        if len(tuple(sample_shape)) != 0:
            raise NotImplementedError("sample_shape other than () is not used by the reference")
        be, eng = self._be, self._be.eng
        ns, g_ny, H, T = self.mean.shape
        if be.post_token != self._token or be.version != self._model._synced_version:
            # another model call or an append replaced the engine's cached posterior: rebuild it
            self._model._sync()
            eng.posterior(self._x)
            be.post_token += 1
            self._token = be.post_token
        if base_samples is None:  # GPyTorch draws randn(*batch, q, 1); a 1x1 covariance uses the unclamped sqrt (A.6)
            eps = torch.randn(ns, g_ny, H * T, 1, dtype=F64, device=self._out).reshape(ns, g_ny, H, T)
            opts = eng.opts(unclamped_sqrt_1x1=True)
        else:
            eps = base_samples.to(eng.device, F64).reshape(ns, g_ny, H, T)
            opts = eng.opts()
        y, jl = eng.sample(eps, H, opts)
        self.jitter_level = jl.to(self._out)
        eng.raise_on_status()  # NanError / NotPSDError where GPyTorch raises, a warning where it falls back to the eigen root
        return y.to(self._out)


class ExactGP(_Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood
        self._backend: Optional[_Backend] = None
        self._synced_version = -1

    # -- hyper-parameters as the reference sets them after construction (GP_model.py:121-143) ------------------
    def _hypers(self, ns: int, g_ny: int, d: int, T: int):
        base = self.covar_module.base_kernel
        raw = (base.lengthscale, self.covar_module.outputscale, self.likelihood.noise, self.likelihood.task_noises)
        if any(torch.is_tensor(t) and t.is_cuda for t in raw):
            # one device->host copy for all four (they live on the GPU when the reference has set_default_device('cuda'))
            flat = torch.cat([torch.as_tensor(t, dtype=F64).detach().reshape(-1).to(raw[0].device) for t in raw]).cpu()
            parts = torch.split(flat, [int(torch.as_tensor(t).numel()) for t in raw])
            raw = tuple(p.reshape(torch.as_tensor(t).shape) for p, t in zip(parts, raw))
        base_ls, cov_os, lik_noise, lik_task = raw
        ls = torch.as_tensor(base_ls, dtype=F64).detach().cpu()
        ls = ls.reshape(ns, g_ny, d) if ls.numel() == ns * g_ny * d else ls.reshape(1, g_ny, d)
        os_ = torch.as_tensor(cov_os, dtype=F64).detach().cpu()
        os_ = os_.reshape(ns, g_ny) if os_.numel() == ns * g_ny else os_.reshape(1, g_ny)
        nz = torch.as_tensor(lik_noise, dtype=F64).detach().cpu()
        nz = nz.reshape(ns, g_ny, 1) if nz.numel() == ns * g_ny else nz.reshape(1, g_ny, 1)
        tn = torch.as_tensor(lik_task, dtype=F64).detach().cpu()
        tn = tn.reshape(ns, g_ny, T) if tn.numel() == ns * g_ny * T else tn.reshape(1, g_ny, T)
        ls = _uniform_over_samples(ls, "lengthscale")
        os_ = _uniform_over_samples(os_, "outputscale")
        noise = _uniform_over_samples(tn + nz, "noise")  # diagonal of Sigma: task_noises[t] + noise (A.2)
        return ls.numpy().copy(), os_.numpy().copy(), noise.numpy().copy(), float(_SETTINGS["jitter"])

    def _sync(self) -> _Backend:
        """Bring the engine's training set to this model's (train_inputs, train_targets).

        The reference re-builds its model from freshly concatenated tensors at every SQP iteration (agent.py:216-258), so the
        only way to know what changed is to compare -- on the device, with ONE host round trip per model build: every check
        below lands in one small flag tensor that is read back once (the first build of a data set needs a second one)."""
        be = self._backend
        if be is not None and be.version == self._synced_version:
            return be
        if not torch.cuda.is_available():
            raise RuntimeError("sampling_gpmpc_b200.gpytorch_shim needs a CUDA device (B200): there is no CPU fallback")
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        X = self.train_inputs[0].to(dev, F64)
        Y = self.train_targets.to(dev, F64)
        if X.dim() != 4 or Y.dim() != 4:
            raise NotImplementedError("expected batched data (ns, g_ny, n, d) / (ns, g_ny, n, T) as agent.py builds it")
        ns, g_ny, n, d = X.shape
        T = Y.shape[3]
        use_grad = bool(self.covar_module.base_kernel.use_grad)
        if T != (d + 1 if use_grad else 1):
            raise NotImplementedError(f"{T} tasks for input dim {d} (use_grad={use_grad})")
        hyp = self._hypers(ns, g_ny, d, T)
        key_base = (ns, g_ny, d, T, dev.index)

        def shared_prefix():
            # leading block of points that is the same for every sample and output = the real data (agent.py:204-214)
            same = (X == X[:1, :1]).all(3).all(1).all(0) & _nan_equal(Y, Y[:1]).all(3).all(1).all(0)
            return int((~same).to(torch.int32).cumsum(0).eq(0).sum())

        # fast path: ONE backend of this shape exists and its real block is still the head of the data
        cands = [k for k in _BACKENDS if k[:5] == key_base]
        be = _BACKENDS[cands[0]] if len(cands) == 1 and cands[0][5] <= n else None
        flags = None
        if be is not None:
            n_real, old, nh = be.Xs.shape[0], be.nh, n - be.Xs.shape[0]
            Xh, Yh = X[:, :, n_real:], Y[:, :, n_real:]
            checks = [(X[:, :, :n_real] == be.Xs).all(), _nan_equal(Y[:, :, :n_real], be.Ys.unsqueeze(0)).all()]
            prefix_kept = old <= nh
            if prefix_kept and old:
                checks += [(Xh[:, :, :old] == be.Xh).all(), _nan_equal(Yh[:, :, :old], be.Yh).all()]
            n_new = nh - old if prefix_kept else nh
            all_nan = Yh.isnan().any(1).any(0).reshape(-1)  # NaN flag of every hallucinated label slot (nh * T)
            flags = torch.cat([torch.stack(checks).to(torch.uint8), all_nan.to(torch.uint8)]).cpu().numpy()  # the one sync
            if not (flags[0] and flags[1]):
                be = None  # different real data: rebuild below
        if be is None:
            n_same = shared_prefix()
            if n_same < 1:
                raise NotImplementedError("no training block shared by all samples (the reference always has real data)")
            for k in cands:
                del _BACKENDS[k]  # one real data set per shape at a time
            eng = GPEngine(ns, g_ny, d, T, n_same, device=dev)
            eng.set_hypers(*hyp)
            Xs, Ys = X[0, 0, :n_same].contiguous(), Y[0, :, :n_same].contiguous()
            eng.set_real_data(Xs, Ys)
            be = _BACKENDS[key_base + (n_same,)] = _Backend(eng, Xs, Ys, hyp)
            n_real, old, nh = n_same, 0, n - n_same
            Xh, Yh = X[:, :, n_real:], Y[:, :, n_real:]
            keep, n_new = True, nh
            new_nan = Yh.isnan().any(1).any(0).reshape(-1).to(torch.uint8).cpu().numpy() if nh else np.zeros(0, np.uint8)
        else:
            keep = prefix_kept and (old == 0 or bool(flags[2] and flags[3]))
            all_nan = flags[len(checks):]
            if not keep:  # the hallucinated set shrank or changed: start it over (agent.py:261-272)
                n_new = nh
            new_nan = all_nan[(nh - n_new) * T:]
            if any(not np.array_equal(a, b) for a, b in zip(hyp[:3], be.hypers[:3])) or hyp[3] != be.hypers[3]:
                be.eng.set_hypers(*hyp)
                be.eng.set_real_data(be.Xs, be.Ys)  # new hypers: new shared factor, the hallucinated rows go with it
                be.hypers, be.Xh, be.Yh = hyp, None, None
                be.version += 1
                keep, old, n_new = True, 0, nh
                new_nan = Yh.isnan().any(1).any(0).reshape(-1).to(torch.uint8).cpu().numpy() if nh else np.zeros(0, np.uint8)
        eng = be.eng
        eng.set_condition_on_hallucinated(True)
        if not keep:
            eng.reset_hallucinated()
            be.Xh = be.Yh = None
            be.version += 1
            old = 0
        if n_new > 0:
            newX, newY = Xh[:, :, old:].contiguous(), Yh[:, :, old:].contiguous()
            # observation_nan_policy('mask'): a label slot is dropped for every element if it is NaN in any;
            # slots are (point, task): a point may keep its value and lose its derivative slots (agent.py:402)
            active = (1 - np.asarray(new_nan, dtype=np.uint8)).reshape(n_new, T)
            step = max(1, 512 // T)
            for p0 in range(0, n_new, step):
                p1 = min(n_new, p0 + step)
                eng.append_masked(newX[:, :, p0:p1].contiguous(), newY[:, :, p0:p1].contiguous(), active[p0:p1])
            be.Xh = newX if be.Xh is None else torch.cat([be.Xh, newX], 2)
            be.Yh = newY if be.Yh is None else torch.cat([be.Yh, newY], 2)
            be.version += 1
        self._backend, self._synced_version = be, be.version
        return be

    def __call__(self, x: torch.Tensor) -> Posterior:
        be = self._sync()
        out_device = x.device
        x = x.to(be.eng.device, F64)
        mean, var = be.eng.posterior(x)
        be.post_token += 1
        return Posterior(self, be, x, mean, var, be.post_token, out_device)


# ---- gpytorch.settings -------------------------------------------------------------------------------------------
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


def namespace() -> types.SimpleNamespace:
    """The shim's symbols grouped like the gpytorch package (what `install()` registers)."""
    ns = types.SimpleNamespace
    return ns(models=ns(ExactGP=ExactGP),
              means=ns(ConstantMean=ConstantMean, ConstantMeanGrad=ConstantMeanGrad),
              kernels=ns(RBFKernel=RBFKernel, RBFKernelGrad=RBFKernelGrad, ScaleKernel=ScaleKernel),
              likelihoods=ns(MultitaskGaussianLikelihood=MultitaskGaussianLikelihood),
              constraints=ns(GreaterThan=GreaterThan),
              distributions=ns(MultitaskMultivariateNormal=MultitaskMultivariateNormal),
              settings=ns(observation_nan_policy=_noop, fast_computations=_noop, fast_pred_var=_noop,
                          cholesky_jitter=_cholesky_jitter))


def install() -> None:
    """Register the shim as `gpytorch` (+ submodules) in sys.modules."""
    src = namespace()
    g = types.ModuleType("gpytorch")
    g.__doc__ = "sampling_gpmpc_b200.gpytorch_shim (B200-backed subset of gpytorch used by sampling-gpmpc)"
    for name in ("models", "means", "kernels", "likelihoods", "constraints", "distributions", "settings"):
        sub = types.ModuleType(f"gpytorch.{name}")
        for k, v in vars(getattr(src, name)).items():
            setattr(sub, k, v)
        setattr(g, name, sub)
        sys.modules[f"gpytorch.{name}"] = sub
    g.NotPSDError = NotPSDError
    sys.modules["gpytorch"] = g
