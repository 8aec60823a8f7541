"""CPU oracle for the GP arithmetic of sampling-gpmpc's hot path (TEST INFRASTRUCTURE ONLY).

This file is a pure-torch float64 CPU restatement of what GPyTorch 1.13 + linear_operator
compute for the reference's call sites

    src/agent.py:630-641      (model_i(x_input); .sample(base_samples=...))
    src/agent.py:365-376      (prepare_dynamics_set, no base samples)
    src/GP_model.py:50-143    (BatchMultitaskGPModelWithDerivatives[_fromParams])
    benchmarking/simulate_true_reachable_set.py:199-209

under the settings the reference always uses (agent.py:630-638): observation_nan_policy
"mask", fast_computations all off (=> exact Cholesky everywhere), cholesky_jitter = Dyn_gp_jitter,
torch.no_grad().  The op order follows SURVEY.md Appendix A (A.1 kernel, A.2 noise, A.3 posterior,
A.4 NaN mask, A.5 outputs, A.6 sampling / psd_safe_cholesky).  It deliberately keeps the
reference's cost structure: a full re-fit (Cholesky of the whole training block) on every model
call, so it can also serve as the "reference-equivalent CPU torch path" timed by bench.py.

PARITY UNPINNED: gpytorch==1.13 (requirements.txt:6) and linear_operator are third-party
dependencies that are absent from /root/reference and not installable here (no network), and the
reference holds no golden vectors for this path (test/partial_gp_updates.py has no assertions).
The arithmetic below is restated from the published algorithm; tests/golden/ pins *this* oracle
(and the reference's own Agent code run on top of it), not GPyTorch itself.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package never does.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

F64 = torch.float64
MIN_VARIANCE_F64 = 1e-10  # gpytorch.settings.min_variance default for double (A.0 / A.5)
CHOLESKY_MAX_TRIES = 3  # gpytorch.settings.cholesky_max_tries default (A.0)


class NotPSDError(RuntimeError):
    """Mirrors linear_operator.utils.errors.NotPSDError (raised after the jitter ladder fails)."""


class NanError(RuntimeError):
    """Mirrors linear_operator.utils.errors.NanError."""


# --------------------------------------------------------------------------------------------
# A.1  prior covariance
# --------------------------------------------------------------------------------------------
def sq_dist(x1: torch.Tensor, x2: torch.Tensor, x1_eq_x2: bool) -> torch.Tensor:
    """gpytorch.kernels.kernel.sq_dist: ||a||^2 + ||b||^2 - 2 a.b on mean-centred inputs,
    diagonal forced to 0 when x1 is x2, clamped at 0 (Appendix A.1)."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    if x1_eq_x2:
        x2, x2_norm, x2_pad = x1, x1_norm, x1_pad
    else:
        x2 = x2 - adjustment
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    a = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    b = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = a.matmul(b.transpose(-2, -1))
    if x1_eq_x2:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min_(0)


def rbf_kernel(x1: torch.Tensor, x2: torch.Tensor, lengthscale: torch.Tensor) -> torch.Tensor:
    """gpytorch RBFKernel.forward (use_grad=False path of GP_model.py:59-60): exp(-d^2/2) on
    lengthscale-scaled inputs.  x*: (..., n, d); lengthscale: (..., 1, d)."""
    x1_ = x1.div(lengthscale)
    x2_ = x2.div(lengthscale)
    eq = x1.shape == x2.shape and bool(torch.equal(x1, x2))
    return sq_dist(x1_, x2_, eq).div(-2).exp()


def rbf_grad_kernel(x1: torch.Tensor, x2: torch.Tensor, lengthscale: torch.Tensor) -> torch.Tensor:
    """gpytorch RBFKernelGrad.forward (GP_model.py:56-57): joint covariance of value and the d
    partial derivatives, returned in the INTERLEAVED multitask order (scalar index =
    point*(d+1) + task), symmetrised when x1 is x2 (Appendix A.1)."""
    batch_shape = x1.shape[:-2]
    nb = len(batch_shape)
    n1, d = x1.shape[-2:]
    n2 = x2.shape[-2]
    K = torch.zeros(*batch_shape, n1 * (d + 1), n2 * (d + 1), dtype=x1.dtype)

    x1_ = x1.div(lengthscale)
    x2_ = x2.div(lengthscale)
    # (x1 - x2) / l^2, laid out (..., d, n2, n1) -> views below
    outer = x1_.reshape(*batch_shape, n1, 1, d) - x2_.reshape(*batch_shape, 1, n2, d)
    outer = outer / lengthscale.unsqueeze(-2)
    outer = torch.transpose(outer, -1, -2).contiguous()  # (..., n1, d, n2)

    eq = x1.shape == x2.shape and bool(torch.equal(x1, x2))
    K_11 = sq_dist(x1_, x2_, eq).div(-2).exp()
    K[..., :n1, :n2] = K_11

    outer1 = outer.reshape(*batch_shape, n1, n2 * d)
    K[..., :n1, n2:] = outer1 * K_11.repeat([*([1] * (nb + 1)), d])

    outer2 = outer.transpose(-1, -3).reshape(*batch_shape, n2, n1 * d)
    outer2 = outer2.transpose(-1, -2)
    K[..., n1:, :n2] = -outer2 * K_11.repeat([*([1] * nb), d, 1])

    outer3 = outer1.repeat([*([1] * nb), d, 1]) * outer2.repeat([*([1] * (nb + 1)), d])
    eye_over_l2 = torch.eye(d, dtype=x1.dtype).expand(*batch_shape, d, d) / lengthscale.pow(2)
    ones = torch.ones(n1, n2, dtype=x1.dtype)
    # Kronecker (eye/l^2) (x) ones(n1,n2)
    kp = (eye_over_l2[..., :, None, :, None] * ones[:, None, :]).reshape(*batch_shape, d * n1, d * n2)
    K[..., n1:, n2:] = (kp - outer3) * K_11.repeat([*([1] * nb), d, d])

    if eq:
        K = 0.5 * (K.transpose(-1, -2) + K)

    pi1 = torch.arange(n1 * (d + 1)).view(d + 1, n1).t().reshape(n1 * (d + 1))
    pi2 = torch.arange(n2 * (d + 1)).view(d + 1, n2).t().reshape(n2 * (d + 1))
    return K[..., pi1, :][..., :, pi2]


# --------------------------------------------------------------------------------------------
# A.6  psd_safe_cholesky
# --------------------------------------------------------------------------------------------
def psd_safe_cholesky(A: torch.Tensor, jitter: float, max_tries: int = CHOLESKY_MAX_TRIES
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """linear_operator.utils.cholesky.psd_safe_cholesky.  Returns (L, level) where level is the
    per-batch-element index of the first successful try (0 = no jitter, i>0 = jitter*10**(i-1))."""
    L, info = torch.linalg.cholesky_ex(A)
    level = torch.zeros(A.shape[:-2], dtype=torch.int32)
    if not torch.any(info):
        return L, level
    if torch.isnan(A).any():
        raise NanError("cholesky_cpu: matrix contains NaN")
    Aprime = A.clone()
    jitter_prev = 0.0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        failing = info > 0
        level = torch.where(failing, torch.full_like(level, i + 1), level)
        diag_add = (failing.to(A.dtype) * (jitter_new - jitter_prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime.diagonal(dim1=-1, dim2=-2).add_(diag_add)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not torch.any(info):
            return L, level
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


# --------------------------------------------------------------------------------------------
# A.2-A.5  exact GP posterior with derivative tasks and NaN-masked observations
# --------------------------------------------------------------------------------------------
class RefPosterior:
    """What the reference reads off ``model_i(x)`` (a MultitaskMultivariateNormal):
    .mean, .variance, .stddev, .sample(base_samples), .confidence_region()  (agent.py:640-706)."""

    def __init__(self, mean: torch.Tensor, covar: torch.Tensor, H: int, T: int, jitter: float):
        self._mean_flat = mean  # (..., q)
        self.covariance_matrix = covar  # (..., q, q)
        self.H, self.T, self.jitter = H, T, jitter
        self.jitter_level: Optional[torch.Tensor] = None

    @property
    def mean(self) -> torch.Tensor:
        return self._mean_flat.reshape(*self._mean_flat.shape[:-1], self.H, self.T)

    @property
    def variance(self) -> torch.Tensor:
        var = self.covariance_matrix.diagonal(dim1=-1, dim2=-2)
        var = var.clamp_min(MIN_VARIANCE_F64)  # A.5
        return var.reshape(*var.shape[:-1], self.H, self.T)

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev.mul(2)
        return self.mean.sub(s2), self.mean.add(s2)

    def sample(self, base_samples: Optional[torch.Tensor] = None) -> torch.Tensor:
        covar = self.covariance_matrix
        q = covar.shape[-1]
        if base_samples is None:
            # zero_mean_mvn_samples: 1x1 -> unclamped sqrt, else Cholesky root; eps = randn(*batch, q, 1)
            if q == 1:
                root = covar.sqrt()
            else:
                root, self.jitter_level = self._root()
            eps = torch.randn(*covar.shape[:-2], q, 1, dtype=covar.dtype)
            res = root.matmul(eps).squeeze(-1) + self._mean_flat
        else:
            if q == 1:  # LinearOperator._cholesky: 1x1 -> clamp_min(0).sqrt()
                root = covar.clamp_min(0.0).sqrt()
                self.jitter_level = torch.zeros(covar.shape[:-2], dtype=torch.int32)
            else:
                root, self.jitter_level = self._root()
            eps = base_samples.reshape(*covar.shape[:-2], q, 1).to(covar.dtype)
            res = root.matmul(eps).squeeze(-1) + self._mean_flat
        return res.reshape(*res.shape[:-1], self.H, self.T)

    def _root(self):
        """linear_operator root_decomposition(method="cholesky"): psd_safe_cholesky, and on ANY RuntimeError (NotPSDError
        after the ladder, NanError) "Using symeig method": evals, evecs = torch.linalg.eigh(dense) for the WHOLE batch,
        evals.clamp_min(0), root = evecs * evals.sqrt().unsqueeze(-2) (LinearOperator._symeig / root_decomposition).
        jitter_level 4 marks that branch (every batch element)."""
        try:
            return psd_safe_cholesky(self.covariance_matrix, self.jitter)
        except RuntimeError:
            evals, evecs = torch.linalg.eigh(self.covariance_matrix)
            root = evecs * evals.clamp_min(0.0).sqrt().unsqueeze(-2)
            return root, torch.full(self.covariance_matrix.shape[:-2], 4, dtype=torch.int32)


class RefExactGP:
    """Batched exact GP = BatchMultitaskGPModelWithDerivatives_fromParams (GP_model.py:94-143) +
    MultitaskGaussianLikelihood(rank=0) (agent.py:235-240) in eval mode.

    train_x (*batch, n, d); train_y (*batch, n, T) with NaN = unobserved;
    lengthscale (*batch, 1, d); outputscale (*batch,); noise (*batch, 1); task_noises (*batch, T).
    """

    def __init__(self, train_x, train_y, lengthscale, outputscale, noise, task_noises,
                 use_grad: bool, jitter: float):
        self.train_x = train_x.to(F64)
        self.train_y = train_y.to(F64)
        self.batch_shape = self.train_x.shape[:-2]
        self.n, self.d = self.train_x.shape[-2:]
        self.T = self.train_y.shape[-1]
        assert self.T == (self.d + 1 if use_grad else 1)
        self.use_grad = use_grad
        self.lengthscale = lengthscale.to(F64).expand(*self.batch_shape, 1, self.d)
        self.outputscale = outputscale.to(F64).expand(*self.batch_shape)
        self.noise = noise.to(F64).expand(*self.batch_shape, 1)
        self.task_noises = task_noises.to(F64).expand(*self.batch_shape, self.T)
        self.jitter = jitter

    def _k(self, x1, x2):
        base = rbf_grad_kernel if self.use_grad else rbf_kernel
        return base(x1, x2, self.lengthscale) * self.outputscale[..., None, None]  # ScaleKernel

    def obs_mask(self) -> torch.Tensor:
        """A.4: an observation slot counts only if it is non-NaN in EVERY batch element."""
        flat = self.train_y.reshape(-1, self.n * self.T)
        return ~torch.isnan(flat).any(dim=0)

    def __call__(self, x: torch.Tensor) -> RefPosterior:
        x = x.to(F64)
        H = x.shape[-2]
        T = self.T
        mask = self.obs_mask()
        # prior blocks (each evaluated on its own slice, like LazyEvaluatedKernelTensor does)
        K_tt = self._k(self.train_x, self.train_x)[..., mask, :][..., :, mask]
        K_st = self._k(x, self.train_x)[..., :, mask]
        K_ss = self._k(x, x)
        # A.2: likelihood noise on the training block only
        sigma = (self.task_noises[..., None, :] + self.noise[..., None, :]).expand(
            *self.batch_shape, self.n, T).reshape(*self.batch_shape, self.n * T)[..., mask]
        A = K_tt + torch.diag_embed(sigma)
        y = self.train_y.reshape(*self.batch_shape, self.n * T)[..., mask]
        # A.3: mean cache (first factorisation) and covariance path (second factorisation)
        L1, _ = psd_safe_cholesky(A, self.jitter)
        alpha = torch.cholesky_solve(y.unsqueeze(-1), L1)
        mean = K_st.matmul(alpha).squeeze(-1)
        L2, _ = psd_safe_cholesky(A, self.jitter)
        rhs = torch.cholesky_solve(K_st.transpose(-1, -2), L2)
        covar = K_ss + K_st.matmul(rhs.mul(-1))
        return RefPosterior(mean, covar, H, T, self.jitter)


def make_gp_from_params(params: dict, train_x, train_y, batch_shape, use_grad: bool) -> RefExactGP:
    """Hyper-parameter broadcast of BatchMultitaskGPModelWithDerivatives_fromParams
    (GP_model.py:121-143), without the softplus round trip (<=2e-15 relative, A.2)."""
    ag = params["agent"]
    ns, g_ny = batch_shape
    noise = torch.tile(torch.tensor([ag["Dyn_gp_noise"]], dtype=F64), dims=(ns, g_ny, 1))
    val = ag["Dyn_gp_task_noises"]["val"]
    if not use_grad:
        val = val[0]
    task_noises = torch.tile(torch.tensor(val, dtype=F64) * ag["Dyn_gp_task_noises"]["multiplier"],
                             dims=(ns, g_ny, 1))
    lengthscale = torch.tile(torch.tensor(ag["Dyn_gp_lengthscale"]["both"], dtype=F64), dims=(ns, 1, 1, 1))
    outputscale = torch.tile(torch.tensor(ag["Dyn_gp_outputscale"]["both"], dtype=F64), dims=(ns, 1))
    return RefExactGP(train_x, train_y, lengthscale, outputscale, noise, task_noises,
                      use_grad=use_grad, jitter=ag["Dyn_gp_jitter"])
