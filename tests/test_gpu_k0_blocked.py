"""GPU: K0 for training sets in the thousands -- the blocked factorisation (k_factor_real_blocked: 64-column panels, DMMA trailing
update) and the strip-wise inverse (k_invert_real_mma) of csrc/gpmpc_k0.cuh -- against (i) the per-pivot kernels they replace
from m = 768 on (GPMPC_K0_BLOCKED_MIN_M switches between them) and (ii) dense fp64 linear algebra in torch, through the
posterior they feed (fused step kernels read inv(L_oo), the scalar block kernels read L_oo itself)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _data(n_real, d, T, g_ny, grad_obs, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_real, d, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((g_ny, n_real, T), float("nan"), dtype=torch.float64)
    for j in range(g_ny):
        Y[j, :, 0] = torch.sin(X * (1 + 0.2 * j)).sum(1)
        if grad_obs and T > 1:
            Y[j, :, 1:] = (1 + 0.2 * j) * torch.cos(X * (1 + 0.2 * j))
    return X, Y


def _engine(X, Y, ns, d, T, ls, noise, blocked, jitter=1e-6):
    from sampling_gpmpc_b200.engine import GPEngine
    g_ny, n_real = Y.shape[0], X.shape[0]
    old = os.environ.get("GPMPC_K0_BLOCKED_MIN_M")
    os.environ["GPMPC_K0_BLOCKED_MIN_M"] = "1" if blocked else "100000000"
    try:
        eng = GPEngine(ns, g_ny, d, T, n_real)
        eng.set_hypers(np.full((g_ny, d), ls), np.ones(g_ny), np.full((g_ny, T), noise), jitter)
        eng.set_real_data(X, Y)
    finally:
        if old is None:
            del os.environ["GPMPC_K0_BLOCKED_MIN_M"]
        else:
            os.environ["GPMPC_K0_BLOCKED_MIN_M"] = old
    return eng


def _ratio(a, b, s=1.0):
    return float(((a - b).abs() / (RTOL * torch.maximum(b.abs(), torch.tensor(s, device=b.device)))).max())


@pytest.mark.parametrize("n_real,d,T,grad_obs", [(800, 2, 1, False), (1501, 2, 3, False), (333, 2, 3, True), (2003, 3, 1, False),
                                                 (70, 2, 3, True)])
def test_blocked_k0_matches_the_per_pivot_kernels(n_real, d, T, grad_obs):
    """m = 800 / 1501 / 999 (derivative observations) / 2003 / 210: panel counts with and without a ragged last panel, m not a
    multiple of 8 or 16, two GP outputs; posterior through the fused step kernel (reads inv(L_oo)) and through the scalar block
    kernels (read L_oo)."""
    ns, g_ny, H = 6, 2, 1
    X, Y = _data(n_real, d, T, g_ny, grad_obs, 11)
    new = _engine(X, Y, ns, d, T, 0.35, 1e-2, blocked=True)
    ref = _engine(X, Y, ns, d, T, 0.35, 1e-2, blocked=False)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(ns, g_ny, H, d, generator=g, dtype=torch.float64) * 1.6 - 0.8)
    worst = 0.0
    for use_step in (True, False):
        if use_step:
            mn, vn = new.step(x, None)
            mr, vr = ref.step(x, None)
        else:
            new.set_block_kernels(False)
            ref.set_block_kernels(False)
            mn, vn = new.posterior(x)
            mr, vr = ref.posterior(x)
        worst = max(worst, _ratio(mn, mr), _ratio(vn, vr))
    assert new.status() == 0 and ref.status() == 0
    assert worst <= 1.0, f"off by {worst:.3g} x tolerance"


def test_blocked_k0_against_dense_algebra():
    """Value-only model, m = 1100: posterior mean / variance from torch.linalg on the dense kernel matrix (nothing shared with the
    library: no factor layout, no explicit inverse)."""
    n_real, d, ns = 1100, 2, 8
    ls, noise = 0.4, 1e-2
    X, Y = _data(n_real, d, 1, 1, False, 4)
    eng = _engine(X, Y, ns, d, 1, ls, noise, blocked=True)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(ns, 1, 1, d, generator=g, dtype=torch.float64) * 1.6 - 0.8
    mean, var = eng.step(x, None)
    Xs = X / ls
    K = torch.exp(-0.5 * torch.cdist(Xs, Xs) ** 2) + noise * torch.eye(n_real, dtype=torch.float64)
    ks = torch.exp(-0.5 * torch.cdist(x.reshape(ns, d) / ls, Xs) ** 2)  # (ns, n)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(Y[0, :, :1], L)
    mu = (ks @ alpha).reshape(-1)
    v = 1.0 - (ks * torch.cholesky_solve(ks.T, L).T).sum(1)
    assert eng.status() == 0
    assert _ratio(mean.reshape(-1).cpu(), mu) <= 1.0
    assert _ratio(var.reshape(-1).cpu(), v) <= 1.0


def test_blocked_k0_jitter_ladder():
    """A diagonal shifted BELOW zero (noise -5e-3 on a numerically singular kernel matrix): the plain factorisation must fail,
    the ladder's first rung (+1e-2) repairs it -- the blocked kernel reports the same level as the per-pivot one, every CTA
    taking the same decision at every grid barrier, and the handle stays usable."""
    from sampling_gpmpc_b200.engine import ST_TRAIN_NOT_PD
    n_real, d = 900, 2
    X, Y = _data(n_real, d, 1, 1, False, 9)
    st, out = [], []
    for blocked in (True, False):
        eng = _engine(X, Y, 4, d, 1, 0.5, -5e-3, blocked=blocked, jitter=1e-2)
        st.append(eng.status())
        assert st[-1] & ST_TRAIN_NOT_PD == 0
        m, v = eng.step(torch.full((4, 1, 1, d), 0.1, dtype=torch.float64), None)
        assert torch.isfinite(m).all()
        out.append(m)
    assert st[0] == st[1] and st[0] != 0  # TRAIN_JITTER with the same level in both
    assert _ratio(out[0], out[1]) <= 10.0
