"""GPU: the SQP-mode conditioning with the append's Cholesky factorised BESIDE the draw's (k_pm_finish's second CTA row -> st.Lpre,
picked up by k_append; gpmpc_linearise requests it, gpmpc_set_option(h, "prefactor_next", 1) requests it for separate
posterior / append calls).  The prefactored path runs the same fill expression and the same packed Cholesky on the same matrix,
so everything downstream must be BIT-identical to the plain path -- including when the request cannot be honoured (other test
points at the append, a reset in between, masked scalars)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(ns=6, g_ny=2, d=2, T=3, n_real=30, seed=2):
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_real, d, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((g_ny, n_real, T), float("nan"), dtype=torch.float64)
    for j in range(g_ny):
        Y[j, :, 0] = torch.sin(X * (1 + 0.3 * j)).sum(1)
    eng = GPEngine(ns, g_ny, d, T, n_real, cap_points=64)
    eng.set_hypers(np.full((g_ny, d), 0.8), np.ones(g_ny), np.full((g_ny, T), 1e-4), 1e-6)
    eng.set_real_data(X, Y)
    return eng


def _iterate(eng, prefactor, H, variant):
    g = torch.Generator().manual_seed(9)
    ns, g_ny, d, T = eng.ns, eng.g_ny, eng.d, eng.T
    outs = []
    for it in range(4):
        x = (torch.rand(ns, g_ny, H, d, generator=g, dtype=torch.float64) * 1.6 - 0.8).cuda()
        eps = torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).cuda()
        if prefactor:
            eng.set_option("prefactor_next", 1)
        mean, var, y, jl = eng.posterior(x, eps, eng.opts(beta=2.5))
        if variant == "reset" and it == 2:
            eng.reset_hallucinated()          # the factor changes between the model call and the conditioning
        if variant == "other_points" and it == 1:
            x = (x + 0.01).contiguous()       # the append is asked for OTHER points than the model call saw
        if variant == "masked" and it == 1:
            active = np.ones(H, dtype=np.uint8)
            active[H // 2] = 0
            eng.append(x, y, active)
        else:
            eng.append(x, y)
        outs += [mean.clone(), var.clone(), y.clone(), jl.clone()]
    xq = (torch.rand(ns, g_ny, 3, d, generator=g, dtype=torch.float64) * 1.6 - 0.8).cuda()
    m2, v2 = eng.posterior(xq)
    Xh, Yh = eng.export_hallucinated()
    return outs + [m2, v2, Xh, Yh], eng.status(), eng.num_factor_rows


@pytest.mark.parametrize("H", [1, 5, 17])
@pytest.mark.parametrize("variant", ["plain", "reset", "other_points", "masked"])
def test_prefactored_append_is_bit_identical(H, variant):
    a, sa, ra = _iterate(_engine(), True, H, variant)
    b, sb, rb = _iterate(_engine(), False, H, variant)
    assert sa == sb == 0 and ra == rb
    for u, v in zip(a, b):
        assert torch.equal(torch.nan_to_num(u.double(), nan=-7.0), torch.nan_to_num(v.double(), nan=-7.0))
