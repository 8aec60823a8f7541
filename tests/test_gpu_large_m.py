"""GPU: the fused step for training sets beyond k_step's per-warp w array (m in the thousands and up; BASELINE configs[4]
reaches m = 10^4): shared rows by the batched GEMM in k-slabs (k_shared_rows with accumulation), own rows / moments / append by
k_step_big (one CTA per element, w_o in global memory), then the usual finishing kernel.  Held to the scalar substitution
kernels (independent arithmetic: no explicit inverse, no tensor cores) through gpmpc_posterior + gpmpc_append at sizes they
finish quickly, with the path forced on where k_step would normally serve."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _engine(ns, g_ny, d, T, n_real, seed, grad_obs):
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_real, d, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((g_ny, n_real, T), float("nan"), dtype=torch.float64)
    for j in range(g_ny):
        Y[j, :, 0] = torch.sin(X * (1 + 0.2 * j)).sum(1)
        if grad_obs and T > 1:
            Y[j, :, 1:] = (1 + 0.2 * j) * torch.cos(X * (1 + 0.2 * j))
    eng = GPEngine(ns, g_ny, d, T, n_real)
    eng.set_hypers(np.full((g_ny, d), 0.6), np.ones(g_ny), np.full((g_ny, T), 1e-4), 1e-6)
    eng.set_real_data(X, Y)
    return eng


@pytest.mark.parametrize("d,T,n_real,grad_obs,slab_cap", [(2, 3, 203, False, 0), (2, 3, 203, False, 64), (3, 4, 90, True, 128),
                                                           (6, 7, 40, True, 0), (2, 1, 300, False, 72), (2, 3, 1501, False, 0)])
def test_large_m_step_matches_the_scalar_block_kernels(d, T, n_real, grad_obs, slab_cap):
    ns, g_ny, steps = 5, 2, 6
    big = _engine(ns, g_ny, d, T, n_real, 3, grad_obs)
    big.set_option("force_big", 1)
    big.set_option("big_slab_cap", slab_cap)
    ref = _engine(ns, g_ny, d, T, n_real, 3, grad_obs)
    ref.set_block_kernels(False)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(ns, 1, 1, d, generator=g, dtype=torch.float64) * 1.2 - 0.6
    worst = 0.0
    for t in range(steps):
        xx = x.expand(ns, g_ny, 1, d).contiguous()
        eps = torch.randn(ns, g_ny, 1, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
        mb, vb, yb, jb = big.step(xx, eps, big.opts(beta=3.0))
        mr, vr, yr, jr = ref.posterior(xx, eps, ref.opts(beta=3.0))
        ref.append(xx, yb)  # teacher forcing: both condition on the same labels
        assert torch.equal(jb, jr)
        for a, b, s in ((mb, mr, 1.0), (vb, vr, 1.0), (yb, yr, 1.0)):
            worst = max(worst, float(((a - b).abs() / (RTOL * torch.maximum(b.abs(), torch.tensor(s, device=b.device)))).max()))
        x = (x + 0.1 * torch.randn(ns, 1, 1, d, generator=g, dtype=torch.float64)).clamp(-0.9, 0.9)
    assert big.status() == 0 and ref.status() == 0
    assert big.num_factor_rows == ref.num_factor_rows == steps * T
    assert worst <= 1.0, f"off by {worst:.3g} x tolerance"
    # posterior-only query (no draw) through the big path as well
    mb, vb = big.step(xx, None)
    mr, vr = ref.posterior(xx)
    assert float(((mb - mr).abs() / (RTOL * torch.maximum(mr.abs(), torch.tensor(1.0, device=mr.device)))).max()) <= 1.0
    assert float(((vb - vr).abs() / (RTOL * torch.maximum(vr.abs(), torch.tensor(1.0, device=vr.device)))).max()) <= 1.0


def test_m_in_the_thousands_takes_the_slab_path_by_itself():
    """m = 5000: neither k_step's w array nor the one-pass kernel tile fits; no option set."""
    ns, g_ny, d, T = 6, 1, 2, 3
    eng = _engine(ns, g_ny, d, T, 5000, 5, False)
    small = _engine(ns, g_ny, d, T, 5000, 5, False)
    small.set_option("force_block_fallback", 1)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(ns, g_ny, 1, d, generator=g, dtype=torch.float64) * 1.2 - 0.6
    for t in range(2):
        eps = torch.randn(ns, g_ny, 1, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
        l0 = eng.launch_count
        ma, va, ya, _ = eng.step(x, eps, eng.opts(beta=3.0))
        assert eng.launch_count - l0 >= 4  # several GEMM slabs + k_step_big + finish
        mb, vb, yb, _ = small.posterior(x, eps, small.opts(beta=3.0))
        small.append(x, ya)
        assert float((ma - mb).abs().max()) < 1e-8 and float((va - vb).abs().max()) < 1e-8 and float((ya - yb).abs().max()) < 1e-7
        x = (x + 0.05).clamp(-0.9, 0.9)
    assert eng.status() == 0


@pytest.mark.parametrize("n_real,slab_on", [(1300, 1), (1300, 0), (2100, 1)])
def test_mid_m_gemm_in_slabs_with_the_fused_step(n_real, slab_on):
    """1200 < m <= 3600: k_step<WO> takes the shared rows from the batched GEMM with THREE column blocks walked in k-slabs
    (accumulating into Wo) instead of two / one in one pass (option "wo_slab_nb3", default on) -- both forms against the scalar
    substitution kernels."""
    d, T, ns, g_ny, steps = 2, 3, 5, 2, 4
    eng = _engine(ns, g_ny, d, T, n_real, 3, False)
    eng.set_option("wo_slab_nb3", slab_on)
    ref = _engine(ns, g_ny, d, T, n_real, 3, False)
    ref.set_block_kernels(False)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(ns, 1, 1, d, generator=g, dtype=torch.float64) * 1.2 - 0.6
    worst = 0.0
    for t in range(steps):
        xx = x.expand(ns, g_ny, 1, d).contiguous()
        eps = torch.randn(ns, g_ny, 1, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
        mb, vb, yb, jb = eng.step(xx, eps, eng.opts(beta=3.0))
        mr, vr, yr, jr = ref.posterior(xx, eps, ref.opts(beta=3.0))
        ref.append(xx, yb)  # teacher forcing
        assert torch.equal(jb, jr)
        for a, b in ((mb, mr), (vb, vr), (yb, yr)):
            worst = max(worst, float(((a - b).abs() / (RTOL * torch.maximum(b.abs(), torch.tensor(1.0, device=b.device)))).max()))
        x = (x + 0.1 * torch.randn(ns, 1, 1, d, generator=g, dtype=torch.float64)).clamp(-0.9, 0.9)
    assert eng.status() == 0 and ref.status() == 0
    assert worst <= 1.0, f"off by {worst:.3g} x tolerance"
