"""GPU: libgpmpc_b200.so (through the C ABI, GPEngine) against the implementation-independent properties of
tests/gp_properties.py -- the same statements tests/test_oracle_properties.py holds the oracle to, so the two
implementations are each pinned to mathematics (autograd derivatives of the SE kernel, dense linear algebra with
NaN rows deleted, sequential conditioning), not only to one another.  Tolerance 1e-9 * max(|b|, s) unless a
finite-difference truncation error is involved."""
import numpy as np
import pytest
import torch

from tests import gp_properties as P

pytestmark = pytest.mark.gpu
F64 = torch.float64
RTOL = 1e-9


def _engine(X, Y, ls, os_, noise, ns, T=None, jitter=1e-8, g_ny=1):
    """Y (n, T) -> one engine with g_ny identical outputs."""
    from sampling_gpmpc_b200.engine import GPEngine
    n, d = X.shape
    T = Y.shape[1] if T is None else T
    eng = GPEngine(ns, g_ny, d, T, n)
    eng.set_hypers(np.tile(np.asarray(ls), (g_ny, 1)), np.full(g_ny, os_), np.tile(np.asarray(noise)[:T], (g_ny, 1)), jitter)
    eng.set_real_data(X, Y[:, :T].expand(g_ny, n, T).contiguous())
    return eng


@pytest.mark.parametrize("d", [1, 2, 3, 6])
def test_kernel_blocks_read_through_the_posterior_mean_are_autograd_derivatives(d):
    """One training point x1 with all T tasks observed and labels y = (K11 + Sigma) e_b: the posterior mean at x* is
    column b of cov(tasks at x*, tasks at x1).  K11 is diagonal ([os, os/l_a^2]) because r = 0."""
    from sampling_gpmpc_b200.engine import GPEngine
    T = d + 1
    g = torch.Generator().manual_seed(100 + d)
    x1 = torch.rand(1, d, generator=g, dtype=F64) - 0.5
    ls = (0.6 + torch.rand(d, generator=g, dtype=F64)).numpy()
    os_, noise = 0.37, 1e-3 * (1.0 + np.arange(T))
    kdiag = np.concatenate([[os_], os_ / ls ** 2]) + noise
    ns = 7
    xs = torch.rand(ns, d, generator=g, dtype=F64) * 2 - 1
    eng = GPEngine(ns, T, d, T, 1)  # output j carries label e_j
    eng.set_hypers(np.tile(ls, (T, 1)), np.full(T, os_), np.tile(noise, (T, 1)), 1e-8)
    eng.set_real_data(x1, torch.diag(torch.tensor(kdiag)).reshape(T, 1, T))
    x = xs.reshape(ns, 1, 1, d).expand(ns, T, 1, d).contiguous()
    for fused in (False, True):
        mean = (eng.step(x, None)[0] if fused else eng.posterior(x)[0]).cpu()  # (ns, T(b), 1, T(a))
        for s in range(ns):
            want = P.autograd_cov_block(xs[s], x1[0], ls, os_)  # [a, b]
            got = mean[s, :, 0, :].T.numpy()
            assert P.rel_err(got, want.numpy(), os_) < 1e-12, (fused, s)
    assert eng.status() == 0


@pytest.mark.parametrize("d", [1, 2, 3])
def test_derivative_task_means_are_derivatives_of_the_value_mean(d):
    X, Y, xs, ls, os_, noise = P.random_problem(20 + d, n=14, d=d, H=5)
    H = xs.shape[0]
    eng = _engine(X, Y, ls, os_, noise, ns=1)
    mean = eng.posterior(xs.reshape(1, 1, H, d))[0][0, 0].cpu()
    h = 1e-5
    for a in range(d):
        e = torch.zeros(d, dtype=F64)
        e[a] = h
        mp = eng.posterior((xs + e).reshape(1, 1, H, d))[0][0, 0, :, 0].cpu()
        mm = eng.posterior((xs - e).reshape(1, 1, H, d))[0][0, 0, :, 0].cpu()
        assert P.rel_err(mean[:, 1 + a].numpy(), ((mp - mm) / (2 * h)).numpy(), 1.0) < 1e-8


@pytest.mark.parametrize("d,nan_fraction", [(2, 0.0), (2, 0.3), (3, 0.3)])
def test_posterior_and_joint_draw_equal_dense_algebra_with_nan_rows_deleted(d, nan_fraction):
    X, Y, xs, ls, os_, noise = P.random_problem(30 + d, n=10, d=d, H=4, nan_fraction=nan_fraction)
    H, T = xs.shape[0], d + 1
    mean_w, cov_w = P.dense_posterior(X, Y, xs, ls, os_, noise)
    g = torch.Generator().manual_seed(3)
    eps = torch.randn(1, 1, H, T, generator=g, dtype=F64)
    y_w = mean_w + torch.linalg.cholesky(cov_w) @ eps.reshape(-1)
    y_seq = P.sequential_draw(mean_w, cov_w, eps.reshape(-1))
    for mma in (True, False):  # tensor-core block kernels and the scalar substitution kernels
        eng = _engine(X, Y, ls, os_, noise, ns=1)
        eng.set_block_kernels(mma)
        mean, var, y, jl = eng.posterior(xs.reshape(1, 1, H, d), eps)
        assert int(jl.max()) == 0 and eng.status() == 0
        assert P.rel_err(mean.cpu().reshape(-1).numpy(), mean_w.numpy(), np.sqrt(os_)) < RTOL
        assert P.rel_err(var.cpu().reshape(-1).numpy(), cov_w.diagonal().numpy(), os_) < RTOL
        assert P.rel_err(y.cpu().reshape(-1).numpy(), y_w.numpy(), np.sqrt(os_)) < RTOL
        # joint Cholesky draw == scalar-by-scalar conditioning in interleaved order
        assert P.rel_err(y.cpu().reshape(-1).numpy(), y_seq.numpy(), np.sqrt(os_)) < RTOL


def test_value_only_labels_on_the_derivative_model_equal_the_value_only_model():
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(41, n=16, d=d, H=6, with_derivative_labels=False)
    H = xs.shape[0]
    full = _engine(X, Y, ls, os_, noise, ns=1)
    plain = _engine(X, Y, ls, os_, noise, ns=1, T=1)
    assert full.num_real_observed == plain.num_real_observed == 16  # NaN derivative slots are not in the factor
    mf, vf = full.posterior(xs.reshape(1, 1, H, d))
    mp_, vp = plain.posterior(xs.reshape(1, 1, H, d))
    assert P.rel_err(mf[0, 0, :, 0].cpu().numpy(), mp_[0, 0, :, 0].cpu().numpy(), np.sqrt(os_)) < RTOL
    assert P.rel_err(vf[0, 0, :, 0].cpu().numpy(), vp[0, 0, :, 0].cpu().numpy(), os_) < RTOL


def test_posterior_interpolates_as_noise_goes_to_zero_and_reverts_to_the_prior_far_away():
    d = 2
    X, Y, _, _, os_, _ = P.random_problem(61, n=10, d=d, H=1)
    ls = np.full(d, 0.35)
    eng = _engine(X, Y, ls, os_, np.full(d + 1, 1e-12), ns=1)
    mean, var = eng.posterior(X.reshape(1, 1, 10, d))
    assert eng.status() == 0
    assert P.rel_err(mean[0, 0].cpu().numpy(), Y.numpy(), 1.0) < 1e-6
    assert float(var.max()) <= 1e-6 * os_
    far = torch.full((1, 1, 1, d), 60.0, dtype=F64)
    mean, var = eng.posterior(far)
    assert float(mean.abs().max()) < 1e-12
    want = np.concatenate([[os_], os_ / ls ** 2])
    assert P.rel_err(var[0, 0, 0].cpu().numpy(), want, 1.0) < 1e-12


@pytest.mark.parametrize("fused", [True, False])
def test_conditioning_on_sampled_points_equals_dense_algebra_on_the_grown_data_set(fused):
    """The bordered factor after k appends (rank-T append of the fused step kernel, or k_append of the block kernels)
    gives the posterior of the data set [real || sampled points], the sampled labels entering WITH noise."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(71, n=8, d=d, H=5)
    T, ns = d + 1, 3
    eng = _engine(X, Y, ls, os_, noise, ns=ns)
    g = torch.Generator().manual_seed(9)
    offs = 0.05 * torch.randn(ns, 1, d, generator=g, dtype=F64)  # every sample conditions on its own points
    Xc = [X.clone() for _ in range(ns)]
    Yc = [Y.clone() for _ in range(ns)]
    for i in range(4):
        x = (xs[i].reshape(1, 1, d) + offs).reshape(ns, 1, 1, d)
        eps = torch.randn(ns, 1, 1, T, generator=g, dtype=F64)
        if fused:
            _, _, y, _ = eng.step(x, eps)
        else:
            _, _, y, _ = eng.posterior(x, eps)
            eng.append(x, y)
        for s in range(ns):
            # the draw itself: dense posterior of sample s's current data set at its own point
            m_w, c_w = P.dense_posterior(Xc[s], Yc[s], x[s, 0].cpu(), ls, os_, noise)
            y_w = m_w + torch.linalg.cholesky(c_w) @ eps[s, 0, 0]
            assert P.rel_err(y[s, 0, 0].cpu().numpy(), y_w.numpy(), np.sqrt(os_)) < RTOL, (i, s)
            Xc[s] = torch.cat([Xc[s], x[s, 0].cpu()])
            Yc[s] = torch.cat([Yc[s], y[s, 0].cpu()])
    probe = torch.tensor([[0.1, -0.2], [0.4, 0.3]], dtype=F64)
    mean, var = eng.posterior(probe.expand(ns, 1, 2, d).contiguous())
    assert eng.status() == 0
    for s in range(ns):
        m_w, c_w = P.dense_posterior(Xc[s], Yc[s], probe, ls, os_, noise)
        assert P.rel_err(mean[s, 0].cpu().reshape(-1).numpy(), m_w.numpy(), np.sqrt(os_)) < RTOL
        assert P.rel_err(var[s, 0].cpu().reshape(-1).numpy(), c_w.diagonal().numpy(), os_) < RTOL
