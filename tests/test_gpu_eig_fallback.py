"""GPU: GPyTorch's eigen-root fallback of the draw (src/agent.py:641 -> MultivariateNormal.rsample ->
root_decomposition: NotPSDError after the jitter ladder => "Using symeig method" for the WHOLE batch), restated in
oracle/gp_ref.py (RefPosterior._root) and implemented by gpmpc_eig.cuh.  With the car yamls' Dyn_gp_jitter 1e-20 the ladder
is a no-op, so this is the draw whenever a joint posterior covariance is numerically singular.

Eigenvectors are defined up to sign (LAPACK applies no convention), so two eigh implementations agree on root root^T but
not on root: the draws are compared modulo one sign per eigenvector, on the eigen-directions that carry the draw."""
import numpy as np
import pytest
import torch

from oracle import gp_ref
from tests import gp_properties as P

pytestmark = pytest.mark.gpu
F64 = torch.float64


# jitter 0.0: the ladder is an exact no-op (the car yamls' 1e-20 is one too, except that it rescues a pivot that is EXACTLY 0,
# which is what exactly duplicated points tend to produce)
def _problem(seed, ns, g_ny, n_real, d=2, noise=1e-30):
    g = torch.Generator().manual_seed(seed)
    T = d + 1
    X = torch.rand(n_real, d, generator=g, dtype=F64) * 2 - 1
    Y = torch.empty(g_ny, n_real, T, dtype=F64)
    for j in range(g_ny):
        Y[j, :, 0] = torch.sin(X * (1 + 0.3 * j)).sum(1)
        Y[j, :, 1:] = (1 + 0.3 * j) * torch.cos(X * (1 + 0.3 * j))
    ls = np.stack([np.array([1.0, 1.3]) + 0.2 * j for j in range(g_ny)])
    os_ = np.array([0.5 + 0.25 * j for j in range(g_ny)])
    nz = np.full((g_ny, T), noise)
    return X, Y, ls, os_, nz, g


def _engine(X, Y, ls, os_, nz, ns, jitter=0.0):
    from sampling_gpmpc_b200.engine import GPEngine
    g_ny, n_real, T = Y.shape
    eng = GPEngine(ns, g_ny, X.shape[1], T, n_real)
    eng.set_hypers(ls, os_, nz, jitter)
    eng.set_real_data(X, Y)
    return eng


def _oracle(X, Y, ls, os_, nz, ns, jitter=0.0):
    g_ny, n_real, T = Y.shape
    d = X.shape[1]
    return gp_ref.RefExactGP(X.expand(ns, g_ny, n_real, d).clone(), Y.expand(ns, g_ny, n_real, T).clone(),
                             torch.tensor(ls).reshape(1, g_ny, 1, d).expand(ns, g_ny, 1, d),
                             torch.tensor(os_).reshape(1, g_ny).expand(ns, g_ny), torch.zeros(ns, g_ny, 1, dtype=F64),
                             torch.tensor(nz).reshape(1, g_ny, T).expand(ns, g_ny, T), use_grad=True, jitter=jitter)


@pytest.mark.parametrize("mma", [True, False])
def test_joint_draw_falls_back_to_the_eigen_root_for_the_whole_batch(mma):
    from sampling_gpmpc_b200.engine import NotPSDError, ST_SAMPLE_EIG
    ns, g_ny, H, d, T = 3, 2, 5, 2, 3
    X, Y, ls, os_, nz, g = _problem(5, ns, g_ny, n_real=7, noise=1e-8)
    x = torch.rand(ns, g_ny, H, d, generator=g, dtype=F64) * 1.6 - 0.8
    x[1, :, 1:] = x[1, :, :1]  # sample 1 evaluates one point five times: its joint covariance has rank 3 of 15
    eps = torch.randn(ns, g_ny, H, T, generator=g, dtype=F64).clamp(-2.5, 2.5)
    eng = _engine(X, Y, ls, os_, nz, ns)
    eng.set_block_kernels(mma)
    mean, var, y, jl = eng.posterior(x, eps)
    assert eng.status() & ST_SAMPLE_EIG, "the duplicated test points did not break the Cholesky: pick another seed"
    assert (jl.cpu().numpy() == 4).all()  # batch-wide, like root_decomposition
    with pytest.warns(RuntimeWarning, match="eigen root"):
        eng.raise_on_status()
    post = _oracle(X, Y, ls, os_, nz, ns)(x)
    y_ref = post.sample(base_samples=eps)
    assert (post.jitter_level.numpy() == 4).all()
    for s in range(ns):
        for j in range(g_ny):
            assert float((mean[s, j].cpu() - post.mean[s, j]).abs().max()) < 1e-9 * np.sqrt(os_[j]) * 10
            P.compare_modulo_eigenvector_signs(y[s, j].cpu().reshape(-1), (post.mean[s, j].reshape(-1),
                                              post.covariance_matrix[s, j]), eps[s, j].reshape(-1), np.sqrt(os_[j]))
    # root root^T is unique: the oracle's own draw differs from ours at most by those signs, so |y - mean| projected on each
    # eigenvector agrees -- and with eps = 0 both are the mean
    mean0, _, y0, _ = eng.posterior(x, torch.zeros_like(eps))
    assert torch.equal(y0, mean0)
    assert torch.isfinite(y_ref).all() and torch.isfinite(y).all()
    eng.status(clear=True)

    # fallback switched off: psd_safe_cholesky's NotPSDError, failing elements NaN, the others keep their Cholesky draw
    _, _, y2, jl2 = eng.posterior(x, eps, eng.opts(eig_fallback=False))
    jl2 = jl2.cpu().numpy()
    assert (jl2[1] == 4).all() and (jl2[[0, 2]] == 0).all()
    assert torch.isnan(y2[1]).all() and torch.isfinite(y2[[0, 2]]).all()
    L = torch.linalg.cholesky(post.covariance_matrix[[0, 2]])
    want = post.mean[[0, 2]].reshape(2, g_ny, -1) + (L @ eps[[0, 2]].reshape(2, g_ny, -1, 1)).squeeze(-1)
    # (15 x 15 joint covariance of 5 nearby points with derivatives: lambda_min ~ 1e-12, the Cholesky draw itself moves by
    # ~1e-8 under rounding-level perturbations; an eigen-root draw would differ at the 1e-1 level)
    assert float((y2[[0, 2]].cpu().reshape(2, g_ny, -1) - want).abs().max()) < 1e-6
    with pytest.raises(NotPSDError):
        eng.raise_on_status()


def test_fused_step_redoes_the_whole_batch_through_the_eigen_root():
    """H = 1 rollout step (k_step + k_step_finish): one sample sits exactly on a noise-free real point, its 3 x 3 posterior
    covariance is rounding noise around 0 and the unjittered Cholesky fails => every element's draw (and the rows it
    appends) is redone through the eigen root by the launch queued behind the regular one."""
    from sampling_gpmpc_b200.engine import ST_SAMPLE_EIG
    ns, g_ny, d, T = 4, 2, 2, 3
    X, Y, ls, os_, nz, g = _problem(11, ns, g_ny, n_real=2)  # two distant noise-free points: K stays well conditioned
    X[0], X[1] = torch.tensor([-0.6, -0.5], dtype=F64), torch.tensor([0.7, 0.4], dtype=F64)
    x = torch.rand(ns, g_ny, 1, d, generator=g, dtype=F64) * 1.6 - 0.8
    x[2, :, 0] = X[0]
    eps = torch.randn(ns, g_ny, 1, T, generator=g, dtype=F64).clamp(-2.5, 2.5)
    eng = _engine(X, Y, ls, os_, nz, ns)
    mean, var, y, jl = eng.step(x, eps)
    st = eng.status(clear=True)
    assert st & ST_SAMPLE_EIG, "the singular element did not break the Cholesky: pick another seed"
    assert (jl.cpu().numpy() == 4).all()
    post = _oracle(X, Y, ls, os_, nz, ns)(x)
    post.sample(base_samples=eps)
    assert (post.jitter_level.numpy() == 4).all()
    for s in (0, 1, 3):  # well-conditioned elements: the eigen-root draw is far from the Cholesky draw, and resolvable
        for j in range(g_ny):
            P.compare_modulo_eigenvector_signs(y[s, j].cpu().reshape(-1), (post.mean[s, j].reshape(-1),
                                              post.covariance_matrix[s, j]), eps[s, j].reshape(-1), np.sqrt(os_[j]))
    assert float((y[2].cpu() - post.mean[2]).abs().max()) < 1e-6  # the singular element: mean +- sqrt(rounding)
    # the appended rows belong to the eigen-root labels: the next posterior of a well-conditioned element is the oracle's
    # re-fit on [real || (x, y)]
    probe = torch.rand(ns, g_ny, 2, d, generator=g, dtype=F64) - 0.5
    m2, v2 = eng.posterior(probe)
    if True:
        for s in (0, 1, 3):
            Xs = torch.cat([X.expand(g_ny, *X.shape), x[s].cpu()], 1)
            Ys = torch.cat([Y, y[s].cpu()], 1)
            ref = gp_ref.RefExactGP(Xs[None], Ys[None], torch.tensor(ls).reshape(1, g_ny, 1, d),
                                    torch.tensor(os_).reshape(1, g_ny), torch.zeros(1, g_ny, 1, dtype=F64),
                                    torch.tensor(nz).reshape(1, g_ny, T), use_grad=True, jitter=1e-20)(probe[s][None])
            assert float((m2[s].cpu() - ref.mean[0]).abs().max()) < 1e-5  # noise-free K: cond ~1e8, the re-fit itself is only that good
