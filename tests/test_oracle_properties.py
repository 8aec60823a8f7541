"""CPU: oracle/gp_ref.py against properties that do not come from SURVEY.md Appendix A (tests/gp_properties.py):
the restated GPyTorch arithmetic has no golden vectors (parity unpinned), so it is pinned to mathematics instead --
kernel blocks equal autograd derivatives of the SE kernel (signs, interleaved order), the NaN mask equals row
deletion, the joint draw equals sequential conditioning, noise-free data is interpolated, derivative-task means are
the derivatives of the value mean.  tests/test_gpu_properties.py holds the CUDA path to the same statements."""
import numpy as np
import pytest
import torch

from oracle import gp_ref
from tests import gp_properties as P

F64 = torch.float64


def _oracle_gp(X, Y, ls, os_, noise_diag, use_grad=True, jitter=1e-8, batch=(1, 1)):
    """RefExactGP on data tiled over `batch` = (ns, g_ny) like agent.py:204-214 tiles it."""
    ns, g_ny = batch
    T = Y.shape[-1]
    tx = X.expand(ns, g_ny, *X.shape).clone()
    ty = Y.expand(ns, g_ny, *Y.shape).clone()
    lsb = torch.as_tensor(ls, dtype=F64).reshape(1, 1, 1, -1).expand(ns, g_ny, 1, -1)
    return gp_ref.RefExactGP(tx, ty, lsb, torch.full((ns, g_ny), os_, dtype=F64),
                             torch.zeros(ns, g_ny, 1, dtype=F64),
                             torch.as_tensor(noise_diag, dtype=F64).reshape(1, 1, T).expand(ns, g_ny, T),
                             use_grad=use_grad, jitter=jitter)


@pytest.mark.parametrize("d", [1, 2, 3, 6])
def test_grad_kernel_blocks_are_autograd_derivatives_of_the_se_kernel(d):
    g = torch.Generator().manual_seed(d)
    X1 = torch.rand(4, d, generator=g, dtype=F64) * 2 - 1
    X2 = torch.rand(3, d, generator=g, dtype=F64) * 2 - 1
    ls = 0.5 + torch.rand(d, generator=g, dtype=F64)
    os_ = 0.37
    K = gp_ref.rbf_grad_kernel(X1, X2, ls.reshape(1, d)) * os_
    want = P.autograd_cov_matrix(X1, X2, ls, os_)
    assert P.rel_err(K.numpy(), want.numpy(), os_) < 1e-13
    # value-only kernel = the task-0 rows / columns
    K0 = gp_ref.rbf_kernel(X1, X2, ls.reshape(1, d)) * os_
    assert P.rel_err(K0.numpy(), want[:: d + 1, :: d + 1].numpy(), os_) < 1e-13


@pytest.mark.parametrize("d", [2, 3])
def test_prior_covariance_is_symmetric_and_psd(d):
    g = torch.Generator().manual_seed(10 + d)
    X = torch.rand(12, d, generator=g, dtype=F64) * 2 - 1
    ls = 0.8 + torch.rand(d, generator=g, dtype=F64)
    K = gp_ref.rbf_grad_kernel(X, X, ls.reshape(1, d))
    assert torch.equal(K, K.transpose(-1, -2))  # symmetrised when x1 is x2 (A.1)
    # cross-covariance blocks: cov(d_a f(x), d_b f(x')) = cov(d_b f(x'), d_a f(x))
    X2 = torch.rand(5, d, generator=g, dtype=F64)
    K12 = gp_ref.rbf_grad_kernel(X, X2, ls.reshape(1, d))
    K21 = gp_ref.rbf_grad_kernel(X2, X, ls.reshape(1, d))
    assert P.rel_err(K12.numpy(), K21.T.numpy(), 1.0) < 1e-14
    ev = torch.linalg.eigvalsh(K)
    assert ev.min() > -1e-12 * ev.max()


@pytest.mark.parametrize("d", [1, 2, 3])
def test_derivative_task_means_are_derivatives_of_the_value_mean(d):
    X, Y, xs, ls, os_, noise = P.random_problem(20 + d, n=14, d=d, H=5)
    gp = _oracle_gp(X, Y, ls, os_, noise)
    mean = gp(xs.expand(1, 1, *xs.shape)).mean[0, 0]  # (H, T)
    h = 1e-5
    for a in range(d):
        e = torch.zeros(d, dtype=F64)
        e[a] = h
        mp = gp((xs + e).expand(1, 1, *xs.shape)).mean[0, 0][:, 0]
        mm = gp((xs - e).expand(1, 1, *xs.shape)).mean[0, 0][:, 0]
        fd = (mp - mm) / (2 * h)
        assert P.rel_err(mean[:, 1 + a].numpy(), fd.numpy(), 1.0) < 1e-8  # O(h^2) + cancellation in the difference


@pytest.mark.parametrize("d,nan_fraction", [(2, 0.0), (2, 0.3), (3, 0.3)])
def test_posterior_equals_dense_algebra_with_nan_rows_deleted(d, nan_fraction):
    X, Y, xs, ls, os_, noise = P.random_problem(30 + d, n=10, d=d, H=4, nan_fraction=nan_fraction)
    gp = _oracle_gp(X, Y, ls, os_, noise)
    post = gp(xs.expand(1, 1, *xs.shape))
    mean, cov = P.dense_posterior(X, Y, xs, ls, os_, noise)
    assert P.rel_err(post.mean[0, 0].reshape(-1).numpy(), mean.numpy(), np.sqrt(os_)) < 1e-9
    assert P.rel_err(post.covariance_matrix[0, 0].numpy(), cov.numpy(), os_) < 1e-9


def test_value_only_labels_on_the_derivative_model_equal_the_value_only_model():
    """train_data_has_derivatives False (pendulum1D.py:53-54): derivative slots NaN => the value task of the
    derivative model is the plain RBF GP; two different kernel code paths of the oracle must agree."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(41, n=16, d=d, H=6, with_derivative_labels=False)
    full = _oracle_gp(X, Y, ls, os_, noise, use_grad=True)(xs.expand(1, 1, *xs.shape))
    plain = _oracle_gp(X, Y[:, :1], ls, os_, noise[:1], use_grad=False)(xs.expand(1, 1, *xs.shape))
    assert P.rel_err(full.mean[0, 0][:, 0].numpy(), plain.mean[0, 0][:, 0].numpy(), np.sqrt(os_)) < 1e-11
    assert P.rel_err(full.variance[0, 0][:, 0].numpy(), plain.variance[0, 0][:, 0].numpy(), os_) < 1e-11


def test_nan_in_any_batch_element_masks_the_slot_for_every_element():
    """observation_nan_policy('mask') (A.4): the mask is an any() over the batch."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(42, n=9, d=d, H=3)
    gp = _oracle_gp(X, Y, ls, os_, noise, batch=(3, 1))
    gp.train_y[1, 0, 4, 2] = float("nan")  # one slot, one sample
    post = gp(xs.expand(3, 1, *xs.shape))
    Yd = Y.clone()
    Yd[4, 2] = float("nan")
    mean, cov = P.dense_posterior(X, Yd, xs, ls, os_, noise)
    for s in (0, 2):  # the samples whose own label was NOT NaN lose the slot too
        assert P.rel_err(post.mean[s, 0].reshape(-1).numpy(), mean.numpy(), np.sqrt(os_)) < 1e-9
        assert P.rel_err(post.covariance_matrix[s, 0].numpy(), cov.numpy(), os_) < 1e-9


def test_joint_draw_equals_sequential_conditioning():
    """y = mu + chol(Sigma*) eps in interleaved order = scalar-by-scalar noise-free conditioning --
    what licenses appending the sampled point to a bordered factor."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(51, n=12, d=d, H=4)
    post = _oracle_gp(X, Y, ls, os_, noise)(xs.expand(1, 1, *xs.shape))
    g = torch.Generator().manual_seed(5)
    eps = torch.randn(1, 1, 4, d + 1, generator=g, dtype=F64)
    y = post.sample(base_samples=eps)[0, 0].reshape(-1)
    assert int(post.jitter_level.max()) == 0
    want = P.sequential_draw(post.mean[0, 0].reshape(-1), post.covariance_matrix[0, 0], eps.reshape(-1))
    assert P.rel_err(y.numpy(), want.numpy(), np.sqrt(os_)) < 1e-10


def test_posterior_interpolates_as_noise_goes_to_zero():
    d = 2
    X, Y, _, _, os_, _ = P.random_problem(61, n=10, d=d, H=1)
    ls = np.full(d, 0.35)  # short lengthscale: K stays well conditioned without noise
    gp = _oracle_gp(X, Y, ls, os_, np.full(d + 1, 1e-12))
    post = gp(X.expand(1, 1, *X.shape))
    assert P.rel_err(post.mean[0, 0].numpy(), Y.numpy(), 1.0) < 1e-6
    assert float(post.covariance_matrix[0, 0].diagonal().abs().max()) < 1e-9 * os_ * 1e3
    # and far away the prior comes back: mean 0, variance [os, os / l_a^2]
    far = torch.full((1, d), 60.0, dtype=F64)
    pf = gp(far.expand(1, 1, 1, d))
    assert float(pf.mean.abs().max()) < 1e-12
    want = np.concatenate([[os_], os_ / np.asarray(ls) ** 2])
    assert P.rel_err(pf.variance[0, 0, 0].numpy(), want, 1.0) < 1e-12


def test_sequential_conditioning_with_noise_equals_refit():
    """The persistent bordered factor: conditioning on a sampled point = re-fitting on the data set with
    that point appended (its labels enter WITH likelihood noise)."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(71, n=8, d=d, H=3)
    g = torch.Generator().manual_seed(9)
    Xc, Yc = X, Y
    for i in range(3):
        post = _oracle_gp(Xc, Yc, ls, os_, noise)(xs[i:i + 1].expand(1, 1, 1, d))
        y = post.sample(base_samples=torch.randn(1, 1, 1, d + 1, generator=g, dtype=F64))[0, 0]
        Xc, Yc = torch.cat([Xc, xs[i:i + 1]]), torch.cat([Yc, y])
    probe = torch.tensor([[0.1, -0.2]], dtype=F64)
    refit = _oracle_gp(Xc, Yc, ls, os_, noise)(probe.expand(1, 1, 1, d))
    mean, cov = P.dense_posterior(Xc, Yc, probe, ls, os_, noise)
    assert P.rel_err(refit.mean[0, 0].reshape(-1).numpy(), mean.numpy(), np.sqrt(os_)) < 1e-9
    assert P.rel_err(refit.covariance_matrix[0, 0].numpy(), cov.numpy(), os_) < 1e-9


def test_draw_falls_back_to_the_eigen_root_for_the_whole_batch_when_a_cholesky_fails():
    """root_decomposition: NotPSDError after the ladder => symeig for every batch element (linear_operator);
    root root^T = the covariance with its negative eigenvalues clamped."""
    d = 2
    X, Y, xs, ls, os_, noise = P.random_problem(81, n=9, d=d, H=4)
    gp = _oracle_gp(X, Y, ls, os_, noise, jitter=1e-20, batch=(2, 1))
    x = xs.expand(2, 1, *xs.shape).clone()
    x[1, 0, 2] = x[1, 0, 0]  # sample 1 evaluates one point twice: exactly singular joint covariance
    x[1, 0, 3] = x[1, 0, 0]
    post = gp(x)
    g = torch.Generator().manual_seed(2)
    eps = torch.randn(2, 1, 4, d + 1, generator=g, dtype=F64)
    y = post.sample(base_samples=eps)
    assert (post.jitter_level == 4).all()  # both elements, although only sample 1 is singular
    lam, V = torch.linalg.eigh(post.covariance_matrix)
    root = V * lam.clamp_min(0).sqrt().unsqueeze(-2)
    want = post.mean.reshape(2, 1, -1) + (root @ eps.reshape(2, 1, -1, 1)).squeeze(-1)
    assert torch.equal(y.reshape(2, 1, -1), want)
    # with a working ladder (default jitter) the Cholesky draw is kept and only the singular element escalates
    post2 = _oracle_gp(X, Y, ls, os_, noise, jitter=1e-8, batch=(2, 1))(x)
    post2.sample(base_samples=eps)
    assert int(post2.jitter_level[0, 0]) == 0 and 1 <= int(post2.jitter_level[1, 0]) <= 3
