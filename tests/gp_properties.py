"""Shared, implementation-independent statements of what the GP arithmetic of the hot path must satisfy
(SURVEY.md section 4 (ii)).  Nothing here imports the oracle or the CUDA library: the expected values come from
torch.autograd on the closed-form SE kernel and from plain dense linear algebra, so the same checks can be held
against oracle/gp_ref.py (tests/test_oracle_properties.py, CPU) and against libgpmpc_b200.so through the C ABI
(tests/test_gpu_properties.py).

Prior of the reference model (src/GP_model.py:54-60: ScaleKernel(RBFKernelGrad(ard_num_dims=d)), zero mean):
    k(x, x') = os * exp(-1/2 sum_a ((x_a - x'_a) / l_a)^2)
    cov(task ta at x, task tb at x') = D_ta D'_tb k(x, x'),   D_0 = identity, D_a = d/dx_a, D'_b = d/dx'_b
in GPyTorch's interleaved multitask order (scalar index = point * (d+1) + task).
"""
import numpy as np
import torch

F64 = torch.float64


def se_kernel(x, xp, ls, os_):
    return os_ * torch.exp(-0.5 * (((x - xp) / ls) ** 2).sum())


def autograd_cov_block(x, xp, ls, os_):
    """(d+1, d+1) block cov(task ta at x, task tb at xp) by automatic differentiation of the SE kernel."""
    d = x.numel()
    x = x.clone().to(F64).requires_grad_(True)
    xp = xp.clone().to(F64).requires_grad_(True)
    ls = torch.as_tensor(ls, dtype=F64)
    k = se_kernel(x, xp, ls, os_)
    out = torch.zeros(d + 1, d + 1, dtype=F64)
    out[0, 0] = k.detach()
    gx, gxp = torch.autograd.grad(k, (x, xp), create_graph=True)
    out[1:, 0] = gx.detach()
    out[0, 1:] = gxp.detach()
    for a in range(d):
        (row,) = torch.autograd.grad(gx[a], xp, retain_graph=True)
        out[1 + a, 1:] = row
    return out


def autograd_cov_matrix(X1, X2, ls, os_):
    """Dense (n1 (d+1), n2 (d+1)) prior covariance in interleaved order, block by block from autograd."""
    n1, d = X1.shape
    n2 = X2.shape[0]
    T = d + 1
    K = torch.zeros(n1 * T, n2 * T, dtype=F64)
    for i in range(n1):
        for j in range(n2):
            K[i * T:(i + 1) * T, j * T:(j + 1) * T] = autograd_cov_block(X1[i], X2[j], ls, os_)
    return K


def dense_posterior(X, Y, xs, ls, os_, noise_diag):
    """Textbook GP posterior with NaN labels dropped by DELETING their rows / columns: an independent restatement used
    as the expected value of the mask semantics.  X (n,d), Y (n,T) with NaN = unobserved, xs (H,d), noise_diag (T,).
    Returns mean (H*T,), covariance (H*T, H*T) in interleaved order."""
    n, d = X.shape
    T = Y.shape[1]
    Kf = autograd_cov_matrix(X, X, ls, os_)
    Ks = autograd_cov_matrix(xs, X, ls, os_)
    Kss = autograd_cov_matrix(xs, xs, ls, os_)
    if T == 1:  # value-only model: task 0 rows / columns only
        sel = torch.arange(n) * (d + 1)
        Kf = Kf[sel][:, sel]
        Ks = Ks[torch.arange(xs.shape[0]) * (d + 1)][:, sel]
        Kss = Kss[torch.arange(xs.shape[0]) * (d + 1)][:, torch.arange(xs.shape[0]) * (d + 1)]
    y = Y.reshape(-1)
    keep = ~torch.isnan(y)
    A = Kf[keep][:, keep] + torch.diag(torch.as_tensor(noise_diag, dtype=F64).repeat(n)[keep])
    Ks = Ks[:, keep]
    sol = torch.linalg.solve(A, torch.cat([y[keep, None], Ks.T], 1))
    return Ks @ sol[:, 0], Kss - Ks @ sol[:, 1:]


def random_problem(seed, n, d, H, with_derivative_labels=True, nan_fraction=0.0):
    """Well-conditioned random regression problem: X ~ U[-1,1]^d, y = sum sin(x_i) (+ analytic gradient)."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=F64) * 2 - 1
    xs = torch.rand(H, d, generator=g, dtype=F64) * 1.6 - 0.8
    Y = torch.empty(n, d + 1, dtype=F64)
    Y[:, 0] = torch.sin(X).sum(1)
    Y[:, 1:] = torch.cos(X)
    Y = Y + 1e-3 * torch.randn(n, d + 1, generator=g, dtype=F64)
    if not with_derivative_labels:
        Y[:, 1:] = float("nan")
    elif nan_fraction > 0:
        drop = torch.rand(n, d + 1, generator=g) < nan_fraction
        drop[:, 0] &= torch.rand(n, generator=g) < 0.5
        drop[0] = False  # at least one fully observed point
        Y[drop] = float("nan")
    ls = (0.7 + 0.6 * torch.rand(d, generator=g, dtype=F64)).numpy()
    os_ = 0.8
    noise = (1e-4 * (1.0 + torch.arange(d + 1, dtype=F64))).numpy()
    return X, Y, xs, ls, os_, noise


def sequential_draw(mean, cov, eps):
    """y_i = E[f_i | f_<i = y_<i] + sd(f_i | f_<i) eps_i, scalar by scalar in the given order (noise-free
    conditioning of the joint Gaussian on the values already drawn)."""
    q = mean.numel()
    y = torch.zeros(q, dtype=F64)
    for i in range(q):
        if i == 0:
            m, v = mean[0], cov[0, 0]
        else:
            S = cov[:i, :i]
            c = cov[i, :i]
            w = torch.linalg.solve(S, torch.stack([y[:i] - mean[:i], c], 1))
            m = mean[i] + c @ w[:, 0]
            v = cov[i, i] - c @ w[:, 1]
        y[i] = m + torch.sqrt(v) * eps[i]
    return y


def rel_err(a, b, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale))) if a.size else 0.0


def compare_modulo_eigenvector_signs(y, post, eps, scale):
    """y, eps (q,), post = (mean (q,), cov (q,q)) of ONE batch element.  The draw must be mean + sum_k s_k sqrt(l_k) e_k v_k
    with s_k = +-1, i.e. |v_k . (y - mean)| = sqrt(l_k) |e_k| for every k (V is orthonormal, so this pins y - mean up to
    those signs).  Any eigh resolves eigenvalues only to ~eps_machine * l_max absolutely, which moves sqrt(l_k) e_k by up
    to sqrt(eps_machine * l_max) |e_k| for the (near-)null directions: that floor is added to the 1e-9 tolerance."""
    mean, cov = post
    lam, V = torch.linalg.eigh(cov)
    coef = (V.T @ (y - mean)).abs()
    want = lam.clamp_min(0.0).sqrt() * eps.abs()
    # Sigma* = K** - W^T W carries absolute rounding ~eps_machine * outputscale (= scale^2), whatever its own size
    tol = 1e-9 * scale + 4.0 * float(np.sqrt(2.3e-16 * max(float(lam.max()), scale ** 2))) * float(eps.abs().max())
    err = float((coef - want).abs().max())
    assert err <= tol, (err, tol)
    assert float(want.max()) > 1e3 * tol  # the comparison is not vacuous: the draw is far above the floor
