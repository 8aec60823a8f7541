"""GPU parity of the rows either side of the hot path (SURVEY.md 8f and the Dyn_gp_min_data_dist rules of a9/a11):
every kernel through the C ABI against oracle/consumers_ref.py.  All of it is copy / compare / index work, so the
bar is BIT-EXACT (the one exception is stated where it applies)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

F64 = torch.float64


def _engine_with_data(ns, g_ny, d, T, n_real, n_h, seed, nan_real_grads=True, dup_frac=0.3):
    """An engine whose data set has real points (derivative targets NaN if asked) and n_h recorded hallucinated
    points per element, some of them with NaN labels (record-only handle: no factor needed for these kernels)."""
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_real, d, generator=g, dtype=F64) * 2 - 1
    Y = torch.randn(g_ny, n_real, T, generator=g, dtype=F64)
    if nan_real_grads and T > 1:
        Y[:, : n_real // 2, 1:] = float("nan")  # half the real points are value-only (never "fully observed")
    eng = GPEngine(ns, g_ny, d, T, n_real, cap_points=max(n_h, 1))
    eng.set_condition_on_hallucinated(False)
    eng.set_hypers(np.ones((g_ny, d)), np.ones(g_ny), np.full((g_ny, T), 1e-6), 1e-6)
    eng.set_real_data(X, Y)
    Xh = torch.rand(ns, g_ny, n_h, d, generator=g, dtype=F64) * 2 - 1
    Yh = torch.randn(ns, g_ny, n_h, T, generator=g, dtype=F64)
    if n_h:
        Yh[torch.rand(ns, g_ny, n_h, generator=g) < 0.25] = float("nan")
        eng.append(Xh.cuda(), Yh.cuda())
    return eng, X, Y, Xh, Yh, g


def _test_points(X, Xh, H, g, dup_frac=0.4, eps=1e-5):
    """Test inputs: random, but a fraction sits within eps of a real / hallucinated input (the near-duplicates the
    rule exists for), incl. exact duplicates."""
    ns, g_ny, n_h, d = Xh.shape
    x = torch.rand(ns, g_ny, H, d, generator=g, dtype=F64) * 2 - 1
    for s in range(ns):
        for j in range(g_ny):
            for h in range(H):
                r = float(torch.rand(1, generator=g))
                if r < dup_frac / 2:
                    x[s, j, h] = X[int(torch.randint(X.shape[0], (1,), generator=g))] + eps * 0.3 * torch.randn(d, generator=g, dtype=F64)
                elif r < dup_frac and n_h:
                    x[s, j, h] = Xh[s, j, int(torch.randint(n_h, (1,), generator=g))] + (0.0 if r < dup_frac * 0.75 else eps * 0.3) * torch.randn(d, generator=g, dtype=F64)
    return x


@pytest.mark.parametrize("d,T,n_real,n_h,H", [(3, 4, 45, 9, 1), (2, 3, 36, 17, 17), (3, 4, 20, 0, 5), (6, 7, 33, 40, 3)])
def test_min_dist_overwrite_bit_exact(d, T, n_real, n_h, H):
    from oracle import consumers_ref as ref
    ns, g_ny = 11, 2
    eng, X, Y, Xh, Yh, g = _engine_with_data(ns, g_ny, d, T, n_real, n_h, 1)
    x = _test_points(X, Xh, H, g)
    y = torch.randn(ns, g_ny, H, T, generator=g, dtype=F64)
    mean = 0.1 * torch.randn(ns, g_ny, H, T, generator=g, dtype=F64)
    var = torch.rand(ns, g_ny, H, T, generator=g, dtype=F64) * 0.5
    x_train = torch.cat([X.expand(ns, g_ny, n_real, d), Xh], 2)
    y_train = torch.cat([Y.expand(ns, g_ny, n_real, T), Yh], 2)
    # the overwrite itself is index / copy work: bit-exact (no clipping: beta < 0 here, a huge beta in the oracle)
    want = ref.min_dist_overwrite(x, x_train, y_train, y, mean, var, 1e-5, 1e30)
    got = eng.min_dist_overwrite(x.cuda(), None, None, y.cuda().clone(), 1e-5, -1.0).cpu()
    assert (want != y).any() or n_h == 0, "test must exercise the overwrite"
    assert torch.equal(got, want)
    # with the truncation (floating point: mean +- beta sqrt(var)); torch's CPU sqrt is MKL's vdSqrt, which is not
    # correctly rounded (0.6 % of inputs are 1 ulp off), CUDA's is: a few ulp on the clipped entries, nothing else
    want = ref.min_dist_overwrite(x, x_train, y_train, y, mean, var, 1e-5, 2.5)
    got = eng.min_dist_overwrite(x.cuda(), mean.cuda(), var.cuda(), y.cuda().clone(), 1e-5, 2.5).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=4e-16)
    assert (got != want).float().mean() < 0.02


@pytest.mark.parametrize("d,T,n_real,n_h,H,use_h", [(3, 4, 45, 9, 1, True), (2, 3, 36, 17, 17, True), (2, 3, 36, 17, 17, False),
                                                     (6, 7, 33, 40, 3, True)])
def test_filter_new_points_bit_exact(d, T, n_real, n_h, H, use_h):
    from oracle import consumers_ref as ref
    from sampling_gpmpc_b200.rollout import reduce_filter_counts
    ns, g_ny = 11, 2
    eng, X, Y, Xh, Yh, g = _engine_with_data(ns, g_ny, d, T, n_real, n_h, 2)
    x = _test_points(X, Xh, H, g)
    x[:, :, 0] = X[3]  # point 0 duplicates a real input for EVERY sample: the reference drops it
    y = torch.randn(ns, g_ny, H, T, generator=g, dtype=F64)
    X_cond = torch.cat([X.expand(ns, g_ny, n_real, d), Xh], 2) if use_h else X.expand(ns, g_ny, n_real, d)
    y_want, filt, f_all = ref.filter_new_points(x, y, X_cond, 1e-5)
    yg = y.cuda().clone()
    counts = eng.filter_new_points(x.cuda(), yg, 1e-5, use_hallucinated=use_h)
    assert torch.equal(counts.cpu(), filt.sum(0).to(torch.int32))
    assert torch.equal(torch.isnan(yg.cpu()), torch.isnan(y_want))
    assert torch.equal(torch.nan_to_num(yg.cpu()), torch.nan_to_num(y_want))
    flags = reduce_filter_counts(counts, ns, 1)
    assert np.array_equal(flags[0], f_all.numpy()) and flags[0][0]
    assert np.array_equal(flags[1], filt.reshape(-1, H).any(0).numpy())


@pytest.mark.parametrize("ns,nx,nu,H,n_tail,use_K", [(70, 2, 1, 17, 5, False), (20, 4, 2, 15, 9, True), (3, 4, 2, 50, 0, True),
                                                      (1, 2, 1, 1, 3, True), (200, 4, 2, 50, 7, False)])
def test_pack_plin_matches_reference_concat_loop(ns, nx, nu, H, n_tail, use_K):
    """gpmpc_pack_plin vs the reference's per-stage, per-sample np.concatenate (src/solver.py:98-131).  Pure copies:
    bit-exact; with feedback and nu > 1 the u_grad @ K sum may differ from BLAS by one rounding (2e-15 absolute on O(1) data)."""
    from oracle import consumers_ref as ref
    from sampling_gpmpc_b200.engine import GPEngine, GpmpcEnv
    g = torch.Generator().manual_seed(ns + H)
    eng = GPEngine(ns, 1, 2, 3, 4)
    env = GpmpcEnv()
    env.nx, env.nu, env.g_ny, env.d = nx, nu, 1, 2
    K = torch.randn(nu, nx, generator=g, dtype=F64).numpy()
    for i, v in enumerate(K.reshape(-1)):
        env.K_fb[i] = v
    lin = torch.randn(ns, nx, H, 1 + nx + nu, generator=g, dtype=F64)
    x_h = torch.randn(H, ns * nx, generator=g, dtype=F64)
    n_w = max(0, n_tail - nu - 2)
    n_g = n_tail - nu - n_w - (1 if n_tail else 0) if n_tail else 0
    u_h = torch.randn(H, nu if n_tail else 0, generator=g, dtype=F64).numpy()
    xg = torch.randn(H, max(n_g, 0), generator=g, dtype=F64).numpy()
    w = torch.randn(H, n_w, generator=g, dtype=F64).numpy()
    te = [np.array([float(t)]) if n_tail else np.empty(0) for t in range(H)]
    tail = np.hstack([u_h, xg, w, np.stack(te)]) if n_tail else None
    assert tail is None or tail.shape[1] == n_tail
    ln = lin.numpy()
    want = ref.pack_p_lin(ln[:, :, :, [0]], ln[:, :, :, 1:1 + nx], ln[:, :, :, 1 + nx:], x_h.numpy(), u_h, xg, w, te, ns, nx,
                          K if use_K else None)
    got = eng.pack_plin(env, lin.cuda(), x_h.cuda(), None if tail is None else torch.tensor(tail), use_K).cpu().numpy()
    want = np.stack(want)
    assert got.shape == want.shape == (H, ns * (nx * nx + nx * nu + 2 * nx) + n_tail)
    if use_K and nu > 1:
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-15)  # O(1) operands: one rounding of the k-sum
        exact = np.ones(want.shape[1], dtype=bool)
        per = nx * nx + nx * nu + 2 * nx
        for i in range(ns):
            exact[i * per: i * per + nx * nx] = False
        assert np.array_equal(got[:, exact], want[:, exact])
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("ns,nx,H1", [(1, 4, 51), (257, 4, 51), (5000, 2, 31), (100003, 4, 51)])
def test_traj_stats_bit_exact(ns, nx, H1):
    from oracle import consumers_ref as ref
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(ns)
    traj = torch.randn(ns, nx, H1, generator=g, dtype=F64).cumsum(-1)
    mu = traj.mean(0) + 0.01
    eng = GPEngine(4, 1, 2, 3, 4)
    lo, hi, dev = eng.traj_stats(traj.cuda(), mu.cuda())
    w_lo, w_hi, w_dev = ref.traj_stats(traj.numpy(), mu.numpy())
    assert np.array_equal(lo.cpu().numpy(), w_lo) and np.array_equal(hi.cpu().numpy(), w_hi)
    assert np.array_equal(dev.cpu().numpy(), w_dev)
    lo2, hi2 = eng.traj_stats(traj.cuda())
    assert torch.equal(lo2, lo) and torch.equal(hi2, hi)


def _cyclic_equal(a, b):
    a, b = list(a), list(b)
    if sorted(a) != sorted(b):
        return False
    if not a:
        return True
    k = b.index(a[0])
    return a == b[k:] + b[:k]


@pytest.mark.parametrize("ns,H1,shape", [(200, 51, "gauss"), (4000, 51, "gauss"), (50000, 12, "banana"), (3000, 5, "uniform")])
def test_stage_hulls_match_qhull(ns, H1, shape):
    """Per-stage hull vertices: same sample indices, same counter-clockwise order as scipy ConvexHull (qhull), incl.
    stage 0 where every sample starts from the same state (a single point)."""
    from oracle import consumers_ref as ref
    from sampling_gpmpc_b200.engine import GPEngine
    g = torch.Generator().manual_seed(ns + H1)
    nx = 4
    z = torch.randn(ns, nx, H1, generator=g, dtype=F64)
    if shape == "banana":
        z[:, 1] = z[:, 1] * 0.2 + z[:, 0] ** 2
    elif shape == "uniform":
        z = torch.rand(ns, nx, H1, generator=g, dtype=F64)
    traj = z * torch.linspace(0, 1, H1, dtype=F64) + torch.tensor([0.0, 1.95, 0.0, 14.0], dtype=F64)[None, :, None]
    eng = GPEngine(4, 1, 2, 3, 4)
    got = eng.stage_hulls(traj.cuda(), 0, 1)
    want = ref.stage_hulls(traj.numpy(), 0, 1)
    assert len(got) == H1
    assert list(got[0]) == [0]  # all samples coincide at stage 0: the lowest index stands for the point
    for t in range(1, H1):
        assert _cyclic_equal(got[t], want[t]), f"stage {t}: {sorted(got[t])} vs {sorted(want[t])}"


def test_stage_hulls_at_scale_properties():
    """10^6 samples x 51 stages (the car forward-rollout consumer size): every vertex is a sample, the hull contains a
    random subset of all samples, and merging two half-hulls gives the whole hull (the multi-GPU reduction)."""
    from sampling_gpmpc_b200.engine import GPEngine, hull2d
    ns, nx, H1 = 1_000_000, 4, 51
    g = torch.Generator(device="cuda").manual_seed(3)
    traj = torch.randn(ns, nx, H1, generator=g, dtype=F64, device="cuda").cumsum(-1)
    eng = GPEngine(4, 1, 2, 3, 4)
    full = eng.stage_hulls(traj, 0, 1)
    a = eng.stage_hulls(traj[: ns // 2].contiguous(), 0, 1)
    b = eng.stage_hulls(traj[ns // 2:].contiguous(), 0, 1)
    sub = traj[torch.randint(ns, (2000,), device="cuda")].cpu().numpy()
    for t in (1, 25, 50):
        va = np.concatenate([a[t], b[t] + ns // 2])
        pts = traj[torch.as_tensor(va, device="cuda"), :2, t].cpu().numpy()
        merged = va[hull2d(pts)]
        assert _cyclic_equal(merged, full[t])
        hv = traj[torch.as_tensor(full[t].astype(np.int64), device="cuda"), :2, t].cpu().numpy()
        e = np.roll(hv, -1, 0) - hv
        p = sub[:, :2, t]
        cr = e[None, :, 0] * (p[:, None, 1] - hv[None, :, 1]) - e[None, :, 1] * (p[:, None, 0] - hv[None, :, 0])
        assert (cr >= -1e-9 * np.abs(hv).max()).all(), "a sample lies outside its stage's hull"
