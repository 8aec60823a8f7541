"""CPU: the oracle restatements of the rows either side of the hot path (oracle/consumers_ref.py) and the library's
host-only hull helper (gpmpc_hull2d has no device work, so it runs here)."""
import os

import numpy as np
import pytest
import torch

from oracle import consumers_ref as ref


def test_p_lin_layout_matches_the_acados_model_unpacking(golden_dir):
    """src/utils/model.py:34-41 reads p_lin back as [A_i (nx*nx) | B_i (nx*nu) | x_lin_i (nx) | f_i (nx)] per sample,
    then u_lin, xg, w, tilde_eps.  Build p_lin from a golden fixture's linearisation with the reference's loop and
    unpack it with that rule."""
    z = np.load(f"{golden_dir}/pendulum1D_sqp.npz")
    gp_val, y_grad, u_grad, x_h, u_h = z["gp_val_0"], z["y_grad_0"], z["u_grad_0"], z["x_h_0"], z["u_h_0"]
    ns, nx, H, nu = gp_val.shape[0], gp_val.shape[1], gp_val.shape[2], u_grad.shape[3]
    xg, w = np.arange(H * 2, dtype=float).reshape(H, 2), np.zeros((H, nx))
    te = [np.array([0.1 * t]) for t in range(H)]
    p = ref.pack_p_lin(gp_val, y_grad, u_grad, x_h[:H], u_h, xg, w, te, ns, nx)
    per = nx * nx + nx * nu + 2 * nx
    assert len(p) == H and all(v.shape == (ns * per + nu + 2 + nx + 1,) for v in p)
    for stage in (0, H - 1):
        for i in (0, ns - 1):
            blk = p[stage][i * per:(i + 1) * per]
            assert np.array_equal(blk[:nx * nx].reshape(nx, nx), y_grad[i, :, stage, :])
            assert np.array_equal(blk[nx * nx:nx * nx + nx * nu].reshape(nx, nu), u_grad[i, :, stage, :])
            assert np.array_equal(blk[nx * nx + nx * nu:nx * nx + nx * nu + nx], x_h[stage, i * nx:(i + 1) * nx])
            assert np.array_equal(blk[-nx:], gp_val[i, :, stage, 0])
        assert np.array_equal(p[stage][ns * per:ns * per + nu], u_h[stage])
        assert p[stage][-1] == 0.1 * stage


def test_min_dist_oracle_on_a_hand_case():
    """agent.py:666-708 semantics on a case small enough to read: NaN-target points never match, the nearest
    fully observed one does, the result is clipped."""
    x_train = torch.tensor([[[[0.0, 0.0], [1.0, 0.0], [1.0, 1e-6]]]])
    y_train = torch.tensor([[[[5.0, float("nan")], [7.0, 8.0], [9.0, 10.0]]]])
    x = torch.tensor([[[[0.0, 0.0], [1.0, 4e-7], [3.0, 3.0]]]])
    y = torch.full((1, 1, 3, 2), 0.25)
    mean, var = torch.zeros(1, 1, 3, 2), torch.full((1, 1, 3, 2), 4.0)
    out = ref.min_dist_overwrite(x, x_train, y_train, y, mean, var, 1e-5, 2.5)
    assert out[0, 0, 0].tolist() == [0.25, 0.25]   # duplicate of a point with a NaN target: not overwritten
    assert out[0, 0, 1].tolist() == [5.0, 5.0]     # nearest fully observed: [7, 8] clipped to 0 + 2.5*2
    assert out[0, 0, 2].tolist() == [0.25, 0.25]
    newY, filt, f_all = ref.filter_new_points(x, y, x_train, 1e-5)
    assert filt[0, 0].tolist() == [True, True, False] and f_all.tolist() == [True, True, False]
    assert torch.isnan(newY[0, 0, :2]).all() and not torch.isnan(newY[0, 0, 2]).any()


@pytest.fixture(scope="module")
def hull2d():
    import __graft_entry__ as g
    g.build()
    from sampling_gpmpc_b200.engine import hull2d
    return hull2d


@pytest.mark.parametrize("n,kind", [(3, "gauss"), (10, "gauss"), (1000, "gauss"), (20000, "gauss"), (5000, "disc"), (400, "grid")])
def test_hull2d_matches_qhull(hull2d, n, kind):
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(n)
    if kind == "gauss":
        p = rng.normal(size=(n, 2))
    elif kind == "disc":
        a, r = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n))
        p = np.stack([r * np.cos(a), r * np.sin(a)], 1)
    else:  # integer grid: many collinear boundary points, which are NOT vertices (qhull's default, strict turns here)
        g = np.arange(int(np.sqrt(n)), dtype=float)
        p = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)
    got, want = list(hull2d(p)), list(ConvexHull(p).vertices)
    assert sorted(got) == sorted(want)
    k = want.index(got[0])
    assert got == want[k:] + want[:k], "not the same counter-clockwise cycle"


def test_hull2d_degenerate(hull2d):
    assert list(hull2d(np.array([[0.0, 0], [1, 1], [2, 2], [0.5, 0.5]]))) == [0, 2]
    assert list(hull2d(np.array([[1.0, 1], [1, 1], [1, 1]]))) == [0]
    assert list(hull2d(np.zeros((0, 2)))) == []


def test_traj_stats_and_hull_oracle_shapes():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(50, 4, 6)).cumsum(-1)
    lo, hi, dev = ref.traj_stats(X, X.mean(0))
    assert lo.shape == hi.shape == dev.shape == (4, 6) and (dev >= 0).all() and (lo <= hi).all()
    hulls = ref.stage_hulls(X)
    assert len(hulls) == 6 and all(len(h) >= 3 for h in hulls)


def test_save_X_traj_matches_the_reference_pickle_format(tmp_path):
    """data_X_traj_<k>.pkl (simulate_forward_sampling_car.py:157-161) as generate_convex_hull.py:77-84 reads it back."""
    import pickle
    from sampling_gpmpc_b200.rollout import save_X_traj
    traj = torch.arange(10 * 4 * 6, dtype=torch.float64).reshape(10, 4, 6)
    (p,) = save_X_traj(traj, str(tmp_path), 3)
    assert p.endswith("data_X_traj_3.pkl")
    with open(p, "rb") as f:
        back = pickle.load(f)
    assert isinstance(back, np.ndarray) and back.dtype == np.float64 and np.array_equal(back, traj.numpy())
    paths = save_X_traj(traj, str(tmp_path / "jobs"), 0, chunk=4)
    assert [os.path.basename(q) for q in paths] == ["data_X_traj_0.pkl", "data_X_traj_1.pkl", "data_X_traj_2.pkl"]
    parts = [pickle.load(open(q, "rb")) for q in paths]
    assert np.array_equal(np.vstack(parts), traj.numpy())  # the consumer's np.vstack over the job files


def test_reachable_set_ball_matches_the_reference_function():
    """sampling_gpmpc_b200.agent.reachable_set_ball restates src/utils/reachable_set.py:3-39 (host scalar math that
    prepare_dynamics_set reads its acceptance radii from); compared with the reference's own function where the checkout is
    mounted (CPU container; skipped on the GPU box)."""
    import contextlib
    import importlib.util
    import io
    import os
    import pytest
    ref_root = os.environ.get("GPMPC_REFERENCE", "/root/reference")
    path = os.path.join(ref_root, "src", "utils", "reachable_set.py")
    if not os.path.isfile(path):
        pytest.skip("reference checkout not mounted")
    spec = importlib.util.spec_from_file_location("ref_reachable_set", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pytest.importorskip("torch")  # agent.py imports torch at module level
    from sampling_gpmpc_b200.agent import reachable_set_ball
    params = {"optimizer": {"H": 17, "terminal_tightening": {"P": [[10.47241433, 0.2680862], [0.2680862, 8.74083638]],
                                                              "K": [[-18.82703934, -7.32095004]]}},
              "agent": {"tight": {"Lipschitz": 0.96, "dyn_eps": 0.002, "w_bound": 0.0001}}}
    V = np.linspace(1.0, 0.5, 18)
    with contextlib.redirect_stdout(io.StringIO()):
        want_eps, want_ci = mod.get_reachable_set_ball(params, V)
    got_eps, got_ci = reachable_set_ball(params, V)
    assert np.array_equal(np.asarray(got_ci), np.asarray(want_ci))
    assert len(got_eps) == len(want_eps) and all(np.array_equal(a, b) for a, b in zip(got_eps, want_eps))


@pytest.mark.parametrize("kind", ["gauss", "grid", "same"])
def test_merged_shard_hulls_equal_the_hull_of_all_points(hull2d, kind):
    """ForwardRollout.stage_hulls on several ranks: every shard's own hull, the vertices merged by rollout.merge_shard_hulls --
    the same vertex list (global sample indices) as the hull over all samples, also with points repeated across shards (a grid;
    stage 0 of a rollout, where every sample sits on the start state)."""
    from sampling_gpmpc_b200.rollout import merge_shard_hulls, shard_bounds
    rng = np.random.default_rng(4)
    n, world = 1001, 3
    if kind == "gauss":
        p = rng.standard_normal((n, 2))
    elif kind == "grid":
        p = rng.integers(0, 7, size=(n, 2)).astype(np.float64)
    else:
        p = np.tile(np.array([[0.3, -1.2]]), (n, 1))
    want = list(hull2d(p))
    cands = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        local = hull2d(p[lo:hi])                       # positions inside the shard
        c = np.full((16, 3), np.nan)                   # padded like the all-gathered buffer
        c[: len(local), :2] = p[lo:hi][local]
        c[: len(local), 2] = local + lo
        cands.append(c)
    got = list(merge_shard_hulls(np.concatenate(cands[::-1])))  # any rank order
    assert got == want
