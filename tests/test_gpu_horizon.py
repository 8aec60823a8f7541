"""GPU: the fused-horizon rollout (csrc/gpmpc_horizon.cuh: ONE launch for the whole conditioned rollout, a warp owns an
element for all steps, its factor re-read from L2) against the step-wise rollout (one gpmpc_step per horizon step).  Same
arithmetic in the same order => BIT-IDENTICAL trajectories and an identical handle state afterwards (factor rows, beta,
recorded points), for ragged sample counts, both car / pendulum-like shapes, any grouping / staggering of the persistent
grid.  The step-wise path itself is held to the oracle by tests/test_gpu_parity.py."""
import numpy as np
import warnings
import pytest
import torch

pytestmark = pytest.mark.gpu


def _car(ns, steps, seed):
    from sampling_gpmpc_b200 import configs
    params = configs.car_residual_fs(num_dyn_samples=ns, steps=steps, with_derivatives=True)
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(steps, ns, 3, 1, 3, generator=g, dtype=torch.float64).clamp(-3, 3)
    u = torch.stack([0.3 * torch.sin(torch.linspace(0, 6, steps)), 0.5 * torch.cos(torch.linspace(0, 4, steps))], 1).to(torch.float64)
    return params, eps, u


def _pendulum(ns, steps, seed):
    from sampling_gpmpc_b200 import configs
    params = configs.pendulum2D_rollout(num_dyn_samples=ns, steps=steps, min_data_dist=-1)  # (the per-Agent filter is step-wise only)
    params["env"]["train_data_has_derivatives"] = False  # m = 45 value observations: inv(L_oo) of both outputs fits in shared memory
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(steps, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
    u = (2.0 * torch.sin(torch.linspace(0, 3, steps, dtype=torch.float64))).reshape(steps, 1)
    return params, eps, u


def _run(params, eps, u, fused, **options):
    from sampling_gpmpc_b200.rollout import ForwardRollout
    fr = ForwardRollout(params, condition=True)
    fr.use_fused_horizon(bool(fused))
    for k, v in options.items():
        fr.engine.set_option(k, v)
    traj = fr.run(u, eps)
    torch.cuda.synchronize()
    return fr, traj


@pytest.mark.parametrize("make,ns,steps", [(_car, 1, 6), (_car, 7, 12), (_car, 333, 50), (_car, 5000, 50),
                                           (_pendulum, 5, 9), (_pendulum, 1200, 30)])
def test_fused_horizon_is_bit_identical_to_the_step_wise_rollout(make, ns, steps):
    params, eps, u = make(ns, steps, 11)
    fr_s, traj_s = _run(params, eps, u, fused=False)
    fr_f, traj_f = _run(params, eps, u, fused=True)
    launches_f, launches_s = fr_f.engine.launch_count, fr_s.engine.launch_count
    assert fr_s.engine.status() == 0 and fr_f.engine.status() == 0
    assert torch.isfinite(traj_s).all()
    assert torch.equal(traj_f, traj_s)
    assert launches_f <= 5 < launches_s  # K0 (2 launches) + ONE horizon launch + the row tables: it did take the fused path
    # the handle is left in the same state: recorded points, factor (through posterior calls that read every row)
    Xf, Yf = fr_f.engine.export_hallucinated()
    Xs, Ys = fr_s.engine.export_hallucinated()
    assert torch.equal(Xf, Xs) and torch.equal(Yf, Ys)
    assert fr_f.engine.num_factor_rows == fr_s.engine.num_factor_rows == steps * fr_f.T
    g = torch.Generator().manual_seed(3)
    d = fr_f.spec.d
    probe = (torch.rand(ns, 1, 3, d, generator=g, dtype=torch.float64) - 0.5).expand(ns, fr_f.spec.g_ny, 3, d).contiguous()
    pe = torch.randn(ns, fr_f.spec.g_ny, 3, fr_f.T, generator=g, dtype=torch.float64)
    for mma in (True, False):
        outs = []
        for fr in (fr_f, fr_s):
            fr.engine.set_block_kernels(mma)
            outs.append(fr.engine.posterior(probe, pe))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
    # and the step API continues on it
    x1 = probe[:, :, :1].contiguous()
    e1 = pe[:, :, :1].contiguous()
    for a, b in zip(fr_f.engine.step(x1, e1), fr_s.engine.step(x1, e1)):
        assert torch.equal(a, b)


def test_grouping_and_staggering_do_not_change_the_result():
    params, eps, u = _car(901, 30, 5)
    _, ref = _run(params, eps, u, fused=False)
    for opts in ({"hz_groups": 1}, {"hz_groups": 2, "hz_stagger_ns": 300000}, {"hz_stagger_ns": 0}, {"hz_stagger_ns": 2000000}):
        fr, traj = _run(params, eps, u, fused=True, **opts)
        assert torch.equal(traj, ref), opts
        assert fr.engine.status() == 0


def test_repeated_fused_rollouts_on_one_handle():
    """reset + rollout again on the same handle (stale rows of the previous rollout in the partially filled sub-panels)."""
    params, eps, u = _car(257, 20, 9)
    fr_s, ref = _run(params, eps, u, fused=False)
    fr, first = _run(params, eps, u, fused=True)
    g = torch.Generator().manual_seed(1)
    eps2 = torch.randn(eps.shape, generator=g, dtype=torch.float64).clamp(-3, 3)
    second = fr.run(u, eps2).clone()
    third = fr.run(u, eps)
    assert torch.equal(first, ref) and torch.equal(third, ref) and not torch.equal(second, ref)
    assert torch.equal(fr_s.run(u, eps2), second)


def test_shapes_the_fused_kernel_does_not_serve_fall_back_to_the_step_wise_path():
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.rollout import ForwardRollout
    params = configs.pendulum2D_rollout(num_dyn_samples=64, steps=8)  # m = 180: inv(L_oo) does not fit next to the warps
    g = torch.Generator().manual_seed(2)
    eps = torch.randn(8, 64, 2, 1, 4, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
    u = torch.zeros(8, 1, dtype=torch.float64)
    fr = ForwardRollout(params, condition=True)
    traj = fr.run(u, eps)
    assert torch.isfinite(traj).all() and fr.engine.launch_count > 16


def test_failed_ladder_in_the_fused_kernel_is_redone_step_wise_with_the_eigen_root():
    """A draw that fails its (no-op) jitter ladder inside the one-launch kernel: check() repeats the rollout on the step-wise
    path, whose in-stream redo takes GPyTorch's batch-wide eigen root.  Trajectories, status word and the error raised equal
    those of a rollout that ran step-wise from the start (here: the eigen-root draw succeeds, conditioning on the noise-free
    singular point then fails like the reference's re-fit would)."""
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.engine import ST_SAMPLE_EIG
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps = 9, 6
    params = configs.pendulum2D_rollout(num_dyn_samples=ns, steps=steps, min_data_dist=-1)
    params["env"]["train_data_has_derivatives"] = True
    params["agent"]["Dyn_gp_jitter"] = 0.0
    params["agent"]["Dyn_gp_noise"] = 1e-30
    params["agent"]["Dyn_gp_task_noises"]["multiplier"] = 1e-30
    params["agent"]["Dyn_gp_variance_is_zero"] = -1
    # 2 far-apart noise-free real points; the start state [0, 0] with u_0 = 0 sits exactly on the first one: the 4 x 4
    # posterior covariance there is rounding noise around 0
    X = torch.tensor([[0.0, 0.0, 0.0], [1.5, -1.0, 4.0]], dtype=torch.float64)
    Y = torch.zeros(2, 2, 4, dtype=torch.float64)
    Y[:, :, 0] = torch.tensor([[0.1, -0.2], [0.3, 0.05]])
    g = torch.Generator().manual_seed(6)
    eps = torch.randn(steps, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
    u = torch.zeros(steps, 1, dtype=torch.float64)
    u[1:, 0] = torch.linspace(0.5, 2.0, steps - 1)

    def outcome(fr):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            try:
                return ("status", fr.check()), [str(x.message)[:40] for x in w]
            except Exception as e:  # noqa: BLE001  (the TYPE of the error is what is compared)
                return ("raised", type(e).__name__), [str(x.message)[:40] for x in w]

    ref = ForwardRollout(params, condition=True, X_real=X, Y_real=Y)
    ref.use_fused_horizon(False)
    l0 = ref.engine.launch_count
    want = ref.run(u, eps).clone()
    launches_ref = ref.engine.launch_count - l0
    st_ref = ref.engine.status()
    if not st_ref & ST_SAMPLE_EIG:
        pytest.skip("the singular first step did not break the Cholesky on this build")
    want_outcome = outcome(ref)
    for mode in (True, "auto"):
        fr = ForwardRollout(params, condition=True, X_real=X, Y_real=Y)
        fr.use_fused_horizon(mode)
        l0 = fr.engine.launch_count
        got = fr.run(u, eps)
        # the one-launch kernel ran first (horizon + row tables), then the step-wise repeat
        assert fr.engine.launch_count - l0 == launches_ref + 2
        assert fr.engine.get_option("last_rollout_fused") == 0
        assert fr.engine.status() == st_ref
        assert outcome(fr) == want_outcome
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
    assert not torch.isnan(want[:, :, 1]).any()  # the eigen-root draw of the singular step itself is finite


def test_automatic_choice_is_the_one_launch_kernel_for_small_batches_only():
    """ForwardRollout's default: one launch while a warp of the fused kernel gets at most 4 samples, step-wise beyond."""
    from sampling_gpmpc_b200.rollout import ForwardRollout
    for ns, fused in ((40, 1), (2300, 1), (2400, 0)):
        params, eps, u = _car(ns, 6, 3)
        fr = ForwardRollout(params, condition=True)
        l0 = fr.engine.launch_count
        traj = fr.run(u, eps)
        assert fr.engine.get_option("last_rollout_fused") == fused, ns
        assert (fr.engine.launch_count - l0 <= 2) == bool(fused)
        fr2, ref = _run(params, eps, u, fused=False)
        assert torch.equal(traj, ref) and fr.engine.status() == 0
