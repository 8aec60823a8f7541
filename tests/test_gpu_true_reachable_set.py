"""GPU: the pendulum true-reachable-set rollout AS CONFIGURED (benchmarking/simulate_true_reachable_set.py:152-259 on
params_pendulum.yaml: Dyn_gp_variance_is_zero 1.1e-6 and Dyn_gp_min_data_dist 1e-4 both on).  The script builds a new
Agent of num_dyn_samples samples per repeat and update_hallucinated_Dyn_dataset (src/agent.py:164-202) couples the samples
of ONE Agent (NaN labels -> GPyTorch's any-over-batch mask; drop iff filtered for all samples of an output); the product
rolls every repeat out in one batch with those reductions per group (gpmpc_set_grouping).  Compared with the oracle's
restatement of the script, at the yaml's threshold and at thresholds where the filter actually fires."""
import numpy as np
import pytest
import torch

from tests.replay import outputscales, scaled_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _problem(ns, steps, min_dist, seed=4, derivatives=True):
    from sampling_gpmpc_b200 import configs
    params = configs.pendulum2D_rollout(num_dyn_samples=ns, steps=steps, min_data_dist=min_dist)
    params["env"]["train_data_has_derivatives"] = derivatives
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(steps, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
    u = (2.0 * torch.sin(torch.linspace(0, 3, steps, dtype=torch.float64))).reshape(steps, 1)
    return params, eps, u


@pytest.mark.parametrize("min_dist,derivatives", [(1.0e-4, True), (0.02, True), (0.05, False), (0.3, True)])
def test_grouped_rollout_matches_the_oracle_of_the_script(min_dist, derivatives):
    from oracle.rollout_ref import reference_true_reachable_set
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps, agent = 12, 10, 4  # three reference Agents of four samples each
    params, eps, u = _problem(ns, steps, min_dist, derivatives=derivatives)
    fr = ForwardRollout(params, condition=True, agent_size=agent)
    traj = fr.run(u, eps).cpu().numpy()
    assert fr.engine.status() == 0
    ref, datasets = reference_true_reachable_set(params, fr.spec, u, eps, agent, return_datasets=True)
    ratio = scaled_close(traj, ref, float(np.sqrt(outputscales(params).max())), RTOL)
    assert ratio <= 1.0, f"trajectory off by {ratio:.3g} x tolerance"
    # the data sets the Agents end up with: dropped points absent, filtered labels NaN, everything else as sampled
    Xh, Yh = fr.engine.export_hallucinated()
    ps = fr.engine.export_point_states().cpu().numpy()
    Xh, Yh = Xh.cpu().numpy(), Yh.cpu().numpy()
    n_masked = n_dropped = n_nan = 0
    for gi, (Xo, Yo) in enumerate(datasets):
        sl = slice(gi * agent, (gi + 1) * agent)
        state = ps[sl]
        assert (state == state[:1, :1]).all()  # one decision per (Agent, point)
        keep = state[0, 0] != 2
        n_dropped += int((~keep).sum())
        n_masked += int((state[0, 0] == 1).sum())
        Xg, Yg = Xh[sl][:, :, keep], Yh[sl][:, :, keep]
        assert Xg.shape == tuple(Xo.shape), (Xg.shape, Xo.shape)
        assert scaled_close(Xg, Xo.numpy(), 1.0, RTOL) <= 1.0
        nan_o = np.isnan(Yo.numpy())
        assert np.array_equal(np.isnan(Yg), nan_o)
        n_nan += int(nan_o[..., 0].sum())
        assert scaled_close(np.nan_to_num(Yg), np.nan_to_num(Yo.numpy()), float(np.sqrt(outputscales(params).max())), RTOL) <= 1.0
        # masked <=> some label of the Agent is NaN at that point
        assert np.array_equal(state[0, 0][keep] == 1, nan_o.any(axis=(0, 1, 3)))
    if min_dist >= 0.02:
        assert n_masked + n_dropped > 0, "the filter never fired: the test does not test it"
    if min_dist >= 0.3:
        assert n_dropped > 0
    print(f"min_dist {min_dist}: masked {n_masked}, dropped {n_dropped}, NaN labels {n_nan}, worst traj error {ratio:.3g} x tol")


def test_groups_are_independent_of_each_other():
    """An Agent's trajectories do not depend on which other Agents share the batch (bit for bit)."""
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps, agent = 40, 12, 4
    params, eps, u = _problem(ns, steps, 0.05)
    full = ForwardRollout(params, condition=True, agent_size=agent).run(u, eps)
    p2, _, _ = _problem(8, steps, 0.05)
    part = ForwardRollout(p2, condition=True, agent_size=agent).run(u, eps[:, 16:24].contiguous())
    assert torch.equal(part, full[16:24])


def test_block_entry_points_refuse_a_grouped_handle():
    from sampling_gpmpc_b200.engine import GPEngineError
    from sampling_gpmpc_b200.rollout import ForwardRollout
    params, eps, u = _problem(8, 3, 0.05)
    fr = ForwardRollout(params, condition=True, agent_size=4)
    fr.run(u, eps)
    x = torch.zeros(8, 2, 2, 3, dtype=torch.float64)
    with pytest.raises(GPEngineError, match="grouped"):
        fr.engine.posterior(x)
    with pytest.raises(GPEngineError, match="grouped"):
        fr.engine.append(x, torch.zeros(8, 2, 2, 4, dtype=torch.float64))
