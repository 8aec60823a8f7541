"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerance (BASELINE.md / SURVEY.md section 7): |a-b| <= 1e-9 * max(|b|, s) with s = outputscale for
variances and sqrt(outputscale) for means, samples and Jacobians; jitter-ladder decisions must be identical.
"""
import json
import os

import numpy as np
import pytest
import torch

from tests.replay import CASES, load_case, n_calls, outputscales, replay, scaled_close

pytestmark = pytest.mark.gpu

RTOL = 1e-9
REPORT = {}


def _dump_report():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _make_agent(params, z):
    from sampling_gpmpc_b200.agent import Agent
    from sampling_gpmpc_b200.envs import make_env_spec
    spec = make_env_spec(params)
    return Agent(params, spec=spec, X_real=torch.tensor(z["X_real"]), Y_real=torch.tensor(z["Y_real"]),
                 epistimic_random_vector=torch.tensor(z["eps"]))


@pytest.mark.parametrize("case", CASES)
def test_agent_replay_matches_golden(case):
    """Same iterates as the fixture, through the product Agent (block kernels + assembly kernel)."""
    z, params = load_case(case)
    agent = _make_agent(params, z)
    os_ = outputscales(params)
    s_val = float(np.sqrt(os_.max()))
    worst = {"mean": 0.0, "variance": 0.0, "y_sample": 0.0, "gp_val": 0.0, "y_grad": 0.0, "u_grad": 0.0}

    def check(k, ag, res):
        gp_val, y_grad, u_grad = res
        xscale = max(1.0, float(np.abs(z[f"x_h_{k}"]).max()))
        mean = ag.model_i_call.mean.cpu().numpy()
        var = ag.model_i_call.variance.cpu().numpy()
        for j in range(mean.shape[1]):
            worst["mean"] = max(worst["mean"], scaled_close(mean[:, j], z[f"mean_{k}"][:, j], np.sqrt(os_[j]), RTOL))
            worst["variance"] = max(worst["variance"], scaled_close(var[:, j], z[f"variance_{k}"][:, j], os_[j], RTOL))
        if f"jitter_level_{k}" in z.files:
            assert np.array_equal(ag.model_i_call.jitter_level.cpu().numpy(), z[f"jitter_level_{k}"]), \
                f"jitter ladder decisions differ at call {k}"
            ys = ag.model_i_samples.cpu().numpy()
            for j in range(ys.shape[1]):
                worst["y_sample"] = max(worst["y_sample"], scaled_close(ys[:, j], z[f"y_sample_{k}"][:, j], np.sqrt(os_[j]), RTOL))
        worst["gp_val"] = max(worst["gp_val"], scaled_close(gp_val, z[f"gp_val_{k}"], s_val * xscale, RTOL))
        worst["y_grad"] = max(worst["y_grad"], scaled_close(y_grad, z[f"y_grad_{k}"], s_val * xscale, RTOL))
        worst["u_grad"] = max(worst["u_grad"], scaled_close(u_grad, z[f"u_grad_{k}"], s_val * xscale, RTOL))
        assert ag.engine.num_hallucinated == int(z[f"n_halluc_{k}"]) or ag._pending_reset

    replay(agent, z, params, on_call=check)
    status = agent.engine.status()
    REPORT[f"golden/{case}"] = dict(worst, status=status)
    _dump_report()
    assert status & ~0x301 == 0, f"engine status {status:#x}"
    for k, v in worst.items():
        assert v <= 1.0, f"{case}: {k} off by {v:.3g} x tolerance"
    Xh, Yh = agent.Hallcinated_X_train.cpu().numpy(), agent.Hallcinated_Y_train.cpu().numpy()
    assert Xh.shape == z["halluc_X_final"].shape
    np.testing.assert_array_equal(Xh, z["halluc_X_final"])  # inputs are copied, never recomputed: bit-exact
    assert np.array_equal(np.isnan(Yh), np.isnan(z["halluc_Y_final"]))


def _rollout_problem(ns, steps, seed, T3=True):
    """car-residual rollout shapes: m=45 shared, d=2, g_ny=3, T=3 or 1."""
    from sampling_gpmpc_b200 import configs
    params = configs.car_residual_fs(num_dyn_samples=ns, steps=steps, with_derivatives=T3)
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(steps, ns, 3, 1, 3 if T3 else 1, generator=g, dtype=torch.float64).clamp(-3, 3)
    u = torch.stack([0.3 * torch.sin(torch.linspace(0, 6, steps)), 0.5 * torch.cos(torch.linspace(0, 4, steps))], 1).to(torch.float64)
    return params, eps, u


@pytest.mark.parametrize("T3", [True, False])
def test_fused_rollout_matches_oracle_refit(T3):
    """gpmpc_rollout (fused warp kernel, conditioning every step) vs the oracle's full re-fit per step."""
    from oracle.rollout_ref import reference_rollout
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps = 24, 12
    params, eps, u = _rollout_problem(ns, steps, 7, T3)
    fr = ForwardRollout(params, condition=T3)  # T=1 is the script as shipped: value-only model, no conditioning
    traj = fr.run(u, eps).cpu().numpy()
    status = fr.engine.status()
    # oracle: the reference loop (simulate_forward_sampling_car.py:117-138) with true conditioning
    ref = reference_rollout(params, fr.spec, u, eps, condition=T3)
    os_ = outputscales(params)
    ratio = scaled_close(traj, ref, float(np.sqrt(os_.max())) * 14.0, RTOL)
    REPORT[f"rollout/T{3 if T3 else 1}"] = dict(traj=ratio, status=status)
    _dump_report()
    assert status == 0
    assert ratio <= 1.0, f"trajectory off by {ratio:.3g} x tolerance"


def test_step_equals_posterior_plus_append():
    """The fused warp kernel and the general block kernels are two implementations of one recursion."""
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps = 16, 9
    params, eps, u = _rollout_problem(ns, steps, 11, True)
    a = ForwardRollout(params, condition=True)
    b = ForwardRollout(params, condition=True)
    g = torch.Generator().manual_seed(3)
    worst = 0.0
    for t in range(steps):
        x = (torch.rand(ns, 1, 1, 2, generator=g, dtype=torch.float64) - 0.5).expand(ns, 3, 1, 2).contiguous().cuda()
        x = x * torch.tensor([2.0, 1.2], dtype=torch.float64, device="cuda")
        e = eps[t].cuda()
        opts = a.engine.opts(beta=3.0)
        m1, v1, y1, j1 = a.engine.step(x, e, opts)
        m2, v2, y2, j2 = b.engine.posterior(x, e, opts)
        b.engine.append(x, y2)
        assert torch.equal(j1, j2)
        os_ = outputscales(params)
        for j in range(3):
            worst = max(worst, scaled_close(m1[:, j].cpu(), m2[:, j].cpu(), np.sqrt(os_[j]), RTOL),
                        scaled_close(v1[:, j].cpu(), v2[:, j].cpu(), os_[j], RTOL),
                        scaled_close(y1[:, j].cpu(), y2[:, j].cpu(), np.sqrt(os_[j]), RTOL))
    REPORT["step_vs_block"] = worst
    _dump_report()
    assert a.engine.num_factor_rows == b.engine.num_factor_rows == steps * 3
    assert worst <= 1.0


def test_rollout_properties_at_scale():
    """Size-independent checks at a size the oracle cannot run: permutation equivariance over samples
    (bit-exact sample indexing) and determinism, 4096 samples x 50 steps with conditioning."""
    from sampling_gpmpc_b200.rollout import ForwardRollout
    ns, steps = 4096, 50
    params, eps, u = _rollout_problem(ns, steps, 5, True)
    fr = ForwardRollout(params, condition=True)
    t1 = fr.run(u, eps)
    perm = torch.randperm(ns, generator=torch.Generator().manual_seed(1))
    fr2 = ForwardRollout(params, condition=True)
    t2 = fr2.run(u, eps[:, perm])
    assert torch.equal(t1[perm], t2), "sample s's trajectory must depend only on its own base samples"
    assert torch.isfinite(t1).all()
    assert fr.engine.status() == 0
