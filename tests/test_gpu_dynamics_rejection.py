"""GPU: the rejection rollout of the reference Agent -- train_forward_sampling_dynGP / prepare_dynamics_set
(src/agent.py:283-443, switched on by common.dynamics_rejection): forward sampling from the measured state with value-only
conditioning (:365-415), the sample-survival test (:351-394) and the resampling of rejected samples from survivors
(:418-436) -- product Agent (C ABI: gpmpc_posterior, gpmpc_fs_advance, gpmpc_append_masked, gpmpc_truncate_hallucinated)
against the oracle's restatement, on the pendulum1D closed-loop shape with the yaml's tightening constants."""
import copy

import numpy as np
import pytest
import torch

from tests.replay import load_case, outputscales, scaled_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _params():
    z, params = load_case("pendulum1D_sqp")
    params = copy.deepcopy(params)
    params["agent"]["num_dyn_samples"] = 9
    params["optimizer"]["H"] = 8
    # params_pendulum1D_samples.yaml:47-51, :101-102
    params["agent"]["tight"] = {"use": True, "dyn_eps": 0.002, "Lipschitz": 0.96, "w_bound": 0.0001}
    params["optimizer"]["terminal_tightening"]["P"] = [[10.47241433, 0.2680862], [0.2680862, 8.74083638]]
    return z, params


@pytest.mark.parametrize("loosen", [1.0, 40.0])
def test_prepare_dynamics_set_matches_the_oracle(loosen):
    from oracle.agent_ref import RefAgent
    from sampling_gpmpc_b200.agent import Agent, reachable_set_ball
    from sampling_gpmpc_b200.envs import make_env_spec
    z, params = _params()
    spec = make_env_spec(params)
    ns, H, nx = 9, 8, 2
    X = torch.tensor(z["X_real"])
    Y = torch.tensor(z["Y_real"])
    g = torch.Generator().manual_seed(7)
    eps = torch.randn(2, 1, ns, 1, H, 3, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
    gpu = Agent(params, spec=spec, X_real=X, Y_real=Y, epistimic_random_vector=eps)
    ref = RefAgent(params, spec, X, Y, epistimic_random_vector=eps)
    _, ci = reachable_set_ball(params, np.ones(H + 1))
    ci = [c * loosen for c in ci]  # loosen = 40: some samples survive, some are rejected
    gpu.ci_list, ref.ci_list = ci, ci
    # one SQP solve: iterates spread over the samples
    rng = np.random.default_rng(3)
    x_h = np.tile(np.stack([np.linspace(2.3, 3.0, H), np.linspace(1.5, 0.2, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, nx * ns))
    u_h = np.linspace(-2, 2, H).reshape(H, 1)
    # one SQP iteration under teacher forcing: both condition on the ORACLE's draw (the joint 24 x 24 covariance of a smooth
    # iterate is numerically singular, lambda_min ~ 1e-16, and factorises without jitter on rounding noise: the draw itself
    # is not reproducible between any two implementations -- tests/test_gpu_parity.py deals with that; here the subject is
    # the rejection rollout)
    for a in (gpu, ref):
        a.mpc_iteration(0)
        a.train_hallucinated_dynGP(0)
    g_xu = ref.get_g_xu_hat(ref.get_batch_x_hat(x_h, u_h)).contiguous()
    y_ref = ref.sample_gp(g_xu, eps[0][0])
    gpu.sample_gp(g_xu.cuda(), eps[0][0].cuda())
    ref.update_hallucinated_Dyn_dataset(g_xu, y_ref)
    gpu.update_hallucinated_Dyn_dataset(g_xu.cuda(), y_ref.cuda())
    s = float(np.sqrt(outputscales(params).max()))
    # the solver's solution: every sample's predicted state sequence; the measurement is close to sample 0's x_1
    X_soln = np.concatenate([x_h, x_h[-1:] + 0.01], 0)  # (H+1, ns*nx)
    U_soln = u_h
    spread = 0.004 * loosen * rng.standard_normal((H + 1, ns * nx)) * (np.arange(ns * nx) // nx > 4)  # samples 5.. drift away
    X_soln = X_soln + spread
    X_kp1 = X_soln[1, :nx].reshape(nx, 1)
    base = [torch.randn(ns, 1, 1, 3, generator=g, dtype=torch.float64) for _ in range(H)]
    np.random.seed(11)
    left_g = gpu.prepare_dynamics_set(X_soln, U_soln, X_kp1, base_samples=base).cpu().numpy()
    np.random.seed(11)
    left_r = ref.prepare_dynamics_set(X_soln, U_soln, X_kp1, base_samples=base).numpy()
    assert np.array_equal(left_g, left_r)
    if loosen > 1:
        assert 0 < left_r.sum() < ns, "want both survivors and rejected samples: tune the test"
    os_ = outputscales(params)
    s = float(np.sqrt(os_.max()))
    Xg, Yg = gpu.Hallcinated_X_train.cpu().numpy(), gpu.Hallcinated_Y_train.cpu().numpy()
    Xr, Yr = ref.Hallcinated_X_train.numpy(), ref.Hallcinated_Y_train.numpy()
    assert Xg.shape == Xr.shape and np.array_equal(Xg, Xr)  # inputs are copied and permuted, never recomputed
    assert scaled_close(Yg, Yr, s, RTOL) <= 1.0
    assert gpu.engine.num_hallucinated == Xr.shape[2]
    # the restored model: the posterior at the next iterate agrees
    x_h2 = x_h + 0.003 * rng.standard_normal(x_h.shape)
    for a in (gpu, ref):
        a.mpc_iteration(1)
        a.train_hallucinated_dynGP(0)
    g2 = ref.get_g_xu_hat(ref.get_batch_x_hat(x_h2, u_h)).contiguous()
    post = ref.model_i(g2)
    mean, var = gpu.engine.posterior(g2.cuda())
    assert scaled_close(mean.cpu().numpy(), post.mean.numpy(), s, RTOL) <= 1.0
    assert scaled_close(var.cpu().numpy(), post.variance.numpy(), float(os_.max()), RTOL) <= 1.0
    assert gpu.engine.status() & ~0x301 == 0
