"""CPU: the product's base-sample generation (sampling_gpmpc_b200.base_samples -> gpmpc_base_samples in the C-ABI library,
csrc/gpmpc_rng.cuh) is STREAM-IDENTICAL to the reference's loop (src/agent.py:76-104): same values bit for bit, same number
of candidates, and the torch generator is left in the same state.  Covers both of ATen's normal paths (candidates of
fewer than 16 scalars: scalar Box-Muller with the cached second sample; 16 or more: normal_fill, with and without the
recomputed tail) and the shapes of every yaml (g_ny * H * T = 3 car-fs, 8 pendulum rollout, 51 pendulum1D, 180 car, 450
car residual)."""
import numpy as np
import pytest
import torch

from sampling_gpmpc_b200.base_samples import truncated_normal_base_samples


def reference_loop(n_mpc, n_sqp, ns, g_ny, H, T, beta):
    """src/agent.py:76-104 verbatim in structure: one torch.normal per candidate, two torch.all, torch.cat."""
    ret_mpc = torch.empty(n_mpc, n_sqp, ns, g_ny, H, T, dtype=torch.float64)
    calls = 0
    for j in range(n_mpc):
        ret_itrs = torch.empty(n_sqp, ns, g_ny, H, T, dtype=torch.float64)
        for i in range(n_sqp):
            ret = torch.empty(0, g_ny, H, T, dtype=torch.float64)
            while True:
                w = torch.normal(0, 1, size=(1, g_ny, H, T), dtype=torch.float64)
                calls += 1
                if torch.all(w >= -beta) and torch.all(w <= beta):
                    ret = torch.cat([ret, w], dim=0)
                if ret.shape[0] == ns:
                    break
            ret_itrs[i] = ret
        ret_mpc[j] = ret_itrs
    return ret_mpc, calls


@pytest.mark.parametrize("g_ny,H,T,beta", [(3, 1, 1, 1.1), (2, 1, 4, 1.9), (1, 5, 3, 2.0), (1, 16, 1, 2.0), (1, 17, 3, 2.5),
                                            (2, 8, 4, 2.6), (3, 15, 4, 3.0), (3, 50, 3, 3.0)])
def test_product_base_samples_are_stream_identical_to_the_reference_loop(g_ny, H, T, beta):
    n_mpc, n_sqp, ns = 3, 2, 9
    torch.manual_seed(123456)  # experiment.rnd_seed of every yaml (main.py:41-42)
    torch.randn(3)             # something drawn before, so the generator carries a cached normal sample
    state0 = torch.get_rng_state()
    want, calls = reference_loop(n_mpc, n_sqp, ns, g_ny, H, T, beta)
    state_after_loop = torch.get_rng_state()
    next_after_loop = torch.rand(4)

    torch.set_rng_state(state0)
    got = truncated_normal_base_samples(n_mpc, n_sqp, ns, g_ny, H, T, beta)
    assert got.shape == want.shape and got.dtype == torch.float64
    assert torch.equal(got, want)                                   # bit for bit
    assert torch.equal(torch.get_rng_state(), state_after_loop)     # generator left where the loop leaves it
    assert torch.equal(torch.rand(4), next_after_loop)
    assert calls > n_mpc * n_sqp * ns                               # the test does exercise rejections
    assert float(got.abs().max()) <= beta


def test_base_samples_match_the_fixture_the_reference_agent_generated(golden_dir):
    """tests/golden/pendulum1D_sqp.npz holds the epistimic_random_vector the UNMODIFIED reference Agent drew from seed
    123456 (tests/golden/make_golden.py)."""
    import os
    import yaml
    z = np.load(os.path.join(golden_dir, "pendulum1D_sqp.npz"))
    params = yaml.safe_load(str(z["params_yaml"]))
    torch.manual_seed(params["experiment"]["rnd_seed"]["value"])
    ag, opt = params["agent"], params["optimizer"]
    eps = truncated_normal_base_samples(params["common"]["num_MPC_itrs"], opt["SEMPC"]["max_sqp_iter"], ag["num_dyn_samples"],
                                        ag["g_dim"]["ny"], opt["H"], 1 + ag["g_dim"]["nx"] + ag["g_dim"]["nu"], ag["Dyn_gp_beta"])
    assert np.array_equal(eps.numpy()[: z["eps"].shape[0]], z["eps"])


def test_bad_arguments_are_rejected():
    with pytest.raises(RuntimeError):
        truncated_normal_base_samples(1, 1, 2, 1, 1, 1, -1.0)
