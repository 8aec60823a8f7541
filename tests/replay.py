"""Shared test helper: replay a golden fixture's SQP iterates through an Agent-like object.

Mirrors what src/solver.py:84-94 (SQP loop) and benchmarking/simulate_forward_sampling_car.py:118-130
(rollout loop) do around the hot path, for any object exposing the reference Agent's methods.
"""
import os

import numpy as np
import torch
import yaml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["pendulum1D_sqp", "car_residual_truedyn", "car_residual_sqp", "car_residual_sqp_jit", "car_residual_fs",
         "pendulum2D_sqp", "car_sqp"]
# car_residual_sqp runs at the yaml's Dyn_gp_jitter 1e-20: every draw goes through GPyTorch's eigen-root fallback, whose
# eigenvector signs no two eigh implementations share -- comparable call by call (modulo signs), not free-running
EIGEN_ROOT = ["car_residual_sqp"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = yaml.safe_load(str(z["params_yaml"]))
    return z, params


def n_calls(z):
    return len([k for k in z.files if k.startswith("gp_val_")])


def replay(agent, z, params, on_call=None):
    """Feed the fixture's iterates through ``agent``; returns the list of per-call outputs."""
    fs = params["env"]["use_model_without_derivatives"]
    n_sqp = params["optimizer"]["SEMPC"]["max_sqp_iter"] if not fs else 1
    outs = []
    for k in range(n_calls(z)):
        mpc, sqp = divmod(k, n_sqp)
        agent.mpc_iteration(mpc)
        x_h, u_h = z[f"x_h_{k}"], z[f"u_h_{k}"]
        if fs:
            agent.train_hallucinated_dynGP(1, use_model_without_derivatives=True)
            bx = agent.get_batch_x_hat(x_h, u_h)
            res = agent.dyn_fg_jacobians(bx, 1)
        else:
            agent.train_hallucinated_dynGP(sqp)
            bx = agent.get_batch_x_hat(x_h, u_h)
            res = agent.dyn_fg_jacobians(bx, sqp)
        outs.append(res)
        if on_call is not None:
            on_call(k, agent, res)
    return outs


def scaled_close(a, b, scale, rtol=1e-9):
    """|a-b| <= rtol*max(|b|, scale)  (BASELINE.md parity gate; scale = outputscale for variances,
    sqrt(outputscale) for means / samples / Jacobians).  Returns the worst ratio err/allowed."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    allowed = rtol * np.maximum(np.abs(b), scale)
    return float(np.max(np.abs(a - b) / allowed)) if a.size else 0.0


def outputscales(params):
    os_ = np.asarray(params["agent"]["Dyn_gp_outputscale"]["both"], dtype=np.float64).reshape(-1)
    return os_
