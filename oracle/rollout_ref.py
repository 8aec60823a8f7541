"""CPU oracle for the forward-rollout callers (TEST INFRASTRUCTURE ONLY).

Restates the loop of benchmarking/simulate_forward_sampling_car.py:117-138 on top of oracle/agent_ref.py:
per step a full model re-fit (``train_hallucinated_dynGP``), ``dyn_fg_jacobians`` on the current states and
next state := gp_val.  ``condition=True`` is the iterative-conditioning variant
(benchmarking/simulate_true_reachable_set.py:179-259): the model is re-fitted on [real || hallucinated].
Parity unpinned for the GPyTorch arithmetic (see gp_ref.py); the loop itself follows the reference file.
"""
from __future__ import annotations

import numpy as np
import torch

from .agent_ref import RefAgent

F64 = torch.float64


def reference_rollout(params: dict, spec, u_ff: torch.Tensor, eps: torch.Tensor, condition: bool,
                      X_real=None, Y_real=None) -> np.ndarray:
    """u_ff (steps, nu), eps (steps, ns, g_ny, 1, T) -> X_traj (ns, nx, steps+1) float64."""
    ns, nx = params["agent"]["num_dyn_samples"], params["agent"]["dim"]["nx"]
    steps = u_ff.shape[0]
    if X_real is None:
        X_real, Y_real = spec.initial_training_data(params)
    # epistimic_random_vector[mpc_iter][sqp_iter]: the script reads sqp index 1 (:119,:124)
    erv = torch.zeros(steps, 2, *eps.shape[1:], dtype=F64)
    erv[:, 1] = eps
    agent = RefAgent(params, spec, X_real, Y_real, epistimic_random_vector=erv)
    fs = params["env"]["use_model_without_derivatives"]
    assert fs == (not condition) or not fs, "value-only model never conditions (agent.py:221-226)"
    fb = params["agent"].get("feedback", {}).get("use", False)
    K = np.asarray(params["optimizer"]["terminal_tightening"]["K"], dtype=np.float64)
    x_equi = np.asarray(params["env"]["goal_state"], dtype=np.float64)
    x_h = np.tile(np.asarray(params["env"]["start"], dtype=np.float64), (1, ns))
    X_traj = np.empty((ns, nx, steps + 1))
    u_np = u_ff.numpy()
    for t in range(steps):
        agent.train_hallucinated_dynGP(1, use_model_without_derivatives=fs)
        agent.mpc_iteration(t)
        u_h = u_np[t].reshape(1, -1)
        if fb:
            bx = agent.get_batch_x_hat_u_diff(x_h, -(x_equi - x_h.reshape(1, ns, -1)) @ K.T + np.tile(u_h[:, None, :], (ns, 1)))
        else:
            bx = agent.get_batch_x_hat(x_h, u_h)
        gp_val, _, _ = agent.dyn_fg_jacobians(bx, 1)
        X_traj[:, :, t] = bx[:, 0, 0, :nx].numpy()
        x_h = gp_val[:, :, 0, 0].reshape(1, -1)
    X_traj[:, :, steps] = gp_val[:, :, 0, 0]
    return X_traj
