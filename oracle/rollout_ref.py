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


def reference_true_reachable_set(params: dict, spec, u_ff: torch.Tensor, eps: torch.Tensor, group_size: int,
                                 X_real=None, Y_real=None, return_datasets: bool = False):
    """benchmarking/simulate_true_reachable_set.py:152-259 for ns_total / group_size repeats: every repeat is a NEW Agent of
    ``group_size`` samples (:172); per horizon step a full re-fit on [real || hallucinated] (:182), the posterior at the
    current state / input of every sample (:208), a joint draw (:209; here with the given base samples eps
    (steps, ns_total, g_ny, 1, T) instead of the script's internal torch.randn), the zero-variance rule (:213-228), the
    truncation (:230-236), ``update_hallucinated_Dyn_dataset`` WITH its min-distance filter (:239 -> agent.py:164-202),
    next state := sampled values (:251-258).  -> X_traj (ns_total, nx, steps+1)."""
    import copy
    ag = params["agent"]
    ns_total, steps = eps.shape[1], u_ff.shape[0]
    nx, nu, g_ny = ag["dim"]["nx"], ag["dim"]["nu"], ag["g_dim"]["ny"]
    assert g_ny == nx, "the script feeds the sampled values back as the next state"
    if X_real is None:
        X_real, Y_real = spec.initial_training_data(params)
    beta, var0 = ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"]
    X_traj = np.empty((ns_total, nx, steps + 1))
    datasets = []
    for g0 in range(0, ns_total, group_size):
        gs = min(group_size, ns_total - g0)
        p = copy.deepcopy(params)
        p["agent"]["num_dyn_samples"] = gs
        agent = RefAgent(p, spec, X_real, Y_real)
        x = torch.tensor(params["env"]["start"], dtype=F64).expand(gs, nx).clone()
        for i in range(steps):
            agent.train_hallucinated_dynGP(i)
            X_inp = torch.zeros(gs, g_ny, 1, nx + nu, dtype=F64)
            X_inp[:, :, 0, :nx] = x[:, None, :]
            X_inp[:, :, 0, nx:] = u_ff[i]
            post = agent.model_i(X_inp)
            Y = post.sample(base_samples=eps[i, g0:g0 + gs])
            zero = (post.variance <= var0).all(dim=-1, keepdim=True).tile(1, 1, 1, nx + nu + 1)
            num = torch.zeros_like(post.variance)
            num[zero] = 1
            Y = num * post.mean + (1 - num) * Y
            Y = torch.max(Y, post.mean - beta * torch.sqrt(post.variance))
            Y = torch.min(Y, post.mean + beta * torch.sqrt(post.variance))
            agent.update_hallucinated_Dyn_dataset(X_inp, Y)
            X_traj[g0:g0 + gs, :, i] = x.numpy()
            x = Y[:, :, 0, 0].clone()
        X_traj[g0:g0 + gs, :, steps] = x.numpy()
        datasets.append((agent.Hallcinated_X_train.clone(), agent.Hallcinated_Y_train.clone()))
    return (X_traj, datasets) if return_datasets else X_traj
