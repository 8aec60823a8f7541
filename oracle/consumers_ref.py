"""CPU oracle for the formats either side of the hot path (TEST INFRASTRUCTURE ONLY).

Restates, loop for loop, what the reference does with the hot path's outputs:
  * ``pack_p_lin``      -- src/solver.py:98-131 (the per-stage acados parameter vector, sample by sample);
  * ``min_dist_*``      -- src/agent.py:666-708 and :166-191 as free functions on explicit arrays;
  * ``traj_stats``      -- extra/approx_sampling_mpc/README.md:19-27 (Delta_k = max_n |x_k^n - x_k^mu|) + the box;
  * ``stage_hulls``     -- benchmarking/generate_convex_hull.py:88-100 (scipy.spatial.ConvexHull per stage, the same
                           third-party routine the reference calls: scipy/qhull, importable here and on the GPU box).
Pinned: p_lin by construction from the golden fixtures' gp_val / y_grad / u_grad (tests/test_oracle_golden.py);
the hull against qhull itself.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch


def pack_p_lin(gp_val, y_grad, u_grad, x_h, u_h, xg, w, tilde_eps_list, ns, nx, K=None):
    """solver.py:84-131.  gp_val (ns,nx,H,1), y_grad (ns,nx,H,nx), u_grad (ns,nx,H,nu), x_h (H, ns*nx), u_h (H,nu),
    xg (H,·), w (H,·), tilde_eps_list[stage] -> list over stages of p_lin vectors."""
    H = x_h.shape[0]
    if K is not None:
        y_grad = y_grad + u_grad @ K  # solver.py:90
    out = []
    for stage in range(H):
        p_lin = np.empty(0)
        for i in range(ns):
            p_lin = np.concatenate([
                p_lin,
                y_grad[i, :, stage, :].reshape(-1),
                u_grad[i, :, stage, :].reshape(-1),
                x_h[stage, i * nx: nx * (i + 1)],
                gp_val[i, :, stage, :].reshape(-1),
            ])
        p_lin = np.hstack([p_lin, u_h[stage], xg[stage], w[stage], tilde_eps_list[stage]])
        out.append(p_lin)
    return out


def min_dist_overwrite(x_input, x_train, y_train, y_sample, mean, variance, min_distance, beta):
    """src/agent.py:666-708 on explicit tensors: x_input (ns,g_ny,H,d), x_train (ns,g_ny,n,d), y_train (ns,g_ny,n,T)."""
    H, T = x_input.shape[2], y_train.shape[3]
    dist = x_input[:, :, None, :, :] - x_train[:, :, :, None, :]
    y_train_isnan = torch.any(torch.isnan(y_train), dim=3).unsqueeze(-1).tile(1, 1, 1, H)
    dist_norm = torch.linalg.vector_norm(dist, dim=-1)
    dist_norm[y_train_isnan] = torch.tensor(float("inf"))
    dist_too_small = torch.any(dist_norm <= min_distance, dim=2).unsqueeze(-1).tile(1, 1, 1, T)
    _, idx = torch.min(dist_norm, dim=2)
    A, B, _, _ = y_train.shape
    E = idx.shape[2]
    i1 = torch.arange(A).view(A, 1, 1).expand(A, B, E)
    i2 = torch.arange(B).view(1, B, 1).expand(A, B, E)
    closest = y_train[i1, i2, idx, :]
    y = torch.where(dist_too_small, closest, y_sample)
    y_max = mean + beta * torch.sqrt(variance)
    y_min = mean - beta * torch.sqrt(variance)
    return torch.min(torch.max(y, y_min), y_max)


def filter_new_points(newX, newY, X_cond, min_distance):
    """src/agent.py:166-191: returns (newY with NaN labels, filter_these_out (ns,g_ny,H), filter_these_out_all (H,))."""
    dist = newX[:, :, None, :, :] - X_cond[:, :, :, None, :]
    dist_norm = torch.linalg.vector_norm(dist, dim=-1)
    filt = torch.any(dist_norm <= min_distance, dim=2)
    newY = newY.clone()
    newY[filt.unsqueeze(-1).tile(1, 1, 1, newY.shape[-1])] = torch.nan
    f_all = torch.any(torch.all(filt, dim=0), dim=0)
    return newY, filt, f_all


def traj_stats(X_traj: np.ndarray, x_mu: np.ndarray):
    """X_traj (ns,nx,H1), x_mu (nx,H1) -> box_min, box_max, Delta = max_n |x^n - x^mu| (all (nx,H1))."""
    return X_traj.min(0), X_traj.max(0), np.abs(X_traj - x_mu[None]).max(0)


def stage_hulls(X_traj: np.ndarray, i0: int = 0, i1: int = 1):
    """generate_convex_hull.py:88-100: per stage the hull vertices (sample indices, qhull's CCW order) of the
    (i0, i1) point cloud.  Stages whose cloud is degenerate for qhull (< 3 distinct points) return the distinct
    points' lowest indices."""
    from scipy.spatial import ConvexHull
    from scipy.spatial import QhullError
    out = []
    for t in range(X_traj.shape[2]):
        pts = X_traj[:, [i0, i1], t]
        try:
            out.append(np.asarray(ConvexHull(pts).vertices, dtype=np.int64))
        except (QhullError, ValueError):
            _, first = np.unique(pts, axis=0, return_index=True)
            out.append(np.sort(first))
    return out
