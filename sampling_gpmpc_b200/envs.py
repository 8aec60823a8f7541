"""Environment hooks used inside the GP-sampling hot path, as plain data.

The reference passes an ``env_model`` object whose methods are called from ``Agent.dyn_fg_jacobians``
(src/agent.py:532-564).  Only five of those hooks touch the path; for all four shipped systems they
reduce to index lists, two small constant matrices and (for the residual car) one multiply by the
velocity.  ``EnvSpec`` holds exactly that, so the CUDA assembly kernel can apply them without any
per-call Python:

  get_g_xu_hat            -> ``g_idx_inputs``   (pendulum1D.py:165-170, car_model_residual.py:132-137)
  get_f_known_jacobian /
  known_dyn               -> ``F_known``        f(x,u) = F_known @ [x;u]; the Jacobian is F_known itself
                                                (pendulum1D.py:137-188, car_model_residual.py:101-161,
                                                 car_model.py:101-161, pendulum.py:149-156 (f == 0))
  B_d                     -> ``B_d``            (pendulum1D.py:26-28, car_model_residual.py:26)
  pad_g                   -> ``pad_g``          (pendulum1D.py:15, car_model_residual.py:15)
  transform_sensitivity   -> ``transform``      identity, or [v g, v dg/dphi, g, v dg/ddelta]
                                                (car_model_residual.py:211-224)

``prior_data`` / ``initial_training_data`` restate the one-off training-set generators (setup, not
hot path) so synthetic workloads of the reference's shapes can be built where /root/reference is
absent (the GPU box).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch

TRANSFORM_IDENTITY = 0
TRANSFORM_CAR_RESIDUAL = 1


@dataclass
class EnvSpec:
    name: str
    nx: int
    nu: int
    g_ny: int
    g_nx: int
    g_nu: int
    g_idx_inputs: Tuple[int, ...]
    pad_g: Tuple[int, ...]
    B_d: np.ndarray  # (nx, g_ny)
    F_known: np.ndarray  # (nx, nx+nu)
    transform: int = TRANSFORM_IDENTITY
    dt: float = 0.0
    phys: dict = field(default_factory=dict)

    @property
    def d(self) -> int:
        return self.g_nx + self.g_nu

    # ---- unknown dynamics g and its analytic gradient (get_prior_data) -------------------------
    def prior_data(self, X: torch.Tensor) -> torch.Tensor:
        """(n, d) GP inputs -> (g_ny, n, 1+d): value and gradient of the unknown dynamics."""
        X = X.to(torch.float64)
        n = X.shape[0]
        dt = self.dt
        Y = torch.zeros(self.g_ny, n, 1 + self.d, dtype=torch.float64)
        if self.name == "Pendulum1D":  # pendulum1D.py:58-84,127-135
            l, g = self.phys["l"], self.phys["g"]
            th, u = X[:, 0], X[:, 1]
            Y[0, :, 0] = -g * torch.sin(th) * dt / l + u * dt
            Y[0, :, 1] = (-g * torch.cos(th) / l) * dt
            Y[0, :, 2] = dt
        elif self.name == "pendulum":  # pendulum.py:64-92,136-145
            l, g = self.phys["l"], self.phys["g"]
            x1, x2, u = X[:, 0], X[:, 1], X[:, 2]
            Y[0, :, 0] = x1 + x2 * dt
            Y[0, :, 1] = 1.0
            Y[0, :, 2] = dt
            Y[1, :, 0] = x2 - g * torch.sin(x1) * dt / l + u * dt / (l * l)
            Y[1, :, 1] = (-g * torch.cos(x1) / l) * dt
            Y[1, :, 2] = 1.0
            Y[1, :, 3] = dt / (l * l)
        elif self.name == "bicycle_Bdx":  # car_model_residual.py:62-99,167-182
            lf, lr = self.phys["lf"], self.phys["lr"]
            phi, delta = X[:, 0], X[:, 1]
            beta_in = (lr * torch.tan(delta)) / (lf + lr)
            beta = torch.atan(beta_in)
            term = ((lr / (torch.cos(delta) ** 2)) / (lf + lr)) / (1 + beta_in ** 2)
            Y[0, :, 0] = torch.cos(phi + beta) * dt
            Y[1, :, 0] = torch.sin(phi + beta) * dt
            Y[2, :, 0] = torch.sin(beta) * dt / lr
            Y[0, :, 1] = -torch.sin(phi + beta) * dt
            Y[0, :, 2] = -torch.sin(phi + beta) * dt * term
            Y[1, :, 1] = torch.cos(phi + beta) * dt
            Y[1, :, 2] = torch.cos(phi + beta) * dt * term
            Y[2, :, 2] = torch.cos(beta) * dt * term / lr
        elif self.name == "bicycle":  # car_model.py:62-99,163-178
            lf, lr = self.phys["lf"], self.phys["lr"]
            phi, v, delta = X[:, 0], X[:, 1], X[:, 2]
            beta_in = (lr * torch.tan(delta)) / (lf + lr)
            beta = torch.atan(beta_in)
            term = ((lr / (torch.cos(delta) ** 2)) / (lf + lr)) / (1 + beta_in ** 2)
            Y[0, :, 0] = v * torch.cos(phi + beta) * dt
            Y[1, :, 0] = v * torch.sin(phi + beta) * dt
            Y[2, :, 0] = v * torch.sin(beta) * dt / lr
            Y[0, :, 1] = -v * torch.sin(phi + beta) * dt
            Y[0, :, 2] = torch.cos(phi + beta) * dt
            Y[0, :, 3] = -v * torch.sin(phi + beta) * dt * term
            Y[1, :, 1] = v * torch.cos(phi + beta) * dt
            Y[1, :, 2] = torch.sin(phi + beta) * dt
            Y[1, :, 3] = v * torch.cos(phi + beta) * dt * term
            Y[2, :, 2] = torch.sin(beta) * dt / lr
            Y[2, :, 3] = v * torch.cos(beta) * dt * term / lr
        else:
            raise ValueError(f"no analytic prior data for env {self.name!r}")
        return Y

    def initial_training_data(self, params: dict):
        """Grid of real measurements (initial_training_data of each env; e.g. pendulum1D.py:30-56)."""
        opt, env = params["optimizer"], params["env"]
        nxd, nud = env["n_data_x"], env["n_data_u"]
        lin = lambda a, b, n: torch.linspace(a, b, n, dtype=torch.float64)
        if self.name == "Pendulum1D":
            axes = [lin(opt["x_min"][0], opt["x_max"][0], nxd), lin(opt["u_min"][0], opt["u_max"][0], nud)]
        elif self.name == "pendulum":
            axes = [lin(opt["x_min"][0], opt["x_max"][0], nxd), lin(opt["x_min"][1], opt["x_max"][1], nxd),
                    lin(opt["u_min"][0], opt["u_max"][0], nud)]
        elif self.name == "bicycle_Bdx":  # dphi = ddelta = 0 in the reference (car_model_residual.py:41-48)
            axes = [lin(opt["x_min"][2], opt["x_max"][2], nxd), lin(opt["u_min"][0], opt["u_max"][0], nud)]
        elif self.name == "bicycle":  # cell-centred grid (car_model.py:36-49)
            dphi = (opt["x_max"][2] - opt["x_min"][2]) / nxd
            dv = (opt["x_max"][3] - opt["x_min"][3]) / nxd
            dd = (opt["u_max"][0] - opt["u_min"][0]) / nud
            axes = [lin(opt["x_min"][2] + dphi / 2, opt["x_max"][2] - dphi / 2, nxd),
                    lin(opt["x_min"][3] + dv / 2, opt["x_max"][3] - dv / 2, nxd),
                    lin(opt["u_min"][0] + dd / 2, opt["u_max"][0] - dd / 2, nud)]
        else:
            raise ValueError(self.name)
        grids = torch.meshgrid(*axes, indexing="ij")
        X = torch.stack([g.reshape(-1) for g in grids], dim=1)
        Y = self.prior_data(X)
        if not env["train_data_has_derivatives"]:
            Y[:, :, 1:] = float("nan")
        return X, Y


def make_env_spec(params: dict) -> EnvSpec:
    """EnvSpec for the yaml's ``env.dynamics`` (main.py:13-16 name mapping)."""
    ag, env, opt = params["agent"], params["env"], params["optimizer"]
    nx, nu = ag["dim"]["nx"], ag["dim"]["nu"]
    g_ny, g_nx, g_nu = ag["g_dim"]["ny"], ag["g_dim"]["nx"], ag["g_dim"]["nu"]
    dt = float(opt["dt"])
    name = env["dynamics"]
    F = np.zeros((nx, nx + nu))
    phys = dict(env.get("params", {}))
    if name == "Pendulum1D":
        F[0, 0], F[0, 1], F[1, 1] = 1.0, dt, 1.0  # theta+ = theta + omega dt ; omega+ = omega
        B_d = np.array([[0.0], [1.0]])
        return EnvSpec(name, nx, nu, g_ny, g_nx, g_nu, (0, 2), (0, 1, 3), B_d, F, TRANSFORM_IDENTITY, dt, phys)
    if name == "pendulum":
        return EnvSpec(name, nx, nu, g_ny, g_nx, g_nu, (0, 1, 2), (0, 1, 2, 3), np.eye(nx, g_ny), F,
                       TRANSFORM_IDENTITY, dt, phys)
    if name in ("bicycle_Bdx", "bicycle"):
        F[0, 0] = F[1, 1] = F[2, 2] = F[3, 3] = 1.0
        F[3, 5] = dt  # V+ = V + acc dt
        if name == "bicycle_Bdx":
            return EnvSpec(name, nx, nu, g_ny, g_nx, g_nu, (2, 4), (0, 3, 4, 5), np.eye(nx, g_ny), F,
                           TRANSFORM_CAR_RESIDUAL, dt, phys)
        return EnvSpec(name, nx, nu, g_ny, g_nx, g_nu, (2, 3, 4), (0, 3, 4, 5), np.eye(nx, g_ny), F,
                       TRANSFORM_IDENTITY, dt, phys)
    raise ValueError(f"unknown env.dynamics {name!r}")


def env_spec_from_env_model(env_model, params: dict) -> EnvSpec:
    """Build the spec from a reference ``env_model`` instance (duck-typed), checking that its known
    dynamics really are linear by probing ``get_f_known_jacobian`` once."""
    spec = make_env_spec(params)
    assert tuple(env_model.g_idx_inputs) == spec.g_idx_inputs and tuple(env_model.pad_g) == spec.pad_g
    xu = torch.randn(2, spec.nx, 3, spec.nx + spec.nu, dtype=torch.float64)
    xu = xu[:, :1].expand(2, spec.nx, 3, spec.nx + spec.nu).contiguous()
    J = env_model.get_f_known_jacobian(xu).detach().cpu().to(torch.float64)
    F = torch.tensor(spec.F_known)
    assert torch.allclose(J[..., 1:], F[None, :, None, :].expand_as(J[..., 1:]))
    assert torch.allclose(J[..., 0], torch.einsum("ij,shj->sih", F, xu[:, 0].cpu()))
    assert torch.allclose(torch.as_tensor(env_model.B_d).detach().cpu().to(torch.float64), torch.tensor(spec.B_d))
    return spec
