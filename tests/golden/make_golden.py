"""Generate tests/golden/*.npz by running the UNMODIFIED reference Agent + env classes.

Needs /root/reference (read-only checkout of manish-pra/sampling-gpmpc).  gpytorch is not
installable here, so ``oracle.ref_gpytorch.install()`` provides the `gpytorch` symbols the reference
imports, backed by oracle/gp_ref.py; everything above that line -- src/agent.py, src/GP_model.py and
src/environments/*.py -- is the reference's own code, imported as is.  Each fixture stores the
inputs fed to the Agent (yaml overrides, SQP iterates, base samples) and what it returned.

    python tests/golden/make_golden.py            # rewrites the .npz files next to this script
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GPMPC_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
from oracle import ref_gpytorch  # noqa: E402

ref_gpytorch.install()
sys.path.insert(0, REF)
from src.agent import Agent  # noqa: E402
from src.environments.pendulum import Pendulum as pendulum  # noqa: E402
from src.environments.car_model_residual import CarKinematicsModel as bicycle_Bdx  # noqa: E402
from src.environments.car_model import CarKinematicsModel as bicycle  # noqa: E402
from src.environments.pendulum1D import Pendulum as Pendulum1D  # noqa: E402

ENVS = {"pendulum": pendulum, "bicycle_Bdx": bicycle_Bdx, "bicycle": bicycle, "Pendulum1D": Pendulum1D}

# fixture name -> (yaml, overrides (dotted), number of MPC steps, sqp iterations per step)
CASES = {
    "pendulum1D_sqp": ("params_pendulum1D_samples", {"agent.num_dyn_samples": 12, "common.num_MPC_itrs": 3}, 3, 1),
    # as shipped: one sample overwritten by the true dynamics, the model is still evaluated (agent.py:583-616)
    "car_residual_truedyn": ("params_car_residual", {"optimizer.H": 10, "optimizer.SEMPC.max_sqp_iter": 2,
                                                     "common.num_MPC_itrs": 2}, 2, 2),
    # sampling switched on as SURVEY.md 8(d) config 3 suggests, at the yaml's own jitter 1e-20 (params_car_residual.yaml:51):
    # the ladder is a no-op, the joint q x q Cholesky fails on the near-duplicate iterates and the draw goes through
    # root_decomposition's eigen root for the whole batch (jitter_level 4 in the fixture marks those calls)
    "car_residual_sqp": ("params_car_residual", {"agent.num_dyn_samples": 4, "agent.true_dyn_as_sample": False,
                                                 "optimizer.H": 10, "optimizer.SEMPC.max_sqp_iter": 3,
                                                 "common.num_MPC_itrs": 2}, 2, 3),
    # the same with a working ladder (1e-9): Cholesky draws, comparable sample by sample
    "car_residual_sqp_jit": ("params_car_residual", {"agent.num_dyn_samples": 4, "agent.true_dyn_as_sample": False,
                                                     "agent.Dyn_gp_jitter": 1.0e-9,
                                                     "optimizer.H": 10, "optimizer.SEMPC.max_sqp_iter": 3,
                                                     "common.num_MPC_itrs": 2}, 2, 3),
    "car_residual_fs": ("params_car_residual_fs", {"agent.num_dyn_samples": 8, "common.num_MPC_itrs": 6}, 6, 1),
    "pendulum2D_sqp": ("params_pendulum", {"agent.num_dyn_samples": 4, "optimizer.H": 8, "common.num_MPC_itrs": 2}, 2, 3),
    # params_car.yaml is stale w.r.t. src/agent.py:32: supply the missing key
    "car_sqp": ("params_car", {"env.use_model_without_derivatives": False, "agent.num_dyn_samples": 3, "optimizer.H": 6, "optimizer.SEMPC.max_sqp_iter": 2,
                               "common.num_MPC_itrs": 2}, 2, 2),
}


def load_params(name, overrides):
    with open(os.path.join(REF, "params", name + ".yaml")) as f:
        params = yaml.load(f, Loader=yaml.FullLoader)
    params["common"]["use_cuda"] = False
    params["env"]["i"], params["env"]["name"] = 1, 0
    for key, val in overrides.items():
        node = params
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = val
    return params


def synthetic_iterates(params, rng, n_calls):
    """SQP-iterate-like (x_h, u_h) sequences inside the yaml's box constraints; consecutive calls
    differ by a small perturbation so later calls land next to already-hallucinated points."""
    opt, ag = params["optimizer"], params["agent"]
    H, ns, nx, nu = opt["H"], ag["num_dyn_samples"], ag["dim"]["nx"], ag["dim"]["nu"]
    lo = np.array(opt["x_min"], dtype=float)
    hi = np.array(opt["x_max"], dtype=float)
    ulo, uhi = np.array(opt["u_min"], dtype=float), np.array(opt["u_max"], dtype=float)
    t = np.linspace(0.15, 0.85, H)[:, None]
    base_x = lo + (hi - lo) * t
    base_u = ulo + (uhi - ulo) * (0.5 + 0.3 * np.sin(3.0 * t))
    x_h = np.tile(base_x, (1, ns)) + 0.02 * np.tile(hi - lo, ns) * rng.standard_normal((H, nx * ns))
    u_h = base_u.copy()
    out = []
    for _ in range(n_calls):
        out.append((x_h.copy(), u_h.copy()))
        x_h = x_h + 0.002 * np.tile(hi - lo, ns) * rng.standard_normal(x_h.shape)
        u_h = u_h + 0.002 * (uhi - ulo) * rng.standard_normal(u_h.shape)
    return out


def run_case(name):
    yaml_name, overrides, n_mpc, n_sqp = CASES[name]
    params = load_params(yaml_name, overrides)
    torch.manual_seed(params["experiment"]["rnd_seed"]["value"])
    rng = np.random.default_rng(20260101)
    env_model = ENVS[params["env"]["dynamics"]](params)
    with contextlib.redirect_stdout(io.StringIO()):
        agent = Agent(params, env_model)
    out = {
        "yaml": np.array(yaml_name), "overrides": np.array(yaml.dump(overrides)),
        "params_yaml": np.array(yaml.dump(params)),  # fully resolved config, so tests need no /root/reference
        "X_real": agent.Dyn_gp_X_train.numpy(), "Y_real": agent.Dyn_gp_Y_train.numpy(),
        "eps": agent.epistimic_random_vector.numpy()[:n_mpc],
    }
    fs = params["env"]["use_model_without_derivatives"]
    iters = synthetic_iterates(params, rng, n_mpc * n_sqp)
    k = 0
    for mpc in range(n_mpc):
        agent.mpc_iteration(mpc)
        for sqp in range(n_sqp):
            x_h, u_h = iters[k]
            with contextlib.redirect_stdout(io.StringIO()):
                if fs:  # benchmarking/simulate_forward_sampling_car.py:118-130 (sqp index 1)
                    agent.train_hallucinated_dynGP(1, use_model_without_derivatives=True)
                    bx = agent.get_batch_x_hat(x_h, u_h)
                    gp_val, y_grad, u_grad = agent.dyn_fg_jacobians(bx, 1)
                else:  # src/solver.py:84-94
                    agent.train_hallucinated_dynGP(sqp)
                    bx = agent.get_batch_x_hat(x_h, u_h)
                    gp_val, y_grad, u_grad = agent.dyn_fg_jacobians(bx, sqp)
            out[f"x_h_{k}"], out[f"u_h_{k}"] = x_h, u_h
            out[f"gp_val_{k}"], out[f"y_grad_{k}"], out[f"u_grad_{k}"] = gp_val, y_grad, u_grad
            out[f"mean_{k}"] = agent.model_i_call.mean.numpy()
            out[f"variance_{k}"] = agent.model_i_call.variance.numpy()
            if agent.model_i_call.jitter_level is not None:
                out[f"y_sample_{k}"] = agent.model_i_samples.numpy()
                out[f"jitter_level_{k}"] = agent.model_i_call.jitter_level.numpy()
            out[f"n_halluc_{k}"] = np.array(agent.Hallcinated_X_train.shape[2])
            k += 1
    out["halluc_X_final"] = agent.Hallcinated_X_train.numpy()
    out["halluc_Y_final"] = agent.Hallcinated_Y_train.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "calls", k, "final hallucinated points", agent.Hallcinated_X_train.shape[2],
          "max jitter level", max([int(out[f"jitter_level_{i}"].max()) for i in range(k) if f"jitter_level_{i}" in out] or [-1]))


if __name__ == "__main__":
    for case in (sys.argv[1:] or CASES):
        run_case(case)
