#!/usr/bin/env python
"""The UNMODIFIED reference Agent on the B200: sampling-gpmpc's own src/agent.py, src/GP_model.py and
src/environments/*.py, imported as they are, with `import gpytorch` resolving to the product's shim
(sampling_gpmpc_b200.gpytorch_shim -> libgpmpc_b200.so), driven the way the reference drives them:

    SQP seam           src/solver.py:56-96         train_hallucinated_dynGP(sqp_iter); get_batch_x_hat; dyn_fg_jacobians
    car rollout loop   benchmarking/simulate_forward_sampling_car.py:117-138   (sqp index 1, value-only model)

acados is not installable, so the iterates (x_h, u_h) the QP would produce are replayed from the golden fixtures
(tests/golden/*.npz) -- those were recorded from this same reference code on the CPU stand-in (tests/golden/make_golden.py),
which makes the fixture's outputs the expected values here.  Runs in its own process (the reference calls
torch.set_default_device and the shim registers itself as `gpytorch` in sys.modules).

    python tests/ref_agent_driver.py <case> [--time REPS]     -> one JSON line

The reference checkout is looked up in $GPMPC_REFERENCE, baseline/_ref/sampling-gpmpc (a git-ignored copy that
__graft_entry__.build() makes where /root/reference is mounted, so that it travels to the GPU box), /root/reference.
Nothing under oracle/ is imported.
"""
import contextlib
import io
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)


def find_reference():
    for cand in (os.environ.get("GPMPC_REFERENCE"), os.path.join(REPO, "baseline", "_ref", "sampling-gpmpc"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "src", "agent.py")):
            return cand
    return None


def scaled_ratio(a, b, scale, rtol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (rtol * np.maximum(np.abs(b), scale)))) if a.size else 0.0


def main():
    import torch
    import yaml
    case = sys.argv[1]
    reps = int(sys.argv[sys.argv.index("--time") + 1]) if "--time" in sys.argv else 0
    ref = find_reference()
    if ref is None:
        print(json.dumps({"case": case, "unavailable": "no reference checkout (baseline/_ref/sampling-gpmpc)"}))
        return 0
    from sampling_gpmpc_b200 import gpytorch_shim as shim
    shim.install()
    try:
        import matplotlib.pyplot  # noqa: F401  (src/agent.py:13 imports it; the hot path never calls it)
    except Exception:
        m, p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        p.rcParams = {}
        m.pyplot = p
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, p
    sys.path.insert(0, ref)
    from src.agent import Agent
    from src.environments.car_model import CarKinematicsModel as bicycle
    from src.environments.car_model_residual import CarKinematicsModel as bicycle_Bdx
    from src.environments.pendulum import Pendulum as pendulum
    from src.environments.pendulum1D import Pendulum as Pendulum1D
    envs = {"pendulum": pendulum, "bicycle_Bdx": bicycle_Bdx, "bicycle": bicycle, "Pendulum1D": Pendulum1D}

    z = np.load(os.path.join(HERE, "golden", case + ".npz"))
    params = yaml.safe_load(str(z["params_yaml"]))
    # the yamls of the closed-loop configs say use_cuda True (params_pendulum1D_samples.yaml:80); --cpu-tensors keeps the
    # reference's tensors on the CPU (use_cuda False, e.g. params_car_residual.yaml): the shim then computes on the GPU and
    # hands CPU tensors back, so the reference's own post-processing keeps working
    on_gpu = "--cpu-tensors" not in sys.argv
    params["common"]["use_cuda"] = on_gpu
    dev = "cuda" if on_gpu else "cpu"
    n_calls = len([k for k in z.files if k.startswith("gp_val_")])
    fs = params["env"]["use_model_without_derivatives"]
    n_sqp = 1 if fs else params["optimizer"]["SEMPC"]["max_sqp_iter"]
    torch.manual_seed(params["experiment"]["rnd_seed"]["value"])
    with contextlib.redirect_stdout(io.StringIO()):
        agent = Agent(params, envs[params["env"]["dynamics"]](params))
    assert agent.use_cuda == on_gpu and agent.Dyn_gp_X_train_batch.is_cuda == on_gpu
    # real data as recorded (the env classes regenerate it to the last ulp or two; 1e-16 input differences are amplified
    # ~1e5 x by the solve) and the fixture's base samples: the reference draws them from the generator of the device it
    # runs on (agent.py:84-93), the fixture holds the CPU stream -- loading them is what
    # simulate_forward_sampling_car.py:78-80 does with its pickled epistemic vectors
    dX = float((agent.Dyn_gp_X_train.cpu() - torch.tensor(z["X_real"], device="cpu")).abs().max())
    agent.Dyn_gp_X_train = torch.tensor(z["X_real"], device=dev)
    agent.Dyn_gp_Y_train = torch.tensor(z["Y_real"], device=dev)
    agent.real_data_batch()
    eps_own_shape = tuple(agent.epistimic_random_vector.shape)
    agent.epistimic_random_vector = torch.tensor(z["eps"], device=dev)

    os_ = np.asarray(params["agent"]["Dyn_gp_outputscale"]["both"], dtype=np.float64).reshape(-1)
    s_val = float(np.sqrt(os_.max()))
    worst = {"mean": 0.0, "variance": 0.0, "y_sample": 0.0, "gp_val": 0.0, "y_grad": 0.0, "u_grad": 0.0}
    jitter_ok, n_h_ok = True, True
    times = []

    def one_pass(check):
        for k in range(n_calls):
            mpc, sqp = divmod(k, n_sqp)
            agent.mpc_iteration(mpc)
            x_h, u_h = z[f"x_h_{k}"], z[f"u_h_{k}"]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                if fs:
                    agent.train_hallucinated_dynGP(1, use_model_without_derivatives=True)
                    gp_val, y_grad, u_grad = agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 1)
                else:
                    agent.train_hallucinated_dynGP(sqp)
                    gp_val, y_grad, u_grad = agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), sqp)
            times.append(time.perf_counter() - t0)  # dyn_fg_jacobians ends with the device->host copies (agent.py:555-557)
            if not check:
                continue
            nonlocal jitter_ok, n_h_ok
            xscale = max(1.0, float(np.abs(x_h).max()))
            mean = agent.model_i_call.mean.cpu().numpy()
            var = agent.model_i_call.variance.cpu().numpy()
            for j in range(mean.shape[1]):
                worst["mean"] = max(worst["mean"], scaled_ratio(mean[:, j], z[f"mean_{k}"][:, j], np.sqrt(os_[j])))
                worst["variance"] = max(worst["variance"], scaled_ratio(var[:, j], z[f"variance_{k}"][:, j], os_[j]))
            if f"jitter_level_{k}" in z.files:
                jl = agent.model_i_call.jitter_level
                jitter_ok = jitter_ok and jl is not None and np.array_equal(jl.cpu().numpy(), z[f"jitter_level_{k}"])
                ys = agent.model_i_samples.cpu().numpy()
                for j in range(ys.shape[1]):
                    worst["y_sample"] = max(worst["y_sample"], scaled_ratio(ys[:, j], z[f"y_sample_{k}"][:, j], np.sqrt(os_[j])))
            worst["gp_val"] = max(worst["gp_val"], scaled_ratio(gp_val, z[f"gp_val_{k}"], s_val * xscale))
            worst["y_grad"] = max(worst["y_grad"], scaled_ratio(y_grad, z[f"y_grad_{k}"], s_val * xscale))
            worst["u_grad"] = max(worst["u_grad"], scaled_ratio(u_grad, z[f"u_grad_{k}"], s_val * xscale))
            n_h_ok = n_h_ok and agent.Hallcinated_X_train.shape[2] == int(z[f"n_halluc_{k}"])

    one_pass(check=True)
    Xh = agent.Hallcinated_X_train.cpu().numpy()
    halluc_ok = Xh.shape == z["halluc_X_final"].shape and bool(np.array_equal(Xh, z["halluc_X_final"]))
    from sampling_gpmpc_b200 import engine as _engine
    out = {"case": case, "reference": ref, "calls": n_calls, "worst_over_tolerance": worst, "jitter_levels_equal": jitter_ok,
           "hallucinated_counts_equal": n_h_ok, "hallucinated_inputs_bit_equal": halluc_ok, "real_X_regenerated_max_abs_diff": dX,
           "eps_shape_generated_by_reference": list(eps_own_shape), "model_class": type(agent.model_i).__mro__[1].__module__,
           "native_library": _engine.LIB_PATH, "device": torch.cuda.get_device_name(0), "reference_tensors_on": dev}
    if reps:
        first = list(times)
        times.clear()
        for _ in range(reps):  # replays restart the hallucinated set at sqp 0 of every MPC step, as the reference does
            one_pass(check=False)
        per_call = np.asarray(times).reshape(reps, n_calls)
        out["ms_per_linearisation"] = {"median": float(np.median(per_call) * 1e3), "min": float(per_call.min() * 1e3),
                                       "first_pass_ms": [round(t * 1e3, 3) for t in first],
                                       "what": "host-observed train_hallucinated_dynGP + get_batch_x_hat + dyn_fg_jacobians of "
                                               "the unmodified reference Agent on the shim, incl. its three D2H copies"}
    print(json.dumps(out))
    ok = all(v <= 1.0 for v in worst.values()) and jitter_ok and n_h_ok and halluc_ok
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
