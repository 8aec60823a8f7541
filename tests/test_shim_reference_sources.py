"""CPU, only where the reference checkout is mounted (/root/reference; skipped on the GPU box): the UNMODIFIED
`src/agent.py` + `src/GP_model.py` + env classes import and build their model against the product's gpytorch shim
(SURVEY.md 8b level B1) -- every constructor, keyword and property setter they use exists -- and the first model call
fails loudly for want of a CUDA device (no CPU fallback).  The arithmetic behind the shim is checked on the GPU by
tests/test_gpu_shim.py against the fixtures this same reference code produced."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPMPC_REFERENCE", "/root/reference")

SCRIPT = r'''
import contextlib, io, sys, types
import torch, yaml
REPO, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, REPO)
from sampling_gpmpc_b200 import gpytorch_shim as shim
shim.install()
try:
    import matplotlib.pyplot  # agent.py:13 imports it; the hot path never calls it
except Exception:
    m, p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    p.rcParams = {}
    m.pyplot = p
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, p
sys.path.insert(0, REF)
from src.agent import Agent
from src.environments.pendulum1D import Pendulum
with open(REF + "/params/params_pendulum1D_samples.yaml") as f:
    params = yaml.load(f, Loader=yaml.FullLoader)
params["common"]["use_cuda"] = False
params["env"]["i"], params["env"]["name"] = 1, 0
params["agent"]["num_dyn_samples"], params["common"]["num_MPC_itrs"] = 5, 2
with contextlib.redirect_stdout(io.StringIO()):
    agent = Agent(params, Pendulum(params))
    agent.mpc_iteration(0)
    agent.train_hallucinated_dynGP(0)
m = agent.model_i
assert isinstance(m, shim.ExactGP), type(m).__mro__
ns, n, d, T = 5, agent.Dyn_gp_X_train.shape[0], 2, 3
assert tuple(m.train_inputs[0].shape) == (ns, 1, n, d) and tuple(m.train_targets.shape) == (ns, 1, n, T)
assert tuple(m.covar_module.base_kernel.lengthscale.shape) == (ns, 1, 1, d)
assert tuple(m.covar_module.outputscale.shape) == (ns, 1)
assert tuple(m.likelihood.noise.shape) == (ns, 1, 1) and tuple(m.likelihood.task_noises.shape) == (ns, 1, T)
ls, os_, noise, jit = m._hypers(ns, 1, d, T)
import numpy as np
assert ls.shape == (1, d) and abs(ls[0, 0] - np.asarray(params["agent"]["Dyn_gp_lengthscale"]["both"]).reshape(-1)[0]) < 1e-15
assert abs(noise[0, 1] - (params["agent"]["Dyn_gp_noise"] + params["agent"]["Dyn_gp_task_noises"]["val"][1]
                          * params["agent"]["Dyn_gp_task_noises"]["multiplier"])) < 1e-18
H = params["optimizer"]["H"]
x_h = torch.zeros(H, 2 * ns).numpy() + 2.5
u_h = torch.zeros(H, 1).numpy()
try:
    with contextlib.redirect_stdout(io.StringIO()):
        agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 0)
except RuntimeError as e:
    assert "CUDA" in str(e) and "no CPU fallback" in str(e), e
    print("OK loud failure without a GPU")
else:
    assert torch.cuda.is_available(), "a model call without a GPU must raise"
    print("OK evaluated on the GPU")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not mounted")
def test_unmodified_reference_agent_builds_its_model_on_the_shim():
    r = subprocess.run([sys.executable, "-c", SCRIPT, REPO, REF], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "OK" in r.stdout
