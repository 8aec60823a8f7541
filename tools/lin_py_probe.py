"""Where the host time of one SQP linearisation goes (pendulum1D shape): python tools/lin_py_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import Agent
params = configs.pendulum1D_sqp()
ns, H, T = 70, 17, 3
agent = Agent(params, generate_base_samples=False)
g = torch.Generator().manual_seed(0)
n_it = 60
agent.epistimic_random_vector = torch.randn(n_it, 1, ns, 1, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).cuda()
rng = np.random.default_rng(0)
x_h = np.tile(np.stack([np.linspace(2.2, 3.1, H), np.linspace(2.0, 0.1, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, 2 * ns))
u_h = np.linspace(-3, 3, H).reshape(H, 1)
T_ = {k: [] for k in ("mpc_iteration", "get_batch_x_hat", "train", "dyn_fg_jacobians", "total")}
pc = time.perf_counter
for i in range(n_it):
    torch.cuda.synchronize()
    t0 = pc(); agent.mpc_iteration(i); t1 = pc()
    xu = agent.get_batch_x_hat(x_h, u_h); t2 = pc()
    agent.train_hallucinated_dynGP(0); t3 = pc()
    agent.dyn_fg_jacobians(xu, 0); t4 = pc()
    for k, v in zip(T_, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t1)):
        T_[k].append(v * 1e6)
print({k: round(float(np.median(v[10:])), 1) for k, v in T_.items()}, "us (median)")
# the pieces of dyn_fg_jacobians
eng = agent.engine
P = {k: [] for k in ("applies", "opts", "linearise_call", "raise_on_status", "numpy_copy_slices")}
for i in range(n_it):
    agent.mpc_iteration(i)
    xu = agent.get_batch_x_hat(x_h, u_h)
    agent.train_hallucinated_dynGP(0)
    torch.cuda.synchronize()
    t0 = pc(); ok = agent._fused_linearisation_applies(xu); t1 = pc()
    ag = params["agent"]; opts = eng.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"]); t2 = pc()
    if not hasattr(agent, "_lin_bufs"):
        agent._lin_bufs = {}
    mean, var, y_gp, jl, _, host = eng.linearise(agent.env_struct, xu, agent.epistimic_random_vector[i][0], opts, agent._pending_reset, agent._lin_bufs); t3 = pc()
    agent._pending_reset = False; agent._data_version += 1; agent._appended_since_train = True
    eng.raise_on_status(); t4 = pc()
    h = host.numpy().copy(); a, b, c = h[:, :, :, [0]], h[:, :, :, 1:3], h[:, :, :, 3:4]; t5 = pc()
    for k, v in zip(P, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        P[k].append(v * 1e6)
print({k: round(float(np.median(v[10:])), 1) for k, v in P.items()}, "us (median)")
