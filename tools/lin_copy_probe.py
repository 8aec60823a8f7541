import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import Agent
params = configs.pendulum1D_sqp()
ns, H, T = 70, 17, 3
agent = Agent(params, generate_base_samples=False)
g = torch.Generator().manual_seed(0)
n_it = 40
agent.epistimic_random_vector = torch.randn(n_it, 1, ns, 1, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).cuda()
rng = np.random.default_rng(0)
x_h = np.tile(np.stack([np.linspace(2.2, 3.1, H), np.linspace(2.0, 0.1, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, 2 * ns))
u_h = np.linspace(-3, 3, H).reshape(H, 1)
eng = agent.engine
opts = eng.opts(params["agent"]["Dyn_gp_beta"], params["agent"]["Dyn_gp_variance_is_zero"])
for mode in ("host_in_host_out", "dev_in_no_copy", "host_in_host_out", "dev_in_no_copy"):
    ms = []
    for i in range(n_it):
        agent.mpc_iteration(i)
        xu = agent.get_batch_x_hat(x_h, u_h)
        xu_d = xu.cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        agent.train_hallucinated_dynGP(0)
        bufs = agent.__dict__.setdefault("_lb_" + mode, {})
        if mode == "host_in_host_out":
            eng.linearise(agent.env_struct, xu, agent.epistimic_random_vector[i][0], opts, agent._pending_reset, bufs)
        else:
            eng.linearise(agent.env_struct, xu_d, agent.epistimic_random_vector[i][0], opts, agent._pending_reset, bufs, copy_to_host=False)
        agent._pending_reset = False
        agent._data_version += 1
        agent._appended_since_train = True
        eng.raise_on_status()
        ms.append((time.perf_counter() - t0) * 1e3)
    print(mode, "median ms %.4f  min %.4f" % (np.median(ms[5:]), np.min(ms[5:])), flush=True)
