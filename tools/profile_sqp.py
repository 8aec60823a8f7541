"""SQP linearisation calls at the pendulum1D / car-residual shapes (for ncu launch lists): python tools/profile_sqp.py [car]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import Agent
params = configs.pendulum1D_sqp()
ns, H, T = 70, 17, 3
agent = Agent(params, generate_base_samples=False)
g = torch.Generator().manual_seed(0)
agent.epistimic_random_vector = torch.randn(6, 1, ns, 1, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).cuda()
rng = np.random.default_rng(0)
x_h = np.tile(np.stack([np.linspace(2.2, 3.1, H), np.linspace(2.0, 0.1, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, 2 * ns))
u_h = np.linspace(-3, 3, H).reshape(H, 1)
for i in range(6):
    agent.mpc_iteration(i)
    agent.train_hallucinated_dynGP(0)
    agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 0)
    x_h = x_h + 0.005 * rng.standard_normal(x_h.shape)
torch.cuda.synchronize()
print("done", agent.engine.status())
