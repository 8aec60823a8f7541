import os, sys
sys.path.insert(0, "/root/repo")
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
ns, steps = 100000, 30
g = torch.Generator().manual_seed(5)
eps = torch.randn(steps, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp_(-2.5, 2.5).cuda()
u = (2.0 * torch.sin(torch.linspace(0, 3, steps, dtype=torch.float64))).reshape(steps, 1).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fr = ForwardRollout(configs.pendulum2D_rollout(ns, steps), condition=True, agent_size=20)
for cap in (0, 110, 74, 50):
    fr.engine.set_option("step_grid_cap", cap)
    for i in range(2):
        torch.cuda.synchronize(); e0.record(); fr.run(u, eps); e1.record(); torch.cuda.synchronize()
    print("cap", cap, "ms", e0.elapsed_time(e1), flush=True)
