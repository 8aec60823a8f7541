"""One conditioned pendulum true-reachable-set rollout (m = 180, 2e5 samples x 30 steps), for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
ns, st = 200000, 30
fr = ForwardRollout(configs.pendulum2D_rollout(ns, st), condition=True, agent_size=20)
g = torch.Generator().manual_seed(5)
eps = torch.randn(st, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp_(-2.5, 2.5).cuda()
u = (2.0 * torch.sin(torch.linspace(0, 3, st, dtype=torch.float64))).reshape(st, 1).cuda()
fr.run(u, eps); torch.cuda.synchronize(); print("ok", fr.engine.status())
