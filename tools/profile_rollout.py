"""One conditioned car rollout (for ncu): python tools/profile_rollout.py [ns] [steps] [T3=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t3 = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) == 1
fr = ForwardRollout(configs.car_residual_fs(ns, steps, with_derivatives=t3), condition=t3)
u, eps = synthetic_inputs(ns, steps, 3 if t3 else 1, 0)
traj = fr.run(u.cuda(), eps.cuda())
torch.cuda.synchronize()
print("done", float(traj.abs().max()), fr.engine.status())
