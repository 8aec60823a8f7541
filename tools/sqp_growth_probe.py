"""SQP-mode cost as the hallucinated set grows (BASELINE configs[2]): car-residual shape g_ny=3, d=2,
T=3, m=45, H=50 (q=150 joint scalars per call, +150 factor rows per SQP iteration).  python tools/sqp_growth_probe.py [ns] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import gp_hypers_from_params
from sampling_gpmpc_b200.engine import GPEngine
from sampling_gpmpc_b200.envs import make_env_spec

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, T, d, g_ny = 50, 3, 2, 3
params = configs.car_residual_fs(ns, 50, with_derivatives=True)
params["agent"]["Dyn_gp_jitter"] = 1e-9
spec = make_env_spec(params)
X, Y = spec.initial_training_data(params)
eng = GPEngine(ns, g_ny, d, T, X.shape[0], cap_points=H * iters)
ls, os_, noise = gp_hypers_from_params(params, g_ny, d, use_grad=True)
eng.set_hypers(ls, os_, noise, 1e-9)
eng.set_real_data(X, Y)
g = torch.Generator(device="cuda").manual_seed(0)
base = torch.stack([torch.linspace(-0.9, 0.9, H), torch.linspace(-0.5, 0.5, H)], 1).to("cuda", torch.float64)
x = (base[None, None] + 0.05 * torch.randn(ns, 1, H, d, generator=g, dtype=torch.float64, device="cuda")).expand(ns, g_ny, H, d).contiguous()
opts = eng.opts(beta=3.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(iters):
    eps = torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64, device="cuda").clamp(-3, 3)
    torch.cuda.synchronize()
    e0.record()
    mean, var, y, jl = eng.posterior(x, eps, opts)
    eng.append(x, y)
    e1.record()
    torch.cuda.synchronize()
    print("iter %2d  factor rows %5d  posterior+append %8.3f ms  max jitter level %d  status %#x"
          % (it, eng.num_factor_rows, e0.elapsed_time(e1), int(jl.max()), eng.status()), flush=True)
    x = (x + 0.03 * torch.randn(ns, 1, H, d, generator=g, dtype=torch.float64, device="cuda")).contiguous()
