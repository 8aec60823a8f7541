"""A few fused steps at m = 1e4 (for ncu): python tools/profile_large_m.py [ns] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200.engine import GPEngine
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, d, T, g_ny = 10000, 2, 3, 2
g = torch.Generator().manual_seed(0)
X = torch.rand(m, d, generator=g, dtype=torch.float64) * 2 - 1
Y = torch.full((g_ny, m, T), float("nan"), dtype=torch.float64)
Y[:, :, 0] = torch.sin(X).sum(1)
eng = GPEngine(ns, g_ny, d, T, m, cap_points=steps)
eng.set_hypers(np.ones((g_ny, d)), np.ones(g_ny), np.full((g_ny, T), 1e-6), 1e-6)
eng.set_real_data(X, Y)
gd = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(ns, 1, 1, d, generator=gd, dtype=torch.float64, device="cuda") * 1.6 - 0.8
for t in range(steps):
    eps = torch.randn(ns, g_ny, 1, T, generator=gd, dtype=torch.float64, device="cuda").clamp_(-3, 3)
    eng.step(x.expand(ns, g_ny, 1, d), eps, eng.opts(beta=3.0), want_moments=False)
    x = (x + 0.05 * torch.randn(ns, 1, 1, d, generator=gd, dtype=torch.float64, device="cuda")).clamp_(-1, 1)
torch.cuda.synchronize()
print("done", eng.status(), eng.launch_count)
