import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from sampling_gpmpc_b200.engine import GPEngine
ns, steps = 10000, 6
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for t in range(steps):
    eps = torch.randn(ns, g_ny, 1, T, generator=gd, dtype=torch.float64, device="cuda").clamp_(-3, 3)
    xx = x.expand(ns, g_ny, 1, d).contiguous()
    torch.cuda.synchronize(); e0.record()
    eng.step(xx, eps, eng.opts(beta=3.0), want_moments=False)
    e1.record(); torch.cuda.synchronize(); ts.append(round(e0.elapsed_time(e1), 1))
print(os.environ.get("GPMPC_B200_LIB", "default"), "m=1e4 ns=1e4 ms per step", ts, flush=True)
