import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200.engine import GPEngine
for n in (180, 1000, 2000, 3000):
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n, 2, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((1, n, 3), float("nan"), dtype=torch.float64); Y[0, :, 0] = torch.sin(X).sum(1)
    e = GPEngine(4, 1, 2, 3, n)
    e.set_hypers(np.ones((1, 2)), np.ones(1), np.full((1, 3), 1e-6), 1e-6)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e.set_real_data(X, Y); torch.cuda.synchronize()
    print("m=%d  K0 (factor + inverse + beta): %.1f ms  status %#x" % (n, (time.perf_counter() - t0) * 1e3, e.status()), flush=True)
    del e
