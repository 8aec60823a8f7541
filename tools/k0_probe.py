"""K0 (real-data factorisation + inverse) time by m: python tools/k0_probe.py [m ...]
(min of three timed calls after one warm-up call; GPMPC_K0_BLOCKED_MIN_M=100000000 selects the per-pivot kernels)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200.engine import GPEngine
for m in [int(a) for a in sys.argv[1:]] or [1000, 2000, 3000, 5000, 10000]:
    g = torch.Generator().manual_seed(0)
    X = torch.rand(m, 2, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((1, m, 3), float("nan"), dtype=torch.float64); Y[0, :, 0] = torch.sin(X).sum(1)
    eng = GPEngine(4, 1, 2, 3, m)
    eng.set_hypers(np.ones((1, 2)), np.ones(1), np.full((1, 3), 1e-6), 1e-6)
    eng.set_real_data(X, Y); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); eng.set_real_data(X, Y); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    kind = "per-pivot" if int(os.environ.get("GPMPC_K0_BLOCKED_MIN_M", "768")) > m else "blocked"
    print(f"m={m:6d}  K0 (factor + inverse, {kind}) {best*1e3:9.1f} ms   status {eng.status()}", flush=True)
    del eng
