import sys, time, torch
sys.path.insert(0, "/root/repo")
from sampling_gpmpc_b200.engine import GPEngine
import numpy as np
ns = 1_000_000
g = torch.Generator(device="cuda").manual_seed(0)
traj = torch.randn(ns, 4, 51, generator=g, dtype=torch.float64, device="cuda").cumsum(2)
eng = GPEngine(4, 1, 2, 3, 5)
for _ in range(2):
    eng.traj_stats(traj); h = eng.stage_hulls(traj, 0, 1)
torch.cuda.synchronize(); t0 = time.perf_counter()
lo, hi = eng.traj_stats(traj); torch.cuda.synchronize(); t1 = time.perf_counter()
h = eng.stage_hulls(traj, 0, 1); t2 = time.perf_counter()
print("traj_stats %.2f ms, stage_hulls %.2f ms, vertices %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(h[-1])))
