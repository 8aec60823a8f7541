"""params_pendulum (true-reachable-set rollout) shape: g_ny=2, d=3, T=4, 45 real points WITH derivative observations
(m = 180), 30 steps: which shared-rows path is faster?  python tools/m180_probe.py  (run once per GPMPC_WO_MIN_M)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sweep
for ns in (20000, 200000):
    r = sweep.run_point(ns, 30, 45, 3, 2, True, 1)
    print(os.environ.get("GPMPC_WO_MIN_M", "default"), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
