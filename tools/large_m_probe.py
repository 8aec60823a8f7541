import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sweep
for (ns, steps, n_real, d) in ((1000, 10, 2000, 2), (1000, 10, 3000, 2), (10000, 10, 3000, 2)):
    r = sweep.run_point(ns, steps, n_real, d, 2, False, 1)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
