"""Per-step device time of the fused step kernel along one conditioned car rollout (kernel experiments):
python tools/step_sweep.py [ns]   -> ms of steps 0,10,20,30,40,49 and the rollout total."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
steps = 50
fr = ForwardRollout(configs.car_residual_fs(ns, steps, with_derivatives=True), condition=True)
u, eps = synthetic_inputs(ns, steps, 3, 0)
u, eps = u.cuda(), eps.cuda()
for _ in range(2):
    fr.run(u, eps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fr.run(u, eps); e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1)
# per-step: drive the engine step by step
eng = fr.engine
eng.reset_hallucinated()
x = torch.zeros(ns, 3, 1, 2, dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
out = []
for t in range(steps):
    x = (torch.rand(ns, 1, 1, 2, generator=g, dtype=torch.float64, device="cuda") - 0.5).expand(ns, 3, 1, 2).contiguous()
    torch.cuda.synchronize()
    e0.record(); eng.step(x, eps[t], fr.opts, want_moments=False); e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1))
b = lambda c: 8 * (c * 45 + c * (c + 1) / 2) * ns * 3 / 1e9
print(os.environ.get("GPMPC_B200_LIB", "default"), "rollout_ms %.1f" % total,
      " ".join("c=%d:%.2fms(%.0fGB/s)" % (3 * t, out[t], b(3 * t) / out[t] * 1e3) for t in (1, 5, 10, 20, 30, 40, 49)), "status", eng.status())
