"""Fused-horizon rollout (k_horizon) vs the step-wise rollout at the bench shape: python tools/hz_probe.py [ns] [configs...]
each config = groups:stagger_ns (stagger -1 = automatic, 0 = none)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
cfgs = sys.argv[2:] or ["step", "4:0", "4:-1", "3:-1", "2:-1"]
steps = 50
fr = ForwardRollout(configs.car_residual_fs(ns, steps, with_derivatives=True), condition=True)
u, eps = synthetic_inputs(ns, steps, 3, 0)
u, eps = u.cuda(), eps.cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for cfg in cfgs:
    if cfg == "auto":
        fr.use_fused_horizon("auto")
        fr.engine.set_option("hz_groups", 0)
        fr.engine.set_option("hz_stagger_ns", -1)
    elif cfg in ("step", "wo"):
        fr.use_fused_horizon(False)
        fr.engine.set_option("force_wo", int(cfg == "wo"))
    else:
        g, s = cfg.split(":")
        fr.use_fused_horizon(True)
        fr.engine.set_option("hz_groups", int(g))
        fr.engine.set_option("hz_stagger_ns", int(s))
    times = []
    for i in range(4):
        torch.cuda.synchronize()
        e0.record(); traj = fr.run(u, eps); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    if ref is None:
        ref = traj.clone()
    same = bool(torch.equal(traj, ref))
    print(f"ns={ns} {cfg:>10s}  ms {' '.join('%.1f' % t for t in times)}  best {min(times):.1f}  M sample-steps/s {ns*steps/min(times)/1e3:.2f}"
          f"  bit-identical {same}  status {fr.engine.status()}", flush=True)
