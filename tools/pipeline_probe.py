"""Pendulum true-reachable-set rollout (m = 180: batched GEMM k_shared_rows + k_step<WO> per step): ONE engine vs the batch
split over TWO engines on two CUDA streams, so that one half's DMMA-bound GEMM runs beside the other half's HBM-bound step.
    python tools/pipeline_probe.py [ns] [caps...]     caps = SMs the step kernel of each half may take (0 = all)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
caps = [int(a) for a in sys.argv[2:]] or [0, 74, 90, 110]
steps = 30
g = torch.Generator().manual_seed(5)
eps = torch.randn(steps, ns, 2, 1, 4, generator=g, dtype=torch.float64).clamp_(-2.5, 2.5).cuda()
u = (2.0 * torch.sin(torch.linspace(0, 3, steps, dtype=torch.float64))).reshape(steps, 1).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n=3):
    best = None
    for _ in range(n):
        torch.cuda.synchronize()
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


fr = ForwardRollout(configs.pendulum2D_rollout(ns, steps), condition=True, agent_size=20)
ms, ref = timed(lambda: fr.run(u, eps))
ref = ref.clone()
print(f"one engine: {ms:.1f} ms  {ns * steps / ms / 1e3:.2f} M sample-steps/s  status {fr.engine.status()}", flush=True)
del fr
torch.cuda.empty_cache()

half = ns // 2
frs = [ForwardRollout(configs.pendulum2D_rollout(half, steps), condition=True, agent_size=20) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
traj = torch.empty_like(ref)
eh = [eps[:, :half].contiguous(), eps[:, half:].contiguous()]


def both():
    cur = torch.cuda.current_stream()
    for i in range(2):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            frs[i].run(u, eh[i], traj[i * half:(i + 1) * half])
    for s in streams:
        cur.wait_stream(s)
    return traj


for cap in caps:
    for f in frs:
        f.engine.set_option("step_grid_cap", cap)
        f.engine.set_option("sr_grid_cap", 148 - cap if cap else 0)  # the GEMM of the other half on exactly the SMs left free
    ms, out = timed(both)
    print(f"two engines, two streams, step kernel on <= {cap or 148} SMs each: {ms:.1f} ms  {ns * steps / ms / 1e3:.2f} M sample-steps/s"
          f"  bit-identical {bool(torch.equal(out, ref))}  status {[f.engine.status() for f in frs]}", flush=True)
