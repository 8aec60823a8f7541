"""Does a small-c half of the batch hide under a large-c half?  Car rollout (bench shape), the batch split into two engines A / B of
ns / 2 samples.  Measured: the full batch in one engine; A's horizon halves alone; then A's SECOND half (c = 75 .. 147, HBM bound)
on one stream beside B's FIRST half (c = 0 .. 72, issue / shared-memory bound) on another, both step kernels capped to `cap`
warps per CTA so that one CTA of each shares every SM.
    python tools/stagger_probe.py [ns] [cap...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
caps = [int(a) for a in sys.argv[2:]] or [7, 6, 8]
steps, half_t = 50, 25
u, eps = synthetic_inputs(ns, steps, 3, 0)
u, eps = u.cuda(), eps.cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    torch.cuda.synchronize(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


fr = ForwardRollout(configs.car_residual_fs(ns, steps, with_derivatives=True), condition=True)
for _ in range(2):
    t_full, ref = timed(lambda: fr.run(u, eps))
ref = ref.clone()
print(f"one engine, {ns} samples: {t_full:.1f} ms", flush=True)
del fr
torch.cuda.empty_cache()

h = ns // 2
frs = [ForwardRollout(configs.car_residual_fs(h, steps, with_derivatives=True), condition=True) for _ in range(2)]
ep = [eps[:, :h].contiguous(), eps[:, h:2 * h].contiguous()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
trajs = [torch.empty(h, ref.shape[1], steps + 1, dtype=torch.float64, device="cuda") for _ in range(2)]


def part(i, first):
    """first / second half of the horizon of engine i (the second continues from the state the first one reached)"""
    f = frs[i]
    if first:
        f.engine.reset_hallucinated()
        out = f.engine.rollout(f.env, f.x0, u[:half_t], ep[i][:half_t], f.opts)
        trajs[i][:, :, :half_t + 1] = out
    else:
        out = f.engine.rollout(f.env, trajs[i][:, :, half_t].contiguous(), u[half_t:], ep[i][half_t:], f.opts)
        trajs[i][:, :, half_t:] = out
    return out


for f in frs:
    f.engine.set_option("rollout_fused", 0)
for rep in range(2):
    a1, _ = timed(lambda: part(0, True))
    a2, _ = timed(lambda: part(0, False))
print(f"half batch alone (16 warps): steps 0-24 {a1:.1f} ms, steps 25-49 {a2:.1f} ms, sum x 2 = {2 * (a1 + a2):.1f} ms; "
      f"bit-identical to the full batch: {bool(torch.equal(trajs[0], ref[:h]))}", flush=True)

for cap in caps:
    for f in frs:
        f.engine.set_option("step_warps_cap", cap)
    part(0, True)  # A at step 25

    def overlapped():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        with torch.cuda.stream(streams[0]):
            part(0, False)
        with torch.cuda.stream(streams[1]):
            part(1, True)
        for s in streams:
            cur.wait_stream(s)

    o, _ = timed(overlapped)
    for f in frs:
        f.engine.set_option("step_warps_cap", 0)
    b2, _ = timed(lambda: part(1, False))
    ok = bool(torch.equal(trajs[0], ref[:h])) and bool(torch.equal(trajs[1], ref[h:2 * h]))
    print(f"cap {cap} warps: A[25-49] || B[0-24] {o:.1f} ms (alone: {a2:.1f} + {a1:.1f}); staggered rollout = {a1:.1f} + {o:.1f} + {b2:.1f} = "
          f"{a1 + o + b2:.1f} ms vs {t_full:.1f}; bit-identical {ok}; status {[f.engine.status() for f in frs]}", flush=True)
    # same with the capped kernels but NOT overlapped (what the cap alone costs)
    for f in frs:
        f.engine.set_option("step_warps_cap", cap)
    c1, _ = timed(lambda: part(0, True))
    c2, _ = timed(lambda: part(0, False))
    for f in frs:
        f.engine.set_option("step_warps_cap", 0)
    print(f"           capped kernels alone: steps 0-24 {c1:.1f} ms, steps 25-49 {c2:.1f} ms", flush=True)
