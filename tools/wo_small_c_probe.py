"""Per-step time of the fused step at small c with the shared rows from the batched GEMM (force_wo) against the in-kernel
product: python tools/wo_small_c_probe.py [ns]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
steps = 50
u, eps = synthetic_inputs(ns, steps, 3, 0)
u, eps = u.cuda(), eps.cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fr = ForwardRollout(configs.car_residual_fs(ns, steps, with_derivatives=True), condition=True)
eng = fr.engine
for wo in (0, 1, 0, 1):
    eng.set_option("force_wo", wo)
    eng.reset_hallucinated()
    g = torch.Generator(device="cuda").manual_seed(0)
    out = []
    for t in range(24):
        x = (torch.rand(ns, 1, 1, 2, generator=g, dtype=torch.float64, device="cuda") - 0.5).expand(ns, 3, 1, 2).contiguous()
        torch.cuda.synchronize()
        e0.record(); eng.step(x, eps[t], fr.opts, want_moments=False); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    print("force_wo", wo, " ".join("c=%d:%.2f" % (3 * t, out[t]) for t in (1, 3, 5, 8, 12, 16, 20, 23)), "sum(0..23) %.1f ms" % sum(out), flush=True)
