"""SQP iterations at the car-residual shape (ns = 20, g_ny = 3, H = 50, T = 3: q = 150, +150 factor rows per iteration) for ncu
launch lists: python tools/profile_sqp_car.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import gp_hypers_from_params
from sampling_gpmpc_b200.engine import GPEngine
from sampling_gpmpc_b200.envs import make_env_spec
its = int(sys.argv[1]) if len(sys.argv) > 1 else 7
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cp = configs.car_residual_fs(NS, 50, with_derivatives=True)
sp = make_env_spec(cp)
Xc, Yc = sp.initial_training_data(cp)
Hc = 50
eng = GPEngine(NS, 3, 2, 3, Xc.shape[0], cap_points=Hc * its)
ls, os_, nz = gp_hypers_from_params(cp, 3, 2, use_grad=True)
eng.set_hypers(ls, os_, nz, 1e-9)
eng.set_real_data(Xc, Yc)
gd = torch.Generator(device="cuda").manual_seed(0)
base = torch.stack([torch.linspace(-0.9, 0.9, Hc), torch.linspace(-0.5, 0.5, Hc)], 1).to("cuda", torch.float64)
xq = (base[None, None] + 0.05 * torch.randn(NS, 1, Hc, 2, generator=gd, dtype=torch.float64, device="cuda")).expand(NS, 3, Hc, 2).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms, chk = [], 0.0
for it in range(its):
    ee = torch.randn(NS, 3, Hc, 3, generator=gd, dtype=torch.float64, device="cuda").clamp(-3, 3)
    torch.cuda.synchronize(); e0.record()
    eng.set_option("prefactor_next", 1)
    _, _, yq, _ = eng.posterior(xq, ee, eng.opts(beta=3.0))
    eng.append(xq, yq)
    e1.record(); torch.cuda.synchronize(); ms.append(round(e0.elapsed_time(e1), 3))
    chk += float(yq.double().sum())
    xq = (xq + 0.03 * torch.randn(NS, 1, Hc, 2, generator=gd, dtype=torch.float64, device="cuda")).contiguous()
torch.cuda.synchronize()
print(os.environ.get("GPMPC_B200_LIB", "default"), "ms per iteration", ms, "checksum %.17g" % chk, "status", eng.status(), eng.num_factor_rows)
