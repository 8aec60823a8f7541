"""Panel-wise packed Cholesky (block_cholesky_packed) against the pivot-wise form it replaces: BIT-equality of everything an
SQP linearisation returns, and its time.  Two builds, two processes:
    bash tools/build_variant.sh pivotwise -DGPMPC_CHOL_PIVOTWISE
    GPMPC_B200_LIB=$PWD/variants_pivotwise.so python tools/chol_ab.py gpurun_out/chol_a.npz
    python tools/chol_ab.py gpurun_out/chol_b.npz gpurun_out/chol_a.npz      # compares with the first run"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.agent import Agent

out = {}
# pendulum1D closed-loop shape through the product Agent (gpmpc_linearise: q = 51, draw + prefactored append)
params, ns, H, nx = configs.pendulum1D_sqp(), 70, 17, 2
agent = Agent(params, generate_base_samples=False)
T, g_ny = agent.in_dim_y, agent.g_ny
g = torch.Generator().manual_seed(0)
n_it = 8
agent.epistimic_random_vector = torch.randn(n_it, 1, ns, g_ny, H, T, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).cuda()
rng = np.random.default_rng(0)
x_h = np.tile(np.stack([np.linspace(2.2, 3.1, H), np.linspace(2.0, 0.1, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, nx * ns))
u_h = np.linspace(-3, 3, H).reshape(H, 1)
ms = []
for i in range(n_it):
    agent.mpc_iteration(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    agent.train_hallucinated_dynGP(0)
    f, fx, fu = agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 0)
    ms.append((time.perf_counter() - t0) * 1e3)
    out[f"p_{i}_f"], out[f"p_{i}_fx"], out[f"p_{i}_fu"] = f.copy(), fx.copy(), fu.copy()
    out[f"p_{i}_mean"] = agent.model_i_call.mean.cpu().numpy().copy()
    out[f"p_{i}_var"] = agent.model_i_call.variance.cpu().numpy().copy()
    x_h = x_h + 0.005 * rng.standard_normal(x_h.shape)
lib = os.environ.get("GPMPC_B200_LIB", "default")
print(f"{lib}: pendulum1D: host-observed ms per linearisation {np.round(ms[2:], 3).tolist()}  median {np.median(ms[2:]):.3f}  "
      f"status {agent.engine.status()}", flush=True)
# car-residual SQP shape through the engine (q = 150; the draw and the append each factorise a 150 x 150 block)
from sampling_gpmpc_b200.agent import gp_hypers_from_params
from sampling_gpmpc_b200.engine import GPEngine
from sampling_gpmpc_b200.envs import make_env_spec
cp = configs.car_residual_fs(20, 50, with_derivatives=True)
sp = make_env_spec(cp)
Xc, Yc = sp.initial_training_data(cp)
Hc, its = 50, 6
eng = GPEngine(20, 3, 2, 3, Xc.shape[0], cap_points=Hc * its)
ls, os_, nz = gp_hypers_from_params(cp, 3, 2, use_grad=True)
eng.set_hypers(ls, os_, nz, 1e-9)
eng.set_real_data(Xc, Yc)
gd = torch.Generator(device="cuda").manual_seed(0)
base = torch.stack([torch.linspace(-0.9, 0.9, Hc), torch.linspace(-0.5, 0.5, Hc)], 1).to("cuda", torch.float64)
xq = (base[None, None] + 0.05 * torch.randn(20, 1, Hc, 2, generator=gd, dtype=torch.float64, device="cuda")).expand(20, 3, Hc, 2).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
per_it = []
for it in range(its + 1):
    ee = torch.randn(20, 3, Hc, 3, generator=gd, dtype=torch.float64, device="cuda").clamp(-3, 3)
    torch.cuda.synchronize(); e0.record()
    mq, vq, yq, jq = eng.posterior(xq, ee, eng.opts(beta=3.0))
    eng.append(xq, yq)
    e1.record(); torch.cuda.synchronize()
    per_it.append(round(e0.elapsed_time(e1), 3))
    out[f"c_{it}_y"], out[f"c_{it}_m"], out[f"c_{it}_v"], out[f"c_{it}_j"] = yq.cpu().numpy(), mq.cpu().numpy(), vq.cpu().numpy(), jq.cpu().numpy()
    if it == 0:
        eng.reset_hallucinated()  # (first call: lazy module load and allocations)
    xq = (xq + 0.03 * torch.randn(20, 1, Hc, 2, generator=gd, dtype=torch.float64, device="cuda")).contiguous()
print(f"{lib}: car SQP shape (q = 150): device ms per iteration {per_it[1:]}  status {eng.status()}", flush=True)
np.savez(sys.argv[1], **out)
if len(sys.argv) > 2:
    ref = np.load(sys.argv[2])
    bad = [k for k in out if not np.array_equal(out[k], ref[k], equal_nan=True)]
    print("bit-identical to", sys.argv[2], ":", not bad, bad[:5], flush=True)
