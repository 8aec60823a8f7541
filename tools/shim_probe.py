"""Cost of the zero-edit boundary (gpytorch shim, SURVEY.md 8b level B1) at the pendulum1D SQP shape: per SQP iteration the
reference builds a new model on [real || previous step's 17 hallucinated points], calls it on 17 test points and samples.
python tools/shim_probe.py  -> host-observed ms per (model build + call + sample), median of the steady iterations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sampling_gpmpc_b200 import configs, gpytorch_shim as shim
from sampling_gpmpc_b200.envs import make_env_spec

G = shim.namespace()
params = configs.pendulum1D_sqp()
ag = params["agent"]
ns, g_ny, d, T, H = 70, 1, 2, 3, 17
bs = torch.Size([ns, g_ny])
X, Y = make_env_spec(params).initial_training_data(params)
Xr, Yr = torch.tile(X, (ns, g_ny, 1, 1)).cuda(), torch.tile(Y, (ns, 1, 1, 1)).cuda()


class Model(G.models.ExactGP):
    def __init__(self, tx, ty, lik):
        super().__init__(tx, ty, lik)
        self.mean_module = G.means.ConstantMeanGrad(batch_shape=bs)
        self.base_kernel = G.kernels.RBFKernelGrad(ard_num_dims=d, batch_shape=bs)
        self.covar_module = G.kernels.ScaleKernel(self.base_kernel, batch_shape=bs)


def build(tx, ty):
    lik = G.likelihoods.MultitaskGaussianLikelihood(num_tasks=T, rank=0, noise_constraint=G.constraints.GreaterThan(0.0), batch_shape=bs)
    m = Model(tx, ty, lik)
    m.likelihood.noise = torch.tile(torch.tensor([ag["Dyn_gp_noise"]], dtype=torch.float64), dims=(ns, g_ny, 1))
    m.likelihood.task_noises = torch.tile(torch.tensor(ag["Dyn_gp_task_noises"]["val"], dtype=torch.float64) * ag["Dyn_gp_task_noises"]["multiplier"], dims=(ns, g_ny, 1))
    m.covar_module.base_kernel.lengthscale = torch.tile(torch.tensor(ag["Dyn_gp_lengthscale"]["both"], dtype=torch.float64), dims=(ns, 1, 1, 1))
    m.covar_module.outputscale = torch.tile(torch.tensor(ag["Dyn_gp_outputscale"]["both"], dtype=torch.float64), dims=(ns, 1))
    return m.eval().cuda()


g = torch.Generator(device="cuda").manual_seed(0)
base = torch.stack([torch.linspace(2.2, 3.1, H), torch.linspace(2.0, 0.1, H)], 1).to("cuda", torch.float64)
hx = hy = None
times = []
for it in range(10):
    x = (base[None, None] + 0.01 * torch.randn(ns, g_ny, H, d, generator=g, dtype=torch.float64, device="cuda")).contiguous()
    eps = torch.randn(ns, g_ny, H, T, generator=g, dtype=torch.float64, device="cuda").clamp(-2.5, 2.5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tx = Xr if hx is None else torch.cat([Xr, hx], 2)
    ty = Yr if hy is None else torch.cat([Yr, hy], 2)
    with G.settings.cholesky_jitter(double_value=ag["Dyn_gp_jitter"]):
        post = build(tx, ty)(x)
        y = post.sample(base_samples=eps)
    mean, var = post.mean, post.variance
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
    hx, hy = x, y  # max_sqp_iter = 1: the next model sees exactly this step's points (agent.py:261-272)
print("shim: model build + call + sample, ms per SQP iteration:", [round(t, 2) for t in times], "median of the last 6: %.2f" % float(np.median(times[4:])))
