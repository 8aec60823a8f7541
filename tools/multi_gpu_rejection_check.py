"""N-GPU check of the one exchange prepare_dynamics_set needs (SURVEY.md 8e "exceptions"), launched with torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_rejection_check.py

Every rank runs the rejection rollout (src/agent.py:331-443) of its own block of the dynamics samples; samples_left and the
hallucinated sets are all-gathered, rank 0's two np.random.choice draws are broadcast and every rank rewrites its rejected
samples (rollout.resample_rejected).  Rank 0 also runs the whole population on its own GPU from the same generator state:
survivors, data sets and the restored model's posterior must be BIT-IDENTICAL (samples are independent; global indices kept)."""
import copy, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from sampling_gpmpc_b200.agent import Agent, reachable_set_ball
from sampling_gpmpc_b200.envs import make_env_spec
from sampling_gpmpc_b200.rollout import gather_padded
from tests.replay import load_case

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
z, params = load_case("pendulum1D_sqp")
params = copy.deepcopy(params)
ns, H, nx = 9, 8, 2
params["agent"]["num_dyn_samples"] = ns
params["optimizer"]["H"] = H
params["agent"]["tight"] = {"use": True, "dyn_eps": 0.002, "Lipschitz": 0.96, "w_bound": 0.0001}  # params_pendulum1D_samples.yaml:47-51
params["optimizer"]["terminal_tightening"]["P"] = [[10.47241433, 0.2680862], [0.2680862, 8.74083638]]
spec = make_env_spec(params)
X, Y = torch.tensor(z["X_real"]), torch.tensor(z["Y_real"])
g = torch.Generator().manual_seed(7)
eps = torch.randn(2, 1, ns, 1, H, 3, generator=g, dtype=torch.float64).clamp(-2.5, 2.5)
rng = np.random.default_rng(3)
x_h = np.tile(np.stack([np.linspace(2.3, 3.0, H), np.linspace(1.5, 0.2, H)], 1), (1, ns)) + 0.01 * rng.standard_normal((H, nx * ns))
u_h = np.linspace(-2, 2, H).reshape(H, 1)
loosen = 40.0
_, ci = reachable_set_ball(params, np.ones(H + 1))
ci = [c * loosen for c in ci]
X_soln = np.concatenate([x_h, x_h[-1:] + 0.01], 0)
X_soln = X_soln + 0.004 * loosen * rng.standard_normal((H + 1, ns * nx)) * (np.arange(ns * nx) // nx > 4)
X_kp1 = X_soln[1, :nx].reshape(nx, 1)
base = [torch.randn(ns, 1, 1, 3, generator=g, dtype=torch.float64) for _ in range(H)]
x_h2 = x_h + 0.003 * rng.standard_normal(x_h.shape)


def run(agent):
    """one SQP iteration, the rejection rollout, the next model; returns (samples_left, X, Y, mean, var) of agent's block"""
    agent.ci_list = ci
    agent.mpc_iteration(0)
    agent.train_hallucinated_dynGP(0)
    agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 0)
    np.random.seed(11)
    left = agent.prepare_dynamics_set(X_soln, u_h, X_kp1, base_samples=base)
    agent.mpc_iteration(1)
    agent.train_hallucinated_dynGP(0)
    g2 = agent.get_g_xu_hat(agent.get_batch_x_hat(x_h2, u_h)).contiguous()
    mean, var = agent.engine.posterior(g2)
    return left, agent.Hallcinated_X_train, agent.Hallcinated_Y_train, mean, var, agent.engine.status()


sh = Agent(params, spec=spec, X_real=X, Y_real=Y, epistimic_random_vector=eps, rank=rank, world_size=world)
out = run(sh)
full = [gather_padded(t.contiguous(), ns, world) for t in out[:5]]
res = {"ok": True}
if rank == 0:
    one = Agent(params, spec=spec, X_real=X, Y_real=Y, epistimic_random_vector=eps)
    ref = run(one)
    names = ["samples_left", "X", "Y", "mean", "var"]
    for n, a, b in zip(names, full, ref[:5]):
        res[n + "_equal"] = bool(torch.equal(torch.nan_to_num(a.double(), nan=-7.0), torch.nan_to_num(b.double(), nan=-7.0)))
    left = ref[0].cpu().numpy()
    res.update(n_gpus=world, survivors=int(left.sum()), rejected=int((left == 0).sum()), status=[out[5], ref[5]])
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
