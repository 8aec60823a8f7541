"""N-GPU check of SURVEY.md 8(e) on real hardware (NCCL over NVLink), launched with torchrun:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py [ns]

Every rank rolls out its contiguous shard of the dynamics samples (no data-path collective), ONE all-gather brings the
trajectories to every rank, and rank 0 also runs the whole population on its own GPU: the gathered array must be
BIT-IDENTICAL to the single-GPU result (samples are independent; global indices are kept), and so must the consumer
reductions on it (stage boxes / max-deviation tightening, per-stage hull vertices).  Also times the all-gather."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.rollout import ForwardRollout
from bench import synthetic_inputs

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20001  # not divisible by the world size on purpose
steps = 50
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
params = configs.car_residual_fs(ns, steps, with_derivatives=True)
u, eps = synthetic_inputs(ns, steps, 3, 0)
u, eps = u.cuda(), eps.cuda()
fr = ForwardRollout(params, condition=True, rank=rank, world_size=world)
traj = fr.run(u, eps)
full = fr.all_gather_trajectories(traj)          # warm-up of the communicator
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    full = fr.all_gather_trajectories(traj)
e1.record(); torch.cuda.synchronize()
gather_ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
dist.all_reduce(gather_ms, op=dist.ReduceOp.MAX)
stats = fr.engine.traj_stats(full)
hulls = fr.engine.stage_hulls(full)
# ... and the same consumers WITHOUT the gather: local reductions, only the reductions travel (3 x 1.6 KB + the hull vertices)
ref_traj = full[0].contiguous()  # a reference trajectory for the max-deviation tightening (identical on every rank)
stats_g = fr.engine.traj_stats(full, ref_traj)
stats_d = fr.stage_boxes(traj, ref_traj)
hulls_d = fr.stage_hulls(traj)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(5):
    fr.stage_boxes(traj, ref_traj); fr.stage_hulls(traj)
e1.record(); torch.cuda.synchronize()
red_ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
dist.all_reduce(red_ms, op=dist.ReduceOp.MAX)
e0.record()
for _ in range(5):
    f2 = fr.all_gather_trajectories(traj); fr.engine.traj_stats(f2, ref_traj); fr.engine.stage_hulls(f2)
e1.record(); torch.cuda.synchronize()
gat_ms = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
dist.all_reduce(gat_ms, op=dist.ReduceOp.MAX)
res = {"ok": True}
res["reduced_boxes_equal_gathered"] = all(bool(torch.equal(a, b)) for a, b in zip(stats_d, stats_g))
res["reduced_hulls_equal_gathered"] = len(hulls_d) == len(hulls) and all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(hulls_d, hulls))
res["consumers_ms"] = {"all_gather_then_reduce": float(gat_ms), "reduce_then_exchange": float(red_ms)}
if rank == 0:
    one = ForwardRollout(params, condition=True)
    ref = one.run(u, eps)
    res["gathered_equals_single_gpu"] = bool(torch.equal(full, ref))
    s1 = one.engine.traj_stats(ref)
    res["traj_stats_equal"] = all(torch.equal(a, b) if torch.is_tensor(a) else np.array_equal(a, b) for a, b in zip(stats, s1))
    h1 = one.engine.stage_hulls(ref)
    res["hulls_equal"] = len(h1) == len(hulls) and all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(hulls, h1))
    res.update(n_gpus=world, ns=ns, steps=steps, shard=[fr.s_lo, fr.s_hi],
               all_gather_ms=float(gather_ms), gathered_MB=full.numel() * 8 / 1e6,
               status=[fr.engine.status(), one.engine.status()])
    res["ok"] = res["gathered_equals_single_gpu"] and res["traj_stats_equal"] and res["hulls_equal"]
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
