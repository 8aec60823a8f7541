#!/usr/bin/env python
"""bench.py -- conditioned sample-steps/s of the GP-sampling hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[3], the metric's own configuration): the car forward rollout of
benchmarking/simulate_forward_sampling_car.py with true iterative conditioning -- every sampled point
(value + 2 derivative tasks) becomes training data of its own dynamics sample -- on the shapes of
params_car_residual_fs.yaml: g_ny=3 outputs, d=2, T=3, m=45 shared real observations, 50 horizon steps,
125 000 dynamics samples per GPU (10^6 over 8 GPUs, weak scaling: samples are independent, no collective on
the data path).  One bench "step" = one whole 50-step rollout of this rank's samples.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference-equivalent CPU torch path (oracle, full re-fit)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "conditioned_sample_steps_per_sec"
UNIT = "sample-steps/s"
HORIZON = 50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ns-per-gpu", type=int, default=125000)
    ap.add_argument("--cpu-samples", type=int, default=192, help="bounded sample for the CPU legs (~10 s of host work)")
    ap.add_argument("--no-extras", action="store_true", help="skip the as-shipped / SQP side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(ns_per_gpu, n_gpus):
    return {
        "workload": "car forward rollout with iterative conditioning (params_car_residual_fs shapes, "
                    "use_model_without_derivatives=False): g_ny=3, d=2, T=3, m=45 shared real observations, "
                    f"{HORIZON} horizon steps, factor grows to c=150 rows per (sample, output)",
        "ns_per_gpu": ns_per_gpu, "ns_total": ns_per_gpu * n_gpus, "horizon": HORIZON,
        "parallelism": f"samples sharded contiguously over {n_gpus} GPU(s), no data-path collective",
        "l2": "per-GPU factor state (~58 GB at 125k samples) >> 126 MB L2: inputs larger than L2, no flush needed",
    }


def synthetic_inputs(ns, horizon, T, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    pin = torch.cuda.is_available()  # the CPU legs must run without a GPU too
    eps = torch.empty(horizon, ns, 3, 1, T, dtype=torch.float64, pin_memory=pin)
    torch.randn(eps.shape, generator=g, dtype=torch.float64, out=eps)
    eps.clamp_(-3.0, 3.0)
    t = torch.linspace(0, 1, horizon, dtype=torch.float64)
    u = torch.stack([0.05 * torch.sin(6.0 * t), 0.3 * torch.cos(4.0 * t)], 1).contiguous()
    u = u.pin_memory() if pin else u
    return u, eps


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rollout(ns, horizon, seed, threads, return_traj=False):
    """The reference-equivalent CPU torch path (oracle: full GP re-fit per step) on a bounded sample."""
    import torch
    from oracle.rollout_ref import reference_rollout
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.envs import make_env_spec
    torch.set_num_threads(threads)
    params = configs.car_residual_fs(ns, horizon, with_derivatives=True)
    u, eps = synthetic_inputs(ns, horizon, 3, seed)
    t0 = time.perf_counter()
    traj = reference_rollout(params, make_env_spec(params), u, eps, condition=True)
    dt = time.perf_counter() - t0
    return (dt, traj) if return_traj else dt


def gpu_parity_against_cpu(ns, horizon, seed, ref_traj, device):
    """The trajectories the CPU leg just computed (oracle, full re-fit per step) against the CUDA path on the same inputs:
    worst |gpu - cpu| / (1e-9 * max(|cpu|, s)), s = the state scale 14 (m/s) * sqrt(max outputscale).  <= 1 passes."""
    import numpy as np
    import torch
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.rollout import ForwardRollout
    params = configs.car_residual_fs(ns, horizon, with_derivatives=True)
    u, eps = synthetic_inputs(ns, horizon, 3, seed)
    fr = ForwardRollout(params, condition=True, device=device)
    traj = fr.run(u.to(device), eps.to(device)).cpu().numpy()
    status = fr.check()
    scale = 14.0 * float(np.sqrt(max(params["agent"]["Dyn_gp_outputscale"]["both"])))
    worst = float(np.max(np.abs(traj - ref_traj) / (1e-9 * np.maximum(np.abs(ref_traj), scale))))
    return {"samples": ns, "steps": horizon, "worst_error_over_tolerance": worst, "passes": bool(worst <= 1.0),
            "tolerance": "1e-9 * max(|cpu|, 14 * sqrt(outputscale))", "engine_status": int(status),
            "what": "CUDA rollout vs the oracle's full-refit CPU rollout of the cpu_baseline leg, same seeded inputs"}


def run_reference(args):
    """--impl reference: times the reference's CPU implementation of the path (gpytorch is not installable,
    so this is the oracle's op-sequence-faithful restatement incl. the full re-fit per conditioning step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ns = args.cpu_samples
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rollout(min(ns, 8), 10, 1, threads)
    times = [cpu_reference_rollout(ns, HORIZON, 100 + i, threads) for i in range(args.steps)]
    dt = sum(times) / len(times)
    value = ns * HORIZON / dt
    sample = f"{ns} samples x {HORIZON} steps per bench step (same shapes, same per-sample work as the GPU arm)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.ns_per_gpu, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def extras(torch, device, cpu_leg=True):
    """Side measurements reported next to the headline (not the bench value): the script as shipped
    (value-only model, no conditioning) and ms per SQP GP linearisation at the pendulum1D shape."""
    import numpy as np
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.agent import Agent
    from sampling_gpmpc_b200.rollout import ForwardRollout
    out = {}
    ns = 1_000_000
    fr = ForwardRollout(configs.car_residual_fs(ns, HORIZON, with_derivatives=False), condition=False, device=device)
    u, eps = synthetic_inputs(ns, HORIZON, 1, 3)
    u, eps = u.to(device), eps.to(device)
    for _ in range(2):
        fr.run(u, eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fr.run(u, eps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["as_shipped_value_only"] = {"sample_steps_per_sec": ns * HORIZON / (ms * 1e-3), "ms_per_rollout": ms, "ns": ns,
                                    "note": "simulate_forward_sampling_car.py as shipped: T=1, model on real data only"}
    del fr, eps
    torch.cuda.empty_cache()

    # the reference's OWN sample counts (params_car_residual_fs.yaml: 200 dynamics samples; the SQP configs 20): the step-wise
    # rollout is launch-latency bound there, ForwardRollout's default takes the one-launch kernel (k_horizon) instead
    try:
        small = {}
        for ns_s in (20, 200, 2000):
            frs = ForwardRollout(configs.car_residual_fs(ns_s, HORIZON, with_derivatives=True), condition=True, device=device)
            us, es = synthetic_inputs(ns_s, HORIZON, 3, 3)
            us, es = us.to(device), es.to(device)
            row = {}
            for label, mode in (("step_wise_ms", False), ("default_ms", "auto")):
                frs.use_fused_horizon(mode)
                best = None
                for _ in range(4):
                    torch.cuda.synchronize()
                    e0.record()
                    tr = frs.run(us, es)
                    e1.record()
                    torch.cuda.synchronize()
                    ms_s = e0.elapsed_time(e1)
                    best = ms_s if best is None else min(best, ms_s)
                row[label] = best
                row[label.replace("_ms", "_one_launch")] = bool(frs.engine.get_option("last_rollout_fused"))
                if mode is False:
                    ref_tr = tr.clone()
            row["bit_identical"] = bool(torch.equal(tr, ref_tr))
            row["default_sample_steps_per_sec"] = ns_s * HORIZON / (row["default_ms"] * 1e-3)
            small[str(ns_s)] = row
            del frs
        out["small_batch_car_rollout"] = dict(small, what="conditioned car rollout, 50 steps, at the reference's own sample counts: "
                                              "host-observed ms per rollout incl. launch latency (best of 4), step-wise vs "
                                              "ForwardRollout's default (one k_horizon launch while a warp gets <= 4 samples)")
    except Exception as exc:  # noqa: BLE001
        out["small_batch_car_rollout"] = {"error": repr(exc)}

    # the other large-batch rollout of the reference: benchmarking/simulate_true_reachable_set.py (2-D pendulum, real data WITH
    # derivative observations: m = 180, g_ny = 2, d = 3, T = 4, 30 steps; 20 samples x 10^4 repeats = 2e5 samples)
    try:
        ns2, st2 = 200_000, 30
        fr = ForwardRollout(configs.pendulum2D_rollout(ns2, st2), condition=True, device=device, agent_size=20)
        g2 = torch.Generator().manual_seed(5)
        eps2 = torch.randn(st2, ns2, 2, 1, 4, generator=g2, dtype=torch.float64).clamp_(-2.5, 2.5).to(device)
        u2 = (2.0 * torch.sin(torch.linspace(0, 3, st2, dtype=torch.float64))).reshape(st2, 1).to(device)
        fr.run(u2, eps2)
        torch.cuda.synchronize()
        e0.record()
        fr.run(u2, eps2)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1)
        wb, wf = fr.engine.last_launch_work()
        out["pendulum_true_reachable_set"] = {
            "sample_steps_per_sec": ns2 * st2 / (ms2 * 1e-3), "ms_per_rollout": ms2, "ns": ns2, "steps": st2,
            "algorithmic_GBps": wb / ms2 / 1e6, "algorithmic_GFLOPs": wf / ms2 / 1e6, "engine_status": fr.engine.status(),
            "shape": "g_ny=2, d=3, T=4, m=180 (derivative observations), conditioning on, params_pendulum.yaml's zero-variance "
                     "rule and min-distance filter (1e-4) on, 10^4 reference Agents of 20 samples (per-Agent reductions on the "
                     "device); shared rows by the batched GEMM (K1a)"}
        del fr, eps2
    except Exception as exc:  # noqa: BLE001
        out["pendulum_true_reachable_set"] = {"error": repr(exc)}
    torch.cuda.empty_cache()

    params = configs.pendulum1D_sqp()
    agent = Agent(params, generate_base_samples=False, device=device)
    g = torch.Generator().manual_seed(0)
    agent.epistimic_random_vector = torch.randn(8, 1, 70, 1, 17, 3, generator=g, dtype=torch.float64).clamp(-2.5, 2.5).to(device)
    H = 17
    rng = np.random.default_rng(0)
    x_h = np.tile(np.stack([np.linspace(2.2, 3.1, H), np.linspace(2.0, 0.1, H)], 1), (1, 70)) + 0.01 * rng.standard_normal((H, 140))
    u_h = np.linspace(-3, 3, H).reshape(H, 1)
    times, dev_times = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iterates = []
    for i in range(8):
        iterates.append(x_h.copy())
        agent.mpc_iteration(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        agent.train_hallucinated_dynGP(0)
        agent.dyn_fg_jacobians(agent.get_batch_x_hat(x_h, u_h), 0)
        e1.record()
        times.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        dev_times.append(e0.elapsed_time(e1))
        x_h = x_h + 0.005 * rng.standard_normal(x_h.shape)
    through_ref = None
    try:  # the UNMODIFIED reference Agent (baseline/_ref copy) on the shim, in its own process
        drv = subprocess.run([sys.executable, os.path.join(REPO, "tests", "ref_agent_driver.py"), "pendulum1D_sqp", "--time", "5"],
                             capture_output=True, text=True, timeout=600)
        d = json.loads([l for l in drv.stdout.splitlines() if l.startswith("{")][-1])
        through_ref = d.get("ms_per_linearisation", d)
        if isinstance(through_ref, dict) and "median" in through_ref:
            through_ref = {"median_ms": through_ref["median"], "min_ms": through_ref["min"], "ns": 12,
                           "worst_error_over_tolerance": max(d["worst_over_tolerance"].values()),
                           "what": through_ref["what"] + " (fixture shape: ns=12, H=17; most of it is the reference's own "
                                   "Python: model construction, tiling, asserts with device syncs)"}
    except Exception as exc:  # noqa: BLE001
        through_ref = {"error": repr(exc)}
    out["sqp_linearisation_ms"] = {"host_observed_ms_incl_d2h": float(np.median(times[2:])),
                                   "through_unmodified_agent": through_ref,
                                   "device_ms": float(np.median(dev_times[2:])),
                                   "shape": "pendulum1D: ns=70, H=17, T=3 (q=51), n_obs=87",
                                   "calls": "train_hallucinated_dynGP + dyn_fg_jacobians (solver.py:84-94)"}
    # SQP mode at the car-residual shape (BASELINE configs[2]): q = 150 joint scalars per call, +150 factor rows per iteration
    try:
        from sampling_gpmpc_b200.agent import gp_hypers_from_params
        from sampling_gpmpc_b200.engine import GPEngine
        from sampling_gpmpc_b200.envs import make_env_spec as _mk
        cp = configs.car_residual_fs(20, 50, with_derivatives=True)
        sp = _mk(cp)
        Xc, Yc = sp.initial_training_data(cp)
        Hc, its = 50, 6
        eng = GPEngine(20, 3, 2, 3, Xc.shape[0], cap_points=Hc * its, device=device)
        ls, os_, nz = gp_hypers_from_params(cp, 3, 2, use_grad=True)
        eng.set_hypers(ls, os_, nz, 1e-9)
        eng.set_real_data(Xc, Yc)
        gd = torch.Generator(device=device).manual_seed(0)
        base = torch.stack([torch.linspace(-0.9, 0.9, Hc), torch.linspace(-0.5, 0.5, Hc)], 1).to(device, torch.float64)
        xq = (base[None, None] + 0.05 * torch.randn(20, 1, Hc, 2, generator=gd, dtype=torch.float64, device=device)).expand(20, 3, Hc, 2).contiguous()
        # untimed first call: CUDA loads each kernel of the 7 MB library lazily at its first launch and the workspaces are
        # allocated (cudaMalloc synchronises) -- 264 ms in round 1's device_ms_by_sqp_iteration[0]; not part of an iteration
        ee = torch.randn(20, 3, Hc, 3, generator=gd, dtype=torch.float64, device=device).clamp(-3, 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, yq, _ = eng.posterior(xq, ee, eng.opts(beta=3.0))
        eng.append(xq, yq)
        torch.cuda.synchronize()
        first_call_ms = (time.perf_counter() - t0) * 1e3
        eng.reset_hallucinated()
        per_it = []
        for it in range(its):
            ee = torch.randn(20, 3, Hc, 3, generator=gd, dtype=torch.float64, device=device).clamp(-3, 3)
            torch.cuda.synchronize()
            e0.record()
            eng.set_option("prefactor_next", 1)  # the append below follows: its Cholesky runs beside the draw's (as in gpmpc_linearise)
            _, _, yq, _ = eng.posterior(xq, ee, eng.opts(beta=3.0))
            eng.append(xq, yq)
            e1.record()
            torch.cuda.synchronize()
            per_it.append(round(e0.elapsed_time(e1), 3))
            xq = (xq + 0.03 * torch.randn(20, 1, Hc, 2, generator=gd, dtype=torch.float64, device=device)).contiguous()
        out["sqp_linearisation_car_ms"] = {"device_ms_by_sqp_iteration": per_it, "factor_rows_before_call": [150 * i for i in range(its)],
                                           "shape": "car residual: ns=20, g_ny=3, H=50, T=3 (q=150), m=45; model call + conditioning",
                                           "engine_status": eng.status(), "first_call_ms_incl_lazy_module_load_and_allocation": round(first_call_ms, 2)}
        del eng
    except Exception as exc:  # noqa: BLE001
        out["sqp_linearisation_car_ms"] = {"error": repr(exc)}
    if cpu_leg:
        # the same calls through the oracle's Agent restatement (full re-fit per call, all host threads)
        from oracle.agent_ref import RefAgent
        from sampling_gpmpc_b200.envs import make_env_spec
        spec = make_env_spec(params)
        Xr, Yr = spec.initial_training_data(params)
        ref = RefAgent(params, spec, Xr, Yr, epistimic_random_vector=agent.epistimic_random_vector.cpu())
        ct = []
        for i in range(5):
            ref.mpc_iteration(i)
            t0 = time.perf_counter()
            ref.train_hallucinated_dynGP(0)
            ref.dyn_fg_jacobians(ref.get_batch_x_hat(iterates[i], u_h), 0)
            ct.append((time.perf_counter() - t0) * 1e3)
        out["sqp_linearisation_ms"]["cpu_port_ms"] = float(np.median(ct[1:]))
        out["sqp_linearisation_ms"]["cpu_threads"] = torch.get_num_threads()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sampling_gpmpc_b200 import configs
    from sampling_gpmpc_b200.rollout import ForwardRollout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ns = args.ns_per_gpu
    params = configs.car_residual_fs(ns * world, HORIZON, with_derivatives=True)
    fr = ForwardRollout(params, condition=True, rank=rank, world_size=world, device=device)
    eng = fr.engine
    u_host, eps_host = synthetic_inputs(ns, HORIZON, 3, 1000 + rank)
    u_dev, eps_dev = u_host.to(device), eps_host.to(device)
    traj = torch.empty((ns, 4, HORIZON + 1), dtype=torch.float64, device=device)
    traj_host = torch.empty(traj.shape, dtype=torch.float64).pin_memory()

    # ---- device-resident timing ("value") ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        fr.run(u_dev, eps_dev, traj)
    barrier()
    eng.set_timing(True)
    launches0 = eng.launch_count
    kern_ms = kern_launches = 0
    work_bytes = work_flops = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            fr.run(u_dev, eps_dev, traj)
            # summed after the loop would need per-step event sets; reading them waits only for this rollout
            ms, nl = eng.rollout_kernel_ms()
            kern_ms += ms
            kern_launches += nl
            b, f = eng.last_launch_work()
            work_bytes += b
            work_flops += f
        e1.record()
        barrier()
    eng.set_timing(False)
    gpu_launches = eng.launch_count - launches0
    t_dev = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total = float(t_dev.item())
    ms_per_step = ms_total / args.steps
    value = world * ns * HORIZON / (ms_per_step * 1e-3)
    status = eng.status()

    # ---- end-to-end through the public call with HOST buffers ("e2e") ----------------------------------
    def e2e_step():
        # the public host-buffer call: base samples pinned host -> device (streamed in behind the first horizon
        # steps), rollout, trajectories device -> pinned host
        fr.run_from_host(u_host, eps_host, traj, traj_host)
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * ns * HORIZON / (float(t_e2e.item()) / args.steps * 1e-3)
    h2d = eps_host.numel() * 8 + u_host.numel() * 8
    d2h = traj_host.numel() * 8

    # ---- the one collective of the path and its consumers (N > 1): rollout -> NCCL all-gather of the trajectories ->
    #      per-stage boxes / max-deviation tightening -> per-stage convex hulls (generate_convex_hull.py:83-104) ---------------
    gathered = None
    if world > 1:
        def gathered_step():
            fr.run(u_dev, eps_dev, traj)
            allt = fr.all_gather_trajectories(traj)  # (ns_total, nx, H+1) on every rank
            lo, hi = eng.traj_stats(allt)
            hulls = eng.stage_hulls(allt, 0, 1)      # ends with the device->host copy of the hull candidates
            return allt, lo, hi, hulls
        gathered_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            allt, lo, hi, hulls = gathered_step()
        e1.record()
        barrier()
        t_g = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
        e0.record()
        for _ in range(args.steps):
            fr.all_gather_trajectories(traj)
        e1.record()
        barrier()
        t_ag = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t_ag, op=dist.ReduceOp.MAX)
        # the same consumers without the gather: every rank reduces its own shard, only the reductions travel
        # (ForwardRollout.stage_boxes / stage_hulls: an all-reduce of 2 x (nx, H+1) doubles, an all-gather of the hull vertices)
        def reduced_step():
            fr.run(u_dev, eps_dev, traj)
            lo_r, hi_r = fr.stage_boxes(traj)
            return lo_r, hi_r, fr.stage_hulls(traj, 0, 1)
        lo_r, hi_r, hulls_r = reduced_step()
        reduced_ok = bool(torch.equal(lo_r, lo)) and bool(torch.equal(hi_r, hi)) and len(hulls_r) == len(hulls) and \
            all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(hulls_r, hulls))
        barrier()
        e0.record()
        for _ in range(args.steps):
            reduced_step()
        e1.record()
        barrier()
        t_r = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
        gathered = {"value": world * ns * HORIZON / (float(t_g.item()) / args.steps * 1e-3), "unit": UNIT,
                    "ms_per_step": float(t_g.item()) / args.steps,
                    "all_gather_ms": float(t_ag.item()) / args.steps, "all_gather_bytes": int(allt.numel() * 8),
                    "hull_vertices_stage_last": int(len(hulls[-1])),
                    "what": "rollout + all_gather_into_tensor of the trajectories + stage boxes + stage hulls, max over ranks",
                    "reduced": {"value": world * ns * HORIZON / (float(t_r.item()) / args.steps * 1e-3), "unit": UNIT,
                                "ms_per_step": float(t_r.item()) / args.steps, "equal_to_gathered": reduced_ok,
                                "what": "rollout + the same consumers WITHOUT the gather: local boxes / hulls per rank, all-reduce "
                                        "of the boxes, all-gather of the hull vertices only (bit-identical results)"}}
        del allt
        # sharded vs single GPU, bit for bit, at a size one GPU holds: rank 0 rolls the WHOLE population out alone
        ns_chk = 2048
        p_chk = configs.car_residual_fs(ns_chk * world, HORIZON, with_derivatives=True)
        u_c, eps_c = synthetic_inputs(ns_chk * world, HORIZON, 3, 4242)  # same seed on every rank: the global base samples
        fr_c = ForwardRollout(p_chk, condition=True, rank=rank, world_size=world, device=device)
        part = fr_c.run(u_c.to(device), eps_c.to(device))
        whole = fr_c.all_gather_trajectories(part)
        if rank == 0:
            fr_1 = ForwardRollout(p_chk, condition=True, device=device)
            single = fr_1.run(u_c.to(device), eps_c.to(device))
            gathered["sharded_bit_identical_to_single_gpu"] = bool(torch.equal(whole, single))
            gathered["self_check_samples"] = ns_chk * world
            del fr_1, single
        del fr_c, part, whole
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = work_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    # DRAM bytes per step launch from the committed ncu capture of this same workload (profiles/, per round)
    traffic, traffic_src = None, None
    from sampling_gpmpc_b200.engine import load_library
    lib_version = load_library().gpmpc_version().decode()
    caps = sorted(f for f in os.listdir(os.path.join(REPO, "profiles")) if "traffic" in f and f.endswith(".json"))
    if caps and ns == 125000:
        cap = json.load(open(os.path.join(REPO, "profiles", caps[-1])))
        if cap.get("library_version") == lib_version:  # never quote a capture of another build
            traffic, traffic_src = cap["traffic_bytes_per_step_launch"], "profiles/" + caps[-1]
        else:
            traffic_src = f"no ncu DRAM capture of library {lib_version!r} committed (latest: profiles/{caps[-1]})"
    roofline = {"bound": "hbm", "kernel": "k_step<2,3> + k_step_finish<3> (fused rollout step: one gpmpc_step call)",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms_per_launch": kern_ms / max(kern_launches, 1),
                "launches_timed": kern_launches,
                "algorithmic_bytes_per_launch": work_bytes / max(kern_launches, 1),
                "algorithmic_gflop_per_launch": work_flops / max(kern_launches, 1) / 1e9,
                "kernel_share_of_step": kern_ms / ms_total}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(ns, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "engine_status": status,
            "state_bytes_per_gpu": int(eng.state_bytes)}
    if gathered is not None:
        line["e2e_gathered"] = gathered
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_rollout(4, 5, 1, threads)  # warm the CPU path
        dt, ref_traj = cpu_reference_rollout(args.cpu_samples, HORIZON, 7, threads, return_traj=True)
        try:
            line["parity_check"] = gpu_parity_against_cpu(args.cpu_samples, HORIZON, 7, ref_traj, device)
        except Exception as exc:  # noqa: BLE001
            line["parity_check"] = {"error": repr(exc)}
        line["cpu_baseline"] = {"value": args.cpu_samples * HORIZON / dt, "unit": UNIT, "cores": threads,
                                "kind": "port",
                                "sample": f"{args.cpu_samples} samples x {HORIZON} steps ({dt:.1f} s), oracle = reference-"
                                          "equivalent CPU torch path with full re-fit per step (gpytorch not installable)"}
    if world == 1 and not args.no_extras:
        del fr, eng, eps_dev, traj
        torch.cuda.empty_cache()
        try:
            line["extra"] = extras(torch, device, cpu_leg=not args.no_cpu_baseline)
        except Exception as exc:  # side measurements must never take the headline down
            line["extra"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
