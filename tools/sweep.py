"""Synthetic scaling sweep (BASELINE.json configs[4], SURVEY.md 8d "config 5"): conditioned sample-steps/s of the
fused step over  ns x horizon x training-set size m x input dim d  on ONE GPU, each point with its algorithmic
GB/s and GFLOP/s (gpmpc_last_launch_work) against the measured HBM / FP64 peaks.

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.json]
    torchrun --nproc-per-node N tools/sweep.py --only-oversize   # N GPUs: the points whose factor state exceeds ONE GPU's HBM,
                                                                  # samples sharded contiguously (no collective on the data path)

Inputs as SURVEY.md states them: X_real ~ U[-1,1]^d (seed 0), y = sum sin(x_i) with the analytic gradient as
derivative observations when --grad-obs (m = n_real * T) else values only (m = n_real); lengthscale 1, outputscale 1,
noise 1e-6, jitter 1e-6; test path = random walk of step 0.05 per sample; eps ~ N(0,1) truncated at beta = 3.
Corners whose per-sample factor state exceeds the HBM budget are skipped and listed as such."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sampling_gpmpc_b200.engine import GPEngine

HBM_BUDGET = 150e9


def peaks():
    p = {"hbm_gbs": 6534.1, "fp64_tflops": 37.1}  # fp64: profiles/r1_fp64_peaks_b200.json (DMMA m8n8k4, measured)
    f = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(f):
        p["hbm_gbs"] = json.load(open(f)).get("hbm_gbs", p["hbm_gbs"])
    return p


def state_bytes(ns, g_ny, T, m, steps):
    mo, P = (m + 7) // 8 * 8, (steps * T + 7) // 8
    return ns * g_ny * 8.0 * 8 * (P * (mo + 8) + 4 * P * (P - 1))


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def run_point(ns_total, steps, n_real, d, g_ny, grad_obs, reps):
    T = d + 1
    per = -(-ns_total // WORLD)
    ns = max(0, min(per, ns_total - RANK * per))  # this rank's contiguous shard
    g = torch.Generator().manual_seed(0)
    X = torch.rand(n_real, d, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.full((g_ny, n_real, T), float("nan"), dtype=torch.float64)
    for j in range(g_ny):
        Y[j, :, 0] = torch.sin(X).sum(1)
        if grad_obs:
            Y[j, :, 1:] = torch.cos(X)
    eng = GPEngine(ns, g_ny, d, T, n_real, cap_points=steps)
    eng.set_hypers(np.ones((g_ny, d)), np.ones(g_ny), np.full((g_ny, T), 1e-6), 1e-6)
    eng.set_real_data(X, Y)
    gd = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.rand(ns, 1, 1, d, generator=gd, dtype=torch.float64, device="cuda") * 1.6 - 0.8
    walk = 0.05 * torch.randn(steps, ns, 1, 1, d, generator=gd, dtype=torch.float64, device="cuda")
    eps = torch.randn(steps, ns, g_ny, 1, T, generator=gd, dtype=torch.float64, device="cuda").clamp_(-3, 3)
    opts = eng.opts(beta=3.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best, work = None, None
    for rep in range(reps + 1):  # first pass = warm-up
        eng.reset_hallucinated()
        x = x0.clone()
        by = fl = 0.0
        torch.cuda.synchronize()
        e0.record()
        for t in range(steps):
            x = (x + walk[t]).clamp_(-1, 1)
            eng.step(x.expand(ns, g_ny, 1, d), eps[t], opts, want_moments=False)
            b_, f_ = eng.last_launch_work()
            by += b_
            fl += f_
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if WORLD > 1:  # the step loop of the slowest rank
            import torch.distributed as dist
            t = torch.tensor([ms, by, fl], dtype=torch.float64, device="cuda")
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, by, fl = float(tm[0]), float(t[1]), float(t[2])
        if rep > 0 and (best is None or ms < best):
            best, work = ms, (by, fl)
    st = eng.status()
    m = eng.num_real_observed
    sb = eng.state_bytes
    del eng
    torch.cuda.empty_cache()
    return {"ns": ns_total, "n_gpus": WORLD, "steps": steps, "n_real": n_real, "m": m, "d": d, "T": T, "g_ny": g_ny, "ms": best,
            "sample_steps_per_s": ns_total * steps / best * 1e3, "alg_GBps": work[0] / best / 1e6,
            "alg_GFLOPs": work[1] / best / 1e6, "state_GB": sb / 1e9, "status": st}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--g-ny", type=int, default=2)
    ap.add_argument("--max-seconds", type=float, default=400.0)
    ap.add_argument("--large-m", action="store_true", help="only the m = 1e4 points small enough for the block-kernel fallback")
    ap.add_argument("--only-oversize", action="store_true", help="only the points that do not fit one GPU (multi-GPU runs)")
    a = ap.parse_args()
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    pk = peaks()
    if a.quick:
        grid = [(1000, 10, 100, 2, False), (1000, 10, 250, 3, True), (200, 10, 1000, 2, False)]
    else:
        grid = []
        for d in (2, 3, 6):
            for n_real, grad in ((100, False), (1000, False), (1000 // (d + 1), True), (10000, False)):
                for steps in (10, 30, 100):
                    for ns in (100, 1000, 10000, 100000, 1000000):
                        grid.append((ns, steps, n_real, d, grad))
    rows, skipped, t_start = [], [], time.time()
    for ns, steps, n_real, d, grad in grid:
        m = n_real * (d + 1 if grad else 1)
        need = state_bytes(ns, a.g_ny, d + 1, m, steps)
        # rough cost model to keep the sweep bounded: skip points that would take > ~20 s
        est_flop = ns * a.g_ny * steps * (d + 1) * (m * m + m * steps * (d + 1) + (steps * (d + 1)) ** 2 / 3.0)
        if m > 4000 and not a.large_m:
            skipped.append({"ns": ns, "steps": steps, "m": m, "d": d,
                            "why": "m = 1e4 (shared rows by the k-slab GEMM, own rows by k_step_big): run separately with --large-m"})
            continue
        if a.large_m and (m <= 4000 or ns * a.g_ny * steps > 400000):
            continue
        if a.only_oversize and need <= HBM_BUDGET:
            continue
        if need > HBM_BUDGET * WORLD:
            skipped.append({"ns": ns, "steps": steps, "m": m, "d": d,
                            "why": f"factor state {need / 1e9:.0f} GB > HBM of {WORLD} GPU(s)"})
            continue
        elapsed = time.time() - t_start
        if WORLD > 1:  # every rank must take the same decision
            import torch.distributed as dist
            te = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            elapsed = float(te[0])
        if est_flop > 2e14 * WORLD or elapsed > a.max_seconds:
            skipped.append({"ns": ns, "steps": steps, "m": m, "d": d, "why": "time budget of the sweep"})
            continue
        try:
            r = run_point(ns, steps, n_real, d, a.g_ny, grad, 1 if ns * steps >= 10 ** 6 else 2)
        except Exception as e:  # noqa: BLE001  (a failing corner must not hide the others)
            skipped.append({"ns": ns, "steps": steps, "m": m, "d": d, "why": f"error: {e}"[:200]})
            torch.cuda.empty_cache()
            continue
        r["hbm_frac"] = r["alg_GBps"] / pk["hbm_gbs"] / WORLD   # fractions of the aggregate peaks of the GPUs used
        r["fp64_frac"] = r["alg_GFLOPs"] / 1e3 / pk["fp64_tflops"] / WORLD
        r["bound"] = "hbm" if r["hbm_frac"] >= r["fp64_frac"] else "fp64"
        rows.append(r)
        if RANK == 0:
          print("ns=%-8d steps=%-4d m=%-6d d=%d T=%d  %9.2f ms  %12.0f sample-steps/s  %7.0f GB/s (%.2f)  %8.0f GFLOP/s (%.3f)  st=%d"
              % (ns, steps, r["m"], d, r["T"], r["ms"], r["sample_steps_per_s"], r["alg_GBps"], r["hbm_frac"],
                 r["alg_GFLOPs"], r["fp64_frac"], r["status"]), flush=True)
    if RANK != 0:
        return
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"peaks": pk, "g_ny": a.g_ny, "rows": rows, "skipped": skipped,
               "n_gpus": WORLD,
               "note": f"{WORLD} GPU(s), samples sharded contiguously, max over ranks; per point best of the timed passes after one warm-up pass; CUDA events around the "
                       "whole step loop (host launch overhead included: small ns x steps points are launch-bound)"},
              open(a.out, "w"), indent=1)
    print(f"{len(rows)} points, {len(skipped)} skipped -> {a.out}")


if __name__ == "__main__":
    main()
