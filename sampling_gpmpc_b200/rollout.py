"""Forward rollouts of sampled GP dynamics -- the large-batch callers of the hot path.

``ForwardRollout`` is the loop of benchmarking/simulate_forward_sampling_car.py:117-138 (and, with
``condition=True``, of benchmarking/simulate_true_reachable_set.py:179-259): per horizon step build
[x, u (+ feedback)], evaluate every sampled dynamics function at its own state, draw, (condition,) and
take the value as the next state.  The whole horizon is queued on one CUDA stream by gpmpc_rollout;
nothing returns to the host until the trajectories are read.

Multi-GPU: the dynamics samples are independent, so each rank owns a contiguous block of
the sample index and there is NO collective on the data path; ``all_gather_trajectories`` is the single
exchange, for consumers that need every trajectory (convex hulls, constraint tightening).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .agent import gp_hypers_from_params
from .engine import GPEngine, hull2d, make_env_struct
from .envs import make_env_spec

F64 = torch.float64


def shard_bounds(ns_global: int, rank: int, world_size: int):
    """Contiguous block of sample indices owned by ``rank`` (global indices are kept, so gathered arrays
    are laid out exactly like the single-GPU result)."""
    per = -(-ns_global // world_size)
    return min(rank * per, ns_global), min((rank + 1) * per, ns_global)


class ForwardRollout:
    def __init__(self, params: dict, condition: bool, rank: int = 0, world_size: int = 1,
                 device: Optional[torch.device] = None, X_real=None, Y_real=None, agent_size: Optional[int] = None):
        """agent_size: benchmarking/simulate_true_reachable_set.py rolls out a NEW Agent of num_dyn_samples samples per
        repeat; here all repeats are one batch of `num_dyn_samples` samples in total and `agent_size` says how many
        consecutive samples form one reference Agent.  It matters iff Dyn_gp_min_data_dist >= 0: the min-distance filter of
        update_hallucinated_Dyn_dataset and GPyTorch's any-over-batch NaN mask couple the samples of one Agent
        (src/agent.py:164-202); default: the whole batch is one Agent (as in the reference's own single-Agent runs)."""
        self.params = params
        ag = params["agent"]
        self.spec = make_env_spec(params)
        self.ns_global = ag["num_dyn_samples"]
        self.rank, self.world_size = rank, world_size
        self.s_lo, self.s_hi = shard_bounds(self.ns_global, rank, world_size)
        self.ns = self.s_hi - self.s_lo
        self.steps = params["common"]["num_MPC_itrs"]
        self.T = 1 if params["env"]["use_model_without_derivatives"] else 1 + self.spec.d
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if X_real is None:
            X_real, Y_real = self.spec.initial_training_data(params)
        Y_real = Y_real[:, :, : self.T] if self.T == 1 else Y_real
        self.engine = GPEngine(self.ns, self.spec.g_ny, self.spec.d, self.T, X_real.shape[0],
                               cap_points=self.steps, device=self.device)
        self.engine.set_condition_on_hallucinated(condition)  # before any allocation: no factor rows if off
        ls, os_, noise = gp_hypers_from_params(params, self.spec.g_ny, self.spec.d, use_grad=self.T > 1)
        self.engine.set_hypers(ls, os_, noise, ag["Dyn_gp_jitter"])
        self.engine.set_real_data(X_real, Y_real.contiguous())
        self.condition = condition
        self.agent_size = None
        if condition and ag["Dyn_gp_min_data_dist"] >= 0.0:
            self.agent_size = int(agent_size) if agent_size else self.ns_global
            if world_size > 1 and self.s_lo % self.agent_size:
                raise ValueError("sharding must not cut a reference Agent in two: ceil(ns / world_size) % agent_size != 0")
            self.engine.set_grouping(self.agent_size, ag["Dyn_gp_min_data_dist"])
        fb = ag.get("feedback", {}).get("use", False)
        K = np.asarray(params["optimizer"]["terminal_tightening"]["K"]) if fb else None
        self.env = make_env_struct(self.spec, K, params["env"]["goal_state"] if fb else None)
        self.opts = self.engine.opts(ag["Dyn_gp_beta"], ag["Dyn_gp_variance_is_zero"])
        self.x0 = torch.tensor(params["env"]["start"], dtype=F64, device=self.device).expand(self.ns, -1).contiguous()
        self.use_fused_horizon("auto")

    def run(self, u_ff: torch.Tensor, eps: torch.Tensor, traj: Optional[torch.Tensor] = None,
            check: bool = False) -> torch.Tensor:
        """u_ff (steps, nu); eps (steps, ns_global or ns, g_ny, 1, T) standard-normal draws (the reference's
        epistimic_random_vector[:, 1]).  Returns this rank's trajectories (ns, nx, steps+1) on the device.  The call
        only queues work; `check=True` (or `check()` later) waits for it and raises NotPSDError where GPyTorch would.

        Small batches (the reference's own sample counts, up to ~2000 samples) run as ONE launch of the fused-horizon
        kernel (use_fused_horizon: automatic by default), which cannot take GPyTorch's batch-wide eigen-root fallback of a
        failed jitter ladder in flight: for those the status word is read right away (one 4-byte copy behind a ~1 ms
        kernel) and a rollout that hit a failed ladder is repeated step-wise, so the result is the step-wise one either way."""
        if eps.shape[1] == self.ns_global and self.world_size > 1:
            eps = eps[:, self.s_lo:self.s_hi]
        self.engine.reset_hallucinated()
        out = self.engine.rollout(self.env, self.x0, u_ff, eps, self.opts, traj)
        if self._fused_mode and self.engine.get_option("last_rollout_fused"):
            from .engine import ST_SAMPLE_NOT_PD
            if self.engine.status() & ST_SAMPLE_NOT_PD:  # (a NaN flagged next to it is the failed draw's own, carried forward)
                self.engine.status(clear=True)
                self.engine.set_option("rollout_fused", 0)
                try:
                    self.engine.reset_hallucinated()
                    self.engine.rollout(self.env, self.x0, u_ff, eps, self.opts, out)
                finally:
                    self.engine.set_option("rollout_fused", self._fused_mode)
        if check:
            self.check()
        return out

    def use_fused_horizon(self, mode="auto"):
        """The one-launch rollout kernel (csrc/gpmpc_horizon.cuh; bit-identical results): "auto" (default) = only where the
        step-wise rollout is launch-latency bound (1.2 - 2.5x faster there; slower at the bench shape),
        True = wherever the shape allows, False = never."""
        self._fused_mode = 2 if mode == "auto" else int(bool(mode))
        self.engine.set_option("rollout_fused", self._fused_mode)

    def check(self) -> int:
        """Synchronises and turns the device status word into the reference's error behaviour (engine.raise_on_status)."""
        return self.engine.raise_on_status()

    def run_from_host(self, u_ff: torch.Tensor, eps_host: torch.Tensor, traj: Optional[torch.Tensor] = None,
                      traj_host: Optional[torch.Tensor] = None, chunk_steps: int = 5) -> torch.Tensor:
        """`run` for base samples that live in pinned host memory (the reference generates them on the host,
        src/agent.py:76-104): eps_host (steps, ns, g_ny, 1, T) pinned.  The upload streams in chunks of `chunk_steps`
        horizon steps on a side stream while the rollout runs; with `traj_host` (pinned) the trajectories are copied
        back behind the last step.  Returns the device trajectories (ns, nx, steps+1)."""
        steps = u_ff.shape[0]
        eps2 = eps_host.reshape(steps, -1)
        if getattr(self, "_eps_dev", None) is None or self._eps_dev.shape != eps2.shape:
            self._eps_dev = torch.empty(eps2.shape, dtype=F64, device=self.device)
            self._copy_stream = torch.cuda.Stream(self.device)
        self.engine.reset_hallucinated()
        out = self.engine.rollout_from_host(self.env, self.x0, u_ff, eps2, self.opts, self._eps_dev, self._copy_stream,
                                            traj, chunk_steps)
        if traj_host is not None:
            traj_host.copy_(out, non_blocking=True)
        return out

    def all_gather_trajectories(self, traj: torch.Tensor) -> torch.Tensor:
        """(ns_local, nx, steps+1) per rank -> (ns_global, nx, steps+1) on every rank: one NCCL all-gather."""
        return gather_padded(traj, self.ns_global, self.world_size)

    # ---- the trajectory consumers WITHOUT the gather: each rank reduces its own shard, only the reductions travel ----------
    def stage_boxes(self, traj: torch.Tensor, ref: Optional[torch.Tensor] = None):
        """Per-stage bounding box [and max-deviation tightening max_n |x_k^n - ref_k|] of ALL ranks' trajectories
        (generate_convex_hull.py:83-87; extra/approx_sampling_mpc/README.md:19-27) from this rank's shard traj
        (ns_local, nx, H1): min / max are exact and order-free, so the local reduction (gpmpc_traj_stats) followed by an
        all-reduce of the (nx, H1) results -- 3 x 1.6 KB at the car shape -- is bit-identical to the reduction over the
        gathered array (1.63 GB at 10^6 samples).  Returns (lo, hi[, max_dev]), the same on every rank."""
        out = self.engine.traj_stats(traj, ref)
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(out[0], op=dist.ReduceOp.MIN)
            for t in out[1:]:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out

    def stage_hulls(self, traj: torch.Tensor, i0: int = 0, i1: int = 1, max_vertices: int = 512):
        """Per-stage convex hull of (traj[:, i0, t], traj[:, i1, t]) over ALL ranks' samples (generate_convex_hull.py:88-104)
        from this rank's shard: hull(union of the shards) = hull(union of the shards' hulls), so every rank takes the hull of its
        own samples (gpmpc_stage_hulls), the hull VERTICES (coordinates + global sample index; tens of points per stage) are
        all-gathered, and the final hull of the few candidates is taken on the host (gpmpc_hull2d: the same monotone chain,
        identical points resolved to the lowest global sample index) -- the same vertex lists as the hull of the gathered
        array.  Returns a list over the stages of GLOBAL sample-index arrays, counter-clockwise, the same on every rank."""
        local = self.engine.stage_hulls(traj, i0, i1, max_vertices)
        if self.world_size == 1:
            return local
        import torch.distributed as dist
        H1 = traj.shape[2]
        cnt = torch.tensor([max(len(v) for v in local)], dtype=torch.int64, device=traj.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
        cmax = int(cnt.item())
        idx = np.zeros((H1, cmax), dtype=np.int64)
        live = np.zeros((H1, cmax), dtype=bool)
        for t, v in enumerate(local):
            idx[t, : len(v)] = v
            live[t, : len(v)] = True
        idx_d = torch.from_numpy(idx).to(traj.device)
        tt = torch.arange(H1, device=traj.device)[:, None].expand(H1, cmax)
        pack = torch.stack([traj[idx_d, i0, tt], traj[idx_d, i1, tt], (idx_d + self.s_lo).to(F64)], dim=-1)  # (H1, cmax, 3)
        pack[~torch.from_numpy(live).to(traj.device)] = float("nan")
        allp = torch.empty((self.world_size, H1, cmax, 3), dtype=F64, device=traj.device)
        dist.all_gather_into_tensor(allp, pack.contiguous())
        allp = allp.permute(1, 0, 2, 3).reshape(H1, self.world_size * cmax, 3).cpu().numpy()
        return [merge_shard_hulls(allp[t]) for t in range(H1)]


def save_X_traj(traj: torch.Tensor, save_dir: str, epistemic_idx: int, chunk: Optional[int] = None) -> list:
    """On-disk hand-off of benchmarking/simulate_forward_sampling_car.py:157-161: ``data_X_traj_<k>.pkl`` = pickle of the
    numpy array (ns, nx, H+1), which generate_convex_hull.py:77-84 and the plotting scripts load and ``np.vstack``.  The
    reference writes one file per job of 4000 samples (2500 jobs); here ONE rollout holds every sample, so ``chunk``
    splits it into files ``k, k+1, ...`` of ``chunk`` samples each for consumers that expect that many (default: one
    file).  The copy goes through one pinned staging buffer.  Returns the paths written."""
    import os
    import pickle
    ns = traj.shape[0]
    chunk = ns if not chunk else int(chunk)
    host = torch.empty(traj.shape, dtype=traj.dtype, pin_memory=traj.is_cuda)
    host.copy_(traj, non_blocking=False)
    arr = host.numpy()
    os.makedirs(save_dir, exist_ok=True)
    paths = []
    for i, lo in enumerate(range(0, ns, chunk)):
        path = os.path.join(save_dir, f"data_X_traj_{epistemic_idx + i}.pkl")
        with open(path, "wb") as f:
            pickle.dump(np.ascontiguousarray(arr[lo:lo + chunk]), f)
        paths.append(path)
    return paths


def gather_padded(local: torch.Tensor, ns_global: int, world_size: int) -> torch.Tensor:
    """All-gather of per-rank sample blocks (shard_bounds layout; the last shard may be shorter) into the
    single-process layout: row s of the result is global sample s on every rank."""
    if world_size == 1:
        return local
    import torch.distributed as dist
    per = -(-ns_global // world_size)
    pad = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world_size * per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return out[:ns_global]



def merge_shard_hulls(cand: np.ndarray) -> np.ndarray:
    """cand (n, 3): x, y, global sample index of the hull vertices of every shard (NaN rows = padding).  hull(union of the
    shards) = hull(union of the shards' hulls): the exact host hull (gpmpc_hull2d, monotone chain) of the candidates, taken in
    the order of their global sample index so that identical points resolve to the lowest one -- exactly what the hull over
    all samples returns.  -> global sample indices of the vertices, counter-clockwise."""
    c = cand[~np.isnan(cand[:, 2])]
    c = c[np.argsort(c[:, 2], kind="stable")]
    pos = hull2d(np.ascontiguousarray(c[:, :2]))
    return c[pos, 2].astype(np.int32)


def resample_rejected(samples_left: torch.Tensor, Xh: torch.Tensor, Yh: torch.Tensor, ns_global: int, rank: int, world_size: int):
    """Survivor resampling of Agent.prepare_dynamics_set (src/agent.py:418-436) over a SHARDED population (SURVEY.md 8e:
    "all-gather of samples_left + index broadcast").  samples_left (ns_local,) int, Xh (ns_local, g_ny, n, d) / Yh
    (ns_local, g_ny, n, T) = this rank's hallucinated sets.  The reference replaces the data of every rejected sample by
    that of a random surviving one -- two independent np.random.choice draws from numpy's global generator, one for the
    inputs and one for the labels, over the WHOLE population: here samples_left and the data sets are all-gathered, rank 0
    makes the two draws exactly as the reference does (its generator is THE generator) and broadcasts them, and every rank
    rewrites its own rejected samples.  Returns (changed, Xh, Yh, active): changed = False if nobody or everybody was
    rejected (nothing to do, like the reference); active (n * T,) uint8 = GPyTorch's any-over-batch slot mask of the new
    label set, over every rank's samples."""
    left = gather_padded(samples_left.reshape(-1, 1), ns_global, world_size).reshape(-1).cpu().numpy()
    if not (left.sum() > 0 and (left == 0).any()):
        return False, Xh, Yh, None
    n_rep = int((left == 0).sum())
    remaining = np.arange(ns_global)[left > 0]
    if world_size > 1:
        import torch.distributed as dist
        pick = torch.zeros(2, n_rep, dtype=torch.int64)
        if rank == 0:
            pick[0] = torch.from_numpy(np.random.choice(remaining, n_rep))
            pick[1] = torch.from_numpy(np.random.choice(remaining, n_rep))
        pick = pick.to(Xh.device)
        dist.broadcast(pick, src=0)
        Xg, Yg = gather_padded(Xh, ns_global, world_size), gather_padded(Yh, ns_global, world_size)
    else:
        pick = torch.stack([torch.from_numpy(np.random.choice(remaining, n_rep)),
                            torch.from_numpy(np.random.choice(remaining, n_rep))]).to(Xh.device)
        Xg, Yg = Xh, Yh
    dead = torch.as_tensor(np.nonzero(left == 0)[0], device=Xh.device)
    Xg, Yg = Xg.clone(), Yg.clone()
    # (advanced indexing reads the right-hand side before it writes: a rejected sample never copies from another rejected one
    # anyway, the sources are survivors)
    Xg[dead] = Xg[pick[0]]
    Yg[dead] = Yg[pick[1]]
    active = (~Yg.isnan().any(1).any(0)).reshape(-1).cpu().numpy().astype(np.uint8)
    lo, hi = shard_bounds(ns_global, rank, world_size)
    return True, Xg[lo:hi].contiguous(), Yg[lo:hi].contiguous(), active


def reduce_filter_counts(counts: torch.Tensor, ns_global: int, world_size: int) -> np.ndarray:
    """counts (g_ny, H) int32 = this shard's samples whose new point h was filtered for output j
    (gpmpc_filter_new_points).  Sum over ranks, then the reference's flags (src/agent.py:186-191 and GPyTorch's
    any-over-batch NaN mask): returns bool (2, H): [0] point filtered for ALL samples of some
    output (drop it), [1] filtered for ANY batch element (keep it in the data set, mask it in the factor)."""
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    c = counts.cpu().numpy()
    return np.stack([(c == ns_global).any(0), (c > 0).any(0)])
