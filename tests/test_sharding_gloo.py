"""CPU, world_size 2 over gloo: the multi-GPU host logic of the hot path.

The dynamics samples are independent, so the N>1 path is: contiguous shards of the sample index (global
indices kept), no data-path collective, ONE all-gather of the trajectories for the consumers, and -- only when
Dyn_gp_min_data_dist >= 0 -- an all-reduce of the per-(output, point) filter counts.  The arithmetic itself needs the GPU
(tests/test_gpu_parity.py); here the sharding, the gather layout and the flag reduction are checked with the
same functions the product calls, on CPU tensors.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sampling_gpmpc_b200.rollout import gather_padded, reduce_filter_counts, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns_global, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(ns_global, rank, world)
        # a "trajectory" whose value encodes its global sample index: (ns_local, nx, H+1)
        nx, H1 = 4, 6
        g = torch.arange(lo, hi, dtype=torch.float64)
        traj = g[:, None, None] * 100 + torch.arange(nx, dtype=torch.float64)[None, :, None] * 10 + torch.arange(H1, dtype=torch.float64)
        full = gather_padded(traj, ns_global, world)
        want = torch.arange(ns_global, dtype=torch.float64)[:, None, None] * 100 + \
            torch.arange(nx, dtype=torch.float64)[None, :, None] * 10 + torch.arange(H1, dtype=torch.float64)
        ok_gather = bool(torch.equal(full, want))
        # min-distance flags: per-sample filter (ns_local, g_ny, H) -> all/any over EVERY sample of every rank
        gen = torch.Generator().manual_seed(5)
        filt_global = torch.rand(ns_global, 2, 7, generator=gen) < 0.5
        filt_global[:, :, 0] = True    # point 0: filtered for all samples -> dropped
        filt_global[:, :, 1] = False   # point 1: filtered for none
        filt_global[:, 0, 2] = True    # point 2: all samples of ONE output -> dropped too (agent.py:186-191)
        counts = filt_global[lo:hi].sum(0).to(torch.int32)  # what gpmpc_filter_new_points hands back per shard
        flags = reduce_filter_counts(counts, ns_global, world)
        ok_flags = np.array_equal(flags[0], filt_global.all(0).any(0).numpy()) and \
            np.array_equal(flags[1], filt_global.reshape(-1, 7).any(0).numpy()) and bool(flags[0][0]) and \
            bool(flags[0][2]) and not flags[0][1]
        q.put((rank, lo, hi, ok_gather, ok_flags))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ns_global", [10, 7])
def test_two_rank_sharding_gather_and_flag_reduction(ns_global):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ns_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # contiguous, disjoint, covering, in rank order
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == ns_global
    assert all(r[3] for r in res), "all-gathered trajectories are not in global sample order"
    assert all(r[4] for r in res), "all/any flag reduction over ranks differs from the single-process result"


def test_shard_bounds_properties():
    for ns in (1, 2, 7, 70, 4000, 10**6):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(ns, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == ns
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) == -(-ns // world)  # weak scaling: the largest shard is ceil(ns / world)


def _resample_worker(rank, world, port, ns_global, q):
    from sampling_gpmpc_b200.rollout import resample_rejected
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        left, Xh, Yh = _resample_case(ns_global)
        lo, hi = shard_bounds(ns_global, rank, world)
        np.random.seed(11 if rank == 0 else 999)  # only rank 0's generator may matter
        changed, Xl, Yl, active = resample_rejected(left[lo:hi].clone(), Xh[lo:hi].clone(), Yh[lo:hi].clone(), ns_global, rank, world)
        q.put((rank, lo, hi, changed, Xl.numpy(), Yl.numpy(), None if active is None else active))
    finally:
        dist.destroy_process_group()


def _resample_case(ns_global):
    g = torch.Generator().manual_seed(21)
    left = torch.ones(ns_global, dtype=torch.int32)
    left[[1, ns_global - 2, ns_global - 1]] = 0          # rejected samples on both ranks
    Xh = torch.rand(ns_global, 2, 5, 3, generator=g, dtype=torch.float64)
    Yh = torch.rand(ns_global, 2, 5, 4, generator=g, dtype=torch.float64)
    Yh[0, 1, 2, 3] = float("nan")                        # a masked label of a SURVIVOR: must reach every rank's slot mask
    Yh[ns_global - 1, 0, 0, 0] = float("nan")            # ... and one of a rejected sample, which disappears with its data
    return left, Xh, Yh


@pytest.mark.parametrize("ns_global", [9, 8])
def test_two_rank_survivor_resampling_equals_single_process(ns_global):
    """prepare_dynamics_set's survivor resampling (src/agent.py:418-436) over a sharded population: all-gather of samples_left
    and of the data sets, rank 0's two np.random.choice draws broadcast -- every rank ends up with exactly its block of what
    ONE process computes from the same generator state, and with the population-wide NaN slot mask."""
    from sampling_gpmpc_b200.rollout import resample_rejected
    left, Xh, Yh = _resample_case(ns_global)
    np.random.seed(11)
    changed, X1, Y1, act1 = resample_rejected(left.clone(), Xh.clone(), Yh.clone(), ns_global, 0, 1)
    assert changed and not torch.equal(X1, Xh)
    # the reference's own statement sequence on the same generator state
    np.random.seed(11)
    remaining = np.arange(ns_global)[left.numpy() > 0]
    dead = np.nonzero(left.numpy() == 0)[0]
    Xr, Yr = Xh.clone(), Yh.clone()
    Xr[dead] = Xr[np.random.choice(remaining, dead.size)]
    Yr[dead] = Yr[np.random.choice(remaining, dead.size)]
    assert torch.equal(X1, Xr) and torch.equal(torch.nan_to_num(Y1, nan=-7.0), torch.nan_to_num(Yr, nan=-7.0))
    assert np.array_equal(act1, (~Yr.isnan().any(1).any(0)).reshape(-1).numpy().astype(np.uint8)) and act1.sum() < act1.size
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_resample_worker, args=(r, world, port, ns_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, lo, hi, ch, Xl, Yl, active in res:
        assert ch
        assert np.array_equal(Xl, X1[lo:hi].numpy())
        assert np.array_equal(np.nan_to_num(Yl, nan=-7.0), np.nan_to_num(Y1[lo:hi].numpy(), nan=-7.0))
        assert np.array_equal(active, act1)
    # nobody rejected / everybody rejected: nothing happens, as in the reference
    for pattern in (torch.ones(ns_global, dtype=torch.int32), torch.zeros(ns_global, dtype=torch.int32)):
        ch, X2, Y2, _ = resample_rejected(pattern, Xh, Yh, ns_global, 0, 1)
        assert not ch and X2 is Xh
