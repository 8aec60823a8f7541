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
