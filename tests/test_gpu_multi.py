"""GPU, only where at least two GPUs are visible (skipped on a one-GPU box): SURVEY.md 8(e) over NCCL -- every rank rolls
out its contiguous shard, one all-gather, and the gathered trajectories / stage boxes / hull vertices are BIT-IDENTICAL to
the same rollout on a single GPU (tools/multi_gpu_check.py; the host-side sharding logic is covered on CPU with gloo)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("ns", [2001])
def test_sharded_rollout_and_all_gather_equal_single_gpu(ns):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tools", "multi_gpu_check.py"), str(ns)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["gathered_equals_single_gpu"] and res["traj_stats_equal"] and res["hulls_equal"]
    # the consumers without the gather (local reduction, then all-reduce of the boxes / all-gather of the hull vertices only)
    assert res["reduced_boxes_equal_gathered"] and res["reduced_hulls_equal_gathered"]
    assert res["status"] == [0, 0]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_rejection_rollout_equals_single_gpu():
    """prepare_dynamics_set over two ranks (all-gather of samples_left / the data sets, rank 0's resampling draws broadcast):
    survivors, data sets and the restored model bit-identical to one GPU (tools/multi_gpu_rejection_check.py)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tools", "multi_gpu_rejection_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert all(res[k + "_equal"] for k in ("samples_left", "X", "Y", "mean", "var")), res
    assert res["survivors"] > 0 and res["rejected"] > 0, "want both survivors and rejected samples"
