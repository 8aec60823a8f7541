"""Base samples of the hot path: ``Agent.random_vector_within_bounds`` (src/agent.py:76-104).

The reference pre-draws every standard-normal vector the MPC run will use -- one candidate per Python iteration,
``torch.normal(0, 1, size=(1, g_ny, H, T))`` from the default generator, kept iff all |w| <= beta -- into
``epistimic_random_vector (n_mpc, n_sqp, ns, g_ny, H, T)``.  Results are only comparable with the reference if these
draws are the SAME numbers, so the generator stream is part of the interface:

* CPU generator (``common.use_cuda: False``, e.g. params_car_residual_fs.yaml, 400 000 candidates): the library restates
  torch's generator and both of ATen's normal paths (csrc/gpmpc_rng.cuh) and runs the whole rejection loop in one host
  call on the bytes of ``torch.get_rng_state()``; values AND the generator's final state are bit-identical to the loop.
* CUDA generator (``use_cuda: True``): one Philox launch per candidate is inherent to reproducing that stream; the loop
  stays, without the reference's quadratic ``torch.cat`` and with one device sync per candidate instead of two.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .engine import load_library

F64 = torch.float64


def truncated_normal_base_samples(n_mpc: int, n_sqp: int, ns: int, g_ny: int, H: int, T: int, beta: float,
                                  device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """-> (n_mpc, n_sqp, ns, g_ny, H, T) float64 on `device`, drawn from torch's DEFAULT generator of that device exactly as
    the reference's loop would (and leaving that generator in the same state)."""
    device = torch.device(device)
    shape = (n_mpc, n_sqp, ns, g_ny, H, T)
    if device.type == "cpu":
        lib = load_library()
        state = torch.get_rng_state()
        st = state.numpy().copy()
        out = np.empty(shape, dtype=np.float64)
        calls = lib.gpmpc_base_samples(st.ctypes.data_as(C.c_void_p), st.size, n_mpc * n_sqp * ns, g_ny * H * T, float(beta),
                                       out.ctypes.data_as(C.c_void_p))
        if calls < 0:
            raise RuntimeError(f"gpmpc_base_samples failed ({calls}): generator state of {st.size} bytes, beta {beta}")
        torch.set_rng_state(torch.from_numpy(st))
        return torch.from_numpy(out)
    out = torch.empty(shape, dtype=F64, device=device)
    for j in range(n_mpc):
        for i in range(n_sqp):
            k = 0
            while k < ns:
                w = torch.normal(0, 1, size=(1, g_ny, H, T), dtype=F64, device=device)
                if bool(w.abs().max() <= beta):  # == all(w >= -beta) and all(w <= beta); one sync
                    out[j, i, k] = w[0]
                    k += 1
    return out
