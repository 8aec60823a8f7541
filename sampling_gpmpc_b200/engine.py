"""ctypes binding of libgpmpc_b200.so (include/gpmpc_b200.h) -- the only way this package computes.

There is deliberately no fallback: importing works anywhere (so CPU-only tooling can inspect the ABI),
but constructing a ``GPEngine`` without the built library or without a CUDA device raises.
torch is used for device memory and streams only; every number comes out of the hand-written kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPMPC_B200_LIB", os.path.join(_HERE, "libgpmpc_b200.so"))  # override: kernel experiments

MAX_D, MAX_T, MAX_NX = 6, 7, 8

ST_TRAIN_JITTER, ST_TRAIN_NOT_PD, ST_SAMPLE_NOT_PD, ST_APPEND_NOT_PD, ST_NAN_INPUT, ST_SAMPLE_EIG = 1, 2, 4, 8, 16, 32
OPT_NO_EIG_FALLBACK = 1


class GpmpcDims(C.Structure):
    _fields_ = [("ns", C.c_int32), ("g_ny", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
                ("n_real", C.c_int32), ("cap_points", C.c_int32)]


class GpmpcSampleOpts(C.Structure):
    _fields_ = [("beta", C.c_double), ("variance_is_zero", C.c_double),
                ("unclamped_sqrt_1x1", C.c_int32), ("flags", C.c_int32)]


class GpmpcEnv(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("g_ny", C.c_int32), ("d", C.c_int32),
                ("g_idx_inputs", C.c_int32 * MAX_D), ("pad_g", C.c_int32 * MAX_NX), ("n_pad", C.c_int32),
                ("transform", C.c_int32), ("B_d", C.c_double * (MAX_NX * MAX_NX)),
                ("F_known", C.c_double * (MAX_NX * 2 * MAX_NX)), ("use_feedback", C.c_int32),
                ("reserved", C.c_int32), ("K_fb", C.c_double * (MAX_NX * MAX_NX)), ("x_equi", C.c_double * MAX_NX)]


# every symbol include/gpmpc_b200.h declares: name -> (restype, argtypes)
_P, _D, _I = C.c_void_p, C.c_void_p, C.c_int32
ABI = {
    "gpmpc_create": (C.c_int, [C.POINTER(GpmpcDims), C.POINTER(_P)]),
    "gpmpc_destroy": (C.c_int, [_P]),
    "gpmpc_last_error": (C.c_char_p, [_P]),
    "gpmpc_set_hypers": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double]),
    "gpmpc_set_real_data": (C.c_int, [_P, _D, _D, _P]),
    "gpmpc_reset_hallucinated": (C.c_int, [_P]),
    "gpmpc_reserve": (C.c_int, [_P, _I, _P]),
    "gpmpc_set_condition_on_hallucinated": (C.c_int, [_P, _I]),
    "gpmpc_posterior": (C.c_int, [_P, _D, _I, _D, _D, _D, C.POINTER(GpmpcSampleOpts), _D, _D, _P]),
    "gpmpc_sample": (C.c_int, [_P, _D, C.POINTER(GpmpcSampleOpts), _D, _D, _P]),
    "gpmpc_append": (C.c_int, [_P, _D, _D, C.POINTER(C.c_uint8), _I, _P]),
    "gpmpc_append_masked": (C.c_int, [_P, _D, _D, C.POINTER(C.c_uint8), _I, _P]),
    "gpmpc_step": (C.c_int, [_P, _D, _D, C.POINTER(GpmpcSampleOpts), _D, _D, _D, _D, _P]),
    "gpmpc_assemble": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _D, _I, _D, _P]),
    "gpmpc_rollout": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _D, _D, C.POINTER(GpmpcSampleOpts), _I, _D, _P]),
    "gpmpc_rollout_gated": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _D, _D, C.POINTER(GpmpcSampleOpts), _I, _D,
                                      C.POINTER(C.c_void_p), _P]),
    "gpmpc_min_dist_overwrite": (C.c_int, [_P, _D, _I, _D, _D, C.c_double, C.c_double, _D, _P]),
    "gpmpc_filter_new_points": (C.c_int, [_P, _D, _I, C.c_double, _I, _D, _D, _P]),
    "gpmpc_pack_plin": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _D, _D, _I, _I, _I, _D, _P]),
    "gpmpc_traj_stats": (C.c_int, [_P, _D, _I, _I, _I, _D, _D, _D, _D, _P]),
    "gpmpc_stage_hulls": (C.c_int, [_P, _D, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "gpmpc_hull2d": (C.c_int, [C.POINTER(C.c_double), _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gpmpc_num_hallucinated": (C.c_int32, [_P]),
    "gpmpc_num_factor_rows": (C.c_int32, [_P]),
    "gpmpc_num_real_observed": (C.c_int32, [_P]),
    "gpmpc_export_hallucinated": (C.c_int, [_P, _D, _D, _P]),
    "gpmpc_status": (C.c_int, [_P, C.POINTER(C.c_uint32), _I, _P]),
    "gpmpc_state_bytes": (C.c_int64, [_P]),
    "gpmpc_last_launch_work": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gpmpc_launch_count": (C.c_int64, [_P]),
    "gpmpc_set_timing": (C.c_int, [_P, _I]),
    "gpmpc_set_block_kernels": (C.c_int, [_P, _I]),
    "gpmpc_rollout_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "gpmpc_version": (C.c_char_p, []),
    "gpmpc_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "gpmpc_get_option": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "gpmpc_set_grouping": (C.c_int, [_P, _I, C.c_double]),
    "gpmpc_truncate_hallucinated": (C.c_int, [_P, _I]),
    "gpmpc_linearise": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _I, _I, _D, C.POINTER(GpmpcSampleOpts), _I, _D, _D, _D, _D, _D, _D, _P]),
    "gpmpc_fs_advance": (C.c_int, [_P, C.POINTER(GpmpcEnv), _D, _D, _D, C.c_double, _D, _D, _D, _D, _P]),
    "gpmpc_export_point_states": (C.c_int, [_P, _D, _P]),
    "gpmpc_base_samples": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p]),
}

_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the C-ABI library and type every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  sampling_gpmpc_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class GPEngineError(RuntimeError):
    pass


class NotPSDError(GPEngineError):
    """Counterpart of linear_operator's NotPSDError (posterior covariance not PD after the jitter ladder)."""


class NanError(GPEngineError):
    """Counterpart of linear_operator's NanError (a NaN reached a Cholesky factorisation)."""


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.dtype in (torch.float64, torch.int32, torch.uint8) and t.is_contiguous(), \
        "gpmpc_b200 takes contiguous CUDA float64/int32 tensors"
    return C.c_void_p(t.data_ptr())


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_env_struct(spec, feedback_K=None, x_equi=None) -> GpmpcEnv:
    e = GpmpcEnv()
    e.nx, e.nu, e.g_ny, e.d = spec.nx, spec.nu, spec.g_ny, spec.d
    for i, v in enumerate(spec.g_idx_inputs):
        e.g_idx_inputs[i] = v
    for i, v in enumerate(spec.pad_g):
        e.pad_g[i] = v
    e.n_pad = len(spec.pad_g)
    e.transform = spec.transform
    for i, v in enumerate(np.asarray(spec.B_d, dtype=np.float64).reshape(-1)):
        e.B_d[i] = v
    for i, v in enumerate(np.asarray(spec.F_known, dtype=np.float64).reshape(-1)):
        e.F_known[i] = v
    if feedback_K is not None:
        e.use_feedback = 1
        for i, v in enumerate(np.asarray(feedback_K, dtype=np.float64).reshape(-1)):
            e.K_fb[i] = v
        for i, v in enumerate(np.asarray(x_equi, dtype=np.float64).reshape(-1)):
            e.x_equi[i] = v
    return e


def hull2d(xy: np.ndarray) -> np.ndarray:
    """Exact 2-D hull of a few host points (n,2) -> positions of the vertices, counter-clockwise (host helper of
    the library; merges per-rank hull vertices of sharded samples)."""
    lib = load_library()
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    pos = np.empty((max(1, xy.shape[0]),), dtype=np.int32)
    n = C.c_int32(0)
    rc = lib.gpmpc_hull2d(xy.ctypes.data_as(C.POINTER(C.c_double)), xy.shape[0], pos.ctypes.data_as(C.POINTER(C.c_int32)),
                          C.byref(n))
    if rc != 0:
        raise GPEngineError(f"gpmpc_hull2d failed ({rc})")
    return pos[: n.value].copy()


class GPEngine:
    """One handle = one Agent's GP on one GPU: shared real-data factor + per-sample bordered factors."""

    def __init__(self, ns: int, g_ny: int, d: int, T: int, n_real: int, cap_points: int = 0,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("sampling_gpmpc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.ns, self.g_ny, self.d, self.T, self.n_real = ns, g_ny, d, T, n_real
        self.B = ns * g_ny
        dims = GpmpcDims(ns, g_ny, d, T, n_real, cap_points)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.gpmpc_create(C.byref(dims), C.byref(h))
        if rc != 0:
            raise GPEngineError(f"gpmpc_create failed ({rc}): {self.lib.gpmpc_last_error(None).decode()}")
        self.h = h

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.gpmpc_destroy(h)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GPEngineError(f"{what} failed ({rc}): {self.lib.gpmpc_last_error(self.h).decode()}")

    # ---- model definition ----------------------------------------------------------------------
    def set_hypers(self, lengthscale, outputscale, noise, jitter: float):
        ls = np.ascontiguousarray(np.asarray(lengthscale, dtype=np.float64).reshape(self.g_ny, self.d))
        os_ = np.ascontiguousarray(np.asarray(outputscale, dtype=np.float64).reshape(self.g_ny))
        nz = np.ascontiguousarray(np.asarray(noise, dtype=np.float64).reshape(self.g_ny, self.T))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self.lib.gpmpc_set_hypers(self.h, dp(ls), dp(os_), dp(nz), float(jitter)), "gpmpc_set_hypers")
        self.outputscale = os_

    def set_real_data(self, X: torch.Tensor, Y: torch.Tensor):
        X = X.to(self.device, torch.float64).contiguous()
        Y = Y.to(self.device, torch.float64).contiguous()
        assert X.shape == (self.n_real, self.d) and Y.shape == (self.g_ny, self.n_real, self.T)
        self._check(self.lib.gpmpc_set_real_data(self.h, _ptr(X), _ptr(Y), _stream(self.device)), "gpmpc_set_real_data")

    def reset_hallucinated(self):
        self._check(self.lib.gpmpc_reset_hallucinated(self.h), "gpmpc_reset_hallucinated")

    def reserve(self, cap_points: int):
        self._check(self.lib.gpmpc_reserve(self.h, int(cap_points), _stream(self.device)), "gpmpc_reserve")

    def set_condition_on_hallucinated(self, on: bool):
        self._check(self.lib.gpmpc_set_condition_on_hallucinated(self.h, int(on)), "gpmpc_set_condition")

    # ---- hot path ------------------------------------------------------------------------------
    @staticmethod
    def opts(beta=-1.0, variance_is_zero=-1.0, unclamped_sqrt_1x1=False, eig_fallback=True) -> GpmpcSampleOpts:
        """eig_fallback (default, GPyTorch's behaviour): a joint Cholesky that fails after the jitter ladder makes the
        whole batch draw through the eigen root; off: failing elements return NaN and raise_on_status raises."""
        return GpmpcSampleOpts(float(beta), float(variance_is_zero), int(unclamped_sqrt_1x1),
                               0 if eig_fallback else OPT_NO_EIG_FALLBACK)

    def _x(self, x: torch.Tensor, H: int) -> torch.Tensor:
        x = x.to(self.device, torch.float64).reshape(self.B, H, self.d)
        return x if x.is_contiguous() else x.contiguous()

    def posterior(self, x: torch.Tensor, eps: Optional[torch.Tensor] = None, opts: Optional[GpmpcSampleOpts] = None):
        """x (ns,g_ny,H,d) -> mean, var (ns,g_ny,H,T) [, y (ns,g_ny,H,T), jitter_level (ns,g_ny)]."""
        H = x.shape[-2]
        xx = self._x(x, H)
        shp = (self.ns, self.g_ny, H, self.T)
        mean = torch.empty(shp, dtype=torch.float64, device=self.device)
        var = torch.empty(shp, dtype=torch.float64, device=self.device)
        y = jl = None
        if eps is not None:
            eps = eps.to(self.device, torch.float64).reshape(self.B, H * self.T).contiguous()
            y = torch.empty(shp, dtype=torch.float64, device=self.device)
            jl = torch.empty((self.ns, self.g_ny), dtype=torch.int32, device=self.device)
            opts = opts or self.opts()
        rc = self.lib.gpmpc_posterior(self.h, _ptr(xx), H, _ptr(mean), _ptr(var), _ptr(eps),
                                      C.byref(opts) if opts is not None else None, _ptr(y), _ptr(jl), _stream(self.device))
        self._check(rc, "gpmpc_posterior")
        return (mean, var) if eps is None else (mean, var, y, jl)

    def sample(self, eps: torch.Tensor, H: int, opts: Optional[GpmpcSampleOpts] = None):
        eps = eps.to(self.device, torch.float64).reshape(self.B, H * self.T).contiguous()
        y = torch.empty((self.ns, self.g_ny, H, self.T), dtype=torch.float64, device=self.device)
        jl = torch.empty((self.ns, self.g_ny), dtype=torch.int32, device=self.device)
        opts = opts or self.opts()
        self._check(self.lib.gpmpc_sample(self.h, _ptr(eps), C.byref(opts), _ptr(y), _ptr(jl), _stream(self.device)), "gpmpc_sample")
        return y, jl

    def append(self, x: torch.Tensor, y: torch.Tensor, point_active: Optional[np.ndarray] = None):
        H = x.shape[-2]
        xx = self._x(x, H)
        yy = y.to(self.device, torch.float64).reshape(self.B, H * self.T).contiguous()
        act = None
        if point_active is not None:
            pa = np.ascontiguousarray(np.asarray(point_active, dtype=np.uint8).reshape(H))
            act = pa.ctypes.data_as(C.POINTER(C.c_uint8))
        self._check(self.lib.gpmpc_append(self.h, _ptr(xx), _ptr(yy), act, H, _stream(self.device)), "gpmpc_append")

    def append_masked(self, x: torch.Tensor, y: torch.Tensor, scalar_active: np.ndarray):
        """append with one flag per new scalar: scalar_active (H, T) bool / uint8 (a point may enter with some tasks only)."""
        H = x.shape[-2]
        xx = self._x(x, H)
        yy = y.to(self.device, torch.float64).reshape(self.B, H * self.T).contiguous()
        sa = np.ascontiguousarray(np.asarray(scalar_active, dtype=np.uint8).reshape(H * self.T))
        self._check(self.lib.gpmpc_append_masked(self.h, _ptr(xx), _ptr(yy), sa.ctypes.data_as(C.POINTER(C.c_uint8)), H,
                                                 _stream(self.device)), "gpmpc_append_masked")

    def step(self, x: torch.Tensor, eps: Optional[torch.Tensor], opts: Optional[GpmpcSampleOpts] = None,
             want_moments: bool = True):
        """Fused H=1 step: x (ns,g_ny,1,d) [, eps (ns,g_ny,1,T)] -> mean, var [, y, jitter_level]."""
        xx = self._x(x, 1)
        shp = (self.ns, self.g_ny, 1, self.T)
        mean = torch.empty(shp, dtype=torch.float64, device=self.device) if want_moments else None
        var = torch.empty(shp, dtype=torch.float64, device=self.device) if want_moments else None
        y = jl = None
        if eps is not None:
            eps = eps.to(self.device, torch.float64).reshape(self.B, self.T).contiguous()
            y = torch.empty(shp, dtype=torch.float64, device=self.device)
            jl = torch.empty((self.ns, self.g_ny), dtype=torch.int32, device=self.device)
            opts = opts or self.opts()
        rc = self.lib.gpmpc_step(self.h, _ptr(xx), _ptr(eps), C.byref(opts) if opts is not None else None,
                                 _ptr(mean), _ptr(var), _ptr(y), _ptr(jl), _stream(self.device))
        self._check(rc, "gpmpc_step")
        return (mean, var) if eps is None else (mean, var, y, jl)

    def assemble(self, env: GpmpcEnv, xu: torch.Tensor, y_gp: torch.Tensor) -> torch.Tensor:
        """xu (ns,nx,H,nx+nu), y_gp (ns,g_ny,H,T) -> (ns,nx,H,1+nx+nu) = [f, df/dx, df/du] (dyn_fg_jacobians)."""
        ns, nx, H, nz = xu.shape
        xu = xu.to(self.device, torch.float64).contiguous()
        y_gp = y_gp.to(self.device, torch.float64).contiguous()
        out = torch.empty((ns, nx, H, 1 + nz), dtype=torch.float64, device=self.device)
        self._check(self.lib.gpmpc_assemble(self.h, C.byref(env), _ptr(xu), _ptr(y_gp), H, _ptr(out), _stream(self.device)),
                    "gpmpc_assemble")
        return out

    def rollout(self, env: GpmpcEnv, x0: torch.Tensor, u_ff: torch.Tensor, eps: torch.Tensor,
                opts: GpmpcSampleOpts, traj: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x0 (ns,nx), u_ff (n_steps,nu), eps (n_steps,ns,g_ny,1,T) -> traj (ns,nx,n_steps+1)."""
        n_steps = u_ff.shape[0]
        x0 = x0.to(self.device, torch.float64).contiguous()
        u_ff = u_ff.to(self.device, torch.float64).contiguous()
        eps = eps.to(self.device, torch.float64).reshape(n_steps, self.B * self.T).contiguous()
        if traj is None:
            traj = torch.empty((self.ns, env.nx, n_steps + 1), dtype=torch.float64, device=self.device)
        rc = self.lib.gpmpc_rollout(self.h, C.byref(env), _ptr(x0), _ptr(u_ff), _ptr(eps), C.byref(opts), n_steps,
                                    _ptr(traj), _stream(self.device))
        self._check(rc, "gpmpc_rollout")
        return traj

    def rollout_from_host(self, env: GpmpcEnv, x0: torch.Tensor, u_ff: torch.Tensor, eps_host: torch.Tensor,
                          opts: GpmpcSampleOpts, eps_dev: torch.Tensor, copy_stream: "torch.cuda.Stream",
                          traj: Optional[torch.Tensor] = None, chunk_steps: int = 5) -> torch.Tensor:
        """The rollout with the base samples in PINNED HOST memory: eps_host (n_steps, B*T) is copied to eps_dev
        (same shape, device) in chunks of `chunk_steps` horizon steps on `copy_stream`; step t waits only for the event
        of its own chunk (gpmpc_rollout_gated), so all but the first chunk's upload overlaps the horizon."""
        n_steps = u_ff.shape[0]
        assert eps_host.is_pinned() and eps_host.shape == eps_dev.shape == (n_steps, self.B * self.T)
        x0 = x0.to(self.device, torch.float64).contiguous()
        u_ff = u_ff.to(self.device, torch.float64, non_blocking=True).contiguous()
        if traj is None:
            traj = torch.empty((self.ns, env.nx, n_steps + 1), dtype=torch.float64, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        copy_stream.wait_stream(cur)  # eps_dev may still be read by work queued earlier on the compute stream
        ready = (C.c_void_p * n_steps)()
        events = []
        with torch.cuda.stream(copy_stream):
            for t0 in range(0, n_steps, chunk_steps):
                t1 = min(n_steps, t0 + chunk_steps)
                eps_dev[t0:t1].copy_(eps_host[t0:t1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
                ready[t0] = ev.cuda_event
        rc = self.lib.gpmpc_rollout_gated(self.h, C.byref(env), _ptr(x0), _ptr(u_ff), _ptr(eps_dev), C.byref(opts),
                                          n_steps, _ptr(traj), ready, _stream(self.device))
        self._check(rc, "gpmpc_rollout_gated")
        self._keep_events = events  # alive until the next call (the waits are already queued)
        return traj

    # ---- data-set rules, p_lin, trajectory consumers ---------------------------------------------
    def min_dist_overwrite(self, x: torch.Tensor, mean, var, y: torch.Tensor, min_dist: float, beta: float):
        """In place on y (ns,g_ny,H,T): nearest fully observed training targets where closer than min_dist, then
        truncation to mean +- beta sqrt(var) (src/agent.py:666-708)."""
        H = x.shape[-2]
        assert y.is_contiguous() and y.shape == (self.ns, self.g_ny, H, self.T)
        rc = self.lib.gpmpc_min_dist_overwrite(self.h, _ptr(self._x(x, H)), H, _ptr(mean), _ptr(var), float(min_dist),
                                               float(beta), _ptr(y), _stream(self.device))
        self._check(rc, "gpmpc_min_dist_overwrite")
        return y

    def filter_new_points(self, x: torch.Tensor, y: torch.Tensor, min_dist: float, use_hallucinated: bool = True):
        """NaNs (in place) the labels y (ns,g_ny,H,T) of new points within min_dist of the element's data set;
        returns counts (g_ny,H) int32 on the device: samples of this shard filtered per (output, point)."""
        H = x.shape[-2]
        assert y.is_contiguous() and y.shape == (self.ns, self.g_ny, H, self.T)
        counts = torch.empty((self.g_ny, H), dtype=torch.int32, device=self.device)
        rc = self.lib.gpmpc_filter_new_points(self.h, _ptr(self._x(x, H)), H, float(min_dist), int(use_hallucinated),
                                              _ptr(y), _ptr(counts), _stream(self.device))
        self._check(rc, "gpmpc_filter_new_points")
        return counts

    def pack_plin(self, env: GpmpcEnv, lin: torch.Tensor, x_h: torch.Tensor, tail: Optional[torch.Tensor],
                  use_feedback_K: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """lin (ns,nx,H,1+nx+nu), x_h (H,ns*nx), tail (H,n_tail) -> p_lin of every stage (H,P) (solver.py:98-131)."""
        ns, nx, H, w = lin.shape
        nu = w - 1 - nx
        n_tail = 0 if tail is None else tail.shape[1]
        P = ns * (nx * nx + nx * nu + 2 * nx) + n_tail
        x_h = x_h.to(self.device, torch.float64).contiguous()
        tail = None if tail is None else tail.to(self.device, torch.float64).contiguous()
        assert x_h.shape == (H, ns * nx) and lin.is_contiguous()
        if out is None:
            out = torch.empty((H, P), dtype=torch.float64, device=self.device)
        rc = self.lib.gpmpc_pack_plin(self.h, C.byref(env), _ptr(lin), _ptr(x_h), _ptr(tail), n_tail, H,
                                      int(use_feedback_K), _ptr(out), _stream(self.device))
        self._check(rc, "gpmpc_pack_plin")
        return out

    def traj_stats(self, traj: torch.Tensor, ref: Optional[torch.Tensor] = None):
        """traj (ns,nx,H1) -> box_min, box_max (nx,H1) [, max_dev (nx,H1) = max_n |traj - ref|]."""
        ns, nx, H1 = traj.shape
        traj = traj.contiguous()
        mk = lambda: torch.empty((nx, H1), dtype=torch.float64, device=self.device)
        lo, hi = mk(), mk()
        dev = mk() if ref is not None else None
        ref = None if ref is None else ref.to(self.device, torch.float64).contiguous()
        rc = self.lib.gpmpc_traj_stats(self.h, _ptr(traj), ns, nx, H1, _ptr(ref), _ptr(lo), _ptr(hi), _ptr(dev), _stream(self.device))
        self._check(rc, "gpmpc_traj_stats")
        return (lo, hi) if ref is None else (lo, hi, dev)

    def stage_hulls(self, traj: torch.Tensor, i0: int = 0, i1: int = 1, max_vertices: int = 512):
        """Per-stage convex hull of (traj[:, i0, t], traj[:, i1, t]): list over t of int32 sample-index arrays,
        counter-clockwise (generate_convex_hull.py:88-100)."""
        ns, nx, H1 = traj.shape
        traj = traj.contiguous()
        idx = np.empty((H1, max_vertices), dtype=np.int32)
        n = np.empty((H1,), dtype=np.int32)
        rc = self.lib.gpmpc_stage_hulls(self.h, _ptr(traj), ns, nx, H1, i0, i1, max_vertices,
                                        idx.ctypes.data_as(C.POINTER(C.c_int32)), n.ctypes.data_as(C.POINTER(C.c_int32)),
                                        _stream(self.device))
        self._check(rc, "gpmpc_stage_hulls")
        return [idx[t, : n[t]].copy() for t in range(H1)]

    # ---- introspection -------------------------------------------------------------------------
    @property
    def num_hallucinated(self) -> int:
        return self.lib.gpmpc_num_hallucinated(self.h)

    @property
    def num_factor_rows(self) -> int:
        return self.lib.gpmpc_num_factor_rows(self.h)

    @property
    def num_real_observed(self) -> int:
        return self.lib.gpmpc_num_real_observed(self.h)

    def export_hallucinated(self):
        n = self.num_hallucinated
        X = torch.empty((self.ns, self.g_ny, n, self.d), dtype=torch.float64, device=self.device)
        Y = torch.empty((self.ns, self.g_ny, n, self.T), dtype=torch.float64, device=self.device)
        if n:
            self._check(self.lib.gpmpc_export_hallucinated(self.h, _ptr(X), _ptr(Y), _stream(self.device)), "gpmpc_export")
        return X, Y

    def export_point_states(self) -> torch.Tensor:
        """(ns, g_ny, num_hallucinated) uint8: 0 in the factor, 1 recorded but masked, 2 dropped (grouped rollouts)."""
        out = torch.zeros((self.ns, self.g_ny, self.num_hallucinated), dtype=torch.uint8, device=self.device)
        if self.num_hallucinated:
            self._check(self.lib.gpmpc_export_point_states(self.h, _ptr(out), _stream(self.device)), "gpmpc_export_point_states")
        return out

    def status(self, clear: bool = False) -> int:
        s = C.c_uint32(0)
        self._check(self.lib.gpmpc_status(self.h, C.byref(s), int(clear), _stream(self.device)), "gpmpc_status")
        return int(s.value)

    def engine_status_ok(self) -> bool:
        """True when no Cholesky failed (a jitter-ladder escalation of the real-data block is not a failure, and neither
        is a draw that GPyTorch-style fell back to the eigen root)."""
        s = self.status()
        if s & ST_SAMPLE_EIG:
            s &= ~ST_SAMPLE_NOT_PD
        return s & (ST_SAMPLE_NOT_PD | ST_TRAIN_NOT_PD | ST_APPEND_NOT_PD | ST_NAN_INPUT) == 0

    def raise_on_status(self):
        """Reads and clears the device status word (one 4-byte copy + stream sync).  Raises NotPSDError where GPyTorch's
        psd_safe_cholesky would (real-data block / conditioning block not PD; a draw not PD with the eigen fallback off);
        warns where linear_operator warns and carries on (draw through the eigen root).  After an ST_APPEND_NOT_PD the
        failing elements' new factor rows were not written: the handle is unusable until reset_hallucinated()."""
        s = self.status(clear=True)
        if s & ST_NAN_INPUT:
            raise NanError(f"NaN in a matrix to be factorised (status {s:#x})")
        if s & ST_SAMPLE_EIG:
            import warnings
            warnings.warn("Cholesky of a posterior covariance failed after the jitter ladder; the batch was drawn "
                          "through the eigen root (GPyTorch: 'Using symeig method')", RuntimeWarning, stacklevel=2)
            s &= ~ST_SAMPLE_NOT_PD
        if s & (ST_SAMPLE_NOT_PD | ST_TRAIN_NOT_PD | ST_APPEND_NOT_PD):
            raise NotPSDError(f"matrix not positive definite after the jitter ladder (status {s:#x})")
        return s

    @property
    def state_bytes(self) -> int:
        return self.lib.gpmpc_state_bytes(self.h)

    def last_launch_work(self):
        b, f = C.c_double(0), C.c_double(0)
        self.lib.gpmpc_last_launch_work(self.h, C.byref(b), C.byref(f))
        return b.value, f.value

    def set_block_kernels(self, mma: bool):
        """SQP-mode model call on the tensor cores (default) or by the scalar substitution kernel (reference semantics)."""
        self._check(self.lib.gpmpc_set_block_kernels(self.h, int(mma)), "gpmpc_set_block_kernels")

    def linearise(self, env: GpmpcEnv, xu: torch.Tensor, eps: torch.Tensor, opts: GpmpcSampleOpts, reset_first: bool, bufs: dict,
                  copy_to_host: bool = True):
        """One SQP GP linearisation in one C call (gpmpc_linearise).  xu (ns,nx,H,nx+nu): pinned CPU tensor or CUDA tensor;
        eps CUDA (ns,g_ny,H,T).  `bufs` caches the output tensors between calls.  Returns (mean, var, y, jl, out, out_host);
        nothing has been waited for."""
        ns, nx, H, nz = xu.shape
        key = (ns, nx, H, nz)
        if bufs.get("key") != key:
            shp = (self.ns, self.g_ny, H, self.T)
            mk = lambda: torch.empty(shp, dtype=torch.float64, device=self.device)
            bufs.update(key=key, mean=mk(), var=mk(), y=mk(),
                        jl=torch.empty((self.ns, self.g_ny), dtype=torch.int32, device=self.device),
                        out=torch.empty((ns, nx, H, 1 + nz), dtype=torch.float64, device=self.device),
                        out_host=torch.empty((ns, nx, H, 1 + nz), dtype=torch.float64, pin_memory=True))
        on_host = not xu.is_cuda
        assert xu.dtype == torch.float64 and xu.is_contiguous() and (xu.is_cuda or xu.is_pinned())
        eps = eps.to(self.device, torch.float64).reshape(self.B, H * self.T)
        eps = eps if eps.is_contiguous() else eps.contiguous()
        b = bufs
        rc = self.lib.gpmpc_linearise(self.h, C.byref(env), C.c_void_p(xu.data_ptr()), int(on_host), H, _ptr(eps), C.byref(opts),
                                      int(reset_first), _ptr(b["mean"]), _ptr(b["var"]), _ptr(b["y"]), _ptr(b["jl"]),
                                      _ptr(b["out"]), C.c_void_p(b["out_host"].data_ptr()) if copy_to_host else None,
                                      _stream(self.device))
        self._check(rc, "gpmpc_linearise")
        return b["mean"], b["var"], b["y"], b["jl"], b["out"], b["out_host"]

    def truncate_hallucinated(self, n_points: int):
        """Forget the hallucinated points from index n_points on (prepare_dynamics_set's forward-sampling set)."""
        self._check(self.lib.gpmpc_truncate_hallucinated(self.h, int(n_points)), "gpmpc_truncate_hallucinated")

    def fs_advance(self, env: GpmpcEnv, xu: torch.Tensor, y: torch.Tensor, x_target: torch.Tensor, c_i: float,
                   u_next: Optional[torch.Tensor], samples_left: torch.Tensor):
        """One step of Agent.prepare_dynamics_set's rejection rollout (src/agent.py:381-415): -> x_next (ns,nx)
        [, xu_next (ns,nx,1,nx+nu)]; samples_left (ns,) int32 is updated in place."""
        ns, nx, _, nz = xu.shape
        xu = xu.to(self.device, torch.float64).contiguous()
        y = y.to(self.device, torch.float64).contiguous()
        x_target = x_target.to(self.device, torch.float64).contiguous()
        assert samples_left.dtype == torch.int32 and samples_left.is_cuda and samples_left.is_contiguous()
        x_next = torch.empty((ns, nx), dtype=torch.float64, device=self.device)
        xu_next = None
        if u_next is not None:
            u_next = u_next.to(self.device, torch.float64).contiguous()
            xu_next = torch.empty((ns, nx, 1, nz), dtype=torch.float64, device=self.device)
        rc = self.lib.gpmpc_fs_advance(self.h, C.byref(env), _ptr(xu), _ptr(y), _ptr(x_target), float(c_i), _ptr(u_next),
                                       _ptr(samples_left), _ptr(x_next), _ptr(xu_next), _stream(self.device))
        self._check(rc, "gpmpc_fs_advance")
        return x_next, xu_next

    def set_grouping(self, group_size: int, min_dist: float):
        """Consecutive blocks of `group_size` samples are one reference Agent each: the min-distance filter of
        update_hallucinated_Dyn_dataset (src/agent.py:164-202) and GPyTorch's any-over-batch NaN mask act per group inside
        gpmpc_step / gpmpc_rollout.  0 switches it off."""
        self._check(self.lib.gpmpc_set_grouping(self.h, int(group_size), float(min_dist)), "gpmpc_set_grouping")

    def set_option(self, name: str, value: int):
        """Tuning switches of the C ABI (gpmpc_set_option): rollout_fused, hz_groups, hz_stagger_ns.  Results do not depend
        on them."""
        self._check(self.lib.gpmpc_set_option(self.h, name.encode(), int(value)), "gpmpc_set_option")

    def get_option(self, name: str) -> int:
        """Reads a gpmpc_set_option switch back; also "last_rollout_fused" (did the last rollout take the one-launch kernel)."""
        v = C.c_int64(0)
        self._check(self.lib.gpmpc_get_option(self.h, name.encode(), C.byref(v)), "gpmpc_get_option")
        return int(v.value)

    def set_timing(self, on: bool):
        self._check(self.lib.gpmpc_set_timing(self.h, int(on)), "gpmpc_set_timing")

    def rollout_kernel_ms(self):
        ms, n = C.c_double(0), C.c_int32(0)
        self._check(self.lib.gpmpc_rollout_kernel_ms(self.h, C.byref(ms), C.byref(n)), "gpmpc_rollout_kernel_ms")
        return ms.value, n.value

    @property
    def launch_count(self) -> int:
        return self.lib.gpmpc_launch_count(self.h)
