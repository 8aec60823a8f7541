"""`gpytorch` as seen by an unmodified sampling-gpmpc checkout when <repo>/sampling_gpmpc_b200/shim is on PYTHONPATH:

    PYTHONPATH=<repo>:<repo>/sampling_gpmpc_b200/shim python sampling-gpmpc/main.py -i 1 -param params_pendulum1D_samples

Importing this package registers sampling_gpmpc_b200.gpytorch_shim (the libgpmpc_b200.so-backed subset of the GPyTorch
API the reference uses, SURVEY.md 8b level B1) under the names `gpytorch`, `gpytorch.models`, `gpytorch.kernels`, ...
"""
import sys as _sys

from sampling_gpmpc_b200 import gpytorch_shim as _shim

_shim.install()                      # replaces sys.modules["gpytorch"] (this module) by the shim's module tree
_mod = _sys.modules["gpytorch"]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
