"""CPU: the C-ABI library builds, loads, and exports every symbol include/gpmpc_b200.h declares."""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from sampling_gpmpc_b200 import engine
    return engine.load_library()


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "gpmpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpmpc_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from sampling_gpmpc_b200 import engine
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(engine.ABI) == declared, "engine.ABI and include/gpmpc_b200.h disagree"
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libgpmpc_b200.so"
    assert b"sm_100a" in lib.gpmpc_version()


def test_struct_layouts_match_header():
    """ctypes mirrors of the header's structs: sizes a C compiler gives for the same declarations."""
    import ctypes as C
    from sampling_gpmpc_b200 import engine
    assert C.sizeof(engine.GpmpcDims) == 6 * 4
    assert C.sizeof(engine.GpmpcSampleOpts) == 2 * 8 + 2 * 4
    ints = 4 + engine.MAX_D + engine.MAX_NX + 2
    ints += ints % 2  # doubles are 8-aligned
    expect = ints * 4 + 8 * (engine.MAX_NX ** 2 + 2 * engine.MAX_NX ** 2) + 8 + 8 * (engine.MAX_NX ** 2 + engine.MAX_NX)
    assert C.sizeof(engine.GpmpcEnv) == expect


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to construct (and never touches the oracle)."""
    import torch
    from sampling_gpmpc_b200.engine import GPEngine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GPEngine(4, 1, 2, 3, 10)
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import sampling_gpmpc_b200.agent, sampling_gpmpc_b200.rollout, "
            "sampling_gpmpc_b200.engine; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % REPO)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_create_fails_loudly_without_device(lib):
    import ctypes as C
    import torch
    from sampling_gpmpc_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dims = engine.GpmpcDims(4, 1, 2, 3, 10, 0)
    h = C.c_void_p()
    rc = lib.gpmpc_create(C.byref(dims), C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.gpmpc_last_error(None)
    bad = engine.GpmpcDims(4, 1, 9, 3, 10, 0)
    assert lib.gpmpc_create(C.byref(bad), C.byref(h)) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the engine (and with it every product path) refuses to run instead of falling back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sampling_gpmpc_b200.engine import GPEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GPEngine(2, 1, 2, 3, 5)
