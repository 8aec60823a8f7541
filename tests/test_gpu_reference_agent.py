"""GPU: the UNMODIFIED reference Agent (src/agent.py + src/GP_model.py + env classes of sampling-gpmpc, from the git-ignored
copy baseline/_ref/sampling-gpmpc that __graft_entry__.build() stages) evaluated on the B200 through the product's
gpytorch shim, driven like src/solver.py:84-94 / simulate_forward_sampling_car.py:117-138 (tests/ref_agent_driver.py), and
compared with the fixtures the same reference code produced on the CPU stand-in.  Tolerance 1e-9 * max(|b|, s), identical
jitter-ladder decisions, bit-identical hallucinated inputs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(REPO, "tests", "ref_agent_driver.py")
# free-running replays are meaningful for the well-conditioned fixtures (tests/test_gpu_parity.py explains the others)
CASES = ["pendulum1D_sqp", "car_residual_truedyn", "car_residual_sqp_jit", "car_residual_fs"]


def _reference_present():
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from ref_agent_driver import find_reference
    return find_reference() is not None


@pytest.mark.skipif(not _reference_present(), reason="no reference checkout (run __graft_entry__.build() where /root/reference is mounted)")
@pytest.mark.parametrize("case,extra", [(c, []) for c in CASES] + [("pendulum1D_sqp", ["--cpu-tensors"]), ("car_residual_fs", ["--cpu-tensors"])])
def test_unmodified_reference_agent_on_the_gpu_matches_its_own_fixtures(case, extra):
    r = subprocess.run([sys.executable, DRIVER, case, *extra], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(line[-1])
    out_dir = os.path.join(REPO, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"reference_agent_{case}{'_cpu_tensors' if extra else ''}.json"), "w") as f:
        json.dump(out, f, indent=1)
    assert "unavailable" not in out, out
    assert out["model_class"].startswith("src.GP_model"), out["model_class"]  # the reference's own model class on the shim
    assert out["jitter_levels_equal"] and out["hallucinated_counts_equal"] and out["hallucinated_inputs_bit_equal"], out
    for q, v in out["worst_over_tolerance"].items():
        assert v <= 1.0, f"{case}: {q} off by {v:.3g} x tolerance ({out})"
    assert r.returncode == 0, r.stderr[-4000:]
