# compute-sanitizer over the kernels added / rewritten in the second half of round 2 (small shapes: the tools slow kernels 10-100x)
set -x
export GPMPC_K0_BLOCKED_MIN_M=1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_k0_blocked.py::test_blocked_k0_matches_the_per_pivot_kernels[70-2-3-True]" tests/test_gpu_dynamics_rejection.py -x -q 2>&1 | tail -4
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/profile_sqp.py 2>&1 | tail -3
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/profile_rollout.py 300 12 2>&1 | tail -3
compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_sqp.py 2>&1 | grep -v "^=========     " | tail -12
compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_gpu_k0_blocked.py::test_blocked_k0_matches_the_per_pivot_kernels[70-2-3-True]" -x -q 2>&1 | grep -v "^=========     " | tail -8
compute-sanitizer --tool synccheck python tools/profile_sqp.py 2>&1 | tail -3
