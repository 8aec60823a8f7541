set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 600 gpurun_out/r2_bench_c.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_c.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_ref_c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_c.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_c.csv > gpurun_out/r2_launches_c.txt; head -5 gpurun_out/r2_launches_c.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_v04.csv python tools/profile_rollout.py 125000 50 1 > gpurun_out/traffic_v04.log 2>&1
python tools/traffic_summary.py gpurun_out/traffic_v04.csv gpurun_out/r2_traffic_v0.4.json "gpmpc_b200 0.4 (sm_100a)" | tail -5
ncu --set full --clock-control none --import-source on -k regex:^k_step$ -s 10 -c 1 -o gpurun_out/r2c_step_c30 -f python tools/profile_rollout.py 20000 50 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_step$ -s 44 -c 1 -o gpurun_out/r2c_step_c132 -f python tools/profile_rollout.py 20000 50 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_sqp_c.csv python tools/profile_sqp.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_sqp_c.csv > gpurun_out/r2_sqp_launches_c.txt; cat gpurun_out/r2_sqp_launches_c.txt | head -12
python tools/k0_probe.py 1000 2000 3000 5000 10000 > gpurun_out/r2_k0_blocked.txt 2>&1; GPMPC_K0_BLOCKED_MIN_M=100000000 python tools/k0_probe.py 1000 2000 3000 >> gpurun_out/r2_k0_blocked.txt 2>&1; cat gpurun_out/r2_k0_blocked.txt
python tools/stagger_probe.py 125000 7 > gpurun_out/r2_stagger_probe.txt 2>&1; python tools/pipeline_probe.py 200000 0 110 74 >> gpurun_out/r2_stagger_probe.txt 2>&1; tail -8 gpurun_out/r2_stagger_probe.txt
