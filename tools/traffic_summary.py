"""ncu csv (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one conditioned car rollout,
tools/profile_rollout.py 125000 50) -> profiles/r<round>_traffic_<tag>.json, the `roofline.traffic` source of bench.py.
    python tools/traffic_summary.py gpurun_out/traffic.csv profiles/r1_traffic_v18.json"""
import csv, json, sys
src, dst = sys.argv[1], sys.argv[2]
version = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
k = {}
for r in rows:
    name = r["Kernel Name"].replace("void ", "").split("(")[0]
    d = k.setdefault(name, {"launches": 0, "time_ms_ncu_serialised": 0.0, "dram_read_GB": 0.0, "dram_write_GB": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["launches"] += 1
        d["time_ms_ncu_serialised"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[unit]
    else:
        gb = v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[unit]
        d["dram_read_GB" if "read" in r["Metric Name"] else "dram_write_GB"] += gb
step = [n for n in k if n.startswith("k_step")]
n_steps = max(k[n]["launches"] for n in step if n.startswith("k_step<"))  # (k_step_finish also has its eigen-redo launches)
traffic = sum(k[n]["dram_read_GB"] + k[n]["dram_write_GB"] for n in step) * 1e9 / n_steps
tot = sum(d["time_ms_ncu_serialised"] for d in k.values())
B, m = 375000, 45
alg = sum(8.0 * (c * m + c * (c + 1) / 2) + 8.0 * (2 + 9) + 8.0 * c / 3 * 2 + 8.0 * (3 * (m + c) + 6 + 3) for c in range(0, 150, 3)) * B / n_steps
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "python tools/profile_rollout.py 125000 50 1 (one 50-step conditioned car rollout, 125000 samples, B200)",
       "kernels": {n: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in d.items()} for n, d in k.items()},
       "step_share_of_rollout_time": sum(k[n]["time_ms_ncu_serialised"] for n in step) / tot,
       "traffic_bytes_per_step_launch": traffic, "algorithmic_bytes_per_step_launch": alg,
       "traffic_over_algorithmic": traffic / alg, "library_version": version}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
