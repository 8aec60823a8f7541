"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per captured launch, SASS opcode mix and the
instructions with most stall samples.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def run(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
print(f"# {rep}\n## raw metrics per captured launch")
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90])
    for w in WANT:
        if w in hdr:
            print(f"  {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
              if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[i] not in ("", "n/a")]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {h.split('stalled_')[1].split('_per_warp')[0]:30s} {v:.1f} %")

sass = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
hdr = None
data, nk = [], 0
for r in sass:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk == 2:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
if data:
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ia]) for r in data)
    ts = sum(int(r[isamp]) for r in data) or 1
    byop, samp = collections.Counter(), collections.Counter()
    for r in data:
        parts = r[isrc].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        byop[op] += int(r[ia])
        samp[op] += int(r[isamp])
    print(f"\n## SASS opcode mix of the first captured launch ({tot} warp-instructions, {len(data)} SASS lines)")
    for op, c in byop.most_common(22):
        print(f"  {op:10s} {100 * c / tot:5.1f}% of instructions   {100 * samp[op] / ts:5.1f}% of stall samples")
    print("\n## top 25 SASS lines by stall samples")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:25]:
        print(f"  {100 * int(r[isamp]) / ts:5.1f}%  x{int(r[ia])}  {r[isrc].strip()[:100]}")
