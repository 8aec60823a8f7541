"""Where the warp-instructions of the first captured kernel go: SASS lines grouped by execution count, with the
CUDA source lines they map to.   python tools/sass_hot.py x.ncu-rep [elements]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data)
print(f"total warp-instructions {tot}  per element {tot / nelem:.1f}")
groups = collections.defaultdict(list)
for idx, r in enumerate(data): groups[int(r[ia])].append(idx)
for k, idxs in sorted(groups.items(), key=lambda kv: -kv[0] * len(kv[1]))[:14]:
    ops = collections.Counter((data[i][isrc].split()[1] if data[i][isrc].split()[0].startswith("@") else data[i][isrc].split()[0]).split(".")[0] for i in idxs)
    print(f"exec/elem={k / nelem:9.2f} lines={len(idxs):4d} share={k * len(idxs) / tot:.3f} sass#{idxs[0]}..{idxs[-1]}  " + " ".join(f"{o}:{n}" for o, n in ops.most_common(8)))
