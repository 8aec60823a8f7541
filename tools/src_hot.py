"""Warp-instructions and stall samples per CUDA source line of the first captured kernel (needs -lineinfo):
python tools/src_hot.py x.ncu-rep [elements] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines, fname, hdr, nfun = [], None, None, 0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        nfun += 1
        if nfun > 1 and lines and r[1] != func: break
        func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) == len(hdr):
        ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if r[ia].isdigit(): lines.append((int(r[ia]), int(r[isamp]) if r[isamp].isdigit() else 0, fname, int(r[0]), r[1].strip()[:105]))
tot = sum(l[0] for l in lines); ts = sum(l[1] for l in lines) or 1
print(f"total {tot} warp-instructions, per element {tot / nelem:.1f}")
for n, sm, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{n / nelem:8.1f} {100 * n / tot:5.1f}%i {100 * sm / ts:5.1f}%s {f[6:-4]:>8}:{ln:<4} {src}")
