"""Developer tool: warp instructions per 64-particle tile attributed to CUDA source lines (needs -lineinfo and
--import-source on).  Usage: python tools/ncu_srclines.py x.ncu-rep tiles [min_per_tile]"""
import csv, io, subprocess, sys
rep, tiles = sys.argv[1], float(sys.argv[2]); mn = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
fname = "?"; tot = 0; out = []
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] == "Line No": continue
    try: n = int(r[7]); smp = int(r[6]); ln = int(r[0])
    except ValueError: continue
    tot += n
    out.append((fname, ln, n / tiles, smp, r[1].strip()[:110]))
ts = sum(o[3] for o in out)
print(f"total {tot / tiles:.1f} warp instructions per tile")
for f, ln, n, smp, src in out:
    if n >= mn: print(f"{f}:{ln:4d} {n:7.1f} {100 * smp / ts:5.1f}%  {src}")
