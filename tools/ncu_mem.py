"""Developer tool: every memory instruction of a kernel with its per-tile counts (executions, L1 tag requests, shared
wavefronts, theoretical L2 sectors).  Usage: python tools/ncu_mem.py x.ncu-rep tiles"""
import csv, io, subprocess, sys
rep, tiles = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[col[k]])
    except ValueError: return 0.0
tot = {}
print("  #    exec/t  thr   tagreq/t  shwave/t shideal/t  l2sec/t  space      source")
for i, r in enumerate(rows[2:]):
    if len(r) != len(hdr) or not r[col["Address Space"]].strip(): continue
    ex = num(r, "Instructions Executed") / tiles
    if ex < 0.01: continue
    sp = r[col["Address Space"]] + "/" + r[col["Access Operation"]]
    v = (ex, num(r, "L1 Tag Requests Global") / tiles, num(r, "L1 Wavefronts Shared") / tiles, num(r, "L2 Theoretical Sectors Global") / tiles + num(r, "L2 Theoretical Sectors Local") / tiles)
    t = tot.setdefault(sp, [0, 0, 0, 0])
    for k in range(4): t[k] += v[k]
    print(f"{i:4d} {ex:8.2f} {num(r, 'Avg. Threads Executed'):5.1f} {v[1]:9.2f} {v[2]:9.2f} {num(r, 'L1 Wavefronts Shared Ideal') / tiles:9.2f} {v[3]:8.2f}  {sp:16s} {r[col['Source']].strip()[:70]}")
print("totals per tile (exec, tag requests, shared wavefronts, L2 sectors):")
for k, t in sorted(tot.items()): print(f"  {k:20s} {t[0]:8.2f} {t[1]:8.2f} {t[2]:8.2f} {t[3]:8.2f}")
