"""Developer tool: dynamic instruction mix by opcode of one kernel from an .ncu-rep (SASS page).
Usage: python tools/ncu_mix.py x.ncu-rep tiles [--lines]   (tiles = 64-particle tiles per launch)"""
import csv, io, subprocess, sys, collections

def main():
    rep = sys.argv[1]; tiles = float(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    mix = collections.Counter(); thr = collections.Counter()
    tot = 0
    for r in body:
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.split(".")[0]
        n = int(r[col["Instructions Executed"]]); mix[op] += n; tot += n
        thr[op] += int(r[col["Thread Instructions Executed"]]) if "Thread Instructions Executed" in col else 0
    print(f"total {tot / tiles:.1f} warp instructions per 64-particle tile ({tot} in all)\n")
    print("| opcode | warp instr / tile | share | avg active threads |\n|---|---|---|---|")
    for op, n in mix.most_common(45):
        print(f"| {op} | {n / tiles:.1f} | {100 * n / tot:.1f} % | {thr[op] / n if n else 0:.1f} |")
    if "--lines" in sys.argv:
        print("\nper-SASS-line executed counts (per tile) with samples:")
        ts = sum(int(r[col["# Samples"]]) for r in body)
        for i, r in enumerate(body):
            n = int(r[col["Instructions Executed"]])
            if n: print(f"#{i:4d} {n / tiles:7.2f} {100 * int(r[col['# Samples']]) / ts:5.2f}%  {r[col['Source']].strip()}")

main()
