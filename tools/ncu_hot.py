"""Developer tool: top stall sites of one kernel from an .ncu-rep (SASS level, with the neighbouring
instruction that produced the stall).  Usage: python tools/ncu_hot.py x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[col["# Samples"]]) for r in body)
    inst = sum(int(r[col["Instructions Executed"]]) for r in body)
    print(f"kernel: {rows[0][1]}\nsamples {tot}, warp instructions {inst}, SASS lines {len(body)}")
    reasons = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_branch_resolving",
               "stall_not_selected", "stall_selected", "stall_no_inst", "stall_barrier", "stall_dispatch", "stall_tex"]
    print("by reason:", ", ".join(f"{k[6:]} {100 * sum(int(r[col[k]]) for r in body) / tot:.1f}%" for k in reasons))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]]))[:top]
    for i in sorted(order):
        r = body[i]
        why = max(reasons, key=lambda k: int(r[col[k]]))
        print(f"{100 * int(r[col['# Samples']]) / tot:5.1f}%  #{i:4d}  exec {int(r[col['Instructions Executed']]):>9d}  {why[6:]:12s} {r[col['Source']].strip()}")


if __name__ == "__main__":
    main()
