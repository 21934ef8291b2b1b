"""Top SASS lines of one kernel of an .ncu-rep by warp-stall samples (reads `ncu --page source --csv`).
Usage: python tools/ncu_hot.py <rep> <kernel-regex> [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
allrows = list(csv.reader(out.splitlines()))
# one table per matching launch ("Kernel Name" row, header row, body); some ncu builds ignore --launch-skip on import
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
k = int(skip)
rows = allrows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(allrows))]
print(rows[0][1][:120])
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
S, E = ix["# Samples"], ix["Instructions Executed"]
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
ts = sum(float(r[S] or 0) for r in body)
te = sum(float(r[E] or 0) for r in body)
print(f"lines {len(body)}  samples {ts:.0f}  warp-instructions {te:.0f}")
tot = {n: sum(float(r[ix[n]] or 0) for r in body) for n in stall_cols}
print("stall totals:", ", ".join(f"{n[6:]} {100 * v / ts:.1f}%" for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -float(body[i][S] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = max(stall_cols, key=lambda n: float(r[ix[n]] or 0))
    print(f"{i:5d} {100 * float(r[S]) / ts:5.1f}% smp {100 * float(r[E]) / te:5.1f}% inst  {st[6:]:<14} {r[ix['Source']].strip()[:90]}")
