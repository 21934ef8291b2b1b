"""Turns gpurun_out/<tag>_launches.csv (launch list) and gpurun_out/<tag>_step.ncu-rep (--set full capture)
into the committed summaries under profiles/:
   profiles/<name>_launches.md   per-kernel totals of ONE predict step (ncu gpu__time_duration, cold, serialised)
   profiles/<name>_kernels.md    per-launch metrics of the captured kernels (time, DRAM bytes, throughputs, stalls)
Usage: python tools/ncu_summary.py <tag> <name> [rep-suffix ...]   (default rep: <tag>_step.ncu-rep)"""
import csv
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, name = sys.argv[1], sys.argv[2]
out_dir = ROOT / "profiles"

# ---- launch list ----
rows = list(csv.reader((ROOT / "gpurun_out" / f"{tag}_launches.csv").read_text().splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ix = {n: i for i, n in enumerate(h)}
agg = OrderedDict()
total = 0.0
for r in rows[hdr + 1:]:
    if len(r) != len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = r[ix["Kernel Name"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
lines = [f"# {name}: launch list of one predict step (batch 64, 640x640)", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/kernel_table.py --ncu`",
         "(cold-cache, serialised launches: the SHARE of a kernel is meaningful, not the absolute).", "",
         f"total {total:.1f} us over {sum(a[0] for a in agg.values())} launches", "", "| kernel | launches | us | share |", "|---|---|---|---|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k[:110]}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")
(out_dir / f"{name}_launches.md").write_text("\n".join(lines) + "\n")

# ---- full-set capture ----
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
     "smsp__cycles_active.avg", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_tma.sum", "sm__inst_executed_pipe_tmem.sum"]
suffixes = sys.argv[3:] or ["step"]
rr = []
for sfx in suffixes:
    rep = ROOT / "gpurun_out" / f"{tag}_{sfx}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    rr = part if not rr else rr + part[2:]
hh, units = rr[0], rr[1]
jx = {n: i for i, n in enumerate(hh)}
cols = [m for m in M if m in jx]
out = [f"# {name}: ncu --set full of the hot kernels of one predict step (batch 64, 640x640)", "",
       "`ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:... python tools/kernel_table.py --ncu`", "",
       "| # | kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | SM % | L1 % | L2 % | tensor pipe % (active / elapsed) | warps active % | warp inst |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for i, r in enumerate(rr[2:]):
    g = lambda m: r[jx[m]] if m in jx else ""
    f = lambda m: float(g(m).replace(",", "")) if g(m) not in ("", "n/a") else float("nan")
    kn = r[jx["Kernel Name"]].replace("void ", "").replace("unnamed>::", "")
    out.append(f"| {i} | `{kn[:70]}` | {g('launch__grid_size')} x {g('launch__block_size')} | {g('launch__registers_per_thread')} | "
               f"{f('gpu__time_duration.sum'):.1f} | {f('dram__bytes_read.sum'):.1f} | {f('dram__bytes_write.sum'):.1f} | "
               f"{f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{f('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {f('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} / {f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {f('smsp__inst_executed.sum'):.3g} |")
out += ["", f"units: time {units[jx['gpu__time_duration.sum']]}, DRAM {units[jx['dram__bytes_read.sum']]}", ""]
(out_dir / f"{name}_kernels.md").write_text("\n".join(out) + "\n")
print("\n".join(lines[:24]))
print("\n".join(out[:40]))
