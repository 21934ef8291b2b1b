"""SASS census of libuyd.so: per kernel, how many tensor-core / TMEM / TMA instructions the shipped cubin holds.
   UTCHMMA / UTCIMMA = tcgen05.mma kind::f16 / kind::i8,  LDTM = tcgen05.ld,  UTMALDG = TMA tensor load,
   HMMA / IMMA = Ampere-style mma.sync,  FFMA2 = packed fp32 FMA.
Usage: python tools/sass_census.py [out.md]   (runs cuobjdump -sass on unina-yolo-dla_b200/build/*.o; no GPU needed)"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OPS = ["UTCHMMA", "UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "HMMA", "IMMA", "FFMA2", "LDSM"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    clean = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        clean.append(re.sub(r"\(.*$", "", n)[:90])
    return clean


def main():
    rows = OrderedDict()
    for obj in sorted((ROOT / "unina-yolo-dla_b200" / "build").glob("*.o")):
        sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        cur = None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = (obj.stem, m.group(1))
                rows[cur] = dict.fromkeys(OPS, 0)
                rows[cur]["total"] = 0
                continue
            if cur is None:
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1).split(".")[0]
            rows[cur]["total"] += 1
            if op in rows[cur]:
                rows[cur][op] += 1
    names = demangle([k[1] for k in rows])
    lines = ["# SASS census of libuyd.so (cuobjdump -sass of every object that is linked into it)", "",
             "Counts of instructions in the shipped sm_100a cubins.  UTCHMMA / UTCIMMA = `tcgen05.mma` (bf16 / int8), LDTM = `tcgen05.ld`,",
             "UTMALDG = TMA tensor load, HMMA / IMMA = `mma.sync`, FFMA2 = packed fp32 FMA, LDSM = `ldmatrix`.", "",
             "| object | kernel | SASS instr | " + " | ".join(OPS) + " |", "|---|---|---|" + "---|" * len(OPS)]
    tot = dict.fromkeys(OPS, 0)
    for (obj, _), name in zip(rows, names):
        r = rows[(obj, _)]
        if not any(r[o] for o in OPS):
            continue
        for o in OPS:
            tot[o] += r[o]
        lines.append(f"| {obj} | `{name}` | {r['total']} | " + " | ".join(str(r[o]) if r[o] else "" for o in OPS) + " |")
    lines.append("| **all** | | | " + " | ".join(f"**{tot[o]}**" for o in OPS) + " |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text)
    print(text)


if __name__ == "__main__":
    main()
