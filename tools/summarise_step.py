"""Per-kernel totals of the LAST of `nsteps` identical steps in an ncu launch list (training / text step lists
under profiles/).   python tools/summarise_step.py <launches.csv> <nsteps> "<header comment>" > profiles/...csv"""
import csv
import re
import sys
from collections import OrderedDict

path, nsteps = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
L = []
for r in rows[hdr + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[mu], v)
    n = re.sub(r"^void\s+", "", r[kn]).replace("tdm::", "").replace("(int)", "").replace("(bool)", "")
    n = re.sub(r"\(.*\)$", "", n).replace(" ", "").replace("conv3x3_tc_kernel", "conv3x3_tc")
    L.append((n[:90], v))
per = len(L) // nsteps
step = L[-per:]
acc: "OrderedDict[str, list[float]]" = OrderedDict()
for n, v in step:
    acc.setdefault(n, []).append(v)
tot = sum(v for _, v in step)
for c in sys.argv[3:]:
    print("# " + c)
print("kernel,launches,total_us,share")
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k},{len(v)},{sum(v):.1f},{sum(v) / tot:.3f}")
print(f"TOTAL,{len(step)},{tot:.1f},1.000")
