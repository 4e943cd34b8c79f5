"""Summarise an ncu --csv metrics log of the sampling step into per-kernel rows and the DRAM-traffic JSON bench.py reads.

    python tools/traffic_step.py <ncu.csv> <batch> <fused 0|1> <launches per step> <out.csv> <out.json> ["comment" ...]
The log holds every matched launch of the command (several identical steps); the LAST step is summarised."""
import csv
import json
import re
import sys
from collections import OrderedDict

path, batch, fused, per, out_csv, out_json = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3])), int(sys.argv[4]), sys.argv[5], sys.argv[6]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
idc, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
launch: "OrderedDict[str, dict]" = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    d = launch.setdefault(r[idc], {"name": r[kn]})
    v = float(r[mv].replace(",", "")) if r[mv] not in ("", "n/a") else float("nan")
    unit = r[mu]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    d[r[mn]] = v * scale
L = list(launch.values())[-per:]


def short(n: str) -> str:
    n = re.sub(r"^void\s+", "", n).replace("tdm::", "").replace("(int)", "").replace("(bool)", "")
    return re.sub(r"\(.*\)$", "", n).replace(" ", "")


metrics = [m for m in L[0] if m != "name"]
with open(out_csv, "w") as f:
    for c in sys.argv[7:]:
        f.write("# " + c + "\n")
    f.write("# units: gpu__time_duration.sum = us; dram bytes = bytes; the rest = % of peak\n")
    f.write("kernel," + ",".join(metrics) + "\n")
    for d in L:
        f.write(short(d["name"]).replace(",", ";") + "," + ",".join(f"{d.get(m, float('nan')):.6g}" for m in metrics) + "\n")
order = (["rb1_fused", "avgpool", "rb2_conv1", "rb2_conv2", "rb3_conv1", "rb3_conv2", "rb4_fused_out_step"] if fused else
         ["rb1_conv1", "rb1_conv2", "avgpool", "rb2_conv1", "rb2_conv2", "rb3_conv1", "rb3_conv2", "rb4_conv1", "rb4_conv2_out_step"])
if len(order) != len(L):
    order = [f"k{i}" for i in range(len(L))]
perk = {n: d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for n, d in zip(order, L)}
json.dump({"batch": batch, "fused": fused, "source": out_csv + " (ncu --metrics, one launch each, last of the captured steps)",
           "per_kernel": perk, "dram_bytes_per_step": sum(perk.values()),
           "kernel_us_under_ncu": {n: d.get("gpu__time_duration.sum") for n, d in zip(order, L)}}, open(out_json, "w"), indent=1)
print(json.dumps({"dram_GB_per_step": sum(perk.values()) / 1e9, **{k: round(v / 1e9, 3) for k, v in perk.items()}}))
