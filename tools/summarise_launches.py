"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launches / mean
duration / share of the profiled time (the form committed under profiles/r01_ncu_launches_*.csv).

    python tools/summarise_launches.py gpurun_out/launches.csv "<header comment>" > profiles/...csv
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("tdm::", "").replace("(int)", "").replace("(bool)", "")
    name = re.sub(r"\(.*\)$", "", name).replace(" ", "")
    return name.replace("conv3x3_tc_kernel", "conv3x3_tc")


def main() -> None:
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    acc: "OrderedDict[str, list[float]]" = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[mu], v)   # -> microseconds
        acc.setdefault(short(r[kn]), []).append(v)
    total = sum(sum(v) for v in acc.values())
    for c in sys.argv[2:]:
        print("# " + c)
    print("kernel,launches,avg_us,share_of_profiled_time")
    for k, v in acc.items():
        print(f"{k},{len(v)},{sum(v) / len(v):.2f},{sum(v) / total:.4f}")


if __name__ == "__main__":
    main()
