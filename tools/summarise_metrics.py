"""Summarise an `ncu --metrics ... --csv` capture (one row per launch and metric) into one line per launch:
    python tools/summarise_metrics.py raw.csv "<header comment>" ... > profiles/xxx.csv
Columns: kernel (template arguments kept), grid, duration_us, dram_read_MB, dram_write_MB, dram_pct, l2_bytes_MB, tc_pipe_pct,
tensor_math_pct, issue_pct."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
acc: "OrderedDict[str, dict]" = OrderedDict()
for r in csv.DictReader(lines):
    d = acc.setdefault(r["ID"], {"kernel": re.sub(r"\(.*\)$", "", r["Kernel Name"].replace("void ", "").replace("tdm::", "").replace("(int)", "")),
                                 "grid": r["Grid Size"].replace(", 1, 1)", "").replace("(", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        d["duration_us"] = v / {"ns": 1e3, "us": 1.0, "ms": 1e-3}.get(u, 1e3)
    elif n.startswith("dram__bytes_read"):
        d["dram_read_MB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    elif n.startswith("dram__bytes_write"):
        d["dram_write_MB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    elif n.startswith("lts__t_bytes"):
        d["l2_bytes_MB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    elif n.startswith("gpu__dram_throughput"):
        d["dram_pct"] = v
    elif n.startswith("sm__pipe_tc_cycles"):
        d["tc_pipe_pct"] = v
    elif n.startswith("sm__pipe_tensor_cycles"):
        d["tensor_math_pct"] = v
    elif n.startswith("smsp__issue_active"):
        d["issue_pct"] = v
for c in sys.argv[2:]:
    print("# " + c)
cols = ["kernel", "grid", "duration_us", "dram_read_MB", "dram_write_MB", "dram_pct", "l2_bytes_MB", "tc_pipe_pct", "tensor_math_pct", "issue_pct"]
print(",".join(cols))
for d in acc.values():
    print(",".join(f"{d.get(c, ''):.1f}" if isinstance(d.get(c), float) else str(d.get(c, "")) for c in cols))
