"""Hot source lines of an ncu report (development aid).

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/source_hotspots.py src.csv [kernel-substring] [top=40]

Per kernel: total stall samples and warp instructions, then the source lines (file:line) ranked by samples with
their instruction counts and dominant stall reasons.
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

rows = list(csv.reader(open(path, newline="")))
fn = fpath = None
hdr = None
per = defaultdict(lambda: defaultdict(lambda: {"samples": 0, "inst": 0, "stalls": defaultdict(int), "src": ""}))
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "":          # SASS rows under a source line: already summed in the line's own row
        continue
    d = dict(zip(hdr, r))
    key = f"{fpath}:{r[0]}"
    e = per[fn][key]
    try:
        e["samples"] += int(d["# Samples"] or 0)
        e["inst"] += int(d["Instructions Executed"] or 0)
    except ValueError:
        continue
    e["src"] = r[1].strip()[:110]
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            e["stalls"][k[6:]] += int(v)

for f, lines in per.items():
    if want not in f:
        continue
    tot_s = sum(e["samples"] for e in lines.values())
    tot_i = sum(e["inst"] for e in lines.values())
    print(f"== {f}\n   samples {tot_s}   warp-instructions {tot_i}")
    for key, e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k}={v}" for k, v in st)
        print(f"   {e['samples']:7d} {100 * e['samples'] / max(tot_s, 1):5.1f}%  inst {e['inst']:9d} {100 * e['inst'] / max(tot_i, 1):5.1f}%  {key:24s} [{sts}]  {e['src']}")
