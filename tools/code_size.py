"""Static SASS size of a kernel by source line (development aid; needs only cuobjdump / nvdisasm).

    python tools/code_size.py <kernel-substring> [top=30] [cubin=unet_fwd]

The instruction cache is a real limit for the warp-specialised fused kernels: every role's loop body is
resident code, and ~50 KB of SASS thrashed it.  This prints instructions per source line, largest first.
"""
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

root = Path(__file__).resolve().parent.parent
lib = root / "tinydiffusionmodels_b200" / "libtdm_b200.so"
want = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cubin = sys.argv[3] if len(sys.argv) > 3 else "unet_fwd"

with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=d, capture_output=True)
    f = next(Path(d).glob(f"{cubin}*.cubin"))
    txt = subprocess.run(["nvdisasm", "-g", str(f)], capture_output=True, text=True).stdout

fn = None
cur = None
cnt = defaultdict(lambda: defaultdict(int))
for line in txt.splitlines():
    if line.startswith(".text."):
        fn = line[6:].rstrip(":")
        cur = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = f"{Path(m.group(1)).name}:{m.group(2)}"
        continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[fn][cur] += 1

src_cache = {}
def src(key):
    name, ln = key.split(":")
    for base in (root / "tinydiffusionmodels_b200" / "csrc",):
        p = base / name
        if p.exists():
            if p not in src_cache:
                src_cache[p] = p.read_text().splitlines()
            try:
                return src_cache[p][int(ln) - 1].strip()[:100]
            except IndexError:
                return ""
    return ""

for f, lines in cnt.items():
    if want not in f:
        continue
    tot = sum(lines.values())
    print(f"== {f}: {tot} instructions = {tot * 16 / 1024:.1f} KB")
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1])[:top]:
        print(f"   {v:5d}  {k or '?':26s} {src(k) if k else ''}")
