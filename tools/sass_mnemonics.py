"""Count the SASS mnemonics that prove tcgen05 / TMEM / bulk-copy use, per kernel of the in-tree library.

    python tools/sass_mnemonics.py [> profiles/rNN_sass_mnemonics.txt]        (cuobjdump only; no GPU needed)

UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc/relinquish, LDTM / STTM =
tcgen05.ld / st, UBLKCP = cp.async.bulk (1-D, TMA engine), LDGSTS = cp.async (B200_PROFILING.md).
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "tinydiffusionmodels_b200" / "libtdm_b200.so"
WANTED = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "LDGSTS")


def kernel_mnemonics(lib: Path = LIB) -> "collections.OrderedDict[str, collections.Counter]":
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"\b(" + "|".join(WANTED) + r")\b")
    out: "collections.OrderedDict[str, collections.Counter]" = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = collections.Counter()
        elif cur and "/*" in line:
            mm = pat.search(line)
            if mm:
                out[cur][mm.group(1)] += 1
    names = list(out)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    pretty = collections.OrderedDict()
    for n, d in zip(names, dem):
        d = re.sub(r"\(int\)|\(bool\)|tdm::|^void ", "", d)
        pretty[re.sub(r"\(.*\)$", "", d)] = out[n]
    return pretty


def main() -> None:
    table = kernel_mnemonics()
    print(f"# cuobjdump -sass {LIB.name}: static counts of tcgen05 / TMEM / bulk-copy instructions per kernel (sm_100a)")
    print("# " + ", ".join(WANTED))
    for name, c in table.items():
        if any(c.values()):
            print(f"{name:72s} " + " ".join(f"{k}={c[k]}" for k in WANTED if c[k]))
    plain = [n for n, c in table.items() if not any(c.values())]
    print(f"# {len(plain)} SIMT kernels without tensor-core / bulk-copy instructions: " + ", ".join(plain))


if __name__ == "__main__":
    sys.exit(main())
