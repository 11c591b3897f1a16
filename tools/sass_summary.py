"""Opcode histogram per kernel of libtribe_b200.so (`cuobjdump -sass`): the evidence that the hot kernels really use the
sm_100a machinery (tcgen05 = UTCHMMA / UTCBAR / LDTM, TMA loads = UTMALDG, TMA stores = UTMASTG, mbarrier = SYNCS) and how
heavy each epilogue is.  No GPU needed.  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "algonauts-2025_b200", "csrc", "libtribe_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UBLKPF", "HMMA", "STG", "LDG", "STS", "LDS", "STL", "LDL", "ATOM", "RED", "MUFU", "MEMBAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o).replace("void ", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            cur[m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM"):
                cur["_" + full] += 1
    names = demangle(list(kernels))
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}   arch: {', '.join(sorted(arch))}   kernels: {len(kernels)}")
    print("instruction counts per kernel (static SASS); columns: total | " + " ".join(KEY))
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), names):
        print(f"{name[:78]:78s} {c['_total']:6d} | " + " ".join(f"{c[k]:5d}" if c[k] else "    ." for k in KEY))
        tot.update(c)
    print(f"{'ALL KERNELS':78s} {tot['_total']:6d} | " + " ".join(f"{tot[k]:5d}" if tot[k] else "    ." for k in KEY))
    print("\nvariants of the tensor-core / TMA opcodes over the whole library:")
    for k, v in sorted(tot.items()):
        if k.startswith("_") and k != "_total":
            print(f"  {k[1:]:40s} {v}")


if __name__ == "__main__":
    sys.exit(main())
