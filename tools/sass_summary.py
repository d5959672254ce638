#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code paths
(/opt/skills/guides/B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA load/store -> UTMALDG/UTMASTG, mma.sync -> HMMA, cp.async -> LDGSTS.
  python tools/sass_summary.py [path/to/libb200.so] > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "gpustack_b200" / "csrc" / "libb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDGSTS", "MUFU.EX2"]
counts = collections.defaultdict(collections.Counter)
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    if name is None:
        continue
    m = re.search(r"\s(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|HMMA|LDGSTS|MUFU)((?:\.[A-Za-z0-9_]+)*)\s", line)
    if not m:
        continue
    op, suffix = m.group(1), m.group(2)
    if op in ("UTCHMMA", "UTCQMMA"):
        op += ".2CTA" if ".2CTA" in suffix else ""
    if op == "MUFU":
        if ".EX2" not in suffix:
            continue
        op = "MUFU.EX2"
    counts[name][op] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
agg = collections.defaultdict(collections.Counter)
for n, d in zip(names, dem):
    d = d.split("(")[0].replace("void ", "").replace("b200::", "")
    agg[d].update(counts[n])
print("# SASS mnemonic counts per kernel of gpustack_b200/csrc/libb200.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma -> UTCHMMA / UTCQMMA(.2CTA); tcgen05.ld/st -> LDTM/STTM; TMA load/store -> UTMALDG/UTMASTG;")
print("# mma.sync -> HMMA; cp.async -> LDGSTS.  Kernels with none of these are omitted.")
print("| kernel | " + " | ".join(KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
tot = collections.Counter()
for d in sorted(agg):
    c = agg[d]
    if not any(c[k] for k in KEYS):
        continue
    tot.update(c)
    print(f"| {d} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
print("| TOTAL | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
