#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (no GPU needed) and writes
profiles/<tag>_sass_extract.md: `tcgen05.mma` = UTCHMMA, `tcgen05.ld/st` = LDTM / STTM, TMA = UTMALDG, `tcgen05.commit` =
UTCBAR, mbarrier = SYNCS.*; HMMA = legacy `mma.sync` (must be 0).   Usage: python tools/sass_extract.py r2b"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2b"
lib = os.path.join(ROOT, "healnet_b200", "libhealnet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "MUFU.EX2", "HFMA2", "HMMA", "SYNCS"]
per, order, cur, i = collections.OrderedDict(), [], None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        n = names[i].replace("(anonymous namespace)::", "").replace("(int)", "").replace("(bool)", "")
        n = re.sub(r"\(.*", "", n).replace("void ", "").replace("hn::", "")
        cur = per.setdefault(n, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["all"] += 1
        for k in keys:
            if op == k or op.startswith(k + ".") or (k == "HMMA" and op.startswith("HMMA")):
                cur[k] += 1
rows = [(n, c) for n, c in per.items() if c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"]]
tot = collections.Counter()
for _, c in per.items():
    tot.update(c)
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_extract.md"), "w") as f:
    f.write(f"# {tag} — SASS evidence of the Blackwell-native kernels in `healnet_b200/libhealnet_b200.so`\n\n")
    f.write("Produced with `python tools/sass_extract.py` (`cuobjdump -sass` on the product build, no GPU needed), counting "
            "mnemonics per kernel: `tcgen05.mma` = `UTCHMMA`, `tcgen05.ld/st` = `LDTM`/`STTM`, TMA `cp.async.bulk.tensor` = "
            "`UTMALDG`, `tcgen05.commit` = `UTCBAR`, mbarrier = `SYNCS.*`; `HMMA` = legacy `mma.sync`, must be 0. Kernels "
            "without tensor-core / TMA instructions (row kernels, the strided fp32 checker, packing) are omitted.\n\n")
    f.write("| kernel | " + " | ".join(keys) + " | all instructions |\n|---|" + "---:|" * (len(keys) + 1) + "\n")
    for n, c in rows:
        f.write(f"| `{n}` | " + " | ".join(str(c[k]) for k in keys) + f" | {c['all']} |\n")
    f.write("| **library total** | " + " | ".join(str(tot[k]) for k in keys) + " | |\n\n")
    f.write(f"{len(per)} kernels in the library, {len(rows)} of them use tcgen05 / TMEM / TMA.\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_extract.md")).read()[-1500:])
