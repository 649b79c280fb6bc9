#!/usr/bin/env python
"""Instruction mix of the steady-state tile body of attn_small_kernel<32,3,6,split,4>: the SASS between the first
`LDTM.x` of a tile and the first `STTM` of its steady-state path (the P store) (cuobjdump -sass of the object file).
Usage: python tools/sass_tile_mix.py healnet_b200/csrc/build/xattn_small.o [mangled-name substring]"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "attn_small_kernelILi32ELi3ELi6ELb1ELi4E"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = None
body = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and want in fn:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            body.append(m.group(1).strip())
first = next(i for i, x in enumerate(body) if x.startswith("LDTM.x"))
last = next(i for i in range(first, len(body)) if body[i].startswith("STTM"))
mix = collections.Counter()
for ins in body[first:last + 1]:
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    mix[ins.split()[0].split(".")[0] + ("." + ins.split()[0].split(".")[1] if ins.startswith(("MUFU", "IMAD", "VIMNMX", "VIADDMNMX")) and "." in ins.split()[0] else "")] += 1
total = sum(mix.values())
print("steady-state tile body: %d instructions per thread per 64 columns (%.2f per element)" % (total, total / 64))
for k, v in mix.most_common():
    print("  %-22s %4d" % (k, v))
