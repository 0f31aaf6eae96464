#!/usr/bin/env python3
"""Static SASS instruction census per kernel: python tools/sass_count.py <binary-or-.so> [name-substring ...]
Counts every instruction of each matching function (unrolled loop bodies dominate), by opcode."""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pats = sys.argv[2:]
fn, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and fn:
        ins = m.group(1).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        counts[fn][op.split(".")[0]] += 1
for fn, c in counts.items():
    if pats and not any(p in fn for p in pats):
        continue
    tot = sum(c.values())
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
    print("%s\n  total %d: %s" % (dem, tot, ", ".join("%s %d" % kv for kv in c.most_common(16))))
