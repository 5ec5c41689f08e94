#!/usr/bin/env python
"""Per-kernel SASS summary of libdryv_recon.so: static instruction count, opcode histogram, local-memory traffic and the
Blackwell data-movement instructions (UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = tensor copies, SYNCS = mbarrier)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dryv_b200/csrc/libdryv_recon.so"
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
hist = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and fn:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[fn][op.split(".")[0]] += 1
for fn, h in hist.items():
    if want and want not in fn:
        continue
    total = sum(h.values())
    print(f"{fn}: {total} instructions")
    print("   " + "  ".join(f"{k} {v}" for k, v in h.most_common(40)))
