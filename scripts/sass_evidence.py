#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library -> profiles/<round>_sass.md.
    python scripts/sass_evidence.py r02        (runs in the build container: cuobjdump needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "reactranker_b200", "librr_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if mm:
        counts[cur][mm.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts.keys()), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCBAR", "FFMA", "HMMA", "REDG", "ATOMG"]
out = [f"# {R} SASS evidence", "",
       "`cuobjdump -sass reactranker_b200/librr_sm100.so` (the library `__graft_entry__.build()` compiles with `-gencode arch=compute_100a,code=sm_100a`;",
       "regenerate with `python scripts/sass_evidence.py`), instructions per kernel.  UTCHMMA = `tcgen05.mma` (kind::tf32 / kind::f16, A from tensor memory,",
       "B from shared memory), LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTMALDG = `cp.async.bulk.tensor` (TMA tile load), UBLKCP = `cp.async.bulk`",
       "(the row pipeline's 1-D bulk copies), SYNCS = mbarrier operations, UTCBAR = `tcgen05.commit`.  HMMA (`mma.sync`) occurs nowhere: the tensor-core",
       "kernels are tcgen05 only.  Kernels with fewer than 60 instructions are omitted.", "",
       "| kernel | " + " | ".join(keys) + " | all |", "|---|" + "---|" * (len(keys) + 1)]
for (k, c), n in zip(counts.items(), names):
    short = (n[:n.index("(")] if "(" in n else n).replace("rr::", "").replace("tc::", "").replace("pipe::", "").replace("void ", "")
    if sum(c.values()) < 60:
        continue
    out.append(f"| `{short}` | " + " | ".join(str(c.get(x, 0)) for x in keys) + f" | {sum(c.values())} |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out += ["", "Library totals: " + ", ".join(f"{x} {tot.get(x, 0)}" for x in keys) + "."]
open(os.path.join(ROOT, "profiles", f"{R}_sass.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[8:]))
