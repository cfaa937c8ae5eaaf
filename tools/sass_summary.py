"""Writes profiles/r2_sass_opcodes.md: per kernel template, the count of Blackwell-specific SASS opcodes in the in-tree
libbdlru.so (cuobjdump -sass; runs without a GPU).    python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "datamining_recblr_b200", "libbdlru.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
per = collections.defaultdict(collections.Counter)
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        per[name][m.group(1)] += 1
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDGSTS", "MUFU", "HMMA"]
agg = collections.OrderedDict()
for n, c in per.items():
    base = re.sub(r"<.*", "", re.sub(r"\(.*", "", n).replace("void ", "").replace("bdlru::", ""))
    a = agg.setdefault(base, [0, 0, [0] * len(keys)])
    a[0] += 1
    a[1] += sum(c.values())
    a[2] = [x + c.get(k, 0) for x, k in zip(a[2], keys)]
with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.md"), "w") as f:
    f.write("# SASS opcode summary of libbdlru.so (round 2)\n\n`cuobjdump -sass datamining_recblr_b200/libbdlru.so` (all cubins "
            "sm_100a), counted per kernel template (summed over its\ninstantiations) by tools/sass_summary.py.  UTCHMMA = "
            "tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM/STTM = tcgen05.ld/st (TMEM),\nUTCBAR = tcgen05.commit, SYNCS = "
            "mbarrier ops, LDGSTS = cp.async; HMMA (legacy mma.sync) must be 0.\n\n")
    f.write("| kernel template | instantiations | SASS instructions | " + " | ".join(keys) + " |\n|---|---|---|" + "---|" * len(keys) + "\n")
    for base, (n, tot, vals) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{base}` | {n} | {tot} | " + " | ".join(str(v) for v in vals) + " |\n")
    tot = [sum(a[2][i] for a in agg.values()) for i in range(len(keys))]
    f.write("| **total** | | | " + " | ".join(str(v) for v in tot) + " |\n")
