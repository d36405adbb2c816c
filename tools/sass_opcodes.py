#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libpnvo.so (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
    cuobjdump -sass pointnav-vo_b200/csrc/libpnvo.so | python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

OPS = ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "HMMA", "ATOMG", "RED")
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        counts[cur][m.group(1)] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# per-kernel SASS opcode counts of pointnav-vo_b200/csrc/libpnvo.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA bulk-tensor load = UTMALDG, cp.async.bulk = UBLKCP, tcgen05.commit = UTCBAR,")
print("# mbarrier ops = SYNCS, elect.sync = ELECT; HMMA (mma.sync) must be 0 everywhere")
print("%-78s %s" % ("kernel", " ".join("%8s" % o for o in OPS)))
tot = collections.Counter()
for n, d in zip(names, dem):
    c = counts[n]
    if not any(c[o] for o in OPS[:5]):
        continue
    short = d.split("(")[0].replace("pnvo::", "").replace("void ", "")
    print("%-78s %s" % (short[:78], " ".join("%8d" % c[o] for o in OPS)))
    for o in OPS:
        tot[o] += c[o]
print("%-78s %s" % ("TOTAL over the kernels above", " ".join("%8d" % tot[o] for o in OPS)))
print("kernels in the library: %d; with tcgen05 / TMA / bulk-copy instructions: %d" %
      (len(names), sum(1 for n in names if any(counts[n][o] for o in OPS[:5]))))
