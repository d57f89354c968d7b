"""Opcode histogram of the shipped library's SASS (no GPU needed): the evidence that the hot kernels are tcgen05 / TMA code.
    python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gdn_pytorch_b200", "libgdn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per_fn, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and fn:
        op = m.group(1)
        per_fn[fn][op] += 1
        total[op.split(".")[0]] += 1
KEY = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "HMMA", "REDG", "RED", "ATOM", "SHFL", "UCGABAR")
print("# cuobjdump -sass %s : opcode histogram (tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, TMA = UTMALDG, "
      "tcgen05.ld = LDTM, mbarrier = SYNCS; HMMA would be legacy mma.sync)" % os.path.relpath(LIB, ROOT))
print("total instructions: %d in %d kernels" % (sum(total.values()), len(per_fn)))
for k in KEY:
    full = collections.Counter()
    for c in per_fn.values():
        for op, n in c.items():
            if op.split(".")[0] == k:
                full[op] += n
    print("%-12s %6d   %s" % (k, sum(full.values()), ", ".join("%s x%d" % kv for kv in full.most_common(8))))
print()
print("%-70s %8s %8s %8s %8s %8s" % ("kernel", "instrs", "UTCHMMA", "UTMALDG", "LDTM", "HMMA"))
for fn_, c in sorted(per_fn.items(), key=lambda kv: -sum(kv[1].values())):
    g = lambda k: sum(n for op, n in c.items() if op.split(".")[0] == k)
    print("%-70s %8d %8d %8d %8d %8d" % (fn_[:70], sum(c.values()), g("UTCHMMA"), g("UTMALDG"), g("LDTM"), g("HMMA")))
