"""SASS mnemonic counts per kernel: cuobjdump -sass <lib.so> | python tools/sass_summary.py > profiles/rNN_sass_summary.txt
(tcgen05 MMA = UTCHMMA, TMEM = LDTM / STTM, TMA = UTMALDG / UBLKCP, mma.sync = HMMA, mbarrier = SYNCS)."""
import collections
import re
import subprocess
import sys

pat = re.compile(r"\b(UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|SYNCS|ATOMS|CCTL)\b")
counts, cur = collections.OrderedDict(), None
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
    elif cur:
        for k in pat.findall(line.split("/*")[1] if line.count("/*") >= 2 else line):
            counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(__doc__.strip().splitlines()[-1])
tot = collections.Counter()
for (fn, c), name in zip(counts.items(), names):
    if c:
        print(f"{name.split('(')[0][:96]:96s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
        tot.update(c)
print(f"TOTAL over {len(counts)} kernels: " + " ".join(f"{k}={v}" for k, v in sorted(tot.items())))
