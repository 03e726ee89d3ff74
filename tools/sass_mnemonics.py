"""Count, per kernel of libscnym_b200.so, the SASS instructions that prove the tensor-core / bulk-copy / cluster paths:
    cuobjdump -sass scnym_b200/libscnym_b200.so | python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt
"""
import collections
import re
import sys

KEYS = (("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("UBLKCP/UTMA", r"\bUBLKCP|\bUTMA"), ("SYNCS", r"\bSYNCS"),
        ("HMMA", r"\bHMMA"), ("cluster", r"UCGABAR|\bMAPA\b|\.CLUSTER"))
cur = None
cnt = collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for key, pat in KEYS:
        if re.search(pat, line):
            cnt[cur][key] += 1
print("# SASS evidence (cuobjdump -sass scnym_b200/libscnym_b200.so): per kernel, the number of instructions that issue tcgen05.mma")
print("# (UTC*MMA), read TMEM (LDTM), move tiles with bulk copies (UBLKCP), wait on mbarriers (SYNCS), issue legacy mma.sync (HMMA)")
print("# or use the thread-block cluster (barrier / distributed shared memory).  Kernels with none of these are not listed.")
for k, c in cnt.items():
    if sum(c.values()) == 0:
        continue
    short = re.sub(r"^_Z\d+", "", k)[:92]
    print(f"{short:94s} " + "  ".join(f"{n} {c[n]:3d}" for n, _ in KEYS if c[n]))
