"""Opcode histogram of a cuobjdump -sass listing, optionally restricted to an address range (hex)."""
import re
import sys
from collections import Counter

path = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
c = Counter()
for line in open(path):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if not (lo <= a < hi):
        continue
    toks = m.group(2).split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += 1
print(sum(c.values()), "instructions")
for k, v in c.most_common(40):
    print(f"{v:5d} {k}")
