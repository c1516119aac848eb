"""Opcode histogram of a SASS address range: python tools/sass_hist.py build/fir.sass 0x1310 0x51b0"""
import sys, collections
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
c = collections.Counter()
for line in open(sys.argv[1]):
    p = line.split()
    a = int(p[0], 16)
    if lo <= a <= hi:
        op = p[1]
        if op.startswith("@"):
            op = p[2]
        c[op.rstrip(";")] += 1
tot = sum(c.values())
print("total", tot)
for k, v in c.most_common():
    print(f"{v:5d} {k}")
