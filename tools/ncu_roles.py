"""Warp-state samples of a fused-kernel capture, by warp role (SASS address ranges found from marker instructions).
Usage: python tools/ncu_roles.py rep.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:ais_fused"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ins.append((int(r[0], 16), r[1].strip(), r))
    except ValueError:
        pass
base = ins[0][0]
def find(pred, start=0):
    for a, t, r in ins:
        if a - base >= start and pred(t):
            return a - base
    return None
# role boundaries: --bounds 0x2ad0:issuer,0x3900:resolver,... (offsets from the kernel's first instruction; print the
# markers with --markers to find them: the roles' code order changes from build to build)
bounds = [(0, "setup")]
for arg in sys.argv[2:]:
    if arg.startswith("--bounds="):
        for part in arg[9:].split(","):
            off, name = part.split(":")
            bounds.append((int(off, 16), name))
if "--markers" in sys.argv:
    for a, t, r in ins:
        if any(k in t for k in ("NANOSLEEP", "LDTM", "UTCIMMA", "UTMALDG", "BAR.SYNC", "EXIT", "UTCBAR")):
            print(hex(a - base), t[:70])
    sys.exit(0)
def role(off):
    name = bounds[0][1]
    for b, n in bounds:
        if off >= b:
            name = n
    return name
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(collections.Counter)
for a, t, r in ins:
    ro = role(a - base)
    def g(name):
        try:
            return int(r[ci[name]])
        except (ValueError, KeyError):
            return 0
    agg[ro]["samples"] += g("# Samples")
    agg[ro]["inst"] += g("Instructions Executed")
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            agg[ro][h] += g(h)
tot = sum(v["samples"] for v in agg.values()) or 1
print("role boundaries:", [(hex(b), n) for b, n in bounds])
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted([(c, h) for h, c in v.items() if h.startswith("stall_")], reverse=True)[:6]
    print(f"{n:42s} samples {100*v['samples']/tot:5.1f}%  inst {v['inst']/1e6:8.1f}M  " +
          ", ".join(f"{h[6:]} {100*c/max(v['samples'],1):.0f}%" for c, h in st))
pos = [x for x in sys.argv[2:] if not x.startswith("--")]
if pos:
    want = pos[0]
    top = []
    for a, t, r in ins:
        if role(a - base) == want:
            try:
                top.append((int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), hex(a - base), t[:70]))
            except ValueError:
                pass
    for x in sorted(top, reverse=True)[:int(pos[1]) if len(pos) > 1 else 25]:
        print(f"  {100*x[0]/tot:5.2f}% smp  {x[1]/1e6:7.2f}M  {x[2]:>7s}  {x[3]}")
