"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_src_lines.py rep kernel-regex [top-N]"""
import csv, io, subprocess, sys, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None
lines = collections.OrderedDict()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not r or r[0] in ("Function Name", "") or hdr is None:
        continue
    # a source line: fields from the end are stable even if the source text holds commas
    n = len(hdr)
    tail = r[-(n - 4):]            # from 'Warp Stall Sampling (All Samples)' on
    try:
        samples, inst, tinst = int(tail[2]), int(tail[3]), int(tail[4])
    except ValueError:
        continue
    key = (cur, int(r[0]))
    src = ",".join(r[1:len(r) - (n - 4) - 2]).strip()
    a = lines.setdefault(key, [0, 0, 0, src])
    a[0] += inst; a[1] += tinst; a[2] += samples
tot = sum(v[0] for v in lines.values()) or 1
tots = sum(v[2] for v in lines.values()) or 1
print(f"total instructions {tot}, samples {tots}")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[2]/tots:5.1f}% smp  lanes {v[1]/max(v[0],1):4.1f}  {f}:{ln}  {v[3][:100]}")
