"""Hot SASS regions of one kernel in an .ncu-rep: instructions executed and stall samples,
aggregated over address windows.  Usage: python tools/ncu_sass_hot.py rep.ncu-rep kernel_regex [window]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
win = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
h = rows[start]
body = rows[start + 1:end]
ci = h.index("Instructions Executed"); si = h.index("# Samples"); ti = h.index("Thread Instructions Executed"); src = h.index("Source")
tot_i = sum(int(r[ci]) for r in body if r[ci].isdigit()); tot_s = sum(int(r[si]) for r in body if r[si].isdigit())
print(f"total warp-instr {tot_i}, samples {tot_s}, SASS lines {len(body)}")
for k in range(0, len(body), win):
    blk = body[k:k + win]
    ni = sum(int(r[ci]) for r in blk if r[ci].isdigit()); ns = sum(int(r[si]) for r in blk if r[si].isdigit())
    nt = sum(int(r[ti]) for r in blk if r[ti].isdigit())
    if ni * 100 >= tot_i or ns * 100 >= tot_s:
        ops = " ".join(r[src].split()[0] if not r[src].strip().startswith("@") else r[src].split()[1] for r in blk[:win])
        print(f"[{k:4d}] inst {100*ni/tot_i:5.1f}%  samples {100*ns/max(tot_s,1):5.1f}%  lanes {nt/max(ni,1):4.1f}  {ops[:150]}")
