"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) into the few numbers the
roofline discussion needs.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_register_spilling",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    for r in rows:
        print(f"== {r[hdr.index('Kernel Name')][:60]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                print(f"   {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("   stalls (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if not rows:
            return
        h = rows[0]
        try:
            ci = h.index("# Samples") if "# Samples" in h else [i for i, x in enumerate(h) if "Samples" in x][0]
            si = h.index("Source")
        except (ValueError, IndexError):
            print(h)
            return
        body = []
        for r in rows[1:]:
            try:
                body.append((int(r[ci]), r[si].strip()[:110]))
            except (ValueError, IndexError):
                pass
        tot = sum(b[0] for b in body) or 1
        for c, s in sorted(body, reverse=True)[:n]:
            print(f"   {100 * c / tot:5.1f}%  {s}")


if __name__ == "__main__":
    main()
