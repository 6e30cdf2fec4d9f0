"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt [--top-sass 25]
"""
import csv
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_cycles_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("== %s  (id %s)" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        for h, u in zip(hdr, units):
            if any(h.startswith(k) for k in KEYS):
                v = d[h]
                if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.15:
                    continue
                lines.append("  %-90s %s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    # first kernel only
    try:
        hi = next(i for i, r in enumerate(srows) if r and r[0] == "Address")
        sh = srows[hi]
        iS, iE = sh.index("# Samples"), sh.index("Instructions Executed")
        body = []
        for r in srows[hi + 1:]:
            if len(r) != len(sh):
                break
            body.append(r)
        tot_s = sum(int(r[iS] or 0) for r in body) or 1
        tot_e = sum(int(r[iE] or 0) for r in body)
        lines.append("")
        lines.append("-- source page, first kernel: %d SASS instructions, %d warp-instructions executed, %d samples" % (len(body), tot_e, tot_s))
        mix = Counter()
        for r in body:
            t = r[1].split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
            mix[op.split(".")[0]] += int(r[iE] or 0)
        lines.append("   executed-instruction mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot_e, 1)) for k, v in mix.most_common(14)))
        lines.append("   top stall-sample instructions:")
        for r in sorted(body, key=lambda r: -int(r[iS] or 0))[:int(sys.argv[4]) if len(sys.argv) > 4 else 15]:
            lines.append("     %5.1f%%  %s" % (100.0 * int(r[iS] or 0) / tot_s, r[1].strip()[:110]))
    except StopIteration:
        pass
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main()
