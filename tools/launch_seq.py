"""Print the last N launches of an ncu launch list (--metrics gpu__time_duration.sum --csv) in order: kernel, microseconds.

    python tools/launch_seq.py gpurun_out/launches.csv [N]
"""
import csv, sys

def main():
    path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0].replace("amp::<unnamed>::", ""), v, r.get("Grid Size", ""), r.get("Block Size", "")))
    for k, v, g, b in rows[-n:]:
        print("%8.2f us  %-60s grid %s block %s" % (v, k[:60], g, b))
main()
