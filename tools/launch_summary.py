"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) : per-kernel count, total and share.

    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n] [out.txt]
"""
import csv, sys
from collections import OrderedDict

def main():
    path = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v))
    rows = rows[skip:]
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in rows)
    out = ["%d launches, %.1f us total (cold-cache, serialised: compare shares)" % (len(rows), tot)]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%8.1f us  %5.1f %%  x%-4d %7.2f us/launch  %s" % (t, 100 * t / tot, n, t / n, k))
    txt = "\n".join(out)
    print(txt)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(txt + "\n")
main()
