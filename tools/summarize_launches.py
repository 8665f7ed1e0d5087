"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.

The capture starts at process start, so it begins with the bench's set-up (random weights, the synthetic feature maps:
a few hundred-microsecond ATen fills); launches before the first libft3d kernel are reported separately."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name)
        rows.append((name, ns))
    is_ours = lambda k: "ft3d" in k or k.startswith(("conv_", "kmap_", "to_bf16"))
    first = next((i for i, (n, _) in enumerate(rows) if is_ours(n)), 0)
    if first:
        print("set-up before the first libft3d kernel: %d launches, %.3f ms (not in the table)" % (
            first, sum(t for _, t in rows[:first]) / 1e6))
        rows = rows[first:]
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if is_ours(k))
    print("launches %d  total %.3f ms  (libft3d kernels %.1f %% of captured GPU time)" % (len(rows), tot / 1e6, 100 * ours / max(tot, 1)))
    print("%8s %10s %7s  %s" % ("count", "total us", "share", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%8d %10.1f %6.1f%%  %s" % (v[0], v[1] / 1e3, 100 * v[1] / tot, k[:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
