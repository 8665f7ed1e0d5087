"""Offline view of `bench.py --trace N --trace-dump step.csv`: the launches of one step in time order, per stream,
with gaps; and per-kernel totals restricted to a stream."""
import csv
import sys


def load(path):
    rows = list(csv.DictReader(open(path)))
    for r in rows:
        r["ts"], r["dur"] = float(r["ts_us"]), float(r["dur_us"])
        r["short"] = r["name"].replace("void ", "").replace("ft3d::", "").split("(")[0][:48]
    return rows


def main():
    path = sys.argv[1]
    rows = load(path)
    lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    hi = float(sys.argv[3]) if len(sys.argv) > 3 else 1e18
    last_end = {}
    for r in rows:
        gap = r["ts"] - last_end.get(r["stream"], r["ts"])
        last_end[r["stream"]] = r["ts"] + r["dur"]
        if lo <= r["ts"] <= hi:
            print("%9.1f %7.1f gap%7.1f s%-4s g%-10s b%-9s %s" % (r["ts"], r["dur"], gap, r["stream"], r["grid"], r["block"], r["short"]))


if __name__ == "__main__":
    main()
