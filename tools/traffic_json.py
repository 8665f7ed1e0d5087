"""profiles/traffic_<kernel>.json from an ncu per-launch CSV (dram / L2 bytes, duration, tensor-pipe %).

    python tools/traffic_json.py profiles/r02_ncu_conv_os_traffic_per_launch.csv conv_os_kernel > profiles/traffic_conv_os.json
"""
import csv
import json
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3, "%": 1.0}


def main(path, kernel):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    per = defaultdict(dict)
    for r in csv.DictReader(lines):
        if kernel not in r["Kernel Name"]:
            continue
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    n = len(per)
    mean = lambda k: sum(v.get(k, 0.0) for v in per.values()) / max(n, 1)
    tp = [v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for v in per.values()]
    out = {"kernel": kernel, "launches": n,
           "dram_bytes_per_launch": mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum"),
           "dram_read_bytes_per_launch": mean("dram__bytes_read.sum"),
           "dram_write_bytes_per_launch": mean("dram__bytes_write.sum"),
           "l2_bytes_per_launch": mean("lts__t_bytes.sum"),
           "mean_duration_us": mean("gpu__time_duration.sum"),
           "tensor_pipe_pct_mean": sum(tp) / max(n, 1), "tensor_pipe_pct_max": max(tp) if tp else 0.0,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,"
                     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:%s -s 164 -c 82 (the 82 "
                     "launches of one eager training step, nuScenes-shaped batch 8; cold L2, serialised)" % kernel}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
