"""Per-layer sparse-conv timing on real kernel maps (BASELINE.json configs[3]: Conv3d layer sweep).

Builds the stride-1..16 voxel pyramid of a synthetic batch, then times forward / dgrad / wgrad of every distinct
(Cin, Cout, kernel, stride) the SPVCNN uses (or a channel sweep with --sweep), 5 warm-up + 20 timed launches each
(CUDA events), and prints achieved TFLOP/s (algorithmic: 2*pairs*Cin*Cout) and the gather-byte bound.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SPVCNN_K3 = {1: [(32, 32), (128, 96), (96, 96)], 2: [(32, 32), (128, 96), (96, 96)], 4: [(32, 64), (64, 64), (192, 128), (128, 128)],
             8: [(64, 128), (128, 128), (384, 256), (256, 256)], 16: [(128, 256), (256, 256)]}
SPVCNN_K2 = {1: [(32, 32), (96, 96)], 2: [(32, 32), (128, 96)], 4: [(64, 64), (256, 128)], 8: [(128, 128), (256, 256)]}


def timeit(fn, warm=5, it=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nuscenes")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import conv_engine, dataflow
    from fusiontransformer_b200.synthetic import make_scan
    from fusiontransformer_b200.voxel_glue import initial_voxelize
    spf = ft.nn.functional
    dev = torch.device("cuda", 0)
    scans = [make_scan(a.workload, i) for i in range(a.batch)]
    db = dataflow.to_device(dataflow.host_batch_from_scans(scans), dev)
    lidar, *_ = dataflow.voxelize_batch(db)
    z = ft.PointTensor(lidar.F, lidar.C.float())
    x0 = initial_voxelize(z, 1, 1)
    coords = {1: x0.C}
    for s in (2, 4, 8, 16):
        coords[s] = spf.spdownsample(coords[s // 2], s)
    rows = []
    print("workload %s batch %d: voxels per stride %s" % (a.workload, a.batch, {s: c.shape[0] for s, c in coords.items()}))
    hdr = "%-3s %-6s %-9s %8s %9s | %9s %8s | %9s %8s | %9s %8s" % ("k", "stride", "Cin->Cout", "N_out", "pairs", "fwd us", "TF/s", "dgrad us", "TF/s", "wgrad us", "TF/s")
    print(hdr)
    for ks, table in ((3, SPVCNN_K3), (2, SPVCNN_K2)):
        for s, chans in table.items():
            cin_c = coords[s]
            cout_c = coords[s] if ks == 3 else coords[2 * s]
            km = spf.build_kernel_map(cin_c, cout_c, ks, s)
            L = km.num_pairs()
            for cin, cout in chans:
                g = torch.Generator(device=dev).manual_seed(cin * 7 + cout)
                x = torch.randn(cin_c.shape[0], cin, device=dev, generator=g)
                go = torch.randn(cout_c.shape[0], cout, device=dev, generator=g)
                w = torch.nn.Parameter(torch.randn(ks ** 3, cin, cout, device=dev, generator=g) * 0.05)
                if conv_engine.pairs_ok(cin, cout):
                    from fusiontransformer_b200 import ops
                    x16, g16 = ops.to_bf16(x), ops.to_bf16(go)
                    km.ppos, km.pposT
                    t_f = timeit(lambda: conv_engine.pairs_conv(ops.to_bf16(x), km, w, "forward"))
                    t_d = timeit(lambda: conv_engine.pairs_conv(ops.to_bf16(go), km, w, "dgrad"))
                    t_w = timeit(lambda: conv_engine.pairs_wgrad(x16, g16, km, cin, cout, False))
                else:
                    t_f = timeit(lambda: conv_engine.gather_conv(x, km.nbr, km, w, False, False))
                    tbl, flip = (km.nbr, True) if km.symmetric else (km.nbrT, False)
                    t_d = timeit(lambda: conv_engine.gather_conv(go, tbl, km, w, flip, True))
                    t_w = timeit(lambda: conv_engine.wgrad(x, go, km, cin, cout, False))
                fl = 2.0 * L * cin * cout
                tf = lambda us: fl / (us * 1e-6) / 1e12  # noqa: E731
                print("%-3d %-6d %-9s %8d %9d | %9.1f %8.2f | %9.1f %8.2f | %9.1f %8.2f" %
                      (ks, s, "%d->%d" % (cin, cout), cout_c.shape[0], L, t_f, tf(t_f), t_d, tf(t_d), t_w, tf(t_w)))
                rows.append(dict(k=ks, stride=s, cin=cin, cout=cout, n_in=cin_c.shape[0], n_out=cout_c.shape[0], pairs=L,
                                 fwd_us=t_f, dgrad_us=t_d, wgrad_us=t_w, gflop=fl / 1e9))
    if a.json:
        json.dump(dict(workload=a.workload, batch=a.batch, rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
