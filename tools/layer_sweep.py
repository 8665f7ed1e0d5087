"""BASELINE.json configs[3]: sparse Conv3d layer sweep -- C_in/C_out 32..384 -> 32..256, kernel 3 at tensor stride 1 and 2
(+ the model's k2 s2 down / transposed up), 20k-200k active voxels drawn from KITTI-shaped scans; forward, dgrad and
wgrad separately, each with BOTH roofline bounds.

    python tools/layer_sweep.py [--sizes 20000,50000,100000,200000] > profiles/r02_layer_sweep.txt

Times are device time per launch from a CUDA graph of 20 back-to-back launches (no host overhead); forward / dgrad =
ft3d_conv_os (+ its fold launch when the schedule has split tiles); wgrad = the persistent tcgen05 wgrad kernel.
Algorithmic work per SURVEY 8(d): F = 2 L Cin Cout;  bytes(fwd/dgrad) = 2 (N_in red + K red ncols) + 4 N_out ncols + 8 L
(bf16 operands, fp32 result), bytes(wgrad) = 2 (N_in Cin + N_out Cout) + 4 K Cin Cout + 8 L.  frac = achieved / peak for
the bound that binds (the larger of F / peak_tf and B / peak_bw), peaks from MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

CHANNELS = [(32, 32), (64, 64), (96, 96), (128, 128), (192, 128), (256, 256), (384, 256), (128, 256), (32, 64)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20000,50000,100000,200000")
    ap.add_argument("--channels", default="")
    a = ap.parse_args()
    from conv_os_probe import graph_time
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import conv_engine, dataflow, ops
    from fusiontransformer_b200.synthetic import make_scan
    from fusiontransformer_b200.voxel_glue import initial_voxelize
    spf = ft.nn.functional
    dev = torch.device("cuda", 0)
    bw, tf, src = peaks()
    chans = [tuple(int(v) for v in c.split(":")) for c in a.channels.split(",")] if a.channels else CHANNELS
    print("# peaks (%s): HBM %.1f GB/s, bf16 burst %.1f TFLOP/s (kernels timed alone)" % (src, bw, tf))
    print("%-4s %-3s %-9s %8s %9s %6s | %-5s %8s %8s %7s %7s %-6s %6s" % (
        "N", "map", "Cin->Cout", "N_out", "pairs", "passes", "pass", "us", "TF/s", "GB/s", "MB", "bound", "frac"))
    for n_target in [int(v) for v in a.sizes.split(",")]:
        scans, tot = [], 0
        while tot < n_target * 1.05 and len(scans) < 64:
            scans.append(make_scan("kitti", len(scans)))
            tot += len(scans[-1]["points"])
        db = dataflow.to_device(dataflow.host_batch_from_scans(scans), dev)
        lidar, *_ = dataflow.voxelize_batch(db)
        C = lidar.C[:n_target].contiguous()            # whole scans plus a prefix of the last one
        z = ft.PointTensor(torch.zeros(C.shape[0], 4, device=dev), C.float())
        c1 = initial_voxelize(z, 1, 1).C
        c2 = spf.spdownsample(c1, 2)
        maps = {"k3s1": (spf.build_kernel_map(c1, c1, 3, 1), c1, c1), "k3s2": (spf.build_kernel_map(c2, c2, 3, 2), c2, c2),
                "k2dn": (spf.build_kernel_map(c1, c2, 2, 1), c1, c2)}
        for mname, (km, cin_c, cout_c) in maps.items():
            L = km.num_pairs()
            P = km.os_plan("out").host_counts()[0]
            for cin, cout in chans:
                if mname == "k2dn" and (cin, cout) not in ((32, 32), (64, 64), (128, 128), (256, 256)):
                    continue
                g = torch.Generator(device=dev).manual_seed(cin * 7 + cout)
                x16 = ops.to_bf16(torch.randn(km.n_in, cin, device=dev, generator=g))
                g16 = ops.to_bf16(torch.randn(km.n_out, cout, device=dev, generator=g))
                w = torch.nn.Parameter(torch.randn(km.K, cin, cout, device=dev, generator=g) * 0.05)
                flops = 2.0 * L * cin * cout
                by_f = 2.0 * (km.n_in * cin + km.K * cin * cout) + 4.0 * km.n_out * cout + 8.0 * L
                by_d = 2.0 * (km.n_out * cout + km.K * cin * cout) + 4.0 * km.n_in * cin + 8.0 * L
                by_w = 2.0 * (km.n_in * cin + km.n_out * cout) + 4.0 * km.K * cin * cout + 8.0 * L
                runs = [("fwd", lambda: conv_engine.os_conv(x16, km, w, "forward"), by_f),
                        ("dgrad", lambda: conv_engine.os_conv(g16, km, w, "dgrad"), by_d)]
                if cout <= 256:
                    runs.append(("wgrad", lambda: conv_engine.pairs_wgrad(x16, g16, km, cin, cout, False), by_w))
                for pname, fn, by in runs:
                    try:
                        us = graph_time(fn)
                    except Exception as e:  # noqa: BLE001
                        print("%-4dk %-4s %3d->%-3d   %s failed: %s" % (n_target // 1000, mname, cin, cout, pname, str(e)[:60]))
                        continue
                    tfs, gbs = flops / us / 1e6, by / us / 1e3
                    ft_, fb = tfs / tf, gbs / bw
                    bound, frac = ("tensor", ft_) if flops / (tf * 1e12) >= by / (bw * 1e9) else ("hbm", fb)
                    print("%-4s %-4s %3d->%-3d %8d %9d %6d | %-5s %8.1f %8.1f %7.0f %7.1f %-6s %6.3f" % (
                        "%dk" % (n_target // 1000), mname, cin, cout, km.n_out, L, P, pname, us, tfs, gbs, by / 1e6, bound, frac))
        del maps
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
