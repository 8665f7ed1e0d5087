"""Device time of the BatchNorm streaming kernels (csrc/bn.cu) at the bench's layer shapes, per launch, from a CUDA graph
of 20 back-to-back launches; sweeps the CTA cap (FT3D_COL_CTAS) and the rows-in-flight batch is FT3D_ROWBATCH.

    python tools/bn_probe.py > profiles/r02_bn_probe.txt

Algorithmic bytes per element: bn_stats 4 (y); bn_apply 4 (y) + 4 (z) + 2 (z16) [+ 4 res]; bn_bwd_reduce 4 (gz) + 4 (y)
+ 2 (mask); bn_bwd_apply the same reads + 2 (gy16) [+ 4 gres].  GB/s = those bytes / time; of-HBM = / MEASURED_PEAKS hbm.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

SHAPES = [(49964, 32), (49964, 96), (40480, 64), (31751, 128), (20390, 256), (10824, 256), (55312, 64)]


def main():
    from conv_os_probe import graph_time
    from fusiontransformer_b200 import ops
    dev = torch.device("cuda", 0)
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    caps = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "148,296,592,1184".split(","))]
    print("# HBM peak %.0f GB/s; us per launch (GB/s) for CTA caps %s" % (hbm, caps))
    print("%-14s %-14s " % ("rows x C", "kernel") + " ".join("%16s" % ("cap %d" % c) for c in caps))
    for n, c in SHAPES:
        g = torch.Generator(device=dev).manual_seed(n + c)
        y = torch.randn(n, c, device=dev, generator=g)
        gz = torch.randn(n, c, device=dev, generator=g)
        res = torch.randn(n, c, device=dev, generator=g)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        stat = ops.bn_stats(y, 1e-5, 0.1, None, None)
        z, z16 = ops.bn_apply(y, stat, gamma, beta, None, True)
        red, _, _ = ops.bn_bwd_reduce(gz, y, z16, None, stat)
        runs = [("bn_stats", 4, lambda: ops.bn_stats(y, 1e-5, 0.1, None, None)),
                ("bn_apply", 10, lambda: ops.bn_apply(y, stat, gamma, beta, None, True)),
                ("bn_apply+res", 14, lambda: ops.bn_apply(y, stat, gamma, beta, res, True)),
                ("bwd_reduce", 10, lambda: ops.bn_bwd_reduce(gz, y, z16, None, stat)),
                ("bwd_apply", 12, lambda: ops.bn_bwd_apply(gz, y, z16, None, stat, gamma, red, False, True, False)),
                ("bwd_apply+res", 16, lambda: ops.bn_bwd_apply(gz, y, z16, None, stat, gamma, red, False, True, True))]
        for name, bpe, fn in runs:
            cells = []
            for cap in caps:
                os.environ["FT3D_COL_CTAS"] = str(cap)
                us = graph_time(fn)
                gbs = n * c * bpe / us / 1e3
                cells.append("%6.1f (%5.0f %2.0f%%)" % (us, gbs, 100 * gbs / hbm))
            print("%-14s %-14s " % ("%dx%d" % (n, c), name) + " ".join("%16s" % s for s in cells))
    os.environ.pop("FT3D_COL_CTAS", None)


if __name__ == "__main__":
    main()
