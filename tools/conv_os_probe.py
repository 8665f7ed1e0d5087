"""Per-layer timing and in-kernel timeline of the output-stationary convolution (csrc/conv_os.cu) on the real kernel
maps of a synthetic batch, next to the pair-major GEMM + sorted scatter it replaces.

    python tools/conv_os_probe.py [--workload nuscenes] [--batch 8] [--chunk 0]

Times are device time per launch from CUDA graphs of 20 back-to-back launches (no host launch overhead), the input
rewritten between graphs.  The timeline columns come from the kernel's optional trace buffer (globaltimer stamps):
per CTA, microseconds from its start to the end of its producer / MMA / epilogue roles; `passes` = passes per CTA.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LAYERS = [  # (kernel, stride, cin, cout)
    (3, 1, 32, 32), (3, 1, 96, 96), (3, 1, 128, 96), (3, 2, 32, 32), (3, 2, 96, 96), (3, 4, 64, 64), (3, 4, 128, 128),
    (3, 4, 192, 128), (3, 8, 128, 128), (3, 8, 256, 256), (3, 8, 384, 256), (3, 16, 256, 256),
    (2, 1, 32, 32), (2, 4, 64, 64), (2, 8, 128, 128),
]


def graph_time(fn, iters=20, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    torch.cuda.current_stream().wait_stream(s)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nuscenes")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--stats", type=int, default=1)
    ap.add_argument("--cluster", type=int, default=1, help="CTAs per cluster sharing B (1, 2, 4)")
    ap.add_argument("--detail", action="store_true", help="per-CTA breakdown of the slowest CTAs")
    ap.add_argument("--fine", type=int, default=0, help="print the first N per-stage stamps of CTA 0")
    ap.add_argument("--only", default="", help="comma list of layer indices")
    ap.add_argument("--modes", default="tma,ldgsts", help="gather modes to time (ncu captures: one mode)")
    ap.add_argument("--timeline-gather", default="tma", choices=["tma", "ldgsts"], help="gather mode of the traced launch")
    a = ap.parse_args()
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import conv_engine, dataflow, ops
    from fusiontransformer_b200.synthetic import make_scan
    from fusiontransformer_b200.voxel_glue import initial_voxelize
    spf = ft.nn.functional
    dev = torch.device("cuda", 0)
    ops.OS_CHUNK_PASSES = a.chunk
    scans = [make_scan(a.workload, i) for i in range(a.batch)]
    db = dataflow.to_device(dataflow.host_batch_from_scans(scans), dev)
    lidar, *_ = dataflow.voxelize_batch(db)
    z = ft.PointTensor(lidar.F, lidar.C.float())
    x0 = initial_voxelize(z, 1, 1)
    coords = {1: x0.C}
    for s in (2, 4, 8, 16):
        coords[s] = spf.spdownsample(coords[s // 2], s)
    print("workload %s batch %d chunk %d: voxels per stride %s" % (a.workload, a.batch, a.chunk,
                                                                   {s: c.shape[0] for s, c in coords.items()}))
    print("%-16s %7s %7s %6s %5s %4s | %8s %8s %8s %7s | %s" % (
        "layer", "rows", "pairs", "passes", "units", "cap", "os-tma", "os-ldg", "pairs us", "TF/s", "timeline (tma): passes/CTA max,mean; us start->producer, mma, epilogue, stats end (max over CTAs)"))
    maps = {}
    layers = [LAYERS[int(i)] for i in a.only.split(",")] if a.only else LAYERS
    for ks, s, cin, cout in layers:
        key = (ks, s)
        if key not in maps:
            ci = coords[s]
            co = coords[s] if ks == 3 else coords[2 * s]
            maps[key] = spf.build_kernel_map(ci, co, ks, s)
        km = maps[key]
        L = km.num_pairs()
        g = torch.Generator(device=dev).manual_seed(cin * 7 + cout)
        x16 = ops.to_bf16(torch.randn(km.n_in, cin, device=dev, generator=g))
        w = torch.nn.Parameter(torch.randn(ks ** 3, cin, cout, device=dev, generator=g) * 0.05)
        os.environ["FT3D_OS_CLUSTER"] = str(a.cluster)
        plan = km.os_plan("out", 128 * a.cluster)
        P, U, S, cap, NS = plan.host_counts()
        bn = (1e-5, 0.1, None, None) if a.stats else None
        t = {}
        for mode in ("tma", "ldgsts"):
            os.environ["FT3D_OS_GATHER"] = mode
            t[mode] = (graph_time(lambda: conv_engine.os_conv(x16, km, w, "forward", bn=bn))
                       if ((mode == "tma" or a.cluster == 1) and mode in a.modes.split(",")) else float("nan"))
        os.environ["FT3D_OS_GATHER"] = a.timeline_gather
        km.ppos
        tp = graph_time(lambda: ops.conv_reduce_bn(conv_engine.pairs_partial(x16, km, w, "forward")[0], km.ppos, cout,
                                                  1e-5, 0.1, None, None))
        ops.OS_TRACE = []
        conv_engine.os_conv(x16, km, w, "forward", bn=bn)
        torch.cuda.synchronize()
 
        raw = ops.OS_TRACE[0].cpu().numpy().astype(np.int64)
        tr, fine = raw[:148 * 8].reshape(148, 8), raw[148 * 8:]
        if a.fine:
            t00 = tr[0, 0]
            st = fine[:2048].reshape(512, 4)
            n = int((st[:, 0] > 0).sum())
            print("      CTA 0 stages (us since CTA start): issued / landed / committed; landed-issued")
            for i in range(min(n, a.fine)):
                print("        stage %3d: %7.2f %7.2f %7.2f   fill %5.2f" % (
                    i, (st[i, 0] - t00) / 1e3, (st[i, 1] - t00) / 1e3, (st[i, 2] - t00) / 1e3, (st[i, 1] - st[i, 0]) / 1e3))
            un = fine[2048:2048 + 128].reshape(64, 2)
            for i in range(int((un[:, 0] > 0).sum())):
                print("        unit %2d: accumulator ready %7.2f rows written %7.2f" % (
                    i, (un[i, 0] - t00) / 1e3, (un[i, 1] - t00) / 1e3))
        tr = tr[tr[:, 0] > 0]                    # CTAs that ran (the grid is min(tiles, 148))
        ops.OS_TRACE = None
        t0 = tr[:, 0]
        rel = lambda c: (tr[:, c] - t0) / 1e3
        flops = 2.0 * L * cin * cout
        if a.detail:
            order = np.argsort(-(tr[:, 3] - t0))[:4]
            for b in list(order) + [int(np.argmin(tr[:, 3] - t0))]:
                print("      cta %3d: passes %2d units %d split-units %d | producer %.1f mma %.1f epilogue %.1f us" % (
                    b, tr[b, 5], tr[b, 6] & 0xffff, tr[b, 6] >> 32, rel(1)[b], rel(2)[b], rel(3)[b]))
        print("k%d s%-2d %3d->%-3d  %7d %7d %6d %5d %4d | %8.1f %8.1f %8.1f %7.1f | %d,%.1f; %.1f %.1f %.1f %.1f; span %.1f" % (
            ks, s, cin, cout, km.n_out, L, P, U, cap, t["tma"], t["ldgsts"], tp, flops / t["tma"] / 1e6,
            tr[:, 5].max(), tr[:, 5].mean(), rel(1).max(), rel(2).max(), rel(3).max(), rel(4).max(),
            (tr[:, 4].max() - t0.min()) / 1e3))


if __name__ == "__main__":
    main()
