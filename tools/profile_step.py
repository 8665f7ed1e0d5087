"""torch.profiler timeline summary of one training step (GPU kernel time by kernel name, GPU busy vs wall)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nuscenes")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--cprofile", action="store_true", help="host-side cProfile of the step instead of the GPU timeline")
    a = ap.parse_args()
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.dp import GradSync
    from fusiontransformer_b200.spvcnn import Net3DSeg
    from fusiontransformer_b200.synthetic import make_scan
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda", 0)
    B = 8
    scans = [make_scan(a.workload, i) for i in range(B)]
    H, W = scans[0]["image_size"]
    db = dataflow.to_device(dataflow.host_batch_from_scans(scans), dev)
    fmap = torch.randn(B, 96, H, W, device=dev)
    torch.manual_seed(1)
    net = Net3DSeg(fusion="middle").to(dev).train()
    sync = GradSync(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=5e-4, fused=True)

    def step():
        lidar, rc, bidx, labels, _, _ = dataflow.voxelize_batch(db)
        img = ft.nn.functional.lift(fmap, rc, bidx)
        out = net(lidar, img.detach())
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels)
        sync.zero_grad()
        loss.backward()
        sync.finish()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    import time
    if a.cprofile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        for _ in range(a.steps):
            step()
        pr.disable()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("host ms/step to enqueue (cProfile on): %.2f" % ((t1 - t0) / a.steps * 1e3))
        pstats.Stats(pr).sort_stats("tottime").print_stats(a.top)
        return
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    from torch.autograd import DeviceType
    agg = {}
    for e in prof.events():
        if e.device_type == DeviceType.CUDA:
            t = agg.setdefault(e.name, [0.0, 0])
            t[0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
            t[1] += 1
    rows = sorted(((v[0] / a.steps / 1e3, v[1] / a.steps, k) for k, v in agg.items()), reverse=True)
    tot = sum(r[0] for r in rows)
    print("wall ms/step (under profiler) %.2f   GPU kernel ms/step %.2f" % (wall, tot))
    for ms, cnt, key in rows[: a.top]:
        print("%8.3f ms  %6.1f x  %s" % (ms, cnt, key[:110]))
    print("kernel launches per step: %.0f" % sum(r[1] for r in rows))


if __name__ == "__main__":
    main()
