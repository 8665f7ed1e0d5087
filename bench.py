#!/usr/bin/env python
"""Headline benchmark: train scans/s of the FusionTransformer 3D-branch hot path on N B200s (BASELINE.json).

One "step" = one data-parallel training step of the middle-fusion 3D branch on a batch of synthetic scans:
  device-side voxelization + dedup (a1-a3) -> initial_voxelize / kernel maps / SPVCNN sparse-conv UNet (a5-a14)
  -> 2D->3D lift of a [B,96,H,W] image-feature map at the points' img_indices (a15) -> CE loss -> backward
  (dgrad + wgrad of all 49 sparse convs) -> NCCL gradient all-reduce (N>1) -> Adam.
`value`  : scans/s with the raw scans already resident in HBM.
`e2e`    : the same through the public API with the batch in pinned HOST memory (H2D inside the timed region)
           and the loss read back every step.
Default workload: configs[1] (nuScenes-shaped) at every N; the line also carries `configs2_kitti` = configs[2]
(KITTI-shaped, BASELINE.json's data-parallel config) measured at the same N by a child run of this file.
`--impl reference` times the CPU oracle (the reference's algorithm restated on PyTorch-CPU; torchsparse v1.1.0 is
not installable offline) on the box's host cores, one scan per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs[1], configs[2], configs[4]
    "nuscenes": dict(shape="nuscenes", batch=8, desc="configs[1]: nuScenes-shaped scans (~7k points, 1600x900 image), train step, batch 8 per GPU"),
    "kitti": dict(shape="kitti", batch=8, desc="configs[2]: SemanticKITTI-shaped scans (~20k front-camera points, 1226x370 image), train step, batch 8 per GPU"),
    "stress": dict(shape="stress", batch=8, desc="configs[4]: dense-scan stress (120k points/scan, 1226x370 feature map), train step, batch 8 per GPU"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for nm, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def oracle_step_factory(shape: str, fusion: str = "middle"):
    """One-scan training step of the CPU oracle (fwd + bwd + Adam), same stages as the GPU step."""
    from fusiontransformer_b200.synthetic import make_scan
    from oracle import ft_glue as og, ts_ops as ts
    torch.manual_seed(1)
    net = og.Net3DSeg(num_classes=20, dual_head=False, fusion=fusion).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=5e-4)
    scans = [make_scan(shape, i) for i in range(2)]
    H, W = scans[0]["image_size"]
    fmap = torch.randn(1, 96, H, W)

    def step(i):
        s = scans[i % len(scans)]
        vc, keep, inds, inv = og.voxelize_scan(s["points"])
        st = og.collate([dict(coords=vc[inds], feats=s["feats"][keep][inds])])
        img = og.lift(fmap, [s["points_img"][keep][inds]])
        labels = torch.from_numpy(s["seg_labels"][keep][inds])
        opt.zero_grad()
        out = net(ts.SparseTensor(st.F, st.C), img.detach())
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels)
        loss.backward()
        opt.step()
        return loss.item()

    return step


def time_oracle(shape: str, steps: int, warmup: int, budget_s: float = 1e9):
    torch.set_num_threads(os.cpu_count() or 1)
    step = oracle_step_factory(shape)
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        step(warmup + i)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, dt / done, done


def run_reference(args, rank):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    steps = max(1, args.steps)
    sps, spstep, done = time_oracle(wl["shape"], steps, max(0, args.warmup), budget_s=240.0)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "train scans/sec", "value": sps, "unit": "scans/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": "1 scan per step (batch 1) of the same synthetic shape"},
        "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port",
                         "sample": "%d timed single-scan train steps (fwd+bwd+Adam) of the CPU oracle, torch threads=%d" % (done, cores)},
        "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_configs2_child(args, rank, world):
    """`bench.py --workload kitti` at the same N as a child of every rank (own rendezvous port); rank 0 returns the
    child's line condensed, the other ranks None.  A failure is reported in the key, never raised."""
    env = dict(os.environ)
    if world > 1:
        # torchrun's agent serves the parent's rendezvous store; the children make their own on a derived port
        env.pop("TORCHELASTIC_USE_AGENT_STORE", None)
        env["MASTER_ADDR"] = env.get("MASTER_ADDR", "127.0.0.1")
        base = int(env.get("MASTER_PORT", "29500"))
        env["MASTER_PORT"] = str(base + 1 if base < 65000 else base - 1)
    cmd = [sys.executable, os.path.abspath(__file__), "--gpus", str(args.gpus), "--workload", "kitti",
           "--steps", str(args.steps), "--warmup", str(args.warmup), "--no-cpu-baseline", "--no-roofline",
           "--fusion", args.fusion, "--fmap-format", args.fmap_format]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        if rank != 0:
            return None
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                return {"workload": d["config"]["workload"], "value": d["value"], "unit": d["unit"],
                        "ms_per_step": d["ms_per_step"], "e2e": d["e2e"], "n_gpus": d["n_gpus"],
                        "scans_per_gpu": d["config"]["scans_per_gpu"], "parallelism": d["config"]["parallelism"],
                        "cuda_graph": d["config"]["cuda_graph"], "clocks": d.get("clocks"),
                        "note": "same bench.py, same N, --workload kitti, run in a child process before this line's "
                                "own measurement: BASELINE.json configs[2], the data-parallel config"}
        return {"error": "rc %d: %s" % (r.returncode, (r.stderr or "")[-300:])}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:300]} if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: BASELINE.json configs[1] (nuscenes) at --gpus 1, configs[2] (kitti, the data-parallel "
                         "config) at --gpus > 1; 'stress' = configs[4]")
    ap.add_argument("--batch", type=int, default=0, help="scans per GPU per step (default: the workload's)")
    ap.add_argument("--fusion", default="middle", choices=["none", "middle", "early"])
    ap.add_argument("--fmap-format", default="channels_last", choices=["channels_last", "nchw"],
                    help="memory format of the synthetic image-feature map the lift gathers from")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-baseline", action="store_true",
                    help="with the default workload, skip the extra run of configs[2] (KITTI-shaped) at the same N that "
                         "is reported as `configs2_kitti`")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--prefetch-thread", dest="no_prefetch_thread", action="store_false",
                    help="build the next batch's geometry from a worker thread (GIL-bound: slower)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel of the step from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--reuse-plans", action="store_true", help="diagnostic only: time the compute with cached geometry")
    ap.add_argument("--trace", type=int, default=0, metavar="N",
                    help="after the timed runs, profile N more steps with torch.profiler and print the GPU kernel table")
    ap.add_argument("--trace-top", type=int, default=60)
    ap.add_argument("--trace-dump", default="", help="CSV of the last traced step, one row per GPU launch")
    ap.add_argument("--diag", action="store_true", help="print host-side enqueue time per phase (stderr)")
    ap.add_argument("--no-prefetch", dest="prefetch", action="store_false",
                    help="build each batch's geometry inside its own step (host reads stall the launch queue)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    default_workload = args.workload is None
    if args.workload is None:
        args.workload = "nuscenes"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # configs[2] (KITTI-shaped, the data-parallel config of BASELINE.json) at the SAME N, in a child process per rank,
    # before this process touches the GPU: every default line then carries both workloads -- `value` is configs[1] at
    # every N (one workload along the whole 1/2/4/8 curve), `configs2_kitti` is configs[2] at that N.
    configs2 = None
    if default_workload and not args.no_scaling_baseline:
        configs2 = run_configs2_child(args, rank, world)

    import torch.distributed as dist
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import _lib, conv_engine, dataflow
    from fusiontransformer_b200.dp import GradSync
    from fusiontransformer_b200.spvcnn import Net3DSeg
    from fusiontransformer_b200.synthetic import make_scan

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a mismatched collective must fail in minutes, not hold the box for NCCL's 10-minute default
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("FT3D_NCCL_TIMEOUT_S", "180"))))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    nbatches = 4                                              # distinct batches cycled so inputs are never L2-warm
    scans_all = [[make_scan(wl["shape"], (rank * nbatches + b) * B + i) for i in range(B)] for b in range(nbatches)]
    H, W = scans_all[0][0]["image_size"]
    host = [dataflow.host_batch_from_scans(s) for s in scans_all]
    resident = [dataflow.to_device(h, dev) for h in host]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    # stands in for the image branch's output [B,96,H,W] fp32 (image_models_billinear.py:111-116).  The producer is
    # free to choose the memory format: channels-last makes the lift read 384 contiguous bytes per point instead of
    # 96 sectors (the NCHW map ran the gather at 5 % of the HBM peak); --fmap-format nchw restores the reference layout.
    fmap = torch.randn(B, 96, H, W, device=dev, generator=g)
    if args.fmap_format == "channels_last":
        fmap = fmap.contiguous(memory_format=torch.channels_last)

    if conv_engine.mode() == "tc":
        # reduced-precision mode: the point-branch nn.Linear GEMMs (a14, cuBLAS) run on TF32 tensor cores as well
        torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(1)
    net = Net3DSeg(num_classes=20, dual_head=False, fusion=args.fusion).to(dev).train()
    sync = GradSync(net)
    sync.broadcast_parameters(net)
    use_graph = args.graph and conv_engine.mode() == "tc"
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=5e-4, fused=True, capturable=use_graph)

    # The geometry of batch i+1 (upload, voxelization, kernel maps: every host read of a data-dependent size) is
    # built on a high-priority side stream while batch i's convolutions run (fusiontransformer_b200/plan.py).
    from fusiontransformer_b200.plan import Prefetcher
    pre = Prefetcher(dev, threaded=not args.no_prefetch_thread,
                     priority=int(os.environ.get("FT3D_PREFETCH_PRIORITY", "-2"))) if args.prefetch else None

    diag = {} if args.diag else None

    def tick(name, t0):
        if diag is not None:
            diag[name] = diag.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    from fusiontransformer_b200.fused import weight_packer
    from fusiontransformer_b200.losses import seg_loss
    packer = weight_packer(net) if conv_engine.mode() == "tc" else None

    def fwd_bwd(plan):
        if packer is not None:
            packer.pack()                 # all 96 bf16 weight images (the optimizer moved the weights) in one launch
        ex = plan.extras
        img = ft.nn.functional.lift(fmap, ex["rc"], ex["bidx"]) if args.fusion != "none" else None
        out = net(ex["lidar"], None if img is None else img.detach(), plan=plan)
        loss = seg_loss(out["lidar_seg_logit"], ex["labels"])      # CE forward + gradient in one libft3d pass
        sync.zero_grad()
        loss.backward()
        return loss

    # The compute of a step (lift, forward, loss, backward, optimizer) is captured once per capacity set as ONE CUDA
    # graph and replayed on capacity-padded geometry (fusiontransformer_b200/graph.py).  With N > 1 the NCCL gradient
    # exchange and the optimizer run eagerly after the replay.
    gstep, in_graph = None, False
    if use_graph:
        from fusiontransformer_b200.fused import join_side_streams
        from fusiontransformer_b200.graph import GraphedStep
        # N > 1: the NCCL all-reduces are captured too (FT3D_DP_EXCHANGE=graph, default): each bucket is launched on the
        # communication stream from the gradient hooks as soon as its last wgrad has landed, overlapping the rest of the
        # backward chain, and the optimizer step stays inside the graph.  FT3D_DP_EXCHANGE=eager exchanges after the
        # replay instead (one exposed 87 MB all-reduce + an eager optimizer step per step).
        in_graph = world == 1 or os.environ.get("FT3D_DP_EXCHANGE", "graph") == "graph"
        if in_graph:
            def body(plan):
                loss = fwd_bwd(plan)
                sync.finish()
                opt.step()
                return loss
            gstep = GraphedStep(body, modules=[net])
        else:
            sync.deferred = True

            def body(plan):
                loss = fwd_bwd(plan)
                join_side_streams()
                return loss

            def exchange_and_step():
                sync.finish()
                opt.step()
            gstep = GraphedStep(body, modules=[net], after_replay=exchange_and_step)

    cached_plans = {}
    prepare = gstep.prepare if gstep is not None else dataflow.prepare_batch

    def train_step(batch, nxt=None, eager=False):
        t = time.perf_counter()
        if args.reuse_plans:          # diagnostic: geometry of each distinct batch built once (not a valid bench mode)
            plan = cached_plans.get(id(batch))
            if plan is None:
                plan = cached_plans[id(batch)] = dataflow.prepare_batch(batch, dev)
            nxt = None
        elif pre is None:
            plan = dataflow.prepare_batch(batch, dev)
        else:
            if pre._pending is None:
                pre.submit(dataflow.prepare_batch if eager else prepare, batch, dev)
            plan = pre.get()
        t = tick("get_plan", t)
        if gstep is not None and not eager:
            loss = gstep.step(plan)
            t = tick("graph_step", t)
        else:
            loss = fwd_bwd(plan)
            t = tick("forward_backward", t)
            sync.finish()
            opt.step()
            t = tick("optimizer", t)
        if pre is not None and nxt is not None:
            pre.submit(dataflow.prepare_batch if eager else prepare, nxt, dev)
        tick("prefetch_next", t)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_wall = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tw = time.perf_counter()
            fn(i)
            step_wall.append(1e3 * (time.perf_counter() - tw))
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    L = _lib.lib()
    for i in range(args.warmup):
        train_step(resident[i % nbatches], resident[(i + 1) % nbatches])
    # ---- device-resident timing (the `value`)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    L.calls.clear()
    replays0 = gstep.replays if gstep is not None else 0
    ms = timed(lambda i: train_step(resident[i % nbatches], resident[(i + 1) % nbatches]), args.steps)
    if pre is not None and pre._pending is not None:
        pre.get()
    launches = _lib.launch_count(L.calls)
    if gstep is not None:
        launches += gstep.launches_per_replay * (gstep.replays - replays0)
    if diag is not None:
        print("host enqueue ms/step (device-resident loop): " +
              ", ".join("%s %.2f" % (k, 1e3 * v / (args.steps + args.warmup)) for k, v in diag.items()), file=sys.stderr)
        print("host ms per step: " + " ".join("%.1f" % v for v in step_wall), file=sys.stderr)
        step_wall.clear()
        diag.clear()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end timing through the public API: pinned host batch -> H2D -> step -> loss to host
    def e2e_step(i):
        loss = train_step(host[i % nbatches], host[(i + 1) % nbatches])     # pinned host batches: H2D inside the step
        return loss.item()
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    if pre is not None and pre._pending is not None:
        pre.get()
    # sampled across BOTH timed regions (device-resident and end-to-end; 100 ms period): a 20-step region lasts ~0.13 s
    clocks = sampler.stop() if sampler else None
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    if diag is not None:
        print("host enqueue ms/step (e2e loop): " +
              ", ".join("%s %.2f" % (k, 1e3 * v / (args.steps + 2)) for k, v in diag.items()), file=sys.stderr)
    h2d = int(np.mean([h.nbytes() for h in host]))

    # ---- per-entry-point device times + conv work log (separate pass, not part of the reported value)
    roofline, shares = None, None
    if not args.no_roofline:             # every rank runs the pass (the steps contain collectives); rank 0 reports
        pk = peaks()
        L.profile, conv_engine.WORK_LOG = [], []
        # idempotent launches (outputs rewritten with the same values; conv_os's repeats also re-apply the BatchNorm
        # running-statistics momentum update, which training-mode steps never read)
        L.profile_repeat = {"conv_pairs_tc": 4, "conv_os": 4}
        nprof = min(args.steps, 5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(nprof):       # launched kernel by kernel (no graph replay) so that each entry point can be timed
            train_step(resident[i % nbatches], resident[(i + 1) % nbatches], eager=True)
        e1.record()
        torch.cuda.synchronize()
        prof, work = L.profile, conv_engine.WORK_LOG
        L.profile, conv_engine.WORK_LOG, L.profile_repeat = None, None, {}
        tot = {}
        for name, a, b, reps in prof:
            t = tot.setdefault(name, [0.0, 0])
            t[0] += a.elapsed_time(b) / reps
            t[1] += 1
        step_ms = e0.elapsed_time(e1) / nprof
        kernel_only = tot.pop("conv_os_kernel", None)     # conv_os_kernel alone (extra launches, not part of the step)
        shares = {k: round(v[0] / nprof, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])}
        shares["_step_ms_profiled"] = round(step_ms, 3)
        shares["_libft3d_ms"] = round(sum(v[0] for v in tot.values()) / nprof, 3)
        # The dominant kernel is the pair-major tcgen05 GEMM (conv_pairs_tc): forward, dgrad, transposed and k=1
        # convolutions all run through it.  Algorithmic work of one launch (DESIGN.md "Roofline accounting"):
        #   flops = 2 L red ncols ;  bytes = 2 (rows_in red + K red ncols) + 4 rows_out ncols + 8 L   (SURVEY 8(d), bf16
        #   operands, fp32 result).  Both bounds are evaluated per launch; the binding one is reported.
        dom = "conv_os" if "conv_os" in tot else "conv_pairs_tc"
        convs = [w for w in work if w["kind"] == dom]
        if dom in tot and convs:
            t_s = tot[dom][0] * 1e-3
            fl = sum(2.0 * w["pairs"] * w["red"] * w["ncols"] for w in convs)
            by = sum(2.0 * (w["rows_in"] * w["red"] + w["K"] * w["red"] * w["ncols"]) + 4.0 * w["rows"] * w["ncols"]
                     + 8.0 * w["pairs"] for w in convs)
            t_tensor, t_hbm = fl / (pk["tf_sus"] * 1e12), by / (pk["hbm"] * 1e9)
            bound = "tensor" if t_tensor >= t_hbm else "hbm"
            ach, peak, unit = (fl / t_s / 1e12, pk["tf_sus"], "TFLOP/s") if bound == "tensor" else \
                              (by / t_s / 1e9, pk["hbm"], "GB/s")
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % dom)
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            t_k = t_s
            if dom == "conv_os" and kernel_only is not None:
                # `achieved` is quoted on conv_os_kernel alone (the kernel the ncu capture shows); the whole entry point
                # (+ the fold of split tiles and the statistics finalize, two tiny dependent launches) is avg_call_us
                t_k = kernel_only[0] * 1e-3
                ach = ach * t_s / t_k
            # launches differ in which bound binds them (32-channel layers: HBM; 256-channel layers: tensor): the time
            # the roofline allows for the whole set is the sum of the per-launch maxima
            t_roof = sum(max(2.0 * w["pairs"] * w["red"] * w["ncols"] / (pk["tf_sus"] * 1e12),
                             (2.0 * (w["rows_in"] * w["red"] + w["K"] * w["red"] * w["ncols"]) + 4.0 * w["rows"] * w["ncols"]
                              + 8.0 * w["pairs"]) / (pk["hbm"] * 1e9)) for w in convs)
            roofline = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                        "traffic": traffic,
                        "peak_source": ("MEASURED_PEAKS.json (hbm_gbs / bf16_tflops_sustained)" if pk["src"] == "measured"
                                        else "fallback of B200_PROFILING.md"),
                        "launches_per_step": tot[dom][1] / nprof,
                        "avg_launch_us": 1e6 * t_k / tot[dom][1], "avg_call_us": 1e3 * tot[dom][0] / tot[dom][1],
                        "frac_per_launch_bound": t_roof / t_k,
                        "algorithmic_mb_per_launch": by / len(convs) / 1e6,
                        "algorithmic_gflop_per_launch": fl / len(convs) / 1e9,
                        "tflops": fl / t_k / 1e12, "gbs": by / t_k / 1e9,
                        "share_of_step": tot[dom][0] / nprof / (ms / args.steps),
                        "share_of_kernel_time": tot[dom][0] / max(sum(v[0] for v in tot.values()), 1e-9),
                        "note": "time = CUDA events on the launching stream around every " + dom + " call (4 back-to-back "
                                "launches of it between the two events, / 4: a single ~20 us launch would carry ~8 us "
                                "of event overhead), summed over %d separately profiled steps launched kernel by kernel. "
                                "avg_call_us = the whole entry point (conv_os_kernel + fold of split tiles + statistics "
                                "finalize); avg_launch_us / achieved / tflops / gbs = conv_os_kernel alone (the same call "
                                "repeated with the two small launches disabled). frac = achieved / peak for the bound "
                                "that binds the SUM of the launches; frac_per_launch_bound = sum over launches of "
                                "max(flops / tensor peak, bytes / hbm peak) / measured time. share_of_step = entry-point "
                                "time per step / the graph-replayed ms_per_step (kernels of other streams overlap it); "
                                "share_of_kernel_time = / the sum over all libft3d entry points" % nprof}

    if args.trace and world == 1:
        # torch.profiler (CUPTI) timeline of N more steps: per-kernel table, per-stream busy time, idle gaps
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(args.trace):
                train_step(resident[i % nbatches], resident[(i + 1) % nbatches])
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.trace * 1e3
        if pre is not None and pre._pending is not None:
            pre.get()
        tf = tempfile.NamedTemporaryFile(suffix=".json", delete=False).name
        prof.export_chrome_trace(tf)
        evs = [e for e in json.load(open(tf))["traceEvents"]
               if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
        os.unlink(tf)
        agg, streams = {}, {}
        for e in evs:
            a = agg.setdefault(e["name"], [0.0, 0])
            a[0] += e["dur"]
            a[1] += 1
            streams.setdefault(e["args"].get("stream"), []).append((e["ts"], e["ts"] + e["dur"], e["name"]))
        rows = sorted(((v[0] / args.trace / 1e3, v[1] / args.trace, k) for k, v in agg.items()), reverse=True)
        iv = sorted((a, b) for v in streams.values() for a, b, _ in v)
        busy, cur_a, cur_b = 0.0, None, None
        for a, b in iv:
            if cur_b is None or a > cur_b:
                if cur_b is not None:
                    busy += cur_b - cur_a
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        if cur_b is not None:
            busy += cur_b - cur_a
        span = (iv[-1][1] - iv[0][0]) if iv else 0.0
        print("trace: wall %.2f ms/step under profiler, GPU span %.2f ms/step, any-stream busy %.2f ms/step, "
              "sum of kernel time %.2f ms/step, %d launches/step"
              % (wall, span / args.trace / 1e3, busy / args.trace / 1e3, sum(r[0] for r in rows), sum(r[1] for r in rows)),
              file=sys.stderr)
        for sid, v in sorted(streams.items(), key=lambda kv: -sum(b - a for a, b, _ in kv[1])):
            v.sort()
            tot = sum(b - a for a, b, _ in v)
            gaps = sorted(((v[i + 1][0] - v[i][1], v[i][2][:50], v[i + 1][2][:50]) for i in range(len(v) - 1)), reverse=True)
            small = sum(g for g, _, _ in gaps if 0 < g < 50.0)
            print("  stream %s: %d launches/step, busy %.2f ms/step, gaps<50us sum %.2f ms/step" %
                  (sid, len(v) / args.trace, tot / args.trace / 1e3, small / args.trace / 1e3), file=sys.stderr)
            for g, n0, n1 in gaps[:6]:
                print("      gap %8.1f us after %s before %s" % (g, n0, n1), file=sys.stderr)
        if args.trace_dump:
            # the last traced step, launch by launch (all streams): analysed offline (tools/analyze_trace.py)
            cut = iv[0][0] + span * (args.trace - 1) / args.trace
            with open(args.trace_dump, "w") as f:
                f.write("ts_us,dur_us,stream,grid,block,smem,regs,name\n")
                for e in sorted(evs, key=lambda e: e["ts"]):
                    if e["ts"] < cut:
                        continue
                    a_ = e["args"]
                    f.write("%.3f,%.3f,%s,%s,%s,%s,%s,\"%s\"\n" % (
                        e["ts"] - cut, e["dur"], a_.get("stream"), "x".join(map(str, a_.get("grid", []))),
                        "x".join(map(str, a_.get("block", []))), a_.get("shared memory", ""),
                        a_.get("registers per thread", ""), e["name"][:90].replace('"', "'")))
        for ms_, cnt, key in rows[:args.trace_top]:
            print("  %8.3f ms %6.1f x %7.1f us  %s" % (ms_, cnt, 1e3 * ms_ / max(cnt, 1e-9), key[:110]), file=sys.stderr)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, spstep, done = time_oracle(wl["shape"], steps=24, warmup=1, budget_s=15.0)
        cpu_baseline = {"value": sps, "unit": "scans/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": "%d single-scan train steps (fwd+bwd+Adam) of the CPU oracle after 1 warm-up, %.1f s/step"
                                  % (done, spstep)}

    if rank == 0:
        nvox = int(np.mean([len(np.concatenate([s["points"] for s in sc])) for sc in scans_all]))
        line = {
            "metric": "train scans/sec", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if conv_engine.mode() == "tc" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "scans_per_gpu": B, "points_per_batch": nvox, "fusion": args.fusion,
                       "image_hw": [H, W], "feature_map_format": args.fmap_format,
                       "parallelism": "dp%d" % world, "optimizer": "Adam(lr 1e-4, wd 5e-4)",
                       "geometry_prefetch": bool(args.prefetch) and not args.reuse_plans,
                       **({"INVALID_diagnostic": "geometry cached across steps"} if args.reuse_plans else {}),
                       "cuda_graph": ("whole step%s, %d capture(s)" % (
                           "" if world == 1 else (" incl. NCCL exchange" if in_graph else ", exchange after replay"),
                           gstep.captures)) if gstep is not None else "off",
                       "linear_layers": ("libft3d tcgen05 (bf16 operands); 96->20 head on cuBLAS" if conv_engine.mode() == "tc"
                                         else "cuBLAS fp32"),
                       "conv_algo": os.environ.get("FT3D_CONV_ALGO", "os"),
                       "wgrad": "two-stage deterministic" if __import__("fusiontransformer_b200.ops", fromlist=["x"]).wgrad_deterministic() else "persistent, fp32 atomics flush",
                       "l2": "no explicit flush: %d distinct batches are cycled and the step's working set (348 MB of "
                             "weights+Adam state, the activations and the %.1f GB feature map) exceeds the 126 MB L2"
                             % (nbatches, B * 96 * H * W * 4 / 1e9)},
            "e2e": {"value": e2e_value, "unit": "scans/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernel_ms_per_step": shares, "configs2_kitti": configs2,
        }
        print(json.dumps(line), flush=True)
    if pre is not None:
        pre.close()
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if os.environ.get("FT3D_NCCL_TEARDOWN", "clean") == "exit":
            os._exit(0)        # escape hatch only: skip the process-group teardown
        # NCCL work captured in CUDA graphs: the graphs (and the static buffers they reference) are released BEFORE the
        # process group is destroyed -- destroying it first left ProcessGroupNCCL's teardown waiting on the captured work.
        if gstep is not None:
            gstep.graph = None
            gstep.static = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
