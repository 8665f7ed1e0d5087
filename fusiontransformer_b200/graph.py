"""One CUDA graph per training step: static-capacity geometry + replay.

Every data-dependent size of the 3D branch (voxels per stride, pairs per kernel map) is known on the host once the
GeometryPlan of a batch exists (plan.py, built one step ahead on a side stream).  The compute that follows --
lift, 49 fused conv blocks, point<->voxel gathers, loss, backward (dgrad on the main stream, wgrad on a forked
branch), optimizer -- is ~1000 small launches whose cost on the host (~15 us each through Python) exceeds their cost
on the B200.  ``GraphedStep`` therefore

  * copies each batch's plan into ``StaticGeometry``: the same tensors padded to a fixed capacity at fixed addresses.
    Padding rows are inert by construction: map rows are -1 (no pairs, no neighbours), point->voxel indices -1,
    weights/counts 0, labels -100, features 0 -- so every gather/scatter/GEMM kernel runs unchanged on the padded
    shapes and contributes exact zeros for them; the BatchNorm kernels are the only ones that need the real row
    count (the divisor of the statistics, and to keep padding rows at exactly zero after the affine shift), which
    they read from a device int32 looked up by the padded row dimension (ops.ROW_COUNTS);
  * captures the whole step once per capacity set and replays it for every later batch that fits (capacities carry
    ~10 % slack; a batch that does not fit triggers one eager step and a re-capture).

Results equal the eager exact-shape step (tests/test_gpu_graph.py): padding adds only zero terms.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .functional import KernelMap
from .plan import GeometryPlan
from .sparse_tensor import SparseTensor

__all__ = ["StaticGeometry", "GraphedStep"]


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _Slot:
    """One capacity-padded tensor: ``buf[:n]`` holds the batch's rows, ``buf[n:]`` the padding value."""
    __slots__ = ("buf", "pad", "used")

    def __init__(self, cap_rows, tail_shape, dtype, pad, device):
        self.buf = torch.full((cap_rows,) + tuple(tail_shape), pad, dtype=dtype, device=device)
        self.pad, self.used = pad, 0

    def load(self, src: torch.Tensor, copies, fills):
        """Queue ``buf[:n] <- src`` (and the re-padding of rows a larger previous batch left behind) on the batched
        copy / fill lists; StaticGeometry.load issues them as a handful of multi-tensor launches."""
        n = src.shape[0]
        copies.setdefault(self.buf.dtype, ([], []))
        dst, srcs = copies[self.buf.dtype]
        dst.append(self.buf[:n])
        srcs.append(src.to(self.buf.dtype) if src.dtype != self.buf.dtype else src)
        if n < self.used:
            fills.setdefault((self.buf.dtype, self.pad), []).append(self.buf[n:self.used])
        self.used = n


def capacity(n: int, slack: float, granule: int, floor: int = 0, taken: set | None = None) -> int:
    """Row capacity for ``n`` real rows: ``slack`` head-room, rounded up to ``granule``, never below ``floor`` (the
    capacity of the geometry being replaced: capacities only grow, so a stream of batches whose sizes wander converges
    instead of re-capturing for ever), and distinct from every value in ``taken`` (capacities double as the keys of
    ops.ROW_COUNTS, which maps a padded row count to the device scalar holding the real one)."""
    c = _round_up(max(int(n * slack) + 1, floor), granule)
    if taken is not None:
        while c in taken:
            c += granule
        taken.add(c)
    return c


class StaticGeometry:
    """Capacity-padded, fixed-address image of (GeometryPlan + the batch's voxelized inputs)."""

    def __init__(self, plan: GeometryPlan, slack: float = 1.10, granule: int = 128, prev: "StaticGeometry" = None):
        """``prev``: the geometry this one replaces -- no capacity shrinks below its value, so a stream of batches whose
        sizes wander (one larger in points, the next larger at stride 8 ...) converges after a few re-captures instead
        of re-capturing for ever."""
        dev = plan.point_coords.device
        self.device = dev
        used_caps = set()

        def cap_rows(n, floor=0):
            return capacity(n, slack, granule, floor, used_caps)

        from . import conv_engine
        self.full_tables = conv_engine.mode() == "f32"   # exact-precision mode gathers through nbr / nbrT everywhere
        ex = plan.extras
        self.strides = sorted(plan.coord_maps)
        self.n_points_cap = cap_rows(plan.point_coords.shape[0], prev.n_points_cap if prev is not None else 0)
        self.row_caps = {s: cap_rows(plan.coord_maps[s].shape[0],
                                     prev.row_caps.get(s, 0) if prev is not None else 0) for s in self.strides}
        self.slots = {}

        def slot(name, cap, tail, dtype, pad):
            self.slots[name] = _Slot(cap, tail, dtype, pad, dev)
            return self.slots[name].buf

        P = self.n_points_cap
        self.point_coords = slot("point_coords", P, (4,), torch.float32, 0)
        self.idx_query = slot("idx_query", P, (), torch.int32, -1)
        self.counts = slot("counts", self.row_caps[self.strides[0]], (), torch.int32, 0)
        self.feats = slot("feats", P, (ex["lidar"].F.shape[1],), torch.float32, 0)
        self.coords_in = slot("coords_in", P, (4,), torch.int32, 0)
        self.rc = slot("rc", P, (2,), torch.int32, 0)
        self.bidx = slot("bidx", P, (), torch.int32, 0)
        self.labels = slot("labels", P, (), torch.int64, -100)
        self.coord_maps = {s: slot("C%d" % s, self.row_caps[s], (4,), torch.int32, 0) for s in self.strides}
        self.kernel_maps, self.pair_caps, self.pass_caps = {}, {}, {}
        self.use_os = conv_engine.mode() == "tc" and conv_engine.os_enabled()
        for key, km in plan.kernel_maps.items():
            s_in, s_out = self._map_strides(key)
            kpad = km.nbr.shape[1]
            lcap = _round_up(max(int(km.num_pairs() * slack) + 1,
                                 prev.pair_caps.get(key, 0) if prev is not None else 0), 128)
            self.pair_caps[key] = lcap
            want_nbr = self.full_tables or key == self._stem_key()
            nbr = slot(key + ".nbr", self.row_caps[s_out], (kpad,), torch.int32, -1) if want_nbr else None
            skm = KernelMap(nbr, km.K, self.row_caps[s_in], self.row_caps[s_out], km.symmetric)
            if self.full_tables:
                skm._nbrT = slot(key + ".nbrT", self.row_caps[s_in], (kpad,), torch.int32, -1)
            skm._pairs = slot(key + ".pairs", lcap, (2,), torch.int32, 0)
            skm._offsets = slot(key + ".offsets", km.K + 1, (), torch.int32, 0)
            if self.use_os:
                # tile schedules of the output-stationary convolution, padded with empty tiles / passes
                skm.os_cluster = getattr(km, "os_cluster", 1)
                for (side, trows), op in km._os.items():
                    rows = self.row_caps[s_out] if side == "out" else self.row_caps[s_in]
                    tcap = (rows + trows - 1) // trows
                    name = "%s.os_%s%d" % (key, side, trows)
                    n_pass, n_unit, n_slot = op.host_counts()[:3]
                    old = prev.pass_caps.get(name, (0, 0, 0)) if prev is not None else (0, 0, 0)
                    # the unit array holds whole rounds of one unit per CTA cluster (os_plan.cu: placement), so its
                    # length moves in steps of ncl: one spare round on top of the head-room
                    ncl = 148 // (trows // 128)
                    pcap, ucap, scap = (max(int(n_pass * 1.25) + 8, old[0]),
                                        max((-(-int(n_unit * 1.25) // ncl) + 1) * ncl, tcap, old[1]),
                                        max(int(n_slot * 1.5) + 8, old[2]))
                    self.pass_caps[name] = (pcap, ucap, scap)
                    # `num` carries the batch's real unit count: the kernel never walks the padding units
                    skm._os[(side, trows)] = ops.OsPlan(slot(name + ".units", ucap, (8,), torch.int32, 0),
                                                        slot(name + ".split", tcap, (4,), torch.int32, 0),
                                                        slot(name + ".num", 8, (), torch.int32, 0),
                                                        slot(name + ".out_row", tcap * trows, (), torch.int32, -1),
                                                        slot(name + ".pass_k", pcap, (), torch.int32, 0),
                                                        slot(name + ".pass_idx", pcap, (trows,), torch.int32, -1),
                                                        rows, km.K, counts=(pcap, ucap, scap, 0, 0), tile_rows=trows)
            else:
                skm._ppos = slot(key + ".ppos", self.row_caps[s_out], (kpad,), torch.int32, -1)
                skm._pposT = slot(key + ".pposT", self.row_caps[s_in], (kpad,), torch.int32, -1)
            skm._num_pairs = lcap                        # sizes the grids and the partial-row buffers
            self.kernel_maps[key] = skm
        self.p2v = {s: (slot("p2v%d.idx" % s, P, (), torch.int32, -1),
                        slot("p2v%d.cnt" % s, self.row_caps[s], (), torch.int32, 0)) for s in plan.p2v}
        self.v2p = {s: (slot("v2p%d.idx" % s, P, (8,), torch.int32, -1),
                        slot("v2p%d.w" % s, P, (8,), torch.float32, 0)) for s in plan.v2p}
        # real row counts: [points, stride_1, stride_2, ...] on the device, staged through pinned host memory
        self._count_keys = ["points"] + self.strides
        self.counts_dev = torch.zeros(len(self._count_keys), dtype=torch.int32, device=dev)
        # ring of pinned staging rows: an async H2D copy may still be pending when the next batch is loaded
        self._counts_host = torch.zeros((8, len(self._count_keys)), dtype=torch.int32).pin_memory()
        self._ring = 0
        self.row_counts = {self.n_points_cap: self.counts_dev[0:1]}
        for i, s in enumerate(self.strides):
            self.row_counts[self.row_caps[s]] = self.counts_dev[i + 1:i + 2]

    def _stem_key(self):
        return "k3_os%d_s1_d1" % self.strides[0]

    @staticmethod
    def _map_strides(key):
        # "k{ks}_os{s}_s{stride}_d1": input stride s, output stride s*stride
        parts = key.split("_")
        s, st = int(parts[1][2:]), int(parts[2][1:])
        return s, s * st

    # ------------------------------------------------------------------ per-batch
    def fits(self, plan: GeometryPlan) -> bool:
        if plan.point_coords.shape[0] > self.n_points_cap or set(plan.kernel_maps) != set(self.kernel_maps):
            return False
        if any(plan.coord_maps[s].shape[0] > self.row_caps[s] for s in self.strides):
            return False
        for key, km in plan.kernel_maps.items():
            if km.num_pairs() > self.pair_caps[key]:
                return False
            if self.use_os:
                if set(km._os) != set(self.kernel_maps[key]._os):
                    return False
                for (side, trows), op in km._os.items():
                    n_pass, n_unit, n_slot = op.host_counts()[:3]
                    pcap, ucap, scap = self.pass_caps["%s.os_%s%d" % (key, side, trows)]
                    if n_pass > pcap or n_unit > ucap or n_slot > scap:
                        return False
        return True

    def load(self, plan: GeometryPlan):
        """Copy one batch's geometry and inputs into the static buffers (enqueued on the current stream)."""
        ex, S = plan.extras, self.slots
        copies, fills = {}, {}
        S["point_coords"].load(plan.point_coords, copies, fills)
        S["idx_query"].load(plan.idx_query, copies, fills)
        S["counts"].load(plan.counts, copies, fills)
        S["feats"].load(ex["lidar"].F, copies, fills)
        S["coords_in"].load(ex["lidar"].C, copies, fills)
        S["rc"].load(ex["rc"], copies, fills)
        S["bidx"].load(ex["bidx"], copies, fills)
        S["labels"].load(ex["labels"], copies, fills)
        for s in self.strides:
            S["C%d" % s].load(plan.coord_maps[s], copies, fills)
        for key, km in plan.kernel_maps.items():
            L = km.num_pairs()
            if key + ".nbr" in S:
                S[key + ".nbr"].load(km.nbr, copies, fills)
            if key + ".nbrT" in S:
                S[key + ".nbrT"].load(km.nbrT, copies, fills)
            S[key + ".pairs"].load(km.pairs_padded[:L], copies, fills)
            S[key + ".offsets"].load(km.pair_offsets, copies, fills)
            if self.use_os:
                for (side, trows), op in km._os.items():
                    name = "%s.os_%s%d" % (key, side, trows)
                    n_pass, n_unit = op.host_counts()[:2]
                    S[name + ".units"].load(op.units[:n_unit], copies, fills)
                    S[name + ".split"].load(op.split_tiles[:max(op.host_counts()[4], 1)], copies, fills)
                    S[name + ".num"].load(op.num, copies, fills)
                    S[name + ".out_row"].load(op.out_row, copies, fills)
                    S[name + ".pass_k"].load(op.pass_k[:n_pass], copies, fills)
                    S[name + ".pass_idx"].load(op.pass_idx[:n_pass], copies, fills)
            else:
                S[key + ".ppos"].load(km.ppos, copies, fills)
                S[key + ".pposT"].load(km.pposT, copies, fills)
        for s, (idx, cnt) in plan.p2v.items():
            S["p2v%d.idx" % s].load(idx, copies, fills)
            S["p2v%d.cnt" % s].load(cnt, copies, fills)
        for s, (idx, w) in plan.v2p.items():
            S["v2p%d.idx" % s].load(idx, copies, fills)
            S["v2p%d.w" % s].load(w, copies, fills)
        for dst, srcs in copies.values():
            torch._foreach_copy_(dst, srcs, non_blocking=True)
        for (dtype, pad), views in fills.items():
            torch._foreach_zero_(views)
            if pad != 0:
                torch._foreach_add_(views, pad)
        row = self._counts_host[self._ring]
        self._ring = (self._ring + 1) % self._counts_host.shape[0]
        row[0] = plan.point_coords.shape[0]
        for i, s in enumerate(self.strides):
            row[i + 1] = plan.coord_maps[s].shape[0]
        self.counts_dev.copy_(row, non_blocking=True)

    def as_plan(self) -> GeometryPlan:
        """The static buffers dressed as a GeometryPlan (what SPVCNN.backbone(plan=...) consumes)."""
        p = GeometryPlan(point_coords=self.point_coords, sparse_hash=None, idx_query=self.idx_query, counts=self.counts)
        p.coord_maps, p.kernel_maps, p.tables = dict(self.coord_maps), dict(self.kernel_maps), {}
        p.p2v, p.v2p = dict(self.p2v), dict(self.v2p)
        lidar = SparseTensor(self.feats, self.coords_in)      # no back-reference to p: a cycle would park every
                                                                # batch's tensors until the cyclic GC runs
        p.extras.update(lidar=lidar, rc=self.rc, bidx=self.bidx, labels=self.labels)
        return p


class GraphedStep:
    """``step(plan) -> loss`` (a static device scalar): replays the captured training step on the batch described by
    ``plan`` (from dataflow.prepare_batch).  ``body(static_plan) -> loss`` is the eager step written against a plan
    (forward, loss, zero_grad, backward, gradient sync, optimizer); it is captured as is.

    The optimizer must be capture-safe (``torch.optim.Adam(..., fused=True, capturable=True)``).  With
    ``world_size > 1`` pass a ``body`` that stops after backward and do the gradient exchange + optimizer eagerly in
    ``after_replay``."""

    def __init__(self, body, modules=(), after_replay=None, slack: float = 1.10):
        self.body, self.after_replay, self.slack = body, after_replay, slack
        self.static: StaticGeometry | None = None
        self.graph = None
        self.loss = None
        self.captures = 0
        self.replays = 0
        self.launches_per_replay = 0
        self._bns = [m for mod in modules for m in mod.modules()
                     if isinstance(m, torch.nn.BatchNorm1d) and m.track_running_stats]
        self.done = None                    # event: the last replay has finished reading the static buffers
        # warm-up and capture share one stream: autograd's AccumulateGrad nodes remember the stream they were created
        # on, and a mismatch during capture would insert an uncapturable cross-stream wait
        # priorities: geometry prefetch (-2) > this stream's forward/dgrad chain (-1) > side-stream wgrad (0), so the
        # short kernels of the dependent chain are scheduled as soon as a wgrad CTA retires instead of behind its grid
        self.stream = torch.cuda.Stream(priority=int(os.environ.get("FT3D_MAIN_PRIORITY", "-1")))

    def _capture(self, plan: GeometryPlan):
        from . import _lib
        self.graph = None                                   # release the previous private pool first
        prev = None
        if self.static is not None:                          # keep only the numbers: the old buffers are freed first
            import types
            prev = types.SimpleNamespace(n_points_cap=self.static.n_points_cap, row_caps=dict(self.static.row_caps),
                                         pair_caps=dict(self.static.pair_caps), pass_caps=dict(self.static.pass_caps))
        self.static = None
        self.static = StaticGeometry(plan, self.slack, prev=prev)
        self.static.load(plan)
        splan = self.static.as_plan()
        prev_counts = ops.ROW_COUNTS
        ops.ROW_COUNTS = self.static.row_counts
        try:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                loss = self.body(splan)                     # eager, padded: this IS the step for this batch
            cur.wait_stream(self.stream)
            if self.after_replay is not None:
                self.after_replay()
            torch.cuda.synchronize()
            pend = [getattr(b, "_ft3d_pending_batches", 0) for b in self._bns]
            calls0 = _lib.launch_count(_lib.lib().calls)
            g = torch.cuda.CUDAGraph()
            ops.FORCE_REPACK = True                         # weight images must be re-packed inside every replay
            ops.PACK_EPOCH += 1
            try:
                with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):   # NCCL's watchdog
                    # thread polls events while we capture; only this thread's calls must be capture-safe
                    self.loss = self.body(splan)
            finally:
                ops.FORCE_REPACK = False
            self.launches_per_replay = _lib.launch_count(_lib.lib().calls) - calls0
            for b, n in zip(self._bns, pend):               # the capture pass executed nothing
                b._ft3d_pending_batches = n
            self.graph = g
            self.captures += 1
        finally:
            ops.ROW_COUNTS = prev_counts
        return loss

    class Loaded:
        """Token returned by ``prepare``: the batch already sits in the static buffers.  ``inverse`` / ``kept`` are the
        host-side maps of that batch (predictions of the unique voxels -> the original points, reference
        ``data/utils/validate.py:10-11``; the bounds mask of the dataloader) -- exact-size tensors that never enter
        the graph."""

        def __init__(self, inverse=None, kept=None):
            self.inverse, self.kept = inverse, kept

    LOADED = Loaded()      # a token without maps (kept for callers that compare by identity)

    def prepare(self, batch, device="cuda"):
        """For plan.Prefetcher: upload + voxelize + build the exact geometry of ``batch`` and, when it fits the captured
        capacities, copy it into the static buffers -- all on the prefetch stream, so the exact-size tensors never
        touch the compute stream (no record_stream => the allocator can recycle them without cudaMalloc, which would
        synchronise the device behind the running graph).  The copy waits for the previous replay to finish reading."""
        from .dataflow import prepare_batch
        plan = prepare_batch(batch, device)
        if self.static is not None and self.static.fits(plan):
            if self.done is not None:
                torch.cuda.current_stream().wait_event(self.done)
            self.static.load(plan)
            return GraphedStep.Loaded(plan.extras.get("inverse"), plan.extras.get("kept"))
        return plan

    def step(self, plan):
        """``plan``: a GeometryPlan, or the token returned by ``prepare``."""
        if not isinstance(plan, GraphedStep.Loaded):
            if self.static is None or not self.static.fits(plan):
                loss = self._capture(plan)
                self._mark_done()
                return loss
            if self.done is not None:
                torch.cuda.current_stream().wait_event(self.done)
            self.static.load(plan)
        self.graph.replay()
        self.replays += 1
        ops.REPLAY_EPOCH += 1        # the replayed optimizer step moved the weights: cached bf16 images are stale
        for b in self._bns:
            b._ft3d_pending_batches = getattr(b, "_ft3d_pending_batches", 0) + 1
        if self.after_replay is not None:
            self.after_replay()
        self._mark_done()
        return self.loss

    def _mark_done(self):
        self.done = torch.cuda.Event()
        self.done.record()
