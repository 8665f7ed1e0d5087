"""Geometry plan of one SPVCNN forward pass, built one step ahead on a side stream.

Everything integer in the 3D branch depends only on the voxel coordinates of the batch: the sorted-unique
re-voxelization of ``initial_voxelize`` (models/utils.py:15-35), the coordinate sets and hash tables of the five
strides, the nine kernel maps (k3 at strides 1..16, k2s2 between them; models/spvcnn.py:99-124) with their pair
lists and pair-position tables, and the point<->voxel maps at strides 1, 16 and 4 (models/utils.py:40-106).
The reference builds them lazily inside the forward pass, reading each data-dependent size back to the host
(``counts.cpu()`` in torchsparse conv3d; ``torch.unique``), which stalls the launch queue ~15 times per step.

``build_plan`` runs exactly the same builders (same bit-exact results: tests/test_gpu_plan.py) but ahead of time;
``Prefetcher`` runs it for batch i+1 on a high-priority side stream while batch i's convolutions execute, so the
host reads overlap compute and the main stream never synchronises.  ``SPVCNN.backbone(x, plan=...)`` then finds
every map in the caches the reference code already consults (``coord_maps``, ``kernel_maps``,
``z.idx_query/weights/additional_features``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import conv_engine, ops
from .functional import KernelMap, build_kernel_map, spdownsample
from .ops import CoordTable

__all__ = ["GeometryPlan", "build_plan", "Prefetcher"]


@dataclass
class GeometryPlan:
    point_coords: torch.Tensor = None        # [N,4] f32: z.C after initial_voxelize (pres == vres)
    sparse_hash: torch.Tensor = None         # initial_voxelize: sorted unique voxel hashes
    idx_query: torch.Tensor = None           # point -> stride-1 voxel
    counts: torch.Tensor = None
    coord_maps: dict = field(default_factory=dict)     # stride -> [N_s,4] int32
    tables: dict = field(default_factory=dict)         # stride -> CoordTable
    kernel_maps: dict = field(default_factory=dict)    # torchsparse key -> KernelMap
    p2v: dict = field(default_factory=dict)            # stride -> (idx_query int32 [N], counts int32 [N_s])
    v2p: dict = field(default_factory=dict)            # stride -> (idx int32 [N,8], weights f32 [N,8])
    extras: dict = field(default_factory=dict)         # batch-level tensors that travel with the plan

    def tensors(self):
        out = [self.point_coords, self.sparse_hash, self.idx_query, self.counts]
        out += list(self.coord_maps.values())
        for t in self.tables.values():
            out += [t.keys, t.vals]
        for km in self.kernel_maps.values():
            out += [km.nbr, km._nbrT, km._pairs, km._offsets, km._ppos, km._pposT]
            for op in km._os.values():
                out += op.tensors()
        for a, b in list(self.p2v.values()) + list(self.v2p.values()):
            out += [a, b]
        for v in self.extras.values():
            if isinstance(v, torch.Tensor):
                out.append(v)
            elif hasattr(v, "F"):
                out += [v.F, v.C]
        return [t for t in out if isinstance(t, torch.Tensor)]

    def record_stream(self, stream):
        for t in self.tensors():
            t.record_stream(stream)


def _force(km: KernelMap, need_nbrT: bool):
    from . import conv_engine
    km.num_pairs()              # pairs, offsets, ppos + the single host read of this map
    if conv_engine.mode() == "tc" and conv_engine.os_enabled():
        trows = conv_engine.os_tile_rows(km, 0, 0)
        km.os_plan("out", trows)       # tile schedules of the output-stationary convolution (forward / dgrad sides)
        if not km.symmetric:
            km.os_plan("in", trows)
    else:
        km.pposT
    if need_nbrT:
        km.nbrT


def build_plan(coords: torch.Tensor, strides=(1, 2, 4, 8, 16), v2p_strides=(1, 16, 4), p2v_strides=(16, 4),
               fp32_layers: bool = False) -> GeometryPlan:
    """coords: the batch's voxel coordinates [N,4] (x,y,z,b), int or float as collate.py:67 delivers them."""
    plan = GeometryPlan()
    zc = coords.float()
    plan.point_coords = zc
    c_int = torch.floor(zc).int()
    pc_hash = ops.hash_coords(c_int)
    plan.sparse_hash, plan.idx_query, plan.counts, first = ops.unique_sorted(pc_hash)
    c = ops.gather_rows_i32(c_int, first)
    s = strides[0]
    plan.coord_maps[s] = c
    plan.tables[s] = CoordTable(plan.sparse_hash)
    for nxt in strides[1:] + (None,):
        km3 = build_kernel_map(c, c, 3, s, plan.tables[s])
        km3.os_cluster = conv_engine.os_cluster_for_stride(s)
        plan.kernel_maps["k3_os%d_s1_d1" % s] = km3
        _force(km3, need_nbrT=fp32_layers)
        if nxt is None:
            break
        ratio = nxt // s
        cn = spdownsample(c, nxt)
        km2 = build_kernel_map(c, cn, 2, s, plan.tables[s])
        plan.kernel_maps["k2_os%d_s%d_d1" % (s, ratio)] = km2
        _force(km2, need_nbrT=fp32_layers)
        plan.coord_maps[nxt] = cn
        plan.tables[nxt] = CoordTable.from_coords(cn)
        c, s = cn, nxt
    os_plans = [op for km in plan.kernel_maps.values() for op in km._os.values()]
    if os_plans:                # one host read for the (passes, units, scratch slots, cap) of every schedule
        nums = torch.stack([op.num for op in os_plans]).tolist()
        for op, n in zip(os_plans, nums):
            op.counts = tuple(int(v) for v in n)
            op.host_counts()
    for s in v2p_strides:
        plan.v2p[s] = ops.v2p_build(zc, s, plan.tables[s])
    for s in p2v_strides:
        plan.p2v[s] = ops.p2v_build(zc, s, plan.tables[s], plan.coord_maps[s].shape[0])
    return plan


class Prefetcher:
    """Runs ``fn(*args)`` (batch upload + voxelization + build_plan) on a high-priority side stream, from a worker
    thread: the ~15 host reads of data-dependent sizes block that thread (GIL released), not the thread that is
    enqueueing the current step's convolutions.

    ``submit`` queues the work for the NEXT step; ``get`` waits for the worker, makes the consumer's stream wait for
    the side stream and tells the caching allocator which stream now uses the tensors."""

    def __init__(self, device=None, threaded: bool = False, priority: int = -1):
        import queue
        import threading
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device, priority=priority)
        self._pending = None
        self._threaded = threaded
        if threaded:
            self._jobs, self._done = queue.Queue(), queue.Queue()
            self._worker = threading.Thread(target=self._loop, name="ft3d-prefetch", daemon=True)
            self._worker.start()

    def _run(self, fn, args):
        with torch.cuda.stream(self.stream):
            out = fn(*args)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def _loop(self):
        torch.cuda.set_device(self.device)
        while True:
            job = self._jobs.get()
            if job is None:
                return
            try:
                self._done.put((self._run(*job), None))
            except BaseException as e:  # noqa: BLE001 -- re-raised in get()
                self._done.put((None, e))

    def submit(self, fn, *args):
        assert self._pending is None, "one batch in flight"
        # NB no wait on the consumer's stream: fn's inputs must already be complete (resident batches, or pinned host
        # memory that fn uploads itself) -- waiting would chain the plan's host reads behind the running step.
        if self._threaded:
            self._jobs.put((fn, args))
            self._pending = "queued"
        else:
            self._pending = self._run(fn, args)

    def get(self):
        if self._threaded:
            res, err = self._done.get()
            if err is not None:
                self._pending = None
                raise err
            out, ev = res
        else:
            out, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for o in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(o, GeometryPlan):
                o.record_stream(cur)
            elif isinstance(o, torch.Tensor):
                o.record_stream(cur)
        return out

    def close(self):
        if self._threaded and self._worker.is_alive():
            self._jobs.put(None)
            self._worker.join(timeout=5)
