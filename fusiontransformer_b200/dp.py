"""Data-parallel gradient exchange for the 3D branch (the reference's only parallelism).

Reference: DistributedDataParallel(model, find_unused_parameters=True) at
FusionTransformer/modules/TorchpackInterface.py:78-81 -- shard by scan, per-rank BatchNorm statistics, one
gradient allreduce(sum)/world per step.  Here: one process per GPU (torchrun), all gradients live in ONE flat fp32
arena laid out in reverse registration order (the order backward produces them), parameters' ``.grad`` are views
into it, and each bucket's NCCL all-reduce is launched on a side stream the moment its last gradient has been
accumulated, overlapping the remaining dgrad/wgrad kernels.  Parameters that received no gradient (unused heads)
contribute zeros, which is what ``find_unused_parameters=True`` amounts to.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20, process_group=None, overlap: bool = True,
                 side_wgrad: bool = True):
        self.side_wgrad = side_wgrad       # fused.py: weight-gradient kernels run on a second stream
        self.deferred = False              # True: hooks do nothing, finish() exchanges every bucket (graph replay)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        dev = order[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []            # (start, end) element ranges of the arena
        self._bucket_of = {}
        self._pending = []
        off, bstart, bidx = 0, 0, 0
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[p] = bidx
            off += n
            if (off - bstart) * 4 >= bucket_bytes:
                self.buckets.append((bstart, off))
                bstart, bidx = off, bidx + 1
        if off > bstart:
            self.buckets.append((bstart, off))
        self._view = {p: p.grad for p in order}      # the arena view of every parameter (re-adopted if `.grad` is replaced)
        self._need = [0] * len(self.buckets)
        for p in order:
            self._need[self._bucket_of[p]] += 1
        self._seen = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self.overlap = overlap and self.world > 1 and dev.type == "cuda"
        # NCCL averages inside the collective (no extra pass over the 87 MB arena); gloo only sums
        self._avg = (self.world > 1 and dev.type == "cuda" and dist.get_backend(process_group) == "nccl")
        self.comm_stream = torch.cuda.Stream(device=dev) if self.overlap else None
        self._works = []
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)
        for p in self.params:
            p._ft3d_sink = self      # fused backward kernels may accumulate straight into the arena (fused.py)

    def owns(self, p) -> bool:
        """True while ``p.grad`` is still this arena's view (someone may have replaced it, e.g. set_to_none)."""
        g = p.grad
        return g is not None and g.untyped_storage().data_ptr() == self.flat.untyped_storage().data_ptr()

    def _adopt(self, p):
        """``optimizer.zero_grad()`` (set_to_none=True, what the reference's train_step calls) drops the arena views:
        the fused kernels then return fresh gradient tensors and autograd installs them as ``.grad``.  Bring such a
        gradient back into the arena (copy + re-bind the view) so that the exchange reduces what backward produced."""
        if self.owns(p):
            return
        view = self._view[p]
        g = p.grad
        if g is None:
            view.zero_()                 # an unused parameter contributes zeros (find_unused_parameters=True)
        else:
            view.copy_(g)
            p.grad = view

    def note(self, p):
        """A kernel has added p's gradient into the arena (no autograd accumulation => no hook fires)."""
        if self.world > 1:
            self._hook(p)

    @staticmethod
    def _join_side():
        from .fused import join_side_streams
        join_side_streams()

    # -- broadcast initial parameters/buffers from rank 0 (what DDP's constructor does)
    def broadcast_parameters(self, module: torch.nn.Module):
        if self.world == 1:
            return
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, 0, group=self.group)

    def zero_grad(self):
        """Use instead of ``optimizer.zero_grad()``: zeroes the arena in one launch and keeps the ``.grad`` views (an
        ``optimizer.zero_grad()`` is tolerated -- see ``_adopt`` -- at the price of one copy per parameter)."""
        for p in self.params:
            if p.grad is not self._view[p]:
                p.grad = self._view[p]
        self.flat.zero_()
        self._seen = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def _launch(self, b: int):
        if self._launched[b]:
            return
        self._launched[b] = True
        self._join_side()                   # side-stream wgrad kernels of this bucket must have been ordered first
        s, e = self.buckets[b]
        view = self.flat[s:e]
        if self.overlap:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            from .fused import _BranchStream           # gradients of a bucket may have been produced on either stream
            for idx in _BranchStream.used:
                self.comm_stream.wait_stream(_BranchStream.streams[idx])
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group)
        else:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM,
                                               group=self.group, async_op=True))

    def _hook(self, p):
        if self.deferred:
            return
        self._adopt(p)
        b = self._bucket_of[p]
        self._seen[b] += 1
        if self._seen[b] == self._need[b]:
            self._launch(b)

    def finish(self):
        """Call after backward: launches buckets with unused parameters, waits, and averages over ranks."""
        self._join_side()
        for b, launched in enumerate(self._launched):
            if not launched:             # parameters whose gradient never arrived through a hook (unused, or world 1)
                for p in self.params:
                    if self._bucket_of[p] == b:
                        self._adopt(p)
        if self.world == 1:
            return
        if self.deferred:
            # nothing was launched from hooks, and under CUDA-graph replay zero_grad() (which re-arms the buckets) runs
            # only inside the captured graph, not in Python: every bucket is exchanged here, every step
            self._launched = [False] * len(self.buckets)
        for b in range(len(self.buckets)):
            self._launch(b)
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        for w in self._works:
            w.wait()
        self._works = []
        if not self._avg:
            self.flat.mul_(1.0 / self.world)
        # re-arm for the next step here, not only in zero_grad(): a caller that clears gradients with
        # optimizer.zero_grad() (or a graph replay, which runs no Python) never reaches GradSync.zero_grad()
        self._seen = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)


def shard_indices(num_items: int, rank: int, world: int, epoch: int = 0, shuffle: bool = False, seed: int = 0):
    """DistributedSampler semantics of FusionTransformer/data/build.py:62-67: pad to a multiple of world by
    wrapping, then take every world-th index starting at rank."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(num_items, generator=g).tolist()
    else:
        idx = list(range(num_items))
    total = (num_items + world - 1) // world * world
    idx += idx[: total - len(idx)]
    return idx[rank:total:world]
