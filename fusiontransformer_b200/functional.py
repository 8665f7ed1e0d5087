"""torchsparse.nn.functional v1.1.0 operator surface on libft3d.

Each function names the reference call site it serves (file:line under FusionTransformer/).  All take
and return CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import CoordTable
from .sparse_tensor import SparseTensor
from .utils.kernel_region import KernelRegion

__all__ = ["sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights", "conv3d",
           "spdownsample", "KernelMap", "lift", "conv_geometry"]


# ------------------------------------------------------------------------ hashing (models/utils.py:19,46-52,74-80)
def sphash(coords: torch.Tensor, offsets: torch.Tensor | None = None) -> torch.Tensor:
    coords = coords.int() if coords.dtype != torch.int32 else coords
    if offsets is None:
        return ops.hash_coords(coords)
    return ops.kernel_hash(coords, offsets.to(device=coords.device, dtype=torch.int32))


def sphashquery(queries: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Row of each query hash in ``targets`` or -1 (models/utils.py:21,50,80)."""
    table = CoordTable(targets.reshape(-1))
    return table.query(queries.contiguous().view(-1)).view(queries.shape)


def spcount(idx: torch.Tensor, num: int) -> torch.Tensor:
    return ops.count(idx.int() if idx.dtype != torch.int32 else idx, int(num))


# ------------------------------------------------------------------------ point <-> voxel (models/utils.py:24-27,58,87,99)
class _Voxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, idx, cnt):
        ctx.save_for_backward(idx, cnt)
        ctx.n = feat.shape[0]
        return ops.voxelize_fwd(feat, idx, cnt)

    @staticmethod
    def backward(ctx, gout):
        idx, cnt = ctx.saved_tensors
        return ops.voxelize_bwd(gout.contiguous(), idx, cnt, ctx.n), None, None


def spvoxelize(feat: torch.Tensor, idx: torch.Tensor, cnt: torch.Tensor) -> torch.Tensor:
    idx = idx.int() if idx.dtype != torch.int32 else idx
    cnt = cnt.int() if cnt.dtype != torch.int32 else cnt
    return _Voxelize.apply(feat.float().contiguous(), idx.contiguous(), cnt.contiguous())


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, idx, w):
        ctx.save_for_backward(idx, w)
        ctx.m = feat.shape[0]
        return ops.devoxelize_fwd(feat, idx, w)

    @staticmethod
    def backward(ctx, gout):
        idx, w = ctx.saved_tensors
        return ops.devoxelize_bwd(gout.contiguous(), idx, w, ctx.m), None, None


def spdevoxelize(feat: torch.Tensor, idx: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    idx = idx.int() if idx.dtype != torch.int32 else idx
    return _Devoxelize.apply(feat.contiguous(), idx.contiguous(), w.contiguous())


def calc_ti_weights(pc: torch.Tensor, idx_query: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Trilinear weights [8,N] (models/utils.py:81-82)."""
    with torch.no_grad():
        return ops.ti_weights(pc.float().contiguous(), idx_query.long().contiguous(), scale)


# ------------------------------------------------------------------------ kernel maps
class KernelMap:
    """One kernel map of a forward pass, shared by every layer with the same (kernel, stride) key.

    Device-native view: ``nbr`` [n_out, kpad] int32 (output-stationary, -1 = no neighbour).  Derived lazily:
    the input-stationary transpose (dgrad / transposed conv), and the reference-format pair list.
    Indexing ``[0], [1], [2]`` yields the torchsparse v1.1.0 triple ``[pairs [L,2] int32, counts [K] (CPU),
    (n_in, n_out)]`` -- the only place that synchronises with the host.
    """

    def __init__(self, nbr: torch.Tensor, K: int, n_in: int, n_out: int, symmetric: bool):
        self.nbr, self.K, self.n_in, self.n_out, self.symmetric = nbr, K, n_in, n_out, symmetric
        self._nbrT = None
        self._pairs = None
        self._offsets = None
        self._ppos = None
        self._pposT = None
        self._host_offsets = None
        self._event = None
        self._num_pairs = None
        self.aux = {}               # per-map caches owned by the conv kernels (tile schedules, ...)
        self._os = {}               # (side, tile_rows) -> ops.OsPlan of the output-stationary convolution

    @property
    def nbrT(self):
        if self._nbrT is None:
            self._nbrT = ops.kmap_transpose(self.nbr, self.K, self.n_in)
        return self._nbrT

    def _build_pairs(self):
        if self._pairs is None:
            self._pairs, self._offsets, self._ppos = ops.kmap_pairs(self.nbr, self.K)
            self._host_offsets = torch.empty(self.K + 1, dtype=torch.int32, pin_memory=True)
            self._host_offsets.copy_(self._offsets, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()

    @property
    def pairs_padded(self):
        self._build_pairs()
        return self._pairs

    @property
    def pair_offsets(self):
        self._build_pairs()
        return self._offsets

    @property
    def ppos(self):
        """pair position of (output row, offset) -- drives the sorted scatter of the forward conv"""
        self._build_pairs()
        return self._ppos

    @property
    def pposT(self):
        """pair position of (input row, offset) -- sorted scatter of dgrad / transposed conv"""
        if self._pposT is None:
            self._build_pairs()
            self._pposT = ops.kmap_pair_positions(self._pairs, self._offsets, self.K, self.nbr.shape[1], 0,
                                                  self.n_in, self.num_pairs())
        return self._pposT

    def os_plan(self, side: str, tile_rows: int = 128):
        """Tile schedule of conv_os for the rows of one side of the map: ``"out"`` = the map's output rows (forward
        conv, dgrad of a transposed conv; table ``nbr``), ``"in"`` = its input rows (dgrad, transposed conv; table
        ``nbrT``).  A symmetric stride-1 map (nbrT[i,k] == nbr[i,K-1-k]) serves both sides with the "out" schedule
        and mirrored weights (``kflip``)."""
        plan = self._os.get((side, tile_rows))
        if plan is None:
            table = self.nbr if side == "out" else self.nbrT
            plan = ops.conv_os_plan(table, self.K, self._num_pairs, tile_rows)
            self._os[(side, tile_rows)] = plan
        return plan

    def os_args(self, role: str, tile_rows: int = 128):
        """-> (plan, w_transposed, kflip, n_rows) of a convolution role (forward | dgrad | transposed |
        dgrad_transposed)."""
        if role == "forward":
            return self.os_plan("out", tile_rows), False, False, self.n_out
        if role == "dgrad":
            if self.symmetric:
                return self.os_plan("out", tile_rows), True, True, self.n_in
            return self.os_plan("in", tile_rows), True, False, self.n_in
        if role == "transposed":
            return self.os_plan("in", tile_rows), False, False, self.n_in
        if role == "dgrad_transposed":
            return self.os_plan("out", tile_rows), True, False, self.n_out
        raise ValueError(role)

    def host_offsets(self):
        self._build_pairs()
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._host_offsets

    def num_pairs(self) -> int:
        if self._num_pairs is None:
            self._num_pairs = int(self.host_offsets()[self.K])
        return self._num_pairs

    def __getitem__(self, i):
        if i == 0:
            return self.pairs_padded[: self.num_pairs()]
        if i == 1:
            off = self.host_offsets()
            return (off[1:] - off[:-1]).clone()
        if i == 2:
            return (self.n_in, self.n_out)
        raise IndexError(i)

    def __len__(self):
        return 3


def spdownsample(coords: torch.Tensor, ratio: int) -> torch.Tensor:
    """Coarser coordinates in ascending-hash order (reached from spvcnn.py:105-123)."""
    coarse, h = ops.coarsen_hash(coords, ratio)
    _, _, _, first = ops.unique_sorted(h)
    return ops.gather_rows_i32(coarse, first)


def _table_for(x: SparseTensor, stride: int, coords: torch.Tensor) -> CoordTable:
    t = x.tables.get(stride)
    if t is None or t.n != coords.shape[0]:
        t = CoordTable.from_coords(coords)
        x.tables[stride] = t
    return t


def build_kernel_map(coords_in, coords_out, kernel_size, cur_stride, table=None) -> KernelMap:
    off = KernelRegion(kernel_size, cur_stride).get_kernel_offset().to(coords_in.device)
    if table is None:
        table = CoordTable.from_coords(coords_in)
    nbr = ops.kmap_build(coords_out, off, table)
    same = coords_in is coords_out
    return KernelMap(nbr, off.shape[0], coords_in.shape[0], coords_out.shape[0], symmetric=(same and kernel_size % 2 == 1))


# ------------------------------------------------------------------------ sparse convolution
def conv_mode() -> str:
    return os.environ.get("FT3D_CONV", "tc")


class _SparseConv(torch.autograd.Function):
    """conv3d arithmetic.  Default (`FT3D_CONV=tc`): pair-major tcgen05 path of csrc/conv_pairs_tc.cu -- the input
    is rounded to bf16 once (and that copy is what is saved for wgrad), partial rows are produced per pair and
    summed per output row by the sorted scatter.  Shapes the tensor-core kernels do not cover (the 4-channel stem
    conv) and `FT3D_CONV=f32` use the fp32 CUDA-core kernels on the same maps."""

    @staticmethod
    def forward(ctx, feats, kernel, kmap: KernelMap, transpose: bool):
        from . import conv_engine
        ctx.kmap, ctx.transpose = kmap, transpose
        cin, cout = kernel.shape[-2], kernel.shape[-1]
        ctx.pairs_path = conv_engine.pairs_ok(cin, cout)
        if ctx.pairs_path:
            x16 = ops.to_bf16(feats)
            ctx.save_for_backward(x16, kernel)
            if conv_engine.os_enabled():
                return conv_engine.os_conv(x16, kmap, kernel, "transposed" if transpose else "forward")[0]
            return conv_engine.pairs_conv(x16, kmap, kernel, role="transposed" if transpose else "forward")
        ctx.save_for_backward(feats, kernel)
        table = kmap.nbrT if transpose else kmap.nbr
        return conv_engine.gather_conv(feats, table, kmap, kernel, kflip=False, w_transposed=False)

    @staticmethod
    def backward(ctx, gout):
        from . import conv_engine
        feats, kernel = ctx.saved_tensors
        kmap, transpose = ctx.kmap, ctx.transpose
        gout = gout.contiguous()
        gin = gw = None
        if ctx.pairs_path:
            g16 = ops.to_bf16(gout)
            if ctx.needs_input_grad[0] and conv_engine.os_enabled():
                gin = conv_engine.os_conv(g16, kmap, kernel, "dgrad_transposed" if transpose else "dgrad")[0]
            elif ctx.needs_input_grad[0]:
                gin = conv_engine.pairs_conv(g16, kmap, kernel, role="dgrad_transposed" if transpose else "dgrad")
            if ctx.needs_input_grad[1]:
                gw = conv_engine.pairs_wgrad(feats, g16, kmap, kernel.shape[-2], kernel.shape[-1], transpose)
            return gin, gw, None, None
        if ctx.needs_input_grad[0]:
            if transpose:
                table, kflip = kmap.nbr, False
            elif kmap.symmetric:
                table, kflip = kmap.nbr, True
            else:
                table, kflip = kmap.nbrT, False
            gin = conv_engine.gather_conv(gout, table, kmap, kernel, kflip=kflip, w_transposed=True)
        if ctx.needs_input_grad[1]:
            gw = conv_engine.wgrad(feats, gout, kmap, kernel.shape[-2], kernel.shape[-1], transpose)
        return gin, gw, None, None


class _DenseConv(torch.autograd.Function):
    """kernel_size 1 convolution (torchsparse: ``F.matmul(kernel)``) on the same tcgen05 kernels with an identity
    gather: out = bf16(F) @ bf16(W), fp32 accumulate; dgrad and wgrad likewise."""

    @staticmethod
    def forward(ctx, feats, kernel):
        from . import conv_engine
        x16 = ops.to_bf16(feats.contiguous())
        ctx.save_for_backward(x16, kernel)
        return conv_engine.dense_conv(x16, kernel, w_transposed=False)

    @staticmethod
    def backward(ctx, gout):
        from . import conv_engine
        x16, kernel = ctx.saved_tensors
        g16 = ops.to_bf16(gout.contiguous())
        gin = conv_engine.dense_conv(g16, kernel, w_transposed=True) if ctx.needs_input_grad[0] else None
        gw = conv_engine.dense_wgrad(x16, g16, kernel.shape[0], kernel.shape[1]) if ctx.needs_input_grad[1] else None
        return gin, gw


def conv_geometry(inputs: SparseTensor, kernel_size: int, stride: int = 1, dilation: int = 1, transpose: bool = False):
    """Coordinate / kernel-map side of torchsparse conv3d v1.1.0: looks up or builds the map and
    returns ``(kmap or None for kernel_size 1, make_out)`` where ``make_out(feats)`` wraps the result features in a
    SparseTensor with the output coordinates, stride and the propagated map caches."""
    C, s = inputs.C, inputs.s
    if dilation != 1:
        raise NotImplementedError("dilation != 1 is never used by the reference (spvcnn.py) and is not implemented")
    if kernel_size == 1 and stride == 1:
        def make_out(feats):
            out = inputs._like(feats)
            out.check()
            return out
        return None, make_out
    if not transpose:
        key = "k%s_os%d_s%d_d%d" % (kernel_size, s, stride, dilation)
        if stride > 1:
            kmap = inputs.kernel_maps.get(key)
            new_c = inputs.coord_maps.get(s * stride) if kmap is not None else None
            if kmap is None or new_c is None:
                new_c = spdownsample(C, stride * s)
                table = _table_for(inputs, s, C)
                kmap = build_kernel_map(C, new_c, kernel_size, s, table)

            def make_out(feats):
                out = SparseTensor(feats, new_c, s * stride)
                out.coord_maps = dict(inputs.coord_maps)
                out.kernel_maps = dict(inputs.kernel_maps)
                out.tables = inputs.tables          # tables are keyed by stride and never invalidated within a pass
                out.check()
                out.kernel_maps[key] = kmap
                return out
            return kmap, make_out
        kmap = inputs.kernel_maps.get(key)
        if kmap is None:
            table = _table_for(inputs, s, C)
            kmap = build_kernel_map(C, C, kernel_size, s, table)
            inputs.kernel_maps[key] = kmap

        def make_out(feats):
            out = inputs._like(feats)
            out.check()
            return out
        return kmap, make_out
    orig = int(s / stride)
    kmap = inputs.kernel_maps["k%s_os%d_s%d_d%d" % (kernel_size, orig, stride, dilation)]

    def make_out(feats):
        out = SparseTensor(feats, inputs.coord_maps[orig], orig)
        out.coord_maps, out.kernel_maps, out.tables = inputs.coord_maps, inputs.kernel_maps, inputs.tables
        out.check()
        return out
    return kmap, make_out


def conv3d(inputs: SparseTensor, kernel: torch.Tensor, kernel_size: int, bias=None, stride: int = 1,
           dilation: int = 1, transpose: bool = False) -> SparseTensor:
    """torchsparse.nn.functional.conv3d v1.1.0; all 49 spnn.Conv3d of models/spvcnn.py."""
    kmap, make_out = conv_geometry(inputs, kernel_size, stride, dilation, transpose)
    F = inputs.F
    if kmap is None:
        from . import conv_engine
        if conv_engine.pairs_ok(kernel.shape[0], kernel.shape[1]):
            out = make_out(_DenseConv.apply(F, kernel))
        else:
            out = make_out(F.matmul(kernel))
    else:
        out = make_out(_SparseConv.apply(F, kernel, kmap, transpose))
    if bias is not None:
        out.F = out.F + bias
    return out


# ------------------------------------------------------------------------ 2D -> 3D lift
class _Lift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap, rc, bidx):
        ctx.save_for_backward(rc, bidx)
        ctx.shape = tuple(fmap.shape)
        ctx.cl = fmap.is_contiguous(memory_format=torch.channels_last) and not fmap.is_contiguous()
        return ops.lift_fwd(fmap, rc, bidx)

    @staticmethod
    def backward(ctx, gout):
        rc, bidx = ctx.saved_tensors
        return ops.lift_bwd(gout.contiguous(), rc, bidx, ctx.shape, ctx.cl), None, None


def lift(feature_map: torch.Tensor, img_indices, batch_index: torch.Tensor | None = None) -> torch.Tensor:
    """2D->3D feature lift (models/image_models_billinear.py:117-124): ``feats[p] = X[b(p), :, row(p), col(p)]``.

    ``img_indices`` is either the reference's list of per-sample ``[N_i,2]`` (row, col) arrays (numpy or
    tensor), concatenated here in sample order, or a single ``[N,2]`` tensor with ``batch_index`` [N].
    One launch for the whole batch; works on NCHW and channels-last maps (strides are passed through).
    """
    dev = feature_map.device
    if isinstance(img_indices, (list, tuple)):
        rcs, bs = [], []
        for b, idx in enumerate(img_indices):
            t = torch.as_tensor(idx)
            rcs.append(t.to(torch.int32))
            bs.append(torch.full((t.shape[0],), b, dtype=torch.int32))
        rc = torch.cat(rcs, 0).to(dev, non_blocking=True)
        bidx = torch.cat(bs, 0).to(dev, non_blocking=True)
    else:
        rc = img_indices.to(device=dev, dtype=torch.int32)
        bidx = batch_index.to(device=dev, dtype=torch.int32)
    return _Lift.apply(feature_map, rc.contiguous(), bidx.contiguous())
