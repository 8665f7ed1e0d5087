"""Thin torch-tensor wrappers over the libft3d C ABI (include/ft3d.h).

PyTorch is used only for device memory (caching allocator) and the current stream; every wrapper
hands raw device pointers, sizes and the stream to one C entry point.  No wrapper has a CPU or
PyTorch fallback: a non-CUDA tensor or a missing library raises.
"""
from __future__ import annotations

import torch

from ._lib import Ft3dError, lib

KPAD = {27: 32, 8: 8}


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_dev = torch.cuda.current_device


def _stream() -> int:
    """Raw cudaStream_t of torch's current stream (the C accessor: ~0.2 us, vs ~3 us for torch.cuda.current_stream())."""
    return _raw_stream(_cur_dev())


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise Ft3dError("%s must be a CUDA tensor: libft3d has no CPU fallback" % name)
    if t.dtype != dtype:
        raise Ft3dError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _p(t):
    return 0 if t is None else t.data_ptr()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------- hashing / tables
def hash_coords(coords: torch.Tensor) -> torch.Tensor:
    coords = _chk(coords, torch.int32, "coords")
    n = coords.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=coords.device)
    lib().hash(coords.data_ptr(), n, out.data_ptr(), _stream())
    return out


def kernel_hash(coords: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    coords = _chk(coords, torch.int32, "coords")
    offsets = _chk(offsets, torch.int32, "offsets")
    n, k = coords.shape[0], offsets.shape[0]
    out = torch.empty((k, n), dtype=torch.int64, device=coords.device)
    lib().kernel_hash(coords.data_ptr(), n, offsets.data_ptr(), k, out.data_ptr(), _stream())
    return out


class CoordTable:
    """Open-addressing table hash(coord) -> row; built once per coordinate set and reused by every
    map that looks voxels up in it (kernel maps, point->voxel, voxel->point)."""

    __slots__ = ("keys", "vals", "cap", "n")

    def __init__(self, hashes: torch.Tensor):
        hashes = _chk(hashes, torch.int64, "hashes")
        self.n = hashes.numel()
        self.cap = int(lib().table_capacity(self.n))
        self.keys = torch.empty(self.cap, dtype=torch.int64, device=hashes.device)
        self.vals = torch.empty(self.cap, dtype=torch.int32, device=hashes.device)
        lib().table_build(hashes.data_ptr(), self.n, self.keys.data_ptr(), self.vals.data_ptr(), self.cap, _stream())

    @classmethod
    def from_coords(cls, coords: torch.Tensor) -> "CoordTable":
        return cls(hash_coords(coords))

    def query(self, queries: torch.Tensor) -> torch.Tensor:
        q = _chk(queries, torch.int64, "queries")
        out = torch.empty_like(q)
        lib().table_query(q.data_ptr(), q.numel(), self.keys.data_ptr(), self.vals.data_ptr(), self.cap,
                          out.data_ptr(), _stream())
        return out


# ----------------------------------------------------------------------------- unique / quantize
def unique_sorted(keys: torch.Tensor):
    """(unique ascending, inverse int32, counts int32, first-occurrence row int32) of int64 keys."""
    keys = _chk(keys, torch.int64, "keys")
    n = keys.numel()
    dev = keys.device
    uniq = torch.empty(n, dtype=torch.int64, device=dev)
    inverse = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    first = torch.empty(n, dtype=torch.int32, device=dev)
    num = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = lib().unique_workspace(n)
    ws = _ws(wsb, dev)
    lib().unique(keys.data_ptr(), n, uniq.data_ptr(), inverse.data_ptr(), counts.data_ptr(), first.data_ptr(),
                 num.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    m = int(num.item())            # one host read per coordinate level (sizes the next allocation)
    return uniq[:m], inverse, counts[:m], first[:m]


def scale_coords(points: torch.Tensor, scan_id: torch.Tensor, num_scans: int, scale: float, full_scale: int,
                 rot: torch.Tensor | None = None, transl_u: torch.Tensor | None = None):
    """a1.  ``rot`` f32 [num_scans,3,3] / ``transl_u`` f64 [num_scans,3] (device): the per-scan draws of the
    augmentation branch (utils/augment.py), applied with numpy's arithmetic by ``ft3d_augment_scale_coords``."""
    points = _chk(points, torch.float32, "points")
    scan_id = _chk(scan_id, torch.int32, "scan_id")
    n = points.shape[0]
    dev = points.device
    coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    if rot is not None or transl_u is not None:
        if rot is not None:
            rot = _chk(rot, torch.float32, "rot")
            if tuple(rot.shape) != (num_scans, 3, 3):
                raise Ft3dError("scale_coords: rot must be [num_scans, 3, 3]")
        if transl_u is not None:
            transl_u = _chk(transl_u, torch.float64, "transl_u")
            if tuple(transl_u.shape) != (num_scans, 3):
                raise Ft3dError("scale_coords: transl_u must be [num_scans, 3]")
        ws = torch.empty(num_scans * 6, dtype=torch.float32, device=dev)
        lib().augment_scale_coords(points.data_ptr(), scan_id.data_ptr(), n, num_scans, float(scale), int(full_scale),
                                   _p(rot), _p(transl_u), coords.data_ptr(), keep.data_ptr(), ws.data_ptr(), _stream())
        return coords, keep.bool()
    mins = torch.empty(num_scans * 3, dtype=torch.float32, device=dev)
    lib().scale_coords(points.data_ptr(), scan_id.data_ptr(), n, num_scans, float(scale), int(full_scale),
                       coords.data_ptr(), keep.data_ptr(), mins.data_ptr(), _stream())
    return coords, keep.bool()


def quantize(coords: torch.Tensor, num_scans: int):
    """GPU sparse_quantize over a batch of scans: (inds int32 [U], inverse int32 [n], per-scan unique counts)."""
    coords = _chk(coords, torch.int32, "coords")
    n = coords.shape[0]
    dev = coords.device
    inds = torch.empty(n, dtype=torch.int32, device=dev)
    inverse = torch.empty(n, dtype=torch.int32, device=dev)
    scan_counts = torch.empty(num_scans, dtype=torch.int32, device=dev)
    num = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _ws(lib().quantize_workspace(n, num_scans), dev)
    lib().quantize(coords.data_ptr(), n, num_scans, inds.data_ptr(), inverse.data_ptr(), scan_counts.data_ptr(),
                   num.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    u = int(num.item())
    return inds[:u], inverse, scan_counts


def coarsen_hash(coords: torch.Tensor, ratio: int):
    coords = _chk(coords, torch.int32, "coords")
    n = coords.shape[0]
    coarse = torch.empty_like(coords)
    h = torch.empty(n, dtype=torch.int64, device=coords.device)
    lib().coarsen_hash(coords.data_ptr(), n, int(ratio), coarse.data_ptr(), h.data_ptr(), _stream())
    return coarse, h


def gather_rows_i32(src: torch.Tensor, first: torch.Tensor) -> torch.Tensor:
    src = _chk(src, torch.int32, "src")
    first = _chk(first, torch.int32, "first")
    m, w = first.numel(), src.shape[1]
    out = torch.empty((m, w), dtype=torch.int32, device=src.device)
    lib().gather_rows_i32(src.data_ptr(), first.data_ptr(), m, w, out.data_ptr(), _stream())
    return out


# ----------------------------------------------------------------------------- kernel maps
def kmap_build(coords_q: torch.Tensor, offsets: torch.Tensor, table: CoordTable) -> torch.Tensor:
    coords_q = _chk(coords_q, torch.int32, "coords_q")
    offsets = _chk(offsets, torch.int32, "offsets")
    k = offsets.shape[0]
    kpad = KPAD[k]
    n_out = coords_q.shape[0]
    nbr = torch.empty((n_out, kpad), dtype=torch.int32, device=coords_q.device)
    lib().kmap_build(coords_q.data_ptr(), n_out, offsets.data_ptr(), k, table.keys.data_ptr(),
                     table.vals.data_ptr(), table.cap, nbr.data_ptr(), kpad, _stream())
    return nbr


def kmap_pairs(nbr: torch.Tensor, k: int):
    """Reference-format pair list (upper-bound sized), device-side prefix offsets [K+1] and the pair-position
    table ppos [n_out,kpad] (position of pair (row, offset) in the list, -1 if absent)."""
    nbr = _chk(nbr, torch.int32, "nbr")
    n_out, kpad = nbr.shape
    dev = nbr.device
    pairs = torch.empty((max(n_out * k, 1), 2), dtype=torch.int32, device=dev)
    offsets = torch.empty(k + 1, dtype=torch.int32, device=dev)
    ppos = torch.empty((n_out, kpad), dtype=torch.int32, device=dev)
    ws = _ws(lib().kmap_pairs_workspace(n_out, kpad), dev)
    lib().kmap_pairs(nbr.data_ptr(), n_out, k, kpad, pairs.data_ptr(), offsets.data_ptr(), ppos.data_ptr(),
                     ws.data_ptr(), ws.numel(), _stream())
    return pairs, offsets, ppos


def kmap_pair_positions(pairs, offsets, k: int, kpad: int, col: int, n_rows: int, max_pairs: int) -> torch.Tensor:
    out = torch.empty((n_rows, kpad), dtype=torch.int32, device=pairs.device)
    lib().kmap_pair_positions(pairs.data_ptr(), offsets.data_ptr(), k, kpad, int(col), n_rows, int(max_pairs),
                              out.data_ptr(), _stream())
    return out


def kmap_transpose(nbr: torch.Tensor, k: int, n_in: int) -> torch.Tensor:
    nbr = _chk(nbr, torch.int32, "nbr")
    n_out, kpad = nbr.shape
    out = torch.empty((n_in, kpad), dtype=torch.int32, device=nbr.device)
    lib().kmap_transpose(nbr.data_ptr(), n_out, k, kpad, out.data_ptr(), n_in, _stream())
    return out


# ----------------------------------------------------------------------------- point <-> voxel
def count(idx: torch.Tensor, m: int) -> torch.Tensor:
    idx = _chk(idx, torch.int32, "idx")
    out = torch.empty(m, dtype=torch.int32, device=idx.device)
    lib().count(idx.data_ptr(), idx.numel(), out.data_ptr(), m, _stream())
    return out


def deterministic() -> bool:
    """``FT3D_DETERMINISTIC=1``: every floating-point reduction of the training step runs in a fixed order -- the
    point<->voxel scatter-adds as sorted segmented sums (``_segments`` + ``ft3d_segsum_rows``), the weight gradients as
    two-stage / single-owner sums (``wgrad_deterministic``), BatchNorm sums and conv_os as always.  Two steps from the
    same state give bit-identical gradients (tests/test_gpu_fidelity.py); slower, so opt-in."""
    import os
    return os.environ.get("FT3D_DETERMINISTIC", "0") not in ("0", "")


def _segments(dest, m: int):
    """CSR view of a scatter: ``dest`` int32 [E] destination row of every contribution (< 0 or >= m: dropped) ->
    (order int32 [E]: contributions sorted by destination, stable; offsets int32 [m+1])."""
    key = torch.where((dest < 0) | (dest >= m), torch.full_like(dest, m), dest)
    skey, order = torch.sort(key, stable=True)
    offsets = torch.searchsorted(skey, torch.arange(m + 1, dtype=skey.dtype, device=dest.device))
    return order.to(torch.int32), offsets.to(torch.int32)


def voxelize_fwd(feat, idx, cnt):
    feat = _chk(feat, torch.float32, "feat")
    idx = _chk(idx, torch.int32, "idx")
    cnt = _chk(cnt, torch.int32, "cnt")
    n, c = feat.shape
    m = cnt.numel()
    out = torch.empty((m, c), dtype=torch.float32, device=feat.device)
    if deterministic() and n > 0 and m > 0:
        order, offsets = _segments(idx, m)
        lib().segsum_rows(feat.data_ptr(), order.data_ptr(), None, offsets.data_ptr(), cnt.data_ptr(), m, c,
                          out.data_ptr(), _stream())
        return out
    lib().voxelize_fwd(feat.data_ptr(), idx.data_ptr(), cnt.data_ptr(), n, m, c, out.data_ptr(), _stream())
    return out


def voxelize_bwd(gout, idx, cnt, n):
    gout = _chk(gout, torch.float32, "gout")
    m, c = gout.shape
    gin = torch.empty((n, c), dtype=torch.float32, device=gout.device)
    lib().voxelize_bwd(gout.data_ptr(), idx.data_ptr(), cnt.data_ptr(), n, m, c, gin.data_ptr(), _stream())
    return gin


def devoxelize_fwd(feat, idx, w):
    feat = _chk(feat, torch.float32, "feat")
    idx = _chk(idx, torch.int32, "idx")
    w = _chk(w, torch.float32, "w")
    m, c = feat.shape
    n = idx.shape[0]
    out = torch.empty((n, c), dtype=torch.float32, device=feat.device)
    lib().devoxelize_fwd(feat.data_ptr(), idx.data_ptr(), w.data_ptr(), n, m, c, out.data_ptr(), _stream())
    return out


def devoxelize_bwd(gout, idx, w, m):
    gout = _chk(gout, torch.float32, "gout")
    n, c = gout.shape
    gfeat = torch.empty((m, c), dtype=torch.float32, device=gout.device)
    if deterministic() and n > 0 and m > 0:
        order, offsets = _segments(idx.reshape(-1), m)          # contribution e = 8 * point + corner
        rows = torch.div(order, 8, rounding_mode="floor").to(torch.int32)
        wts = w.reshape(-1)[order.long()].contiguous()
        lib().segsum_rows(gout.data_ptr(), rows.data_ptr(), wts.data_ptr(), offsets.data_ptr(), None, m, c,
                          gfeat.data_ptr(), _stream())
        return gfeat
    lib().devoxelize_bwd(gout.data_ptr(), idx.data_ptr(), w.data_ptr(), n, m, c, gfeat.data_ptr(), _stream())
    return gfeat


def ti_weights(pc, idx_query, scale):
    pc = _chk(pc, torch.float32, "pc")
    idx_query = _chk(idx_query, torch.int64, "idx_query")
    if pc.shape[1] != 4:
        raise Ft3dError("calc_ti_weights expects point coordinates [N,4]")
    n = pc.shape[0]
    w = torch.empty((8, n), dtype=torch.float32, device=pc.device)
    lib().ti_weights(pc.data_ptr(), idx_query.data_ptr(), n, float(scale), w.data_ptr(), _stream())
    return w


def v2p_build(pc, stride, table: CoordTable):
    pc = _chk(pc, torch.float32, "pc")
    n = pc.shape[0]
    idx = torch.empty((n, 8), dtype=torch.int32, device=pc.device)
    w = torch.empty((n, 8), dtype=torch.float32, device=pc.device)
    lib().v2p_build(pc.data_ptr(), n, int(stride), table.keys.data_ptr(), table.vals.data_ptr(), table.cap,
                    idx.data_ptr(), w.data_ptr(), _stream())
    return idx, w


def p2v_build(pc, stride, table: CoordTable, m):
    pc = _chk(pc, torch.float32, "pc")
    n = pc.shape[0]
    idx = torch.empty(n, dtype=torch.int32, device=pc.device)
    cnt = torch.empty(m, dtype=torch.int32, device=pc.device)
    lib().p2v_build(pc.data_ptr(), n, int(stride), table.keys.data_ptr(), table.vals.data_ptr(), table.cap,
                    idx.data_ptr(), cnt.data_ptr(), m, _stream())
    return idx, cnt


# ----------------------------------------------------------------------------- lift
def _fmap_strides(fmap):
    if fmap.dim() != 4:
        raise Ft3dError("feature map must be [B,C,H,W]")
    return fmap.stride(0), fmap.stride(1), fmap.stride(2), fmap.stride(3)


def lift_fwd(fmap, rc, bidx):
    if not (fmap.is_cuda and fmap.dtype == torch.float32):
        raise Ft3dError("feature map must be a CUDA float32 tensor")
    rc = _chk(rc, torch.int32, "rc")
    bidx = _chk(bidx, torch.int32, "bidx")
    b, c, h, w = fmap.shape
    sb, sc, sh, sw = _fmap_strides(fmap)
    n = rc.shape[0]
    out = torch.empty((n, c), dtype=torch.float32, device=fmap.device)
    lib().lift_fwd(fmap.data_ptr(), sb, sc, sh, sw, b, c, h, w, rc.data_ptr(), bidx.data_ptr(), n, out.data_ptr(),
                   _stream())
    return out


def lift_bwd(gout, rc, bidx, shape, channels_last):
    gout = _chk(gout, torch.float32, "gout")
    b, c, h, w = shape
    mf = torch.channels_last if channels_last else torch.contiguous_format
    gmap = torch.empty(shape, dtype=torch.float32, device=gout.device, memory_format=mf).zero_()
    sb, sc, sh, sw = _fmap_strides(gmap)
    lib().lift_bwd(gout.data_ptr(), sb, sc, sh, sw, b, c, h, w, rc.data_ptr(), bidx.data_ptr(), rc.shape[0],
                   gmap.data_ptr(), _stream())
    return gmap


# ----------------------------------------------------------------------------- sparse convolution
def conv_gather_f32(inp, nbr, k, kflip, w, w_transposed):
    """out[j] = sum_k inp[nbr[j,k]] @ (W[k] or W[k]^T), fp32 CUDA-core path."""
    inp = _chk(inp, torch.float32, "inp")
    nbr = _chk(nbr, torch.int32, "nbr")
    w = _chk(w, torch.float32, "w")
    n_out, kpad = nbr.shape
    cin, cout = w.shape[-2], w.shape[-1]
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    if inp.shape[1] != red:
        raise Ft3dError("conv: feature width %d != %d" % (inp.shape[1], red))
    out = torch.empty((n_out, ncols), dtype=torch.float32, device=inp.device)
    lib().conv_gather_f32(inp.data_ptr(), nbr.data_ptr(), n_out, k, kpad, int(kflip), red, ncols, w.data_ptr(),
                          int(w_transposed), out.data_ptr(), _stream())
    return out


def conv_wgrad_f32(a, b, pairs, pair_offsets, k, ca, cin, cout, max_pairs):
    a = _chk(a, torch.float32, "a")
    b = _chk(b, torch.float32, "b")
    gw = torch.zeros((k, cin, cout), dtype=torch.float32, device=a.device)
    fn = lib().conv_wgrad_f32_det if deterministic() else lib().conv_wgrad_f32
    fn(a.data_ptr(), b.data_ptr(), pairs.data_ptr(), pair_offsets.data_ptr(), k, int(ca), cin, cout, int(max_pairs),
       gw.data_ptr(), _stream())
    return gw


_PACK_CACHE = {}
FORCE_REPACK = False     # graph.py sets this while capturing: the pack launch must be part of every replay
PACK_EPOCH = 0           # bumped per capture; images packed by pack_all() inside the capture are not packed twice
REPLAY_EPOCH = 0         # bumped by graph.GraphedStep after every replay: the captured optimizer step moved the weights
                         # on the device without touching Tensor._version, so every cached image is one step old


class WeightPacker:
    """All bf16 weight images of a set of conv kernels (both orientations) in ONE launch per step.

    Built once (the fp32 parameters and the images keep their addresses); ``pack()`` refreshes every image and marks
    the per-tensor cache entries fresh, so the per-layer ``packed_weights`` calls of that step are cache hits."""

    def __init__(self, kernels):
        import struct
        import weakref
        self.entries = []
        recs, chunk = [], 0
        for p in kernels:
            w = p.detach()
            w3 = w.unsqueeze(0) if w.dim() == 2 else w
            k, cin, cout = w3.shape
            for wt in (False, True):
                red, ncols = (cout, cin) if wt else (cin, cout)
                nbytes = int(lib().conv_packed_bytes(k, red, ncols))
                img = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
                recs.append(struct.pack("<QQiiiiq", w3.data_ptr(), img.data_ptr(), k, cin, cout, int(wt), chunk))
                chunk += nbytes // 16
                self.entries.append((weakref.ref(p), wt, w3.data_ptr(), img))
        assert lib().conv_pack_desc_bytes() == 40
        self.total = chunk
        self.n = len(recs)
        dev = kernels[0].device if kernels else None
        self.desc = (torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev) if recs else None)

    def pack(self):
        if not self.n:
            return
        lib().conv_pack_weights_multi(self.desc.data_ptr(), self.n, self.total, _stream())
        for ref, wt, ptr, img in self.entries:
            p = ref()
            if p is not None:
                _PACK_CACHE[(id(p), wt)] = (ref, p._version, ptr, img, PACK_EPOCH, REPLAY_EPOCH)


def packed_weights(w: torch.Tensor, w_transposed: bool, owner=None) -> torch.Tensor:
    """bf16 swizzled weight image for the tcgen05 kernels.  Cached per owning tensor object and version counter:
    repacked only after an in-place update (optimizer step) or when a different tensor is passed; entries die
    with their owner (weak reference), so a recycled device address can never alias a stale image."""
    import weakref
    w = _chk(w, torch.float32, "w")
    owner = w if owner is None else owner
    key = (id(owner), bool(w_transposed))
    hit = _PACK_CACHE.get(key)
    if (hit is not None and hit[0]() is owner and hit[1] == owner._version and hit[2] == w.data_ptr()
            and hit[5] == REPLAY_EPOCH and (not FORCE_REPACK or hit[4] == PACK_EPOCH)):
        return hit[3]
    k, cin, cout = w.shape
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    nbytes = lib().conv_packed_bytes(k, red, ncols)
    reuse = hit is not None and hit[0]() is owner and hit[3].numel() == nbytes
    img = hit[3] if reuse else torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    lib().conv_pack_weights(w.data_ptr(), k, cin, cout, int(w_transposed), img.data_ptr(), _stream())
    if len(_PACK_CACHE) > 512:
        for kk in [kk for kk, v in _PACK_CACHE.items() if v[0]() is None]:
            del _PACK_CACHE[kk]
    _PACK_CACHE[key] = (weakref.ref(owner), owner._version, w.data_ptr(), img, -1, REPLAY_EPOCH)
    return img


# ----------------------------------------------------------------------------- pair-major tensor-core path
def to_bf16(x: torch.Tensor) -> torch.Tensor:
    x = _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lib().to_bf16(x.data_ptr(), x.numel(), out.data_ptr(), _stream())
    return out


def conv_pairs_tc(x16, pairs, offsets, k, gather_col, max_pairs, w, w_transposed, owner=None):
    """partial[p] = x16[pairs[p][gather_col]] @ (W[k(p)] | W[k(p)]^T) for every pair; identity gather if pairs is None."""
    x16 = _chk(x16, torch.bfloat16, "x16")
    cin, cout = w.shape[-2], w.shape[-1]
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    if x16.shape[1] != red:
        raise Ft3dError("conv: feature width %d != %d" % (x16.shape[1], red))
    img = packed_weights(w, w_transposed, owner)
    partial = torch.empty((max(int(max_pairs), 1), ncols), dtype=torch.float32, device=x16.device)
    lib().conv_pairs_tc(x16.data_ptr(), _p(pairs), _p(offsets), k, int(gather_col), int(max_pairs), red, ncols,
                        img.data_ptr(), partial.data_ptr(), _stream())
    return partial


def conv_reduce(partial, ppos, ncols):
    n_rows, kpad = ppos.shape
    out = torch.empty((n_rows, ncols), dtype=torch.float32, device=partial.device)
    lib().conv_reduce(partial.data_ptr(), ppos.data_ptr(), n_rows, kpad, ncols, out.data_ptr(), _stream())
    return out


# ----------------------------------------------------------------------------- output-stationary tensor-core path
OS_CHUNK_PASSES = 0     # passes per work unit of conv_os; 0 = chosen per schedule (csrc/os_plan.cu)


class OsPlan:
    """Schedule of one side of a kernel map for ``conv_os`` (csrc/os_plan.cu): 128-row tiles of output rows sorted by
    occupancy mask, the offsets ("passes") each tile visits with their gather indices, and the work units (tiles, or
    pass ranges of heavy tiles) in longest-first order."""

    __slots__ = ("units", "split_tiles", "num", "out_row", "pass_k", "pass_idx", "T", "n_rows", "K", "counts",
                 "tile_rows")

    def __init__(self, units, split_tiles, num, out_row, pass_k, pass_idx, n_rows, K, counts=None, tile_rows=128):
        self.units, self.split_tiles, self.num = units, split_tiles, num
        self.out_row, self.pass_k, self.pass_idx = out_row, pass_k, pass_idx
        self.tile_rows = tile_rows      # 128 x CTAs of the cluster that shares a tile's weight blocks
        self.T, self.n_rows, self.K = out_row.shape[0] // tile_rows, n_rows, K
        self.counts = counts            # host copy of num[:5] = (passes, units, scratch slots, cap, split tiles)

    def host_counts(self):
        """(passes, units, scratch slots, cap, split tiles) on the host (one read, cached); validates the allocation."""
        if self.counts is None:
            self.counts = tuple(int(v) for v in self.num.tolist()[:5])
        P, U = self.counts[0], self.counts[1]
        if P > self.pass_k.shape[0] or U > self.units.shape[0]:
            raise Ft3dError("conv_os schedule: %d passes / %d units exceed the allocation (%d / %d)"
                            % (P, U, self.pass_k.shape[0], self.units.shape[0]))
        return self.counts

    def passes(self) -> int:
        return self.host_counts()[0]

    def scratch_slots(self) -> int:
        return self.host_counts()[2]

    def tensors(self):
        return [self.units, self.split_tiles, self.num, self.out_row, self.pass_k, self.pass_idx]


def conv_os_plan(table: torch.Tensor, k: int, max_pairs: int | None = None, tile_rows: int = 128) -> OsPlan:
    """``table`` int32 [n_rows, kpad]: the neighbour table seen from the rows the convolution PRODUCES."""
    table = _chk(table, torch.int32, "table")
    n_rows, kpad = table.shape
    dev = table.device
    T = (n_rows + tile_rows - 1) // tile_rows
    cap = max(T * k, 1)
    if max_pairs is not None:
        cap = max(min(cap, int(max_pairs)), 1)         # every pass holds at least one pair
    chunk = int(OS_CHUNK_PASSES)
    ucap = 4 * T + 148                                 # a tile is split into at most 4 units; the placement of the
                                                       # units over the CTA clusters pads the last round (os_plan.cu)
    units = torch.empty((ucap, 8), dtype=torch.int32, device=dev)
    split_tiles = torch.empty((max(T, 1), 4), dtype=torch.int32, device=dev)
    out_row = torch.empty(T * tile_rows, dtype=torch.int32, device=dev)
    pass_k = torch.empty(cap, dtype=torch.int32, device=dev)
    pass_idx = torch.empty((cap, tile_rows), dtype=torch.int32, device=dev)
    num = torch.empty(8, dtype=torch.int32, device=dev)
    ws = _ws(lib().conv_os_plan_workspace(n_rows, ucap), dev)
    lib().conv_os_plan(table.data_ptr(), n_rows, k, kpad, tile_rows, cap, ucap, chunk, units.data_ptr(),
                       split_tiles.data_ptr(), out_row.data_ptr(), pass_k.data_ptr(), pass_idx.data_ptr(), num.data_ptr(),
                       ws.data_ptr(), ws.numel(), _stream())
    return OsPlan(units, split_tiles, num, out_row, pass_k, pass_idx, n_rows, k, tile_rows=tile_rows)


_OS_SCRATCH = {}


def os_scratch(device, nbytes: int):
    """Per-CTA statistics rows + partial tiles of split tiles for conv_os; one buffer per (device, stream) -- the
    launches of a stream are ordered, so every layer can share it -- grown on demand."""
    key = (device.index, _stream())
    hit = _OS_SCRATCH.get(key)
    if hit is None or hit.numel() < nbytes:
        hit = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=device)
        _OS_SCRATCH[key] = hit
    return hit


OS_TRACE = None      # when a list: conv_os appends its trace tensor (int64: [148, 8] per-CTA records, then CTA 0's
                     # per-stage stamps, csrc/conv_os.cu); tools/conv_os_probe.py


def conv_os(x16, plan: OsPlan, w, w_transposed: bool, kflip: bool, n_out: int, owner=None, bn=None,
            scratch_slots: int | None = None):
    """Output-stationary tcgen05 convolution.  ``bn`` = (eps, momentum, running_mean, running_var) adds the
    BatchNorm training statistics of the result: returns (y [n_out,ncols] f32, stat [2,ncols] f32 or None)."""
    x16 = _chk(x16, torch.bfloat16, "x16")
    cin, cout = w.shape[-2], w.shape[-1]
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    if x16.shape[1] != red:
        raise Ft3dError("conv: feature width %d != %d" % (x16.shape[1], red))
    img = packed_weights(w, w_transposed, owner)
    dev = x16.device
    out = torch.empty((n_out, ncols), dtype=torch.float32, device=dev)
    stat = None
    eps = momentum = 0.0
    rm = rv = None
    if bn is not None:
        eps, momentum, rm, rv = bn
        stat = torch.empty((2, ncols), dtype=torch.float32, device=dev)
    slots = plan.scratch_slots() if scratch_slots is None else scratch_slots
    ws = os_scratch(dev, lib().conv_os_workspace(ncols, slots, plan.tile_rows))
    trace = None
    if OS_TRACE is not None:
        trace = torch.zeros((148 * 8 + 2048 + 128,), dtype=torch.int64, device=dev)
        OS_TRACE.append(trace)
    lib().conv_os(x16.data_ptr(), x16.shape[0], plan.units.data_ptr(), plan.split_tiles.data_ptr(),
                  plan.num.data_ptr(), plan.out_row.data_ptr(),
                  plan.pass_k.data_ptr(), plan.pass_idx.data_ptr(), plan.units.shape[0], plan.T, plan.tile_rows, slots,
                  plan.K,
                  int(kflip), red, ncols, img.data_ptr(), out.data_ptr(), n_out, _valid(n_out), float(eps),
                  float(momentum), _p(stat), _p(rm), _p(rv), ws.data_ptr(), ws.numel(), _p(trace), _stream())
    return out, stat


# ----------------------------------------------------------------------------- fused BatchNorm / ReLU / residual
# Static-shape execution (graph.py): activations are padded to a capacity; the number of real rows of every padded
# row dimension lives in a device int32 and is looked up here by that dimension (capacities are made distinct).
ROW_COUNTS = {}


def _valid(n_rows: int) -> int:
    t = ROW_COUNTS.get(n_rows)
    return 0 if t is None else t.data_ptr()


_BN_SCRATCH = {}


def bn_scratch(device):
    """Partial-row workspace of the deterministic column reductions; one per (device, stream): kernels on a stream
    are ordered, so every layer can share it."""
    key = (device.index, _stream())
    hit = _BN_SCRATCH.get(key)
    if hit is None:
        hit = torch.empty(int(lib().bn_workspace(1024)), dtype=torch.uint8, device=device)
        _BN_SCRATCH[key] = hit
    return hit


def conv_reduce_bn(partial, ppos, ncols, eps, momentum, running_mean, running_var):
    """sorted scatter + BatchNorm training statistics: returns (y [n,ncols] f32, stat [2,ncols] f32)."""
    n_rows, kpad = ppos.shape
    dev = partial.device
    out = torch.empty((n_rows, ncols), dtype=torch.float32, device=dev)
    stat = torch.empty((2, ncols), dtype=torch.float32, device=dev)
    ws = bn_scratch(dev)
    lib().conv_reduce_bn(partial.data_ptr(), ppos.data_ptr(), n_rows, kpad, ncols, out.data_ptr(), float(eps),
                         float(momentum), stat.data_ptr(), _p(running_mean), _p(running_var), _valid(n_rows),
                         ws.data_ptr(), ws.numel(), _stream())
    return out, stat


def bn_stats(y, eps, momentum, running_mean, running_var):
    n, c = y.shape
    stat = torch.empty((2, c), dtype=torch.float32, device=y.device)
    ws = bn_scratch(y.device)
    lib().bn_stats(y.data_ptr(), n, c, float(eps), float(momentum), stat.data_ptr(), _p(running_mean), _p(running_var),
                   _valid(n), ws.data_ptr(), ws.numel(), _stream())
    return stat


def col_sum(x, into=None):
    """Column sums of an fp32 [N,C] tensor; with ``into`` they are ADDED to that tensor (gradient arena)."""
    n, c = x.shape
    out = into if into is not None else torch.empty((c,), dtype=torch.float32, device=x.device)
    ws = bn_scratch(x.device)
    lib().col_sum(x.data_ptr(), n, c, out.data_ptr(), int(into is not None), ws.data_ptr(), ws.numel(), _stream())
    return out


def bn_apply(y, stat, gamma, beta, res, relu: bool, want_f32: bool = True, want_bf16: bool = True, extra_cols: int = 0):
    """``extra_cols`` > 0: the outputs are allocated [n, c + extra_cols] and written into their left c columns (row
    pitch c + extra_cols): the caller fills the rest (``copy_cols``) -- a concatenation without a concatenation pass."""
    n, c = y.shape
    ld = c + extra_cols
    z = torch.empty((n, ld), dtype=torch.float32, device=y.device) if want_f32 else None
    z16 = torch.empty((n, ld), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    lib().bn_apply(y.data_ptr(), n, c, stat.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(res), int(relu), _p(z),
                   _p(z16), ld, _valid(n), _stream())
    return z, z16


def copy_cols(src, src16, dst, dst16, col0: int):
    """dst[:, col0:col0+c] = src (and dst16[...] = src16, or bf16(src) when src16 is None) in one launch."""
    n, c = src.shape
    lib().copy_cols(src.data_ptr(), _p(src16), n, c, dst.data_ptr(), _p(dst16), dst.shape[1], int(col0), _stream())


def bn_bwd_reduce(gz, y, z16, z, stat, dgamma_into=None, dbeta_into=None, ldg: int = 0):
    """-> (red [2,C], dgamma [C], dbeta [C]); with ``*_into`` the sums are ADDED to those tensors (gradient arena)
    and (red, None, None) is returned.  ``ldg``: row pitch of gz and of the saved output (0 = C)."""
    n, c = y.shape
    dev = y.device
    red = torch.empty((2, c), dtype=torch.float32, device=dev)
    ws = bn_scratch(dev)
    if dgamma_into is not None:
        lib().bn_bwd_reduce(gz.data_ptr(), y.data_ptr(), _p(z16), _p(z), n, c, stat.data_ptr(), red.data_ptr(),
                            dgamma_into.data_ptr(), dbeta_into.data_ptr(), 1, ldg, _valid(n), ws.data_ptr(), ws.numel(),
                            _stream())
        return red, None, None
    dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
    lib().bn_bwd_reduce(gz.data_ptr(), y.data_ptr(), _p(z16), _p(z), n, c, stat.data_ptr(), red.data_ptr(),
                        dgb[0].data_ptr(), dgb[1].data_ptr(), 0, ldg, _valid(n), ws.data_ptr(), ws.numel(), _stream())
    return red, dgb[0], dgb[1]


def bn_bwd_apply(gz, y, z16, z, stat, gamma, red, want_f32: bool, want_bf16: bool, want_res: bool, ldg: int = 0):
    n, c = gz.shape[0], stat.shape[1]
    dev = gz.device
    gy = torch.empty((n, c), dtype=torch.float32, device=dev) if want_f32 else None
    gy16 = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    gres = torch.empty((n, c), dtype=torch.float32, device=dev) if want_res else None
    lib().bn_bwd_apply(gz.data_ptr(), _p(y), _p(z16), _p(z), n, c, stat.data_ptr(), gamma.data_ptr(), _p(red), _p(gy),
                       _p(gy16), _p(gres), ldg, _valid(n), _stream())
    return gy, gy16, gres


_WGRAD_SCRATCH = {}


def wgrad_deterministic() -> bool:
    """``FT3D_WGRAD=det`` (or ``FT3D_DETERMINISTIC=1``): two-stage split-K weight gradient -- every segment of the
    persistent kernel's schedule flushes to its own slot and a second launch adds the slots of each (offset, Cin block)
    in order: no atomics, bit-identical from run to run.  Costs ~20-40 MB of extra traffic per layer (+11 % step time on
    the bench workload, measured), so the default is the same kernel flushing with red.global.add.f32."""
    import os
    v = os.environ.get("FT3D_WGRAD")
    if v is not None:
        return v == "det"
    return os.environ.get("FT3D_DETERMINISTIC", "0") not in ("0", "")


def conv_wgrad_pairs_tc(a16, b16, pairs, offsets, k, ca, cin, cout, max_pairs, into=None, stream=None):
    """``into``: an fp32 tensor of k*cin*cout elements that the gradient is ACCUMULATED into (a gradient arena);
    otherwise a fresh tensor is returned."""
    a16 = _chk(a16, torch.bfloat16, "a16")
    b16 = _chk(b16, torch.bfloat16, "b16")
    st = _stream() if stream is None else stream
    if wgrad_deterministic():
        dev = a16.device
        gw = into if into is not None else torch.empty((k, cin, cout), dtype=torch.float32, device=dev)
        if int(max_pairs) == 0:
            if into is None:
                gw.zero_()
            return gw
        need = int(lib().conv_wgrad_det_workspace(k, cin, cout, int(max_pairs), int(pairs is not None)))
        key = (dev.index, st)
        ws = _WGRAD_SCRATCH.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(int(need * 1.25), 1 << 20), dtype=torch.uint8, device=dev)
            _WGRAD_SCRATCH[key] = ws
        lib().conv_wgrad_pairs_tc_det(a16.data_ptr(), b16.data_ptr(), _p(pairs), _p(offsets), k, int(ca), cin, cout,
                                      int(max_pairs), gw.data_ptr(), int(into is not None), ws.data_ptr(), ws.numel(), st)
        return gw
    gw = into if into is not None else torch.zeros((k, cin, cout), dtype=torch.float32, device=a16.device)
    lib().conv_wgrad_pairs_tc(a16.data_ptr(), b16.data_ptr(), _p(pairs), _p(offsets), k, int(ca), cin, cout,
                              int(max_pairs), gw.data_ptr(), st)
    return gw
