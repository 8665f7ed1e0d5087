"""Output-stationary tcgen05 convolution (csrc/conv_os.cu) and its tile schedule (csrc/os_plan.cu).

The schedule is product-internal (the reference has no such object), so it is checked by its defining properties
against the neighbour table it was built from; the arithmetic is checked against the CPU oracle's sparse convolution
(torchsparse semantics) per layer at the north-star 5e-3, in both gather modes (TMA gather4 and
16-byte cp.async), with and without the BatchNorm-statistics epilogue."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ft():
    import fusiontransformer_b200 as ft
    return ft


@pytest.fixture(scope="module")
def geom(small_batch, ft):
    from oracle import ft_glue as og, ts_ops as ts
    z = ts.PointTensor(small_batch["feats"], small_batch["coords"].float())
    x0 = og.initial_voxelize(z, 1, 1)
    spf = ft.nn.functional
    C = x0.C.cuda()
    C2 = spf.spdownsample(C, 2)
    return dict(Co=x0.C, C=C, C2=C2, km3=spf.build_kernel_map(C, C, 3, 1), km2=spf.build_kernel_map(C, C2, 2, 1))


def _check_plan(plan, table, K):
    table = table.cpu().numpy()
    n = table.shape[0]
    P, U, S, cap, NS = plan.host_counts()
    slots, out_row = plan.units.cpu().numpy()[:U], plan.out_row.cpu().numpy()
    # placement (os_assign_kernel): slot j * ncl + c is the unit of cluster c in round j; the last round is padded
    # with holes (tile -1, no passes).  Round by round the longest unit goes to the least loaded cluster.
    ncl = 148 // (plan.tile_rows // 128)
    assert U % ncl == 0
    holes = slots[:, 2] < 0
    assert np.all(slots[holes][:, 1] == 0) and not holes[:U - ncl].any()
    units = slots[~holes]
    load = np.zeros(ncl, np.int64)
    sizes_in_order = []
    for j in range(U // ncl):
        rnd = slots[j * ncl:(j + 1) * ncl]
        order = np.lexsort((np.arange(ncl), load))                     # clusters by load so far, ties by index
        got = rnd[order]                                               # units in the order they were handed out
        live = got[:, 2] >= 0
        assert not live[np.argmin(live):].any() if not live.all() else True   # holes go to the most loaded clusters
        sizes_in_order += [int(v) for v in got[live][:, 1]]
        load += rnd[:, 1]
    assert np.all(np.diff(sizes_in_order) <= 0)                        # longest units first
    assert load.max() <= int(np.ceil(load.sum() / ncl)) + int(units[:, 1].max())     # within one unit of the mean
    pass_k, pass_idx = plan.pass_k.cpu().numpy()[:P], plan.pass_idx.cpu().numpy()[:P]
    R = plan.tile_rows
    T = (n + R - 1) // R
    assert out_row.shape == (T * R,) and 1 <= cap <= 32 and pass_idx.shape[1] == R
    live = out_row[out_row >= 0]
    assert np.array_equal(np.sort(live), np.arange(n))                 # every row produced exactly once
    assert np.all(out_row[n:] == -1)                                   # empty slots only behind the last row
    assert np.all(units[:, 1] <= np.maximum(cap, 7)) and units[:, 1].sum() == P and np.all(units[:, 3] <= 4)
    occ = table[:, :K] >= 0
    pairs_seen, slots_seen = 0, 0
    by_tile = {}
    for un in units:
        by_tile.setdefault(int(un[2]), []).append(un)
    assert sorted(by_tile) == list(range(T))                           # every tile has its unit(s)
    for t in range(T):
        us = sorted(by_tile[t], key=lambda un: un[4])
        assert [int(un[4]) for un in us] == list(range(len(us))) and all(int(un[3]) == len(us) for un in us)
        assert all(int(un[5]) == int(us[0][5]) for un in us)
        if len(us) > 1:
            slots_seen += len(us)
            sizes = [int(un[1]) for un in us]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1 and sum(sizes) > cap   # even, non-empty split
        first, npass = int(us[0][0]), sum(int(un[1]) for un in us)
        for a, b in zip(us, us[1:]):
            assert int(b[0]) == int(a[0]) + int(a[1])                  # disjoint, consecutive pass ranges
        rows = out_row[t * R:(t + 1) * R]
        ok = rows >= 0
        union = np.flatnonzero(occ[rows[ok]].any(0))
        assert np.array_equal(pass_k[first:first + npass], union)      # ascending offsets, exactly the used ones
        for j, k in enumerate(union):
            idx = pass_idx[first + j]
            assert np.array_equal(idx[ok], table[rows[ok], k]) and np.all(idx[~ok] == -1)
            pairs_seen += int((idx >= 0).sum())
    assert pairs_seen == int(occ.sum()) and slots_seen == S
    split = plan.split_tiles.cpu().numpy()[:NS]
    want = [(t, len(by_tile[t]), int(by_tile[t][0][5])) for t in range(T) if len(by_tile[t]) > 1]
    assert [tuple(int(v) for v in row[:3]) for row in split] == want     # split tiles listed in tile order
    return pairs_seen / max(1, R * P)


def test_schedule_properties(ft, geom):
    from fusiontransformer_b200 import ops
    km3, km2 = geom["km3"], geom["km2"]
    e3 = _check_plan(ops.conv_os_plan(km3.nbr, 27), km3.nbr, 27)
    e2 = _check_plan(ops.conv_os_plan(km2.nbr, 8), km2.nbr, 8)
    eT = _check_plan(ops.conv_os_plan(km2.nbrT, 8), km2.nbrT, 8)
    assert e3 > 0.45 and eT > 0.9, (e3, e2, eT)       # mask sorting keeps the tiles dense (random order: ~0.15)
    # cluster schedules: tiles of 256 / 512 rows shared by 2 / 4 CTAs
    for trows in (256, 512):
        _check_plan(ops.conv_os_plan(km3.nbr, 27, tile_rows=trows), km3.nbr, 27)
        _check_plan(ops.conv_os_plan(km2.nbrT, 8, tile_rows=trows), km2.nbrT, 8)
    # deterministic: the same table gives the same schedule
    a, b = ops.conv_os_plan(km3.nbr, 27), ops.conv_os_plan(km3.nbr, 27)
    assert torch.equal(a.out_row, b.out_row) and torch.equal(a.units[:a.host_counts()[1]], b.units[:b.host_counts()[1]])
    # a forced small chunk size splits many tiles; the schedule stays consistent
    monkey_chunk = ops.OS_CHUNK_PASSES
    try:
        ops.OS_CHUNK_PASSES = 2
        p2 = ops.conv_os_plan(km3.nbr, 27)
        _check_plan(p2, km3.nbr, 27)
        assert p2.host_counts()[2] > 0 and p2.host_counts()[3] == 2
    finally:
        ops.OS_CHUNK_PASSES = monkey_chunk
    # ragged / tiny tables
    for n in (1, 127, 128, 129):
        tbl = km3.nbr[:n].clone()
        tbl[tbl >= n] = -1
        _check_plan(ops.conv_os_plan(tbl, 27), tbl, 27)


@pytest.mark.parametrize("gather", ["tma", "ldgsts"])
@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 96), (96, 128), (128, 256), (192, 128), (384, 256), (256, 384)])
def test_conv_os_k3_forward_dgrad_vs_oracle(ft, geom, monkeypatch, gather, cin, cout):
    from oracle import ts_ops as ts
    from fusiontransformer_b200 import conv_engine, ops
    monkeypatch.setenv("FT3D_OS_GATHER", gather)
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16")
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    C = geom["Co"]
    n = C.shape[0]
    feats = torch.randn(n, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / (cin * 27) ** 0.5
    fo = feats.clone().requires_grad_(True)
    yo = ts.conv3d(ts.SparseTensor(fo, C, 1), w, 3).F
    gsel = torch.randn(yo.shape, generator=g)
    (yo * gsel).sum().backward()
    km = geom["km3"]
    wg = w.cuda()
    y, stat = conv_engine.os_conv(ops.to_bf16(feats.cuda()), km, wg, "forward")
    assert stat is None and rel_l2(y, yo) < 5e-3
    gin, _ = conv_engine.os_conv(ops.to_bf16(gsel.cuda()), km, wg, "dgrad")
    assert rel_l2(gin, fo.grad) < 5e-3
    # two launches give bit-identical rows (no atomics, fixed summation order)
    y2, _ = conv_engine.os_conv(ops.to_bf16(feats.cuda()), km, wg, "forward")
    assert torch.equal(y, y2)


@pytest.mark.parametrize("gather", ["tma", "ldgsts"])
def test_conv_os_down_up_vs_oracle(ft, geom, monkeypatch, gather):
    from oracle import ts_ops as ts
    from fusiontransformer_b200 import conv_engine, ops
    monkeypatch.setenv("FT3D_OS_GATHER", gather)
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16")
    g = torch.Generator().manual_seed(5)
    C = geom["Co"]
    feats = torch.randn(C.shape[0], 64, generator=g)
    w = torch.randn(8, 64, 128, generator=g) / 23.0
    wt = torch.randn(8, 128, 96, generator=g) / 8.0
    fo, wo, wto = feats.clone().requires_grad_(True), w.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    xo = ts.SparseTensor(fo, C, 1)
    xo.check()
    yo = ts.conv3d(xo, wo, 2, stride=2)
    zo = ts.conv3d(yo, wto, 2, stride=2, transpose=True)
    yo.F.retain_grad()
    gsel = torch.randn(zo.F.shape, generator=g)
    (zo.F * gsel).sum().backward()
    km = geom["km2"]
    y, _ = conv_engine.os_conv(ops.to_bf16(feats.cuda()), km, w.cuda(), "forward")
    assert rel_l2(y, yo.F) < 5e-3
    z, _ = conv_engine.os_conv(ops.to_bf16(yo.F.detach().cuda()), km, wt.cuda(), "transposed")
    assert rel_l2(z, zo.F) < 5e-3
    gy, _ = conv_engine.os_conv(ops.to_bf16(gsel.cuda()), km, wt.cuda(), "dgrad_transposed")
    assert rel_l2(gy, yo.F.grad) < 5e-3
    gx, _ = conv_engine.os_conv(ops.to_bf16(yo.F.grad.cuda()), km, w.cuda(), "dgrad")
    assert rel_l2(gx, fo.grad) < 5e-3


@pytest.mark.parametrize("gather", ["tma", "ldgsts"])
@pytest.mark.parametrize("cout", [32, 96, 256])
def test_conv_os_statistics_epilogue(ft, geom, monkeypatch, gather, cout):
    """BatchNorm training statistics from the epilogue == nn.BatchNorm1d's on the rows the kernel wrote; the
    self-resetting counter survives repeated launches; running statistics are updated once per launch."""
    from fusiontransformer_b200 import conv_engine, ops
    monkeypatch.setenv("FT3D_OS_GATHER", gather)
    g = torch.Generator().manual_seed(cout)
    n = geom["C"].shape[0]
    feats = torch.randn(n, 64, generator=g).cuda() + 0.5
    w = (torch.randn(27, 64, cout, generator=g) / 40).cuda()
    x16 = ops.to_bf16(feats)
    for rep in range(3):
        rm, rv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
        y, stat = conv_engine.os_conv(x16, geom["km3"], w, "forward", bn=(1e-5, 0.1, rm, rv))
        mean, var = y.double().mean(0), y.double().var(0, unbiased=False)
        assert rel_l2(stat[0], mean) < 1e-5 and rel_l2(stat[1], 1.0 / torch.sqrt(var + 1e-5)) < 1e-5
        assert rel_l2(rm, 0.1 * mean) < 1e-5
        assert rel_l2(rv, 0.9 + 0.1 * y.double().var(0, unbiased=True)) < 1e-5
    y0, _ = conv_engine.os_conv(x16, geom["km3"], w, "forward")
    assert torch.equal(y0, y)                               # the epilogue does not change the rows


@pytest.mark.parametrize("gather", ["tma", "ldgsts"])
@pytest.mark.parametrize("chunk", [1, 3])
def test_conv_os_split_tiles_fold_deterministically(ft, geom, monkeypatch, gather, chunk):
    """Heavy tiles are split into units over disjoint pass ranges whose partial rows are folded by whichever unit
    finishes last, in unit order: same rows as the unsplit schedule up to fp32 re-association, identical from launch
    to launch, statistics included."""
    from fusiontransformer_b200 import ops
    monkeypatch.setenv("FT3D_OS_GATHER", gather)
    g = torch.Generator().manual_seed(11)
    n = geom["C"].shape[0]
    feats = torch.randn(n, 64, generator=g).cuda()
    w = (torch.randn(27, 64, 96, generator=g) / 40).cuda()
    x16 = ops.to_bf16(feats)
    km = geom["km3"]
    base, stat0 = ops.conv_os(x16, ops.conv_os_plan(km.nbr, 27), w, False, False, n, bn=(1e-5, 0.1, None, None))
    monkeypatch.setattr(ops, "OS_CHUNK_PASSES", chunk)
    plan = ops.conv_os_plan(km.nbr, 27)
    assert plan.host_counts()[2] > 0
    outs = [ops.conv_os(x16, plan, w, False, False, n, bn=(1e-5, 0.1, None, None)) for _ in range(3)]
    assert rel_l2(outs[0][0], base) < 1e-6 and rel_l2(outs[0][1], stat0) < 1e-5
    for y, st in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(st, outs[0][1])


def test_conv_os_ragged_and_isolated_rows(ft, monkeypatch):
    """Row counts around the tile size, voxels without any neighbour but themselves, a single voxel."""
    from oracle import ts_ops as ts
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16")
    for n in (1, 127, 128, 129, 300):
        g = torch.Generator().manual_seed(n)
        lin = torch.randperm(12 ** 3, generator=g)[:n]
        C = torch.stack([lin // 144, (lin // 12) % 12, lin % 12, torch.zeros_like(lin)], 1).int()
        feats = torch.randn(n, 64, generator=g)
        w = torch.randn(27, 64, 64, generator=g) / 40
        yo = ts.conv3d(ts.SparseTensor(feats, C, 1), w, 3)
        yg = ft.nn.functional.conv3d(ft.SparseTensor(feats.cuda(), C.cuda(), 1), w.cuda(), 3)
        assert rel_l2(yg.F, yo.F) < 5e-3, n


def test_conv_os_matches_pair_major_path(ft, geom, monkeypatch):
    """Both tensor-core algorithms compute the same sums of the same bf16 products (different fp32 summation order)."""
    g = torch.Generator().manual_seed(3)
    n = geom["C"].shape[0]
    feats = torch.randn(n, 96, generator=g).cuda()
    w = (torch.randn(27, 96, 96, generator=g) / 50).cuda()
    outs = {}
    for algo in ("os", "pairs"):
        monkeypatch.setenv("FT3D_CONV_ALGO", algo)
        x = ft.SparseTensor(feats, geom["C"], 1)
        outs[algo] = ft.nn.functional.conv3d(x, w, 3).F
    assert rel_l2(outs["os"], outs["pairs"]) < 1e-5


@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 128), (256, 256), (384, 256)])
def test_wgrad_two_stage_is_deterministic_and_matches_atomic(ft, geom, monkeypatch, cin, cout):
    """Two-stage split-K weight gradient: bit-identical across launches, equal to the atomic kernel up to fp32
    re-association, accumulates into a gradient arena, and handles the dense (identity pair list) case."""
    from fusiontransformer_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    km = geom["km3"]
    n = geom["C"].shape[0]
    a16 = ops.to_bf16(torch.randn(n, cin, generator=g).cuda())
    b16 = ops.to_bf16(torch.randn(n, cout, generator=g).cuda())
    L = km.num_pairs()
    monkeypatch.setenv("FT3D_WGRAD", "det")
    outs = [ops.conv_wgrad_pairs_tc(a16, b16, km.pairs_padded, km.pair_offsets, 27, 0, cin, cout, L) for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    arena = torch.ones(27, cin, cout, device="cuda")
    ops.conv_wgrad_pairs_tc(a16, b16, km.pairs_padded, km.pair_offsets, 27, 0, cin, cout, L, into=arena)
    assert rel_l2(arena - 1.0, outs[0]) < 1e-6
    monkeypatch.setenv("FT3D_WGRAD", "atomic")
    ref = ops.conv_wgrad_pairs_tc(a16, b16, km.pairs_padded, km.pair_offsets, 27, 0, cin, cout, L)
    assert rel_l2(outs[0], ref) < 1e-5
    monkeypatch.setenv("FT3D_WGRAD", "det")
    d0 = ops.conv_wgrad_pairs_tc(a16, b16, None, None, 1, 0, cin, cout, n)
    want = a16.float().t() @ b16.float()
    assert rel_l2(d0[0], want) < 1e-5


@pytest.mark.parametrize("cluster,pair", [(2, "1"), (2, "0"), (4, "0")])
@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 128), (256, 256), (384, 256), (256, 384)])
def test_conv_os_cluster_multicast_matches_single_cta(ft, geom, monkeypatch, cluster, pair, cin, cout):
    """Thread-block clusters share one schedule tile's weight blocks.  2 CTAs, default: a CTA pair issuing
    tcgen05.mma.cta_group::2 (M = 256), each CTA holding half of the weight block's columns; FT3D_OS_PAIR=0 and 4 CTAs:
    independent MMAs per CTA fed by a multicast bulk copy.  Same rows as the single-CTA kernel up to fp32
    re-association of the split-tile fold, deterministic, statistics included, forward and dgrad."""
    monkeypatch.setenv("FT3D_OS_PAIR", pair)
    from oracle import ts_ops as ts
    from fusiontransformer_b200 import conv_engine, ops
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16")
    g = torch.Generator().manual_seed(cin + cout + cluster)
    C = geom["Co"]
    n = C.shape[0]
    feats = torch.randn(n, cin, generator=g)
    w = torch.randn(27, cin, cout, generator=g) / (cin * 27) ** 0.5
    fo = feats.clone().requires_grad_(True)
    yo = ts.conv3d(ts.SparseTensor(fo, C, 1), w, 3).F
    gsel = torch.randn(yo.shape, generator=g)
    (yo * gsel).sum().backward()
    km, wg = geom["km3"], w.cuda()
    x16, g16 = ops.to_bf16(feats.cuda()), ops.to_bf16(gsel.cuda())
    monkeypatch.setenv("FT3D_OS_CLUSTER", "1")
    y1, st1 = conv_engine.os_conv(x16, km, wg, "forward", bn=(1e-5, 0.1, None, None))
    monkeypatch.setenv("FT3D_OS_CLUSTER", str(cluster))
    outs = [conv_engine.os_conv(x16, km, wg, "forward", bn=(1e-5, 0.1, None, None)) for _ in range(3)]
    assert rel_l2(outs[0][0], yo) < 5e-3 and rel_l2(outs[0][0], y1) < 1e-5 and rel_l2(outs[0][1], st1) < 1e-4
    for y, st in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(st, outs[0][1])
    gin, _ = conv_engine.os_conv(g16, km, wg, "dgrad")
    assert rel_l2(gin, fo.grad) < 5e-3
