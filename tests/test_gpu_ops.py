"""GPU parity of every libft3d operator against the CPU oracle, called through the Python boundary -> C ABI.

Integer artefacts (hashes, voxel order, inverse maps, kernel maps, idx_query) are bit-exact; fp32 paths are held
to rel-L2 <= 1e-5; the bf16 tensor-core convolution to rel-L2 <= 5e-3 per layer (north_star tolerance).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ft():
    import fusiontransformer_b200 as ft
    return ft


@pytest.fixture(scope="module")
def voxels(small_batch):
    """Oracle stride-1 voxels + the same on the GPU."""
    from oracle import ft_glue as og, ts_ops as ts
    z = ts.PointTensor(small_batch["feats"], small_batch["coords"].float())
    x0 = og.initial_voxelize(z, 1, 1)
    return dict(z=z, x0=x0, C=x0.C.cuda())


# ----------------------------------------------------------------------------------------------- hashing
def test_sphash_bit_exact(ft, small_batch):
    from oracle import ts_ops as ts
    c = small_batch["coords"].int()
    c = torch.cat([c, torch.tensor([[0, 0, 0, 0], [-1, 5, 4095, 7], [2 ** 20, -3, 9, 3]], dtype=torch.int32)])
    got = ft.nn.functional.sphash(c.cuda())
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), ts.sphash(c))
    assert got[-3].item() == 0x0D25767F9DCE13F1 and got[-2].item() == 0x0601D450FE50BE7A   # hand-computed KATs
    for ks, stride in ((3, 1), (3, 4), (2, 2), (2, 16)):
        off = ts.KernelRegion(ks, stride).get_kernel_offset()
        off_g = ft.utils.KernelRegion(ks, stride).get_kernel_offset()
        assert torch.equal(off, off_g)
        assert torch.equal(ft.nn.functional.sphash(c.cuda(), off_g.cuda()).cpu(), ts.sphash(c, off))
    assert ft.nn.functional.sphash(torch.zeros(0, 4, dtype=torch.int32).cuda()).shape == (0,)


def test_sphashquery_and_count(ft, voxels):
    from oracle import ts_ops as ts
    spf = ft.nn.functional
    h = ts.sphash(voxels["x0"].C)
    g = torch.Generator().manual_seed(0)
    q = torch.cat([h[torch.randint(0, h.numel(), (5000,), generator=g)],
                   torch.randint(0, 2 ** 60, (5000,), generator=g)]).view(4, 2500)
    want = ts.sphashquery(q, h)
    got = spf.sphashquery(q.cuda(), h.cuda())
    assert got.shape == q.shape and torch.equal(got.cpu(), want)
    assert (want == -1).any() and (want >= 0).any()
    idx = want.view(-1).int()
    assert torch.equal(spf.spcount(idx.cuda(), h.numel()).cpu(), ts.spcount(idx, h.numel()))
    # empty target
    assert torch.all(spf.sphashquery(q.cuda(), h[:0].cuda()) == -1)


# ----------------------------------------------------------------------------------------------- quantize
def test_sparse_quantize_bit_exact_golden(ft):
    gold = np.load(os.path.join(GOLD, "quantize_small.npz"))
    coords = gold["coords"].astype(np.int64)
    inds, labels, inv = ft.utils.sparse_quantize(coords, np.zeros((len(coords), 1), np.float32),
                                                 np.arange(len(coords)), return_index=True, return_invs=True)
    assert isinstance(inds, np.ndarray)
    np.testing.assert_array_equal(inds, gold["inds"])
    np.testing.assert_array_equal(inv, gold["inverse"])
    # index-only and coordinate-returning call forms
    np.testing.assert_array_equal(ft.utils.sparse_quantize(coords), gold["inds"])
    uc = ft.utils.sparse_quantize(coords, np.zeros((len(coords), 1), np.float32))[0]
    np.testing.assert_array_equal(uc, coords[gold["inds"]])


def test_sparse_quantize_collisions_and_edge_cases(ft):
    from oracle import ts_ops as ts
    rng = np.random.default_rng(3)
    c = rng.integers(0, 5, size=(4000, 3))                      # heavy duplication
    lab = rng.integers(0, 20, size=4000)
    oi, ol, ov = ts.sparse_quantize(c, np.zeros((4000, 1), np.float32), lab, return_index=True, return_invs=True)
    gi, gl, gv = ft.utils.sparse_quantize(c, np.zeros((4000, 1), np.float32), lab, return_index=True, return_invs=True)
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gv, ov)
    np.testing.assert_array_equal(gl, ol)                       # ignore_label on multiply-hit voxels
    one = np.array([[7, 8, 9]])
    assert ft.utils.sparse_quantize(one).tolist() == [0]
    gi, gv = ft.utils.sparse_quantize(np.zeros((0, 3), np.int64), return_index=True, return_invs=True)
    assert len(gi) == 0 and len(gv) == 0


def test_batch_quantize_matches_per_scan_reference(ft):
    """Device-side a1+a2+a3 (scale, bounds filter, dedup, batch) == the reference's per-scan numpy pipeline."""
    from fusiontransformer_b200.synthetic import make_batch
    from oracle import ft_glue as og
    scans = make_batch("nuscenes", 3)
    pts = torch.from_numpy(np.concatenate([s["points"] for s in scans])).cuda()
    sid = torch.cat([torch.full((len(s["points"]),), i, dtype=torch.int32) for i, s in enumerate(scans)]).cuda()
    vc, kept, inds, inv, counts = ft.utils.sparse_quantize_batch(pts, sid, len(scans))
    start, ustart = 0, 0
    for i, s in enumerate(scans):
        ovc, okeep, oinds, oinv = og.voxelize_scan(s["points"])
        m = len(ovc)
        np.testing.assert_array_equal(vc[start:start + m, :3].cpu().numpy(), ovc)
        assert torch.all(vc[start:start + m, 3] == i)
        u = int(counts[i])
        assert u == len(oinds)
        np.testing.assert_array_equal(inds[ustart:ustart + u].cpu().numpy() - start, oinds)
        np.testing.assert_array_equal(inv[start:start + m].cpu().numpy(), oinv)
        start += m
        ustart += u


# ----------------------------------------------------------------------------------------------- voxelize glue
def test_initial_voxelize_bit_exact(ft, small_batch, voxels):
    from fusiontransformer_b200.voxel_glue import initial_voxelize
    gold = np.load(os.path.join(GOLD, "kmap_small.npz"))
    z = ft.PointTensor(small_batch["feats"].cuda(), small_batch["coords"].float().cuda())
    x0 = initial_voxelize(z, 1, 1)
    np.testing.assert_array_equal(x0.C.cpu().numpy(), gold["coords"])
    np.testing.assert_array_equal(z.additional_features["idx_query"][1].cpu().numpy(), gold["idx_query"])
    assert torch.equal(x0.C.cpu(), voxels["x0"].C)
    assert rel_l2(x0.F, voxels["x0"].F) < 1e-6
    assert torch.all(z.additional_features["counts"][1] == 1)


def test_reference_glue_on_operator_api_matches_fused(ft, small_batch):
    """The reference's models/utils.py formulation (hash -> unique -> query -> count -> voxelize), restated on the
    operator-level API, gives the same tensors as the fused single-launch builders."""
    import fusiontransformer_b200.voxel_glue as vg
    spf = ft.nn.functional
    feats, coords = small_batch["feats"].cuda(), small_batch["coords"].float().cuda()
    # reference formulation (models/utils.py:15-35)
    pc_hash = spf.sphash(torch.floor(coords).int())
    sparse_hash = torch.unique(pc_hash)
    idx_query = spf.sphashquery(pc_hash, sparse_hash)
    counts = spf.spcount(idx_query.int(), len(sparse_hash))
    ic = torch.round(spf.spvoxelize(torch.floor(coords), idx_query, counts)).int()
    iF = spf.spvoxelize(feats, idx_query, counts)
    z = ft.PointTensor(feats, coords)
    x0 = vg.initial_voxelize(z, 1, 1)
    assert torch.equal(ic, x0.C) and torch.equal(idx_query.int(), z.additional_features["idx_query"][1])
    assert rel_l2(iF, x0.F) < 1e-6
    # voxel_to_point at stride 4 (models/utils.py:71-87)
    c4 = spf.spdownsample(spf.spdownsample(x0.C, 2), 4)
    x4 = ft.SparseTensor(torch.randn(c4.shape[0], 32, device="cuda"), c4, 4)
    off = ft.utils.KernelRegion(2, 4, 1).get_kernel_offset().cuda()
    old_hash = spf.sphash(torch.cat([torch.floor(z.C[:, :3] / 4).int() * 4, z.C[:, -1].int().view(-1, 1)], 1), off)
    iq = spf.sphashquery(old_hash, spf.sphash(x4.C))
    w = spf.calc_ti_weights(z.C, iq, scale=4).transpose(0, 1).contiguous()
    iq = iq.transpose(0, 1).contiguous()
    ref = spf.spdevoxelize(x4.F, iq, w)
    z1 = vg.voxel_to_point(x4, z)
    assert torch.equal(z.idx_query[4].long(), iq)
    assert torch.allclose(z.weights[4], w, atol=1e-6)
    assert rel_l2(z1.F, ref) < 1e-6
    # point_to_voxel at stride 4 (models/utils.py:46-58)
    ph = spf.sphash(torch.cat([torch.floor(z.C[:, :3] / 4).int() * 4, z.C[:, -1].int().view(-1, 1)], 1))
    iq2 = spf.sphashquery(ph, spf.sphash(x4.C))
    cnt2 = spf.spcount(iq2.int(), x4.C.shape[0])
    ref2 = spf.spvoxelize(z1.F, iq2, cnt2)
    got2 = vg.point_to_voxel(x4, z1)
    assert torch.equal(z1.additional_features["idx_query"][4].long(), iq2)
    assert torch.equal(z1.additional_features["counts"][4], cnt2)
    assert rel_l2(got2.F, ref2) < 1e-5


def test_point_voxel_ops_vs_oracle(ft, voxels, small_batch):
    from oracle import ft_glue as og, ts_ops as ts
    spf = ft.nn.functional
    z, x0 = voxels["z"], voxels["x0"]
    c8 = ts.spdownsample(ts.spdownsample(ts.spdownsample(x0.C, 2), 4), 8)
    g = torch.Generator().manual_seed(2)
    x8 = ts.SparseTensor(torch.randn(c8.shape[0], 48, generator=g, requires_grad=True), c8, 8)
    zp = ts.PointTensor(torch.randn(z.C.shape[0], 48, generator=g, requires_grad=True), z.C)
    zo = og.voxel_to_point(x8, zp)
    idx, w = zp.idx_query[8], zp.weights[8]
    zo.F.sum().backward()
    # GPU
    xg = x8.F.detach().cuda().requires_grad_(True)
    got = spf.spdevoxelize(xg, idx.cuda(), w.cuda())
    assert rel_l2(got, zo.F) < 1e-5
    got.sum().backward()
    assert rel_l2(xg.grad, x8.F.grad) < 1e-5
    wg = spf.calc_ti_weights(zp.C.cuda(), idx.t().contiguous().cuda(), scale=8)
    assert torch.allclose(wg.t().cpu(), w, atol=1e-6)
    # voxelize fwd/bwd
    xo = og.point_to_voxel(x8, zp)
    iq, cnt = zp.additional_features["idx_query"][8], zp.additional_features["counts"][8]
    (xo.F * torch.arange(48.0)).sum().backward()
    pg = zp.F.detach().cuda().requires_grad_(True)
    gotv = spf.spvoxelize(pg, iq.cuda(), cnt.cuda())
    assert rel_l2(gotv, xo.F) < 1e-5
    (gotv * torch.arange(48.0, device="cuda")).sum().backward()
    assert rel_l2(pg.grad, zp.F.grad) < 1e-5
    # odd channel count exercises the scalar path
    f5 = torch.randn(z.C.shape[0], 5, generator=g)
    assert rel_l2(spf.spvoxelize(f5.cuda(), iq.cuda(), cnt.cuda()), ts.spvoxelize(f5, iq, cnt)) < 1e-5


# ----------------------------------------------------------------------------------------------- kernel maps
def test_kernel_maps_bit_exact_golden(ft):
    gold = np.load(os.path.join(GOLD, "kmap_small.npz"))
    spf = ft.nn.functional
    C = torch.from_numpy(gold["coords"]).cuda()
    km = spf.build_kernel_map(C, C, 3, 1)
    np.testing.assert_array_equal(km.nbr[:, :27].t().cpu().numpy(), gold["nbr_k3"])
    assert torch.all(km.nbr[:, 27:] == -1)
    np.testing.assert_array_equal(km[0].cpu().numpy(), gold["pairs_k3"])
    np.testing.assert_array_equal(km[1].numpy(), gold["counts_k3"])
    assert km[2] == (C.shape[0], C.shape[0])
    c2 = spf.spdownsample(C, 2)
    np.testing.assert_array_equal(c2.cpu().numpy(), gold["coords_s2"])
    km2 = spf.build_kernel_map(C, c2, 2, 1)
    np.testing.assert_array_equal(km2.nbr.t().cpu().numpy(), gold["nbr_k2"])
    np.testing.assert_array_equal(km2[0].cpu().numpy(), gold["pairs_k2"])
    np.testing.assert_array_equal(km2[1].numpy(), gold["counts_k2"])
    # input-stationary transpose: every fine voxel has exactly one (offset, parent)
    nT = km2.nbrT
    assert nT.shape == (C.shape[0], 8) and torch.all((nT >= 0).sum(1) == 1)
    p = km2[0].long()
    k_of_pair = torch.repeat_interleave(torch.arange(8), km2[1].long()).cuda()
    assert torch.equal(nT[p[:, 0], k_of_pair].long(), p[:, 1])
    # symmetric-map shortcut used by dgrad: transpose of a k3 stride-1 map is its column flip
    assert torch.equal(km.nbrT[:, :27], km.nbr[:, :27].flip(1))


def test_kernel_maps_all_strides_vs_oracle(ft, voxels):
    from oracle import ts_ops as ts
    spf = ft.nn.functional
    co, cg = voxels["x0"].C, voxels["C"]
    for s in (1, 2, 4, 8):
        nbr_o, pairs_o, counts_o = ts.build_kernel_map(co, co, 3, s)
        km = spf.build_kernel_map(cg, cg, 3, s)
        assert torch.equal(km.nbr[:, :27].t().cpu().long(), nbr_o)
        assert torch.equal(km[0].cpu(), pairs_o) and torch.equal(km[1], counts_o)
        co2 = ts.spdownsample(co, 2 * s)
        cg2 = spf.spdownsample(cg, 2 * s)
        assert torch.equal(cg2.cpu(), co2)
        nbr2_o, pairs2_o, counts2_o = ts.build_kernel_map(co, co2, 2, s)
        km2 = spf.build_kernel_map(cg, cg2, 2, s)
        assert torch.equal(km2.nbr.t().cpu().long(), nbr2_o) and torch.equal(km2[0].cpu(), pairs2_o)
        assert km2.num_pairs() == co.shape[0]
        co, cg = co2, cg2


# ----------------------------------------------------------------------------------------------- convolution
def _conv_case(voxels, cin, cout, ks, stride, seed):
    from oracle import ts_ops as ts
    g = torch.Generator().manual_seed(seed)
    C = voxels["x0"].C
    n = C.shape[0]
    feats = torch.randn(n, cin, generator=g)
    w = torch.randn(ks ** 3, cin, cout, generator=g) / (cin * ks ** 3) ** 0.5
    return C, feats, w


@pytest.mark.parametrize("mode,tol", [("f32", 1e-5), ("tc", 5e-3)])
@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 96), (96, 128), (128, 256), (192, 128), (384, 256), (256, 384), (4, 32)])
def test_conv_k3_forward_backward(ft, voxels, monkeypatch, mode, tol, cin, cout):
    from oracle import ts_ops as ts
    monkeypatch.setenv("FT3D_CONV", mode)
    C, feats, w = _conv_case(voxels, cin, cout, 3, 1, cin * 1000 + cout)
    fo = feats.clone().requires_grad_(True)
    wo = w.clone().requires_grad_(True)
    xo = ts.SparseTensor(fo, C, 1)
    yo = ts.conv3d(xo, wo, 3)
    gsel = torch.randn(yo.F.shape, generator=torch.Generator().manual_seed(9))
    (yo.F * gsel).sum().backward()
    fg = feats.cuda().requires_grad_(True)
    wg = torch.nn.Parameter(w.cuda())
    xg = ft.SparseTensor(fg, C.cuda(), 1)
    yg = ft.nn.functional.conv3d(xg, wg, 3)
    assert rel_l2(yg.F, yo.F) < tol
    (yg.F * gsel.cuda()).sum().backward()
    assert rel_l2(fg.grad, fo.grad) < tol
    assert rel_l2(wg.grad, wo.grad) < tol
    # per-offset weight gradients individually (catches offset permutation bugs hidden by the global norm)
    for k in (0, 13, 26):
        assert rel_l2(wg.grad[k], wo.grad[k]) < 4 * tol


@pytest.mark.parametrize("mode,tol", [("f32", 1e-5), ("tc", 5e-3)])
def test_conv_down_up_forward_backward(ft, voxels, monkeypatch, mode, tol):
    from oracle import ts_ops as ts
    monkeypatch.setenv("FT3D_CONV", mode)
    C, feats, w = _conv_case(voxels, 32, 64, 2, 2, 5)
    g = torch.Generator().manual_seed(6)
    wt = torch.randn(8, 64, 96, generator=g) / 8.0
    fo, wo, wto = feats.clone().requires_grad_(True), w.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    xo = ts.SparseTensor(fo, C, 1)
    xo.check()
    yo = ts.conv3d(xo, wo, 2, stride=2)
    zo = ts.conv3d(yo, wto, 2, stride=2, transpose=True)
    gsel = torch.randn(zo.F.shape, generator=g)
    (zo.F * gsel).sum().backward()
    fg, wg, wtg = feats.cuda().requires_grad_(True), torch.nn.Parameter(w.cuda()), torch.nn.Parameter(wt.cuda())
    xg = ft.SparseTensor(fg, C.cuda(), 1)
    xg.check()
    yg = ft.nn.functional.conv3d(xg, wg, 2, stride=2)
    assert yg.s == 2 and torch.equal(yg.C.cpu(), yo.C)
    zg = ft.nn.functional.conv3d(yg, wtg, 2, stride=2, transpose=True)
    assert zg.s == 1 and torch.equal(zg.C.cpu(), C)
    assert rel_l2(yg.F, yo.F) < tol and rel_l2(zg.F, zo.F) < 2 * tol
    (zg.F * gsel.cuda()).sum().backward()
    assert rel_l2(wtg.grad, wto.grad) < 2 * tol
    assert rel_l2(wg.grad, wo.grad) < 3 * tol
    assert rel_l2(fg.grad, fo.grad) < 3 * tol


def test_conv_k1_and_ragged_sizes(ft, monkeypatch):
    """k=1 conv == matmul; tiles that are not a multiple of 128 rows; single-voxel and empty inputs."""
    from oracle import ts_ops as ts
    for mode, tol in (("f32", 1e-5), ("tc", 5e-3)):
        monkeypatch.setenv("FT3D_CONV", mode)
        for n in (1, 127, 129, 300):
            g = torch.Generator().manual_seed(n)
            lin = torch.randperm(12 ** 3, generator=g)[:n]
            C = torch.stack([lin // 144, (lin // 12) % 12, lin % 12, torch.zeros_like(lin)], 1).int()
            feats = torch.randn(C.shape[0], 64, generator=g)
            w = torch.randn(27, 64, 64, generator=g) / 40
            yo = ts.conv3d(ts.SparseTensor(feats, C, 1), w, 3)
            yg = ft.nn.functional.conv3d(ft.SparseTensor(feats.cuda(), C.cuda(), 1), w.cuda(), 3)
            assert rel_l2(yg.F, yo.F) < tol, (mode, n)
        w1 = torch.randn(64, 32)
        y1 = ft.nn.functional.conv3d(ft.SparseTensor(feats.cuda(), C.cuda(), 1), w1.cuda(), 1)
        assert rel_l2(y1.F, feats @ w1) < tol          # tc mode: bf16 operands on the dense tcgen05 path


# ----------------------------------------------------------------------------------------------- lift
def test_lift_forward_backward(ft):
    from oracle import ft_glue as og
    g = torch.Generator().manual_seed(0)
    B, Cc, H, W = 2, 96, 37, 123
    fmap = torch.randn(B, Cc, H, W, generator=g)
    idx = [torch.stack([torch.randint(0, H, (n,), generator=g), torch.randint(0, W, (n,), generator=g)], 1).numpy()
           for n in (700, 333)]
    fo = fmap.clone().requires_grad_(True)
    want = og.lift(fo, idx)
    gsel = torch.randn(want.shape, generator=g)
    (want * gsel).sum().backward()
    for cl in (False, True):
        fg = fmap.cuda()
        if cl:
            fg = fg.contiguous(memory_format=torch.channels_last)
        fg.requires_grad_(True)
        got = ft.nn.functional.lift(fg, idx)
        assert torch.equal(got.cpu(), want.detach())                  # a pure gather: bit-exact
        (got * gsel.cuda()).sum().backward()
        assert rel_l2(fg.grad, fo.grad) < 1e-6                        # duplicate pixels accumulate
    # odd channel count -> scalar path
    f7 = torch.randn(1, 7, 5, 6, generator=g)
    i7 = [np.array([[0, 0], [4, 5], [2, 3]])]
    assert torch.equal(ft.nn.functional.lift(f7.cuda(), i7).cpu(), og.lift(f7, i7))
