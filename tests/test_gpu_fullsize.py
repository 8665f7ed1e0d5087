"""BASELINE.json's full sizes (batch 8 of nuScenes-shaped scans; one 120k-point stress scan) through properties that
do not need the oracle to finish: dedup round trips and idempotence, kernel-map symmetries, convolution linearity and
dense-GEMM equivalence of the centre offset, lift/scatter adjointness.  Integer properties are exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def big():
    """Device-side voxelization of a full bench batch + of one stress scan."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.synthetic import make_scan
    out = {}
    for name, shape, nscan in (("bench", "nuscenes", 8), ("stress", "stress", 1)):
        scans = [make_scan(shape, 100 + i) for i in range(nscan)]
        db = dataflow.to_device(dataflow.host_batch_from_scans(scans), torch.device("cuda", 0))
        lidar, rc, bidx, labels, inv, kept = dataflow.voxelize_batch(db)
        out[name] = dict(scans=scans, db=db, lidar=lidar, rc=rc, bidx=bidx, inv=inv, kept=kept)
    return out


@pytest.mark.parametrize("name", ["bench", "stress"])
def test_quantize_round_trip_and_idempotence(big, name):
    import fusiontransformer_b200 as ft
    b = big[name]
    C = b["lidar"].C                                       # [u,4] unique voxels (x,y,z,batch)
    u = C.shape[0]
    assert u > (40000 if name == "bench" else 60000)      # 120k points of a stress scan fall into ~70k voxels
    # no duplicates: the 4-tuple is unique
    key = ((C[:, 3].long() * 4096 + C[:, 0].long()) * 4096 + C[:, 1].long()) * 4096 + C[:, 2].long()
    assert torch.unique(key).numel() == u
    # inverse map reconstructs every kept point's voxel: voxel(point) == C[offset(scan) + inverse]
    pts, sid = b["db"].points, b["db"].scan_id
    vc, kept, inds, inv, counts = ft.utils.sparse_quantize_batch(pts, sid.int(), len(b["scans"]))
    ustart = torch.cumsum(counts, 0) - counts
    row = ustart.to(vc.device)[vc[:, 3].long()] + inv.to(vc.device).long()
    assert torch.equal(C[row][:, :3].int(), vc[:, :3].int())
    # `inds` are first occurrences: voxel coordinates at inds equal the unique list, in order
    assert torch.equal(vc[inds.long()][:, :3].int(), C[:, :3].int())
    # idempotence: quantizing the unique voxels of one scan again is the identity up to the reference's key order
    one = C[C[:, 3] == 0][:, :3].cpu().numpy().astype(np.int64)
    i2, v2 = ft.utils.sparse_quantize(one, return_index=True, return_invs=True)
    assert len(i2) == len(one) and np.array_equal(np.sort(i2), np.arange(len(one)))
    assert np.array_equal(one[i2][v2], one)


@pytest.mark.parametrize("name", ["bench", "stress"])
def test_kernel_map_symmetries_at_full_size(big, name):
    import fusiontransformer_b200 as ft
    spf = ft.nn.functional
    C = big[name]["lidar"].C
    for s in (1, 2, 4, 8):
        n = C.shape[0]
        km = spf.build_kernel_map(C, C, 3, s)
        nbr = km.nbr[:, :27]
        pairs, counts = km[0].long(), km[1].long()
        assert counts.sum().item() == pairs.shape[0] == km.num_pairs() == int((nbr >= 0).sum())
        # centre offset is the identity map
        assert torch.equal(nbr[:, 13].long().cpu(), torch.arange(n))
        # pair (i -> j) at offset k  <=>  pair (j -> i) at offset 26 - k
        rows = torch.arange(n, device=nbr.device).view(-1, 1).expand(-1, 27)
        valid = nbr >= 0
        kk = torch.arange(27, device=nbr.device).view(1, -1).expand(n, -1)
        i, j, k = nbr[valid].long(), rows[valid], kk[valid]
        assert torch.equal(nbr[i, 26 - k].long(), j)
        # each output appears at most once per offset, pairs are offset-major and out-ascending inside an offset
        off = torch.cumsum(counts, 0) - counts
        kp = torch.repeat_interleave(torch.arange(27), counts).to(pairs.device)
        assert torch.equal(nbr[pairs[:, 1], kp].long(), pairs[:, 0])
        seg_start = torch.zeros(pairs.shape[0], dtype=torch.bool)
        seg_start[off[counts > 0]] = True
        d = (pairs[1:, 1] - pairs[:-1, 1]).cpu()
        assert bool(torch.all((d > 0) | seg_start[1:]))
        # stride-2 downsample: every fine voxel has exactly one parent
        C2 = spf.spdownsample(C, 2 * s)
        km2 = spf.build_kernel_map(C, C2, 2, s)
        assert km2.num_pairs() == n
        assert torch.equal(torch.sort(km2[0][:, 0].long()).values.cpu(), torch.arange(n))
        C = C2


@pytest.mark.parametrize("mode,tol", [("f32", 2e-5), ("tc", 5e-3)])
def test_conv_linearity_and_centre_offset_at_full_size(big, monkeypatch, mode, tol):
    import fusiontransformer_b200 as ft
    monkeypatch.setenv("FT3D_CONV", mode)
    spf = ft.nn.functional
    C = big["bench"]["lidar"].C
    n = C.shape[0]
    g = torch.Generator(device="cuda").manual_seed(5)
    cin, cout = 64, 96
    x1 = torch.randn(n, cin, device="cuda", generator=g)
    x2 = torch.randn(n, cin, device="cuda", generator=g)
    w = torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    conv = lambda x, ww: spf.conv3d(ft.SparseTensor(x, C), ww, 3).F
    y1, y2, y12 = conv(x1, w), conv(x2, w), conv(x1 + 2.0 * x2, w)
    assert rel_l2(y12, y1 + 2.0 * y2) < (tol if mode == "f32" else 1e-2)      # bf16 rounds x1 + 2 x2 once more
    # with only the centre offset non-zero the sparse conv is a dense GEMM over the rows
    wc = torch.zeros_like(w)
    wc[13] = w[13]
    ref = x1.double() @ w[13].double() if mode == "f32" else \
        x1.bfloat16().double() @ w[13].bfloat16().double()
    assert rel_l2(conv(x1, wc), ref) < tol
    # weight linearity is exact in both modes up to fp32 summation: conv(x, a W) == a conv(x, W) for a power of two
    assert torch.equal(conv(x1, 4.0 * w), 4.0 * y1)


def test_lift_and_its_adjoint_at_full_size(big):
    import fusiontransformer_b200 as ft
    b = big["bench"]
    H, W = b["scans"][0]["image_size"]
    B = len(b["scans"])
    n = b["rc"].shape[0]
    g = torch.Generator(device="cuda").manual_seed(9)
    fmap = torch.randn(B, 96, H, W, device="cuda", generator=g).requires_grad_(True)
    out = ft.nn.functional.lift(fmap, b["rc"], b["bidx"])
    rc, bi = b["rc"].long(), b["bidx"].long()
    assert torch.equal(out, fmap.detach().permute(0, 2, 3, 1)[bi, rc[:, 0], rc[:, 1]])       # pure gather: exact
    # <lift(X), G> == <X, lift^T(G)>  (adjointness of forward and backward)
    G = torch.randn(n, 96, device="cuda", generator=g)
    (out * G).sum().backward()
    lhs = (out.detach().double() * G.double()).sum()
    rhs = (fmap.detach().double() * fmap.grad.double()).sum()
    assert abs(lhs - rhs).item() < 1e-6 * abs(lhs).item() + 1e-3
    # ones scattered back count the points of every pixel: the total is n * 96
    fmap.grad = None
    ft.nn.functional.lift(fmap, b["rc"], b["bidx"]).sum().backward()
    assert fmap.grad.sum().item() == n * 96
