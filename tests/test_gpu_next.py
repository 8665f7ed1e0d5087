"""SURVEY 8(f) "next" rows: lift straight from the low-resolution image-feature map (rank 2), loss and IoU metric
on the device (rank 4).  Oracles are the reference's own expressions in plain torch (they are executable here:
image_models_billinear.py:8-23,117-124; SemanticTorchpackTrainer.py:70-108; metric.py:37-58)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------ host logic (CPU)
@pytest.mark.parametrize("inp,out", [(24, 370), (24, 1226), (24, 900), (40, 1600), (24, 384), (7, 5), (1, 9)])
def test_nearest_source_index_is_atens_rule(inp, out):
    """Index map of nn.Upsample(size) (legacy 'nearest') recovered by upsampling an index ramp."""
    from fusiontransformer_b200.image_lift import nearest_source_index
    ramp = torch.arange(inp, dtype=torch.float32).view(1, 1, inp, 1)
    want = torch.nn.Upsample((out, 1))(ramp).view(-1).long()
    got = nearest_source_index(torch.arange(out), inp, out)
    assert torch.equal(got, want)


# ------------------------------------------------------------------ lift from the low-resolution map
@pytest.mark.gpu
@pytest.mark.parametrize("hw,HW", [((24, 24), (370, 1226)), ((24, 40), (900, 1600))])
def test_lift_nearest_equals_lift_of_upsampled_map(hw, HW):
    from fusiontransformer_b200.image_lift import lift_nearest
    torch.manual_seed(0)
    B, C = 3, 96
    src = torch.randn(B, C, *hw)
    rng = np.random.default_rng(1)
    idx = [np.stack([rng.integers(0, HW[0], n), rng.integers(0, HW[1], n)], 1) for n in (700, 1, 1300)]
    idx[0][:4] = [[0, 0], [HW[0] - 1, HW[1] - 1], [0, HW[1] - 1], [HW[0] - 1, 0]]
    # reference: materialise, permute, index per sample, concatenate (image_models_billinear.py:117-124)
    xr = src.clone().requires_grad_(True)
    up = torch.nn.Upsample(HW)(xr).permute(0, 2, 3, 1)
    ref = torch.cat([up[i][idx[i][:, 0], idx[i][:, 1]] for i in range(B)], 0)
    xg = src.cuda().requires_grad_(True)
    got = lift_nearest(xg, idx, HW)
    assert torch.equal(got.cpu(), ref.detach())             # a pure gather: bit-exact
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w.cuda()).sum().backward()
    assert rel_l2(xg.grad, xr.grad) < 1e-6


@pytest.mark.gpu
def test_bilinear_lift_head_matches_reference_module():
    from fusiontransformer_b200.image_lift import BilinearLiftHead
    torch.manual_seed(0)
    head = BilinearLiftHead(768, 96, (370, 1226)).cuda().train()
    assert set(head.state_dict()) == {"stem.0.weight", "stem.0.bias", "stem.2.weight", "stem.2.bias",
                                      "stem.2.running_mean", "stem.2.running_var", "stem.2.num_batches_tracked"}
    x = torch.randn(2, 768, 24, 24, device="cuda")
    rng = np.random.default_rng(0)
    idx = [np.stack([rng.integers(0, 370, 500), rng.integers(0, 1226, 500)], 1) for _ in range(2)]
    with torch.no_grad():
        head.eval()
        full = head(x).permute(0, 2, 3, 1)
        ref = torch.cat([full[i][idx[i][:, 0], idx[i][:, 1]] for i in range(2)], 0)
        got = head.lift(x, idx)
    assert torch.equal(got, ref)


# ------------------------------------------------------------------ loss + metric on the device
def _torch_loss(logits, labels, weight, teacher, lam):
    ce = F.cross_entropy(logits, labels, weight=weight)
    if lam == 0:
        return ce
    kl = F.kl_div(F.log_softmax(logits, dim=1), F.softmax(teacher.detach(), dim=1), reduction="none").sum(1).mean()
    return (1 - lam) * ce + lam * kl


@pytest.mark.gpu
@pytest.mark.parametrize("weighted,lam,ignored", [(False, 0.0, False), (True, 0.0, True), (True, 0.3, True),
                                                  (False, 1.0, False)])
def test_seg_loss_matches_torch(weighted, lam, ignored):
    from fusiontransformer_b200.losses import seg_loss
    torch.manual_seed(2)
    n, c = 20011, 20
    logits = (3 * torch.randn(n, c, dtype=torch.float64)).requires_grad_(True)
    teacher = 3 * torch.randn(n, c, dtype=torch.float64)
    labels = torch.randint(0, c, (n,))
    if ignored:
        labels[::7] = -100
    weight = torch.rand(c, dtype=torch.float64) + 0.5 if weighted else None
    want = _torch_loss(logits, labels, weight, teacher, lam)
    want.backward()
    lg = logits.detach().float().cuda().requires_grad_(True)
    got = seg_loss(lg, labels.cuda(), None if weight is None else weight.float().cuda(),
                   teacher.float().cuda() if lam > 0 else None, lam)
    (2.0 * got).backward()                                   # a non-unit upstream gradient
    assert got.dim() == 0 and abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel_l2(lg.grad, 2.0 * logits.grad) < 1e-5


@pytest.mark.gpu
def test_seg_loss_all_rows_ignored_is_nan_like_torch():
    from fusiontransformer_b200.losses import seg_loss
    lg = torch.randn(64, 20, device="cuda")
    labels = torch.full((64,), -100, device="cuda")
    assert torch.isnan(seg_loss(lg, labels)) and torch.isnan(F.cross_entropy(lg, labels))


@pytest.mark.gpu
def test_seg_iou_matches_reference_metric():
    from fusiontransformer_b200.losses import SegIoU
    torch.manual_seed(3)
    n, c = 30000, 20
    m = SegIoU(c, ignore_index=0, name="seg_iou_3d")
    mat = torch.zeros(c, c, dtype=torch.int64)
    for step in range(3):
        logits = torch.randn(n + step, c)
        labels = torch.randint(0, c, (n + step,))
        m.update_dict({"lidar_seg_logit": logits.cuda()}, {"seg_label": labels.cuda()})
        keep = labels != 0                                   # metric.py:46-51
        inds = c * labels[keep] + logits.argmax(1)[keep]
        mat += torch.bincount(inds, minlength=c * c).reshape(c, c)
    assert torch.equal(m.mat.cpu(), mat)
    h = mat.float()
    iou = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
    assert torch.allclose(m.iou.cpu(), iou, equal_nan=True)
    assert abs(m.global_avg - iou.mean().item()) < 1e-6 or (np.isnan(m.global_avg) and torch.isnan(iou.mean()))


# ------------------------------------------------------------------ vectors produced by the reference's own code
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["nus", "kitti"])
def test_device_voxelization_matches_reference_augment_and_scale(tag):
    """a1 on the GPU (ft3d_scale_coords) against augmentation_3d.py + dataloader :220,:225 run by the reference."""
    import fusiontransformer_b200 as ft
    g = np.load(os.path.join(GOLD, "ref_voxelize.npz"))
    pts = torch.from_numpy(g[tag + "_points"]).cuda()
    sid = torch.zeros(len(pts), dtype=torch.int32, device="cuda")
    vc, kept, inds, inv, counts = ft.utils.sparse_quantize_batch(pts, sid, 1)
    keep = g[tag + "_keep"]
    np.testing.assert_array_equal(kept.cpu().numpy(), np.nonzero(keep)[0])
    np.testing.assert_array_equal(vc[:, :3].cpu().numpy(), g[tag + "_coords"][keep])
    np.testing.assert_array_equal(vc[inds.long()][inv.long()][:, :3].cpu().numpy(), g[tag + "_coords"][keep])


@pytest.mark.gpu
def test_device_augmentation_matches_reference_augment_and_scale():
    """a1 with the augmentation branch on, on the GPU: four scans with four different parameter sets in ONE batch
    (ft3d_augment_scale_coords with the host-side draws of utils/augment.py) against the reference function run on each
    scan under the same seed -- voxel coordinates and bounds mask bit-exact."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import ops
    from fusiontransformer_b200.utils import augment
    g = np.load(os.path.join(GOLD, "ref_augment.npz"))
    tags = ["a", "b", "c", "d"]
    pts, sid, rots, us = [], [], [], []
    for i, tag in enumerate(tags):
        kw = dict(noisy_rot=float(g[tag + "_noisy_rot"]), flip_x=float(g[tag + "_flip_x"]), flip_y=float(g[tag + "_flip_y"]),
                  rot_z=float(g[tag + "_rot_z"]), transl=bool(g[tag + "_transl"]))
        np.random.seed(int(g[tag + "_seed"]))
        rot, u = augment.draw(**kw)
        rots.append(np.eye(3, dtype=np.float32) if rot is None else rot)          # identity: exact under the fma chain
        us.append(np.zeros(3) if u is None else u)                                # zero factor: no translation
        pts.append(g[tag + "_points"])
        sid.append(np.full(len(pts[-1]), i, np.int32))
    points = torch.from_numpy(np.concatenate(pts)).cuda()
    scan_id = torch.from_numpy(np.concatenate(sid)).cuda()
    rot = torch.from_numpy(np.stack(rots)).cuda()
    tu = torch.from_numpy(np.stack(us)).cuda()
    coords, keep = ops.scale_coords(points, scan_id, len(tags), 20.0, 4096, rot=rot, transl_u=tu)
    coords, keep = coords.cpu().numpy(), keep.cpu().numpy()
    o = 0
    for i, tag in enumerate(tags):
        n = len(pts[i])
        np.testing.assert_array_equal(coords[o:o + n, :3], g[tag + "_coords"], err_msg=tag)
        np.testing.assert_array_equal(coords[o:o + n, 3], i)
        np.testing.assert_array_equal(keep[o:o + n], g[tag + "_keep"], err_msg=tag)
        o += n
    # and through the batch API: same coordinates after the bounds filter
    vc, kept, inds, inv, counts = ft.utils.sparse_quantize_batch(points, scan_id, len(tags), rot=rot, transl_u=tu)
    np.testing.assert_array_equal(vc.cpu().numpy(), coords[keep])


@pytest.mark.gpu
def test_device_seg_iou_matches_reference_metric_fixture():
    """(f)4: losses.SegIoU against FusionTransformer/models/metric.py:SegIoU run by the reference on the same data."""
    from fusiontransformer_b200.losses import SegIoU
    g = np.load(os.path.join(GOLD, "ref_segiou.npz"))
    m = SegIoU(20, ignore_index=0, name="seg_iou_3d")
    for step in range(2):
        m.update_dict({"lidar_seg_logit": torch.from_numpy(g["logits%d" % step]).cuda()},
                      {"seg_label": torch.from_numpy(g["labels%d" % step]).cuda()})
    np.testing.assert_array_equal(m.mat.cpu().numpy(), g["mat"])
    np.testing.assert_allclose(m.iou.cpu().numpy(), g["iou"], rtol=1e-6, equal_nan=True)
    pred = torch.from_numpy(g["logits1"]).cuda().argmax(1)
    np.testing.assert_array_equal(pred[torch.from_numpy(g["inverse_map"]).cuda()].cpu().numpy(), g["pred_points"])


@pytest.mark.gpu
def test_device_batch_path_matches_reference_collate_fixture():
    """a3 on the GPU: dataflow.voxelize_batch (scale + bounds + dedup + select + batch column for the whole batch in
    two libft3d calls) against the tensors FusionTransformer/data/collate.py:6-86 `collate_scn_base` produced from the
    per-scan CPU pipeline (tests/golden/ref_collate.npz, generated by running the reference's collate): coordinates,
    features and voxel labels of the batch, row for row."""
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.synthetic import make_scan
    g = np.load(os.path.join(GOLD, "ref_collate.npz"))
    scans = [make_scan("nuscenes", 20 + i) for i in range(3)]
    db = dataflow.to_device(dataflow.host_batch_from_scans(scans), "cuda")
    lidar, rc, bidx, labels, inv, kept = dataflow.voxelize_batch(db)
    np.testing.assert_array_equal(lidar.C.cpu().numpy(), g["C"])
    np.testing.assert_array_equal(lidar.F.cpu().numpy(), g["F"])
    np.testing.assert_array_equal(labels.cpu().numpy(), g["seg_label"])
    np.testing.assert_array_equal(bidx.cpu().numpy(), g["C"][:, 3])
    off = 0
    for i in range(3):                       # per-scan pieces of the fixture, in batch order
        n = len(g["coords%d" % i])
        np.testing.assert_array_equal(lidar.C[off:off + n, :3].cpu().numpy(), g["coords%d" % i])
        off += n
    assert off == lidar.C.shape[0]
