"""Host-side logic that needs no GPU: containers, alias installation, KernelRegion, model key parity."""
import os
import sys

import numpy as np
import pytest
import torch


def test_alias_exposes_reference_import_surface():
    import fusiontransformer_b200 as ft
    ft.install_as_torchsparse()
    import torchsparse
    import torchsparse.nn as spnn
    import torchsparse.nn.functional as spf
    from torchsparse.point_tensor import PointTensor
    from torchsparse.sparse_tensor import SparseTensor
    from torchsparse.utils import sparse_quantize
    ns = {}
    exec("from torchsparse.utils.kernel_region import *\nfrom torchsparse.utils.helpers import *", ns)
    assert "torch" in ns and "KernelRegion" in ns       # FusionTransformer/models/utils.py relies on the leaked torch
    for name in ("sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights", "conv3d"):
        assert hasattr(spf, name)
    for name in ("Conv3d", "BatchNorm", "ReLU"):
        assert hasattr(spnn, name)
    assert callable(torchsparse.cat) and callable(sparse_quantize)
    st = SparseTensor(coords=torch.zeros(3, 4, dtype=torch.int32), feats=torch.zeros(3, 4))
    st.check()
    assert st.s == 1 and 1 in st.coord_maps
    pt = PointTensor(st.F, st.C.float())
    assert pt.additional_features == {"idx_query": {}, "counts": {}}


@pytest.mark.skipif(not os.path.exists("/root/reference/FusionTransformer/models/spvcnn.py"), reason="reference tree absent")
def test_reference_spvcnn_builds_on_alias_with_same_state_dict():
    import fusiontransformer_b200 as ft
    ft.install_as_torchsparse()
    sys.path.insert(0, "/root/reference")
    try:
        from FusionTransformer.models.spvcnn import SPVCNN as Ref
    finally:
        sys.path.pop(0)
    from fusiontransformer_b200.spvcnn import SPVCNN
    ref, ours = Ref(), SPVCNN()
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys())
    assert all(sr[k].shape == so[k].shape for k in sr)
    assert sum(p.numel() for p in ours.parameters()) == 21776160


def test_kernel_region_matches_oracle():
    from fusiontransformer_b200.utils import KernelRegion
    from oracle import ts_ops as ts
    for ks in (1, 2, 3):
        for s in (1, 2, 8):
            assert torch.equal(KernelRegion(ks, s).get_kernel_offset(), ts.KernelRegion(ks, s).get_kernel_offset())


def test_conv3d_parameter_layout():
    from fusiontransformer_b200 import nn as spnn
    assert spnn.Conv3d(32, 64, 3).kernel.shape == (27, 32, 64)
    assert spnn.Conv3d(32, 64, 2, stride=2, transpose=True).kernel.shape == (8, 32, 64)
    assert spnn.Conv3d(32, 64, 1).kernel.shape == (32, 64)
    c = spnn.Conv3d(16, 8, 3)
    assert c.kernel.abs().max() <= 1.0 / (16 * 27) ** 0.5 + 1e-7


def test_graph_capacities_only_grow_and_stay_distinct():
    """graph.capacity: head-room, granule, monotone across re-captures (the KITTI-shaped batches used to re-capture on
    most steps because a capacity could shrink), unique keys for ops.ROW_COUNTS."""
    from fusiontransformer_b200.graph import capacity
    assert capacity(1000, 1.10, 128) == 1152 and capacity(1000, 1.10, 128) % 128 == 0
    assert capacity(900, 1.10, 128, floor=1152) == 1152            # a smaller batch does not shrink the buffers
    assert capacity(2000, 1.10, 128, floor=1152) == 2304
    taken = set()
    caps = [capacity(n, 1.10, 128, 0, taken) for n in (1000, 1001, 1002, 5000)]
    assert len(set(caps)) == 4 and caps[:3] == [1152, 1280, 1408]
    # a wandering stream converges: after each size has been seen once nothing grows any more
    sizes = [(50000, 20000), (47000, 22000), (52000, 19000), (49000, 21500)] * 3
    caps, grows = (0, 0), 0
    for a, b in sizes:
        if a > caps[0] or b > caps[1]:                             # StaticGeometry.fits
            caps = (capacity(a, 1.10, 128, caps[0]), capacity(b, 1.10, 128, caps[1]))
            grows += 1
    assert grows <= 2


def test_synthetic_scans_follow_the_survey_shapes_and_are_deterministic():
    """SURVEY 8(d) generator: KITTI-shaped ~18-22k front-camera points in a 1226x370 image, nuScenes-shaped ~6-9k in
    1600x900, stress 120k +-2 %; pixel indices strictly inside the image; same seed -> same scan."""
    import numpy as np
    from fusiontransformer_b200.synthetic import make_scan
    want = {"kitti": (18000, 22000, (370, 1226)), "nuscenes": (6000, 9000, (900, 1600)),
            "stress": (117600, 122400, (370, 1226))}
    for shape, (lo, hi, hw) in want.items():
        a, b, c = make_scan(shape, 7), make_scan(shape, 7), make_scan(shape, 8)
        n = len(a["points"])
        assert lo <= n <= hi and tuple(a["image_size"]) == hw
        assert a["points"].dtype == np.float32 and a["feats"].shape == (n, 4) and a["points_img"].shape == (n, 2)
        assert a["points_img"].min() >= 0 and a["points_img"][:, 0].max() < hw[0] and a["points_img"][:, 1].max() < hw[1]
        assert np.array_equal(a["feats"][:, :3], a["points"])              # feats = (x, y, z, intensity)
        assert 0 <= a["seg_labels"].min() and a["seg_labels"].max() < 20
        assert all(np.array_equal(a[k], b[k]) for k in ("points", "feats", "points_img", "seg_labels"))
        assert not np.array_equal(a["points"], c["points"])
        if shape == "kitti":
            assert a["points"][:, 0].min() > 0                             # front camera: x > 0
