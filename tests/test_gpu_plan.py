"""GeometryPlan (plan.py): geometry built ahead of time == geometry built lazily inside the forward pass."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_plan_maps_equal_lazy_maps(small_batch):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.plan import build_plan
    from fusiontransformer_b200.spvcnn import Net3DSeg
    coords, feats = small_batch["coords"].cuda(), small_batch["feats"].cuda()
    plan = build_plan(coords)
    torch.manual_seed(1)
    net = Net3DSeg(fusion="none").cuda().eval()
    taps = {}
    with torch.no_grad():
        net(ft.SparseTensor(feats, coords), taps=taps)
    for s, name in ((1, "x0"), (2, "x1"), (4, "x2"), (8, "x3"), (16, "x4")):
        assert torch.equal(plan.coord_maps[s], taps[name].C), s
    lazy = taps["x4"].kernel_maps
    assert set(plan.kernel_maps) == set(lazy)
    for key, km in plan.kernel_maps.items():
        assert torch.equal(km.nbr, lazy[key].nbr), key
        assert torch.equal(km[0], lazy[key][0]) and torch.equal(km[1], lazy[key][1]) and km[2] == lazy[key][2], key
    z = taps["z3"]
    for s, (idx, w) in plan.v2p.items():
        assert torch.equal(idx, z.idx_query[s]) and torch.equal(w, z.weights[s]), s
    for s, (idx, cnt) in plan.p2v.items():
        assert torch.equal(idx, z.additional_features["idx_query"][s]), s
        assert torch.equal(cnt, z.additional_features["counts"][s]), s


@pytest.mark.parametrize("fusion", ["none", "middle"])
def test_planned_forward_backward_equals_lazy(monkeypatch, small_batch, fusion):
    """Same kernels, same maps: only the order of the fp32 atomics (point->voxel averaging, wgrad) differs between
    two runs, so the comparison is made in the exact-precision mode where that noise is not amplified by bf16
    rounding flips (DESIGN.md "Tolerances")."""
    import fusiontransformer_b200 as ft
    monkeypatch.setenv("FT3D_CONV", "f32")
    from fusiontransformer_b200.plan import Prefetcher, build_plan
    from fusiontransformer_b200.spvcnn import Net3DSeg
    coords, feats = small_batch["coords"].cuda(), small_batch["feats"].cuda()
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 96, generator=g).cuda() if fusion != "none" else None
    labels = torch.randint(0, 20, (n,), generator=g).cuda()
    torch.manual_seed(1)
    net = Net3DSeg(fusion=fusion).cuda().train()
    net.dropout.p = 0.0
    pre = Prefetcher()
    pre.submit(build_plan, coords)
    plan = pre.get()                                   # built on the side stream, consumed on the current one
    res = []
    for p in (None, plan):
        net.zero_grad(set_to_none=True)
        out = net(ft.SparseTensor(feats, coords), img, plan=p)["lidar_seg_logit"]
        torch.nn.functional.cross_entropy(out, labels).backward()
        res.append((out.detach().clone(), [q.grad.clone() for q in net.parameters() if q.grad is not None]))
    assert (res[0][0] - res[1][0]).norm() <= 1e-3 * res[1][0].norm()
    gmax = max(b.norm().item() for b in res[1][1])
    for a, b in zip(res[0][1], res[1][1]):
        assert (a - b).norm().item() <= 2e-2 * max(b.norm().item(), 1e-4 * gmax)
