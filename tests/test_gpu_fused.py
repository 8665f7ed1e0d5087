"""Fused conv + BatchNorm + ReLU (+ shortcut) blocks (csrc/bn.cu, fused.py) vs the oracle's module-by-module chain.

The oracle executes spvcnn.py:22-79 literally (Conv3d, nn.BatchNorm1d, ReLU as separate modules); the product runs
each chain as one autograd node.  f32 mode is held to 1e-4 (fp32 summation order only), tc mode to the per-layer
5e-3 of the north star with the oracle computing on bf16-rounded conv operands (oracle.ts_ops.OPERAND_DTYPE).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def voxels(small_batch):
    from oracle import ft_glue as og, ts_ops as ts
    z = ts.PointTensor(small_batch["feats"], small_batch["coords"].float())
    return og.initial_voxelize(z, 1, 1)


def _mirror(block_o, build):
    """The product's module with the oracle's parameters."""
    m = build()
    m.load_state_dict(block_o.state_dict())
    return m.cuda()


def _check_grads(mo, mg, tol, what):
    for (name, po), (_, pg) in zip(mo.named_parameters(), mg.named_parameters()):
        assert pg.grad is not None, (what, name)
        scale = max(po.grad.norm().item(), 1e-6)
        err = (pg.grad.double().cpu() - po.grad.double()).norm().item() / scale
        assert err < tol, (what, name, err)


def _check_buffers(mo, mg, tol, what):
    bo, bg = dict(mo.named_buffers()), dict(mg.state_dict())
    for name, b in bo.items():
        if name.endswith("num_batches_tracked"):
            assert int(bg[name]) == int(b), (what, name)
        else:
            assert rel_l2(bg[name], b) < tol, (what, name)


CASES = [  # (kind, cin, cout)
    ("k3", 32, 64), ("k3", 96, 96), ("k3", 4, 32), ("down", 32, 32), ("res_identity", 64, 64), ("res_proj", 64, 128),
]


@pytest.mark.parametrize("mode,tol", [("f32", 1e-4), ("tc", 5e-3)])
@pytest.mark.parametrize("kind,cin,cout", CASES)
def test_fused_block_matches_module_chain(monkeypatch, voxels, mode, tol, kind, cin, cout):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import spvcnn as sp
    from oracle import ft_glue as og, ts_ops as ts
    monkeypatch.setenv("FT3D_CONV", mode)
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16" if mode == "tc" else None)
    torch.manual_seed(cin * 7 + cout)
    if kind == "k3":
        mo = og._Block(cin, cout, 3, 1)
        mg = _mirror(mo, lambda: sp.BasicConvolutionBlock(cin, cout, ks=3, stride=1))
    elif kind == "down":
        mo = og._Block(cin, cout, 2, 2)
        mg = _mirror(mo, lambda: sp.BasicConvolutionBlock(cin, cout, ks=2, stride=2))
    else:
        mo = og.ResidualBlock(cin, cout, 3, 1)
        mg = _mirror(mo, lambda: sp.ResidualBlock(cin, cout, ks=3, stride=1))
    for p in list(mo.parameters()):          # non-trivial affine parameters
        if p.dim() == 1:
            p.data.uniform_(0.5, 1.5)
    mg.load_state_dict(mo.state_dict())
    ft.fuse(mg)
    assert any(getattr(m, "_ft3d_fused", False) for m in mg.modules())
    mo.train(), mg.train()
    C = voxels.C
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(C.shape[0], cin, generator=g)
    fo = feats.clone().requires_grad_(True)
    xo = ts.SparseTensor(fo, C, 1)
    xo.check()
    yo = mo(xo)
    gsel = torch.randn(yo.F.shape, generator=g)
    (yo.F * gsel).sum().backward()
    fg = feats.cuda().requires_grad_(True)
    xg = ft.SparseTensor(fg, C.cuda(), 1)
    xg.check()
    yg = mg(xg)
    assert torch.equal(yg.C.cpu(), yo.C.int()) and yg.s == yo.s
    assert rel_l2(yg.F, yo.F) < tol
    if mode == "tc" and cin % 16 == 0:
        assert yg.F16 is not None and torch.equal(yg.F16, yg.F.to(torch.bfloat16))   # the next conv's operand
    (yg.F * gsel.cuda()).sum().backward()
    assert rel_l2(fg.grad, fo.grad) < 4 * tol
    _check_grads(mo, mg, 6 * tol, kind)
    _check_buffers(mo, mg, max(tol, 1e-5), kind)


@pytest.mark.parametrize("mode,tol", [("f32", 1e-4), ("tc", 5e-3)])
def test_fused_down_up_chain(monkeypatch, voxels, mode, tol):
    """stride-2 conv block followed by the transposed block that reuses its map (spvcnn.py:22-50)."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import spvcnn as sp
    from oracle import ft_glue as og, ts_ops as ts
    monkeypatch.setenv("FT3D_CONV", mode)
    monkeypatch.setattr(ts, "OPERAND_DTYPE", "bf16" if mode == "tc" else None)
    torch.manual_seed(11)
    mo = torch.nn.Sequential(og._Block(32, 64, 2, 2), og._Block(64, 96, 2, 2, transpose=True))
    mg = torch.nn.Sequential(sp.BasicConvolutionBlock(32, 64, ks=2, stride=2),
                             sp.BasicDeconvolutionBlock(64, 96, ks=2, stride=2))
    mg.load_state_dict(mo.state_dict())
    mg = ft.fuse(mg.cuda())
    mo.train(), mg.train()
    C = voxels.C
    g = torch.Generator().manual_seed(6)
    feats = torch.randn(C.shape[0], 32, generator=g)
    fo = feats.clone().requires_grad_(True)
    xo = ts.SparseTensor(fo, C, 1)
    xo.check()
    yo = mo(xo)
    gsel = torch.randn(yo.F.shape, generator=g)
    (yo.F * gsel).sum().backward()
    fg = feats.cuda().requires_grad_(True)
    xg = ft.SparseTensor(fg, C.cuda(), 1)
    xg.check()
    yg = mg(xg)
    assert yg.s == 1 and torch.equal(yg.C.cpu(), C.int())
    assert rel_l2(yg.F, yo.F) < 2 * tol
    (yg.F * gsel.cuda()).sum().backward()
    assert rel_l2(fg.grad, fo.grad) < 6 * tol
    _check_grads(mo, mg, 8 * tol, "down_up")


def test_fused_eval_mode_uses_running_statistics(monkeypatch, voxels):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import spvcnn as sp
    from oracle import ft_glue as og, ts_ops as ts
    monkeypatch.setenv("FT3D_CONV", "f32")
    monkeypatch.setattr(ts, "OPERAND_DTYPE", None)
    torch.manual_seed(3)
    mo = og.ResidualBlock(32, 64, 3, 1)
    for name, b in mo.named_buffers():
        if name.endswith("running_mean"):
            b.uniform_(-0.5, 0.5)
        elif name.endswith("running_var"):
            b.uniform_(0.5, 2.0)
    mg = sp.ResidualBlock(32, 64, ks=3, stride=1)
    mg.load_state_dict(mo.state_dict())
    mg = ft.fuse(mg.cuda())
    mo.eval(), mg.eval()
    C = voxels.C
    feats = torch.randn(C.shape[0], 32, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        yo = mo(ts.SparseTensor(feats, C, 1))
        yg = mg(ft.SparseTensor(feats.cuda(), C.cuda(), 1))
    assert rel_l2(yg.F, yo.F) < 1e-5


@pytest.mark.parametrize("mode,tol", [("f32", 2e-4), ("tc", 1e-2)])
def test_fused_model_matches_unfused_model(monkeypatch, small_batch, mode, tol):
    """Whole 3D branch: fused and unfused execution of the same parameters on the GPU agree (same kernels and
    rounding points; only the BatchNorm reduction order differs)."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.spvcnn import Net3DSeg
    monkeypatch.setenv("FT3D_CONV", mode)
    torch.manual_seed(1)
    a = Net3DSeg(fusion="middle").cuda().train()
    b = Net3DSeg(fusion="middle").cuda().train()
    b.load_state_dict(a.state_dict())
    ft.unfuse(b)
    a.dropout.p = b.dropout.p = 0.0
    coords, feats = small_batch["coords"].cuda(), small_batch["feats"].cuda()
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 96, generator=g).cuda()
    labels = torch.randint(0, 20, (n,), generator=g).cuda()
    outs = []
    for net in (a, b):
        out = net(ft.SparseTensor(feats, coords), img)["lidar_seg_logit"]
        torch.nn.functional.cross_entropy(out, labels).backward()
        outs.append(out)
    assert rel_l2(outs[0], outs[1]) < tol
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k]) == 1, k
        elif "running_" in k:
            assert rel_l2(sa[k], sb[k]) < max(tol, 1e-4), k
    if mode == "f32":
        gmax = max(p.grad.norm().item() for p in b.parameters() if p.grad is not None)
        for (name, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            err = (pa.grad - pb.grad).norm().item() / max(pb.grad.norm().item(), 1e-4 * gmax)
            assert err < 5e-2, (name, err)


def test_gradient_arena_sink_equals_autograd_accumulation(monkeypatch, voxels):
    """dp.GradSync arena: fused backward kernels add parameter gradients straight into the flat buffer; the result
    equals ordinary autograd accumulation into fresh .grad tensors (two blocks deep, so that bf16 rounding flips of
    the ~50-BatchNorm network do not enter: DESIGN.md "Tolerances")."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import spvcnn as sp
    from fusiontransformer_b200.dp import GradSync
    monkeypatch.setenv("FT3D_CONV", "tc")
    torch.manual_seed(1)

    def build():
        return torch.nn.Sequential(sp.BasicConvolutionBlock(32, 64, ks=2, stride=2), sp.ResidualBlock(64, 128),
                                   sp.BasicDeconvolutionBlock(128, 32, ks=2, stride=2))
    a, b = ft.fuse(build().cuda().train()), ft.fuse(build().cuda().train())
    b.load_state_dict(a.state_dict())
    sync = GradSync(a)
    C = voxels.C.cuda()
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(C.shape[0], 32, generator=g).cuda()
    gsel = torch.randn(C.shape[0], 32, generator=g).cuda()
    for _ in range(2):                                   # second pass: the arena is re-zeroed, not re-allocated
        sync.zero_grad()
        b.zero_grad(set_to_none=True)
        for net in (a, b):
            x = ft.SparseTensor(feats, C, 1)
            x.check()
            (net(x).F * gsel).sum().backward()
        sync.finish()
    for (name, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert sync.owns(pa), name
        assert (pa.grad - pb.grad).norm().item() <= 1e-3 * pb.grad.norm().item() + 1e-7, name


def test_multi_pack_equals_per_layer_pack(monkeypatch):
    """ops.WeightPacker (all images, one launch) writes the same bytes as ft3d_conv_pack_weights per image."""
    from fusiontransformer_b200 import ops
    from fusiontransformer_b200 import spvcnn as sp
    from fusiontransformer_b200.fused import weight_packer
    monkeypatch.setenv("FT3D_CONV", "tc")
    torch.manual_seed(2)
    net = torch.nn.Sequential(sp.BasicConvolutionBlock(32, 64, ks=2, stride=2), sp.ResidualBlock(64, 128),
                              sp.ResidualBlock(128, 96)).cuda()
    packer = weight_packer(net)
    assert packer.n == 2 * sum(1 for m in net.modules() if hasattr(m, "kernel"))
    packer.pack()
    for ref, wt, _, img in packer.entries:
        p = ref()
        w = p.detach().unsqueeze(0) if p.dim() == 2 else p.detach()
        ops._PACK_CACHE.clear()
        single = ops.packed_weights(w.contiguous(), wt)
        assert torch.equal(single, img), (tuple(p.shape), wt)


@pytest.mark.parametrize("mode,tol", [("f32", 1e-4), ("tc", 5e-3)])
@pytest.mark.parametrize("inc,outc,sink", [(32, 256, False), (256, 128, True), (96, 256, True), (128, 96, False)])
def test_fused_point_mlp_matches_torch(monkeypatch, mode, tol, inc, outc, sink):
    """Linear -> BatchNorm1d -> ReLU of the point branch (models/spvcnn.py:164-180) as one node: forward, input
    gradient and all four parameter gradients (the weight gradient runs on the tcgen05 wgrad kernel in tc mode, the
    bias gradient on the column-sum kernel) against the plain torch chain on the CPU in fp64."""
    from fusiontransformer_b200 import spvcnn as sp
    from fusiontransformer_b200.dp import GradSync
    monkeypatch.setenv("FT3D_CONV", mode)
    # library GEMMs in fp32 in both modes: what is under test are the libft3d kernels (tc mode: the bf16 tensor-core
    # weight gradient).  TF32 forward rounding would flip ReLU masks the fp64 reference cannot reproduce.
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(3)
    ref = torch.nn.Sequential(torch.nn.Linear(inc, outc), torch.nn.BatchNorm1d(outc), torch.nn.ReLU(True)).double()
    with torch.no_grad():
        ref[1].weight.uniform_(0.5, 1.5)
        ref[1].bias.uniform_(-0.5, 0.5)
        if mode == "tc":     # the forward GEMM runs on bf16 operands (a14 on tcgen05): use bf16-representable values so
            # that the fp64 reference sees the same operands and the ReLU masks agree; what remains is the fp32
            # accumulation and the bf16 rounding of the backward operand gy
            ref[0].weight.copy_(ref[0].weight.bfloat16().double())
    mlp = sp._point_mlp(inc, outc)
    mlp.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    from fusiontransformer_b200.fused import fuse
    mlp = fuse(mlp).cuda().train()
    assert getattr(mlp, "_ft3d_fused", False)
    if sink:
        GradSync(mlp)
    n = 5000
    x = torch.randn(n, inc, dtype=torch.float64)
    if mode == "tc":
        x = x.bfloat16().double()
    w = torch.randn(n, outc, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    out_ref = ref(xr)                       # one training-mode forward: the running statistics move once
    (out_ref * w).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    out = mlp(xg)
    (out * w.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel_l2(out, out_ref.detach()) < 1e-4
    gtol = max(tol, 1e-3)           # a few of the n*outc ReLU masks flip under fp32 rounding of the pre-activation
    assert rel_l2(xg.grad, xr.grad) < gtol
    for (name, po), (_, pg) in zip(ref.named_parameters(), mlp.named_parameters()):
        # the Linear bias gradient is analytically zero (BatchNorm removes a per-channel shift): what is compared there
        # is fp32 summation noise over n rows, so it is scaled by the size of the summed gradient instead
        scale = max(po.grad.norm().item(), 1e-2 * w.norm().item())
        err = (pg.grad.double().cpu() - po.grad).norm().item() / scale
        assert err < gtol, (name, err)
    assert rel_l2(mlp[1].running_var, ref[1].running_var) < 1e-4


@pytest.mark.parametrize("mode,tol", [("f32", 1e-4), ("tc", 5e-3)])
def test_deconv_cat_in_place_equals_cat(monkeypatch, small_batch, mode, tol):
    """torchsparse.cat([deconv(y), skip]) (models/spvcnn.py:212-228) with the BatchNorm epilogue writing straight into
    the concatenation buffer: same forward as block + torch.cat, same gradients for y, the skip and the parameters."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import spvcnn as sp
    from fusiontransformer_b200.fused import deconv_cat, fuse
    monkeypatch.setenv("FT3D_CONV", mode)
    spf = ft.nn.functional
    torch.manual_seed(2)
    C = small_batch["coords"].int().cuda()
    x = ft.SparseTensor(torch.randn(C.shape[0], 32, device="cuda"), C, 1)
    x.check()
    down = sp.BasicConvolutionBlock(32, 64, ks=2, stride=2).cuda().train()
    up_a = sp.BasicDeconvolutionBlock(64, 96, ks=2, stride=2).cuda().train()
    up_b = sp.BasicDeconvolutionBlock(64, 96, ks=2, stride=2).cuda().train()
    up_b.load_state_dict(up_a.state_dict())
    fuse(down), fuse(up_a), fuse(up_b)
    outs = []
    for up, inplace in ((up_a, True), (up_b, False)):
        xf = x.F.clone().requires_grad_(True)
        xin = ft.SparseTensor(xf, C, 1)
        xin.check()
        skip_f = torch.randn(C.shape[0], 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)).requires_grad_(True)
        skip = ft.SparseTensor(skip_f, C, 1)
        y = down(xin)
        out = deconv_cat(up, y, skip) if inplace else ft.cat([up(y), skip])
        assert out.F.shape == (C.shape[0], 96 + 32)
        w = torch.randn(out.F.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
        (out.F * w).sum().backward()
        outs.append((out.F.detach(), None if out.F16 is None else out.F16.float(), xf.grad, skip_f.grad,
                     up.net[0].kernel.grad.clone(), up.net[1].weight.grad.clone()))
        for p in list(down.parameters()):
            p.grad = None
    a, b = outs
    assert rel_l2(a[0], b[0]) < 1e-6 and rel_l2(a[3], b[3]) < 1e-6            # same kernels, same rows
    if a[1] is not None:                                                      # bf16 twin of the whole concatenation
        # (the second pass sees `down`'s updated running mean as the pivot of its shifted sums: rows agree to 1e-7,
        # which may flip a handful of bf16 roundings)
        assert a[1].shape == a[0].shape and rel_l2(a[1], b[0].bfloat16().float()) < 1e-5
    assert rel_l2(a[2], b[2]) < 1e-4 and rel_l2(a[4], b[4]) < 1e-4 and rel_l2(a[5], b[5]) < 1e-4


@pytest.mark.parametrize("producer", ["bn_stats", "conv_os", "conv_reduce"])
def test_statistics_survive_a_large_mean(monkeypatch, voxels, producer):
    """BatchNorm statistics are shifted sums around the running mean (csrc/bn_common.cuh): a channel whose mean is
    1000x its standard deviation (mean 100, std 0.1) keeps its variance once the running mean has found the
    neighbourhood -- raw fp32 E[y^2] - mean^2 returns noise there (nn.BatchNorm1d / cuDNN use Welford and do not)."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import conv_engine, ops
    g = torch.Generator().manual_seed(7)
    C = voxels.C.int()
    n, c = C.shape[0], 64
    y = (100.0 + 0.1 * torch.randn(n, c, generator=g)).cuda()
    if producer == "bn_stats":
        run = lambda rm, rv: (ops.bn_stats(y, 1e-5, 0.1, rm, rv), y)
    else:
        # an identity-like convolution (only the centre offset carries weight) reproduces y, offset included
        monkeypatch.setenv("FT3D_CONV_ALGO", "os" if producer == "conv_os" else "pairs")
        km = ft.nn.functional.build_kernel_map(C.cuda(), C.cuda(), 3, 1)
        w = torch.zeros(27, c, c)
        w[13] = torch.eye(c)
        w = w.cuda()
        x16 = ops.to_bf16(y)
        if producer == "conv_os":
            def run(rm, rv):
                out, stat = conv_engine.os_conv(x16, km, w, "forward", bn=(1e-5, 0.1, rm, rv))
                return stat, out
        else:
            km.ppos

            def run(rm, rv):
                out, stat = ops.conv_reduce_bn(conv_engine.pairs_partial(x16, km, w, "forward")[0], km.ppos, c, 1e-5, 0.1,
                                               rm, rv)
                return stat, out
    rm = torch.full((c,), 99.9, device="cuda")
    rv = torch.ones(c, device="cuda")
    stat, out = run(rm, rv)
    mean, var = out.double().mean(0), out.double().var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    assert (stat[0].double() - mean).abs().max() < 1e-4
    assert ((stat[1].double() - rstd) / rstd).abs().max() < 2e-3
    assert (rm.double() - (0.9 * 99.9 + 0.1 * mean)).abs().max() < 1e-4
