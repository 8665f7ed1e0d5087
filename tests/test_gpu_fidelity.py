"""Gradient fidelity of the bf16 tensor-core mode, measured instead of argued (VERDICT r1 "weak" 2).

1. Conditioning: the fp64 CPU oracle is the ground truth.  Against it we measure, tensor by tensor, the parameter
   gradients of (a) the fp32 oracle, (b) the oracle evaluating the product's arithmetic specification (conv / linear
   operands rounded to bf16, fp32 accumulation) and (c) the GPU in tc mode.  The 50-BatchNorm network amplifies any
   2^-9 perturbation of the forward activations at random initialisation, so (b) itself is far from fp64; the claim
   under test is that the GPU is no further from the truth than its own specification is.
2. Convergence: 200 optimiser steps on a fixed batch from the same seed in exact fp32 mode and in tc mode -- the two
   loss curves must stay together, i.e. the bf16 gradients train like the fp32 ones.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _param_grads(net):
    return {n: p.grad.detach().double().cpu() for n, p in net.named_parameters() if p.grad is not None}


def _errors(got, truth):
    gmax = max(v.norm().item() for v in truth.values())
    out = []
    for n, t in truth.items():
        if t.norm().item() < 1e-4 * gmax:        # analytically-zero gradients (Linear bias under BatchNorm): noise
            continue
        out.append(((got[n] - t).norm() / t.norm()).item())
    return np.array(out)


def test_tc_gradients_are_as_close_to_fp64_truth_as_their_arithmetic_spec(monkeypatch, small_batch):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.spvcnn import Net3DSeg
    from oracle import ft_glue as og, ts_ops as ts
    coords, feats = small_batch["coords"], small_batch["feats"]
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 96, generator=g)
    labels = torch.randint(0, 20, (n,), generator=g)
    torch.set_num_threads(os.cpu_count() or 1)

    def oracle_run(dtype, operand):
        monkeypatch.setattr(ts, "OPERAND_DTYPE", operand)
        torch.manual_seed(1)
        net = og.Net3DSeg(num_classes=20, dual_head=False, fusion="middle").to(dtype).train()
        net.dropout.p = 0.0
        out = net(ts.SparseTensor(feats.to(dtype), coords), img.to(dtype))
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels)
        loss.backward()
        return loss.item(), _param_grads(net), net

    l64, g64, _ = oracle_run(torch.float64, None)
    l32, g32, o32 = oracle_run(torch.float32, None)
    lbf, gbf, _ = oracle_run(torch.float32, "bf16")
    monkeypatch.setenv("FT3D_CONV", "tc")
    net = Net3DSeg(num_classes=20, dual_head=False, fusion="middle")
    net.load_state_dict(o32.state_dict())
    net = net.cuda().train()
    net.dropout.p = 0.0
    out = net(ft.SparseTensor(feats.cuda(), coords.cuda()), img.cuda())
    loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels.cuda())
    loss.backward()
    ggpu = _param_grads(net)
    e32, ebf, egpu = _errors(g32, g64), _errors(gbf, g64), _errors(ggpu, g64)
    print("\nloss: fp64 %.6f  fp32 %.6f  bf16-spec %.6f  gpu-tc %.6f" % (l64, l32, lbf, loss.item()))
    for name, e in (("oracle fp32", e32), ("oracle bf16-spec", ebf), ("GPU tc", egpu)):
        print("  %-18s gradient rel-L2 vs fp64 truth: median %.3e  p90 %.3e  worst %.3e" %
              (name, np.median(e), np.percentile(e, 90), e.max()))
    # the loss itself is well conditioned: all three agree with the truth
    assert abs(l32 - l64) < 1e-5 * l64 and abs(lbf - l64) < 2e-3 * l64 and abs(loss.item() - l64) < 2e-3 * l64
    # fp32 arithmetic reproduces the truth; bf16 operand rounding ALONE (CPU, no GPU involved) already moves the
    # gradients by orders of magnitude more -- the network's conditioning, not a kernel property
    assert np.median(e32) < 1e-3
    assert np.median(ebf) > 20 * np.median(e32)
    # the GPU is no further from the truth than the specification it implements
    assert np.median(egpu) < 1.5 * np.median(ebf) + 1e-3
    assert np.percentile(egpu, 90) < 1.5 * np.percentile(ebf, 90) + 1e-3


def _train_curve(mode, steps, monkeypatch):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.dp import GradSync
    from fusiontransformer_b200.spvcnn import Net3DSeg
    from fusiontransformer_b200.synthetic import make_scan
    monkeypatch.setenv("FT3D_CONV", mode)
    torch.manual_seed(1)
    net = Net3DSeg(num_classes=20, dual_head=False, fusion="middle").cuda().train()
    net.dropout.p = 0.0
    sync = GradSync(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    hb = dataflow.host_batch_from_scans([make_scan("nuscenes", i) for i in range(2)])
    plan = dataflow.prepare_batch(hb, "cuda")
    ex = plan.extras
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn(ex["lidar"].C.shape[0], 96, device="cuda", generator=g)
    # learnable labels: a function of the voxel position, so that the loss can actually go down
    labels = ((ex["lidar"].C[:, 0] // 64 + ex["lidar"].C[:, 1] // 64) % 20).long()
    losses = []
    for _ in range(steps):
        plan = dataflow.prepare_batch(hb, "cuda")
        out = net(plan.extras["lidar"], img, plan=plan)
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels)
        sync.zero_grad()
        loss.backward()
        sync.finish()
        opt.step()
        losses.append(loss.item())
    return np.array(losses)


def test_tc_mode_trains_like_fp32_mode(monkeypatch):
    steps = 200
    lf = _train_curve("f32", steps, monkeypatch)
    lt = _train_curve("tc", steps, monkeypatch)
    # The two trajectories are chaotic (and the fp32 run itself differs from run to run through its atomics), so in the
    # steep part of the descent -- the loss falls 5x per 20 steps -- a shift of three steps is a 25 % difference at a
    # fixed step.  The curves are therefore compared by WHEN they reach a loss level, and pointwise only where that is
    # meaningful: at the start (same weights) and at the end (both overfitted).
    def first_below(curve, level):
        sm = np.convolve(curve, np.ones(5) / 5, mode="valid")
        idx = np.flatnonzero(sm < level)
        return int(idx[0]) if len(idx) else len(curve)

    w = 20
    mf, mt = lf.reshape(-1, w).mean(1), lt.reshape(-1, w).mean(1)
    hits = [(lv, first_below(lf, lv), first_below(lt, lv)) for lv in (1.0, 0.3, 0.1, 0.03, 0.01)]
    print("\nloss, mean of each %d-step window\n  f32: %s\n  tc : %s\n  steps to reach (level, f32, tc): %s" %
          (w, np.round(mf, 4), np.round(mt, 4), hits))
    assert lf[-w:].mean() < 0.01 * lf[:5].mean() and lt[-w:].mean() < 0.01 * lt[:5].mean()   # both train (>100x)
    assert abs(mt[0] - mf[0]) < 0.02 * mf[0]                  # same start: first window within 2 %
    for lv, sf, st in hits:
        assert sf < steps and st < steps and abs(st - sf) <= max(6, 0.15 * sf), (lv, sf, st)
    assert abs(mt[-1] - mf[-1]) < 0.25 * mf[-1] + 2e-4        # same plateau


def test_deterministic_mode_gives_bit_identical_steps(monkeypatch, small_batch):
    """FT3D_DETERMINISTIC=1: the point<->voxel scatter-adds run as sorted segmented sums, the weight gradients as
    two-stage / single-owner sums; conv_os and the BatchNorm sums are order-fixed anyway.  Two training steps from the
    same state (tensor-core mode, fused blocks) give bit-identical loss, logits and parameter gradients -- and agree
    with the default (atomic) mode to rounding."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.fused import fuse, join_side_streams
    from fusiontransformer_b200.spvcnn import Net3DSeg
    coords, feats = small_batch["coords"].cuda(), small_batch["feats"].cuda()
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    img = torch.randn(n, 96, generator=g).cuda()
    labels = torch.randint(0, 20, (n,), generator=g).cuda()
    monkeypatch.setenv("FT3D_CONV", "tc")
    torch.manual_seed(1)
    net = Net3DSeg(num_classes=20, dual_head=False, fusion="middle").cuda().train()
    net.dropout.p = 0.0
    fuse(net)
    state = {k: v.clone() for k, v in net.state_dict().items()}

    def step():
        net.load_state_dict(state)                       # running statistics too
        for p in net.parameters():
            p.grad = None
        out = net(ft.SparseTensor(feats, coords), img)
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], labels)
        loss.backward()
        join_side_streams()
        torch.cuda.synchronize()
        return loss.detach().clone(), out["lidar_seg_logit"].detach().clone(), \
            {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

    monkeypatch.setenv("FT3D_DETERMINISTIC", "1")
    la, za, ga = step()
    lb, zb, gb = step()
    assert torch.equal(la, lb) and torch.equal(za, zb)
    for k in ga:
        assert torch.equal(ga[k], gb[k]), k
    monkeypatch.setenv("FT3D_DETERMINISTIC", "0")
    lc, zc, gc = step()
    assert abs(lc.item() - la.item()) < 1e-4 * abs(la.item())
    a = torch.cat([v.flatten() for v in ga.values()]).double()
    c = torch.cat([gc[k].flatten() for k in ga]).double()
    assert (a @ c / (a.norm() * c.norm())).item() > 0.99       # same gradients up to summation order (+ chaos)


@pytest.mark.parametrize("c", [4, 32, 96])
def test_segmented_sums_match_the_atomic_scatters(monkeypatch, c):
    """ft3d_segsum_rows (deterministic) against ft3d_voxelize_fwd / ft3d_devoxelize_bwd (atomic): same sums up to
    summation order, with dropped (-1) contributions, empty voxels and repeated launches bit-identical."""
    from fusiontransformer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(c)
    n, m = 5000, 1300
    idx = torch.randint(-1, m, (n,), device="cuda", generator=g, dtype=torch.int32)
    cnt = torch.bincount(idx[idx >= 0].long(), minlength=m).int()
    feat = torch.randn(n, c, device="cuda", generator=g)
    monkeypatch.setenv("FT3D_DETERMINISTIC", "0")
    want = ops.voxelize_fwd(feat, idx, cnt)
    monkeypatch.setenv("FT3D_DETERMINISTIC", "1")
    got = [ops.voxelize_fwd(feat, idx, cnt) for _ in range(2)]
    assert torch.equal(got[0], got[1]) and (got[0] - want).abs().max() < 1e-5
    idx8 = torch.randint(-1, m, (n, 8), device="cuda", generator=g, dtype=torch.int32)
    w8 = torch.rand(n, 8, device="cuda", generator=g)
    monkeypatch.setenv("FT3D_DETERMINISTIC", "0")
    want = ops.devoxelize_bwd(feat, idx8, w8, m)
    monkeypatch.setenv("FT3D_DETERMINISTIC", "1")
    got = [ops.devoxelize_bwd(feat, idx8, w8, m) for _ in range(2)]
    assert torch.equal(got[0], got[1]) and (got[0] - want).abs().max() < 1e-4 * max(1.0, want.abs().max().item())
