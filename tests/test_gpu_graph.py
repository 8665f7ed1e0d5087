"""GraphedStep (graph.py): the training step replayed as one CUDA graph on capacity-padded geometry equals the eager
exact-shape step, for batches of different sizes served by the same captured graph."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_batches():
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.synthetic import make_scan
    # the first batch is the largest, so that the later (smaller) ones replay its graph
    sets = [(0, 1, 2), (3, 4), (5, 6, 7)]
    sets.sort(key=lambda ids: -sum(len(make_scan("nuscenes", i)["points"]) for i in ids))
    return [dataflow.host_batch_from_scans([make_scan("nuscenes", i) for i in ids]) for ids in sets]


def _trainer(mode, optimize=True):
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200.dp import GradSync
    from fusiontransformer_b200.spvcnn import Net3DSeg
    torch.manual_seed(1)
    net = Net3DSeg(fusion="middle").cuda().train()
    net.dropout.p = 0.0
    sync = GradSync(net)
    # f32: SGD is linear in the gradient, so replay == eager to rounding.  Adam's first steps are lr * sign(g) and turn
    # the rounding noise of (mathematically) zero gradients into +-lr parameter changes -- used for the tc run, which
    # is compared at the north-star tolerance anyway, to exercise the capturable optimizer the bench uses.
    opt = (torch.optim.SGD(net.parameters(), lr=0.05) if mode == "f32"
           else torch.optim.Adam(net.parameters(), lr=1e-3, fused=True, capturable=True))
    g = torch.Generator(device="cuda").manual_seed(7)
    fmap = torch.randn(3, 96, 90, 160, device="cuda", generator=g)

    def body(plan):
        ex = plan.extras
        rc = torch.stack([ex["rc"][:, 0] % 90, ex["rc"][:, 1] % 160], 1).contiguous()
        img = ft.nn.functional.lift(fmap, rc, ex["bidx"])
        out = net(ex["lidar"], img.detach(), plan=plan)
        loss = torch.nn.functional.cross_entropy(out["lidar_seg_logit"], ex["labels"])
        sync.zero_grad()
        loss.backward()
        sync.finish()
        if optimize:
            opt.step()
        return loss
    return net, body


@pytest.mark.parametrize("mode,tol", [("f32", 1e-4), ("tc", 1e-2)])
def test_graph_replay_equals_eager_gradients(monkeypatch, mode, tol):
    """Same weights, three batches of different size through ONE captured graph: loss and every parameter gradient
    equal the eager exact-shape step (padding contributes exact zeros)."""
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.graph import GraphedStep
    monkeypatch.setenv("FT3D_CONV", mode)
    batches = _host_batches()
    net_e, body_e = _trainer(mode, optimize=False)
    net_g, body_g = _trainer(mode, optimize=False)
    net_g.load_state_dict(net_e.state_dict())
    gs = GraphedStep(body_g, modules=[net_g])
    for i, hb in enumerate(batches):
        le = body_e(dataflow.prepare_batch(hb, "cuda")).item()
        lg = gs.step(dataflow.prepare_batch(hb, "cuda")).item()
        assert abs(le - lg) <= tol * max(1.0, abs(le)), (i, le, lg)
        gmax = max(p.grad.norm().item() for p in net_e.parameters())
        worst = 0.0
        for (name, pe), (_, pg) in zip(net_e.named_parameters(), net_g.named_parameters()):
            err = (pe.grad - pg.grad).norm().item() / max(pe.grad.norm().item(), 1e-4 * gmax)
            worst = max(worst, err)
            if mode == "f32":
                assert err < 2e-2, (i, name, err)       # fp32 summation order, amplified by ~50 BatchNorms
        if mode == "tc":                                 # bf16 rounding flips: direction check (DESIGN.md Tolerances)
            a = torch.cat([p.grad.flatten() for p in net_e.parameters()])
            b = torch.cat([p.grad.flatten() for p in net_g.parameters()])
            assert (a @ b / (a.norm() * b.norm())).item() > 0.9
    assert gs.captures == 1 and gs.replays == len(batches) - 1, (gs.captures, gs.replays)
    se, sg = net_e.state_dict(), net_g.state_dict()
    for k in se:
        if k.endswith("num_batches_tracked"):
            assert int(se[k]) == int(sg[k]) == len(batches), k
        elif "running_" in k:
            assert ((se[k] - sg[k]).norm() / se[k].norm().clamp_min(1e-12)).item() < max(tol, 1e-3), k


def test_graph_replay_trains(monkeypatch):
    """The captured step includes the (capturable) optimizer: replays keep updating the weights and the loss on a
    fixed batch goes down."""
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.graph import GraphedStep
    monkeypatch.setenv("FT3D_CONV", "tc")
    hb = _host_batches()[0]
    net, body = _trainer("tc")
    gs = GraphedStep(body, modules=[net])
    losses = [gs.step(dataflow.prepare_batch(hb, "cuda")).item() for _ in range(6)]
    assert gs.captures == 1 and gs.replays == 5
    assert losses[-1] < losses[0] - 0.05, losses


def test_graph_recaptures_when_a_batch_does_not_fit(monkeypatch):
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.graph import GraphedStep
    monkeypatch.setenv("FT3D_CONV", "tc")
    batches = _host_batches()[::-1]                       # smallest first: the later ones overflow its capacities
    net, body = _trainer("tc")
    gs = GraphedStep(body, modules=[net], slack=1.0)
    for hb in batches:
        loss = gs.step(dataflow.prepare_batch(hb, "cuda")).item()
        assert np.isfinite(loss)
    assert gs.captures >= 2


def test_eval_between_replays_sees_the_updated_weights(monkeypatch):
    """train(replay) -> eval -> train(replay) -> eval: the optimizer step inside a replay moves the weights on the
    device without bumping Tensor._version, so the cached bf16 weight images must be invalidated by the replay itself
    (ops.REPLAY_EPOCH); every eval must equal one run with freshly packed images."""
    import fusiontransformer_b200 as ft
    from fusiontransformer_b200 import dataflow, ops
    from fusiontransformer_b200.graph import GraphedStep
    monkeypatch.setenv("FT3D_CONV", "tc")
    hb = _host_batches()[0]
    net, body = _trainer("tc")
    gs = GraphedStep(body, modules=[net])

    def evaluate():
        plan = dataflow.prepare_batch(hb, "cuda")
        net.eval()
        with torch.no_grad():
            out = net(plan.extras["lidar"], torch.zeros(plan.extras["lidar"].C.shape[0], 96, device="cuda"), plan=plan)
        net.train()
        return out["lidar_seg_logit"].clone()

    for _ in range(2):
        gs.step(dataflow.prepare_batch(hb, "cuda"))
    for _ in range(2):
        gs.step(dataflow.prepare_batch(hb, "cuda"))        # a replay: weights move, _version does not
        got = evaluate()
        ops._PACK_CACHE.clear()                              # force fresh images
        want = evaluate()
        # (the point->voxel scatter sums with fp32 atomics: equal up to summation order; a stale image differs by ~1e-2)
        assert ((got - want).norm() / want.norm()).item() < 1e-3


def test_prepare_token_carries_the_point_maps(monkeypatch):
    """``GraphedStep.prepare`` (the prefetch path) loads the batch into the static buffers and returns a token; the
    token keeps the host-side maps of that batch -- ``inverse`` (unique voxel -> original points, reference
    data/utils/validate.py:10-11) and ``kept`` -- so predictions of a replayed step can be mapped back to the points."""
    from fusiontransformer_b200 import dataflow
    from fusiontransformer_b200.graph import GraphedStep
    monkeypatch.setenv("FT3D_CONV", "tc")
    batches = _host_batches()
    net, body = _trainer("tc", optimize=False)
    gs = GraphedStep(body, modules=[net])
    gs.step(dataflow.prepare_batch(batches[0], "cuda"))       # capture
    for hb in batches[:2]:
        ref = dataflow.prepare_batch(hb, "cuda")
        tok = gs.prepare(hb, "cuda")
        if not isinstance(tok, GraphedStep.Loaded):            # does not fit the captured capacities: a plan comes back
            assert "inverse" in tok.extras
            gs.step(tok)
            continue
        assert torch.equal(tok.inverse, ref.extras["inverse"]) and torch.equal(tok.kept, ref.extras["kept"])
        loss = gs.step(tok)
        assert torch.isfinite(loss).all()
