"""The reference's own model files, EXECUTED (not merely imported), pin the oracle's restated topology.

CPU only.  Two layers of evidence:
  * dev container (where /root/reference exists): FusionTransformer/models/spvcnn.py + utils.py + middle_fusion.py /
    early_fusion.py run unmodified on the oracle-backed `torchsparse` alias (oracle/ref_alias.py); the oracle's
    restatement of the same model (oracle/ft_glue.py) must give bit-identical logits and parameter gradients.
  * everywhere (also the GPU box): the oracle model must reproduce tests/golden/ref_model_small.npz, the committed
    outputs of that reference run (tests/golden/make_reference_model_golden.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import ft_glue as og
from oracle import ref_alias
from oracle import ts_ops as ts

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _inputs():
    gold = np.load(os.path.join(GOLD, "ref_model_small.npz"))
    from tests.golden.make_golden import model_small_img_feats
    coords = torch.from_numpy(gold["coords"]).long()
    feats = torch.from_numpy(gold["feats"])
    return gold, coords, feats, model_small_img_feats(coords.shape[0])


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)
    ref_alias.uninstall()


@pytest.mark.parametrize("fusion", ["middle", "early"])
def test_oracle_model_reproduces_reference_run_fixture(fusion):
    gold, coords, feats, img = _inputs()
    torch.manual_seed(1)
    o = og.Net3DSeg(num_classes=20, dual_head=False, fusion=fusion).eval()
    with torch.no_grad():
        logits = o(ts.SparseTensor(feats, coords), img)["lidar_seg_logit"]
    assert np.array_equal(logits.numpy(), gold[fusion + "_eval_logits"])
    o.train()
    o.dropout.p = 0.0
    p = o(ts.SparseTensor(feats, coords), img)
    loss = torch.nn.functional.cross_entropy(p["lidar_seg_logit"], torch.from_numpy(gold["labels"]))
    loss.backward()
    assert np.array_equal(p["lidar_seg_logit"].detach().numpy(), gold[fusion + "_train_logits"])
    assert loss.item() == float(gold[fusion + "_train_loss"])
    assert np.array_equal(o.linear.weight.grad.numpy(), gold[fusion + "_grad_linear_weight"])
    assert np.array_equal(o.stem[0].kernel.grad.numpy(), gold[fusion + "_grad_stem0_kernel"])
    assert np.array_equal(o.up4[1][1].net[3].kernel.grad.numpy(), gold[fusion + "_grad_up4_last_kernel"])


@pytest.mark.skipif(not ref_alias.available(), reason="reference tree not present on this box")
@pytest.mark.parametrize("fusion", ["middle", "early"])
def test_reference_model_files_execute_and_match_oracle(fusion):
    from tests.golden.make_reference_model_golden import reference_net
    _, coords, feats, img = _inputs()
    r, o = reference_net(fusion)
    assert type(r).__module__.startswith("FusionTransformer.models.") and type(r).__mro__[1].__name__ == "SPVCNN"
    assert [n for n, _ in r.named_parameters()] == [n for n, _ in o.named_parameters()]
    for net in (r, o):
        net.train()
        net.dropout.p = 0.0
    a = r(ts.SparseTensor(feats, coords), img)["lidar_seg_logit"]
    b = o(ts.SparseTensor(feats, coords), img)["lidar_seg_logit"]
    assert torch.equal(a, b)
    a.square().mean().backward()
    b.square().mean().backward()
    po = dict(o.named_parameters())
    for n, p in r.named_parameters():
        if p.grad is None:
            assert po[n].grad is None, n
        else:
            assert torch.equal(p.grad, po[n].grad), n


@pytest.mark.skipif(not ref_alias.available(), reason="reference tree not present on this box")
def test_reference_lidar_only_backbone_matches_oracle():
    """models/lidar_model.py:4-22 (LidarSeg = SPVCNN + one head): the reference SPVCNN.forward :191-233."""
    spv, _, _ = ref_alias.load_reference_models()
    _, coords, feats, _ = _inputs()
    torch.manual_seed(1)
    o = og.SPVCNN().eval()
    r = spv.SPVCNN().eval()
    r.load_state_dict(o.state_dict(), strict=True)
    with torch.no_grad():
        assert torch.equal(r(ts.SparseTensor(feats, coords)), o(ts.SparseTensor(feats, coords)))
