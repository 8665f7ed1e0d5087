"""Model-level golden vectors produced by EXECUTING THE REFERENCE'S OWN MODEL FILES on the CPU:
    python -m tests.golden.make_reference_model_golden

FusionTransformer/models/spvcnn.py (SPVCNN, ResidualBlock, Basic(De)ConvolutionBlock), models/utils.py
(initial_voxelize, point_to_voxel, voxel_to_point), models/middle_fusion.py:10-88 and models/early_fusion.py:9-87
(Net3DSeg) are imported unmodified from /root/reference; their `torchsparse` imports resolve to oracle/ref_alias.py,
i.e. the operators are the oracle's restatement of torchsparse v1.1.0 (absent here) while every line of topology and
glue that runs is the reference's.  The outputs are committed as tests/golden/ref_model_small.npz because
/root/reference does not exist on the GPU box.  tests/test_reference_topology.py holds the oracle's restated model
(oracle/ft_glue.py) bit-exactly to the same run.
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_net(fusion: str, seed: int = 1):
    """The reference Net3DSeg (middle / early) carrying the oracle model's seeded parameters."""
    from oracle import ft_glue as og
    from oracle import ref_alias
    spv, utl, mid = ref_alias.load_reference_models()
    torch.manual_seed(seed)
    o = og.Net3DSeg(num_classes=20, dual_head=False, fusion=fusion)
    if fusion == "middle":
        r = mid.Net3DSeg(num_classes=20, dual_head=False)
    elif fusion == "early":
        import importlib
        r = importlib.import_module("FusionTransformer.models.early_fusion").Net3DSeg(num_classes=20, dual_head=False)
    else:
        raise ValueError(fusion)
    r.load_state_dict(o.state_dict(), strict=True)          # identical parameter names and shapes
    return r, o


def gen():
    from oracle import ts_ops as ts
    from tests.golden.make_golden import _voxels, model_small_img_feats
    torch.set_num_threads(1)                                   # fixed summation order for the fixture
    st, _, _ = _voxels(1)
    img = model_small_img_feats(st.C.shape[0])
    out = dict(coords=st.C.numpy().astype(np.int32), feats=st.F.numpy())
    for fusion in ("middle", "early"):
        r, _ = reference_net(fusion)
        r.eval()
        with torch.no_grad():
            p = r(ts.SparseTensor(st.F, st.C), img)
        out[fusion + "_eval_logits"] = p["lidar_seg_logit"].numpy()
        r.train()
        r.dropout.p = 0.0
        labels = (torch.arange(st.C.shape[0]) * 7 % 20)
        p = r(ts.SparseTensor(st.F, st.C), img)
        loss = torch.nn.functional.cross_entropy(p["lidar_seg_logit"], labels)
        loss.backward()
        out[fusion + "_train_logits"] = p["lidar_seg_logit"].detach().numpy()
        out[fusion + "_train_loss"] = np.float64(loss.item())
        out[fusion + "_grad_linear_weight"] = r.linear.weight.grad.numpy()
        out[fusion + "_grad_up4_last_kernel"] = r.up4[1][1].net[3].kernel.grad.numpy()
        out[fusion + "_grad_stem0_kernel"] = r.stem[0].kernel.grad.numpy()
    out["labels"] = labels.numpy()
    return out


def main():
    d = gen()
    np.savez_compressed(os.path.join(HERE, "ref_model_small.npz"), **d)
    print({k: getattr(v, "shape", v) for k, v in d.items()})


if __name__ == "__main__":
    main()
