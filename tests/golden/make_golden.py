"""Generates the committed golden fixtures from the CPU oracle:  python -m tests.golden.make_golden

PARITY UNPINNED: the reference has no fixtures and its arithmetic (torchsparse v1.1.0) cannot run here, so these
vectors freeze the ORACLE's behaviour (itself pinned by tests/test_oracle.py against dense conv3d, big-integer FNV
and algebraic properties).  The GPU parity tests compare libft3d with the same vectors on the GPU box, where neither
/root/reference nor a second implementation exists.
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _scan(shape="nuscenes", scan_id=0):
    from fusiontransformer_b200.synthetic import make_scan
    return make_scan(shape, scan_id)


def gen_quantize_small():
    from oracle import ft_glue as og
    s = _scan()
    vc, keep, inds, inv = og.voxelize_scan(s["points"])
    return dict(points=s["points"], coords=vc.astype(np.int32), keep=keep, inds=inds.astype(np.int64),
                inverse=inv.astype(np.int64))


def _voxels(nscan=2):
    from oracle import ft_glue as og
    from oracle import ts_ops as ts
    items = []
    for i in range(nscan):
        s = _scan(scan_id=i)
        vc, keep, inds, _ = og.voxelize_scan(s["points"])
        items.append(dict(coords=vc[inds], feats=s["feats"][keep][inds]))
    st = og.collate(items)
    z = ts.PointTensor(st.F, st.C.float())
    x0 = og.initial_voxelize(z, 1, 1)
    return st, z, x0


def gen_kmap_small():
    from oracle import ts_ops as ts
    st, z, x0 = _voxels()
    nbr, pairs, counts = ts.build_kernel_map(x0.C, x0.C, 3, 1)
    c2 = ts.spdownsample(x0.C, 2)
    nbr2, pairs2, counts2 = ts.build_kernel_map(x0.C, c2, 2, 1)
    return dict(in_coords=st.C.numpy().astype(np.int32), coords=x0.C.numpy(), hash=ts.sphash(x0.C).numpy(),
                idx_query=z.additional_features["idx_query"][1].numpy(),
                nbr_k3=nbr.numpy().astype(np.int32), pairs_k3=pairs.numpy(), counts_k3=counts.numpy(),
                coords_s2=c2.numpy(), nbr_k2=nbr2.numpy().astype(np.int32), pairs_k2=pairs2.numpy(),
                counts_k2=counts2.numpy())


def model_small_img_feats(n):
    """Lifted image features of the model fixture: a fixed low-discrepancy pattern (no RNG dependence)."""
    i = torch.arange(n, dtype=torch.float64).view(-1, 1)
    c = torch.arange(96, dtype=torch.float64).view(1, -1)
    return torch.sin(0.37 * i + 1.3 * c).float()


def gen_model_small():
    """Eval-mode forward of the middle-fusion 3D branch (seeded weights) on one small scan."""
    from oracle import ft_glue as og
    from oracle import ts_ops as ts
    torch.manual_seed(1)
    torch.set_num_threads(1)     # fixed summation order for the fixture
    net = og.Net3DSeg(num_classes=20, dual_head=False, fusion="middle").eval()
    st, _, _ = _voxels(1)
    img = model_small_img_feats(st.C.shape[0])
    with torch.no_grad():
        out = net(ts.SparseTensor(st.F, st.C), img)
    return dict(coords=st.C.numpy().astype(np.int32), feats=st.F.numpy(),
                logits=out["lidar_seg_logit"].numpy(), lidar_feats_mean=out["lidar_feats"].mean(0).numpy())


def main():
    for name in ("quantize_small", "kmap_small", "model_small"):
        d = globals()["gen_" + name]()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
