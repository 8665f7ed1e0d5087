"""Golden vectors produced by the REFERENCE'S OWN CODE, for the parts of the hot path that live in the reference tree
and can be executed in the build container:   python -m tests.golden.make_reference_golden

  * a1  FusionTransformer/data/utils/augmentation_3d.py:4-53 `augment_and_scale_3d` (numpy only), followed by the two
        dataloader lines that turn its output into voxel coordinates (semantic_kitti_dataloader.py:220 cast, :225
        bounds mask -- inside a Dataset.__getitem__ that needs the KITTI files, so they are restated literally here);
  * a3  FusionTransformer/data/collate.py:6-86 `collate_scn_base` (its SparseTensor import resolves to this package's
        container through install_as_torchsparse());
  * a16 FusionTransformer/data/utils/validate.py:10-11 `map_sparse_to_org`;
  * (f)4 FusionTransformer/models/metric.py:26-82 `SegIoU` (update_dict / iou), torch only.

Everything else on the path (sparse_quantize, the operators, the convolution) lives in torchsparse v1.1.0, which is
not in the tree and not installable here: those vectors come from the oracle (make_golden.py) and stay "unpinned".
/root/reference does not exist on the GPU box, hence the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def gen_ref_voxelize():
    sys.path.insert(0, REF)
    from FusionTransformer.data.utils.augmentation_3d import augment_and_scale_3d
    from fusiontransformer_b200.synthetic import make_scan
    out = {}
    for tag, shape, sid in (("nus", "nuscenes", 3), ("kitti", "kitti", 5)):
        pts = make_scan(shape, sid)["points"]
        if tag == "kitti":                        # push part of the scan outside the 4096-voxel receptive field
            pts = pts.copy()
            pts[::17, 0] += 250.0
        coords_f = augment_and_scale_3d(pts, 20, 4096)                    # the reference function, no augmentation
        coords = coords_f.astype(np.int64)                                # semantic_kitti_dataloader.py:220
        idxs = (coords.min(1) >= 0) * (coords.max(1) < 4096)              # semantic_kitti_dataloader.py:225
        out.update({tag + "_points": pts, tag + "_coords_float": coords_f, tag + "_coords": coords, tag + "_keep": idxs})
    return out


def gen_ref_augment():
    """a1 with the augmentation branch on (augmentation_3d.py:22-51; the values of the commented-out training configs,
    config/FusionTransformerConfig.py:44-47,66-69): the reference function under a seeded global numpy generator."""
    sys.path.insert(0, REF)
    from FusionTransformer.data.utils.augmentation_3d import augment_and_scale_3d
    from fusiontransformer_b200.synthetic import make_scan
    out = {}
    cases = (("a", "nuscenes", 7, 101, dict(noisy_rot=0.1, flip_x=0.5, rot_z=6.2831, transl=True)),
             ("b", "nuscenes", 8, 202, dict(noisy_rot=0.1, flip_y=0.5, rot_z=6.2831, transl=True)),
             ("c", "nuscenes", 9, 303, dict(rot_z=6.2831)),
             ("d", "nuscenes", 10, 404, dict(transl=True)))
    for tag, shape, sid, seed, kw in cases:
        pts = make_scan(shape, sid)["points"][:3000].copy()
        np.random.seed(seed)
        coords_f = augment_and_scale_3d(pts, 20, 4096, **kw)
        coords = coords_f.astype(np.int64)
        idxs = (coords.min(1) >= 0) * (coords.max(1) < 4096)
        out.update({tag + "_points": pts, tag + "_seed": np.int64(seed), tag + "_coords_float": coords_f,
                    tag + "_coords": coords.astype(np.int32), tag + "_keep": idxs})
        for k in ("noisy_rot", "flip_x", "flip_y", "rot_z"):
            out[tag + "_" + k] = np.float64(kw.get(k, 0.0))
        out[tag + "_transl"] = np.bool_(kw.get("transl", False))
    return out


def gen_ref_segiou():
    sys.path.insert(0, REF)
    from FusionTransformer.data.utils.validate import map_sparse_to_org
    from FusionTransformer.models.metric import SegIoU
    g = torch.Generator().manual_seed(11)
    m = SegIoU(20, ignore_index=0, name="seg_iou_3d")
    out = {}
    for step, n in enumerate((4099, 2500)):
        logits = torch.randn(n, 20, generator=g)
        labels = torch.randint(0, 20, (n,), generator=g)
        m.update_dict({"lidar_seg_logit": logits}, {"seg_label": labels})
        out["logits%d" % step], out["labels%d" % step] = logits.numpy(), labels.numpy()
    out["mat"] = m.mat.numpy()
    out["iou"] = m.iou.numpy()
    inv = torch.randint(0, 2500, (6000,), generator=g)
    out["inverse_map"] = inv.numpy()
    out["pred_points"] = map_sparse_to_org(torch.from_numpy(out["logits1"]).argmax(1), inv).numpy()
    return out


def gen_ref_collate():
    """a3: FusionTransformer/data/collate.py:6-86 `collate_scn_base`, imported from the reference tree.  Its one
    dependency, torchsparse's SparseTensor container, is provided by this package's alias (a plain holder of .C/.F)."""
    sys.path.insert(0, REF)
    import fusiontransformer_b200 as ft
    ft.install_as_torchsparse()
    from FusionTransformer.data.collate import collate_scn_base
    from fusiontransformer_b200.synthetic import make_scan
    from oracle import ft_glue as og
    dicts, out = [], {}
    for i in range(3):
        s = make_scan("nuscenes", 20 + i)
        vc, keep, inds, inv = og.voxelize_scan(s["points"])
        d = dict(voxel_coords=vc, coords=vc[inds], feats=s["feats"][keep][inds], seg_label=s["seg_labels"][keep][inds],
                 img=np.zeros((3, 4, 4), np.float32), img_indices=s["points_img"][keep][inds], seq="00", filename="%06d" % i)
        dicts.append(d)
        out["coords%d" % i], out["feats%d" % i], out["labels%d" % i] = d["coords"], d["feats"], d["seg_label"]
    batch = collate_scn_base(dicts, output_orig=False)
    out["C"], out["F"], out["seg_label"] = batch["lidar"].C.numpy(), batch["lidar"].F.numpy(), batch["seg_label"].numpy()
    assert isinstance(batch["img_indices"], list) and len(batch["img_indices"]) == 3
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit("the reference tree is needed to regenerate these fixtures")
    np.savez_compressed(os.path.join(HERE, "ref_voxelize.npz"), **gen_ref_voxelize())
    np.savez_compressed(os.path.join(HERE, "ref_segiou.npz"), **gen_ref_segiou())
    np.savez_compressed(os.path.join(HERE, "ref_collate.npz"), **gen_ref_collate())
    np.savez_compressed(os.path.join(HERE, "ref_augment.npz"), **gen_ref_augment())
    for f in ("ref_voxelize.npz", "ref_segiou.npz", "ref_collate.npz", "ref_augment.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
