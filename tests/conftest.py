import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def collate_scans(shape, batch, first=0):
    """Reference dataloader + collate pipeline (oracle restatement) on synthetic scans.
    Returns dict(coords int64 [N,4], feats f32 [N,4], img_indices list, inverse list, scans list)."""
    from fusiontransformer_b200.synthetic import make_batch
    from oracle import ft_glue as og
    scans = make_batch(shape, batch, first)
    items, img_idx, inv = [], [], []
    for s in scans:
        vc, keep, inds, invs = og.voxelize_scan(s["points"])
        items.append(dict(coords=vc[inds], feats=s["feats"][keep][inds]))
        img_idx.append(s["points_img"][keep][inds])
        inv.append(invs)
    st = og.collate(items)
    return dict(coords=st.C, feats=st.F, img_indices=img_idx, inverse=inv, scans=scans)


@pytest.fixture(scope="session")
def small_batch():
    return collate_scans("nuscenes", 2)


@pytest.fixture(scope="session")
def kitti_scan():
    return collate_scans("kitti", 1)
