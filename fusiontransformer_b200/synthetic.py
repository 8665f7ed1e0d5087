"""Deterministic synthetic LiDAR scans shaped like the reference's preprocessed pickles.

Stands in for FusionTransformer/data/semantic_kitti/preprocess.py:93-162 (which
emits ``points, feats(=x,y,z,intensity), seg_labels, points_img`` per front-camera
scan) because no dataset is available offline.  Pure numpy, CPU only: this is data
generation, not part of the timed path.  SURVEY.md section 8(d) fixes the recipe.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # beams, elevation range (deg), azimuth step (deg), intrinsics, image (W, H)
    "kitti": dict(beams=64, elev=(2.0, -24.8), az_step=0.15, fx=718.856, cx=607.19, cy=185.22, img=(1226, 370)),
    "nuscenes": dict(beams=32, elev=(10.0, -30.0), az_step=0.2, fx=1266.4, cx=816.3, cy=491.5, img=(1600, 900)),
    "stress": dict(beams=128, elev=(2.0, -24.8), az_step=0.0489, fx=718.856, cx=607.19, cy=185.22, img=(1226, 370)),
}
CONFIG_ID = {"kitti": 1, "nuscenes": 2, "stress": 5}


def _ray_boxes(d, lo, hi):
    """Slab test of unit rays ``d`` [R,3] from the origin against boxes [B,3] lo/hi -> t [R] (inf = miss)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d[:, None, :]
        t0 = lo[None] * inv
        t1 = hi[None] * inv
    tn = np.minimum(t0, t1).max(-1)
    tf = np.maximum(t0, t1).min(-1)
    hit = (tf >= tn) & (tn > 0)
    return np.where(hit, tn, np.inf).min(1) if lo.shape[0] else np.full(d.shape[0], np.inf)


def make_scan(shape: str = "kitti", scan_id: int = 0, num_classes: int = 20):
    """One scan: dict(points [n,3] f32, feats [n,4] f32, points_img [n,2] int64 (row,col), seg_labels [n] int64)."""
    p = SHAPES[shape]
    rng = np.random.default_rng(1000 * CONFIG_ID[shape] + scan_id)
    elev = np.deg2rad(np.linspace(p["elev"][0], p["elev"][1], p["beams"]))
    az = np.deg2rad(np.arange(-90.0, 90.0, p["az_step"]))          # front half-plane only (x > 0)
    E, A = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)

    t = np.full(d.shape[0], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = -1.73 / d[:, 2]
        t = np.where((tg > 0), np.minimum(t, tg), t)
        for wy in (rng.uniform(6, 12), -rng.uniform(6, 12)):
            tw = wy / d[:, 1]
            t = np.where(tw > 0, np.minimum(t, tw), t)
        tf = rng.uniform(50, 75) / d[:, 0]                           # building front closing the street
        t = np.where(tf > 0, np.minimum(t, tf), t)
    nb = int(rng.integers(10, 31))
    cars = rng.random(nb) < 0.6
    size = np.where(cars[:, None], np.array([4.0, 1.8, 1.5]), np.array([0.3, 0.3, 4.0]))
    rad, ang = rng.uniform(5, 60, nb), rng.uniform(-np.pi / 3, np.pi / 3, nb)
    ctr = np.stack([rad * np.cos(ang), rad * np.sin(ang), -1.73 + size[:, 2] / 2], 1)
    t = np.minimum(t, _ray_boxes(d, ctr - size / 2, ctr + size / 2))
    t = t + rng.normal(0.0, 0.02, t.shape)
    ok = np.isfinite(t) & (t > 1.0) & (t < 80.0)
    pts = (d[ok] * t[ok, None]).astype(np.float32)

    # pinhole camera looking along +x: u = cx - fx*y/x, v = cy - fx*z/x ; strictly inside the image
    W, H = p["img"]
    u = p["cx"] - p["fx"] * pts[:, 1] / pts[:, 0]
    v = p["cy"] - p["fx"] * pts[:, 2] / pts[:, 0]
    inside = (pts[:, 0] > 0) & (u > 0) & (u < W - 1) & (v > 0) & (v < H - 1)
    pts, u, v = pts[inside], u[inside], v[inside]
    n = pts.shape[0]
    feats = np.concatenate([pts, rng.random((n, 1), dtype=np.float32)], 1).astype(np.float32)
    return dict(points=pts, feats=feats,
                points_img=np.stack([v, u], 1).astype(np.int64),
                seg_labels=rng.integers(0, num_classes, n).astype(np.int64),
                image_size=(H, W))


def make_batch(shape: str, batch: int, first_scan: int = 0):
    return [make_scan(shape, first_scan + i) for i in range(batch)]
