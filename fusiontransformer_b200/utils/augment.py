"""Host side of the augmentation branch of a1 (data/utils/augmentation_3d.py:22-41,48-51).

The reference draws a dozen random numbers per scan from numpy's global generator and applies them to every point.
Here the draws stay on the host -- in the reference's order, so a seeded run sees the same numbers -- and the per-point
arithmetic runs on the device (``ops.scale_coords(..., rot=, transl_u=)`` -> ``ft3d_augment_scale_coords``).
"""
import numpy as np
import torch


def draw(noisy_rot: float = 0.0, flip_x: float = 0.0, flip_y: float = 0.0, rot_z: float = 0.0, transl: bool = False,
         rng=np.random):
    """One scan's random numbers: -> (rot float32 [3,3] or None, u float64 [3] or None).

    Order of the draws (it defines which numbers a seeded generator hands out): 9 normals for the rotation noise, one
    integer per enabled flip (x, then y), one uniform for the angle about z, and -- after everything that shapes the
    rotation -- 3 uniforms for the translation."""
    rot = None
    if noisy_rot > 0 or flip_x > 0 or flip_y > 0 or rot_z > 0:
        rot = np.eye(3, dtype=np.float32)
        if noisy_rot > 0:
            rot += rng.randn(3, 3) * noisy_rot               # float64 noise accumulated into the float32 matrix
        if flip_x > 0:
            rot[0][0] *= rng.randint(0, 2) * 2 - 1
        if flip_y > 0:
            rot[1][1] *= rng.randint(0, 2) * 2 - 1
        if rot_z > 0:
            theta = rng.rand() * rot_z
            c, s = np.cos(theta), np.sin(theta)
            rot = rot.dot(np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32))
    u = rng.rand(3) if transl else None
    return rot, u


def draw_batch(num_scans: int, device, rng=np.random, **params):
    """The draws of ``num_scans`` consecutive scans as device tensors: (rot f32 [S,3,3] | None, u f64 [S,3] | None)."""
    rots, us = [], []
    for _ in range(num_scans):
        r, u = draw(rng=rng, **params)
        rots.append(r)
        us.append(u)
    rot = None if rots[0] is None else torch.from_numpy(np.stack(rots).astype(np.float32)).to(device)
    tu = None if us[0] is None else torch.from_numpy(np.stack(us).astype(np.float64)).to(device)
    return rot, tu
