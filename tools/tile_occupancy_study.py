"""Design study for the output-stationary conv: how many (128-row tile, offset) combinations are non-empty
under different row orders?  CPU only (numpy); prints efficiency = pairs / (128 * sum_tiles popcount(union mask)).

    python tools/tile_occupancy_study.py [nuscenes|kitti|stress] [batch]
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from fusiontransformer_b200.synthetic import make_batch  # noqa: E402


def voxelize(batch, scale=20, full=4096):
    out = []
    for b, s in enumerate(batch):
        p = s["points"] * scale
        c = (p - p.min(0)).astype(np.int64)
        ok = ((c >= 0) & (c < full)).all(1)
        c = np.unique(c[ok], axis=0)
        out.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(out)


def key_of(c):
    return ((c[:, 3] * 4096 + c[:, 0]) * 4096 + c[:, 1]) * 4096 + c[:, 2]


def masks_k3(c, stride):
    keys = key_of(c)
    order = np.argsort(keys)
    sk = keys[order]
    m = np.zeros(len(c), np.int64)
    nbr = np.full((len(c), 27), -1, np.int64)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                q = c.copy()
                q[:, 0] += dx * stride; q[:, 1] += dy * stride; q[:, 2] += dz * stride
                ok = ((q[:, :3] >= 0) & (q[:, :3] < 4096)).all(1)
                qk = key_of(q)
                pos = np.searchsorted(sk, qk)
                pos[pos >= len(sk)] = 0
                hit = ok & (sk[pos] == qk)
                m |= hit.astype(np.int64) << k
                nbr[hit, k] = order[pos[hit]]
                k += 1
    return m, nbr


def popcount(x):
    x = x.copy(); n = np.zeros_like(x)
    while x.any():
        n += x & 1; x >>= 1
    return n


def eff(m, order, tile=128):
    ms = m[order]
    pad = (-len(ms)) % tile
    ms = np.concatenate([ms, np.zeros(pad, np.int64)]).reshape(-1, tile)
    union = np.bitwise_or.reduce(ms, 1)
    return popcount(m).sum() / (tile * popcount(union).sum()), popcount(union).mean()


def morton(c, stride):
    x = (c[:, :3] // stride).astype(np.int64)
    code = np.zeros(len(c), np.int64)
    for b in range(12):
        for a in range(3):
            code |= ((x[:, a] >> b) & 1) << (3 * b + a)
    return code + (c[:, 3].astype(np.int64) << 40)


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "nuscenes"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    c = voxelize(make_batch(shape, B))
    stride = 1
    while len(c) > 200 and stride <= 16:
        m, nbr = masks_k3(c, stride)
        L = popcount(m).sum()
        n = len(c)
        rnd = np.random.default_rng(0).permutation(n)
        res = {
            "hash(random)": eff(m, rnd),
            "morton": eff(m, np.argsort(morton(c, stride), kind="stable")),
            "mask": eff(m, np.argsort(m, kind="stable")),
            "popcnt,mask": eff(m, np.lexsort((m, popcount(m)))),
            "mask sans centre, 64-row": eff(m, np.argsort(m, kind="stable"), 64),
        }
        print("stride %2d  N=%7d  L=%8d  L/N=%.2f  distinct masks=%d" % (stride, n, L, L / n, len(np.unique(m))))
        for k, (e, u) in res.items():
            print("     %-26s efficiency %.3f   offsets/tile %.1f" % (k, e, u))
        # coarsen
        cc = c.copy(); cc[:, :3] = cc[:, :3] // (2 * stride) * (2 * stride)
        c = np.unique(cc, axis=0)
        stride *= 2


if __name__ == "__main__":
    main()


def reorder_bits(m, rare_msb=True):
    """Permute mask bits so that the rarest offsets become the most significant sort digits."""
    freq = np.array([((m >> k) & 1).sum() for k in range(27)])
    order = np.argsort(freq if not rare_msb else -freq, kind="stable")   # order[0] -> bit 0 (LSB)
    out = np.zeros_like(m)
    for newbit, k in enumerate(order):
        out |= ((m >> k) & 1) << newbit
    return out
