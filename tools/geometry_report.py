"""Achieved GB/s of the integer / gather stages of one training step (hashing + dedup, kernel-map building, point-voxel
gathers, lift) from a bench.py JSON line: algorithmic bytes of SURVEY 8(d) -- evaluated on the sizes of the same
synthetic batch, which the CPU oracle recomputes here -- divided by the per-step device time bench.py recorded for the
entry points of each stage (`kernel_ms_per_step`: CUDA events around every call in the kernel-by-kernel pass, so each
call carries a few microseconds of event overhead and small workloads are understated).

    python tools/geometry_report.py nuscenes profiles/r01_bench_final_n1.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sizes(shape, batch=8):
    from fusiontransformer_b200.synthetic import make_scan
    from oracle import ft_glue as og, ts_ops as ts
    items, n_raw = [], 0
    for i in range(batch):
        s = make_scan(shape, i)
        n_raw += len(s["points"])
        vc, keep, inds, _ = og.voxelize_scan(s["points"])
        items.append(dict(coords=vc[inds], feats=s["feats"][keep][inds]))
    st = og.collate(items)
    cur, stride, lv = st.C.int(), 1, []
    for lvl in range(5):
        _, _, counts = ts.build_kernel_map(cur, cur, 3, stride)
        lv.append(dict(stride=stride, N=int(cur.shape[0]), L3=int(counts.sum())))
        if lvl < 4:
            cur = ts.spdownsample(cur, stride * 2)
            stride *= 2
    return n_raw, lv


def main():
    shape, path = sys.argv[1], sys.argv[2]
    text = open(path).read()
    try:
        line = json.loads(text)                          # one (possibly indented) JSON object
    except json.JSONDecodeError:
        line = [json.loads(l) for l in text.splitlines() if l.startswith("{")][-1]      # a log with the JSON line in it
    ms = line["kernel_ms_per_step"]
    peak = 6547.5
    n_raw, lv = sizes(shape)
    N = [l["N"] for l in lv]
    L = [l["L3"] for l in lv]
    u = N[0]
    print("workload %s: %d raw points, voxels per stride %s, k3 pairs per stride %s" % (shape, n_raw, N, L))
    # quantize 16 n + 20 u ; map build 28 N_in + 16 N_out + 8 K N_out + 8 L ; p2v N1(4+4C)+Ns(4+4C) ;
    # v2p N1*64 + Ns*4C + N1*4C ; lift N1 (8 + 2*96*4)
    b_quant = 16 * n_raw + 20 * u + (16 * u + 20 * u)                   # dataloader dedup + initial_voxelize re-dedup
    b_map = sum(28 * n + 16 * n + 8 * 27 * n + 8 * l for n, l in zip(N, L))            # five k3 maps
    b_map += sum(28 * N[i] + 16 * N[i + 1] + 8 * 8 * N[i + 1] + 8 * N[i] for i in range(4))   # four k2s2 maps
    chans = {0: 32, 4: 256, 2: 128}                                     # p2v at strides 1, 16, 4 (spvcnn.py:201,209,221)
    b_p2v = sum(u * (4 + 4 * c) + N[i] * (4 + 4 * c) for i, c in chans.items())
    b_v2p = sum(u * 64 + N[i] * 4 * c + u * 4 * c for i, c in ((0, 32), (4, 256), (2, 128), (0, 96)))
    b_lift = u * (8 + 2 * 96 * 4)
    rows = [
        ("quantize + dedup (scale_coords, quantize, hash, unique, gather_rows)", b_quant,
         sum(ms.get(k, 0) for k in ("scale_coords", "quantize", "hash", "unique", "gather_rows_i32"))),
        ("kernel maps (table_build, coarsen_hash, kmap_build, kmap_pairs, kmap_pair_positions)", b_map,
         sum(ms.get(k, 0) for k in ("table_build", "coarsen_hash", "kmap_build", "kmap_pairs", "kmap_pair_positions"))),
        ("point->voxel (p2v_build, voxelize_fwd)", b_p2v, sum(ms.get(k, 0) for k in ("p2v_build", "voxelize_fwd"))),
        ("voxel->point (v2p_build, devoxelize_fwd)", b_v2p, sum(ms.get(k, 0) for k in ("v2p_build", "devoxelize_fwd"))),
        ("lift (lift_fwd)", b_lift, ms.get("lift_fwd", 0)),
    ]
    print("%-88s %10s %9s %9s %7s" % ("stage", "alg. MB", "ms/step", "GB/s", "of HBM"))
    for name, b, t in rows:
        gbs = b / 1e9 / (t * 1e-3) if t else float("nan")
        print("%-88s %10.1f %9.3f %9.1f %6.1f%%" % (name, b / 1e6, t, gbs, 100 * gbs / peak))
    print("step %.2f ms; these stages run on the prefetch stream (quantize, maps) or inside the graph (gathers, lift)"
          % line["ms_per_step"])


if __name__ == "__main__":
    main()
