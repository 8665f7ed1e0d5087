"""CPU checks that pin the oracle (PARITY UNPINNED by the reference: no tests/goldens exist there).

The oracle is validated against (1) an independent big-integer FNV implementation and hard-coded known answers,
(2) dense torch.nn.functional.conv3d / conv_transpose3d on fully occupied cubes, (3) kernel-map invariants,
(4) algebraic properties of the point<->voxel glue, (5) the committed golden fixtures.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ft_glue as og
from oracle import ts_ops as ts

M64 = (1 << 64) - 1
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def fnv1a_words(words):
    h = 14695981039346656037
    for w in words:
        h ^= (w & 0xFFFFFFFF)
        h = (h * 1099511628211) & M64
    return (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFF)


def fnv1_cols(cols):
    h = 14695981039346656037
    for c in cols:
        h = (h * 1099511628211) & M64
        h ^= (c & M64)
    return h


def test_sphash_known_answers():
    # literals hand-computed with python big integers from the public FNV-1a-64 basis/prime
    # (one 32-bit word per xor/multiply step, then the 60-bit fold of SURVEY App. A.2)
    kat = {(0, 0, 0, 0): 0x0D25767F9DCE13F1, (1, 2, 3, 0): 0x0E7A5A775165719A, (-1, 5, 4095, 7): 0x0601D450FE50BE7A}
    coords = torch.tensor([list(k) for k in kat], dtype=torch.int32)
    got = ts.sphash(coords).tolist()
    for (c, want), g in zip(kat.items(), got):
        assert g == want == fnv1a_words(c), c
    kat3 = {(0, 0, 0): 0xD94D12186C0F2FB7, (1, 2, 3): 0xD949AA186C0C492B, (4095, 0, 17): 0xDA29F8186CCADD31}
    got3 = ts.fnv_hash_vec(np.array([list(k) for k in kat3])).tolist()
    for (c, want), g in zip(kat3.items(), got3):
        assert g == want == fnv1_cols(c), c


def test_sphash_literal_constants():
    # literal values computed once with python big ints (see fnv1a_words); guards the restatement itself
    c = torch.tensor([[0, 0, 0, 0], [1, 2, 3, 0], [100, 200, 300, 1]], dtype=torch.int32)
    ref = [fnv1a_words(tuple(r)) for r in c.tolist()]
    assert ts.sphash(c).tolist() == ref
    off = ts.KernelRegion(3, 2).get_kernel_offset()
    got = ts.sphash(c, off)
    assert got.shape == (27, 3)
    for k in range(27):
        for i in range(3):
            x, y, z, b = c[i].tolist()
            dx, dy, dz = off[k].tolist()
            assert got[k, i].item() == fnv1a_words((x + dx, y + dy, z + dz, b))


def test_fnv_hash_vec_matches_bigint():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 4096, size=(64, 3))
    got = ts.fnv_hash_vec(a)
    for r, g in zip(a.tolist(), got.tolist()):
        assert g == fnv1_cols(r)


def test_sparse_quantize_semantics():
    rng = np.random.default_rng(1)
    c = rng.integers(0, 6, size=(500, 3))
    inds, lab, inv = ts.sparse_quantize(c, np.zeros((500, 1), np.float32), np.arange(500), return_index=True,
                                        return_invs=True)
    keys = ts.fnv_hash_vec(c)
    assert np.all(np.diff(keys[inds].astype(np.uint64)) > 0)            # ascending key order
    assert np.array_equal(keys[inds][inv], keys)                         # inverse map
    for g, i in enumerate(inds):                                         # first occurrence
        assert i == np.nonzero(keys == keys[i])[0][0]
    assert np.array_equal(c[inds][inv], c)


def test_kernel_region_order():
    o3 = ts.KernelRegion(3, 1).get_kernel_offset().tolist()
    assert o3[0] == [-1, -1, -1] and o3[1] == [0, -1, -1] and o3[13] == [0, 0, 0] and o3[26] == [1, 1, 1]
    o2 = ts.KernelRegion(2, 4).get_kernel_offset().tolist()
    assert o2 == [[0, 0, 0], [0, 0, 4], [0, 4, 0], [0, 4, 4], [4, 0, 0], [4, 0, 4], [4, 4, 0], [4, 4, 4]]


def _cube(D, C, seed=0, batch=1):
    g = torch.Generator().manual_seed(seed)
    xs = torch.stack(torch.meshgrid(torch.arange(D), torch.arange(D), torch.arange(D), indexing="ij"), -1).reshape(-1, 3)
    coords = torch.cat([torch.cat([xs, torch.full((xs.shape[0], 1), b)], 1) for b in range(batch)], 0).int()
    perm = torch.randperm(coords.shape[0], generator=g)
    coords = coords[perm]
    feats = torch.randn(coords.shape[0], C, generator=g)
    return coords, feats


def _dense(coords, feats, D, batch=1):
    vol = torch.zeros(batch, feats.shape[1], D, D, D)
    c = coords.long()
    vol[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = feats
    return vol


def test_conv_k3_matches_dense_conv3d():
    D, Cin, Cout = 6, 5, 7
    coords, feats = _cube(D, Cin, batch=2)
    w = torch.randn(27, Cin, Cout, generator=torch.Generator().manual_seed(3))
    x = ts.SparseTensor(feats, coords, 1)
    y = ts.conv3d(x, w, 3)
    wd = w.view(3, 3, 3, Cin, Cout).permute(4, 3, 2, 1, 0).contiguous()   # [co,ci,dx,dy,dz] <- k=(dz,dy,dx)
    ref = F.conv3d(_dense(coords, feats, D, 2), wd, padding=1)
    c = y.C.long()
    got = ref[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]]
    assert torch.allclose(y.F, got, atol=1e-4, rtol=1e-4)


def test_conv_k2s2_and_transpose_match_dense():
    D, Cin, Cout = 8, 4, 6
    coords, feats = _cube(D, Cin)
    g = torch.Generator().manual_seed(5)
    w = torch.randn(8, Cin, Cout, generator=g)
    x = ts.SparseTensor(feats, coords, 1)
    x.check()
    y = ts.conv3d(x, w, 2, stride=2)
    assert y.s == 2 and y.C.shape[0] == (D // 2) ** 3
    wd = w.view(2, 2, 2, Cin, Cout).permute(4, 3, 0, 1, 2).contiguous()   # [co,ci,dx,dy,dz] <- k=dx*4+dy*2+dz
    ref = F.conv3d(_dense(coords, feats, D), wd, stride=2)
    c = (y.C.long() // 2)
    assert torch.allclose(y.F, ref[0][:, c[:, 0], c[:, 1], c[:, 2]].t(), atol=1e-4, rtol=1e-4)
    # k2s2 map has exactly N_fine pairs (each fine voxel has one parent)
    assert y.kernel_maps["k2_os1_s2_d1"][0].shape[0] == coords.shape[0]
    # coarse coordinates ascend in hash
    h = ts.sphash(y.C)
    assert torch.all(h[1:] > h[:-1])
    # transposed conv back to stride 1
    wt = torch.randn(8, Cout, 3, generator=g)
    z = ts.conv3d(y, wt, 2, stride=2, transpose=True)
    assert z.s == 1 and torch.equal(z.C, coords)
    wtd = wt.view(2, 2, 2, Cout, 3).permute(3, 4, 0, 1, 2).contiguous()   # conv_transpose3d weight [ci,co,...]
    vol = torch.zeros(1, Cout, D // 2, D // 2, D // 2)
    vol[0][:, c[:, 0], c[:, 1], c[:, 2]] = y.F.t()
    reft = F.conv_transpose3d(vol, wtd, stride=2)
    cf = coords.long()
    assert torch.allclose(z.F, reft[0][:, cf[:, 0], cf[:, 1], cf[:, 2]].t(), atol=1e-4, rtol=1e-4)


def test_kernel_map_invariants(small_batch):
    coords = small_batch["coords"].int()
    x0 = og.initial_voxelize(ts.PointTensor(small_batch["feats"], coords.float()), 1, 1)
    nbr, pairs, counts = ts.build_kernel_map(x0.C, x0.C, 3, 1)
    n = x0.C.shape[0]
    assert counts.sum().item() == pairs.shape[0]
    assert torch.equal(nbr[13], torch.arange(n))                     # centre offset = identity
    for k in range(27):                                              # map(k) mirrors map(26-k)
        j = torch.nonzero(nbr[k] >= 0).flatten()
        assert torch.equal(nbr[26 - k][nbr[k][j]], j)
    # pair list is offset-major, out-ascending
    cur = 0
    for k in range(27):
        seg = pairs[cur:cur + counts[k]]
        assert torch.all(seg[1:, 1] > seg[:-1, 1])
        cur += counts[k]


def test_glue_properties(small_batch):
    coords, feats = small_batch["coords"], small_batch["feats"]
    z = ts.PointTensor(feats, coords.float())
    x0 = og.initial_voxelize(z, 1, 1)
    h = ts.sphash(x0.C)
    assert torch.all(h[1:] > h[:-1])                                 # ascending hash
    assert x0.C.shape[0] == coords.shape[0]                          # unique input => permutation
    idx = z.additional_features["idx_query"][1]
    assert torch.equal(x0.C[idx].long(), coords)
    assert torch.allclose(x0.F[idx], feats)
    z0 = og.voxel_to_point(x0, z)
    w = z.weights[1]
    assert torch.allclose(w[:, 0], torch.ones_like(w[:, 0]), atol=1e-6) and torch.all(w[:, 1:] == 0)  # one-hot at stride 1
    assert torch.allclose(z0.F, feats, atol=1e-6)
    x1 = og.point_to_voxel(x0, z0)
    assert torch.allclose(x1.F, x0.F, atol=1e-6)                     # p2v o v2p == id at stride 1


def test_conv_gradcheck_fp64():
    coords, feats = _cube(3, 2)
    feats = feats.double().requires_grad_(True)
    w = torch.randn(27, 2, 3, dtype=torch.float64, requires_grad=True)
    _, pairs, counts = ts.build_kernel_map(coords, coords, 3, 1)
    n = coords.shape[0]
    assert torch.autograd.gradcheck(lambda f, k: ts.sparseconv(f, k, pairs, counts, (n, n), False), (feats, w))


def test_lift_matches_plain_indexing():
    g = torch.Generator().manual_seed(0)
    fmap = torch.randn(2, 6, 9, 11, generator=g)
    idx = [torch.stack([torch.randint(0, 9, (20,), generator=g), torch.randint(0, 11, (20,), generator=g)], 1).numpy()
           for _ in range(2)]
    out = og.lift(fmap, idx)
    for b in range(2):
        for p in range(20):
            r, c = idx[b][p]
            assert torch.equal(out[b * 20 + p], fmap[b, :, r, c])


@pytest.mark.parametrize("name", ["quantize_small", "kmap_small", "model_small"])
def test_golden_fixtures(name):
    """The committed fixtures were produced by tests/golden/make_golden.py from this oracle; they freeze its
    behaviour so that GPU parity tests and later oracle edits are checked against a fixed artefact."""
    from tests.golden import make_golden
    path = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(path), "run python -m tests.golden.make_golden"
    want = np.load(path)
    got = getattr(make_golden, "gen_" + name)()
    for k in want.files:
        if want[k].dtype.kind == "f":
            np.testing.assert_allclose(got[k], want[k], rtol=2e-4, atol=2e-5, err_msg=k)
        else:
            np.testing.assert_array_equal(got[k], want[k], err_msg=k)


# ---- vectors produced by the reference's own code (tests/golden/make_reference_golden.py) -------------------------
@pytest.mark.parametrize("tag", ["nus", "kitti"])
def test_oracle_voxelize_matches_reference_augment_and_scale(tag):
    """a1: the oracle's restatement against FusionTransformer/data/utils/augmentation_3d.py run on the same points."""
    from oracle import ft_glue as og
    g = np.load(os.path.join(GOLD, "ref_voxelize.npz"))
    vc, keep, inds, inv = og.voxelize_scan(g[tag + "_points"])
    np.testing.assert_array_equal(keep, g[tag + "_keep"])
    np.testing.assert_array_equal(vc, g[tag + "_coords"][g[tag + "_keep"]])
    if tag == "kitti":
        assert 0 < keep.sum() < len(keep)                 # the bounds filter is exercised
    np.testing.assert_array_equal(vc[inds][inv], vc)      # a2 round trip on the reference's coordinates


def _aug_params(g, tag):
    return dict(noisy_rot=float(g[tag + "_noisy_rot"]), flip_x=float(g[tag + "_flip_x"]), flip_y=float(g[tag + "_flip_y"]),
                rot_z=float(g[tag + "_rot_z"]), transl=bool(g[tag + "_transl"]))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_oracle_augmentation_matches_reference_augment_and_scale(tag):
    """a1 with the augmentation branch on (augmentation_3d.py:22-51): the oracle's restatement, and the product's
    host-side draws (utils/augment.py), against the reference function run under the same seeded numpy generator --
    bit-identical float32 coordinates, identical draws."""
    from fusiontransformer_b200.utils import augment
    from oracle import ft_glue as og
    g = np.load(os.path.join(GOLD, "ref_augment.npz"))
    kw = _aug_params(g, tag)
    np.random.seed(int(g[tag + "_seed"]))
    rot, u = og.augment_draws(**kw)
    coords = og.augment_and_scale(g[tag + "_points"].copy(), 20, 4096, rot, u)
    assert coords.dtype == np.float32
    np.testing.assert_array_equal(coords, g[tag + "_coords_float"])
    np.random.seed(int(g[tag + "_seed"]))
    rot_p, u_p = augment.draw(**kw)
    assert (rot is None) == (rot_p is None) and (u is None) == (u_p is None)
    if rot is not None:
        assert rot_p.dtype == np.float32
        np.testing.assert_array_equal(rot_p, rot)
    if u is not None:
        np.testing.assert_array_equal(u_p, u)


def test_oracle_unmap_matches_reference_map_sparse_to_org():
    from oracle import ft_glue as og
    g = np.load(os.path.join(GOLD, "ref_segiou.npz"))
    pred = torch.from_numpy(g["logits1"]).argmax(1)
    np.testing.assert_array_equal(og.map_sparse_to_org(pred, torch.from_numpy(g["inverse_map"])).numpy(), g["pred_points"])


def test_oracle_collate_matches_reference_collate_scn_base():
    """a3: batch-index column, concatenation order and dtypes of FusionTransformer/data/collate.py:36-67."""
    from oracle import ft_glue as og
    g = np.load(os.path.join(GOLD, "ref_collate.npz"))
    st = og.collate([dict(coords=g["coords%d" % i], feats=g["feats%d" % i]) for i in range(3)])
    assert st.C.dtype == torch.int64 and tuple(st.C.shape) == tuple(g["C"].shape)
    np.testing.assert_array_equal(st.C.numpy(), g["C"])
    np.testing.assert_array_equal(st.F.numpy(), g["F"])


# ---- random occupancy: the sparse convolutions equal dense conv3d on the zero-filled grid, read at the active sites
def _random_voxels(D, frac, C, seed, batch=2):
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand(batch, D, D, D, generator=g) < frac
    b, x, y, z = torch.nonzero(occ, as_tuple=True)
    perm = torch.randperm(b.numel(), generator=g)                     # arbitrary input order, as after a dataloader
    coords = torch.stack([x, y, z, b], 1)[perm].int()
    feats = torch.randn(coords.shape[0], C, generator=g)
    return coords, feats


@pytest.mark.parametrize("seed,frac", [(0, 0.08), (1, 0.3), (2, 0.7)])
def test_sparse_conv_k3_random_occupancy_matches_dense(seed, frac):
    D, Cin, Cout = 9, 3, 5
    coords, feats = _random_voxels(D, frac, Cin, seed)
    w = torch.randn(27, Cin, Cout, generator=torch.Generator().manual_seed(seed + 10))
    y = ts.conv3d(ts.SparseTensor(feats, coords, 1), w, 3)
    assert torch.equal(y.C, coords)                                   # submanifold: outputs = inputs, same order
    wd = w.view(3, 3, 3, Cin, Cout).permute(4, 3, 2, 1, 0).contiguous()
    ref = F.conv3d(_dense(coords, feats, D, 2), wd, padding=1)
    c = coords.long()
    assert torch.allclose(y.F, ref[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("seed,frac", [(3, 0.1), (4, 0.5)])
def test_strided_and_transposed_conv_random_occupancy_match_dense(seed, frac):
    D, Cin, Cmid = 8, 4, 6
    coords, feats = _random_voxels(D, frac, Cin, seed, batch=1)
    g = torch.Generator().manual_seed(seed + 20)
    w = torch.randn(8, Cin, Cmid, generator=g)
    x = ts.SparseTensor(feats, coords, 1)
    x.check()                                                         # registers coord_maps[1] (models/utils.py:31)
    y = ts.conv3d(x, w, 2, stride=2)
    # coarse sites = unique parents of the active fine voxels, in ascending hash order
    parents = torch.unique((coords[:, :3].long() // 2) * 2, dim=0)
    assert y.C.shape[0] == parents.shape[0]
    h = ts.sphash(y.C)
    assert torch.all(h[1:] > h[:-1])
    wd = w.view(2, 2, 2, Cin, Cmid).permute(4, 3, 0, 1, 2).contiguous()
    ref = F.conv3d(_dense(coords, feats, D), wd, stride=2)
    c = y.C.long() // 2
    assert torch.allclose(y.F, ref[0][:, c[:, 0], c[:, 1], c[:, 2]].t(), atol=1e-4, rtol=1e-4)
    # transposed conv writes only at the ORIGINAL fine sites (cached encoder coordinates), values = dense transposed conv
    wt = torch.randn(8, Cmid, 3, generator=g)
    z = ts.conv3d(y, wt, 2, stride=2, transpose=True)
    assert torch.equal(z.C, coords)
    vol = torch.zeros(1, Cmid, D // 2, D // 2, D // 2)
    vol[0][:, c[:, 0], c[:, 1], c[:, 2]] = y.F.t()
    reft = F.conv_transpose3d(vol, wt.view(2, 2, 2, Cmid, 3).permute(3, 4, 0, 1, 2).contiguous(), stride=2)
    cf = coords.long()
    assert torch.allclose(z.F, reft[0][:, cf[:, 0], cf[:, 1], cf[:, 2]].t(), atol=1e-4, rtol=1e-4)


def test_sparse_quantize_properties_hypothesis():
    """A.1 invariants on arbitrary integer clouds: inds = first occurrence of each distinct voxel, keys ascending,
    inverse reconstructs every row, counts>1 voxels get ignore_label."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 6), st.integers(0, 6), st.integers(0, 6)), min_size=1, max_size=200))
    def check(rows):
        c = np.asarray(rows, dtype=np.int64)
        lab = np.arange(len(c)) % 20
        inds, labels, inv = ts.sparse_quantize(c, np.zeros((len(c), 1), np.float32), lab, return_index=True,
                                               return_invs=True)
        uniq = c[inds]
        assert len({tuple(r) for r in uniq}) == len(uniq) == len({tuple(r) for r in c})
        assert np.array_equal(uniq[inv], c)
        keys = ts.fnv_hash_vec(uniq)
        assert np.all(keys[1:] > keys[:-1])
        for j, i in enumerate(inds):                                   # first occurrence
            assert i == min(k for k in range(len(c)) if tuple(c[k]) == tuple(uniq[j]))
        cnt = np.bincount(inv, minlength=len(inds))
        assert np.array_equal(labels[cnt > 1], np.full((cnt > 1).sum(), -100))
        assert np.array_equal(labels[cnt == 1], lab[inds][cnt == 1])
    check()
