"""The C-ABI boundary: libft3d.so loads on a CPU-only box and exports every symbol include/ft3d.h declares."""
import ctypes
import os

import pytest

from fusiontransformer_b200 import _lib


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for must in ("ft3d_hash", "ft3d_quantize", "ft3d_kmap_build", "ft3d_kmap_pairs", "ft3d_conv_os",
                 "ft3d_conv_os_plan", "ft3d_conv_gather_f32", "ft3d_devoxelize_fwd", "ft3d_lift_fwd",
                 "ft3d_last_error", "ft3d_version"):
        assert must in protos, must
    res, args = protos["ft3d_hash"]
    assert res is ctypes.c_int and args == [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]


def test_library_exports_every_declared_symbol():
    if not _lib.LIB_PATH.exists():
        from fusiontransformer_b200.build import build_library
        build_library()
    cdll = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _lib.parse_header():
        assert hasattr(cdll, name), "libft3d.so does not export %s" % name
    cdll.ft3d_version.restype = ctypes.c_int
    assert cdll.ft3d_version() == 100


def test_host_only_entry_points():
    L = _lib.lib()
    assert L.table_capacity(1000) == 2048 and L.table_capacity(0) == 1024
    assert L.conv_packed_bytes(27, 96, 128) == 27 * 2 * 128 * 128
    assert L.unique_workspace(1000) > 1000 * 8
    # argument validation happens before any CUDA call, so it is testable without a GPU
    with pytest.raises(_lib.Ft3dError, match="bad arguments"):          # K > 32
        L.conv_os(256, 10, 256, 256, 256, 256, 256, 256, 8, 1, 128, 0, 40, 0, 64, 64, 256, 256, 128, None, 0.0, 0.0, None,
                  None, None, None, 0, None, None)
    with pytest.raises(_lib.Ft3dError, match="unsupported shape"):      # red % 16 != 0
        L.conv_os(256, 10, 256, 256, 256, 256, 256, 256, 8, 1, 128, 0, 27, 0, 20, 64, 256, 256, 128, None, 0.0, 0.0, None,
                  None, None, None, 0, None, None)
    with pytest.raises(_lib.Ft3dError, match="bad arguments"):          # tile_rows must be 128 x {1, 2, 4}
        L.conv_os(256, 10, 256, 256, 256, 256, 256, 256, 8, 1, 384, 0, 27, 0, 64, 64, 256, 256, 128, None, 0.0, 0.0, None,
                  None, None, None, 0, None, None)
    assert L.conv_os_plan_workspace(1000, 32) > 5 * 1000 * 4
    assert L.conv_os_workspace(128, 4, 256) >= 4 * 256 * 128 * 4


def test_no_cpu_fallback():
    import torch
    from fusiontransformer_b200 import ops
    with pytest.raises(_lib.Ft3dError, match="CUDA tensor"):
        ops.hash_coords(torch.zeros(4, 4, dtype=torch.int32))


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "fusiontransformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_error_convention_of_the_hot_path_entry_points():
    """0 = ok, non-zero + thread-local message otherwise; every check runs before the first CUDA call, so unsupported
    shapes are rejected (never routed to another implementation) even on a box without a GPU."""
    L = _lib.lib()
    p = 256          # any non-null 16-byte aligned "pointer": validation does not dereference
    with pytest.raises(_lib.Ft3dError, match="unsupported shape"):
        L.conv_pairs_tc(p, p, p, 27, 0, 1000, 20, 64, p, p, None)                 # red not a multiple of 16
    with pytest.raises(_lib.Ft3dError, match="unsupported shape"):
        L.conv_pairs_tc(p, p, p, 27, 0, 1000, 64, 320, p, p, None)                # ncols: <= 256 or 384 only
    with pytest.raises(_lib.Ft3dError, match="identity gather needs K == 1"):
        L.conv_pairs_tc(p, None, None, 27, 0, 1000, 64, 64, p, p, None)
    with pytest.raises(_lib.Ft3dError, match="16-byte aligned"):
        L.conv_pairs_tc(p + 4, p, p, 27, 0, 1000, 64, 64, p, p, None)
    with pytest.raises(_lib.Ft3dError, match="unsupported shape"):
        L.conv_wgrad_pairs_tc(p, p, p, p, 27, 0, 64, 512, 1000, p, None)          # cout <= 256
    with pytest.raises(_lib.Ft3dError, match="at least one row"):
        L.bn_stats(p, 0, 64, 1e-5, 0.1, p, None, None, None, p, 1 << 20, None)
    with pytest.raises(_lib.Ft3dError, match="workspace too small"):
        L.bn_stats(p, 10, 64, 1e-5, 0.1, p, None, None, None, p, 16, None)
    with pytest.raises(_lib.Ft3dError, match="bad arguments"):
        L.seg_loss(p, p, 100, 100, -100, None, None, 0.0, None, p, p, p, 1 << 20, None)   # > 64 classes
    with pytest.raises(_lib.Ft3dError, match="needs teacher logits"):
        L.seg_loss(p, p, 100, 20, -100, None, None, 0.5, None, p, p, p, 1 << 20, None)
    # zero-size work is a no-op that succeeds without touching the device
    L.conv_pairs_tc(p, p, p, 27, 0, 0, 64, 64, p, p, None)
    L.bn_apply(p, 0, 64, p, p, p, None, 1, p, None, 0, None, None)
    with pytest.raises(_lib.Ft3dError, match="output pitch"):
        L.bn_apply(p, 10, 64, p, p, p, None, 1, p, None, 32, None, None)            # pitch smaller than the row
    L.confusion_update(p, p, 0, 20, -100, None, p, None)
    with pytest.raises(_lib.Ft3dError, match="too small"):                          # load factor <= 1/2, whatever n
        L.table_build(p, 100, p, p, 128, None)
    with pytest.raises(_lib.Ft3dError, match="power of two"):
        L.table_build(p, 10, p, p, 100, None)
    assert int(L.seg_loss_workspace()) >= 3 * 1024 * 8 and int(L.bn_workspace(256)) > 0


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/ft3d.h compiles as C99 and a C program links against libft3d.so (examples/abi_probe.c)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_probe")
    libdir = os.path.dirname(str(_lib.LIB_PATH))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "abi_probe.c"), "-L", libdir, "-lft3d", "-Wl,-rpath," + libdir,
                    "-o", exe], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "ft3d version 100" in out and "table_capacity(1000) = 2048" in out
    assert "bad shape -> rc 1" in out and "unsupported shape" in out
