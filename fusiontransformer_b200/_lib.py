"""ctypes binding of libft3d.so -- the C ABI declared in include/ft3d.h.

The header is the single source of truth: its prototypes are parsed here to set ctypes
argtypes/restypes, so a signature change cannot silently desynchronise the Python side.  There is
no fallback: if the library is missing or a symbol is absent, importing a product op raises.
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "ft3d.h"
LIB_PATH = PKG / "libft3d.so"

_SCALARS = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t, "float": ctypes.c_float, "double": ctypes.c_double,
    "ft3d_stream_t": ctypes.c_void_p,
}
_PROTO = re.compile(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**)\s*(ft3d_[A-Za-z0-9_]+)\s*\(([^)]*)\)\s*;", re.M | re.S)


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if re.match(r"const\s+char\s*\*", decl) else ctypes.c_void_p
    base = decl.replace("const", "").strip().split()[0]
    return _SCALARS[base]


def parse_header(path: Path = HEADER):
    """Returns {symbol: (restype, [argtypes])} for every prototype in the header."""
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for ret, name, args in _PROTO.findall(text):
        args = args.strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) but keep pointer stars
                m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a, re.S)
                argtypes.append(_ctype(m.group(1) if m and m.group(1).strip() else a))
        out[name] = (_ctype(ret), argtypes)
    return out


class Ft3dError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not LIB_PATH.exists():
            raise Ft3dError(
                "libft3d.so not found at %s -- build it with `python -m fusiontransformer_b200.build` "
                "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        self.cdll = ctypes.CDLL(str(LIB_PATH))
        self.protos = parse_header()
        self.calls = {}          # entry point -> number of calls (bench.py turns this into a launch count)
        self.profile = None      # when a list: (entry point, start event, end event, repeats) appended per call
        self.profile_repeat = {}  # while profiling: idempotent entry point -> back-to-back launches between the events
        for name, (res, args) in self.protos.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError as e:
                raise Ft3dError("libft3d.so does not export %s (declared in include/ft3d.h)" % name) from e
            fn.restype, fn.argtypes = res, args
            if res is ctypes.c_int and name != "ft3d_version":
                setattr(self, name[len("ft3d_"):], self._checked(fn, name))
            else:
                setattr(self, name[len("ft3d_"):], fn)

    def _checked(self, fn, name):
        last_error = self.cdll.ft3d_last_error
        short = name[len("ft3d_"):]

        def call(*a):
            self.calls[short] = self.calls.get(short, 0) + 1
            prof = self.profile
            if prof is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = fn(*a)
            if prof is not None:
                reps = self.profile_repeat.get(short, 1)
                for _ in range(reps - 1):        # amortises the ~8 us of event overhead around a single short launch
                    fn(*a)
                e1.record()
                prof.append((short, e0, e1, reps))
                if short == "conv_os" and reps > 1:
                    # the dominant kernel on its own: the same call again with its fold / finalize launches disabled
                    # (FT3D_OS_DEBUG=8; the complete call above has already produced the result, these rewrite it)
                    import os
                    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    os.environ["FT3D_OS_DEBUG"] = "8"
                    try:
                        e2.record()
                        for _ in range(reps):
                            fn(*a)
                        e3.record()
                    finally:
                        os.environ.pop("FT3D_OS_DEBUG", None)
                    prof.append(("conv_os_kernel", e2, e3, reps))
            if rc != 0:
                last_error.restype = ctypes.c_char_p
                raise Ft3dError("%s failed (%d): %s" % (name, rc, (last_error() or b"").decode()))
        call.__name__ = name
        return call


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


# CUDA kernels launched by one call of each entry point (fixed by csrc/*.cu; cub passes counted from its
# onesweep/scan implementations).  Used only to report `gpu_launches`.
LAUNCHES = {
    "hash": 1, "kernel_hash": 1, "scale_coords": 3, "quantize": 14, "unique": 9, "coarsen_hash": 1,
    "gather_rows_i32": 1, "table_build": 2, "table_query": 1, "kmap_build": 1, "kmap_pairs": 3, "kmap_transpose": 2,
    "count": 2, "voxelize_fwd": 2, "voxelize_bwd": 1, "devoxelize_fwd": 1, "devoxelize_bwd": 2, "ti_weights": 1,
    "v2p_build": 1, "p2v_build": 2, "lift_fwd": 1, "lift_bwd": 1, "conv_gather_f32": 1, "conv_wgrad_f32": 1,
    "conv_pack_weights": 1, "to_bf16": 1, "conv_pairs_tc": 1,
    "conv_reduce": 1, "conv_wgrad_pairs_tc": 1, "kmap_pair_positions": 3, "conv_reduce_bn": 2, "bn_stats": 2,
    "bn_apply": 1, "col_sum": 2, "seg_loss": 3, "confusion_update": 1, "bn_bwd_reduce": 2, "conv_pack_weights_multi": 1, "bn_bwd_apply": 1,
    "conv_os": 3, "conv_os_plan": 13, "conv_wgrad_pairs_tc_det": 2,
}


def launch_count(calls: dict) -> int:
    return sum(LAUNCHES.get(k, 1) * v for k, v in calls.items())
