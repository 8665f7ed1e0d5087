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

        def call(*a):
            rc = fn(*a)
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
