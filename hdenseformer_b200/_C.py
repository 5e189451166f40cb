"""ctypes binding of libhdf_b200.so (include/hdf_b200.h).  The product path has no fallback:
if the library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhdf_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "hdf_b200.h")

F32, BF16 = 0, 1

_lib = None

_CTYPE = {
    "int": C.c_int, "long long": C.c_longlong, "float": C.c_float, "size_t": C.c_size_t,
    "unsigned": C.c_uint, "unsigned long long": C.c_ulonglong,
}


class HDFError(RuntimeError):
    pass


def declared_symbols(header: str = HEADER):
    """Parse `name -> (restype, [argtypes])` from the C header (single source of truth)."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t|const char\*|unsigned long long)\s+(hdf_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(C.c_void_p)
                else:
                    ty = " ".join(a.split()[:-1])
                    argtypes.append(_CTYPE[ty])
        restype = {"int": C.c_int, "size_t": C.c_size_t, "const char*": C.c_char_p, "unsigned long long": C.c_ulonglong}[ret]
        out[name] = (restype, argtypes)
    return out


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HDFError(f"{LIB_PATH} not found: build it with `python -m hdenseformer_b200.build` "
                       "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in declared_symbols().items():
        fn = getattr(lib, name)  # AttributeError if the header declares a symbol the .so lacks
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().hdf_last_error_string().decode()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise HDFError(f"{what} failed (status {rc}): {last_error()}")


_inited = set()


def init(device: int):
    if device not in _inited:
        check(load().hdf_init(device), "hdf_init")
        _inited.add(device)
