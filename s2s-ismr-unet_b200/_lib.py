"""ctypes binding of libs2s_unet.so, generated from include/s2s_unet.h.

The prototypes are parsed from the public header so the Python side cannot drift from the C ABI.
There is no CPU fallback: if the library is missing and cannot be built, or a call fails, an
exception is raised (`S2SError`).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "s2s_unet.h"
LIB_PATH = PKG / "lib" / "libs2s_unet.so"


class S2SError(RuntimeError):
    pass


class UnetCfg(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("filters", C.c_int32),
                ("n_blocks", C.c_int32), ("ct_kernel", C.c_int32), ("pool", C.c_int32), ("bn", C.c_int32),
                ("head", C.c_int32), ("max_batch", C.c_int32), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
                ("precision", C.c_int32), ("act", C.c_int32)]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("arena", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int32 * 4),
                ("offset", C.c_int64), ("count", C.c_int64)]


class AdamCfg(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double)]


_SCALARS = {"int": C.c_int, "float": C.c_float, "double": C.c_double, "size_t": C.c_size_t, "int64_t": C.c_int64, "int32_t": C.c_int32}
_PTR_TO_PTR = C.POINTER(C.c_void_p)


def _ctype(decl: str):
    """C parameter declaration (without the name) -> ctypes type."""
    d = decl.replace("const", "").strip()
    d = re.sub(r"\s+", " ", d)
    if d in _SCALARS:
        return _SCALARS[d]
    if d == "char*":
        return C.c_char_p
    if d.endswith("**"):
        return _PTR_TO_PTR
    if d == "s2s_unet_cfg*":
        return C.POINTER(UnetCfg)
    if d == "s2s_tensor_desc*":
        return C.POINTER(TensorDesc)
    if d == "s2s_adam_cfg*":
        return C.POINTER(AdamCfg)
    if d in ("int*",):
        return C.POINTER(C.c_int)
    if d == "size_t*":
        return C.POINTER(C.c_size_t)
    if d == "int64_t*":
        return C.POINTER(C.c_int64)
    if d.endswith("*"):
        return C.c_void_p          # device / host buffers, streams, events, handles
    raise ValueError(f"unmapped C type in header: {decl!r}")


def parse_header(path: Path = HEADER):
    """[(name, restype, [argtypes])] for every function declared in the header."""
    src = path.read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = []
    for m in re.finditer(r"\b(int|const\s+char\s*\*)\s+(s2s_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = re.sub(r"\s+", " ", args).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\s*\b\w+$", "", a) if not a.endswith("*") else a   # drop the parameter name
                a = a.replace(" *", "*").replace("* ", "*")
                argtypes.append(_ctype(a))
        protos.append((name, C.c_int if ret == "int" else C.c_char_p, argtypes))
    return protos


_lib = None


def _build_if_needed() -> None:
    from . import build as _b
    _b.build()


def load():
    """Load (building first if the sources are newer) and bind every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("S2S_LIB"):        # experiment hook: load another build of the same ABI (same-box A/B of two builds)
        LIB_PATH = Path(os.environ["S2S_LIB"]).resolve()
    elif os.environ.get("S2S_NO_BUILD") != "1":
        try:
            _build_if_needed()
        except Exception as e:   # no nvcc on the box: fall through to the prebuilt .so
            if not LIB_PATH.exists():
                raise S2SError(f"libs2s_unet.so is missing and could not be built: {e}") from e
    if not LIB_PATH.exists():
        raise S2SError(f"{LIB_PATH} not found: build it with `python s2s-ismr-unet_b200/build.py` "
                       "(the CUDA extension is mandatory, there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, restype, argtypes in parse_header():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise S2SError(f"libs2s_unet.so does not export {name} declared in {HEADER.name}") from e
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().s2s_last_error()
        raise S2SError(f"s2s error {rc}: {msg.decode() if msg else '?'}")


def call(name: str, *args):
    check(getattr(load(), name)(*args))
