"""ctypes binding of the exvae_b200 C ABI (include/exvae_b200.h).

The prototypes are parsed from the header itself, so the Python binding cannot drift from the
declared ABI.  There is NO fallback: if ``csrc/libexvae_b200.so`` is missing or does not
export a declared symbol, importing the product raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "exvae_b200.h")
LIB_PATH = os.path.join(_HERE, "csrc", "libexvae_b200.so")

EXVAE_OK = 0
ACT_NONE, ACT_SIGMOID, ACT_HARDTANH, ACT_RELU = 0, 1, 2, 3

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "exvae_stream_t": ctypes.c_void_p,
}


class ExvaeError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Return {symbol: (restype, [argtypes])} for every EXVAE_API declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"EXVAE_API\s+(const char\*|int|size_t)\s+(exvae_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p}[ret]
        argtypes = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                ty = a.replace("const ", "").split()[0]
                argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ExvaeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. "
                "Run `python -m exemplar_vae_b200.build` (needs nvcc); there is no CPU fallback.")
        self._dll = ctypes.CDLL(LIB_PATH)
        self.profile = None
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:
                raise ExvaeError(f"{LIB_PATH} does not export {name} declared in include/exvae_b200.h") from e
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, self._wrap(name, fn) if restype is ctypes.c_int else fn)
        if self.exvae_abi_version() != 1:
            raise ExvaeError("exvae_b200 ABI version mismatch between header and library")

    def _wrap(self, name, fn):
        """Optional per-entry-point device timing: when ``self.profile`` is a list, every call is
        bracketed by CUDA events on the current stream (bench.py uses this for the roofline)."""
        def call(*args):
            prof = self.profile
            if prof is None:
                return fn(*args)
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            prof.append((name, e0, e1, args))
            return rc
        call.__name__ = name
        return call

    def check(self, rc: int, what: str = ""):
        if rc != EXVAE_OK:
            msg = self.exvae_error_string(rc)
            raise ExvaeError(f"{what or 'exvae call'} failed with code {rc}: {msg.decode() if msg else '?'}")


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
