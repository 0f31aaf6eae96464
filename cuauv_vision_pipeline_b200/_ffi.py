"""cffi ABI-mode binding of libb200vision.so, in the style of the reference's own binding
(core/bindings/camera_message_framework.py:11-70: `ffi.cdef(...)` mirroring the C header, then
`ffi.dlopen(...)`, int status codes wrapped on the Python side).

The cdef text is taken from include/b200vision.h itself, so the header is the single source of
truth for the ABI.  There is no fallback: if the shared library is missing and cannot be built,
or no CUDA device is usable, importing / creating a context raises.
"""
import os
import re

import cffi

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "b200vision.h")
LIB_PATH = os.path.join(_HERE, "lib", "libb200vision.so")


def _cdef_text():
    with open(HEADER) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # comments
    lines = []
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("#define BV_VERSION"):
            lines.append(s)
        elif s.startswith("#") or s.startswith('extern "C"') or s == "}":
            continue
        else:
            lines.append(line)
    return "\n".join(lines)


ffi = cffi.FFI()
ffi.cdef(_cdef_text())


def _load():
    # build when the library is missing, and re-build when a source is newer than it and nvcc is here (a stale binary
    # would silently run old kernels); on a box without nvcc the prebuilt library that travelled with the tree is used
    from . import build as _build
    if not os.path.exists(LIB_PATH) or (_build.needs_build() and os.path.exists(_build.NVCC)):
        _build.build()
    return ffi.dlopen(LIB_PATH)


lib = _load()


class BVError(RuntimeError):
    """A bv_* call returned a negative bv_status."""

    def __init__(self, status, message):
        super().__init__("b200vision error %d: %s" % (status, message))
        self.status = status


def check(status):
    if status != 0:
        raise BVError(status, ffi.string(lib.bv_last_error()).decode("utf-8", "replace"))
    return status


def exported_symbols():
    """Function names declared in include/b200vision.h (tests check the .so exports each one)."""
    txt = _cdef_text()
    return sorted(set(re.findall(r"\b(bv_[a-z0-9_]+|process_frame)\s*\(", txt)))
