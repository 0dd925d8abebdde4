"""Test helper: run the library's kernels under the cuemu host emulation (tools/cuemu).

TEST INFRASTRUCTURE.  `install()` builds tools/cuemu/_build/libqbx_emu.so from the unmodified
sources of quiqbox.jl_b200/csrc (g++, CUDA execution model emulated with coroutines) and makes the
Python host mirror talk to it instead of libqbx.so; `uninstall()` restores the product binding.
Only tests/ call this; the product path (quiqbox.jl_b200/lib.py) knows nothing about it."""
import ctypes as C
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

_saved = []


def build():
    spec = importlib.util.spec_from_file_location("qbx_build_emu", os.path.join(ROOT, "tools", "cuemu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def install():
    from quiqbox_b200 import lib
    path = build()
    L = C.CDLL(path)
    for name, args in lib.SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.qbx_last_error.restype = C.c_char_p
    L.qbx_last_error.argtypes = []
    _saved.append(lib._lib)
    lib._lib = L
    return L


def uninstall():
    from quiqbox_b200 import lib
    lib._lib = _saved.pop() if _saved else None
