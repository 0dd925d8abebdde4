"""Runs the small-case parity checks of tests/test_gpu_parity.py against the cuemu build of the
library compiled with -fsanitize=address,undefined (device memory is host heap there, so an
out-of-bounds load or store of a kernel, a misaligned vector access or signed overflow in an
index computation is reported with a stack trace instead of silently reading garbage on the GPU).

    CXX=/usr/bin/g++ QBX_EMU_SANITIZE=address,undefined python tools/cuemu/build_emu.py
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/cuemu/sanitize_check.py
TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from quiqbox_b200 import lib  # noqa: E402

L = C.CDLL(os.path.join(HERE, "_build_san", "libqbx_emu.so"))
for name, args in lib.SIGNATURES.items():
    fn = getattr(L, name)
    fn.argtypes = args
    fn.restype = C.c_int
L.qbx_last_error.restype = C.c_char_p
L.qbx_last_error.argtypes = []
lib._lib = L

import test_gpu_parity as P  # noqa: E402
from molecules import h2, h2o  # noqa: E402

checks = [("boys", lambda: (P.test_boys_golden_points_generic_kernel(), P.test_boys_table_vs_oracle())),
          ("generic", lambda: (P.test_primitive_golden_eris_any_l(), P.test_lih_tensor_symmetry_and_one_body())),
          ("tensor H2", lambda: P.test_full_tensor_vs_oracle("H2/STO-3G", h2(1.4), "STO-3G")),
          ("tensor H2O 6-31G", lambda: P.test_full_tensor_vs_oracle("H2O/6-31G", h2o(), "6-31G")),
          ("tensor H2O cc-pVDZ", lambda: P.test_full_tensor_vs_oracle("H2O/cc-pVDZ", h2o(), "cc-pVDZ")),
          ("fock 6-31G", lambda: P.test_fock_build_modes_vs_oracle_getGcore("6-31G")),
          ("fock cc-pVDZ", lambda: P.test_fock_build_modes_vs_oracle_getGcore("cc-pVDZ")),
          ("sharded", P.test_sharded_partial_G_sums_to_full),
          ("scf", lambda: (P.test_hoh_sto3g_scf(), P.test_h2o2_631g_scf())),
          ("boundary", P.test_boundary_errors),
          ("irregular", P.test_irregular_basis_falls_back_to_generic_kernels)]
checks += [(f"synthetic {c}", (lambda c=c: P.test_synthetic_class_batch_vs_oracle(c))) for c in P.CLASSES]
only = sys.argv[1:]
for name, fn in checks:
    if only and not any(o in name for o in only):
        continue
    t = time.time()
    fn()
    print(f"ok  {name}  ({time.time() - t:.1f} s)", flush=True)
print("sanitize_check: all clean")
