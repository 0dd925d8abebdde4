"""CPU-side checks of the boundary: libqbx.so loads and exports exactly the symbols
include/qbx.h declares (no compute call is made -- there is no GPU here), the product
package never reaches into oracle/, and host-side basis logic."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import quiqbox_b200 as qb
from quiqbox_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return L.load()


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "qbx.h")).read()
    declared = set(re.findall(r"\b(qbx_[a-z0-9_]+)\s*\(", hdr))
    assert declared >= {"qbx_init", "qbx_basis_create", "qbx_eri_tensor", "qbx_eri_quartets", "qbx_eri_store",
                        "qbx_fock_build", "qbx_fock_build_device", "qbx_one_body", "qbx_boys", "qbx_prim_batch"}
    # the WHOLE dynamic symbol table (functions and data, any binding), not just the qbx_* names
    nm = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in nm.splitlines() if ln.strip()}
    assert declared == exported, (declared - exported, sorted(exported - declared)[:20])
    for name in declared:
        assert isinstance(getattr(built, name), ctypes._CFuncPtr)
    assert set(L.SIGNATURES) | {"qbx_last_error"} == declared


def test_argument_errors_need_no_device(built):
    # argument validation happens before any CUDA call
    assert built.qbx_basis_create(0, None, None, None, 0, None, None, None, None) == 1
    assert b"null or empty" in built.qbx_last_error()
    assert built.qbx_boys(1, None, 200, 0, None) == 1
    assert built.qbx_prim_batch(0, 1, 0, 0, 1, 10, 0, None, None, None, 0, None, None) == 1


def test_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.QbxError):
        qb.DeviceBasis(qb.genGaussTypeOrbSeq((0, 0, 0), "H", "STO-3G"))


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "quiqbox.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "qbx_oracle" not in txt and "import oracle" not in txt, f


def test_basis_flattening_and_component_order():
    assert qb.SubshellXYZs(2) == [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]   # Lexicons.jl:21
    bs = qb.genGaussTypeOrbSeq((0.0, 0.0, 0.0), "O", "cc-pVDZ")
    assert len(bs) == 3 + 6 + 6                                  # 3 s, 2 p shells, 1 Cartesian d shell
    m = qb.MultiOrbitalData.from_orbitals(bs)
    # general contraction: the two 9-term s functions and the 1-term s share primitives (indexGetOrbCore!)
    assert m.nprim == 9 + 3 * 4 + 6 * 1 and m.nbf == 15
    assert m.bf_off[-1] == 9 + 9 + 1 + 3 * 4 + 3 * 1 + 6
    sp = qb.genGaussTypeOrbSeq((0.0, 0.0, 0.0), "Li", "3-21G")
    assert [sum(g.ang) for g in sp] == [0, 0, 1, 1, 1, 0, 1, 1, 1]   # SP shells: S first, then P (:531-541)


def test_nuclear_cluster_sorting():
    c = qb.NuclearCluster(["O", "H", "H"], [(0, 0, 0), (1, 0, 0), (-1, 0, 0)])
    assert c.syms == ["H", "H", "O"] and c.coords[0] == (-1.0, 0.0, 0.0)        # Particles.jl:29-55
    assert qb.nucRepulsion(c) == pytest.approx(8 + 8 + 0.5)


def test_communicator_single_rank_needs_no_nccl(built):
    """qbx_comm_init(0, 1, NULL) is legal without NCCL or a GPU; a multi-rank communicator needs the 128-byte id."""
    r, n = ctypes.c_int(-1), ctypes.c_int(-1)
    assert built.qbx_comm_destroy() == 0
    assert built.qbx_comm_info(ctypes.byref(r), ctypes.byref(n)) == 0 and (r.value, n.value) == (0, 1)
    assert built.qbx_comm_init(0, 1, None) == 0
    assert built.qbx_comm_info(ctypes.byref(r), ctypes.byref(n)) == 0 and (r.value, n.value) == (0, 1)
    assert built.qbx_comm_init(0, 2, None) == 1 and b"bad argument" in built.qbx_last_error()
    assert built.qbx_comm_init(3, 2, None) == 1
    assert built.qbx_comm_unique_id(None) == 1
    assert built.qbx_comm_destroy() == 0
