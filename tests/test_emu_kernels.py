"""The library's kernels under host emulation of the CUDA execution model (tools/cuemu), CPU tier.

The build container has no GPU.  So that `pytest -m "not gpu"` still exercises the *kernel logic*
(recurrences of all 21 shell classes, the warp-cooperative and general-contraction kernels, Schwarz
task lists, work queues, J/K digestion with its warp reductions and atomics, scatter), the
unmodified sources of quiqbox.jl_b200/csrc are compiled with g++ against a coroutine emulation of
blocks / warps / shuffles / barriers, and the SAME parity checks the B200 box runs
(tests/test_gpu_parity.py) are repeated here on the small cases.  This is a functional check of the
CUDA sources, not a CPU fallback and not a performance path: the product binding
(quiqbox.jl_b200/lib.py) only ever loads libqbx.so and fails loudly without a GPU
(tests/test_abi_cpu.py); the GPU parity tests proper remain `-m gpu`.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import emu
import oracle
import quiqbox_b200 as qb
import test_gpu_parity as P
from molecules import h2, h2o, water_cluster

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module", autouse=True)
def emulated_library():
    emu.install()
    yield
    emu.uninstall()


def test_emulated_library_is_not_the_product_binding():
    from quiqbox_b200 import lib
    assert lib.LIB_PATH.endswith("libqbx.so") and "cuemu" not in lib.LIB_PATH
    pkg = os.path.join(os.path.dirname(HERE), "quiqbox.jl_b200")
    for root, _, files in os.walk(pkg):                       # nothing in the product knows about the emulation
        if os.path.basename(root) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read().lower()
                assert "cuemu" not in src and "emu.install" not in src and "libqbx_emu" not in src, f


def test_boys_kernels():
    P.test_boys_golden_points_generic_kernel()
    P.test_boys_table_vs_oracle()


def test_generic_kernels_golden_primitives_and_one_body():
    P.test_primitive_golden_eris_any_l()
    P.test_lih_tensor_symmetry_and_one_body()


@pytest.mark.parametrize("name,mol,basis", [("H2/STO-3G", h2(1.4), "STO-3G"), ("H2O/6-31G", h2o(), "6-31G"),
                                            ("H2O/cc-pVDZ", h2o(), "cc-pVDZ")])
def test_full_tensor_vs_oracle(name, mol, basis):
    P.test_full_tensor_vs_oracle(name, mol, basis)


@pytest.mark.parametrize("basis", ["6-31G", "cc-pVDZ"])
def test_fock_build_modes_vs_oracle_getGcore(basis):
    P.test_fock_build_modes_vs_oracle_getGcore(basis)


def test_sharded_partial_G_sums_to_full():
    P.test_sharded_partial_G_sums_to_full()


def test_scf_golden_energies():
    P.test_hoh_sto3g_scf()                      # HartreeFock-test.jl:92, 152
    P.test_h2o2_631g_scf()                      # HartreeFock-test.jl:296


def test_boundary_errors():
    P.test_boundary_errors()


@pytest.mark.parametrize("cls", P.CLASSES)
def test_synthetic_class_batch_vs_oracle(cls):
    P.test_synthetic_class_batch_vs_oracle(cls)


def test_irregular_basis_falls_back_to_generic_kernels():
    P.test_irregular_basis_falls_back_to_generic_kernels()


def _run_emulated(code, env=None, timeout=900):
    boot = "import sys\nsys.path[:0] = [%r, %r]\nimport emu\nemu.install()\n" % (os.path.dirname(HERE), HERE)
    r = subprocess.run([sys.executable, "-c", boot + code], env=dict(os.environ, **(env or {})), capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


_TENSOR_CODE = r'''
import numpy as np
import oracle, quiqbox_b200 as qb
from molecules import h2o
nuc, xyz = h2o()
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
T = qb.elecRepulsions(bs)
ref = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs)).eri_tensor()
err = float(np.max(np.abs(T - ref)))
print("ERR", err)
assert err < 1e-12
'''


@pytest.mark.parametrize("order", ["forward", "reverse"])
def test_cooperative_kernel_all_classes_vs_oracle(order):
    """QBX_COOP_MIN_ACC=0: every s/p/d class through the warp-cooperative kernel.  Run with the lanes
    of a warp (and the warps of a block) scheduled in ascending and in descending order: a missing
    __syncwarp / __syncthreads between a write and another lane's read shows up in one of the two."""
    _run_emulated(_TENSOR_CODE, {"QBX_COOP_MIN_ACC": "0", "QBX_EMU_LANE_ORDER": order})



def test_general_contraction_sharing_on_off():
    """(H2O)2/cc-pVDZ with Schwarz screening: the group kernels (QBX_GC=1, default) and the plain
    class kernels (QBX_GC=0) must give the same Fock matrix, and both must match the oracle tensor."""
    code = r'''
import numpy as np
import oracle, quiqbox_b200 as qb
from molecules import water_cluster
nuc, xyz = water_cluster(2)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
db = qb.DeviceBasis(bs)
n = db.nbf
rng = np.random.RandomState(3)
D = rng.uniform(-1, 1, (n, n)); D = (D + D.T) / 2
G = qb.DeviceERI(db, mode="stored", screen_tol=1e-12).getGcore(2 * D, [D])[0]
np.save(sys.argv[1] if len(sys.argv) > 1 else "/dev/null", G)
idx = rng.randint(0, n, size=(200, 4))
ref = oracle.OracleBasis(db.data).eri_list(idx, canonical=True)
assert np.max(np.abs(qb.elecRepulsionList(db, idx) - ref)) < 1e-10
print("GSUM %.15e" % float(np.sum(G * D)))
'''
    outs = [_run_emulated(code, {"QBX_GC": gc}) for gc in ("1", "0")]
    vals = [float(o.split("GSUM")[1]) for o in outs]
    assert abs(vals[0] - vals[1]) < 1e-8 * max(1.0, abs(vals[0]))


_FOCK_CODE = r'''
import numpy as np
import oracle, quiqbox_b200 as qb
from molecules import water_cluster
nuc, xyz = water_cluster(2)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "6-31G") for s, c in zip(nuc, xyz)), [])
db = qb.DeviceBasis(bs)
n = db.nbf
rng = np.random.RandomState(9)
DJ = rng.uniform(-1, 1, (n, n)); DJ = (DJ + DJ.T) / 2
DK = rng.uniform(-1, 1, (n, n)); DK = (DK + DK.T) / 2
Gref = oracle.getGcore(oracle.OracleBasis(db.data).eri_tensor(canonical=True), DJ, DK)
for mode in ("stored", "direct"):
    G = qb.DeviceERI(db, mode=mode, screen_tol=1e-13).getGcore(DJ, [DK])[0]
    err = float(np.max(np.abs(G - Gref)))
    print(mode, "ERR", err)
    assert err < 1e-10
print("NQ", db.info()["n_quartets"])
assert db.info()["n_quartets"] == 13652        # Schwarz bounds do not depend on which kernel computed the diagonals
'''


@pytest.mark.parametrize("env", [{}, {"QBX_EMU_LANE_ORDER": "reverse"}, {"QBX_DIGEST_SPREAD": "1"}, {"QBX_DIGEST_SPREAD": "3"},
                                 {"QBX_DIGEST_GROUP": "0"},
                                 {"QBX_GC": "0"}, {"QBX_EMU_SMS": "148"},
                                 {"QBX_GC": "0", "QBX_EMU_LANE_ORDER": "reverse", "QBX_EMU_SMS": "148"}],
                         ids=["default", "reverse-lane-order", "task-order-blocks", "3-block-slots-ragged-grid", "per-quartet-group-digest", "no-general-contraction",
                              "148-SMs-short-lists", "no-general-contraction-reverse-148"])
def test_fock_build_switches_vs_oracle(env):
    """(H2O)2/6-31G with Schwarz screening (ragged rows: most warps straddle several (bra pair, C)
    runs): the A/B switches must all give the oracle's G and the same surviving quartets.  With 148
    emulated SMs the lists are short against the device, which turns on the task-granular hand-out of
    the heaviest group tasks (one task per warp, lanes split the ket primitives)."""
    _run_emulated(_FOCK_CODE, env)


def test_deadlock_detector_reports_divergent_barrier(tmp_path):
    """The emulator must abort (not hang) on a barrier that not all live lanes reach."""
    src = tmp_path / "dl.cpp"
    src.write_text('#include <cuda_runtime.h>\n'
                   'static void k(int *p) { if (threadIdx.x < 16) __syncwarp(); else __syncthreads(); p[0] = 1; }\n'
                   'int main() { int v = 0; int *p = &v; cuemu::launch(dim3(1), dim3(32), 0, [&] { k(p); }); return 0; }\n')
    cu = os.path.join(os.path.dirname(HERE), "tools", "cuemu")
    exe = tmp_path / "dl"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(cu, "include"), str(src),
                           os.path.join(cu, "cuemu_rt.cpp"), "-o", str(exe), "-pthread"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "DEADLOCK" in r.stderr


@pytest.mark.parametrize("order", ["forward", "reverse"])
def test_emulator_warp_semantics(tmp_path, order):
    """The shim's __shfl*_sync / __ballot_sync / __all_sync / exited-lane / __syncthreads behaviour against what
    CUDA documents (tests/cuemu_unit/warp_semantics.cpp): a green emulated parity test is only worth as much
    as the emulation."""
    import importlib.util
    root = os.path.dirname(HERE)
    cu = os.path.join(root, "tools", "cuemu")
    spec = importlib.util.spec_from_file_location("qbx_build_emu_t", os.path.join(cu, "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    src = tmp_path / "ws.cpp"
    src.write_text(mod.transform(open(os.path.join(HERE, "cuemu_unit", "warp_semantics.cpp")).read()))
    exe = tmp_path / "ws"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-fpermissive", "-I", os.path.join(cu, "include"), str(src),
                           os.path.join(cu, "cuemu_rt.cpp"), "-o", str(exe), "-pthread"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, env=dict(os.environ, QBX_EMU_LANE_ORDER=order))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
