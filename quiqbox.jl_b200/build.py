"""Builds quiqbox.jl_b200/libqbx.so (sm_100a only) from csrc/ with nvcc, in-tree.

21 per-class translation units (class_inst.cu compiled with -DQLA.. macros) + 4 host/generic
units, compiled in parallel, then linked into one shared library whose dynamic symbol table is
exactly include/qbx.h (-fvisibility=hidden + QBX_API; tests/test_abi_cpu.py compares the two).  Incremental: a unit is rebuilt when its sources are newer than its
object.  `python quiqbox.jl_b200/build.py [-j N] [--force]`.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqbx.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-ccbin", "/usr/bin/g++", "-I", CSRC]
# tuning constants for A/B builds on the GPU box, e.g. QBX_NVCC_DEFS="-DQBX_ERI_THREADS=128 -DCOOP_WARPS=8"
# (then --force; the defaults are the product)
FLAGS += os.environ.get("QBX_NVCC_DEFS", "").split()

CLASSES = [(a, b, c, d) for a in range(3) for b in range(a + 1) for c in range(3) for d in range(c + 1)
           if (a * (a + 1) // 2 + b) >= (c * (c + 1) // 2 + d)]
HEADERS = ["boys.cuh", "qbx_internal.h", "eri_class.cuh", "digest.cuh", "engine.h", "handle.h", "linalg.h", "../../include/qbx.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(job):
    src, obj, defs = job
    cmd = [NVCC] + FLAGS + defs + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def build(jobs=None, force=False, verbose=True):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    # a library that is newer than every source is up to date even when the object files are not there
    # (they do not travel to the GPU box: .gpurunignore)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + hdrs
    if not force and os.path.exists(LIB) and not _newer(LIB, srcs):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    work, objs = [], []
    for name in ("api", "generic", "engine", "eri_coop", "eri_group", "pool", "comm", "linalg", "scf"):
        src, obj = os.path.join(CSRC, name + ".cu"), os.path.join(OBJ, name + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            work.append((src, obj, []))
    src = os.path.join(CSRC, "class_inst.cu")
    # biggest classes first so the pool drains evenly
    for (a, b, c, d) in sorted(CLASSES, key=lambda t: -sum(t)):
        obj = os.path.join(OBJ, f"class_{a}{b}{c}{d}.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            work.append((src, obj, [f"-DQLA={a}", f"-DQLB={b}", f"-DQLC={c}", f"-DQLD={d}"]))
    if work:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for obj, rc, out in ex.map(_compile, work):
                if verbose:
                    print(("ok   " if rc == 0 else "FAIL ") + os.path.basename(obj), flush=True)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed for {obj}:\n{out}")
    if work or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, "-ccbin", "/usr/bin/g++"] + objs + ["-ldl", "-Xlinker", "--version-script=" + os.path.join(CSRC, "libqbx.map")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print("linked", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    j = None
    if "-j" in sys.argv:
        j = int(sys.argv[sys.argv.index("-j") + 1])
    build(j, "--force" in sys.argv)
