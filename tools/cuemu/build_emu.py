"""Builds tools/cuemu/_build/libqbx_emu.so: the library's own sources (quiqbox.jl_b200/csrc, unmodified)
compiled with g++ against the cuemu host emulation of the CUDA execution model.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see include/cuda_runtime.h).  The only source rewriting is
syntactic, done on a scratch copy of csrc/:
    K<<<grid, block, smem, stream>>>(args);   ->  cuemu::launch(dim3(grid), dim3(block), smem, [&] { K(args); });
    extern __shared__ T name[];               ->  T *name = (T *)cuemu::dyn_smem();
    __shared__ T x[...];                      ->  static T x[...];      (blocks run one after the other)
`python tools/cuemu/build_emu.py [-j N] [--force]`.
"""
import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "quiqbox.jl_b200", "csrc")
SAN = os.environ.get("QBX_EMU_SANITIZE", "")           # e.g. "address,undefined": a separate build directory
OUT = os.path.join(HERE, "_build" + ("_san" if SAN else ""))
SRC = os.path.join(OUT, "src")
LIB = os.path.join(OUT, "libqbx_emu.so")
CXX = os.environ.get("CXX", "g++")
FLAGS = ["-std=c++17", "-O1", "-g0", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes", "-fpermissive", "-w",
         "-march=native", "-I", os.path.join(HERE, "include"), "-I", SRC, "-I", CSRC]

CLASSES = [(a, b, c, d) for a in range(3) for b in range(a + 1) for c in range(3) for d in range(c + 1)
           if (a * (a + 1) // 2 + b) >= (c * (c + 1) // 2 + d)]
if SAN:
    FLAGS = [f for f in FLAGS if f not in ("-O1", "-g0")] + ["-O1", "-g", "-fno-omit-frame-pointer", f"-fsanitize={SAN}"]
UNITS = ["api", "generic", "engine", "eri_coop", "eri_group", "pool", "comm", "linalg", "scf"]

_launch = re.compile(r"([A-Za-z_]\w*(?:\s*<[^<>;(){}]*>)?)\s*<<<")


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes text[i] == open_ch"""
    depth = 0
    while i < len(text):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced launch expression")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(text):
    out, pos = "", 0
    while True:
        m = _launch.search(text, pos)
        if not m:
            out += text[pos:]
            break
        out += text[pos:m.start()]
        kern = m.group(1)
        cfg_end = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():cfg_end])
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        j = cfg_end + 3
        while text[j].isspace():
            j += 1
        assert text[j] == "(", "launch without argument list"
        k = _match(text, j, "(", ")")
        args = text[j + 1:k - 1]
        out += f"cuemu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&] {{ {kern}({args}); }})"
        pos = k
    text = out
    text = re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;", r"\1 *\2 = (\1 *)cuemu::dyn_smem();", text)
    text = re.sub(r"(?<![\w])__shared__\s+", "static ", text)
    return text


def _stage():
    os.makedirs(SRC, exist_ok=True)
    changed = False
    for name in sorted(os.listdir(CSRC)):
        if not name.endswith((".cu", ".cuh", ".h")):
            continue
        with open(os.path.join(CSRC, name)) as f:
            t = transform(f.read())
        dst = os.path.join(SRC, name[:-3] + ".cpp" if name.endswith(".cu") else name)
        old = open(dst).read() if os.path.exists(dst) else None
        if old != t:
            with open(dst, "w") as f:
                f.write(t)
            changed = True
    return changed


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(job):
    src, obj, defs = job
    r = subprocess.run([CXX] + FLAGS + defs + ["-c", src, "-o", obj], capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def build(jobs=None, force=False, verbose=False):
    _stage()
    hdrs = [os.path.join(SRC, h) for h in os.listdir(SRC) if h.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "thrust", "execution_policy.h"),
             os.path.join(ROOT, "include", "qbx.h")]
    work, objs = [], []
    rt_src, rt_obj = os.path.join(HERE, "cuemu_rt.cpp"), os.path.join(OUT, "cuemu_rt.o")
    objs.append(rt_obj)
    if force or _newer(rt_obj, [rt_src] + hdrs):
        work.append((rt_src, rt_obj, []))
    for name in UNITS:
        src, obj = os.path.join(SRC, name + ".cpp"), os.path.join(OUT, name + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            work.append((src, obj, []))
    src = os.path.join(SRC, "class_inst.cpp")
    for (a, b, c, d) in sorted(CLASSES, key=lambda t: -sum(t)):
        obj = os.path.join(OUT, f"class_{a}{b}{c}{d}.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            work.append((src, obj, [f"-DQLA={a}", f"-DQLB={b}", f"-DQLC={c}", f"-DQLD={d}"]))
    if work:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for obj, rc, out in ex.map(_compile, work):
                if verbose:
                    print(("ok   " if rc == 0 else "FAIL ") + os.path.basename(obj), flush=True)
                if rc != 0:
                    raise RuntimeError(f"{CXX} failed for {obj}:\n{out[-6000:]}")
    if work or not os.path.exists(LIB):
        r = subprocess.run([CXX, "-shared", "-pthread", "-o", LIB] + ([f"-fsanitize={SAN}"] if SAN else []) + objs + ["-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print("linked", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    j = None
    if "-j" in sys.argv:
        j = int(sys.argv[sys.argv.index("-j") + 1])
    build(j, "--force" in sys.argv, verbose=True)
