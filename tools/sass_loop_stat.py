"""Static inner-loop instruction mix of eri_group_kernel<0/1> from two cubins: python tools/sass_loop_stat.py OLD.cubin NEW.cubin
(nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I quiqbox.jl_b200/csrc -cubin -o X.cubin quiqbox.jl_b200/csrc/eri_group.cu)."""
import re, subprocess, sys
def kernel_sass(cubin, pat):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks:
        if re.match(pat, b): return b
    raise SystemExit("kernel not found")
def instrs(sass):
    res = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m: res.append((int(m.group(1), 16), m.group(2)))
    return res
def innermost_loop_with(ins, needle):
    # backward branches define loops; pick the smallest loop containing an instruction with `needle`
    best = None
    for addr, txt in ins:
        m = re.search(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [t for a, t in ins if tgt <= a <= addr]
                if any(needle in t for t in body) and (best is None or len(body) < len(best)): best = body
    return best
for name, cubin in (("round-1 (9-column table)", sys.argv[1] if len(sys.argv) > 2 else "/tmp/old_group.cubin"), ("now (2 columns + recursion)", sys.argv[2] if len(sys.argv) > 2 else "/tmp/new_group.cubin")):
    for la in (0, 1):
        s = kernel_sass(cubin, r".*eri_group_kernelILi%dE" % la)
        body = innermost_loop_with(instrs(s), "MUFU.RSQ64H")
        ops = {}
        for t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        fp64 = sum(ops.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
        print(f"{name:30s} eri_group_kernel<{la}> bra-primitive loop: {len(body):4d} instr, FP64 {fp64:3d} (DFMA {ops.get('DFMA',0)}, DMUL {ops.get('DMUL',0)}, DADD {ops.get('DADD',0)}), LDS {ops.get('LDS',0)}, LDG {ops.get('LDG',0)}, LDL+STL {ops.get('LDL',0)+ops.get('STL',0)}")
