"""Extract Gaussian-format basis-set text (public Basis Set Exchange data) for H..Ne from
the reference's data file lib/BasisSets/BasisSets.jl into a JSON fixture that travels with
this repo (the reference tree does not exist on the GPU box).  Data only; run once here:
    python tools/extract_basis.py
"""
import json, re, sys
src = open("/root/reference/lib/BasisSets/BasisSets.jl").read()
names = re.search(r"AtomicGTOrbSetNames.*?\[(.*?)\]\)", src, re.S).group(1)
names = re.findall(r'"([^"]+)"', names)
body = src[src.index("const AtomicGTOrbSetTexts"):]
# each family starts at a line "#<name>"
out = {}
for k, name in enumerate(names):
    start = body.index("#" + name + "\n")
    end = body.index("#" + names[k + 1] + "\n") if k + 1 < len(names) else len(body)
    chunk = body[start:end]
    entries = re.findall(r'"""\n(.*?)"""|(\bnothing\b)', chunk, re.S)
    fam = {}
    for z, (txt, nothing) in enumerate(entries[:10], start=1):
        if txt:
            sym = txt.split()[0]
            fam[sym] = txt
    out[name] = fam
json.dump(out, open("quiqbox.jl_b200/data/basis_sets.json", "w"), indent=0)
print({k: list(v) for k, v in out.items()})
