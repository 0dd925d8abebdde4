import sys, numpy as np
sys.path[:0] = ['.', 'tests']
import oracle, quiqbox_b200 as qb
from molecules import water_cluster
nuc, xyz = water_cluster(3)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
n = len(bs)
ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
T = ob.eri_tensor()
Td = qb.elecRepulsions(bs)
d = np.abs(T - Td)
bad = np.argwhere(d > 1e-9)
print("n bad", len(bad), "of", d.size)
lab = ["%d:l%d:K%d:%s" % (i, sum(g.ang), len(g.xpns), "".join(map(str, g.ang))) for i, g in enumerate(bs)]
import collections
cnt = collections.Counter()
for b in bad:
    cnt[tuple(sorted((sum(bs[i].ang), len(bs[i].xpns)) for i in b))] += 1
for k, v in cnt.most_common(12): print(k, v)
i = tuple(bad[np.argmax(d[tuple(bad.T)])])
print("worst", i, [lab[k] for k in i], "oracle", T[i], "class", Td[i], "generic", qb.elecRepulsionList(bs, [list(i)])[0])
# atoms of the worst
at = lambda f: [j for j, (s, c) in enumerate(zip(nuc, xyz)) if tuple(c) == bs[f].center][0]
print("atoms", [at(k) for k in i])
