import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import oracle_py as O
from hmmufotu_b200 import synth, engine
db = synth.make_db(n_leaves=60, cs_len=700, occupancy=0.85, seed=11)
otree = O.OracleTree(db); up, dn = otree.messages(); db.seq = otree.seqs()
ohmm = O.OracleHmm(db)
gdb = engine.Database(db, messages=(up, dn))
rs = synth.make_reads(db, 300, read_len=200, seed=5, drop_seed_prob=0.05)
n = 64
L = db.cs_len
codes = np.full((n, L), -2, dtype=np.int8); start = np.zeros(n, dtype=np.int32); end = np.zeros(n, dtype=np.int32)
for i in range(n):
    o = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
    codes[i] = O.digitize(o["align"]); start[i], end[i] = o["coords"][4]-1, o["coords"][5]-1
b = engine.Batch(gdb, n, 260); b.set_aligned(codes, start, end)
p = engine.SepParams(); b.sep(p)
ns, ids, dist = b.seeds(with_dist=True)
nk, est = b.estimates()
nbad = 0
for i in range(n):
    for s in range(ns[i]):
        o = otree.estimate(codes[i], start[i], end[i], int(ids[i, s]), float(dist[i, s]))
        # find gpu est row for this seed (filter order) -> search
        g = [e for e in est[i, :nk[i]] if e["cNode"] == ids[i, s]]
        if not g: continue
        g = g[0]
        rel = abs(g["estLoglik"] - o.loglik) / abs(o.loglik)
        if rel > 1e-12 or g["wnr"] != o.wnr or abs(g["ratio"] - o.ratio) > 1e-15:
            nbad += 1
            if nbad < 15: print(i, s, ids[i, s], "gpu", g["estLoglik"], g["wnr"], g["ratio"], "orc", o.loglik, o.wnr, o.ratio, "rel", rel, "W", end[i]-start[i]+1)
print("nbad", nbad)
