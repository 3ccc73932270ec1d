"""Diagnostic: how many bases of an aligned read sit in non-match columns (the seed scan's sparse list), per read and per
256-read tile, for a bench shape.  python scripts/gpu_inserts.py [scale]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hmmufotu_b200 import engine

scale = sys.argv[1] if len(sys.argv) > 1 else "full"
cfg = bench.SCALES[scale]
n = 2048
db, sets = bench.build_inputs(cfg, n, read_seed=1)
gdb = engine.Database(db, device=0)
paired = bool(cfg.get("paired"))
max_len = max(int(np.max(np.diff(x.offsets))) for x in sets) + 8
b = engine.Batch(gdb, n, max_len, paired=paired)
b.upload(*sets)
b.alignSeq()
codes = b.aligned_codes()
rows = b.merged()
match = np.zeros(db.cs_len, dtype=bool)
match[np.asarray(db.p2cs[1:db.K + 1]) - 1] = True
ns, cols = [], []
for i in range(n):
    s, e = rows["csStart"][i] - 1, rows["csEnd"][i] - 1
    c = np.nonzero((codes[i, s:e + 1] >= 0) & ~match[s:e + 1])[0] + s
    ns.append(len(c)); cols.append(c)
ns = np.array(ns)
print("scale", scale, "reads", n, "with inserts", int((ns > 0).sum()), "mean ns", ns.mean(), "max", ns.max(), "hist", np.bincount(ns)[:12].tolist())
for t in range(0, n, 256):
    allc = np.concatenate(cols[t:t + 256])
    print("tile", t // 256, "entries", len(allc), "distinct columns", len(np.unique(allc)))
seq = gdb.seqs()
leaf = np.ones(db.n_nodes, dtype=bool); leaf[db.parent[db.parent >= 0]] = False
allc = np.unique(np.concatenate(cols))
occ_leaf = (seq[leaf][:, allc] >= 0).mean(); occ_int = (seq[~leaf][:, allc] >= 0).mean()
print("node occupancy at those columns: leaves %.4f internal %.4f" % (occ_leaf, occ_int))
vals, cnt = np.unique(seq[~leaf][:, allc], return_counts=True)
print("internal node codes there:", dict(zip(vals.tolist(), cnt.tolist())))
