"""K4 work statistics at the C3 shape: placements kept per read, outer iterations and EM iterations per placement."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hmmufotu_b200 import engine
scale = sys.argv[1] if len(sys.argv) > 1 else "full"
cfg = bench.SCALES[scale]
n = 2048
db, sets = bench.build_inputs(cfg, n, read_seed=1)
gdb = engine.Database(db, device=0)
max_len = max(int(np.max(np.diff(x.offsets))) for x in sets) + 8
b = engine.Batch(gdb, n, max_len, paired=bool(cfg.get("paired")))
b.upload(*sets)
p = engine.SepParams()
b.alignSeq(); b.getSeed(p); b.estimateSeq(p); b.placeSeq(p)
nk, pl = b.places()
W = []
outer, em = [], []
for i in range(n):
    q = pl[i, :nk[i]]
    outer += list(q["outerIters"]); em += list(q["nPlaces"])
    if nk[i]: W.append(int(q["end"][0] - q["start"][0] + 1))
outer, em = np.array(outer), np.array(em)
print("scale %s: kept per read mean %.2f (min %d max %d); window mean %.0f; outer iters mean %.2f max %d; EM iters per placement mean %.1f max %d; EM per outer %.1f" %
      (scale, nk.mean(), nk.min(), nk.max(), np.mean(W), outer.mean(), outer.max(), em.mean(), em.max(), em.sum() / max(1, outer.sum())))
print("hist outer", np.bincount(outer)[:12].tolist())
