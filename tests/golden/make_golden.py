"""Mint the golden vectors of tests/golden/ from the CPU oracle (oracle/liboracle.so).

The reference's own tests hold no numeric goldens for this path and the reference cannot be compiled in this
image, so these vectors pin the ORACLE (regression), not the reference: "parity unpinned".

    python tests/golden/make_golden.py        # rewrites tests/golden/sep_small.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle_py as O  # noqa: E402
from hmmufotu_b200 import synth  # noqa: E402

DB_ARGS = dict(n_leaves=24, cs_len=320, occupancy=0.85, seed=42)
READ_ARGS = dict(n=12, read_len=120, seed=9, drop_seed_prob=0.1)


def make(dg: bool):
    args = dict(DB_ARGS)
    if dg:
        args.update(dg_alpha=0.6, dg_k=4)
    db = synth.make_db(**args)
    otree = O.OracleTree(db)
    db.seq = otree.seqs()
    ohmm = O.OracleHmm(db)
    rs = synth.make_reads(db, **READ_ARGS)
    out = []
    for i in range(rs.n):
        a = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        seq = O.digitize(a["align"])
        s, e = int(a["coords"][4]) - 1, int(a["coords"][5]) - 1
        ids, dist = otree.get_seed(seq, s, e, 50, order=1)
        rows = otree.assign(seq, s, e, ids, dist)
        out.append(dict(read=rs.read_str(i), vpaths=rs.vpaths[i, :rs.n_paths[i]].tolist(), coords=a["coords"].tolist(),
                        cost=a["cost"].hex() if hasattr(a["cost"], "hex") else float(a["cost"]).hex(), trace=a["trace"],
                        fallback=a["fallback"], seeds=ids.tolist(), dist=[float(d).hex() for d in dist],
                        places=[dict(c=r.c, p=r.p, a=r.a, ratio=r.ratio, wnr=r.wnr, loglik=r.loglik, loglikTrue=r.loglikTrue,
                                     estLoglik=r.estLoglik, qPlace=r.qPlace, qTaxon=r.qTaxon, outerIters=r.outerIters) for r in rows]))
    return dict(db_args=args, read_args=READ_ARGS, reads=out)


if __name__ == "__main__":
    gold = dict(note="oracle-minted (parity unpinned vs the reference); costs and distances are exact hex doubles",
                plain=make(False), dgamma=make(True))
    with open(os.path.join(HERE, "sep_small.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "sep_small.json"))
