"""CPU: the oracle against the committed golden vectors (tests/golden/sep_small.json, minted by
tests/golden/make_golden.py from the oracle itself: a regression pin, the reference has no goldens)."""
import json
import os

import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sep_small.json")))


@pytest.mark.parametrize("name", ["plain", "dgamma"])
def test_oracle_reproduces_golden(name):
    g = GOLD[name]
    db = synth.make_db(**g["db_args"])
    otree = O.OracleTree(db)
    ohmm = O.OracleHmm(db)
    for i, r in enumerate(g["reads"]):
        vp = np.array(r["vpaths"], dtype=np.int32).reshape(-1, 6)
        a = ohmm.align(r["read"], vp)
        assert list(a["coords"]) == r["coords"] and a["trace"] == r["trace"] and a["fallback"] == r["fallback"], i
        assert float(a["cost"]).hex() == r["cost"], i
        seq = O.digitize(a["align"])
        s, e = r["coords"][4] - 1, r["coords"][5] - 1
        ids, dist = otree.get_seed(seq, s, e, 50, order=1)
        assert ids.tolist() == r["seeds"] and [float(d).hex() for d in dist] == r["dist"], i
        rows = otree.assign(seq, s, e, ids, dist)
        assert len(rows) == len(r["places"])
        for o, w in zip(rows, r["places"]):
            assert (o.c, o.p, o.a, o.outerIters) == (w["c"], w["p"], w["a"], w["outerIters"])
            np.testing.assert_allclose([o.ratio, o.wnr, o.loglik, o.loglikTrue, o.estLoglik, o.qPlace],
                                       [w["ratio"], w["wnr"], w["loglik"], w["loglikTrue"], w["estLoglik"], w["qPlace"]], rtol=1e-12)
