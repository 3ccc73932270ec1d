"""CPU: the oracle restatement against independent properties (no GPU)."""
import math

import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import synth


@pytest.fixture(scope="module")
def small():
    db = synth.make_db(n_leaves=30, cs_len=400, occupancy=0.85, seed=3)
    return dict(db=db, hmm=O.OracleHmm(db), reads=synth.make_reads(db, 40, read_len=150, seed=5, drop_seed_prob=0.1))


@pytest.mark.parametrize("mtype", range(6))
def test_transition_matrices_vs_expm(mtype):
    """P(t) of every model family against scipy.linalg.expm of the model's rate matrix."""
    from scipy.linalg import expm
    m = synth.gg97_model(mtype)
    om = O.OracleModel(m)
    pi = om.pi()
    # recover the generator from the oracle itself: Q = dP/dt at 0 (central difference), then check the semigroup
    h = 1e-6
    Q = (om.Pr(h) - om.Pr(0.0 if mtype != synth.GTR else 1e-300)) / h
    assert np.allclose(Q.sum(axis=1), 0, atol=1e-5)
    for t in (1e-3, 0.05, 0.7, 3.0):
        P = om.Pr(t)
        assert np.allclose(P.sum(axis=1), 1.0, atol=1e-12)
        assert (P >= 0).all()
        assert np.allclose(pi[:, None] * P, (pi[:, None] * P).T, atol=1e-12)          # detailed balance
        assert np.allclose(P @ P, om.Pr(2 * t), atol=1e-11)                              # Chapman-Kolmogorov
        assert np.allclose(P, expm(Q * t), atol=2e-5)                                   # generator (finite-difference Q)
    assert np.allclose(om.Pr(0.0), np.eye(4), atol=1e-15)
    assert np.allclose(pi @ om.Pr(1.3), pi, atol=1e-12)                                  # stationarity


def test_gtr_scaling_is_trace():
    """GTR Q is scaled so that -trace(Q) = 1 (src/DNASubModel.cpp:123-126), not by the mean rate."""
    om = O.OracleModel(synth.gg97_model(synth.GTR))
    h = 1e-7
    Q = (om.Pr(h) - np.eye(4)) / h
    assert abs(np.trace(Q) + 1.0) < 1e-5


def test_model_text_roundtrip():
    for t in range(6):
        m = synth.gg97_model(t)
        a, b = O.OracleModel(m), O.OracleModel(text=synth.MODEL_NAMES[t] + "\n" + synth.model_text(m))
        assert np.allclose(a.Pr(0.37), b.Pr(0.37), atol=1e-15)


def test_hmm_text_roundtrip(small, tmp_path):
    path = str(tmp_path / "synth.hmm")
    synth.write_hmm(small["db"], path)
    a, b = small["hmm"].params(), O.OracleHmm(path=path).params()
    for k in ("eM", "eI", "T", "entry", "exit"):
        x, y = a[k], b[k]
        fin = np.isfinite(x)
        assert np.array_equal(fin, np.isfinite(y)), k
        assert np.array_equal(x[fin][1 if k == "eM" else 0:], y[fin][1 if k == "eM" else 0:]), k  # COMPO row is free-form
    assert np.array_equal(a["p2cs"], b["p2cs"]) and np.array_equal(a["cs2p"], b["cs2p"])


def test_wing_retraction_bounds(small):
    p = small["hmm"].params()
    assert np.isinf(p["entry"][0]) and np.isinf(p["exit"][0])
    assert (p["entry"][1:] >= 0).all() and (p["exit"][1:] >= 0).all()
    # retraction only adds probability: costs never above the uniform local-mode cost
    assert (p["entry"][2:] <= p["entry"][1] + 1e-12).all()


def test_banded_equals_full_when_band_covers_optimum(small):
    rs, h = small["reads"], small["hmm"]
    same = 0
    for i in range(rs.n):
        if rs.n_paths[i] == 0:
            continue
        a = h.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        f = h.align(rs.read_str(i), np.zeros((0, 6), dtype=np.int32))
        assert a["valid"] and f["valid"]
        assert f["cost"] <= a["cost"] + 1e-9          # the full matrix can only do better
        same += a["cost"] == f["cost"] and a["trace"] == f["trace"]
        assert a["cells"] < f["cells"] or a["fallback"]
    assert same >= 0.8 * (rs.n_paths > 0).sum()        # seeds are exact matches: the band nearly always holds the optimum


def test_alignment_string_shape(small):
    rs, h, db = small["reads"], small["hmm"], small["db"]
    for i in range(rs.n):
        a = h.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        al = a["align"]
        cs, ce = a["coords"][4], a["coords"][5]
        assert len(al) == db.cs_len
        assert set(al[:cs - 1]) <= {"."} and set(al[ce:]) <= {"."}
        assert a["trace"][0] == "B" and a["trace"][-1] == "E"
        n_m, n_i = a["trace"].count("M"), a["trace"].count("I")
        assert n_m + n_i == a["coords"][1] - a["coords"][0] + 1
        # every match base is written; insert bases are written only where a non-profile column exists
        written = sum(c not in ".-" for c in al)
        assert n_m <= written <= n_m + 2 * n_i
        assert sum(c.isupper() for c in al) == n_m


def test_global_mode_aligns_whole_read(small):
    rs, h = small["reads"], small["hmm"]
    for i in range(10):
        a = h.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        assert a["coords"][0] == 1 and a["coords"][1] == len(rs.read_str(i))


def test_merge_and_digitize():
    L = 12
    a = dict(coords=np.array([1, 4, 2, 5, 3, 6], dtype=np.int32), cost=1.5, align="..ACgT......")
    b = dict(coords=np.array([1, 3, 7, 9, 8, 10], dtype=np.int32), cost=2.0, align=".......T-A..")
    m = O.merge(L, a, b)
    assert m["align"] == "..ACgT.T-A.." and m["cost"] == 3.5
    assert list(m["coords"]) == [1, 4, 2, 9, 3, 10]
    assert list(O.digitize("AcgtN-.")) == [0, 1, 2, 3, 0, -2, -2]


def test_felsenstein_messages_vs_brute_force():
    """All-roots evaluation on a 4-leaf tree against explicit enumeration of the internal states."""
    db = synth.make_db(n_leaves=4, cs_len=6, occupancy=1.0, seed=1, leaf_del=0.0, leaf_ins=0.0)
    t = O.OracleTree(db, infer_internal=False)
    up, dn = t.messages()
    om = O.OracleModel(db.model)
    pi = om.pi()
    children = db.children
    P = {x: om.Pr(db.blen[x]) for x in range(1, db.n_nodes)}

    def cond(x, j):  # conditional likelihood vector of the subtree below x, by recursion on probabilities
        if not children[x]:
            v = np.zeros(4); v[db.seq[x, j]] = 1.0
            return v
        v = np.ones(4)
        for c in children[x]:
            v *= P[c] @ cond(c, j)
        return v
    for x in range(1, db.n_nodes):
        for j in range(db.cs_len):
            want = cond(x, j)
            got = np.exp(up[x, j])
            assert np.allclose(got, want, rtol=1e-12, atol=1e-300)
    # the two messages of an edge combine to the same tree likelihood on every edge
    for j in range(db.cs_len):
        liks = []
        for x in range(1, db.n_nodes):
            a = np.exp(up[x, j]); b = np.exp(dn[x, j])
            liks.append(float(pi @ ((P[x] @ a) * b)))
        assert np.allclose(liks, liks[0], rtol=1e-11)


def test_sep_invariants(small):
    db = small["db"]
    t = O.OracleTree(db)
    seqs = t.seqs()
    assert (seqs[~db.is_leaf] >= 0).all()                 # inferred internal sequences never hold a gap
    h, rs = small["hmm"], small["reads"]
    a = h.align(rs.read_str(0), rs.vpaths[0, :rs.n_paths[0]])
    seq = O.digitize(a["align"])
    s, e = a["coords"][4] - 1, a["coords"][5] - 1
    ids, dist = t.get_seed(seq, s, e, 50, order=1)
    assert (np.diff(dist) >= 0).all() and db.root not in ids
    d, N = t.pdist(int(ids[0]), seq, s, e)
    assert dist[0] == d / N
    lit_ids, lit_dist = t.get_seed(seq, s, e, 50, order=0)  # std::sort: same multiset of distances
    assert np.array_equal(np.sort(lit_dist), np.sort(dist))
    rows = t.assign(seq, s, e, ids, dist)
    W = e - s + 1
    for r in rows:
        assert r.loglik == rows[0].loglik and abs(r.loglik - W) < 1e-9   # literal placeSeq return: W * (1 + log(sum pi))
        assert 0 <= r.ratio <= 1 and 0 <= r.wnr <= 1
        assert abs(r.loglikTrue - r.estLoglik) < 30                      # optimised and estimated placements agree roughly
        assert r.qPlace == rows[0].qPlace
    n = len(rows)
    want_q = 250.0 if n == 1 else -10 * math.log10(1 - 1 / n)
    assert abs(rows[0].qPlace - want_q) < 1e-9


def test_equal_key_sort_permutation():
    """The element std::sort(rbegin, rend) leaves first on n equal keys (identity up to 16)."""
    for n in range(1, 17):
        assert O.equal_sort_perm(n)[0] == 0
    assert [int(O.equal_sort_perm(n)[0]) for n in (17, 20, 33, 50)] == [15, 18, 31, 25]
