"""The CPU oracle against the REFERENCE ITSELF: tests/golden/ref70 holds databases built by the reference's own
hmmufotu-build and assignment TSVs written by the reference's own hmmufotu (compiled in place from
/root/reference/src against stand-in Eigen/Boost headers: oracle/Makefile.ref, tests/golden/make_ref_golden.py).
Every row of those TSVs is recomputed here by the restatement in oracle/ from the same database files and reads."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import oracle_py as O  # noqa: E402
import refdb  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ref70")
HOST_TEST = os.path.join(ROOT, "hmmufotu_b200", "bin", "hu_host_test")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcom(s):
    return "".join(COMP.get(c, c) for c in reversed(s))


# <db>.msa / .csfm do not depend on the substitution model and <db>.hmm only in its DATE line: the GTR + dGamma4 database
# (BASELINE.json configs[1]) is committed as its .ptu.gz alone and paired with gg_70_otus_GTR's other files
SHARED = {"gg_70_otus_GTR_dG": "gg_70_otus_GTR"}


@pytest.fixture(scope="module", params=["gg_70_otus_GTR", "gg_70_otus_HKY85_dG", "gg_70_otus_GTR_dG"])
def refcase(request, tmp_path_factory):
    name = request.param
    d = tmp_path_factory.mktemp(name)
    src = SHARED.get(name, name)
    for ext in (".msa", ".hmm"):
        shutil.copy(os.path.join(GOLD, src + ext), d / (name + ext))
    db = refdb.load_ptu(os.path.join(GOLD, name + ".ptu.gz"))
    otree = O.OracleTree(db, evaluate=False)
    up, dn = np.ascontiguousarray(db.up), np.ascontiguousarray(db.dn)
    O.lib().orc_tree_set_messages(otree.h, up.ctypes.data_as(O.C.c_void_p), dn.ctypes.data_as(O.C.c_void_p))
    ohmm = O.OracleHmm(path=os.path.join(GOLD, src + ".hmm"))
    return dict(name=name, dir=str(d), db=db, otree=otree, ohmm=ohmm, hmm_path=os.path.join(GOLD, src + ".hmm"))


def seeds_of(case, fasta, seed, mate=None, single=False):
    """Host CSFM seed search (product host code, rand() in the reference's order) -> per sequence VPath arrays."""
    cmd = [HOST_TEST, "seeds", os.path.join(case["dir"], case["name"]), fasta, str(seed)] + ([mate] if mate else []) + (["--single"] if single else [])
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    return [np.array([[int(x) for x in p.split(",")] for p in ln.split("\t")[2:]], dtype=np.int32).reshape(-1, 6) for ln in out]


def eq_g(ref: str, val: float) -> bool:
    """the reference prints doubles with operator<< at precision 6 (%g); the sign of a NaN is not compared"""
    if val != val:
        return ref.lstrip("-") == "nan"
    return ref == "%g" % val


def check_rows(case, tsv, aligns):
    """aligns: {read id: oracle alignment dict}; every TSV column the reference writes is compared."""
    hdr, rows = refdb.load_tsv(tsv)
    col = {h: i for i, h in enumerate(hdr)}
    db, otree = case["db"], case["otree"]
    n_checked = n_tied = 0
    for rid, a in aligns.items():
        r = rows[rid]
        assert [int(r[col[c]]) for c in ("seq_start", "seq_end", "hmm_start", "hmm_end", "CS_start", "CS_end")] == [int(x) for x in a["coords"]], rid
        assert r[col["cost"]] == "%g" % a["cost"], rid
        assert r[col["alignment"]] == a["align"], rid
        seq = O.digitize(a["align"])
        s, e = int(a["coords"][4]) - 1, int(a["coords"][5]) - 1
        ids, dist = otree.get_seed(seq, s, e, 50, order=0)       # literal std::sort order
        want = otree.assign(seq, s, e, ids, dist)
        assert len(want) > 0, rid
        # Placements whose estimated log-likelihoods agree to rounding are ordered by rounding noise (libm exp/log
        # here and in the shim vs Eigen's packet functions in a stock build): compare as a set then.
        est = sorted(w.estLoglik for w in want)
        tied = any(abs(x - y) <= 1e-11 * abs(x) for x, y in zip(est, est[1:]))
        best = want[0]
        if tied:
            n_tied += 1
            assert r[col["branch_id"]] in {"%d->%d" % (w.c, w.p) for w in want}, rid
            continue
        n_checked += 1
        assert r[col["branch_id"]] == "%d->%d" % (best.c, best.p), rid
        assert eq_g(r[col["branch_ratio"]], best.ratio), (rid, r[col["branch_ratio"]], best.ratio)
        assert r[col["taxon_id"]] == str(best.a), rid
        assert r[col["taxon_anno"]] == db.anno[best.a], rid
        anno_dist = (best.wuv * best.ratio if best.a == best.c else (1 - best.ratio) * best.wuv) + best.wnr
        assert eq_g(r[col["anno_dist"]], anno_dist), (rid, r[col["anno_dist"]], anno_dist)
        assert eq_g(r[col["loglik"]], best.loglik), (rid, r[col["loglik"]], best.loglik)       # the literal constant W * log(sum pi e)
        assert eq_g(r[col["Q_placement"]], best.qPlace), (rid, r[col["Q_placement"]], best.qPlace)
        assert eq_g(r[col["Q_taxon"]], best.qTaxon), (rid, r[col["Q_taxon"]], best.qTaxon)
    return n_checked, n_tied


def test_single_end_rows_match_the_reference(refcase):
    name = refcase["name"]
    total = 0
    for k in range(4):
        fa = os.path.join(GOLD, "%s_se%d.fasta" % (name, k))
        if not os.path.exists(fa):
            continue
        reads = refdb.read_fasta(fa)
        vps = seeds_of(refcase, fa, k)
        assert len(vps) == len(reads)
        aligns = {rid: refcase["ohmm"].align(seq.upper(), vps[i]) for i, (rid, _, seq) in enumerate(reads)}
        n, t = check_rows(refcase, os.path.join(GOLD, "%s_se%d.tsv" % (name, k)), aligns)
        assert n + t == len(reads) and n >= 0.9 * len(reads)
        total += n
        print("%s se%d: %d rows identical, %d with rounding-tied estimates" % (name, k, n, t))
    assert total >= 57
    if name == "gg_70_otus_GTR_dG":
        assert len(refcase["db"].dg_rates) == 4 and refcase["db"].model["type"] == refdb.MODEL_TYPES["GTR"]


def test_paired_end_rows_match_the_reference(refcase):
    name = refcase["name"]
    f1, f2 = (os.path.join(GOLD, "%s_pe0_%d.fasta" % (name, m)) for m in (1, 2))
    fwd, rev = refdb.read_fasta(f1), refdb.read_fasta(f2)
    vps = seeds_of(refcase, f1, 7, f2)
    assert len(vps) == 2 * len(fwd)
    aligns = {}
    L = refcase["ohmm"].L
    for i, ((rid, _, s1), (_, _, s2)) in enumerate(zip(fwd, rev)):
        a = refcase["ohmm"].align(s1.upper(), vps[2 * i])
        b = refcase["ohmm"].align(revcom(s2.upper()), vps[2 * i + 1])
        aligns[rid] = O.merge(L, a, b)
    n, t = check_rows(refcase, os.path.join(GOLD, name + "_pe0.tsv"), aligns)
    assert n + t == len(fwd) and n >= 0.9 * len(fwd)


def test_single_mode_matches_the_reference(refcase):
    """--single: NGCL alignment mode (src/hmmufotu.cpp:358, src/BandedHMMP7.cpp:561-583)"""
    name = refcase["name"]
    tsv = os.path.join(GOLD, "%s_se0.single.tsv" % name)
    if not os.path.exists(tsv):
        pytest.fail("missing golden " + tsv)
    fa = os.path.join(GOLD, name + "_se0.fasta")
    reads = refdb.read_fasta(fa)
    vps = seeds_of(refcase, fa, 0, single=True)
    ohmm = O.OracleHmm(path=refcase["hmm_path"], mode=O.NGCL)
    aligns = {rid: ohmm.align(seq.upper(), vps[i]) for i, (rid, _, seq) in enumerate(reads)}
    n, t = check_rows(refcase, tsv, aligns)
    assert n + t == len(reads) and n >= 0.9 * len(reads)


REF_BIN = os.path.join(ROOT, "oracle", "_ref")
REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF_BIN, "hmmufotu-build")) and os.path.isdir(REF_TEST)),
                    reason="needs the reference tree and oracle/_ref (this container only)")
@pytest.mark.parametrize("model,extra", [("TN93", []), ("F81", ["-V", "-k", "3"]), ("K80", []), ("JC69", ["-V", "-k", "5"])])
def test_other_models_built_and_run_by_the_reference(model, extra, tmp_path):
    """The four substitution models without a committed golden: database built, reads drawn and assigned by the
    reference's own programs on the spot, every TSV column recomputed by the oracle."""
    db_name = "gg_70_otus_" + model
    run = lambda *c: subprocess.run(c, cwd=str(tmp_path), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    run(os.path.join(REF_BIN, "hmmufotu-build"), os.path.join(REF_TEST, "70_otus.fasta"), os.path.join(REF_TEST, "70_otus.tree"),
        "-a", os.path.join(REF_TEST, "70_otus_taxonomy.txt"), "-n", db_name, "-s", model, *extra)
    run(os.path.join(REF_BIN, "hmmufotu-sim"), db_name, "r.fasta", "-N", "32", "-S", "5", "-r", "250")
    run(os.path.join(REF_BIN, "hmmufotu"), db_name, "r.fasta", "-o", "r.tsv", "-s", "1", "-S", "5", "-p", "1")
    db = refdb.load_ptu(str(tmp_path / (db_name + ".ptu")))
    assert db.model["type"] == refdb.MODEL_TYPES[model] and (len(db.dg_rates) > 0) == bool(extra)
    otree = O.OracleTree(db, evaluate=False)
    up, dn = np.ascontiguousarray(db.up), np.ascontiguousarray(db.dn)
    O.lib().orc_tree_set_messages(otree.h, up.ctypes.data_as(O.C.c_void_p), dn.ctypes.data_as(O.C.c_void_p))
    case = dict(name=db_name, dir=str(tmp_path), db=db, otree=otree, ohmm=O.OracleHmm(path=str(tmp_path / (db_name + ".hmm"))))
    fa = str(tmp_path / "r.fasta")
    reads = refdb.read_fasta(fa)
    vps = seeds_of(case, fa, 5)
    aligns = {rid: case["ohmm"].align(seq.upper(), vps[i]) for i, (rid, _, seq) in enumerate(reads)}
    n, t = check_rows(case, str(tmp_path / "r.tsv"), aligns)
    assert n + t == len(reads) and n >= 0.9 * len(reads)


@pytest.mark.parametrize("variant", ["weighted", "ml", "height"])
def test_run_options_match_the_reference(refcase, variant):
    """-m weighted, --ML and --prior height (src/hmmufotu.cpp:229-241) against the reference's TSVs of the same reads"""
    name = refcase["name"]
    tsv = os.path.join(GOLD, "%s_se0.%s.tsv" % (name, variant))
    if not os.path.exists(tsv):
        pytest.fail("missing golden " + tsv)
    fa = os.path.join(GOLD, name + "_se0.fasta")
    reads = refdb.read_fasta(fa)
    vps = seeds_of(refcase, fa, 0)
    aligns = {rid: refcase["ohmm"].align(seq.upper(), vps[i]) for i, (rid, _, seq) in enumerate(reads)}
    refcase["otree"].set_options(weighted=variant == "weighted", prior_height=variant == "height", only_ml=variant == "ml")
    try:
        n, t = check_rows(refcase, tsv, aligns)
    finally:
        refcase["otree"].set_options()
    assert n + t == len(reads) and n >= 0.9 * len(reads)
