"""The C++ host layer (hmmufotu_b200/host): formats, CSFM seed index, command line.  CPU tests, plus one
end-to-end run of the hmmufotu executable on a GPU."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "hmmufotu_b200", "bin")


@pytest.fixture(scope="module")
def exes():
    from hmmufotu_b200 import build
    build.build_host()
    return os.path.join(BIN, "hmmufotu"), os.path.join(BIN, "hu_host_test")


@pytest.fixture(scope="module")
def dbfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("db")
    db = synth.make_db(n_leaves=40, cs_len=600, occupancy=0.85, seed=21, dg_alpha=0.7)
    otree = O.OracleTree(db)
    db.seq = otree.seqs()
    up, dn = otree.messages()
    prefix = str(d / "synth")
    synth.write_db(db, up, dn, prefix)
    reads = synth.make_reads(db, 60, read_len=180, seed=4, drop_seed_prob=0.0)
    synth.write_fasta(reads, prefix + "_reads.fa")
    return dict(db=db, otree=otree, prefix=prefix, reads=reads, dir=str(d))


def test_csfm_known_answers(exes):
    """Same shape as the reference's only known-answer test (test/CSFMIndex_test.cpp:19-63)."""
    out = subprocess.run([exes[1], "csfm"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "count(ATC) = 6" in out.stdout and out.stdout.strip().endswith("OK")


def test_csfm_matches_the_reference_index_hit_for_hit(exes):
    """The host index rebuilt from <db>.msa against the reference's own CSFMIndex loaded from the <db>.csfm its
    hmmufotu-build wrote (golden: oracle/_ref/csfm_probe, see tests/golden/make_ref_golden.py): count() and every
    locate() hit (start, end, gapped CS string) of 123 patterns, 1 551 hits."""
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    with open(os.path.join(g, "csfm_locate.txt")) as fin:
        out = subprocess.run([exes[1], "locate", os.path.join(g, "gg_70_otus_GTR.msa")], stdin=fin, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    want = open(os.path.join(g, "csfm_locate.tsv")).read()
    assert out.stdout == want
    probe = os.path.join(ROOT, "oracle", "_ref", "csfm_probe")
    if os.path.exists(probe):  # the reference's index itself, live, where oracle/_ref is built
        with open(os.path.join(g, "csfm_locate.txt")) as fin:
            ref = subprocess.run([probe, os.path.join(g, "gg_70_otus_GTR.csfm")], stdin=fin, capture_output=True, text=True)
        assert ref.returncode == 0 and ref.stdout == want


REF_TEST = "/root/reference/test"


def _skeleton(exe, *args):
    out = subprocess.run([exe] + list(args), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out.stdout


def test_builder_host_half_on_handwritten_inputs(exes, tmp_path):
    """SURVEY.md 8(f-1), host half (hu_build.cpp): FASTA -> pruned MSA with counts, position-based weights and consensus,
    and Newick + MSA + taxonomy -> tree skeleton (ids, edges in file order, branch lengths, names, annotations, heights,
    leaf sequences), against what the reference's own hmmufotu-build wrote for the same inputs (tests/golden/build:
    quoted label, negative and missing lengths, partly canonical taxonomy, lower case, N, nothing to prune)."""
    g = os.path.join(ROOT, "tests", "golden", "build")
    msa = str(tmp_path / "hand.msa")
    out = subprocess.run([exes[1], "msabuild", os.path.join(g, "hand.fasta"), "hand", msa], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert open(msa, "rb").read() == open(os.path.join(g, "hand.msa"), "rb").read()
    want = _skeleton(exes[1], "ptuskel", os.path.join(g, "hand.ptu"))
    got = _skeleton(exes[1], "treebuild", os.path.join(g, "hand.fasta"), os.path.join(g, "hand.tree"), os.path.join(g, "hand_taxonomy.txt"))
    assert got == want
    assert "[k__Bac;k__Bac;p__Foo;c__Bar;s__zed]" in got      # the reference's literal annotation path (prefix repeated)
    assert "1.0000000000000001e-05" in got                      # leaf branches of length <= 0 lifted to BRANCH_EPS


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TEST, "70_otus.fasta")), reason="needs the reference's test inputs")
def test_builder_host_half_on_70_otus(exes, tmp_path):
    """The same on BASELINE.json configs[0]'s inputs: <db>.msa byte for byte (125 x 7 682 -> 1 486 columns) and all 249
    nodes of <db>.ptu field by field, doubles to the last bit."""
    import gzip
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    for db in ("gg_70_otus_GTR", "gg_70_otus_HKY85_dG"):
        msa = str(tmp_path / (db + ".msa"))
        out = subprocess.run([exes[1], "msabuild", os.path.join(REF_TEST, "70_otus.fasta"), db, msa], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert open(msa, "rb").read() == open(os.path.join(g, db + ".msa"), "rb").read()
    ptu = str(tmp_path / "ref.ptu")
    with gzip.open(os.path.join(g, "gg_70_otus_GTR.ptu.gz"), "rb") as f, open(ptu, "wb") as o:
        o.write(f.read())
    want = _skeleton(exes[1], "ptuskel", ptu)
    got = _skeleton(exes[1], "treebuild", os.path.join(REF_TEST, "70_otus.fasta"), os.path.join(REF_TEST, "70_otus.tree"),
                    os.path.join(REF_TEST, "70_otus_taxonomy.txt"))
    assert got == want and got.startswith("nodes 249 csLen 1486 root 0")


REF_DM = "/root/reference/data/gg_97_otus.dm"


@pytest.mark.skipif(not (os.path.exists(REF_DM) and os.path.exists(os.path.join(REF_TEST, "70_otus.fasta"))), reason="needs the reference's prior file and test inputs")
def test_builder_profile_training(exes, tmp_path):
    """BandedHMMP7::build (hu_build.cpp) against the .hmm files the reference's hmmufotu-build wrote: match columns, MAP and
    CONS annotations, match / insert emissions, effective sequence number and -- with the transition counts taken the way the
    reference's GCC build takes them (its loop reads a shadowed, never-assigned variable; the disassembly shows it as M) --
    every transition line: identical text except the DATE tag, on 70_otus (K = 1 285) and on the hand-written alignment.
    `--intended` (next state as the loop evidently means it) changes transition lines only and still writes distributions."""
    g70, gb = os.path.join(ROOT, "tests", "golden", "ref70"), os.path.join(ROOT, "tests", "golden", "build")
    strip = lambda path: [ln for ln in open(path).read().split("\n") if not ln.startswith("DATE")]
    for fasta, name, want in ((os.path.join(REF_TEST, "70_otus.fasta"), "gg_70_otus_GTR", os.path.join(g70, "gg_70_otus_GTR.hmm")),
                              (os.path.join(gb, "hand.fasta"), "hand", os.path.join(gb, "hand.hmm"))):
        out = str(tmp_path / (name + ".hmm"))
        r = subprocess.run([exes[1], "hmmbuild", fasta, name, REF_DM, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert strip(out) == strip(want)
    out2 = str(tmp_path / "intended.hmm")
    r = subprocess.run([exes[1], "hmmbuild", os.path.join(REF_TEST, "70_otus.fasta"), "gg_70_otus_GTR", REF_DM, out2, "--intended"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = strip(str(tmp_path / "gg_70_otus_GTR.hmm")), strip(out2)
    assert len(a) == len(b)
    body = a.index("HMM\t\tA\tC\tG\tT") + 2
    ndiff = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            assert i >= body and (i - body) % 3 == 2, (i, x, y)       # only the third line of a node: the transitions
            ndiff += 1
            t = [0.0 if v == "*" else np.exp(-float(v)) for v in y.split()]
            last = i + 3 >= len(a) - 2                                    # node K: M->D is forced to 0 after the estimate, not renormalised
            assert (last or abs(sum(t[0:3]) - 1) < 1e-4) and abs(sum(t[3:5]) - 1) < 1e-4 and abs(sum(t[5:7]) - 1) < 1e-4
    assert ndiff > 100


def test_builder_root_block_and_shape_estimate(exes, tmp_path):
    """What the builder derives from the evaluated messages on the host (hu_build.cpp): the root log-likelihood block
    (loglik(root, j) over the children's messages, discrete-Gamma mixture included) recomputed from the messages stored in the
    reference-built databases equals the stored block (GTR and HKY85+dG4), and the shape estimate from the per-site mutation
    counts of the fixed-rate GTR messages -- including the literal root quirk: the root's own block still holds the
    INVALID_LOGLIK constant when hmmufotu-build asks, so the root infers state 0 everywhere -- is the alpha the reference's
    builder stores when run with -V on the same inputs (recorded; re-run live where the reference is available)."""
    import gzip
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    alpha = {}
    for db in ("gg_70_otus_GTR", "gg_70_otus_HKY85_dG"):
        ptu = str(tmp_path / (db + ".ptu"))
        with gzip.open(os.path.join(g, db + ".ptu.gz"), "rb") as f, open(ptu, "wb") as o:
            o.write(f.read())
        out = _skeleton(exes[1], "rootll", ptu).split()
        assert float(out[1]) < 1e-12, out
        alpha[db] = out[3]
    assert alpha["gg_70_otus_GTR"] == "1.8236165938119167"      # oracle/_ref/hmmufotu-build ... -s GTR -V -k 4 stores exactly this
    builder = os.path.join(ROOT, "oracle", "_ref", "hmmufotu-build")
    if os.path.exists(builder) and os.path.exists(os.path.join(REF_TEST, "70_otus.fasta")):
        r = subprocess.run([builder, os.path.join(REF_TEST, "70_otus.fasta"), os.path.join(REF_TEST, "70_otus.tree"), "-a",
                            os.path.join(REF_TEST, "70_otus_taxonomy.txt"), "-n", "gtrv", "-s", "GTR", "-V", "-k", "4"],
                           cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-500:]
        stored = _skeleton(exes[1], "dgshow", str(tmp_path / "gtrv.ptu")).split("\n")
        assert stored[0].split()[1] == alpha["gg_70_otus_GTR"]
        ours = _skeleton(exes[1], "dgamma", "4", alpha["gg_70_otus_GTR"]).split("\n")
        assert np.allclose([float(x) for x in ours[2].split()], [float(x) for x in stored[2].split()], rtol=1e-12, atol=0)


@pytest.mark.skipif(not (os.path.exists(REF_DM) and os.path.exists(os.path.join(REF_TEST, "70_otus.fasta"))), reason="needs the reference's data files and test inputs")
def test_builder_program_glue(exes, tmp_path):
    """hmmufotu-build of this repository end to end on BASELINE.json configs[0]'s inputs, with the one device step (all-roots
    evaluation + ancestral sequences) replaced by its TEST HOOK --messages-from (the messages of the reference-built
    database; no arithmetic stands in for the device's): <db>.msa byte-identical, <db>.hmm identical but DATE, <db>.ptu with
    the reference's skeleton, node records and edge blocks byte for byte (the files agree up to the root block / hash-ordered
    index / model text at the very end), and the result loads as a database (`hmmufotu --check-db`)."""
    import gzip
    builder = os.path.join(BIN, "hmmufotu-build")
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    for db, sm, extra in (("gg_70_otus_GTR", "gg_97_otus_GTR.sm", []), ("gg_70_otus_HKY85_dG", "gg_97_otus_HKY85.sm", ["-V", "-k", "4"])):
        ref_ptu = str(tmp_path / (db + ".ref.ptu"))
        with gzip.open(os.path.join(g, db + ".ptu.gz"), "rb") as f, open(ref_ptu, "wb") as o:
            o.write(f.read())
        r = subprocess.run([builder, os.path.join(REF_TEST, "70_otus.fasta"), os.path.join(REF_TEST, "70_otus.tree"), "-a",
                            os.path.join(REF_TEST, "70_otus_taxonomy.txt"), "-n", db, "--dm", REF_DM, "--sm", os.path.join(os.path.dirname(REF_DM), sm),
                            "--messages-from", ref_ptu] + extra, cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-800:]
        assert open(tmp_path / (db + ".msa"), "rb").read() == open(os.path.join(g, db + ".msa"), "rb").read()
        assert open(tmp_path / (db + ".csfm"), "rb").read() == open(os.path.join(g, db + ".csfm"), "rb").read()
        strip = lambda path: [ln for ln in open(path).read().split("\n") if not ln.startswith("DATE")]
        assert strip(str(tmp_path / (db + ".hmm"))) == strip(os.path.join(g, db + ".hmm"))
        ours = str(tmp_path / (db + ".ptu"))
        assert _skeleton(exes[1], "ptuskel", ours) == _skeleton(exes[1], "ptuskel", ref_ptu)
        a, b = open(ours, "rb").read(), open(ref_ptu, "rb").read()
        same = next((i for i in range(0, min(len(a), len(b)), 4096) if a[i:i + 4096] != b[i:i + 4096]), min(len(a), len(b)))
        assert same > 0.99 * len(b), (same, len(a), len(b))          # node records and every edge block (messages) are the reference's bytes
        assert float(_skeleton(exes[1], "rootll", ours).split()[1]) < 1e-12
        if extra:
            assert _skeleton(exes[1], "dgshow", ours).split("\n")[0] == _skeleton(exes[1], "dgshow", ref_ptu).split("\n")[0]
        chk = subprocess.run([exes[0], str(tmp_path / db), "--check-db"], capture_output=True, text=True)
        assert chk.returncode == 0 and "Database loaded" in chk.stderr, chk.stderr[-500:]
    # the hand-written inputs with this repository's own prior file and the model dumped from the reference-built database
    b = os.path.join(ROOT, "tests", "golden", "build")
    sm = str(tmp_path / "hand.sm")
    assert subprocess.run([exes[1], "smdump", os.path.join(b, "hand.ptu"), sm]).returncode == 0
    r = subprocess.run([builder, os.path.join(b, "hand.fasta"), os.path.join(b, "hand.tree"), "-a", os.path.join(b, "hand_taxonomy.txt"), "-n", "hand",
                        "--dm", os.path.join(b, "test_prior.dm"), "--sm", sm, "--messages-from", os.path.join(b, "hand.ptu")], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert _skeleton(exes[1], "ptucmp", str(tmp_path / "hand.ptu"), os.path.join(b, "hand.ptu")).split() == ["max_rel_diff", "0", "seq_equal", "1"]


def test_builder_csfm_file_is_the_references(exes, tmp_path):
    """<db>.csfm in libcds' serialisation (sampled suffix array, RRR bitmap of the sampled ranks, pointerless wavelet tree over
    the BWT with RRR-compressed levels; hu_build.cpp) written from <db>.msa: byte-identical to the file the reference's builder
    wrote, for 70_otus (544 055 bytes) and the hand-written alignment; and the reference's own loader answers from it exactly
    as from its own file (oracle/_ref/csfm_probe, where built)."""
    for msa, want in ((os.path.join(ROOT, "tests", "golden", "ref70", "gg_70_otus_GTR.msa"), os.path.join(ROOT, "tests", "golden", "ref70", "gg_70_otus_GTR.csfm")),
                      (os.path.join(ROOT, "tests", "golden", "build", "hand.msa"), os.path.join(ROOT, "tests", "golden", "build", "hand.csfm"))):
        out = str(tmp_path / (os.path.basename(msa) + ".csfm"))
        r = subprocess.run([exes[1], "csfmwrite", msa, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == open(want, "rb").read()
    probe = os.path.join(ROOT, "oracle", "_ref", "csfm_probe")
    if os.path.exists(probe):
        g = os.path.join(ROOT, "tests", "golden", "ref70")
        with open(os.path.join(g, "csfm_locate.txt")) as fin:
            ref = subprocess.run([probe, str(tmp_path / "gg_70_otus_GTR.msa.csfm")], stdin=fin, capture_output=True, text=True)
        assert ref.returncode == 0 and ref.stdout == open(os.path.join(g, "csfm_locate.tsv")).read()
    # and back: the index read from the reference's file (RRR bitmaps and wavelet tree decoded, full suffix array recovered from
    # the BWT) is the index built from the alignment, field for field
    for stem in (os.path.join(ROOT, "tests", "golden", "ref70", "gg_70_otus_GTR"), os.path.join(ROOT, "tests", "golden", "build", "hand")):
        assert _skeleton(exes[1], "csfmload", stem + ".csfm", stem + ".msa").strip() == "identical"


def _sim_statistics(fasta, ptu, gap_frac):
    """z-scores of a keep-gap simulation against the model it was drawn from: total log-probability of the sampled bases
    (per column: normalised exp(log(P(v rc) e^U) + log(P(v (1-rc)) e^V)) at the read's own branch and branching point),
    number of gap columns (weighted gap fraction per column), plus the sample moments of start and window length."""
    import oracle_py as O
    from refdb import read_fasta
    model = O.OracleModel(ptu.model)
    s_obs = s_exp = s_var = 0.0
    g_obs = g_exp = g_var = 0.0
    starts, lens = [], []
    for rid, desc, seq in read_fasta(fasta):
        f = dict(kv.split("=", 1) for kv in desc.strip(";").split(";") if "=" in kv)
        c, p = (int(x) for x in f["branchID"].split("->"))
        rc, a, b = float(f["branchPoint"]), int(f["csStart"]), int(f["csEnd"])
        assert ptu.parent[c] == p and len(seq) == ptu.cs_len and 0.0 <= rc < 1.0 and 0 <= a <= b < ptu.cs_len
        assert set(seq[:a]) <= {"."} and set(seq[b + 1:]) <= {"."}
        starts.append(a); lens.append(b - a)
        v = ptu.blen[c]
        P1, P2 = model.Pr(v * rc), model.Pr(v * (1 - rc))
        U, V = ptu.up[c, a:b + 1], ptu.dn[c, a:b + 1]
        lu = np.log(np.exp(U - U.max(1, keepdims=True)) @ P1.T) + U.max(1, keepdims=True)
        lv = np.log(np.exp(V - V.max(1, keepdims=True)) @ P2.T) + V.max(1, keepdims=True)
        w = lu + lv
        w = np.exp(w - w.max(1, keepdims=True))
        pr = w / w.sum(1, keepdims=True)
        win = np.frombuffer(seq[a:b + 1].encode(), dtype=np.uint8)
        gap = win == ord("-")
        gf = gap_frac[a:b + 1]
        g_obs += gap.sum(); g_exp += gf.sum(); g_var += (gf * (1 - gf)).sum()
        code = np.full(win.shape, -1)
        for k, ch in enumerate("ACGT"):
            code[win == ord(ch)] = k
        assert ((code >= 0) | gap).all()
        nz = ~gap
        lp = np.log(np.maximum(pr[nz], 1e-300))
        s_obs += lp[np.arange(nz.sum()), code[nz]].sum()
        e1 = (pr[nz] * lp).sum(1)
        s_exp += e1.sum()
        s_var += ((pr[nz] * lp * lp).sum(1) - e1 * e1).sum()
    return (s_obs - s_exp) / np.sqrt(s_var), (g_obs - g_exp) / np.sqrt(g_var), np.array(starts), np.array(lens)


def test_simulator_draws_from_the_tree_model(exes, tmp_path):
    """hmmufotu-sim of this repository (statistical parity: the reference draws from Boost's generators):
    3 000 keep-gap reads from the reference-built GTR database follow the model they claim to be drawn from -- sampled bases,
    gap columns, window starts and lengths -- within 4.5 sigma, and the reference's own simulator passes the same checks
    where it is built."""
    import gzip
    from refdb import load_ptu
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    with gzip.open(os.path.join(g, "gg_70_otus_GTR.ptu.gz"), "rb") as f, open(tmp_path / "db.ptu", "wb") as o:
        o.write(f.read())
    with open(os.path.join(g, "gg_70_otus_GTR.msa"), "rb") as f, open(tmp_path / "db.msa", "wb") as o:
        o.write(f.read())
    ptu = load_ptu(str(tmp_path / "db.ptu"))
    # weighted gap fraction per column from the .msa's auxiliary blocks: ... resWCount (4 x csLen f64), gapWCount (csLen f64) at the end
    raw = open(tmp_path / "db.msa", "rb").read()
    L = ptu.cs_len
    gapw = np.frombuffer(raw[-8 * L:], dtype="<f8")
    resw = np.frombuffer(raw[-8 * 5 * L:-8 * L], dtype="<f8").reshape(L, 4).sum(1)
    gap_frac = gapw / (gapw + resw)
    sims = [(os.path.join(BIN, "hmmufotu-sim"), "ours")]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "hmmufotu-sim")):
        sims.append((os.path.join(ROOT, "oracle", "_ref", "hmmufotu-sim"), "reference"))
    for exe, tag in sims:
        out = str(tmp_path / (tag + ".fa"))
        r = subprocess.run([exe, str(tmp_path / "db"), out, "-N", "3000", "-S", "11", "-k"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-500:]
        zb, zg, starts, lens = _sim_statistics(out, ptu, gap_frac)
        assert abs(zb) < 4.5 and abs(zg) < 4.5, (tag, zb, zg)
        assert abs(lens.mean() - 500) < 4.5 * 30 / np.sqrt(len(lens)) + 1.0 and abs(lens.std() - 30) < 3.0, (tag, lens.mean(), lens.std())
        assert starts.min() >= 0 and (starts + lens).max() < L
        # uniform start, cut where the window would leave the alignment: mean (L - 500) / 2 within sampling error
        assert abs(starts.mean() - (L - 500) / 2) < 4.5 * (L - 500) / np.sqrt(12 * len(starts)) + 2.0, (tag, starts.mean())
    # un-aligned single-end and paired-end output: read length, mate = reverse complement of the insert
    o1, o2 = str(tmp_path / "pe_1.fa"), str(tmp_path / "pe_2.fa")
    r = subprocess.run([sims[0][0], str(tmp_path / "db"), o1, o2, "-N", "50", "-S", "5", "-r", "150"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    from refdb import read_fasta
    a, b = read_fasta(o1), read_fasta(o2)
    assert len(a) == len(b) == 50 and all(x[0] == y[0] and len(x[2]) <= 150 and set(x[2]) <= set("ACGT") for x, y in zip(a, b))


def test_inspect_output_is_the_references(exes, tmp_path):
    """hmmufotu-inspect of this repository on the reference-built databases: summary lines, substitution model, discrete
    Gamma model, tree (Newick), node annotations and node sequences are the text the reference's hmmufotu-inspect writes
    (committed under tests/golden/inspect; compared live as well where oracle/_ref is built)."""
    import gzip, hashlib
    g, gi = os.path.join(ROOT, "tests", "golden", "ref70"), os.path.join(ROOT, "tests", "golden", "inspect")
    ours, ref = os.path.join(BIN, "hmmufotu-inspect"), os.path.join(ROOT, "oracle", "_ref", "hmmufotu-inspect")
    for db in ("gg_70_otus_GTR", "gg_70_otus_HKY85_dG"):
        for ext in ("msa", "csfm", "hmm"):
            with open(os.path.join(g, db + "." + ext), "rb") as f, open(tmp_path / (db + "." + ext), "wb") as o:
                o.write(f.read())
        with gzip.open(os.path.join(g, db + ".ptu.gz"), "rb") as f, open(tmp_path / (db + ".ptu"), "wb") as o:
            o.write(f.read())
    r = subprocess.run([ours, "gg_70_otus_GTR", "-sm", "-dg", "-t", "t.nwk", "-a", "a.txt", "-s", "s.fa"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(gi, "gg_70_otus_GTR.stdout.txt")).read()
    assert open(tmp_path / "t.nwk").read() == open(os.path.join(gi, "gg_70_otus_GTR.tree")).read()
    assert open(tmp_path / "a.txt").read() == open(os.path.join(gi, "gg_70_otus_GTR.anno.txt")).read()
    assert hashlib.md5(open(tmp_path / "s.fa", "rb").read()).hexdigest() == "7bed962f6aaa472b95c277a548f477dc"
    r = subprocess.run([ours, "gg_70_otus_HKY85_dG", "-sm", "-dg", "-t", "dt.nwk", "-a", "da.txt", "-s", "ds.fa", "-n", "--use-dbname"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(gi, "gg_70_otus_HKY85_dG.stdout.txt")).read()
    assert hashlib.md5(open(tmp_path / "ds.fa", "rb").read()).hexdigest() == "936fc6a380eecfd5fbf788fa46584b2b"
    if os.path.exists(ref):
        rr = subprocess.run([ref, "gg_70_otus_HKY85_dG", "-sm", "-dg", "-t", "rt.nwk", "-a", "ra.txt", "-s", "rs.fa", "-n", "--use-dbname"], cwd=str(tmp_path), capture_output=True, text=True)
        assert rr.returncode == 0 and rr.stdout == r.stdout
        for a, b in (("dt.nwk", "rt.nwk"), ("da.txt", "ra.txt"), ("ds.fa", "rs.fa")):
            assert open(tmp_path / a, "rb").read() == open(tmp_path / b, "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "tsv_probe")), reason="needs oracle/_ref/tsv_probe (the reference's own TSV reader)")
def test_assignment_tsv_is_read_by_the_references_consumers(tmp_path):
    """the reference's downstream programs (hmmufotu-sum, hmmufotu-jplace) read the assignment TSV with
    readProgInfo() + TSVScanner and fetch fields by name.  The only lines in which this repository's TSV differs from the
    reference's are the two comment lines in front (tests/golden/ours_cli/header.txt: as written by hmmufotu_b200/bin/hmmufotu
    on a B200; the rows are the reference's line for line, tests/test_gpu_ref.py): with them in front, the reference's own reader
    (oracle/_ref/tsv_probe) accepts the file and returns every named field exactly as from the reference's TSV."""
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    probe = os.path.join(ROOT, "oracle", "_ref", "tsv_probe")
    ref_tsv = os.path.join(g, "gg_70_otus_GTR_se0.tsv")
    body = [ln for ln in open(ref_tsv) if not ln.startswith("#")]
    hdr = open(os.path.join(ROOT, "tests", "golden", "ours_cli", "header.txt")).readlines()
    assert hdr[2] == body[0]                                    # same column header
    ours = tmp_path / "ours.tsv"
    ours.write_text("".join(hdr[:2] + body))
    a = subprocess.run([probe, str(ours)], capture_output=True, text=True)
    b = subprocess.run([probe, ref_tsv], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout == b.stdout and a.stdout.count("\n") == 64


def test_builder_discrete_gamma(exes, tmp_path):
    """DiscreteGammaModel(K, alpha) as the builder derives it (src/DiscreteGammaModel.cpp:40-55, literal: Gamma(shape alpha,
    SCALE alpha) quantiles, rates not multiplied by K) with this repository's own incomplete-gamma code: against the block the
    reference's builder stored in the HKY85+dG4 database (1e-12 relative; Boost's and these series differ in the last bits)
    and against scipy; the moment estimate of the shape against numpy."""
    import gzip
    from scipy.special import gammainc
    from scipy.stats import gamma as G
    ptu = str(tmp_path / "dg.ptu")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "ref70", "gg_70_otus_HKY85_dG.ptu.gz"), "rb") as f, open(ptu, "wb") as o:
        o.write(f.read())
    ref = _skeleton(exes[1], "dgshow", ptu).split("\n")
    K, alpha = int(ref[0].split()[0]), float(ref[0].split()[1])
    for k, a in ((K, alpha), (8, 0.31), (2, 5.5), (4, 40.0)):
        got = _skeleton(exes[1], "dgamma", str(k), repr(a)).split("\n")
        b = np.array([float(x) for x in got[1].split()])
        r = np.array([float(x) for x in got[2].split()])
        if (k, a) == (K, alpha):
            assert np.allclose(b[:-1], [float(x) for x in ref[1].split()][:-1], rtol=1e-12, atol=0)
            assert np.allclose(r, [float(x) for x in ref[2].split()], rtol=1e-12, atol=0)
        wb = G.ppf(np.arange(k) / k, a, scale=a)
        assert np.allclose(b[:-1], wb, rtol=1e-10, atol=1e-300) and np.isinf(b[-1])
        wr = np.diff(np.append(gammainc(a + 1, wb * a), 1.0))
        assert np.allclose(r, wr, rtol=1e-9, atol=1e-15)
    x = [3.0, 0.0, 7.0, 1.0, 12.0, 2.0, 0.0, 5.0]
    got = float(_skeleton(exes[1], "dgshape", *[repr(v) for v in x]))
    m, s2 = np.mean(x), np.var(x, ddof=1)
    assert abs(got - m * m / (s2 - m)) < 1e-12 * abs(got)
    assert np.isinf(float(_skeleton(exes[1], "dgshape", "4")))


def test_batched_seed_search_equals_the_one_by_one_loop(exes):
    """findSeedsBatch (backward searches on several threads, rand() picks replayed in file order) returns the paths of the
    one-by-one findSeeds loop -- which tests/test_ref_pin.py pins to the reference's alignments -- for every thread count,
    single-end, paired-end (forward read, then its reverse-complemented mate) and --single."""
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    db = os.path.join(g, "gg_70_otus_GTR")
    cases = [[os.path.join(g, "gg_70_otus_GTR_se%d.fasta" % k), str(k)] for k in range(4)]
    cases.append([os.path.join(g, "gg_70_otus_GTR_pe0_1.fasta"), "7", os.path.join(g, "gg_70_otus_GTR_pe0_2.fasta")])
    cases.append([os.path.join(g, "gg_70_otus_GTR_se0.fasta"), "0", "--single"])
    for c in cases:
        want = _skeleton(exes[1], "seeds", db, *c)
        assert want.count("\n") >= 64
        for t in (1, 3, 8):
            assert _skeleton(exes[1], "seeds", db, *(c + ["--threads", str(t)])) == want, (c, t)


def test_command_line_staging_equals_the_seed_driver(exes, tmp_path):
    """The read loop of the hmmufotu command line up to the upload (`--stage-only`, no device): batches of reads, the
    batched seed search on -p threads, forward reads then mates -- what it would hand to hu_batch_upload per read equals the
    one-by-one seed search in file order, for single-end and paired-end input, several batch sizes and thread counts."""
    import gzip
    g = os.path.join(ROOT, "tests", "golden", "ref70")
    for ext in ("msa", "hmm"):
        with open(os.path.join(g, "gg_70_otus_GTR." + ext), "rb") as f, open(tmp_path / ("db." + ext), "wb") as o:
            o.write(f.read())
    with gzip.open(os.path.join(g, "gg_70_otus_GTR.ptu.gz"), "rb") as f, open(tmp_path / "db.ptu", "wb") as o:
        o.write(f.read())
    db = str(tmp_path / "db")
    lens = {}
    for fa in ("gg_70_otus_GTR_se1.fasta", "gg_70_otus_GTR_pe0_1.fasta", "gg_70_otus_GTR_pe0_2.fasta"):
        rid = None
        for ln in open(os.path.join(g, fa)):
            if ln.startswith(">"):
                rid = (fa, ln[1:].split()[0]); lens[rid] = 0
            else:
                lens[rid] += len(ln.strip())
    for reads, seed in (([os.path.join(g, "gg_70_otus_GTR_se1.fasta")], "1"),
                        ([os.path.join(g, "gg_70_otus_GTR_pe0_1.fasta"), os.path.join(g, "gg_70_otus_GTR_pe0_2.fasta")], "7")):
        want = _skeleton(exes[1], "seeds", os.path.join(g, "gg_70_otus_GTR"), reads[0], seed, *reads[1:]).strip().split("\n")
        for batch, p in ((64, 1), (7, 4), (4096, 8)):
            # first configuration: the index rebuilt from the .msa; the others: <db>.csfm as written by the reference's builder
            if batch == 7:
                with open(os.path.join(g, "gg_70_otus_GTR.csfm"), "rb") as f, open(tmp_path / "db.csfm", "wb") as o:
                    o.write(f.read())
            got = _skeleton(exes[0], db, *reads, "-s", "1", "-S", seed, "--batch", str(batch), "-p", str(p), "--stage-only").strip().split("\n")
            assert len(got) == len(want) == 64 * len(reads)
            for i, (a, b) in enumerate(zip(got, want)):
                fa = a.split("\t")
                assert fa[:-1] == b.split("\t"), (batch, p, a, b)
                assert int(fa[-1]) == lens[(os.path.basename(reads[i % len(reads)]), fa[0])]   # staged code length = read length


def test_ptu_and_msa_roundtrip_bytes(exes, dbfiles):
    """load -> save reproduces the files byte for byte (the reference's IO tests only check the stream state)."""
    for kind in ("ptu", "msa"):
        src = dbfiles["prefix"] + "." + kind
        dst = os.path.join(dbfiles["dir"], "copy." + kind)
        out = subprocess.run([exes[1], kind, src, dst], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        a, b = open(src, "rb").read(), open(dst, "rb").read()
        if kind == "msa":
            assert a == b
        else:  # the substitution model is text: digits may be printed differently, everything else is binary
            cut = a.index(b"GTR\n# DNA")
            assert a[:cut] == b[:cut]
            # the model text carries 15 significant digits (Eigen's FullPrecision print, as the reference writes it) and
            # the human-only Q block is recomputed from what was read: save(load(x)) is a fixed point from the second
            # generation on, and everything before the model text from the first
            dst2 = os.path.join(dbfiles["dir"], "copy2." + kind)
            dst3 = os.path.join(dbfiles["dir"], "copy3." + kind)
            assert subprocess.run([exes[1], kind, dst, dst2]).returncode == 0
            assert subprocess.run([exes[1], kind, dst2, dst3]).returncode == 0
            c2 = open(dst2, "rb").read()
            assert b[:cut] == c2[:cut] and c2 == open(dst3, "rb").read()
            assert "nodes %d csLen %d root 0 model GTR dG 4" % (dbfiles["db"].n_nodes, dbfiles["db"].cs_len) in out.stdout


def test_hmm_roundtrip_matches_oracle(exes, dbfiles):
    src = dbfiles["prefix"] + ".hmm"
    dst = os.path.join(dbfiles["dir"], "copy.hmm")
    out = subprocess.run([exes[1], "hmm", src, dst], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    a, b = O.OracleHmm(path=src).params(), O.OracleHmm(path=dst).params()
    for k in ("eI", "T", "entry", "exit", "p2cs", "cs2p"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["eM"][1:], b["eM"][1:])
    K, L, chk = out.stdout.split()[1], out.stdout.split()[3], float(out.stdout.split()[5])
    assert (int(K), int(L)) == (dbfiles["db"].K, dbfiles["db"].cs_len)
    want = (a["entry"][1:] + a["exit"][1:] + a["eM"][1:, 0]).sum()
    assert abs(chk - want) < 1e-9 * abs(want)       # wing retraction on the host equals the oracle's


def test_check_db(exes, dbfiles):
    out = subprocess.run([exes[0], dbfiles["prefix"], "--check-db"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "csLen %d, profile size %d, %d nodes, model GTR + dGamma" % (dbfiles["db"].cs_len, dbfiles["db"].K, dbfiles["db"].n_nodes) in out.stderr
    bad = subprocess.run([exes[0], os.path.join(dbfiles["dir"], "nope"), "--check-db"], capture_output=True, text=True)
    assert bad.returncode != 0 and "Unable to open" in bad.stderr
    usage = subprocess.run([exes[0]], capture_output=True, text=True)
    assert usage.returncode != 0 and "Usage:" in usage.stderr


def test_seed_search_finds_the_source_columns(exes, dbfiles):
    """findSeeds (host half of alignSeq) against the generator's own exact-match paths: same read interval,
    profile interval within the few columns by which sequences holding the same 20-mer can differ."""
    out = subprocess.run([exes[1], "seeds", dbfiles["prefix"], dbfiles["prefix"] + "_reads.fa", "7"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rs = dbfiles["reads"]
    lines = out.stdout.strip().split("\n")
    assert len(lines) == rs.n
    same = 0
    for i, ln in enumerate(lines):
        f = ln.split("\t")
        n = int(f[1])
        assert n >= rs.n_paths[i]                     # the index sees every occurrence, the generator only its own
        got = [tuple(int(x) for x in p.split(",")) for p in f[2:]]
        for g in got:
            assert g[0] >= 1 and g[0] <= g[1] and g[2] >= 1 and g[2] <= g[3]
        if n == rs.n_paths[i] and all(g[2:4] == tuple(rs.vpaths[i, k, 2:4]) and abs(g[0] - rs.vpaths[i, k, 0]) <= 3 for k, g in enumerate(got)):
            same += 1
    assert same >= 0.8 * rs.n


@pytest.mark.gpu
def test_cli_end_to_end(exes, dbfiles):
    """hmmufotu <db> reads.fa -s 1 -S 0: every TSV row against the oracle fed with the same seeds."""
    tsv = os.path.join(dbfiles["dir"], "out.tsv")
    out = subprocess.run([exes[0], dbfiles["prefix"], dbfiles["prefix"] + "_reads.fa", "-o", tsv, "-s", "1", "-S", "0", "--batch", "32"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    seeds = subprocess.run([exes[1], "seeds", dbfiles["prefix"], dbfiles["prefix"] + "_reads.fa", "0"], capture_output=True, text=True).stdout.strip().split("\n")
    rows = [ln.rstrip("\n").split("\t") for ln in open(tsv) if not ln.startswith("#")]
    assert rows[0][:3] == ["id", "description", "seq_start"] and rows[0][-1] == "Q_taxon"
    rs, db, otree = dbfiles["reads"], dbfiles["db"], dbfiles["otree"]
    ohmm = O.OracleHmm(db)
    assert len(rows) - 1 == rs.n
    for i in range(rs.n):
        r = rows[1 + i]
        vp = np.array([[int(x) for x in p.split(",")] for p in seeds[i].split("\t")[2:]], dtype=np.int32).reshape(-1, 6)
        a = ohmm.align(rs.read_str(i), vp)
        assert r[0] == "read%d" % i and r[1] == "src_leaf=%d" % rs.src_leaf[i]
        assert [int(x) for x in r[2:8]] == list(a["coords"])
        assert r[8] == "%g" % a["cost"]
        assert r[9] == a["align"]
        seq = O.digitize(a["align"])
        s, e = a["coords"][4] - 1, a["coords"][5] - 1
        ids, dist = otree.get_seed(seq, s, e, 50, order=0)
        want = otree.assign(seq, s, e, ids, dist)
        est = sorted(w.estLoglik for w in want)
        tied = any(abs(x - y) <= 1e-11 * abs(x) for x, y in zip(est, est[1:]))
        if not tied:
            assert r[10] == "%d->%d" % (want[0].c, want[0].p), i
            assert r[12] == str(want[0].a)
            assert r[15] == "%g" % want[0].loglik and r[16] == "%g" % want[0].qPlace


def test_cli_rejects_what_the_reference_rejects(exes):
    """option validation of src/hmmufotu.cpp:318-348 (no GPU needed: the checks come before the database is opened)"""
    def err(*opts):
        p = subprocess.run([exes[0], "nodb", "noreads.fa", *opts], capture_output=True, text=True)
        assert p.returncode != 0
        return p.stderr
    assert "even number" in err("-C", "--num-segment", "3")
    assert "--chimera-lod must be non-negative" in err("-C", "--chimera-lod", "-1")
    assert "--chimera-err must be positive" in err("-C", "--chimera-err", "0")
    assert "Unsupported prior" in err("--prior", "depth")
    assert "Unsupported branch estimating method" in err("-m", "both")
    assert "-L|--seed-len" in err("-L", "99")
