"""The product path against the REFERENCE ITSELF on a GPU: the hmmufotu command line of this repository (C++ host
layer over the C ABI, all arithmetic on the device) is run on the databases the reference's hmmufotu-build wrote
and on the reads the reference's hmmufotu assigned (tests/golden/ref70, see tests/golden/make_ref_golden.py); the
assignment TSV must be the reference's, line for line."""
import gzip
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "ref70")
CLI = os.path.join(ROOT, "hmmufotu_b200", "bin", "hmmufotu")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cli_built():
    """The command line is a build product (not tracked): make sure it exists and is current."""
    from hmmufotu_b200 import build
    build.build_host()


# gg_70_otus_GTR_dG (BASELINE.json configs[1]: GTR + dGamma4) is committed as its .ptu.gz alone: .msa / .csfm do not depend
# on the substitution model and .hmm only in its DATE line, so it is paired with gg_70_otus_GTR's other files
SHARED = {"gg_70_otus_GTR_dG": "gg_70_otus_GTR"}


@pytest.fixture(scope="module", params=["gg_70_otus_GTR", "gg_70_otus_HKY85_dG", "gg_70_otus_GTR_dG"])
def refdb_dir(request, tmp_path_factory):
    name = request.param
    d = tmp_path_factory.mktemp(name)
    for ext in (".msa", ".hmm", ".csfm"):
        shutil.copy(os.path.join(GOLD, SHARED.get(name, name) + ext), d / (name + ext))
    with gzip.open(os.path.join(GOLD, name + ".ptu.gz"), "rb") as f, open(d / (name + ".ptu"), "wb") as g:
        shutil.copyfileobj(f, g)
    return name, str(d)


def data_lines(path):
    return [ln.rstrip("\n") for ln in open(path) if not ln.startswith("#")]


def run_case(name, d, tag, seed, paired, opts=(), variant="", batch=64):
    reads = [os.path.join(GOLD, "%s_%s_%d.fasta" % (name, tag, m)) for m in (1, 2)] if paired else [os.path.join(GOLD, "%s_%s.fasta" % (name, tag))]
    out = os.path.join(d, tag + variant + ".tsv")
    p = subprocess.run([CLI, os.path.join(d, name)] + reads + ["-o", out, "-s", "1", "-S", str(seed), "--batch", str(batch)] + list(opts), capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    got, want = data_lines(out), data_lines(os.path.join(GOLD, "%s_%s%s.tsv" % (name, tag, variant)))
    assert got[0] == want[0]                      # column header
    assert len(got) == len(want) == 65
    bad = [(g.split("\t")[0], i) for i, (g, w) in enumerate(zip(got, want)) if g != w]
    return got, want, bad


def test_single_end_tsv_is_the_references(refdb_dir):
    name, d = refdb_dir
    for k in range(4):
        if not os.path.exists(os.path.join(GOLD, "%s_se%d.fasta" % (name, k))):
            continue
        got, want, bad = run_case(name, d, "se%d" % k, k, False)
        assert not bad, "rows differ from the reference's TSV: %s\n%s\n%s" % (bad[:5], got[bad[0][1]][-300:], want[bad[0][1]][-300:])


def test_paired_end_tsv_is_the_references(refdb_dir):
    name, d = refdb_dir
    got, want, bad = run_case(name, d, "pe0", 7, True)
    assert not bad, "rows differ from the reference's TSV: %s\n%s\n%s" % (bad[:5], got[bad[0][1]][-300:], want[bad[0][1]][-300:])


@pytest.mark.parametrize("variant,opts", [("weighted", ["-m", "weighted"]), ("ml", ["--ML"]), ("height", ["--prior", "height"]), ("single", ["--single"])])
def test_run_options_tsv_is_the_references(refdb_dir, variant, opts):
    name, d = refdb_dir
    if not os.path.exists(os.path.join(GOLD, "%s_se0.%s.tsv" % (name, variant))):
        pytest.fail("missing golden for %s %s" % (name, variant))
    got, want, bad = run_case(name, d, "se0", 0, False, opts, "." + variant)
    assert not bad, "rows differ from the reference's TSV: %s\n%s\n%s" % (bad[:5], got[bad[0][1]][-300:], want[bad[0][1]][-300:])


def test_chimera_check_changes_no_row(refdb_dir):
    """-C: the reference's segment log-odds are differences of the constant placeSeq() value, i.e. 0, so no read is
    ever flagged and its TSV is the one without -C (verified on the reference binary when the goldens were minted)."""
    name, d = refdb_dir
    got, want, bad = run_case(name, d, "se0", 0, False, ["-C", "--chimera-out", os.path.join(d, "chi.tsv")])
    assert not bad
    assert [ln for ln in open(os.path.join(d, "chi.tsv")) if not ln.startswith("#")][1:] == []   # header line only


def test_chimera_info_tsv_is_the_references(refdb_dir):
    """-C --chimera-info (src/hmmufotu.cpp:653-691, 736-745): the 5' / 3' segment placements over the read's common seeds
    (estimateSeq on the segment's columns, filterPlacements(--chimera-err), placeSeq, std::sort by log-likelihood) and the
    log-odds, five extra columns per row, as the reference's binary writes them"""
    name, d = refdb_dir
    if not os.path.exists(os.path.join(GOLD, "%s_se0.chimera.tsv" % name)):
        pytest.skip("the --chimera-info golden exists for the GTR database only")
    got, want, bad = run_case(name, d, "se0", 0, False, ["-C", "--chimera-info"], ".chimera")
    assert not bad, "rows differ from the reference's TSV: %s\n%s\n%s" % (bad[:5], got[bad[0][1]][-400:], want[bad[0][1]][-400:])


@pytest.mark.parametrize("batch", [7, 16])
def test_pipelined_read_loop_keeps_file_order(refdb_dir, batch):
    """the command line's read loop is a pipeline (reader -> two workspaces per device -> writer): with batches much smaller
    than the file, several batches are in flight on different workspaces and finish out of order; the TSV must still be the
    reference's, rows in file order, and the seed search's rand() stream the one of a one-by-one loop"""
    name, d = refdb_dir
    reads = os.path.join(GOLD, name + "_se0.fasta")
    out = os.path.join(d, "se0.b%d.tsv" % batch)
    p = subprocess.run([CLI, os.path.join(d, name), reads, "-o", out, "-s", "1", "-S", "0", "--batch", str(batch), "-p", "3"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert data_lines(out) == data_lines(os.path.join(GOLD, name + "_se0.tsv"))


def test_two_devices_write_the_same_tsv(refdb_dir):
    """--gpus 2: the database is built on the first device and cloned device to device; batches are dealt to both"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    name, d = refdb_dir
    got, want, bad = run_case(name, d, "pe0", 7, True, ["--gpus", "2"], batch=8)
    assert not bad


def test_device_tree_evaluation_matches_the_reference_builder(refdb_dir):
    """hu_db_create() without messages (the path the benchmark's synthetic database takes): all-roots messages and
    inferred internal sequences evaluated on the device from the leaves, against what the reference's hmmufotu-build
    stored in its .ptu (src/hmmufotu-build.cpp:455-470)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, HERE)
    import oracle_py as O
    import refdb
    from types import SimpleNamespace
    from hmmufotu_b200 import engine
    name, d = refdb_dir
    ref = refdb.load_ptu(os.path.join(d, name + ".ptu"))
    hp = O.OracleHmm(path=os.path.join(d, name + ".hmm")).params()
    is_leaf = np.ones(ref.n_nodes, dtype=bool)
    is_leaf[ref.parent[ref.parent >= 0]] = False
    seq = ref.seq.copy()
    seq[~is_leaf] = 0                                   # internal rows are placeholders: the device infers them
    ns = SimpleNamespace(K=hp["eM"].shape[0] - 1, cs_len=ref.cs_len, n_nodes=ref.n_nodes, root=ref.root, eM=hp["eM"], eI=hp["eI"],
                         T=hp["T"], p2cs=hp["p2cs"], parent=ref.parent, blen=ref.blen, height=ref.height, seq=seq, taxon=ref.taxon,
                         model=ref.model, dg_rates=ref.dg_rates)
    gdb = engine.Database(ns, messages=None, infer_internal=True)
    up, dn = gdb.messages()
    nonroot = np.arange(ref.n_nodes) != ref.root
    for got, want in ((up, ref.up), (dn, ref.dn)):
        got, want = got[nonroot], want[nonroot]
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-9, atol=0)
    assert np.array_equal(gdb.seqs(), ref.seq)          # leaves untouched, internal nodes as the reference inferred them
    gdb.close()


def test_alignment_fasta_is_the_references(refdb_dir):
    """-a: the aligned reads as the reference writes them (src/hmmufotu.cpp:709-715, 60 columns per line)"""
    name, d = refdb_dir
    want = os.path.join(GOLD, name + "_se0.aln.fasta")
    if not os.path.exists(want):
        pytest.skip("the alignment golden exists for the GTR database only")
    aln = os.path.join(d, "se0.aln.fasta")
    got, ref, bad = run_case(name, d, "se0", 0, False, ["-a", aln])
    assert not bad
    assert open(aln).read() == open(want).read()


def test_strand_autodetection_is_the_references(refdb_dir):
    """-s 0: the strand pre-test (src/hmmufotu.cpp:501-542) aligns the first reads in both orientations, drawing from
    the same rand() stream as the seed search, then fixes the strand for the run"""
    name, d = refdb_dir
    want = os.path.join(GOLD, name + "_se1.strand0.tsv")
    if not os.path.exists(want):
        pytest.skip("the strand golden exists for the GTR database only")
    out = os.path.join(d, "se1.strand0.tsv")
    p = subprocess.run([CLI, os.path.join(d, name), os.path.join(GOLD, name + "_se1.fasta"), "-o", out, "-s", "0", "-S", "1", "--batch", "64"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    got, ref = data_lines(out), data_lines(want)
    assert len(got) == len(ref)
    # the reverse-complemented reads align as short fragments (windows of a few dozen columns): all 50 seeds
    # pass the filter with estimated log-likelihoods that differ by rounding noise, and the reported placement is
    # a position in that order (third bullet).  Alignments must be identical; placements mostly are.
    aln_bad = [g.split("\t")[0] for g, w in zip(got, ref) if g.split("\t")[:10] != w.split("\t")[:10]]
    assert not aln_bad, aln_bad[:8]
    same = sum(g == w for g, w in zip(got, ref))
    assert same >= 0.7 * len(ref), same


def test_fastq_input_gives_the_same_rows(refdb_dir):
    """FASTQ reads (format guessed from the file name, src/SeqIO.cpp): the reference writes the rows of the FASTA run
    (checked on its binary when the golden was minted), and so must this command line"""
    name, d = refdb_dir
    fq = os.path.join(GOLD, name + "_se0.fastq")
    if not os.path.exists(fq):
        pytest.skip("the FASTQ golden exists for the GTR database only")
    out = os.path.join(d, "se0.fq.tsv")
    p = subprocess.run([CLI, os.path.join(d, name), fq, "-o", out, "-s", "1", "-S", "0", "--batch", "64"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert data_lines(out) == data_lines(os.path.join(GOLD, name + "_se0.tsv"))


def test_builder_on_the_device(tmp_path):
    """hmmufotu-build of this repository with the all-roots evaluation on the device (no --messages-from hook) on the
    hand-written inputs of tests/golden/build: messages within 1e-9 of the reference builder's .ptu (1.97e-11 when first run,
    profiles/r1d_cli_builder_check.txt), inferred sequences equal, skeleton identical."""
    from hmmufotu_b200 import build
    build.build_host()
    b = os.path.join(HERE, "golden", "build")
    tool = os.path.join(ROOT, "hmmufotu_b200", "bin", "hu_host_test")
    sm = str(tmp_path / "hand.sm")
    assert subprocess.run([tool, "smdump", os.path.join(b, "hand.ptu"), sm]).returncode == 0
    r = subprocess.run([os.path.join(ROOT, "hmmufotu_b200", "bin", "hmmufotu-build"), os.path.join(b, "hand.fasta"), os.path.join(b, "hand.tree"),
                        "-a", os.path.join(b, "hand_taxonomy.txt"), "-n", "hand", "--dm", os.path.join(b, "test_prior.dm"), "--sm", sm],
                       cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    out = subprocess.run([tool, "ptucmp", str(tmp_path / "hand.ptu"), os.path.join(b, "hand.ptu")], capture_output=True, text=True).stdout.split()
    assert float(out[1]) < 1e-9 and out[3] == "1", out
    sk = lambda f: subprocess.run([tool, "ptuskel", f], capture_output=True, text=True).stdout
    assert sk(str(tmp_path / "hand.ptu")) == sk(os.path.join(b, "hand.ptu"))


def _export_builder_inputs(db, tmp_path):
    """Builder inputs of a reference-built database, derived from the database itself with this repository's hmmufotu-inspect
    (the reference's own 70_otus input files are not part of this repository): the leaves' aligned sequences, the tree in Newick
    form with nodes named by id, the leaves' taxonomy, the substitution model.  Rebuilding from them gives the same node ids,
    parents, branch lengths and sequences (checked on the CPU through the --messages-from hook: messages identical)."""
    import gzip
    import shutil
    d = str(tmp_path)
    for ext in ("msa", "hmm", "csfm"):
        shutil.copy(os.path.join(GOLD, "%s.%s" % (db, ext)), os.path.join(d, "%s.%s" % (db, ext)))
    ref_ptu = os.path.join(d, db + ".ptu")
    with gzip.open(os.path.join(GOLD, db + ".ptu.gz"), "rb") as f, open(ref_ptu, "wb") as o:
        o.write(f.read())
    bind = os.path.join(ROOT, "hmmufotu_b200", "bin")
    tree, anno, seqs, sm = (os.path.join(d, n) for n in ("in.tree", "in.anno", "in.fasta", "in.sm"))
    r = subprocess.run([os.path.join(bind, "hmmufotu-inspect"), os.path.join(d, db), "-t", tree, "-a", anno, "-s", seqs], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    leaves = {ln[1:].split()[0] for ln in open(seqs) if ln.startswith(">")}
    lines = [ln for ln in open(anno) if ln.split("\t")[0] in leaves]
    with open(anno, "w") as o:
        o.writelines(lines)
    assert subprocess.run([os.path.join(bind, "hu_host_test"), "smdump", ref_ptu, sm]).returncode == 0
    return ref_ptu, seqs, tree, anno, sm


@pytest.mark.parametrize("db,extra", [("gg_70_otus_GTR", []), ("gg_70_otus_HKY85_dG", ["-V", "-k", "4"])])
def test_builder_on_the_device_at_70_otus_size(tmp_path, db, extra):
    """hmmufotu-build of this repository as ONE program run with the all-roots evaluation on the device, on
    inputs the size of BASELINE.json configs[0] / [1] (125 leaves, 249 nodes, 1 486 columns; derived from the reference-built
    databases, see _export_builder_inputs): every message within 1e-9 of what the reference's builder stored in its .ptu
    (1.4e-14 when first run), the inferred ancestral sequences equal, with -V the discrete-Gamma shape estimated from the
    device's fixed-rate evaluation equal to the reference's digit for digit (alpha = 1.8310984095294136), and the result
    loads as a database."""
    from hmmufotu_b200 import build
    build.build_host()
    bind = os.path.join(ROOT, "hmmufotu_b200", "bin")
    ref_ptu, seqs, tree, anno, sm = _export_builder_inputs(db, tmp_path)
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([os.path.join(bind, "hmmufotu-build"), seqs, tree, "-a", anno, "-n", "re70", "--dm", os.path.join(HERE, "golden", "build", "test_prior.dm"),
                        "--sm", sm] + extra, cwd=str(out), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    tool = os.path.join(bind, "hu_host_test")
    cmp_ = subprocess.run([tool, "ptucmp", str(out / "re70.ptu"), ref_ptu], capture_output=True, text=True).stdout.split()
    assert float(cmp_[1]) < 1e-9 and cmp_[3] == "1", cmp_
    if extra:
        dg = lambda f: subprocess.run([tool, "dgshow", f], capture_output=True, text=True).stdout.split("\n")[0]
        assert dg(str(out / "re70.ptu")) == dg(ref_ptu) != ""
    chk = subprocess.run([os.path.join(bind, "hmmufotu"), str(out / "re70"), "--check-db"], capture_output=True, text=True)
    assert chk.returncode == 0 and "Database loaded" in chk.stderr, chk.stderr[-500:]


def test_per_call_facades_write_the_references_rows(refdb_dir):
    """SURVEY.md 8(b), "class surface to preserve": the reference's own per-read loop (src/hmmufotu.cpp:603-751) written with
    ITS calls -- alignSeq, getSeed, estimateSeq, filterPlacements, placeSeq, calcQValues, std::sort, places[0] -- over the
    per-call facades of host/hu_facade.cpp (BandedHMMP7::calcViterbiScores / buildViterbiTrace / buildGlobalAlign and
    PTUnrooted::estimateSeq / placeSeq underneath), every stage on the device through the C ABI, one read at a time: the
    rows are the reference's TSV rows, single-end and paired-end."""
    name, d = refdb_dir
    tool = os.path.join(ROOT, "hmmufotu_b200", "bin", "hu_host_test")
    for tag, seed, paired in (("se0", 0, False), ("pe0", 7, True)):
        reads = [os.path.join(GOLD, "%s_%s_%d.fasta" % (name, tag, m)) for m in (1, 2)] if paired else [os.path.join(GOLD, "%s_%s.fasta" % (name, tag))]
        p = subprocess.run([tool, "facade", os.path.join(d, name), reads[0], str(seed)] + reads[1:], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        got = [ln for ln in p.stdout.split("\n") if ln]
        want = data_lines(os.path.join(GOLD, "%s_%s.tsv" % (name, tag)))[1:]
        assert len(got) == len(want) == 64
        bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
        assert not bad, "rows differ from the reference's TSV: %s\n%s\n%s" % (bad[:5], got[bad[0]][-300:], want[bad[0]][-300:])
