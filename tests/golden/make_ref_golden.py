#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF (oracle/_ref: HmmUFOtu v1.5.1 compiled in place from /root/reference/src
against the stand-in Eigen/Boost headers of oracle/shim, see oracle/Makefile.ref).

    python tests/golden/make_ref_golden.py          # needs /root/reference and a built oracle/_ref

Writes tests/golden/ref70/:
  gg_70_otus_<MODEL>.{msa,csfm,hmm,ptu.gz}   database built by the reference's hmmufotu-build from its own test inputs
                                              (test/70_otus.fasta + .tree + taxonomy: BASELINE.json configs[0])
  se<k>.fasta, se<k>.tsv                      reads drawn by the reference's hmmufotu-sim and the assignment TSV of
                                              the reference's hmmufotu (-s 1 -S <k> -p 1)
  pe<k>_{1,2}.fasta, pe<k>.tsv                the same for paired-end reads
  se0.aln.fasta                               the -a alignment output of se0
  csfm_locate.{txt,tsv}                       patterns (k-mers cut from se1's reads, plus misses) and what the reference's own
                                              CSFMIndex (oracle/_ref/csfm_probe on gg_70_otus_GTR.csfm) counts and locates for them
  ../build/hand.{msa,csfm,hmm,ptu}                    the reference's hmmufotu-build on the hand-written inputs next to them (quoted label,
                                              negative and missing branch lengths, partly canonical taxonomy, lower case, N,
                                              nothing to prune): what tests/test_host.py compares the host half of the builder with
  se0.{weighted,ml,height,single}.tsv         se0 again with -m weighted / --ML / --prior height / --single (NGCL alignment mode)
  gg_70_otus_GTR_dG.ptu.gz, _se0, _pe0        (round 2, extra()) BASELINE.json configs[1]: GTR + dGamma4 database, SE and 2 x 150 PE reads;
                                              variant TSVs for both discrete-Gamma databases
Every file holds at most 64 reads: up to 64 OpenMP tasks are deferred and then run in creation order, so the
reference consumes rand() (CSFM seed picks) in file order, as the batched driver does.
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SRC = "/root/reference/test"
OUT = os.path.join(HERE, "ref70")


def run(*cmd, **kw):
    print("+", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, **kw)


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-j8", "-s"])
    os.chdir(OUT)
    for model, extra in (("GTR", []), ("HKY85", ["-V", "-k", "4"])):
        db = "gg_70_otus_" + model + ("_dG" if extra else "")
        run(os.path.join(REF, "hmmufotu-build"), os.path.join(SRC, "70_otus.fasta"), os.path.join(SRC, "70_otus.tree"),
            "-a", os.path.join(SRC, "70_otus_taxonomy.txt"), "-n", db, "-s", model, *extra)
        with open(db + ".ptu", "rb") as f, gzip.GzipFile(db + ".ptu.gz", "wb", compresslevel=9, mtime=0) as g:
            shutil.copyfileobj(f, g)
        n_files = 4 if model == "GTR" else 1
        for k in range(n_files):
            tag = "%s_se%d" % (db, k)
            args = ["-r", "250"] if k % 2 == 0 else []
            run(os.path.join(REF, "hmmufotu-sim"), db, tag + ".fasta", "-N", "64", "-S", str(k), *args)
            run(os.path.join(REF, "hmmufotu"), db, tag + ".fasta", "-o", tag + ".tsv", "-s", "1", "-S", str(k), "-p", "1")
        if model == "GTR":  # the other run options of the per-read path, on the first read file
            tag = "%s_se0" % db
            for name, opts in (("weighted", ["-m", "weighted"]), ("ml", ["--ML"]), ("height", ["--prior", "height"]), ("single", ["--single"])):
                run(os.path.join(REF, "hmmufotu"), db, tag + ".fasta", "-o", "%s.%s.tsv" % (tag, name), "-s", "1", "-S", "0", "-p", "1", *opts)
            run(os.path.join(REF, "hmmufotu"), db, tag + ".fasta", "-o", os.devnull, "-a", tag + ".aln.fasta", "-s", "1", "-S", "0", "-p", "1")
        tag = "%s_pe0" % db
        run(os.path.join(REF, "hmmufotu-sim"), db, tag + "_1.fasta", tag + "_2.fasta", "-N", "64", "-S", "7", "-r", "150")
        run(os.path.join(REF, "hmmufotu"), db, tag + "_1.fasta", tag + "_2.fasta", "-o", tag + ".tsv", "-s", "1", "-S", "7", "-p", "1")
        os.remove(db + ".ptu")
    extra()
    import random
    random.seed(5)
    seqs = [l.strip() for l in open("gg_70_otus_GTR_se1.fasta") if not l.startswith(">")]
    pats = []
    for q in seqs[:40]:
        for k in (20, 12, 8):  # 20 = the seed length of the run path; shorter ones have many hits
            i = random.randrange(0, len(q) - k)
            pats.append(q[i:i + k])
    pats += ["ACGT", "GGGGGGGGGGGGGGGGGGGG", "TTTTTTTTTT"]
    with open("csfm_locate.txt", "w") as f:
        f.write("\n".join(pats) + "\n")
    with open("csfm_locate.txt") as fin, open("csfm_locate.tsv", "w") as fout:
        subprocess.run([os.path.join(REF, "csfm_probe"), "gg_70_otus_GTR.csfm"], stdin=fin, stdout=fout, check=True)
    bdir = os.path.join(HERE, "build")
    run(os.path.join(REF, "hmmufotu-build"), "hand.fasta", "hand.tree", "-a", "hand_taxonomy.txt", "-n", "hand", "-s", "GTR", cwd=bdir)
    idir = os.path.join(HERE, "inspect")  # what the reference's hmmufotu-inspect prints for the two databases
    os.makedirs(idir, exist_ok=True)
    for db, extra in (("gg_70_otus_GTR", []), ("gg_70_otus_HKY85_dG", ["-n", "--use-dbname"])):
        with gzip.open(db + ".ptu.gz", "rb") as f, open(db + ".ptu", "wb") as o:
            shutil.copyfileobj(f, o)
        with open(os.path.join(idir, db + ".stdout.txt"), "w") as so:
            subprocess.run([os.path.join(REF, "hmmufotu-inspect"), db, "-sm", "-dg", "-t", os.path.join(idir, db + ".tree"), "-a", os.path.join(idir, db + ".anno.txt"),
                            "-s", os.path.join(idir, db + ".seq.fa")] + extra, stdout=so, check=True)
        os.remove(db + ".ptu")
        os.remove(os.path.join(idir, db + ".seq.fa"))  # large: the test keeps its md5
    for f in ("gg_70_otus_HKY85_dG.tree", "gg_70_otus_HKY85_dG.anno.txt"):
        os.remove(os.path.join(idir, f))
    for f in sorted(os.listdir(OUT)):
        print("%10d  %s" % (os.path.getsize(f), f))


def extra():
    """Round 2: BASELINE.json configs[1]'s database -- 70_otus with GTR + dGamma4 (-s GTR -V -k 4) -- with 2 x 150 bp PE reads
    and 250 bp SE reads, and the run-option variants (-m weighted, --ML, --prior height, --single) for BOTH discrete-Gamma
    databases.  <db>.msa / .csfm do not depend on the model and <db>.hmm only in its DATE line, so gg_70_otus_GTR_dG keeps
    its .ptu.gz only and the tests pair it with gg_70_otus_GTR's other three files."""
    os.chdir(OUT)
    db = "gg_70_otus_GTR_dG"
    run(os.path.join(REF, "hmmufotu-build"), os.path.join(SRC, "70_otus.fasta"), os.path.join(SRC, "70_otus.tree"),
        "-a", os.path.join(SRC, "70_otus_taxonomy.txt"), "-n", db, "-s", "GTR", "-V", "-k", "4")
    strip = lambda path: [ln for ln in open(path).read().split("\n") if not ln.startswith("DATE") and not ln.startswith("NAME")]
    assert open(db + ".csfm", "rb").read() == open("gg_70_otus_GTR.csfm", "rb").read()
    assert strip(db + ".hmm") == strip("gg_70_otus_GTR.hmm")
    for tag, k, args in (("se0", 0, ["-r", "250"]),):
        run(os.path.join(REF, "hmmufotu-sim"), db, "%s_%s.fasta" % (db, tag), "-N", "64", "-S", str(k), *args)
        run(os.path.join(REF, "hmmufotu"), db, "%s_%s.fasta" % (db, tag), "-o", "%s_%s.tsv" % (db, tag), "-s", "1", "-S", str(k), "-p", "1")
    tag = "%s_pe0" % db
    run(os.path.join(REF, "hmmufotu-sim"), db, tag + "_1.fasta", tag + "_2.fasta", "-N", "64", "-S", "7", "-r", "150")
    run(os.path.join(REF, "hmmufotu"), db, tag + "_1.fasta", tag + "_2.fasta", "-o", tag + ".tsv", "-s", "1", "-S", "7", "-p", "1")
    with open(db + ".ptu", "rb") as f, gzip.GzipFile(db + ".ptu.gz", "wb", compresslevel=9, mtime=0) as g:
        shutil.copyfileobj(f, g)
    for d in (db, "gg_70_otus_HKY85_dG"):
        if not os.path.exists(d + ".ptu"):
            with gzip.open(d + ".ptu.gz", "rb") as f, open(d + ".ptu", "wb") as o:
                shutil.copyfileobj(f, o)
        tag = "%s_se0" % d
        for name, opts in (("weighted", ["-m", "weighted"]), ("ml", ["--ML"]), ("height", ["--prior", "height"]), ("single", ["--single"])):
            run(os.path.join(REF, "hmmufotu"), d, tag + ".fasta", "-o", "%s.%s.tsv" % (tag, name), "-s", "1", "-S", "0", "-p", "1", *opts)
        os.remove(d + ".ptu")
    for ext in (".msa", ".csfm", ".hmm"):
        os.remove(db + ext)
    # -C --chimera-info: the segment placements in the output (src/hmmufotu.cpp:653-691, 736-745)
    d = "gg_70_otus_GTR"
    with gzip.open(d + ".ptu.gz", "rb") as f, open(d + ".ptu", "wb") as o:
        shutil.copyfileobj(f, o)
    run(os.path.join(REF, "hmmufotu"), d, d + "_se0.fasta", "-o", d + "_se0.chimera.tsv", "-s", "1", "-S", "0", "-p", "1", "-C", "--chimera-info")
    os.remove(d + ".ptu")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--extra":
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-j8", "-s"])
        extra()
    else:
        main()
