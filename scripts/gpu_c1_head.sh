#!/bin/bash
# debugging aid for scripts/gpu_c1_cli.sh: the first 256 reads of a simulated C1 file on ONE thread by both programs,
# stderr of this repository's run and the first differing rows
out=gpurun_out; mkdir -p $out
d=/tmp/c1h; rm -rf $d; mkdir -p $d
G=tests/golden/ref70; REF=oracle/_ref; CLI=hmmufotu_b200/bin/hmmufotu
python -c "from hmmufotu_b200 import build; build.build_host()" || exit 1
db=gg_70_otus_GTR
for ext in msa hmm csfm; do cp $G/$db.$ext $d/$db.$ext; done
gunzip -c $G/$db.ptu.gz > $d/$db.ptu
$REF/hmmufotu-sim $d/$db $d/se.fasta -N 256 -S 11 -r 250 > /dev/null 2>&1
$REF/hmmufotu $d/$db $d/se.fasta -o $d/ref1.tsv -s 1 -S 0 -p 1 > $d/ref1.log 2>&1; echo "ref rc=$?"
$CLI $d/$db $d/se.fasta -o $d/our1.tsv -s 1 -S 0 -p 1 --batch 64 > $d/our1.log 2>&1; echo "ours rc=$?"
tail -5 $d/our1.log
python - $d <<'PY' | tee $out/c1_head.txt
import sys
d = sys.argv[1]
rows = lambda f: [l.rstrip("\n") for l in open(f) if not l.startswith("#")][1:]
a, b = rows(d + "/ref1.tsv"), rows(d + "/our1.tsv")
ia = {x.split("\t")[0]: x for x in a}
same = sum(ia.get(y.split("\t")[0]) == y for y in b)
print("rows: reference %d, ours %d, identical %d" % (len(a), len(b), same))
print("reference order of ids:", [x.split("\t")[0] for x in a[:6]], "ours:", [y.split("\t")[0] for y in b[:6]])
k = 0
for y in b:
    x = ia.get(y.split("\t")[0])
    if x != y and k < 3:
        k += 1
        fx, fy = (x or "").split("\t"), y.split("\t")
        print("diff id", fy[0], [(i, u[:40], v[:40]) for i, (u, v) in enumerate(zip(fx, fy)) if u != v][:6], len(fx), len(fy))
PY
