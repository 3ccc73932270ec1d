#!/bin/bash
# BASELINE.json configs[0] and [1], program against program on the same box: the reference's own hmmufotu (oracle/_ref,
# compiled in place from its sources, -p all host threads) and this repository's hmmufotu (C++ host layer -> C ABI -> CUDA),
# FASTA in, TSV out, wall clock of the whole command (database load included) and of the read loop alone.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_c1_cli.sh [nReads]'
n=${1:-10000}
out=gpurun_out; mkdir -p $out
d=/tmp/c1; rm -rf $d; mkdir -p $d
G=tests/golden/ref70
REF=oracle/_ref
CLI=hmmufotu_b200/bin/hmmufotu
python -c "from hmmufotu_b200 import build; build.build_host()" || exit 1
thr=$(nproc)
report=$out/c1_cli.txt; : > $report
for db in gg_70_otus_GTR gg_70_otus_GTR_dG; do
  src=gg_70_otus_GTR
  for ext in msa hmm csfm; do cp $G/$src.$ext $d/$db.$ext; done
  gunzip -c $G/$db.ptu.gz > $d/$db.ptu
  if [ $db = gg_70_otus_GTR ]; then
    $REF/hmmufotu-sim $d/$db $d/se.fasta -N $n -S 11 -r 250 > /dev/null 2>&1 || { echo "sim failed"; exit 1; }
    reads="$d/se.fasta"; what="C1: 70_otus GTR, $n x 250 bp SE"
  else
    $REF/hmmufotu-sim $d/$db $d/pe_1.fasta $d/pe_2.fasta -N $n -S 12 -r 150 > /dev/null 2>&1 || { echo "sim failed"; exit 1; }
    reads="$d/pe_1.fasta $d/pe_2.fasta"; what="C2: 70_otus GTR+dGamma4, $n x 2x150 bp PE"
  fi
  t0=$(date +%s.%N); $REF/hmmufotu $d/$db $reads -o $d/$db.ref.tsv -s 1 -S 0 -p $thr > $d/ref.log 2>&1; rc1=$?; t1=$(date +%s.%N)
  t2=$(date +%s.%N); $CLI $d/$db $reads -o $d/$db.our.tsv -s 1 -S 0 -p $thr --batch 4096 > $d/our.log 2>&1; rc2=$?; t3=$(date +%s.%N)
  # the reference consumes rand() in task-execution order, so its seed picks (hence a few alignments) differ with the thread
  # count and from run to run; the first 256 reads are run again on ONE thread by both programs; rows are matched by read id
  python - "$d" "$db" "$reads" <<'PY'
import sys
d, db, reads = sys.argv[1], sys.argv[2], sys.argv[3].split()
for f in reads:
    out = f.replace(".fasta", ".head.fasta")
    k = 0
    with open(f) as i, open(out, "w") as o:
        for ln in i:
            if ln.startswith(">"):
                k += 1
                if k > 256: break
            o.write(ln)
PY
  heads=$(for f in $reads; do echo ${f%.fasta}.head.fasta; done)
  $REF/hmmufotu $d/$db $heads -o $d/$db.ref1.tsv -s 1 -S 0 -p 1 > /dev/null 2>&1
  $CLI $d/$db $heads -o $d/$db.our1.tsv -s 1 -S 0 -p 1 --batch 64 > /dev/null 2>&1
  # this program again on the read file taken ten times (start-up and database load weigh less)
  big=""; for f in $reads; do for k in 1 2 3 4 5 6 7 8 9 10; do cat $f; done > ${f%.fasta}.x10.fasta; big="$big ${f%.fasta}.x10.fasta"; done
  t4=$(date +%s.%N); $CLI $d/$db $big -o $d/$db.big.tsv -s 1 -S 0 -p $thr --batch 8192 > /dev/null 2>&1; rc3=$?; t5=$(date +%s.%N)
  python -c "print('  ours on %d reads (the file ten times): %.2f s = %.0f reads/s (rc %s)' % (10 * $n, $t5 - $t4, 10 * $n / ($t5 - $t4), '$rc3'))" | tee -a $report
  python - "$d" "$db" "$what" "$n" "$thr" $t0 $t1 $t2 $t3 $rc1 $rc2 <<'PY' | tee -a $report
import sys
d, db, what, n, thr = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
t0, t1, t2, t3 = map(float, sys.argv[6:10])
rows = lambda f: [l.rstrip("\n") for l in open(f) if not l.startswith("#")][1:]
a, b = rows("%s/%s.ref1.tsv" % (d, db)), rows("%s/%s.our1.tsv" % (d, db))
ida1 = {x.split("\t")[0]: x for x in a}   # the reference's rows come out in task order (reads 65.. before 1..64 on one thread), ours in file order
same1 = sum(ida1.get(y.split("\t")[0]) == y for y in b)
ra, rb = rows("%s/%s.ref.tsv" % (d, db)), rows("%s/%s.our.tsv" % (d, db))
ida = {x.split("\t")[0]: x for x in ra}
same = sum(ida.get(y.split("\t")[0]) == y for y in rb)
print("%s | reference hmmufotu -p %d: %.1f s = %.1f reads/s (rc %s) | this repository on 1 x B200: %.2f s = %.0f reads/s, database load included (rc %s) | "
      "rows: reference %d, ours %d, identical to the %d-thread reference run %d (its rand() order is not reproducible) | first 256 reads on one thread: %d of %d rows identical"
      % (what, thr, t1 - t0, n / (t1 - t0), sys.argv[10], t3 - t2, n / (t3 - t2), sys.argv[11], len(ra), len(rb), thr, same, same1, len(a)))
PY
done
