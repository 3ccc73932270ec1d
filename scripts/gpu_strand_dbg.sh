#!/bin/bash
d=/tmp/sdbg; mkdir -p $d; G=tests/golden/ref70; n=gg_70_otus_GTR
cp $G/$n.msa $G/$n.hmm $G/$n.csfm $d/; gunzip -c $G/$n.ptu.gz > $d/$n.ptu
hmmufotu_b200/bin/hmmufotu $d/$n $G/${n}_se1.fasta -o gpurun_out/strand0_ours.tsv -s 0 -S 1 --batch 64
hmmufotu_b200/bin/hmmufotu $d/$n $G/${n}_se1.fasta -o gpurun_out/strand2_ours.tsv -s 2 -S 1 --batch 64
echo done
