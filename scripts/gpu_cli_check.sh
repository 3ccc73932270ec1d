#!/bin/bash
# a few seconds of GPU: the command line on a reference-built database (TSV vs the reference's), then the builder's device path
G=tests/golden/ref70; B=tests/golden/build; X=/tmp/hu_x; O=gpurun_out; mkdir -p $X $O
cp $G/gg_70_otus_GTR.msa $G/gg_70_otus_GTR.hmm $X/; gunzip -c $G/gg_70_otus_GTR.ptu.gz > $X/gg_70_otus_GTR.ptu
hmmufotu_b200/bin/hmmufotu $X/gg_70_otus_GTR $G/gg_70_otus_GTR_se0.fasta -o $X/se0.tsv -s 1 -S 0 --batch 64 2> $O/cli_se0.err; echo "se0 rc=$?"
echo "se0 differing rows: $(diff <(grep -v '^#' $X/se0.tsv) <(grep -v '^#' $G/gg_70_otus_GTR_se0.tsv) | grep -c '^<')" | tee $O/cli_check.txt
hmmufotu_b200/bin/hu_host_test smdump $B/hand.ptu $X/hand.sm
( cd $X && /root/repo/hmmufotu_b200/bin/hmmufotu-build /root/repo/$B/hand.fasta /root/repo/$B/hand.tree -a /root/repo/$B/hand_taxonomy.txt -n hand --dm /root/repo/$B/test_prior.dm --sm hand.sm 2> /root/repo/$O/build_hand.err; echo "build rc=$?" )
echo "builder device path: $(hmmufotu_b200/bin/hu_host_test ptucmp $X/hand.ptu $B/hand.ptu 2>&1)" | tee -a $O/cli_check.txt
hmmufotu_b200/bin/hmmufotu $X/gg_70_otus_GTR $G/gg_70_otus_GTR_pe0_1.fasta $G/gg_70_otus_GTR_pe0_2.fasta -o $X/pe0.tsv -s 1 -S 7 --batch 64 -p 4 2> $O/cli_pe0.err; echo "pe0 rc=$?"
echo "pe0 differing rows: $(diff <(grep -v '^#' $X/pe0.tsv) <(grep -v '^#' $G/gg_70_otus_GTR_pe0.tsv) | grep -c '^<')" | tee -a $O/cli_check.txt
