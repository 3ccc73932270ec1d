// Test driver of the host layer (CPU only), in the spirit of the reference's test/ programs:
//   csfm               the known-answer test of /root/reference/test/CSFMIndex_test.cpp:19-63
//   hmm  IN OUT        load a .hmm, save it (bHmm_IO_test.cpp); prints K L and a checksum of the costs
//   ptu  IN OUT        load a .ptu, save it (PTU_IO_test.cpp:18-59); prints a summary
//   msa  IN OUT        load a .msa, save it (MSAIO_test.cpp)
//   locate DB.msa      count() and sorted locate() hits of every pattern on stdin (compared with the reference's own index)
//   msabuild FASTA NAME OUT   the builder's MSA: load an aligned FASTA, prune, weights, consensus, save (hu_build.cpp)
//   treebuild FASTA TREE [TAX] / ptuskel DB.ptu   the tree skeleton built from Newick + MSA + taxonomy, or read from a database
//   hmmbuild FASTA NAME PRIOR.dm OUT   the builder's profile training (hu_build.cpp)
//   rootll DB.ptu      root log-likelihood block recomputed from the stored messages; shape estimate from the messages
//   dgamma K ALPHA / dgshow DB.ptu   discrete-Gamma break points and rates, built here or as stored in a database
//   csfmload DB.csfm DB.msa     the index read from the reference's file = the index built from the alignment?
//   csfmwrite DB.msa OUT.csfm   the index file as the reference's hmmufotu loads it (libcds serialisation)
//   seeds DB FASTA S   CSFM seed search of every read with srand(S): prints the ViterbiAlignPaths
#include "hu_host.h"
#include <limits>
#include <algorithm>
#include <stdexcept>

#include <stdlib.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <iostream>

using namespace std;
using namespace EGriceLab::HmmUFOtu;

static int csfmTest() {
	MSA msa;
	msa.alphabet = "DNA"; msa.name = "toy"; msa.numSeq = 4; msa.csLen = 8; msa.isPruned = true;
	const char* rows[4] = {"ATCGNATC", "-TCGAATC", "AT-GCATC", "ATCG-ATC"};
	for (int i = 0; i < 4; ++i) { msa.seqNames.push_back(string("seq") + char('1' + i)); msa.concatMSA += rows[i]; }
	msa.CS = "ATCGAATC";
	CSFMIndex csfm;
	csfm.build(msa);
	int rc = 0;
	const int32_t n = csfm.count("ATC");
	cout << "count(ATC) = " << n << endl;
	// ATC occurs at columns 1-3 in rows 1 and 4, and at columns 6-8 in all four rows
	if (n != 6) rc = 1;
	const vector<CSLoc> locs = csfm.locate("ATC");
	for (size_t i = 0; i < locs.size(); ++i) {
		cout << "loc " << locs[i].start << "-" << locs[i].end << " " << locs[i].CS << endl;
		if (!((locs[i].start == 1 && locs[i].end == 3) || (locs[i].start == 6 && locs[i].end == 8))) rc = 1;
	}
	const CSLoc gap = csfm.locateFirst("TGC"); // row 3: T(2) -(3) G(4) C(5): the CS string shows the skipped column
	cout << "TGC " << gap.start << "-" << gap.end << " " << gap.CS << endl;
	if (!(gap.start == 2 && gap.end == 5 && gap.CS == "T-GC")) rc = 1;
	if (csfm.count("GGGG") != 0 || csfm.locateOne("GGGG").isValid()) rc = 1;
	srand(1);
	const CSLoc one = csfm.locateOne("GAATC");
	cout << "GAATC " << one.start << "-" << one.end << endl;
	if (!(one.start == 4 && one.end == 8)) rc = 1;
	cout << (rc ? "FAILED" : "OK") << endl;
	return rc;
}

int main(int argc, char** argv) {
	if (argc < 2) { cerr << "usage: hu_host_test csfm | locate DB.msa | msabuild FASTA NAME OUT | treebuild FASTA TREE [TAX] | ptuskel DB.ptu | hmm IN OUT | ptu IN OUT | msa IN OUT | seeds DB FASTA SEED [MATE-FASTA]" << endl; return 2; }
	const string cmd = argv[1];
	if (cmd == "csfm") return csfmTest();
	if (cmd == "hmm" && argc == 4) {
		ifstream in(argv[2]);
		BandedHMMP7 hmm;
		in >> hmm;
		if (!in || in.bad()) { cerr << "cannot load " << argv[2] << endl; return 1; }
		ofstream out(argv[3]);
		out << hmm;
		double sum = 0;
		for (int k = 1; k <= hmm.K; ++k) sum += hmm.entryPr_cost[k] + hmm.exitPr_cost[k] + hmm.E_M_cost[(size_t) k * 4];
		cout.precision(17);
		cout << "K " << hmm.K << " L " << hmm.L << " checksum " << sum << endl;
		return out.good() ? 0 : 1;
	}
	if (cmd == "ptu" && argc == 4) {
		ifstream in(argv[2], ios_base::binary);
		if (loadProgInfo(in).bad()) return 1;
		PTUnrooted ptu;
		ptu.load(in);
		if (!in || in.bad()) { cerr << "cannot load " << argv[2] << endl; return 1; }
		ofstream out(argv[3], ios_base::binary);
		saveProgInfo(out);
		ptu.save(out);
		cout << "nodes " << ptu.numNodes() << " csLen " << ptu.csLen << " root " << ptu.root << " model " << ptu.model.modelType()
			 << " dG " << (ptu.hasDG ? ptu.dG.K : 0) << endl;
		return out.good() ? 0 : 1;
	}
	if (cmd == "msa" && argc == 4) {
		ifstream in(argv[2], ios_base::binary);
		if (loadProgInfo(in).bad()) return 1;
		MSA msa;
		msa.load(in);
		if (!in || in.bad()) { cerr << "cannot load " << argv[2] << endl; return 1; }
		ofstream out(argv[3], ios_base::binary);
		saveProgInfo(out);
		msa.save(out);
		cout << "numSeq " << msa.numSeq << " csLen " << msa.csLen << endl;
		return out.good() ? 0 : 1;
	}
	if (cmd == "msabuild" && argc == 5) { // msabuild FASTA NAME OUT.msa: the builder's first step (load, name, prune, save)
		ifstream fa(argv[2]);
		MSA msa;
		if (!fa || msa.loadFasta(fa) < 0) { cerr << "cannot load " << argv[2] << endl; return 1; }
		msa.name = argv[3];
		msa.prune();
		ofstream out(argv[4], ios_base::binary);
		saveProgInfo(out);
		msa.save(out);
		cout << "numSeq " << msa.numSeq << " csLen " << msa.csLen << " pruned " << msa.isPruned << endl;
		return out.good() ? 0 : 1;
	}
	if ((cmd == "treebuild" && (argc == 4 || argc == 5)) || (cmd == "ptuskel" && argc == 3)) {
		// treebuild FASTA TREE [TAXONOMY] / ptuskel DB.ptu: the tree skeleton (ids, edges in file order, branch lengths, names,
		// annotations, heights, leaf sequences), built here or read from a database, as exact text
		PTUnrooted tree;
		if (cmd == "treebuild") {
			ifstream fa(argv[2]), tr(argv[3]), an;
			if (argc == 5) an.open(argv[4]);
			MSA msa;
			NewickTree nt;
			if (!fa || msa.loadFasta(fa) < 0 || !tr || !nt.read(tr) || (argc == 5 && !an)) { cerr << "cannot load the inputs" << endl; return 1; }
			msa.prune();
			if (!buildTree(tree, nt, msa, argc == 5 ? &an : NULL)) { cerr << "cannot build the tree" << endl; return 1; }
		} else {
			ifstream in(argv[2], ios_base::binary);
			if (loadProgInfo(in).bad()) return 1;
			tree.load(in);
			if (!in) { cerr << "cannot load " << argv[2] << endl; return 1; }
			sort(tree.msaIndex.begin(), tree.msaIndex.end()); // the file keeps hash-map order
		}
		char buf[64];
		cout << "nodes " << tree.numNodes() << " csLen " << tree.csLen << " root " << tree.root << "\n";
		for (size_t i = 0; i < tree.numNodes(); ++i) {
			cout << i << "\t" << tree.parent[i];
			snprintf(buf, sizeof buf, "\t%.17g", tree.blen[i]); cout << buf;
			snprintf(buf, sizeof buf, "\t%.17g", tree.height[i]); cout << buf;
			snprintf(buf, sizeof buf, "\t%.17g", tree.annoDist[i]); cout << buf;
			cout << "\t[" << tree.name[i] << "]\t[" << tree.anno[i] << "]\t[" << tree.seqName[i] << "]\t";
			for (size_t k = 0; k < tree.neighbors[i].size(); ++k) cout << (k ? "," : "") << tree.neighbors[i][k];
			cout << "\t";
			if (tree.neighbors[i].size() == 1 && !tree.seqName[i].empty()) { // leaf sequences come from the MSA; inferred ones need the evaluation
				unsigned long long h = 1469598103934665603ull;
				for (int j = 0; j < tree.csLen; ++j) { h ^= (unsigned char) tree.seq[i * (size_t) tree.csLen + j]; h *= 1099511628211ull; }
				cout << h;
			}
			cout << "\n";
		}
		cout << "msaIndex";
		for (size_t i = 0; i < tree.msaIndex.size(); ++i) cout << " " << tree.msaIndex[i].first << ":" << tree.msaIndex[i].second;
		cout << "\n";
		return 0;
	}
	if ((cmd == "dgamma" && argc == 4) || (cmd == "dgshow" && argc == 3)) { // dgamma K ALPHA: built here; dgshow DB.ptu: as stored
		DiscreteGammaModel dg;
		if (cmd == "dgamma") dg.build(atoi(argv[2]), atof(argv[3]));
		else {
			ifstream in(argv[2], ios_base::binary);
			PTUnrooted tree;
			if (loadProgInfo(in).bad()) return 1;
			tree.load(in);
			if (!in || !tree.hasDG) { cerr << "no discrete Gamma model in " << argv[2] << endl; return 1; }
			dg = tree.dG;
		}
		char buf[64];
		snprintf(buf, sizeof buf, "%d %.17g", dg.K, dg.alpha); cout << buf << "\n";
		for (int i = 0; i <= dg.K; ++i) { snprintf(buf, sizeof buf, "%.17g", dg.b[i]); cout << (i ? " " : "") << buf; }
		cout << "\n";
		for (int i = 0; i < dg.K; ++i) { snprintf(buf, sizeof buf, "%.17g", dg.r[i]); cout << (i ? " " : "") << buf; }
		cout << "\n";
		return 0;
	}
	if (cmd == "dgshape" && argc >= 3) { // dgshape X1 X2 ...: moment estimate of the shape from per-site mutation counts
		vector<double> x;
		for (int i = 2; i < argc; ++i) x.push_back(atof(argv[i]));
		char buf[64];
		snprintf(buf, sizeof buf, "%.17g", DiscreteGammaModel::estimateShape(x)); cout << buf << "\n";
		return 0;
	}
	if (cmd == "hmmbuild" && (argc == 6 || (argc == 7 && string(argv[6]) == "--intended"))) { // hmmbuild FASTA NAME PRIOR.dm OUT.hmm [--intended]: the builder's profile (symfrac 0.5, the reference's default)
		ifstream fa(argv[2]), dm(argv[4]);
		MSA msa;
		if (!fa || !dm || msa.loadFasta(fa) < 0) { cerr << "cannot load the inputs" << endl; return 1; }
		msa.name = argv[3];
		msa.prune();
		BandedHMMP7 hmm;
		if (!hmm.build(msa, 0.5, dm, argc == 6)) { cerr << "cannot train the profile" << endl; return 1; }
		ofstream out(argv[5]);
		out << hmm;
		cout << "K " << hmm.getProfileSize() << " L " << hmm.getCSLen() << endl;
		return out.good() ? 0 : 1;
	}
	if (cmd == "rootll" && argc == 3) { // rootll DB.ptu: the root block recomputed from the stored messages vs the stored one; alpha from the messages
		ifstream in(argv[2], ios_base::binary);
		PTUnrooted tree;
		if (loadProgInfo(in).bad()) return 1;
		tree.load(in);
		if (!in) { cerr << "cannot load " << argv[2] << endl; return 1; }
		const vector<double> stored = tree.rootLoglik;
		tree.updateRootLoglik();
		double worst = 0;
		for (size_t i = 0; i < stored.size(); ++i) {
			const double d = fabs(tree.rootLoglik[i] - stored[i]) / max(1.0, fabs(stored[i]));
			if (!(d <= worst)) worst = d;
		}
		char buf[96];
		snprintf(buf, sizeof buf, "max_rel_diff %.3g alpha %.17g", worst, DiscreteGammaModel::estimateShape(tree.numMutations()));
		cout << buf << endl;
		return 0;
	}
	if (cmd == "smdump" && argc == 4) { // smdump DB.ptu OUT.sm: the substitution model of a database as a model file
		ifstream in(argv[2], ios_base::binary);
		PTUnrooted tree;
		if (loadProgInfo(in).bad()) return 1;
		tree.load(in);
		if (!in) { cerr << "cannot load " << argv[2] << endl; return 1; }
		ofstream out(argv[3]);
		out << tree.model;
		return out.good() ? 0 : 1;
	}
	if (cmd == "ptucmp" && argc == 4) { // ptucmp A.ptu B.ptu: largest relative difference of the messages, inferred sequences equal?
		PTUnrooted t[2];
		for (int q = 0; q < 2; ++q) {
			ifstream in(argv[2 + q], ios_base::binary);
			if (loadProgInfo(in).bad()) return 1;
			t[q].load(in);
			if (!in) { cerr << "cannot load " << argv[2 + q] << endl; return 1; }
		}
		if (t[0].numNodes() != t[1].numNodes() || t[0].csLen != t[1].csLen) { cout << "shape differs" << endl; return 1; }
		double worst = 0;
		for (int d = 0; d < 2; ++d) {
			const vector<double>& a = d ? t[0].dn : t[0].up;
			const vector<double>& b = d ? t[1].dn : t[1].up;
			for (size_t i = 0; i < a.size(); ++i) {
				if (a[i] == b[i]) continue; // also -inf against -inf
				const double e = fabs(a[i] - b[i]) / max(1.0, fabs(b[i]));
				if (!(e <= worst)) worst = e;
			}
		}
		char buf[64];
		snprintf(buf, sizeof buf, "%.3g", worst);
		cout << "max_rel_diff " << buf << " seq_equal " << (t[0].seq == t[1].seq ? 1 : 0) << endl;
		return 0;
	}
	if (cmd == "csfmwrite" && argc == 4) { // csfmwrite DB.msa OUT.csfm: the index file in libcds' serialisation
		ifstream in(argv[2], ios_base::binary);
		if (loadProgInfo(in).bad()) return 1;
		MSA msa;
		msa.load(in);
		if (!in) { cerr << "cannot load " << argv[2] << endl; return 1; }
		CSFMIndex csfm;
		csfm.build(msa);
		ofstream out(argv[3], ios_base::binary);
		saveProgInfo(out);
		csfm.save(out, msa);
		return out.good() ? 0 : 1;
	}
	if (cmd == "csfmload" && argc == 4) { // csfmload DB.csfm DB.msa: the index read from the file against the one built from the alignment
		ifstream in(argv[2], ios_base::binary), min(argv[3], ios_base::binary);
		CSFMIndex loaded, built;
		MSA msa;
		if (loadProgInfo(in).bad() || !loaded.load(in)) { cerr << "cannot load " << argv[2] << endl; return 1; }
		if (loadProgInfo(min).bad() || !msa.load(min)) { cerr << "cannot load " << argv[3] << endl; return 1; }
		built.build(msa);
		cout << (loaded == built ? "identical" : "DIFFERENT") << endl;
		return loaded == built ? 0 : 1;
	}
	if (cmd == "locate" && argc == 3) { // locate DB.msa < patterns: count and sorted hits, the format of oracle/ref_csfm_probe.cpp
		ifstream in(argv[2], ios_base::binary);
		if (loadProgInfo(in).bad()) return 1;
		MSA msa;
		msa.load(in);
		if (!in) { cerr << "cannot load " << argv[2] << endl; return 1; }
		CSFMIndex csfm;
		csfm.build(msa);
		string pat;
		while (getline(cin, pat)) {
			if (pat.empty()) continue;
			const vector<CSLoc> locs = csfm.locate(pat);
			vector<string> items;
			for (size_t i = 0; i < locs.size(); ++i)
				items.push_back(to_string(locs[i].start) + "," + to_string(locs[i].end) + "," + locs[i].CS);
			sort(items.begin(), items.end());
			cout << pat << "\t" << csfm.count(pat);
			for (size_t i = 0; i < items.size(); ++i) cout << "\t" << items[i];
			cout << "\n";
		}
		return 0;
	}
	int seedThreads = 0; // --threads N (last): the batched two-phase search of findSeedsBatch instead of the one-by-one loop
	if (argc > 6 && string(argv[argc - 2]) == "--threads") { seedThreads = atoi(argv[argc - 1]); argc -= 2; }
	const bool single = argc > 5 && string(argv[argc - 1]) == "--single"; // NGCL mode: no 3' seed (src/HmmUFOtu_main.cpp:66)
	if (single) argc--;
	if (cmd == "seeds" && (argc == 5 || argc == 6)) { // seeds DB FASTA SEED [MATE-FASTA]: rand() is drawn pair by pair, forward read first
		const string db = argv[2];
		ifstream msaIn((db + ".msa").c_str(), ios_base::binary), hmmIn((db + ".hmm").c_str());
		if (loadProgInfo(msaIn).bad()) return 1;
		MSA msa;
		msa.load(msaIn);
		BandedHMMP7 hmm;
		hmmIn >> hmm;
		if (!msaIn || !hmmIn) { cerr << "cannot load " << db << endl; return 1; }
		CSFMIndex csfm;
		csfm.build(msa);
		srand((unsigned) atoi(argv[4]));
		ifstream fa(argv[3]);
		SeqIO io(&fa, SeqIO::guessFormat(argv[3]));
		ifstream mfa;
		if (argc == 6) mfa.open(argv[5]);
		SeqIO mio(&mfa, SeqIO::guessFormat(argc == 6 ? argv[5] : argv[3]));
		if (seedThreads > 0) { // batches of 4096 reads (pairs: forward read, then its mate), as the command line stages them
			bool more = true;
			while (more) {
				vector<PrimarySeq> rs;
				while (rs.size() < 4096 * (size_t) (argc == 6 ? 2 : 1) && io.hasNext() && (argc == 5 || mio.hasNext())) {
					rs.push_back(io.nextSeq());
					if (argc == 6) rs.push_back(mio.nextSeq().revcom());
				}
				more = !rs.empty();
				vector<const PrimarySeq*> ptr;
				for (size_t q = 0; q < rs.size(); ++q) ptr.push_back(&rs[q]);
				const vector<vector<BandedHMMP7::ViterbiAlignPath> > all = findSeedsBatch(hmm, csfm, ptr, 20, 50, single ? BandedHMMP7::NGCL : BandedHMMP7::GLOBAL, seedThreads);
				for (size_t q = 0; q < rs.size(); ++q) {
					const vector<BandedHMMP7::ViterbiAlignPath>& v = all[q];
					cout << rs[q].getId() << "\t" << v.size();
					for (size_t i = 0; i < v.size(); ++i)
						cout << "\t" << v[i].start << "," << v[i].end << "," << v[i].from << "," << v[i].to << "," << v[i].nIns << "," << v[i].nDel;
					cout << "\n";
				}
			}
			return 0;
		}
		while (io.hasNext() && (argc == 5 || mio.hasNext())) {
			vector<PrimarySeq> rs(1, io.nextSeq());
			if (argc == 6) rs.push_back(mio.nextSeq().revcom()); // src/hmmufotu.cpp:609
			for (size_t q = 0; q < rs.size(); ++q) {
				const vector<BandedHMMP7::ViterbiAlignPath> v = findSeeds(hmm, csfm, rs[q], 20, 50, single ? BandedHMMP7::NGCL : BandedHMMP7::GLOBAL);
				cout << rs[q].getId() << "\t" << v.size();
				for (size_t i = 0; i < v.size(); ++i)
					cout << "\t" << v[i].start << "," << v[i].end << "," << v[i].from << "," << v[i].to << "," << v[i].nIns << "," << v[i].nDel;
				cout << endl;
			}
		}
		return 0;
	}
	if (cmd == "facade" && (argc == 5 || argc == 6)) {
		// facade DB FASTA SEED [MATE-FASTA]: the reference's per-read loop (src/hmmufotu.cpp:603-751) written with ITS OWN
		// calls -- alignSeq, getSeed, estimateSeq, filterPlacements, placeSeq, calcQValues, std::sort, places[0] -- one read
		// at a time over the per-call facades of hu_facade.cpp (device needed).  Prints the TSV data rows.
		const string db = argv[2];
		ifstream msaIn((db + ".msa").c_str(), ios_base::binary), hmmIn((db + ".hmm").c_str()), ptuIn((db + ".ptu").c_str(), ios_base::binary);
		if (loadProgInfo(msaIn).bad() || loadProgInfo(ptuIn).bad()) return 1;
		MSA msa;
		msa.load(msaIn);
		BandedHMMP7 hmm;
		hmmIn >> hmm;
		PTUnrooted ptu;
		ptu.load(ptuIn);
		if (!msaIn || !hmmIn || !ptuIn) { cerr << "cannot load " << db << endl; return 1; }
		CSFMIndex csfm;
		csfm.build(msa);
		const BandedHMMP7::align_mode mode = single ? BandedHMMP7::NGCL : BandedHMMP7::GLOBAL;
		hmm.setSequenceMode(mode);
		hmm.wingRetract();
		try {
			DeviceDB dev(hmm, ptu, 0);
			srand((unsigned) atoi(argv[4]));
			ifstream fa(argv[3]);
			SeqIO io(&fa, SeqIO::guessFormat(argv[3]));
			ifstream mfa;
			if (argc == 6) mfa.open(argv[5]);
			SeqIO mio(&mfa, SeqIO::guessFormat(argc == 6 ? argv[5] : argv[3]));
			const double inf = numeric_limits<double>::infinity();
			while (io.hasNext() && (argc == 5 || mio.hasNext())) {
				const PrimarySeq fwdRead = io.nextSeq();
				BandedHMMP7::HmmAlignment aln = alignSeq(hmm, csfm, fwdRead, 20, 50, mode);
				bool isChimera = false;
				if (argc == 6) {
					const PrimarySeq revRead = mio.nextSeq().revcom();
					const BandedHMMP7::HmmAlignment revAln = alignSeq(hmm, csfm, revRead, 20, 50, mode);
					if (!(aln.csStart <= revAln.csStart && aln.csEnd <= revAln.csEnd)) isChimera = true;
					else aln.merge(revAln);
				}
				if (isChimera) continue; // written to --chimera-out only
				const DigitalSeq seq = digitalSeq(aln.align);
				vector<PTUnrooted::PTLoc> seeds = getSeed(ptu, seq, aln.csStart - 1, aln.csEnd - 1, inf, inf);
				if (seeds.size() > 50) seeds.erase(seeds.end() - (seeds.size() - 50), seeds.end());
				vector<PTUnrooted::PTPlacement> places = estimateSeq(ptu, seq, seeds, "unweighted");
				filterPlacements(places, 20.0);
				placeSeq(ptu, seq, places, inf);
				calcQValues(places, PTUnrooted::UNIFORM);
				std::sort(places.rbegin(), places.rend(), compareByQPlace);
				cout << fwdRead.getId() << "\t" << fwdRead.getDesc() << "\t" << aln << "\t" << places[0] << endl;
			}
		} catch (const std::exception& e) {
			cerr << "facade: " << e.what() << endl;
			return 1;
		}
		return 0;
	}
	cerr << "bad arguments" << endl;
	return 2;
}
