// hmmufotu-inspect over this repository's host layer (the reference's src/hmmufotu-inspect.cpp:60-280): summary of the four
// database files, and optionally the substitution model, the discrete Gamma model, the tree (Newick, nodes named by id),
// the nodes' taxonomy annotation and the node sequences (FASTA, 60 columns).  Host-side only.
//
//   hmmufotu-inspect DB [-sm] [-dg] [-t TREE-OUT] [-a ANNO-OUT] [-s SEQ-OUT [-n]] [--use-dbname]
#include "hu_host.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

using namespace std;
using namespace EGriceLab::HmmUFOtu;

static void usage(const char* prog) {
	cerr << "Usage:    " << prog << "  <DB-NAME> [options]" << endl
		 << "Options:    -sm  FLAG           : report the DNA substitution model of the database" << endl
		 << "            -dg  FLAG           : report the discrete Gamma model of the database, if it has one" << endl
		 << "            -t|--tree  FILE     : write the phylogenetic tree to FILE in Newick format" << endl
		 << "            -a|--anno  FILE     : write the taxonomy annotation of the tree nodes to FILE" << endl
		 << "            -s|--seq  FILE      : write the sequences of the tree's leaves to FILE in FASTA format" << endl
		 << "            -n|--node  FLAG     : with -s: every node's sequence, not just the leaves'" << endl
		 << "            --use-dbname  FLAG  : prefix node names with DB-NAME" << endl;
}

static void writeNewick(ostream& out, const PTUnrooted& t, long node, const string& prefix) { // convertToNewickTree + NewickTree::write
	bool first = true, any = false;
	for (size_t c = 0; c < t.neighbors[node].size(); ++c) {
		const long ch = t.neighbors[node][c];
		if (t.parent[ch] != node) continue;
		out << (first ? "(" : ",");
		writeNewick(out, t, ch, prefix);
		first = false; any = true;
	}
	if (any) out << ')';
	out << prefix << node << ':' << (t.parent[node] < 0 ? 0.0 : t.blen[node]);
}

int main(int argc, char* argv[]) {
	vector<string> pos;
	map<string, string> kv;
	for (int i = 1; i < argc; ++i) {
		const string a = argv[i];
		if (a.size() > 1 && a[0] == '-') {
			const bool flag = a == "-sm" || a == "-dg" || a == "-n" || a == "--node" || a == "--use-dbname" || a == "-h" || a == "--help" || a == "-v";
			if (flag || i + 1 >= argc) kv[a] = "";
			else kv[a] = argv[++i];
		} else
			pos.push_back(a);
	}
	if (kv.count("-h") || kv.count("--help") || pos.size() != 1) { usage(argv[0]); return pos.size() == 1 ? EXIT_SUCCESS : EXIT_FAILURE; }
	const string db = pos[0];
	const string treeFn = kv.count("-t") ? kv["-t"] : kv.count("--tree") ? kv["--tree"] : "";
	const string annoFn = kv.count("-a") ? kv["-a"] : kv.count("--anno") ? kv["--anno"] : "";
	const string seqFn = kv.count("-s") ? kv["-s"] : kv.count("--seq") ? kv["--seq"] : "";
	const bool leafOnly = !(kv.count("-n") || kv.count("--node"));
	const string prefix = kv.count("--use-dbname") ? db + "_" : "";
	ostringstream ver;
	ver << "v" << progVer[0] << "." << progVer[1] << "." << progVer[2];

	ifstream msaIn((db + ".msa").c_str(), ios_base::binary), csfmIn((db + ".csfm").c_str(), ios_base::binary), hmmIn((db + ".hmm").c_str()), ptuIn((db + ".ptu").c_str(), ios_base::binary);
	if (!msaIn.is_open()) { cerr << "Unable to open MSA data '" << db << ".msa': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!csfmIn.is_open()) { cerr << "Unable to open CSFM-index '" << db << ".csfm': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!hmmIn.is_open()) { cerr << "Unable to open HMM profile '" << db << ".hmm': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!ptuIn.is_open()) { cerr << "Unable to open PTU data '" << db << ".ptu': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	ofstream treeOut, annoOut, seqOut;
	if (!treeFn.empty()) { treeOut.open(treeFn.c_str()); if (!treeOut.is_open()) { cerr << "Unable to write to tree file '" << treeFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; } }
	if (!annoFn.empty()) { annoOut.open(annoFn.c_str()); if (!annoOut.is_open()) { cerr << "Unable to write to tree file '" << annoFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; } }
	if (!seqFn.empty()) { seqOut.open(seqFn.c_str()); if (!seqOut.is_open()) { cerr << "Unable to write to tree file '" << seqFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; } }

	MSA msa;
	if (loadProgInfo(msaIn).bad() || !msa.load(msaIn)) { cerr << "Failed to load MSA data '" << db << ".msa'" << endl; return EXIT_FAILURE; }
	const int csLen = (int) msa.getCSLen();
	cout << "MSA loaded. Version: " << ver.str() << " # of seq: " << msa.numSeq << " CS length: " << csLen << endl;
	CSFMIndex csfm;
	if (loadProgInfo(csfmIn).bad() || !csfm.load(csfmIn)) { cerr << "Failed to load CSFM-index '" << db << ".csfm'" << endl; return EXIT_FAILURE; }
	cout << "CSFM-index loaded. Version: " << ver.str() << " Concatenated length: " << csfm.getConcatLen() << " CS length: " << csfm.getCSLen() << endl;
	if (csfm.getCSLen() != csLen) { cerr << "Error: Unmatched CS length between CSFM-index and MSA data" << endl; return EXIT_FAILURE; }
	BandedHMMP7 hmm;
	hmmIn >> hmm;
	if (hmmIn.bad() || hmmIn.fail()) { cerr << "Unable to read HMM profile '" << db << ".hmm'" << endl; return EXIT_FAILURE; }
	cout << "HMM profile read. Name: " << hmm.name << " Alphabet: DNA Profile size: " << hmm.getProfileSize() << endl;
	if (hmm.getProfileSize() > csLen) { cerr << "Error: HMM profile size is found greater than the MSA CS length" << endl; return EXIT_FAILURE; }
	PTUnrooted ptu;
	if (loadProgInfo(ptuIn).bad() || !ptu.load(ptuIn)) { cerr << "Unable to load Phylogenetic tree data '" << db << ".ptu'" << endl; return EXIT_FAILURE; }
	size_t nLeaves = 0;
	for (size_t i = 0; i < ptu.numNodes(); ++i) nLeaves += ptu.neighbors[i].size() == 1;
	cout << "Phylogenetic tree loaded. Vesion: " << ver.str() << " Root ID: " << ptu.root << " # of leaves: " << nLeaves << " # of nodes: " << ptu.numNodes()
	     << " # of branches: " << ptu.numNodes() << " # of sites: " << ptu.getCSLen() << endl; // numBranches() is numNodes() there (.h:532-534)
	// treeLoglik(): sum over sites of log(pi . exp(root block)), scaled as dot_product_scaled (src/PhyloTreeUnrooted.h:1341-1349, 1505-1510)
	double ll = 0;
	const double* pi = ptu.model.getPi();
	for (int j = 0; j < ptu.getCSLen(); ++j) {
		const double* v = &ptu.rootLoglik[(size_t) j * 4];
		const double maxV = max(max(v[0], v[1]), max(v[2], v[3]));
		const double minExp = -1021 / 2, scale = maxV != -INFINITY && maxV < minExp ? minExp - maxV : 0;
		double d = 0;
		for (int k = 0; k < 4; ++k) d += pi[k] * ::exp(v[k] + scale);
		ll += ::log(d) - scale;
	}
	cout << "Overall tree log-likelihood: " << ll << endl;
	if (kv.count("-sm")) cout << ptu.model;
	if (kv.count("-dg") && ptu.hasDG)
		cout << "Discrete Gamma Model is enabled for this tree" << endl << "# of categories used: " << ptu.dG.K << " Shape parameter: " << ptu.dG.alpha << endl;
	if (treeOut.is_open()) { writeNewick(treeOut, ptu, ptu.root, prefix); treeOut << ';'; }
	if (annoOut.is_open())
		for (size_t i = 0; i < ptu.numNodes(); ++i) annoOut << prefix << i << "\t" << ptu.getAnno((long) i) << endl;
	if (seqOut.is_open())
		for (size_t i = 0; i < ptu.numNodes(); ++i) {
			if (leafOnly && ptu.neighbors[i].size() != 1) continue;
			string s((size_t) csLen, '-');
			for (int j = 0; j < csLen; ++j) { const int8_t b = ptu.seq[i * (size_t) csLen + j]; if (b >= 0) s[j] = DNA::decode(b); }
			seqOut << ">" << prefix << i << (ptu.getAnno((long) i).empty() ? "" : " " + ptu.getAnno((long) i)) << endl;
			for (size_t p = 0; p < s.size(); p += 60) { seqOut.write(s.data() + p, (streamsize) min((size_t) 60, s.size() - p)); seqOut.put('\n'); }
		}
	return EXIT_SUCCESS;
}
