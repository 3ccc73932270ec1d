// hmmufotu-build over the C ABI (the reference's src/hmmufotu-build.cpp:308-502): aligned FASTA + Newick tree
// [+ taxonomy] -> <db>.msa, <db>.csfm, <db>.hmm, <db>.ptu.  The host half (hu_build.cpp: MSA pruning / weights / consensus, profile
// training, tree skeleton, discrete Gamma, root block, shape estimate) is checked file by file against the reference's
// builder on the CPU; the likelihood evaluation at every rooting and the ancestral sequences come from the device
// (hu_db_create with hu_tree_desc.up == NULL, hu_db_get_messages / hu_db_get_seq; checked against the reference
// builder's .ptu on the GPU).  All four database files are written: <db>.msa, <db>.csfm, <db>.hmm, <db>.ptu.
//
//   hmmufotu-build MSA.fasta TREE.newick -n DBNAME --dm PRIOR.dm --sm MODEL.sm [-a TAXONOMY] [-r ROOTNAME]
//                  [-f SYMFRAC] [-V [-k K]] [--intended-hmm] [--device N]
//   --messages-from DB.ptu  TEST HOOK: take the evaluated messages and inferred sequences from an existing database
//                           instead of the device (no arithmetic replaces the device's; lets the CPU suite run the glue)
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
	cerr << "Usage:    " << prog << "  <MSA-FILE> <TREE-FILE> -n DBNAME --dm PRIOR-FILE --sm MODEL-FILE [options]" << endl
		 << "MSA-FILE  FILE                 : multiple sequence alignment in FASTA format" << endl
		 << "TREE-FILE  FILE                : phylogenetic tree of the MSA sequences in Newick format" << endl
		 << "Options:    -n  STR            : database name (prefix of the files written)" << endl
		 << "            --dm  FILE         : Dirichlet priors for the profile training (the reference's data/gg_97_otus.dm)" << endl
		 << "            --sm  FILE         : DNA substitution model (the reference's data/gg_97_otus_<TYPE>.sm)" << endl
		 << "            -a|--anno  FILE    : tab-delimited taxonomy annotation of the tree's leaves" << endl
		 << "            -r|--root  STR     : annotation of nodes without any named ancestor [cellular_organisms]" << endl
		 << "            -f|--symfrac  DBL  : weighted symbol fraction of a match column [0.5]" << endl
		 << "            -V|--var           : among-site rate variation by a discrete Gamma model, shape estimated from the tree" << endl
		 << "            -k  INT            : number of discrete Gamma categories, 2..8 [4]" << endl
		 << "            --intended-hmm     : count profile transitions by the next state instead of as the reference's build does" << endl
		 << "            --device  INT      : CUDA device [0]" << endl;
}

struct Opts {
	vector<string> pos;
	map<string, string> kv;
	bool has(const string& a, const string& b = "") const { return kv.count(a) || (!b.empty() && kv.count(b)); }
	string get(const string& a, const string& b, const string& d) const { return kv.count(a) ? kv.find(a)->second : (!b.empty() && kv.count(b)) ? kv.find(b)->second : d; }
};

#define HU_CHECK(call) do { if ((call) != HU_OK) { cerr << "Error: " << hu_last_error() << endl; return EXIT_FAILURE; } } while (0)

// the all-roots evaluation: messages and (optionally) inferred internal sequences into `tree`
static int evaluate(PTUnrooted& tree, const BandedHMMP7& hmm, int device, bool wantSeq, const PTUnrooted* from) {
	const size_t per = (size_t) tree.csLen * 4, n = tree.numNodes();
	if (from) { // test hook: an existing database with the same skeleton supplies what the device would
		if (from->numNodes() != n || from->csLen != tree.csLen) { cerr << "Error: --messages-from does not fit this tree" << endl; return EXIT_FAILURE; }
		tree.up = from->up; tree.dn = from->dn;
		if (wantSeq)
			for (size_t id = 0; id < n; ++id)
				if (tree.neighbors[id].size() != 1) memcpy(&tree.seq[id * (size_t) tree.csLen], &from->seq[id * (size_t) tree.csLen], tree.csLen);
		return EXIT_SUCCESS;
	}
	hu_db* db = NULL;
	const hu_hmm_params hp = hmm.params();
	hu_tree_desc td = tree.treeDesc();
	td.up = NULL; td.dn = NULL; td.inferInternalSeq = wantSeq ? 1 : 0;
	const hu_model_desc md = tree.modelDesc();
	HU_CHECK(hu_db_create(&db, &hp, &td, &md, device));
	tree.up.assign(n * per, 0.0);
	tree.dn.assign(n * per, 0.0);
	HU_CHECK(hu_db_get_messages(db, 0, (int64_t) n, tree.up.data(), tree.dn.data()));
	if (wantSeq) {
		vector<int8_t> seq(n * (size_t) tree.csLen);
		HU_CHECK(hu_db_get_seq(db, 0, (int64_t) n, seq.data()));
		for (size_t id = 0; id < n; ++id) // leaves keep what the MSA gave them
			if (tree.neighbors[id].size() != 1) memcpy(&tree.seq[id * (size_t) tree.csLen], &seq[id * (size_t) tree.csLen], tree.csLen);
	}
	hu_db_destroy(db);
	return EXIT_SUCCESS;
}

int main(int argc, char* argv[]) {
	Opts o;
	static const char* const flags[] = {"-V", "--var", "--intended-hmm", "-h", "--help", "-v"};
	for (int i = 1; i < argc; ++i) {
		const string a = argv[i];
		if (a.size() > 1 && a[0] == '-' && !isdigit((unsigned char) a[1])) {
			bool isFlag = false;
			for (size_t f = 0; f < sizeof flags / sizeof flags[0]; ++f) isFlag |= a == flags[f];
			if (isFlag || i + 1 >= argc) o.kv[a] = "";
			else o.kv[a] = argv[++i];
		} else
			o.pos.push_back(a);
	}
	if (o.has("-h", "--help")) { usage(argv[0]); return EXIT_SUCCESS; }
	if (o.pos.size() != 2 || !o.has("-n") || !o.has("--dm") || !o.has("--sm")) { cerr << "Error:" << endl; usage(argv[0]); return EXIT_FAILURE; }
	const string seqFn = o.pos[0], treeFn = o.pos[1], dbName = o.get("-n", "", "");
	const double symfrac = atof(o.get("-f", "--symfrac", "0.5").c_str());
	const bool isVar = o.has("-V", "--var");
	const int K = atoi(o.get("-k", "", "4").c_str());
	const int device = atoi(o.get("--device", "", "0").c_str());
	const string rootName = o.get("-r", "--root", "cellular_organisms");
	if (!(symfrac > 0 && symfrac < 1)) { cerr << "-f|--symfrac must be between 0 and 1" << endl; return EXIT_FAILURE; }
	if (isVar && !(K >= 2 && K <= 8)) { cerr << "-k must be between 2 and 8" << endl; return EXIT_FAILURE; }

	ifstream seqIn(seqFn.c_str()), treeIn(treeFn.c_str()), dmIn(o.get("--dm", "", "").c_str()), smIn(o.get("--sm", "", "").c_str()), annoIn;
	if (!seqIn.is_open()) { cerr << "Unable to open '" << seqFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!treeIn.is_open()) { cerr << "Unable to open '" << treeFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!dmIn.is_open()) { cerr << "Unable to open the HMM prior file '" << o.get("--dm", "", "") << "'" << endl; return EXIT_FAILURE; }
	if (!smIn.is_open()) { cerr << "Unable to open the substitution model file '" << o.get("--sm", "", "") << "'" << endl; return EXIT_FAILURE; }
	if (o.has("-a", "--anno")) {
		annoIn.open(o.get("-a", "--anno", "").c_str());
		if (!annoIn.is_open()) { cerr << "Unable to open '" << o.get("-a", "--anno", "") << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	}

	// ---- MSA (:327-339) ----
	MSA msa;
	if (msa.loadFasta(seqIn) < 0 || msa.numSeq == 0) { cerr << "Unable to load MSA from '" << seqFn << "'" << endl; return EXIT_FAILURE; }
	msa.name = dbName;
	msa.prune();
	cerr << "MSA database created for " << msa.numSeq << " X " << msa.csLen << " aligned sequences" << endl;

	// ---- profile (:351-364) ----
	BandedHMMP7 hmm;
	if (!hmm.build(msa, symfrac, dmIn, !o.has("--intended-hmm"))) { cerr << "Unable to train the HMM profile (prior file '" << o.get("--dm", "", "") << "')" << endl; return EXIT_FAILURE; }
	cerr << "Banded HMM profile trained: " << hmm.getProfileSize() << " match columns" << endl;

	// ---- tree skeleton (:366-392) ----
	NewickTree nt;
	if (!nt.read(treeIn)) { cerr << "Unable to read Newick tree in '" << treeFn << "'" << endl; return EXIT_FAILURE; }
	PTUnrooted tree;
	if (!buildTree(tree, nt, msa, annoIn.is_open() ? &annoIn : NULL, rootName)) { cerr << "Unable to build the phylogenetic tree (unnamed or duplicated sequences?)" << endl; return EXIT_FAILURE; }
	size_t nLeaves = 0;
	for (size_t id = 0; id < tree.numNodes(); ++id) nLeaves += tree.neighbors[id].size() == 1;
	if (tree.msaIndex.size() != nLeaves) {
		cerr << "Unmatched MSA and Tree. Found " << tree.msaIndex.size() << " leaf sequences from MSA but expecting " << nLeaves << " leaves in the Phylogenetic Tree " << endl;
		return EXIT_FAILURE;
	}
	cerr << "Phylogenetic Tree constructed with total " << tree.numNodes() << " nodes" << endl;

	// ---- substitution model (:394-414) ----
	{
		string text((istreambuf_iterator<char>(smIn)), istreambuf_iterator<char>()), tag, type;
		istringstream scan(text);
		while (scan >> tag) if (tag == "Type:") { scan >> type; break; }
		static const char* const names[6] = {"GTR", "TN93", "HKY85", "F81", "K80", "JC69"};
		int t = -1;
		for (int i = 0; i < 6; ++i) if (type == names[i]) t = i;
		if (t < 0) { cerr << "Unable to load DNA Substitution Model from: '" << o.get("--sm", "", "") << "'" << endl; return EXIT_FAILURE; }
		tree.model.desc.type = t;
		istringstream again(text);
		again >> tree.model;
		if (again.bad()) { cerr << "Unable to load DNA Substitution Model from: '" << o.get("--sm", "", "") << "'" << endl; return EXIT_FAILURE; }
	}

	// ---- evaluation (:416-466) ----
	PTUnrooted given;
	const PTUnrooted* from = NULL;
	if (o.has("--messages-from")) {
		ifstream in(o.get("--messages-from", "", "").c_str(), ios_base::binary);
		if (loadProgInfo(in).bad() || !given.load(in)) { cerr << "Unable to load '" << o.get("--messages-from", "", "") << "'" << endl; return EXIT_FAILURE; }
		from = &given;
	}
	tree.hasDG = false;
	if (isVar && !from) { // fixed rates first, at the original root: the shape is estimated from those messages (:419-436)
		if (evaluate(tree, hmm, device, false, NULL) != EXIT_SUCCESS) return EXIT_FAILURE;
	}
	if (isVar) {
		double alpha;
		if (from) alpha = given.hasDG ? given.dG.alpha : INFINITY; // the hook's database was evaluated with ITS shape: keep them together
		else alpha = DiscreteGammaModel::estimateShape(tree.numMutations());
		if (alpha == INFINITY) cerr << "Unable to estimate the shape parameter with less than 2 alignment sites" << endl;
		else if (!(alpha > 0)) cerr << "Unable to estimate the shape parameter with near invariant rates, reducing to fixed rate model" << endl;
		else {
			cerr << "Estimated alpha = " << alpha << endl;
			tree.dG.build(K, alpha);
			tree.hasDG = true;
		}
	}
	if (evaluate(tree, hmm, device, true, from) != EXIT_SUCCESS) return EXIT_FAILURE;
	tree.updateRootLoglik();

	// ---- files (:468-502) ----
	const string msaFn = dbName + ".msa", csfmFn = dbName + ".csfm", hmmFn = dbName + ".hmm", ptuFn = dbName + ".ptu";
	ofstream msaOut(msaFn.c_str(), ios_base::binary), csfmOut(csfmFn.c_str(), ios_base::binary), hmmOut(hmmFn.c_str()), ptuOut(ptuFn.c_str(), ios_base::binary);
	if (!msaOut.is_open() || !csfmOut.is_open() || !hmmOut.is_open() || !ptuOut.is_open()) { cerr << "Unable to write the database files of '" << dbName << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	saveProgInfo(msaOut);
	msa.save(msaOut);
	{ // the seed index in libcds' serialisation, for the reference's hmmufotu (this repository's rebuilds it from the .msa)
		CSFMIndex csfm;
		csfm.build(msa);
		saveProgInfo(csfmOut);
		csfm.save(csfmOut, msa);
	}
	hmmOut << hmm;
	saveProgInfo(ptuOut);
	tree.save(ptuOut);
	if (!msaOut || !csfmOut || !hmmOut || !ptuOut) { cerr << "Unable to save the database files: " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	cerr << "Database '" << dbName << "' saved" << endl;
	return EXIT_SUCCESS;
}
