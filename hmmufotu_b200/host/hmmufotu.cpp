// hmmufotu -- taxonomy assignment of 16S reads: banded-HMM alignment + SEP phylogenetic placement on B200 GPUs.
// Command line, defaults and TSV output follow /root/reference/src/hmmufotu.cpp (options :77-109, defaults
// :37-57, database loading :445-498, strand pre-test :501-545, read loop :603-751).  The per-read work of the
// reference's OpenMP tasks is batched through include/hmmufotu_b200.h; the CSFM seed search stays on the
// host, in file order, so that its rand() stream is the reference's for a given -S.
#include "hu_host.h"

#include <unistd.h>

#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fstream>
#include <limits>
#include <sstream>
#include <algorithm>
#include <atomic>
#include <map>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

using namespace std;
using namespace EGriceLab::HmmUFOtu;

static const double inf = numeric_limits<double>::infinity();
static const int DEFAULT_READ_STRAND = 0, MIN_STRAND_TEST = 100, MAX_STRAND_TEST = 1000;
static const double STRAND_CONFIDENCE = 0.9;
static const int DEFAULT_SEED_LEN = 20, MAX_SEED_LEN = 25, MIN_SEED_LEN = 15, DEFAULT_SEED_REGION = 50;

static void printUsage(const string& prog) {
	cerr << "Usage:    " << prog << "  <HmmUFOtu-DB> <READ-FILE1> [READ-FILE2] [options]" << endl
		 << "READ-FILE1  FILE                 : sequence read file for the assembled/forward read" << endl
		 << "READ-FILE2  FILE                 : sequence read file for the reverse read" << endl
		 << "Options:    -o  FILE             : write the assignment output to FILE instead of stdout" << endl
		 << "            -a  FILE             : in addition to the assignment output, write the read alignment in fasta format" << endl
		 << "            --fmt  STR           : read file format (applied to all read files), supported format: 'fasta', 'fastq'" << endl
		 << "            -L|--seed-len  INT   : seed length used for banded-Hmm search [" << DEFAULT_SEED_LEN << "]" << endl
		 << "            -R  INT              : size of 5'/3' seed region for finding seed matches for CSFM-index [" << DEFAULT_SEED_REGION << "]" << endl
		 << "            --single  FLAG       : assume READ-FILE1 is single-end read instead of assembled read (partial-local HMM setting)" << endl
		 << "            -s|--strand  INT     : strand of reads/mates, 1 for 1st-strand, 2 for 2nd-strand, 0 for auto-detection [" << DEFAULT_READ_STRAND << "]" << endl
		 << "            -t|--test  INT       : use first # reads to detect the strandness of input reads/mates [" << MIN_STRAND_TEST << "]" << endl
		 << "            -i|--ignore  FLAG    : ignore forward/reverse orientation check" << endl
		 << "            -N  INT              : max # of seed nodes used in the 'Seed' stage of SEP algorithm [50]" << endl
		 << "            -d|--max-diff  DBL   : max p-dist difference allowed for sub-optimal seeds [inf]" << endl
		 << "            -H|--max-height  DBL : max height of seed nodes to be used in the SEP algorithm [inf]" << endl
		 << "            -e|--err  DBL        : max placement error used in the 'Estimate' stage of SEP algorithm [20]" << endl
		 << "            -m|--method  STR     : branch length estimating method during the estimated-placement stage, 'unweighted' or 'weighted' [unweighted]" << endl
		 << "            --ML  FLAG           : use maximum likelihood in phylogenetic placement, do not calculate posterior p-values" << endl
		 << "            --prior  STR         : method for calculating prior probability of a placement, either 'uniform' or 'height' [uniform]" << endl
		 << "            --true-loglik  FLAG  : report the root log-likelihood of each placement instead of the reference's constant (DESIGN.md 5)" << endl
		 << "            -C|--chimera  FLAG   : enable the chimera checking procedure (segment re-estimation); see DESIGN.md 5 for what it can report" << endl
		 << "            --num-segment  INT   : number of segments used in chimera checking procedure [2]" << endl
		 << "            --chimera-err  DBL   : max placement error used in the 'Estimate' stage of chimera SEP algorithm" << endl
		 << "            --chimera-lod  DBL   : min log-odd required for defining a chimera read [0]" << endl
		 << "            --chimera-out  FILE  : keep assignment output of chimera reads (bad mate orientation) in FILE" << endl
		 << "            --chimera-info  FLAG : report detailed chimera information (segment placements, log-odd) in assignment outputs" << endl
		 << "            -S|--seed  INT       : random seed used for CSFM-index seed searches, for debug only" << endl
		 << "            --gpu  INT           : first CUDA device to use [0]" << endl
		 << "            --gpus  INT          : number of devices (--gpu, --gpu + 1, ...): the database is loaded once, copied device to device, and the batches of reads are dealt to the devices; rows come out in file order [1]" << endl
		 << "            --batch  INT         : reads per GPU batch [4096]" << endl
		 << "            --align-only  FLAG   : only align the reads, do not place them" << endl
		 << "            --check-db  FLAG     : load the database files, print a summary and exit (no GPU needed)" << endl
		 << "            -v  FLAG             : enable verbose information" << endl
		 << "            --version            : show program version and exit" << endl
		 << "            -h|--help            : print this message and exit" << endl;
}

// CommandOptions (src/util/CommandOptions.cpp:32-52): "-x value" pairs, flags, positional arguments
struct CommandOptions {
	vector<string> mainOpts;
	map<string, string> opts;
	string cmd;
	CommandOptions(int argc, char** argv) {
		for (int i = 0; i < argc; ++i) cmd += string(i ? " " : "") + argv[i];
		for (int i = 1; i < argc; ++i) {
			const string a = argv[i];
			if (a.size() > 1 && a[0] == '-' && !(isdigit((unsigned char) a[1]) || a[1] == '.')) {
				if (i + 1 < argc && !(strlen(argv[i + 1]) > 1 && argv[i + 1][0] == '-' && !isdigit((unsigned char) argv[i + 1][1]))) opts[a] = argv[++i];
				else opts[a] = "";
			} else
				mainOpts.push_back(a);
		}
	}
	bool hasOpt(const string& o) const { return opts.count(o) != 0; }
	const char* getOptStr(const string& o) const { return opts.find(o)->second.c_str(); }
	string get(const string& a, const string& b, const string& dflt) const { return hasOpt(a) ? opts.find(a)->second : hasOpt(b) ? opts.find(b)->second : dflt; }
};

struct ReadBatchHost { // one batch of (pairs of) reads staged for hu_batch_upload
	vector<PrimarySeq> fwd, rev;
	vector<int8_t> codes;
	vector<int64_t> offsets;
	vector<hu_vpath> vpaths;
	vector<int32_t> nPaths;
};

static void stage(const PrimarySeq& read, const vector<BandedHMMP7::ViterbiAlignPath>& vps, ReadBatchHost& hb) {
	static int8_t table[256];
	static bool tableMade = false; // built by the reader thread, the only caller
	if (!tableMade) {
		for (int c = 0; c < 256; ++c) table[c] = DNA::encode((char) ::toupper(c));
		tableMade = true;
	}
	const size_t n = read.length(), at = hb.codes.size();
	hb.codes.resize(at + n);
	int8_t* out = hb.codes.data() + at;
	int8_t bad = 0;
	for (size_t i = 0; i < n; ++i) { out[i] = table[(unsigned char) read.seq[i]]; bad |= out[i]; }
	if (bad < 0) throw invalid_argument("read '" + read.id + "' holds a character outside the nucleotide alphabet");
	hb.offsets.push_back((int64_t) hb.codes.size());
	for (int i = 0; i < 2; ++i) {
		hu_vpath v = {0, 0, 0, 0, 0, 0};
		if (i < (int) vps.size()) { v.start = vps[i].start; v.end = vps[i].end; v.from = vps[i].from; v.to = vps[i].to; v.nIns = vps[i].nIns; v.nDel = vps[i].nDel; }
		hb.vpaths.push_back(v);
	}
	hb.nPaths.push_back((int32_t) vps.size());
}

// One batch of the read loop (src/hmmufotu.cpp:603-626, batched): up to batchSize reads (pairs), the seed search in file
// order (a pair draws rand() for the forward read first, :621-626), everything staged the way the device wants it: all
// forward reads, then all mates.  Returns the number of reads (pairs).
static int readBatch(SeqIO& fwdSeqI, SeqIO& revSeqI, bool paired, int rStrand, int batchSize, const BandedHMMP7& hmm, const CSFMIndex& csfm,
		int seedLen, int seedRegion, BandedHMMP7::align_mode mode, int nThreads, ReadBatchHost& hb, int& maxLen) {
	hb.offsets.push_back(0);
	maxLen = 1;
	while ((int) hb.fwd.size() < batchSize && fwdSeqI.hasNext() && (!paired || revSeqI.hasNext())) {
		PrimarySeq f = fwdSeqI.nextSeq();
		if (rStrand == 2 && !paired) f = f.revcom(); // wrong strand for single-strand reads, :618-619
		hb.fwd.push_back(f);
		if (paired) hb.rev.push_back(revSeqI.nextSeq().revcom());
		maxLen = max(maxLen, (int) f.length());
		if (paired) maxLen = max(maxLen, (int) hb.rev.back().length());
	}
	const int n = (int) hb.fwd.size();
	ReadBatchHost mates;
	mates.offsets.push_back(0);
	vector<const PrimarySeq*> inOrder;
	for (int i = 0; i < n; ++i) {
		inOrder.push_back(&hb.fwd[i]);
		if (paired) inOrder.push_back(&hb.rev[i]);
	}
	// backward searches on -p threads, rand() picks in this order (findSeedsBatch): the paths of the one-by-one loop
	const vector<vector<BandedHMMP7::ViterbiAlignPath> > seeds = findSeedsBatch(hmm, csfm, inOrder, seedLen, seedRegion, mode, nThreads);
	for (int i = 0; i < n; ++i) {
		stage(hb.fwd[i], seeds[paired ? 2 * i : i], hb);
		if (paired) stage(hb.rev[i], seeds[2 * i + 1], mates);
	}
	if (paired) {
		const int64_t base = hb.offsets.back();
		hb.codes.insert(hb.codes.end(), mates.codes.begin(), mates.codes.end());
		for (size_t i = 1; i < mates.offsets.size(); ++i) hb.offsets.push_back(base + mates.offsets[i]);
		hb.vpaths.insert(hb.vpaths.end(), mates.vpaths.begin(), mates.vpaths.end());
		hb.nPaths.insert(hb.nPaths.end(), mates.nPaths.begin(), mates.nPaths.end());
	}
	return n;
}

#define HU_CHECK(call) do { if ((call) != HU_OK) { cerr << "Error: " << hu_last_error() << endl; return EXIT_FAILURE; } } while (0)

int main(int argc, char* argv[]) {
	CommandOptions cmdOpts(argc, argv);
	if (cmdOpts.hasOpt("-h") || cmdOpts.hasOpt("--help")) { printUsage(argv[0]); return EXIT_SUCCESS; }
	if (cmdOpts.hasOpt("--version")) {
		cout << argv[0] << ": v" << progVer[0] << "." << progVer[1] << "." << progVer[2] << endl << "Package: " << hu_version() << endl;
		return EXIT_SUCCESS;
	}
	const bool checkDb = cmdOpts.hasOpt("--check-db");
	if (!(cmdOpts.mainOpts.size() == 2 || cmdOpts.mainOpts.size() == 3 || (checkDb && cmdOpts.mainOpts.size() == 1))) {
		cerr << "Error:" << endl;
		printUsage(argv[0]);
		return EXIT_FAILURE;
	}
	const string dbName = cmdOpts.mainOpts[0];
	string fwdFn = cmdOpts.mainOpts.size() > 1 ? cmdOpts.mainOpts[1] : "", revFn = cmdOpts.mainOpts.size() > 2 ? cmdOpts.mainOpts[2] : "";
	const string outFn = cmdOpts.get("-o", "-o", ""), alnFn = cmdOpts.get("-a", "-a", ""), chiFn = cmdOpts.get("--chimera-out", "--chimera-out", "");
	int seedLen = atoi(cmdOpts.get("-L", "--seed-len", "20").c_str());
	int seedRegion = atoi(cmdOpts.get("-R", "-R", "50").c_str());
	int rStrand = atoi(cmdOpts.get("-s", "--strand", "0").c_str());
	int nTest = atoi(cmdOpts.get("-t", "--test", "100").c_str());
	const bool ignoreOrient = cmdOpts.hasOpt("-i") || cmdOpts.hasOpt("--ignore");
	const bool alignOnly = cmdOpts.hasOpt("--align-only");
	const bool verbose = cmdOpts.hasOpt("-v");
	const BandedHMMP7::align_mode mode = cmdOpts.hasOpt("--single") ? BandedHMMP7::NGCL : BandedHMMP7::GLOBAL; // :358
	hu_sep_params sp;
	hu_sep_params_default(&sp);
	sp.maxNSeed = atoi(cmdOpts.get("-N", "-N", "50").c_str());
	if (cmdOpts.hasOpt("-d") || cmdOpts.hasOpt("--max-diff")) sp.maxDiff = atof(cmdOpts.get("-d", "--max-diff", "inf").c_str());
	if (cmdOpts.hasOpt("-H") || cmdOpts.hasOpt("--max-height")) sp.maxHeight = atof(cmdOpts.get("-H", "--max-height", "inf").c_str());
	sp.maxError = atof(cmdOpts.get("-e", "--err", "20").c_str());
	sp.useTrueLoglik = cmdOpts.hasOpt("--true-loglik");
	// -C|--chimera and its options, src/hmmufotu.cpp:248-260, 325-340.  The segment re-estimation compares placeSeq() return
	// values of the SAME window (best vs. alternative branch), and that value is the constant W * log(sum pi e) (SURVEY.md
	// F6), so chimera_lod is exactly 0 and `chimera_lod > --chimera-lod (>= 0)` never holds: with -C the reference writes the
	// same assignment rows (checked against the reference binary, tests/test_gpu_ref.py).  --chimera-info adds the segment
	// placements to every row; they are computed here as there: per segment, estimateSeq over the read's common seeds
	// restricted to the segment's columns, filterPlacements(--chimera-err), placeSeq, then std::sort by log-likelihood.
	const bool checkChimera = cmdOpts.hasOpt("-C") || cmdOpts.hasOpt("--chimera");
	const bool chimeraInfo = checkChimera && cmdOpts.hasOpt("--chimera-info");
	int numSeg = 2;
	double maxChimeraError = 20.0 / 2; // literal: maxError / numSeg of the DEFAULTS, evaluated before the options are parsed (:147)
	if (checkChimera) {
		numSeg = atoi(cmdOpts.get("--num-segment", "--num-segment", "2").c_str());
		if (!(2 <= numSeg && numSeg <= 6)) { cerr << "--num-segment must be in range [2, 6]" << endl; return EXIT_FAILURE; }
		if (numSeg % 2) { cerr << "--num-segment must be an even number" << endl; return EXIT_FAILURE; }
		if (cmdOpts.hasOpt("--chimera-err")) {
			maxChimeraError = atof(cmdOpts.get("--chimera-err", "--chimera-err", "0").c_str());
			if (!(maxChimeraError > 0)) { cerr << "--chimera-err must be positive" << endl; return EXIT_FAILURE; }
		}
		if (cmdOpts.hasOpt("--chimera-lod") && !(atof(cmdOpts.get("--chimera-lod", "--chimera-lod", "0").c_str()) >= 0)) { cerr << "--chimera-lod must be non-negative" << endl; return EXIT_FAILURE; }
	}
	{ // -m|--method, --ML, --prior: src/hmmufotu.cpp:229-246, 335-338
		const string estMethod = cmdOpts.get("-m", "--method", "unweighted");
		if (estMethod != "unweighted" && estMethod != "weighted") { cerr << "Unsupported branch estimating method '" << estMethod << "'" << endl; return EXIT_FAILURE; }
		sp.estMethod = estMethod == "weighted" ? HU_EST_WEIGHTED : HU_EST_UNWEIGHTED;
		sp.onlyML = cmdOpts.hasOpt("--ML");
		const string prior = cmdOpts.get("--prior", "--prior", "uniform");
		if (prior != "uniform" && prior != "height") { cerr << "Unsupported prior specified, check the --prior option" << endl; return EXIT_FAILURE; }
		sp.prior = prior == "height" ? HU_PRIOR_HEIGHT : HU_PRIOR_UNIFORM;
	}
	unsigned seed = (unsigned) time(NULL);
	if (cmdOpts.hasOpt("-S") || cmdOpts.hasOpt("--seed")) seed = (unsigned) atoi(cmdOpts.get("-S", "--seed", "0").c_str());
	const int device = atoi(cmdOpts.get("--gpu", "--gpu", "0").c_str());
	const int batchSize = atoi(cmdOpts.get("--batch", "--batch", "4096").c_str());
	// -p|--process (src/hmmufotu.cpp:269-272, 343-346): here the threads of the host-side seed search; the assignment does not depend on it
	const long nCpu = sysconf(_SC_NPROCESSORS_ONLN);
	const int nThreads = (cmdOpts.hasOpt("-p") || cmdOpts.hasOpt("--process")) ? atoi(cmdOpts.get("-p", "--process", "1").c_str()) : (int) (nCpu > 0 ? nCpu : 1);
	string seqFmt = cmdOpts.get("--fmt", "--fmt", "");
	// option validation, :318-348
	if (!(MIN_SEED_LEN <= seedLen && seedLen <= MAX_SEED_LEN)) { cerr << "-L|--seed-len must be in range [" << MIN_SEED_LEN << ", " << MAX_SEED_LEN << "]" << endl; return EXIT_FAILURE; }
	if (seedRegion < seedLen) { cerr << "-R cannot be smaller than -L" << endl; return EXIT_FAILURE; }
	if (!(0 <= rStrand && rStrand <= 2)) { cerr << "-s|--strand must be 0, 1, or 2" << endl; return EXIT_FAILURE; }
	if (nTest < MIN_STRAND_TEST && rStrand == 0 && !cmdOpts.hasOpt("-t") && !cmdOpts.hasOpt("--test")) nTest = MIN_STRAND_TEST;
	if (nTest > MAX_STRAND_TEST) nTest = MAX_STRAND_TEST;
	if (sp.maxNSeed < 1 || sp.maxNSeed > HU_MAX_SEED) { cerr << "-N must be in range [1, " << HU_MAX_SEED << "]" << endl; return EXIT_FAILURE; }
	if (!(sp.maxDiff >= 0)) { cerr << "-d must be non-negative" << endl; return EXIT_FAILURE; }
	if (!(sp.maxError > 0)) { cerr << "-e|--err must be positive" << endl; return EXIT_FAILURE; }
	if (batchSize < 1) { cerr << "--batch must be positive" << endl; return EXIT_FAILURE; }
	if (nThreads < 1) { cerr << "-p|--process must be positive" << endl; return EXIT_FAILURE; }
	if (seqFmt.empty() && !fwdFn.empty()) seqFmt = SeqIO::guessFormat(fwdFn);
	if (!seqFmt.empty() && seqFmt != "fasta" && seqFmt != "fastq") { cerr << "Unsupported sequence format '" << seqFmt << "'" << endl; return EXIT_FAILURE; }
	srand(seed);

	// ---- database, :352-383, :445-498 ----
	const string msaFn = dbName + ".msa", hmmFn = dbName + ".hmm", ptuFn = dbName + ".ptu";
	ifstream msaIn(msaFn.c_str(), ios_base::binary), hmmIn(hmmFn.c_str()), ptuIn(ptuFn.c_str(), ios_base::binary);
	if (!msaIn.is_open()) { cerr << "Unable to open MSA data '" << msaFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!hmmIn.is_open()) { cerr << "Unable to open HMM profile '" << hmmFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!ptuIn.is_open()) { cerr << "Unable to open PTU data '" << ptuFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (loadProgInfo(msaIn).bad()) return EXIT_FAILURE;
	MSA msa;
	msa.load(msaIn);
	if (!msaIn) { cerr << "Failed to load MSA data '" << msaFn << "'" << endl; return EXIT_FAILURE; }
	const int csLen = (int) msa.getCSLen();
	BandedHMMP7 hmm;
	hmmIn >> hmm;
	if (hmmIn.bad() || hmmIn.fail()) { cerr << "Unable to read HMM profile '" << hmmFn << "'" << endl; return EXIT_FAILURE; }
	if (hmm.getProfileSize() > csLen) { cerr << "Error: HMM profile size is found greater than the MSA CS length" << endl; return EXIT_FAILURE; }
	// <db>.csfm (libcds serialisation, src/hmmufotu.cpp:470-480) when the database has one, else the same index rebuilt from
	// the MSA (databases written by tools that do not produce the file): the two are equal field for field
	CSFMIndex csfm;
	{
		ifstream csfmIn((dbName + ".csfm").c_str(), ios_base::binary);
		if (csfmIn.is_open()) {
			if (loadProgInfo(csfmIn).bad() || !csfm.load(csfmIn)) { cerr << "Unable to load CSFM-index '" << dbName << ".csfm'" << endl; return EXIT_FAILURE; }
		} else
			csfm.build(msa);
	}
	if (csfm.getCSLen() != csLen) { cerr << "Error: Unmatched CS length between CSFM-index and MSA data" << endl; return EXIT_FAILURE; }
	if (loadProgInfo(ptuIn).bad()) return EXIT_FAILURE;
	PTUnrooted ptu;
	ptu.load(ptuIn, false); // everything but the messages: they go from the file to the device(s) edge by edge (below)
	if (!ptuIn || ptuIn.bad()) { cerr << "Unable to load Phylogenetic tree data '" << ptuFn << "'" << endl; return EXIT_FAILURE; }
	if (ptu.getCSLen() != csLen || hmm.getCSLen() != csLen) { cerr << "Error: Unmatched CS length between the database files" << endl; return EXIT_FAILURE; }
	hmm.setSequenceMode(mode);
	hmm.wingRetract();
	if (verbose || checkDb)
		cerr << "Database loaded: csLen " << csLen << ", profile size " << hmm.getProfileSize() << ", " << ptu.numNodes() << " nodes, model "
			 << ptu.model.modelType() << (ptu.hasDG ? " + dGamma" : "") << ", " << msa.numSeq << " MSA rows" << endl;
	if (checkDb) return EXIT_SUCCESS;
	if (cmdOpts.hasOpt("--stage-only")) {
		// test hook, no device needed: run the read loop's staging (readBatch) and print what would be uploaded per read,
		// in the format of `hu_host_test seeds` (pairs: forward read, then its mate)
		if (rStrand != 1 && rStrand != 2) { cerr << "--stage-only needs -s 1 or -s 2 (the strand pre-test runs on the device)" << endl; return EXIT_FAILURE; }
		const bool pairedS = !revFn.empty();
		if (rStrand == 2 && pairedS) std::swap(fwdFn, revFn);
		ifstream f1(fwdFn.c_str()), f2;
		if (pairedS) f2.open(revFn.c_str());
		if (!f1.is_open() || (pairedS && !f2.is_open())) { cerr << "Unable to open the seq files" << endl; return EXIT_FAILURE; }
		SeqIO s1(&f1, seqFmt), s2(&f2, seqFmt);
		while (s1.hasNext() && (!pairedS || s2.hasNext())) {
			ReadBatchHost hb;
			int maxLen = 1;
			const int n = readBatch(s1, s2, pairedS, rStrand, batchSize, hmm, csfm, seedLen, seedRegion, mode, nThreads, hb, maxLen);
			for (int i = 0; i < n; ++i)
				for (int m = 0; m < (pairedS ? 2 : 1); ++m) {
					const int q = m * n + i;
					cout << (m ? hb.rev[i] : hb.fwd[i]).getId() << "\t" << hb.nPaths[q];
					for (int k = 0; k < hb.nPaths[q]; ++k) {
						const hu_vpath& v = hb.vpaths[(size_t) q * 2 + k];
						cout << "\t" << v.start << "," << v.end << "," << v.from << "," << v.to << "," << v.nIns << "," << v.nDel;
					}
					cout << "\t" << (hb.offsets[q + 1] - hb.offsets[q]) << "\n";
				}
		}
		return EXIT_SUCCESS;
	}

	// ---- devices: the replica is built once, on the first device, from the messages streamed out of the .ptu (never the
	//      whole 4 x csLen x 2 (N - 1) doubles in host memory), then copied device to device (peer path) ----
	const int nGpus = atoi(cmdOpts.get("--gpus", "--gpus", "1").c_str());
	if (nGpus < 1) { cerr << "--gpus must be positive" << endl; return EXIT_FAILURE; }
	vector<hu_db*> dbs(nGpus, (hu_db*) NULL);
	{
		const hu_hmm_params hp = hmm.params();
		const hu_tree_desc td = ptu.treeDesc();
		const hu_model_desc md = ptu.modelDesc();
		HU_CHECK(hu_db_create_deferred(&dbs[0], &hp, &td, &md, device));
		struct Sink : PTUnrooted::MessageSink { // a few hundred MB of edges per host -> device transfer
			hu_db* db; size_t per, cap;
			vector<int64_t> nodes; vector<int32_t> dirs; vector<double> buf;
			bool put(long node, int dir, const double* m) {
				nodes.push_back(node); dirs.push_back(dir);
				buf.insert(buf.end(), m, m + per);
				return nodes.size() < cap || flush();
			}
			bool flush() {
				if (nodes.empty()) return true;
				const int rc = hu_db_put_messages(db, (int64_t) nodes.size(), nodes.data(), dirs.data(), buf.data());
				nodes.clear(); dirs.clear(); buf.clear();
				return rc == HU_OK;
			}
		} sink;
		sink.db = dbs[0]; sink.per = (size_t) csLen * 4; sink.cap = max<size_t>(1, ((size_t) 256 << 20) / (sink.per * sizeof(double)));
		sink.buf.reserve(sink.cap * sink.per);
		if (!ptu.streamMessages(ptuIn, sink)) { cerr << "Unable to load the messages of '" << ptuFn << "': " << hu_last_error() << endl; return EXIT_FAILURE; }
		HU_CHECK(hu_db_finish(dbs[0]));
		for (int g = 1; g < nGpus; ++g) HU_CHECK(hu_db_clone(&dbs[g], dbs[0], device + g));
		if (verbose) cerr << "Database on " << nGpus << " device(s)" << endl;
	}
	hu_db* db = dbs[0];

	// ---- strand pre-test, :501-542: align the first nTest reads in both orientations ----
	const bool paired = !revFn.empty();
	if (rStrand == 0) {
		ifstream testIn(fwdFn.c_str());
		if (!testIn.is_open()) { cerr << "Unable to test forward seq file '" << fwdFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; }
		SeqIO testSeqI(&testIn, seqFmt);
		ReadBatchHost hb;
		hb.offsets.push_back(0);
		int n = 0, maxLen = 1;
		for (; n < nTest && testSeqI.hasNext(); ++n) { // fwd then rev per read: the order alignSeq draws rand()
			// Literal behaviour: the reference's PrimarySeq::revcom() works IN PLACE (src/PrimarySeq.h:229), so
			// `PrimarySeq revRead = fwdRead.revcom();` (src/hmmufotu.cpp:525) leaves BOTH reads reverse-complemented.
			// The two alignments differ only through the rand() draws of the seed search, "fwd < rev" rarely holds,
			// and the test settles on strand 2 (then every single-end read is reverse-complemented, :617-618).
			const PrimarySeq f = testSeqI.nextSeq(), r = f.revcom();
			stage(r, findSeeds(hmm, csfm, r, seedLen, seedRegion, mode), hb);
			stage(r, findSeeds(hmm, csfm, r, seedLen, seedRegion, mode), hb);
			maxLen = max(maxLen, (int) f.length());
		}
		double fwdScore = 0, revScore = 0;
		if (n > 0) {
			hu_batch* tb = NULL;
			HU_CHECK(hu_batch_create(&tb, db, 2 * n, maxLen, 0));
			const hu_read_batch rb = {2 * n, 0, hb.codes.data(), hb.offsets.data(), hb.vpaths.data(), hb.nPaths.data()};
			HU_CHECK(hu_batch_upload(tb, &rb));
			HU_CHECK(hu_align_batch(tb, mode));
			vector<hu_align_row> rows(2 * n);
			HU_CHECK(hu_batch_get_align(tb, rows.data(), NULL, 0, NULL));
			for (int i = 0; i < n; ++i) (rows[2 * i].cost < rows[2 * i + 1].cost ? fwdScore : revScore)++;
			hu_batch_destroy(tb);
		}
		if (fwdScore >= (fwdScore + revScore) * STRAND_CONFIDENCE) rStrand = 1;
		else if (revScore >= (fwdScore + revScore) * STRAND_CONFIDENCE) rStrand = 2;
		else { cerr << "Failed to determine read strandness. Try larger -t|--test or determine manually" << endl; return EXIT_FAILURE; }
		if (verbose) cerr << "Read strand determined as " << rStrand << endl;
	}
	if (rStrand == 2 && paired) std::swap(fwdFn, revFn);

	// ---- inputs / outputs ----
	ifstream fwdIn(fwdFn.c_str()), revIn;
	if (!fwdIn.is_open()) { cerr << "Unable to open seq file '" << fwdFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (paired) {
		revIn.open(revFn.c_str());
		if (!revIn.is_open()) { cerr << "Unable to open reverse seq file '" << revFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	}
	SeqIO fwdSeqI(&fwdIn, seqFmt), revSeqI(&revIn, seqFmt);
	ofstream outFile, alnFile, chiFile;
	if (!outFn.empty()) { outFile.open(outFn.c_str()); if (!outFile.is_open()) { cerr << "Unable to write to '" << outFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; } }
	if (!alnFn.empty()) { alnFile.open(alnFn.c_str()); if (!alnFile.is_open()) { cerr << "Unable to write to '" << alnFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; } }
	if (!chiFn.empty()) { chiFile.open(chiFn.c_str()); if (!chiFile.is_open()) { cerr << "Unable to write to '" << chiFn << "' " << ::strerror(errno) << endl; return EXIT_FAILURE; } }
	ostream& out = outFn.empty() ? cout : outFile;
	static const string CHIMERA_TSV_HEADER = "seg5_taxon_id\tseg3_taxon_id\tseg5_taxon_anno\tseg3_taxon_anno\tchimera_lod"; // :57
	const string header = string("id\tdescription\t") + BandedHMMP7::HmmAlignment::TSV_HEADER + (chimeraInfo ? "\t" + CHIMERA_TSV_HEADER + "\t" : "\t") +
	                      PTUnrooted::PTPlacement::TSV_HEADER; // :590-600
	writeProgInfo(out, string(" taxonomy assignment generated by ") + argv[0]);
	out << "# command: " << cmdOpts.cmd << endl << header << endl;
	if (chiFile.is_open()) {
		writeProgInfo(chiFile, string(" taxonomy assignment generated by ") + argv[0]);
		chiFile << "# command: " << cmdOpts.cmd << endl << header << endl;
	}

	// ---- the read loop: a pipeline ----
	// reader (this thread's helper): FASTA / FASTQ parsing and the CSFM seed search in FILE ORDER (its rand() stream is the
	//   reference's for a given -S), -p threads inside findSeedsBatch;
	// workers: two per device, each with its own workspace: upload, the five stages, read-back, then the TSV text of the batch
	//   (aligned strings, placement columns) -- so the kernels of one batch overlap the host work of the next;
	// writer (this thread): the batches' text in file order.
	struct Job {
		long index;
		int n, maxLen;
		ReadBatchHost hb;
		string outText, chiText, alnText, errText;
		bool failed;
		Job() : index(0), n(0), maxLen(1), failed(false) {}
	};
	struct Queue {
		mutex m; condition_variable cv;
		deque<shared_ptr<Job> > q; bool closed; size_t cap;
		Queue() : closed(false), cap(0) {}
		void push(const shared_ptr<Job>& j) { unique_lock<mutex> l(m); cv.wait(l, [&] { return cap == 0 || q.size() < cap; }); q.push_back(j); cv.notify_all(); }
		shared_ptr<Job> pop() { unique_lock<mutex> l(m); cv.wait(l, [&] { return !q.empty() || closed; }); if (q.empty()) return shared_ptr<Job>(); shared_ptr<Job> j = q.front(); q.pop_front(); cv.notify_all(); return j; }
		void close() { unique_lock<mutex> l(m); closed = true; cv.notify_all(); }
	};
	const int workersPerGpu = 2, nWorkers = nGpus * workersPerGpu;
	Queue todo, done;
	todo.cap = (size_t) nWorkers + 2;
	const bool wantAln = alnFile.is_open(), wantChi = chiFile.is_open();
	auto work = [&](int wk) {
		hu_db* wdb = dbs[wk / workersPerGpu];
		hu_batch* batch = NULL;
		hu_batch* segBatch = NULL; // --chimera-info: (read, segment) pairs as the rows of a second workspace
		int batchMaxLen = 0;
		for (;;) {
			shared_ptr<Job> job = todo.pop();
			if (!job) break;
			Job& J = *job;
			ostringstream tOut, tChi, tAln, err;
			const int n = J.n;
			do {
#define HU_JOB(call) if ((call) != HU_OK) { err << "Error: " << hu_last_error() << endl; J.failed = true; break; }
				if (!batch || J.maxLen > batchMaxLen) {
					if (batch) hu_batch_destroy(batch);
					batch = NULL;
					batchMaxLen = max(J.maxLen + J.maxLen / 4, 64);
					HU_JOB(hu_batch_create(&batch, wdb, batchSize, batchMaxLen, paired ? 1 : 0));
				}
				const hu_read_batch rb = {n, paired ? 1 : 0, J.hb.codes.data(), J.hb.offsets.data(), J.hb.vpaths.data(), J.hb.nPaths.data()};
				HU_JOB(hu_batch_upload(batch, &rb));
				if (alignOnly) { HU_JOB(hu_align_batch(batch, mode)); }
				else { HU_JOB(hu_assign_batch(batch, mode, &sp)); }
				const int nSeq = n * (paired ? 2 : 1);
				vector<hu_align_row> rows(nSeq);
				vector<int64_t> traceOff(nSeq + 1);
				vector<char> traces((size_t) nSeq * (batchMaxLen + hmm.getProfileSize() + 8));
				HU_JOB(hu_batch_get_align(batch, rows.data(), traces.data(), (int64_t) traces.size(), traceOff.data()));
				vector<hu_place_row> best(n);
				if (!alignOnly) HU_JOB(hu_batch_get_best(batch, best.data()));
				// ---- --chimera-info (src/hmmufotu.cpp:653-691): segment placements over the common seeds ----
				vector<PTUnrooted::PTPlacement> seg5(n), seg3(n);
				vector<char> segDone(n, 0);
				if (chimeraInfo && !alignOnly) {
					vector<hu_align_row> merged(n);
					HU_JOB(hu_batch_get_merged(batch, merged.data()));
					vector<int8_t> codes((size_t) n * csLen);
					HU_JOB(hu_batch_get_aligned_codes(batch, codes.data()));
					vector<int32_t> nSeeds(n), ids((size_t) n * sp.maxNSeed);
					HU_JOB(hu_batch_get_seeds(batch, nSeeds.data(), ids.data(), NULL));
					const int nRows = n * numSeg;
					if (!segBatch) HU_JOB(hu_batch_create(&segBatch, wdb, batchSize * numSeg, batchMaxLen, 0));
					vector<int8_t> codes2((size_t) nRows * csLen);
					vector<int32_t> st(nRows), en(nRows), nSeeds2(nRows), ids2((size_t) nRows * HU_MAX_SEED, 0);
					for (int i = 0; i < n; ++i) {
						const bool ok = !(merged[i].flags & (HU_ALIGN_INVALID | HU_ALIGN_CHIMERA));
						const int segLen = (merged[i].csEnd - merged[i].csStart + 1) / numSeg;
						for (int sg = 0; sg < numSeg; ++sg) {
							const int q = i * numSeg + sg;
							memcpy(&codes2[(size_t) q * csLen], &codes[(size_t) i * csLen], csLen);
							const int segStart = merged[i].csStart + sg * segLen, segEnd = segStart + segLen - 1; // 1-based, :660-661
							st[q] = ok && segLen > 0 ? segStart - 1 : -1; en[q] = ok && segLen > 0 ? segEnd - 1 : -1;
							nSeeds2[q] = ok ? max(nSeeds[i], 0) : 0;
							for (int k = 0; k < nSeeds2[q]; ++k) ids2[(size_t) q * HU_MAX_SEED + k] = ids[(size_t) i * sp.maxNSeed + k];
						}
					}
					hu_sep_params sp2 = sp;
					sp2.maxNSeed = HU_MAX_SEED; // row pitch of the setters / getters below
					sp2.maxError = maxChimeraError;
					HU_JOB(hu_batch_set_aligned(segBatch, nRows, codes2.data(), st.data(), en.data()));
					HU_JOB(hu_seed_batch(segBatch, &sp2)); // sizes the seed rows of the workspace; replaced by the common seeds
					HU_JOB(hu_batch_set_seeds(segBatch, nSeeds2.data(), ids2.data()));
					HU_JOB(hu_estimate_batch(segBatch, &sp2));
					HU_JOB(hu_place_batch(segBatch, &sp2));
					vector<int32_t> nk(nRows);
					vector<hu_place_row> rows2((size_t) nRows * HU_MAX_SEED);
					HU_JOB(hu_batch_get_places(segBatch, nk.data(), rows2.data()));
					for (int i = 0; i < n; ++i) {
						vector<PTUnrooted::PTPlacement> p5, p3;
						for (int sg = 0; sg < numSeg; ++sg)
							for (int k = 0; k < nk[i * numSeg + sg]; ++k)
								(sg < numSeg / 2 ? p5 : p3).push_back(ptu.toPlacement(rows2[(size_t) (i * numSeg + sg) * HU_MAX_SEED + k]));
						if (p5.empty() || p3.empty()) continue;
						std::sort(p5.rbegin(), p5.rend(), compareByLoglik); // :674-675
						std::sort(p3.rbegin(), p3.rend(), compareByLoglik);
						seg5[i] = p5[0]; seg3[i] = p3[0]; segDone[i] = 1;
					}
				}
#undef HU_JOB
				for (int i = 0; i < n; ++i) {
					const PrimarySeq& f = J.hb.fwd[i];
					if (rows[i].flags & HU_ALIGN_INVALID) { err << "Warning: read '" << f.getId() << "' could not be aligned, skipped" << endl; continue; }
					BandedHMMP7::HmmAlignment al = hmm.buildGlobalAlign(f, rows[i], traces.data() + traceOff[i]);
					bool isChimera = false;
					if (paired) {
						if (rows[n + i].flags & HU_ALIGN_INVALID) { err << "Warning: mate of '" << f.getId() << "' could not be aligned, skipped" << endl; continue; }
						const BandedHMMP7::HmmAlignment revAln = hmm.buildGlobalAlign(J.hb.rev[i], rows[n + i], traces.data() + traceOff[n + i]);
						if (!ignoreOrient && !(al.csStart <= revAln.csStart && al.csEnd <= revAln.csEnd)) {
							err << "Warning: Bad orientation of forward/reverse read detected, treating as chimera" << endl;
							isChimera = true;
						} else
							al.merge(revAln);
					}
					PTUnrooted::PTPlacement bestPlace;
					bestPlace.taxonName = "NULL";
					if (!alignOnly && !isChimera && best[i].nPlaces > 0) bestPlace = ptu.toPlacement(best[i]);
					if (isChimera && !wantChi) continue; // chimera reads go to --chimera-out only, :693-706
					if (!isChimera && wantAln) {
						tAln << ">" << f.getId() << " " << f.getDesc() << ";csStart=" << al.csStart << ";csEnd=" << al.csEnd << ";" << endl;
						for (size_t o = 0; o < al.align.length(); o += 60) // SeqIO::writeFastaSeq, DEFAULT_MAX_LINE = 60 (src/SeqIO.cpp:121-132)
							tAln << al.align.substr(o, 60) << '\n';
					}
					ostream& dst = isChimera ? (ostream&) tChi : (ostream&) tOut;
					dst << f.getId() << "\t" << f.getDesc() << "\t" << al << "\t";
					if (chimeraInfo) {
						// chimera_lod = (best5 - alt5) + (best3 - alt3), where alt is the SAME window placed on the other segment's
						// branch: placeSeq returns the same constant for both, so the sum is exactly 0 (:676-689)
						const double chimeraLod = segDone[i] ? (seg5[i].loglik - seg5[i].loglik) + (seg3[i].loglik - seg3[i].loglik) : numeric_limits<double>::quiet_NaN();
						dst << seg5[i].getTaxonId() << "\t" << seg3[i].getTaxonId() << "\t" << (segDone[i] ? seg5[i].taxonName : string("NULL")) << "\t"
						    << (segDone[i] ? seg3[i].taxonName : string("NULL")) << "\t" << chimeraLod << "\t";
					}
					dst << bestPlace << endl;
				}
			} while (false);
			J.outText = tOut.str(); J.chiText = tChi.str(); J.alnText = tAln.str(); J.errText = err.str();
			vector<PrimarySeq>().swap(J.hb.fwd); vector<PrimarySeq>().swap(J.hb.rev);
			done.push(job);
		}
		if (batch) hu_batch_destroy(batch);
		if (segBatch) hu_batch_destroy(segBatch);
	};
	vector<thread> workers;
	for (int wk = 0; wk < nWorkers; ++wk) workers.push_back(thread(work, wk));
	atomic<long> nJobs(-1); // set by the reader when the input is exhausted
	thread reader([&] {
		long k = 0;
		while (fwdSeqI.hasNext() && (!paired || revSeqI.hasNext())) {
			shared_ptr<Job> job(new Job());
			job->index = k++;
			job->n = readBatch(fwdSeqI, revSeqI, paired, rStrand, batchSize, hmm, csfm, seedLen, seedRegion, mode, nThreads, job->hb, job->maxLen);
			todo.push(job);
		}
		nJobs = k;
		todo.close();
	});
	bool anyFailed = false;
	{
		map<long, shared_ptr<Job> > parked; // batches that finished ahead of their turn
		long next = 0, nDone = 0;
		for (;;) {
			const long total = nJobs.load();
			if (total >= 0 && next >= total) break;
			shared_ptr<Job> job;
			{
				unique_lock<mutex> l(done.m);
				done.cv.wait_for(l, chrono::milliseconds(50), [&] { return !done.q.empty(); });
				if (!done.q.empty()) { job = done.q.front(); done.q.pop_front(); }
			}
			if (job) parked[job->index] = job;
			while (parked.count(next)) {
				Job& J = *parked[next];
				cerr << J.errText;
				if (J.failed) anyFailed = true;
				out << J.outText;
				if (wantChi) chiFile << J.chiText;
				if (wantAln) alnFile << J.alnText;
				nDone += J.n;
				if (verbose) cerr << nDone << " reads processed" << endl;
				parked.erase(next++);
			}
		}
	}
	reader.join();
	for (size_t wk = 0; wk < workers.size(); ++wk) workers[wk].join();
	for (int g = 0; g < nGpus; ++g) hu_db_destroy(dbs[g]);
	return anyFailed ? EXIT_FAILURE : EXIT_SUCCESS;
}
