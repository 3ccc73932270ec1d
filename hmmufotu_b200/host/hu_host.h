// Host-side C++ layer above the C ABI: the reference's class surface for the per-read path, free of Eigen
// and Boost.  Names and method meanings follow /root/reference/src (file:line cited per declaration) so that
// the hmmufotu driver (hmmufotu.cpp in this directory) reads like the reference's.  No hot-path arithmetic
// lives here: Viterbi, seeding scan, estimate and place run on the GPU through include/hmmufotu_b200.h.
#ifndef HU_HOST_H_
#define HU_HOST_H_

#include <stdint.h>
#include <string.h>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "../../include/hmmufotu_b200.h"

namespace EGriceLab {
namespace HmmUFOtu {

extern const std::string progName;                 // "HmmUFOtu", src/util/ProgEnv.cpp
extern const int progVer[3];                       // 1.5.1
std::istream& loadProgInfo(std::istream& in);      // src/util/ProgEnv.cpp:30-65 (binary header check)
std::ostream& saveProgInfo(std::ostream& out);     // src/util/ProgEnv.cpp:24-28
std::ostream& writeProgInfo(std::ostream& out, const std::string& info); // text header, ProgEnv.cpp:101-104
bool loadString(std::string& dest, std::istream& in);                    // src/util/StringUtils.h:187-194
void saveString(const std::string& src, std::ostream& out);              // src/util/StringUtils.h:213-218

// ---- alphabet / sequences (src/DegenAlphabet.cpp:43-64, src/IUPACNucl.cpp:33-71, src/PrimarySeq.h) ----
struct DNA {
	static int8_t encode(char c);                  // ACGT -> 0..3, IUPAC degenerate -> first member, gap -2, else -1
	static char decode(int8_t b) { return "ACGT"[b & 3]; }
	static bool isSymbol(char c) { return encode(c) >= 0; }
	static bool isGap(char c) { return c == '-' || c == '.' || c == '_'; }
	static char complement(char c);                // IUPAC complement, case preserved
};

class PrimarySeq {
public:
	PrimarySeq() {}
	PrimarySeq(const std::string& id, const std::string& seq, const std::string& desc = "", const std::string& qual = "")
		: id(id), seq(seq), desc(desc), qual(qual) {}
	const std::string& getId() const { return id; }
	const std::string& getDesc() const { return desc; }
	const std::string& getSeq() const { return seq; }
	size_t length() const { return seq.length(); }
	PrimarySeq revcom() const;                     // src/PrimarySeq.cpp
	std::string id, seq, desc, qual;
};

class SeqIO {                                      // src/SeqIO.cpp: streaming FASTA / FASTQ reader
public:
	SeqIO(std::istream* in, const std::string& fmt) : in(in), fmt(fmt) {}
	bool hasNext();
	PrimarySeq nextSeq();
	static std::string guessFormat(const std::string& fn);
private:
	std::istream* in;
	std::string fmt;
};

// ---- CSLoc + CSFM index (src/CSLoc.h, src/CSFMIndex.{h,cpp}) ----
struct CSLoc {
	CSLoc() : start(0), end(0) {}
	CSLoc(int start, int end, const std::string& CS) : start(start), end(end), CS(CS) {}
	bool isValid() const { return start > 0 && start <= end && !CS.empty(); }
	int start, end;                                // 1-based consensus columns of the first / last pattern char
	std::string CS;                                // the pattern with '-' for every skipped column
};

struct MSA {                                       // src/MSA.{h,cpp}: only what the run path touches
	std::string alphabet, name, CS;
	unsigned numSeq, csLen;
	bool isPruned;
	std::vector<std::string> seqNames;
	std::string concatMSA;                         // numSeq * csLen chars, row-major
	std::string auxBlocks;                         // start/end/len, counts and weights exactly as loaded (written back by save)
	std::istream& load(std::istream& in);          // src/MSA.cpp:362-430 (whole record is parsed)
	std::ostream& save(std::ostream& out) const;   // src/MSA.cpp:295-360 (built: the builder's counts and weights; loaded: as read)
	// database builder, host half (hu_build.cpp): what hmmufotu-build does to the alignment before anything else
	long loadFasta(std::istream& in);              // src/MSA.cpp:140-164 (+ counts, position-based weights, consensus)
	MSA& prune();                                  // src/MSA.cpp:87-138: columns without a residue are dropped
	void updateCounts();                           // src/MSA.cpp:226-293: raw counts, sequence weights, weighted counts
	void calculateCS();                            // src/MSA.cpp:211-226
	std::vector<int> startIdx, endIdx, lenIdx, resCount, gapCount; // resCount[j * 4 + b], as Eigen's column-major 4 x csLen
	std::vector<double> seqWeight, resWCount, gapWCount;
	unsigned getCSLen() const { return csLen; }
	char residualAt(unsigned i, unsigned j) const { return concatMSA[(size_t) i * csLen + j]; }
};

// FM index over the gap-stripped concatenation of the MSA rows with a map back to consensus columns.
// The reference serialises libcds structures into <db>.csfm; this class rebuilds the same index from <db>.msa
// at start-up (the suffix array of a text with a unique terminator is unique, so SA intervals and SA values,
// hence locateOne()'s rand() pick, are the reference's).
class CSFMIndex {
public:
	CSFMIndex() : csLen(0), concatLen(0), qlen(0) {}
	CSFMIndex& build(const MSA& msa);              // src/CSFMIndex.cpp:232-368
	int getCSLen() const { return csLen; }
	int32_t getConcatLen() const { return concatLen; }
	int32_t count(const std::string& pattern) const;                 // src/CSFMIndex.cpp:79-99
	CSLoc locateOne(const std::string& pattern) const;               // src/CSFMIndex.cpp:121-147 (uses rand())
	CSLoc locateFirst(const std::string& pattern) const;             // src/CSFMIndex.cpp:101-119
	std::vector<CSLoc> locate(const std::string& pattern) const;     // src/CSFMIndex.cpp:53-77
	// locateOne() in two halves, so that the backward searches of a batch of reads can run on many threads while the
	// rand() picks stay in file order: the SA interval of a pattern (no random draw) ...
	bool interval(const std::string& pattern, int32_t& start, int32_t& end) const { return search(pattern, start, end); }
	bool interval(const char* pattern, size_t len, int32_t& start, int32_t& end) const { return search(pattern, len, start, end); } // no temporary string (the batch search calls this ~30 times per read)
	// ... and the random hit inside a non-empty interval (the rand() call and what follows it in locateOne)
	CSLoc locateAt(const std::string& pattern, int32_t start, int32_t end) const;
	// <db>.csfm as the reference's hmmufotu loads it (src/CSFMIndex.cpp:176-230): the index in libcds' serialisation --
	// sampled suffix array + RRR bitmap of the sampled positions + pointerless wavelet tree over the BWT with RRR-compressed
	// levels (hu_build.cpp).  msa: the alignment this index was built from.
	std::ostream& save(std::ostream& out, const MSA& msa) const;
	// ... and read back (src/CSFMIndex.cpp:198-230): the RRR bitmaps and the wavelet tree are decoded into this class's own
	// rank blocks, and the full suffix array is recovered from the BWT by LF steps (seconds where a suffix sort of a large
	// alignment takes minutes).  The result is the index build() makes from the same alignment, field for field.
	std::istream& load(std::istream& in);
	bool operator==(const CSFMIndex& o) const;
private:
	bool search(const std::string& pattern, int32_t& start, int32_t& end) const { return search(pattern.data(), pattern.length(), start, end); }
	bool search(const char* pattern, size_t len, int32_t& start, int32_t& end) const;
	int32_t occ(int b, int32_t i) const;           // # of symbol b in BWT[0..i]
	std::string extractCS(int32_t start, const std::string& pattern) const; // src/CSFMIndex.cpp:262-273
	int csLen;
	int32_t concatLen;
	int32_t C[6];
	struct RankBlock {                             // 64 BWT positions in one cache line: counts before the block, one bit mask per symbol
		uint32_t cnt[5], pad;
		uint64_t mask[5];
	};
	std::vector<RankBlock> rank;                   // BWT symbols 0..4 (0 = separator / terminator)
	int qlen;                                      // the backward search of every qlen-mer, tabulated (0: no table)
	std::vector<int32_t> qtab;                     // [4^qlen][2]: SA interval of the qlen-mer whose LAST symbol is the most significant digit; start > end: none
	void buildQTable();                            // called when the rank blocks are made (build(), load())
	void fillQ(int depth, size_t idx, int32_t start, int32_t end);
	std::vector<int32_t> sa;                       // full suffix array (1-based text positions as in the reference)
	std::vector<uint16_t> concat2CS;               // 1-based text position -> 1-based consensus column (0: separator)
};

// ---- banded plan-7 profile (src/BandedHMMP7.{h,cpp}) ----
class BandedHMMP7 {
public:
	enum align_mode { GLOBAL, LOCAL, NGCL, CGNL };                   // .h:163-168
	struct ViterbiAlignPath {                                        // .h:229-249
		ViterbiAlignPath() : start(0), end(0), from(0), to(0), nIns(0), nDel(0) {}
		ViterbiAlignPath(int start, int end, int from, int to, int nIns, int nDel)
			: start(start), end(end), from(from), to(to), nIns(nIns), nDel(nDel) {}
		bool isValid() const { return start > 0 && start <= end && from > 0 && from <= to && nIns >= 0 && nDel >= 0; }
		int start, end, from, to, nIns, nDel;
	};
	struct HmmAlignment {                                            // .h:74-134
		HmmAlignment() : K(0), L(0), seqStart(0), seqEnd(0), hmmStart(0), hmmEnd(0), csStart(0), csEnd(0), cost(0) {}
		bool isValid() const;
		HmmAlignment& merge(const HmmAlignment& other);              // cpp:1188-1213
		int K, L, seqStart, seqEnd, hmmStart, hmmEnd, csStart, csEnd;
		double cost;
		std::string align;
		static const std::string TSV_HEADER;
	};
	// ViterbiScores (.h:173-199) holds the reference's four dense DP matrices; here it is the handle of one device alignment:
	// calcViterbiScores() runs the banded (or full) Viterbi pass of ONE read on the attached device and keeps its result,
	// buildViterbiTrace() hands it out as the reference's ViterbiAlignTrace (.h:251-270)
	struct ViterbiScores {
		ViterbiScores() : seq(NULL) { memset(&row, 0, sizeof row); }
		explicit ViterbiScores(const PrimarySeq& s) : seq(&s) { memset(&row, 0, sizeof row); }
		const PrimarySeq* seq;
		hu_align_row row;
		std::string trace;
	};
	struct ViterbiAlignTrace {
		ViterbiAlignTrace() : minScore(0), alnStart(0), alnEnd(0), alnFrom(0), alnTo(0) {}
		double minScore;
		int alnStart, alnEnd, alnFrom, alnTo;
		std::string alnTrace;
	};
	ViterbiScores& calcViterbiScores(const PrimarySeq& seq, ViterbiScores& vs) const;                                              // cpp:748-780
	ViterbiScores& calcViterbiScores(const PrimarySeq& seq, ViterbiScores& vs, const std::vector<ViterbiAlignPath>& vpaths) const; // cpp:782-892
	void buildViterbiTrace(const ViterbiScores& vs, ViterbiAlignTrace& vtrace) const;                                              // cpp:943-1006
	HmmAlignment buildGlobalAlign(const PrimarySeq& seq, const ViterbiScores& vs, const ViterbiAlignTrace& vtrace) const;           // cpp:1008-1081
	mutable class DeviceDB* dev;                   // set by DeviceDB: where the per-call facades run
	BandedHMMP7() : dev(NULL), K(0), L(0), wingRetracted(false), mode(GLOBAL) {}
	int getProfileSize() const { return K; }
	int getCSLen() const { return L; }
	int getProfileLoc(int cs) const { return cs2ProfileIdx[cs]; }
	int getCSLoc(int k) const { return profile2CSIdx[k]; }
	void setSequenceMode(align_mode m) { mode = m; }                  // cpp:561-583 (costs derived on the device side)
	align_mode getSequenceMode() const { return mode; }
	void wingRetract();                                              // cpp:1083-1120 (via hu_hmm_finalize)
	ViterbiAlignPath buildAlignPath(const CSLoc& csLoc, int csFrom, int csTo) const; // cpp:894-941
	HmmAlignment buildGlobalAlign(const PrimarySeq& seq, const hu_align_row& row, const char* trace) const; // cpp:1008-1081
	hu_hmm_params params() const;                                    // view for hu_db_create / hu_trace_to_align
	// database builder (hu_build.cpp): profile training from the pruned MSA, src/BandedHMMP7.cpp:386-541 -- match columns by
	// weighted symbol fraction, weighted emission / transition counts, effective sequence number by bisection to a mean
	// relative entropy of 1 bit, posterior means under the Dirichlet priors of a .dm file.  The reference's transition loop reads
	// a variable its inner loop shadows (:446-463, undefined behaviour).  asCompiled = true reproduces what the reference's
	// GCC -O2 build does with it (oracle/_ref, read off its disassembly: the never-assigned variable is taken as M, so every
	// counted transition goes to a match state) and with it the reference builder's <db>.hmm, every line but DATE;
	// asCompiled = false counts the state of the next non-gap position, which is what the loop evidently means.
	bool build(const MSA& msa, double symfrac, std::istream& priorIn, bool asCompiled = true);
	friend std::istream& operator>>(std::istream& in, BandedHMMP7& hmm);        // cpp:100-246
	friend std::ostream& operator<<(std::ostream& out, const BandedHMMP7& hmm); // cpp:324-378
	std::string name;
	int K, L;
	std::vector<double> E_M_cost, E_I_cost, Tmat_cost, entryPr_cost, exitPr_cost; // [k*4+b], [k*9+a*3+b], [k]
	std::vector<int32_t> cs2ProfileIdx, profile2CSIdx;
	std::vector<std::pair<std::string, std::string> > optTags;
	std::vector<std::string> locMap;                                 // per-position text after the MAP column (CONS ...)
	bool wingRetracted;
	align_mode mode;
};
std::ostream& operator<<(std::ostream& out, const BandedHMMP7::HmmAlignment& aln); // cpp:1215-1221

// ---- substitution models (src/DNASubModel.h, GTR/TN93/HKY85/F81/K80/JC69, DNASubModelFactory) ----
class DNASubModel {
public:
	DNASubModel();
	std::string modelType() const;
	void Pr(double v, double* P16) const;          // row-major 4x4 through hu_model_Pr
	const double* getPi() const { return desc.pi; }
	std::istream& read(std::istream& in);          // each model's read()
	std::ostream& write(std::ostream& out) const;  // each model's write()
	hu_model_desc desc;
};
std::istream& operator>>(std::istream& in, DNASubModel& m);
std::ostream& operator<<(std::ostream& out, const DNASubModel& m);

struct DiscreteGammaModel {                        // src/DiscreteGammaModel.{h,cpp}
	int K;
	double alpha;
	std::vector<double> b, r;
	double rate(int k) const { return r[k]; }
	std::istream& load(std::istream& in);          // cpp:57-72
	std::ostream& save(std::ostream& out) const;   // cpp:74-90
	// builder side (hu_build.cpp): break points and rates from (K, alpha), literally as cpp:40-55 -- quantiles of
	// Gamma(shape alpha, SCALE alpha) and r_i = P(alpha+1, b_{i+1} alpha) - P(alpha+1, b_i alpha), not multiplied by K
	DiscreteGammaModel& build(int K, double alpha);
	static double estimateShape(const std::vector<double>& numMut); // cpp:92-98: m^2 / (s^2 - m), inf for < 2 sites
};
double gammaP(double a, double x);                 // regularised lower incomplete gamma function P(a, x)
double gammaQuantile(double shape, double scale, double p);

// ---- unrooted phylogenetic tree as stored in <db>.ptu (src/PhyloTreeUnrooted.{h,cpp}) ----
typedef std::vector<int8_t> DigitalSeq;            // src/DigitalSeq.h: an aligned read as codes over the consensus columns (gap < 0)
DigitalSeq digitalSeq(const std::string& align);   // src/DigitalSeq.cpp:41-48

class PTUnrooted {
public:
	enum PRIOR_TYPE { UNIFORM, HEIGHT };
	struct PTLoc { int start, end; long id; double dist; };          // .h:390-405
	struct PTPlacement {                                             // .h:410-510
		PTPlacement();
		std::string getId() const;
		long getTaxonId() const { return aNode; }
		int start, end;
		long cNode, pNode, aNode;
		double wuv, ratio, wnr, loglik, height, annoDist, qPlace, qTaxon;
		std::string taxonName;
		// what the device row holds beyond the reference's fields (p-distance of the seed, estimated log-likelihood, taxon id,
		// iteration counts): kept so that a placement can travel back to the device through the per-call facades below
		hu_place_row raw;
		const PTUnrooted* owner;
		hu_place_row toRow() const;                 // raw with the visible fields written over it
		static const std::string TSV_HEADER;
	};
	PTUnrooted() : dev(NULL), csLen(0), root(0), hasDG(false) {}
	// class surface of the per-read path (src/PhyloTreeUnrooted.h:870-900), one placement per call, on the attached device
	PTPlacement estimateSeq(const DigitalSeq& seq, const PTLoc& loc, const std::string& method = "unweighted") const; // cpp:849-877
	PTPlacement& placeSeq(const DigitalSeq& seq, PTPlacement& place, double maxHeight) const;                            // cpp:925-954
	mutable class DeviceDB* dev;                   // set by DeviceDB: where the per-call facades run
	// cpp:496-535.  keepMessages = false reads everything but the messages (98 GB for a gg_97-sized tree): they are then
	// streamed from the file to the device(s), edge by edge, by streamMessages()
	std::istream& load(std::istream& in, bool keepMessages = true);
	struct MessageSink {
		virtual ~MessageSink() {}
		virtual bool put(long node, int dir, const double* logMsg) = 0; // dir 0: msg(node -> parent), 1: msg(parent -> node)
		virtual bool flush() { return true; }
	};
	bool streamMessages(std::istream& in, MessageSink& sink) const;
	std::streampos edgePos;                        // where the edge records start in the file load() read
	std::ostream& save(std::ostream& out) const;   // cpp:537-567
	size_t numNodes() const { return parent.size(); }
	int getCSLen() const { return csLen; }
	double getHeight(long id) const { return height[id]; }
	const std::string& getAnno(long id) const { return anno[id]; }
	hu_tree_desc treeDesc() const;                 // view for hu_db_create (messages included)
	// builder side (hu_build.cpp), from the evaluated messages up[] = msg(node -> parent):
	void updateRootLoglik();                       // cpp:298-301 with loglik(root, j), cpp:320-346
	std::vector<double> numMutations() const;      // cpp:1008-1016 per site, as hmmufotu-build feeds estimateShape (:431-436)
	hu_model_desc modelDesc() const;
	PTPlacement toPlacement(const hu_place_row& row) const;
	int csLen;
	long root;
	std::vector<int32_t> parent;
	std::vector<double> blen, height, annoDist;
	std::vector<std::string> name, anno;
	std::vector<std::string> seqName;              // DigitalSeq name of a node (the MSA row id for leaves, empty otherwise)
	std::vector<int8_t> seq;                       // numNodes x csLen
	std::vector<int32_t> taxon;                    // id of the annotation string (equal strings share an id)
	std::vector<double> up, dn;                    // [node][col][4]
	std::vector<std::vector<long> > neighbors;     // file order, for save()
	std::vector<double> rootLoglik;
	std::vector<std::pair<unsigned, long> > msaIndex;
	DNASubModel model;
	bool hasDG;
	DiscreteGammaModel dG;
};
std::ostream& operator<<(std::ostream& out, const PTUnrooted::PTPlacement& place); // .h:1611-1617

// ---- database builder, host half (src/hmmufotu-build.cpp:308-392; hu_build.cpp) ----
struct NewickTree {                                // src/NewickTree.h: name, branch length to the parent, children in file order
	NewickTree() : length(0) {}
	std::string name;
	double length;
	std::vector<NewickTree> children;
	bool read(std::istream& in);                   // the grammar of src/NewickTree.h:172-236
};
// PTUnrooted(NewickTree) + loadMSA + fixBranchLength + loadAnnotation + formatName + annotate + calcNodeHeight
// (src/PhyloTreeUnrooted.cpp:131-296, 974-1006): everything of <db>.ptu that does not need the likelihood evaluation
// (node ids, edges in file order, branch lengths, leaf sequences, names, annotations, heights).  Messages and the
// inferred internal sequences come from hu_db_create() with hu_tree_desc.up == NULL.
bool buildTree(PTUnrooted& tree, const NewickTree& nt, const MSA& msa, std::istream* annoIn,
		const std::string& rootName = "cellular_organisms");

// ---- the per-read pipeline, batched (src/HmmUFOtu_main.{h,cpp}) ----
// Host half of alignSeq (src/HmmUFOtu_main.cpp:39-84): the CSFM seed search, in call order (it draws rand()).
std::vector<BandedHMMP7::ViterbiAlignPath> findSeeds(const BandedHMMP7& hmm, const CSFMIndex& csfm, const PrimarySeq& read,
		int seedLen, int seedRegion, BandedHMMP7::align_mode mode);
// The same for a batch of reads, in the given order, with the backward searches on nThreads threads: phase 1 (parallel)
// finds, per read, the SA intervals of the seed attempts up to the first non-empty one at each end; phase 2 (sequential,
// file order) replays findSeeds' control flow on those intervals and draws rand() exactly as the one-by-one loop does
// (an attempt whose hit turns out unusable continues with on-the-spot searches).  Same paths, same rand() stream.
std::vector<std::vector<BandedHMMP7::ViterbiAlignPath> > findSeedsBatch(const BandedHMMP7& hmm, const CSFMIndex& csfm,
		const std::vector<const PrimarySeq*>& reads, int seedLen, int seedRegion, BandedHMMP7::align_mode mode, int nThreads);

// ---- the reference's free functions (src/HmmUFOtu_main.h:70-113), one read per call, over the C ABI (hu_facade.cpp) ----
// One database replica on one device plus a one-read workspace; attaching it to the host objects is what makes their
// per-call methods and the free functions below run: every stage goes through include/hmmufotu_b200.h (no host arithmetic).
class DeviceDB {
public:
	DeviceDB(const BandedHMMP7& hmm, const PTUnrooted& ptu, int device); // throws std::runtime_error with hu_last_error()
	~DeviceDB();
	hu_db* db;
	hu_batch* one(int maxLen);                     // the one-read workspace, grown on demand
	const BandedHMMP7& hmm;
	const PTUnrooted& ptu;
private:
	hu_batch* batch;
	int batchLen;
	DeviceDB(const DeviceDB&);
	DeviceDB& operator=(const DeviceDB&);
};
BandedHMMP7::HmmAlignment alignSeq(const BandedHMMP7& hmm, const CSFMIndex& csfm, const PrimarySeq& read, int seedLen, int seedRegion,
		BandedHMMP7::align_mode mode);                                                                                    // cpp:39-105
// all the device keeps of the sorted candidate vector is its first HU_MAX_SEED (64) entries: callers truncate to -N <= 64
std::vector<PTUnrooted::PTLoc> getSeed(const PTUnrooted& ptu, const DigitalSeq& seq, int start, int end, double maxDiff, double maxHeight); // cpp:127-152
std::vector<PTUnrooted::PTPlacement> estimateSeq(const PTUnrooted& ptu, const DigitalSeq& seq, const std::vector<PTUnrooted::PTLoc>& locs,
		const std::string& method);                                                                                       // cpp:154-160
std::vector<PTUnrooted::PTPlacement>& filterPlacements(std::vector<PTUnrooted::PTPlacement>& places, double maxError);            // cpp:162-173
std::vector<PTUnrooted::PTPlacement>& placeSeq(const PTUnrooted& ptu, const DigitalSeq& seq, std::vector<PTUnrooted::PTPlacement>& places,
		double maxHeight);                                                                                                // cpp:175-180
void calcQValues(std::vector<PTUnrooted::PTPlacement>& places, PTUnrooted::PRIOR_TYPE type);                                      // cpp:182-216
bool compareByLoglik(const PTUnrooted::PTPlacement& lhs, const PTUnrooted::PTPlacement& rhs);                                     // .h:1627-1637
bool compareByQPlace(const PTUnrooted::PTPlacement& lhs, const PTUnrooted::PTPlacement& rhs);

} // namespace HmmUFOtu
} // namespace EGriceLab

#endif
