/*
 * oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the hmmufotu per-read hot path.
 *
 * This directory is a plain C++ fp64 restatement of the reference algorithm
 * (Grice-Lab/HmmUFOtu v1.5.1).  Nothing under oracle/ is shipped, linked into or
 * called by the product (hmmufotu_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF: the reference's own sources compile in place
 * (from /root/reference/src, nothing copied) against the stand-in Eigen/Boost headers in
 * oracle/shim/ (recipe: oracle/Makefile.ref -> oracle/_ref/{hmmufotu,hmmufotu-build,hmmufotu-sim}).
 * tests/golden/ref70/ holds databases written by that hmmufotu-build and assignment TSVs written by
 * that hmmufotu (tests/golden/make_ref_golden.py); tests/test_ref_pin.py recomputes every TSV
 * column with this restatement from the same files: all 448 rows are identical.  The reference's
 * own tests hold no numeric golden vectors for this path, so these are the pins,
 * next to the independent properties in tests/test_oracle.py (P(t) vs scipy.linalg.expm, detailed
 * balance, Felsenstein pruning vs brute force, banded == full Viterbi).
 *
 * Floating-point conventions chosen where Eigen's evaluation order is not
 * visible from the reference source (documented, not verifiable here):
 *   - Matrix4d row . Vector4d (strided row, scalar path) : (c0+c1)+(c2+c3)
 *   - Vector4d . Vector4d (packet path)                  : (c0+c2)+(c1+c3)
 *   - 4x4 product U*diag*U^-1                            : sequential k=0..3
 *   - maxCoeff/minCoeff                                  : first strict extremum
 *   - exp/log                                            : glibc
 */
#ifndef HU_ORACLE_H_
#define HU_ORACLE_H_

#include <cstdint>
#include <string>
#include <vector>
#include <limits>

namespace orc {

static const double INF = std::numeric_limits<double>::infinity();
static const double NEG_INF = -std::numeric_limits<double>::infinity();

/* ---------------------------------------------------------------- HMM ---- */
enum AlignMode { GLOBAL = 0, LOCAL = 1, NGCL = 2, CGNL = 3 }; /* src/BandedHMMP7.h:163-168 */
enum { ST_M = 0, ST_I = 1, ST_D = 2 };                        /* src/BandedHMMP7.h:157 */

struct Hmm {
	int K = 0;  /* profile size */
	int L = 0;  /* consensus length (MAXL) */
	/* costs, index [k*4+b] and [k*9+a*3+b]; k=0 is the begin state */
	std::vector<double> eM, eI, T;
	std::vector<double> entryPr, exitPr, entry, exit_; /* probs and costs, size K+1 */
	std::vector<int> cs2p, p2cs;                        /* size 65537 / 65537 as in the reference */
	double cNN = INF, cNB = INF, cEC = INF, cCC = INF;  /* T_SP_cost entries used by Viterbi */
	bool wingRetracted = false;
	std::string name;

	void setProfileSize(int K);
	/* after all costs are set: extend_index, adjustProfileLocalMode, wingRetract
	 * (src/BandedHMMP7.cpp:105-108, 701-705, 721-733, 1083-1120) */
	void finalize();
	void wingRetract();
	void setSequenceMode(int mode); /* src/BandedHMMP7.cpp:561-583 */
	int loadText(const std::string& path); /* src/BandedHMMP7.cpp:100-246 */
};

struct VPath { int start, end, from, to, nIns, nDel; }; /* src/BandedHMMP7.h:229-249 */

struct VScores { /* src/BandedHMMP7.h:193-227; column-major (row=read pos, col=profile pos) */
	int K, L;
	std::vector<double> M, I, D, S;
	VScores(int K, int L) : K(K), L(L) { reset(); }
	void reset();
	double& m(int i, int j) { return M[(size_t) j * (L + 1) + i]; }
	double& ii(int i, int j) { return I[(size_t) j * (L + 1) + i]; }
	double& d(int i, int j) { return D[(size_t) j * (L + 1) + i]; }
	double& s(int i, int j) { return S[(size_t) j * (L + 1) + i]; }
};

struct VTrace { /* src/BandedHMMP7.h:251-270 */
	double minScore = INF;
	int alnStart = 0, alnEnd = 0, alnFrom = 0, alnTo = 0;
	std::string trace;
};

struct HmmAlignment { /* src/BandedHMMP7.h:74-134 */
	int K = 0, L = 0;
	int seqStart = 0, seqEnd = 0, hmmStart = 0, hmmEnd = 0, csStart = 0, csEnd = 0;
	double cost = 0;
	std::string align;
	bool usedFallback = false;   /* oracle-side diagnostics, not in the reference */
	long bandCells = 0;          /* number of cell updates executed */
	std::string trace;
	bool isValid() const;
	HmmAlignment& merge(const HmmAlignment& other); /* src/BandedHMMP7.cpp:1188-1213 */
};

int8_t encodeBase(char c);                  /* src/DegenAlphabet.cpp:43-64, src/IUPACNucl.cpp:33-48 */
void calcViterbiFull(const Hmm&, const std::string& read, VScores&);                      /* src/BandedHMMP7.cpp:748-780 */
long calcViterbiBanded(const Hmm&, const std::string& read, VScores&, const std::vector<VPath>&); /* :782-892 */
void buildViterbiTrace(const Hmm&, VScores&, VTrace&);                                   /* :943-1006 */
HmmAlignment buildGlobalAlign(const Hmm&, const std::string& read, const VTrace&);        /* :1008-1081 */
/* alignSeq after the host-side seed search (src/HmmUFOtu_main.cpp:86-104) */
HmmAlignment alignWithPaths(const Hmm&, const std::string& read, const std::vector<VPath>&);
std::vector<int8_t> digitalSeq(const std::string& align); /* src/DigitalSeq.cpp:41-48 */

/* --------------------------------------------------------- subst models ---- */
enum ModelType { GTR = 0, TN93 = 1, HKY85 = 2, F81 = 3, K80 = 4, JC69 = 5 };

struct Model { /* src/DNASubModel.h, GTR/TN93/HKY85/F81/K80/JC69 .h/.cpp */
	int type = GTR;
	double pi[4] = {0.25, 0.25, 0.25, 0.25};
	double R[16] = {0};                /* GTR symmetric rates, row-major */
	double kr = 1, ky = 1, kappa = 1, beta = 1;
	/* derived (GTR) */
	double Q[16], lambda[4], U[16], U_1[16];
	void setup();                      /* GTR::setQfromParams (src/GTR.cpp:124-145) + K80 setBeta */
	void Pr(double v, double P[16]) const; /* row-major 4x4 */
	int loadText(std::istream& in);    /* each model's read() */
};

struct DGamma { int K = 0; double alpha = 0; std::vector<double> b, r; }; /* src/DiscreteGammaModel.cpp:57-72 */

/* ----------------------------------------------------------------- tree ---- */
struct Tree {
	long nNodes = 0;
	int csLen = 0;
	long root = 0;
	std::vector<long> parent;           /* -1 for root */
	std::vector<double> blen;           /* length of edge (id, parent[id]) */
	std::vector<double> height;
	std::vector<int8_t> seq;            /* nNodes x csLen */
	std::vector<std::string> anno;
	/* messages "before convolution", [node][col][4]:
	 * up[id] = msg(id -> parent), dn[id] = msg(parent -> id)   (src/PhyloTreeUnrooted.cpp:595-603) */
	std::vector<double> up, dn;
	Model model;
	DGamma dg;                          /* dg.K == 0 -> no discrete Gamma */
	/* run options of src/hmmufotu.cpp:229-241: -m weighted, --prior height, --ML */
	int estWeighted = 0, priorHeight = 0, onlyML = 0;
	/* sparse storage (bench.py's CPU baseline at gg_97 scale, where all 397 k messages are 98 GB): up/dn hold the rows of the
	 * nodes with slotOf[id] >= 0 only; empty slotOf = dense */
	std::vector<long> slotOf;
	bool hasMsg(long id) const { return slotOf.empty() || slotOf[id] >= 0; }
	const double* U(long id) const { return &up[(size_t) (slotOf.empty() ? id : slotOf[id]) * csLen * 4]; }
	const double* V(long id) const { return &dn[(size_t) (slotOf.empty() ? id : slotOf[id]) * csLen * 4]; }
	const int8_t* s(long id) const { return &seq[(size_t) id * csLen]; }
	/* all-roots Felsenstein evaluation (src/hmmufotu-build.cpp:455-459,
	 * src/PhyloTreeUnrooted.cpp:320-374): fills up/dn from leaf sequences */
	/* columns [j0, j1) only (column stripes are independent: one thread each); alloc: (re)size up/dn first */
	void evaluateAll(const std::vector<std::vector<long> >& children, int j0 = 0, int j1 = -1, bool alloc = true);
};

struct PTLoc { int start, end; long id; double dist; };      /* src/PhyloTreeUnrooted.h:390-405 */
inline bool operator<(const PTLoc& a, const PTLoc& b) { return a.dist < b.dist; }

struct Placement { /* src/PhyloTreeUnrooted.h:410-510 */
	int start = 0, end = 0;
	long c = -1, p = -1, a = -1;
	double wuv = 0, ratio = 0, wnr = 0, loglik = 0, height = 0, qPlace = 0, qTaxon = 0;
	double loglikTrue = 0;   /* intended root loglik; not in the reference */
	int outerIters = 0;      /* diagnostics */
};

double pDist(const int8_t* s1, const int8_t* s2, int start, int end, long* d = 0, long* N = 0); /* src/SeqUtils.cpp:37-54 */
/* seedOrder: 0 = libstdc++ std::sort (literal), 1 = stable (dist, id) */
std::vector<PTLoc> getSeed(const Tree&, const int8_t* seq, int start, int end, double maxDiff, double maxHeight, int seedOrder);
Placement estimateSeq(const Tree&, const int8_t* seq, const PTLoc& loc);           /* src/PhyloTreeUnrooted.cpp:849-877 */
void sortByLoglik(std::vector<Placement>& places);
void filterPlacements(std::vector<Placement>&, double maxError);                   /* src/HmmUFOtu_main.cpp:162-173 */
void placeSeq(const Tree&, const int8_t* seq, Placement&, double maxHeight);       /* src/PhyloTreeUnrooted.cpp:879-954 */
void calcQValues(const Tree&, std::vector<Placement>&);                            /* src/HmmUFOtu_main.cpp:182-216 */
void sortByQPlace(std::vector<Placement>&);                                        /* src/hmmufotu.cpp:730 */

/* primitives (src/PhyloTreeUnrooted.h:1488-1529) */
void dpsMat(const double P[16], const double V[4], double Y[4]);
double dpsVec(const double pi[4], const double V[4]);

} // namespace orc

#endif
