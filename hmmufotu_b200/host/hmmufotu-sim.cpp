// hmmufotu-sim over this repository's host layer (the reference's src/hmmufotu-sim.cpp:313-424): draws
// reads from a database's tree model -- a branch uniformly over the non-root nodes, a branching point U(0,1) on it, a
// consensus window (uniform start, N(mean, sd) columns, or a row of a BED file), per column a gap with the column's weighted
// gap fraction, otherwise a base from the normalised likelihood of the branching point,
//   exp( log(P(v rc) e^{msg(c->p)_j}) + log(P(v (1-rc)) e^{msg(p->c)_j}) ).
// Host-side arithmetic only (4x4 matrices through hu_model_Pr), no device.  The reference draws from Boost's mt11213b and
// Boost's distributions; this program uses std::mt19937_64 and the standard ones, so parity is statistical, not bit-wise
// (tests/test_host.py checks both programs against the model's own probabilities).
//
//   hmmufotu-sim DB SEQ-OUT [MATE-OUT] -N NUM [-k] [-d MAXDIST] [-m MEAN] [-s SD] [-l MIN] [-u MAX] [-r READLEN]
//                [-R REGIONS.bed] [--prefix STR] [-S SEED]
#include "hu_host.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>

using namespace std;
using namespace EGriceLab::HmmUFOtu;

static void usage(const char* prog) {
	cerr << "Usage:    " << prog << "  <HmmUFOtu-DB> <SEQ-OUT> [MATE-OUT] <-N NUM-READS> [options]" << endl
		 << "Options:    -N  LONG            : number of reads/pairs to generate" << endl
		 << "            -k|--keep-gap FLAG  : keep simulated gaps, so that the reads come out aligned (no MATE-OUT then)" << endl
		 << "            -d|--max-dist  DBL  : largest height of a branching point [inf]" << endl
		 << "            -m|--mean-size  DBL : mean amplicon size in consensus columns [500]" << endl
		 << "            -s|--sd-size  DBL   : standard deviation of the amplicon size [30]" << endl
		 << "            -l|--min-size  DBL  : smallest amplicon size, 0 for no limit [0]" << endl
		 << "            -u|--max-size  DBL  : largest amplicon size, 0 for no limit [0]" << endl
		 << "            -r|--read-len  INT  : read length, -1 for the whole amplicon [-1]" << endl
		 << "            -R|--region  FILE   : BED file of consensus regions to draw from (replaces -m -s -l -u)" << endl
		 << "            --prefix  STR       : prefix of the read ids [r]" << endl
		 << "            -S|--seed  INT      : random seed" << endl;
}

// log(X exp(V)) per row with the reference's underflow guard (src/PhyloTreeUnrooted.h:1495-1503)
static void dotProductScaled(const double* P16, const double* V, double* Y) {
	const double minExp = -1021 / 2;
	const double maxV = max(max(V[0], V[1]), max(V[2], V[3]));
	const double scale = maxV != -INFINITY && maxV < minExp ? minExp - maxV : 0;
	double e[4];
	for (int k = 0; k < 4; ++k) e[k] = ::exp(V[k] + scale);
	for (int i = 0; i < 4; ++i) {
		double d = 0;
		for (int k = 0; k < 4; ++k) d += P16[i * 4 + k] * e[k];
		Y[i] = ::log(d) - scale;
	}
}

static string fmtDouble(double v) { // boost::lexical_cast<string>(double): 17 significant digits
	char buf[64];
	snprintf(buf, sizeof buf, "%.17g", v);
	return buf;
}

int main(int argc, char* argv[]) {
	vector<string> pos;
	map<string, string> kv;
	for (int i = 1; i < argc; ++i) {
		const string a = argv[i];
		if (a.size() > 1 && a[0] == '-' && !isdigit((unsigned char) a[1])) {
			const bool flag = a == "-k" || a == "--keep-gap" || a == "-h" || a == "--help" || a == "-v";
			if (flag || i + 1 >= argc) kv[a] = "";
			else kv[a] = argv[++i];
		} else
			pos.push_back(a);
	}
	struct { const map<string, string>& kv; bool has(const string& a, const string& b) const { return kv.count(a) || kv.count(b); }
	         string get(const string& a, const string& b, const string& d) const { return kv.count(a) ? kv.find(a)->second : kv.count(b) ? kv.find(b)->second : d; } } o = {kv};
	if (o.has("-h", "--help") || pos.size() < 2 || pos.size() > 3 || !kv.count("-N")) { usage(argv[0]); return pos.size() < 2 ? EXIT_FAILURE : EXIT_SUCCESS; }
	const string db = pos[0], outFn = pos[1];
	const bool keepGap = o.has("-k", "--keep-gap");
	const string mateFn = pos.size() == 3 && !keepGap ? pos[2] : "";
	const long N = atol(kv["-N"].c_str());
	const double maxDist = o.has("-d", "--max-dist") ? atof(o.get("-d", "--max-dist", "").c_str()) : INFINITY;
	const double meanSize = atof(o.get("-m", "--mean-size", "500").c_str()), sdSize = atof(o.get("-s", "--sd-size", "30").c_str());
	const double minSize = atof(o.get("-l", "--min-size", "0").c_str()), maxSize = atof(o.get("-u", "--max-size", "0").c_str());
	const int readLen = atoi(o.get("-r", "--read-len", "-1").c_str());
	const string prefix = o.get("--prefix", "--prefix", "r");
	const unsigned long seed = o.has("-S", "--seed") ? strtoul(o.get("-S", "--seed", "0").c_str(), NULL, 10) : (unsigned long) time(NULL);
	if (N <= 0) { cerr << "-N must be positive" << endl; return EXIT_FAILURE; }
	if (!(sdSize >= 0) || !(meanSize > 0)) { cerr << "-m must be positive and -s non-negative" << endl; return EXIT_FAILURE; }

	// ---- database ----
	ifstream msaIn((db + ".msa").c_str(), ios_base::binary), ptuIn((db + ".ptu").c_str(), ios_base::binary);
	if (!msaIn.is_open() || !ptuIn.is_open()) { cerr << "Unable to open the database files of '" << db << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	MSA msa;
	PTUnrooted ptu;
	if (loadProgInfo(msaIn).bad() || !msa.load(msaIn)) { cerr << "Failed to load MSA data '" << db << ".msa'" << endl; return EXIT_FAILURE; }
	if (loadProgInfo(ptuIn).bad() || !ptu.load(ptuIn)) { cerr << "Unable to load Phylogenetic tree data '" << db << ".ptu'" << endl; return EXIT_FAILURE; }
	const int csLen = ptu.getCSLen();
	if ((int) msa.csLen != csLen) { cerr << "Error: Unmatched CS length between the database files" << endl; return EXIT_FAILURE; }
	msa.updateCounts(); // weighted gap fractions per column (the file carries the same numbers in its auxiliary blocks)
	vector<double> gapWFrac(csLen);
	for (int j = 0; j < csLen; ++j) {
		double res = 0;
		for (int b = 0; b < 4; ++b) res += msa.resWCount[(size_t) j * 4 + b];
		gapWFrac[j] = msa.gapWCount[j] / (res + msa.gapWCount[j]);
	}
	vector<pair<int, int> > loci; // 0-based inclusive windows from a BED file (start, end]
	if (o.has("-R", "--region")) {
		ifstream bed(o.get("-R", "--region", "").c_str());
		if (!bed.is_open()) { cerr << "Unable to open '" << o.get("-R", "--region", "") << "'" << endl; return EXIT_FAILURE; }
		string line;
		while (getline(bed, line)) {
			if (line.empty() || line[0] == '#') continue;
			istringstream ls(line);
			string chr;
			int s, e;
			if (!(ls >> chr >> s >> e)) continue;
			if (!(0 <= s && s < e && e <= csLen)) { cerr << "Region (" << s << "," << e << "] is not in the consensus range, ignored" << endl; continue; }
			loci.push_back(make_pair(s + 1, e)); // as the reference stores them: CSLoc(start + 1, end)
		}
	}
	ofstream seqOut(outFn.c_str()), mateOut;
	if (!seqOut.is_open()) { cerr << "Unable to write to '" << outFn << "': " << ::strerror(errno) << endl; return EXIT_FAILURE; }
	if (!mateFn.empty()) { mateOut.open(mateFn.c_str()); if (!mateOut.is_open()) { cerr << "Unable to write to '" << mateFn << "'" << endl; return EXIT_FAILURE; } }

	// ---- draws ----
	mt19937_64 rng(seed);
	uniform_real_distribution<double> u01(0.0, 1.0);
	const size_t numNodes = ptu.numNodes();
	vector<double> nodePr(numNodes, 1.0);
	if (maxDist != INFINITY)
		for (size_t i = 0; i < numNodes; ++i) if (ptu.getHeight((long) i) > maxDist) nodePr[i] = 0;
	discrete_distribution<size_t> nodeDist(nodePr.begin(), nodePr.end());
	uniform_int_distribution<int> locDist(0, loci.empty() ? csLen - 1 : (int) loci.size() - 1);
	normal_distribution<double> sizeDist(meanSize, sdSize > 0 ? sdSize : 1e-300);
	const size_t per = (size_t) csLen * 4;
	double P1[16], P2[16];
	for (long n = 1; n <= N;) {
		const size_t id = nodeDist(rng);
		if ((long) id == ptu.root) continue; // no parent branch
		const long pid = ptu.parent[id];
		const double v = ptu.blen[id], rc = u01(rng);
		if (ptu.getHeight((long) id) + v * rc > maxDist) continue;
		int start, end;
		if (loci.empty()) {
			start = locDist(rng);
			double len = sizeDist(rng);
			if (len < minSize) len = minSize;
			if (maxSize > 0 && len > maxSize) len = maxSize;
			end = start + (int) len;
			if (!(end < csLen)) continue;
		} else {
			const pair<int, int>& l = loci[locDist(rng)];
			start = l.first; end = l.second;
			if (!(end < csLen)) end = csLen - 1; // the reference indexes past the last column here; stay inside
		}
		ptu.model.Pr(v * rc, P1);
		ptu.model.Pr(v * (1 - rc), P2);
		string seq;
		if (keepGap) seq.append(start, '.');
		for (int j = start; j <= end; ++j) {
			if (u01(rng) <= gapWFrac[j]) { if (keepGap) seq.push_back('-'); continue; }
			double a[4], b[4], w[4], mx = -INFINITY;
			dotProductScaled(P1, &ptu.up[id * per + (size_t) j * 4], a);
			dotProductScaled(P2, &ptu.dn[id * per + (size_t) j * 4], b);
			for (int k = 0; k < 4; ++k) { w[k] = a[k] + b[k]; mx = max(mx, w[k]); }
			double tot = 0;
			for (int k = 0; k < 4; ++k) { w[k] = ::exp(w[k] - mx); tot += w[k]; }
			double x = u01(rng) * tot;
			int base = 0;
			while (base < 3 && x >= w[base]) { x -= w[base]; base++; }
			seq.push_back(DNA::decode((int8_t) base));
		}
		if (keepGap) seq.append(csLen - 1 - end, '.');
		ostringstream desc;
		desc << "branchID=" << id << "->" << pid << ";taxonID=" << (rc <= 0.5 ? (long) id : pid) << ";taxonName=\"" << ptu.getAnno(rc <= 0.5 ? (long) id : pid)
		     << "\";branchPoint=" << fmtDouble(rc) << ";csStart=" << start << ";csEnd=" << end << ";seqLen=" << seq.length() << ";";
		ostringstream rid;
		rid << prefix << n;
		const PrimarySeq insert(rid.str(), seq, desc.str());
		const size_t cut = readLen < 0 ? string::npos : (size_t) readLen; // insert.trunc(0, readLen)
		seqOut << ">" << insert.getId() << " " << insert.getDesc() << "\n" << insert.getSeq().substr(0, cut) << "\n";
		if (mateOut.is_open()) mateOut << ">" << insert.getId() << " " << insert.getDesc() << "\n" << insert.revcom().getSeq().substr(0, cut) << "\n";
		n++;
	}
	return seqOut.good() ? EXIT_SUCCESS : EXIT_FAILURE;
}
