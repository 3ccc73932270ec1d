/*
 * oracle_hmm.cpp -- CPU ORACLE (test infrastructure, never shipped): banded plan-7 HMM
 * Viterbi, literal restatement of /root/reference/src/BandedHMMP7.{h,cpp} and the
 * alignment half of src/HmmUFOtu_main.cpp.  PARITY UNPINNED, see oracle.h.
 */
#include "oracle.h"
#include <algorithm>
#include <cassert>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <map>

namespace orc {

/* src/DegenAlphabet.cpp:43-64 with the IUPAC map of src/IUPACNucl.cpp:33-48:
 * ACGT -> 0..3, degenerate code -> first member, "-._" -> -2, anything else -1 */
int8_t encodeBase(char c) {
	switch (c) {
	case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
	case 'U': return 3;
	case 'M': case 'R': case 'W': case 'V': case 'H': case 'D': case 'N': return 0;
	case 'S': case 'Y': case 'B': return 1;
	case 'K': return 2;
	case '-': case '.': case '_': return -2;
	default: return -1;
	}
}

/* src/DigitalSeq.cpp:41-48 : upper-case, keep only valid characters */
std::vector<int8_t> digitalSeq(const std::string& align) {
	std::vector<int8_t> out;
	out.reserve(align.size());
	for (size_t i = 0; i < align.size(); ++i) {
		char c = (char) ::toupper(align[i]);
		int8_t b = encodeBase(c);
		if (b != -1)
			out.push_back(b);
	}
	return out;
}

/* ------------------------------------------------------------ parameters ---- */
static const int kMaxProfile = 65536; /* src/BandedHMMP7.h:279-280 */

void Hmm::setProfileSize(int k) { /* src/BandedHMMP7.cpp:551-558, 585-609 */
	K = k;
	eM.assign((size_t) 4 * (K + 1), INF);
	eI.assign((size_t) 4 * (K + 1), INF);
	T.assign((size_t) 9 * (K + 1), INF);
	entryPr.assign(K + 1, 0.0);
	exitPr.assign(K + 1, 0.0);
	entry.assign(K + 1, INF);
	exit_.assign(K + 1, INF);
	if (cs2p.empty()) cs2p.assign(kMaxProfile + 1, 0); /* "uninitialised, assume zero": SURVEY App. A */
	if (p2cs.empty()) p2cs.assign(kMaxProfile + 1, 0);
	wingRetracted = false;
}

void Hmm::finalize() {
	/* extend_index, src/BandedHMMP7.cpp:701-705 */
	for (int i = p2cs[K] + 1; i <= L && i < kMaxProfile; ++i)
		cs2p[i] = K;
	/* resetProbByCost (src/BandedHMMP7.h:862-869) then adjustProfileLocalMode (:721-733):
	 * entryPr(1..K) = Tmat[0](M,M) = exp(-Tmat_cost[0](M,M)); exitPr(1..K) = exp(-Tmat_cost[K](M,M)) */
	double ent = ::exp(-T[0 * 9 + ST_M * 3 + ST_M]);
	double ext = ::exp(-T[(size_t) K * 9 + ST_M * 3 + ST_M]);
	entryPr[0] = 0;
	exitPr[0] = 0;
	for (int k = 1; k <= K; ++k) {
		entryPr[k] = ent;
		exitPr[k] = ext;
	}
	for (int k = 0; k <= K; ++k) {
		entry[k] = -::log(entryPr[k]);
		exit_[k] = -::log(exitPr[k]);
	}
	wingRetracted = false;
	wingRetract();
}

void Hmm::wingRetract() { /* src/BandedHMMP7.cpp:1083-1120 */
	if (wingRetracted)
		return;
	for (int j = 2; j <= K; ++j) {
		double cost = 0;
		cost += T[0 * 9 + ST_M * 3 + ST_D];
		for (int i = 1; i < j - 1; ++i)
			cost += T[(size_t) i * 9 + ST_D * 3 + ST_D];
		cost += T[(size_t) (j - 1) * 9 + ST_D * 3 + ST_M];
		entryPr[j] += ::exp(-cost);
		if (entryPr[j] > 1)
			entryPr[j] = 1;
	}
	for (int i = 1; i <= K - 1; ++i) {
		double cost = 0;
		cost += T[(size_t) i * 9 + ST_M * 3 + ST_D];
		for (int j = i + 1; j < K; ++j)
			cost += T[(size_t) j * 9 + ST_D * 3 + ST_D];
		cost += T[(size_t) K * 9 + ST_D * 3 + ST_M];
		exitPr[i] += ::exp(-cost);
		if (exitPr[i] > 1)
			exitPr[i] = 1;
	}
	for (int k = 0; k <= K; ++k) {
		entry[k] = -::log(entryPr[k]);
		exit_[k] = -::log(exitPr[k]);
	}
	wingRetracted = true;
}

void Hmm::setSequenceMode(int mode) { /* src/BandedHMMP7.cpp:561-583, src/BandedHMMP7Bg.cpp:33-35 */
	const int MIN_BG_K = 350;
	double p1 = K >= MIN_BG_K ? K / (K + 1.0) : MIN_BG_K / (MIN_BG_K + 1.0);
	double term = 1 - p1;
	double tNN = 0, tCC = 0;
	switch (mode) {
	case GLOBAL: tNN = tCC = 0; break;
	case LOCAL: tNN = tCC = term; break;
	case NGCL: tNN = 0; tCC = term; break;
	case CGNL: tNN = term; tCC = 0; break;
	}
	double tNB = 1.0 - tNN;
	double tEC = 1.0;
	cNN = -::log(tNN);
	cCC = -::log(tCC);
	cNB = -::log(tNB);
	cEC = -::log(tEC);
}

static bool isIntegerTag(const std::string& s) { int i; return sscanf(s.c_str(), "%d", &i) == 1; }
static double hmmValueOf(const std::string& s) { return s != "*" ? ::atof(s.c_str()) : INF; }

int Hmm::loadText(const std::string& path) { /* src/BandedHMMP7.cpp:100-246 */
	std::ifstream in(path.c_str());
	if (!in) return -1;
	std::string line;
	int k = 0;
	std::map<std::string, std::string> optTags;
	cs2p.assign(kMaxProfile + 1, 0);
	p2cs.assign(kMaxProfile + 1, 0);
	while (std::getline(in, line)) {
		if (line == "//") {
			finalize();
			return 0;
		}
		std::istringstream iss(line);
		std::string tag, tmp;
		/* literal: `!isspace(line[0])` also holds for an empty line (line[0] == '\0'), i.e. the line
		 * remainder left behind by the `in >>` reads of a transition row is a header line with an empty tag */
		if (line.empty() || !isspace((unsigned char) line[0])) {
			iss >> tag;
			if (tag.substr(0, 6) == "HMMER3") {
				if (tag.length() < 8 || tag[7] < 'f') return -2;
			} else if (tag == "NAME") {
				iss >> name;
			} else if (tag == "LENG") {
				iss >> K;
				setProfileSize(K);
			} else if (tag == "ALPH") {
				std::string abc; iss >> abc;
				if (abc != "DNA") return -3;
			} else if (tag == "MAXL") {
				iss >> L;
			} else if (tag == "HMM") {
				std::getline(in, tmp);
			} else {
				std::string val; std::getline(iss, val);
				size_t b = val.find_first_not_of(" \t");
				size_t e = val.find_last_not_of(" \t");
				if (!tag.empty()) optTags[tag] = b == std::string::npos ? "" : val.substr(b, e - b + 1);
			}
		} else {
			iss >> tag;
			if (tag == "COMPO" || isIntegerTag(tag)) {
				double emit[4];
				for (int i = 0; i < 4; ++i) iss >> emit[i];
				if (tag == "COMPO") {
					for (int i = 0; i < 4; ++i) eM[0 * 4 + i] = emit[i];
				} else {
					for (int i = 0; i < 4; ++i) eM[(size_t) k * 4 + i] = emit[i];
					if (optTags["MAP"] != "yes") return -4;
					iss >> tmp;
					int map = atoi(tmp.c_str());
					cs2p[map] = k;
					p2cs[k] = map;
				}
				for (int i = 0; i < 4; ++i) in >> eI[(size_t) k * 4 + i];
			} else {
				/* non-COMPO begin line: the insert emissions follow on the stream */
				for (int i = 0; i < 4; ++i) in >> eI[(size_t) k * 4 + i];
			}
			static const int order[7][2] = {{ST_M, ST_M}, {ST_M, ST_I}, {ST_M, ST_D}, {ST_I, ST_M}, {ST_I, ST_I}, {ST_D, ST_M}, {ST_D, ST_D}};
			for (int t = 0; t < 7; ++t) {
				in >> tmp;
				T[(size_t) k * 9 + order[t][0] * 3 + order[t][1]] = hmmValueOf(tmp);
			}
			k++;
		}
	}
	return -5; /* reached EOF without '//' */
}

/* ------------------------------------------------------------- Viterbi ---- */
void VScores::reset() {
	size_t n = (size_t) (L + 1) * (K + 1);
	M.assign(n, INF);
	I.assign(n, INF);
	D.assign(n, INF);
	S.assign((size_t) (L + 1) * (K + 2), INF);
}

static void prepare(const Hmm& h, VScores& vs) { /* src/BandedHMMP7.cpp:735-746 */
	vs.m(0, 0) = vs.ii(0, 0) = vs.d(0, 0) = INF;
	for (int i = 1; i <= vs.L; i++)
		vs.m(i, 0) = i == 1 ? 0 : h.cNN * (i - 1);
	for (int i = 0; i <= vs.L; i++)
		vs.m(i, 0) += h.cNB;
	for (int i = 0; i <= vs.L; i++)
		vs.ii(i, 0) = vs.m(i, 0);
}

static inline double min3(double a, double b, double c) { return std::min(a, std::min(b, c)); }
static inline double min4(double a, double b, double c, double d) { return std::min(a, min3(b, c, d)); }

static inline void updCell(const Hmm& h, const std::vector<int8_t>& x, VScores& vs, int i, int j, bool useB) {
	const int K = h.K;
	const double* Tp = &h.T[(size_t) (j - 1) * 9];
	const double* Tj = &h.T[(size_t) j * 9];
	int b = x[i - 1];
	double mm = vs.m(i - 1, j - 1) + Tp[ST_M * 3 + ST_M];
	double im = vs.ii(i - 1, j - 1) + Tp[ST_I * 3 + ST_M];
	double dm = vs.d(i - 1, j - 1) + Tp[ST_D * 3 + ST_M];
	double best = useB ? min4(vs.m(i, 0) + h.entry[j], mm, im, dm) : min3(mm, im, dm);
	vs.m(i, j) = h.eM[(size_t) j * 4 + b] + best;
	vs.ii(i, j) = h.eI[(size_t) j * 4 + b] + std::min(vs.m(i - 1, j) + Tj[ST_M * 3 + ST_I], vs.ii(i - 1, j) + Tj[ST_I * 3 + ST_I]);
	if (j > 1 && j < K)
		vs.d(i, j) = std::min(vs.m(i, j - 1) + Tp[ST_M * 3 + ST_D], vs.d(i, j - 1) + Tp[ST_D * 3 + ST_D]);
}

static void finishS(const Hmm& h, VScores& vs) { /* src/BandedHMMP7.cpp:772-779 / 883-891 */
	const int K = h.K, L = vs.L;
	for (int j = 0; j <= K; ++j)
		for (int i = 0; i <= L; ++i)
			vs.s(i, j) = vs.m(i, j);
	for (int i = 0; i <= L; ++i)
		vs.s(i, K + 1) = vs.ii(i, K);
	for (int j = 0; j <= K; ++j)
		for (int i = 0; i <= L; ++i)
			vs.s(i, j) += h.exit_[j];
	for (int i = 0; i <= L; ++i)
		vs.s(i, K + 1) += h.T[(size_t) K * 9 + ST_I * 3 + ST_M];
	for (size_t n = 0; n < vs.S.size(); ++n)
		vs.S[n] += h.cEC;
	for (int i = 1; i < L; ++i)
		for (int j = 0; j <= K + 1; ++j)
			vs.s(i, j) += h.cCC * (L - i);
}

static std::vector<int8_t> encodeRead(const std::string& read) {
	std::vector<int8_t> x(read.size());
	for (size_t i = 0; i < read.size(); ++i) {
		x[i] = encodeBase(read[i]);
		assert(x[i] >= 0 && "reads must be upper-case nucleotides (SURVEY App. D.3)");
	}
	return x;
}

void calcViterbiFull(const Hmm& h, const std::string& read, VScores& vs) {
	std::vector<int8_t> x = encodeRead(read);
	const int L = vs.L;
	prepare(h, vs);
	for (int j = 1; j <= h.K; ++j)
		for (int i = 1; i <= L; ++i)
			updCell(h, x, vs, i, j, true);
	finishS(h, vs);
}

long calcViterbiBanded(const Hmm& h, const std::string& read, VScores& vs, const std::vector<VPath>& vpaths) {
	const double kMinGapFrac = 0.2;
	std::vector<int8_t> x = encodeRead(read);
	const int L = vs.L, K = h.K;
	long cells = 0;
	if (vpaths.empty())
		return 0;
	prepare(h, vs);
	for (size_t p = 0; p < vpaths.size(); ++p) {
		const VPath& vp = vpaths[p];
		bool first = p == 0;
		int upQLen = first ? vp.from - 1 : vp.from - vpaths[p - 1].to;
		if (upQLen < 0) upQLen = 0;
		int up_start = first ? (int) (vp.start - upQLen * (1 + kMinGapFrac)) : vpaths[p - 1].end;
		if (up_start < 1) up_start = 1;
		int up_from = first ? (int) (vp.from - upQLen * (1 + kMinGapFrac)) : vpaths[p - 1].to;
		if (up_from < 1) up_from = 1;
		for (int j = up_start; j <= vp.start; ++j)
			for (int i = up_from; i <= vp.from; ++i) { updCell(h, x, vs, i, j, true); cells++; }
		for (int j = vp.start; j <= vp.end; ++j)
			for (int i = vp.from; i <= vp.to; ++i) {
				int dist = (i - vp.from) - (j - vp.start);
				if (!(dist <= vp.nIns && dist >= -vp.nDel))
					continue;
				updCell(h, x, vs, i, j, true); cells++;
			}
	}
	int last_end = vpaths.back().end;
	int last_to = vpaths.back().to;
	int downQLen = L - last_to;
	int down_end = (int) (last_end + downQLen * (1 + kMinGapFrac));
	int down_to = (int) (last_to + downQLen * (1 + kMinGapFrac));
	if (down_end > K) down_end = K;
	if (down_to > L) down_to = L;
	for (int j = last_end; j <= down_end; ++j)
		for (int i = last_to; i <= down_to; ++i) { updCell(h, x, vs, i, j, false); cells++; }
	finishS(h, vs);
	return cells;
}

/* src/BandedHMMP7.h:722-788 : strict '<', index defaults to 0 */
static char whichMin4(double b, double m, double i, double d, const char* st = "BMID") {
	int idx = 0; double mn = INF;
	if (b < mn) { idx = 0; mn = b; }
	if (m < mn) { idx = 1; mn = m; }
	if (i < mn) { idx = 2; mn = i; }
	if (d < mn) { idx = 3; mn = d; }
	return st[idx];
}
static char whichMin2(double a, double b, const char* st) {
	int idx = 0; double mn = INF;
	if (a < mn) { idx = 0; mn = a; }
	if (b < mn) { idx = 1; mn = b; }
	return st[idx];
}

void buildViterbiTrace(const Hmm& h, VScores& vs, VTrace& vt) { /* src/BandedHMMP7.cpp:943-1006 */
	const int K = h.K, L = vs.L;
	/* S.minCoeff(&row,&col): column-major visit, first strict minimum */
	int minRow = 0, minCol = 0;
	double mn = vs.s(0, 0);
	for (int j = 0; j <= K + 1; ++j)
		for (int i = 0; i <= L; ++i)
			if (vs.s(i, j) < mn) { mn = vs.s(i, j); minRow = i; minCol = j; }
	vt.minScore = mn;
	if (vt.minScore == INF)
		return;
	char s = minCol <= K ? 'M' : 'I';
	int i = minRow;
	int j = minCol <= K ? minCol : K;
	vt.alnEnd = j;
	vt.alnTo = minRow;
	vt.trace.push_back('E');
	while (i >= 1 && j >= 0) {
		vt.trace.push_back(s);
		if (s == 'M') {
			const double* Tp = &h.T[(size_t) (j - 1) * 9];
			s = j > 1 ? whichMin4(vs.m(i, 0) + h.entry[j], vs.m(i - 1, j - 1) + Tp[ST_M * 3 + ST_M],
					vs.ii(i - 1, j - 1) + Tp[ST_I * 3 + ST_M], vs.d(i - 1, j - 1) + Tp[ST_D * 3 + ST_M])
				: whichMin2(vs.m(i, 0) + h.entry[j], vs.ii(i - 1, j - 1) + Tp[ST_I * 3 + ST_M], "BI");
			i--; j--;
		} else if (s == 'I') {
			const double* Tj = &h.T[(size_t) j * 9];
			s = j > 0 ? whichMin2(vs.m(i - 1, j) + Tj[ST_M * 3 + ST_I], vs.ii(i - 1, j) + Tj[ST_I * 3 + ST_I], "MI")
				: whichMin2(vs.m(i, 0) + h.T[ST_M * 3 + ST_I], vs.ii(i - 1, j) + Tj[ST_I * 3 + ST_I], "BI");
			i--;
		} else if (s == 'D') {
			const double* Tp = &h.T[(size_t) (j - 1) * 9];
			s = whichMin2(vs.m(i, j - 1) + Tp[ST_M * 3 + ST_D], vs.d(i, j - 1) + Tp[ST_D * 3 + ST_D], "MD");
			j--;
		} else
			break;
	}
	vt.alnStart = j + 1;
	vt.alnFrom = i + 1;
	if (vt.trace[vt.trace.size() - 1] != 'B')
		vt.trace.push_back('B');
	std::reverse(vt.trace.begin(), vt.trace.end());
}

enum PadMode { PAD_LEFT, PAD_RIGHT, PAD_JUSTIFIED };
static std::string getPaddingSeq(int L, const std::string& insert, char padCh, PadMode mode) { /* src/BandedHMMP7.cpp:1137-1186 */
	if (insert.empty())
		return std::string(L, padCh);
	std::string pad;
	int n = (int) insert.length();
	switch (mode) {
	case PAD_LEFT:
		if (n >= L) pad.append(insert.substr(0, L));
		else { pad.append(insert); pad.append(L - n, padCh); }
		break;
	case PAD_RIGHT:
		if (n >= L) pad.append(insert.substr(n - L, L));
		else { pad.append(L - n, padCh); pad.append(insert); }
		break;
	case PAD_JUSTIFIED:
		if (n >= L) {
			pad.append(insert.substr(0, (int) ::floor(L / 2.0)));
			pad.append(insert.substr(n - (int) ::ceil(L / 2.0), (int) ::ceil(L / 2.0)));
		} else { /* literal: the head of the insert is duplicated */
			pad.append(insert.substr(0, (int) ::floor(n / 2.0)));
			pad.append(L - n, padCh);
			pad.append(insert.substr(0, (int) ::ceil(n / 2.0)));
		}
		break;
	}
	assert((int) pad.length() == L);
	return pad;
}

HmmAlignment buildGlobalAlign(const Hmm& h, const std::string& seq, const VTrace& vt) { /* src/BandedHMMP7.cpp:1008-1081 */
	HmmAlignment aln;
	const int L = h.L;
	const std::string seqN = seq.substr(0, vt.alnFrom - 1);
	const std::string seqC = seq.substr(vt.alnTo, (size_t) (long) (L - vt.alnTo)); /* literal: length L - alnTo with L = csLen */
	int csStart = h.p2cs[vt.alnStart];
	int csEnd = h.p2cs[vt.alnEnd];
	int j = 0, k = 0;
	std::string insert;
	const std::string& tr = vt.trace;
	for (size_t t = 0; t < tr.size(); ++t) {
		switch (tr[t]) {
		case 'B':
			aln.align.append(getPaddingSeq(csStart - 1, seqN, '.', PAD_RIGHT));
			j = vt.alnFrom;
			k = vt.alnStart;
			break;
		case 'M':
			if (k > 1 && t > 1 && h.p2cs[k] - h.p2cs[k - 1] > 1)
				aln.align.append(getPaddingSeq(h.p2cs[k] - h.p2cs[k - 1] - 1, insert, '-', PAD_JUSTIFIED));
			insert.clear();
			aln.align.push_back(seq.at(j - 1));
			j++; k++;
			break;
		case 'I':
			insert.clear();
			while (t < tr.size() && tr[t] == 'I') {
				insert.push_back((char) ::tolower(seq.at(j - 1)));
				j++; t++;
			}
			t--;
			break;
		case 'D':
			if (k > 1 && h.p2cs[k] - h.p2cs[k - 1] > 1)
				aln.align.append(h.p2cs[k] - h.p2cs[k - 1] - 1, '-');
			aln.align.push_back('-');
			k++;
			break;
		case 'E':
			aln.align.append(getPaddingSeq(L - csEnd, seqC, '.', PAD_LEFT));
			break;
		}
	}
	aln.K = h.K; aln.L = L;
	aln.seqStart = vt.alnFrom; aln.seqEnd = vt.alnTo;
	aln.hmmStart = vt.alnStart; aln.hmmEnd = vt.alnEnd;
	aln.csStart = csStart; aln.csEnd = csEnd;
	aln.cost = vt.minScore;
	aln.trace = vt.trace;
	return aln;
}

bool HmmAlignment::isValid() const { /* src/BandedHMMP7.h:89-94 */
	return 0 < seqStart && seqStart <= seqEnd && 0 < hmmStart && hmmStart <= hmmEnd && hmmEnd <= K &&
		0 < csStart && csStart <= csEnd && csEnd <= L && cost >= 0 && cost != INF && (size_t) L == align.length();
}

HmmAlignment& HmmAlignment::merge(const HmmAlignment& o) {
	if (K == o.K && L == o.L) {
		if (o.seqStart < seqStart) seqStart = o.seqStart;
		if (o.seqEnd > seqEnd) seqEnd = o.seqEnd;
		if (o.hmmStart < hmmStart) hmmStart = o.hmmStart;
		if (o.hmmEnd > hmmEnd) hmmEnd = o.hmmEnd;
		if (o.csStart < csStart) csStart = o.csStart;
		if (o.csEnd > csEnd) csEnd = o.csEnd;
		cost += o.cost;
		for (int i = 0; i < L; ++i)
			if (align[i] == '.' && o.align[i] != '.')
				align[i] = o.align[i];
	}
	return *this;
}

HmmAlignment alignWithPaths(const Hmm& h, const std::string& read, const std::vector<VPath>& vpaths) { /* src/HmmUFOtu_main.cpp:86-104 */
	VScores vs(h.K, (int) read.size());
	VTrace vt;
	bool fallback = false;
	long cells = 0;
	if (!vpaths.empty()) {
		cells = calcViterbiBanded(h, read, vs, vpaths);
		double mn = INF;
		for (size_t n = 0; n < vs.S.size(); ++n) mn = std::min(mn, vs.S[n]);
		if (mn == INF) {
			vs.reset();
			calcViterbiFull(h, read, vs);
			fallback = true;
			cells += (long) read.size() * h.K;
		}
	} else {
		calcViterbiFull(h, read, vs);
		fallback = true;
		cells = (long) read.size() * h.K;
	}
	buildViterbiTrace(h, vs, vt);
	assert(vt.minScore != INF);
	HmmAlignment aln = buildGlobalAlign(h, read, vt);
	aln.usedFallback = fallback;
	aln.bandCells = cells;
	return aln;
}

} // namespace orc
