// Host-side C++ layer: formats, alphabet, FASTA/FASTQ, the CSFM seed index and the reference's class surface
// over the C ABI.  See hu_host.h.  Citations are relative to /root/reference/src.
#include "hu_host.h"
#include <atomic>

#include <pthread.h>

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cctype>
#include <cfloat>
#include <cmath>
#include <iomanip>
#include <limits>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace EGriceLab {
namespace HmmUFOtu {

using namespace std;

const string progName = "HmmUFOtu";
const int progVer[3] = {1, 5, 1};
static const double inf = numeric_limits<double>::infinity();

// ------------------------------------------------------------------ binary helpers ----
istream& loadProgInfo(istream& in) {
	char buf[8];
	in.read(buf, 8);
	int v[3] = {0, 0, 0};
	in.read((char*) v, sizeof v);
	if (in.bad() || !in || string(buf, 8) != progName) {
		cerr << "Not an valid database file of " << progName << endl;
		in.setstate(ios_base::badbit);
		return in;
	}
	if (!(make_pair(progVer[0], make_pair(progVer[1], progVer[2])) >= make_pair(v[0], make_pair(v[1], v[2])))) {
		cerr << "You are using an old version of " << progName << " to read a newer database file" << endl;
		in.setstate(ios_base::badbit);
	}
	return in;
}
ostream& saveProgInfo(ostream& out) {
	out.write(progName.data(), progName.length());
	out.write((const char*) progVer, sizeof progVer);
	return out;
}
ostream& writeProgInfo(ostream& out, const string& info) {
	out << "# " << progName << " v" << progVer[0] << "." << progVer[1] << "." << progVer[2] << info << endl;
	return out;
}
bool loadString(string& dest, istream& in) {
	size_t len = 0;
	in.read((char*) &len, sizeof len);
	if (!in || len > ((size_t) 1 << 40)) { in.setstate(ios_base::badbit); return false; }
	dest.resize(len);
	if (len) in.read(&dest[0], len);
	return (bool) in;
}
void saveString(const string& src, ostream& out) {
	size_t len = src.length();
	out.write((const char*) &len, sizeof len);
	out.write(src.data(), len);
}

// ------------------------------------------------------------------ alphabet ----
int8_t DNA::encode(char c) {
	switch (c) {
	case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': case 'U': return 3;
	case 'M': case 'R': case 'W': case 'V': case 'H': case 'D': case 'N': return 0; // first member of the degenerate set
	case 'S': case 'Y': case 'B': return 1;
	case 'K': return 2;
	case '-': case '.': case '_': return -2;
	default: return -1;
	}
}
char DNA::complement(char c) {
	static const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
	static const char* to = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
	const char* p = strchr(from, c);
	return p && c ? to[p - from] : c;
}
PrimarySeq PrimarySeq::revcom() const {
	PrimarySeq r(*this);
	std::reverse(r.seq.begin(), r.seq.end());
	for (size_t i = 0; i < r.seq.size(); ++i) r.seq[i] = DNA::complement(r.seq[i]);
	std::reverse(r.qual.begin(), r.qual.end());
	return r;
}

// ------------------------------------------------------------------ SeqIO ----
bool SeqIO::hasNext() {
	int c = in->peek();
	return c != EOF && c == (fmt == "fastq" ? '@' : '>');
}
PrimarySeq SeqIO::nextSeq() { // src/SeqIO.cpp:85-125
	string id, seq, desc, qual, line;
	in->get();
	*in >> id;
	while (::isspace(in->peek()) && in->peek() != '\n') in->get();
	getline(*in, desc);
	if (fmt == "fastq") {
		getline(*in, seq);
		getline(*in, line);
		getline(*in, qual);
	} else
		while (in->peek() != EOF && in->peek() != '>') {
			getline(*in, line);
			seq += line;
		}
	while (!seq.empty() && (seq[seq.size() - 1] == '\r')) seq.erase(seq.size() - 1);
	return PrimarySeq(id, seq, desc, qual);
}
string SeqIO::guessFormat(const string& fn) { // src/hmmufotu.cpp:301-316
	string f = fn;
	for (size_t i = 0; i < f.size(); ++i) f[i] = (char) ::tolower(f[i]);
	const char* fq[] = {".fastq", ".fq"};
	for (int i = 0; i < 2; ++i)
		if (f.size() >= strlen(fq[i]) && f.compare(f.size() - strlen(fq[i]), strlen(fq[i]), fq[i]) == 0) return "fastq";
	return "fasta";
}

// ------------------------------------------------------------------ MSA ----
istream& MSA::load(istream& in) {
	loadString(alphabet, in);
	loadString(name, in);
	in.read((char*) &numSeq, sizeof(unsigned));
	in.read((char*) &csLen, sizeof(unsigned));
	loadString(CS, in);
	in.read((char*) &isPruned, sizeof(bool));
	if (!in || numSeq > (1u << 28) || csLen > (1u << 20)) { in.setstate(ios_base::badbit); return in; }
	seqNames.resize(numSeq);
	for (unsigned i = 0; i < numSeq; ++i) loadString(seqNames[i], in);
	loadString(concatMSA, in);
	if (concatMSA.size() != (size_t) numSeq * csLen) { in.setstate(ios_base::badbit); return in; }
	// auxiliary index, counts and weights: present in the file, not used on the run path
	const size_t skip = (size_t) 3 * numSeq * sizeof(int) + (size_t) 4 * csLen * sizeof(int) + (size_t) csLen * sizeof(int) +
			(size_t) numSeq * sizeof(double) + (size_t) 4 * csLen * sizeof(double) + (size_t) csLen * sizeof(double);
	auxBlocks.resize(skip);
	in.read(&auxBlocks[0], (streamsize) skip);
	if (in.gcount() != (streamsize) skip) in.setstate(ios_base::badbit);
	return in;
}
ostream& MSA::save(ostream& out) const {
	saveString(alphabet, out);
	saveString(name, out);
	out.write((const char*) &numSeq, sizeof(unsigned));
	out.write((const char*) &csLen, sizeof(unsigned));
	saveString(CS, out);
	out.write((const char*) &isPruned, sizeof(bool));
	for (unsigned i = 0; i < numSeq; ++i) saveString(seqNames[i], out);
	saveString(concatMSA, out);
	if (!auxBlocks.empty()) { // a loaded database: the builder's counts and sequence weights go back unchanged
		out.write(auxBlocks.data(), (streamsize) auxBlocks.size());
		return out;
	}
	if (seqWeight.size() == numSeq && resWCount.size() == (size_t) 4 * csLen && numSeq > 0) { // built here (hu_build.cpp)
		out.write((const char*) startIdx.data(), sizeof(int) * numSeq);
		out.write((const char*) endIdx.data(), sizeof(int) * numSeq);
		out.write((const char*) lenIdx.data(), sizeof(int) * numSeq);
		out.write((const char*) resCount.data(), sizeof(int) * resCount.size());
		out.write((const char*) gapCount.data(), sizeof(int) * gapCount.size());
		out.write((const char*) seqWeight.data(), sizeof(double) * numSeq);
		out.write((const char*) resWCount.data(), sizeof(double) * resWCount.size());
		out.write((const char*) gapWCount.data(), sizeof(double) * gapWCount.size());
		return out;
	}
	vector<int> start(numSeq), end(numSeq), len(numSeq), res((size_t) 4 * csLen, 0), gap(csLen, 0);
	for (unsigned i = 0; i < numSeq; ++i) {
		int s = -1, e = -1, n = 0;
		for (unsigned j = 0; j < csLen; ++j) {
			const int8_t b = DNA::encode((char) ::toupper(residualAt(i, j)));
			if (b >= 0) { if (s < 0) s = j; e = j; n++; res[(size_t) j * 4 + b]++; }
			else gap[j]++;
		}
		start[i] = s; end[i] = e; len[i] = n;
	}
	out.write((const char*) start.data(), sizeof(int) * numSeq);
	out.write((const char*) end.data(), sizeof(int) * numSeq);
	out.write((const char*) len.data(), sizeof(int) * numSeq);
	out.write((const char*) res.data(), sizeof(int) * res.size());
	out.write((const char*) gap.data(), sizeof(int) * gap.size());
	vector<double> w(numSeq, 1.0), resW(res.begin(), res.end()), gapW(gap.begin(), gap.end());
	out.write((const char*) w.data(), sizeof(double) * numSeq);
	out.write((const char*) resW.data(), sizeof(double) * resW.size());
	out.write((const char*) gapW.data(), sizeof(double) * gapW.size());
	return out;
}

// ------------------------------------------------------------------ CSFM index ----
// Suffix array by prefix doubling with radix passes; text symbols 0..4, last symbol is the unique terminator.
static void buildSuffixArray(const vector<uint8_t>& text, vector<int32_t>& sa) {
	const int32_t n = (int32_t) text.size();
	sa.resize(n);
	vector<int32_t> rnk(n), tmp(n), cnt, sa2(n);
	for (int32_t i = 0; i < n; ++i) { sa[i] = i; rnk[i] = text[i]; }
	// the final '\0' must sort before every other suffix that starts with a separator: give it rank -1 by shifting
	for (int32_t i = 0; i < n; ++i) rnk[i] += 1;
	rnk[n - 1] = 0;
	int32_t maxRank = 6;
	for (int32_t k = 0;; k = k ? k * 2 : 1) {
		// sort by (rnk[i], rnk[i+k]) with two counting passes (k = 0: by rnk only)
		if (k > 0) {
			cnt.assign(maxRank + 2, 0);
			for (int32_t i = 0; i < n; ++i) cnt[(i + k < n ? rnk[i + k] + 1 : 0) + 1]++;
			for (int32_t i = 1; i < (int32_t) cnt.size(); ++i) cnt[i] += cnt[i - 1];
			for (int32_t i = 0; i < n; ++i) sa2[cnt[i + k < n ? rnk[i + k] + 1 : 0]++] = i;
		} else
			for (int32_t i = 0; i < n; ++i) sa2[i] = i;
		cnt.assign(maxRank + 2, 0);
		for (int32_t i = 0; i < n; ++i) cnt[rnk[i] + 1]++;
		for (int32_t i = 1; i < (int32_t) cnt.size(); ++i) cnt[i] += cnt[i - 1];
		for (int32_t i = 0; i < n; ++i) sa[cnt[rnk[sa2[i]]]++] = sa2[i];
		tmp[sa[0]] = 0;
		int32_t r = 0;
		for (int32_t i = 1; i < n; ++i) {
			const int32_t a = sa[i - 1], b = sa[i];
			const int32_t a2 = k > 0 ? (a + k < n ? rnk[a + k] : -1) : 0, b2 = k > 0 ? (b + k < n ? rnk[b + k] : -1) : 0;
			if (rnk[a] != rnk[b] || a2 != b2) r++;
			tmp[b] = r;
		}
		rnk.swap(tmp);
		maxRank = r + 1;
		if (r == n - 1) break;
		if (k >= n) break;
	}
}

CSFMIndex& CSFMIndex::build(const MSA& msa) {
	if (msa.csLen > 65535) throw runtime_error("CSFMIndex cannot handle MSA with consensus length longer than 65535");
	csLen = msa.csLen;
	// concatenated gap-stripped rows, symbols 1..4, one separator (0) after every row, a final terminator
	vector<uint8_t> text;
	concat2CS.assign(1, 0);                       // 1-based text positions; [0] unused
	for (unsigned i = 0; i < msa.numSeq; ++i) {
		for (unsigned j = 0; j < msa.csLen; ++j) {
			const char c = msa.residualAt(i, j);
			if (DNA::isGap(c)) continue;
			const int8_t b = DNA::encode((char) ::toupper(c));
			text.push_back((uint8_t) ((b < 0 ? 0 : b) + 1));
			concat2CS.push_back((uint16_t) (j + 1));
		}
		text.push_back(0);
		concat2CS.push_back(0);
	}
	concatLen = (int32_t) text.size();
	text.push_back(0);                           // terminator
	concat2CS.push_back(0);
	const int32_t N = concatLen + 1;
	memset(C, 0, sizeof C);
	vector<int32_t> freq(6, 0);
	for (int32_t i = 0; i < N; ++i) freq[text[i]]++;
	for (int b = 0; b < 5; ++b) C[b + 1] = C[b] + freq[b];
	buildSuffixArray(text, sa);
	rank.assign((size_t) N / 64 + 2, RankBlock());
	uint32_t run[5] = {0, 0, 0, 0, 0};
	for (int32_t i = 0; i < N; ++i) {
		RankBlock& rb = rank[(size_t) i / 64];
		if (i % 64 == 0)
			for (int b = 0; b < 5; ++b) { rb.cnt[b] = run[b]; rb.mask[b] = 0; }
		const int b = sa[i] == 0 ? 0 : text[sa[i] - 1]; // BWT symbol
		rb.mask[b] |= 1ull << (i % 64);
		run[b]++;
	}
	for (int32_t i = 0; i < N; ++i) sa[i] += 1; // the reference's SA holds 1-based text positions
	buildQTable();
	return *this;
}

int32_t CSFMIndex::occ(int b, int32_t i) const { // symbols b in BWT[0..i]: one cache line and one popcount
	if (i < 0) return 0;
	const RankBlock& rb = rank[(size_t) i / 64];
	const int r = i % 64;
	return (int32_t) (rb.cnt[b] + __builtin_popcountll(rb.mask[b] & (r == 63 ? ~0ull : ((1ull << (r + 1)) - 1ull))));
}

// The first qlen steps of every backward search, tabulated: most attempts of the seed search die within log4(text length) steps
// (a 20-mer with one mismatch near its end matches nothing), so one table lookup settles them.  qlen = that many symbols, at
// most 11 (33 MB).  Depth-first over the q-mers, last symbol first, with the interval carried along; empty branches stay empty.
void CSFMIndex::fillQ(int depth, size_t idx, int32_t start, int32_t end) {
	for (int b = 1; b <= 4; ++b) {
		int32_t s2, e2;
		if (depth == 0) { s2 = C[b]; e2 = C[b + 1] - 1; }
		else { s2 = C[b] + occ(b, start - 1); e2 = C[b] + occ(b, end) - 1; }
		if (s2 > e2) continue;
		const size_t id2 = idx * 4 + (size_t) (b - 1);
		if (depth + 1 == qlen) { qtab[2 * id2] = s2; qtab[2 * id2 + 1] = e2; }
		else fillQ(depth + 1, id2, s2, e2);
	}
}
void CSFMIndex::buildQTable() {
	qlen = 0;
	qtab.clear();
	if (rank.empty() || concatLen < 256) return;
	int q = 1;
	for (long long v = 4; v < concatLen && q < 11; v *= 4) q++;
	qlen = q;
	qtab.assign((size_t) 2 << (2 * q), 0);
	for (size_t i = 0; i < qtab.size(); i += 2) qtab[i] = 1; // (1, 0): no match
	fillQ(0, 0, 0, 0);
}

// backward search, src/CSFMIndex.cpp:79-99 / 121-135: [start, end] in SA coordinates
bool CSFMIndex::search(const char* pattern, size_t len, int32_t& start, int32_t& end) const {
	if (len == 0 || rank.empty()) return false;
	start = 0;
	end = concatLen;
	bool first = true;
	const char* c = pattern + len;
	if (qlen > 0 && len >= (size_t) qlen) { // the last qlen symbols from the table
		size_t idx = 0;
		for (int k = 0; k < qlen; ++k) {
			const int8_t e = DNA::encode(*--c);
			if (e < 0) return false; // the plain loop would meet this symbol, or run out of matches before it: false either way
			idx = idx * 4 + (size_t) e;
		}
		start = qtab[2 * idx];
		end = qtab[2 * idx + 1];
		first = false;
	}
	for (; c != pattern && start <= end;) {
		const int8_t e = DNA::encode(*--c);
		if (e < 0) return false;
		const int b = e + 1;
		if (first) {
			start = C[b];
			end = C[b + 1] - 1;
			first = false;
		} else {
			start = C[b] + occ(b, start - 1);
			end = C[b] + occ(b, end) - 1;
		}
	}
	return start <= end;
}
int32_t CSFMIndex::count(const string& pattern) const {
	int32_t s, e;
	return search(pattern, s, e) ? e - s + 1 : 0;
}
string CSFMIndex::extractCS(int32_t start, const string& pattern) const {
	string cs;
	for (int32_t i = start; i < start + (int32_t) pattern.length(); ++i) {
		if (i > start && concat2CS[i] - concat2CS[i - 1] > 1) cs.append(concat2CS[i] - concat2CS[i - 1] - 1, '-');
		cs.push_back(pattern[i - start]);
	}
	return cs;
}
CSLoc CSFMIndex::locateAt(const string& pattern, int32_t s, int32_t e) const {
	const int32_t i = s + rand() % (e - s + 1);  // src/CSFMIndex.cpp:139
	const int32_t pos = sa[i];
	return CSLoc(concat2CS[pos], concat2CS[pos + pattern.length() - 1], extractCS(pos, pattern));
}
CSLoc CSFMIndex::locateOne(const string& pattern) const {
	int32_t s, e;
	if (!search(pattern, s, e)) return CSLoc();
	return locateAt(pattern, s, e);
}
CSLoc CSFMIndex::locateFirst(const string& pattern) const {
	int32_t s, e;
	if (!search(pattern, s, e)) return CSLoc();
	const int32_t pos = sa[s];
	return CSLoc(concat2CS[pos], concat2CS[pos + pattern.length() - 1], extractCS(pos, pattern));
}
vector<CSLoc> CSFMIndex::locate(const string& pattern) const {
	vector<CSLoc> locs;
	int32_t s, e;
	if (!search(pattern, s, e)) return locs;
	for (int32_t i = s; i <= e; ++i) {
		const int32_t pos = sa[i];
		locs.push_back(CSLoc(concat2CS[pos], concat2CS[pos + pattern.length() - 1], extractCS(pos, pattern)));
	}
	return locs;
}

// ------------------------------------------------------------------ BandedHMMP7 ----
const string BandedHMMP7::HmmAlignment::TSV_HEADER = "seq_start\tseq_end\thmm_start\thmm_end\tCS_start\tCS_end\tcost\talignment";

static double hmmValueOf(const string& s) { return s != "*" ? ::atof(s.c_str()) : inf; }
static bool isInteger(const string& s) { int v; return sscanf(s.c_str(), "%d", &v) == 1; }
static string trim(const string& s) {
	const size_t b = s.find_first_not_of(" \t"), e = s.find_last_not_of(" \t");
	return b == string::npos ? "" : s.substr(b, e - b + 1);
}
static const int kTransOrder[7] = {0, 1, 2, 3, 4, 6, 8}; // m->m m->i m->d i->m i->i d->m d->d in a*3+b

istream& operator>>(istream& in, BandedHMMP7& hmm) {
	string line;
	int k = 0;
	map<string, string> tags;
	while (getline(in, line)) {
		if (line == "//") {
			// extend_index (cpp:701-705): columns after the last match column map to K
			for (int i = hmm.profile2CSIdx[hmm.K] + 1; i <= hmm.L; ++i) hmm.cs2ProfileIdx[i] = hmm.K;
			hmm.wingRetracted = false;
			hmm.wingRetract(); // adjustProfileLocalMode + wingRetract
			return in;
		}
		istringstream iss(line);
		string tag, tmp;
		if (line.empty() || !isspace((unsigned char) line[0])) { // header section (an empty line has line[0] == '\0')
			iss >> tag;
			if (tag.substr(0, 6) == "HMMER3") {
				if (tag.length() < 8 || tag[7] < 'f') {
					cerr << "Obsolete HMM file version: " << tag << ", must be HMMER3/f or higher" << endl;
					in.setstate(ios_base::badbit);
					return in;
				}
			} else if (tag == "NAME") iss >> hmm.name;
			else if (tag == "LENG") {
				iss >> hmm.K;
				if (hmm.K < 1 || hmm.K > 65535) { in.setstate(ios_base::badbit); return in; }
				hmm.E_M_cost.assign((size_t) 4 * (hmm.K + 1), inf);
				hmm.E_I_cost.assign((size_t) 4 * (hmm.K + 1), inf);
				hmm.Tmat_cost.assign((size_t) 9 * (hmm.K + 1), inf);
				hmm.profile2CSIdx.assign(hmm.K + 1, 0);
				hmm.locMap.assign(hmm.K + 1, "");
			} else if (tag == "ALPH") {
				string abc;
				iss >> abc;
				if (abc != "DNA") throw invalid_argument("Not allowed alphabet '" + abc + "' in the HMM input file! Must be DNA");
			} else if (tag == "MAXL") {
				iss >> hmm.L;
				hmm.cs2ProfileIdx.assign(hmm.L + 2, 0); // non-match columns: 0 (uninitialised in the reference)
			} else if (tag == "HMM") getline(in, tmp);
			else {
				string val;
				getline(iss, val);
				if (!tag.empty()) { tags[tag] = trim(val); hmm.optTags.push_back(make_pair(tag, trim(val))); }
			}
		} else {
			iss >> tag;
			if (hmm.K == 0 || hmm.L == 0) { in.setstate(ios_base::badbit); return in; }
			if (tag == "COMPO" || isInteger(tag)) {
				double e[4];
				for (int i = 0; i < 4; ++i) { iss >> tmp; e[i] = hmmValueOf(tmp); }
				if (k > hmm.K) { in.setstate(ios_base::badbit); return in; }
				for (int i = 0; i < 4; ++i) hmm.E_M_cost[(size_t) k * 4 + i] = e[i];
				if (tag != "COMPO") {
					if (tags["MAP"] != "yes") {
						cerr << "Error: HMM file must has the MAP flag set to 'yes'" << endl;
						in.setstate(ios_base::badbit);
						return in;
					}
					iss >> tmp;
					const int mapCol = atoi(tmp.c_str());
					if (mapCol < 1 || mapCol > hmm.L) { in.setstate(ios_base::badbit); return in; }
					hmm.cs2ProfileIdx[mapCol] = k;
					hmm.profile2CSIdx[k] = mapCol;
					string rest;
					getline(iss, rest);
					hmm.locMap[k] = trim(rest);
				}
			}
			for (int i = 0; i < 4; ++i) { in >> tmp; hmm.E_I_cost[(size_t) k * 4 + i] = hmmValueOf(tmp); }
			for (int t = 0; t < 7; ++t) { in >> tmp; hmm.Tmat_cost[(size_t) k * 9 + kTransOrder[t]] = hmmValueOf(tmp); }
			k++;
		}
	}
	in.setstate(ios::failbit);
	return in;
}

static void hmmPrintValue(ostream& out, double v) { if (v == inf) out << "*"; else out << v; }

ostream& operator<<(ostream& out, const BandedHMMP7& hmm) {
	out << "HMMER3/f\t" << progName << "-v" << progVer[0] << "." << progVer[1] << "." << progVer[2] << endl; // hmmVersion as hmmufotu-build sets it
	out << "NAME\t" << hmm.name << endl;
	out << "LENG\t" << hmm.K << endl;
	out << "ALPH\tDNA" << endl;
	bool hasMaxl = false;
	for (size_t i = 0; i < hmm.optTags.size(); ++i) hasMaxl |= hmm.optTags[i].first == "MAXL";
	if (!hasMaxl) out << "MAXL  " << hmm.L << endl;
	for (size_t i = 0; i < hmm.optTags.size(); ++i) out << hmm.optTags[i].first << "  " << hmm.optTags[i].second << endl;
	out << "HMM\t\tA\tC\tG\tT" << endl << "\t\tm->m\tm->i\tm->d\ti->m\ti->i\td->m\td->d" << endl;
	for (int k = 0; k <= hmm.K; ++k) {
		if (k == 0) out << "\tCOMPO";
		else out << "\t" << k;
		for (int i = 0; i < 4; ++i) hmmPrintValue(out << "\t", hmm.E_M_cost[(size_t) k * 4 + i]);
		if (k > 0) {
			out << "\t" << hmm.profile2CSIdx[k];
			if (!hmm.locMap[k].empty()) out << "\t" << hmm.locMap[k];
		}
		out << endl << "\t";
		for (int i = 0; i < 4; ++i) hmmPrintValue(out << "\t", hmm.E_I_cost[(size_t) k * 4 + i]);
		out << endl << "\t";
		for (int t = 0; t < 7; ++t) hmmPrintValue(out << "\t", hmm.Tmat_cost[(size_t) k * 9 + kTransOrder[t]]);
		out << endl;
	}
	out << "//" << endl;
	return out;
}

void BandedHMMP7::wingRetract() {
	if (wingRetracted) return;
	entryPr_cost.assign(K + 1, inf);
	exitPr_cost.assign(K + 1, inf);
	if (hu_hmm_finalize(K, Tmat_cost.data(), entryPr_cost.data(), exitPr_cost.data()) != HU_OK) throw runtime_error(hu_last_error());
	wingRetracted = true;
}

hu_hmm_params BandedHMMP7::params() const {
	hu_hmm_params p;
	p.K = K; p.L = L;
	p.eM = E_M_cost.data(); p.eI = E_I_cost.data(); p.T = Tmat_cost.data();
	p.entry = entryPr_cost.data(); p.exit = exitPr_cost.data(); p.p2cs = profile2CSIdx.data();
	return p;
}

BandedHMMP7::ViterbiAlignPath BandedHMMP7::buildAlignPath(const CSLoc& csLoc, int csFrom, int csTo) const {
	int start = 0, end = 0, from = 0, to = 0, nIns = 0, nDel = 0;
	int i = csFrom, j = csLoc.start;
	for (string::const_iterator it = csLoc.CS.begin(); it != csLoc.CS.end(); ++it) {
		const int k = getProfileLoc(j);
		const bool nonGap = DNA::isSymbol(*it);
		if (from == 0 && nonGap) from = i;
		if (nonGap) to = i;
		if (k != 0) {            // a non-D loc on profile
			if (start == 0) start = k;
			end = k;
			if (!nonGap) nDel++; // a deletion
		} else if (nonGap)
			nIns++;              // an insertion
		j++;
		if (nonGap) i++;
	}
	(void) csTo;
	return ViterbiAlignPath(start, end, from, to, nIns, nDel);
}

BandedHMMP7::HmmAlignment BandedHMMP7::buildGlobalAlign(const PrimarySeq& seq, const hu_align_row& row, const char* trace) const {
	HmmAlignment aln;
	aln.K = K; aln.L = L;
	aln.seqStart = row.seqStart; aln.seqEnd = row.seqEnd; aln.hmmStart = row.hmmStart; aln.hmmEnd = row.hmmEnd;
	aln.csStart = row.csStart; aln.csEnd = row.csEnd; aln.cost = row.cost;
	aln.align.resize(L);
	hu_hmm_params p = params();
	if (hu_trace_to_align(&p, seq.getSeq().c_str(), (int32_t) seq.length(), &row, trace, &aln.align[0]) != HU_OK)
		throw runtime_error(hu_last_error());
	return aln;
}

bool BandedHMMP7::HmmAlignment::isValid() const {
	return 0 < seqStart && seqStart <= seqEnd && 0 < hmmStart && hmmStart <= hmmEnd && hmmEnd <= K && 0 < csStart && csStart <= csEnd &&
			csEnd <= L && cost >= 0 && cost != inf && (size_t) L == align.length();
}
BandedHMMP7::HmmAlignment& BandedHMMP7::HmmAlignment::merge(const HmmAlignment& o) {
	if (!(K == o.K && L == o.L)) return *this;
	seqStart = std::min(seqStart, o.seqStart); seqEnd = std::max(seqEnd, o.seqEnd);
	hmmStart = std::min(hmmStart, o.hmmStart); hmmEnd = std::max(hmmEnd, o.hmmEnd);
	csStart = std::min(csStart, o.csStart); csEnd = std::max(csEnd, o.csEnd);
	cost += o.cost;
	for (int i = 0; i < L; ++i)
		if (align[i] == '.' && o.align[i] != '.') align[i] = o.align[i];
	return *this;
}
ostream& operator<<(ostream& out, const BandedHMMP7::HmmAlignment& aln) {
	out << aln.seqStart << "\t" << aln.seqEnd << "\t" << aln.hmmStart << "\t" << aln.hmmEnd << "\t" << aln.csStart << "\t" << aln.csEnd
			<< "\t" << aln.cost << "\t" << aln.align;
	return out;
}

// ------------------------------------------------------------------ substitution models ----
static const char* kModelNames[6] = {"GTR", "TN93", "HKY85", "F81", "K80", "JC69"};
DNASubModel::DNASubModel() {
	memset(&desc, 0, sizeof desc);
	desc.type = HU_GTR;
	for (int i = 0; i < 4; ++i) desc.pi[i] = 0.25;
	desc.kr = desc.ky = desc.kappa = desc.beta = 1;
}
string DNASubModel::modelType() const { return kModelNames[desc.type]; }
void DNASubModel::Pr(double v, double* P16) const {
	if (hu_model_Pr(&desc, v, P16) != HU_OK) throw runtime_error(hu_last_error());
}
istream& DNASubModel::read(istream& in) {
	string line, tag, value;
	while (in >> tag) {
		if (tag[0] == '#') { getline(in, line); continue; }
		if (tag == "Type:") {
			in >> value;
			if (value != modelType()) {
				cerr << "Unmatched Model Type! Trying to read in a " << value << " model into a " << modelType() << " object" << endl;
				in.setstate(ios_base::badbit);
				return in;
			}
			if (desc.type == HU_JC69) { getline(in, line); break; }
		} else if (tag == "pi:") for (int i = 0; i < 4; ++i) in >> desc.pi[i];
		else if (tag == "R:" && desc.type == HU_GTR) for (int i = 0; i < 16; ++i) in >> desc.R[i];
		else if (tag == "Q:" && desc.type == HU_GTR) { for (int i = 0; i <= 4; ++i) getline(in, line); break; }
		else if (tag == "kr:") in >> desc.kr;
		else if (tag == "ky:") in >> desc.ky;
		else if (tag == "kappa:") { in >> desc.kappa; if (desc.type == HU_K80) { getline(in, line); break; } }
		else if (tag == "beta:") { in >> desc.beta; getline(in, line); break; }
		else {
			cerr << "Un-recognized line found in " << modelType() << " Model input: tag: " << tag << endl;
			in.setstate(ios_base::badbit);
			return in;
		}
	}
	return in;
}
// Eigen's `m.format(IOFormat(FullPrecision))` as the reference's model writers use it: 15 significant digits
// (std::numeric_limits<double>::digits10), every coefficient right-aligned to the widest one, ' ' between
// columns and '\n' between rows; the stream's precision is restored afterwards (Eigen/src/Core/IO.h).
static void writeEigenFull(ostream& out, const double* v, int rows, int cols) {
	const streamsize oldPrec = out.precision(15);
	size_t width = 0;
	for (int i = 0; i < rows * cols; ++i) {
		ostringstream ss;
		ss.copyfmt(out);
		ss << v[i];
		width = max(width, ss.str().length());
	}
	for (int i = 0; i < rows; ++i) {
		for (int j = 0; j < cols; ++j) {
			if (j) out << " ";
			out.width((streamsize) width);
			out << v[i * cols + j];
		}
		if (i < rows - 1) out << "\n";
	}
	out.precision(oldPrec);
}

ostream& DNASubModel::write(ostream& out) const { // GTR.cpp:83-90, TN93.cpp:79-86, HKY85.cpp, F81.cpp, K80.cpp, JC69.cpp
	out << "# DNA Substitution Model" << endl << "Type: " << modelType() << endl;
	if (desc.type != HU_K80 && desc.type != HU_JC69) {
		out << "pi: ";
		writeEigenFull(out, desc.pi, 1, 4);
		out << endl;
	}
	if (desc.type == HU_GTR) {
		out << "R:" << endl;
		writeEigenFull(out, desc.R, 4, 4);
		out << endl;
		// Q (src/GTR.cpp:124-135): R * diag(pi) off the diagonal, rows summing to 0, scaled so that -trace = 1
		double Q[16], tr = 0;
		for (int i = 0; i < 4; ++i) {
			double rs = 0;
			for (int j = 0; j < 4; ++j) { Q[i * 4 + j] = i == j ? 0 : desc.R[i * 4 + j] * desc.pi[j]; rs += Q[i * 4 + j]; }
			Q[i * 4 + i] = -rs;
			tr += Q[i * 4 + i];
		}
		for (int i = 0; i < 16; ++i) Q[i] /= -tr;
		out << "Q:" << endl;
		writeEigenFull(out, Q, 4, 4);
		out << endl;
	} else {
		// the scalar parameters go through std::setprecision(DBL_DIG), which the reference leaves set on the stream
		out << setprecision(15);
		if (desc.type == HU_TN93) out << "kr: " << desc.kr << " ky: " << desc.ky << " beta: " << desc.beta << endl;
		else if (desc.type == HU_HKY85) out << "kappa: " << desc.kappa << " beta: " << desc.beta << endl;
		else if (desc.type == HU_F81) out << "beta: " << desc.beta << endl;
		else if (desc.type == HU_K80) out << "kappa: " << desc.kappa << endl;
	}
	return out;
}
istream& operator>>(istream& in, DNASubModel& m) { return m.read(in); }
ostream& operator<<(ostream& out, const DNASubModel& m) { return m.write(out); }

istream& DiscreteGammaModel::load(istream& in) {
	in.read((char*) &K, sizeof(int));
	in.read((char*) &alpha, sizeof(double));
	if (!in || K < 1 || K > HU_MAX_DG) { in.setstate(ios_base::badbit); return in; }
	b.resize(K + 1);
	r.resize(K);
	in.read((char*) b.data(), sizeof(double) * (K + 1));
	in.read((char*) r.data(), sizeof(double) * K);
	return in;
}
ostream& DiscreteGammaModel::save(ostream& out) const {
	out.write((const char*) &K, sizeof(int));
	out.write((const char*) &alpha, sizeof(double));
	out.write((const char*) b.data(), sizeof(double) * (K + 1));
	out.write((const char*) r.data(), sizeof(double) * K);
	return out;
}

// ------------------------------------------------------------------ PTUnrooted ----
const string PTUnrooted::PTPlacement::TSV_HEADER = "branch_id\tbranch_ratio\ttaxon_id\ttaxon_anno\tanno_dist\tloglik\tQ_placement\tQ_taxon";

istream& PTUnrooted::load(istream& in, bool keepMessages) {
	size_t nNodes = 0;
	in.read((char*) &nNodes, sizeof(size_t));
	in.read((char*) &csLen, sizeof(int));
	if (!in || nNodes < 2 || nNodes > ((size_t) 1 << 30) || csLen < 1 || csLen > 65535) { in.setstate(ios_base::badbit); return in; }
	parent.assign(nNodes, -1);
	blen.assign(nNodes, 0.0); height.assign(nNodes, 0.0); annoDist.assign(nNodes, 0.0);
	name.resize(nNodes); anno.resize(nNodes); seqName.resize(nNodes); neighbors.assign(nNodes, vector<long>());
	seq.assign(nNodes * (size_t) csLen, -2);
	for (size_t i = 0; i < nNodes; ++i) { // PTUNode::load, cpp:99-114
		long id;
		in.read((char*) &id, sizeof(long));
		if (!in || id < 0 || (size_t) id >= nNodes) { in.setstate(ios_base::badbit); return in; }
		loadString(name[id], in);
		bool hasAbc;
		in.read((char*) &hasAbc, sizeof(bool));
		string tmp, codes;
		if (hasAbc) loadString(tmp, in);
		loadString(seqName[id], in); // DigitalSeq name
		loadString(codes, in);       // int8 codes
		if (!codes.empty()) {
			if (codes.size() != (size_t) csLen) { in.setstate(ios_base::badbit); return in; }
			memcpy(&seq[(size_t) id * csLen], codes.data(), csLen);
		}
		loadString(anno[id], in);
		in.read((char*) &annoDist[id], sizeof(double));
	}
	size_t nEdges = 0;
	in.read((char*) &nEdges, sizeof(size_t));
	if (!in || nEdges != 2 * (nNodes - 1)) { in.setstate(ios_base::badbit); return in; }
	const size_t per = (size_t) csLen * 4;
	// keepMessages = false: the 4 x csLen doubles of every directed edge (98 GB for a gg_97-sized database) are skipped here
	// and handed to the device(s) edge by edge later (streamMessages); only the records around them are read
	up.clear(); dn.clear();
	if (keepMessages) {
		up.assign(nNodes * per, 0.0);
		dn.assign(nNodes * per, 0.0);
	}
	edgePos = in.tellg();
	vector<double> buf(keepMessages ? per : 0);
	vector<pair<long, long> > edges(nEdges);
	vector<vector<double> > pending; // messages whose direction is known only once every parent flag is read
	pending.reserve(nEdges);
	vector<double> lens(nEdges);
	vector<char> flags(nEdges);
	for (size_t e = 0; e < nEdges; ++e) { // loadEdge, cpp:595-619 + PTUBranch::load, cpp:1068-1083
		long id1, id2;
		bool isPar;
		double length;
		size_t N;
		in.read((char*) &id1, sizeof(long));
		in.read((char*) &id2, sizeof(long));
		in.read((char*) &isPar, sizeof(bool));
		in.read((char*) &length, sizeof(double));
		in.read((char*) &N, sizeof(size_t));
		if (!in || id1 < 0 || id2 < 0 || (size_t) id1 >= nNodes || (size_t) id2 >= nNodes || N != per) { in.setstate(ios_base::badbit); return in; }
		if (keepMessages) in.read((char*) buf.data(), sizeof(double) * per);
		else in.seekg((streamoff) (sizeof(double) * per), ios_base::cur);
		neighbors[id1].push_back(id2);
		if (isPar) { // id1 is the parent of id2: this record is msg(parent -> child) = dn[id2]
			parent[id2] = (int32_t) id1;
			blen[id2] = length;
			if (keepMessages) memcpy(&dn[(size_t) id2 * per], buf.data(), sizeof(double) * per);
		} else if (keepMessages) { // msg(child id1 -> parent id2) = up[id1]
			memcpy(&up[(size_t) id1 * per], buf.data(), sizeof(double) * per);
		}
	}
	in.read((char*) &root, sizeof(long));
	rootLoglik.resize(per);
	in.read((char*) rootLoglik.data(), sizeof(double) * per);
	for (size_t i = 0; i < nNodes; ++i) {
		long id;
		double h;
		in.read((char*) &id, sizeof(long));
		in.read((char*) &h, sizeof(double));
		if (!in || id < 0 || (size_t) id >= nNodes) { in.setstate(ios_base::badbit); return in; }
		height[id] = h;
	}
	unsigned nIdx = 0;
	in.read((char*) &nIdx, sizeof(unsigned));
	msaIndex.resize(nIdx);
	for (unsigned i = 0; i < nIdx; ++i) {
		in.read((char*) &msaIndex[i].first, sizeof(unsigned));
		in.read((char*) &msaIndex[i].second, sizeof(long));
	}
	string type;
	in >> type;
	in.ignore();
	int t = -1;
	for (int i = 0; i < 6; ++i) if (type == kModelNames[i]) t = i;
	if (!in || t < 0) { in.setstate(ios_base::badbit); return in; }
	model.desc.type = t;
	in >> model;
	in.read((char*) &hasDG, sizeof(bool));
	if (hasDG) dG.load(in);
	if (!in || parent[root] != -1) { in.setstate(ios_base::badbit); return in; }
	// annotation strings -> ids (equal getTaxonName() <=> equal id)
	map<string, int32_t> ids;
	taxon.resize(nNodes);
	for (size_t i = 0; i < nNodes; ++i) {
		map<string, int32_t>::iterator it = ids.find(anno[i]);
		if (it == ids.end()) it = ids.insert(make_pair(anno[i], (int32_t) ids.size())).first;
		taxon[i] = it->second;
	}
	return in;
}

ostream& PTUnrooted::save(ostream& out) const {
	const size_t nNodes = numNodes(), per = (size_t) csLen * 4;
	out.write((const char*) &nNodes, sizeof(size_t));
	out.write((const char*) &csLen, sizeof(int));
	for (size_t i = 0; i < nNodes; ++i) {
		const long id = (long) i;
		out.write((const char*) &id, sizeof(long));
		saveString(name[i], out);
		const bool hasAbc = false; // PTUNode::save -> DigitalSeq::save(out) with withAbc = false (src/DigitalSeq.h)
		out.write((const char*) &hasAbc, sizeof(bool));
		saveString(i < seqName.size() ? seqName[i] : string(), out);
		saveString(string((const char*) &seq[i * (size_t) csLen], csLen), out);
		saveString(anno[i], out);
		out.write((const char*) &annoDist[i], sizeof(double));
	}
	const size_t nEdges = 2 * (nNodes - 1);
	out.write((const char*) &nEdges, sizeof(size_t));
	for (size_t u = 0; u < nNodes; ++u)
		for (size_t k = 0; k < neighbors[u].size(); ++k) {
			const long id1 = (long) u, id2 = neighbors[u][k];
			const bool isPar = parent[id2] == (int32_t) id1;
			const double length = isPar ? blen[id2] : blen[id1];
			out.write((const char*) &id1, sizeof(long));
			out.write((const char*) &id2, sizeof(long));
			out.write((const char*) &isPar, sizeof(bool));
			out.write((const char*) &length, sizeof(double));
			out.write((const char*) &per, sizeof(size_t));
			out.write((const char*) (isPar ? &dn[(size_t) id2 * per] : &up[(size_t) id1 * per]), sizeof(double) * per);
		}
	out.write((const char*) &root, sizeof(long));
	vector<double> rl(rootLoglik);
	rl.resize(per, 0.0);
	out.write((const char*) rl.data(), sizeof(double) * per);
	for (size_t i = 0; i < nNodes; ++i) {
		const long id = (long) i;
		out.write((const char*) &id, sizeof(long));
		out.write((const char*) &height[i], sizeof(double));
	}
	const unsigned nIdx = (unsigned) msaIndex.size();
	out.write((const char*) &nIdx, sizeof(unsigned));
	for (unsigned i = 0; i < nIdx; ++i) {
		out.write((const char*) &msaIndex[i].first, sizeof(unsigned));
		out.write((const char*) &msaIndex[i].second, sizeof(long));
	}
	out << model.modelType() << endl << model;
	out.write((const char*) &hasDG, sizeof(bool));
	if (hasDG) dG.save(out);
	return out;
}

hu_tree_desc PTUnrooted::treeDesc() const {
	hu_tree_desc td;
	td.nNodes = (int64_t) numNodes(); td.csLen = csLen; td.root = (int32_t) root;
	td.parent = parent.data(); td.blen = blen.data(); td.height = height.data(); td.seq = seq.data(); td.taxon = taxon.data();
	td.up = up.empty() ? NULL : up.data(); td.dn = dn.empty() ? NULL : dn.data(); td.inferInternalSeq = 0;
	return td;
}
hu_model_desc PTUnrooted::modelDesc() const {
	hu_model_desc md = model.desc;
	md.dgK = hasDG ? dG.K : 0;
	for (int k = 0; k < md.dgK; ++k) md.dgRates[k] = dG.r[k];
	return md;
}
// Second pass over the edge records of a .ptu read with keepMessages = false: every message goes to `sink` as
// (node, direction, 4 x csLen log-domain doubles); direction 0 = msg(node -> parent), 1 = msg(parent -> node).
bool PTUnrooted::streamMessages(istream& in, MessageSink& sink) const {
	const size_t nNodes = parent.size(), nEdges = 2 * (nNodes - 1), per = (size_t) csLen * 4;
	in.clear();
	in.seekg(edgePos);
	vector<double> buf(per);
	for (size_t e = 0; e < nEdges; ++e) {
		long id1, id2;
		bool isPar;
		double length;
		size_t N;
		in.read((char*) &id1, sizeof(long));
		in.read((char*) &id2, sizeof(long));
		in.read((char*) &isPar, sizeof(bool));
		in.read((char*) &length, sizeof(double));
		in.read((char*) &N, sizeof(size_t));
		if (!in || N != per) return false;
		in.read((char*) buf.data(), sizeof(double) * per);
		if (!in) return false;
		if (!sink.put(isPar ? id2 : id1, isPar ? 1 : 0, buf.data())) return false;
	}
	return sink.flush();
}

PTUnrooted::PTPlacement PTUnrooted::toPlacement(const hu_place_row& r) const {
	PTPlacement p;
	p.start = r.start; p.end = r.end; p.cNode = r.cNode; p.pNode = r.pNode; p.aNode = r.aNode;
	p.wuv = r.wuv; p.ratio = r.ratio; p.wnr = r.wnr; p.loglik = r.loglik; p.height = r.height;
	p.annoDist = r.annoDist; p.qPlace = r.qPlace; p.qTaxon = r.qTaxon;
	p.taxonName = r.aNode >= 0 ? anno[r.aNode] : "NULL";
	p.raw = r;
	p.owner = this;
	return p;
}
PTUnrooted::PTPlacement::PTPlacement() : start(0), end(0), cNode(-1), pNode(-1), aNode(-1), wuv(0), ratio(0), wnr(0), loglik(0), height(0),
		annoDist(0), qPlace(0), qTaxon(0), owner(NULL) {
	memset(&raw, 0, sizeof raw);
	raw.cNode = raw.pNode = raw.aNode = raw.taxon = -1;
}
hu_place_row PTUnrooted::PTPlacement::toRow() const {
	hu_place_row r = raw;
	r.start = start; r.end = end; r.cNode = (int32_t) cNode; r.pNode = (int32_t) pNode; r.aNode = (int32_t) aNode;
	r.wuv = wuv; r.ratio = ratio; r.wnr = wnr; r.loglik = loglik; r.height = height;
	r.annoDist = annoDist; r.qPlace = qPlace; r.qTaxon = qTaxon;
	return r;
}
string PTUnrooted::PTPlacement::getId() const {
	if (cNode < 0 || pNode < 0) return "NULL";
	ostringstream os;
	os << cNode << "->" << pNode;
	return os.str();
}
ostream& operator<<(ostream& out, const PTUnrooted::PTPlacement& p) {
	out << p.getId() << "\t" << p.ratio << "\t" << p.getTaxonId() << "\t" << p.taxonName << "\t" << p.annoDist << "\t" << p.loglik
			<< "\t" << p.qPlace << "\t" << p.qTaxon;
	return out;
}

// ------------------------------------------------------------------ seed search ----
vector<BandedHMMP7::ViterbiAlignPath> findSeeds(const BandedHMMP7& hmm, const CSFMIndex& csfm, const PrimarySeq& read,
		int seedLen, int seedRegion, BandedHMMP7::align_mode mode) {
	vector<BandedHMMP7::ViterbiAlignPath> vpaths;
	const int N = (int) read.length();
	const int regionLen = seedRegion < N ? seedRegion : N;
	for (int seedFrom = 0; seedFrom + seedLen - 1 < regionLen; ++seedFrom) { // 5' seed
		const int seedTo = seedFrom + seedLen - 1;
		const CSLoc loc = csfm.locateOne(read.getSeq().substr(seedFrom, seedLen));
		if (loc.isValid()) {
			const BandedHMMP7::ViterbiAlignPath vp = hmm.buildAlignPath(loc, seedFrom + 1, seedTo + 1);
			if (vp.isValid()) { vpaths.push_back(vp); break; }
		}
	}
	if (mode == BandedHMMP7::GLOBAL && (vpaths.empty() || N >= 2 * regionLen)) { // 3' seed
		for (int seedTo = N - 1; seedTo - seedLen + 1 >= N - regionLen; --seedTo) {
			const int seedFrom = seedTo - seedLen + 1;
			if (seedFrom < 0) break;
			const CSLoc loc = csfm.locateOne(read.getSeq().substr(seedFrom, seedLen));
			if (loc.isValid()) {
				const BandedHMMP7::ViterbiAlignPath vp = hmm.buildAlignPath(loc, seedFrom + 1, seedTo + 1);
				if (vp.isValid()) { vpaths.push_back(vp); break; }
			}
		}
	}
	return vpaths;
}

namespace {
// Phase 1 of findSeedsBatch: the intervals of the attempts at each end of a read, in attempt order, (1, 0) = empty interval.
// They live in ONE flat array that the calling thread owns and keeps between batches (maxAtt slots per read and end): the
// worker threads allocate nothing and touch no fresh pages -- with a vector per read the page faults and heap growth of eight
// threads serialised in the kernel (one address space), and eight threads were no faster than one.
typedef pair<int32_t, int32_t> SaRange;
struct SeedIntervals {
	const SaRange* at5; const SaRange* at3;
	int n5, n3;
};
void speculate(const CSFMIndex& csfm, const PrimarySeq& read, int seedLen, int seedRegion, BandedHMMP7::align_mode mode, SaRange* at5, SaRange* at3,
		int maxAtt, SeedIntervals& si) {
	const int N = (int) read.length();
	const int regionLen = seedRegion < N ? seedRegion : N;
	int32_t s, e;
	si.at5 = at5; si.at3 = at3; si.n5 = si.n3 = 0;
	for (int seedFrom = 0; seedFrom + seedLen - 1 < regionLen && si.n5 < maxAtt; ++seedFrom) {
		const bool hit = csfm.interval(read.getSeq().data() + seedFrom, (size_t) seedLen, s, e);
		at5[si.n5++] = hit ? make_pair(s, e) : make_pair((int32_t) 1, (int32_t) 0);
		if (hit) break; // nearly always the seed that is kept
	}
	if (mode != BandedHMMP7::GLOBAL) return;
	for (int seedTo = N - 1; seedTo - seedLen + 1 >= N - regionLen && si.n3 < maxAtt; --seedTo) {
		const int seedFrom = seedTo - seedLen + 1;
		if (seedFrom < 0) break;
		const bool hit = csfm.interval(read.getSeq().data() + seedFrom, (size_t) seedLen, s, e);
		at3[si.n3++] = hit ? make_pair(s, e) : make_pair((int32_t) 1, (int32_t) 0);
		if (hit) break;
	}
}
// locateOne() with the interval taken from phase 1 when it has it
CSLoc locateCached(const CSFMIndex& csfm, const string& seq, int from, int len, const SaRange* cache, int nCached, size_t attempt) {
	int32_t s, e;
	if (attempt < (size_t) nCached) { s = cache[attempt].first; e = cache[attempt].second; }
	else if (!csfm.interval(seq.data() + from, (size_t) len, s, e)) return CSLoc();
	if (s > e) return CSLoc();
	return csfm.locateAt(seq.substr(from, len), s, e); // the pattern as a string only for the attempt that hits
}
struct SpecArgs {
	const CSFMIndex* csfm; const vector<const PrimarySeq*>* reads; SeedIntervals* out; SaRange* flat;
	int seedLen, seedRegion, maxAtt; BandedHMMP7::align_mode mode; size_t n;
	std::atomic<size_t>* next; // reads are handed out 64 at a time: a thread that loses its core holds up 64 reads, not an eighth of the batch
};
void* specThread(void* p) {
	const SpecArgs* a = (const SpecArgs*) p;
	for (;;) {
		const size_t b = a->next->fetch_add(64);
		if (b >= a->n) break;
		const size_t e = b + 64 < a->n ? b + 64 : a->n;
		for (size_t i = b; i < e; ++i)
			speculate(*a->csfm, *(*a->reads)[i], a->seedLen, a->seedRegion, a->mode, a->flat + 2 * i * (size_t) a->maxAtt, a->flat + (2 * i + 1) * (size_t) a->maxAtt,
			          a->maxAtt, a->out[i]);
	}
	return NULL;
}

// Worker threads that live as long as the process: a batch is some tens of milliseconds of work, and threads created per batch
// spend that long being created, scheduled and spread over the cores (measured: eight fresh threads were no faster than one).
class SpecPool {
public:
	SpecPool() : job(NULL), gen(0), pending(0) { pthread_mutex_init(&mu, NULL); pthread_cond_init(&work, NULL); pthread_cond_init(&done, NULL); pthread_mutex_init(&callers, NULL); }
	// runs specThread(&a) on nThreads threads (the caller is one of them) and returns when all are through
	void run(const SpecArgs& a, int nThreads) {
		pthread_mutex_lock(&callers); // one batch at a time
		pthread_mutex_lock(&mu);
		while ((int) th.size() < nThreads - 1) {
			pthread_t t;
			Start* st = new Start; st->pool = this; st->seen = gen;
			if (pthread_create(&t, NULL, loop, st) != 0) { delete st; break; }
			pthread_detach(t);
			th.push_back(t);
		}
		const int helpers = min((int) th.size(), nThreads - 1);
		job = &a; quota = helpers; pending = helpers; gen++;
		pthread_cond_broadcast(&work);
		pthread_mutex_unlock(&mu);
		specThread((void*) &a);
		pthread_mutex_lock(&mu);
		while (pending > 0) pthread_cond_wait(&done, &mu);
		job = NULL;
		pthread_mutex_unlock(&mu);
		pthread_mutex_unlock(&callers);
	}
private:
	struct Start { SpecPool* pool; unsigned long seen; };
	static void* loop(void* p) {
		Start* st = (Start*) p;
		SpecPool* self = st->pool;
		unsigned long seen = st->seen;
		delete st;
		for (;;) {
			pthread_mutex_lock(&self->mu);
			while (self->gen == seen || self->quota <= 0) { if (self->gen != seen) seen = self->gen; pthread_cond_wait(&self->work, &self->mu); }
			seen = self->gen;
			self->quota--; // this thread takes one of the batch's helper slots
			const SpecArgs* a = self->job;
			pthread_mutex_unlock(&self->mu);
			specThread((void*) a);
			pthread_mutex_lock(&self->mu);
			if (--self->pending == 0) pthread_cond_signal(&self->done);
			pthread_mutex_unlock(&self->mu);
		}
		return NULL;
	}
	pthread_mutex_t mu, callers;
	pthread_cond_t work, done;
	vector<pthread_t> th;
	const SpecArgs* job;
	unsigned long gen;
	int quota, pending;
};
SpecPool& specPool() { static SpecPool* p = new SpecPool; return *p; } // never destroyed: its threads outlive main()'s statics
}

vector<vector<BandedHMMP7::ViterbiAlignPath> > findSeedsBatch(const BandedHMMP7& hmm, const CSFMIndex& csfm,
		const vector<const PrimarySeq*>& reads, int seedLen, int seedRegion, BandedHMMP7::align_mode mode, int nThreads) {
	const size_t n = reads.size();
	const int maxAtt = max(1, seedRegion - seedLen + 1); // attempts per end: one per seed start inside the region
	static thread_local vector<SaRange> flat;            // kept (and its pages with it) from batch to batch
	static thread_local vector<SeedIntervals> spec;
	if (flat.size() < 2 * n * (size_t) maxAtt) flat.resize(2 * n * (size_t) maxAtt);
	if (spec.size() < n) spec.resize(n);
	// phase 1: no random draws, any order
	nThreads = max(1, min(nThreads, (int) ((n + 63) / 64)));
	std::atomic<size_t> next(0);
	const SpecArgs job = {&csfm, &reads, spec.data(), flat.data(), seedLen, seedRegion, maxAtt, mode, n, &next};
	if (nThreads > 1) specPool().run(job, nThreads);
	else specThread((void*) &job);
	// phase 2: findSeeds' control flow, read by read, rand() drawn as the one-by-one loop draws it
	vector<vector<BandedHMMP7::ViterbiAlignPath> > out(n);
	for (size_t r = 0; r < n; ++r) {
		const PrimarySeq& read = *reads[r];
		vector<BandedHMMP7::ViterbiAlignPath>& vpaths = out[r];
		const int N = (int) read.length();
		const int regionLen = seedRegion < N ? seedRegion : N;
		size_t attempt = 0;
		for (int seedFrom = 0; seedFrom + seedLen - 1 < regionLen; ++seedFrom, ++attempt) { // 5' seed
			const int seedTo = seedFrom + seedLen - 1;
			const CSLoc loc = locateCached(csfm, read.getSeq(), seedFrom, seedLen, spec[r].at5, spec[r].n5, attempt);
			if (loc.isValid()) {
				const BandedHMMP7::ViterbiAlignPath vp = hmm.buildAlignPath(loc, seedFrom + 1, seedTo + 1);
				if (vp.isValid()) { vpaths.push_back(vp); break; }
			}
		}
		if (mode == BandedHMMP7::GLOBAL && (vpaths.empty() || N >= 2 * regionLen)) { // 3' seed
			attempt = 0;
			for (int seedTo = N - 1; seedTo - seedLen + 1 >= N - regionLen; --seedTo, ++attempt) {
				const int seedFrom = seedTo - seedLen + 1;
				if (seedFrom < 0) break;
				const CSLoc loc = locateCached(csfm, read.getSeq(), seedFrom, seedLen, spec[r].at3, spec[r].n3, attempt);
				if (loc.isValid()) {
					const BandedHMMP7::ViterbiAlignPath vp = hmm.buildAlignPath(loc, seedFrom + 1, seedTo + 1);
					if (vp.isValid()) { vpaths.push_back(vp); break; }
				}
			}
		}
	}
	return out;
}

} // namespace HmmUFOtu
} // namespace EGriceLab
