// Host half of the database builder (the reference's hmmufotu-build, src/hmmufotu-build.cpp:308-392): the alignment
// (FASTA -> pruned MSA with counts, position-based sequence weights and consensus) and the tree skeleton (Newick ->
// node ids, edges, leaf sequences, taxonomy names, annotations, heights).  Everything here is integer / string work
// plus a few sums of doubles, decided before any likelihood is evaluated; the all-roots messages and the inferred
// internal sequences are evaluated on the device by hu_db_create() (hu_tree_desc.up == NULL).  Checked against the
// files the reference's own builder wrote (tests/test_host.py: <db>.msa byte for byte, <db>.ptu field by field).
#include "hu_host.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <stdio.h>
#include <algorithm>
#include <sstream>

using namespace std;

namespace EGriceLab {
namespace HmmUFOtu {

// ------------------------------------------------------------------ MSA ----
long MSA::loadFasta(istream& in) { // src/MSA.cpp:140-164
	alphabet = "dna";
	numSeq = 0; csLen = 0; isPruned = false;
	seqNames.clear(); concatMSA.clear(); CS.clear(); auxBlocks.clear();
	SeqIO io(&in, "fasta");
	while (io.hasNext()) {
		const PrimarySeq seq = io.nextSeq();
		if (csLen != 0 && seq.length() != csLen) return -1; // not an alignment
		csLen = (unsigned) seq.length();
		numSeq++;
		seqNames.push_back(seq.getId());
		concatMSA.append(seq.getSeq()); // case preserved
	}
	updateCounts();
	calculateCS();
	return numSeq;
}

void MSA::updateCounts() {
	// updateRawCounts, src/MSA.cpp:226-254
	startIdx.assign(numSeq, 0); endIdx.assign(numSeq, 0); lenIdx.assign(numSeq, 0);
	resCount.assign((size_t) 4 * csLen, 0); gapCount.assign(csLen, 0);
	for (unsigned i = 0; i < numSeq; ++i) {
		int start = -1, end = -1, len = 0;
		for (unsigned j = 0; j < csLen; ++j) {
			const char c = (char) ::toupper(residualAt(i, j));
			const int8_t b = DNA::encode(c);
			if (b >= 0) {
				if (start == -1) start = (int) j;
				if ((int) j > end) end = (int) j;
				len++;
				resCount[(size_t) j * 4 + b]++;
			} else if (DNA::isGap(c))
				gapCount[j]++;
		}
		startIdx[i] = start; endIdx[i] = end; lenIdx[i] = len;
	}
	// updateSeqWeight, src/MSA.cpp:256-278: position-based weights, 1 / (#distinct residues x count of this residue)
	// summed over the row's residues in column order, divided by the row's length, rescaled to sum = numSeq
	vector<int> pssw((size_t) 4 * csLen);
	for (unsigned j = 0; j < csLen; ++j) {
		int distinct = 0;
		for (int b = 0; b < 4; ++b) distinct += resCount[(size_t) j * 4 + b] != 0;
		for (int b = 0; b < 4; ++b) pssw[(size_t) j * 4 + b] = distinct * resCount[(size_t) j * 4 + b];
	}
	seqWeight.assign(numSeq, 0.0);
	double total = 0;
	for (unsigned i = 0; i < numSeq; ++i) {
		double w = 0;
		for (unsigned j = 0; j < csLen; ++j) {
			const int8_t b = DNA::encode((char) ::toupper(residualAt(i, j)));
			if (b >= 0) w += 1.0 / pssw[(size_t) j * 4 + b];
		}
		if (lenIdx[i] > 0) w /= lenIdx[i];
		seqWeight[i] = w;
		total += w;
	}
	const double scale = numSeq / total;
	for (unsigned i = 0; i < numSeq; ++i) seqWeight[i] *= scale;
	// updateWeightedCounts, src/MSA.cpp:280-293 (rows outermost, so a column's sum runs over rows in file order)
	resWCount.assign((size_t) 4 * csLen, 0.0); gapWCount.assign(csLen, 0.0);
	for (unsigned i = 0; i < numSeq; ++i)
		for (unsigned j = 0; j < csLen; ++j) {
			const char c = (char) ::toupper(residualAt(i, j));
			const int8_t b = DNA::encode(c);
			if (b >= 0) resWCount[(size_t) j * 4 + b] += seqWeight[i];
			else if (DNA::isGap(c)) gapWCount[j] += seqWeight[i];
		}
}

void MSA::calculateCS() { // src/MSA.cpp:211-226
	if (CS.length() == csLen) return;
	CS.clear();
	for (unsigned j = 0; j < csLen; ++j) {
		int best = 0; // first maximum, as Eigen's maxCoeff(&row) visits a column
		for (int b = 1; b < 4; ++b) if (resWCount[(size_t) j * 4 + b] > resWCount[(size_t) j * 4 + best]) best = b;
		CS.push_back(resWCount[(size_t) j * 4 + best] >= gapWCount[j] ? DNA::decode((int8_t) best) : '-');
	}
}

MSA& MSA::prune() { // src/MSA.cpp:87-138
	if (isPruned || numSeq == 0) return *this;
	if (resCount.size() != (size_t) 4 * csLen) updateCounts();
	vector<char> drop(csLen, 0);
	unsigned nDrop = 0;
	for (unsigned j = 0; j < csLen; ++j) {
		int n = 0;
		for (int b = 0; b < 4; ++b) n += resCount[(size_t) j * 4 + b];
		if (n == 0) { drop[j] = 1; nDrop++; }
	}
	if (nDrop == 0) return *this; // nothing to do: isPruned stays false, as in the reference
	string pruned;
	pruned.reserve(concatMSA.length() - (size_t) numSeq * nDrop);
	for (size_t i = 0; i < concatMSA.length(); ++i)
		if (!drop[i % csLen]) pruned.push_back(concatMSA[i]);
	concatMSA.swap(pruned);
	if (!CS.empty()) {
		string cs;
		for (size_t j = 0; j < CS.length(); ++j)
			if (!drop[j]) cs.push_back(CS[j]);
		CS.swap(cs);
	}
	csLen -= nDrop;
	updateCounts();
	isPruned = true;
	return *this;
}

// ------------------------------------------------------------------ <db>.csfm in libcds' serialisation ----
namespace {
// libcds packs fields LSB first into 32-bit words (libcdsBasics.h: set_field / set_var_field)
struct BitWords {
	vector<uint32_t> w;
	explicit BitWords(size_t nWords) : w(nWords, 0u) {}
	void put(size_t pos, unsigned len, uint32_t x) { // bits pos .. pos+len-1 <- x
		for (unsigned b = 0; b < len; ++b)
			if ((x >> b) & 1u) w[(pos + b) / 32] |= 1u << ((pos + b) % 32);
	}
	void setBit(size_t pos) { w[pos / 32] |= 1u << (pos % 32); }
	bool bit(size_t pos) const { return (w[pos / 32] >> (pos % 32)) & 1u; }
};
unsigned bitsOf(uint32_t n) { unsigned b = 0; while (n) { b++; n >>= 1; } return b; }
size_t uintLen(uint64_t e, uint64_t n) { return (size_t) ((e * n + 31) / 32); }

// table_offset(15) of libcds (TableOffsetRRR.cpp): a 15-bit block is stored as (class = popcount, offset = its index in
// the enumeration of that class: positions chosen in increasing order, lowest first position first); an offset takes
// bits(C(15, class) - 1) bits
struct RrrTable {
	enum { U = 15 };
	vector<uint16_t> offsetOf; // [block value]
	unsigned logBinom[U + 1];
	unsigned next;
	void gen(unsigned cls, unsigned placed, unsigned from, uint32_t value) {
		if (cls == placed) { offsetOf[value] = (uint16_t) next++; return; }
		for (unsigned i = from; i < U; ++i) gen(cls, placed + 1, i + 1, value | (1u << i));
	}
	RrrTable() : offsetOf(1u << U, 0) {
		uint32_t binom[U + 1][U + 1];
		for (int i = 0; i <= U; ++i)
			for (int j = 0; j <= U; ++j) binom[i][j] = j == 0 || j == i ? 1 : (j > i ? 0 : binom[i - 1][j - 1] + binom[i - 1][j]);
		for (unsigned c = 0; c <= U; ++c) {
			logBinom[c] = (c == 0 || c == U) ? 0 : bitsOf(binom[U][c] - 1);
			next = 0;
			gen(c, 0, 0, 0);
		}
	}
};
}
namespace {
// BitSequenceRRR::build + save (BitSequenceRRR.cpp:51-87, 379-402): classes in C (4 bits each), offsets in O; the
// samplings are rebuilt by the loader
void saveRRR(ostream& out, const BitWords& bm, size_t len, uint32_t sampleRate, const RrrTable& E) {
	const uint32_t cLen = (uint32_t) (len / 15 + (len % 15 != 0)), cFieldBits = bitsOf(15);
	BitWords C(uintLen(cLen, cFieldBits));
	vector<uint16_t> blocks(cLen);
	uint64_t ones = 0;
	uint32_t oBits = 0;
	for (uint32_t i = 0; i < cLen; ++i) {
		uint32_t v = 0;
		const size_t lo = (size_t) i * 15, hi = min(len - 1, (size_t) (i + 1) * 15 - 1);
		for (size_t p = lo; p <= hi; ++p) if (bm.bit(p)) v |= 1u << (p - lo);
		blocks[i] = (uint16_t) v;
		const unsigned cls = (unsigned) __builtin_popcount(v);
		C.put((size_t) i * cFieldBits, cFieldBits, cls);
		ones += cls;
		oBits += E.logBinom[cls];
	}
	const uint32_t oLen = (uint32_t) uintLen(1, oBits);
	BitWords O(oLen);
	size_t oPos = 0;
	for (uint32_t i = 0; i < cLen; ++i) {
		const unsigned nb = E.logBinom[__builtin_popcount(blocks[i])];
		O.put(oPos, nb, E.offsetOf[blocks[i]]);
		oPos += nb;
	}
	const uint32_t hdr = 2; // RRR02_HDR
	const uint64_t length = len;
	out.write((const char*) &hdr, 4);
	out.write((const char*) &length, 8);
	out.write((const char*) &ones, 8);
	out.write((const char*) &cLen, 4);
	out.write((const char*) &cFieldBits, 4);
	out.write((const char*) &oLen, 4);
	out.write((const char*) &oBits, 4);
	out.write((const char*) &sampleRate, 4);
	out.write((const char*) C.w.data(), (streamsize) (C.w.size() * 4));
	out.write((const char*) O.w.data(), (streamsize) (O.w.size() * 4));
}
}

ostream& CSFMIndex::save(ostream& out, const MSA& msa) const {
	const int32_t N = concatLen + 1;
	saveString("IUPACNucl", out);                 // abc->getName()
	const char gapCh = '-';
	out.write(&gapCh, 1);
	const uint16_t cs16 = (uint16_t) csLen;
	out.write((const char*) &cs16, 2);
	out.write((const char*) &concatLen, 4);
	int32_t C256[256];
	memset(C256, 0, sizeof C256);
	for (int b = 0; b < 6; ++b) C256[b] = C[b];
	out.write((const char*) C256, sizeof C256);
	saveString(" " + msa.CS, out);                // csSeq: dummy position 0
	vector<double> ident(csLen + 1, 0.0);         // identityAt(j) = largest raw residue count / numSeq
	for (int j = 0; j < csLen; ++j) {
		int cnt[4] = {0, 0, 0, 0};
		for (unsigned i = 0; i < msa.numSeq; ++i) {
			const int8_t b = DNA::encode((char) ::toupper(msa.residualAt(i, j)));
			if (b >= 0) cnt[b]++;
		}
		ident[j + 1] = max(max(cnt[0], cnt[1]), max(cnt[2], cnt[3])) / (double) msa.numSeq;
	}
	out.write((const char*) ident.data(), (streamsize) (ident.size() * sizeof(double)));
	// concat2CS of the file: 0-based text positions (this class keeps them 1-based with a dummy entry in front)
	out.write((const char*) (concat2CS.data() + 1), (streamsize) ((size_t) N * sizeof(uint16_t)));
	// every 4th text position is sampled (SA_SAMPLE_RATE), in suffix-array order; a bitmap marks the sampled ranks
	vector<uint32_t> sampled;
	BitWords B((size_t) N / 32 + 1);
	for (int32_t i = 0; i < N; ++i) {
		const int32_t v = sa[i] - 1; // 0-based suffix position
		if (v % 4 == 0) { sampled.push_back((uint32_t) v); B.setBit(i); }
	}
	sampled.resize((size_t) N / 4 + 1, 0u);
	out.write((const char*) sampled.data(), (streamsize) ((size_t) (concatLen / 4) * sizeof(uint32_t)));
	const RrrTable E;
	saveRRR(out, B, (size_t) N, 8, E);            // saIdx, RRR_SAMPLE_RATE
	// WaveletTreeNoptrs over the BWT symbols 0..4 (WaveletTreeNoptrs.cpp:157-225, 245-257), MapperNone, RRR levels
	vector<uint8_t> sym(N);
	uint32_t maxV = 0;
	for (int32_t i = 0; i < N; ++i) {
		int b = 0;
		for (int q = 0; q < 5; ++q) if ((rank[(size_t) i / 64].mask[q] >> (i % 64)) & 1ull) b = q;
		sym[i] = (uint8_t) b;
		maxV = max(maxV, (uint32_t) b);
	}
	const uint32_t height = bitsOf(maxV);
	vector<uint32_t> OCC(maxV + 2, 0u);
	for (int32_t i = 0; i < N; ++i) OCC[sym[i] + 1]++;
	for (uint32_t i = 1; i <= maxV + 1; ++i)     // symbols that never occur are appended once
		if (OCC[i] == 0) { OCC[i]++; sym.push_back((uint8_t) (i - 1)); }
	const uint64_t newN = sym.size(), length = (uint64_t) N;
	for (uint32_t i = 1; i <= maxV + 1; ++i) OCC[i] += OCC[i - 1];
	vector<BitWords> levels(height, BitWords((size_t) newN / 32 + 1));
	{ // build_level: level by level, a stable split of every node's range by the level's bit (most significant first)
		vector<uint8_t> cur(sym), nxt(sym.size());
		vector<pair<size_t, size_t> > ranges(1, make_pair((size_t) 0, (size_t) newN)), nextRanges;
		for (uint32_t level = 0; level < height; ++level) {
			const unsigned bitIx = height - level - 1;
			nextRanges.clear();
			for (size_t r = 0; r < ranges.size(); ++r) {
				const size_t off = ranges[r].first, len = ranges[r].second;
				size_t nl = 0;
				for (size_t i = 0; i < len; ++i) nl += !((cur[off + i] >> bitIx) & 1u);
				size_t l = off, rr = off + nl;
				for (size_t i = 0; i < len; ++i) {
					if ((cur[off + i] >> bitIx) & 1u) { levels[level].setBit(off + i); nxt[rr++] = cur[off + i]; }
					else nxt[l++] = cur[off + i];
				}
				nextRanges.push_back(make_pair(off, nl));
				nextRanges.push_back(make_pair(off + nl, len - nl));
			}
			cur.swap(nxt);
			ranges.swap(nextRanges);
		}
	}
	const uint32_t wtHdr = 3, mapHdr = 2;         // WVTREE_NOPTRS_HDR, MAPPER_NONE_HDR
	out.write((const char*) &wtHdr, 4);
	out.write((const char*) &newN, 8);
	out.write((const char*) &length, 8);
	out.write((const char*) &maxV, 4);
	out.write((const char*) &height, 4);
	out.write((const char*) &mapHdr, 4);
	for (uint32_t level = 0; level < height; ++level) saveRRR(out, levels[level], (size_t) newN, 8, E);
	out.write((const char*) OCC.data(), (streamsize) (OCC.size() * 4));
	return out;
}

namespace {
// BitSequenceRRR::load: header, classes, offsets -> the plain bitmap
bool loadRRR(istream& in, vector<uint8_t>& bits, const RrrTable& E, const vector<vector<uint16_t> >& valueOf) {
	uint32_t hdr = 0, cLen = 0, cFieldBits = 0, oLen = 0, oBits = 0, sampleRate = 0;
	uint64_t length = 0, ones = 0;
	in.read((char*) &hdr, 4); in.read((char*) &length, 8); in.read((char*) &ones, 8);
	in.read((char*) &cLen, 4); in.read((char*) &cFieldBits, 4); in.read((char*) &oLen, 4); in.read((char*) &oBits, 4); in.read((char*) &sampleRate, 4);
	if (!in || hdr != 2 || cFieldBits != 4 || length > ((uint64_t) 1 << 40) || cLen != length / 15 + (length % 15 != 0)) return false;
	BitWords C(uintLen(cLen, cFieldBits)), O(oLen);
	in.read((char*) C.w.data(), (streamsize) (C.w.size() * 4));
	in.read((char*) O.w.data(), (streamsize) (O.w.size() * 4));
	if (!in) return false;
	bits.assign(length, 0);
	size_t oPos = 0;
	for (uint32_t i = 0; i < cLen; ++i) {
		unsigned cls = 0;
		for (unsigned b = 0; b < 4; ++b) cls |= (unsigned) C.bit((size_t) i * 4 + b) << b;
		if (cls > 15) return false;
		const unsigned nb = E.logBinom[cls];
		uint32_t off = 0;
		for (unsigned b = 0; b < nb; ++b) off |= (uint32_t) O.bit(oPos + b) << b;
		oPos += nb;
		if (off >= valueOf[cls].size()) return false;
		const uint16_t v = valueOf[cls][off];
		for (unsigned b = 0; b < 15 && (size_t) i * 15 + b < length; ++b) bits[(size_t) i * 15 + b] = (v >> b) & 1u;
	}
	return true;
}
}

istream& CSFMIndex::load(istream& in) {
	string abcName;
	char gapCh;
	uint16_t cs16 = 0;
	loadString(abcName, in);
	in.read(&gapCh, 1);
	in.read((char*) &cs16, 2);
	in.read((char*) &concatLen, 4);
	int32_t C256[256];
	in.read((char*) C256, sizeof C256);
	if (!in || concatLen < 1) { in.setstate(ios_base::badbit); return in; }
	csLen = cs16;
	for (int b = 0; b < 6; ++b) C[b] = C256[b];
	const int32_t N = concatLen + 1;
	string csSeq;
	loadString(csSeq, in);
	in.ignore((streamsize) ((size_t) (csLen + 1) * sizeof(double))); // csIdentity: not used on the run path
	concat2CS.assign((size_t) N + 1, 0);                              // 1-based here, 0-based in the file
	in.read((char*) (concat2CS.data() + 1), (streamsize) ((size_t) N * sizeof(uint16_t)));
	vector<uint32_t> sampled((size_t) concatLen / 4);                 // every 4th text position, in suffix-array order
	in.read((char*) sampled.data(), (streamsize) (sampled.size() * sizeof(uint32_t)));
	const RrrTable E;
	vector<vector<uint16_t> > valueOf(16);
	for (uint32_t v = 0; v < (1u << 15); ++v) {
		vector<uint16_t>& list = valueOf[__builtin_popcount(v)];
		if (list.size() <= E.offsetOf[v]) list.resize((size_t) E.offsetOf[v] + 1);
		list[E.offsetOf[v]] = (uint16_t) v;
	}
	vector<uint8_t> bm;
	if (!loadRRR(in, bm, E, valueOf) || bm.size() != (size_t) N) { in.setstate(ios_base::badbit); return in; } // saIdx
	// the row of text position 0 (its BWT symbol is the wrapped-around terminator): the sampled rank whose value is 0, or
	// the one sampled rank the file does not carry (it holds concatLen / 4 values for ceil(N / 4) sampled rows)
	int32_t wrapRow = -1;
	{
		size_t k = 0, missing = (size_t) -1;
		for (int32_t i = 0; i < N && wrapRow < 0; ++i)
			if (bm[i]) {
				if (k < sampled.size()) { if (sampled[k] == 0) wrapRow = i; }
				else missing = (size_t) i;
				k++;
			}
		if (wrapRow < 0 && missing != (size_t) -1) wrapRow = (int32_t) missing;
		if (wrapRow < 0) { in.setstate(ios_base::badbit); return in; }
	}
	uint32_t wtHdr = 0, maxV = 0, height = 0, mapHdr = 0;
	uint64_t newN = 0, length = 0;
	in.read((char*) &wtHdr, 4); in.read((char*) &newN, 8); in.read((char*) &length, 8); in.read((char*) &maxV, 4); in.read((char*) &height, 4); in.read((char*) &mapHdr, 4);
	if (!in || wtHdr != 3 || mapHdr != 2 || length != (uint64_t) N || newN < length || maxV > 4 || height > 3) { in.setstate(ios_base::badbit); return in; }
	// the wavelet tree, undone level by level: a level's bit is the symbol's next bit (most significant first), and the
	// elements of every node move on in a stable split
	vector<uint8_t> sym(newN, 0);
	{
		vector<uint32_t> orig(newN), next(newN);
		for (uint64_t i = 0; i < newN; ++i) orig[i] = (uint32_t) i;
		vector<pair<size_t, size_t> > ranges(1, make_pair((size_t) 0, (size_t) newN)), nextRanges;
		for (uint32_t level = 0; level < height; ++level) {
			if (!loadRRR(in, bm, E, valueOf) || bm.size() != newN) { in.setstate(ios_base::badbit); return in; }
			nextRanges.clear();
			for (size_t r = 0; r < ranges.size(); ++r) {
				const size_t off = ranges[r].first, len = ranges[r].second;
				size_t nl = 0;
				for (size_t i = 0; i < len; ++i) nl += !bm[off + i];
				size_t l = off, rr = off + nl;
				for (size_t i = 0; i < len; ++i) {
					if (bm[off + i]) { sym[orig[off + i]] |= (uint8_t) (1u << (height - level - 1)); next[rr++] = orig[off + i]; }
					else next[l++] = orig[off + i];
				}
				nextRanges.push_back(make_pair(off, nl));
				nextRanges.push_back(make_pair(off + nl, len - nl));
			}
			orig.swap(next);
			ranges.swap(nextRanges);
		}
	}
	in.ignore((streamsize) ((size_t) (maxV + 2) * 4)); // OCC
	if (!in) { in.setstate(ios_base::badbit); return in; }
	// rank blocks over the BWT
	rank.assign((size_t) N / 64 + 2, RankBlock());
	uint32_t run[5] = {0, 0, 0, 0, 0};
	for (int32_t i = 0; i < N; ++i) {
		RankBlock& rb = rank[(size_t) i / 64];
		if (i % 64 == 0)
			for (int b = 0; b < 5; ++b) { rb.cnt[b] = run[b]; rb.mask[b] = 0; }
		const int b = sym[i] > 4 ? 0 : sym[i];
		rb.mask[b] |= 1ull << (i % 64);
		run[b]++;
	}
	buildQTable();
	// the suffix array from the BWT: the lone terminator is the smallest suffix (row 0, text position N - 1); an LF step
	// leads from the row of position p to the row of position p - 1
	// leads from the row of position p to the row of position p - 1.  Separators and the terminator share symbol 0, for
	// which the usual rank formula is off: the rows that start with a separator are ordered by what FOLLOWS them, i.e. by the
	// rows of the terminator (row 0) and of the sequence starts -- exactly the rows whose BWT symbol is 0, the wrapped row of
	// text position 0 aside -- so the k-th such row (k = 0 for row 0) leads to row k + 1.
	vector<int32_t> zeroRows;
	for (int32_t i = 0; i < N; ++i)
		if ((sym[i] > 4 ? 0 : sym[i]) == 0 && i != wrapRow) zeroRows.push_back(i);
	sa.assign(N, 0);
	int32_t row = 0;
	sa[0] = N; // 1-based text positions, as build() keeps them
	for (int32_t p = N - 1; p >= 1; --p) {
		const int c = sym[row] > 4 ? 0 : sym[row];
		if (c == 0) {
			const vector<int32_t>::const_iterator it = lower_bound(zeroRows.begin(), zeroRows.end(), row);
			if (it == zeroRows.end() || *it != row) { in.setstate(ios_base::badbit); return in; }
			row = (int32_t) (it - zeroRows.begin()) + 1;
		} else
			row = C[c] + occ(c, row) - 1;
		if (row < 0 || row >= N || sa[row] != 0) { in.setstate(ios_base::badbit); return in; }
		sa[row] = p;
	}
	return in;
}

bool CSFMIndex::operator==(const CSFMIndex& o) const {
	if (csLen != o.csLen || concatLen != o.concatLen || memcmp(C, o.C, sizeof C) != 0 || sa != o.sa || concat2CS != o.concat2CS || rank.size() != o.rank.size()) return false;
	for (size_t i = 0; i + 1 < rank.size(); ++i) // the last block is never read
		if ((size_t) i * 64 <= (size_t) concatLen && (memcmp(rank[i].cnt, o.rank[i].cnt, sizeof rank[i].cnt) != 0 || memcmp(rank[i].mask, o.rank[i].mask, sizeof rank[i].mask) != 0)) return false;
	return true;
}

// ------------------------------------------------------------------ discrete Gamma ----
// The reference takes these from Boost.Math (gamma_p, quantile(gamma_distribution)); Boost is not a dependency here.
double gammaP(double a, double x) {
	if (!(x > 0)) return 0;
	if (x == INFINITY) return 1;
	const double lg = ::lgamma(a);
	if (x < a + 1) { // series: P = e^-x x^a / Gamma(a) * sum_n x^n / (a (a+1) ... (a+n))
		double term = 1 / a, sum = term;
		for (int n = 1; n < 100000; ++n) {
			term *= x / (a + n);
			sum += term;
			if (::fabs(term) < ::fabs(sum) * 1e-17) break;
		}
		return sum * ::exp(-x + a * ::log(x) - lg);
	}
	// continued fraction for Q = 1 - P (modified Lentz)
	const double tiny = 1e-300;
	double b = x + 1 - a, c = 1 / tiny, d = 1 / b, h = d;
	for (int n = 1; n < 100000; ++n) {
		const double an = -n * (n - a);
		b += 2;
		d = an * d + b; if (::fabs(d) < tiny) d = tiny;
		c = b + an / c; if (::fabs(c) < tiny) c = tiny;
		d = 1 / d;
		const double del = d * c;
		h *= del;
		if (::fabs(del - 1) < 1e-16) break;
	}
	return 1 - ::exp(-x + a * ::log(x) - lg) * h;
}

double gammaQuantile(double shape, double scale, double p) {
	if (!(p > 0)) return 0;
	if (!(p < 1)) return INFINITY;
	// bracket, then bisection polished by Newton steps on P(shape, y) = p (density e^-y y^(shape-1) / Gamma(shape))
	double lo = 0, hi = shape > 1 ? shape : 1;
	while (gammaP(shape, hi) < p) hi *= 2;
	double y = 0.5 * (lo + hi);
	const double lg = ::lgamma(shape);
	for (int it = 0; it < 200; ++it) {
		const double f = gammaP(shape, y) - p;
		if (f > 0) hi = y; else lo = y;
		const double dens = ::exp(-y + (shape - 1) * ::log(y) - lg);
		double next = dens > 0 ? y - f / dens : 0.5 * (lo + hi);
		if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
		if (::fabs(next - y) <= 1e-16 * ::fabs(y)) { y = next; break; }
		y = next;
	}
	return y * scale;
}

DiscreteGammaModel& DiscreteGammaModel::build(int K_, double alpha_) {
	K = K_; alpha = alpha_;
	b.assign(K + 1, 0.0); r.assign(K, 0.0);
	for (int i = 0; i < K; ++i) b[i] = gammaQuantile(alpha, alpha, i / (double) K); // setBreaks, cpp:40-45
	b[K] = INFINITY;
	for (int i = 0; i < K; ++i)                                                     // setRates, cpp:47-55
		r[i] = b[i + 1] != INFINITY ? gammaP(alpha + 1, b[i + 1] * alpha) - gammaP(alpha + 1, b[i] * alpha)
		                            : 1 - gammaP(alpha + 1, b[i] * alpha);
	return *this;
}

double DiscreteGammaModel::estimateShape(const vector<double>& X) {
	if (X.size() < 2) return INFINITY;
	double m = 0;
	for (size_t i = 0; i < X.size(); ++i) m += X[i];
	m /= (double) X.size();
	double ss = 0;
	for (size_t i = 0; i < X.size(); ++i) ss += (X[i] - m) * (X[i] - m);
	const double s = ss / (double) (X.size() - 1);
	return m * m / (s - m);
}

// ------------------------------------------------------------------ profile training ----
namespace {
struct Dirichlet { // a density (L = 1) or a mixture of L densities over K categories (src/math/DirichletDensity, DirichletMixture)
	int K, L;
	vector<double> q, alpha; // alpha[i * L + j]
	Dirichlet() : K(0), L(0) {}
	bool read(istream& in) { // DirichletDensity::read, cpp:135-162 / DirichletMixture::read, cpp:254-285
		string line;
		if (!getline(in, line)) return false;
		const bool mix = line == "Dirichlet Mixture Model";
		if (!mix && line != "Dirichlet Density Model") return false;
		getline(in, line); // training cost
		getline(in, line);
		L = 1;
		if (mix ? sscanf(line.c_str(), "K: %d L: %d", &K, &L) != 2 : sscanf(line.c_str(), "K: %d", &K) != 1) return false;
		q.assign(L, 1.0);
		alpha.assign((size_t) K * L, 0.0);
		if (mix) {
			getline(in, line); // "Mixture coefficients:"
			for (int j = 0; j < L; ++j) in >> q[j];
			getline(in, line);
		}
		getline(in, line); // "alpha:"
		for (int i = 0; i < K; ++i)
			for (int j = 0; j < L; ++j) in >> alpha[(size_t) i * L + j];
		getline(in, line);
		return !in.fail();
	}
	// posterior mean of the category probabilities given observed (weighted) counts v[0..K)
	void meanPostP(double* v) const {
		double dataSum = 0;
		for (int i = 0; i < K; ++i) dataSum += v[i];
		if (L == 1) { // DirichletDensity::meanPostP, cpp:25-27
			double aSum = 0;
			for (int i = 0; i < K; ++i) aSum += alpha[i];
			for (int i = 0; i < K; ++i) v[i] = (v[i] + alpha[i]) / (dataSum + aSum);
			return;
		}
		vector<double> logB(L), aSum(L, 0.0), X(K, 0.0); // DirichletMixture::meanPostP, cpp:45-61
		double mx = -INFINITY;
		for (int j = 0; j < L; ++j) {
			double s1 = 0, t1 = 0, s0 = 0, t0 = 0;
			for (int i = 0; i < K; ++i) {
				const double a = alpha[(size_t) i * L + j];
				s1 += ::lgamma(a + v[i]); t1 += a + v[i];
				s0 += ::lgamma(a); t0 += a;
			}
			aSum[j] = t0;
			logB[j] = (s1 - ::lgamma(t1)) - (s0 - ::lgamma(t0));
			mx = max(mx, logB[j]);
		}
		double tot = 0;
		for (int i = 0; i < K; ++i) {
			for (int j = 0; j < L; ++j) X[i] += q[j] * ::exp(logB[j] - mx) * (alpha[(size_t) i * L + j] + v[i]) / (aSum[j] + dataSum);
			tot += X[i];
		}
		for (int i = 0; i < K; ++i) v[i] = X[i] / tot;
	}
};
struct P7Counts { // probabilities or counts: emissions [k * 4 + b], transitions [k * 9 + a * 3 + b] with M, I, D = 0, 1, 2
	int K;
	vector<double> EM, EI, T;
	void scale(double r) {
		for (size_t i = 0; i < EM.size(); ++i) { EM[i] *= r; EI[i] *= r; }
		for (size_t i = 0; i < T.size(); ++i) T[i] *= r;
	}
	void estimate(const Dirichlet& ME, const Dirichlet& IE, const Dirichlet& MT, const Dirichlet& IT, const Dirichlet& DT) { // estimateParams, cpp:280-315
		for (int k = 0; k <= K; ++k) {
			double* t = &T[(size_t) k * 9];
			MT.meanPostP(t);                 // M -> M, I, D
			IT.meanPostP(t + 3);             // I -> M, I
			double d[2] = {t[6], t[8]};      // D -> M, D
			DT.meanPostP(d);
			t[6] = d[0]; t[8] = d[1];
			ME.meanPostP(&EM[(size_t) k * 4]);
			IE.meanPostP(&EI[(size_t) k * 4]);
		}
		T[6] = 1; T[8] = 0;                                   // T[0](D, M) = 1, T[0](D, D) = 0
		double* tk = &T[(size_t) K * 9];
		tk[2] = 0; tk[6] = 1; tk[8] = 0;                      // T[K](M, D) = 0, T[K](D, M) = 1, T[K](D, D) = 0
	}
	double meanRelativeEntropy() const { // cpp:317-322 against the uniform background a fresh profile carries, in bits
		double ent = 0;
		for (int k = 1; k <= K; ++k) {
			double e = 0;
			for (int b = 0; b < 4; ++b) {
				const double p = EM[(size_t) k * 4 + b];
				if (p > 0) e += p * ::log(p / 0.25);
			}
			ent += e / ::log(2.0);
		}
		return ent / K;
	}
};
inline int matchingState(const vector<int32_t>& cs2p, int loc, int b) { // src/BandedHMMP7.h:713-716: M, I, D = 0, 1, 2; 3 = nothing here
	const bool isPos = cs2p[loc] != cs2p[loc - 1];
	return isPos && b >= 0 ? 0 : isPos && b < 0 ? 2 : !isPos && b >= 0 ? 1 : 3;
}
}

bool BandedHMMP7::build(const MSA& msa, double symfrac, istream& priorIn, bool asCompiled) {
	if (msa.numSeq == 0 || msa.csLen == 0 || !(symfrac > 0 && symfrac < 1)) return false;
	if (msa.seqWeight.size() != msa.numSeq) return false; // counts and weights come from MSA::loadFasta / prune
	Dirichlet ME, IE, MT, IT, DT;
	{ // BandedHMMP7Prior, src/BandedHMMP7Prior.cpp:20-45
		string head;
		while (getline(priorIn, head)) {
			Dirichlet* d = head.compare(0, 15, "Match emission:") == 0 ? &ME : head.compare(0, 16, "Insert emission:") == 0 ? &IE :
			               head.compare(0, 17, "Match transition:") == 0 ? &MT : head.compare(0, 18, "Insert transition:") == 0 ? &IT :
			               head.compare(0, 18, "Delete transition:") == 0 ? &DT : NULL;
			if (d && !d->read(priorIn)) return false;
		}
		if (ME.K != 4 || IE.K != 4 || MT.K != 3 || IT.K != 2 || DT.K != 2) return false;
	}
	name = msa.name;
	L = (int) msa.csLen;
	const unsigned N = msa.numSeq;
	// match columns: weighted symbol fraction >= symfrac (cpp:405-411)
	cs2ProfileIdx.assign(L + 1, 0);
	profile2CSIdx.assign(1, 0);
	int k = 0;
	for (int j = 1; j <= L; ++j) {
		double res = 0;
		for (int b = 0; b < 4; ++b) res += msa.resWCount[(size_t) (j - 1) * 4 + b];
		if (res / (res + msa.gapWCount[j - 1]) >= symfrac) { profile2CSIdx.push_back(j); ++k; }
		cs2ProfileIdx[j] = k;
	}
	K = k;
	if (K < 1) return false;
	P7Counts c;
	c.K = K;
	c.EM.assign((size_t) (K + 1) * 4, 0.0); c.EI.assign((size_t) (K + 1) * 4, 0.0); c.T.assign((size_t) (K + 1) * 9, 0.0);
	// observed weighted counts (cpp:428-478)
	for (int j = 1; j <= L; ++j) {
		k = cs2ProfileIdx[j];
		for (unsigned i = 0; i < N; ++i) {
			const int b = DNA::encode((char) ::toupper(msa.residualAt(i, j - 1)));
			const double w = msa.seqWeight[i];
			const int sm = matchingState(cs2ProfileIdx, j, b);
			if (sm == 3) continue;
			if (sm == 0) { c.EM[b] += w; c.EM[(size_t) k * 4 + b] += w; } // M0 collects the composition
			else if (sm == 1) c.EI[(size_t) k * 4 + b] += w;
			int jN, smN = 3;
			for (jN = j + 1; jN <= L; ++jN) { // the next position of this row that is in some state
				smN = matchingState(cs2ProfileIdx, jN, DNA::encode((char) ::toupper(msa.residualAt(i, jN - 1))));
				if (smN != 3) break;
			}
			if (!(jN <= L && smN != 3)) continue;
			if (asCompiled) smN = 0; // the variable the reference tests and indexes with here is never assigned: GCC reads it as M
			if ((sm == 1 && smN == 2) || (sm == 2 && smN == 1)) continue; // no I->D or D->I
			c.T[(size_t) k * 9 + sm * 3 + smN] += w;
		}
	}
	for (unsigned i = 0; i < N; ++i) { // B -> first state, last state -> E (cpp:466-476)
		const double w = msa.seqWeight[i];
		const int start = msa.startIdx[i], end = msa.endIdx[i];
		if (start < 0) continue;
		const int smS = matchingState(cs2ProfileIdx, start + 1, DNA::encode((char) ::toupper(msa.residualAt(i, start))));
		c.T[0 * 3 + smS] += w;
		const int smE = matchingState(cs2ProfileIdx, end + 1, DNA::encode((char) ::toupper(msa.residualAt(i, end))));
		c.T[(size_t) K * 9 + smE * 3 + 0] += w;
	}
	// effective sequence number: bisection on [0, N] to a mean relative entropy of DEFAULT_ERE = 1 bit
	// (RelativeEntropyTargetFunc, cpp:1122-1135; RootFinder::rootBisection, src/math/RootFinder.cpp:22-78)
	const double nSeq = N;
	struct Target {
		const P7Counts& c; double nSeq; const Dirichlet &ME, &IE, &MT, &IT, &DT;
		double operator()(double x) const {
			if (x > nSeq) return 0;
			P7Counts t = c;
			t.scale(x / nSeq);
			t.estimate(ME, IE, MT, IT, DT);
			return t.meanRelativeEntropy() - 1.0;
		}
	} f = {c, nSeq, ME, IE, MT, IT, DT};
	double effN = nSeq;
	{
		double xl = 0, xr = nSeq, x = NAN;
		double fxl = f(xl), fxr = f(xr);
		if (fxl * fxr < 0) {
			for (;;) {
				x = (xl + xr) / 2;
				const double fx = f(x);
				if (fx == 0) break;
				if (xr - xl < 1e-10 + 1e-10 * x) break;
				if (fxl > 0) { if (fx > 0) { xl = x; fxl = fx; } else { xr = x; fxr = fx; } }
				else { if (fx < 0) { xl = x; fxl = fx; } else { xr = x; fxr = fx; } }
			}
		}
		if (!::isnan(x)) effN = x;
	}
	c.scale(effN / nSeq);
	c.estimate(ME, IE, MT, IT, DT);
	// costs
	E_M_cost.assign(c.EM.size(), 0.0); E_I_cost.assign(c.EI.size(), 0.0); Tmat_cost.assign(c.T.size(), 0.0);
	for (size_t i = 0; i < c.EM.size(); ++i) { E_M_cost[i] = -::log(c.EM[i]); E_I_cost[i] = -::log(c.EI[i]); }
	for (size_t i = 0; i < c.T.size(); ++i) Tmat_cost[i] = -::log(c.T[i]);
	entryPr_cost.assign(K + 1, INFINITY); exitPr_cost.assign(K + 1, INFINITY);
	wingRetracted = false;
	// tags (cpp:496-538)
	char value[128];
	optTags.clear();
	snprintf(value, sizeof value, "%d", L); optTags.push_back(make_pair(string("MAXL"), string(value)));
	optTags.push_back(make_pair(string("RF"), string("no")));
	optTags.push_back(make_pair(string("MM"), string("no")));
	optTags.push_back(make_pair(string("CONS"), string("yes")));
	optTags.push_back(make_pair(string("CS"), string("no")));
	optTags.push_back(make_pair(string("MAP"), string("yes")));
	snprintf(value, sizeof value, "%u", N); optTags.push_back(make_pair(string("NSEQ"), string(value)));
	snprintf(value, sizeof value, "%g", effN); optTags.push_back(make_pair(string("EFFN"), string(value)));
	locMap.assign(K + 1, "");
	for (k = 1; k <= K; ++k) {
		const int map = profile2CSIdx[k];
		int bb = 0; // MSA::CSBaseAt, src/MSA.cpp:52-57: the most frequent base by weighted count (first maximum), never a gap
		for (int b = 1; b < 4; ++b) if (msa.resWCount[(size_t) (map - 1) * 4 + b] > msa.resWCount[(size_t) (map - 1) * 4 + bb]) bb = b;
		char ch = DNA::decode((int8_t) bb);
		if (msa.resWCount[(size_t) (map - 1) * 4 + bb] / N < 0.9) ch = (char) ::tolower(ch); // wIdentityAt < CONS_THRESHOLD
		locMap[k] = string(1, ch) + "\t-\t-\t-";
	}
	time_t raw;
	time(&raw);
	strftime(value, sizeof value, "%c", localtime(&raw));
	optTags.push_back(make_pair(string("DATE"), string(value)));
	return true;
}

// ------------------------------------------------------------------ after the evaluation ----
namespace {
const double kMinLoglikExp = -1021 / 2; // MIN_LOGLIK_EXP = DBL_MIN_EXP / 2 (integer division), src/PhyloTreeUnrooted.cpp:68
// dot_product_scaled(X, V), src/PhyloTreeUnrooted.h:1495-1503: log(X exp(V)) row by row, V lifted when it would underflow
void dotProductScaled(const double* P16, const double* V, double* Y) {
	const double maxV = max(max(V[0], V[1]), max(V[2], V[3]));
	const double scale = maxV != -INFINITY && maxV < kMinLoglikExp ? kMinLoglikExp - maxV : 0;
	double e[4];
	for (int k = 0; k < 4; ++k) e[k] = ::exp(V[k] + scale);
	for (int i = 0; i < 4; ++i) {
		double d = 0;
		for (int k = 0; k < 4; ++k) d += P16[i * 4 + k] * e[k];
		Y[i] = ::log(d) - scale;
	}
}
}

void PTUnrooted::updateRootLoglik() {
	const size_t per = (size_t) csLen * 4;
	rootLoglik.assign(per, 0.0);
	const int K = hasDG ? dG.K : 1;
	vector<vector<double> > P; // [child][rate] 4x4
	vector<long> kids;
	for (size_t c = 0; c < neighbors[root].size(); ++c) {
		const long ch = neighbors[root][c];
		if (parent[ch] != root) continue;
		kids.push_back(ch);
		for (int k = 0; k < K; ++k) {
			P.push_back(vector<double>(16));
			model.Pr(blen[ch] * (hasDG ? dG.rate(k) : 1.0), P.back().data());
		}
	}
	vector<double> mat((size_t) 4 * K);
	for (int j = 0; j < csLen; ++j) {
		fill(mat.begin(), mat.end(), 0.0);
		for (size_t c = 0; c < kids.size(); ++c)
			for (int k = 0; k < K; ++k) {
				double y[4];
				dotProductScaled(P[c * K + k].data(), &up[(size_t) kids[c] * per + (size_t) j * 4], y);
				for (int i = 0; i < 4; ++i) mat[(size_t) i * K + k] += y[i];
			}
		double* out = &rootLoglik[(size_t) j * 4];
		if (!hasDG) { for (int i = 0; i < 4; ++i) out[i] = mat[i]; continue; }
		for (int i = 0; i < 4; ++i) { // row_mean_exp_scaled, .h:1521-1529
			double mx = -INFINITY;
			for (int k = 0; k < K; ++k) mx = max(mx, mat[(size_t) i * K + k]);
			const double scale = mx != -INFINITY && mx < kMinLoglikExp ? kMinLoglikExp - mx : 0;
			double sum = 0;
			for (int k = 0; k < K; ++k) sum += ::exp(mat[(size_t) i * K + k] + scale);
			out[i] = ::log(sum / K) - scale;
		}
	}
}

vector<double> PTUnrooted::numMutations() const {
	// estimateNumMutations(j): edges whose two ends infer different states, a node's state being the arg-max (first maximum)
	// of its message towards the parent.  When hmmufotu-build asks, the root's own block still holds the INVALID_LOGLIK
	// constant (initRootLoglik fills it, evaluate() does not cache the root), so the root "infers" state 0 at every site.
	const size_t per = (size_t) csLen * 4;
	vector<double> n(csLen, 0.0);
	vector<int8_t> state(numNodes());
	for (int j = 0; j < csLen; ++j) {
		for (size_t id = 0; id < numNodes(); ++id) {
			int best = 0;
			if ((long) id != root) {
				const double* v = &up[id * per + (size_t) j * 4];
				for (int k = 1; k < 4; ++k) if (v[k] > v[best]) best = k;
			}
			state[id] = (int8_t) best;
		}
		int cnt = 0;
		for (size_t id = 0; id < numNodes(); ++id)
			if ((long) id != root && state[id] != state[parent[id]]) cnt++;
		n[j] = cnt;
	}
	return n;
}

// ------------------------------------------------------------------ Newick ----
// tree = subtree ';'    subtree = -('(' subtree % ',' ')') >> -label >> -(':' double)    skipper: white space
// label = +(printable - space - "()[]':;,")  |  '\'' +(printable - '\'') '\''            (src/NewickTree.h:172-236)
namespace {
struct NewickParser {
	const string& s;
	size_t i;
	explicit NewickParser(const string& s) : s(s), i(0) {}
	void skip() { while (i < s.size() && ::isspace((unsigned char) s[i])) ++i; }
	static bool labelChar(char c) {
		return ::isprint((unsigned char) c) && !::isspace((unsigned char) c) && c != '(' && c != ')' && c != '[' && c != ']' &&
		       c != '\'' && c != ':' && c != ';' && c != ',';
	}
	bool subtree(NewickTree& node) {
		skip();
		if (i < s.size() && s[i] == '(') { // optional descendant list: when it does not parse, it matches nothing
			const size_t save = i;
			++i;
			vector<NewickTree> kids;
			bool ok = true;
			for (;;) {
				kids.push_back(NewickTree());
				if (!subtree(kids.back())) { ok = false; break; }
				skip();
				if (i < s.size() && s[i] == ',') { ++i; continue; }
				break;
			}
			skip();
			if (ok && i < s.size() && s[i] == ')') { ++i; node.children.swap(kids); }
			else i = save;
		}
		skip();
		if (i < s.size() && s[i] == '\'') { // quoted label
			size_t j = i + 1;
			while (j < s.size() && s[j] != '\'' && ::isprint((unsigned char) s[j])) ++j;
			if (j < s.size() && s[j] == '\'' && j > i + 1) { node.name = s.substr(i + 1, j - i - 1); i = j + 1; }
		} else {
			size_t j = i;
			while (j < s.size() && labelChar(s[j])) ++j;
			if (j > i) { node.name = s.substr(i, j - i); i = j; }
		}
		skip();
		if (i < s.size() && s[i] == ':') { // optional branch length
			const size_t save = i;
			++i;
			skip();
			const char* b = s.c_str() + i;
			char* e = NULL;
			const double v = ::strtod(b, &e);
			if (e != b) { node.length = v; i += (size_t) (e - b); }
			else i = save;
		}
		return true;
	}
};
}

bool NewickTree::read(istream& in) {
	string text, line;
	while (getline(in, line)) text += line + "\n";
	NewickParser p(text);
	NewickTree t;
	if (!p.subtree(t)) return false;
	p.skip();
	if (!(p.i < text.size() && text[p.i] == ';')) return false;
	*this = t;
	return true;
}

// ------------------------------------------------------------------ tree skeleton ----
namespace {
const char* const kTaxonSep = ";: "; // src/PhyloTreeUnrooted.cpp:88

// boost::split(fields, s, is_any_of(seps), token_compress_on): runs of separators count as one; a leading or
// trailing run leaves an empty token at that end
vector<string> splitCompress(const string& s, const char* seps) {
	vector<string> out;
	string cur;
	bool inSep = false;
	for (size_t i = 0; i < s.size(); ++i) {
		if (::strchr(seps, s[i]) != NULL) {
			if (!inSep) { out.push_back(cur); cur.clear(); inSep = true; }
		} else {
			cur.push_back(s[i]);
			inSep = false;
		}
	}
	out.push_back(cur);
	return out;
}
bool startsWith(const string& s, const string& prefix) { return s.size() >= prefix.size() && s.compare(0, prefix.size(), prefix) == 0; }
bool isCanonical(const string& t) { // src/PhyloTreeUnrooted.h:1564-1574
	static const char* const pre[8] = {"d__", "k__", "p__", "c__", "o__", "f__", "g__", "s__"};
	if (t.length() <= 3) return false;
	for (int i = 0; i < 8; ++i) if (startsWith(t, pre[i])) return true;
	return false;
}
bool isCanonicalAt(const string& t, size_t level) { // .h:1531-1552, 1576-1578: levels beyond species have an empty prefix
	static const char* const pre[7] = {"k__", "p__", "c__", "o__", "f__", "g__", "s__"};
	return level < 7 ? startsWith(t, pre[level]) : true;
}
bool isFullCanonical(const string& taxon) { // src/PhyloTreeUnrooted.cpp:956-963
	const vector<string> f = splitCompress(taxon, kTaxonSep);
	for (size_t l = 0; l < f.size(); ++l) if (!isCanonicalAt(f[l], l)) return false;
	return true;
}
bool isPartialCanonical(const string& taxon) { // :965-972
	const vector<string> f = splitCompress(taxon, kTaxonSep);
	for (size_t l = 0; l < f.size(); ++l) if (!isCanonical(f[l])) return false;
	return true;
}
string formatTaxonName(const string& taxon) { // :974-986
	if (taxon.empty()) return taxon;
	const vector<string> f = splitCompress(taxon, kTaxonSep);
	string out;
	bool first = true;
	for (size_t l = 0; l < f.size(); ++l)
		if (isCanonical(f[l])) { if (!first) out += ";"; out += f[l]; first = false; }
	return out;
}
}

bool buildTree(PTUnrooted& tree, const NewickTree& nt, const MSA& msa, istream* annoIn, const string& rootName) {
	// node ids: explicit-stack DFS, id at pop time, children pushed in file order => the LAST child is numbered first
	// (src/PhyloTreeUnrooted.cpp:131-155); a second identical walk adds the edges (:157-181)
	vector<const NewickTree*> byId;
	map<const NewickTree*, long> idOf;
	{
		vector<const NewickTree*> S(1, &nt);
		while (!S.empty()) {
			const NewickTree* v = S.back();
			S.pop_back();
			idOf[v] = (long) byId.size();
			byId.push_back(v);
			for (size_t c = 0; c < v->children.size(); ++c) S.push_back(&v->children[c]);
		}
	}
	const size_t n = byId.size();
	if (n < 2) return false;
	tree = PTUnrooted();
	tree.csLen = (int) msa.csLen;
	tree.root = 0;
	tree.parent.assign(n, -1);
	tree.blen.assign(n, 0.0); tree.height.assign(n, 0.0); tree.annoDist.assign(n, 0.0);
	tree.name.resize(n); tree.anno.resize(n); tree.seqName.resize(n);
	tree.neighbors.assign(n, vector<long>());
	tree.seq.assign(n * (size_t) msa.csLen, -2);
	for (size_t i = 0; i < n; ++i) tree.name[i] = byId[i]->name;
	{
		vector<const NewickTree*> S(1, &nt);
		while (!S.empty()) {
			const NewickTree* v = S.back();
			S.pop_back();
			const long u = idOf[v];
			for (size_t c = 0; c < v->children.size(); ++c) {
				const NewickTree* w = &v->children[c];
				const long k = idOf[w];
				tree.neighbors[u].push_back(k); // addEdge(u, child), src/PhyloTreeUnrooted.h:577
				tree.neighbors[k].push_back(u);
				tree.parent[k] = (int32_t) u;
				tree.blen[k] = w->length;
				S.push_back(w);
			}
		}
	}
	// loadMSA (:183-221): leaves (one neighbour) take the row with their ORIGINAL name; rows nobody claims are ignored
	map<string, unsigned> rowOf;
	for (unsigned i = 0; i < msa.numSeq; ++i) {
		if (rowOf.count(msa.seqNames[i])) return false; // non-unique sequence name
		rowOf[msa.seqNames[i]] = i;
	}
	tree.msaIndex.clear();
	for (size_t id = 0; id < n; ++id) {
		if (tree.neighbors[id].size() != 1) continue;
		const map<string, unsigned>::const_iterator it = rowOf.find(tree.name[id]);
		if (it == rowOf.end()) continue;
		tree.seqName[id] = msa.seqNames[it->second];
		for (unsigned j = 0; j < msa.csLen; ++j) // DigitalSeq(abc, name, row): src/DigitalSeq.cpp:41-48
			tree.seq[id * (size_t) msa.csLen + j] = DNA::encode((char) ::toupper(msa.residualAt(it->second, j)));
		tree.msaIndex.push_back(make_pair(it->second, (long) id));
	}
	sort(tree.msaIndex.begin(), tree.msaIndex.end());
	// fixBranchLength (:289-296): only LEAF branches of length <= 0 are lifted, to BRANCH_EPS = 1e-5
	for (size_t id = 0; id < n; ++id)
		if (tree.parent[id] >= 0 && tree.neighbors[id].size() == 1 && tree.blen[id] <= 0) tree.blen[id] = 1e-5;
	// loadAnnotation (:223-240): "name <TAB> taxonomy" lines replace the names they match
	if (annoIn) {
		map<string, string> annoOf;
		string line;
		while (getline(*annoIn, line)) {
			istringstream ls(line);
			string nm, an;
			getline(ls, nm, '\t');
			getline(ls, an, '\t');
			annoOf[nm] = an;
		}
		for (size_t id = 0; id < n; ++id) {
			const map<string, string>::const_iterator it = annoOf.find(tree.name[id]);
			if (it != annoOf.end()) tree.name[id] = it->second;
		}
	}
	for (size_t id = 0; id < n; ++id) tree.name[id] = formatTaxonName(tree.name[id]); // formatName, .h:1554-1557
	// annotate (:988-1006): climb to the first ancestor-or-self with a full canonical lineage, collecting partial
	// canonical names and branch lengths on the way
	for (size_t id = 0; id < n; ++id) {
		vector<string> path;
		long p = (long) id;
		while (!isFullCanonical(tree.name[p]) && tree.parent[p] >= 0) {
			tree.annoDist[id] += tree.blen[p];
			if (isPartialCanonical(tree.name[p])) path.push_back(tree.name[p]);
			p = tree.parent[p];
		}
		if (isFullCanonical(tree.name[p])) path.push_back(tree.name[p]);
		reverse(path.begin(), path.end());
		string a;
		for (size_t k = 0; k < path.size(); ++k) { if (k) a += ";"; a += path[k]; }
		tree.anno[id] = !path.empty() ? a : rootName;
	}
	// calcNodeHeight (:274-287): per node, the smallest accumulated distance over the leaves below it
	vector<char> seen(n, 0);
	for (size_t leaf = 0; leaf < n; ++leaf) {
		if (tree.neighbors[leaf].size() != 1) continue;
		double h = 0;
		for (long node = (long) leaf; node >= 0; node = tree.parent[node]) {
			if (!seen[node] || h < tree.height[node]) { tree.height[node] = h; seen[node] = 1; }
			if (tree.parent[node] >= 0) h += tree.blen[node];
		}
	}
	// annotation strings -> ids (equal strings share an id), as PTUnrooted::load leaves them
	map<string, int32_t> ids;
	tree.taxon.resize(n);
	for (size_t i = 0; i < n; ++i) {
		map<string, int32_t>::iterator it = ids.find(tree.anno[i]);
		if (it == ids.end()) it = ids.insert(make_pair(tree.anno[i], (int32_t) ids.size())).first;
		tree.taxon[i] = it->second;
	}
	return true;
}

} // namespace HmmUFOtu
} // namespace EGriceLab
