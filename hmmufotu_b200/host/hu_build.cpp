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
