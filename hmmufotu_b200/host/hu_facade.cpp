// The reference's per-read call surface over the C ABI: the free functions of src/HmmUFOtu_main.h:70-113 and the class
// methods they are built on (BandedHMMP7::calcViterbiScores / buildViterbiTrace / buildGlobalAlign, PTUnrooted::estimateSeq /
// placeSeq), one read (one placement vector) per call.  Every function stages its arguments into a one-read device
// workspace, runs the corresponding stage of include/hmmufotu_b200.h and reads the result back: no arithmetic of the path
// happens on the host (the batched driver, hmmufotu.cpp, is the fast way to use the device; this file is the drop-in one).
#include "hu_host.h"

#include <math.h>
#include <stdexcept>
#include <limits>

using namespace std;

namespace EGriceLab {
namespace HmmUFOtu {

namespace {
void check(int rc, const char* what) {
	if (rc != HU_OK) throw runtime_error(string(what) + ": " + hu_last_error());
}
DeviceDB& deviceOf(DeviceDB* d, const char* who) {
	if (!d) throw logic_error(string(who) + ": no device attached (construct a DeviceDB over the profile and the tree first)");
	return *d;
}
hu_sep_params sepParams() {
	hu_sep_params p;
	hu_sep_params_default(&p);
	p.maxNSeed = HU_MAX_SEED;
	p.maxError = numeric_limits<double>::infinity();
	return p;
}
// window + aligned codes of one read into the workspace (the stages read the window from there)
void stageAligned(hu_batch* b, const DigitalSeq& seq, int start, int end) {
	const int32_t s = start, e = end;
	check(hu_batch_set_aligned(b, 1, seq.data(), &s, &e), "hu_batch_set_aligned");
}
} // namespace

DigitalSeq digitalSeq(const string& align) { // src/DigitalSeq.cpp:41-48: upper case, encode, gaps -> negative
	DigitalSeq d(align.size());
	for (size_t i = 0; i < align.size(); ++i) d[i] = DNA::encode((char) ::toupper(align[i]));
	return d;
}

bool compareByLoglik(const PTUnrooted::PTPlacement& lhs, const PTUnrooted::PTPlacement& rhs) { return lhs.loglik < rhs.loglik; }
bool compareByQPlace(const PTUnrooted::PTPlacement& lhs, const PTUnrooted::PTPlacement& rhs) { return lhs.qPlace < rhs.qPlace; }

// ------------------------------------------------------------------ device attachment ----
DeviceDB::DeviceDB(const BandedHMMP7& hmm_, const PTUnrooted& ptu_, int device) : db(NULL), hmm(hmm_), ptu(ptu_), batch(NULL), batchLen(0) {
	const hu_hmm_params hp = hmm.params();
	const hu_tree_desc td = ptu.treeDesc();
	const hu_model_desc md = ptu.modelDesc();
	check(hu_db_create(&db, &hp, &td, &md, device), "hu_db_create");
	hmm.dev = this;
	ptu.dev = this;
}
DeviceDB::~DeviceDB() {
	if (hmm.dev == this) hmm.dev = NULL;
	if (ptu.dev == this) ptu.dev = NULL;
	if (batch) hu_batch_destroy(batch);
	if (db) hu_db_destroy(db);
}
hu_batch* DeviceDB::one(int maxLen) {
	if (!batch || maxLen > batchLen) {
		if (batch) hu_batch_destroy(batch);
		batch = NULL;
		batchLen = max(maxLen + maxLen / 4, 64);
		check(hu_batch_create(&batch, db, 1, batchLen, 0), "hu_batch_create");
	}
	return batch;
}

// ------------------------------------------------------------------ alignment ----
BandedHMMP7::ViterbiScores& BandedHMMP7::calcViterbiScores(const PrimarySeq& seq, ViterbiScores& vs, const vector<ViterbiAlignPath>& vpaths) const {
	DeviceDB& d = deviceOf(dev, "BandedHMMP7::calcViterbiScores");
	const int N = (int) seq.length();
	vector<int8_t> codes(N);
	for (int i = 0; i < N; ++i) {
		codes[i] = DNA::encode((char) ::toupper(seq.seq[i]));
		if (codes[i] < 0) throw invalid_argument("read '" + seq.id + "' holds a character outside the nucleotide alphabet");
	}
	const int64_t offsets[2] = {0, N};
	hu_vpath vp[2];
	memset(vp, 0, sizeof vp);
	const int32_t np = (int32_t) min<size_t>(vpaths.size(), 2);
	for (int i = 0; i < np; ++i) { vp[i].start = vpaths[i].start; vp[i].end = vpaths[i].end; vp[i].from = vpaths[i].from; vp[i].to = vpaths[i].to; vp[i].nIns = vpaths[i].nIns; vp[i].nDel = vpaths[i].nDel; }
	hu_batch* b = d.one(N);
	const hu_read_batch rb = {1, 0, codes.data(), offsets, vp, &np};
	check(hu_batch_upload(b, &rb), "hu_batch_upload");
	check(hu_align_batch(b, (int) mode), "hu_align_batch"); // banded pass, full-matrix fallback when the band holds no path (HmmUFOtu_main.cpp:86-96)
	vector<char> trace((size_t) N + K + 16);
	int64_t off[2];
	check(hu_batch_get_align(b, &vs.row, trace.data(), (int64_t) trace.size(), off), "hu_batch_get_align");
	vs.seq = &seq;
	vs.trace.assign(trace.data() + off[0]);
	return vs;
}
BandedHMMP7::ViterbiScores& BandedHMMP7::calcViterbiScores(const PrimarySeq& seq, ViterbiScores& vs) const {
	return calcViterbiScores(seq, vs, vector<ViterbiAlignPath>()); // no path: the full matrix
}
void BandedHMMP7::buildViterbiTrace(const ViterbiScores& vs, ViterbiAlignTrace& vtrace) const {
	vtrace.minScore = (vs.row.flags & HU_ALIGN_INVALID) && vs.row.traceLen == 0 ? numeric_limits<double>::infinity() : vs.row.cost;
	vtrace.alnStart = vs.row.hmmStart; vtrace.alnEnd = vs.row.hmmEnd;
	vtrace.alnFrom = vs.row.seqStart; vtrace.alnTo = vs.row.seqEnd;
	vtrace.alnTrace = vs.trace;
}
BandedHMMP7::HmmAlignment BandedHMMP7::buildGlobalAlign(const PrimarySeq& seq, const ViterbiScores& vs, const ViterbiAlignTrace& vtrace) const {
	return buildGlobalAlign(seq, vs.row, vtrace.alnTrace.c_str());
}

BandedHMMP7::HmmAlignment alignSeq(const BandedHMMP7& hmm, const CSFMIndex& csfm, const PrimarySeq& read, int seedLen, int seedRegion,
		BandedHMMP7::align_mode mode) {
	// src/HmmUFOtu_main.cpp:39-105: seeds from the CSFM index (host, draws rand()), banded Viterbi with the full-matrix
	// fallback, trace, aligned string
	BandedHMMP7::ViterbiScores vs(read);
	BandedHMMP7::ViterbiAlignTrace vt;
	hmm.calcViterbiScores(read, vs, findSeeds(hmm, csfm, read, seedLen, seedRegion, mode));
	hmm.buildViterbiTrace(vs, vt);
	return hmm.buildGlobalAlign(read, vs, vt);
}

// ------------------------------------------------------------------ SEP ----
vector<PTUnrooted::PTLoc> getSeed(const PTUnrooted& ptu, const DigitalSeq& seq, int start, int end, double maxDiff, double maxHeight) {
	DeviceDB& d = deviceOf(ptu.dev, "getSeed");
	hu_batch* b = d.one(64);
	stageAligned(b, seq, start, end);
	hu_sep_params p = sepParams();
	p.maxDiff = maxDiff; p.maxHeight = maxHeight;
	check(hu_seed_batch(b, &p), "hu_seed_batch");
	int32_t n = 0;
	vector<int32_t> ids(HU_MAX_SEED);
	check(hu_batch_get_seeds(b, &n, ids.data(), NULL), "hu_batch_get_seeds");
	vector<PTUnrooted::PTLoc> locs;
	const int csLen = ptu.getCSLen();
	for (int i = 0; i < n; ++i) {
		PTUnrooted::PTLoc l;
		l.start = start; l.end = end; l.id = ids[i];
		// PTLoc.dist = SeqUtils::pDist(node seq, read) over the window (src/SeqUtils.cpp:37-54): a ratio of two small counts,
		// recounted here from the node's sequence for the <= 64 seeds the device picked
		long dd = 0, nn = 0;
		const int8_t* s = &ptu.seq[(size_t) ids[i] * csLen];
		for (int c = start; c <= end; ++c)
			if (s[c] >= 0 && seq[c] >= 0) { nn++; dd += s[c] != seq[c]; }
		l.dist = (double) dd / (double) nn;
		locs.push_back(l);
	}
	return locs;
}

vector<PTUnrooted::PTPlacement> estimateSeq(const PTUnrooted& ptu, const DigitalSeq& seq, const vector<PTUnrooted::PTLoc>& locs, const string& method) {
	DeviceDB& d = deviceOf(ptu.dev, "estimateSeq");
	vector<PTUnrooted::PTPlacement> places;
	if (locs.empty()) return places;
	if (locs.size() > HU_MAX_SEED) throw invalid_argument("estimateSeq: more than HU_MAX_SEED seeds");
	hu_batch* b = d.one(64);
	stageAligned(b, seq, locs[0].start, locs[0].end);
	hu_sep_params p = sepParams();
	p.estMethod = method == "weighted" ? HU_EST_WEIGHTED : HU_EST_UNWEIGHTED;
	const int32_t n = (int32_t) locs.size();
	vector<int32_t> ids(HU_MAX_SEED, 0);
	for (int i = 0; i < n; ++i) ids[i] = (int32_t) locs[i].id;
	check(hu_batch_set_seeds(b, &n, ids.data()), "hu_batch_set_seeds");
	check(hu_estimate_batch(b, &p), "hu_estimate_batch");
	int32_t nk = 0;
	vector<hu_place_row> rows(HU_MAX_SEED);
	check(hu_batch_get_estimates(b, &nk, rows.data()), "hu_batch_get_estimates"); // maxError = inf: all of them, sorted
	places.resize(n);
	for (int i = 0; i < n; ++i) { // back into the order of locs, which is what the reference's loop returns
		int at = -1;
		for (int k = 0; k < nk; ++k) if (rows[k].cNode == ids[i]) { at = k; break; }
		if (at < 0) throw runtime_error("estimateSeq: a seed came back without an estimate");
		places[i] = ptu.toPlacement(rows[at]);
	}
	return places;
}
PTUnrooted::PTPlacement PTUnrooted::estimateSeq(const DigitalSeq& seq, const PTLoc& loc, const string& method) const {
	return EGriceLab::HmmUFOtu::estimateSeq(*this, seq, vector<PTLoc>(1, loc), method)[0];
}

namespace {
// places <-> device rows of the one-read workspace
void putRows(hu_batch* b, const vector<PTUnrooted::PTPlacement>& places, bool asPlaces) {
	const int32_t n = (int32_t) places.size();
	vector<hu_place_row> rows(HU_MAX_SEED);
	for (int i = 0; i < n; ++i) rows[i] = places[i].toRow();
	check(asPlaces ? hu_batch_set_places(b, &n, rows.data()) : hu_batch_set_estimates(b, &n, rows.data()), "hu_batch_set_*");
}
const PTUnrooted& ownerOf(const vector<PTUnrooted::PTPlacement>& places, const char* who) {
	if (places.empty() || !places[0].owner) throw logic_error(string(who) + ": placements that did not come from estimateSeq()");
	return *places[0].owner;
}
} // namespace

vector<PTUnrooted::PTPlacement>& filterPlacements(vector<PTUnrooted::PTPlacement>& places, double maxError) {
	if (places.empty()) return places;
	if (places.size() > HU_MAX_SEED) throw invalid_argument("filterPlacements: more than HU_MAX_SEED placements");
	const PTUnrooted& ptu = ownerOf(places, "filterPlacements");
	DeviceDB& d = deviceOf(ptu.dev, "filterPlacements");
	hu_batch* b = d.one(64);
	const DigitalSeq none((size_t) ptu.getCSLen(), (int8_t) -2); // the filter reads the window, not the read
	stageAligned(b, none, places[0].start, places[0].end);
	for (size_t i = 0; i < places.size(); ++i) places[i].raw.estLoglik = places[i].loglik; // the key of the sort is the visible loglik
	putRows(b, places, false);
	hu_sep_params p = sepParams();
	p.maxError = maxError;
	check(hu_filter_batch(b, &p), "hu_filter_batch"); // std::sort(rbegin, rend, compareByLoglik) replayed on the device + the maxError cut
	int32_t nk = 0;
	vector<hu_place_row> rows(HU_MAX_SEED);
	check(hu_batch_get_estimates(b, &nk, rows.data()), "hu_batch_get_estimates");
	vector<PTUnrooted::PTPlacement> kept(nk);
	for (int k = 0; k < nk; ++k) kept[k] = ptu.toPlacement(rows[k]);
	places.swap(kept);
	return places;
}

vector<PTUnrooted::PTPlacement>& placeSeq(const PTUnrooted& ptu, const DigitalSeq& seq, vector<PTUnrooted::PTPlacement>& places, double maxHeight) {
	if (places.empty()) return places;
	if (places.size() > HU_MAX_SEED) throw invalid_argument("placeSeq: more than HU_MAX_SEED placements");
	DeviceDB& d = deviceOf(ptu.dev, "placeSeq");
	hu_batch* b = d.one(64);
	stageAligned(b, seq, places[0].start, places[0].end);
	putRows(b, places, false); // kept, in the given order
	hu_sep_params p = sepParams();
	p.maxHeight = maxHeight;
	check(hu_place_batch(b, &p), "hu_place_batch");
	int32_t nk = 0;
	vector<hu_place_row> rows(HU_MAX_SEED);
	check(hu_batch_get_places(b, &nk, rows.data()), "hu_batch_get_places");
	for (int k = 0; k < nk && k < (int) places.size(); ++k) places[k] = ptu.toPlacement(rows[k]);
	return places;
}
PTUnrooted::PTPlacement& PTUnrooted::placeSeq(const DigitalSeq& seq, PTPlacement& place, double maxHeight) const {
	vector<PTPlacement> one(1, place);
	EGriceLab::HmmUFOtu::placeSeq(*this, seq, one, maxHeight);
	place = one[0];
	return place;
}

void calcQValues(vector<PTUnrooted::PTPlacement>& places, PTUnrooted::PRIOR_TYPE type) {
	if (places.empty()) return;
	if (places.size() > HU_MAX_SEED) throw invalid_argument("calcQValues: more than HU_MAX_SEED placements");
	const PTUnrooted& ptu = ownerOf(places, "calcQValues");
	DeviceDB& d = deviceOf(ptu.dev, "calcQValues");
	hu_batch* b = d.one(64);
	const DigitalSeq none((size_t) ptu.getCSLen(), (int8_t) -2);
	stageAligned(b, none, places[0].start, places[0].end);
	putRows(b, places, true);
	hu_sep_params p = sepParams();
	p.prior = type == PTUnrooted::HEIGHT ? HU_PRIOR_HEIGHT : HU_PRIOR_UNIFORM;
	check(hu_finalize_batch(b, &p), "hu_finalize_batch");
	int32_t nk = 0;
	vector<hu_place_row> rows(HU_MAX_SEED);
	check(hu_batch_get_places(b, &nk, rows.data()), "hu_batch_get_places");
	for (int k = 0; k < nk && k < (int) places.size(); ++k) places[k] = ptu.toPlacement(rows[k]);
}

} // namespace HmmUFOtu
} // namespace EGriceLab
