/*
 * oracle_capi.cpp -- CPU ORACLE (test infrastructure, never shipped): flat C entry points
 * over oracle.h so that tests/ and bench.py's cpu_baseline / --impl reference leg can drive
 * the restatement through ctypes.  PARITY UNPINNED, see oracle.h.
 *
 * orc_assign() restates the per-read body of /root/reference/src/hmmufotu.cpp:641-733.
 */
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <atomic>
#include <thread>

using namespace orc;

namespace {
struct OrcPlaceRow { /* mirrored by tests/oracle_py.py */
	int32_t start, end;
	int64_t c, p, a;
	double wuv, ratio, wnr, loglik, loglikTrue, height, qPlace, qTaxon, estLoglik;
	int32_t outerIters, pad;
};

void fillRow(OrcPlaceRow& r, const Placement& pl, double est) {
	r.start = pl.start; r.end = pl.end; r.c = pl.c; r.p = pl.p; r.a = pl.a;
	r.wuv = pl.wuv; r.ratio = pl.ratio; r.wnr = pl.wnr; r.loglik = pl.loglik; r.loglikTrue = pl.loglikTrue;
	r.height = pl.height; r.qPlace = pl.qPlace; r.qTaxon = pl.qTaxon; r.estLoglik = est;
	r.outerIters = pl.outerIters; r.pad = 0;
}
} // namespace

extern "C" {

/* ------------------------------------------------------------------ HMM ---- */
void* orc_hmm_from_costs(int K, int L, const double* eM, const double* eI, const double* T, const int* p2cs) {
	Hmm* h = new Hmm();
	h->L = L;
	h->setProfileSize(K);
	std::copy(eM, eM + (size_t) 4 * (K + 1), h->eM.begin());
	std::copy(eI, eI + (size_t) 4 * (K + 1), h->eI.begin());
	std::copy(T, T + (size_t) 9 * (K + 1), h->T.begin());
	for (int k = 1; k <= K; ++k) {
		h->p2cs[k] = p2cs[k];
		h->cs2p[p2cs[k]] = k;
	}
	h->finalize();
	return h;
}

void* orc_hmm_load(const char* path) {
	Hmm* h = new Hmm();
	if (h->loadText(path) != 0) { delete h; return NULL; }
	return h;
}

void orc_hmm_free(void* h) { delete (Hmm*) h; }
void orc_hmm_set_mode(void* h, int mode) { ((Hmm*) h)->setSequenceMode(mode); }
int orc_hmm_K(void* h) { return ((Hmm*) h)->K; }
int orc_hmm_L(void* h) { return ((Hmm*) h)->L; }

void orc_hmm_get(void* hp, double* eM, double* eI, double* T, double* entry, double* exit_, int* p2cs, int* cs2p) {
	Hmm* h = (Hmm*) hp;
	if (eM) std::copy(h->eM.begin(), h->eM.end(), eM);
	if (eI) std::copy(h->eI.begin(), h->eI.end(), eI);
	if (T) std::copy(h->T.begin(), h->T.end(), T);
	if (entry) std::copy(h->entry.begin(), h->entry.end(), entry);
	if (exit_) std::copy(h->exit_.begin(), h->exit_.end(), exit_);
	if (p2cs) std::copy(h->p2cs.begin(), h->p2cs.begin() + h->K + 1, p2cs);
	if (cs2p) std::copy(h->cs2p.begin(), h->cs2p.begin() + h->L + 1, cs2p);
}

/* alignSeq after the seed search.  vpaths: nPaths x {start,end,from,to,nIns,nDel}.
 * coords: {seqStart,seqEnd,hmmStart,hmmEnd,csStart,csEnd}; align: L chars; trace: NUL-terminated.
 * returns 0, or -1 when the trace buffer is too small */
int orc_align(void* hp, const char* read, int n, const int* vp, int nPaths, int* coords, double* cost,
		char* align, char* trace, int traceCap, int* usedFallback, long* cells) {
	Hmm* h = (Hmm*) hp;
	std::vector<VPath> paths(nPaths);
	for (int i = 0; i < nPaths; ++i) {
		VPath p = {vp[i * 6], vp[i * 6 + 1], vp[i * 6 + 2], vp[i * 6 + 3], vp[i * 6 + 4], vp[i * 6 + 5]};
		paths[i] = p;
	}
	HmmAlignment a = alignWithPaths(*h, std::string(read, n), paths);
	coords[0] = a.seqStart; coords[1] = a.seqEnd; coords[2] = a.hmmStart; coords[3] = a.hmmEnd;
	coords[4] = a.csStart; coords[5] = a.csEnd;
	*cost = a.cost;
	if (align) memcpy(align, a.align.data(), std::min((size_t) h->L, a.align.size()));
	if (usedFallback) *usedFallback = a.usedFallback;
	if (cells) *cells = a.bandCells;
	if (trace) {
		if ((int) a.trace.size() + 1 > traceCap) return -1;
		memcpy(trace, a.trace.c_str(), a.trace.size() + 1);
	}
	return a.isValid() ? 0 : 1;
}

/* PE merge of two aligned strings (src/BandedHMMP7.cpp:1188-1213); coords/cost updated in place */
void orc_merge(int L, int* coords, double* cost, char* align, const int* ocoords, double ocost, const char* oalign) {
	HmmAlignment a, b;
	a.K = b.K = 1; a.L = b.L = L;
	a.seqStart = coords[0]; a.seqEnd = coords[1]; a.hmmStart = coords[2]; a.hmmEnd = coords[3]; a.csStart = coords[4]; a.csEnd = coords[5];
	b.seqStart = ocoords[0]; b.seqEnd = ocoords[1]; b.hmmStart = ocoords[2]; b.hmmEnd = ocoords[3]; b.csStart = ocoords[4]; b.csEnd = ocoords[5];
	a.cost = *cost; b.cost = ocost;
	a.align.assign(align, L); b.align.assign(oalign, L);
	a.merge(b);
	coords[0] = a.seqStart; coords[1] = a.seqEnd; coords[2] = a.hmmStart; coords[3] = a.hmmEnd; coords[4] = a.csStart; coords[5] = a.csEnd;
	*cost = a.cost;
	memcpy(align, a.align.data(), L);
}

/* DigitalSeq of an aligned string (src/DigitalSeq.cpp:41-48); returns length written */
int orc_digitize(const char* align, int n, int8_t* out) {
	std::vector<int8_t> v = digitalSeq(std::string(align, n));
	std::copy(v.begin(), v.end(), out);
	return (int) v.size();
}

/* ---------------------------------------------------------------- model ---- */
void* orc_model_create(int type, const double* pi, const double* R, double kr, double ky, double kappa, double beta) {
	Model* m = new Model();
	m->type = type;
	for (int i = 0; i < 4; ++i) m->pi[i] = pi[i];
	if (R) for (int i = 0; i < 16; ++i) m->R[i] = R[i];
	m->kr = kr; m->ky = ky; m->kappa = kappa; m->beta = beta;
	m->setup();
	return m;
}
void* orc_model_load_text(const char* text) {
	Model* m = new Model();
	std::istringstream in(text);
	if (m->loadText(in) != 0) { delete m; return NULL; }
	return m;
}
void orc_model_free(void* m) { delete (Model*) m; }
void orc_model_Pr(void* m, double v, double* P) { ((Model*) m)->Pr(v, P); }
void orc_model_pi(void* m, double* pi) { for (int i = 0; i < 4; ++i) pi[i] = ((Model*) m)->pi[i]; }
int orc_model_type(void* m) { return ((Model*) m)->type; }

/* ----------------------------------------------------------------- tree ---- */
struct OrcTree {
	Tree t;
	std::vector<std::vector<long> > children;
	std::vector<double> rootMsg;
};

void* orc_tree_create(long nNodes, int csLen, long root, const int* parent, const double* blen, const double* height,
		const int8_t* seq, const int* taxon, void* model, int dgK, const double* dgRates) {
	OrcTree* o = new OrcTree();
	Tree& t = o->t;
	t.nNodes = nNodes; t.csLen = csLen; t.root = root;
	t.parent.assign(parent, parent + nNodes);
	t.blen.assign(blen, blen + nNodes);
	t.height.assign(height, height + nNodes);
	t.seq.assign(seq, seq + (size_t) nNodes * csLen);
	t.anno.resize(nNodes);
	for (long i = 0; i < nNodes; ++i) {
		char buf[32];
		snprintf(buf, sizeof buf, "taxon%d", taxon ? taxon[i] : (int) i);
		t.anno[i] = buf;
	}
	t.model = *(Model*) model;
	t.dg.K = dgK;
	if (dgK) t.dg.r.assign(dgRates, dgRates + dgK);
	o->children.resize(nNodes);
	for (long i = 0; i < nNodes; ++i)
		if (t.parent[i] >= 0) o->children[t.parent[i]].push_back(i);
	return o;
}
void orc_tree_free(void* o) { delete (OrcTree*) o; }

/* all-roots evaluation; with inferInternal the internal-node sequences (and the root's) are replaced
 * by the arg-max of their upward message (src/PhyloTreeUnrooted.cpp:1085-1093) */
void orc_tree_evaluate_mt(void* op, int inferInternal, int nThreads);
void orc_tree_evaluate(void* op, int inferInternal) { orc_tree_evaluate_mt(op, inferInternal, 1); }

void orc_tree_evaluate_mt(void* op, int inferInternal, int nThreads) {
	OrcTree* o = (OrcTree*) op;
	Tree& t = o->t;
	if (nThreads <= 1)
		t.evaluateAll(o->children);
	else { /* column stripes are independent */
		t.up.assign((size_t) t.nNodes * t.csLen * 4, 0.0);
		t.dn.assign((size_t) t.nNodes * t.csLen * 4, 0.0);
		std::vector<std::thread> th;
		const int per = (t.csLen + nThreads - 1) / nThreads;
		for (int i = 0; i < nThreads; ++i) {
			const int j0 = i * per, j1 = std::min(t.csLen, j0 + per);
			if (j0 >= j1) break;
			th.push_back(std::thread([&t, o, j0, j1]() { t.evaluateAll(o->children, j0, j1, false); }));
		}
		for (size_t i = 0; i < th.size(); ++i) th[i].join();
	}
	/* root message: loglik(root, j), src/PhyloTreeUnrooted.cpp:320-346 */
	const int KK = std::max(t.dg.K, 1);
	o->rootMsg.assign((size_t) t.csLen * 4, 0.0);
	std::vector<double> mat((size_t) 4 * KK), PP;
	for (int j = 0; j < t.csLen; ++j) {
		std::fill(mat.begin(), mat.end(), 0.0);
		for (size_t ci = 0; ci < o->children[t.root].size(); ++ci) {
			long ch = o->children[t.root][ci];
			for (int k = 0; k < KK; ++k) {
				double P[16], c[4];
				t.model.Pr(t.blen[ch] * (t.dg.K ? t.dg.r[k] : 1.0), P);
				dpsMat(P, t.U(ch) + (size_t) j * 4, c);
				for (int i = 0; i < 4; ++i) mat[(size_t) k * 4 + i] += c[i];
			}
		}
		double* out = &o->rootMsg[(size_t) j * 4];
		if (t.dg.K) {
			for (int i = 0; i < 4; ++i) {
				double mx = mat[i];
				for (int k = 1; k < KK; ++k) mx = std::max(mx, mat[(size_t) k * 4 + i]);
				double sc = (mx != NEG_INF && mx < -510.0) ? -510.0 - mx : 0;
				double s = ::exp(mat[i] + sc);
				for (int k = 1; k < KK; ++k) s = s + ::exp(mat[(size_t) k * 4 + i] + sc);
				out[i] = ::log(s / KK) - sc;
			}
		} else
			for (int i = 0; i < 4; ++i) out[i] = mat[i];
	}
	if (inferInternal) {
		for (long x = 0; x < t.nNodes; ++x) {
			if (o->children[x].empty()) continue;
			const double* m = x == t.root ? &o->rootMsg[0] : t.U(x);
			for (int j = 0; j < t.csLen; ++j) {
				const double* v = m + (size_t) j * 4;
				int b = 0;
				for (int k = 1; k < 4; ++k) if (v[k] > v[b]) b = k;
				t.seq[(size_t) x * t.csLen + j] = (int8_t) b;
			}
		}
	}
}

void orc_tree_get(void* op, double* up, double* dn, int8_t* seq) {
	Tree& t = ((OrcTree*) op)->t;
	if (up) std::copy(t.up.begin(), t.up.end(), up);
	if (dn) std::copy(t.dn.begin(), t.dn.end(), dn);
	if (seq) std::copy(t.seq.begin(), t.seq.end(), seq);
}
void orc_tree_set_messages(void* op, const double* up, const double* dn) {
	Tree& t = ((OrcTree*) op)->t;
	size_t n = (size_t) t.nNodes * t.csLen * 4;
	t.up.assign(up, up + n);
	t.dn.assign(dn, dn + n);
}

/* messages of a subset of the nodes only (ids[i] -> row i of up / dn); every other node has none */
void orc_tree_set_sparse_messages(void* op, long nSel, const long* ids, const double* up, const double* dn) {
	Tree& t = ((OrcTree*) op)->t;
	const size_t row = (size_t) t.csLen * 4;
	t.slotOf.assign(t.nNodes, -1);
	for (long i = 0; i < nSel; ++i) t.slotOf[ids[i]] = i;
	t.up.assign(up, up + (size_t) nSel * row);
	t.dn.assign(dn, dn + (size_t) nSel * row);
}

/* getSeed + truncation to maxN (src/HmmUFOtu_main.cpp:127-152, src/hmmufotu.cpp:646-647) */
int orc_get_seed(void* op, const int8_t* seq, int start, int end, double maxDiff, double maxHeight, int seedOrder, int maxN,
		long* ids, double* dist) {
	Tree& t = ((OrcTree*) op)->t;
	std::vector<PTLoc> s = getSeed(t, seq, start, end, maxDiff, maxHeight, seedOrder);
	if ((int) s.size() > maxN) s.resize(maxN);
	for (size_t i = 0; i < s.size(); ++i) { ids[i] = s[i].id; dist[i] = s[i].dist; }
	return (int) s.size();
}

long orc_pdist(void* op, long node, const int8_t* seq, int start, int end, long* N) {
	Tree& t = ((OrcTree*) op)->t;
	long d = 0;
	pDist(t.s(node), seq, start, end, &d, N);
	return d;
}

void orc_estimate(void* op, const int8_t* seq, int start, int end, long id, double dist, OrcPlaceRow* out) {
	Tree& t = ((OrcTree*) op)->t;
	PTLoc loc = {start, end, id, dist};
	Placement pl = estimateSeq(t, seq, loc);
	fillRow(*out, pl, pl.loglik);
}

void orc_place(void* op, const int8_t* seq, OrcPlaceRow* io, double maxHeight) {
	Tree& t = ((OrcTree*) op)->t;
	Placement pl;
	pl.start = io->start; pl.end = io->end; pl.c = io->c; pl.p = io->p; pl.a = io->a;
	pl.wuv = io->wuv; pl.ratio = io->ratio; pl.wnr = io->wnr; pl.loglik = io->loglik;
	double est = io->loglik;
	placeSeq(t, seq, pl, maxHeight);
	fillRow(*io, pl, est);
}

/* per-read SEP body, src/hmmufotu.cpp:645-733.  seeds given (from orc_get_seed or the device);
 * useTrue != 0 substitutes the intended root loglik for the literal constant.
 * rows: up to nSeeds entries, post-sort order; returns n (number of placements kept); rows[0] is bestPlace */
int orc_assign(void* op, const int8_t* seq, int start, int end, int nSeeds, const long* ids, const double* dist,
		double maxError, double maxHeight, int useTrue, OrcPlaceRow* rows) {
	Tree& t = ((OrcTree*) op)->t;
	std::vector<Placement> places;
	for (int i = 0; i < nSeeds; ++i) {
		PTLoc loc = {start, end, ids[i], dist[i]};
		places.push_back(estimateSeq(t, seq, loc));
	}
	filterPlacements(places, maxError);
	std::vector<double> est(places.size());
	for (size_t i = 0; i < places.size(); ++i) {
		est[i] = places[i].loglik;
		places[i].outerIters = (int) i; /* temporarily: index into est[] */
	}
	std::vector<int> iters(places.size());
	for (size_t i = 0; i < places.size(); ++i) {
		int idx = places[i].outerIters;
		placeSeq(t, seq, places[i], maxHeight);
		iters[i] = places[i].outerIters;
		places[i].outerIters = idx;
		if (useTrue) places[i].loglik = places[i].loglikTrue;
	}
	if (t.onlyML) sortByLoglik(places);
	else {
		calcQValues(t, places);
		sortByQPlace(places);
	}
	for (size_t i = 0; i < places.size(); ++i) {
		int idx = places[i].outerIters;
		places[i].outerIters = iters[idx];
		fillRow(rows[i], places[i], est[idx]);
	}
	return (int) places.size();
}

void orc_tree_set_options(void* op, int estWeighted, int priorHeight, int onlyML) {
	Tree& t = ((OrcTree*) op)->t;
	t.estWeighted = estWeighted; t.priorHeight = priorHeight; t.onlyML = onlyML;
}

/* permutation std::sort(rbegin, rend) leaves on an all-equal array of n elements: out[i] = old index now at i */
void orc_equal_sort_perm(int n, int* out) {
	std::vector<std::pair<double, int> > v(n);
	for (int i = 0; i < n; ++i) v[i] = std::make_pair(0.0, i);
	struct Cmp { static bool lt(const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; } };
	std::sort(v.rbegin(), v.rend(), Cmp::lt);
	for (int i = 0; i < n; ++i) out[i] = v[i].second;
}

/* The per-read body of the reference's task loop (src/hmmufotu.cpp:610-748) for a list of reads, one std::thread per core
 * pulling reads from a shared counter (the reference uses one OpenMP task per read).  paired != 0: sequences nReads..2 nReads-1
 * are the already reverse-complemented mates (:609); orientation test and merge as :629-639 (a failed test = chimera: the
 * read is not placed, out[r].c = -3).  The tree may hold a column slice [col0, col0 + csLen) of the consensus; aligned reads
 * are cut to it.  seedOrder 0 = std::sort as in the reference.  out: one best-placement row per read (c = -1: invalid
 * alignment, -2: a seed's messages are not in a sparse tree); coords: 6 ints + cost; nearTie (may be NULL): 1 when two kept
 * estimated log-likelihoods agree to 1e-11 relative, i.e. their order -- hence, under SURVEY.md F6, the reported placement --
 * is decided by rounding noise. */
int orc_run_reads(void* hp, void* op, int nReads, int paired, const char* reads, const long* offsets, const int* vpaths, const int* nPaths,
		int col0, int maxNSeed, double maxError, int seedOrder, int useTrue, int nThreads, OrcPlaceRow* out, int* coords, double* cost,
		int* nearTie) {
	Hmm* h = (Hmm*) hp;
	Tree& t = ((OrcTree*) op)->t;
	std::atomic<int> next(0), bad(0);
	auto alignOne = [&](int s) {
		std::vector<VPath> paths(nPaths[s]);
		for (int i = 0; i < nPaths[s]; ++i) {
			const int* v = vpaths + ((size_t) s * 2 + i) * 6;
			VPath p = {v[0], v[1], v[2], v[3], v[4], v[5]};
			paths[i] = p;
		}
		return alignWithPaths(*h, std::string(reads + offsets[s], offsets[s + 1] - offsets[s]), paths);
	};
	auto work = [&]() {
		for (;;) {
			const int r = next.fetch_add(1);
			if (r >= nReads) return;
			if (nearTie) nearTie[r] = 0;
			HmmAlignment a = alignOne(r);
			bool chimera = false;
			if (paired) {
				HmmAlignment b = alignOne(nReads + r);
				if (!a.isValid() || !b.isValid() || !(a.csStart <= b.csStart && a.csEnd <= b.csEnd)) chimera = true;
				else a.merge(b);
			}
			if (coords) {
				int* c = coords + (size_t) r * 6;
				c[0] = a.seqStart; c[1] = a.seqEnd; c[2] = a.hmmStart; c[3] = a.hmmEnd; c[4] = a.csStart; c[5] = a.csEnd;
			}
			if (cost) cost[r] = a.cost;
			if (chimera) { memset(&out[r], 0, sizeof(OrcPlaceRow)); out[r].c = -3; continue; }
			std::vector<int8_t> seq = digitalSeq(a.align);
			const int s = a.csStart - 1 - col0, e = a.csEnd - 1 - col0;
			if (!a.isValid() || s < 0 || e >= t.csLen) { bad++; memset(&out[r], 0, sizeof(OrcPlaceRow)); out[r].c = -1; continue; }
			const int8_t* sq = seq.data() + col0;
			std::vector<PTLoc> seeds = getSeed(t, sq, s, e, INF, INF, seedOrder);
			if ((int) seeds.size() > maxNSeed) seeds.resize(maxNSeed);
			bool missing = false;
			for (size_t i = 0; i < seeds.size(); ++i) missing |= !t.hasMsg(seeds[i].id);
			if (missing) { bad++; memset(&out[r], 0, sizeof(OrcPlaceRow)); out[r].c = -2; continue; }
			std::vector<Placement> places;
			for (size_t i = 0; i < seeds.size(); ++i) places.push_back(estimateSeq(t, sq, seeds[i]));
			filterPlacements(places, maxError);
			if (nearTie)
				for (size_t i = 0; i + 1 < places.size(); ++i)
					if (std::fabs(places[i].loglik - places[i + 1].loglik) <= 1e-11 * std::fabs(places[i].loglik)) nearTie[r] = 1;
			for (size_t i = 0; i < places.size(); ++i) {
				placeSeq(t, sq, places[i], INF);
				if (useTrue) places[i].loglik = places[i].loglikTrue;
			}
			calcQValues(t, places);
			sortByQPlace(places);
			fillRow(out[r], places[0], 0.0);
			out[r].pad = (int) places.size();
		}
	};
	std::vector<std::thread> th;
	for (int i = 0; i < std::max(1, nThreads); ++i) th.push_back(std::thread(work));
	for (size_t i = 0; i < th.size(); ++i) th[i].join();
	return bad.load();
}

int orc_sizeof_place_row() { return (int) sizeof(OrcPlaceRow); }

/* ---- the real std::sort of this toolchain's libstdc++, for pinning the device's replay of it ---- */
/* std::sort(v.rbegin(), v.rend(), byKey) as filterPlacements / the final ordering call it (src/HmmUFOtu_main.cpp:164,
 * src/hmmufotu.cpp:725-731): perm[k] = original index of the element left at position k */
void orc_stdsort_desc_perm(const double* keys, int n, int* perm) {
	std::vector<std::pair<double, int> > v(n);
	for (int i = 0; i < n; ++i) v[i] = std::make_pair(keys[i], i);
	struct Cmp { static bool lt(const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; } };
	std::sort(v.rbegin(), v.rend(), Cmp::lt);
	for (int i = 0; i < n; ++i) perm[i] = v[i].second;
}
/* std::sort(locs.begin(), locs.end()) as getSeed calls it (src/HmmUFOtu_main.cpp:139) on integer keys that order like the
 * p-distances: the candidate positions of the first `want` elements */
void orc_stdsort_keys(const unsigned int* keys, int n, int want, int* outPos) {
	std::vector<PTLoc> v(n);
	for (int i = 0; i < n; ++i) { PTLoc l = {0, 0, i, (double) keys[i]}; v[i] = l; }
	std::sort(v.begin(), v.end());
	for (int i = 0; i < want && i < n; ++i) outPos[i] = (int) v[i].id;
}
/* McIlroy's adversary ("A killer adversary for quicksort", 1999) run against this libstdc++'s std::sort: a key vector on
 * which its median-of-3 quicksort degenerates, so that __introsort_loop spends its depth budget and falls back to
 * __partial_sort (heap sort).  keys: n values (a permutation of 0..n-1 up to ties among the last few) */
namespace {
struct Adversary {
	std::vector<int> val;
	int nsolid, candidate, gas;
	bool lt(int x, int y) {
		if (val[x] == gas && val[y] == gas) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
		if (val[x] == gas) candidate = x; else if (val[y] == gas) candidate = y;
		return val[x] < val[y];
	}
};
Adversary* g_adv = NULL;
bool advLess(int a, int b) { return g_adv->lt(a, b); }
}
void orc_sort_killer(int n, unsigned int* keys) {
	Adversary adv;
	adv.val.assign(n, n - 1); adv.nsolid = 0; adv.candidate = 0; adv.gas = n - 1;
	std::vector<int> ptr(n);
	for (int i = 0; i < n; ++i) ptr[i] = i;
	g_adv = &adv;
	std::sort(ptr.begin(), ptr.end(), advLess);
	g_adv = NULL;
	for (int i = 0; i < n; ++i) keys[i] = (unsigned int) adv.val[i];
}

} // extern "C"
