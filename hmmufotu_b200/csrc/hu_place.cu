// K4: the "Place" stage of SEP -- PTUnrooted::placeSeq on the private 4-node tree {v, u, r, n}
// (src/PhyloTreeUnrooted.cpp:879-954), i.e. the joint branch-length optimisation (cpp:800-847) built on
// Felsenstein's single-branch EM (cpp:749-798) and the dGamma-mixed message rebuilds (cpp:320-346),
// then calcQValues + the final sort (src/HmmUFOtu_main.cpp:182-216, src/hmmufotu.cpp:725-733).
//
// One warp, or one four-warp CTA, per (read, placement), persistent over the item space.  Everything is done in the linear domain
// (hu_internal.cuh, expNorm): with e^U_j, e^V_j the max-normalised exponentials of the two stored messages,
//   msg(r->n)_j ~ mean_k (P(lu r_k) e^U_j) * (P(lv r_k) e^V_j)           (the per-column scales cancel below)
//   rho_j = e^{A_j - B_j} = pi.(msg * n_j) / ((pi.msg) (pi.n_j))          A, B of cpp:769-770
// so a message rebuild costs 8 exp + a few dozen FMAs per column and no log, and since rho_j does not change
// inside one EM call (the reference recomputes A_j, B_j every iteration) an EM iteration is one fused
// multiply-add and one reciprocal per column plus a warp reduction:  p <- mean_j p0 / (rho_j q0 + p0).
// Same iterates, same stopping rules, same caps as the reference; the values differ from it only by rounding
// (checked to 1e-9 relative against the oracle in tests/test_gpu_parity.py).
#include "hu_batch.cuh"
#include "hu_stdsort.cuh"
#include <stdlib.h>

namespace hu {

#define HU_PLACE_WARPS 4
#define HU_PLACE_THREADS (HU_PLACE_WARPS * 32)
#define HU_PLACE_CAPW 1536 // columns of rho kept in shared memory per group of warps (12 KB); wider windows use a global slab
#define HU_BRANCH_EPS 1e-5 // src/PhyloTreeUnrooted.cpp:70
#define HU_MAX_ITER 100    // src/PhyloTreeUnrooted.cpp:71

struct PlaceArgs {
	int nReads, maxNSeed, capW, useTrue;
	int ahead, pad;    // L2 prefetch distance of the rebuild passes, in column strides
	double maxHeight;
	const int8_t* aligned;
	const int* winStart;
	const int* winEnd;
	const hu_place_row* est;
	const int* order;
	const int* nKept;
	hu_place_row* places;
	int* cursor;
	double* slab;      // [warp slots][csLen]
};

// per-warp shared memory: P[3][KK][16] = P(lu r_k), P(lv r_k), P(ln r_k); NP[KK][5][4] = P(ln r_k) applied to the
// leaf vector of base b / gap (linear); sized by the actual number of rate categories KK
__host__ __device__ inline int placeWarpDoubles(int KK) { return 3 * KK * 16 + KK * 20; }

// register double buffer over the columns of one lane: the loads of column c+32 are in flight while column c
// is being worked on (the rebuild loops were stalled on these loads: 12 warps per SM cannot hide HBM latency)
struct ColPrefetch {
	double eu[4], ev[4];
	int b;
	__device__ __forceinline__ void init(const double* U, const double* V, const int8_t* x, int c, int W) {
		b = 0;
		if (c < W) { ld4(U + (size_t) c * 4, eu); ld4(V + (size_t) c * 4, ev); b = x[c]; }
	}
	// hands out the buffered column and starts loading column `cn`
	__device__ __forceinline__ int next(const double* U, const double* V, const int8_t* x, int cn, int W, double* ou, double* ov, int ahead) {
		for (int k = 0; k < 4; ++k) { ou[k] = eu[k]; ov[k] = ev[k]; }
		const int ob = b;
		if (cn < W) { ld4(U + (size_t) cn * 4, eu); ld4(V + (size_t) cn * 4, ev); b = x[cn]; }
		// the columns after those: on their way from DRAM to the L2 (one request per 128-byte line; no register is held for them)
		if (ahead > 0 && (threadIdx.x & 3) == 0 && cn + ahead < W) {
			asm volatile("prefetch.global.L2 [%0];" ::"l"(U + (size_t) (cn + ahead) * 4));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(V + (size_t) (cn + ahead) * 4));
		}
		return ob;
	}
};

// The transition matrices of one phase, by ONE out-of-line routine: lane 8 j + k builds matrix j (0: lu, 1: lv,
// 2: ln) for rate category k.  Inlining the six-model switch at every call site made the kernel 180 KB of
// code and the warps, each in a different phase, missed the instruction cache (ncu r1b: stall_no_instruction
// 4.1 cycles per issue).  The model is read from its global-memory copy: taking the address of the by-value
// kernel parameter would force a local-memory copy of it.
__device__ __noinline__ void fillP(const Model* __restrict__ mg, double* __restrict__ wsP, int KK, double l0, double l1, double l2, int mask) {
	const int lane = threadIdx.x & 31;
	const int job = lane >> 3, k = lane & 7;
	if (job < 3 && k < KK && ((mask >> job) & 1)) {
		const double len = job == 0 ? l0 : job == 1 ? l1 : l2;
		modelPr(*mg, len * (mg->dgK ? mg->dgR[k] : 1.0), wsP + (job * KK + k) * 16);
	}
}

__device__ __forceinline__ double warpSum(double v) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

#define HU_RHO_SKIP 1e300 // stands for a skipped column (NaN in the reference's test): its EM term is ~1e-300

// 1/x: hardware fp64 reciprocal approximation (about 2^-20 relative) + one Newton step -> below 1e-12 relative,
// no branches.  x is rho q0 + p0 with 0 <= rho <= 1e300, 0 < p0, q0 <= 1.
__device__ __forceinline__ double rcpFast(double x) {
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	return fma(r, fma(-x, r, 1.0), r);
}

// An item is worked on by a GROUP of WPI warps (1 or 4) of a four-warp CTA.  Group-wide sums in a fixed order (lane tree inside
// a warp, then the warp totals pairwise): every thread of the group gets the same bits, so the iteration control that depends
// on them stays uniform.  `red` is double buffered: one barrier per sum.
struct PlaceRed {
	double d[2][HU_PLACE_WARPS];
	int i[2][HU_PLACE_WARPS];
};
template <int WPI>
__device__ __forceinline__ void groupSync(int group) {
	if (WPI == 1) __syncwarp();
	else if (WPI == 4 || group == 0) asm volatile("bar.sync 1, %0;" ::"n"(WPI * 32) : "memory");
	else asm volatile("bar.sync 2, %0;" ::"n"(WPI * 32) : "memory");
}
template <int WPI>
__device__ __forceinline__ double groupSum(double v, PlaceRed& red, int& flip, int group, int gwarp) {
	v = warpSum(v);
	if (WPI == 1) return v;
	if ((threadIdx.x & 31) == 0) red.d[flip][gwarp] = v;
	groupSync<WPI>(group);
	const double s = WPI == 2 ? red.d[flip][0] + red.d[flip][1] : (red.d[flip][0] + red.d[flip][1]) + (red.d[flip][2] + red.d[flip][3]);
	flip ^= 1;
	return s;
}
template <int WPI>
__device__ __forceinline__ int groupSumInt(int v, PlaceRed& red, int& flip, int group, int gwarp) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if (WPI == 1) { __syncwarp(); return v; }
	if ((threadIdx.x & 31) == 0) red.i[flip][gwarp] = v;
	groupSync<WPI>(group);
	const int s = WPI == 2 ? red.i[flip][0] + red.i[flip][1] : (red.i[flip][0] + red.i[flip][1]) + (red.i[flip][2] + red.i[flip][3]);
	flip ^= 1;
	return s;
}

// Felsenstein EM for one branch given rho_j, src/PhyloTreeUnrooted.cpp:749-798:
// p <- (1/N) sum_j e^B p0 / (e^A q0 + e^B p0) = (1/N) sum_j p0 / (rho_j q0 + p0); the whole group works on one branch
template <int WPI>
__device__ __noinline__ double emBranch(const double* __restrict__ rho, int W, int nValid, double wInit, double cap, int* itersOut, PlaceRed& red, int& flip,
		int group, int gtid) {
	const int NT = WPI * 32;
	double q0 = exp(-wInit), p0 = 1 - q0;
	double p = p0, q = q0;
	int it = 0;
	for (it = 0; it < HU_MAX_ITER && p >= 0 && p <= 1; ++it) {
		double s = 0;
		if (p0 > 1e-200 && q0 >= 0) {
			double s1 = 0; // two accumulators: independent dependency chains
			int c = gtid;
			for (; c + NT < W; c += 2 * NT) {
				s = fma(p0, rcpFast(fma(rho[c], q0, p0)), s);
				s1 = fma(p0, rcpFast(fma(rho[c + NT], q0, p0)), s1);
			}
			if (c < W) s = fma(p0, rcpFast(fma(rho[c], q0, p0)), s);
			s += s1;
		} else { // degenerate branch length: plain division keeps the reference's NaN / inf behaviour
			for (int c = gtid; c < W; c += NT) {
				const double rh = rho[c];
				if (rh < HU_RHO_SKIP) s += p0 / (rh * q0 + p0);
			}
		}
		p = groupSum<WPI>(s, red, flip, group, gtid >> 5) / (double) nValid;
		q = 1 - p;
		if (fabs(log(q) - log(q0)) < HU_BRANCH_EPS) { it++; break; }
		p0 = p; q0 = q;
	}
	*itersOut += it;
	double w = -log(q);
	if (w > cap) w = cap;
	return w;
}

// WPI warps per (read, placement), persistent over the item space.  WPI = 1 (amplicon reads, merged pairs): one warp per item,
// four items per CTA, no block barriers.  WPI = 4 (long reads, windows of thousands of columns): the columns of an item are
// spread over the four warps, so that a pass over the window is W / 128 columns per thread and the items in flight -- hence
// the message windows that are re-streamed 2 iters times -- are four times fewer (they stay in L2 between the passes); the
// price is one named barrier per EM iteration.  Measured (profiles/r2_summary.md): C3 6.2 ms (WPI 1) vs 8.9 (WPI 4); C5 9.2 vs 5.7.
template <int WPI>
__global__ void __launch_bounds__(HU_PLACE_THREADS, 4) placeKernel(PlaceArgs a, TreeDev t, Model m) {
	extern __shared__ __align__(16) unsigned char smemRaw[];
	constexpr int NT = WPI * 32, GROUPS = HU_PLACE_WARPS / WPI;
	__shared__ PlaceRed redAll[GROUPS];
	__shared__ int sItem[GROUPS];
	const int group = (threadIdx.x >> 5) / WPI, gtid = threadIdx.x - group * NT, gwarp = gtid >> 5;
	PlaceRed& red = redAll[group];
	const int KK = m.dgK > 0 ? m.dgK : 1;
	const int groupDoubles = placeWarpDoubles(KK) + a.capW;
	double* wsP = reinterpret_cast<double*>(smemRaw) + (size_t) group * groupDoubles; // P[3][KK][16] = P(lu r_k), P(lv r_k), P(ln r_k)
	double* wsNP = wsP + 3 * KK * 16;                    // NP[KK][5][4] = P(ln r_k) applied to the leaf vector of base b / gap
	double* rhoS = wsP + placeWarpDoubles(KK);
	const int slot = blockIdx.x * GROUPS + group;
	const int total = a.nReads * a.maxNSeed;
	int flip = 0;
	double piSq = 0; // pi . pi: (pi . n) for a gap column of the read
	for (int k = 0; k < 4; ++k) piSq += m.pi[k] * m.pi[k];
	for (;;) {
		groupSync<WPI>(group);
		if (gtid == 0) sItem[group] = atomicAdd(a.cursor, 1);
		groupSync<WPI>(group);
		const int item = sItem[group];
		if (item >= total) return;
		const int r = item / a.maxNSeed, idx = item % a.maxNSeed;
		if (a.winStart[r] < 0 || idx >= a.nKept[r]) continue;
		const hu_place_row e = a.est[(size_t) r * HU_MAX_SEED + a.order[(size_t) r * HU_MAX_SEED + idx]];
		const int start = e.start, end = e.end, W = end - start + 1;
		const int c = e.cNode, pN = e.pNode;
		const double w0 = t.blen[c];
		const double* U = t.up + ((size_t) c * t.csLen + start) * 4;
		const double* V = t.dn + ((size_t) c * t.csLen + start) * 4;
		const int8_t* x = a.aligned + (size_t) r * t.csLen + start;
		double* rho = W <= a.capW ? rhoS : a.slab + (size_t) slot * t.csLen;
		double lu = w0 * e.ratio, lv = w0 * (1 - e.ratio), ln = e.wnr; // cpp:906-912
		double wur0 = lu, wvr0 = lv, wnr0 = ln;
		const double w0j = wur0 + wvr0;
		double wur = wur0, wvr = wvr0, wnr = wnr0;
		int iter = 0, emIters = 0;
		for (iter = 0; iter < HU_MAX_ITER && 0 <= wur && wur <= w0j; ++iter) {
			// ---- msg(r->n) from u and v, then optimise the n-r branch (cap 1) ----
			groupSync<WPI>(group);
			if (gwarp == 0) fillP(t.modelDev, wsP, KK, lu, lv, ln, 3);
			groupSync<WPI>(group);
			int nValid = 0;
			ColPrefetch pf;
			pf.init(U, V, x, gtid, W);
			for (int cI = gtid; cI < W; cI += NT) {
				double eu[4], ev[4], ya[4], yb[4], g[4] = {0, 0, 0, 0};
				const int b = pf.next(U, V, x, cI + NT, W, eu, ev, a.ahead * NT);
				for (int k = 0; k < KK; ++k) {
					matVec4(wsP + (0 * KK + (k)) * 16, eu, ya);
					matVec4(wsP + (1 * KK + (k)) * 16, ev, yb);
					for (int i = 0; i < 4; ++i) g[i] += ya[i] * yb[i];
				}
				// A = log pi.(g n), B = log pi.g + log pi.n; the 1/K of the rate mixture cancels
				const double num = b >= 0 ? m.pi[b] * g[b] : ((m.pi[0] * m.pi[0] * g[0] + m.pi[2] * m.pi[2] * g[2]) + (m.pi[1] * m.pi[1] * g[1] + m.pi[3] * m.pi[3] * g[3]));
				const double den = dotPi(m, g) * (b >= 0 ? m.pi[b] : piSq);
				const double rh = num / den;
				nValid += rh == rh;
				rho[cI] = rh == rh ? fmin(rh, HU_RHO_SKIP) : HU_RHO_SKIP;
			}
			nValid = groupSumInt<WPI>(nValid, red, flip, group, gwarp); // its barrier also publishes rho
			wnr = emBranch<WPI>(rho, W, nValid, ln, 1.0, &emIters, red, flip, group, gtid);
			ln = wnr;
			// ---- msg(r->u) from v and n, then optimise the u-r branch (cap w0) ----
			groupSync<WPI>(group);
			if (gwarp == 0) fillP(t.modelDev, wsP, KK, lu, lv, ln, 4);
			groupSync<WPI>(group);
			for (int q = gtid; q < KK * 5; q += NT) {
				const int k = q / 5, b = q % 5;
				double n[4], y[4];
				leafLin(m, b < 4 ? b : -1, n);
				matVec4(wsP + (2 * KK + (k)) * 16, n, y);
				for (int i = 0; i < 4; ++i) wsNP[(k * 5 + b) * 4 + i] = y[i];
			}
			groupSync<WPI>(group);
			nValid = 0;
			pf.init(U, V, x, gtid, W);
			for (int cI = gtid; cI < W; cI += NT) {
				double eu[4], ev[4], yb[4], g[4] = {0, 0, 0, 0};
				const int b = pf.next(U, V, x, cI + NT, W, eu, ev, a.ahead * NT);
				for (int k = 0; k < KK; ++k) {
					matVec4(wsP + (1 * KK + (k)) * 16, ev, yb);
					const double* np = wsNP + (k * 5 + (b >= 0 ? b : 4)) * 4;
					for (int i = 0; i < 4; ++i) g[i] += yb[i] * np[i];
				}
				double z[4];
				for (int i = 0; i < 4; ++i) z[i] = g[i] * eu[i];
				const double rh = dotPi(m, z) / (dotPi(m, g) * dotPi(m, eu));
				nValid += rh == rh;
				rho[cI] = rh == rh ? fmin(rh, HU_RHO_SKIP) : HU_RHO_SKIP;
			}
			nValid = groupSumInt<WPI>(nValid, red, flip, group, gwarp);
			wur = emBranch<WPI>(rho, W, nValid, lu, w0j, &emIters, red, flip, group, gtid);
			lu = wur;
			wvr = w0j - wur;
			lv = wvr; // msg(r->v) is rebuilt by the reference but never read afterwards
			if (fabs(wur - wur0) < HU_BRANCH_EPS && fabs(wnr - wnr0) < HU_BRANCH_EPS) { // ::abs resolves to the double overload here
				iter++;
				break;
			}
			wur0 = wur; wvr0 = wvr; wnr0 = wnr;
		}
		// ---- intended root log-likelihood: sum_j log(pi . exp(loglik(r, j))) ----
		// One more pass over the two message windows: only when it is asked for (the literal placeSeq value, SURVEY.md F6,
		// needs none of it)
		double tl = 0;
		if (a.useTrue) {
			groupSync<WPI>(group);
			if (gwarp == 0) fillP(t.modelDev, wsP, KK, lu, lv, ln, 7);
			groupSync<WPI>(group);
			for (int q = gtid; q < KK * 5; q += NT) {
				const int k = q / 5, b = q % 5;
				double n[4], y[4];
				leafLin(m, b < 4 ? b : -1, n);
				matVec4(wsP + (2 * KK + (k)) * 16, n, y);
				for (int i = 0; i < 4; ++i) wsNP[(k * 5 + b) * 4 + i] = y[i];
			}
			groupSync<WPI>(group);
			if (gwarp == 0)
				tl = windowScale(t.upS, t.upP32, c, t.csLen, t.nWords, start, end, gtid)
				   + windowScale(t.dnS, t.dnP32, c, t.csLen, t.nWords, start, end, gtid);
			ColPrefetch pf;
			pf.init(U, V, x, gtid, W);
			for (int cI = gtid; cI < W; cI += NT) {
				double eu[4], ev[4], ya[4], yb[4], g[4] = {0, 0, 0, 0};
				const int b = pf.next(U, V, x, cI + NT, W, eu, ev, a.ahead * NT);
				for (int k = 0; k < KK; ++k) {
					matVec4(wsP + (0 * KK + (k)) * 16, eu, ya);
					matVec4(wsP + (1 * KK + (k)) * 16, ev, yb);
					const double* np = wsNP + (k * 5 + (b >= 0 ? b : 4)) * 4;
					for (int i = 0; i < 4; ++i) g[i] += (ya[i] * yb[i]) * np[i];
				}
				tl += log(dotPi(m, g) / (double) KK);
			}
			tl = groupSum<WPI>(tl, red, flip, group, gwarp);
		}
		if (gtid == 0) {
			hu_place_row row = e;
			// literal return value: treeLoglik over a root message of INVALID_LOGLIK = 1, tabulated per W
			row.loglik = t.litTab[W];
			row.loglikTrue = a.useTrue ? tl : 0.0; // not evaluated unless hu_sep_params.useTrueLoglik is set
			row.wnr = ln;
			row.ratio = lu / w0;
			row.height = t.height[c] + lu;
			row.aNode = (row.ratio <= 0.5 || t.height[pN] > a.maxHeight) ? c : pN;
			row.taxon = t.taxon[row.aNode];
			row.outerIters = iter;
			row.nPlaces = emIters; // diagnostic until finalizeKernel overwrites it with the placement count
			a.places[(size_t) r * HU_MAX_SEED + idx] = row;
		}
	}
}

// calcQValues (UNIFORM prior) + sort by qPlace + bestPlace; one thread per read.
__global__ void finalizeKernel(int nReads, const int* __restrict__ winStart, const int* __restrict__ nKept,
		hu_place_row* __restrict__ places, int useTrue, int priorHeight, int onlyML, hu_place_row* __restrict__ best) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nReads) return;
	const int n = winStart[r] < 0 ? 0 : nKept[r];
	hu_place_row* pl = places + (size_t) r * HU_MAX_SEED;
	hu_place_row out;
	if (n == 0) {
		out.start = out.end = 0; out.cNode = out.pNode = out.aNode = out.taxon = -1;
		out.dist = out.wuv = out.ratio = out.wnr = out.loglik = out.loglikTrue = out.estLoglik = 0;
		out.height = out.annoDist = out.qPlace = out.qTaxon = 0;
		out.outerIters = 0; out.nPlaces = 0;
		best[r] = out;
		return;
	}
	const double MAX_Q = 250.0, LOG10 = log(10.0);
	int tax[HU_MAX_SEED];
	double taxPP[HU_MAX_SEED];
	int nTax = 0;
	double norm = -INFINITY, mx = -INFINITY;
	for (int i = 0; i < n; ++i) {
		if (useTrue) pl[i].loglik = pl[i].loglikTrue;
		pl[i].annoDist = (pl[i].aNode == pl[i].cNode ? pl[i].wuv * pl[i].ratio : (1 - pl[i].ratio) * pl[i].wuv) + pl[i].wnr;
		pl[i].nPlaces = n;
		pl[i].qPlace = pl[i].qTaxon = 0;
	}
	double key[HU_MAX_SEED];
	int perm[HU_MAX_SEED];
	if (onlyML) { // --ML (src/hmmufotu.cpp:725-727): std::sort(rbegin, rend, compareByLoglik), no Q values
		for (int i = 0; i < n; ++i) key[i] = pl[i].loglik;
		stdSortDescendingPerm(key, n, perm); // libstdc++'s introsort replayed: ties and NaN keys land where the reference leaves them
		best[r] = pl[perm[0]];
		return;
	}
	for (int i = 0; i < n; ++i) {
		// PTPlacement::logPriorPr (src/PhyloTreeUnrooted.cpp:1166-1177): -0 (uniform) or -(annoDist - wnr + height)
		const double prior = priorHeight ? -(pl[i].annoDist - pl[i].wnr + pl[i].height) : -0.0;
		const double p = pl[i].loglik + prior;
		pl[i].qPlace = p; // parked: the posterior log-probability, replaced by the Q value below
		mx = fmax(mx, p);
		int tIdx = -1;
		for (int k = 0; k < nTax; ++k) if (tax[k] == pl[i].taxon) { tIdx = k; break; }
		if (tIdx < 0) { tax[nTax] = pl[i].taxon; taxPP[nTax] = p; nTax++; }
		else { // add_scaled, src/math/Stats.h:233-236
			const double sc = fmax(taxPP[tIdx], p);
			taxPP[tIdx] = log(exp(taxPP[tIdx] - sc) + exp(p - sc)) + sc;
		}
		const double sc = fmax(norm, p);
		norm = log(exp(norm - sc) + exp(p - sc)) + sc;
	}
	double sum = 0;
	for (int i = 0; i < n; ++i) sum += exp(pl[i].qPlace - mx);
	for (int i = 0; i < n; ++i) {
		const double pr = exp(pl[i].qPlace - mx) / sum;
		double q = -10.0 * log(1 - pr) / LOG10; // p2q
		if (q > MAX_Q) q = MAX_Q;
		// A placement whose ratio is 0/0 makes the height prior, hence every posterior, NaN.  On x86 that NaN reaches the
		// output with its sign bit cleared (the prior's unary minus flips the default NaN); the device's canonical NaN
		// has it set, and the TSV shows the sign.
		if (q != q) q = __longlong_as_double(0x7FF8000000000000LL);
		pl[i].qPlace = q;
		key[i] = q;
		double pt = 0;
		for (int k = 0; k < nTax; ++k) if (tax[k] == pl[i].taxon) pt = taxPP[k];
		double qt = -10.0 * log(1 - exp(pt - norm)) / LOG10;
		if (qt > MAX_Q) qt = MAX_Q;
		if (qt != qt) qt = __longlong_as_double(0x7FF8000000000000LL);
		pl[i].qTaxon = qt;
	}
	// std::sort(places.rbegin(), places.rend(), compareByQPlace); bestPlace = places[0] (src/hmmufotu.cpp:729-733).  Under
	// SURVEY.md F6 all keys are equal and the winner is whatever introsort's data movement leaves first: replayed literally
	// (hu_stdsort.cuh), which also covers partial ties and NaN keys (they compare like equal keys).
	stdSortDescendingPerm(key, n, perm);
	best[r] = pl[perm[0]];
}

int configPlaceKernels() {
	const size_t smem = (size_t) HU_PLACE_WARPS * (placeWarpDoubles(HU_MAX_DG) + HU_PLACE_CAPW) * sizeof(double);
	HU_CUDA(cudaFuncSetAttribute(placeKernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	HU_CUDA(cudaFuncSetAttribute(placeKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	HU_CUDA(cudaFuncSetAttribute(placeKernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	return HU_OK;
}

int launchPlace(hu_batch* b, const hu_sep_params* p) {
	hu_db* db = b->db;
	PlaceArgs a;
	a.nReads = b->nReads; a.maxNSeed = p->maxNSeed; a.capW = HU_PLACE_CAPW; a.useTrue = p->useTrueLoglik;
	a.maxHeight = p->maxHeight;
	a.ahead = 2; a.pad = 0;
	{ const char* e = getenv("HU_PLACE_AHEAD"); if (e) a.ahead = atoi(e); } // measurements
	a.aligned = b->vit.aligned; a.winStart = b->winStart; a.winEnd = b->winEnd;
	a.est = b->est; a.order = b->order; a.nKept = b->nKept; a.places = b->places;
	a.cursor = b->workCount + 1;
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	int ctasPerSm = 4;
	{ const char* e = getenv("HU_PLACE_CTAS"); if (e && atoi(e) >= 1 && atoi(e) <= 4) ctasPerSm = atoi(e); } // measurements
	const int grid = sms * ctasPerSm;
	const int KKh = db->model.dgK > 0 ? db->model.dgK : 1;
	// long reads (windows of thousands of columns): four warps per item; amplicon reads and merged pairs: one warp per item
	const bool wideItems = b->maxSeqLen > 700;
	int wpi = wideItems ? 4 : 1;
	{ const char* e = getenv("HU_PLACE_WPI"); if (e && (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4)) wpi = atoi(e); } // measurements
	const int groups = HU_PLACE_WARPS / wpi;
	const size_t smem = (size_t) groups * (placeWarpDoubles(KKh) + HU_PLACE_CAPW) * sizeof(double);
	// global slab for windows wider than the shared-memory capacity, one row per group of warps
	const size_t slabBytes = (size_t) grid * HU_PLACE_WARPS * db->tree.csLen * sizeof(double);
	if (!b->placeSlab) {
		void* q = NULL;
		HU_CUDA(cudaMalloc(&q, slabBytes));
		b->allocs.push_back(q);
		b->placeSlab = (double*) q;
	}
	a.slab = b->placeSlab;
	HU_CUDA(cudaMemsetAsync(a.cursor, 0, sizeof(int), b->stream)); // the stage can run on its own (per-call facades), not only after hu_estimate_batch
	if (wpi == 4) placeKernel<4><<<grid, HU_PLACE_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	else if (wpi == 2) placeKernel<2><<<grid, HU_PLACE_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	else placeKernel<1><<<grid, HU_PLACE_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	HU_CUDA(cudaEventRecord(b->ev[6], b->stream));
	b->launches += 1;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

int launchFinalize(hu_batch* b, const hu_sep_params* p) {
	finalizeKernel<<<(b->nReads + 63) / 64, 64, 0, b->stream>>>(b->nReads, b->winStart, b->nKept, b->places, p->useTrueLoglik,
			p->prior == HU_PRIOR_HEIGHT, p->onlyML, b->best);
	HU_CUDA(cudaEventRecord(b->ev[7], b->stream));
	b->launches += 1;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
