// K3: the "Estimate" stage of SEP -- PTUnrooted::estimateSeq for every (read, seed)
// (src/PhyloTreeUnrooted.cpp:849-877) followed by filterPlacements (src/HmmUFOtu_main.cpp:162-173).
//
// HBM-bound: an item streams the two message windows of its seed edge (linear form, 32 B per column each,
// hu_internal.cuh) exactly once.  estimateStagedKernel is a persistent producer/consumer pipeline, three CTAs per
// SM: one producer lane walks the item list and issues cp.async.bulk copies (hu_pipe.cuh) of 128-column chunks
// of U and V into a 3-stage shared-memory ring, plus the read's window and the item record of the NEXT item; four
// consumer warps take one column per thread and chunk.  Per column
//   g_j = (P(wur) E^U_j) * (P(wvr) E^V_j)              (= e^{R_j} of the reference up to the column scale)
// pass 1 counts the columns whose arg-max state differs from the read's (wnr, cpp:1018-1032) and parks g_j in
// shared memory; pass 2 sums log(pi . (g_j * P(wnr) n_j)) from there (cpp:872-874).  The per-column scales enter
// only through their sum over the window, taken from 32-column prefix sums (TreeDev::upP32) plus the two ragged
// ends.  36 fp64 operations and 4 shared-memory reads per column in pass 1, no exp.
//
// Round-2 shape (profiles/r1d_summary.md found the round-1 kernel bound by shared-memory wavefronts: every thread re-read
// the item's two 4x4 transition matrices, 256 B, for every column next to the 64 B of message data): the matrices live in
// REGISTERS for the whole item (64 registers; hence fewer threads per CTA and more CTAs per SM), g_j is kept as two 16-byte
// halves (128-bit stores / loads), pi is folded into P(wnr) n.  Windows wider than the shared-memory g buffer keep g in a
// per-CTA slab in global memory (L2-resident) and read the read's window from global memory: any width is streamed from
// HBM once, and there is no separate wide-window kernel any more.
#include "hu_batch.cuh"
#include <stdlib.h>
#include "hu_pipe.cuh"
#include "hu_stdsort.cuh"

namespace hu {

#define HU_EST_CONS 128                   // consumer threads (4 warps), one column of a chunk each
#define HU_EST_THREADS (HU_EST_CONS + 32) // + the producer warp
#define HU_EST_CH 128                     // columns per ring stage
#define HU_EST_STAGES 3
#define HU_EST_CAPW 1408                  // g buffer (columns) with three CTAs per SM: single-end amplicon reads
#define HU_EST_CAPW_BIG 2560              // ... with two CTAs per SM: merged paired-end windows

struct EstArgs {
	int nReads, maxNSeed, weighted, capW; // weighted: estimateBranchLengthWeighted instead of ...Unweighted; capW: widest window whose g fits shared memory
	int csLen, ahead; // ahead: L2 prefetch distance of the producer, in ring chunks (0: none)
	const int8_t* aligned;
	const int* winStart;
	const int* winEnd;
	const int* seedIds;
	const int* nSeeds;
	hu_place_row* est;
	EstItem* items;   // [nReads * maxNSeed]
	EstItemP* itemsP; // [nReads * maxNSeed]: distances, ratio and the two transition matrices of an item
	double2* slab;    // [grid][2][csLen]: g of windows wider than capW
};

// DNASubModel::Pr(v) by a half warp: `l` = 0..15 is this lane's matrix entry (row l / 4, column l % 4).  GTR
// spreads the four exponentials and the sixteen 4-term products over the lanes (same operand order as modelPr,
// so the same bits); the closed-form models are evaluated by the half warp's first lane.
__device__ __forceinline__ void modelPrHalfWarp(const Model& m, double v, double* P, int l, unsigned int mask) {
	if (m.type == HU_GTR) {
		const double ek = exp(m.lambda[l & 3] * v);
		const int i = l >> 2, j = l & 3, base = (threadIdx.x & 31) & ~3;
		double sacc = 0;
		for (int k = 0; k < 4; ++k) {
			const double e = __shfl_sync(mask, ek, base + k);
			sacc += m.U[i * 4 + k] * e * m.Ui[k * 4 + j];
		}
		P[l] = v == 0 ? ((i == j) ? 1.0 : 0.0) : sacc;
	} else if (l == 0)
		modelPr(m, v, P);
}

// ---- item list: one record per (read, seed), so that the streaming kernel has no dependent loads and no
// per-item serial prologue.  One warp per item: the p-distances of the read to the child and parent sequences
// (SeqUtils::pDist over the window), the attachment ratio and the two transition matrices P(wur), P(wvr). ----
__device__ __forceinline__ unsigned int pdistWord(uint32_t xw, uint32_t sw, uint32_t keep) {
	// four bases per word: both valid (code >= 0, i.e. sign bit clear) -> N += 1 (high half), differ -> d += 1 (low half)
	unsigned int acc = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t xb = (xw >> (8 * k)) & 0xffu, sb = (sw >> (8 * k)) & 0xffu;
		const bool ok = ((keep >> (8 * k)) & 1u) && !((xb | sb) & 0x80u);
		if (ok) acc += 0x10000u + (xb != sb);
	}
	return acc;
}

__global__ void __launch_bounds__(256, 4) estPlanKernel(EstArgs a, TreeDev t, Model m) {
	const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (i >= a.nReads * a.maxNSeed) return;
	const int r = i / a.maxNSeed, s = i % a.maxNSeed;
	EstItem it;
	it.rs = r * HU_MAX_SEED + s; it.u = 0; it.v = 0; it.start = 0; it.W = 0; it.kind = 0; it.w0 = 0;
	const int ws = a.winStart[r];
	if (ws >= 0 && s < a.nSeeds[r]) {
		const int u = a.seedIds[(size_t) r * HU_MAX_SEED + s];
		it.u = u; it.v = t.parent[u]; it.start = ws; it.W = a.winEnd[r] - ws + 1; it.w0 = t.blen[u];
		it.kind = 1;
		// the three byte windows are read as aligned 32-bit words (4 columns per lane and load); the rows start at multiples
		// of csLen, so the three rows share their alignment remainder only when csLen % 4 == 0 -- each gets its own shift
		const int8_t* x = a.aligned + (size_t) r * t.csLen + ws;
		const int8_t* su = t.seq + (size_t) it.u * t.csLen + ws;
		const int8_t* sv = t.seq + (size_t) it.v * t.csLen + ws;
		unsigned int cu = 0, cv = 0; // (N << 16) | d
		if ((((uintptr_t) x ^ (uintptr_t) su) & 3) == 0 && (((uintptr_t) x ^ (uintptr_t) sv) & 3) == 0) {
			const int off = (int) ((uintptr_t) x & 3);
			const uint32_t* xw = reinterpret_cast<const uint32_t*>(x - off);
			const uint32_t* uw = reinterpret_cast<const uint32_t*>(su - off);
			const uint32_t* vw = reinterpret_cast<const uint32_t*>(sv - off);
			const int nWords = (off + it.W + 3) >> 2;
			for (int w = lane; w < nWords; w += 32) {
				uint32_t keep = 0x01010101u; // byte k of the word is column 4 w + k - off: inside [0, W)?
				const int c0 = 4 * w - off;
				if (c0 < 0 || c0 + 3 >= it.W) {
					keep = 0;
					for (int k = 0; k < 4; ++k) if (c0 + k >= 0 && c0 + k < it.W) keep |= 1u << (8 * k);
				}
				const uint32_t xv = __ldg(xw + w);
				cu += pdistWord(xv, __ldg(uw + w), keep);
				cv += pdistWord(xv, __ldg(vw + w), keep);
			}
		} else
			for (int c = lane; c < it.W; c += 32) {
				const int b = x[c];
				if (b >= 0) {
					const int bu = su[c], bv = sv[c];
					if (bu >= 0) cu += 0x10000u + (bu != b);
					if (bv >= 0) cv += 0x10000u + (bv != b);
				}
			}
		for (int o = 16; o > 0; o >>= 1) {
			cu += __shfl_xor_sync(0xffffffffu, cu, o);
			cv += __shfl_xor_sync(0xffffffffu, cv, o);
		}
		const double cDist = (double) (cu & 0xffffu) / (double) (cu >> 16);
		const double pDist = (double) (cv & 0xffffu) / (double) (cv >> 16);
		double ratio = cDist / (cDist + pDist);
		if (isnan(ratio)) ratio = 0.5;
		const double wur = it.w0 * ratio, wvr = it.w0 - wur;
		EstItemP& ip = a.itemsP[i];
		modelPrHalfWarp(m, lane < 16 ? wur : wvr, ip.P[lane >> 4], lane & 15, 0xffffffffu);
		// everything else the streaming kernel would have to fetch from global memory
		double sc = windowScale(t.upS, t.upP32, it.u, t.csLen, t.nWords, ws, ws + it.W - 1, lane)
		          + windowScale(t.dnS, t.dnP32, it.u, t.csLen, t.nWords, ws, ws + it.W - 1, lane);
		for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
		if (lane == 0) {
			ip.cDist = cDist; ip.ratio = ratio; ip.sumScale = sc;
			ip.taxU = t.taxon[it.u]; ip.taxV = t.taxon[it.v];
		}
	}
	if (lane == 0) a.items[i] = it;
}

struct EstSmem {
	double ring[HU_EST_STAGES][2][HU_EST_CH * 4]; // U chunk, V chunk
	EstItemP rec[2];                              // cDist, ratio, P(wur), P(wvr) of the staged items (bulk-copied)
	EstItem meta[2];
	int hoff[4];                                  // [0..1]: offset of the window inside hdr (alignment remainder); rest: padding
	double P2[16];                                // P(wnr)
	double NpPi[5][4];                            // pi * (P(wnr) applied to the read's leaf vector), per base / gap
	unsigned int redB[HU_EST_CONS / 32];
	double redC[HU_EST_CONS / 32];
	double redW[2][HU_EST_CONS / 32];
	unsigned long long full[HU_EST_STAGES], empty[HU_EST_STAGES], hfull[2], hempty[2];
	// followed in dynamic shared memory by g01[capW], g23[capW] (g_j of the current item as two 16-byte halves) and
	// hdr[2][capW + 32] (the read's window, double buffered, 16-byte alignment slack on both sides)
};
static_assert(sizeof(EstItemP) % 16 == 0, "item records are bulk-copied");
static_assert(sizeof(EstSmem) % 16 == 0, "the arrays behind EstSmem are targets of 16-byte bulk copies");
__host__ __device__ inline size_t estSmemBytes(int capW) { return sizeof(EstSmem) + (size_t) 2 * capW * sizeof(double2) + (size_t) 2 * (capW + 32); }

// one message column out of a ring stage: two 128-bit reads; every other group of four threads reads the halves
// in the opposite order so that a quarter-warp covers all 32 banks.  Those threads get (v2, v3, v0, v1) and
// multiply by the matrix whose column pairs are exchanged the same way (done once per item, when the matrices are
// loaded into registers): the same products, summed as (p2 + p3) + (p0 + p1) instead of (p0 + p1) + (p2 + p3) --
// identical bits, no selects.
__device__ __forceinline__ void ldsColumn(const double* __restrict__ chunk, int tid, double* v) {
	const int sw = (tid >> 2) & 1;
	const double2 a = *reinterpret_cast<const double2*>(chunk + 4 * tid + 2 * sw);
	const double2 b = *reinterpret_cast<const double2*>(chunk + 4 * tid + 2 * (1 - sw));
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// The consumer side of one item.  G / X: where g_j and the read's window live -- shared memory (WIDE = false: typed shared
// pointers, so that the accesses are LDS / STS) or, for a window wider than the shared-memory buffer, the CTA's slab in global
// memory and the read's row in global memory.
template <bool WEIGHTED, bool WIDE>
__device__ __forceinline__ void estConsumeItem(EstSmem& sm, const EstArgs& a, const Model& m, const EstItem& it, int h, double2* __restrict__ g01,
		double2* __restrict__ g23, const int8_t* __restrict__ x, int& st, int& ph, int tid, int warp, int lane, int piMax, double wGap) {
	const int W = it.W;
	const int sw = (tid >> 2) & 1;
	// the item's two transition matrices, in registers for the whole item (column pairs exchanged for the threads that read
	// their message columns half-swapped)
	double P0[16], P1[16];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			P0[i * 4 + j] = sm.rec[h].P[0][i * 4 + (j ^ (2 * sw))];
			P1[i * 4 + j] = sm.rec[h].P[1][i * 4 + (j ^ (2 * sw))];
		}
	double ll = 0;
	// ---- pass 1: the unweighted branch-length estimate; g_j parked ----
	unsigned int diff = 0;
	double dW = 0, nW = 0; // weighted variant: sums of w1 w2 over mismatching / all columns (cpp:1034-1052)
	for (int c0 = 0; c0 < W; c0 += HU_EST_CH) {
		mbarWait(&sm.full[st], ph);
		const int c = c0 + tid;
		if (c < W) {
			double eu[4], ev[4], ya[4], yb[4], g[4];
			ldsColumn(sm.ring[st][0], tid, eu);
			ldsColumn(sm.ring[st][1], tid, ev);
			// matVec4, not a fused chain: the arg-max below is a discrete decision, and on exactly tied columns (uniform pi,
			// uninformative messages: K80, JC69) the rounding of these sums decides it -- same products, same order of
			// additions as dot_product_scaled in the oracle (a fused chain changed 4 of 22 kept placements of a K80 read)
			matVec4(P0, eu, ya); // rows are not permuted: ya, yb and g are in state order for every thread
			matVec4(P1, ev, yb);
#pragma unroll
			for (int k = 0; k < 4; ++k) g[k] = ya[k] * yb[k];
			g01[c] = make_double2(g[0], g[1]);
			g23[c] = make_double2(g[2], g[3]);
			const int b = x[c];
			const int st1 = argmax4(g);
			const bool mis = st1 != (b >= 0 ? b : piMax);
			diff += mis;
			if (WEIGHTED) { // inferWeight at the arg-max: max / sum for the tree side, 1 (base) or pi_max / sum(pi) (gap) for the read
				const double ww = fmax(fmax(g[0], g[1]), fmax(g[2], g[3])) / ((g[0] + g[1]) + (g[2] + g[3])) * (b >= 0 ? 1.0 : wGap);
				if (mis) dW += ww;
				nW += ww;
			}
		}
		__syncwarp();
		if (lane == 0) mbarArrive(&sm.empty[st]);
		if (++st == HU_EST_STAGES) { st = 0; ph ^= 1; }
	}
	for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
	if (lane == 0) sm.redB[warp] = diff;
	consumerSync(HU_EST_CONS);
	diff = 0;
#pragma unroll
	for (int k = 0; k < HU_EST_CONS / 32; ++k) diff += sm.redB[k];
	double wnr = (double) diff / (double) W;
	if (WEIGHTED) { // block sums of the two weights (uniform branch)
		for (int o = 16; o > 0; o >>= 1) { dW += __shfl_xor_sync(0xffffffffu, dW, o); nW += __shfl_xor_sync(0xffffffffu, nW, o); }
		if (lane == 0) { sm.redW[0][warp] = dW; sm.redW[1][warp] = nW; }
		consumerSync(HU_EST_CONS);
		dW = nW = 0;
		for (int k = 0; k < HU_EST_CONS / 32; ++k) { dW += sm.redW[0][k]; nW += sm.redW[1][k]; }
		wnr = dW / nW;
		consumerSync(HU_EST_CONS); // redW is reused by the next item
	}
	// ---- pi * P(wnr) applied to the read's leaf vector: depends on the base only ----
	if (warp == 0) {
		if (lane < 16) modelPrHalfWarp(m, wnr, sm.P2, lane, 0x0000ffffu);
		__syncwarp();
		if (lane < 5) {
			double n[4], y[4];
			leafLin(m, lane < 4 ? lane : -1, n);
			matVec4(sm.P2, n, y);
			for (int k = 0; k < 4; ++k) sm.NpPi[lane][k] = m.pi[k] * y[k];
		}
	}
	consumerSync(HU_EST_CONS);
	// ---- pass 2: treeLoglik(pi, UPr + VPr + NPr) ----
	// sum of logs as the log of a running product: one log per thread instead of one per column (the product of
	// the <= 12 factors of a thread is flushed long before it could underflow)
	double prod = 1.0;
	for (int c = tid; c < W; c += HU_EST_CONS) {
		const int b = x[c];
		const double* np = sm.NpPi[b >= 0 ? b : 4];
		const double2 a01 = g01[c], a23 = g23[c];
		prod *= fma(a23.y, np[3], fma(a23.x, np[2], fma(a01.y, np[1], a01.x * np[0])));
		if (prod < 1e-150) { ll += log(prod); prod = 1.0; }
	}
	ll += log(prod);
	for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
	if (lane == 0) sm.redC[warp] = ll;
	consumerSync(HU_EST_CONS);
	if (tid == 0) {
		const double cDist = sm.rec[h].cDist, ratio = sm.rec[h].ratio;
		const int taxon = ratio <= 0.5 ? sm.rec[h].taxU : sm.rec[h].taxV;
		ll = sm.rec[h].sumScale;
		mbarArrive(&sm.hempty[h]); // every consumer is past its last read of this item's staged window and record
		for (int k = 0; k < HU_EST_CONS / 32; ++k) ll += sm.redC[k];
		const int r = it.rs / HU_MAX_SEED;
		hu_place_row row;
		row.start = it.start; row.end = it.start + W - 1;
		row.cNode = it.u; row.pNode = it.v; row.aNode = ratio <= 0.5 ? it.u : it.v;
		row.taxon = taxon;
		row.dist = cDist;
		row.wuv = it.w0; row.ratio = ratio; row.wnr = wnr;
		row.loglik = ll; row.loglikTrue = 0; row.estLoglik = ll;
		row.height = 0; row.annoDist = 0; row.qPlace = 0; row.qTaxon = 0;
		row.outerIters = 0; row.nPlaces = 0;
		a.est[(size_t) r * HU_MAX_SEED + (it.rs % HU_MAX_SEED)] = row;
	}
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(HU_EST_THREADS, 3) estimateStagedKernel(EstArgs a, TreeDev t, Model m) {
	extern __shared__ __align__(128) unsigned char estRaw[];
	EstSmem& sm = *reinterpret_cast<EstSmem*>(estRaw);
	double2* const gS = reinterpret_cast<double2*>(estRaw + sizeof(EstSmem));
	int8_t* const hdrS = reinterpret_cast<int8_t*>(gS + (size_t) 2 * a.capW);
	const int capW = a.capW, hdrPitch = a.capW + 32;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int total = a.nReads * a.maxNSeed;
	const int nBatches = (total + 31) / 32;
	if (tid == 0) {
		for (int i = 0; i < HU_EST_STAGES; ++i) { mbarInit(&sm.full[i], 1); mbarInit(&sm.empty[i], HU_EST_CONS / 32); }
		for (int i = 0; i < 2; ++i) { mbarInit(&sm.hfull[i], 1); mbarInit(&sm.hempty[i], 1); }
		mbarInitFence();
	}
	__syncthreads();
	if (warp == HU_EST_CONS / 32) {
		// ================= producer warp: lane l holds item 32 * batch + l; lane 0 issues =================
		int st = 0, ph = 0, h = 0, hph = 0;
		EstItem cur;
		cur.kind = 0;
		int batch = blockIdx.x;
		if (batch < nBatches && batch * 32 + lane < total) cur = a.items[batch * 32 + lane];
		for (; batch < nBatches; batch += gridDim.x) {
			EstItem nxt; // the records of the next batch are in flight while this one is issued
			nxt.kind = 0;
			const int nb = batch + gridDim.x;
			if (nb < nBatches && nb * 32 + lane < total) nxt = a.items[nb * 32 + lane];
			unsigned int todo = __ballot_sync(0xffffffffu, cur.kind == 1);
			while (todo) {
				const int l = __ffs(todo) - 1;
				todo &= todo - 1;
				EstItem it;
				it.rs = __shfl_sync(0xffffffffu, cur.rs, l); it.u = __shfl_sync(0xffffffffu, cur.u, l);
				it.v = __shfl_sync(0xffffffffu, cur.v, l); it.start = __shfl_sync(0xffffffffu, cur.start, l);
				it.W = __shfl_sync(0xffffffffu, cur.W, l); it.kind = 1;
				it.w0 = __shfl_sync(0xffffffffu, cur.w0, l);
				if (lane == 0) {
					const int r = it.rs / HU_MAX_SEED;
					mbarWaitRelaxed(&sm.hempty[h], hph ^ 1);
					sm.meta[h] = it;
					uint32_t bytes = 0;
					const int8_t* src = a.aligned + (size_t) r * t.csLen + it.start;
					const int off = (int) ((uintptr_t) src & 15);
					if (it.W <= capW) bytes = (uint32_t) ((off + it.W + 15) & ~15); // wider windows: the consumers read the read's codes from global memory
					sm.hoff[h] = off;
					mbarArriveExpectTx(&sm.hfull[h], bytes + (uint32_t) sizeof(EstItemP));
					if (bytes) bulkLoad(hdrS + (size_t) h * hdrPitch, src - off, bytes, &sm.hfull[h]);
					bulkLoad(&sm.rec[h], a.itemsP + ((size_t) r * a.maxNSeed + it.rs % HU_MAX_SEED), (uint32_t) sizeof(EstItemP), &sm.hfull[h]);
					const double* U = t.up + ((size_t) it.u * t.csLen + it.start) * 4;
					const double* V = t.dn + ((size_t) it.u * t.csLen + it.start) * 4;
					for (int c0 = 0; c0 < it.W; c0 += HU_EST_CH) {
						const uint32_t nb32 = (uint32_t) min(HU_EST_CH, it.W - c0) * 32u;
						// the chunk that will take this stage next time round is asked into the L2 now: the ring alone keeps too few
						// bytes in flight per SM to cover the DRAM latency (ncu r2c: DRAM 61 % busy, consumers waiting)
						const int ca = c0 + a.ahead * HU_EST_CH;
						if (a.ahead > 0 && ca < it.W) {
							const uint32_t na = (uint32_t) min(HU_EST_CH, it.W - ca) * 32u;
							bulkPrefetchL2(U + (size_t) ca * 4, na);
							bulkPrefetchL2(V + (size_t) ca * 4, na);
						}
						mbarWaitRelaxed(&sm.empty[st], ph ^ 1);
						mbarArriveExpectTx(&sm.full[st], 2 * nb32);
						bulkLoad(sm.ring[st][0], U + (size_t) c0 * 4, nb32, &sm.full[st]);
						bulkLoad(sm.ring[st][1], V + (size_t) c0 * 4, nb32, &sm.full[st]);
						if (++st == HU_EST_STAGES) { st = 0; ph ^= 1; }
					}
				}
				h ^= 1;
				if (h == 0) hph ^= 1;
			}
			cur = nxt;
		}
		if (lane == 0) { // end of the item stream
			mbarWaitRelaxed(&sm.hempty[h], hph ^ 1);
			sm.meta[h].kind = 0;
			mbarArrive(&sm.hfull[h]);
		}
		return;
	}
	// ================= consumer warps =================
	int piMax = 0; // argmax of log(pi): state of a gap column of the read
	for (int k = 1; k < 4; ++k) if (m.logPi[k] > m.logPi[piMax]) piMax = k;
	const double wGap = m.pi[piMax] / ((m.pi[0] + m.pi[1]) + (m.pi[2] + m.pi[3])); // inferWeight(log pi) at its arg-max
	double2* const slab = a.slab + (size_t) blockIdx.x * 2 * a.csLen;
	int st = 0, ph = 0, h = 0, hph = 0;
	for (;;) {
		mbarWait(&sm.hfull[h], hph);
		const EstItem it = sm.meta[h];
		if (it.kind == 0) break;
		if (it.W <= capW)
			estConsumeItem<WEIGHTED, false>(sm, a, m, it, h, gS, gS + capW, hdrS + (size_t) h * hdrPitch + sm.hoff[h], st, ph, tid, warp, lane, piMax, wGap);
		else
			estConsumeItem<WEIGHTED, true>(sm, a, m, it, h, slab, slab + a.csLen, a.aligned + (size_t) (it.rs / HU_MAX_SEED) * t.csLen + it.start, st, ph, tid, warp, lane,
			                               piMax, wGap);
		h ^= 1;
		if (h == 0) hph ^= 1;
	}
}

// filterPlacements: std::sort(places.rbegin(), places.rend(), compareByLoglik) and keep the placements within maxError
// of the best (src/HmmUFOtu_main.cpp:162-173).  The sort is libstdc++'s introsort, replayed move for move
// (hu_stdsort.cuh): exactly equal estimated log-likelihoods come out in the order the reference leaves them in.
__global__ void filterKernel(int nReads, const int* __restrict__ winStart, const int* __restrict__ nSeeds,
		const hu_place_row* __restrict__ est, double maxError, int* __restrict__ order, int* __restrict__ nKept) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nReads) return;
	const int n = winStart[r] < 0 ? 0 : max(nSeeds[r], 0);
	const hu_place_row* e = est + (size_t) r * HU_MAX_SEED;
	double key[HU_MAX_SEED];
	int o[HU_MAX_SEED];
	for (int i = 0; i < n; ++i) key[i] = e[i].estLoglik;
	stdSortDescendingPerm(key, n, o);
	int keep = 0;
	if (n > 0) {
		const double best = key[o[0]];
		while (keep < n && !(best - key[o[keep]] > maxError)) keep++;
	}
	for (int i = 0; i < n; ++i) order[(size_t) r * HU_MAX_SEED + i] = o[i];
	nKept[r] = keep;
}

int configEstimateKernels() {
	HU_CUDA(cudaFuncSetAttribute(estimateStagedKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) estSmemBytes(HU_EST_CAPW_BIG)));
	HU_CUDA(cudaFuncSetAttribute(estimateStagedKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) estSmemBytes(HU_EST_CAPW_BIG)));
	return HU_OK;
}

// filterPlacements alone (the estimates are already on the device: hu_estimate_batch ran with another maxError, or
// hu_batch_set_estimates supplied them)
int launchFilter(hu_batch* b, const hu_sep_params* p) {
	filterKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, b->winStart, b->seed.nSeeds, b->est, p->maxError, b->order, b->nKept);
	b->launches += 1;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

int launchEstimate(hu_batch* b, const hu_sep_params* p) {
	hu_db* db = b->db;
	EstArgs a;
	a.nReads = b->nReads; a.maxNSeed = p->maxNSeed; a.weighted = p->estMethod == HU_EST_WEIGHTED;
	// merged paired-end windows (1.5 - 2.5 k columns) get the larger g buffer (two CTAs per SM); single-end amplicon reads
	// and reads so long that every window spills to the global slab anyway get the small one (three CTAs per SM)
	a.capW = (b->paired || (b->maxSeqLen > 400 && b->maxSeqLen <= 700)) ? HU_EST_CAPW_BIG : HU_EST_CAPW;
	a.csLen = db->tree.csLen; a.ahead = HU_EST_STAGES;
	{ const char* e = getenv("HU_EST_AHEAD"); if (e) a.ahead = atoi(e); } // measurements
	a.aligned = b->vit.aligned; a.winStart = b->winStart; a.winEnd = b->winEnd;
	a.seedIds = b->seed.seedIds; a.nSeeds = b->seed.nSeeds;
	a.est = b->est; a.items = b->estItems; a.itemsP = b->estItemsP;
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	HU_CUDA(cudaMemsetAsync(b->workCount, 0, 4 * sizeof(int), b->stream));
	const int total = a.nReads * a.maxNSeed;
	const size_t smem = estSmemBytes(a.capW);
	const int perSm = std::max(1, std::min(3, (int) ((227 * 1024) / (smem + 1024))));
	const int grid = sms * perSm;
	if (!b->estSlab || b->estSlabCtas < grid) {
		void* q = NULL;
		HU_CUDA(cudaMalloc(&q, (size_t) grid * 2 * a.csLen * sizeof(double2)));
		b->allocs.push_back(q);
		b->estSlab = q; b->estSlabCtas = grid;
	}
	a.slab = (double2*) b->estSlab;
	estPlanKernel<<<(total + 7) / 8, 256, 0, b->stream>>>(a, db->tree, db->model);
	if (a.weighted) estimateStagedKernel<true><<<grid, HU_EST_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	else estimateStagedKernel<false><<<grid, HU_EST_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	filterKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, b->winStart, a.nSeeds, b->est, p->maxError, b->order, b->nKept);
	HU_CUDA(cudaEventRecord(b->ev[5], b->stream));
	b->launches += 3;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
