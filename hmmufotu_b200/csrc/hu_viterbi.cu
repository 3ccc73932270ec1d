// K1: batched banded plan-7 Viterbi, traceback and aligned-code emission.
//
// One warp per sequence.  The "band" of the reference is a chain of boxes joined at their corners
// (upstream rectangle, seed strip, inter-seed rectangle, seed strip, downstream rectangle;
// src/BandedHMMP7.cpp:782-892).  Each box is swept in 32-column stripes: lane l owns profile column
// j0+32s+l, rows advance as an anti-diagonal wavefront (row i = i0+t-l at step t), the (i,j-1) and
// (i-1,j-1) dependencies arrive by warp shuffle, the (i-1,j) dependency stays in registers, and the
// right-most column of a stripe is handed to the next stripe through shared memory.  All arithmetic is
// fp64 adds and mins in the reference's operand order, so costs and argmins are bit-identical to it.
// Instead of the reference's four dense (N+1)x(K+2) matrices, only a 4-bit back-pointer per cell is
// stored (2 bits M: B/M/I/D, 1 bit I: M/I, 1 bit D: M/D), evaluated with exactly the candidate sets and
// first-strict-minimum tie rule buildViterbiTrace uses (src/BandedHMMP7.cpp:943-1006, .h:722-788).
//
// Back-pointer layout of a box: the four pointer bits of a cell (2 bits for DP_M, 1 for DP_I, 1 for DP_D) are compare
// outcomes, i.e. predicates; each wavefront step turns them into four warp ballots (bit l = lane l = column jb + l of the
// running 32-column stripe) and lane 0 stores the four words with one 128-bit store: [stripe][step t][4], where
// t = (i - i0) + ((j - j0) & 31).  No per-lane bit assembly, no per-lane stores.
//
// Round 2: the inner loop of a plain rectangle (sweepBoxLean) carries only what a cell needs -- the round-1 loop spent 170
// instructions per wavefront step, most of them index arithmetic and predicates (profiles/r1d_summary.md); the general loop
// (seed strips, boxes that contain read position 1 or profile column 1, LOCAL / CGNL modes) is kept as sweepBox.
#include "hu_batch.cuh"

namespace hu {

struct Cell { double M, I, D; };

__device__ __forceinline__ double col0M(int i, double cNN, double cNB) {
	// prepareViterbiScores, src/BandedHMMP7.cpp:735-746: DP_M(i,0) = (i==1 ? 0 : cNN*(i-1)) + cNB, DP_M(0,0) = inf
	if (i < 1) return INFINITY;
	return (i == 1 ? 0.0 : cNN * (double) (i - 1)) + cNB;
}

struct Best { double v; int j, i; };
__device__ __forceinline__ bool better(double v, int j, int i, const Best& b) {
	// S.minCoeff(&row,&col): first strict minimum in column-major visiting order
	return v < b.v || (v == b.v && (j < b.j || (j == b.j && i < b.i)));
}

// per-column profile parameters held in registers by the lane that owns the column
struct ColParams {
	double tMMp, tIMp, tDMp, tMDp, tDDp; // Tmat_cost[j-1](M,M) (I,M) (D,M) (M,D) (D,D)
	double tMI, tII, tIMj;               // Tmat_cost[j](M,I) (I,I) (I,M)
	double ent, ext;
	double eM0, eM1, eM2, eM3, eI0, eI1, eI2, eI3;
	bool dOn;                            // 1 < j < K
};
__device__ __forceinline__ void loadCol(const HmmDev& h, int jj, ColParams& p) {
	const int Kp = h.Kp;
	p.tMMp = __ldg(h.tr + T_MM * Kp + jj - 1); p.tIMp = __ldg(h.tr + T_IM * Kp + jj - 1); p.tDMp = __ldg(h.tr + T_DM * Kp + jj - 1);
	p.tMDp = __ldg(h.tr + T_MD * Kp + jj - 1); p.tDDp = __ldg(h.tr + T_DD * Kp + jj - 1);
	p.tMI = __ldg(h.tr + T_MI * Kp + jj); p.tII = __ldg(h.tr + T_II * Kp + jj); p.tIMj = __ldg(h.tr + T_IM * Kp + jj);
	p.ent = __ldg(h.ent + jj); p.ext = __ldg(h.ext + jj);
	p.eM0 = __ldg(h.eM + jj); p.eM1 = __ldg(h.eM + Kp + jj); p.eM2 = __ldg(h.eM + 2 * Kp + jj); p.eM3 = __ldg(h.eM + 3 * Kp + jj);
	p.eI0 = __ldg(h.eI + jj); p.eI1 = __ldg(h.eI + Kp + jj); p.eI2 = __ldg(h.eI + 2 * Kp + jj); p.eI3 = __ldg(h.eI + 3 * Kp + jj);
	p.dOn = jj > 1 && jj < h.K;
}
// the lean loop's subset: DP_D exists for 1 < j < K only (src/BandedHMMP7.cpp:823-825); +inf transition costs make the
// recurrence produce the +inf (and the "M" back-pointer bit) that the general code selects explicitly
struct ColLean { double tMMp, tIMp, tDMp, tMDp, tDDp, tMI, tII; };
__device__ __forceinline__ void loadColLean(const HmmDev& h, int jj, ColLean& p) {
	const int Kp = h.Kp;
	const bool dOn = jj > 1 && jj < h.K;
	p.tMMp = __ldg(h.tr + T_MM * Kp + jj - 1); p.tIMp = __ldg(h.tr + T_IM * Kp + jj - 1); p.tDMp = __ldg(h.tr + T_DM * Kp + jj - 1);
	p.tMDp = dOn ? __ldg(h.tr + T_MD * Kp + jj - 1) : INFINITY; p.tDDp = dOn ? __ldg(h.tr + T_DD * Kp + jj - 1) : INFINITY;
	p.tMI = __ldg(h.tr + T_MI * Kp + jj); p.tII = __ldg(h.tr + T_II * Kp + jj);
}
__device__ __forceinline__ void storeEmisCol(const HmmDev& h, int jj, double* emis, int lane) {
	const int Kp = h.Kp;
	emis[0 * 32 + lane] = __ldg(h.eM + jj); emis[1 * 32 + lane] = __ldg(h.eM + Kp + jj); emis[2 * 32 + lane] = __ldg(h.eM + 2 * Kp + jj); emis[3 * 32 + lane] = __ldg(h.eM + 3 * Kp + jj);
	emis[4 * 32 + lane] = INFINITY;
	emis[5 * 32 + lane] = __ldg(h.eI + jj); emis[6 * 32 + lane] = __ldg(h.eI + Kp + jj); emis[7 * 32 + lane] = __ldg(h.eI + 2 * Kp + jj); emis[8 * 32 + lane] = __ldg(h.eI + 3 * Kp + jj);
	emis[9 * 32 + lane] = INFINITY;
}

struct SeqConst { // per-sequence constants of the special states (setSequenceMode, src/BandedHMMP7.cpp:561-583)
	int N, K;
	double cNN, cNB, cEC, cCC;
	bool sAllRows; // cCC finite: rows i < N can hold the minimum of S (NGCL / LOCAL); GLOBAL: only row N
	bool nnInf;    // T_SP(N,N) = 0 (GLOBAL / NGCL): the begin state is reachable from read position 1 only
};
__device__ __forceinline__ void setSeqConst(SeqConst& q, int N, int K, double cNN, double cNB, double cEC, double cCC) {
	q.N = N; q.K = K; q.cNN = cNN; q.cNB = cNB; q.cEC = cEC; q.cCC = cCC;
	q.sAllRows = cCC < INFINITY;
	q.nnInf = !(cNN < INFINITY);
}
// DP_M(i,0) + entry cost: the B candidate of cell (i, j)
__device__ __forceinline__ double entryCand(const SeqConst& q, int i, double ent) {
	if (q.nnInf) return i == 1 ? (0.0 + q.cNB) + ent : INFINITY;
	return col0M(i, q.cNN, q.cNB) + ent;
}
// emission tables of the lane's column in shared memory: [M|I][base 0..3, 4 = outside the band (+inf)][lane]
__device__ __forceinline__ void storeEmis(double* emis, const ColParams& p, int lane) {
	emis[0 * 32 + lane] = p.eM0; emis[1 * 32 + lane] = p.eM1; emis[2 * 32 + lane] = p.eM2; emis[3 * 32 + lane] = p.eM3;
	emis[4 * 32 + lane] = INFINITY;
	emis[5 * 32 + lane] = p.eI0; emis[6 * 32 + lane] = p.eI1; emis[7 * 32 + lane] = p.eI2; emis[8 * 32 + lane] = p.eI3;
	emis[9 * 32 + lane] = INFINITY;
}

// One cell (i, j): DP_M / DP_I / DP_D (src/BandedHMMP7.cpp:812-825) and its 4-bit back-pointer.
// diag = (i-1,j-1), up = (i-1,j), left = (i,j-1).  eM / eI are the emission costs of the read base (+inf for a
// cell outside the band, which therefore comes out as +inf like the reference's untouched matrix entries).
template <bool USE_B>
__device__ __forceinline__ uint32_t cellUpdate(const ColParams& p, double cB, double eM, double eI, bool first, bool dOk,
		double dgM, double dgI, double dgD, double upM, double upI, double lM, double lD,
		double& nM, double& nI, double& nD) {
	// candidates of DP_M; the trace (src/BandedHMMP7.cpp:962-971) always ranks B too, and only B and I at j == 1
	const double cM = dgM + p.tMMp, cI = dgI + p.tIMp, cD = dgD + p.tDMp;
	double mn = cB;
	uint32_t pm = 0;
	if (cM < mn) { mn = cM; pm = 1; }
	if (cI < mn) { mn = cI; pm = 2; }
	if (cD < mn) { mn = cD; pm = 3; }
	if (first) pm = (cI < cB) ? 2 : 0;
	// useB at :812-819; no B in the downstream box :866-872
	nM = eM + (USE_B ? mn : fmin(cM, fmin(cI, cD)));
	// DP_I: :820-822
	const double iM = upM + p.tMI, iI = upI + p.tII;
	const bool pi = iI < iM;
	nI = eI + (pi ? iI : iM);
	// DP_D: :823-825 (only 1 < j < K)
	const double dM = lM + p.tMDp, dD = lD + p.tDDp;
	const bool pd = dD < dM;
	nD = dOk ? (pd ? dD : dM) : INFINITY;
	return pm | ((uint32_t) pi << 2) | ((uint32_t) (pd && dOk) << 3);
}

// S(i,j) = DP_M(i,j) + exit(j) (+ the I_K -> E column) + special-state costs (src/BandedHMMP7.cpp:883-891);
// only rows that can hold a finite S are visited (row N alone when T_SP(C,C) = 0, i.e. GLOBAL)
__device__ __forceinline__ void cellScore(const ColParams& p, const SeqConst& q, int i, int j, double nM, double nI, Best& best) {
	const double rowPen = i < q.N ? q.cCC * (double) (q.N - i) : 0.0;
	double sv = (nM + p.ext) + q.cEC;
	if (i < q.N) sv += rowPen;
	if (better(sv, j, i, best)) { best.v = sv; best.j = j; best.i = i; }
	if (j == q.K) {
		double s2 = (nI + p.tIMj) + q.cEC;
		if (i < q.N) s2 += rowPen;
		if (better(s2, q.K + 1, i, best)) { best.v = s2; best.j = q.K + 1; best.i = i; }
	}
}

// the four pointer bits of the 32 cells of a wavefront step as four ballots, stored by lane 0
__device__ __forceinline__ void storePtrBits(uint4* __restrict__ dst, bool b0, bool b1, bool b2, bool b3, int lane) {
	const uint32_t w0 = __ballot_sync(0xffffffffu, b0), w1 = __ballot_sync(0xffffffffu, b1);
	const uint32_t w2 = __ballot_sync(0xffffffffu, b2), w3 = __ballot_sync(0xffffffffu, b3);
	if (lane == 0) *dst = make_uint4(w0, w1, w2, w3);
}
// lean boxes store the two DP_M compare outcomes as they are: a = (I candidate < M candidate), b = (D candidate < the
// smaller of the two); decoded here into the B/M/I/D index of the general layout
__device__ __forceinline__ uint32_t leanPm(uint32_t a, uint32_t b) { return b ? 3u : (a ? 2u : 1u); }
__device__ __forceinline__ bool isLeanBox(const VRegion& R, bool nnInf) { return !R.strip && nnInf && R.i0 >= 2 && R.j0 >= 2; }

struct WarpCtx {
	const int8_t* x;   // read codes, x[i-1]
	SeqConst q;
	double* bnd;       // shared: [3][bndPitch]: M, I, D of the column left of the running stripe, rows i0-1 .. i1 (+ 32 slack entries)
	int bndPitch;
	double* emis;      // shared: [2][5][32] emission tables of the current stripe
	uint16_t* xbox;    // shared: [maxLen + 64] per-box emission offsets of the lean sweep
	Cell* win;         // shared: 2x2 window around the current chain corner, [a*2+b] = cell (ic-1+a, jc-1+b)
	Cell* winNew;      // shared: window being captured around the end corner of the running box
	int ic, jc;        // chain corner of win
};

// Sweep one box.  Returns the number of cell updates.
__device__ long long sweepBox(const HmmDev& h, const VRegion& Rsm, bool hasNext, WarpCtx& wsm, uint32_t* __restrict__ ptrBase, Best& best) {
	const int lane = threadIdx.x & 31;
	const VRegion R = Rsm; // registers: the box lives in shared memory, which the loop below also writes
	WarpCtx w = wsm;
	const int rows = R.i1 - R.i0 + 1, cols = R.j1 - R.j0 + 1;
	long long cells = 0;
	// ---- window around the end corner: start from what is known before this box ----
	if (lane < 4) {
		const int a = lane >> 1, b = lane & 1;
		const int ci = R.i1 - 1 + a, cj = R.j1 - 1 + b;
		Cell c;
		c.M = c.I = c.D = INFINITY;
		const int oa = ci - (w.ic - 1), ob = cj - (w.jc - 1);
		if (oa >= 0 && oa < 2 && ob >= 0 && ob < 2) c = w.win[oa * 2 + ob];
		else if (cj == 0) { c.M = c.I = col0M(ci, w.q.cNN, w.q.cNB); }
		w.winNew[lane] = c;
	}
	// ---- left boundary column (j0-1) for stripe 0 ----
	for (int r = lane; r <= rows; r += 32) {
		const int i = R.i0 - 1 + r;
		Cell c;
		c.M = c.I = c.D = INFINITY;
		if (R.j0 - 1 == 0) c.M = c.I = col0M(i, w.q.cNN, w.q.cNB);
		else if (r == 0) c = w.win[0];
		else if (r == 1) c = w.win[2];
		w.bnd[r] = c.M; w.bnd[w.bndPitch + r] = c.I; w.bnd[2 * w.bndPitch + r] = c.D;
	}
	__syncwarp();
	const Cell top0 = w.win[1]; // (i0-1, j0)
	const int nStripes = (cols + 31) / 32;
	for (int s = 0; s < nStripes; ++s) {
		const int jb = R.j0 + 32 * s;
		const int j = jb + lane;
		const int ncols = min(32, R.j1 - jb + 1);
		const bool colOn = lane < ncols;
		const int jj = colOn ? j : R.j1; // clamp for loads
		ColParams cp;
		loadCol(h, jj, cp);
		__syncwarp();
		storeEmis(w.emis, cp, lane);
		__syncwarp();
		// state: cur = (i-1, j) own column; diag = (i-1, j-1)
		double curM = INFINITY, curI = INFINITY, curD = INFINITY;
		if (s == 0 && lane == 0) { curM = top0.M; curI = top0.I; curD = top0.D; }
		double dgM = INFINITY, dgI = INFINITY, dgD = INFINITY;
		if (lane == 0) { dgM = w.bnd[0]; dgI = w.bnd[w.bndPitch]; dgD = w.bnd[2 * w.bndPitch]; }
		uint4* ptrStripe = reinterpret_cast<uint4*>(ptrBase + R.ptrOff) + (size_t) s * R.rowWords; // [step t] of this stripe
		const int steps = rows + ncols - 1;
		for (int t = 0; t < steps; ++t) {
			const int i = R.i0 + t - lane;
			double lM = __shfl_up_sync(0xffffffffu, curM, 1);
			double lI = __shfl_up_sync(0xffffffffu, curI, 1);
			double lD = __shfl_up_sync(0xffffffffu, curD, 1);
			if (lane == 0) {
				const int r = i - (R.i0 - 1);
				if (r <= rows) { lM = w.bnd[r]; lI = w.bnd[w.bndPitch + r]; lD = w.bnd[2 * w.bndPitch + r]; }
				else lM = lI = lD = INFINITY;
			}
			const bool on = colOn && (unsigned) (i - R.i0) < (unsigned) rows;
			bool pred = on;
			if (R.strip) {
				const int d = (i - R.pFrom) - (j - R.pStart);
				pred = on && d <= R.nIns && d >= -R.nDel;
			}
			// straight-line cell update: a lane outside the band looks up the +inf emission row, so its cell
			// comes out +inf like an untouched entry of the reference's matrices
			const int ic = min(max(i, 1), w.q.N);
			const int bb = pred ? (int) w.x[ic - 1] : 4;
			const double eM = w.emis[bb * 32 + lane], eI = w.emis[(5 + bb) * 32 + lane];
			const double cB = entryCand(w.q, i, cp.ent);
			double nM, nI, nD;
			uint32_t code = R.useB ? cellUpdate<true>(cp, cB, eM, eI, j == 1, cp.dOn && pred, dgM, dgI, dgD, curM, curI, lM, lD, nM, nI, nD)
			                       : cellUpdate<false>(cp, cB, eM, eI, j == 1, cp.dOn && pred, dgM, dgI, dgD, curM, curI, lM, lD, nM, nI, nD);
			cells += pred;
			// the end corner is recomputed by the next box: its S candidate is taken there
			if (pred && (i == w.q.N || w.q.sAllRows) && !(hasNext && i == R.i1 && j == R.j1)) cellScore(cp, w.q, i, j, nM, nI, best);
			storePtrBits(ptrStripe + t, code & 1u, code & 2u, code & 4u, code & 8u, lane);
			if (on) {
				if (pred && i >= R.i1 - 1 && j >= R.j1 - 1) {
					Cell c; c.M = nM; c.I = nI; c.D = nD;
					w.winNew[(i - (R.i1 - 1)) * 2 + (j - (R.j1 - 1))] = c;
				}
				if (lane == 31) {
					const int r = i - (R.i0 - 1);
					w.bnd[r] = nM; w.bnd[w.bndPitch + r] = nI; w.bnd[2 * w.bndPitch + r] = nD;
				}
				curM = nM; curI = nI; curD = nD;
			}
			dgM = lM; dgI = lI; dgD = lD;
		}
		__syncwarp();
		if (lane == 0) { w.bnd[0] = INFINITY; w.bnd[w.bndPitch] = INFINITY; w.bnd[2 * w.bndPitch] = INFINITY; }
		__syncwarp();
	}
	// commit the window
	__syncwarp();
	if (lane < 4) w.win[lane] = w.winNew[lane];
	wsm.ic = R.i1; wsm.jc = R.j1;
	__syncwarp();
	return cells;
}

// A plain rectangle (no seed strip) whose rows are all >= 2 and whose columns are all >= 2, in GLOBAL / NGCL mode: the
// begin state cannot be entered inside the box (DP_M(i, 0) = +inf for i >= 2), so the B candidate of every cell is +inf --
// it still wins the back-pointer when all three predecessors are +inf, exactly as whichMin does on four +inf candidates --
// and useB makes no difference.  Same arithmetic, same operand order, same first-strict-minimum rules as cellUpdate; what
// is gone is the per-step bookkeeping:
//  * cells outside the rows of the box run the same code: their emission offset (a per-box table, xbox) selects the +inf
//    emission row and their D state is forced to +inf; their back-pointer nibbles are never read;
//  * the S candidates and the corner window are looked at only in the last steps of a stripe (TAIL);
//  * a lane's two emission costs are two shared-memory loads off one address; the left-boundary column is an array of
//    (M, I, D) triples read by lane 0 with immediate offsets; the back-pointer word of 8 steps is assembled by a funnel
//    shift and stored by all lanes at once; the steps before the tail run in blocks of 8, fully unrolled.
#ifndef HU_LEAN_UNROLL
#define HU_LEAN_UNROLL 4
#endif
struct LeanState {
	double curM, curI, curD, dgM, dgI, dgD;
};
// shared-memory accesses by 32-bit shared address (the pointers of WarpCtx travel through a struct and the compiler
// otherwise falls back to generic addressing: an address-space conversion per access)
__device__ __forceinline__ uint32_t sAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ double ldsF64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void stsF64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint32_t ldsU16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// One wavefront step.  bndIn: shared address of the boundary triple lane 0 reads at this step (row offset t + 1);
// xoA: shared address of this lane's emission offset for this step; emisA: shared address of this lane's slot in the
// emission table; d = t - lane: the lane's row offset; rowsEff: rows of the box (0 for a lane beyond the stripe's columns).
template <bool TAIL, bool CORNER>
__device__ __forceinline__ void leanStep(LeanState& q, const ColLean& cp, uint32_t bndIn, uint32_t bndOut, uint32_t xoA, uint32_t emisA, uint4* __restrict__ ptrStep, int d, int lane,
		uint32_t rowsEff, const HmmDev& h, const VRegion& R, bool hasNext, WarpCtx& w, int j, int jj, Best& best) {
	double lM = __shfl_up_sync(0xffffffffu, q.curM, 1);
	double lI = __shfl_up_sync(0xffffffffu, q.curI, 1);
	double lD = __shfl_up_sync(0xffffffffu, q.curD, 1);
	// lane 0 takes its left neighbour from the boundary column: predicated loads straight into the shuffle's registers
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %3, 0;\n\t@p ld.shared.f64 %0, [%4];\n\t@p ld.shared.f64 %1, [%4+8];\n\t@p ld.shared.f64 %2, [%4+16];\n\t}"
	             : "+d"(lM), "+d"(lI), "+d"(lD) : "r"(lane), "r"(bndIn));
	const bool pred = (uint32_t) d < rowsEff;
	const uint32_t e = emisA + ldsU16(xoA); // byte offset of the base's emission row (the +inf row outside the box)
	const double eM = ldsF64(e), eI = ldsF64(e + 5 * 32 * 8);
	// DP_M: first strict minimum of (M, I, D); the B candidate is +inf here and a cell whose three predecessors are all +inf is
	// never on a traced path (every state the traceback visits has a finite cost), so its pointer needs no B case
	const double cM = q.dgM + cp.tMMp, cI = q.dgI + cp.tIMp, cD = q.dgD + cp.tDMp;
	const bool pa = cI < cM;
	double mn = pa ? cI : cM;
	const bool pb = cD < mn;
	mn = pb ? cD : mn;
	const double nM = eM + mn;
	// DP_I
	const double iM = q.curM + cp.tMI, iI = q.curI + cp.tII;
	const bool pi = iI < iM;
	const double nI = eI + (pi ? iI : iM);
	// DP_D (+inf transition costs where the column has no D state; +inf outside the rows of the box)
	const double dM = lM + cp.tMDp, dD = lD + cp.tDDp;
	const bool pd = dD < dM;
	const double nD = pred ? (pd ? dD : dM) : INFINITY;
	storePtrBits(ptrStep, pa, pb, pi, pd, lane);
	// lane 31 hands its column to the next stripe: boundary entry of row offset d + 1 (bndOut: its shared address; in the banded
	// sweep the same array, 31 entries behind lane 0's read)
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %0, 31;\n\tsetp.lt.and.u32 p, %1, %2, p;\n\t@p st.shared.f64 [%3], %4;\n\t@p st.shared.f64 [%3+8], %5;\n\t@p st.shared.f64 [%3+16], %6;\n\t}"
	             :: "r"(lane), "r"((uint32_t) d), "r"(rowsEff), "r"(bndOut), "d"(nM), "d"(nI), "d"(nD) : "memory");
	if (TAIL && pred) {
		const int i = R.i0 + d;
		// the end corner is recomputed by the next box: its S candidate is taken there
		if ((i == w.q.N || w.q.sAllRows) && !(hasNext && i == R.i1 && j == R.j1)) {
			ColParams full; // exit cost etc.: only touched here
			full.ext = __ldg(h.ext + jj); full.tIMj = __ldg(h.tr + T_IM * h.Kp + jj);
			cellScore(full, w.q, i, j, nM, nI, best);
		}
		if (CORNER && i >= R.i1 - 1 && j >= R.j1 - 1) {
			Cell c; c.M = nM; c.I = nI; c.D = nD;
			w.winNew[(i - (R.i1 - 1)) * 2 + (j - (R.j1 - 1))] = c;
		}
	}
	q.curM = nM; q.curI = nI; q.curD = nD;
	q.dgM = lM; q.dgI = lI; q.dgD = lD;
}

// The lean step inside the full-matrix pass: the region (1..N, 1..K) is not a lean box (it holds row 1 and column 1), so its
// back-pointers are in the general encoding: the DP_M index 1 / 2 / 3 instead of the two raw compare outcomes.
template <bool TAIL>
__device__ __forceinline__ void leanStepFull(LeanState& q, const ColLean& cp, uint32_t bndIn, uint32_t bndOut, uint32_t xoA, uint32_t emisA, uint4* __restrict__ ptrStep, int d, int lane,
		uint32_t rowsEff, const HmmDev& h, const VRegion& R, WarpCtx& w, int j, int jj, Best& best) {
	double lM = __shfl_up_sync(0xffffffffu, q.curM, 1);
	double lI = __shfl_up_sync(0xffffffffu, q.curI, 1);
	double lD = __shfl_up_sync(0xffffffffu, q.curD, 1);
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %3, 0;\n\t@p ld.shared.f64 %0, [%4];\n\t@p ld.shared.f64 %1, [%4+8];\n\t@p ld.shared.f64 %2, [%4+16];\n\t}"
	             : "+d"(lM), "+d"(lI), "+d"(lD) : "r"(lane), "r"(bndIn));
	const bool pred = (uint32_t) d < rowsEff;
	const uint32_t e = emisA + ldsU16(xoA);
	const double eM = ldsF64(e), eI = ldsF64(e + 5 * 32 * 8);
	const double cM = q.dgM + cp.tMMp, cI = q.dgI + cp.tIMp, cD = q.dgD + cp.tDMp;
	const bool pa = cI < cM;
	double mn = pa ? cI : cM;
	const bool pb = cD < mn;
	mn = pb ? cD : mn;
	const double nM = eM + mn;
	const double iM = q.curM + cp.tMI, iI = q.curI + cp.tII;
	const bool pi = iI < iM;
	const double nI = eI + (pi ? iI : iM);
	const double dM = lM + cp.tMDp, dD = lD + cp.tDDp;
	const bool pd = dD < dM;
	const double nD = pred ? (pd ? dD : dM) : INFINITY;
	// general encoding of the DP_M pointer: 1 (M), 2 (I), 3 (D) -> bit 0 = b | !a, bit 1 = a | b; a traced cell is never all +inf
	storePtrBits(ptrStep, pb || !pa, pa || pb, pi, pd, lane);
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %0, 31;\n\tsetp.lt.and.u32 p, %1, %2, p;\n\t@p st.shared.f64 [%3], %4;\n\t@p st.shared.f64 [%3+8], %5;\n\t@p st.shared.f64 [%3+16], %6;\n\t}"
	             :: "r"(lane), "r"((uint32_t) d), "r"(rowsEff), "r"(bndOut), "d"(nM), "d"(nI), "d"(nD) : "memory");
	if (TAIL && pred) {
		const int i = R.i0 + d;
		if (i == w.q.N || w.q.sAllRows) {
			ColParams full;
			full.ext = __ldg(h.ext + jj); full.tIMj = __ldg(h.tr + T_IM * h.Kp + jj);
			cellScore(full, w.q, i, j, nM, nI, best);
		}
	}
	q.curM = nM; q.curI = nI; q.curD = nD;
	q.dgM = lM; q.dgI = lI; q.dgD = lD;
}

__device__ long long sweepBoxLean(const HmmDev& h, const VRegion& Rsm, bool hasNext, WarpCtx& wsm, uint32_t* __restrict__ ptrBase, Best& best) {
	const int lane = threadIdx.x & 31;
	const VRegion R = Rsm;
	WarpCtx w = wsm;
	const int rows = R.i1 - R.i0 + 1, cols = R.j1 - R.j0 + 1;
	double* const bnd = w.bnd; // [rows + 33][3]: (M, I, D) of the column left of the running stripe, row i0-1+r at entry r
	// ---- window around the end corner: start from what is known before this box (as sweepBox) ----
	if (lane < 4) {
		const int a = lane >> 1, b = lane & 1;
		const int ci = R.i1 - 1 + a, cj = R.j1 - 1 + b;
		Cell c;
		c.M = c.I = c.D = INFINITY;
		const int oa = ci - (w.ic - 1), ob = cj - (w.jc - 1);
		if (oa >= 0 && oa < 2 && ob >= 0 && ob < 2) c = w.win[oa * 2 + ob];
		w.winNew[lane] = c;
	}
	// ---- left boundary column (j0 - 1 >= 1): only its rows i0-1 and i0 are known (the chain corner); 32 slack entries of
	//      +inf behind row i1 so that lane 0 can read one entry per step to the end of a stripe ----
	for (int r = lane; r <= rows + 32; r += 32) {
		Cell c;
		c.M = c.I = c.D = INFINITY;
		if (r == 0) c = w.win[0];
		else if (r == 1) c = w.win[2];
		bnd[3 * r] = c.M; bnd[3 * r + 1] = c.I; bnd[3 * r + 2] = c.D;
	}
	// ---- emission offsets per wavefront row offset d = t - lane in [-31, rows + 31): the base's row of the emission table
	//      inside the box, the +inf row outside ----
	uint16_t* const xbox = w.xbox + 32;
	for (int d = -32 + lane; d < rows + 32; d += 32)
		xbox[d] = (uint16_t) ((d >= 0 && d < rows ? (int) w.x[R.i0 - 1 + d] : 4) * 256);
	__syncwarp();
	const Cell top0 = w.win[1]; // (i0-1, j0)
	const int nStripes = (cols + 31) / 32;
	for (int s = 0; s < nStripes; ++s) {
		const int jb = R.j0 + 32 * s;
		const int j = jb + lane;
		const int ncols = min(32, R.j1 - jb + 1);
		const bool colOn = lane < ncols;
		const int jj = colOn ? j : R.j1; // clamp for loads
		ColLean cp;
		loadColLean(h, jj, cp);
		__syncwarp();
		storeEmisCol(h, jj, w.emis, lane);
		__syncwarp();
		const double* emisLane = w.emis + lane;
		LeanState q;
		q.curM = q.curI = q.curD = INFINITY; // (i-1, j): row i0-1 of the own column
		if (s == 0 && lane == 0) { q.curM = top0.M; q.curI = top0.I; q.curD = top0.D; }
		q.dgM = q.dgI = q.dgD = INFINITY;    // (i-1, j-1)
		if (lane == 0) { q.dgM = bnd[0]; q.dgI = bnd[1]; q.dgD = bnd[2]; }
		uint4* ptrStripe = reinterpret_cast<uint4*>(ptrBase + R.ptrOff) + (size_t) s * R.rowWords;
		const int steps = rows + ncols - 1;
		// running addresses: boundary triple of row offset t + 1 (24 bytes per entry), emission offset of row offset t - lane
		uint32_t bndIn = sAddr(bnd) + 24, xoA = sAddr(xbox - lane);
		const uint32_t emisA = sAddr(emisLane), rowsEff = colOn ? (uint32_t) rows : 0u;
		uint4* ptrStep = ptrStripe;
		int d = -lane;
		// steps that cannot touch row i1 - 1 or i1 (S candidates, corner window), in unrolled blocks
		const int bulk = w.q.sAllRows ? 0 : (max(rows - 2, 0) / HU_LEAN_UNROLL * HU_LEAN_UNROLL);
		int t = 0;
		for (; t < bulk; t += HU_LEAN_UNROLL) {
#pragma unroll
			for (int k = 0; k < HU_LEAN_UNROLL; ++k)
				leanStep<false, true>(q, cp, bndIn + 24 * k, bndIn + 24 * k - 744, xoA + 2 * k, emisA, ptrStep + k, d + k, lane, rowsEff, h, R, hasNext, w, j, jj, best);
			bndIn += 24 * HU_LEAN_UNROLL; xoA += 2 * HU_LEAN_UNROLL; ptrStep += HU_LEAN_UNROLL; d += HU_LEAN_UNROLL;
		}
		for (; t < steps; ++t) {
			leanStep<true, true>(q, cp, bndIn, bndIn - 744, xoA, emisA, ptrStep, d, lane, rowsEff, h, R, hasNext, w, j, jj, best);
			bndIn += 24; xoA += 2; ptrStep += 1; d += 1;
		}
		__syncwarp();
		if (lane == 0) { bnd[0] = INFINITY; bnd[1] = INFINITY; bnd[2] = INFINITY; }
		__syncwarp();
	}
	__syncwarp();
	if (lane < 4) w.win[lane] = w.winNew[lane];
	wsm.ic = R.i1; wsm.jc = R.j1;
	__syncwarp();
	return lane == 0 ? (long long) rows * cols : 0;
}

__device__ __forceinline__ void initWindow(WarpCtx& w, int i0, int j0) {
	const int lane = threadIdx.x & 31;
	if (lane < 4) {
		const int a = lane >> 1, b = lane & 1;
		const int ci = i0 - 1 + a, cj = j0 - 1 + b;
		Cell c;
		c.M = c.I = c.D = INFINITY;
		if (cj == 0) c.M = c.I = col0M(ci, w.q.cNN, w.q.cNB);
		w.win[lane] = c;
	}
	w.ic = i0; w.jc = j0;
	__syncwarp();
}

// the 4-bit back-pointer of cell (i, j) in the general encoding: bits 0-1 = B/M/I/D index of DP_M, bit 2 = DP_I, bit 3 = DP_D
__device__ __forceinline__ uint32_t ptrAt(const VRegion* regs, int nReg, const uint32_t* __restrict__ ptrBase, int i, int j, bool nnInf) {
	for (int r = nReg - 1; r >= 0; --r) {
		const VRegion& R = regs[r];
		if (i >= R.i0 && i <= R.i1 && j >= R.j0 && j <= R.j1) {
			const int ln = (j - R.j0) & 31, st = (j - R.j0) >> 5;
			const int t = (i - R.i0) + ln; // wavefront step of the cell inside its stripe
			const uint4 wv = reinterpret_cast<const uint4*>(ptrBase + R.ptrOff)[(size_t) st * R.rowWords + t];
			const uint32_t b0 = (wv.x >> ln) & 1u, b1 = (wv.y >> ln) & 1u, b2 = (wv.z >> ln) & 1u, b3 = (wv.w >> ln) & 1u;
			const uint32_t pm = isLeanBox(R, nnInf) ? leanPm(b0, b1) : (b0 | (b1 << 1));
			return pm | (b2 << 2) | (b3 << 3);
		}
	}
	return 0;
}

// getPaddingSeq(..., JUSTIFIED) of an insert into a gap of Lg columns (src/BandedHMMP7.cpp:1168-1181),
// written as base codes; the head duplication for short inserts is the reference's literal behaviour.
__device__ void justifiedPad(int8_t* out, int Lg, const int8_t* ins, int n) {
	if (n == 0) {
		for (int c = 0; c < Lg; ++c) out[c] = HU_CODE_GAP;
		return;
	}
	if (n >= Lg) {
		const int fl = Lg / 2, ce = Lg - fl;
		for (int c = 0; c < fl; ++c) out[c] = ins[c];
		for (int c = 0; c < ce; ++c) out[fl + c] = ins[n - ce + c];
	} else {
		const int fl = n / 2, ce = n - fl;
		for (int c = 0; c < fl; ++c) out[c] = ins[c];
		for (int c = 0; c < Lg - n; ++c) out[fl + c] = HU_CODE_GAP;
		for (int c = 0; c < ce; ++c) out[fl + Lg - n + c] = ins[c];
	}
}

// After the sweeps: S argmin -> traceback -> trace string (reversed) -> aligned codes -> result row.
__device__ void finishSeq(const HmmDev& h, const VRegion* regs, int nReg, const uint32_t* __restrict__ ptrBase, WarpCtx& w,
		Best best, long long cells, int flags, const int8_t* __restrict__ gx, char* __restrict__ trace, int8_t* __restrict__ alignedRow,
		hu_align_row* __restrict__ row) {
	const int lane = threadIdx.x & 31;
	if (lane != 0) return;
	const int K = h.K;
	row->cells = cells;
	row->cost = best.v;
	if (best.v == INFINITY) { // unreachable after the full pass unless the profile forbids the read
		row->flags = flags | HU_ALIGN_INVALID;
		row->traceLen = 0;
		row->seqStart = row->seqEnd = row->hmmStart = row->hmmEnd = row->csStart = row->csEnd = 0;
		return;
	}
	char s = best.j <= K ? 'M' : 'I';
	int i = best.i, j = best.j <= K ? best.j : K;
	const int alnEnd = j, alnTo = best.i;
	int pos = 0;
	trace[pos++] = 'E';
	char last = 'E';
	while (i >= 1 && j >= 0) {
		trace[pos++] = s;
		last = s;
		if (s == 'M') {
			const uint32_t c = ptrAt(regs, nReg, ptrBase, i, j, w.q.nnInf) & 3u;
			s = "BMID"[c];
			i--; j--;
		} else if (s == 'I') {
			if (j > 0) {
				const uint32_t c = (ptrAt(regs, nReg, ptrBase, i, j, w.q.nnInf) >> 2) & 1u;
				s = c ? 'I' : 'M';
			} else { // column 0: whichMin(DP_M(i,0)+T[0](M,I), DP_I(i-1,0)+T[0](I,I)) in "BI"
				const double a = col0M(i, w.q.cNN, w.q.cNB) + __ldg(h.tr + T_MI * h.Kp);
				const double b = col0M(i - 1, w.q.cNN, w.q.cNB) + __ldg(h.tr + T_II * h.Kp);
				s = b < a ? 'I' : 'B';
			}
			i--;
		} else if (s == 'D') {
			const uint32_t c = (ptrAt(regs, nReg, ptrBase, i, j, w.q.nnInf) >> 3) & 1u;
			s = c ? 'D' : 'M';
			j--;
		} else
			break;
	}
	const int alnStart = j + 1, alnFrom = i + 1;
	if (last != 'B') trace[pos++] = 'B';
	const int csStart = __ldg(h.p2cs + alnStart), csEnd = __ldg(h.p2cs + alnEnd);
	row->seqStart = alnFrom; row->seqEnd = alnTo;
	row->hmmStart = alnStart; row->hmmEnd = alnEnd;
	row->csStart = csStart; row->csEnd = csEnd;
	row->traceLen = pos;
	const bool valid = 0 < alnFrom && alnFrom <= alnTo && 0 < alnStart && alnStart <= alnEnd && alnEnd <= K &&
			0 < csStart && csStart <= csEnd && csEnd <= h.L && best.v >= 0;
	if (!valid) flags |= HU_ALIGN_INVALID;
	row->flags = flags;
	if (!valid || !alignedRow) return;
	// buildGlobalAlign (src/BandedHMMP7.cpp:1008-1081) as codes, inside [csStart, csEnd] only: the pads
	// outside are '.' whenever the whole read is aligned (always so in GLOBAL mode)
	int rj = alnFrom, k = alnStart; // 1-based read position / profile position
	int insFrom = 0, insLen = 0;    // pending insert = x[insFrom-1 .. insFrom+insLen-2]
	int t = 0;
	for (int q = pos - 1; q >= 0; --q, ++t) {
		const char st = trace[q];
		if (st == 'M') {
			if (k > 1 && t > 1) {
				const int g = __ldg(h.p2cs + k) - __ldg(h.p2cs + k - 1) - 1;
				if (g > 0) justifiedPad(alignedRow + __ldg(h.p2cs + k - 1), g, gx + insFrom - 1, insLen);
			}
			insLen = 0;
			alignedRow[__ldg(h.p2cs + k) - 1] = gx[rj - 1];
			rj++; k++;
		} else if (st == 'I') {
			insFrom = rj;
			insLen = 0;
			while (q >= 0 && trace[q] == 'I') { insLen++; rj++; q--; t++; }
			q++; t--;
		} else if (st == 'D') {
			if (k > 1) {
				const int g = __ldg(h.p2cs + k) - __ldg(h.p2cs + k - 1) - 1;
				for (int c = 0; c < g; ++c) alignedRow[__ldg(h.p2cs + k - 1) + c] = HU_CODE_GAP;
			}
			alignedRow[__ldg(h.p2cs + k) - 1] = HU_CODE_GAP;
			k++;
		}
	}
}

// smem per warp: codes (maxLen, padded to 8) | bnd 3*(maxLen+2) doubles | win 4 cells | winNew 4 cells | regions
__host__ __device__ inline size_t viterbiWarpSmem(int maxLen) {
	const size_t codes = ((size_t) maxLen + 15) / 16 * 16;
	return codes + (size_t) 3 * (maxLen + 2 + 32) * sizeof(double) + 320 * sizeof(double) + 8 * sizeof(Cell) + 5 * sizeof(VRegion) + 16 +
	       ((size_t) 2 * (maxLen + 64) + 15) / 16 * 16;
}

struct WarpSmem {
	int8_t* codes;
	double* bnd;
	double* emis;
	Cell* win;
	Cell* winNew;
	VRegion* regs;
	uint16_t* xbox;
};
__device__ __forceinline__ WarpSmem carve(unsigned char* base, int maxLen) {
	WarpSmem s;
	const size_t codes = ((size_t) maxLen + 15) / 16 * 16;
	s.bnd = reinterpret_cast<double*>(base);
	s.emis = s.bnd + 3 * (maxLen + 2 + 32);
	s.win = reinterpret_cast<Cell*>(s.emis + 320);
	s.winNew = s.win + 4;
	s.regs = reinterpret_cast<VRegion*>(s.winNew + 4);
	s.codes = reinterpret_cast<int8_t*>(s.regs + 5);
	s.xbox = reinterpret_cast<uint16_t*>(s.codes + codes);
	return s;
}

// Pass 1: the banded plans.  Sequences without a usable plan, or whose band holds no finite path, are
// queued for pass 2 (src/HmmUFOtu_main.cpp:86-96).
#ifndef HU_VIT_REGS
#define HU_VIT_REGS 96 // 3 CTAs of 7 warps per SM
#endif
__global__ void __maxnreg__(HU_VIT_REGS) viterbiBandedKernel(HmmDev h, VitArgs a) {
	extern __shared__ __align__(16) unsigned char smemRaw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	WarpSmem sm = carve(smemRaw + (size_t) warp * ((viterbiWarpSmem(a.maxLen) + 15) / 16 * 16), a.maxLen);
	// persistent: every warp takes the next sequence off a shared cursor, so that long and short reads, banded and failed
	// bands balance over the resident warps (a fixed 8-reads-per-CTA grid left the last wave half empty)
	for (;;) {
		int seq = 0;
		if (lane == 0) seq = atomicAdd(a.bandCursor, 1);
		seq = __shfl_sync(0xffffffffu, seq, 0);
		if (seq >= a.nSeq) return;
		const VPlan& plan = a.plans[seq];
		const int N = plan.N;
		const int8_t* gx = a.codes + plan.codeOff;
		if (plan.nReg == 0) {
			if (lane == 0) {
				const int slot = atomicAdd(a.fullCount, 1);
				a.fullList[slot] = seq;
				a.rows[seq].flags = plan.flags;
				a.rows[seq].cells = 0;
			}
			continue;
		}
		__syncwarp();
		for (int i = lane; i < N; i += 32) sm.codes[i] = gx[i];
		if (lane < plan.nReg) sm.regs[lane] = plan.reg[lane];
		__syncwarp();
		WarpCtx w;
		w.x = sm.codes;
		setSeqConst(w.q, N, h.K, a.cNN, a.cNB, a.cEC, a.cCC);
		w.bnd = sm.bnd; w.bndPitch = a.maxLen + 2 + 32; w.emis = sm.emis;
		w.win = sm.win; w.winNew = sm.winNew; w.xbox = sm.xbox;
		initWindow(w, sm.regs[0].i0, sm.regs[0].j0);
		Best best; best.v = INFINITY; best.j = 0x7fffffff; best.i = 0x7fffffff;
		long long cells = 0;
		for (int r = 0; r < plan.nReg; ++r) {
			const VRegion& R = sm.regs[r];
			cells += isLeanBox(R, w.q.nnInf) ? sweepBoxLean(h, R, r + 1 < plan.nReg, w, a.ptrs, best) : sweepBox(h, R, r + 1 < plan.nReg, w, a.ptrs, best);
		}
		// warp argmin
		for (int o = 16; o > 0; o >>= 1) {
			Best ob;
			ob.v = __shfl_xor_sync(0xffffffffu, best.v, o);
			ob.j = __shfl_xor_sync(0xffffffffu, best.j, o);
			ob.i = __shfl_xor_sync(0xffffffffu, best.i, o);
			if (better(ob.v, ob.j, ob.i, best)) best = ob;
			cells += __shfl_xor_sync(0xffffffffu, cells, o);
		}
		if (best.v == INFINITY) { // banded version failed
			if (lane == 0) {
				const int slot = atomicAdd(a.fullCount, 1);
				a.fullList[slot] = seq;
				a.rows[seq].flags = plan.flags | HU_ALIGN_FALLBACK;
				a.rows[seq].cells = cells;
			}
			continue;
		}
		__syncwarp();
		int8_t* arow = a.aligned ? a.aligned + (size_t) seq * h.L : NULL;
		finishSeq(h, sm.regs, plan.nReg, a.ptrs, w, best, cells, plan.flags, gx, a.traces + plan.traceOff, arow, a.rows + seq);
		__syncwarp();
	}
}

// Pass 2: full Viterbi (src/BandedHMMP7.cpp:748-780) for the queued sequences.  One CTA per sequence, persistent:
// the 32-column stripes of the N x K matrix are dealt round-robin to the warps, which run as a software
// pipeline -- stripe s+1 trails stripe s by one warp-width of rows, the last column of a stripe travels
// through a ring of shared-memory buffers, a per-stripe progress counter orders the hand-over.  Same cell
// code as the banded pass (cellUpdate), so costs and back-pointers are bit-identical to a single-warp sweep.
__host__ __device__ inline size_t viterbiFullSmem(int maxLen, int nW) {
	return ((size_t) maxLen + 15) / 16 * 16 + ((size_t) 2 * (maxLen + 64) + 15) / 16 * 16 + (size_t) nW * (3 * (maxLen + 2 + 32) + 320) * sizeof(double) +
			2048 * sizeof(int) + (size_t) nW * (sizeof(Best) + sizeof(long long)) + sizeof(VRegion) + 64;
}

// Round 2: a stripe beyond the first (columns >= 33) runs its first 32 steps -- the only ones that can touch read position 1,
// where the begin state enters -- with the general cell code and everything after that with the lean step of the banded sweep
// (same arithmetic; see leanStep); the boundary columns are (M, I, D) triples; the hand-over between consecutive stripes is
// checked and published once per block of four steps instead of every step.
__global__ void __launch_bounds__(256) viterbiFullKernel(HmmDev h, VitArgs a) {
	extern __shared__ __align__(16) unsigned char smemRaw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nW = blockDim.x >> 5;
	const int pitchE = a.maxLen + 2 + 32; // entries (rows 0 .. N, + 32 of slack) per boundary array
	double* bndAll = reinterpret_cast<double*>(smemRaw); // [nW][pitchE][3]
	double* emis = bndAll + (size_t) nW * 3 * pitchE + (size_t) warp * 320;
	volatile int* progress = reinterpret_cast<volatile int*>(bndAll + (size_t) nW * (3 * pitchE + 320));
	Best* wBest = reinterpret_cast<Best*>(const_cast<int*>(progress) + 2048);
	long long* wCells = reinterpret_cast<long long*>(wBest + nW);
	VRegion* reg = reinterpret_cast<VRegion*>(wCells + nW);
	int* sIdx = reinterpret_cast<int*>(reg + 1);
	int8_t* codes = reinterpret_cast<int8_t*>(sIdx + 4);
	uint16_t* xboxBase = reinterpret_cast<uint16_t*>(codes + ((size_t) a.maxLen + 15) / 16 * 16);
	const int count = *a.fullCount;
	uint32_t* slab = a.fullPtrs + (size_t) blockIdx.x * a.fullSlabWords;
	const int K = h.K;
	const int nStripes = (K + 31) / 32;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) *sIdx = atomicAdd(a.fullCursor, 1);
		__syncthreads();
		const int idx = *sIdx;
		if (idx >= count) return;
		const int seq = a.fullList[idx];
		const VPlan& plan = a.plans[seq];
		const int N = plan.N;
		const int8_t* gx = a.codes + plan.codeOff;
		for (int i = threadIdx.x; i < N; i += blockDim.x) codes[i] = gx[i];
		for (int i = threadIdx.x; i < nStripes; i += blockDim.x) progress[i] = 0;
		// emission offsets per row offset d = i - 1 in [-32, N + 32) (leanStep), and the parts of the boundary arrays no stripe
		// writes: row 0 and the slack behind row N are +inf
		uint16_t* const xbox = xboxBase + 32;
		for (int d = -32 + (int) threadIdx.x; d < N + 32; d += blockDim.x) xbox[d] = (uint16_t) ((d >= 0 && d < N ? (int) gx[d] : 4) * 256);
		for (int k = threadIdx.x; k < nW * 33; k += blockDim.x) {
			const int wq = k / 33, e = k % 33;
			double* p = bndAll + ((size_t) wq * pitchE + (e == 32 ? 0 : N + 1 + e)) * 3;
			p[0] = p[1] = p[2] = INFINITY;
		}
		if (threadIdx.x == 0) {
			VRegion R;
			R.j0 = 1; R.j1 = K; R.i0 = 1; R.i1 = N;
			R.strip = 0; R.pStart = R.pFrom = R.nIns = R.nDel = 0; R.useB = 1;
			R.rowWords = N + 31; R.pad = 0; // steps per stripe slot
			R.ptrOff = 0;
			*reg = R;
		}
		__syncthreads();
		WarpCtx w;
		w.x = codes;
		setSeqConst(w.q, N, K, a.cNN, a.cNB, a.cEC, a.cCC);
		w.bnd = NULL; w.bndPitch = 0; w.emis = NULL; w.win = NULL; w.winNew = NULL; w.xbox = NULL; w.ic = w.jc = 0;
		const SeqConst& q = w.q;
		const VRegion R = *reg;
		Best best; best.v = INFINITY; best.j = 0x7fffffff; best.i = 0x7fffffff;
		long long cells = 0;
		for (int s = warp; s < nStripes; s += nW) {
			const int jb = 1 + 32 * s;
			const int j = jb + lane;
			const int ncols = min(32, K - jb + 1);
			const bool colOn = lane < ncols;
			const int jj = colOn ? j : K;
			ColParams cp;
			loadCol(h, jj, cp);
			ColLean cl;
			cl.tMMp = cp.tMMp; cl.tIMp = cp.tIMp; cl.tDMp = cp.tDMp; cl.tMI = cp.tMI; cl.tII = cp.tII;
			cl.tMDp = cp.dOn ? cp.tMDp : INFINITY; cl.tDDp = cp.dOn ? cp.tDDp : INFINITY;
			__syncwarp();
			storeEmis(emis, cp, lane);
			__syncwarp();
			const double* in = bndAll + (size_t) ((s + nW - 1) % nW) * 3 * pitchE; // written by stripe s-1: entry i = row i
			double* out = bndAll + (size_t) (s % nW) * 3 * pitchE;
			LeanState st;
			st.curM = st.curI = st.curD = INFINITY; // row 0
			st.dgM = st.dgI = st.dgD = INFINITY;    // (0, jb-1)
			uint4* ptrStripe = reinterpret_cast<uint4*>(slab) + (size_t) s * (N + 31);
			const int steps = N + ncols - 1;
			// ---- general steps: the whole first stripe (column 1, column 0 to its left), the first 32 steps of the others (row 1) ----
			const int tG = (s == 0 || !q.nnInf) ? steps : min(32, steps);
			int t = 0;
			for (; t < tG; ++t) {
				const int i = 1 + t - lane;
				if (s > 0) { // lane 0 needs row t+1 of the previous stripe's last column
					const int need = min(t + 1, N);
					while (progress[s - 1] < need) { }
					__threadfence_block();
				}
				double lM = __shfl_up_sync(0xffffffffu, st.curM, 1);
				double lI = __shfl_up_sync(0xffffffffu, st.curI, 1);
				double lD = __shfl_up_sync(0xffffffffu, st.curD, 1);
				if (lane == 0) {
					if (i <= N) {
						if (s == 0) { lM = lI = col0M(i, q.cNN, q.cNB); lD = INFINITY; }
						else { lM = in[3 * i]; lI = in[3 * i + 1]; lD = in[3 * i + 2]; }
					} else
						lM = lI = lD = INFINITY;
				}
				const bool on = colOn && i >= 1 && i <= N;
				uint32_t code = 0;
				double nM = INFINITY, nI = INFINITY, nD = INFINITY;
				if (on) {
					const int bb = codes[i - 1];
					code = cellUpdate<true>(cp, entryCand(q, i, cp.ent), emis[bb * 32 + lane], emis[(5 + bb) * 32 + lane], j == 1, cp.dOn,
							st.dgM, st.dgI, st.dgD, st.curM, st.curI, lM, lD, nM, nI, nD);
					if (i == N || q.sAllRows) cellScore(cp, q, i, j, nM, nI, best);
				}
				storePtrBits(ptrStripe + t, code & 1u, code & 2u, code & 4u, code & 8u, lane);
				if (on) {
					if (lane == 31) {
						out[3 * i] = nM; out[3 * i + 1] = nI; out[3 * i + 2] = nD;
						__threadfence_block();
						progress[s] = i;
					}
					st.curM = nM; st.curI = nI; st.curD = nD;
				} else if (i > N) { st.curM = st.curI = st.curD = INFINITY; }
				st.dgM = lM; st.dgI = lI; st.dgD = lD;
			}
			// ---- lean steps (rows >= 2, columns >= 33; leanStep decodes to the general pointer encoding through isLeanBox = false
			//      here, so the pointer bits are stored in the general form below) ----
			const uint32_t inA = sAddr(in), outA = sAddr(out), emisA = sAddr(emis + lane), xoA = sAddr(xbox - lane);
			const uint32_t rowsEff = colOn ? (uint32_t) N : 0u;
			for (; !q.sAllRows && t + 3 <= N - 3; t += 4) { // lane 0 stays below row N - 1: no S candidates in these steps (GLOBAL: row N only)
				const int need = min(t + 4, N);
				while (progress[s - 1] < need) { }
				__threadfence_block();
#pragma unroll
				for (int k = 0; k < 4; ++k)
					leanStepFull<false>(st, cl, inA + 24 * (t + k + 1), outA + 24 * (t + k - 30), xoA + 2 * (t + k), emisA, ptrStripe + t + k, t + k - lane, lane, rowsEff, h, R, w, j, jj, best);
				if (lane == 31 && t + 3 - 30 >= 1) {
					__threadfence_block();
					progress[s] = t + 3 - 30;
				}
			}
			for (; t < steps; ++t) {
				const int need = min(t + 1, N);
				while (progress[s - 1] < need) { }
				__threadfence_block();
				leanStepFull<true>(st, cl, inA + 24 * (t + 1), outA + 24 * (t - 30), xoA + 2 * t, emisA, ptrStripe + t, t - lane, lane, rowsEff, h, R, w, j, jj, best);
				if (lane == 31 && t - 30 >= 1 && t - 30 <= N) {
					__threadfence_block();
					progress[s] = t - 30;
				}
			}
			if (lane == 0) cells += (long long) N * ncols;
		}
		for (int o = 16; o > 0; o >>= 1) {
			Best ob;
			ob.v = __shfl_xor_sync(0xffffffffu, best.v, o);
			ob.j = __shfl_xor_sync(0xffffffffu, best.j, o);
			ob.i = __shfl_xor_sync(0xffffffffu, best.i, o);
			if (better(ob.v, ob.j, ob.i, best)) best = ob;
			cells += __shfl_xor_sync(0xffffffffu, cells, o);
		}
		if (lane == 0) { wBest[warp] = best; wCells[warp] = cells; }
		__syncthreads();
		if (warp == 0) {
			for (int k = 1; k < nW; ++k) {
				if (better(wBest[k].v, wBest[k].j, wBest[k].i, best)) best = wBest[k];
				cells += wCells[k];
			}
			int8_t* arow = a.aligned ? a.aligned + (size_t) seq * h.L : NULL;
			const int flags = a.rows[seq].flags;
			const long long prior = a.rows[seq].cells;
			finishSeq(h, reg, 1, slab, w, best, cells + prior, flags, gx, a.traces + plan.traceOff, arow, a.rows + seq);
		}
	}
}

// PE: orientation test (src/hmmufotu.cpp:629-639) and HmmAlignment::merge (src/BandedHMMP7.cpp:1188-1213).
// Mate rows live at nReads + r.  The forward row becomes the merged row.
__global__ void mergeMatesKernel(int nReads, int L, const hu_align_row* __restrict__ rows, hu_align_row* __restrict__ merged,
		int8_t* __restrict__ aligned) {
	const int r = blockIdx.x;
	if (r >= nReads) return;
	const hu_align_row a = rows[r], b = rows[nReads + r];
	hu_align_row m = a;
	m.cells = a.cells + b.cells;
	m.flags = a.flags | (b.flags & (HU_ALIGN_INVALID | HU_ALIGN_FALLBACK));
	const bool bad = (m.flags & HU_ALIGN_INVALID) || !(a.csStart <= b.csStart && a.csEnd <= b.csEnd);
	if (bad) {
		m.flags |= HU_ALIGN_CHIMERA;
	} else {
		m.seqStart = min(a.seqStart, b.seqStart); m.seqEnd = max(a.seqEnd, b.seqEnd);
		m.hmmStart = min(a.hmmStart, b.hmmStart); m.hmmEnd = max(a.hmmEnd, b.hmmEnd);
		m.csStart = min(a.csStart, b.csStart); m.csEnd = max(a.csEnd, b.csEnd);
		m.cost = a.cost + b.cost;
		int8_t* x = aligned + (size_t) r * L;
		const int8_t* y = aligned + (size_t) (nReads + r) * L;
		for (int c = b.csStart - 1 + threadIdx.x; c < b.csEnd; c += blockDim.x)
			if (x[c] == HU_CODE_PAD && y[c] != HU_CODE_PAD) x[c] = y[c];
	}
	if (threadIdx.x == 0) merged[r] = m;
}

__global__ void copyRowsKernel(int n, const hu_align_row* __restrict__ rows, hu_align_row* __restrict__ merged) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n) merged[r] = rows[r];
}

#define HU_VIT_SMEM_MAX (227 * 1024)
int configViterbiKernels() {
	HU_CUDA(cudaFuncSetAttribute(viterbiBandedKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_VIT_SMEM_MAX));
	HU_CUDA(cudaFuncSetAttribute(viterbiFullKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_VIT_SMEM_MAX));
	return HU_OK;
}

int launchViterbi(hu_batch* b, int mode) {
	hu_db* db = b->db;
	const HmmDev& h = db->hmm;
	// setSequenceMode (src/BandedHMMP7.cpp:561-583, src/BandedHMMP7Bg.cpp:33-35)
	const double Kbg = h.K >= 350 ? h.K : 350;
	const double term = 1.0 - Kbg / (Kbg + 1.0);
	double tNN = 0, tCC = 0;
	switch (mode) {
	case HU_GLOBAL: tNN = tCC = 0; break;
	case HU_LOCAL: tNN = tCC = term; break;
	case HU_NGCL: tNN = 0; tCC = term; break;
	case HU_CGNL: tNN = term; tCC = 0; break;
	default: setError("unknown alignment mode %d", mode); return HU_ERR_ARG;
	}
	VitArgs a = b->vit;
	a.cNN = -log(tNN); a.cCC = -log(tCC); a.cNB = -log(1.0 - tNN); a.cEC = -log(1.0);
	const int nSeq = b->nSeq;
	a.nSeq = nSeq;
	const size_t perWarp = (viterbiWarpSmem(a.maxLen) + 15) / 16 * 16;
	const int wpb = a.warpsPerBlock;
	const size_t smem = perWarp * wpb;
	int fullWarps = 8; // the full-matrix pass pipelines the stripes of ONE sequence over the warps of a CTA: as many as its boundary buffers allow
	while (fullWarps > 1 && viterbiFullSmem(a.maxLen, fullWarps) > 200 * 1024) fullWarps--;
	const size_t smemFull = viterbiFullSmem(a.maxLen, fullWarps);
	if (smem > HU_VIT_SMEM_MAX || smemFull > HU_VIT_SMEM_MAX) { setError("maxSeqLen %d too long for the Viterbi shared-memory layout", a.maxLen); return HU_ERR_ARG; }
	HU_CUDA(cudaMemsetAsync(a.fullCount, 0, 3 * sizeof(int), b->stream)); // count + cursor of the full-matrix queue, cursor of the banded pass
	HU_CUDA(cudaMemsetAsync(a.aligned, HU_CODE_PAD & 0xff, (size_t) nSeq * h.L, b->stream));
	viterbiBandedKernel<<<std::min((nSeq + wpb - 1) / wpb, b->vitGrid), wpb * 32, smem, b->stream>>>(h, a);
	viterbiFullKernel<<<b->fullSlots, fullWarps * 32, smemFull, b->stream>>>(h, a);
	b->launches += 2;
	if (b->paired) {
		mergeMatesKernel<<<b->nReads, 128, 0, b->stream>>>(b->nReads, h.L, a.rows, b->merged, a.aligned);
	} else
		copyRowsKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, a.rows, b->merged);
	b->launches += 1;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
