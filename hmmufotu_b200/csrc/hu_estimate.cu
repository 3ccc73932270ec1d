// K3: the "Estimate" stage of SEP -- PTUnrooted::estimateSeq for every (read, seed)
// (src/PhyloTreeUnrooted.cpp:849-877) followed by filterPlacements (src/HmmUFOtu_main.cpp:162-173).
//
// One warp per (read, seed), persistent over the item space; no block-level synchronisation.  The two message
// windows of the seed edge (linear form: 32 B of max-normalised exponentials + 8 B scale per column, see
// hu_internal.cuh) are streamed from HBM with 128-bit loads, columns across lanes.  Per column
//   g_j = (P(wur) E^U_j) * (P(wvr) E^V_j)              (= e^{R_j} of the reference up to the column scale)
// pass 1 counts the columns whose arg-max state differs from the read's (wnr, cpp:1018-1032); pass 2 re-reads
// the window (L1/L2-resident by then) and sums log(pi . (g_j * P(wnr) n_j)) + scales (cpp:872-874):
// 64 fused multiply-adds and one log per column, no exp.
#include "hu_batch.cuh"

namespace hu {

#define HU_EST_WARPS 8

struct EstArgs {
	int nReads, maxNSeed, pad0, pad1;
	const int8_t* aligned;
	const int* winStart;
	const int* winEnd;
	const int* seedIds;
	const int* nSeeds;
	hu_place_row* est;
	int* cursor;
};

struct EstShared {
	double P[3][16]; // P(wur), P(wvr), P(wnr)
	double Np[5][4]; // P(wnr) applied to the read's leaf vector, per base / gap
};

__global__ void __launch_bounds__(HU_EST_WARPS * 32) estimateKernel(EstArgs a, TreeDev t, Model m) {
	__shared__ EstShared sh[HU_EST_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	EstShared& ws = sh[warp];
	const int total = a.nReads * a.maxNSeed;
	int piMax = 0; // argmax of log(pi): state of a gap column of the read
	for (int k = 1; k < 4; ++k) if (m.logPi[k] > m.logPi[piMax]) piMax = k;
	for (;;) {
		int item = 0;
		if (lane == 0) item = atomicAdd(a.cursor, 1);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (item >= total) return;
		const int r = item / a.maxNSeed, s = item % a.maxNSeed;
		if (a.winStart[r] < 0 || s >= a.nSeeds[r]) continue;
		const int start = a.winStart[r], end = a.winEnd[r], W = end - start + 1;
		const int u = a.seedIds[(size_t) r * HU_MAX_SEED + s];
		const int v = t.parent[u];
		const int8_t* x = a.aligned + (size_t) r * t.csLen + start;
		const int8_t* su = t.seq + (size_t) u * t.csLen + start;
		const int8_t* sv = t.seq + (size_t) v * t.csLen + start;
		// ---- p-distances to the child and parent sequences (SeqUtils::pDist); counts < 2^16 packed in pairs ----
		unsigned int cu = 0, cv = 0; // (N << 16) | d
		for (int c = lane; c < W; c += 32) {
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
		const double w0 = t.blen[u];
		const double wur = w0 * ratio, wvr = w0 - wur;
		__syncwarp();
		if (lane < 2) modelPr(m, lane == 0 ? wur : wvr, ws.P[lane]);
		__syncwarp();
		const double* U = t.up + ((size_t) u * t.csLen + start) * 4;
		const double* V = t.dn + ((size_t) u * t.csLen + start) * 4;
		const double* US = t.upS + (size_t) u * t.csLen + start;
		const double* VS = t.dnS + (size_t) u * t.csLen + start;
		// ---- pass 1: the unweighted branch-length estimate ----
		int diff = 0;
		double sumScale = 0;
		for (int c = lane; c < W; c += 32) {
			double eu[4], ev[4], ya[4], yb[4], g[4];
			ld4(U + (size_t) c * 4, eu);
			ld4(V + (size_t) c * 4, ev);
			sumScale += __ldg(US + c) + __ldg(VS + c);
			matVec4(ws.P[0], eu, ya);
			matVec4(ws.P[1], ev, yb);
			for (int k = 0; k < 4; ++k) g[k] = ya[k] * yb[k];
			const int b = x[c];
			diff += argmax4(g) != (b >= 0 ? b : piMax);
		}
		for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
		const double wnr = (double) diff / (double) W;
		// ---- P(wnr) applied to the read's leaf vector: depends on the base only ----
		if (lane == 0) modelPr(m, wnr, ws.P[2]);
		__syncwarp();
		if (lane < 5) {
			double n[4], y[4];
			leafLin(m, lane < 4 ? lane : -1, n);
			matVec4(ws.P[2], n, y);
			for (int k = 0; k < 4; ++k) ws.Np[lane][k] = y[k];
		}
		__syncwarp();
		// ---- pass 2: treeLoglik(pi, UPr + VPr + NPr) ----
		double ll = 0;
		for (int c = lane; c < W; c += 32) {
			double eu[4], ev[4], ya[4], yb[4], z[4];
			ld4(U + (size_t) c * 4, eu);
			ld4(V + (size_t) c * 4, ev);
			matVec4(ws.P[0], eu, ya);
			matVec4(ws.P[1], ev, yb);
			const int b = x[c];
			const double* np = ws.Np[b >= 0 ? b : 4];
			for (int k = 0; k < 4; ++k) z[k] = (ya[k] * yb[k]) * np[k];
			ll += log(dotPi(m, z));
		}
		ll += sumScale;
		for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
		if (lane == 0) {
			hu_place_row row;
			row.start = start; row.end = end;
			row.cNode = u; row.pNode = v; row.aNode = ratio <= 0.5 ? u : v;
			row.taxon = t.taxon[row.aNode];
			row.dist = cDist;
			row.wuv = w0; row.ratio = ratio; row.wnr = wnr;
			row.loglik = ll; row.loglikTrue = 0; row.estLoglik = ll;
			row.height = 0; row.annoDist = 0; row.qPlace = 0; row.qTaxon = 0;
			row.outerIters = 0; row.nPlaces = 0;
			a.est[(size_t) r * HU_MAX_SEED + s] = row;
		}
	}
}

// filterPlacements: order by estimated loglik (descending; equal logliks keep seed order) and keep the
// placements within maxError of the best.
__global__ void filterKernel(int nReads, const int* __restrict__ winStart, const int* __restrict__ nSeeds,
		const hu_place_row* __restrict__ est, double maxError, int* __restrict__ order, int* __restrict__ nKept) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nReads) return;
	const int n = winStart[r] < 0 ? 0 : nSeeds[r];
	int* o = order + (size_t) r * HU_MAX_SEED;
	const hu_place_row* e = est + (size_t) r * HU_MAX_SEED;
	for (int i = 0; i < n; ++i) { // insertion sort, stable
		const double v = e[i].estLoglik;
		int p = i;
		while (p > 0 && e[o[p - 1]].estLoglik < v) { o[p] = o[p - 1]; p--; }
		o[p] = i;
	}
	int keep = 0;
	if (n > 0) {
		const double best = e[o[0]].estLoglik;
		while (keep < n && !(best - e[o[keep]].estLoglik > maxError)) keep++;
	}
	nKept[r] = keep;
}

int launchEstimate(hu_batch* b, const hu_sep_params* p) {
	hu_db* db = b->db;
	EstArgs a;
	a.nReads = b->nReads; a.maxNSeed = p->maxNSeed; a.pad0 = a.pad1 = 0;
	a.aligned = b->vit.aligned; a.winStart = b->winStart; a.winEnd = b->winEnd;
	a.seedIds = b->seed.seedIds; a.nSeeds = b->seed.nSeeds;
	a.est = b->est; a.cursor = b->workCount;
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	HU_CUDA(cudaMemsetAsync(b->workCount, 0, 4 * sizeof(int), b->stream));
	const int grid = sms * 4;
	estimateKernel<<<grid, HU_EST_WARPS * 32, 0, b->stream>>>(a, db->tree, db->model);
	filterKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, b->winStart, a.nSeeds, b->est, p->maxError, b->order, b->nKept);
	HU_CUDA(cudaEventRecord(b->ev[5], b->stream));
	b->launches += 2;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
