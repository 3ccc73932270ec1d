// K3: the "Estimate" stage of SEP -- PTUnrooted::estimateSeq for every (read, seed)
// (src/PhyloTreeUnrooted.cpp:849-877) followed by filterPlacements (src/HmmUFOtu_main.cpp:162-173).
//
// One CTA per (read, seed), persistent over the item space.  The two 4 x W message windows U = msg(u->v) and
// V = msg(v->u) are streamed once from HBM with 128-bit loads (32 contiguous bytes per column); the
// per-column R_j = log(P(wur) e^U_j) + log(P(wvr) e^V_j) is parked in shared memory between the
// branch-length pass (wnr = fraction of columns whose arg-max state differs from the read's) and the
// log-likelihood pass, so every exp/log is evaluated once.
#include "hu_batch.cuh"

namespace hu {

#define HU_EST_THREADS 128
#define HU_EST_CAPW 1536 // columns of g kept in shared memory (48 KB); wider windows recompute g in pass 2

struct EstArgs {
	int nReads, maxNSeed, capW, pad;
	const int8_t* aligned;
	const int* winStart;
	const int* winEnd;
	const int* seedIds;
	const int* nSeeds;
	hu_place_row* est;
	int* cursor;
};

template <typename T>
__device__ __forceinline__ T blockSum(T v, T* red) { // red: HU_EST_THREADS/32 + 1 slots
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	T s = red[0];
	for (int w = 1; w < HU_EST_THREADS / 32; ++w) s += red[w];
	return s;
}

__global__ void __launch_bounds__(HU_EST_THREADS) estimateKernel(EstArgs a, TreeDev t, Model m) {
	extern __shared__ __align__(16) double sR[]; // [capW][4]
	__shared__ double sP[3][16];
	__shared__ double sNp[5][4];
	__shared__ double redD[HU_EST_THREADS / 32];
	__shared__ int redI[HU_EST_THREADS / 32];
	__shared__ int sItem;
	const int tid = threadIdx.x;
	const int total = a.nReads * a.maxNSeed;
	int piMax = 0; // argmax of log(pi): state of a gap column of the read
	for (int k = 1; k < 4; ++k) if (m.logPi[k] > m.logPi[piMax]) piMax = k;
	for (;;) {
		__syncthreads();
		if (tid == 0) sItem = atomicAdd(a.cursor, 1);
		__syncthreads();
		const int item = sItem;
		if (item >= total) return;
		const int r = item / a.maxNSeed, s = item % a.maxNSeed;
		if (a.winStart[r] < 0 || s >= a.nSeeds[r]) continue;
		const int start = a.winStart[r], end = a.winEnd[r], W = end - start + 1;
		const int u = a.seedIds[(size_t) r * HU_MAX_SEED + s];
		const int v = t.parent[u];
		const int8_t* x = a.aligned + (size_t) r * t.csLen;
		const int8_t* su = t.seq + (size_t) u * t.csLen;
		const int8_t* sv = t.seq + (size_t) v * t.csLen;
		// ---- p-distances to the child and parent sequences (SeqUtils::pDist) ----
		int du = 0, nu = 0, dv = 0, nv = 0;
		for (int j = start + tid; j <= end; j += HU_EST_THREADS) {
			const int b = x[j];
			if (b >= 0) {
				const int cu = su[j], cv = sv[j];
				if (cu >= 0) { nu++; du += cu != b; }
				if (cv >= 0) { nv++; dv += cv != b; }
			}
		}
		du = blockSum(du, redI); nu = blockSum(nu, redI); dv = blockSum(dv, redI); nv = blockSum(nv, redI);
		const double cDist = (double) du / (double) nu;
		const double pDist = (double) dv / (double) nv;
		double ratio = cDist / (cDist + pDist);
		if (isnan(ratio)) ratio = 0.5;
		const double w0 = t.blen[u];
		const double wur = w0 * ratio, wvr = w0 - wur;
		if (tid == 0) modelPr(m, wur, sP[0]);
		if (tid == 32) modelPr(m, wvr, sP[1]);
		__syncthreads();
		const double* U = t.up + ((size_t) u * t.csLen + start) * 4;
		const double* V = t.dn + ((size_t) u * t.csLen + start) * 4;
		const double* US = t.upS + (size_t) u * t.csLen + start;
		const double* VS = t.dnS + (size_t) u * t.csLen + start;
		const bool store = W <= a.capW;
		// ---- pass 1: g_j = (P(wur) e^U_j) * (P(wvr) e^V_j) (= e^{R_j} up to the column scale) and the
		//      unweighted branch-length estimate (cpp:1018-1032).  Linear domain, see expNorm().
		int diff = 0;
		double sumScale = 0;
		for (int c = tid; c < W; c += HU_EST_THREADS) {
			double eu[4], ev[4], ya[4], yb[4], g[4];
			ld4(U + (size_t) c * 4, eu);
			ld4(V + (size_t) c * 4, ev);
			sumScale += __ldg(US + c) + __ldg(VS + c);
			matVec4(sP[0], eu, ya);
			matVec4(sP[1], ev, yb);
			for (int k = 0; k < 4; ++k) g[k] = ya[k] * yb[k];
			if (store) {
				reinterpret_cast<double2*>(sR + (size_t) c * 4)[0] = make_double2(g[0], g[1]);
				reinterpret_cast<double2*>(sR + (size_t) c * 4)[1] = make_double2(g[2], g[3]);
			}
			const int b = x[start + c];
			diff += argmax4(g) != (b >= 0 ? b : piMax);
		}
		diff = blockSum(diff, redI);
		const double wnr = (double) diff / (double) W;
		// ---- P(wnr) applied to the read's leaf vector: depends on the base only ----
		if (tid == 0) modelPr(m, wnr, sP[2]);
		__syncthreads();
		if (tid < 5) {
			double n[4], y[4];
			leafLin(m, tid < 4 ? tid : -1, n);
			matVec4(sP[2], n, y);
			for (int k = 0; k < 4; ++k) sNp[tid][k] = y[k];
		}
		__syncthreads();
		// ---- pass 2: treeLoglik(pi, UPr + VPr + NPr) (cpp:872-874) = sum_j log(pi . (g_j * n_j)) + scales ----
		double ll = 0;
		for (int c = tid; c < W; c += HU_EST_THREADS) {
			double g[4];
			if (store) {
				const double2 p0 = reinterpret_cast<const double2*>(sR + (size_t) c * 4)[0];
				const double2 p1 = reinterpret_cast<const double2*>(sR + (size_t) c * 4)[1];
				g[0] = p0.x; g[1] = p0.y; g[2] = p1.x; g[3] = p1.y;
			} else {
				double eu[4], ev[4], ya[4], yb[4];
				ld4(U + (size_t) c * 4, eu);
				ld4(V + (size_t) c * 4, ev);
				matVec4(sP[0], eu, ya);
				matVec4(sP[1], ev, yb);
				for (int k = 0; k < 4; ++k) g[k] = ya[k] * yb[k];
			}
			const int b = x[start + c];
			const double* np = sNp[b >= 0 ? b : 4];
			double z[4];
			for (int k = 0; k < 4; ++k) z[k] = g[k] * np[k];
			ll += log(dotPi(m, z));
		}
		ll = blockSum(ll + sumScale, redD);
		if (tid == 0) {
			hu_place_row row;
			row.start = start; row.end = end;
			row.cNode = u; row.pNode = v; row.aNode = ratio <= 0.5 ? u : v;
			row.taxon = t.taxon[row.aNode];
			row.wuv = w0; row.ratio = ratio; row.wnr = wnr;
			row.loglik = ll; row.loglikTrue = 0; row.estLoglik = ll;
			row.height = 0; row.annoDist = 0; row.qPlace = 0; row.qTaxon = 0;
			row.outerIters = 0; row.nPlaces = 0;
			row.dist = cDist;
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
	a.nReads = b->nReads; a.maxNSeed = p->maxNSeed; a.capW = HU_EST_CAPW; a.pad = 0;
	a.aligned = b->vit.aligned; a.winStart = b->winStart; a.winEnd = b->winEnd;
	a.seedIds = b->seed.seedIds; a.nSeeds = b->seed.nSeeds;
	a.est = b->est; a.cursor = b->workCount;
	const size_t smem = (size_t) HU_EST_CAPW * 4 * sizeof(double);
	static bool attr = false;
	if (!attr) {
		HU_CUDA(cudaFuncSetAttribute(estimateKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		attr = true;
	}
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	HU_CUDA(cudaMemsetAsync(b->workCount, 0, 4 * sizeof(int), b->stream));
	const int grid = std::min(sms * 4, b->nReads * p->maxNSeed);
	estimateKernel<<<grid, HU_EST_THREADS, smem, b->stream>>>(a, db->tree, db->model);
	filterKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, b->winStart, a.nSeeds, b->est, p->maxError, b->order, b->nKept);
	HU_CUDA(cudaEventRecord(b->ev[5], b->stream));
	b->launches += 2;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
