// C ABI: batch workspace, host-side band planning, staging, stage drivers and result read-back.
#include "hu_batch.cuh"
#include "hu_stdsort.cuh"
#include <string.h>
#include <ctype.h>
#include <utility>

namespace hu {

#define HU_TRY(x) do { int rc_ = (x); if (rc_ != HU_OK) return rc_; } while (0)

template <typename T>
static int balloc(hu_batch* b, T** p, size_t n) {
	void* q = NULL;
	cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
	if (e != cudaSuccess) {
		setError("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
		return HU_ERR_NOMEM;
	}
	b->allocs.push_back(q);
	*p = (T*) q;
	return HU_OK;
}
template <typename T>
static int palloc(hu_batch* b, T** p, size_t n) {
	void* q = NULL;
	cudaError_t e = cudaMallocHost(&q, std::max<size_t>(n, 1) * sizeof(T));
	if (e != cudaSuccess) {
		setError("cudaMallocHost of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
		return HU_ERR_NOMEM;
	}
	b->pinned.push_back(q);
	*p = (T*) q;
	return HU_OK;
}

// Band of one sequence as a chain of boxes (src/BandedHMMP7.cpp:794-881).  Returns false when the
// ViterbiAlignPaths do not chain monotonically (the reference would sweep overlapping boxes in place).
static bool planBand(int N, int K, const hu_vpath* vp, int nPaths, VPlan* plan) {
	const double kMinGapFrac = 0.2; // src/BandedHMMP7.h
	plan->nReg = 0;
	for (int p = 0; p < nPaths; ++p) {
		const hu_vpath& v = vp[p];
		if (!(v.start >= 1 && v.start <= v.end && v.end <= K && v.from >= 1 && v.from <= v.to && v.to <= N && v.nIns >= 0 && v.nDel >= 0)) return false;
		if (p > 0 && !(v.start >= vp[p - 1].end && v.from >= vp[p - 1].to)) return false;
	}
	int n = 0;
	for (int p = 0; p < nPaths; ++p) {
		const hu_vpath& v = vp[p];
		int upQLen = p == 0 ? v.from - 1 : v.from - vp[p - 1].to;
		if (upQLen < 0) upQLen = 0;
		int up_start = p == 0 ? (int) (v.start - upQLen * (1 + kMinGapFrac)) : vp[p - 1].end;
		if (up_start < 1) up_start = 1;
		int up_from = p == 0 ? (int) (v.from - upQLen * (1 + kMinGapFrac)) : vp[p - 1].to;
		if (up_from < 1) up_from = 1;
		VRegion r;
		memset(&r, 0, sizeof r);
		r.j0 = up_start; r.j1 = v.start; r.i0 = up_from; r.i1 = v.from; r.useB = 1;
		plan->reg[n++] = r;
		memset(&r, 0, sizeof r);
		r.j0 = v.start; r.j1 = v.end; r.i0 = v.from; r.i1 = v.to; r.useB = 1;
		r.strip = 1; r.pStart = v.start; r.pFrom = v.from; r.nIns = v.nIns; r.nDel = v.nDel;
		plan->reg[n++] = r;
	}
	const hu_vpath& last = vp[nPaths - 1];
	const int downQLen = N - last.to;
	int down_end = (int) (last.end + downQLen * (1 + kMinGapFrac));
	int down_to = (int) (last.to + downQLen * (1 + kMinGapFrac));
	if (down_end > K) down_end = K;
	if (down_to > N) down_to = N;
	VRegion r;
	memset(&r, 0, sizeof r);
	r.j0 = last.end; r.j1 = down_end; r.i0 = last.to; r.i1 = down_to; r.useB = 0;
	plan->reg[n++] = r;
	plan->nReg = n;
	return true;
}

__global__ void windowsKernel(int n, const hu_align_row* __restrict__ merged, int* __restrict__ winStart, int* __restrict__ winEnd) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const hu_align_row a = merged[r];
	const bool bad = a.flags & (HU_ALIGN_INVALID | HU_ALIGN_CHIMERA);
	winStart[r] = bad ? -1 : a.csStart - 1; // src/hmmufotu.cpp:645
	winEnd[r] = bad ? -1 : a.csEnd - 1;
}

// ---- traces: the device keeps them end-to-start in fixed slots; they are handed back packed and start-to-end ----
// offsets = exclusive scan of (traceLen + 1) over the sequences (one CTA; nSeq is a few thousand)
__global__ void traceScanKernel(int nSeq, const hu_align_row* __restrict__ rows, long long* __restrict__ offs) {
	__shared__ long long part[1024];
	const int tid = threadIdx.x, per = (nSeq + blockDim.x - 1) / blockDim.x;
	const int s0 = min(nSeq, tid * per), s1 = min(nSeq, s0 + per);
	long long sum = 0;
	for (int s = s0; s < s1; ++s) sum += rows[s].traceLen + 1;
	part[tid] = sum;
	__syncthreads();
	for (int o = 1; o < blockDim.x; o <<= 1) {
		const long long v = tid >= o ? part[tid - o] : 0;
		__syncthreads();
		part[tid] += v;
		__syncthreads();
	}
	long long run = part[tid] - sum;
	for (int s = s0; s < s1; ++s) { offs[s] = run; run += rows[s].traceLen + 1; }
	if (tid == blockDim.x - 1) offs[nSeq] = part[tid];
}
__global__ void tracePackKernel(int nSeq, const hu_align_row* __restrict__ rows, const VPlan* __restrict__ plans,
		const char* __restrict__ slots, const long long* __restrict__ offs, char* __restrict__ packed) {
	const int s = blockIdx.x;
	const int n = rows[s].traceLen;
	const char* src = slots + plans[s].traceOff;
	char* dst = packed + offs[s];
	for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[n - 1 - i];
	if (threadIdx.x == 0) dst[n] = 0;
}

// fixed-size assignment rows of the read-sharding driver (hmmufotu_b200/driver.py ROW_DT, 96 bytes), built on the device so
// that the gather of the rows (the one collective of the path) can run on device buffers
struct PackedRow {
	long long read;
	int seqStart, seqEnd, hmmStart, hmmEnd, csStart, csEnd, flags, nPlaces;
	double cost;
	int cNode, pNode, aNode, taxon;
	double ratio, annoDist, loglik, qPlace, qTaxon;
};
static_assert(sizeof(PackedRow) == 104, "driver.ROW_DT");
__global__ void packRowsKernel(int n, long long read0, const hu_align_row* __restrict__ merged, const hu_place_row* __restrict__ best,
		PackedRow* __restrict__ out) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const hu_align_row a = merged[r];
	const hu_place_row b = best[r];
	PackedRow o;
	o.read = read0 + r;
	o.seqStart = a.seqStart; o.seqEnd = a.seqEnd; o.hmmStart = a.hmmStart; o.hmmEnd = a.hmmEnd; o.csStart = a.csStart; o.csEnd = a.csEnd;
	o.flags = a.flags; o.nPlaces = b.nPlaces; o.cost = a.cost;
	o.cNode = b.cNode; o.pNode = b.pNode; o.aNode = b.aNode; o.taxon = b.taxon;
	o.ratio = b.ratio; o.annoDist = b.annoDist; o.loglik = b.loglik; o.qPlace = b.qPlace; o.qTaxon = b.qTaxon;
	out[r] = o;
}

} // namespace hu

using namespace hu;

extern "C" {

void* hu_host_alloc(size_t bytes) {
	void* q = NULL;
	if (cudaMallocHost(&q, bytes ? bytes : 1) != cudaSuccess) {
		setError("cudaMallocHost of %zu bytes failed", bytes);
		cudaGetLastError();
		return NULL;
	}
	return q;
}
void hu_host_free(void* p) {
	if (p) cudaFreeHost(p);
}

void hu_sep_params_default(hu_sep_params* p) {
	p->maxNSeed = 50;
	p->useTrueLoglik = 0;
	p->maxDiff = INFINITY;
	p->maxHeight = INFINITY;
	p->maxError = 20;
	p->equalSortPerm = NULL;
	p->seedOrder = HU_SEED_ORDER_STDSORT;
	p->estMethod = HU_EST_UNWEIGHTED;
	p->prior = HU_PRIOR_UNIFORM;
	p->onlyML = 0;
}

int hu_batch_create(hu_batch** out, hu_db* db, int32_t maxReads, int32_t maxSeqLen, int32_t paired) {
	if (!out || !db || maxReads < 1 || maxSeqLen < 1) { setError("hu_batch_create: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(db->device));
	if (std::min((long long) db->hmm.L, (long long) maxSeqLen * (paired ? 2 : 1)) > HU_SEED_MAX_N) {
		setError("hu_batch_create: reads of %d bases over %d consensus columns exceed the %d compared columns for which the seed keys are exact",
		         maxSeqLen * (paired ? 2 : 1), db->hmm.L, HU_SEED_MAX_N);
		return HU_ERR_ARG;
	}
	hu_batch* b = new hu_batch();
	memset(&b->vit, 0, sizeof b->vit);
	memset(&b->seed, 0, sizeof b->seed);
	memset(&b->stats, 0, sizeof b->stats);
	b->db = db; b->maxReads = maxReads; b->maxSeqLen = maxSeqLen; b->paired = paired ? 1 : 0;
	b->maxSeq = maxReads * (paired ? 2 : 1);
	b->nReads = b->nSeq = 0; b->launches = 0;
	b->dPtrs = NULL; b->ptrsCap = 0; b->placeSlab = NULL; b->litIds = NULL; b->litBits = NULL;
	b->posOfSlotH = b->nodeOfPosH = NULL; b->nCandH = 0; b->heightCached = 0;
	b->estSlab = NULL; b->estSlabCtas = 0;
	*out = b;
	HU_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	for (int i = 0; i < 16; ++i) HU_CUDA(cudaEventCreate(&b->ev[i]));
	const int K = db->hmm.K, L = db->hmm.L;
	const TreeDev& t = db->tree;
	b->codesCap = (long long) b->maxSeq * maxSeqLen;
	HU_TRY(balloc(b, &b->dCodes, (size_t) b->codesCap));
	HU_TRY(balloc(b, &b->dPlans, (size_t) b->maxSeq));
	b->tracesCap = (long long) b->maxSeq * (maxSeqLen + K + 4);
	HU_TRY(balloc(b, &b->dTraces, (size_t) b->tracesCap));
	HU_TRY(balloc(b, &b->dTracesPacked, (size_t) b->tracesCap));
	HU_TRY(balloc(b, &b->dTraceOffs, (size_t) b->maxSeq + 1));
	HU_TRY(palloc(b, &b->hTraceOffs, (size_t) b->maxSeq + 1));
	HU_TRY(balloc(b, &b->vit.aligned, (size_t) b->maxSeq * L + 64)); // + slack for 16-byte rounded bulk copies
	HU_TRY(balloc(b, &b->workItems, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->estItems, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->estItemsP, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->vit.rows, (size_t) b->maxSeq));
	HU_TRY(balloc(b, &b->merged, (size_t) maxReads));
	HU_TRY(balloc(b, &b->vit.fullCount, 4));
	b->vit.fullCursor = b->vit.fullCount + 1;
	b->vit.bandCursor = b->vit.fullCount + 2;
	HU_TRY(balloc(b, &b->vit.fullList, (size_t) b->maxSeq));
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	// Viterbi CTA shape: the number of warps (= sequences in flight) per CTA that puts most warps on an SM, given the per-warp
	// shared memory (boundary column, emission tables, codes: grows with maxSeqLen) and the kernel's 96 registers per thread
	const size_t perWarp = ((size_t) (maxSeqLen + 15) / 16 * 16 + (size_t) 3 * (maxSeqLen + 2 + 32) * 8 + 320 * 8 + 8 * 24 + 5 * sizeof(VRegion) + 16 +
	                        ((size_t) 2 * (maxSeqLen + 64) + 15) / 16 * 16 + 15) / 16 * 16;
	if (perWarp > 200 * 1024) { setError("maxSeqLen %d too long for the Viterbi shared-memory layout", maxSeqLen); return HU_ERR_ARG; }
	int wpb = 1, bestWarps = 0, bestCtas = 1;
	for (int w = 1; w <= HU_VIT_WARPS; ++w) {
		if (perWarp * w > 200 * 1024) break;
		const int bySmem = (int) ((227 * 1024) / (perWarp * w + 1024)), byRegs = 65536 / (96 * 32 * w);
		const int ctas = std::min(bySmem, byRegs);
		if (ctas * w >= bestWarps) { bestWarps = ctas * w; wpb = w; bestCtas = ctas; }
	}
	b->vit.warpsPerBlock = wpb;
	b->vitGrid = sms * std::max(1, bestCtas);
	b->vit.maxLen = maxSeqLen;
	b->fullSlots = std::min(sms * 2, b->maxSeq); // CTAs of the full-matrix pass, one back-pointer slab each
	b->vit.fullSlabWords = (size_t) ((K + 31) / 32) * (maxSeqLen + 31) * 4; // [stripe][step][4 ballot words]
	HU_TRY(balloc(b, &b->vit.fullPtrs, b->vit.fullSlabWords * b->fullSlots));
	HU_TRY(balloc(b, &b->winStart, (size_t) maxReads));
	HU_TRY(balloc(b, &b->winEnd, (size_t) maxReads));
	const int nReadsPad = (maxReads + HU_SCAN_READS - 1) / HU_SCAN_READS * HU_SCAN_READS;
	HU_TRY(balloc(b, &b->seed.rplanes, (size_t) t.nPWords * nReadsPad * 4));
	HU_TRY(balloc(b, &b->seed.rValid, (size_t) nReadsPad));
	HU_TRY(balloc(b, &b->seed.rWordLo, (size_t) nReadsPad));
	HU_TRY(balloc(b, &b->seed.rWordHi, (size_t) nReadsPad));
	HU_TRY(balloc(b, &b->seed.nSparse, (size_t) nReadsPad));
	b->seed.sparseCap = maxSeqLen * (paired ? 2 : 1); // a read cannot have more bases in non-match columns than bases
	HU_TRY(balloc(b, &b->seed.sparse, (size_t) maxReads * b->seed.sparseCap));
	HU_TRY(balloc(b, &b->seed.keys, (size_t) maxReads * t.scanPitch));
	HU_CUDA(cudaMemset(b->seed.keys, 0xff, (size_t) maxReads * t.scanPitch * sizeof(uint32_t))); // tail positions stay excluded
	b->mmScanRan = 0; b->mmB = NULL; b->mmRange = NULL; b->mmMode = b->mmTileExt = b->mmCols = b->mmColReads = NULL;
	if (t.mmA) {
		const int numNt = (maxReads + HU_MM_BN - 1) / HU_MM_BN;
		HU_TRY(balloc(b, &b->mmB, (size_t) numNt * (t.mmChunks + 1) * HU_MM_B_BYTES));
		HU_TRY(balloc(b, &b->mmRange, (size_t) numNt));
		HU_TRY(balloc(b, &b->mmMode, (size_t) nReadsPad));
		HU_TRY(balloc(b, &b->mmTileExt, (size_t) numNt));
		HU_TRY(balloc(b, &b->mmCols, (size_t) 3 * (t.csLen + 1) + (size_t) (t.scanPitch + 3) / 4)); // + one byte per candidate position (mmPosExclKernel)
		HU_TRY(balloc(b, &b->mmColReads, (size_t) maxReads * 64));
	}
	HU_TRY(balloc(b, &b->seed.seedIds, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->seed.nSeeds, (size_t) maxReads));
	HU_TRY(balloc(b, &b->est, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->places, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->order, (size_t) maxReads * HU_MAX_SEED));
	HU_TRY(balloc(b, &b->nKept, (size_t) maxReads));
	HU_TRY(balloc(b, &b->best, (size_t) maxReads));
	HU_TRY(balloc(b, &b->workCount, 8));
	HU_CUDA(cudaMemset(b->workCount, 0, 8 * sizeof(int)));
	HU_CUDA(cudaMemset(b->seed.nSeeds, 0, (size_t) maxReads * sizeof(int)));
	HU_CUDA(cudaMemset(b->nKept, 0, (size_t) maxReads * sizeof(int)));
	HU_TRY(palloc(b, &b->hPlansPinned, (size_t) b->maxSeq));
	return HU_OK;
}

int hu_batch_destroy(hu_batch* b) {
	if (!b) return HU_OK;
	cudaSetDevice(b->db->device);
	cudaStreamSynchronize(b->stream);
	for (size_t i = 0; i < b->allocs.size(); ++i) cudaFree(b->allocs[i]);
	for (size_t i = 0; i < b->pinned.size(); ++i) cudaFreeHost(b->pinned[i]);
	if (b->dPtrs) cudaFree(b->dPtrs);
	for (int i = 0; i < 16; ++i) cudaEventDestroy(b->ev[i]);
	cudaStreamDestroy(b->stream);
	delete b;
	return HU_OK;
}

int hu_batch_upload(hu_batch* b, const hu_read_batch* rb) {
	if (!b || !rb || rb->nReads < 1 || rb->nReads > b->maxReads) { setError("hu_batch_upload: bad batch size"); return HU_ERR_ARG; }
	if ((rb->paired != 0) != (b->paired != 0)) { setError("hu_batch_upload: pairedness differs from the workspace"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	const int nSeq = rb->nReads * (rb->paired ? 2 : 1);
	const int K = b->db->hmm.K;
	const long long nCodes = rb->offsets[nSeq] - rb->offsets[0];
	if (nCodes > b->codesCap) { setError("hu_batch_upload: %lld bases exceed the workspace capacity %lld", nCodes, b->codesCap); return HU_ERR_ARG; }
	long long ptrWords = 0, traceOff = 0;
	VPlan* plans = b->hPlansPinned;
	b->hTraceOff.resize(nSeq + 1);
	for (int s = 0; s < nSeq; ++s) {
		VPlan& p = plans[s];
		memset(&p, 0, sizeof p);
		const long long len = rb->offsets[s + 1] - rb->offsets[s];
		if (len < 1 || len > b->maxSeqLen) { setError("hu_batch_upload: sequence %d has length %lld (workspace maxSeqLen %d)", s, len, b->maxSeqLen); return HU_ERR_ARG; }
		p.N = (int) len;
		p.codeOff = rb->offsets[s] - rb->offsets[0];
		p.traceOff = traceOff;
		b->hTraceOff[s] = traceOff;
		traceOff += len + K + 4;
		const int np = rb->nPaths ? rb->nPaths[s] : 0;
		if (np < 0 || np > 2) { setError("hu_batch_upload: nPaths[%d] = %d", s, np); return HU_ERR_ARG; }
		if (np == 0) p.flags = HU_ALIGN_NOSEED;
		else if (!planBand(p.N, K, rb->vpaths + (size_t) s * 2, np, &p)) { p.nReg = 0; p.flags = HU_ALIGN_NOSEED | HU_ALIGN_IRREGULAR; }
		for (int r = 0; r < p.nReg; ++r) {
			VRegion& R = p.reg[r];
			R.rowWords = R.i1 - R.i0 + 1 + 31; // wavefront steps per 32-column stripe (hu_viterbi.cu: four ballot words per step)
			R.ptrOff = ptrWords;
			ptrWords += (long long) ((R.j1 - R.j0 + 1 + 31) / 32) * R.rowWords * 4;
		}
	}
	b->hTraceOff[nSeq] = traceOff;
	if (ptrWords > b->ptrsCap) { // grow the back-pointer buffer (synchronous; only when a batch needs more than any before)
		HU_CUDA(cudaStreamSynchronize(b->stream));
		if (b->dPtrs) cudaFree(b->dPtrs);
		b->ptrsCap = ptrWords + ptrWords / 4 + 1024;
		void* q = NULL;
		cudaError_t e = cudaMalloc(&q, (size_t) b->ptrsCap * sizeof(uint32_t));
		if (e != cudaSuccess) { b->dPtrs = NULL; b->ptrsCap = 0; setError("cudaMalloc of the back-pointer buffer (%lld words) failed", ptrWords); return HU_ERR_NOMEM; }
		b->dPtrs = (uint32_t*) q;
	}
	b->nReads = rb->nReads; b->nSeq = nSeq;
	HU_CUDA(cudaMemcpyAsync(b->dCodes, rb->codes + rb->offsets[0], (size_t) nCodes, cudaMemcpyHostToDevice, b->stream));
	HU_CUDA(cudaMemcpyAsync(b->dPlans, plans, (size_t) nSeq * sizeof(VPlan), cudaMemcpyHostToDevice, b->stream));
	b->vit.plans = b->dPlans; b->vit.codes = b->dCodes; b->vit.ptrs = b->dPtrs; b->vit.traces = b->dTraces;
	b->launches = 0;
	b->h2dBytes = nCodes + (long long) nSeq * sizeof(VPlan);
	return HU_OK;
}

int hu_batch_set_aligned(hu_batch* b, int32_t n, const int8_t* codes, const int32_t* start, const int32_t* end) {
	if (!b || n < 1 || n > b->maxReads || !codes || !start || !end) { setError("hu_batch_set_aligned: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	const int L = b->db->hmm.L;
	b->nReads = b->nSeq = n;
	HU_CUDA(cudaMemcpyAsync(b->vit.aligned, codes, (size_t) n * L, cudaMemcpyHostToDevice, b->stream));
	HU_CUDA(cudaMemcpyAsync(b->winStart, start, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, b->stream));
	HU_CUDA(cudaMemcpyAsync(b->winEnd, end, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, b->stream));
	b->launches = 0;
	b->h2dBytes = (long long) n * L + 8LL * n;
	return HU_OK;
}

// ---- stage inputs supplied by the caller: what the reference's free functions take as arguments when they are called one
// by one (src/HmmUFOtu_main.h:70-113); host/hu_facade.cpp builds those functions on these ----
int hu_batch_set_seeds(hu_batch* b, const int32_t* nSeeds, const int32_t* ids) {
	if (!b || !nSeeds || !ids || b->nReads < 1) { setError("hu_batch_set_seeds: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	const int mx = b->seed.maxNSeed > 0 ? b->seed.maxNSeed : HU_MAX_SEED;
	std::vector<int> all((size_t) b->nReads * HU_MAX_SEED, 0);
	for (int r = 0; r < b->nReads; ++r) {
		if (nSeeds[r] < 0 || nSeeds[r] > HU_MAX_SEED) { setError("hu_batch_set_seeds: %d seeds for read %d", nSeeds[r], r); return HU_ERR_ARG; }
		for (int s = 0; s < nSeeds[r]; ++s) {
			const int id = ids[(size_t) r * mx + s];
			if (id < 0 || id >= b->db->tree.nNodes || id == b->db->tree.root) { setError("hu_batch_set_seeds: node %d is not a seed candidate", id); return HU_ERR_ARG; }
			all[(size_t) r * HU_MAX_SEED + s] = id;
		}
	}
	HU_CUDA(cudaMemcpyAsync(b->seed.nSeeds, nSeeds, (size_t) b->nReads * sizeof(int), cudaMemcpyHostToDevice, b->stream));
	HU_CUDA(cudaMemcpyAsync(b->seed.seedIds, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}
static int setRows(hu_batch* b, hu_place_row* dst, const int32_t* n, const hu_place_row* rows, bool identityOrder) {
	HU_CUDA(cudaSetDevice(b->db->device));
	const int mx = b->seed.maxNSeed > 0 ? b->seed.maxNSeed : HU_MAX_SEED;
	std::vector<hu_place_row> all((size_t) b->nReads * HU_MAX_SEED);
	std::vector<int> ord((size_t) b->nReads * HU_MAX_SEED, 0);
	memset(all.data(), 0, all.size() * sizeof(hu_place_row));
	for (int r = 0; r < b->nReads; ++r) {
		if (n[r] < 0 || n[r] > HU_MAX_SEED) { setError("%d rows for read %d", n[r], r); return HU_ERR_ARG; }
		for (int i = 0; i < n[r]; ++i) { all[(size_t) r * HU_MAX_SEED + i] = rows[(size_t) r * mx + i]; ord[(size_t) r * HU_MAX_SEED + i] = i; }
	}
	HU_CUDA(cudaMemcpyAsync(dst, all.data(), all.size() * sizeof(hu_place_row), cudaMemcpyHostToDevice, b->stream));
	if (identityOrder) {
		HU_CUDA(cudaMemcpyAsync(b->order, ord.data(), ord.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
		HU_CUDA(cudaMemcpyAsync(b->nKept, n, (size_t) b->nReads * sizeof(int), cudaMemcpyHostToDevice, b->stream));
	}
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}
int hu_batch_set_estimates(hu_batch* b, const int32_t* n, const hu_place_row* rows) {
	if (!b || !n || !rows || b->nReads < 1) { setError("hu_batch_set_estimates: bad arguments"); return HU_ERR_ARG; }
	// rows in the caller's order become the estimate vector; they count as kept, in that order, until hu_filter_batch runs
	int rc = setRows(b, b->est, n, rows, true);
	if (rc == HU_OK) HU_CUDA(cudaMemcpy(b->seed.nSeeds, n, (size_t) b->nReads * sizeof(int), cudaMemcpyHostToDevice));
	return rc;
}
int hu_batch_set_places(hu_batch* b, const int32_t* n, const hu_place_row* rows) {
	if (!b || !n || !rows || b->nReads < 1) { setError("hu_batch_set_places: bad arguments"); return HU_ERR_ARG; }
	int rc = setRows(b, b->places, n, rows, false);
	if (rc == HU_OK) HU_CUDA(cudaMemcpy(b->nKept, n, (size_t) b->nReads * sizeof(int), cudaMemcpyHostToDevice));
	return rc;
}

int hu_align_batch(hu_batch* b, int mode) {
	if (!b || b->nSeq < 1 || !b->vit.plans) { setError("hu_align_batch: no reads uploaded"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	b->launches = 0;
	HU_CUDA(cudaEventRecord(b->ev[0], b->stream));
	HU_TRY(launchViterbi(b, mode));
	windowsKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, b->merged, b->winStart, b->winEnd);
	b->launches += 1;
	HU_CUDA(cudaEventRecord(b->ev[1], b->stream));
	return HU_OK;
}

int hu_seed_batch(hu_batch* b, const hu_sep_params* p) {
	if (!b || !p || b->nReads < 1) { setError("hu_seed_batch: empty batch"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	return launchSeed(b, p);
}
int hu_estimate_batch(hu_batch* b, const hu_sep_params* p) {
	if (!b || !p || b->nReads < 1) { setError("hu_estimate_batch: empty batch"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	return launchEstimate(b, p);
}
int hu_filter_batch(hu_batch* b, const hu_sep_params* p) {
	if (!b || !p || b->nReads < 1) { setError("hu_filter_batch: empty batch"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	return launchFilter(b, p);
}
int hu_place_batch(hu_batch* b, const hu_sep_params* p) {
	if (!b || !p || b->nReads < 1) { setError("hu_place_batch: empty batch"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	return launchPlace(b, p);
}
int hu_finalize_batch(hu_batch* b, const hu_sep_params* p) {
	if (!b || !p || b->nReads < 1) { setError("hu_finalize_batch: empty batch"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(b->db->device));
	return launchFinalize(b, p);
}
int hu_assign_batch(hu_batch* b, int mode, const hu_sep_params* p) {
	HU_TRY(hu_align_batch(b, mode));
	HU_TRY(hu_seed_batch(b, p));
	HU_TRY(hu_estimate_batch(b, p));
	HU_TRY(hu_place_batch(b, p));
	return hu_finalize_batch(b, p);
}
int hu_batch_sync(hu_batch* b) {
	if (!b) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}

int hu_batch_get_align(hu_batch* b, hu_align_row* rows, char* traces, int64_t traceCap, int64_t* traceOff) {
	if (!b || !rows) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	const bool wantTraces = traces && traceOff && b->vit.plans;
	if (wantTraces) {
		// reversed and packed on the device: the copy below moves sum(traceLen + 1) bytes, not the fixed slots
		traceScanKernel<<<1, 1024, 0, b->stream>>>(b->nSeq, b->vit.rows, b->dTraceOffs);
		tracePackKernel<<<b->nSeq, 128, 0, b->stream>>>(b->nSeq, b->vit.rows, b->dPlans, b->dTraces, b->dTraceOffs, b->dTracesPacked);
		b->launches += 2;
		HU_CUDA(cudaMemcpyAsync(b->hTraceOffs, b->dTraceOffs, (size_t) (b->nSeq + 1) * sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
	}
	HU_CUDA(cudaMemcpyAsync(rows, b->vit.rows, (size_t) b->nSeq * sizeof(hu_align_row), cudaMemcpyDeviceToHost, b->stream));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	if (wantTraces) {
		const long long total = b->hTraceOffs[b->nSeq];
		if (total > traceCap) { setError("hu_batch_get_align: trace buffer too small (%lld bytes needed)", total); return HU_ERR_ARG; }
		HU_CUDA(cudaMemcpyAsync(traces, b->dTracesPacked, (size_t) total, cudaMemcpyDeviceToHost, b->stream));
		for (int s = 0; s <= b->nSeq; ++s) traceOff[s] = b->hTraceOffs[s];
		HU_CUDA(cudaStreamSynchronize(b->stream));
		b->d2hBytes = total + (long long) (b->nSeq + 1) * 8 + (long long) b->nSeq * sizeof(hu_align_row);
	} else
		b->d2hBytes = (long long) b->nSeq * sizeof(hu_align_row);
	return HU_OK;
}

int hu_batch_pack_rows(hu_batch* b, int64_t read0, void* devRows) {
	if (!b || !devRows || b->nReads < 1) { setError("hu_batch_pack_rows: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	packRowsKernel<<<(b->nReads + 127) / 128, 128, 0, b->stream>>>(b->nReads, read0, b->merged, b->best, (PackedRow*) devRows);
	b->launches += 1;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}
void* hu_batch_stream(hu_batch* b) { return b ? (void*) b->stream : NULL; }

#define HU_GET(dst, src, count)                                                                                   \
	do {                                                                                                          \
		if (dst) HU_CUDA(cudaMemcpyAsync(dst, src, (size_t) (count) * sizeof(*(dst)), cudaMemcpyDeviceToHost, b->stream)); \
	} while (0)

int hu_batch_get_merged(hu_batch* b, hu_align_row* rows) {
	if (!b) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_GET(rows, b->merged, b->nReads);
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}
int hu_batch_get_aligned_codes(hu_batch* b, int8_t* codes) {
	if (!b) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_GET(codes, b->vit.aligned, (size_t) b->nReads * b->db->hmm.L);
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}
int hu_batch_get_seeds(hu_batch* b, int32_t* nSeeds, int32_t* ids, double* dist) {
	if (!b) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	std::vector<int> n(b->nReads), all((size_t) b->nReads * HU_MAX_SEED);
	HU_CUDA(cudaMemcpyAsync(n.data(), b->seed.nSeeds, n.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
	HU_CUDA(cudaMemcpyAsync(all.data(), b->seed.seedIds, all.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
	std::vector<hu_place_row> est;
	if (dist) {
		est.resize((size_t) b->nReads * HU_MAX_SEED);
		HU_CUDA(cudaMemcpyAsync(est.data(), b->est, est.size() * sizeof(hu_place_row), cudaMemcpyDeviceToHost, b->stream));
	}
	HU_CUDA(cudaStreamSynchronize(b->stream));
	const int mx = b->seed.maxNSeed;
	for (int r = 0; r < b->nReads; ++r) {
		if (nSeeds) nSeeds[r] = n[r];
		for (int s = 0; s < mx; ++s) {
			if (ids) ids[(size_t) r * mx + s] = s < n[r] ? all[(size_t) r * HU_MAX_SEED + s] : -1;
			if (dist) dist[(size_t) r * mx + s] = s < n[r] ? est[(size_t) r * HU_MAX_SEED + s].dist : NAN; // valid after hu_estimate_batch
		}
	}
	return HU_OK;
}
static int getRows(hu_batch* b, const hu_place_row* src, bool ordered, int32_t* nKept, hu_place_row* rows) {
	HU_CUDA(cudaSetDevice(b->db->device));
	std::vector<int> n(b->nReads), ord((size_t) b->nReads * HU_MAX_SEED);
	std::vector<hu_place_row> all((size_t) b->nReads * HU_MAX_SEED);
	HU_CUDA(cudaMemcpyAsync(n.data(), b->nKept, n.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
	HU_CUDA(cudaMemcpyAsync(ord.data(), b->order, ord.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
	HU_CUDA(cudaMemcpyAsync(all.data(), src, all.size() * sizeof(hu_place_row), cudaMemcpyDeviceToHost, b->stream));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	const int mx = b->seed.maxNSeed;
	for (int r = 0; r < b->nReads; ++r) {
		if (nKept) nKept[r] = n[r];
		for (int i = 0; i < n[r] && rows; ++i)
			rows[(size_t) r * mx + i] = all[(size_t) r * HU_MAX_SEED + (ordered ? ord[(size_t) r * HU_MAX_SEED + i] : i)];
	}
	return HU_OK;
}
int hu_batch_get_estimates(hu_batch* b, int32_t* nKept, hu_place_row* rows) {
	if (!b) return HU_ERR_ARG;
	return getRows(b, b->est, true, nKept, rows);
}
int hu_batch_get_places(hu_batch* b, int32_t* nKept, hu_place_row* rows) {
	if (!b) return HU_ERR_ARG;
	return getRows(b, b->places, false, nKept, rows);
}
int hu_batch_get_best(hu_batch* b, hu_place_row* rows) {
	if (!b || !rows) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_GET(rows, b->best, b->nReads);
	HU_CUDA(cudaStreamSynchronize(b->stream));
	return HU_OK;
}

int hu_batch_get_seed_keys(hu_batch* b, int32_t read, uint32_t* keys, int32_t* nCand) {
	if (!b || read < 0 || read >= b->nReads) { setError("hu_batch_get_seed_keys: bad read index"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	const int n = b->seed.nCand > 0 ? b->seed.nCand : b->db->tree.nCand;
	if (nCand) *nCand = n;
	if (keys) HU_CUDA(cudaMemcpy(keys, b->seed.keys + (size_t) read * b->db->tree.scanPitch, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return HU_OK;
}

int hu_debug_stdsort(const double* keys, int32_t n, int32_t* perm) {
	if (!keys || !perm || n < 0) { setError("hu_debug_stdsort: bad arguments"); return HU_ERR_ARG; }
	stdSortDescendingPerm(keys, n, perm);
	return HU_OK;
}

int hu_batch_mark(hu_batch* b, int id) {
	if (!b || id < 0 || id > 7) { setError("hu_batch_mark: bad mark id"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_CUDA(cudaEventRecord(b->ev[8 + id], b->stream));
	return HU_OK;
}
int hu_batch_elapsed(hu_batch* from, int idFrom, hu_batch* to, int idTo, float* ms) {
	if (!from || !to || !ms || idFrom < 0 || idFrom > 7 || idTo < 0 || idTo > 7) { setError("hu_batch_elapsed: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(from->db->device));
	HU_CUDA(cudaEventSynchronize(to->ev[8 + idTo]));
	HU_CUDA(cudaEventElapsedTime(ms, from->ev[8 + idFrom], to->ev[8 + idTo]));
	return HU_OK;
}

int hu_batch_stage_ms(hu_batch* b, float ms[8]) {
	if (!b || !ms) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	HU_CUDA(cudaStreamSynchronize(b->stream));
	static const int from[6] = {0, 2, 3, 4, 5, 6}, to[6] = {1, 3, 4, 5, 6, 7};
	for (int i = 0; i < 8; ++i) ms[i] = 0;
	for (int i = 0; i < 6; ++i)
		if (cudaEventElapsedTime(&ms[i], b->ev[from[i]], b->ev[to[i]]) != cudaSuccess) { ms[i] = -1; cudaGetLastError(); }
	return HU_OK;
}

int hu_batch_get_stats(hu_batch* b, hu_batch_stats* s) {
	if (!b || !s) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(b->db->device));
	memset(s, 0, sizeof *s);
	std::vector<hu_align_row> rows(b->nSeq);
	std::vector<int> ws(b->nReads), we(b->nReads), ns(b->nReads), nk(b->nReads);
	HU_CUDA(cudaStreamSynchronize(b->stream));
	if (b->vit.plans) HU_CUDA(cudaMemcpy(rows.data(), b->vit.rows, rows.size() * sizeof(hu_align_row), cudaMemcpyDeviceToHost));
	HU_CUDA(cudaMemcpy(ws.data(), b->winStart, ws.size() * sizeof(int), cudaMemcpyDeviceToHost));
	HU_CUDA(cudaMemcpy(we.data(), b->winEnd, we.size() * sizeof(int), cudaMemcpyDeviceToHost));
	HU_CUDA(cudaMemcpy(ns.data(), b->seed.nSeeds, ns.size() * sizeof(int), cudaMemcpyDeviceToHost));
	HU_CUDA(cudaMemcpy(nk.data(), b->nKept, nk.size() * sizeof(int), cudaMemcpyDeviceToHost));
	if (b->vit.plans)
		for (int i = 0; i < b->nSeq; ++i) s->cells += rows[i].cells;
	const long long nodes = b->db->tree.nNodes - 1;
	for (int r = 0; r < b->nReads; ++r) {
		if (ws[r] < 0) continue;
		const long long W = we[r] - ws[r] + 1;
		s->seedPairs += nodes;
		s->seedBytes += nodes * W + W;
		s->estItems += ns[r];
		s->estBytes += 65 * W * ns[r];
		s->placeItems += nk[r];
		s->placeBytes += 64 * W * nk[r];
	}
	s->launches = b->launches;
	int hf = 0;
	HU_CUDA(cudaMemcpy(&hf, b->workCount + 4, sizeof(int), cudaMemcpyDeviceToHost));
	s->seedHeapSorts = hf;
	HU_CUDA(cudaMemcpy(&hf, b->workCount + 5, sizeof(int), cudaMemcpyDeviceToHost));
	s->seedScanFlags = hf;
	if (b->mmScanRan && b->mmRange) {
		const int numNt = (b->nReads + HU_MM_BN - 1) / HU_MM_BN;
		std::vector<int2> rg(numNt);
		std::vector<int> ext(numNt);
		HU_CUDA(cudaMemcpy(rg.data(), b->mmRange, rg.size() * sizeof(int2), cudaMemcpyDeviceToHost));
		HU_CUDA(cudaMemcpy(ext.data(), b->mmTileExt, ext.size() * sizeof(int), cudaMemcpyDeviceToHost));
		for (int i = 0; i < numNt; ++i)
			if (rg[i].x <= rg[i].y)
				s->seedMacs += (int64_t) (rg[i].y - rg[i].x + 1 + (ext[i] ? 1 : 0)) * HU_MM_BK * HU_MM_BN * (int64_t) b->db->tree.posPitch;
	}
	return HU_OK;
}

// buildGlobalAlign (src/BandedHMMP7.cpp:1008-1081) with getPaddingSeq (cpp:1137-1186), on the host.
int hu_trace_to_align(const hu_hmm_params* hmm, const char* read, int32_t readLen, const hu_align_row* row,
		const char* trace, char* align) {
	if (!hmm || !read || !row || !trace || !align) { setError("hu_trace_to_align: NULL argument"); return HU_ERR_ARG; }
	const int L = hmm->L;
	const int* p2cs = hmm->p2cs;
	std::string out;
	out.reserve(L);
	const std::string seq(read, readLen);
	const std::string seqN = seq.substr(0, row->seqStart - 1);
	const std::string seqC = row->seqEnd < readLen ? seq.substr(row->seqEnd) : std::string();
	const int csStart = row->csStart, csEnd = row->csEnd;
	int j = 0, k = 0;
	std::string insert;
	const int tl = row->traceLen;
	for (int t = 0; t < tl; ++t) {
		switch (trace[t]) {
		case 'B': { // right-justified N-terminal remainder
			const int len = csStart - 1, n = (int) seqN.size();
			if (n == 0) out.append(len, '.');
			else if (n >= len) out.append(seqN, n - len, len);
			else { out.append(len - n, '.'); out.append(seqN); }
			j = row->seqStart; k = row->hmmStart;
			break;
		}
		case 'M':
			if (k > 1 && t > 1 && p2cs[k] - p2cs[k - 1] > 1) {
				const int len = p2cs[k] - p2cs[k - 1] - 1, n = (int) insert.size();
				if (n == 0) out.append(len, '-');
				else if (n >= len) { // both halves of a too-long insert
					out.append(insert, 0, len / 2);
					out.append(insert, n - (len - len / 2), len - len / 2);
				} else { // literal: the head of the insert is written on both sides of the fill
					out.append(insert, 0, n / 2);
					out.append(len - n, '-');
					out.append(insert, 0, n - n / 2);
				}
			}
			insert.clear();
			out.push_back(seq.at(j - 1));
			j++; k++;
			break;
		case 'I':
			insert.clear();
			while (t < tl && trace[t] == 'I') { insert.push_back((char) tolower(seq.at(j - 1))); j++; t++; }
			t--;
			break;
		case 'D':
			if (k > 1 && p2cs[k] - p2cs[k - 1] > 1) out.append(p2cs[k] - p2cs[k - 1] - 1, '-');
			out.push_back('-');
			k++;
			break;
		case 'E': { // left-justified C-terminal remainder
			const int len = L - csEnd, n = (int) seqC.size();
			if (n == 0) out.append(len, '.');
			else if (n >= len) out.append(seqC, 0, len);
			else { out.append(seqC); out.append(len - n, '.'); }
			break;
		}
		default:
			setError("hu_trace_to_align: bad trace state '%c'", trace[t]);
			return HU_ERR_ARG;
		}
	}
	if ((int) out.size() != L) { setError("hu_trace_to_align: aligned length %zu != L %d", out.size(), L); return HU_ERR_STATE; }
	memcpy(align, out.data(), L);
	return HU_OK;
}

} // extern "C"
