// K2: the "Seed" stage of SEP.  getSeed() computes SeqUtils::pDist(node.seq, read, start, end) = d/N for
// every non-root node (src/HmmUFOtu_main.cpp:127-152, src/SeqUtils.cpp:37-54) and keeps the maxNSeed
// closest (src/hmmufotu.cpp:646-647).
//
// Node sequences are held as three bit planes (hi, lo, valid) of 32 consensus columns per word, slots
// contiguous per (word, plane), so a CTA streams a 512-slot tile with coalesced 128-bit loads and compares
// it against 8 reads of the batch kept in shared memory: d = popc(((nh^rh)|(nl^rl)) & nv & rv),
// N = popc(nv & rv).  Ungapped (inferred, internal-node) sequences skip the valid plane.  The result per
// (read, node) is the exact order-preserving integer key floor(d*2^31/N) (distinct rationals with
// denominators < 2^13 differ by more than 2^-26, so the key orders exactly like the reference's doubles),
// packed with the node id, and a per-read radix select returns the maxNSeed smallest in (dist, id) order.
#include "hu_batch.cuh"

namespace hu {

// aligned codes -> bit planes of the window, [word][read][4]
__global__ void packReadsKernel(SeedArgs a, int L, int nWords) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	const int w = blockIdx.y;
	if (r >= a.nReadsPad) return;
	uint32_t hi = 0, lo = 0, va = 0;
	if (r < a.nReads && a.winStart[r] >= 0) {
		const int s = a.winStart[r], e = a.winEnd[r];
		const int c0 = max(w * 32, s), c1 = min(min(w * 32 + 31, e), L - 1);
		const int8_t* x = a.aligned + (size_t) r * L;
		for (int c = c0; c <= c1; ++c) {
			const int v = x[c];
			if (v >= 0) {
				const int b = c - w * 32;
				va |= 1u << b;
				hi |= (uint32_t) ((v >> 1) & 1) << b;
				lo |= (uint32_t) (v & 1) << b;
			}
		}
		if (va) atomicAdd(a.rValid + r, __popc(va));
	}
	reinterpret_cast<uint4*>(a.rplanes)[(size_t) w * a.nReadsPad + r] = make_uint4(hi, lo, va, 0u);
}

__device__ __forceinline__ unsigned long long makeKey(int d, int N, int node, bool excluded) {
	uint32_t q = 0xFFFFFFFFu;
	if (!excluded && N > 0) q = (uint32_t) __double2ll_rd(__ddiv_rn((double) d * 2147483648.0, (double) N));
	return ((unsigned long long) q << 32) | (uint32_t) (excluded ? 0x7FFFFFFF : node);
}

template <bool GAP>
__global__ void __launch_bounds__(128) seedScanKernel(SeedArgs a, TreeDev t, int tile0) {
	extern __shared__ uint4 sR[]; // [word - gLo][8]
	__shared__ int sLo, sHi;
	const int g = blockIdx.x;             // read group
	const int tile = tile0 + blockIdx.y;  // slot tile
	const int r0 = g * HU_SCAN_READS;
	if (threadIdx.x == 0) {
		int lo = 0x7fffffff, hi = -1;
		for (int r = r0; r < r0 + HU_SCAN_READS; ++r)
			if (r < a.nReads && a.winStart[r] >= 0) {
				lo = min(lo, a.winStart[r] >> 5);
				hi = max(hi, a.winEnd[r] >> 5);
			}
		sLo = lo; sHi = hi;
	}
	__syncthreads();
	const int gLo = sLo, gHi = sHi;
	if (gHi < 0) return;
	const uint4* rp = reinterpret_cast<const uint4*>(a.rplanes);
	for (int i = threadIdx.x; i < (gHi - gLo + 1) * HU_SCAN_READS; i += blockDim.x)
		sR[i] = rp[(size_t) (gLo + i / HU_SCAN_READS) * a.nReadsPad + r0 + (i % HU_SCAN_READS)];
	__syncthreads();
	const int slot = tile * HU_SCAN_TILE + threadIdx.x * 4;
	int d[HU_SCAN_READS][4];
	int n[GAP ? HU_SCAN_READS : 1][4];
#pragma unroll
	for (int r = 0; r < HU_SCAN_READS; ++r)
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			d[r][s] = 0;
			if (GAP) n[r][s] = 0;
		}
	const size_t pitch = t.scanPitch;
	for (int w = gLo; w <= gHi; ++w) {
		const uint4 nh = __ldg(reinterpret_cast<const uint4*>(t.planes + ((size_t) w * 3 + 0) * pitch + slot));
		const uint4 nl = __ldg(reinterpret_cast<const uint4*>(t.planes + ((size_t) w * 3 + 1) * pitch + slot));
		uint4 nv = make_uint4(~0u, ~0u, ~0u, ~0u);
		if (GAP) nv = __ldg(reinterpret_cast<const uint4*>(t.planes + ((size_t) w * 3 + 2) * pitch + slot));
		const uint32_t h4[4] = {nh.x, nh.y, nh.z, nh.w}, l4[4] = {nl.x, nl.y, nl.z, nl.w}, v4[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
		for (int r = 0; r < HU_SCAN_READS; ++r) {
			const uint4 q = sR[(w - gLo) * HU_SCAN_READS + r];
#pragma unroll
			for (int s = 0; s < 4; ++s) {
				const uint32_t m = GAP ? (v4[s] & q.z) : q.z;
				d[r][s] += __popc(((h4[s] ^ q.x) | (l4[s] ^ q.y)) & m);
				if (GAP) n[r][s] += __popc(m);
			}
		}
	}
	int node[4];
	bool excl[4];
#pragma unroll
	for (int s = 0; s < 4; ++s) {
		node[s] = __ldg(t.scanNode + slot + s);
		excl[s] = node[s] < 0 || !(__ldg(t.height + max(node[s], 0)) <= a.maxHeight);
	}
#pragma unroll
	for (int r = 0; r < HU_SCAN_READS; ++r) {
		const int rr = r0 + r;
		if (rr >= a.nReads || a.winStart[rr] < 0) continue;
		const int nv = GAP ? 0 : a.rValid[rr];
		unsigned long long k[4];
#pragma unroll
		for (int s = 0; s < 4; ++s) k[s] = makeKey(d[r][s], GAP ? n[r][s] : nv, node[s], excl[s]);
		ulonglong2* dst = reinterpret_cast<ulonglong2*>(a.keys + (size_t) rr * pitch + slot);
		dst[0] = make_ulonglong2(k[0], k[1]);
		dst[1] = make_ulonglong2(k[2], k[3]);
	}
}

// ---- per-read selection of the maxNSeed smallest (key, id) ----
#define HU_SEL_THREADS 256
#define HU_SEL_CAND 2048
#define HU_SEL_BINS 2048

__global__ void __launch_bounds__(HU_SEL_THREADS) seedSelectKernel(SeedArgs a, int scanPitch) {
	__shared__ unsigned int hist[HU_SEL_BINS];
	__shared__ unsigned long long cand[HU_SEL_CAND];
	__shared__ unsigned int sCount, sValid, sDigit, sBelow;
	__shared__ unsigned int sMaxQ;
	const int r = blockIdx.x;
	const int tid = threadIdx.x;
	if (a.winStart[r] < 0) {
		if (tid == 0) a.nSeeds[r] = 0;
		return;
	}
	const unsigned long long* keys = a.keys + (size_t) r * scanPitch;
	const ulonglong2* keys2 = reinterpret_cast<const ulonglong2*>(keys);
	const int n2 = scanPitch / 2;
	const int shifts[6] = {53, 42, 31, 20, 9, 0};
	const int widths[6] = {11, 11, 11, 11, 11, 9};
	unsigned long long prefixVal = 0, prefixMask = 0; // elements with (c & prefixMask) == prefixVal are undecided
	unsigned int below = 0;                          // elements known to be smaller than every undecided one
	unsigned int want = 0;
	unsigned long long bound = 0;                     // gather everything < bound
	for (int level = 0; level < 6; ++level) {
		for (int i = tid; i < HU_SEL_BINS; i += HU_SEL_THREADS) hist[i] = 0;
		if (tid == 0) { sValid = 0; sMaxQ = 0; }
		__syncthreads();
		const int sh = shifts[level];
		const unsigned int dm = (1u << widths[level]) - 1u;
		unsigned int valid = 0, maxq = 0;
		for (int i = tid; i < n2; i += HU_SEL_THREADS) {
			const ulonglong2 kk = keys2[i];
			const unsigned long long c2[2] = {kk.x, kk.y};
#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const unsigned long long c = c2[e];
				const unsigned int q = (unsigned int) (c >> 32);
				if (q == 0xFFFFFFFFu) continue;
				if (level == 0) { valid++; maxq = max(maxq, q); }
				if ((c & prefixMask) == prefixVal) atomicAdd(&hist[(unsigned int) (c >> sh) & dm], 1u);
			}
		}
		if (level == 0) {
			atomicAdd(&sValid, valid);
			atomicMax(&sMaxQ, maxq);
		}
		__syncthreads();
		if (level == 0) want = min((unsigned int) a.maxNSeed, sValid);
		if (want == 0) {
			if (tid == 0) a.nSeeds[r] = 0;
			return;
		}
		if (tid == 0) { // serial scan of 2048 bins: negligible next to the key passes
			unsigned int cum = below, dgt = 0;
			for (; dgt < HU_SEL_BINS; ++dgt) {
				if (cum + hist[dgt] >= want) break;
				cum += hist[dgt];
			}
			sDigit = dgt; sBelow = cum; sCount = cum + hist[dgt];
		}
		__syncthreads();
		const unsigned int dgt = sDigit;
		below = sBelow;
		prefixVal |= (unsigned long long) dgt << sh;
		prefixMask |= (unsigned long long) dm << sh;
		bound = prefixVal + (1ull << sh);
		const unsigned int total = sCount;
		__syncthreads();
		if (total <= HU_SEL_CAND || level == 5) break;
	}
	// gather everything below the bound (count <= HU_SEL_CAND) and sort it
	if (tid == 0) sCount = 0;
	__syncthreads();
	for (int i = tid; i < n2; i += HU_SEL_THREADS) {
		const ulonglong2 kk = keys2[i];
		if (kk.x < bound && (unsigned int) (kk.x >> 32) != 0xFFFFFFFFu) { const unsigned int p = atomicAdd(&sCount, 1u); if (p < HU_SEL_CAND) cand[p] = kk.x; }
		if (kk.y < bound && (unsigned int) (kk.y >> 32) != 0xFFFFFFFFu) { const unsigned int p = atomicAdd(&sCount, 1u); if (p < HU_SEL_CAND) cand[p] = kk.y; }
	}
	__syncthreads();
	const unsigned int cnt = min(sCount, (unsigned int) HU_SEL_CAND);
	unsigned int m = 1;
	while (m < cnt) m <<= 1;
	for (unsigned int i = cnt + tid; i < m; i += HU_SEL_THREADS) cand[i] = ~0ull;
	__syncthreads();
	for (unsigned int k = 2; k <= m; k <<= 1)
		for (unsigned int j = k >> 1; j > 0; j >>= 1) {
			for (unsigned int i = tid; i < m; i += HU_SEL_THREADS) {
				const unsigned int ixj = i ^ j;
				if (ixj > i) {
					const unsigned long long x = cand[i], y = cand[ixj];
					const bool up = (i & k) == 0;
					if ((x > y) == up) { cand[i] = y; cand[ixj] = x; }
				}
			}
			__syncthreads();
		}
	// optional maxDiff cut (src/HmmUFOtu_main.cpp:141-150), on the key scale
	unsigned int nOut = min(want, cnt);
	if (a.maxDiff < INFINITY && nOut > 0) {
		const double best = (double) (cand[0] >> 32) / 2147483648.0, worst = (double) sMaxQ / 2147483648.0;
		if (worst < best + a.maxDiff) {
			unsigned int keep = 0;
			while (keep < nOut && (double) (cand[keep] >> 32) / 2147483648.0 - best <= a.maxDiff) keep++;
			nOut = keep;
		}
	}
	for (unsigned int i = tid; i < nOut; i += HU_SEL_THREADS) a.seedIds[(size_t) r * HU_MAX_SEED + i] = (int) (cand[i] & 0x7FFFFFFFull);
	if (tid == 0) a.nSeeds[r] = (int) nOut;
}

int launchSeed(hu_batch* b, const hu_sep_params* p) {
	hu_db* db = b->db;
	const TreeDev& t = db->tree;
	if (p->maxNSeed < 1 || p->maxNSeed > HU_MAX_SEED) { setError("maxNSeed %d out of range 1..%d", p->maxNSeed, HU_MAX_SEED); return HU_ERR_ARG; }
	SeedArgs a = b->seed;
	a.nReads = b->nReads;
	a.nReadsPad = (b->nReads + HU_SCAN_READS - 1) / HU_SCAN_READS * HU_SCAN_READS;
	a.maxNSeed = p->maxNSeed;
	a.maxDiff = p->maxDiff;
	a.maxHeight = p->maxHeight;
	a.aligned = b->vit.aligned;
	a.winStart = b->winStart;
	a.winEnd = b->winEnd;
	b->seed = a;
	HU_CUDA(cudaEventRecord(b->ev[2], b->stream));
	HU_CUDA(cudaMemsetAsync(a.rValid, 0, (size_t) a.nReadsPad * sizeof(int), b->stream));
	packReadsKernel<<<dim3((a.nReadsPad + 127) / 128, t.nWords), 128, 0, b->stream>>>(a, t.csLen, t.nWords);
	const int groups = a.nReadsPad / HU_SCAN_READS;
	const size_t smem = (size_t) t.nWords * HU_SCAN_READS * sizeof(uint4);
	static bool attr = false;
	if (!attr) {
		HU_CUDA(cudaFuncSetAttribute(seedScanKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		HU_CUDA(cudaFuncSetAttribute(seedScanKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		attr = true;
	}
	if (smem > 200 * 1024) { setError("consensus too long for the seed-scan read tile (%d words)", t.nWords); return HU_ERR_ARG; }
	const int gapTiles = t.nGapSlots / HU_SCAN_TILE, allTiles = t.scanPitch / HU_SCAN_TILE;
	// grid.x (fast) runs over read groups so that consecutive CTAs reuse the same node tile from L2
	if (gapTiles > 0) seedScanKernel<true><<<dim3(groups, gapTiles), 128, smem, b->stream>>>(a, t, 0);
	if (allTiles > gapTiles) seedScanKernel<false><<<dim3(groups, allTiles - gapTiles), 128, smem, b->stream>>>(a, t, gapTiles);
	HU_CUDA(cudaEventRecord(b->ev[3], b->stream));
	seedSelectKernel<<<b->nReads, HU_SEL_THREADS, 0, b->stream>>>(a, t.scanPitch);
	HU_CUDA(cudaEventRecord(b->ev[4], b->stream));
	b->launches += 2 + (gapTiles > 0) + (allTiles > gapTiles);
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
