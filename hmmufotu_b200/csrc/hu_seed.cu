// K2: the "Seed" stage of SEP.  getSeed() computes SeqUtils::pDist(node.seq, read, start, end) = d/N for
// every non-root node (src/HmmUFOtu_main.cpp:127-152, src/SeqUtils.cpp:37-54) and keeps the maxNSeed
// closest (src/hmmufotu.cpp:646-647).
//
// d and N are sums over consensus columns, so the columns are split exactly into
//   * the K match columns of the profile: an aligned read has nearly all of its bases there.  Node and read
//     sequences are held as three bit planes (hi, lo, valid) over PROFILE positions, 32 per word, node slots
//     contiguous per (word, plane): a CTA streams a 512-slot tile with coalesced 128-bit loads against 8 reads
//     kept in shared memory, d += popc(((nh^rh)|(nl^rl)) & nv & rv), N += popc(nv & rv).  A 250-base read spans
//     ~9 words instead of the ~43 words its 1 400 consensus columns would take.  Ungapped (inferred,
//     internal-node) sequences skip the valid plane.
//   * the non-match columns, where a read only has the few insert bases buildGlobalAlign could place: these are
//     listed per read and added from a slot-major transposed copy of the node sequences in the scan epilogue.
// The result per (read, node) is the exact order-preserving integer key floor(d*2^31/N): two distinct rationals
// d1/N1 != d2/N2 differ by at least 1/(N1 N2), so their keys differ whenever N1 N2 <= 2^31, i.e. for N <= 46 340
// compared columns, and equal rationals give equal keys and equal doubles -- keys order, and tie, exactly like the
// reference's d/N doubles.  N is bounded by the bases a read has in its window; hu_batch_create refuses workspaces
// whose reads could exceed the bound (HU_SEED_MAX_N);
// a per-read radix select over (key, node id) returns the maxNSeed smallest in (dist, id) order.
#include <stdlib.h>
#include <string.h>
#include "hu_batch.cuh"
#include "hu_stdsort.cuh"

namespace hu {

// aligned codes -> bit planes over profile positions, [word][read][4]; one thread per (word, read)
__global__ void packReadsKernel(SeedArgs a, int L, int K, const int* __restrict__ p2cs) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	const int w = blockIdx.y;
	if (r >= a.nReadsPad) return;
	uint32_t hi = 0, lo = 0, va = 0;
	if (r < a.nReads && a.winStart[r] >= 0) {
		const int s = a.winStart[r], e = a.winEnd[r];
		const int8_t* x = a.aligned + (size_t) r * L;
		for (int b = 0; b < 32; ++b) {
			const int k = w * 32 + b + 1;
			if (k > K) break;
			const int c = __ldg(p2cs + k) - 1;
			if (c < s || c > e) continue;
			const int v = x[c];
			if (v >= 0) {
				va |= 1u << b;
				hi |= (uint32_t) ((v >> 1) & 1) << b;
				lo |= (uint32_t) (v & 1) << b;
			}
		}
		if (va) {
			atomicAdd(a.rValid + r, __popc(va));
			atomicMin(a.rWordLo + r, w);
			atomicMax(a.rWordHi + r, w);
		}
	}
	reinterpret_cast<uint4*>(a.rplanes)[(size_t) w * a.nReadsPad + r] = make_uint4(hi, lo, va, 0u);
}

// bases of the read in non-match columns of its window: (column << 8) | base; one warp per read
__global__ void sparseListKernel(SeedArgs a, int L, const int* __restrict__ colK) {
	const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (r >= a.nReads) return;
	int n = 0;
	if (a.winStart[r] >= 0) {
		const int s = a.winStart[r], e = a.winEnd[r];
		const int8_t* x = a.aligned + (size_t) r * L;
		uint32_t* out = a.sparse + (size_t) r * a.sparseCap;
		for (int c0 = s; c0 <= e; c0 += 32) {
			const int c = c0 + lane;
			const bool hit = c <= e && x[c] >= 0 && __ldg(colK + c) == 0;
			const unsigned m = __ballot_sync(0xffffffffu, hit);
			if (hit) {
				const int pos = n + __popc(m & ((1u << lane) - 1u));
				if (pos < a.sparseCap) out[pos] = ((uint32_t) c << 8) | (uint32_t) x[c];
			}
			n += __popc(m);
		}
	}
	// more non-match bases than the workspace was sized for (only possible through hu_batch_set_aligned with
	// rows denser than maxSeqLen): the read is reported unplaced rather than scored on a truncated list
	if (lane == 0) a.nSparse[r] = n <= a.sparseCap ? n : -1;
}

// Seed key floor(d * 2^31 / N) for 0 <= d <= N < 2^16, exactly and without a division or 64-bit integer
// arithmetic: inv2N[N] is 2^31 / N correctly rounded, so d * inv2N[N] is within 2^-21 of the quotient.  A quotient
// that is not an integer has a fractional part in [1/N, 1 - 1/N], far from both ends, and an integer quotient
// is lifted over the rounding error by the 2^-20 bias: the floor of the biased product is the exact key
// (checked on the host for every d <= N < 3000 and sampled N up to 65535, tests/test_abi.py).
#define HU_KEY_BIAS 9.5367431640625e-07 // 2^-20
__device__ __forceinline__ uint32_t makeKey(int d, int N, bool excluded, const double* __restrict__ inv2N) {
	const double q = __dadd_rn(__dmul_rn((double) d, __ldg(inv2N + N)), HU_KEY_BIAS);
	return (excluded || N <= 0) ? 0xFFFFFFFFu : __double2uint_rd(q); // 0/0: the reference's NaN distance; such nodes sort last here
}

template <bool GAP>
__global__ void __launch_bounds__(128) seedScanKernel(SeedArgs a, TreeDev t, int tile0) {
	extern __shared__ uint4 sR[]; // [word - gLo][8]
	__shared__ int sLo, sHi;
	__shared__ int sNs[HU_SCAN_READS], sNv[HU_SCAN_READS]; // per read: # sparse bases (-2: no such read), # valid bases
	const int g = blockIdx.x;             // read group
	const int tile = tile0 + blockIdx.y;  // slot tile
	const int r0 = g * HU_SCAN_READS;
	if (threadIdx.x >= 32 && threadIdx.x < 32 + HU_SCAN_READS) {
		const int rr = r0 + (int) threadIdx.x - 32;
		const bool ok = rr < a.nReads && a.winStart[rr] >= 0;
		sNs[threadIdx.x - 32] = ok ? a.nSparse[rr] : -2;
		sNv[threadIdx.x - 32] = ok ? a.rValid[rr] : 0;
	}
	if (threadIdx.x == 0) {
		int lo = 0x7fffffff, hi = -1;
		for (int r = r0; r < r0 + HU_SCAN_READS; ++r)
			if (r < a.nReads && a.winStart[r] >= 0 && a.rWordHi[r] >= 0) {
				lo = min(lo, a.rWordLo[r]);
				hi = max(hi, a.rWordHi[r]);
			}
		sLo = lo; sHi = hi;
	}
	__syncthreads();
	const int gLo = sLo, gHi = sHi;
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
	int pos[4];
	bool excl[4];
#pragma unroll
	for (int s = 0; s < 4; ++s) {
		const int node = __ldg(t.scanNode + slot + s);
		pos[s] = __ldg(a.posOfSlot + slot + s);
		excl[s] = node < 0 || !(__ldg(t.height + max(node, 0)) <= a.maxHeight);
	}
#pragma unroll
	for (int r = 0; r < HU_SCAN_READS; ++r) {
		const int rr = r0 + r;
		const int ns = sNs[r];
		if (ns == -2) continue; // uniform over the CTA
		int dd[4], nn[4];
#pragma unroll
		for (int s = 0; s < 4; ++s) { dd[s] = d[r][s]; nn[s] = GAP ? n[r][s] : sNv[r]; }
		// insert bases of the read in non-match columns
		const uint32_t* sp = a.sparse + (size_t) rr * a.sparseCap;
		for (int i = 0; i < ns; ++i) {
			const uint32_t e = __ldg(sp + i);
			const int rb = (int) (e & 0xff);
			const char4 nb = *reinterpret_cast<const char4*>(t.seqT + (size_t) (e >> 8) * pitch + slot);
			const int b4[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
			for (int s = 0; s < 4; ++s)
				if (b4[s] >= 0) { nn[s]++; dd[s] += b4[s] != rb; }
		}
		// the key vector of a read is in getSeed's candidate order (node id), not in slot order
		uint32_t* krow = a.keys + (size_t) rr * pitch;
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			const uint32_t key = makeKey(dd[s], nn[s], excl[s] || ns < 0, t.inv2N);
			if (pos[s] >= 0) krow[pos[s]] = key;
		}
	}
}

// Measured and dropped in round 2 (profiles/r2_summary.md): a scan tiled by candidate position (a CTA owns 1024 consecutive
// positions of 8 reads, scans the slots of both classes that map there, parks the keys in shared memory and writes whole
// 128-byte lines).  It removes the partial-sector key writes (26 GB of DRAM traffic for a 6.5 GB key matrix) but the class
// mix of a position range varies, the quads of a tile do not divide evenly over the threads, and the kernel is bound by
// POPC issue, not by DRAM: 15.6 ms against 8.7 ms.

// ---- per-read selection of the maxNSeed smallest (key, node id) ----
#define HU_SEL_THREADS 256
#define HU_SEL_CAND 2048
#define HU_SEL_BINS 2048

__global__ void __launch_bounds__(HU_SEL_THREADS) seedSelectKernel(SeedArgs a, int scanPitch, const int* __restrict__ nodeOfPos) {
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
	const uint4* keys4 = reinterpret_cast<const uint4*>(a.keys + (size_t) r * scanPitch);
	const int4* node4 = reinterpret_cast<const int4*>(nodeOfPos);
	const int n4 = scanPitch / 4;
	// composite value c = (key << 32) | node id, refined 11 bits at a time from the top
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
		for (int i = tid; i < n4; i += HU_SEL_THREADS) {
			const uint4 kk = keys4[i];
			const unsigned int q4[4] = {kk.x, kk.y, kk.z, kk.w};
			int4 nd = make_int4(0, 0, 0, 0);
			if (sh < 32) nd = __ldg(node4 + i); // node ids only matter once the key bits are exhausted
			const int n4v[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const unsigned int q = q4[e];
				if (q == 0xFFFFFFFFu) continue;
				if (level == 0) { valid++; maxq = max(maxq, q); }
				const unsigned long long c = ((unsigned long long) q << 32) | (unsigned int) n4v[e];
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
	for (int i = tid; i < n4; i += HU_SEL_THREADS) {
		const uint4 kk = keys4[i];
		const unsigned int q4[4] = {kk.x, kk.y, kk.z, kk.w};
		const unsigned int qb = (unsigned int) (bound >> 32);
		if (min(min(q4[0], q4[1]), min(q4[2], q4[3])) > qb) continue;
		const int4 nd = __ldg(node4 + i);
		const int n4v[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			if (q4[e] == 0xFFFFFFFFu) continue;
			const unsigned long long c = ((unsigned long long) q4[e] << 32) | (unsigned int) n4v[e];
			if (c < bound) {
				const unsigned int p = atomicAdd(&sCount, 1u);
				if (p < HU_SEL_CAND) cand[p] = c;
			}
		}
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

// ---- the seeds std::sort leaves in front (HU_SEED_ORDER_STDSORT) ----
// getSeed() sorts all candidates with libstdc++'s unstable introsort and the caller keeps the first maxNSeed
// (src/HmmUFOtu_main.cpp:139, src/hmmufotu.cpp:646-647).  p-distances are small rationals, so the maxNSeed-th
// value is nearly always a tie group cut in two, and WHICH of its members survive is decided by the sort's
// data movement.  This kernel replays that movement exactly, but only on the part of the array that can
// reach the first maxNSeed positions: __introsort_loop (bits/stl_algo.h) partitions [first, last) around the
// median of (first+1, mid, last-1), recurses into the right part and loops on the left part, so a sub-range
// that starts at or after position maxNSeed is never looked at again.  One CTA per read; the Hoare partition
// (__unguarded_partition) is evaluated in parallel from two bitmaps (elements >= pivot, elements <= pivot):
// the k-th stop of the left scan is swapped with the k-th stop of the right scan while the former lies left
// of the latter, which prefix counts over the bitmaps decide for every element at once.
// Shape of a read's replay (phase clocks: profiles/r1d_summary.md): level 1 works straight on the key vector (flag pass
// with 128-bit loads, then only the left part [0, cut) is materialised as (key, position) pairs, row-wise, from a
// rank-ordered list of the right stops that move); later levels swap in place in the CTA's scratch row; once the live range
// is <= HU_LIT_SMEM pairs and nothing is pending outside it, it moves into shared memory for good; ranges <= HU_LIT_SMALL
// are finished by one warp (litWarpIntrosort); the final insertion sort is a rank computation over <= maxNSeed + 17 pairs.
#define HU_LIT_THREADS 512
#define HU_LIT_UNROLL 8 // key loads in flight per lane in the flag passes
#define HU_LIT_VEC 8    // 128-bit key loads a lane has in flight in the first flag pass
#define HU_LIT_ROWS 4   // 32-position rows a warp has in flight while it materialises the first left part
#define HU_LIT_SWAPS 4  // swaps a thread has in flight in the swap phase of the later levels
#define HU_LIT_SMALL 256 // ranges up to this size (<= 1024: one bitmap word per lane) are finished by one warp in shared memory
#define HU_LIT_SMEM 8192 // once the range that holds the wanted prefix is this small it moves to shared memory for good

#ifdef HU_PROF // build-time instrumentation (HU_PROF=1 python -m hmmufotu_b200.build --force): phase clocks of CTA 0
__device__ unsigned long long huProfSeed[16];
#define LITP_MARK(i) do { if (blockIdx.x == 0 && tid == 0) { const long long now_ = clock64(); atomicAdd(&huProfSeed[i], (unsigned long long) (now_ - pt_)); pt_ = now_; } } while (0)
#define LITP_COUNT(i) do { if (blockIdx.x == 0 && tid == 0) atomicAdd(&huProfSeed[i], 1ull); } while (0)
extern "C" void hu_prof_seed(unsigned long long* out) { cudaMemcpyFromSymbol(out, huProfSeed, sizeof(huProfSeed)); }
#else
#define LITP_MARK(i) do {} while (0)
#define LITP_COUNT(i) do {} while (0)
#endif

typedef unsigned long long lit_t;               // (key << 32) | candidate position; only the key is compared
#define LIT_KEY(v) ((uint32_t) ((v) >> 32))
struct LitLess { __device__ __forceinline__ bool operator()(lit_t a, lit_t b) const { return LIT_KEY(a) < LIT_KEY(b); } };

// loads the compiler may not sink towards their first use: the flag passes issue a batch of these back to back, so
// that HU_LIT_UNROLL requests per lane are in flight (plain loads were re-serialised into load / use pairs)
__device__ __forceinline__ uint32_t ldKeyNow(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint4 ldKey4Now(const uint32_t* p) {
	uint4 v;
	asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ldPairKeyNow(const unsigned long long* p) { // high word of a (key, position) pair
	uint32_t v;
	asm volatile("ld.u32 %0, [%1+4];" : "=r"(v) : "l"(p)); // generic address: the later levels also run on a shared-memory copy
	return v;
}

// Batches of loads as ONE asm statement: ptxas otherwise re-serialises a run of independent loads into load / use
// pairs when registers are tight, and the passes that use these are bound by load latency.
__device__ __forceinline__ void ldKey4x8(const uint32_t* p, uint4 (&v)[8]) { // eight 128-bit loads, 8192 bytes apart
	asm volatile(
		"ld.global.v4.u32 {%0, %1, %2, %3}, [%32];\n\t"
		"ld.global.v4.u32 {%4, %5, %6, %7}, [%32+8192];\n\t"
		"ld.global.v4.u32 {%8, %9, %10, %11}, [%32+16384];\n\t"
		"ld.global.v4.u32 {%12, %13, %14, %15}, [%32+24576];\n\t"
		"ld.global.v4.u32 {%16, %17, %18, %19}, [%32+32768];\n\t"
		"ld.global.v4.u32 {%20, %21, %22, %23}, [%32+40960];\n\t"
		"ld.global.v4.u32 {%24, %25, %26, %27}, [%32+49152];\n\t"
		"ld.global.v4.u32 {%28, %29, %30, %31}, [%32+57344];"
		: "=r"(v[0].x), "=r"(v[0].y), "=r"(v[0].z), "=r"(v[0].w), "=r"(v[1].x), "=r"(v[1].y), "=r"(v[1].z), "=r"(v[1].w),
		  "=r"(v[2].x), "=r"(v[2].y), "=r"(v[2].z), "=r"(v[2].w), "=r"(v[3].x), "=r"(v[3].y), "=r"(v[3].z), "=r"(v[3].w),
		  "=r"(v[4].x), "=r"(v[4].y), "=r"(v[4].z), "=r"(v[4].w), "=r"(v[5].x), "=r"(v[5].y), "=r"(v[5].z), "=r"(v[5].w),
		  "=r"(v[6].x), "=r"(v[6].y), "=r"(v[6].z), "=r"(v[6].w), "=r"(v[7].x), "=r"(v[7].y), "=r"(v[7].z), "=r"(v[7].w)
		: "l"(p));
}
__device__ __forceinline__ void ldPairKeyx8(const unsigned long long* p, uint32_t (&v)[8]) { // high words of eight pairs, 4096 bytes apart (generic address)
	asm volatile(
		"ld.u32 %0, [%8+4];\n\t"
		"ld.u32 %1, [%8+4100];\n\t"
		"ld.u32 %2, [%8+8196];\n\t"
		"ld.u32 %3, [%8+12292];\n\t"
		"ld.u32 %4, [%8+16388];\n\t"
		"ld.u32 %5, [%8+20484];\n\t"
		"ld.u32 %6, [%8+24580];\n\t"
		"ld.u32 %7, [%8+28676];"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
		: "l"(p));
}
__device__ __forceinline__ void ldPairx4(const unsigned long long* p0, const unsigned long long* p1, const unsigned long long* p2,
		const unsigned long long* p3, unsigned long long (&v)[4]) {
	asm volatile(
		"ld.u64 %0, [%4];\n\t"
		"ld.u64 %1, [%5];\n\t"
		"ld.u64 %2, [%6];\n\t"
		"ld.u64 %3, [%7];"
		: "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}
static_assert(HU_LIT_THREADS == 512 && HU_LIT_VEC == 8 && HU_LIT_UNROLL == 8 && HU_LIT_SWAPS == 4,
              "the batch loaders above spell out these strides and counts");

template <typename Arr>
__device__ __forceinline__ void litSwap(Arr& w, int x, int y) { const lit_t t = w[x]; w[x] = w[y]; w[y] = t; }
template <typename Arr>
__device__ __forceinline__ void litMedianToFirst(Arr& w, int result, int a, int b, int c) { // __move_median_to_first
	const uint32_t qa = LIT_KEY(w[a]), qb = LIT_KEY(w[b]), qc = LIT_KEY(w[c]);
	if (qa < qb) {
		if (qb < qc) litSwap(w, result, b);
		else if (qa < qc) litSwap(w, result, c);
		else litSwap(w, result, a);
	} else if (qa < qc) litSwap(w, result, a);
	else if (qb < qc) litSwap(w, result, c);
	else litSwap(w, result, b);
}
// n-th (0-based, from bit 0) set bit of a word: five popc steps (the __fns intrinsic is a long dependent sequence)
__device__ __forceinline__ int litNthBit(uint32_t bits, int n) {
	int pos = 0, c;
	c = __popc(bits & 0xFFFFu); if (n >= c) { n -= c; pos += 16; bits >>= 16; }
	c = __popc(bits & 0xFFu); if (n >= c) { n -= c; pos += 8; bits >>= 8; }
	c = __popc(bits & 0xFu); if (n >= c) { n -= c; pos += 4; bits >>= 4; }
	c = __popc(bits & 0x3u); if (n >= c) { n -= c; pos += 2; bits >>= 2; }
	c = (int) (bits & 1u); if (n >= c) pos += 1;
	return pos;
}
// k-th (0-based) set bit of a bitmap with exclusive per-word prefix counts
__device__ __forceinline__ int litSelect(const uint32_t* bits, const uint32_t* pre, int nWords, uint32_t k) {
	int lo = 0, hi = nWords - 1;
	while (lo < hi) { // last word with pre[w] <= k
		const int mid = (lo + hi + 1) >> 1;
		if (pre[mid] <= k) lo = mid; else hi = mid - 1;
	}
	return lo * 32 + litNthBit(bits[lo], (int) (k - pre[lo]));
}

// number of swaps of a partition, this thread's words [w0, w1): left stops whose rank-mate on the right lies to their
// right.  The test is monotone along the array, so a word whose highest left stop passes counts whole and the
// first word that does not ends the walk.
__device__ __forceinline__ uint32_t litCountSwaps(const uint32_t* faW, const uint32_t* fbW, const uint32_t* preA, const uint32_t* preB,
		int w0, int w1, uint32_t totB) {
	uint32_t myK = 0;
	for (int wd = w0; wd < w1; ++wd) {
		uint32_t bits = faW[wd];
		if (!bits) continue;
		const int hb = 31 - __clz(bits);
		if (totB - (preB[wd] + __popc(fbW[wd] & (0xFFFFFFFFu >> (31 - hb)))) > preA[wd] + __popc(bits) - 1) { myK += __popc(bits); continue; }
		while (bits) {
			const int bit = __ffs(bits) - 1;
			bits &= bits - 1;
			const uint32_t rankA = preA[wd] + __popc(faW[wd] & ((1u << bit) - 1u));
			const uint32_t bIncl = preB[wd] + __popc(fbW[wd] & (0xFFFFFFFFu >> (31 - bit)));
			if (totB - bIncl > rankA) myK++;
			else break;
		}
		break;
	}
	return myK;
}

// __introsort_loop (bits/stl_algo.h) on sbuf[0, m), m <= 1024, by one warp: the same literal data movement as the
// CTA-wide levels (Hoare partition evaluated from the two bitmaps), with one bitmap word per lane, prefix counts by
// warp scan and no block barrier.  Sub-ranges that start at or beyond `want` (global position lo + l) are dropped.
// wsc: 128 words of shared scratch (fa, fb words and their exclusive prefix counts, for cross-lane selects).
__device__ __noinline__ void litWarpIntrosort(lit_t* sbuf, uint32_t* wsc, int m, int depth, int lo, int want, int* overflow) {
	const int lane = threadIdx.x & 31;
	int st[24][3], top = 0;
	st[0][0] = 0; st[0][1] = m; st[0][2] = depth; top = 1;
	while (top > 0) {
		top--;
		int l = st[top][0], h = st[top][1], dep = st[top][2];
		while (h - l > 16 && lo + l < want) {
			if (dep == 0) { // depth budget spent: std::__partial_sort(first, last, last) = heap sort of the range, literally (hu_stdsort.cuh)
				if (lane == 0) { ssHeapSort(sbuf + l, h - l, LitLess()); *overflow = 1; }
				__syncwarp();
				break;
			}
			dep--;
			if (lane == 0) litMedianToFirst(sbuf, l, l + 1, l + (h - l) / 2, h - 1);
			__syncwarp();
			const uint32_t pv = LIT_KEY(sbuf[l]);
			const int base = l + 1, cnt = h - base, nWs = (cnt + 31) / 32;
			uint32_t myA = 0, myB = 0;
			for (int r = 0; r < nWs; ++r) {
				const int t = r * 32 + lane;
				const bool ok = t < cnt;
				const uint32_t v = ok ? LIT_KEY(sbuf[base + t]) : 0u;
				const uint32_t ba = __ballot_sync(0xffffffffu, ok && !(v < pv));
				const uint32_t bb = __ballot_sync(0xffffffffu, ok && !(pv < v));
				if (lane == r) { myA = ba; myB = bb; }
			}
			uint32_t xa = __popc(myA), xb = __popc(myB);
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
				if (lane >= o) { xa += ya; xb += yb; }
			}
			const uint32_t totA = __shfl_sync(0xffffffffu, xa, 31), totB = __shfl_sync(0xffffffffu, xb, 31);
			wsc[lane] = myA; wsc[32 + lane] = myB; wsc[64 + lane] = xa - __popc(myA); wsc[96 + lane] = xb - __popc(myB);
			__syncwarp();
			// one position per lane and round (a lane walking the 32 bits of its own word would leave the warp idle):
			// K = number of left stops whose rank-mate on the right lies to their right
			const uint32_t lt = (1u << lane) - 1u, le = 0xFFFFFFFFu >> (31 - lane);
			uint32_t K = 0;
			for (int r = 0; r < nWs; ++r) {
				const uint32_t fa = wsc[r], fb = wsc[32 + r];
				const bool q = ((fa >> lane) & 1u) && totB - (wsc[96 + r] + __popc(fb & le)) > wsc[64 + r] + __popc(fa & lt);
				K += __popc(__ballot_sync(0xffffffffu, q));
			}
			const int AK = K < totA ? base + litSelect(wsc, wsc + 64, nWs, K) : h;
			int cut = AK;
			if (K >= 1) cut = min(AK, base + litSelect(wsc + 32, wsc + 96, nWs, totB - K));
			// swap k: k-th left stop <-> k-th right stop (from the right); the pairs are disjoint
			for (int r = 0; r < nWs; ++r) {
				const uint32_t fa = wsc[r];
				const uint32_t k = wsc[64 + r] + __popc(fa & lt);
				if (((fa >> lane) & 1u) && k < K) litSwap(sbuf, base + r * 32 + lane, base + litSelect(wsc + 32, wsc + 96, nWs, totB - 1 - k));
			}
			__syncwarp();
			if (lo + cut < want && top < 24) { st[top][0] = cut; st[top][1] = h; st[top][2] = dep; top++; }
			h = cut;
		}
	}
	__syncwarp();
}

__global__ void __launch_bounds__(HU_LIT_THREADS) seedSortLiteralKernel(SeedArgs a, int n, int pitch, const int* __restrict__ nodeOfPos,
		lit_t* __restrict__ scratch, int maxWords) {
	extern __shared__ __align__(16) uint32_t smLit[];
	uint32_t* faW = smLit;              // elements >= pivot  (!(x < pivot))
	uint32_t* fbW = faW + maxWords;     // elements <= pivot  (!(pivot < x))
	uint32_t* preA = fbW + maxWords;    // exclusive prefix counts per word
	uint32_t* preB = preA + maxWords;
	// shared-memory mode (ranges <= HU_LIT_SMEM): four bitmaps of HU_LIT_SMEM / 32 words, then the range itself
	lit_t* const sdata = reinterpret_cast<lit_t*>(smLit + 4 * (HU_LIT_SMEM / 32));
	__shared__ lit_t sbuf[HU_LIT_SMALL];
	__shared__ uint32_t wsc[128];
	__shared__ uint32_t red[HU_LIT_THREADS / 32][2];
	__shared__ uint32_t sTot[2], sK, sMaxQ;
	__shared__ int sCut, sLo, sHi, sDepth, sTop, sOverflow, sBound;
	__shared__ int stack[64][3];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nWarps = HU_LIT_THREADS / 32;
	lit_t* w = scratch + (size_t) blockIdx.x * pitch;
#ifdef HU_PROF
	long long pt_ = clock64();
#endif
	// Reads are taken from a queue: the cost of a replay varies with where the pivots fall (the first left part is anything from
	// a tenth to nine tenths of the vector), and with 28 reads per CTA dealt round-robin the slowest CTA set the stage's time.
	__shared__ int sRead;
	for (int it = 0;; ++it) {
		__syncthreads();
		int r = blockIdx.x + it * gridDim.x;
		if (a.litCursor) {
			if (tid == 0) sRead = atomicAdd(a.litCursor, 1);
			__syncthreads();
			r = sRead;
		}
		if (r >= a.nReads) break;
		LITP_MARK(7);
		LITP_COUNT(8);
		if (a.winStart[r] < 0 || n < 1) {
			if (tid == 0) a.nSeeds[r] = 0;
			continue;
		}
		const uint32_t* keys = a.keys + (size_t) r * pitch;
		const int want = min(a.maxNSeed, n);
		faW = smLit; fbW = faW + maxWords; preA = fbW + maxWords; preB = preA + maxWords;
		lit_t* wp = w;            // where positions live: the scratch row, or (shared-memory mode) sdata - smLo
		bool inSm = false;
		int smLo = 0, smHi = 0;
		// Level 1 straight from the key vector when the array is large: flags, cut, then only the left part
		// [0, cut) is materialised as (key, position) pairs -- the right part is never needed (cut >= want).
		const bool direct = n > 4 * HU_LIT_SMALL;
		uint32_t mx = 0;
		if (!direct)
			for (int i = tid; i < n; i += HU_LIT_THREADS) {
				const uint32_t v = keys[i];
				w[i] = ((lit_t) v << 32) | (uint32_t) i;
				if (v != 0xFFFFFFFFu) mx = max(mx, v);
			}
		if (tid == 0) { sMaxQ = 0; sTop = 0; sOverflow = 0; sBound = n; }
		__syncthreads();
		int depth0 = 0;
		for (int v = n; v > 1; v >>= 1) depth0++; // std::__lg(n)
		depth0 *= 2;
		bool started = false;
		if (direct) {
			// __move_median_to_first(first, first+1, mid, last-1): positions 0 and mS trade places (virtually)
			const int ia = 1, ib = n / 2, ic = n - 1;
			const uint32_t qa = keys[ia], qb = keys[ib], qc = keys[ic];
			int mS;
			if (qa < qb) mS = qb < qc ? ib : (qa < qc ? ic : ia);
			else mS = qa < qc ? ia : (qb < qc ? ic : ib);
			const uint32_t pv = keys[mS], k0v = keys[0];
#define VKEY(p) ((p) == mS ? k0v : keys[p])
#define VPOS(p) ((p) == mS ? 0 : (p))
			// Bit p of the two bitmaps is position p itself (bit 0, the pivot's slot, stays clear), so that the key vector
			// can be read with aligned 128-bit loads: HU_LIT_VEC of them in flight per lane (the pass is bound by load
			// latency).  A lane's four flags form a nibble; three exchanges assemble the words of eight lanes.
			const int base = 0, nW = (n + 31) / 32, nCh = (n + 127) / 128;
			// The scan's epilogue has made this pass already if it knew the pivot (hu_seedmm.cu): bit p of the read's two bitmaps =
			// (key at p >= pivot), (key at p <= pivot) for every position.  What differs here is local: bit 0 (the pivot's slot)
			// is clear, position mS holds what was first, and nothing exists at or beyond n.
			const bool fromScan = a.litBitA != NULL && a.litOk[r] && a.litPv[r] == pv; // uniform over the CTA
			if (fromScan) {
				if (tid == 0 && a.litFromScan) atomicAdd(a.litFromScan, 1);
				const uint32_t* ga = a.litBitA + (size_t) r * (size_t) (pitch >> 5);
				const uint32_t* gb = a.litBitB + (size_t) r * (size_t) (pitch >> 5);
				for (int wd = tid; wd < nW; wd += HU_LIT_THREADS) {
					uint32_t xa = ga[wd], xb = gb[wd];
					if (wd == (mS >> 5)) {
						const uint32_t bit = 1u << (mS & 31);
						xa = !(k0v < pv) ? (xa | bit) : (xa & ~bit);
						xb = !(pv < k0v) ? (xb | bit) : (xb & ~bit);
					}
					if (wd == 0) { xa &= ~1u; xb &= ~1u; }
					if (wd == nW - 1 && (n & 31)) { const uint32_t keep = (1u << (n & 31)) - 1u; xa &= keep; xb &= keep; }
					faW[wd] = xa; fbW[wd] = xb;
				}
			} else {
#define LIT_FLAG_CHUNK(ch, val) { \
				const uint32_t q_[4] = {(val).x, (val).y, (val).z, (val).w}; \
				uint32_t na_ = 0, nb_ = 0; \
				_Pragma("unroll") \
				for (int c = 0; c < 4; ++c) { \
					const int pp = (ch) * 128 + 4 * lane + c; \
					const uint32_t v = pp == mS ? k0v : q_[c]; \
					const bool ok = pp < n && pp > 0; \
					if (ok && v != 0xFFFFFFFFu) mx = max(mx, v); \
					na_ |= (uint32_t) (ok && !(v < pv)) << c; \
					nb_ |= (uint32_t) (ok && !(pv < v)) << c; \
				} \
				uint32_t xa_ = na_ << (4 * (lane & 7)), xb_ = nb_ << (4 * (lane & 7)); \
				_Pragma("unroll") \
				for (int o = 1; o < 8; o <<= 1) { \
					xa_ |= __shfl_xor_sync(0xffffffffu, xa_, o); \
					xb_ |= __shfl_xor_sync(0xffffffffu, xb_, o); \
				} \
				const int wd_ = (ch) * 4 + (lane >> 3); \
				if ((lane & 7) == 0 && wd_ < nW) { faW[wd_] = xa_; fbW[wd_] = xb_; } }
			int ch0 = warp;
			for (; ch0 + (HU_LIT_VEC - 1) * nWarps < nCh; ch0 += nWarps * HU_LIT_VEC) {
				uint4 vv[HU_LIT_VEC];
				ldKey4x8(keys + (size_t) ch0 * 128 + 4 * lane, vv);
#pragma unroll
				for (int u = 0; u < HU_LIT_VEC; ++u) LIT_FLAG_CHUNK(ch0 + u * nWarps, vv[u])
			}
			for (; ch0 < nCh; ch0 += nWarps) {
				const uint4 v1 = ldKey4Now(keys + (size_t) ch0 * 128 + 4 * lane);
				LIT_FLAG_CHUNK(ch0, v1)
			}
#undef LIT_FLAG_CHUNK
			}
			if (pv != 0xFFFFFFFFu) mx = max(mx, pv);
			__syncthreads();
			LITP_MARK(0);
			const int seg = (nW + HU_LIT_THREADS - 1) / HU_LIT_THREADS;
			const int w0 = tid * seg, w1 = min(nW, w0 + seg);
			uint32_t sa = 0, sb = 0;
			for (int wd = w0; wd < w1; ++wd) { sa += __popc(faW[wd]); sb += __popc(fbW[wd]); }
			uint32_t xa = sa, xb = sb;
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
				if (lane >= o) { xa += ya; xb += yb; }
			}
			if (lane == 31) { red[warp][0] = xa; red[warp][1] = xb; }
			__syncthreads();
			uint32_t oa = 0, ob = 0;
			for (int k = 0; k < warp; ++k) { oa += red[k][0]; ob += red[k][1]; }
			uint32_t ra = oa + xa - sa, rb = ob + xb - sb;
			for (int wd = w0; wd < w1; ++wd) { preA[wd] = ra; preB[wd] = rb; ra += __popc(faW[wd]); rb += __popc(fbW[wd]); }
			if (tid == HU_LIT_THREADS - 1) { sTot[0] = oa + xa; sTot[1] = ob + xb; }
			if (tid == 0) sK = 0;
			__syncthreads();
			const uint32_t totA = sTot[0], totB = sTot[1];
			const uint32_t myK = litCountSwaps(faW, fbW, preA, preB, w0, w1, totB);
			if (myK) atomicAdd(&sK, myK);
			__syncthreads();
			const uint32_t K = sK;
			if (tid == 0) {
				const int AK = K < totA ? base + litSelect(faW, preA, nW, K) : n;
				int cut = AK;
				if (K >= 1) cut = min(AK, base + litSelect(fbW, preB, nW, totB - K));
				sCut = cut;
			}
			__syncthreads();
			LITP_MARK(1);
			const int cut = sCut;
			if (cut >= want) {
				// materialise [0, cut): a left stop of rank k takes the k-th right stop (from the right), the rest stay
				if (tid == 0) w[0] = ((lit_t) pv << 32) | (uint32_t) mS;
				// (a) the K right stops that move, listed by their rank from the right: posB[k] = bit index of B_k.  The
				//     list lives in the part of this CTA's scratch row that level 1 never fills (K <= n - cut).
				uint32_t* posB = reinterpret_cast<uint32_t*>(w + cut);
				for (int wd = nW - 1 - tid; wd >= 0; wd -= HU_LIT_THREADS) {
					uint32_t bits = fbW[wd];
					uint32_t rr = totB - preB[wd] - __popc(bits); // rank from the right of the word's highest set bit
					if (rr >= K) continue;                        // uniform enough: words left of B_{K-1} carry nothing
					while (bits) {
						const int bit = 31 - __clz(bits);
						bits &= ~(1u << bit);
						if (rr < K) posB[rr] = (uint32_t) (wd * 32 + bit);
						rr++;
					}
				}
				__syncthreads();
				// (b) rows of 32 positions, lane = position: what stays is copied, a left stop of rank k takes B_k.
				//     Consecutive lanes read consecutive list entries and neighbouring keys, and write one 256-byte line.
				const int nWc = (cut + 31) / 32;
				//     The list entries of the next batch of rows are requested before the keys of this one.
				int pos[HU_LIT_ROWS]; // > 0: stays there; < 0: a left stop, takes the right stop in pk
				uint32_t pk[HU_LIT_ROWS];
				uint32_t lp[HU_LIT_ROWS];
#define LIT_ROW_PLAN(row, u) { \
					const int wd_ = min((row), nWc - 1); \
					const uint32_t fa_ = faW[wd_]; \
					pos[u] = base + wd_ * 32 + lane; \
					const bool mine_ = pos[u] < cut && ((fa_ >> lane) & 1u); \
					lp[u] = mine_ ? 2u * (uint32_t) cut + preA[wd_] + __popc(fa_ & ((1u << lane) - 1u)) : 0u; /* word offset from w; posB = w + cut */ \
					if (mine_) pos[u] = -pos[u]; }
#pragma unroll
				for (int u = 0; u < HU_LIT_ROWS; ++u) LIT_ROW_PLAN(warp + u * nWarps, u)
#pragma unroll
				for (int u = 0; u < HU_LIT_ROWS; ++u) pk[u] = ldKeyNow(reinterpret_cast<const uint32_t*>(w) + lp[u]);
				for (int wdq = warp; wdq < nWc; wdq += nWarps * HU_LIT_ROWS) {
					int srcp[HU_LIT_ROWS];
#pragma unroll
					for (int u = 0; u < HU_LIT_ROWS; ++u) srcp[u] = pos[u] > 0 ? pos[u] : base + (int) pk[u];
#pragma unroll
					for (int u = 0; u < HU_LIT_ROWS; ++u) LIT_ROW_PLAN(wdq + nWarps * HU_LIT_ROWS + u * nWarps, u)
#pragma unroll
					for (int u = 0; u < HU_LIT_ROWS; ++u) pk[u] = ldKeyNow(reinterpret_cast<const uint32_t*>(w) + lp[u]);
					uint32_t kv[HU_LIT_ROWS];
#pragma unroll
					for (int u = 0; u < HU_LIT_ROWS; ++u) kv[u] = ldKeyNow(keys + min(srcp[u], n - 1));
#pragma unroll
					for (int u = 0; u < HU_LIT_ROWS; ++u) {
						const int wd = wdq + u * nWarps;
						const int pp = base + wd * 32 + lane;
						if (wd < nWc && pp < cut && pp > 0) { // position 0 holds the pivot
							const int pq = srcp[u];
							w[pp] = ((lit_t) (pq == mS ? k0v : kv[u]) << 32) | (uint32_t) VPOS(pq);
						}
					}
				}
#undef LIT_ROW_PLAN
				if (tid == 0) {
					stack[0][0] = 0; stack[0][1] = cut; stack[0][2] = depth0 - 1;
					sTop = 1;
					sBound = cut;
				}
				started = true;
			} else {
				// tiny cut: fall back to the in-place path on a full copy
				for (int i = tid; i < n; i += HU_LIT_THREADS) w[i] = ((lit_t) keys[i] << 32) | (uint32_t) i;
			}
#undef VKEY
#undef VPOS
		}
		__syncthreads();
		LITP_MARK(2);
		atomicMax(&sMaxQ, mx);
		if (!started && tid == 0) {
			stack[0][0] = 0; stack[0][1] = n; stack[0][2] = depth0;
			sTop = 1;
		}
		__syncthreads();
		while (sTop > 0) {
			__syncthreads();
			if (tid == 0) { sTop--; sLo = stack[sTop][0]; sHi = stack[sTop][1]; sDepth = stack[sTop][2]; }
			__syncthreads();
			// __introsort_loop on [sLo, sHi); sub-ranges beyond the wanted prefix are dropped
			while (sHi - sLo > 16 && sLo < want) {
				const int lo = sLo, hi = sHi;
				if (sDepth == 0) {
					// depth budget (2 lg n) spent: __introsort_loop finishes the range with std::__partial_sort(first, last, last),
					// i.e. make_heap + sort_heap.  Replayed literally by one thread (hu_stdsort.cuh); a median-of-3 adversary is
					// needed to get here, so speed does not matter, the element order does.
					if (tid == 0) { ssHeapSort(wp + lo, hi - lo, LitLess()); sOverflow = 1; }
					__syncthreads();
					break;
				}
				__syncthreads();
				if (!inSm && sTop == 0 && hi - lo <= HU_LIT_SMEM) {
					// nothing is pending outside [lo, hi): every later range lies inside it, and it fits in shared memory.
					// From here on no level waits for global memory.
					for (int i = tid; i < hi - lo; i += HU_LIT_THREADS) sdata[i] = w[lo + i];
					faW = smLit; fbW = faW + HU_LIT_SMEM / 32; preA = fbW + HU_LIT_SMEM / 32; preB = preA + HU_LIT_SMEM / 32;
					wp = sdata - lo;
					inSm = true; smLo = lo; smHi = hi;
					__syncthreads();
				}
				if (inSm && hi - lo <= HU_LIT_SMALL) {
					if (warp == 0) litWarpIntrosort(wp + lo, wsc, hi - lo, sDepth, lo, want, &sOverflow);
					if (tid == 0) sHi = lo; // range finished
					__syncthreads();
					LITP_MARK(6);
					LITP_COUNT(10);
					continue;
				}
				if (hi - lo <= HU_LIT_SMALL) {
					// the rest of this range, every level of it, by one warp on a shared-memory copy
					const int m = hi - lo;
					for (int i = tid; i < m; i += HU_LIT_THREADS) sbuf[i] = w[lo + i];
					__syncthreads();
					if (warp == 0) litWarpIntrosort(sbuf, wsc, m, sDepth, lo, want, &sOverflow);
					if (tid == 0) sHi = lo; // range finished
					__syncthreads();
					for (int i = tid; i < m; i += HU_LIT_THREADS) w[lo + i] = sbuf[i];
					__syncthreads();
					LITP_MARK(6);
					LITP_COUNT(10);
					continue;
				}
				// __move_median_to_first(lo, lo+1, mid, hi-1): every thread reads the three candidates and w[lo] itself
				// (no hand-over through one thread and a barrier); the exchange of w[lo] and w[mS] is applied to the
				// flag pass virtually and written by thread 0 on the side, well before the swap phase reads it
				int mS;
				uint32_t pv, k0v;
				lit_t wmKeep, w0Keep;
				{
					const int ia = lo + 1, ib = lo + (hi - lo) / 2, ic = hi - 1;
					const lit_t wa = wp[ia], wb = wp[ib], wc = wp[ic], w0 = wp[lo];
					const uint32_t qa = LIT_KEY(wa), qb = LIT_KEY(wb), qc = LIT_KEY(wc);
					if (qa < qb) mS = qb < qc ? ib : (qa < qc ? ic : ia);
					else mS = qa < qc ? ia : (qb < qc ? ic : ib);
					const lit_t wm = mS == ia ? wa : (mS == ib ? wb : wc);
					pv = LIT_KEY(wm);
					k0v = LIT_KEY(w0);
					wmKeep = wm; w0Keep = w0;
				}
				const int base = lo + 1, cnt = hi - base, nW = (cnt + 31) / 32;
#define LIT_FLAG_ROW(wd, val) { \
					const int t_ = (wd) * 32 + lane; \
					const uint32_t v_ = base + t_ == mS ? k0v : (val); \
					const uint32_t ba_ = __ballot_sync(0xffffffffu, t_ < cnt && !(v_ < pv)); \
					const uint32_t bb_ = __ballot_sync(0xffffffffu, t_ < cnt && !(pv < v_)); \
					if (lane == 0) { faW[wd] = ba_; fbW[wd] = bb_; } }
				int wd0 = warp;
				for (; wd0 + (HU_LIT_UNROLL - 1) * nWarps < nW - 1; wd0 += nWarps * HU_LIT_UNROLL) { // every row of the batch is whole
					uint32_t vv[HU_LIT_UNROLL];
					ldPairKeyx8(wp + base + wd0 * 32 + lane, vv);
#pragma unroll
					for (int u = 0; u < HU_LIT_UNROLL; ++u) LIT_FLAG_ROW(wd0 + u * nWarps, vv[u])
				}
				for (; wd0 < nW; wd0 += nWarps) {
					const uint32_t v1 = ldPairKeyNow(wp + min(base + wd0 * 32 + lane, hi - 1));
					LIT_FLAG_ROW(wd0, v1)
				}
#undef LIT_FLAG_ROW
				__syncthreads();
				LITP_MARK(3);
				LITP_COUNT(9);
				if (tid == 0) { wp[lo] = wmKeep; wp[mS] = w0Keep; } // every thread has read the four elements by now
				// exclusive prefix counts per word
				const int seg = (nW + HU_LIT_THREADS - 1) / HU_LIT_THREADS;
				const int w0 = tid * seg, w1 = min(nW, w0 + seg);
				uint32_t sa = 0, sb = 0;
				for (int wd = w0; wd < w1; ++wd) { sa += __popc(faW[wd]); sb += __popc(fbW[wd]); }
				uint32_t xa = sa, xb = sb; // inclusive scan over threads
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
					if (lane >= o) { xa += ya; xb += yb; }
				}
				if (lane == 31) { red[warp][0] = xa; red[warp][1] = xb; }
				__syncthreads();
				uint32_t oa = 0, ob = 0;
				for (int k = 0; k < warp; ++k) { oa += red[k][0]; ob += red[k][1]; }
				uint32_t ra = oa + xa - sa, rb = ob + xb - sb;
				for (int wd = w0; wd < w1; ++wd) { preA[wd] = ra; preB[wd] = rb; ra += __popc(faW[wd]); rb += __popc(fbW[wd]); }
				if (tid == HU_LIT_THREADS - 1) { sTot[0] = oa + xa; sTot[1] = ob + xb; }
				if (tid == 0) sK = 0;
				__syncthreads();
				const uint32_t totA = sTot[0], totB = sTot[1];
				// K = number of swaps: the k-th left stop (k-th element >= pivot) lies left of the k-th right stop
				const uint32_t myK = litCountSwaps(faW, fbW, preA, preB, w0, w1, totB);
				if (myK) atomicAdd(&sK, myK);
				__syncthreads();
				const uint32_t K = sK;
				if (tid == 0) {
					const int AK = K < totA ? base + litSelect(faW, preA, nW, K) : hi;
					int cut = AK;
					if (K >= 1) cut = min(AK, base + litSelect(fbW, preB, nW, totB - K));
					sCut = cut;
				}
				__syncthreads();
				LITP_MARK(4);
				const bool full = sCut < want; // the right part is needed too: swap instead of copy
				// swap k: k-th left stop <-> k-th right stop.  Each thread takes a contiguous run of k and walks the
				// two bitmaps from its first pair (fa forwards, fb backwards).
				{
					const uint32_t per = (K + HU_LIT_THREADS - 1) / HU_LIT_THREADS;
					const uint32_t k0 = (uint32_t) tid * per, k1 = min(K, k0 + per);
					if (k0 < k1) {
						int pa = litSelect(faW, preA, nW, k0);            // bit index of the k0-th left stop
						int pb = litSelect(fbW, preB, nW, totB - 1 - k0); // bit index of the k0-th right stop (from the right)
						for (uint32_t k = k0; k < k1; k += HU_LIT_SWAPS) {
							// positions of a batch first (bitmap walk in shared memory), then every load, then every store: the
							// pairs are disjoint, but the compiler cannot know and would wait for each swap in turn
							int ppv[HU_LIT_SWAPS], pqv[HU_LIT_SWAPS];
#pragma unroll
							for (int u = 0; u < HU_LIT_SWAPS; ++u) {
								ppv[u] = base + pa; pqv[u] = base + pb;
								if (k + u + 1 < k1) {
									// next set bit of fa after pa
									int wd = pa >> 5;
									uint32_t bits = faW[wd] & ~(0xFFFFFFFFu >> (31 - (pa & 31)));
									while (!bits) bits = faW[++wd];
									pa = wd * 32 + __ffs(bits) - 1;
									// previous set bit of fb before pb
									wd = pb >> 5;
									bits = fbW[wd] & ((1u << (pb & 31)) - 1u);
									while (!bits) bits = fbW[--wd];
									pb = wd * 32 + 31 - __clz(bits);
								}
							}
							if (full) { // rare (the cut fell inside the wanted prefix): plain swaps
								for (int u = 0; u < HU_LIT_SWAPS; ++u)
									if (k + u < k1) litSwap(wp, ppv[u], pqv[u]);
								continue;
							}
							lit_t vq[HU_LIT_SWAPS];
							ldPairx4(wp + pqv[0], wp + pqv[1], wp + pqv[2], wp + pqv[3], vq); // past k1: the last pair again
#pragma unroll
							for (int u = 0; u < HU_LIT_SWAPS; ++u)
								if (k + u < k1) wp[ppv[u]] = vq[u];
						}
					}
				}
				__syncthreads();
				if (tid == 0) {
					sDepth--;
					if (sCut < want && sHi - sCut > 1 && sTop < 64) { stack[sTop][0] = sCut; stack[sTop][1] = sHi; stack[sTop][2] = sDepth; sTop++; }
					if (sCut >= want) sBound = min(sBound, sCut); // [sCut, sHi) was left stale (copies, not swaps)
					sHi = sCut;
				}
				__syncthreads();
				LITP_MARK(5);
			}
			__syncthreads();
		}
		LITP_MARK(11);
		// __final_insertion_sort restricted to what can reach the first `want` positions: blocks are at most 16
		// long and ordered among themselves, and an insertion never crosses a block boundary.  Nothing at or beyond a
		// dropped cut may be touched.
		const int P = min(min(n, sBound), want + 17);
		for (int i = tid; i < P; i += HU_LIT_THREADS) sbuf[i] = (inSm && i >= smLo && i < smHi) ? wp[i] : w[i];
		__syncthreads();
		{
			// insertion sort is a stable sort: every element goes straight to (# smaller keys) + (# equal keys before it)
			lit_t mine = 0;
			int rank = 0;
			if (tid < P) {
				mine = sbuf[tid];
				const uint32_t k = LIT_KEY(mine);
				for (int j = 0; j < P; ++j) {
					const uint32_t kj = LIT_KEY(sbuf[j]);
					rank += (kj < k) || (kj == k && j < tid);
				}
			}
			__syncthreads();
			if (tid < P) sbuf[rank] = mine;
			__syncthreads();
		}
		if (tid == 0) {
			int nOut = min(want, P);
			while (nOut > 0 && LIT_KEY(sbuf[nOut - 1]) == 0xFFFFFFFFu) nOut--; // undefined distances (0/0) are not seeds
			if (a.maxDiff < INFINITY && nOut > 0) { // src/HmmUFOtu_main.cpp:141-150, on the key scale
				const double best = (double) LIT_KEY(sbuf[0]) / 2147483648.0, worst = (double) sMaxQ / 2147483648.0;
				if (worst < best + a.maxDiff) {
					int keep = 0;
					while (keep < nOut && (double) LIT_KEY(sbuf[keep]) / 2147483648.0 - best <= a.maxDiff) keep++;
					nOut = keep;
				}
			}
			sCut = nOut;
			a.nSeeds[r] = nOut;
			if (sOverflow && a.heapFallbacks) atomicAdd(a.heapFallbacks, 1); // diagnostic: reads whose replay took the heap-sort path
		}
		__syncthreads();
		if (tid < sCut) a.seedIds[(size_t) r * HU_MAX_SEED + tid] = nodeOfPos[(uint32_t) sbuf[tid]];
	}
}

#define HU_LIT_SMEM_MAX (200 * 1024)

// per-device kernel attributes (hu_db_create calls this with the device current)
int configSeedKernels() {
	HU_CUDA(cudaFuncSetAttribute(seedScanKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	HU_CUDA(cudaFuncSetAttribute(seedScanKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	HU_CUDA(cudaFuncSetAttribute(seedSortLiteralKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_LIT_SMEM_MAX));
	return HU_OK;
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
	a.heapFallbacks = b->workCount + 4;
	a.litFromScan = b->workCount + 5;
	a.litCursor = b->workCount + 6;
	HU_CUDA(cudaMemsetAsync(a.heapFallbacks, 0, 3 * sizeof(int), b->stream));
	// The key vector of a read is getSeed's candidate vector: the non-root nodes in id order that pass the height filter
	// (src/HmmUFOtu_main.cpp:131-137).  The literal order replays std::sort on exactly that vector, so with -H/--max-height
	// the candidates are compacted (positions = ranks among the passing nodes); the (dist, id) order keeps one position per
	// non-root node and marks the filtered ones excluded.
	const bool literal = p->seedOrder == HU_SEED_ORDER_STDSORT;
	a.posOfSlot = t.posOfSlot; a.nodeOfPos = t.nodeOfPos; a.nCand = t.nCand;
	if (literal && p->maxHeight < INFINITY) {
		if (!b->posOfSlotH || b->heightCached != p->maxHeight) {
			std::vector<int> posOfNode(t.nNodes, -1), posOfSlot(t.scanPitch, -1), nodeOfPos(t.scanPitch, 0x7fffffff);
			int pos = 0;
			for (int x = 0; x < t.nNodes; ++x)
				if (x != t.root && db->heightHost[x] <= p->maxHeight) { posOfNode[x] = pos; nodeOfPos[pos] = x; pos++; }
			for (int sl = 0; sl < t.scanPitch; ++sl)
				if (db->scanNodeHost[sl] >= 0) posOfSlot[sl] = posOfNode[db->scanNodeHost[sl]];
			if (!b->posOfSlotH) {
				void* q = NULL;
				HU_CUDA(cudaMalloc(&q, (size_t) 2 * t.scanPitch * sizeof(int)));
				b->allocs.push_back(q);
				b->posOfSlotH = (int*) q;
				b->nodeOfPosH = b->posOfSlotH + t.scanPitch;
			}
			HU_CUDA(cudaStreamSynchronize(b->stream)); // a previous run may still read the old map
			HU_CUDA(cudaMemcpy(b->posOfSlotH, posOfSlot.data(), (size_t) t.scanPitch * sizeof(int), cudaMemcpyHostToDevice));
			HU_CUDA(cudaMemcpy(b->nodeOfPosH, nodeOfPos.data(), (size_t) t.scanPitch * sizeof(int), cudaMemcpyHostToDevice));
			b->nCandH = pos;
			b->heightCached = p->maxHeight;
		}
		a.posOfSlot = b->posOfSlotH; a.nodeOfPos = b->nodeOfPosH; a.nCand = b->nCandH;
	}
	// Level 1 of the literal order's flag pass rides in the tensor-core scan's epilogue (hu_seedmm.cu) when the key vector is the
	// plain candidate vector, large enough for the direct level 1, and the largest key is not needed (no --max-diff):
	// HU_LIT_BITS=0 keeps the pass in seedSortLiteralKernel (A/B measurements, parity tests between the two).
	a.litPv = NULL; a.litOk = NULL; a.litBitA = a.litBitB = NULL; a.litStride = 0;
	{
		const char* e = getenv("HU_LIT_BITS");
		const bool want = !(e && !strcmp(e, "0"));
		if (want && literal && t.mmA && b->mmB && a.posOfSlot == t.posOfSlot && !(p->maxDiff < INFINITY) && a.nCand > 4 * HU_LIT_SMALL &&
		    t.scanPitch % 128 == 0) {
			const int stride = (b->maxReads + 31) / 32 * 32;
			const size_t words = (size_t) (t.scanPitch / 32) * stride;
			if (!b->litBits) {
				void* q = NULL;
				HU_CUDA(cudaMalloc(&q, (2 * words + 2 * (size_t) stride) * sizeof(uint32_t)));
				b->allocs.push_back(q);
				b->litBits = (uint32_t*) q;
			}
			a.litBitA = b->litBits; a.litBitB = b->litBits + words;
			a.litPv = b->litBits + 2 * words; a.litOk = (int*) (b->litBits + 2 * words + stride);
			a.litStride = stride;
		}
	}
	b->seed = a;
	HU_CUDA(cudaEventRecord(b->ev[2], b->stream));
	HU_CUDA(cudaMemsetAsync(a.rValid, 0, (size_t) a.nReadsPad * sizeof(int), b->stream));
	HU_CUDA(cudaMemsetAsync(a.rWordLo, 0x7f, (size_t) a.nReadsPad * sizeof(int), b->stream));
	HU_CUDA(cudaMemsetAsync(a.rWordHi, 0xff, (size_t) a.nReadsPad * sizeof(int), b->stream));
	packReadsKernel<<<dim3((a.nReadsPad + 127) / 128, t.nPWords), 128, 0, b->stream>>>(a, t.csLen, db->hmm.K, db->hmm.p2cs);
	sparseListKernel<<<(b->nReads + 3) / 4, 128, 0, b->stream>>>(a, t.csLen, t.colK);
	// the scan: on the tensor cores (hu_seedmm.cu) when the key vector is the plain candidate vector; the bit-plane kernel for
	// the compacted vector of the literal order under -H, for profiles the packed accumulator cannot hold, or on request
	static int scanPref = -1; // HU_SEED_SCAN=popc forces the bit-plane kernel (A/B measurements, parity tests between the two)
	{ const char* e = getenv("HU_SEED_SCAN"); scanPref = (e && !strcmp(e, "popc")) ? 1 : 0; }
	int scanLaunches = 0;
	if (t.mmA && b->mmB && a.posOfSlot == t.posOfSlot && scanPref == 0) {
		int rcM = launchSeedMm(b, a, &scanLaunches);
		if (rcM != HU_OK) return rcM;
		b->mmScanRan = 1;
	} else {
		b->mmScanRan = 0;
		const int groups = a.nReadsPad / HU_SCAN_READS;
		const size_t smem = (size_t) t.nPWords * HU_SCAN_READS * sizeof(uint4);
		if (smem > 200 * 1024) { setError("profile too long for the seed-scan read tile (%d words)", t.nPWords); return HU_ERR_ARG; }
		const int gapTiles = t.nGapSlots / HU_SCAN_TILE, allTiles = t.scanPitch / HU_SCAN_TILE;
		// grid.x (fast) runs over read groups so that consecutive CTAs reuse the same node tile from L2
		if (gapTiles > 0) seedScanKernel<true><<<dim3(groups, gapTiles), 128, smem, b->stream>>>(a, t, 0);
		if (allTiles > gapTiles) seedScanKernel<false><<<dim3(groups, allTiles - gapTiles), 128, smem, b->stream>>>(a, t, gapTiles);
		scanLaunches = (gapTiles > 0) + (allTiles > gapTiles);
	}
	HU_CUDA(cudaEventRecord(b->ev[3], b->stream));
	if (getenv("HU_LIT_DRY")) a.litBitA = NULL; // measurements: the epilogue made the flags but did not store them
	if (literal) {
		int sms = 0;
		HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
		const int maxWords = (t.nCand + 31) / 32 + 1;
		// the four bitmaps of the CTA-wide levels, or (shared-memory mode) four small bitmaps and the range itself
		const size_t smemLit = std::max((size_t) 4 * maxWords * sizeof(uint32_t),
		                                (size_t) 4 * (HU_LIT_SMEM / 32) * sizeof(uint32_t) + (size_t) std::min(t.nCand, HU_LIT_SMEM) * sizeof(lit_t));
		if (smemLit > HU_LIT_SMEM_MAX) { setError("tree too large for the shared-memory bitmaps of the literal seed order (%d nodes; limit about %d)", t.nNodes, (int) (HU_LIT_SMEM_MAX / 16 * 32)); return HU_ERR_ARG; }
		static int litCtas = 0; // resident CTAs per SM: 2 fills the SM's shared memory; 1 leaves room for another stream's kernels
		if (!litCtas) { const char* e = getenv("HU_LIT_CTAS_PER_SM"); litCtas = e ? std::max(1, std::min(2, atoi(e))) : 2; }
		const int grid = std::min(b->nReads, sms * litCtas);
		if (!b->litIds) {
			void* q = NULL;
			HU_CUDA(cudaMalloc(&q, (size_t) sms * 2 * t.scanPitch * sizeof(unsigned long long)));
			b->allocs.push_back(q);
			b->litIds = (unsigned long long*) q;
		}
		seedSortLiteralKernel<<<grid, HU_LIT_THREADS, smemLit, b->stream>>>(a, a.nCand, t.scanPitch, a.nodeOfPos, b->litIds, maxWords);
	} else
		seedSelectKernel<<<b->nReads, HU_SEL_THREADS, 0, b->stream>>>(a, t.scanPitch, t.nodeOfPos);
	HU_CUDA(cudaEventRecord(b->ev[4], b->stream));
	b->launches += 3 + scanLaunches;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu

// ---- diagnostics (tests): the literal-order kernel on caller-supplied key rows, without a database ----
extern "C" int hu_debug_literal_order(int device, const uint32_t* keys, int32_t nRows, int32_t n, int32_t want, int32_t* outPos, int32_t* nOut,
		int32_t* heapFallbacks) {
	using namespace hu;
	if (!keys || !outPos || !nOut || nRows < 1 || n < 1 || want < 1 || want > HU_MAX_SEED) { setError("hu_debug_literal_order: bad arguments"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(device));
	if (configSeedKernels() != HU_OK) return HU_ERR_CUDA;
	const int pitch = (n + 511) / 512 * 512;
	uint32_t* dKeys = NULL;
	int *dWin = NULL, *dIds = NULL, *dN = NULL, *dPos = NULL, *dHeap = NULL;
	lit_t* dScratch = NULL;
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
	const int grid = std::min((int) nRows, sms * 2);
	HU_CUDA(cudaMalloc(&dKeys, (size_t) nRows * pitch * sizeof(uint32_t)));
	HU_CUDA(cudaMemset(dKeys, 0xff, (size_t) nRows * pitch * sizeof(uint32_t)));
	HU_CUDA(cudaMemcpy2D(dKeys, (size_t) pitch * sizeof(uint32_t), keys, (size_t) n * sizeof(uint32_t), (size_t) n * sizeof(uint32_t), nRows, cudaMemcpyHostToDevice));
	HU_CUDA(cudaMalloc(&dWin, (size_t) nRows * sizeof(int)));
	HU_CUDA(cudaMemset(dWin, 0, (size_t) nRows * sizeof(int)));
	HU_CUDA(cudaMalloc(&dIds, (size_t) nRows * HU_MAX_SEED * sizeof(int)));
	HU_CUDA(cudaMalloc(&dN, (size_t) nRows * sizeof(int)));
	HU_CUDA(cudaMalloc(&dHeap, sizeof(int)));
	HU_CUDA(cudaMemset(dHeap, 0, sizeof(int)));
	HU_CUDA(cudaMalloc(&dScratch, (size_t) grid * pitch * sizeof(lit_t)));
	std::vector<int> ident(pitch);
	for (int i = 0; i < pitch; ++i) ident[i] = i;
	HU_CUDA(cudaMalloc(&dPos, (size_t) pitch * sizeof(int)));
	HU_CUDA(cudaMemcpy(dPos, ident.data(), (size_t) pitch * sizeof(int), cudaMemcpyHostToDevice));
	SeedArgs a;
	memset(&a, 0, sizeof a);
	a.nReads = nRows; a.maxNSeed = want; a.maxDiff = INFINITY; a.maxHeight = INFINITY;
	a.winStart = dWin; a.winEnd = dWin; a.keys = dKeys; a.seedIds = dIds; a.nSeeds = dN; a.heapFallbacks = dHeap;
	a.nodeOfPos = dPos; a.nCand = n;
	const int maxWords = (n + 31) / 32 + 1;
	const size_t smemLit = std::max((size_t) 4 * maxWords * sizeof(uint32_t),
	                                (size_t) 4 * (HU_LIT_SMEM / 32) * sizeof(uint32_t) + (size_t) std::min((int) n, HU_LIT_SMEM) * sizeof(lit_t));
	int rc = HU_OK;
	if (smemLit > HU_LIT_SMEM_MAX) { setError("hu_debug_literal_order: n too large"); rc = HU_ERR_ARG; }
	else {
		seedSortLiteralKernel<<<grid, HU_LIT_THREADS, smemLit>>>(a, n, pitch, dPos, dScratch, maxWords);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { setError("hu_debug_literal_order: %s", cudaGetErrorString(e)); rc = HU_ERR_CUDA; }
	}
	if (rc == HU_OK) {
		std::vector<int> ids((size_t) nRows * HU_MAX_SEED);
		cudaMemcpy(ids.data(), dIds, ids.size() * sizeof(int), cudaMemcpyDeviceToHost);
		cudaMemcpy(nOut, dN, (size_t) nRows * sizeof(int), cudaMemcpyDeviceToHost);
		if (heapFallbacks) cudaMemcpy(heapFallbacks, dHeap, sizeof(int), cudaMemcpyDeviceToHost);
		for (int r = 0; r < nRows; ++r)
			for (int k = 0; k < want; ++k) outPos[(size_t) r * want + k] = k < nOut[r] ? ids[(size_t) r * HU_MAX_SEED + k] : -1;
	}
	cudaFree(dKeys); cudaFree(dWin); cudaFree(dIds); cudaFree(dN); cudaFree(dHeap); cudaFree(dScratch); cudaFree(dPos);
	return rc;
}
