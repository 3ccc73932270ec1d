// K2, tensor-core form: the all-node p-distance scan of getSeed() (src/HmmUFOtu_main.cpp:127-139, src/SeqUtils.cpp:37-54)
// as an int8 contraction on the 5th-generation tensor cores (hu_mm.cuh), with the seed key made in the epilogue.
//
//   database creation : node sequences -> operand A images (rows = getSeed's candidate positions, i.e. the non-root nodes in
//                       id order; K = 4 bytes per profile match column) + one extension chunk per tile (below)
//   per batch         : aligned reads -> operand B images (the read's bases inside its window), chunk range per read tile
//   seedMmKernel      : S = A . B^T per (position, read) -> d, N -> key floor(d 2^31 / N), written in candidate order:
//                       a warp's 32 lanes hold 32 consecutive positions of one read = one 128-byte line per store;
//                       with the literal seed order (std::sort replay, hu_seed.cu) also the first flag pass of that replay:
//                       per read the bitmaps (key >= pivot), (key <= pivot) over the candidate positions (KeyEpi<true>)
//
// Read bases in NON-match columns (insertions buildGlobalAlign could place; pDist counts them like any column).  They are
// few per read (about one for a 250-base amplicon read) but the node side is dense there for inferred sequences, which have
// a base in every column.  Three tiers, all exact:
//   1. the default.  Nodes are split by class: 1 = the sequence has no gap at all (inferred, internal nodes), 0 = the others
//      (leaves: a non-match column is a gap for nearly all of them).  Per non-match column c the class-1 nodes have a majority
//      base g_c.  With every class-1 node at its default and every class-0 node at a gap, an entry (c, b) of a read adds
//      class * (16384 + (b == g_c ? 3 : -1)) to S: rank one per entry, so it rides in the contraction as an extension
//      chunk: A row = class * (-128, 1) x 64, B row = (-128, +3 | -1) per entry.  No epilogue work.
//   2. the exceptions.  Per column the positions that deviate from their class default are listed at database creation
//      (a fraction of a per cent).  For the reads of the batch that have an entry in a column, the (exception position, read)
//      pairs are recomputed from the bit planes and overwrite the tile kernel's key (seedMmExceptionKernel).
//   3. reads with an entry in a column whose exception list is long (a non-match column that many leaves occupy), or with more
//      than 64 entries: the epilogue leaves (d << 16 | N) and seedMmFixupKernel adds the columns from the transposed sequences.
// Keys are bit-identical to seedScanKernel's (tests/test_gpu_mm.py); that kernel remains for profiles longer than
// HU_MM_MAX_N - 64 match columns and for the compacted candidate vector of the literal seed order under -H/--max-height.
#include <stdlib.h>
#include <string.h>
#include "hu_batch.cuh"
#include "hu_mm.cuh"

namespace hu {

#define HU_KEY_BIAS 9.5367431640625e-07 // 2^-20 (see makeKey in hu_seed.cu)
__device__ __forceinline__ uint32_t mmMakeKey(int d, int N, bool excluded, const double* __restrict__ inv2N) {
	const double q = __dadd_rn(__dmul_rn((double) d, __ldg(inv2N + N)), HU_KEY_BIAS);
	return (excluded || N <= 0) ? 0xFFFFFFFFu : __double2uint_rd(q);
}

// ---- operand A: one thread per (position, chunk); chunk `stride - 1` is the extension chunk ----
__global__ void __launch_bounds__(HU_MM_BM) mmPackNodesKernel(const int8_t* __restrict__ seq, int csLen, int K, const int* __restrict__ p2cs,
		const int* __restrict__ posNode, const int8_t* __restrict__ posClass, int stride, int8_t* __restrict__ A) {
	const int mt = blockIdx.x, kc = blockIdx.y, row = threadIdx.x;
	const int p = mt * HU_MM_BM + row;
	const int node = __ldg(posNode + p);
	uint32_t w[32];
	if (kc == stride - 1) {
		const uint32_t v = (node >= 0 && posClass[p]) ? 0x01800180u : 0u; // (-128, 1, -128, 1)
#pragma unroll
		for (int i = 0; i < 32; ++i) w[i] = v;
	} else {
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			const int k = kc * 32 + i + 1;
			w[i] = (node >= 0 && k <= K) ? mmBaseWord(seq[(size_t) node * csLen + __ldg(p2cs + k) - 1]) : 0u;
		}
	}
	uint8_t* img = reinterpret_cast<uint8_t*>(A) + ((size_t) mt * stride + kc) * HU_MM_A_BYTES;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		*reinterpret_cast<uint4*>(img + mmTileOffset(row, c * 16)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// ---- operand B: one thread per (read, chunk); rows past the batch and reads without a window are zero ----
__global__ void __launch_bounds__(128) mmPackReadsKernel(SeedArgs a, int L, int K, const int* __restrict__ p2cs, int stride, int8_t* __restrict__ B) {
	const int r = blockIdx.x * 128 + threadIdx.x, kc = blockIdx.y;
	const bool ok = r < a.nReads && a.winStart[r] >= 0;
	const int s = ok ? a.winStart[r] : 0, e = ok ? a.winEnd[r] : -1;
	const int8_t* x = a.aligned + (size_t) (ok ? r : 0) * L;
	uint32_t w[32];
#pragma unroll
	for (int i = 0; i < 32; ++i) {
		const int k = kc * 32 + i + 1;
		uint32_t v = 0u;
		if (ok && k <= K) {
			const int c = __ldg(p2cs + k) - 1;
			if (c >= s && c <= e) v = mmBaseWord(x[c]);
		}
		w[i] = v;
	}
	const int nt = r / HU_MM_BN, row = r - nt * HU_MM_BN;
	uint8_t* img = reinterpret_cast<uint8_t*>(B) + ((size_t) nt * stride + kc) * HU_MM_B_BYTES;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		*reinterpret_cast<uint4*>(img + mmTileOffset(row, c * 16)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// ---- the extension chunk of operand B and the read's tier; one thread per read row ----
// mode: 0 no entry, 1 entries in the extension chunk (tiers 1 + 2), 2 raw (tier 3), -1 sparse list overflow (read unplaced)
__global__ void __launch_bounds__(128) mmPackExtKernel(SeedArgs a, TreeDev t, int stride, int8_t* __restrict__ B, int* __restrict__ mode,
		int* __restrict__ tileExt, int* __restrict__ colCount) {
	const int r = blockIdx.x * 128 + threadIdx.x;
	const int nt = r / HU_MM_BN, row = r - nt * HU_MM_BN;
	const int ns = (r < a.nReads && a.winStart[r] >= 0) ? a.nSparse[r] : 0;
	const uint32_t* sp = a.sparse + (size_t) (r < a.nReads ? r : 0) * a.sparseCap;
	int m = ns < 0 ? -1 : (ns == 0 ? 0 : (ns <= 64 ? 1 : 2));
	for (int i = 0; m == 1 && i < ns; ++i)
		if (__ldg(t.mmColDefault + (sp[i] >> 8)) > 3) m = 2; // a column with a long exception list
	uint32_t w[32];
#pragma unroll
	for (int i = 0; i < 32; ++i) w[i] = 0u;
	if (m == 1) {
#pragma unroll
		for (int i = 0; i < 32; ++i) { // two entries per word: (-128, +3 | -1) each
			uint32_t v = 0u;
#pragma unroll
			for (int hlf = 0; hlf < 2; ++hlf) {
				const int e = 2 * i + hlf;
				if (e < ns) {
					const uint32_t ent = sp[e];
					const int col = (int) (ent >> 8), rb = (int) (ent & 0xff);
					const uint32_t pair = 0x80u | ((rb == (int) __ldg(t.mmColDefault + col) ? 0x03u : 0xffu) << 8);
					v |= pair << (16 * hlf);
					if (__ldg(t.mmExcPtr + col + 1) > __ldg(t.mmExcPtr + col)) atomicAdd(colCount + col, 1);
				}
			}
			w[i] = v;
		}
		atomicOr(tileExt + nt, 1);
	}
	if (r < a.nReadsPad) mode[r] = m;
	uint8_t* img = reinterpret_cast<uint8_t*>(B) + ((size_t) nt * stride + (stride - 1)) * HU_MM_B_BYTES;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		*reinterpret_cast<uint4*>(img + mmTileOffset(row, c * 16)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// chunk range of a read tile: the profile words (32 positions = one chunk) any of its reads has a base in
__global__ void __launch_bounds__(HU_MM_BN) mmRangeKernel(SeedArgs a, int2* __restrict__ range) {
	__shared__ int sLo, sHi;
	if (threadIdx.x == 0) { sLo = 0x7fffffff; sHi = -1; }
	__syncthreads();
	const int r = blockIdx.x * HU_MM_BN + threadIdx.x;
	if (r < a.nReads && a.winStart[r] >= 0) {
		const bool any = a.rWordHi[r] >= 0; // a read without a base in a match column still gets its keys (chunk 0 is all zero for it)
		atomicMin(&sLo, any ? a.rWordLo[r] : 0);
		atomicMax(&sHi, any ? a.rWordHi[r] : 0);
	}
	__syncthreads();
	if (threadIdx.x == 0) range[blockIdx.x] = make_int2(sLo, sHi);
}

// position -> 1 if it is not a candidate of this run (padding, or the node lies above --max-height): once per launch, so that the
// tile kernel's epilogue has one byte to load per position instead of a node id and then its height
__global__ void mmPosExclKernel(const int* __restrict__ nodeOfPos, const double* __restrict__ height, double maxHeight, int nNodes, int pitch,
		uint8_t* __restrict__ posExcl) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= pitch) return;
	const int node = __ldg(nodeOfPos + p);
	posExcl[p] = (node < 0 || node >= nNodes || !(__ldg(height + node) <= maxHeight)) ? 1 : 0;
}

// ---- epilogue of seedMmKernel: accumulator -> seed key ----
// Integer form of the key.  2^31 = Q N + R (Q = floor(2^31 / N), R < N), so floor(d 2^31 / N) = d Q + floor(d R / N) with
// d R < N^2 < 2^24; the small quotient is a multiply-high: floor(t / N) = (t m) >> k for m = ceil(2^k / N) whenever
// t (m N - 2^k) < 2^k, which k = max(32, 24 + bits(N)) guarantees for t < N^2 (m < 2^32; N = 1 has R = 0 and takes m = 0).
// One 16-byte table entry per N: {Q, R, m, k - 32}; N = 0 never reads it (S == 0 <=> N == 0 -> excluded).  The same value as
// makeKey()'s correctly rounded floating-point form in hu_seed.cu (both are the exact floor; tests/test_gpu_mm.py).
// 32 x 32 bit matrix across a warp: lane i holds row i, afterwards lane j holds column j (five butterfly exchanges)
__device__ __forceinline__ uint32_t warpBitTranspose(uint32_t x, int lane) {
#pragma unroll
	for (int s = 16; s > 0; s >>= 1) {
		const uint32_t m0 = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u; // bits b with (b & s) == 0
		const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
		x = (lane & s) ? ((x & ~m0) | ((y >> s) & m0)) : ((x & m0) | ((y & m0) << s));
	}
	return x;
}

// LIT: the epilogue also makes level 1 of the literal seed order's flag pass (hu_seed.cu, seedSortLiteralKernel) while the keys
// are in registers: per read the two bitmaps (key >= pivot, key <= pivot) over the candidate positions, the pivot being known
// ahead of the scan (seedLitPivotKernel).  A thread (= position) collects its flags against the 32 reads of the chunk in one
// word per bitmap, the warp transposes the 32 x 32 bits, and lane j stores the word of read j: bit i = position 32 w + i.
// (A ballot per key does the same without the transpose, but 2 votes + 1 shuffle per key made the epilogue the kernel's
// bottleneck: 3.2 -> 5.4 ms.)
template <bool LIT>
struct KeyEpi {
	uint32_t* keys;
	int pitch, nReads, nNodes;
	const int* mode;
	const uint8_t* posExcl; // [pitch] 1: the position is not a candidate of this run (padding, or above --max-height)
	const uint4* keyTab; // [HU_MM_MAX_N + 1]
	const uint32_t* litPv; // LIT: pivot per read
	uint32_t* litBitA;     // LIT: [litStride][pitch / 32]
	uint32_t* litBitB;
	int litStride;
	bool excl; // per thread: this position is not a candidate (padding, or above --max-height)
	bool tilePlain; // per warp and tile: no per-key tests needed
	static constexpr int NCH = HU_MM_EPI_COLS / 32; // 32-read chunks per warp and tile
	static constexpr int NGRP = HU_MM_EPI_WARPS / 4; // groups of four warps (the four TMEM lane quarters) that share a column group
	static constexpr int XCH_WORDS = HU_MM_EPI_COLS * 4 * 2; // per group and buffer: [bitmap][read][lane quarter]
	static constexpr int SMEM_BYTES = LIT ? HU_MM_EPI_WARPS * HU_MM_EPI_COLS * 4 + 2 * NGRP * XCH_WORDS * 4 : 0;
	uint32_t* pvSm;      // LIT: the pivots of the 128 reads this warp handles in the current tile
	uint32_t fa[NCH], fb[NCH]; // LIT: this position's flag words against the 32-read chunks of the tile
	uint32_t* xch;       // LIT: this group's two exchange buffers (the four lane-quarter warps assemble 16-byte bitmap stores)
	int xchFlip;
	__device__ __forceinline__ void init(void* sm, int tid, int) {
		const int ew = max((tid >> 5) - 2, 0); // epilogue warps are warps 2, 3, ...
		pvSm = reinterpret_cast<uint32_t*>(sm) + ew * HU_MM_EPI_COLS;
		xch = reinterpret_cast<uint32_t*>(sm) + HU_MM_EPI_WARPS * HU_MM_EPI_COLS + (ew >> 2) * 2 * XCH_WORDS;
		xchFlip = 0;
	}
	// the per-tile inputs of the next tile, loaded one tile ahead
	uint32_t nxExcl;
	int nxMode[NCH];
	uint32_t nxPv[NCH];
	__device__ __forceinline__ void prefetch(int mt, int nt, int row) {
		const int p = mt * HU_MM_BM + row;
		nxExcl = p < pitch ? (uint32_t) __ldg(posExcl + p) : 1u;
		const int lane = threadIdx.x & 31, h = ((int) (threadIdx.x >> 5) - 2) >> 2;
		const int rb = nt * HU_MM_BN + h * HU_MM_EPI_COLS;
#pragma unroll
		for (int i = 0; i < NCH; ++i) {
			const int rr = rb + i * 32 + lane;
			nxMode[i] = rr < nReads ? __ldg(mode + rr) : 2;
			if (LIT) nxPv[i] = rr < abs(litStride) ? __ldg(litPv + rr) : 0u;
		}
	}
	__device__ __forceinline__ void begin(int mt, int, int row) {
		const int p = mt * HU_MM_BM + row;
		excl = nxExcl != 0u;
		// the common case, decided once per tile: every position of this warp is a candidate, every one of its reads exists and is
		// in tier 1 or 2
		bool odd = excl || p >= pitch;
#pragma unroll
		for (int i = 0; i < NCH; ++i) odd = odd || nxMode[i] == 2 || nxMode[i] < 0;
		tilePlain = !__any_sync(0xffffffffu, odd);
		if (LIT) {
			const int lane = threadIdx.x & 31;
			__syncwarp(); // the previous tile's flags() of this warp are done with the old pivots
#pragma unroll
			for (int i = 0; i < NCH; ++i) pvSm[i * 32 + lane] = nxPv[i];
			__syncwarp();
		}
	}
	__device__ __forceinline__ uint32_t key(int S) const {
		const int N = (S + 4096) >> 14;
		const int d = (16387 * N - S) >> 2; // (3 N - T) / 4 with T = S - 16384 N
		const uint4 e = __ldg(keyTab + N);
		const uint32_t q = __umulhi((uint32_t) d * e.y, e.z) >> e.w;
		return S == 0 ? 0xFFFFFFFFu : (uint32_t) d * e.x + q;
	}
	// k[j]: this position's key for read r0 + j.  The flag words of the tile's four chunks are transposed together after the last
	// one: eight independent butterfly chains hide the shuffle latency that one chain per chunk left exposed (two epilogue warps
	// per scheduler).  (Measured and dropped: the key table's first 1024 entries in shared memory.)
	__device__ __forceinline__ void flags(const uint32_t (&k)[32], int p, int r0, int lane) {
		const int c = (r0 & (HU_MM_EPI_COLS - 1)) >> 5;
		const uint4* pvp = reinterpret_cast<const uint4*>(pvSm + c * 32); // the same address in every lane: one broadcast per four pivots
		uint32_t xa = 0u, xb = 0u;
#pragma unroll
		for (int q = 0; q < 8; ++q) {
			const uint4 pv = pvp[q];
			const uint32_t pvs[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int j = 4 * q + i;
				if (k[j] >= pvs[i]) xa |= 1u << j;
				if (k[j] <= pvs[i]) xb |= 1u << j;
			}
		}
#pragma unroll
		for (int i = 0; i < NCH; ++i)
			if (i == c) { fa[i] = xa; fb[i] = xb; }
		if (c != NCH - 1) return;
#pragma unroll
		for (int i = 0; i < NCH; ++i) { fa[i] = warpBitTranspose(fa[i], lane); fb[i] = warpBitTranspose(fb[i], lane); }
		// lane j now holds, per chunk, the word of read j for this warp's 32 positions.  The four warps of the group hold the four
		// words of the tile (128 positions) of the same reads: they meet in shared memory and each thread stores 16 bytes of one
		// read's bitmap (4-byte stores, 32 sectors per instruction, cost the kernel 0.5 ms).  Two buffers, one barrier per tile:
		// a buffer is written again two tiles later, after the barrier of the tile in between, which every reader has passed.
		const int ew = (int) (threadIdx.x >> 5) - 2, q = (int) (threadIdx.x >> 5) & 3, g = ew >> 2;
		uint32_t* xb_ = xch + xchFlip * XCH_WORDS;
		xchFlip ^= 1;
#pragma unroll
		for (int i = 0; i < NCH; ++i) {
			xb_[(i * 32 + lane) * 4 + q] = fa[i];
			xb_[HU_MM_EPI_COLS * 4 + (i * 32 + lane) * 4 + q] = fb[i];
		}
		asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
		// thread t of the group's 128: bitmap t / 64... (HU_MM_EPI_COLS reads per bitmap), read t % HU_MM_EPI_COLS
		const int t = (ew & 3) * 32 + lane; // 0..127 within the group, any order
#pragma unroll
		for (int u = t; u < 2 * HU_MM_EPI_COLS; u += 128) {
			const int which = u / HU_MM_EPI_COLS, rd = u - which * HU_MM_EPI_COLS;
			const int rr = (r0 - 32 * (NCH - 1)) + rd;
			const uint4 w4 = *reinterpret_cast<const uint4*>(xb_ + which * HU_MM_EPI_COLS * 4 + rd * 4);
			if (rr < abs(litStride) && litStride > 0 && p < pitch) { // litStride < 0: a dry run for measurements (HU_LIT_DRY)
				uint32_t* dst = (which ? litBitB : litBitA) + (size_t) rr * (size_t) (pitch >> 5) + (size_t) ((p >> 7) << 2); // the tile's four words
				*reinterpret_cast<uint4*>(dst) = w4;
			}
		}
	}
	__device__ __forceinline__ void operator()(int mt, int nt, int row, int col0, const uint32_t (&v)[32], int lane) {
		const int p = mt * HU_MM_BM + row;
		const int r0 = nt * HU_MM_BN + col0;
		uint32_t* out = keys + (size_t) r0 * pitch + p;
		uint32_t k[32];
		if (tilePlain) { // the common case: no per-key tests
#pragma unroll
			for (int j = 0; j < 32; ++j) { k[j] = key((int) v[j]); *out = k[j]; out += pitch; }
		} else {
			const int rl = r0 + lane;
			const int md = rl < nReads ? __ldg(mode + rl) : 0;
			const unsigned rawMask = __ballot_sync(0xffffffffu, md == 2);  // finished by seedMmFixupKernel
			const unsigned badMask = __ballot_sync(0xffffffffu, md < 0);   // sparse list overflow: the read stays unplaced
			const int nr = min(32, nReads - r0);
#pragma unroll
			for (int j = 0; j < 32; ++j) {
				const int S = (int) v[j];
				if ((rawMask >> j) & 1u) {
					const int N = (S + 4096) >> 14;
					k[j] = ((uint32_t) ((16387 * N - S) >> 2) << 16) | (uint32_t) N;
				} else
					k[j] = (excl || ((badMask >> j) & 1u)) ? 0xFFFFFFFFu : key(S);
				if (j < nr && p < pitch) out[(size_t) j * pitch] = k[j];
			}
		}
		if (LIT) flags(k, p, r0, lane);
	}
};

// the key of one (position, read) pair from the bit planes (match columns) and the transposed sequences (the read's entries in
// non-match columns), as seedScanKernel computes it; bit-identical to the tile kernel's key for the reads of tiers 1 and 2
__device__ __forceinline__ uint32_t mmPairKey(const SeedArgs& a, const TreeDev& t, int p, int r) {
	const size_t pitch = t.scanPitch;
	const uint32_t* pl = t.planes + __ldg(t.mmSlotOfPos + p);
	const bool excl = !(__ldg(t.height + __ldg(t.nodeOfPos + p)) <= a.maxHeight);
	int d = 0, n = 0;
	const int w1 = a.rWordHi[r];
	for (int w = a.rWordLo[r]; w <= w1; ++w) {
		const uint32_t nh = __ldg(pl + ((size_t) w * 3 + 0) * pitch), nl = __ldg(pl + ((size_t) w * 3 + 1) * pitch), nv = __ldg(pl + ((size_t) w * 3 + 2) * pitch);
		const uint4 q = reinterpret_cast<const uint4*>(a.rplanes)[(size_t) w * a.nReadsPad + r];
		const uint32_t m = nv & q.z;
		d += __popc(((nh ^ q.x) | (nl ^ q.y)) & m);
		n += __popc(m);
	}
	const int ns = a.nSparse[r];
	const uint32_t* sp = a.sparse + (size_t) r * a.sparseCap;
	for (int k = 0; k < ns; ++k) {
		const uint32_t ent = sp[k];
		const int b = t.seqTP[(size_t) (ent >> 8) * t.posPitch + p];
		if (b >= 0) { n++; d += b != (int) (ent & 0xff); }
	}
	return mmMakeKey(d, n, excl, t.inv2N);
}

// Level-1 pivot of the literal seed order per read, ahead of the scan: std::sort's first __move_median_to_first looks at the
// keys at positions 1, n / 2 and n - 1 (bits/stl_algo.h; seedSortLiteralKernel does the same on the finished key vector and
// only uses the bitmaps if it arrives at the same pivot).  Four lanes per read, three of them compute one key each.
__global__ void __launch_bounds__(128) seedLitPivotKernel(SeedArgs a, TreeDev t, const int* __restrict__ mode, int n) {
	const int tid = blockIdx.x * 128 + threadIdx.x;
	const int r = tid >> 2, which = tid & 3;
	const bool ok = r < a.nReads && a.winStart[r] >= 0 && (mode[r] == 0 || mode[r] == 1);
	const int pos = which == 0 ? 1 : (which == 1 ? n / 2 : n - 1);
	uint32_t q = 0u;
	if (ok && which < 3) q = mmPairKey(a, t, pos, r);
	const int l0 = threadIdx.x & 28;
	const uint32_t qa = __shfl_sync(0xffffffffu, q, l0), qb = __shfl_sync(0xffffffffu, q, l0 + 1), qc = __shfl_sync(0xffffffffu, q, l0 + 2);
	if (which == 0 && r < a.litStride) {
		uint32_t pv;
		if (qa < qb) pv = qb < qc ? qb : (qa < qc ? qc : qa);
		else pv = qa < qc ? qa : (qb < qc ? qc : qb);
		a.litPv[r] = pv;
		a.litOk[r] = ok ? 1 : 0;
	}
}

// tier 3: (d << 16 | N) -> + the read's non-match columns -> key; one CTA per read (the others leave at once), four positions per thread
__global__ void __launch_bounds__(256) seedMmFixupKernel(SeedArgs a, TreeDev t, int pitch, const int* __restrict__ mode) {
	const int r = blockIdx.x;
	if (mode[r] != 2) return;
	const int ns = a.nSparse[r];
	const uint32_t* sp = a.sparse + (size_t) r * a.sparseCap;
	for (int p = threadIdx.x * 4; p < t.posPitch; p += 1024) { // positions past the last tile keep their 0xFFFFFFFF
		uint4* kp = reinterpret_cast<uint4*>(a.keys + (size_t) r * pitch + p);
		const uint4 raw = *kp;
		int dd[4] = {(int) (raw.x >> 16), (int) (raw.y >> 16), (int) (raw.z >> 16), (int) (raw.w >> 16)};
		int nn[4] = {(int) (raw.x & 0xffff), (int) (raw.y & 0xffff), (int) (raw.z & 0xffff), (int) (raw.w & 0xffff)};
		for (int i = 0; i < ns; ++i) {
			const uint32_t e = __ldg(sp + i);
			const int rb = (int) (e & 0xff);
			const char4 nb = *reinterpret_cast<const char4*>(t.seqTP + (size_t) (e >> 8) * t.posPitch + p);
			const int b4[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
			for (int s = 0; s < 4; ++s)
				if (b4[s] >= 0) { nn[s]++; dd[s] += b4[s] != rb; }
		}
		uint32_t key[4];
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			const int node = __ldg(t.nodeOfPos + p + s);
			const bool excl = node < 0 || node >= t.nNodes || !(__ldg(t.height + node) <= a.maxHeight);
			key[s] = mmMakeKey(dd[s], nn[s], excl, t.inv2N);
		}
		*kp = make_uint4(key[0], key[1], key[2], key[3]);
	}
}

// ---- tier 2: column -> reads of the batch with an entry there (counting sort on the column), then the pairs ----
// colStart: where a column's reads start in colReads; workStart: the (column, exception) work items of the columns with reads
__global__ void __launch_bounds__(1024) mmColScanKernel(int* __restrict__ colCount, int* __restrict__ colStart, int* __restrict__ workStart,
		const int* __restrict__ excPtr, int csLen) {
	__shared__ int part[1024], partW[1024];
	const int per = (csLen + 1023) / 1024, c0 = threadIdx.x * per;
	int s = 0, sw = 0;
	for (int c = c0; c < min(c0 + per, csLen); ++c) {
		s += colCount[c];
		if (colCount[c] > 0) sw += __ldg(excPtr + c + 1) - __ldg(excPtr + c);
	}
	part[threadIdx.x] = s; partW[threadIdx.x] = sw;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0, vw = threadIdx.x >= o ? partW[threadIdx.x - o] : 0;
		__syncthreads();
		part[threadIdx.x] += v; partW[threadIdx.x] += vw;
		__syncthreads();
	}
	int run = part[threadIdx.x] - s, runW = partW[threadIdx.x] - sw;
	for (int c = c0; c < min(c0 + per, csLen); ++c) {
		const int n = colCount[c];
		colStart[c] = run; workStart[c] = runW;
		colCount[c] = 0; // becomes the fill cursor
		run += n;
		if (n > 0) runW += __ldg(excPtr + c + 1) - __ldg(excPtr + c);
	}
	if (threadIdx.x == 1023) { colStart[csLen] = part[1023]; workStart[csLen] = partW[1023]; }
}
__global__ void __launch_bounds__(128) mmColFillKernel(SeedArgs a, TreeDev t, const int* __restrict__ mode, int* __restrict__ colCount,
		const int* __restrict__ colStart, int* __restrict__ colReads) {
	const int r = blockIdx.x * 128 + threadIdx.x;
	if (r >= a.nReads || mode[r] != 1) return;
	const int ns = a.nSparse[r];
	const uint32_t* sp = a.sparse + (size_t) r * a.sparseCap;
	for (int i = 0; i < ns; ++i) {
		const int col = (int) (sp[i] >> 8);
		if (__ldg(t.mmExcPtr + col + 1) > __ldg(t.mmExcPtr + col)) colReads[colStart[col] + atomicAdd(colCount + col, 1)] = r;
	}
}
// a warp per (column, exception position) work item, grid-stride; lane = one of the column's reads: the pair's d and N from the
// bit planes (match columns) and the transposed sequences (all of the read's non-match entries), as seedScanKernel computes them
__global__ void __launch_bounds__(128) seedMmExceptionKernel(SeedArgs a, TreeDev t, const int* __restrict__ colStart, const int* __restrict__ workStart,
		const int* __restrict__ colReads) {
	const int lane = threadIdx.x & 31;
	const int nWarps = gridDim.x * 4;
	const int total = workStart[t.csLen];
	const size_t pitch = t.scanPitch;
	for (int item = blockIdx.x * 4 + (threadIdx.x >> 5); item < total; item += nWarps) {
		int lo = 0, hi = t.csLen; // the column of this item: the last one whose work starts at or before it
		while (hi - lo > 1) {
			const int mid = (lo + hi) >> 1;
			if (workStart[mid] <= item) lo = mid; else hi = mid;
		}
		const int col = lo;
		const int p = __ldg(t.mmExcPos + __ldg(t.mmExcPtr + col) + item - workStart[col]);
		for (int i = colStart[col] + lane; i < colStart[col + 1]; i += 32) {
			const int r = colReads[i];
			const uint32_t k = mmPairKey(a, t, p, r);
			a.keys[(size_t) r * pitch + p] = k;
			if (a.litBitA && a.litOk[r]) {
				// the pair's two flags of the literal order's level 1 follow the key.  Only this pair's bit is at stake (a pair listed
				// under two columns gets the same key twice), so the bit is read plainly and only a changed one costs an atomic.
				const uint32_t pv = a.litPv[r], bit = 1u << (p & 31);
				const size_t o = (size_t) r * (size_t) (t.scanPitch >> 5) + (size_t) (p >> 5);
				const bool fa = k >= pv, fb = k <= pv;
				if (((a.litBitA[o] & bit) != 0u) != fa) { if (fa) atomicOr(a.litBitA + o, bit); else atomicAnd(a.litBitA + o, ~bit); }
				if (((a.litBitB[o] & bit) != 0u) != fb) { if (fb) atomicOr(a.litBitB + o, bit); else atomicAnd(a.litBitB + o, ~bit); }
			}
		}
	}
}

// seq[node][col] -> seqTP[col][position] (32 x 32 tiles through shared memory); gap code for padding positions
__global__ void mmTransposeKernel(const int8_t* __restrict__ seq, int csLen, const int* __restrict__ posNode, int posPitch, int8_t* __restrict__ seqTP) {
	__shared__ int8_t tile[32][33];
	const int p0 = blockIdx.x * 32, col0 = blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int p = p0 + r, col = col0 + threadIdx.x;
		const int node = p < posPitch ? posNode[p] : -1;
		tile[r][threadIdx.x] = (node >= 0 && col < csLen) ? seq[(size_t) node * csLen + col] : (int8_t) -2;
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int col = col0 + r, p = p0 + threadIdx.x;
		if (col < csLen && p < posPitch) seqTP[(size_t) col * posPitch + p] = tile[threadIdx.x][r];
	}
}

int configSeedMmKernels() {
	HU_CUDA(cudaFuncSetAttribute(seedMmKernel<KeyEpi<false> >, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_MM_SMEM));
	HU_CUDA(cudaFuncSetAttribute(seedMmKernel<KeyEpi<true> >, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_MM_SMEM));
	return HU_OK;
}

// host side of the database build: classes, per-column defaults and exception lists.  hs: node sequences [nNodes][csLen],
// posNode: position -> node (-1 padding).  colDefault: -1 match column, 0..3 the class-1 majority base, 4 long exception list
void seedMmColumns(const int8_t* hs, int csLen, const int* colK, const std::vector<int>& posNode, int nCand,
		std::vector<int8_t>& posClass, std::vector<int8_t>& colDefault, std::vector<int>& excPtr, std::vector<int>& excPos) {
	posClass.assign(posNode.size(), 0);
	std::vector<int> hist((size_t) csLen * 4, 0);
	for (int p = 0; p < nCand; ++p) {
		const int8_t* s = hs + (size_t) posNode[p] * csLen;
		bool full = true;
		for (int c = 0; c < csLen && full; ++c) full = s[c] >= 0;
		posClass[p] = full;
		if (full)
			for (int c = 0; c < csLen; ++c)
				if (!colK[c]) hist[(size_t) c * 4 + (s[c] & 3)]++;
	}
	colDefault.assign(csLen, -1);
	for (int c = 0; c < csLen; ++c)
		if (!colK[c]) {
			int g = 0;
			for (int b = 1; b < 4; ++b)
				if (hist[(size_t) c * 4 + b] > hist[(size_t) c * 4 + g]) g = b;
			colDefault[c] = (int8_t) g;
		}
	std::vector<int> count(csLen, 0);
	for (int pass = 0; pass < 2; ++pass) {
		if (pass == 1) {
			const int limit = std::max(64, nCand / 64);
			excPtr.assign(csLen + 1, 0);
			for (int c = 0; c < csLen; ++c) {
				if (colDefault[c] >= 0 && count[c] > limit) { colDefault[c] = 4; count[c] = 0; }
				excPtr[c + 1] = excPtr[c] + count[c];
			}
			excPos.assign(std::max(excPtr[csLen], 1), 0);
			std::fill(count.begin(), count.end(), 0);
		}
		for (int p = 0; p < nCand; ++p) {
			const int8_t* s = hs + (size_t) posNode[p] * csLen;
			const bool full = posClass[p];
			for (int c = 0; c < csLen; ++c) {
				const int g = colDefault[c];
				if (g < 0 || g > 3) continue;
				if (full ? s[c] != g : s[c] >= 0) {
					if (pass == 1) excPos[excPtr[c] + count[c]] = p;
					count[c]++;
				}
			}
		}
	}
}

// device side of the database build: operand A images, position-ordered transpose
int buildSeedMm(hu_db* db, const int8_t* seq, int K, int8_t* A, int8_t* seqTP, const int* posNode, const int8_t* posClass) {
	TreeDev& t = db->tree;
	mmPackNodesKernel<<<dim3(t.mmTiles, t.mmChunks + 1), HU_MM_BM>>>(seq, t.csLen, K, db->hmm.p2cs, posNode, posClass, t.mmChunks + 1, A);
	mmTransposeKernel<<<dim3(t.posPitch / 32, (t.csLen + 31) / 32), dim3(32, 8)>>>(seq, t.csLen, posNode, t.posPitch, seqTP);
	HU_CUDA(cudaGetLastError());
	HU_CUDA(cudaDeviceSynchronize());
	return HU_OK;
}

// the scan of one batch; a: the stage's arguments with rplanes / rValid / rWordLo / rWordHi / nSparse already made
int launchSeedMm(hu_batch* b, const SeedArgs& a, int* launches) {
	hu_db* db = b->db;
	const TreeDev& t = db->tree;
	const int numNt = (a.nReads + HU_MM_BN - 1) / HU_MM_BN, stride = t.mmChunks + 1;
	int* colCount = b->mmCols;
	int* colStart = b->mmCols + t.csLen + 1;
	int* workStart = b->mmCols + 2 * (t.csLen + 1);
	HU_CUDA(cudaMemsetAsync(colCount, 0, (size_t) (t.csLen + 1) * sizeof(int), b->stream));
	HU_CUDA(cudaMemsetAsync(b->mmTileExt, 0, (size_t) numNt * sizeof(int), b->stream));
	mmPackReadsKernel<<<dim3(numNt * (HU_MM_BN / 128), t.mmChunks), 128, 0, b->stream>>>(a, t.csLen, db->hmm.K, db->hmm.p2cs, stride, b->mmB);
	mmPackExtKernel<<<numNt * (HU_MM_BN / 128), 128, 0, b->stream>>>(a, t, stride, b->mmB, b->mmMode, b->mmTileExt, colCount);
	mmRangeKernel<<<numNt, HU_MM_BN, 0, b->stream>>>(a, b->mmRange);
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	MmArgs g = {t.mmA, b->mmB, b->mmRange, t.mmTiles, numNt, stride, b->mmTileExt, stride - 1};
	if (a.litBitA) // level-1 flags of the literal seed order ride in the epilogue: the pivots first
		seedLitPivotKernel<<<(a.nReads * 4 + 127) / 128, 128, 0, b->stream>>>(a, t, b->mmMode, a.nCand);
	const int gridMm = std::min(sms, t.mmTiles * numNt);
	uint8_t* posExcl = reinterpret_cast<uint8_t*>(b->mmCols + 3 * (t.csLen + 1)); // the tail of the per-column work arrays
	mmPosExclKernel<<<(t.scanPitch + 255) / 256, 256, 0, b->stream>>>(t.nodeOfPos, t.height, a.maxHeight, t.nNodes, t.scanPitch, posExcl);
#define HU_MM_LAUNCH(LIT) { \
		KeyEpi<LIT> epi; \
		epi.keys = a.keys; epi.pitch = t.scanPitch; epi.nReads = a.nReads; epi.nNodes = t.nNodes; epi.mode = b->mmMode; \
		epi.posExcl = posExcl; epi.keyTab = t.mmKeyTab; \
		epi.excl = false; epi.tilePlain = false; epi.pvSm = NULL; epi.xch = NULL; epi.xchFlip = 0; \
		epi.litPv = a.litPv; epi.litBitA = a.litBitA; epi.litBitB = a.litBitB; \
		epi.litStride = (LIT && getenv("HU_LIT_DRY")) ? -a.litStride : a.litStride; \
		seedMmKernel<KeyEpi<LIT> ><<<gridMm, HU_MM_THREADS, HU_MM_SMEM, b->stream>>>(g, epi); }
	if (a.litBitA) HU_MM_LAUNCH(true) else HU_MM_LAUNCH(false)
#undef HU_MM_LAUNCH
	seedMmFixupKernel<<<a.nReads, 256, 0, b->stream>>>(a, t, t.scanPitch, b->mmMode);
	mmColScanKernel<<<1, 1024, 0, b->stream>>>(colCount, colStart, workStart, t.mmExcPtr, t.csLen);
	mmColFillKernel<<<(a.nReads + 127) / 128, 128, 0, b->stream>>>(a, t, b->mmMode, colCount, colStart, b->mmColReads);
	seedMmExceptionKernel<<<sms * 8, 128, 0, b->stream>>>(a, t, colStart, workStart, b->mmColReads);
	*launches = a.litBitA ? 10 : 9;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
