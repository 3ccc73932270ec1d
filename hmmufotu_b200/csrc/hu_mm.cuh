// The tensor-core form of the seed scan (K2).  getSeed()'s p-distance d/N between a read and a node
// (src/SeqUtils.cpp:37-54) is two match counts over the window's columns, i.e. a contraction over profile positions:
// a base is the int8 vector (sh, sl, sh*sl, -128) with sh, sl = +-1 the signs of its two code bits (the first three
// components are a vertex of a regular tetrahedron: equal bases give +3, different bases -1), a gap is the zero vector.
// The dot product of a node row with a read row over the window is then
//     S = sum_k a_k . b_k = (3 M - D) + 16384 N,     M matches, D mismatches, N = M + D columns valid in both,
// one s32 accumulator per (node, read) from which N = (S + 4096) >> 14 and d = D = (3 N - (S - 16384 N)) / 4 follow exactly
// for N <= 4095.  nodes x (4 x positions) . reads is an int8 GEMM: tcgen05.mma kind::i8 with the accumulator tile
// (128 nodes x 256 reads, s32) in tensor memory, both operands K-major in shared memory under the 128-byte swizzle.
//
// Operand images.  The operands are stored in global memory as the shared-memory images of their tiles: tile (t, c) =
// rows 128 t .. (A) or 256 t .. (B), K bytes 128 c .. 128 c + 127 (32 profile positions), 8-row groups of 1024 bytes, row
// r of a group at 128 (r & 7), its 16-byte chunk j at chunk position j ^ (r & 7) (Swizzle<3,4,3>).  A stage of the
// pipeline is then ONE cp.async.bulk per operand (16 KB + 32 KB), no tensor map.
//
// Kernel: persistent, one CTA per SM, 2 + HU_MM_EPI_WARPS warps: warp 0 lane 0 issues the bulk copies (4-stage ring of
// full/empty mbarriers), warp 1 lane 0 issues the MMAs (4 x K=32 per stage) and commits them to the stage's empty barrier
// and, after the last chunk, to the accumulator's full barrier; the other warps are the epilogue: each reads its quarter of
// the TMEM lanes (nodes) and HU_MM_EPI_COLS of the columns (reads) with tcgen05.ld, turns S into the seed key and writes it
// -- lane = node, so a warp's store of one read's keys is one full 128-byte line of the key matrix.  Two accumulators
// (2 x 256 columns = all of TMEM) let the MMAs of the next tile run under the epilogue of this one.  16 epilogue warps of 96
// registers rather than 8 of 168: the epilogue's dependent chains (key table lookup, multiply-high) need the warps to hide
// them, and it is the longer side of the kernel once it also makes the literal seed order's flags (hu_seedmm.cu).
#pragma once
#include "hu_pipe.cuh"

#define HU_MM_BM 128
#define HU_MM_BN 256
#define HU_MM_BK 128 // K bytes per chunk: 32 profile positions x 4 components
#define HU_MM_STAGES 4
#define HU_MM_A_BYTES (HU_MM_BM * HU_MM_BK)
#define HU_MM_B_BYTES (HU_MM_BN * HU_MM_BK)
#define HU_MM_STAGE_BYTES (HU_MM_A_BYTES + HU_MM_B_BYTES)
#define HU_MM_EPI_WARPS 16 // epilogue warps: a multiple of four (TMEM lane quarter = warp id % 4), each takes HU_MM_EPI_COLS columns of the accumulator
#define HU_MM_EPI_COLS (HU_MM_BN / (HU_MM_EPI_WARPS / 4))
#define HU_MM_THREADS (64 + 32 * HU_MM_EPI_WARPS)
#define HU_MM_SMEM (HU_MM_STAGES * HU_MM_STAGE_BYTES + 1024) // + slack to align the ring to the swizzle period
#define HU_MM_MAX_N 4095 // valid columns per (node, read) for which the packed accumulator decodes exactly

namespace hu {

// byte offset of (row, K byte) inside a tile image
__host__ __device__ __forceinline__ uint32_t mmTileOffset(int row, int kb) {
	return (uint32_t) ((row >> 3) * 1024 + (row & 7) * 128 + ((((kb >> 4) ^ (row & 7)) & 7) << 4) + (kb & 15));
}
// the int8 quadruple of a base code (0..3), little endian: (sh, sl, sh * sl, -128); 0 for a gap
__host__ __device__ __forceinline__ uint32_t mmBaseWord(int code) {
	if (code < 0) return 0u;
	const int sh = (code & 2) ? -1 : 1, sl = (code & 1) ? -1 : 1;
	return (uint32_t) (sh & 0xff) | ((uint32_t) (sl & 0xff) << 8) | ((uint32_t) ((sh * sl) & 0xff) << 16) | 0x80000000u;
}

struct MmArgs {
	const int8_t* A;   // [numMt][nChunks][HU_MM_A_BYTES]
	const int8_t* B;   // [numNt][nChunks][HU_MM_B_BYTES]
	const int2* range; // [numNt]: first and last chunk any read of the tile has a base in (first > last: none)
	int numMt, numNt, nChunks;
	const int* tileExt; // [numNt] or NULL: the tile also takes the extension chunk `extChunk` (after its range)
	int extChunk;
};

#ifdef __CUDACC__
// ---- tcgen05 plumbing (PTX ISA, sm_100a) ----
__device__ __forceinline__ void tmemAlloc(uint32_t* dst, uint32_t cols) { // whole warp
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(dst)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmemFree(uint32_t addr, uint32_t cols) { // whole warp
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// the mbarrier receives one arrival when every tcgen05.mma this thread has issued so far is complete
__device__ __forceinline__ void tcCommit(unsigned long long* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> s32, one CTA, issued by one thread
__device__ __forceinline__ void tcMmaI8(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t iDesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
	             ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(iDesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 columns of 32-bit cells: thread = lane of the warp's TMEM quarter, v[j] = column j
__device__ __forceinline__ void tmemLd32(uint32_t addr, uint32_t (&v)[32]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
	               "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
	               "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
	               "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
	             : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmemLdWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major operand tile under the 128-byte swizzle: start address / 16 in bits 0-13,
// leading byte offset (unused by this layout; 1) in 16-29, stride byte offset = 1024 bytes between 8-row groups in 32-45,
// descriptor version 1 (sm_100) in 46-47, layout type 2 (SWIZZLE_128B) in 61-63
__device__ __forceinline__ uint64_t mmSmemDesc(uint32_t saddr) {
	return (uint64_t) ((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D s32 (bits 4-5 = 2), A and B signed int8 (bits 7-9, 10-12 = 1), both K-major (15, 16 = 0),
// N / 8 in bits 17-22, M / 16 in bits 24-28
#define HU_MM_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (HU_MM_BN >> 3) << 17) | ((uint32_t) (HU_MM_BM >> 4) << 24))

// Epi: __device__ void operator()(int mt, int nt, int row, int col0, const uint32_t (&v)[32], int lane) -- v[j] is the
// accumulator of (node row 128 mt + row, read 256 nt + col0 + j); begin(mt, nt, row) is called once per tile and thread, and
// prefetch(mt, nt, row) one tile ahead of it (it may only load into registers of the functor);
// init(smem, tid, nThreads) once per CTA by every thread, with Epi::SMEM_BYTES of shared memory of its own.
// (Measured and dropped, profiles/r2_summary.md: the tcgen05.ld of the next chunk issued ahead of this chunk's epilogue.)
template <class Epi>
__global__ void __launch_bounds__(HU_MM_THREADS, 1) seedMmKernel(MmArgs g, Epi epi) {
	extern __shared__ uint8_t mmSmem[];
	__shared__ unsigned long long full[HU_MM_STAGES], empty[HU_MM_STAGES], accFull[2], accEmpty[2];
	__shared__ uint32_t sTmem;
	__shared__ __align__(16) uint8_t epiSm[Epi::SMEM_BYTES > 0 ? Epi::SMEM_BYTES : 16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	epi.init(epiSm, threadIdx.x, HU_MM_THREADS); // published by the barrier below
	uint8_t* ring = mmSmem + ((1024u - (smemAddr(mmSmem) & 1023u)) & 1023u);
	if (threadIdx.x == 0) {
		for (int s = 0; s < HU_MM_STAGES; ++s) { mbarInit(&full[s], 1); mbarInit(&empty[s], 1); }
		for (int s = 0; s < 2; ++s) { mbarInit(&accFull[s], 1); mbarInit(&accEmpty[s], HU_MM_EPI_WARPS); }
		mbarInitFence();
	}
	if (warp == 1) tmemAlloc(&sTmem, 512);
	tcFenceBefore();
	__syncthreads();
	tcFenceAfter();
	const uint32_t tmem = sTmem;
	const int nItems = g.numMt * g.numNt;
	if (warp == 0) {
		if (lane == 0) {
			int stage = 0;
			uint32_t ph = 0;
			for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
				const int mt = item / g.numNt, nt = item - mt * g.numNt;
				const int2 rg = g.range[nt];
				const int last = (rg.x <= rg.y && g.tileExt && g.tileExt[nt]) ? rg.y + 1 : rg.y;
				for (int kk = rg.x; kk <= last; ++kk) {
					const int kc = kk > rg.y ? g.extChunk : kk;
					mbarWaitRelaxed(&empty[stage], ph ^ 1u);
					uint8_t* st = ring + (size_t) stage * HU_MM_STAGE_BYTES;
					mbarArriveExpectTx(&full[stage], HU_MM_STAGE_BYTES);
					bulkLoad(st, g.A + ((size_t) mt * g.nChunks + kc) * HU_MM_A_BYTES, HU_MM_A_BYTES, &full[stage]);
					bulkLoad(st + HU_MM_A_BYTES, g.B + ((size_t) nt * g.nChunks + kc) * HU_MM_B_BYTES, HU_MM_B_BYTES, &full[stage]);
					if (++stage == HU_MM_STAGES) { stage = 0; ph ^= 1u; }
				}
			}
		}
	} else if (warp == 1) {
		if (lane == 0) {
			int stage = 0, acc = 0;
			uint32_t ph = 0, accPh = 0;
			for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
				const int mt = item / g.numNt, nt = item - mt * g.numNt;
				const int2 rg = g.range[nt];
				if (rg.x > rg.y) continue;
				mbarWait(&accEmpty[acc], accPh ^ 1u);
				tcFenceAfter();
				const uint32_t d = tmem + (uint32_t) acc * HU_MM_BN;
				const int last = (g.tileExt && g.tileExt[nt]) ? rg.y + 1 : rg.y;
				for (int kc = rg.x; kc <= last; ++kc) {
					mbarWait(&full[stage], ph);
					tcFenceAfter();
					const uint32_t sa = smemAddr(ring + (size_t) stage * HU_MM_STAGE_BYTES);
					const uint64_t da = mmSmemDesc(sa), db = mmSmemDesc(sa + HU_MM_A_BYTES);
#pragma unroll
					for (int k = 0; k < HU_MM_BK / 32; ++k) // K = 32 int8 per instruction: 32 bytes further along the swizzled rows
						tcMmaI8(d, da + (uint64_t) (k * 2), db + (uint64_t) (k * 2), HU_MM_IDESC, (kc > rg.x || k > 0) ? 1u : 0u);
					tcCommit(&empty[stage]);
					if (++stage == HU_MM_STAGES) { stage = 0; ph ^= 1u; }
				}
				tcCommit(&accFull[acc]);
				acc ^= 1;
				if (acc == 0) accPh ^= 1u;
			}
		}
	} else {
		const int q = warp & 3;         // the TMEM lane quarter this warp may read (warp id % 4)
		const int h = (warp - 2) >> 2;  // which group of HU_MM_EPI_COLS columns of the accumulator
		int acc = 0;
		uint32_t accPh = 0;
		// items whose read tile has no window are skipped (by the other two roles as well)
		int item = blockIdx.x;
		while (item < nItems && g.range[item % g.numNt].x > g.range[item % g.numNt].y) item += gridDim.x;
		if (item < nItems) epi.prefetch(item / g.numNt, item % g.numNt, q * 32 + lane);
		while (item < nItems) {
			const int mt = item / g.numNt, nt = item - mt * g.numNt;
			epi.begin(mt, nt, q * 32 + lane); // takes what prefetch() loaded
			item += gridDim.x;
			while (item < nItems && g.range[item % g.numNt].x > g.range[item % g.numNt].y) item += gridDim.x;
			// the per-tile loads of the NEXT item are issued now and land while this tile is turned into keys (begin()'s loads
			// used to sit in front of every tile: 14 % of the kernel's stall samples)
			if (item < nItems) epi.prefetch(item / g.numNt, item % g.numNt, q * 32 + lane);
			mbarWait(&accFull[acc], accPh);
			tcFenceAfter();
			const uint32_t ta = tmem + ((uint32_t) (q * 32) << 16) + (uint32_t) (acc * HU_MM_BN + h * HU_MM_EPI_COLS);
#pragma unroll 1
			for (int c = 0; c < HU_MM_EPI_COLS / 32; ++c) {
				uint32_t v[32];
				tmemLd32(ta + (uint32_t) (c * 32), v);
				tmemLdWait();
				if (c == HU_MM_EPI_COLS / 32 - 1) { // the accumulator is in registers: hand it back to the MMA warp
					tcFenceBefore();
					__syncwarp();
					if (lane == 0) mbarArrive(&accEmpty[acc]);
				}
				epi(mt, nt, q * 32 + lane, h * HU_MM_EPI_COLS + c * 32, v, lane);
			}
			acc ^= 1;
			if (acc == 0) accPh ^= 1u;
		}
	}
	tcFenceBefore();
	__syncthreads();
	if (warp == 1) tmemFree(tmem, 512);
}
#endif

} // namespace hu
