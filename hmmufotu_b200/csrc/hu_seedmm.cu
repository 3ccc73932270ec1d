// K2, tensor-core form: the all-node p-distance scan of getSeed() (src/HmmUFOtu_main.cpp:127-139, src/SeqUtils.cpp:37-54)
// as an int8 contraction on the 5th-generation tensor cores (hu_mm.cuh), with the seed key made in the epilogue.
//
//   database creation : node sequences -> operand A images (rows = getSeed's candidate positions, i.e. the non-root nodes in
//                       id order; K = 4 bytes per profile match column), and a position-ordered transpose of the sequences
//   per batch         : aligned reads -> operand B images (the read's bases inside its window), chunk range per read tile
//   seedMmKernel      : S = A . B^T per (position, read) -> d, N -> key floor(d 2^31 / N), written in candidate order:
//                       a warp's 32 lanes hold 32 consecutive positions of one read = one 128-byte line per store
//   seedMmFixupKernel : reads with bases in NON-match columns (insertions buildGlobalAlign could place): the epilogue leaves
//                       (d << 16 | N) for them and this pass adds those columns from the transposed sequences, then makes the key
// Keys are bit-identical to seedScanKernel's (tests/test_gpu_mm.py); that kernel remains for profiles longer than
// HU_MM_MAX_N match columns and for the compacted candidate vector of the literal seed order under -H/--max-height.
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

// ---- operand A: one thread per (position, chunk) ----
__global__ void __launch_bounds__(HU_MM_BM) mmPackNodesKernel(const int8_t* __restrict__ seq, int csLen, int K, const int* __restrict__ p2cs,
		const int* __restrict__ posNode, int nChunks, int8_t* __restrict__ A) {
	const int mt = blockIdx.x, kc = blockIdx.y, row = threadIdx.x;
	const int node = __ldg(posNode + mt * HU_MM_BM + row);
	uint32_t w[32];
#pragma unroll
	for (int i = 0; i < 32; ++i) {
		const int k = kc * 32 + i + 1;
		w[i] = (node >= 0 && k <= K) ? mmBaseWord(seq[(size_t) node * csLen + __ldg(p2cs + k) - 1]) : 0u;
	}
	uint8_t* img = reinterpret_cast<uint8_t*>(A) + ((size_t) mt * nChunks + kc) * HU_MM_A_BYTES;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		*reinterpret_cast<uint4*>(img + mmTileOffset(row, c * 16)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// ---- operand B: one thread per (read, chunk); rows past the batch and reads without a window are zero ----
__global__ void __launch_bounds__(128) mmPackReadsKernel(SeedArgs a, int L, int K, const int* __restrict__ p2cs, int nChunks, int8_t* __restrict__ B) {
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
	uint8_t* img = reinterpret_cast<uint8_t*>(B) + ((size_t) nt * nChunks + kc) * HU_MM_B_BYTES;
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
	if (r < a.nReads && a.winStart[r] >= 0 && a.rWordHi[r] >= 0) {
		atomicMin(&sLo, a.rWordLo[r]);
		atomicMax(&sHi, a.rWordHi[r]);
	}
	__syncthreads();
	if (threadIdx.x == 0) range[blockIdx.x] = make_int2(sLo, sHi);
}

// ---- epilogue of seedMmKernel: accumulator -> seed key ----
// Integer form of the key.  2^31 = Q N + R (Q = floor(2^31 / N), R < N), so floor(d 2^31 / N) = d Q + floor(d R / N) with
// d R < N^2 < 2^24; the small quotient is a multiply-high: floor(t / N) = (t m) >> k for m = ceil(2^k / N) whenever
// t (m N - 2^k) < 2^k, which k = max(32, 24 + bits(N)) guarantees for t < N^2 (m < 2^32; N = 1 has R = 0 and takes m = 0).
// One 16-byte table entry per N: {Q, R, m, k - 32}; N = 0 never reads it (S == 0 <=> N == 0 -> excluded).  The same value as
// makeKey()'s correctly rounded floating-point form in hu_seed.cu (both are the exact floor; tests/test_gpu_mm.py).
struct KeyEpi {
	uint32_t* keys;
	int pitch, nReads, nNodes;
	const int* nSparse;
	const int* nodeOfPos;
	const double* height;
	const uint4* keyTab; // [HU_MM_MAX_N + 1]
	double maxHeight;
	bool excl; // per thread: this position is not a candidate (padding, or above --max-height)
	__device__ __forceinline__ void begin(int mt, int, int row) {
		const int p = mt * HU_MM_BM + row;
		const int node = p < pitch ? __ldg(nodeOfPos + p) : -1;
		excl = node < 0 || node >= nNodes || !(__ldg(height + node) <= maxHeight);
	}
	__device__ __forceinline__ uint32_t key(int S) const {
		const int N = (S + 4096) >> 14;
		const int d = (16387 * N - S) >> 2; // (3 N - T) / 4 with T = S - 16384 N
		const uint4 e = __ldg(keyTab + N);
		const uint32_t q = __umulhi((uint32_t) d * e.y, e.z) >> e.w;
		return S == 0 ? 0xFFFFFFFFu : (uint32_t) d * e.x + q;
	}
	__device__ __forceinline__ void operator()(int mt, int nt, int row, int col0, const uint32_t (&v)[32], int lane) const {
		const int p = mt * HU_MM_BM + row;
		const int r0 = nt * HU_MM_BN + col0;
		const int rl = r0 + lane;
		const int ns = rl < nReads ? __ldg(nSparse + rl) : 0;
		const unsigned rawMask = __ballot_sync(0xffffffffu, ns > 0);   // finished by seedMmFixupKernel
		const unsigned badMask = __ballot_sync(0xffffffffu, ns < 0);   // sparse list overflow: the read stays unplaced
		const bool plain = !__any_sync(0xffffffffu, excl || p >= pitch) && rawMask == 0u && badMask == 0u && r0 + 32 <= nReads;
		uint32_t* out = keys + (size_t) r0 * pitch + p;
		if (plain) { // the common case: no per-key tests
#pragma unroll
			for (int j = 0; j < 32; ++j) { *out = key((int) v[j]); out += pitch; }
			return;
		}
		if (p >= pitch) return;
		const int nr = min(32, nReads - r0);
#pragma unroll
		for (int j = 0; j < 32; ++j) {
			const int S = (int) v[j];
			uint32_t k;
			if ((rawMask >> j) & 1u) {
				const int N = (S + 4096) >> 14;
				k = ((uint32_t) ((16387 * N - S) >> 2) << 16) | (uint32_t) N;
			} else
				k = (excl || ((badMask >> j) & 1u)) ? 0xFFFFFFFFu : key(S);
			if (j < nr) out[(size_t) j * pitch] = k;
		}
	}
};

// reads with bases in non-match columns: (d << 16 | N) -> + those columns -> key; four positions per thread
__global__ void __launch_bounds__(256) seedMmFixupKernel(SeedArgs a, TreeDev t, int pitch) {
	const int r = blockIdx.y;
	const int ns = a.nSparse[r];
	if (ns <= 0) return;
	const int p = (blockIdx.x * 256 + threadIdx.x) * 4;
	if (p >= t.posPitch) return; // positions past the last tile keep their 0xFFFFFFFF
	uint4* kp = reinterpret_cast<uint4*>(a.keys + (size_t) r * pitch + p);
	const uint4 raw = *kp;
	int dd[4] = {(int) (raw.x >> 16), (int) (raw.y >> 16), (int) (raw.z >> 16), (int) (raw.w >> 16)};
	int nn[4] = {(int) (raw.x & 0xffff), (int) (raw.y & 0xffff), (int) (raw.z & 0xffff), (int) (raw.w & 0xffff)};
	const uint32_t* sp = a.sparse + (size_t) r * a.sparseCap;
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
	HU_CUDA(cudaFuncSetAttribute(seedMmKernel<KeyEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, HU_MM_SMEM));
	return HU_OK;
}

bool seedMmUsable(const hu_db* db) { return db->tree.mmA != NULL; }

// database side: operand A images, position-ordered transpose.  posNode: position -> node id, -1 padding, posPitch entries
int buildSeedMm(hu_db* db, const int8_t* seq, int K, int8_t* A, int8_t* seqTP, const int* posNode) {
	TreeDev& t = db->tree;
	mmPackNodesKernel<<<dim3(t.mmTiles, t.mmChunks), HU_MM_BM>>>(seq, t.csLen, K, db->hmm.p2cs, posNode, t.mmChunks, A);
	mmTransposeKernel<<<dim3(t.posPitch / 32, (t.csLen + 31) / 32), dim3(32, 8)>>>(seq, t.csLen, posNode, t.posPitch, seqTP);
	HU_CUDA(cudaGetLastError());
	HU_CUDA(cudaDeviceSynchronize());
	return HU_OK;
}

// the scan of one batch; a: the stage's arguments with rValid / rWordLo / rWordHi / nSparse already made
int launchSeedMm(hu_batch* b, const SeedArgs& a, int* launches) {
	hu_db* db = b->db;
	const TreeDev& t = db->tree;
	const int numNt = (a.nReads + HU_MM_BN - 1) / HU_MM_BN;
	mmPackReadsKernel<<<dim3(numNt * (HU_MM_BN / 128), t.mmChunks), 128, 0, b->stream>>>(a, t.csLen, db->hmm.K, db->hmm.p2cs, t.mmChunks, b->mmB);
	mmRangeKernel<<<numNt, HU_MM_BN, 0, b->stream>>>(a, b->mmRange);
	int sms = 0;
	HU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device));
	MmArgs g = {t.mmA, b->mmB, b->mmRange, t.mmTiles, numNt, t.mmChunks};
	KeyEpi epi;
	epi.keys = a.keys; epi.pitch = t.scanPitch; epi.nReads = a.nReads; epi.nNodes = t.nNodes; epi.nSparse = a.nSparse;
	epi.nodeOfPos = t.nodeOfPos; epi.height = t.height; epi.keyTab = t.mmKeyTab; epi.maxHeight = a.maxHeight; epi.excl = false;
	seedMmKernel<KeyEpi><<<std::min(sms, t.mmTiles * numNt), HU_MM_THREADS, HU_MM_SMEM, b->stream>>>(g, epi);
	seedMmFixupKernel<<<dim3((t.posPitch / 4 + 255) / 256, a.nReads), 256, 0, b->stream>>>(a, t, t.scanPitch);
	*launches = 4;
	HU_CUDA(cudaGetLastError());
	return HU_OK;
}

} // namespace hu
