// Batch workspace shared by the stage kernels (K1 Viterbi, K2 seed scan/select, K3 estimate, K4 place).
#pragma once
#include "hu_internal.cuh"

#define HU_CODE_GAP (-2) // '-' in an aligned string (DegenAlphabet gap code, src/DegenAlphabet.cpp:43-64)
#define HU_CODE_PAD (-3) // '.' padding outside the aligned region; any negative code is a gap to the SEP stages
#define HU_VIT_WARPS 7   // max warps (= sequences) per Viterbi CTA: 3 CTAs of 7 warps fit an SM for reads up to ~270 bases
#define HU_SCAN_TILE 512 // node slots per seed-scan CTA
#define HU_SCAN_READS 8  // reads per seed-scan CTA
#define HU_SEED_MAX_N 46340 // compared columns per (read, node) for which the integer seed keys are exact (hu_seed.cu)

namespace hu {

struct VRegion {
	int j0, j1, i0, i1;          // box: profile columns j0..j1, read rows i0..i1 (1-based, inclusive)
	int strip, pStart, pFrom, nIns, nDel; // seed strip: only cells with -nDel <= (i-pFrom)-(j-pStart) <= nIns
	int useB;                    // begin-state entry allowed (false in the downstream box)
	int rowWords;                // wavefront steps per 32-column stripe slot = rows + 31 (four back-pointer ballot words per step)
	int pad;
	long long ptrOff;            // word offset of this box in the back-pointer buffer
};
struct VPlan {
	int N, nReg, flags, pad;
	long long codeOff, traceOff;
	VRegion reg[5];
};
struct VitArgs {
	int nSeq, maxLen, warpsPerBlock, pad;
	const VPlan* plans;
	const int8_t* codes;
	uint32_t* ptrs;
	char* traces;
	int8_t* aligned;     // [nSeq][L]
	hu_align_row* rows;  // [nSeq]
	int* fullCount;      // [0] number of queued sequences, [1] cursor
	int* fullCursor;
	int* bandCursor;     // [2] next sequence of the banded pass
	int* fullList;
	uint32_t* fullPtrs;
	size_t fullSlabWords;
	double cNN, cNB, cEC, cCC;
};

struct SeedArgs {
	int nReads, maxNSeed;
	double maxDiff, maxHeight;
	const int8_t* aligned;  // [nReads][L]
	const int* winStart;    // 0-based inclusive window per read; winStart < 0: read is not placed
	const int* winEnd;
	uint32_t* rplanes;      // [nPWords][nReadsPad][4]: hi, lo, valid, 0 (profile coordinates)
	int nReadsPad;
	int* rValid;            // popcount of the valid plane per read
	int* rWordLo;           // first / last profile word with a read base
	int* rWordHi;
	int* nSparse;           // read bases in non-match columns: count and (column << 8 | base) list
	uint32_t* sparse;
	int sparseCap;
	uint32_t* keys;         // [nReads][scanPitch]: floor(d * 2^31 / N), 0xFFFFFFFF = excluded
	int* seedIds;           // [nReads][HU_MAX_SEED]
	int* nSeeds;            // [nReads]
	const int* posOfSlot;   // slot -> position in this run's candidate vector (TreeDev's, or the height-filtered one), -1: not a candidate
	const int* nodeOfPos;   // position -> node id
	int nCand;              // number of candidates of this run
	int* heapFallbacks;     // diagnostic counter: reads whose literal replay reached introsort's heap-sort fallback
	// level 1 of the literal seed order, flags made by the scan's epilogue while the keys are in registers (hu_seedmm.cu):
	uint32_t* litPv;        // [litStride] the level-1 pivot key per read (median of the keys at positions 1, n/2, n-1), computed ahead of the scan
	int* litOk;             // [litStride] 1: litPv and the read's two bitmaps are valid
	uint32_t* litBitA;      // [litStride][scanPitch / 32]: bit p%32 of word p/32 of a read's row = (key at position p >= pivot); NULL: not in use
	uint32_t* litBitB;      // same layout: key <= pivot
	int litStride;          // reads per bitmap word row (maxReads rounded up to 32)
	int* litFromScan;       // diagnostic counter: reads whose first flag pass came from the scan
	int* litCursor;         // next read of the literal order's persistent CTAs (NULL: reads dealt round-robin)
};

// one (read, seed) of the estimate stage, resolved once so that the streaming kernel has no dependent loads
struct EstItem {
	int rs;       // read * HU_MAX_SEED + seed slot
	int u, v;     // seed node and its parent
	int start, W; // window
	int kind;     // 0: nothing to do (also the end-of-stream mark), 1: an item
	double w0;    // branch length of (u, v)
};

struct EstItemP {
	double cDist, ratio; // pDist(child seq, read); attachment ratio cDist / (cDist + pDist)
	double sumScale;     // sum over the window of the per-column scales of both messages
	int taxU, taxV;      // taxon ids of the child and the parent node (the row's taxon is one of them)
	double P[2][16];     // P(w0 ratio) and P(w0 (1 - ratio)), row-major
};

} // namespace hu

struct hu_batch {
	hu_db* db;
	cudaStream_t stream;
	cudaEvent_t ev[16];
	int maxReads, maxSeqLen, paired, maxSeq;
	int nReads, nSeq;
	long long launches;
	// ---- K1 ----
	hu::VitArgs vit;
	int8_t* dCodes;
	long long codesCap;
	hu::VPlan* dPlans;
	uint32_t* dPtrs;
	long long ptrsCap;
	char* dTraces;
	long long tracesCap;
	char* dTracesPacked;       // traces start-to-end, NUL-terminated, back to back (built by hu_batch_get_align)
	long long* dTraceOffs;     // [nSeq + 1]
	long long* hTraceOffs;     // pinned
	long long d2hBytes;        // bytes the last getter moved to the host
	int fullSlots;
	int vitGrid;           // resident CTAs of the persistent banded kernel
	hu_align_row* merged;  // [maxReads]
	std::vector<hu::VPlan> hPlans;
	std::vector<long long> hTraceOff;
	// ---- SEP ----
	int* winStart;
	int* winEnd;
	hu::SeedArgs seed;
	hu_place_row* est;     // [maxReads][HU_MAX_SEED] estimateSeq results, seed order
	int* order;            // [maxReads][HU_MAX_SEED] filterPlacements order (indices into est)
	int* nKept;            // [maxReads]
	hu_place_row* places;  // [maxReads][HU_MAX_SEED] placeSeq results, filter order
	hu_place_row* best;    // [maxReads]
	int* workItems;        // items of K3 whose window is wider than the staged kernel takes
	hu::EstItem* estItems; // [maxReads][HU_MAX_SEED] item records of K3
	hu::EstItemP* estItemsP;
	int* workCount;
	unsigned long long* litIds; // (key, candidate position) work rows of the literal seed order, one per resident CTA
	uint32_t* litBits;     // level-1 bitmaps of the literal seed order written by the tensor-core scan (allocated on first use), then litPv, litOk
	int* posOfSlotH;       // candidate map of the literal seed order under -H/--max-height (built on first use per height)
	int* nodeOfPosH;
	int nCandH;
	double heightCached;
	int8_t* mmB;           // K2 on the tensor cores: operand images of the reads, [ceil(maxReads / 256)][mmChunks][32 KB]
	int2* mmRange;         // chunk range per read tile
	int* mmMode;           // per read: 0 no non-match base, 1 extension chunk, 2 raw + fix-up, -1 unplaced
	int* mmTileExt;        // per read tile: takes the extension chunk
	int* mmCols;           // [3][csLen + 1]: per-column counts / cursors, starts of mmColReads, starts of the exception work items
	int* mmColReads;       // reads with an entry in a column that has exceptions, grouped by column
	int mmScanRan;         // the last seed stage ran the tensor-core scan
	double* placeSlab;     // global rho slabs for windows wider than the shared-memory capacity
	void* estSlab;         // K3: per-CTA g slabs for windows wider than its shared-memory buffer
	int estSlabCtas;
	long long h2dBytes;
	// stats
	hu_batch_stats stats;
	int maxW;              // widest window of the current batch (host-side bound for shared memory sizing)
	std::vector<void*> allocs;
	std::vector<void*> pinned;
	// pinned staging
	int8_t* hCodes;
	hu::VPlan* hPlansPinned;
};

namespace hu {
int launchViterbi(hu_batch* b, int mode);
int launchSeed(hu_batch* b, const hu_sep_params* p);
int launchSeedMm(hu_batch* b, const SeedArgs& a, int* launches);
int launchEstimate(hu_batch* b, const hu_sep_params* p);
int launchFilter(hu_batch* b, const hu_sep_params* p);
int launchPlace(hu_batch* b, const hu_sep_params* p);
int launchFinalize(hu_batch* b, const hu_sep_params* p);
} // namespace hu
