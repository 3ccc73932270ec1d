/*
 * hmmufotu_b200.h -- C ABI of the B200-native per-read assignment path of HmmUFOtu.
 *
 * The reference (Grice-Lab/HmmUFOtu v1.5.1, pure C++/Eigen/OpenMP) has no FFI; its de-facto boundary for
 * this path is the set of free functions declared in src/HmmUFOtu_main.h:70-113, called once per read
 * from an OpenMP task (src/hmmufotu.cpp:610-748).  Each entry point below is the batched replacement
 * of one of those calls; the citation says which.  Everything crossing the boundary is a plain pointer,
 * a size or a POD struct; no exception crosses it; every function returns HU_OK or a negative status and
 * leaves a message for hu_last_error().  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with HU_ERR_CUDA.
 *
 * Conventions: profile positions k, read positions i and consensus columns in hu_vpath / hu_align_row are
 * 1-based as in the reference; window bounds (start, end) handed to the SEP stages are 0-based inclusive
 * consensus columns (csStart-1, csEnd-1: src/hmmufotu.cpp:645).  Base codes: A,C,G,T = 0..3 (degenerate
 * IUPAC codes map to their first member on the host, src/DegenAlphabet.cpp:43-64), any negative value = gap.
 */
#ifndef HMMUFOTU_B200_H_
#define HMMUFOTU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HU_OK 0
#define HU_ERR_ARG -1
#define HU_ERR_CUDA -2
#define HU_ERR_NOMEM -3
#define HU_ERR_STATE -4

#define HU_MAX_SEED 64 /* upper bound for maxNSeed (reference default 50, src/hmmufotu.cpp:48) */
#define HU_MAX_DG 8    /* discrete-Gamma categories (reference allows 2..8, src/hmmufotu-build.cpp:55-62) */

/* BandedHMMP7::align_mode, src/BandedHMMP7.h:163-168 */
enum { HU_GLOBAL = 0, HU_LOCAL = 1, HU_NGCL = 2, HU_CGNL = 3 };
/* DNASubModel subclasses, src/DNASubModelFactory.cpp */
enum { HU_GTR = 0, HU_TN93 = 1, HU_HKY85 = 2, HU_F81 = 3, HU_K80 = 4, HU_JC69 = 5 };

typedef struct hu_db hu_db;       /* one database replica on one device */
typedef struct hu_batch hu_batch; /* device workspace + stream for one batch of reads */

/* Profile parameters as BandedHMMP7 holds them after operator>> + wingRetract
 * (src/BandedHMMP7.cpp:100-246, 1083-1120; members src/BandedHMMP7.h:499-564).  All costs are -ln(prob). */
typedef struct hu_hmm_params {
	int32_t K;            /* profile size */
	int32_t L;            /* consensus length (MAXL) */
	const double* eM;     /* E_M_cost(b,k) at [k*4+b], k = 0..K */
	const double* eI;     /* E_I_cost(b,k) at [k*4+b] */
	const double* T;      /* Tmat_cost[k](a,b) at [k*9+a*3+b], a,b in {M=0,I=1,D=2} */
	const double* entry;  /* entryPr_cost(k), k = 0..K, wing-retracted */
	const double* exit;   /* exitPr_cost(k) */
	const int32_t* p2cs;  /* profile2CSIdx[k], k = 0..K */
} hu_hmm_params;

/* adjustProfileLocalMode + wingRetract (src/BandedHMMP7.cpp:721-733, 1083-1120): fills entry/exit costs
 * (K+1 doubles each) from the transition costs.  Host-only. */
int hu_hmm_finalize(int32_t K, const double* T, double* entry, double* exit_);

/* DNASubModel parameters (src/GTR.h, TN93.h, HKY85.h, F81.h, K80.h, JC69.h) + DiscreteGammaModel rates
 * (src/DiscreteGammaModel.h:75-77; dgK = 0: no discrete Gamma). */
typedef struct hu_model_desc {
	int32_t type;
	int32_t dgK;
	double pi[4];
	double R[16]; /* GTR symmetric exchangeabilities, row-major */
	double kr, ky, kappa, beta;
	double dgRates[HU_MAX_DG];
} hu_model_desc;

/* DNASubModel::Pr(v) -> row-major 4x4 (host evaluation of the same code the kernels run). */
int hu_model_Pr(const hu_model_desc* m, double v, double* P16);

/* PhyloTreeUnrooted as loaded from a .ptu (src/PhyloTreeUnrooted.cpp:496-535), flattened. */
typedef struct hu_tree_desc {
	int64_t nNodes;
	int32_t csLen;
	int32_t root;
	const int32_t* parent; /* parent id, -1 for the root */
	const double* blen;    /* length of branch (id, parent[id]) */
	const double* height;  /* node2height */
	const int8_t* seq;     /* nNodes x csLen node sequences (leaf: observed, internal: inferred) */
	const int32_t* taxon;  /* annotation-string id per node: equal ids <=> equal getTaxonName() strings */
	/* messages "before convolution", [node][column][4]: up = loglik(id -> parent), dn = loglik(parent -> id)
	 * (src/PhyloTreeUnrooted.cpp:595-603).  Both NULL: evaluate them on the device from the leaf
	 * sequences, as hmmufotu-build does by re-rooting at every node (src/hmmufotu-build.cpp:455-459). */
	const double* up;
	const double* dn;
	int32_t inferInternalSeq; /* with up == NULL: replace internal rows of seq by inferSeq() (cpp:1085-1093) */
} hu_tree_desc;

int hu_db_create(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device);
int hu_db_destroy(hu_db* db);
/* The same replica built in steps, for databases whose messages do not fit in host memory (98 GB at gg_97 scale):
 * hu_db_create_deferred takes everything but the messages (tree->up / dn are ignored), hu_db_put_messages streams them in
 * -- n messages of 4 x csLen log-domain doubles each; nodes[i] / dirs[i] name the directed edge (0: node -> parent,
 * 1: parent -> node), in any order, as PTUnrooted::load meets them in the .ptu (src/PhyloTreeUnrooted.cpp:595-603) -- and
 * hu_db_finish completes the tables derived from them. */
int hu_db_create_deferred(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device);
int hu_db_put_messages(hu_db* db, int64_t n, const int64_t* nodes, const int32_t* dirs, const double* logMsg);
int hu_db_finish(hu_db* db);
/* a replica of `src` on another device, copied device to device (peer path: NVLink where available) */
int hu_db_clone(hu_db** out, hu_db* src, int device);
/* read back device-resident pieces (tests, DB writer): n nodes starting at node0; any pointer may be NULL */
int hu_db_get_messages(hu_db* db, int64_t node0, int64_t n, double* up, double* dn);
int hu_db_get_seq(hu_db* db, int64_t node0, int64_t n, int8_t* seq);
/* messages of the listed nodes over consensus columns [col0, col0 + nCols): [i][column][4], log domain */
int hu_db_get_messages_of(hu_db* db, int64_t nIds, const int64_t* ids, int32_t col0, int32_t nCols, double* up, double* dn);
int64_t hu_db_bytes(hu_db* db);

/* BandedHMMP7::ViterbiAlignPath, src/BandedHMMP7.h:229-249 */
typedef struct hu_vpath { int32_t start, end, from, to, nIns, nDel; } hu_vpath;

/* A batch of reads as the host-side seed stage hands them over (src/HmmUFOtu_main.cpp:39-84).
 * paired != 0: sequences 0..nReads-1 are the forward reads and nReads..2*nReads-1 the already
 * reverse-complemented mates (src/hmmufotu.cpp:609), so codes/offsets/vpaths/nPaths hold 2*nReads entries. */
typedef struct hu_read_batch {
	int32_t nReads;
	int32_t paired;
	const int8_t* codes;    /* concatenated base codes 0..3 */
	const int64_t* offsets; /* nSeq+1 offsets into codes */
	const hu_vpath* vpaths; /* 2 per sequence */
	const int32_t* nPaths;  /* 0..2 per sequence; 0 = full Viterbi (src/HmmUFOtu_main.cpp:96) */
} hu_read_batch;

#define HU_ALIGN_FALLBACK 1  /* banded pass found no path, full Viterbi was run (src/HmmUFOtu_main.cpp:89-93) */
#define HU_ALIGN_NOSEED 2    /* no VPath given: full Viterbi */
#define HU_ALIGN_IRREGULAR 4 /* VPaths not monotone: treated as HU_ALIGN_NOSEED (documented deviation) */
#define HU_ALIGN_CHIMERA 8   /* PE orientation test failed (src/hmmufotu.cpp:629-639): not placed */
#define HU_ALIGN_INVALID 16  /* !HmmAlignment::isValid() */

/* BandedHMMP7::HmmAlignment (src/BandedHMMP7.h:74-134) minus the aligned string, which is rebuilt from
 * the trace on the host (hu_trace_to_align) or kept as codes on the device. */
typedef struct hu_align_row {
	int32_t seqStart, seqEnd, hmmStart, hmmEnd, csStart, csEnd;
	int32_t flags;
	int32_t traceLen; /* length of the "B[MID]*E" trace */
	double cost;
	int64_t cells;    /* DP cell updates executed for this sequence */
} hu_align_row;

/* PTUnrooted::PTPlacement (src/PhyloTreeUnrooted.h:410-510) as a POD row. */
typedef struct hu_place_row {
	int32_t start, end;      /* 0-based window */
	int32_t cNode, pNode, aNode;
	int32_t taxon;           /* taxon id of aNode */
	double dist;             /* PTLoc.dist of the seed: pDist(cNode.seq, read) over the window */
	double wuv, ratio, wnr;
	double loglik;           /* literal return value of placeSeq (constant per read) or the root loglik */
	double loglikTrue;       /* sum_j log(pi . exp(loglik(r, j))): evaluated by hu_place_batch only when useTrueLoglik is set, else 0 */
	double estLoglik;        /* estimateSeq() loglik */
	double height, annoDist, qPlace, qTaxon;
	int32_t outerIters, nPlaces;
} hu_place_row;

typedef struct hu_sep_params {
	int32_t maxNSeed;    /* src/hmmufotu.cpp:48  (50) */
	int32_t useTrueLoglik; /* 0: literal (default), 1: intended root loglik */
	double maxDiff;      /* src/hmmufotu.cpp:50  (inf) */
	double maxHeight;    /* src/hmmufotu.cpp:52  (inf) */
	double maxError;     /* src/hmmufotu.cpp:49  (20) */
	const int32_t* equalSortPerm; /* ignored (kept for layout compatibility): the std::sort calls of filterPlacements and of
	                               * the final ordering (src/HmmUFOtu_main.cpp:164, src/hmmufotu.cpp:725-731) are replayed on
	                               * the device move for move, ties and NaN keys included */
	int32_t seedOrder;   /* HU_SEED_ORDER_STDSORT (default): the seeds libstdc++'s std::sort leaves in front, ties
	                      * included (src/HmmUFOtu_main.cpp:139), also under maxHeight (the candidate vector is then the
	                      * height-filtered one, as in getSeed); HU_SEED_ORDER_ID: deterministic (dist, id) order.
	                      * Limits of the seed stage (HU_ERR_ARG beyond them): consensus length < 65536 columns; the
	                      * literal order keeps four bitmaps of the candidate vector in shared memory: <= ~409 000 nodes */
	int32_t estMethod;   /* -m|--method (src/hmmufotu.cpp:229-232): HU_EST_UNWEIGHTED (default) or HU_EST_WEIGHTED
	                      * (estimateBranchLengthWeighted, src/PhyloTreeUnrooted.cpp:1034-1052) */
	int32_t prior;       /* --prior (src/hmmufotu.cpp:236-241): HU_PRIOR_UNIFORM (default) or HU_PRIOR_HEIGHT
	                      * (PTPlacement::logPriorPr, src/PhyloTreeUnrooted.cpp:1166-1177) */
	int32_t onlyML;      /* --ML (src/hmmufotu.cpp:233-235, 725-727): order by log-likelihood, no Q values */
} hu_sep_params;
#define HU_SEED_ORDER_STDSORT 0
#define HU_SEED_ORDER_ID 1
#define HU_EST_UNWEIGHTED 0
#define HU_EST_WEIGHTED 1
#define HU_PRIOR_UNIFORM 0
#define HU_PRIOR_HEIGHT 1
void hu_sep_params_default(hu_sep_params* p);

/* ---- batch workspace ---- */
int hu_batch_create(hu_batch** out, hu_db* db, int32_t maxReads, int32_t maxSeqLen, int32_t paired);
int hu_batch_destroy(hu_batch* b);

/* host -> device staging of one batch (asynchronous on the batch stream; host buffers must stay valid
 * until hu_batch_sync) */
int hu_batch_upload(hu_batch* b, const hu_read_batch* reads);
/* alternative entry for the SEP stages: aligned DigitalSeq rows (n x csLen codes) + windows */
int hu_batch_set_aligned(hu_batch* b, int32_t n, const int8_t* codes, const int32_t* start, const int32_t* end);

/* Stage inputs supplied by the caller instead of by the previous stage: what the reference's free functions take as
 * arguments when called one by one (src/HmmUFOtu_main.h:70-113); rows are nReads x maxNSeed as the getters return them.
 * set_seeds: the seeds estimateSeq() works on; set_estimates: the vector filterPlacements() / placeSeq() work on (kept,
 * in the given order, until hu_filter_batch runs); set_places: the vector calcQValues() / the final sort work on. */
int hu_batch_set_seeds(hu_batch* b, const int32_t* nSeeds, const int32_t* ids);
int hu_batch_set_estimates(hu_batch* b, const int32_t* n, const hu_place_row* rows);
int hu_batch_set_places(hu_batch* b, const int32_t* n, const hu_place_row* rows);

/* alignSeq() for every sequence of the batch: banded Viterbi, fallback, trace, aligned codes, PE merge
 * (src/HmmUFOtu_main.cpp:86-104, src/BandedHMMP7.cpp:748-1081, 1188-1213).  Device only. */
int hu_align_batch(hu_batch* b, int mode);
/* getSeed() + truncation (src/HmmUFOtu_main.cpp:127-152, src/hmmufotu.cpp:646-647) */
int hu_seed_batch(hu_batch* b, const hu_sep_params* p);
/* estimateSeq() over the seeds + filterPlacements() (src/HmmUFOtu_main.cpp:154-173) */
int hu_estimate_batch(hu_batch* b, const hu_sep_params* p);
/* filterPlacements() alone over the estimates on the device (src/HmmUFOtu_main.cpp:162-173) */
int hu_filter_batch(hu_batch* b, const hu_sep_params* p);
/* placeSeq() over the kept placements (src/HmmUFOtu_main.cpp:175-180, src/PhyloTreeUnrooted.cpp:879-954) */
int hu_place_batch(hu_batch* b, const hu_sep_params* p);
/* calcQValues() + sort + bestPlace (src/HmmUFOtu_main.cpp:182-216, src/hmmufotu.cpp:725-733) */
int hu_finalize_batch(hu_batch* b, const hu_sep_params* p);
/* all five stages back to back on the batch stream */
int hu_assign_batch(hu_batch* b, int mode, const hu_sep_params* p);
int hu_batch_sync(hu_batch* b);

/* device -> host results (each call synchronises the batch stream).  hu_batch_get_align packs the traces on the device
 * (start-to-end, NUL-terminated, back to back; traceOff[s] = offset of sequence s, traceOff[nSeq] = total bytes) and copies
 * exactly those bytes; page-locked `rows` / `traces` buffers (hu_host_alloc) make the copies asynchronous DMA. */
int hu_batch_get_align(hu_batch* b, hu_align_row* rows /* nSeq */, char* traces, int64_t traceCap, int64_t* traceOff /* nSeq+1 */);
int hu_batch_get_merged(hu_batch* b, hu_align_row* rows /* nReads */);
int hu_batch_get_aligned_codes(hu_batch* b, int8_t* codes /* nReads x csLen */);
int hu_batch_get_seeds(hu_batch* b, int32_t* nSeeds, int32_t* ids /* nReads x maxNSeed */, double* dist);
int hu_batch_get_estimates(hu_batch* b, int32_t* nKept, hu_place_row* rows /* nReads x maxNSeed, filter order */);
int hu_batch_get_places(hu_batch* b, int32_t* nKept, hu_place_row* rows /* nReads x maxNSeed */);
int hu_batch_get_best(hu_batch* b, hu_place_row* rows /* nReads */);

/* The read-sharding driver's fixed-size assignment row (104 bytes: int64 read; int32 seqStart, seqEnd, hmmStart, hmmEnd,
 * csStart, csEnd, flags, nPlaces; double cost; int32 cNode, pNode, aNode, taxon; double ratio, annoDist, loglik, qPlace,
 * qTaxon) for every read of the batch, written by a kernel on the batch stream into DEVICE memory `devRows` (nReads rows):
 * the gather of these rows over the ranks is the one collective of the path (ncclAllGather on device buffers). */
int hu_batch_pack_rows(hu_batch* b, int64_t read0, void* devRows);
/* the batch's cudaStream_t, for ordering a caller's collective after the batch's kernels */
void* hu_batch_stream(hu_batch* b);

/* diagnostics (tests): the key vector of one read after hu_seed_batch -- floor(d * 2^31 / N) per candidate, candidates =
 * the non-root nodes in id order that pass maxHeight (0xFFFFFFFF: undefined distance); nCand = their number */
int hu_batch_get_seed_keys(hu_batch* b, int32_t read, uint32_t* keys, int32_t* nCand);
/* the device's literal-order kernel on caller-supplied key rows (nRows x n), positions of the first `want` out */
int hu_debug_literal_order(int device, const uint32_t* keys, int32_t nRows, int32_t n, int32_t want, int32_t* outPos /* nRows x want */,
		int32_t* nOut /* nRows */, int32_t* heapFallbacks /* 1, may be NULL */);
/* the library's restatement of std::sort(rbegin, rend) on n keys (host): perm[k] = index of the element left at k */
int hu_debug_stdsort(const double* keys, int32_t n, int32_t* perm);

/* user timing marks: record CUDA event `id` (0..7) on the batch stream; elapsed ms between two marks
 * (possibly of two workspaces of the same device), both of which must have completed */
int hu_batch_mark(hu_batch* b, int id);
int hu_batch_elapsed(hu_batch* from, int idFrom, hu_batch* to, int idTo, float* ms);

/* per-stage device time of the last run of each stage (CUDA events on the batch stream), ms:
 * [0] align [1] seed scan [2] seed select [3] estimate [4] place [5] finalize; and work counters */
int hu_batch_stage_ms(hu_batch* b, float ms[8]);
typedef struct hu_batch_stats {
	int64_t cells;        /* Viterbi cell updates */
	int64_t seedPairs;    /* (read, node) pairs scanned */
	int64_t seedBytes;    /* algorithmic bytes of the scan: sum_reads (nNodes-1) * W + W */
	int64_t estItems;     /* (read, seed) items */
	int64_t estBytes;     /* 65 * W per item */
	int64_t placeItems;   /* (read, placement) items */
	int64_t placeBytes;   /* 64 * W per item */
	int64_t launches;     /* kernels launched since the last hu_batch_upload / set_aligned */
	int64_t seedHeapSorts; /* reads whose literal seed order reached introsort's heap-sort fallback (replayed; diagnostic) */
	int64_t seedMacs;     /* int8 multiply-accumulates of the tensor-core seed scan (tiles x chunks visited); 0: the bit-plane scan ran */
	int64_t seedScanFlags; /* reads whose literal seed order took its first flag pass from the scan's epilogue (diagnostic) */
} hu_batch_stats;
int hu_batch_get_stats(hu_batch* b, hu_batch_stats* s);

/* buildGlobalAlign() on the host (src/BandedHMMP7.cpp:1008-1081): trace + read -> aligned string of L chars */
int hu_trace_to_align(const hu_hmm_params* hmm, const char* read, int32_t readLen, const hu_align_row* row,
		const char* trace, char* align /* L bytes, not NUL-terminated */);

/* page-locked host memory for the staging buffers of hu_batch_upload / the getters */
void* hu_host_alloc(size_t bytes);
void hu_host_free(void* p);

const char* hu_last_error(void);
const char* hu_version(void);

#ifdef __cplusplus
}
#endif
#endif
