// Internal declarations shared by the kernels and the C ABI (hmmufotu_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/hmmufotu_b200.h"
#include "hu_mm.cuh"

#define HU_HD __host__ __device__ __forceinline__

namespace hu {

// ---------------------------------------------------------------- error plumbing ----
void setError(const char* fmt, ...);
#define HU_CUDA(call)                                                                      \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                           \
		if (e_ != cudaSuccess) {                                                           \
			hu::setError("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return HU_ERR_CUDA;                                                            \
		}                                                                                  \
	} while (0)

// ---------------------------------------------------------------- substitution model ----
// Derived form of hu_model_desc; Pr() runs identically on host and device.
struct Model {
	int type;
	int dgK;
	double pi[4];
	double logPi[4];
	double kr, ky, kappa, beta;
	double lambda[4];   // GTR: eigenvalues of the trace-scaled Q
	double U[16], Ui[16]; // GTR: eigenvectors / inverse, row-major
	double dgR[HU_MAX_DG];
};
int buildModel(const hu_model_desc* d, Model* m); // host: eigen-decomposition etc.

// DNASubModel::Pr(v): GTR src/GTR.h:116-121; TN93 src/TN93.h:113-154; HKY85 src/HKY85.h:111-153;
// F81 src/F81.h:110-118; K80 src/K80.h:109-118; JC69 src/JC69.h:97-101.  Row-major P[i*4+j].
HU_HD void modelPr(const Model& m, double v, double* P) {
	switch (m.type) {
	case HU_GTR: {
		if (v == 0) {
			for (int i = 0; i < 16; ++i) P[i] = (i % 5 == 0) ? 1.0 : 0.0;
			return;
		}
		double e[4];
		for (int k = 0; k < 4; ++k) e[k] = exp(m.lambda[k] * v);
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) {
				double s = 0;
				for (int k = 0; k < 4; ++k) s += m.U[i * 4 + k] * e[k] * m.Ui[k * 4 + j];
				P[i * 4 + j] = s;
			}
		return;
	}
	case HU_TN93:
	case HU_HKY85: {
		const double kR = m.type == HU_TN93 ? m.kr : m.kappa, kY = m.type == HU_TN93 ? m.ky : m.kappa;
		const double a = m.pi[0], c = m.pi[1], g = m.pi[2], t = m.pi[3];
		const double pr = a + g, py = c + t;
		const double e = exp(-m.beta * v);
		const double eR = exp(-(1 + pr * (kR - 1)) * m.beta * v);
		const double eY = exp(-(1 + py * (kY - 1)) * m.beta * v);
		const double tv = 1 - e;
		P[0] = (a * (pr + py * e) + g * eR) / pr;  P[1] = c * tv;  P[2] = (g * (pr + py * e) - g * eR) / pr;  P[3] = t * tv;
		P[4] = a * tv;  P[5] = (c * (py + pr * e) + t * eY) / py;  P[6] = g * tv;  P[7] = (t * (py + pr * e) - t * eY) / py;
		P[8] = (a * (pr + py * e) - a * eR) / pr;  P[9] = c * tv;  P[10] = (g * (pr + py * e) + a * eR) / pr;  P[11] = t * tv;
		P[12] = a * tv;  P[13] = (c * (py + pr * e) - c * eY) / py;  P[14] = g * tv;  P[15] = (t * (py + pr * e) + c * eY) / py;
		if (P[2] < 0) P[2] = 0;
		if (P[7] < 0) P[7] = 0;
		if (P[8] < 0) P[8] = 0;
		if (P[13] < 0) P[13] = 0;
		return;
	}
	case HU_F81: {
		const double e = exp(-m.beta * v);
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) P[i * 4 + j] = (i == j ? e : 0.0) + m.pi[j] * (1 - e);
		return;
	}
	case HU_K80: {
		const double e = exp(-4 * m.beta * v);
		const double eV = exp(-2 * (1 + m.kappa) * m.beta * v);
		const double dg = (1.0 + e + 2 * eV) / 4, ti = (1.0 + e - 2 * eV) / 4, tv = (1.0 - e) / 4;
		for (int i = 0; i < 16; ++i) P[i] = tv;
		for (int i = 0; i < 4; ++i) P[i * 4 + i] = dg;
		P[2] = P[8] = P[7] = P[13] = ti;
		return;
	}
	default: { // JC69
		const double x = exp(-4 * v / 3);
		const double off = (1 - x) / 4, dg = (1 + 3 * x) / 4;
		for (int i = 0; i < 16; ++i) P[i] = off;
		for (int i = 0; i < 4; ++i) P[i * 4 + i] = dg;
		return;
	}
	}
}

// ---------------------------------------------------------------- scaled log-space primitives ----
// src/PhyloTreeUnrooted.h:1488-1529.  MIN_LOGLIK_EXP = DBL_MIN_EXP / 2 = -510 (src/PhyloTreeUnrooted.cpp:68).
#define HU_MIN_LOGLIK_EXP (-510.0)

HU_HD double scaleOf(double mx) { return (mx != -INFINITY && mx < HU_MIN_LOGLIK_EXP) ? HU_MIN_LOGLIK_EXP - mx : 0.0; }
HU_HD double max4(const double* v) { return fmax(fmax(v[0], v[1]), fmax(v[2], v[3])); }
HU_HD int argmax4(const double* v) { // maxCoeff(&idx): first strict maximum
	int b = 0;
	double m = v[0];
	if (v[1] > m) { m = v[1]; b = 1; }
	if (v[2] > m) { m = v[2]; b = 2; }
	if (v[3] > m) { b = 3; }
	return b;
}
// Y = dot_product_scaled(P, V): Y_i = log(sum_k P_ik exp(V_k + s)) - s
HU_HD void dpsMat(const double* P, const double* V, double* Y) {
	const double s = scaleOf(max4(V));
	const double e0 = exp(V[0] + s), e1 = exp(V[1] + s), e2 = exp(V[2] + s), e3 = exp(V[3] + s);
	for (int i = 0; i < 4; ++i)
		Y[i] = log((P[i * 4] * e0 + P[i * 4 + 1] * e1) + (P[i * 4 + 2] * e2 + P[i * 4 + 3] * e3)) - s;
}
HU_HD double dpsVec(const double* pi, const double* V) {
	const double s = scaleOf(max4(V));
	return log((pi[0] * exp(V[0] + s) + pi[2] * exp(V[2] + s)) + (pi[1] * exp(V[1] + s) + pi[3] * exp(V[3] + s))) - s;
}
// leaf message of a base code (src/PhyloTreeUnrooted.h:1431-1437)
HU_HD void leafMsg(const Model& m, int b, double* Y) {
	if (b >= 0) {
		Y[0] = Y[1] = Y[2] = Y[3] = -INFINITY;
		Y[b] = 0.0;
	} else
		for (int k = 0; k < 4; ++k) Y[k] = m.logPi[k];
}

// ---- linear-domain forms used by the estimate / place kernels ----
// A message column V (log domain) enters every SEP formula only as exp(V + s): dot_product_scaled(P, V)_i =
// log(sum_k P_ik e^{V_k+s}) - s.  With s = -max(V) (any s gives the same value up to rounding; the reference
// uses s = 0 unless max(V) < -510) the 4-vector e = exp(V - max V) and the scalar max V carry everything:
//   exp(dps(P, V)_i) = e^{max V} * (P e)_i.
// The kernels therefore do their products and sums on e and P e and take one log per column where a
// log-likelihood is needed, instead of the exp/log round trip per matrix row of the literal code.
HU_HD double expNorm(const double* V, double* e) {
	double mx = max4(V);
	if (mx == -INFINITY) mx = 0.0;
	for (int k = 0; k < 4; ++k) e[k] = exp(V[k] - mx);
	return mx;
}
HU_HD void matVec4(const double* P, const double* e, double* y) {
	for (int i = 0; i < 4; ++i) y[i] = (P[i * 4] * e[0] + P[i * 4 + 1] * e[1]) + (P[i * 4 + 2] * e[2] + P[i * 4 + 3] * e[3]);
}
// exp of the leaf message of a base code: one-hot, or pi for a gap
HU_HD void leafLin(const Model& m, int b, double* y) {
	if (b >= 0) {
		y[0] = y[1] = y[2] = y[3] = 0.0;
		y[b] = 1.0;
	} else
		for (int k = 0; k < 4; ++k) y[k] = m.pi[k];
}
HU_HD double dotPi(const Model& m, const double* g) { return (m.pi[0] * g[0] + m.pi[2] * g[2]) + (m.pi[1] * g[1] + m.pi[3] * g[3]); }

// row_mean_exp_scaled (src/PhyloTreeUnrooted.h:1521-1529) over KK category columns held in mat[k][4]
HU_HD void rmes(const double (*mat)[4], int KK, double* Y) {
	for (int i = 0; i < 4; ++i) {
		double mx = mat[0][i];
		for (int k = 1; k < KK; ++k) mx = fmax(mx, mat[k][i]);
		const double sc = scaleOf(mx);
		double s = 0;
		for (int k = 0; k < KK; ++k) s += exp(mat[k][i] + sc);
		Y[i] = log(s / KK) - sc;
	}
}

#ifdef __CUDACC__
__device__ __forceinline__ void ld4(const double* p, double* v) { // one 4-vector message column, 2 x 128-bit
	const double2 a = __ldg(reinterpret_cast<const double2*>(p));
	const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(double* p, const double* v) {
	reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
	reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
#endif

// ---------------------------------------------------------------- device-resident database ----
struct HmmDev {
	int K, L, Kp;            // Kp: padded row pitch of the per-column tables (multiple of 16)
	const double* eM;        // [4][Kp]
	const double* eI;        // [4][Kp]
	const double* tr;        // [7][Kp]: MM MI MD IM II DM DD
	const double* ent;       // [Kp]
	const double* ext;       // [Kp]
	const int* p2cs;         // [K+1]
};
enum { T_MM = 0, T_MI = 1, T_MD = 2, T_IM = 3, T_II = 4, T_DM = 5, T_DD = 6 };

struct TreeDev {
	int nNodes, csLen, root;
	int nWords;              // ceil(csLen / 32)
	const int* parent;
	const double* blen;
	const double* height;
	const int8_t* seq;       // [nNodes][csLen]
	const int* taxon;
	const double* up;        // [nNodes][csLen][4]: exp(msg(id -> parent) - scale), max-normalised
	const double* dn;        //                     exp(msg(parent -> id) - scale)
	const double* upS;       // [nNodes][csLen]: the scales (max of the log-domain column)
	const double* dnS;
	const double* upP32;     // [nNodes][nWords + 1]: prefix sums of the scales at 32-column granularity,
	const double* dnP32;     //   P32[b] = sum of the scales of columns < 32 b (window sums without streaming the scales)
	const double* litTab;    // [csLen + 1]: W-fold sequential sum of log(pi . e^1): placeSeq's literal return value
	const struct Model* modelDev; // copy of the model in global memory, for out-of-line routines
	// seed scan, in PROFILE coordinates: bit (k-1)%32 of word (k-1)/32 is match column p2cs[k] (k = 1..K);
	// [word][plane][nodeSlot], planes hi, lo, valid; nodeSlot = position in the scan order
	const uint32_t* planes;
	int nPWords;             // ceil(K / 32)
	const int8_t* seqT;      // [csLen][scanPitch]: node codes transposed into slot order (non-match columns are read)
	const int* colK;         // [csLen]: profile position of a match column, 0 otherwise
	int nScan;               // number of scan slots (non-root nodes), padded to scanPitch
	int scanPitch;
	const int* scanNode;     // slot -> node id
	const int* posOfSlot;    // slot -> position in getSeed's candidate vector (non-root nodes in id order), -1 for padding
	const int* nodeOfPos;    // position -> node id, padded to scanPitch with 0x7fffffff
	int nCand;               // number of candidates (nNodes - 1)
	int nGapSlots;           // slots [0, nGapSlots) hold nodes whose sequence has gaps (need the valid plane)
	const double* inv2N;     // [65536]: 2^31 / N correctly rounded (seed keys without a division), inv2N[0] = 0
	// seed scan on the tensor cores (hu_mm.cuh, hu_seedmm.cu); mmA == NULL: not built (profile longer than HU_MM_MAX_N)
	const int8_t* mmA;       // [mmTiles][mmChunks][16 KB]: operand images, rows = candidate positions (non-root nodes in id order)
	const uint4* mmKeyTab;   // [HU_MM_MAX_N + 1]: {Q, R, m, k - 32} per N, the integer form of the seed key (hu_seedmm.cu)
	const int8_t* seqTP;     // [csLen][posPitch]: node codes transposed into position order (non-match columns are read)
	const int8_t* mmColDefault; // [csLen]: -1 match column, 0..3 majority base of the gap-free nodes, 4: long exception list
	const int* mmExcPtr;     // [csLen + 1], mmExcPos: positions that deviate from their class default in a non-match column
	const int* mmExcPos;
	const int* mmSlotOfPos;  // [posPitch]: position -> slot of the bit planes
	int mmTiles, mmChunks;   // ceil(nCand / 128), ceil(K / 32) (+ 1 extension chunk per tile image)
	int posPitch;            // 128 mmTiles
};

#ifdef __CUDACC__
// sum of the scales of columns start..end of one node row, spread over the lanes of a warp
__device__ __forceinline__ double windowScale(const double* __restrict__ S, const double* __restrict__ P32, int node, int csLen,
		int nWords, int start, int end, int lane) {
	const double* s = S + (size_t) node * csLen;
	const int b0 = (start + 31) >> 5, b1 = (end + 1) >> 5;
	double v = 0;
	if (b0 <= b1) {
		const double* p = P32 + (size_t) node * (nWords + 1);
		if (lane == 0) v = __ldg(p + b1) - __ldg(p + b0);
		int c = start + lane;
		if (c < 32 * b0) v += __ldg(s + c);
		c = 32 * b1 + lane;
		if (c <= end) v += __ldg(s + c);
	} else
		for (int c = start + lane; c <= end; c += 32) v += __ldg(s + c);
	return v;
}
#endif

// per-device kernel attributes (dynamic shared memory limits): set for the current device by hu_db_create
int configViterbiKernels();
int configSeedKernels();
int configEstimateKernels();
int configPlaceKernels();
int configSeedMmKernels();

} // namespace hu

struct hu_db {
	int device;
	hu::Model model;
	hu::HmmDev hmm;
	hu::TreeDev tree;
	std::vector<void*> allocs;
	std::vector<size_t> allocBytes; // size of allocs[i] (hu_db_clone copies the replica allocation by allocation)
	int64_t bytes;
	int deferred;                   // created without messages: hu_db_put_messages + hu_db_finish complete it
	// host copies used by the facades / stage planning
	int K, L;
	std::vector<int> p2cs;
	std::vector<double> heightHost;
	std::vector<int> scanNodeHost; // slot -> node id of the seed scan order (-1: padding)
	hu_db() : device(0), bytes(0), deferred(0), K(0), L(0) {}
};

namespace hu {
int buildSeedMm(hu_db* db, const int8_t* seq, int K, int8_t* A, int8_t* seqTP, const int* posNode, const int8_t* posClass);
void seedMmColumns(const int8_t* hs, int csLen, const int* colK, const std::vector<int>& posNode, int nCand,
		std::vector<int8_t>& posClass, std::vector<int8_t>& colDefault, std::vector<int>& excPtr, std::vector<int>& excPos);
}
