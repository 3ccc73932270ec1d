// Database replica on one device: profile tables, tree, node sequences (+ bit planes for the seed
// scan), substitution model and the 4 x csLen conditional-likelihood messages of every directed edge.
// Mirrors what hmmufotu loads from <db>.hmm / <db>.ptu (src/hmmufotu.cpp:457-494) and, when the
// messages are not supplied, what hmmufotu-build computes by re-rooting at every node
// (src/hmmufotu-build.cpp:455-459, src/PhyloTreeUnrooted.cpp:320-374).
#include "hu_internal.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <cmath>

namespace hu {

static thread_local char g_err[512] = "";
void setError(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

// ---------------------------------------------------------------- model (host) ----
// Symmetric 4x4 eigen-solver (cyclic Jacobi rotations).  The reference calls Eigen::EigenSolver on
// the trace-scaled Q (src/GTR.cpp:124-145); P(t) does not depend on the eigenvector basis.
static void symEig4(double a[4][4], double w[4], double v[4][4]) {
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) v[i][j] = i == j;
	for (int sweep = 0; sweep < 100; ++sweep) {
		double off = 0;
		for (int i = 0; i < 4; ++i)
			for (int j = i + 1; j < 4; ++j) off += fabs(a[i][j]);
		if (off < 1e-290) break;
		for (int p = 0; p < 3; ++p)
			for (int q = p + 1; q < 4; ++q) {
				if (a[p][q] == 0.0) continue;
				const double tau = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
				const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
				const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
				for (int k = 0; k < 4; ++k) { // A <- A J
					const double x = a[k][p], y = a[k][q];
					a[k][p] = c * x - s * y;
					a[k][q] = s * x + c * y;
				}
				for (int k = 0; k < 4; ++k) { // A <- J^T A
					const double x = a[p][k], y = a[q][k];
					a[p][k] = c * x - s * y;
					a[q][k] = s * x + c * y;
				}
				for (int k = 0; k < 4; ++k) {
					const double x = v[k][p], y = v[k][q];
					v[k][p] = c * x - s * y;
					v[k][q] = s * x + c * y;
				}
			}
	}
	for (int i = 0; i < 4; ++i) w[i] = a[i][i];
}

int buildModel(const hu_model_desc* d, Model* m) {
	memset(m, 0, sizeof *m);
	m->type = d->type;
	m->dgK = d->dgK;
	if (d->type < HU_GTR || d->type > HU_JC69) { setError("unknown substitution model type %d", d->type); return HU_ERR_ARG; }
	if (d->dgK < 0 || d->dgK > HU_MAX_DG) { setError("discrete Gamma K=%d out of range", d->dgK); return HU_ERR_ARG; }
	for (int i = 0; i < 4; ++i) m->pi[i] = d->pi[i];
	m->kr = d->kr; m->ky = d->ky; m->kappa = d->kappa; m->beta = d->beta;
	if (d->type == HU_K80) { // src/K80.h:97-100
		m->beta = 1.0 / (2.0 * d->kappa);
		for (int i = 0; i < 4; ++i) m->pi[i] = 0.25;
	}
	if (d->type == HU_JC69)
		for (int i = 0; i < 4; ++i) m->pi[i] = 0.25;
	for (int i = 0; i < 4; ++i) m->logPi[i] = log(m->pi[i]);
	for (int k = 0; k < d->dgK; ++k) m->dgR[k] = d->dgRates[k];
	if (d->type == HU_GTR) {
		// Q = R * diag(pi), diagonal = -row sum, scaled so that -trace(Q) = 1 (src/DNASubModel.cpp:123-126)
		double Q[4][4];
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) Q[i][j] = i == j ? 0.0 : d->R[i * 4 + j] * m->pi[j];
		double tr = 0;
		for (int i = 0; i < 4; ++i) {
			double rs = 0;
			for (int j = 0; j < 4; ++j) rs += Q[i][j];
			Q[i][i] = -rs;
			tr += Q[i][i];
		}
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) Q[i][j] /= -tr;
		// reversible: S = D^1/2 Q D^-1/2 is symmetric, Q = D^-1/2 W diag(w) W^T D^1/2
		double S[4][4], W[4][4], w[4], sq[4];
		for (int i = 0; i < 4; ++i) sq[i] = sqrt(m->pi[i]);
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) S[i][j] = sq[i] * Q[i][j] / sq[j];
		for (int i = 0; i < 4; ++i)
			for (int j = i + 1; j < 4; ++j) S[i][j] = S[j][i] = 0.5 * (S[i][j] + S[j][i]);
		symEig4(S, w, W);
		for (int k = 0; k < 4; ++k) m->lambda[k] = w[k];
		for (int i = 0; i < 4; ++i)
			for (int k = 0; k < 4; ++k) {
				m->U[i * 4 + k] = W[i][k] / sq[i];
				m->Ui[k * 4 + i] = W[i][k] * sq[i];
			}
	}
	return HU_OK;
}

// ---------------------------------------------------------------- tree evaluation kernels ----
// per-edge transition matrices for every rate category: PP[node][k][16] = P(blen[node] * rate_k)
__global__ void edgePrKernel(Model m, const double* __restrict__ blen, int nNodes, int KK, double* __restrict__ PP) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= nNodes * KK) return;
	const int node = idx / KK, k = idx % KK;
	double P[16];
	modelPr(m, blen[node] * (m.dgK ? m.dgR[k] : 1.0), P);
	for (int i = 0; i < 16; ++i) PP[(size_t) idx * 16 + i] = P[i];
}

// Messages live on the device in the linear form of hu_internal.cuh (expNorm): E = exp(V - max V) (4 doubles per
// column) and the scale max V (1 double per column).  dot_product_scaled(P, V)_i = log((P E)_i) + scale.
__device__ __forceinline__ void convMsg(const double* P, const double* E, double sc, double* y) {
	double t[4];
	matVec4(P, E, t);
	for (int i = 0; i < 4; ++i) y[i] = log(t[i]) + sc;
}
__device__ __forceinline__ void storeMsg(double* E, double* S, size_t idx, const double* o) {
	double e[4];
	const double mx = expNorm(o, e);
	st4(E + idx * 4, e);
	S[idx] = mx;
}

// Upward pass for the nodes of one depth level: up[x] = loglik(x, j) with x's parent as root side.
__global__ void upLevelKernel(Model m, const int* __restrict__ levelNodes, int nLevel, int csLen,
		const int* __restrict__ childStart, const int* __restrict__ childList, const int8_t* __restrict__ seq,
		const double* __restrict__ PP, int KK, double* __restrict__ up, double* __restrict__ upS, int isRootLevel,
		double* __restrict__ rootMsg) {
	const int x = levelNodes[blockIdx.y];
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= csLen) return;
	const int c0 = childStart[x], c1 = childStart[x + 1];
	double mat[HU_MAX_DG][4];
	for (int k = 0; k < KK; ++k) mat[k][0] = mat[k][1] = mat[k][2] = mat[k][3] = 0.0;
	for (int c = c0; c < c1; ++c) {
		const int ch = childList[c];
		double v[4], y[4];
		const size_t idx = (size_t) ch * csLen + j;
		ld4(up + idx * 4, v);
		const double sc = upS[idx];
		for (int k = 0; k < KK; ++k) {
			convMsg(PP + ((size_t) ch * KK + k) * 16, v, sc, y);
			for (int i = 0; i < 4; ++i) mat[k][i] += y[i];
		}
	}
	double o[4];
	const bool leaf = c0 == c1;
	if (!leaf && m.dgK) rmes(mat, KK, o);
	else for (int i = 0; i < 4; ++i) o[i] = mat[0][i];
	if (leaf) {
		double n[4];
		leafMsg(m, seq[(size_t) x * csLen + j], n);
		for (int i = 0; i < 4; ++i) o[i] += n[i];
	}
	if (isRootLevel) st4(rootMsg + (size_t) j * 4, o); // log domain: only its arg-max is used (inferSeq)
	else storeMsg(up, upS, (size_t) x * csLen + j, o);
}

// Downward pass: dn[x] = loglik(parent, j) seen from x (siblings' up messages + the parent's own dn)
__global__ void dnLevelKernel(Model m, const int* __restrict__ levelNodes, int nLevel, int csLen, int root,
		const int* __restrict__ parent, const int* __restrict__ childStart, const int* __restrict__ childList,
		const double* __restrict__ PP, int KK, const double* __restrict__ up, const double* __restrict__ upS,
		double* __restrict__ dn, double* __restrict__ dnS) {
	const int x = levelNodes[blockIdx.y];
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= csLen) return;
	const int pa = parent[x];
	double mat[HU_MAX_DG][4];
	for (int k = 0; k < KK; ++k) mat[k][0] = mat[k][1] = mat[k][2] = mat[k][3] = 0.0;
	for (int c = childStart[pa]; c < childStart[pa + 1]; ++c) {
		const int sib = childList[c];
		if (sib == x) continue;
		double v[4], y[4];
		const size_t idx = (size_t) sib * csLen + j;
		ld4(up + idx * 4, v);
		const double sc = upS[idx];
		for (int k = 0; k < KK; ++k) {
			convMsg(PP + ((size_t) sib * KK + k) * 16, v, sc, y);
			for (int i = 0; i < 4; ++i) mat[k][i] += y[i];
		}
	}
	if (pa != root) {
		double v[4], y[4];
		const size_t idx = (size_t) pa * csLen + j;
		ld4(dn + idx * 4, v);
		const double sc = dnS[idx];
		for (int k = 0; k < KK; ++k) {
			convMsg(PP + ((size_t) pa * KK + k) * 16, v, sc, y);
			for (int i = 0; i < 4; ++i) mat[k][i] += y[i];
		}
	}
	const int deg = childStart[pa + 1] - childStart[pa] + (pa == root ? 0 : 1);
	double o[4];
	if (deg != 1 && m.dgK) rmes(mat, KK, o);
	else for (int i = 0; i < 4; ++i) o[i] = mat[0][i];
	storeMsg(dn, dnS, (size_t) x * csLen + j, o);
}

// log-domain messages supplied by the host -> linear form, in place over a staging chunk
__global__ void toLinearKernel(const double* __restrict__ logMsg, size_t nCols, double* __restrict__ E, double* __restrict__ S) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nCols) return;
	double v[4];
	ld4(logMsg + i * 4, v);
	storeMsg(E, S, i, v);
}
// and back, for hu_db_get_messages: V_k = log(E_k) + scale
__global__ void toLogKernel(const double* __restrict__ E, const double* __restrict__ S, size_t nCols, double* __restrict__ logMsg) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nCols) return;
	double e[4], v[4];
	ld4(E + i * 4, e);
	const double sc = S[i];
	for (int k = 0; k < 4; ++k) v[k] = log(e[k]) + sc;
	st4(logMsg + i * 4, v);
}

// prefix sums of the per-column scales at 32-column granularity, one warp per node row:
// P32[row][b] = sum of S[row][c] over c < 32 b.  The SEP kernels need the scales only summed over a window.
__global__ void scalePrefixKernel(const double* __restrict__ S, long long nRows, int csLen, int nWords, double* __restrict__ P32) {
	const long long row = (long long) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (row >= nRows) return;
	const double* s = S + (size_t) row * csLen;
	double* out = P32 + (size_t) row * (nWords + 1);
	double run = 0;
	if (lane == 0) out[0] = 0;
	for (int b = 0; b < nWords; ++b) {
		const int c = b * 32 + lane;
		double v = c < csLen ? s[c] : 0.0;
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		run += v;
		if (lane == 0) out[b + 1] = run;
	}
}

// placeSeq()'s literal return value for a window of W columns: the reference adds log(pi . e^1) W times
// (src/PhyloTreeUnrooted.cpp:707-712 over a root message of INVALID_LOGLIK)
__global__ void litTabKernel(Model m, int csLen, double* __restrict__ tab) {
	if (blockIdx.x || threadIdx.x) return;
	const double ones[4] = {1.0, 1.0, 1.0, 1.0};
	const double c1 = dpsVec(m.pi, ones);
	double lit = 0;
	tab[0] = 0;
	for (int j = 1; j <= csLen; ++j) { lit += c1; tab[j] = lit; }
}

// inferSeq(): arg-max of the upward message for internal nodes (src/PhyloTreeUnrooted.cpp:1085-1093)
__global__ void inferSeqKernel(int nNodes, int csLen, int root, const int* __restrict__ childStart,
		const double* __restrict__ up, const double* __restrict__ rootMsg, int8_t* __restrict__ seq) {
	const int x = blockIdx.y;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= csLen || childStart[x] == childStart[x + 1]) return;
	double v[4];
	ld4((x == root ? rootMsg : up + (size_t) x * csLen * 4) + (size_t) j * 4, v);
	seq[(size_t) x * csLen + j] = (int8_t) argmax4(v);
}

// node sequences -> bit planes over the match columns, [word][plane][slot] (hi, lo, valid)
__global__ void packPlanesKernel(const int8_t* __restrict__ seq, int csLen, int K, const int* __restrict__ p2cs,
		const int* __restrict__ scanNode, int nScan, int scanPitch, uint32_t* __restrict__ planes) {
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	const int w = blockIdx.y;
	if (slot >= scanPitch) return;
	uint32_t hi = 0, lo = 0, va = 0;
	const int node = slot < nScan ? scanNode[slot] : -1;
	if (node >= 0) {
		const int8_t* s = seq + (size_t) node * csLen;
		for (int b = 0; b < 32; ++b) {
			const int k = w * 32 + b + 1;
			if (k > K) break;
			const int c = s[p2cs[k] - 1];
			if (c >= 0) {
				va |= 1u << b;
				hi |= (uint32_t) ((c >> 1) & 1) << b;
				lo |= (uint32_t) (c & 1) << b;
			}
		}
	}
	planes[((size_t) w * 3 + 0) * scanPitch + slot] = hi;
	planes[((size_t) w * 3 + 1) * scanPitch + slot] = lo;
	planes[((size_t) w * 3 + 2) * scanPitch + slot] = va;
}

// seq[node][col] -> seqT[col][slot] (32 x 32 tiles through shared memory)
__global__ void transposeSeqKernel(const int8_t* __restrict__ seq, int csLen, const int* __restrict__ scanNode, int scanPitch,
		int8_t* __restrict__ seqT) {
	__shared__ int8_t tile[32][33];
	const int slot0 = blockIdx.x * 32, col0 = blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int slot = slot0 + r, col = col0 + threadIdx.x;
		const int node = slot < scanPitch ? scanNode[slot] : -1;
		tile[r][threadIdx.x] = (node >= 0 && col < csLen) ? seq[(size_t) node * csLen + col] : (int8_t) -2;
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += blockDim.y) {
		const int col = col0 + r, slot = slot0 + threadIdx.x;
		if (col < csLen && slot < scanPitch) seqT[(size_t) col * scanPitch + slot] = tile[threadIdx.x][r];
	}
}

template <typename T>
static int devAlloc(hu_db* db, T** p, size_t n) {
	void* q = NULL;
	cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
	if (e != cudaSuccess) {
		setError("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
		return HU_ERR_NOMEM;
	}
	db->allocs.push_back(q);
	db->allocBytes.push_back(std::max<size_t>(n, 1) * sizeof(T));
	db->bytes += n * sizeof(T);
	*p = (T*) q;
	return HU_OK;
}
#define HU_TRY(x) do { int rc_ = (x); if (rc_ != HU_OK) return rc_; } while (0)

template <typename T>
static int upload(hu_db* db, const T** dst, const T* src, size_t n) {
	T* p;
	HU_TRY(devAlloc(db, &p, n));
	HU_CUDA(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
	*dst = p;
	return HU_OK;
}

static int finishMessages(hu_db* db);

static int createImpl(hu_db* db, const hu_hmm_params* hp, const hu_tree_desc* td, const hu_model_desc* md, bool deferMessages) {
	HU_TRY(buildModel(md, &db->model));
	// ---- profile: SoA tables with a padded pitch ----
	const int K = hp->K, Kp = ((K + 1 + 15) / 16) * 16;
	db->K = K; db->L = hp->L;
	db->p2cs.assign(hp->p2cs, hp->p2cs + K + 1);
	std::vector<double> eM((size_t) 4 * Kp, INFINITY), eI((size_t) 4 * Kp, INFINITY), tr((size_t) 7 * Kp, INFINITY), ent(Kp, INFINITY), ext(Kp, INFINITY);
	static const int tIdx[7] = {0 * 3 + 0, 0 * 3 + 1, 0 * 3 + 2, 1 * 3 + 0, 1 * 3 + 1, 2 * 3 + 0, 2 * 3 + 2};
	for (int k = 0; k <= K; ++k) {
		for (int b = 0; b < 4; ++b) {
			eM[(size_t) b * Kp + k] = hp->eM[(size_t) k * 4 + b];
			eI[(size_t) b * Kp + k] = hp->eI[(size_t) k * 4 + b];
		}
		for (int t = 0; t < 7; ++t) tr[(size_t) t * Kp + k] = hp->T[(size_t) k * 9 + tIdx[t]];
		ent[k] = hp->entry[k];
		ext[k] = hp->exit[k];
	}
	db->hmm.K = K; db->hmm.L = hp->L; db->hmm.Kp = Kp;
	HU_TRY(upload(db, &db->hmm.eM, eM.data(), eM.size()));
	HU_TRY(upload(db, &db->hmm.eI, eI.data(), eI.size()));
	HU_TRY(upload(db, &db->hmm.tr, tr.data(), tr.size()));
	HU_TRY(upload(db, &db->hmm.ent, ent.data(), ent.size()));
	HU_TRY(upload(db, &db->hmm.ext, ext.data(), ext.size()));
	HU_TRY(upload(db, &db->hmm.p2cs, hp->p2cs, (size_t) K + 1));
	// ---- tree ----
	const int n = (int) td->nNodes, csLen = td->csLen;
	if (td->nNodes <= 1 || td->nNodes > (1 << 30) || csLen <= 0 || csLen > 65535) { setError("bad tree dimensions"); return HU_ERR_ARG; }
	if (hp->L != csLen) { setError("profile MAXL %d != tree csLen %d", hp->L, csLen); return HU_ERR_ARG; }
	TreeDev& t = db->tree;
	t.nNodes = n; t.csLen = csLen; t.root = td->root; t.nWords = (csLen + 31) / 32;
	HU_TRY(upload(db, &t.parent, td->parent, (size_t) n));
	HU_TRY(upload(db, &t.blen, td->blen, (size_t) n));
	HU_TRY(upload(db, &t.height, td->height, (size_t) n));
	HU_TRY(upload(db, &t.taxon, td->taxon, (size_t) n));
	db->heightHost.assign(td->height, td->height + n);
	int8_t* seq;
	HU_TRY(devAlloc(db, &seq, (size_t) n * csLen + 64)); // + slack: bulk copies of a window are rounded to 16 bytes
	HU_CUDA(cudaMemset(seq + (size_t) n * csLen, 0xfe, 64));
	HU_CUDA(cudaMemcpy(seq, td->seq, (size_t) n * csLen, cudaMemcpyHostToDevice));
	t.seq = seq;
	double *up, *dn, *upS, *dnS;
	const size_t msgN = (size_t) n * csLen * 4;
	HU_TRY(devAlloc(db, &up, msgN));
	HU_TRY(devAlloc(db, &dn, msgN));
	HU_TRY(devAlloc(db, &upS, msgN / 4));
	HU_TRY(devAlloc(db, &dnS, msgN / 4));
	t.up = up; t.dn = dn; t.upS = upS; t.dnS = dnS;
	// children in CSR form (neighbour order = ascending id, as built from a parent vector)
	std::vector<int> childStart(n + 1, 0), childList(std::max(0, n - 1)), depth(n, -1);
	for (int x = 0; x < n; ++x)
		if (td->parent[x] >= 0) childStart[td->parent[x] + 1]++;
	for (int x = 0; x < n; ++x) childStart[x + 1] += childStart[x];
	{
		std::vector<int> fill(childStart.begin(), childStart.end() - 1);
		for (int x = 0; x < n; ++x)
			if (td->parent[x] >= 0) childList[fill[td->parent[x]]++] = x;
	}
	if ((td->up == NULL) != (td->dn == NULL)) { setError("up and dn must both be given or both be NULL"); return HU_ERR_ARG; }
	db->deferred = deferMessages ? 1 : 0;
	if (deferMessages) {
		// hu_db_put_messages fills the messages later; rows nobody supplies (the root's) stay defined
		HU_CUDA(cudaMemset(up, 0, msgN * sizeof(double)));
		HU_CUDA(cudaMemset(dn, 0, msgN * sizeof(double)));
		HU_CUDA(cudaMemset(upS, 0, msgN / 4 * sizeof(double)));
		HU_CUDA(cudaMemset(dnS, 0, msgN / 4 * sizeof(double)));
	} else if (td->up) {
		// staged conversion of the .ptu's log-domain messages, 16 M columns (512 MB) at a time
		const size_t nCols = msgN / 4, chunk = (size_t) 16 << 20;
		double* stage;
		HU_CUDA(cudaMalloc(&stage, std::min(chunk, nCols) * 4 * sizeof(double)));
		for (int dir = 0; dir < 2; ++dir)
			for (size_t o = 0; o < nCols; o += chunk) {
				const size_t c = std::min(chunk, nCols - o);
				HU_CUDA(cudaMemcpy(stage, (dir ? td->dn : td->up) + o * 4, c * 4 * sizeof(double), cudaMemcpyHostToDevice));
				toLinearKernel<<<(unsigned) ((c + 255) / 256), 256>>>(stage, c, (dir ? dn : up) + o * 4, (dir ? dnS : upS) + o);
			}
		HU_CUDA(cudaGetLastError());
		HU_CUDA(cudaDeviceSynchronize());
		cudaFree(stage);
	} else {
		// depth levels by breadth-first traversal from the root
		std::vector<int> order;
		order.reserve(n);
		order.push_back(td->root);
		depth[td->root] = 0;
		for (size_t h = 0; h < order.size(); ++h) {
			const int x = order[h];
			for (int c = childStart[x]; c < childStart[x + 1]; ++c) {
				depth[childList[c]] = depth[x] + 1;
				order.push_back(childList[c]);
			}
		}
		if ((int) order.size() != n) { setError("parent vector does not describe a tree rooted at %d", td->root); return HU_ERR_ARG; }
		const int maxDepth = depth[order.back()];
		std::vector<int> levelStart(maxDepth + 2, 0);
		for (int x = 0; x < n; ++x) levelStart[depth[x] + 1]++;
		for (int d = 0; d <= maxDepth; ++d) levelStart[d + 1] += levelStart[d];
		int *dChildStart, *dChildList, *dOrder;
		double *dPP, *dRootMsg;
		const int KK = std::max(1, db->model.dgK);
		// scratch (freed below): not part of the replica
		HU_CUDA(cudaMalloc(&dChildStart, (n + 1) * sizeof(int)));
		HU_CUDA(cudaMalloc(&dChildList, std::max(1, n - 1) * sizeof(int)));
		HU_CUDA(cudaMalloc(&dOrder, n * sizeof(int)));
		HU_CUDA(cudaMalloc(&dPP, (size_t) n * KK * 16 * sizeof(double)));
		HU_CUDA(cudaMalloc(&dRootMsg, (size_t) csLen * 4 * sizeof(double)));
		HU_CUDA(cudaMemcpy(dChildStart, childStart.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice));
		HU_CUDA(cudaMemcpy(dChildList, childList.data(), (n - 1) * sizeof(int), cudaMemcpyHostToDevice));
		HU_CUDA(cudaMemcpy(dOrder, order.data(), n * sizeof(int), cudaMemcpyHostToDevice)); // BFS order is level-sorted
		edgePrKernel<<<(n * KK + 127) / 128, 128>>>(db->model, t.blen, n, KK, dPP);
		const int TB = 128;
		const dim3 gx((csLen + TB - 1) / TB);
		for (int d = maxDepth; d >= 0; --d) {
			const int cnt = levelStart[d + 1] - levelStart[d];
			for (int o = 0; o < cnt; o += 65535) {
				const int c = std::min(65535, cnt - o);
				upLevelKernel<<<dim3(gx.x, c), TB>>>(db->model, dOrder + levelStart[d] + o, c, csLen, dChildStart, dChildList, seq, dPP, KK, up, upS, d == 0, dRootMsg);
			}
		}
		if (td->inferInternalSeq)
			for (int o = 0; o < n; o += 65535) {
				// grid.y indexes nodes o..o+c-1 through pointer offsets
				const int c = std::min(65535, n - o);
				inferSeqKernel<<<dim3(gx.x, c), TB>>>(c, csLen, td->root - o, dChildStart + o, up + (size_t) o * csLen * 4, dRootMsg, seq + (size_t) o * csLen);
			}
		for (int d = 1; d <= maxDepth; ++d) {
			const int cnt = levelStart[d + 1] - levelStart[d];
			for (int o = 0; o < cnt; o += 65535) {
				const int c = std::min(65535, cnt - o);
				dnLevelKernel<<<dim3(gx.x, c), TB>>>(db->model, dOrder + levelStart[d] + o, c, csLen, td->root, t.parent, dChildStart, dChildList, dPP, KK, up, upS, dn, dnS);
			}
		}
		HU_CUDA(cudaGetLastError());
		HU_CUDA(cudaDeviceSynchronize());
		cudaFree(dChildStart); cudaFree(dChildList); cudaFree(dOrder); cudaFree(dPP); cudaFree(dRootMsg);
		// the root has no parent edge: keep its rows defined
		HU_CUDA(cudaMemset(up + (size_t) td->root * csLen * 4, 0, (size_t) csLen * 4 * sizeof(double)));
		HU_CUDA(cudaMemset(dn + (size_t) td->root * csLen * 4, 0, (size_t) csLen * 4 * sizeof(double)));
		HU_CUDA(cudaMemset(upS + (size_t) td->root * csLen, 0, (size_t) csLen * sizeof(double)));
		HU_CUDA(cudaMemset(dnS + (size_t) td->root * csLen, 0, (size_t) csLen * sizeof(double)));
	}
	// ---- derived tables of the SEP kernels ----
	{
		double *upP, *dnP, *lit;
		Model* md;
		HU_TRY(devAlloc(db, &upP, (size_t) n * (t.nWords + 1)));
		HU_TRY(devAlloc(db, &dnP, (size_t) n * (t.nWords + 1)));
		HU_TRY(devAlloc(db, &lit, (size_t) csLen + 1));
		HU_TRY(devAlloc(db, &md, 1));
		litTabKernel<<<1, 32>>>(db->model, csLen, lit);
		HU_CUDA(cudaMemcpy(md, &db->model, sizeof(Model), cudaMemcpyHostToDevice));
		HU_CUDA(cudaGetLastError());
		HU_CUDA(cudaDeviceSynchronize());
		t.upP32 = upP; t.dnP32 = dnP; t.litTab = lit; t.modelDev = md;
		if (!deferMessages) HU_TRY(finishMessages(db));
	}
	// ---- scan order + bit planes (after sequence inference) ----
	std::vector<int8_t> hseq;
	const int8_t* hs = td->seq;
	if (!td->up && td->inferInternalSeq) {
		hseq.resize((size_t) n * csLen);
		HU_CUDA(cudaMemcpy(hseq.data(), seq, hseq.size(), cudaMemcpyDeviceToHost));
		hs = hseq.data();
	}
	std::vector<int> gapNodes, fullNodes;
	for (int x = 0; x < n; ++x) {
		if (x == td->root) continue;
		const int8_t* s = hs + (size_t) x * csLen;
		bool gap = false; // a gap in any MATCH column: the node needs the valid plane
		for (int k = 1; k <= K && !gap; ++k) gap = s[hp->p2cs[k] - 1] < 0;
		(gap ? gapNodes : fullNodes).push_back(x);
	}
	// gap block padded to a multiple of the scan tile so that every tile is homogeneous
	const int TILE = 512;
	std::vector<int> scanNode;
	scanNode.insert(scanNode.end(), gapNodes.begin(), gapNodes.end());
	while (scanNode.size() % TILE) scanNode.push_back(-1);
	t.nGapSlots = (int) scanNode.size();
	scanNode.insert(scanNode.end(), fullNodes.begin(), fullNodes.end());
	t.nScan = (int) scanNode.size();
	while (scanNode.size() % TILE) scanNode.push_back(-1);
	t.scanPitch = (int) scanNode.size();
	// padding slots (inside the gap block and at the tail) carry node -1 and all-zero planes
	HU_TRY(upload(db, &t.scanNode, scanNode.data(), scanNode.size()));
	db->scanNodeHost = scanNode;
	{
		// getSeed() visits the non-root nodes in id order: that is the order of the per-read key vector
		std::vector<int> posOfNode(n, -1), posOfSlot(scanNode.size(), -1), nodeOfPos(scanNode.size(), 0x7fffffff);
		int pos = 0;
		for (int x = 0; x < n; ++x)
			if (x != td->root) { posOfNode[x] = pos; nodeOfPos[pos] = x; pos++; }
		for (size_t sl = 0; sl < scanNode.size(); ++sl)
			if (scanNode[sl] >= 0) posOfSlot[sl] = posOfNode[scanNode[sl]];
		t.nCand = pos;
		HU_TRY(upload(db, &t.posOfSlot, posOfSlot.data(), posOfSlot.size()));
		HU_TRY(upload(db, &t.nodeOfPos, nodeOfPos.data(), nodeOfPos.size()));
	}
	{
		std::vector<double> inv(65536, 0.0);
		for (int i = 1; i < 65536; ++i) inv[i] = ldexp(1.0 / (double) i, 31); // scaling by 2^31 is exact
		HU_TRY(upload(db, &t.inv2N, inv.data(), inv.size()));
	}
	uint32_t* planes;
	t.nPWords = (K + 31) / 32;
	HU_TRY(devAlloc(db, &planes, (size_t) t.nPWords * 3 * t.scanPitch));
	int8_t* seqT;
	HU_TRY(devAlloc(db, &seqT, (size_t) csLen * t.scanPitch));
	std::vector<int> colK(csLen, 0);
	for (int k = 1; k <= K; ++k) {
		if (hp->p2cs[k] < 1 || hp->p2cs[k] > csLen) { setError("profile2CSIdx[%d] = %d outside 1..%d", k, hp->p2cs[k], csLen); return HU_ERR_ARG; }
		colK[hp->p2cs[k] - 1] = k;
	}
	HU_TRY(upload(db, &t.colK, colK.data(), colK.size()));
	{
		packPlanesKernel<<<dim3((t.scanPitch + 127) / 128, t.nPWords), 128>>>(seq, csLen, K, db->hmm.p2cs, t.scanNode, t.nScan, t.scanPitch, planes);
		transposeSeqKernel<<<dim3(t.scanPitch / 32, (csLen + 31) / 32), dim3(32, 8)>>>(seq, csLen, t.scanNode, t.scanPitch, seqT);
		HU_CUDA(cudaGetLastError());
		HU_CUDA(cudaDeviceSynchronize());
	}
	t.planes = planes;
	t.seqT = seqT;
	// ---- the tensor-core form of the scan: operand images in candidate order (hu_seedmm.cu) ----
	t.mmA = NULL; t.seqTP = NULL; t.mmKeyTab = NULL; t.mmColDefault = NULL; t.mmExcPtr = t.mmExcPos = t.mmSlotOfPos = NULL; t.mmTiles = t.mmChunks = t.posPitch = 0;
	if (K <= HU_MM_MAX_N - 64) {
		t.mmTiles = (t.nCand + HU_MM_BM - 1) / HU_MM_BM;
		t.mmChunks = (K + 31) / 32;
		t.posPitch = t.mmTiles * HU_MM_BM;
		std::vector<int> posNode(t.posPitch, -1), slotOfPos(t.posPitch, 0);
		for (int x = 0, pos = 0; x < n; ++x)
			if (x != td->root) posNode[pos++] = x;
		for (size_t sl = 0; sl < scanNode.size(); ++sl)
			if (scanNode[sl] >= 0) {
				int pos = scanNode[sl] - (scanNode[sl] > td->root ? 1 : 0);
				slotOfPos[pos] = (int) sl;
			}
		std::vector<int8_t> posClass, colDefault;
		std::vector<int> excPtr, excPos;
		seedMmColumns(hs, csLen, colK.data(), posNode, t.nCand, posClass, colDefault, excPtr, excPos);
		const int* posNodeDev;
		const int8_t* posClassDev;
		HU_TRY(upload(db, &posNodeDev, posNode.data(), posNode.size()));
		HU_TRY(upload(db, &posClassDev, posClass.data(), posClass.size()));
		HU_TRY(upload(db, &t.mmSlotOfPos, slotOfPos.data(), slotOfPos.size()));
		HU_TRY(upload(db, &t.mmColDefault, colDefault.data(), colDefault.size()));
		HU_TRY(upload(db, &t.mmExcPtr, excPtr.data(), excPtr.size()));
		HU_TRY(upload(db, &t.mmExcPos, excPos.data(), excPos.size()));
		int8_t *mmA, *seqTP;
		HU_TRY(devAlloc(db, &mmA, (size_t) t.mmTiles * (t.mmChunks + 1) * HU_MM_A_BYTES));
		HU_TRY(devAlloc(db, &seqTP, (size_t) csLen * t.posPitch));
		HU_TRY(buildSeedMm(db, seq, K, mmA, seqTP, posNodeDev, posClassDev));
		t.mmA = mmA; t.seqTP = seqTP;
		std::vector<uint4> keyTab(HU_MM_MAX_N + 1, make_uint4(0, 0, 0, 0));
		for (uint32_t N = 1; N <= HU_MM_MAX_N; ++N) {
			int bits = 0;
			while ((1u << bits) <= N) ++bits;
			const int k = std::max(32, 24 + bits);
			const unsigned __int128 one = 1;
			const uint64_t m = N == 1 ? 0 : (uint64_t) (((one << k) + N - 1) / N);
			keyTab[N] = make_uint4((uint32_t) ((1ull << 31) / N), (uint32_t) ((1ull << 31) % N), (uint32_t) m, (uint32_t) (k - 32));
		}
		HU_TRY(upload(db, &t.mmKeyTab, keyTab.data(), keyTab.size()));
	}
	return HU_OK;
}

// what depends on the messages: prefix sums of the per-column scales
static int finishMessages(hu_db* db) {
	TreeDev& t = db->tree;
	scalePrefixKernel<<<(t.nNodes + 7) / 8, 256>>>(t.upS, t.nNodes, t.csLen, t.nWords, const_cast<double*>(t.upP32));
	scalePrefixKernel<<<(t.nNodes + 7) / 8, 256>>>(t.dnS, t.nNodes, t.csLen, t.nWords, const_cast<double*>(t.dnP32));
	HU_CUDA(cudaGetLastError());
	HU_CUDA(cudaDeviceSynchronize());
	return HU_OK;
}

} // namespace hu

extern "C" {

const char* hu_last_error(void) { return hu::g_err; }
const char* hu_version(void) { return "hmmufotu-b200 0.1 (reference HmmUFOtu v1.5.1)"; }

int hu_hmm_finalize(int32_t K, const double* T, double* entry, double* exit_) {
	if (K < 1 || !T || !entry || !exit_) { hu::setError("hu_hmm_finalize: bad arguments"); return HU_ERR_ARG; }
	enum { M = 0, I = 1, D = 2 };
#define TC(k, a, b) T[(size_t) (k) *9 + (a) *3 + (b)]
	// adjustProfileLocalMode (src/BandedHMMP7.cpp:721-733): uniform entry/exit probabilities
	std::vector<double> en(K + 1, exp(-TC(0, M, M))), ex(K + 1, exp(-TC(K, M, M)));
	en[0] = ex[0] = 0;
	// wingRetract (src/BandedHMMP7.cpp:1083-1120): fold the B->D..D->M and M->D..D->E paths in
	for (int j = 2; j <= K; ++j) {
		double cost = 0;
		cost += TC(0, M, D);
		for (int i = 1; i < j - 1; ++i) cost += TC(i, D, D);
		cost += TC(j - 1, D, M);
		en[j] += exp(-cost);
		if (en[j] > 1) en[j] = 1;
	}
	for (int i = 1; i <= K - 1; ++i) {
		double cost = 0;
		cost += TC(i, M, D);
		for (int j = i + 1; j < K; ++j) cost += TC(j, D, D);
		cost += TC(K, D, M);
		ex[i] += exp(-cost);
		if (ex[i] > 1) ex[i] = 1;
	}
#undef TC
	for (int k = 0; k <= K; ++k) {
		entry[k] = -log(en[k]);
		exit_[k] = -log(ex[k]);
	}
	return HU_OK;
}

int hu_model_Pr(const hu_model_desc* d, double v, double* P16) {
	hu::Model m;
	int rc = hu::buildModel(d, &m);
	if (rc != HU_OK) return rc;
	hu::modelPr(m, v, P16);
	return HU_OK;
}

static int createDb(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device, bool deferMessages);

int hu_db_create(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device) {
	return createDb(out, hmm, tree, model, device, false);
}
int hu_db_create_deferred(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device) {
	return createDb(out, hmm, tree, model, device, true);
}

// n messages in the log domain, each 4 x csLen doubles; nodes[i] / dirs[i] say whose (0: node -> parent, 1: parent -> node)
int hu_db_put_messages(hu_db* db, int64_t n, const int64_t* nodes, const int32_t* dirs, const double* logMsg) {
	if (!db || !nodes || !dirs || !logMsg || n < 0) { hu::setError("hu_db_put_messages: bad arguments"); return HU_ERR_ARG; }
	if (!db->deferred) { hu::setError("hu_db_put_messages: the database was not created with hu_db_create_deferred"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(db->device));
	const size_t cols = (size_t) db->tree.csLen;
	for (int64_t i = 0; i < n; ++i)
		if (nodes[i] < 0 || nodes[i] >= db->tree.nNodes || (dirs[i] != 0 && dirs[i] != 1)) { hu::setError("hu_db_put_messages: bad node / direction at %lld", (long long) i); return HU_ERR_ARG; }
	const int64_t chunk = std::max<int64_t>(1, (int64_t) ((256u << 20) / (cols * 32)));
	double* stage;
	HU_CUDA(cudaMalloc(&stage, (size_t) std::min(chunk, std::max<int64_t>(n, 1)) * cols * 4 * sizeof(double)));
	for (int64_t i0 = 0; i0 < n; i0 += chunk) {
		const int64_t c = std::min(chunk, n - i0);
		HU_CUDA(cudaMemcpy(stage, logMsg + (size_t) i0 * cols * 4, (size_t) c * cols * 4 * sizeof(double), cudaMemcpyHostToDevice));
		for (int64_t i = 0; i < c; ++i) {
			double* E = const_cast<double*>(dirs[i0 + i] ? db->tree.dn : db->tree.up) + (size_t) nodes[i0 + i] * cols * 4;
			double* S = const_cast<double*>(dirs[i0 + i] ? db->tree.dnS : db->tree.upS) + (size_t) nodes[i0 + i] * cols;
			hu::toLinearKernel<<<(unsigned) ((cols + 255) / 256), 256>>>(stage + (size_t) i * cols * 4, cols, E, S);
		}
		HU_CUDA(cudaGetLastError());
		HU_CUDA(cudaDeviceSynchronize());
	}
	cudaFree(stage);
	return HU_OK;
}
int hu_db_finish(hu_db* db) {
	if (!db) return HU_ERR_ARG;
	HU_CUDA(cudaSetDevice(db->device));
	int rc = hu::finishMessages(db);
	if (rc == HU_OK) db->deferred = 0;
	return rc;
}

// A second replica on another device, copied allocation by allocation over the peer path (NVLink where the devices are
// peers): the database is read from disk and converted once, then replicated at link speed.
int hu_db_clone(hu_db** out, hu_db* src, int device) {
	if (!out || !src) { hu::setError("hu_db_clone: NULL argument"); return HU_ERR_ARG; }
	if (src->deferred) { hu::setError("hu_db_clone: the source is not finished (hu_db_finish)"); return HU_ERR_STATE; }
	HU_CUDA(cudaSetDevice(device));
	{
		int rcA = hu::configViterbiKernels();
		if (rcA == HU_OK) rcA = hu::configSeedKernels();
		if (rcA == HU_OK) rcA = hu::configSeedMmKernels();
		if (rcA == HU_OK) rcA = hu::configEstimateKernels();
		if (rcA == HU_OK) rcA = hu::configPlaceKernels();
		if (rcA != HU_OK) return rcA;
	}
	hu_db* db = new hu_db(*src); // host-side fields; the device pointers are remapped below
	db->device = device;
	db->allocs.clear();
	for (size_t i = 0; i < src->allocs.size(); ++i) {
		void* q = NULL;
		cudaError_t e = cudaMalloc(&q, src->allocBytes[i]);
		if (e == cudaSuccess) e = cudaMemcpyPeer(q, device, src->allocs[i], src->device, src->allocBytes[i]);
		if (e != cudaSuccess) {
			hu::setError("hu_db_clone: allocation %zu of %zu bytes: %s", i, src->allocBytes[i], cudaGetErrorString(e));
			if (q) cudaFree(q);
			for (size_t k = 0; k < db->allocs.size(); ++k) cudaFree(db->allocs[k]);
			delete db;
			return e == cudaErrorMemoryAllocation ? HU_ERR_NOMEM : HU_ERR_CUDA;
		}
		db->allocs.push_back(q);
	}
	HU_CUDA(cudaDeviceSynchronize());
	// every device pointer of the replica points into one of the source's allocations: same offset in the copy
	struct Remap {
		const hu_db* s; const hu_db* d;
		const void* operator()(const void* p) const {
			if (!p) return p;
			for (size_t i = 0; i < s->allocs.size(); ++i) {
				const char* b = (const char*) s->allocs[i];
				if ((const char*) p >= b && (const char*) p < b + s->allocBytes[i]) return (const char*) d->allocs[i] + ((const char*) p - b);
			}
			return NULL;
		}
	} rm = {src, db};
#define HU_RM(field) db->field = (decltype(db->field)) rm(src->field)
	HU_RM(hmm.eM); HU_RM(hmm.eI); HU_RM(hmm.tr); HU_RM(hmm.ent); HU_RM(hmm.ext); HU_RM(hmm.p2cs);
	HU_RM(tree.parent); HU_RM(tree.blen); HU_RM(tree.height); HU_RM(tree.seq); HU_RM(tree.taxon);
	HU_RM(tree.up); HU_RM(tree.dn); HU_RM(tree.upS); HU_RM(tree.dnS); HU_RM(tree.upP32); HU_RM(tree.dnP32);
	HU_RM(tree.litTab); HU_RM(tree.modelDev); HU_RM(tree.planes); HU_RM(tree.seqT); HU_RM(tree.colK);
	HU_RM(tree.scanNode); HU_RM(tree.posOfSlot); HU_RM(tree.nodeOfPos); HU_RM(tree.inv2N);
	HU_RM(tree.mmA); HU_RM(tree.seqTP); HU_RM(tree.mmKeyTab);
	HU_RM(tree.mmColDefault); HU_RM(tree.mmExcPtr); HU_RM(tree.mmExcPos); HU_RM(tree.mmSlotOfPos);
#undef HU_RM
	*out = db;
	return HU_OK;
}

static int createDb(hu_db** out, const hu_hmm_params* hmm, const hu_tree_desc* tree, const hu_model_desc* model, int device, bool deferMessages) {
	if (!out || !hmm || !tree || !model) { hu::setError("hu_db_create: NULL argument"); return HU_ERR_ARG; }
	int nDev = 0;
	cudaError_t e = cudaGetDeviceCount(&nDev);
	if (e != cudaSuccess || nDev == 0) {
		hu::setError("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
		return HU_ERR_CUDA;
	}
	HU_CUDA(cudaSetDevice(device));
	// kernel attributes are per device: set them for this one (no process-wide "already done" flags)
	{
		int rcA = hu::configViterbiKernels();
		if (rcA == HU_OK) rcA = hu::configSeedKernels();
		if (rcA == HU_OK) rcA = hu::configSeedMmKernels();
		if (rcA == HU_OK) rcA = hu::configEstimateKernels();
		if (rcA == HU_OK) rcA = hu::configPlaceKernels();
		if (rcA != HU_OK) return rcA;
	}
	hu_db* db = new hu_db();
	db->device = device;
	int rc = hu::createImpl(db, hmm, tree, model, deferMessages);
	if (rc != HU_OK) {
		hu_db_destroy(db);
		return rc;
	}
	*out = db;
	return HU_OK;
}

int hu_db_destroy(hu_db* db) {
	if (!db) return HU_OK;
	cudaSetDevice(db->device);
	for (size_t i = 0; i < db->allocs.size(); ++i) cudaFree(db->allocs[i]);
	delete db;
	return HU_OK;
}

int hu_db_get_messages(hu_db* db, int64_t node0, int64_t n, double* up, double* dn) {
	if (!db || node0 < 0 || node0 + n > db->tree.nNodes) { hu::setError("hu_db_get_messages: bad range"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(db->device));
	// the device holds (exp(V - max V), max V); hand back V = log(E) + scale, one node at a time
	const size_t cols = (size_t) db->tree.csLen;
	double* stage;
	HU_CUDA(cudaMalloc(&stage, cols * 4 * sizeof(double)));
	for (int dir = 0; dir < 2; ++dir) {
		double* dst = dir ? dn : up;
		if (!dst) continue;
		const double* E = dir ? db->tree.dn : db->tree.up;
		const double* S = dir ? db->tree.dnS : db->tree.upS;
		for (int64_t x = 0; x < n; ++x) {
			hu::toLogKernel<<<(unsigned) ((cols + 255) / 256), 256>>>(E + (size_t) (node0 + x) * cols * 4, S + (size_t) (node0 + x) * cols, cols, stage);
			HU_CUDA(cudaMemcpy(dst + (size_t) x * cols * 4, stage, cols * 4 * sizeof(double), cudaMemcpyDeviceToHost));
		}
	}
	cudaFree(stage);
	return HU_OK;
}

// messages of a list of nodes over a column range, log domain, [i][column - col0][4]: what a sparse copy of the database
// (a test oracle, a partial .ptu writer) needs, without moving whole 4 x csLen rows
int hu_db_get_messages_of(hu_db* db, int64_t nIds, const int64_t* ids, int32_t col0, int32_t nCols, double* up, double* dn) {
	if (!db || !ids || nIds < 0 || col0 < 0 || nCols < 1 || col0 + nCols > db->tree.csLen) { hu::setError("hu_db_get_messages_of: bad range"); return HU_ERR_ARG; }
	for (int64_t i = 0; i < nIds; ++i)
		if (ids[i] < 0 || ids[i] >= db->tree.nNodes) { hu::setError("hu_db_get_messages_of: node %lld out of range", (long long) ids[i]); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(db->device));
	const size_t cols = (size_t) db->tree.csLen, row = (size_t) nCols * 4;
	const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nIds, (int64_t) ((64u << 20) / (row * sizeof(double)))));
	double* stage;
	HU_CUDA(cudaMalloc(&stage, (size_t) chunk * row * sizeof(double)));
	int rc = HU_OK;
	for (int dir = 0; dir < 2 && rc == HU_OK; ++dir) {
		double* dst = dir ? dn : up;
		if (!dst) continue;
		const double* E = dir ? db->tree.dn : db->tree.up;
		const double* S = dir ? db->tree.dnS : db->tree.upS;
		for (int64_t i0 = 0; i0 < nIds; i0 += chunk) {
			const int64_t c = std::min(chunk, nIds - i0);
			for (int64_t i = 0; i < c; ++i)
				hu::toLogKernel<<<(unsigned) ((nCols + 255) / 256), 256>>>(E + ((size_t) ids[i0 + i] * cols + col0) * 4, S + (size_t) ids[i0 + i] * cols + col0,
				                                                        (size_t) nCols, stage + (size_t) i * row);
			if (cudaMemcpy(dst + (size_t) i0 * row, stage, (size_t) c * row * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
				hu::setError("hu_db_get_messages_of: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
				rc = HU_ERR_CUDA;
				break;
			}
		}
	}
	cudaFree(stage);
	return rc;
}

int hu_db_get_seq(hu_db* db, int64_t node0, int64_t n, int8_t* seq) {
	if (!db || node0 < 0 || node0 + n > db->tree.nNodes || !seq) { hu::setError("hu_db_get_seq: bad range"); return HU_ERR_ARG; }
	HU_CUDA(cudaSetDevice(db->device));
	HU_CUDA(cudaMemcpy(seq, db->tree.seq + node0 * db->tree.csLen, (size_t) n * db->tree.csLen, cudaMemcpyDeviceToHost));
	return HU_OK;
}

int64_t hu_db_bytes(hu_db* db) { return db ? db->bytes : 0; }

} // extern "C"
