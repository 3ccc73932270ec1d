/*
 * oracle_phylo.cpp -- CPU ORACLE (test infrastructure, never shipped): substitution models
 * and SEP (Seed-Estimate-Place) numerics, literal restatement of
 * /root/reference/src/{PhyloTreeUnrooted,DNASubModel,GTR,TN93,HKY85,F81,K80,JC69,
 * DiscreteGammaModel,SeqUtils,HmmUFOtu_main}.{h,cpp}.  PARITY UNPINNED, see oracle.h.
 */
#include "oracle.h"
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <istream>
#include <map>
#include <string>

namespace orc {

/* ----------------------------------------------------------- sub models ---- */
static void jacobiSym4(double A[16], double V[16]) {
	/* cyclic Jacobi for a symmetric 4x4; stands in for Eigen::EigenSolver (src/GTR.cpp:137-145):
	 * P(t) is invariant to the eigenvector basis up to rounding */
	for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 64; ++sweep) {
		double off = 0;
		for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += A[p * 4 + q] * A[p * 4 + q];
		if (off < 1e-300) break;
		for (int p = 0; p < 4; ++p)
			for (int q = p + 1; q < 4; ++q) {
				double apq = A[p * 4 + q];
				if (apq == 0) continue;
				double theta = (A[q * 4 + q] - A[p * 4 + p]) / (2 * apq);
				double t = (theta >= 0 ? 1.0 : -1.0) / (::fabs(theta) + ::sqrt(theta * theta + 1));
				double c = 1 / ::sqrt(t * t + 1), s = t * c;
				for (int k = 0; k < 4; ++k) {
					double akp = A[k * 4 + p], akq = A[k * 4 + q];
					A[k * 4 + p] = c * akp - s * akq;
					A[k * 4 + q] = s * akp + c * akq;
				}
				for (int k = 0; k < 4; ++k) {
					double apk = A[p * 4 + k], aqk = A[q * 4 + k];
					A[p * 4 + k] = c * apk - s * aqk;
					A[q * 4 + k] = s * apk + c * aqk;
				}
				for (int k = 0; k < 4; ++k) {
					double vkp = V[k * 4 + p], vkq = V[k * 4 + q];
					V[k * 4 + p] = c * vkp - s * vkq;
					V[k * 4 + q] = s * vkp + c * vkq;
				}
			}
	}
}

void Model::setup() {
	if (type == K80) {
		beta = 1 / (2 * kappa); /* src/K80.h:97-100 */
		for (int i = 0; i < 4; ++i) pi[i] = 0.25;
	}
	if (type == JC69)
		for (int i = 0; i < 4; ++i) pi[i] = 0.25;
	if (type != GTR)
		return;
	/* GTR::setQfromParams, src/GTR.cpp:124-135 */
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			Q[i * 4 + j] = R[i * 4 + j] * pi[j];
	for (int i = 0; i < 4; ++i) {
		Q[i * 4 + i] = 0;
		double rs = (Q[i * 4 + 0] + Q[i * 4 + 1]) + (Q[i * 4 + 2] + Q[i * 4 + 3]);
		Q[i * 4 + i] = -rs;
	}
	/* scale(Q) with pi = Ones, mu = 1 : Q / -trace(Q) * 1  (src/DNASubModel.cpp:123-126) */
	double tr = (1.0 * Q[0] + 1.0 * Q[10]) + (1.0 * Q[5] + 1.0 * Q[15]);
	for (int i = 0; i < 16; ++i) Q[i] = Q[i] / -tr * 1.0;
	/* eigen-decomposition through the similar symmetric matrix S = D^1/2 Q D^-1/2, D = diag(pi) */
	double S[16], W[16], sq[4];
	for (int i = 0; i < 4; ++i) sq[i] = ::sqrt(pi[i]);
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) S[i * 4 + j] = sq[i] * Q[i * 4 + j] / sq[j];
	for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) { double m = 0.5 * (S[i * 4 + j] + S[j * 4 + i]); S[i * 4 + j] = S[j * 4 + i] = m; }
	jacobiSym4(S, W);
	for (int k = 0; k < 4; ++k) lambda[k] = S[k * 4 + k];
	for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) {
		U[i * 4 + k] = W[i * 4 + k] / sq[i];
		U_1[k * 4 + i] = W[i * 4 + k] * sq[i];
	}
}

void Model::Pr(double v, double P[16]) const {
	enum { A = 0, C = 1, G = 2, Tt = 3 };
	switch (type) {
	case GTR: { /* src/GTR.h:116-121 : U * diag(exp(lambda v)) * U^-1, Pr(0) = I */
		if (v == 0) { for (int i = 0; i < 16; ++i) P[i] = (i % 5 == 0) ? 1.0 : 0.0; return; }
		double e[4];
		for (int k = 0; k < 4; ++k) e[k] = ::exp(lambda[k] * v);
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j) {
				double s = (U[i * 4 + 0] * e[0]) * U_1[0 * 4 + j];
				for (int k = 1; k < 4; ++k) s = s + (U[i * 4 + k] * e[k]) * U_1[k * 4 + j];
				P[i * 4 + j] = s;
			}
		return;
	}
	case TN93: case HKY85: { /* src/TN93.h:113-154, src/HKY85.h:111-153 */
		double kR = type == TN93 ? kr : kappa, kY = type == TN93 ? ky : kappa;
		double a = pi[A], c = pi[C], g = pi[G], t = pi[Tt];
		double e = ::exp(-beta * v);
		double eR = ::exp(-(1 + (a + g) * (kR - 1)) * beta * v);
		double eY = ::exp(-(1 + (c + t) * (kY - 1)) * beta * v);
		P[A * 4 + A] = (a * (a + g + (c + t) * e) + g * eR) / (a + g);
		P[A * 4 + C] = c * (1 - e);
		P[A * 4 + G] = (g * (a + g + (c + t) * e) - g * eR) / (a + g);
		P[A * 4 + Tt] = t * (1 - e);
		P[C * 4 + A] = a * (1 - e);
		P[C * 4 + C] = (c * (c + t + (a + g) * e) + t * eY) / (c + t);
		P[C * 4 + G] = g * (1 - e);
		P[C * 4 + Tt] = (t * (c + t + (a + g) * e) - t * eY) / (c + t);
		P[G * 4 + A] = (a * (a + g + (c + t) * e) - a * eR) / (a + g);
		P[G * 4 + C] = c * (1 - e);
		P[G * 4 + G] = (g * (a + g + (c + t) * e) + a * eR) / (a + g);
		P[G * 4 + Tt] = t * (1 - e);
		P[Tt * 4 + A] = a * (1 - e);
		P[Tt * 4 + C] = (c * (c + t + (a + g) * e) - c * eY) / (c + t);
		P[Tt * 4 + G] = g * (1 - e);
		P[Tt * 4 + Tt] = (t * (c + t + (a + g) * e) + c * eY) / (c + t);
		if (P[A * 4 + G] < 0) P[A * 4 + G] = 0;
		if (P[C * 4 + Tt] < 0) P[C * 4 + Tt] = 0;
		if (P[G * 4 + A] < 0) P[G * 4 + A] = 0;
		if (P[Tt * 4 + C] < 0) P[Tt * 4 + C] = 0;
		return;
	}
	case F81: { /* src/F81.h:110-118 */
		double e = ::exp(-beta * v);
		for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
			P[i * 4 + j] = i == j ? e + pi[j] * (1 - e) : pi[j] * (1 - e);
		return;
	}
	case K80: { /* src/K80.h:109-118 */
		double e = ::exp(-4 * beta * v);
		double eV = ::exp(-2 * (1 + kappa) * beta * v);
		double dg = (1.0 + e + 2 * eV) / 4, ti = (1.0 + e - 2 * eV) / 4, tv = (1.0 - e) / 4;
		for (int i = 0; i < 16; ++i) P[i] = tv;
		for (int i = 0; i < 4; ++i) P[i * 4 + i] = dg;
		P[A * 4 + G] = P[G * 4 + A] = P[C * 4 + Tt] = P[Tt * 4 + C] = ti;
		return;
	}
	case JC69: { /* src/JC69.h:97-101 */
		double off = (1 - ::exp(-4 * v / 3)) / 4, dg = (1 + 3 * ::exp(-4 * v / 3)) / 4;
		for (int i = 0; i < 16; ++i) P[i] = off;
		for (int i = 0; i < 4; ++i) P[i * 4 + i] = dg;
		return;
	}
	}
}

int Model::loadText(std::istream& in) { /* <type>\n then the model's read(): src/PhyloTreeUnrooted.cpp:675-685 */
	std::string typeName, tag, line, value;
	in >> typeName;
	in.ignore();
	static const char* names[] = {"GTR", "TN93", "HKY85", "F81", "K80", "JC69"};
	type = -1;
	for (int i = 0; i < 6; ++i) if (typeName == names[i]) type = i;
	if (type < 0) return -1;
	while (in >> tag) {
		if (tag[0] == '#') { std::getline(in, line); continue; }
		if (tag == "Type:") {
			in >> value;
			if (value != typeName) return -2;
			if (type == JC69) { std::getline(in, line); break; }
		} else if (tag == "pi:") {
			for (int i = 0; i < 4; ++i) in >> pi[i];
		} else if (tag == "R:" && type == GTR) {
			for (int i = 0; i < 16; ++i) in >> R[i];
		} else if (tag == "Q:" && type == GTR) {
			for (int i = 0; i <= 4; ++i) std::getline(in, line); /* src/GTR.cpp:66-70 */
			break;
		} else if (tag == "kr:") in >> kr;
		else if (tag == "ky:") in >> ky;
		else if (tag == "kappa:") {
			in >> kappa;
			if (type == K80) { std::getline(in, line); break; }
		} else if (tag == "beta:") { in >> beta; std::getline(in, line); break; }
		else return -3;
	}
	setup();
	return 0;
}

/* -------------------------------------------------------------- primitives ---- */
static const double MIN_LOGLIK_EXP = DBL_MIN_EXP / 2; /* = -510, src/PhyloTreeUnrooted.cpp:68 */
static const double INVALID_LOGLIK = 1;
static const double BRANCH_EPS = 1e-5;
static const int MAX_ITER = 100;

static inline double scaleOf(double maxV) { return maxV != NEG_INF && maxV < MIN_LOGLIK_EXP ? MIN_LOGLIK_EXP - maxV : 0; }
static inline double max4(const double V[4]) { double m = V[0]; for (int k = 1; k < 4; ++k) if (V[k] > m) m = V[k]; return m; }
static inline int argmax4(const double V[4]) { int b = 0; double m = V[0]; for (int k = 1; k < 4; ++k) if (V[k] > m) { m = V[k]; b = k; } return b; }

void dpsMat(const double P[16], const double V[4], double Y[4]) { /* src/PhyloTreeUnrooted.h:1495-1503 */
	double scale = scaleOf(max4(V));
	double e[4];
	for (int k = 0; k < 4; ++k) e[k] = ::exp(V[k] + scale);
	for (int i = 0; i < 4; ++i)
		Y[i] = ::log((P[i * 4 + 0] * e[0] + P[i * 4 + 1] * e[1]) + (P[i * 4 + 2] * e[2] + P[i * 4 + 3] * e[3])) - scale;
}

double dpsVec(const double pi[4], const double V[4]) { /* src/PhyloTreeUnrooted.h:1505-1510 */
	double scale = scaleOf(max4(V));
	double e[4];
	for (int k = 0; k < 4; ++k) e[k] = ::exp(V[k] + scale);
	return ::log((pi[0] * e[0] + pi[2] * e[2]) + (pi[1] * e[1] + pi[3] * e[3])) - scale;
}

static void rmes(const double* X /* [k][4] */, int K, double Y[4]) { /* row_mean_exp_scaled, src/PhyloTreeUnrooted.h:1521-1529 */
	for (int i = 0; i < 4; ++i) {
		double mx = X[i];
		for (int k = 1; k < K; ++k) if (X[k * 4 + i] > mx) mx = X[k * 4 + i];
		double sc = scaleOf(mx);
		double s = ::exp(X[i] + sc);
		for (int k = 1; k < K; ++k) s = s + ::exp(X[k * 4 + i] + sc);
		Y[i] = ::log(s / K) - sc;
	}
}

static void leafLoglik(const Tree& t, int8_t b, double Y[4]) { /* src/PhyloTreeUnrooted.h:1431-1437 */
	if (b >= 0) {
		for (int k = 0; k < 4; ++k) Y[k] = NEG_INF;
		Y[b] = 0;
	} else
		for (int k = 0; k < 4; ++k) Y[k] = ::log(t.model.pi[k]);
}

double pDist(const int8_t* s1, const int8_t* s2, int start, int end, long* dOut, long* nOut) {
	long d = 0, N = 0;
	for (int i = start; i <= end; ++i) {
		int b1 = s1[i], b2 = s2[i];
		if (b1 >= 0 && b2 >= 0) { N++; if (b1 != b2) d++; }
	}
	if (dOut) *dOut = d;
	if (nOut) *nOut = N;
	return static_cast<double>(d) / N;
}

/* --------------------------------------------------------------------- seed ---- */
std::vector<PTLoc> getSeed(const Tree& t, const int8_t* seq, int start, int end, double maxDiff, double maxHeight, int seedOrder) {
	std::vector<PTLoc> locs; /* src/HmmUFOtu_main.cpp:127-152 */
	for (long i = 0; i < t.nNodes; ++i)
		if (t.parent[i] >= 0 && t.height[i] <= maxHeight) {
			PTLoc l = {start, end, i, pDist(t.s(i), seq, start, end)};
			locs.push_back(l);
		}
	assert(!locs.empty());
	if (seedOrder == 0)
		std::sort(locs.begin(), locs.end());
	else
		std::stable_sort(locs.begin(), locs.end()); /* ids are ascending in input -> (dist, id) order */
	double bestDist = locs[0].dist, worstDist = locs[locs.size() - 1].dist;
	if (worstDist < bestDist + maxDiff) {
		std::vector<PTLoc>::iterator good;
		for (good = locs.begin(); good != locs.end(); ++good)
			if (good->dist - bestDist > maxDiff) break;
		locs.erase(good, locs.end());
	}
	return locs;
}

/* ----------------------------------------------------------------- estimate ---- */
Placement estimateSeq(const Tree& t, const int8_t* seq, const PTLoc& loc) { /* src/PhyloTreeUnrooted.cpp:849-877 */
	long u = loc.id, v = t.parent[u];
	double cDist = loc.dist;
	double pD = pDist(t.s(v), seq, loc.start, loc.end);
	double ratio = cDist / (cDist + pD);
	if (std::isnan(ratio)) ratio = 0.5;
	double w0 = t.blen[u];
	const double* U = t.U(u);
	const double* V = t.V(u);
	double wur = w0 * ratio, wvr = w0 - wur;
	double Pu[16], Pv[16], Pn[16];
	t.model.Pr(wur, Pu);
	t.model.Pr(wvr, Pv);
	const int W = loc.end - loc.start + 1;
	std::vector<double> R((size_t) W * 4);
	double d = 0, Nw = 0;
	/* inferWeight(l)(argmax l) (src/PhyloTreeUnrooted.h:1590-1593): exp(l - max l) / sum, i.e. 1 / sum at the arg-max */
	struct Wt { static double top(const double l[4]) { double mx = l[0]; for (int k = 1; k < 4; ++k) if (l[k] > mx) mx = l[k]; double s = 0; for (int k = 0; k < 4; ++k) s += ::exp(l[k] - mx); return 1.0 / s; } };
	for (int j = loc.start; j <= loc.end; ++j) {
		double a[4], b[4], n[4];
		dpsMat(Pu, U + (size_t) j * 4, a);
		dpsMat(Pv, V + (size_t) j * 4, b);
		double* r = &R[(size_t) (j - loc.start) * 4];
		for (int k = 0; k < 4; ++k) r[k] = a[k] + b[k];
		leafLoglik(t, seq[j], n);
		if (!t.estWeighted) { if (argmax4(r) != argmax4(n)) d++; } /* estimateBranchLengthUnweighted, :1018-1032 */
		else { /* estimateBranchLengthWeighted, :1034-1052 */
			const double ww = Wt::top(r) * Wt::top(n);
			if (argmax4(r) != argmax4(n)) d += ww;
			Nw += ww;
		}
	}
	double wnr = t.estWeighted ? d / Nw : d / (loc.end - loc.start + 1);
	t.model.Pr(wnr, Pn);
	double loglik = 0;
	for (int j = loc.start; j <= loc.end; ++j) {
		double n[4], np[4], x[4];
		leafLoglik(t, seq[j], n);
		dpsMat(Pn, n, np);
		const double* r = &R[(size_t) (j - loc.start) * 4];
		for (int k = 0; k < 4; ++k) x[k] = r[k] + np[k];
		loglik += dpsVec(t.model.pi, x);
	}
	Placement pl;
	pl.start = loc.start; pl.end = loc.end;
	pl.c = u; pl.p = v; pl.a = ratio <= 0.5 ? u : v;
	pl.wuv = w0; pl.ratio = ratio; pl.wnr = wnr; pl.loglik = loglik;
	return pl;
}

static bool cmpLoglik(const Placement& a, const Placement& b) { return a.loglik < b.loglik; }
static bool cmpQPlace(const Placement& a, const Placement& b) { return a.qPlace < b.qPlace; }

void filterPlacements(std::vector<Placement>& places, double maxError) { /* src/HmmUFOtu_main.cpp:162-173 */
	assert(!places.empty() && maxError >= 0);
	std::sort(places.rbegin(), places.rend(), cmpLoglik);
	double best = places[0].loglik;
	std::vector<Placement>::iterator good;
	for (good = places.begin(); good != places.end(); ++good)
		if (best - good->loglik > maxError) break;
	places.erase(good, places.end());
}

void sortByQPlace(std::vector<Placement>& places) { std::sort(places.rbegin(), places.rend(), cmpQPlace); }
void sortByLoglik(std::vector<Placement>& places) { std::sort(places.rbegin(), places.rend(), cmpLoglik); } /* --ML, src/hmmufotu.cpp:725-727 */

/* -------------------------------------------------------------------- place ---- */
namespace {
struct Place4 { /* the private 4-node tree {v(0), u(1), r(2), n(3)} of src/PhyloTreeUnrooted.cpp:879-954 */
	const Tree& t;
	int start, end, W;
	const double *U, *V;            /* msg(u->r), msg(v->r), full-length arrays */
	std::vector<double> Nm;          /* msg(n->r), window only */
	std::vector<double> rn, ru;      /* msg(r->n), msg(r->u), window only */
	double lu, lv, ln;
	Place4(const Tree& t, int s, int e) : t(t), start(s), end(e), W(e - s + 1) {}

	/* loglik(r, j) with the two (or three) given children; src/PhyloTreeUnrooted.cpp:320-346 */
	void rebuild(std::vector<double>& out, bool useU, bool useV, bool useN) const {
		const int Kd = t.dg.K;
		std::vector<double> Pu((size_t) 16 * std::max(Kd, 1)), Pv(Pu.size()), Pn(Pu.size());
		for (int k = 0; k < std::max(Kd, 1); ++k) {
			double r = Kd ? t.dg.r[k] : 1.0;
			t.model.Pr(lu * r, &Pu[(size_t) k * 16]);
			t.model.Pr(lv * r, &Pv[(size_t) k * 16]);
			t.model.Pr(ln * r, &Pn[(size_t) k * 16]);
		}
		out.resize((size_t) W * 4);
		std::vector<double> mat((size_t) 4 * std::max(Kd, 1));
		for (int j = start; j <= end; ++j) {
			double* o = &out[(size_t) (j - start) * 4];
			for (int k = 0; k < std::max(Kd, 1); ++k) {
				double acc[4] = {0, 0, 0, 0}, c[4];
				if (useU) { dpsMat(&Pu[(size_t) k * 16], U + (size_t) j * 4, c); for (int i = 0; i < 4; ++i) acc[i] += c[i]; }
				if (useV) { dpsMat(&Pv[(size_t) k * 16], V + (size_t) j * 4, c); for (int i = 0; i < 4; ++i) acc[i] += c[i]; }
				if (useN) { dpsMat(&Pn[(size_t) k * 16], &Nm[(size_t) (j - start) * 4], c); for (int i = 0; i < 4; ++i) acc[i] += c[i]; }
				for (int i = 0; i < 4; ++i) mat[(size_t) k * 4 + i] = acc[i];
			}
			if (Kd) rmes(&mat[0], Kd, o);
			else for (int i = 0; i < 4; ++i) o[i] = mat[i];
		}
	}

	/* single-branch Felsenstein optimiser, src/PhyloTreeUnrooted.cpp:749-798.
	 * Uw: window-indexed message (r->x); Vfull/Vw: the opposite message */
	double optimize(double w0, const std::vector<double>& Uw, const double* Vfull, const std::vector<double>* Vw, double maxL) const {
		double q0 = ::exp(-w0), p0 = 1 - q0, p = p0, q = q0;
		const double* pi = t.model.pi;
		for (int iter = 0; iter < MAX_ITER && p >= 0 && p <= 1; ++iter) {
			p = 0;
			int N = 0;
			for (int j = start; j <= end; ++j) {
				const double* u = &Uw[(size_t) (j - start) * 4];
				const double* v = Vfull ? Vfull + (size_t) j * 4 : &(*Vw)[(size_t) (j - start) * 4];
				double uv[4];
				for (int k = 0; k < 4; ++k) uv[k] = u[k] + v[k];
				double logA = dpsVec(pi, uv);
				double logB = dpsVec(pi, u) + dpsVec(pi, v);
				if (std::isnan(logA) || std::isnan(logB)) continue;
				double scale = std::max(logA, logB);
				logA -= scale; logB -= scale;
				p += ::exp(logB) * p0 / (::exp(logA) * q0 + ::exp(logB) * p0);
				N++;
			}
			p /= N;
			q = 1 - p;
			if (::fabs(::log(q) - ::log(q0)) < BRANCH_EPS) break;
			p0 = p; q0 = q;
		}
		double w = -::log(q);
		if (w > maxL) w = maxL;
		return w;
	}
};
} // namespace

void placeSeq(const Tree& t, const int8_t* seq, Placement& place, double maxHeight) {
	const double ratio0 = place.ratio, wnr0 = place.wnr;
	const long c = place.c, p = place.p;
	const double w0 = t.blen[c]; /* subtree.getBranchLength(u, v), src/PhyloTreeUnrooted.cpp:934 */
	Place4 s(t, place.start, place.end);
	s.U = t.U(c);
	s.V = t.V(c);
	s.lu = w0 * ratio0;          /* :906 */
	s.lv = w0 * (1 - ratio0);    /* :907 */
	s.ln = wnr0;                 /* :912 */
	s.Nm.resize((size_t) s.W * 4);
	for (int j = s.start; j <= s.end; ++j) {
		double n[4];
		leafLoglik(t, seq[j], n);            /* evaluate(r): msg(n->r) = 0 + leaf loglik, :329-343 */
		for (int k = 0; k < 4; ++k) s.Nm[(size_t) (j - s.start) * 4 + k] = 0.0 + n[k];
	}
	/* joint optimisation, src/PhyloTreeUnrooted.cpp:800-847 */
	double wur0 = s.lu, wvr0 = s.lv, wnr0j = s.ln;
	double w0j = wur0 + wvr0;
	double wur = wur0, wvr = wvr0, wnr = wnr0j;
	int iter = 0;
	for (iter = 0; iter < MAX_ITER && 0 <= wur && wur <= w0j; ++iter) {
		s.rebuild(s.rn, true, true, false);                         /* setRoot(n); evaluate(n) */
		wnr = s.optimize(s.ln, s.rn, NULL, &s.Nm, 1);               /* optimizeBranchLength(r, n, .., 1) */
		s.ln = wnr;
		s.rebuild(s.ru, false, true, true);                         /* setRoot(u); evaluate(u) */
		wur = s.optimize(s.lu, s.ru, s.U, NULL, w0j);               /* optimizeBranchLength(r, u, .., w0) */
		s.lu = wur;
		wvr = w0j - wur;
		s.lv = wvr;                                                 /* msg(r->v) rebuild has no observable effect */
		if (::fabs(wur - wur0) < BRANCH_EPS && ::fabs(wnr - wnr0j) < BRANCH_EPS) { /* double overload */
			iter++;
			break;
		}
		wur0 = wur; wvr0 = wvr; wnr0j = wnr;
	}
	place.outerIters = iter;
	/* literal return value: initRootLoglik() then treeLoglik over a matrix of INVALID_LOGLIK */
	double lit = 0;
	const double ones[4] = {INVALID_LOGLIK, INVALID_LOGLIK, INVALID_LOGLIK, INVALID_LOGLIK};
	for (int j = s.start; j <= s.end; ++j) lit += dpsVec(t.model.pi, ones);
	place.loglik = lit;
	/* intended value: sum_j log(pi . exp(loglik(r, j))) */
	std::vector<double> root;
	s.rebuild(root, true, true, true);
	double tl = 0;
	for (int j = 0; j < s.W; ++j) tl += dpsVec(t.model.pi, &root[(size_t) j * 4]);
	place.loglikTrue = tl;
	/* src/PhyloTreeUnrooted.cpp:940-953 */
	place.wnr = s.ln;
	double wurF = s.lu;
	place.ratio = wurF / w0;
	place.height = t.height[c] + wurF;
	if (place.ratio <= 0.5 || t.height[p] > maxHeight) place.a = c;
	else place.a = p;
}

/* ----------------------------------------------------------------- Q values ---- */
static double addScaled(double a, double b) { double s = std::max(a, b); return ::log(::exp(a - s) + ::exp(b - s)) + s; } /* src/math/Stats.h:233-236 */
static double p2q(double p) { return -10 * ::log(p) / ::log(10.0); }                                                     /* :239-241 */

void calcQValues(const Tree& t, std::vector<Placement>& places) { /* src/HmmUFOtu_main.cpp:182-216, UNIFORM prior */
	if (places.empty()) return;
	const int MAX_Q = 250;
	std::vector<double> pp(places.size());
	std::map<std::string, double> ppTaxon;
	double ppTaxNorm = NEG_INF;
	for (size_t i = 0; i < places.size(); ++i) {
		double prior = -0.0; /* PTPlacement::logPriorPr, src/PhyloTreeUnrooted.cpp:1166-1177 */
		if (t.priorHeight) {
			const Placement& pl = places[i];
			const double annoDist = pl.a == pl.c ? pl.wuv * pl.ratio + pl.wnr : (1 - pl.ratio) * pl.wuv + pl.wnr;
			prior = -(annoDist - pl.wnr + pl.height);
		}
		double p = places[i].loglik + prior;
		pp[i] = p;
		const std::string& tax = t.anno[places[i].a];
		if (ppTaxon.find(tax) == ppTaxon.end()) ppTaxon[tax] = p;
		else ppTaxon[tax] = addScaled(ppTaxon[tax], p);
		ppTaxNorm = addScaled(ppTaxNorm, p);
	}
	double mx = pp[0];
	for (size_t i = 1; i < pp.size(); ++i) if (pp[i] > mx) mx = pp[i];
	std::vector<double> pr(pp.size());
	double sum = 0;
	for (size_t i = 0; i < pp.size(); ++i) { pr[i] = ::exp(pp[i] - mx); sum += pr[i]; }
	for (size_t i = 0; i < pp.size(); ++i) {
		pr[i] /= sum;
		double q = p2q(1 - pr[i]);
		places[i].qPlace = q > MAX_Q ? MAX_Q : q;
	}
	for (size_t i = 0; i < places.size(); ++i) {
		double q = p2q(1 - ::exp(ppTaxon[t.anno[places[i].a]] - ppTaxNorm));
		places[i].qTaxon = q > MAX_Q ? MAX_Q : q;
	}
}

/* ------------------------------------------------- all-roots tree evaluation ---- */
void Tree::evaluateAll(const std::vector<std::vector<long> >& children, int j0, int j1, bool alloc) {
	/* Mathematically what src/hmmufotu-build.cpp:455-459 obtains by re-rooting at every node:
	 * up[id] = loglik(id, j) rooted above id; dn[id] = loglik(parent, j) rooted at id.
	 * Summation order over neighbours: children in list order, then the parent. */
	const int Kd = dg.K, KK = std::max(Kd, 1);
	if (j1 < 0) j1 = csLen;
	if (alloc) {
		up.assign((size_t) nNodes * csLen * 4, 0.0);
		dn.assign((size_t) nNodes * csLen * 4, 0.0);
	}
	std::vector<long> order;
	order.reserve(nNodes);
	{
		std::vector<long> st(1, root);
		while (!st.empty()) { long x = st.back(); st.pop_back(); order.push_back(x); for (size_t k = 0; k < children[x].size(); ++k) st.push_back(children[x][k]); }
	}
	/* transition matrices of every edge (id, parent) for every rate category */
	std::vector<double> PP((size_t) nNodes * KK * 16);
	for (long x = 0; x < nNodes; ++x)
		if (x != root)
			for (int k = 0; k < KK; ++k)
				model.Pr(blen[x] * (Kd ? dg.r[k] : 1.0), &PP[((size_t) x * KK + k) * 16]);
	std::vector<double> mat((size_t) 4 * KK);
	/* post-order: up messages */
	for (long oi = nNodes - 1; oi >= 0; --oi) {
		long x = order[oi];
		if (x == root) continue;
		double* out = &up[(size_t) x * csLen * 4];
		bool leaf = children[x].empty();
		for (int j = j0; j < j1; ++j) {
			std::fill(mat.begin(), mat.end(), 0.0);
			for (size_t ci = 0; ci < children[x].size(); ++ci) {
				long ch = children[x][ci];
				for (int k = 0; k < KK; ++k) {
					double c[4];
					dpsMat(&PP[((size_t) ch * KK + k) * 16], &up[((size_t) ch * csLen + j) * 4], c);
					for (int i = 0; i < 4; ++i) mat[(size_t) k * 4 + i] += c[i];
				}
			}
			double* o = out + (size_t) j * 4;
			if (!leaf && Kd) rmes(&mat[0], Kd, o);
			else for (int i = 0; i < 4; ++i) o[i] = mat[i];
			if (leaf) { double n[4]; leafLoglik(*this, seq[(size_t) x * csLen + j], n); for (int i = 0; i < 4; ++i) o[i] += n[i]; }
		}
	}
	/* pre-order: dn[x] = siblings' up messages + dn[parent] */
	for (long oi = 0; oi < nNodes; ++oi) {
		long x = order[oi];
		if (x == root) continue;
		long pa = parent[x];
		double* out = &dn[(size_t) x * csLen * 4];
		bool paLeafLike = (children[pa].size() + (pa == root ? 0 : 1)) == 1; /* isLeaf(): one neighbour */
		for (int j = j0; j < j1; ++j) {
			std::fill(mat.begin(), mat.end(), 0.0);
			for (size_t ci = 0; ci < children[pa].size(); ++ci) {
				long sib = children[pa][ci];
				if (sib == x) continue;
				for (int k = 0; k < KK; ++k) {
					double c[4];
					dpsMat(&PP[((size_t) sib * KK + k) * 16], &up[((size_t) sib * csLen + j) * 4], c);
					for (int i = 0; i < 4; ++i) mat[(size_t) k * 4 + i] += c[i];
				}
			}
			if (pa != root)
				for (int k = 0; k < KK; ++k) {
					double c[4];
					dpsMat(&PP[((size_t) pa * KK + k) * 16], &dn[((size_t) pa * csLen + j) * 4], c);
					for (int i = 0; i < 4; ++i) mat[(size_t) k * 4 + i] += c[i];
				}
			double* o = out + (size_t) j * 4;
			if (!paLeafLike && Kd) rmes(&mat[0], Kd, o);
			else for (int i = 0; i < 4; ++i) o[i] = mat[i];
		}
	}
}

} // namespace orc
