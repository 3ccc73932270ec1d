#ifndef HU_SHIM_BOOST_MATH_GAMMA_DIST
#define HU_SHIM_BOOST_MATH_GAMMA_DIST
#include <algorithm>
#include <limits>
#include "../special_functions/gamma.hpp"
namespace boost { namespace math {
template <class T = double> class gamma_distribution {
public:
	gamma_distribution(T shape, T scale = 1) : k_(shape), th_(scale) {}
	T shape() const { return k_; }
	T scale() const { return th_; }
private:
	T k_, th_;
};
template <class T> inline T cdf(const gamma_distribution<T>& g, T x) { return (T) gamma_p((double) g.shape(), (double) (x / g.scale())); }
template <class T> inline T pdf(const gamma_distribution<T>& g, T x) {
	const double k = g.shape(), th = g.scale();
	return x < 0 ? 0 : (T) std::exp((k - 1) * std::log(x / th) - x / th - ::lgamma(k)) / th;
}
// quantile by bisection + Newton polishing on the regularised incomplete gamma function
template <class T> inline T quantile(const gamma_distribution<T>& g, T p) {
	if (p <= 0) return 0;
	if (p >= 1) return std::numeric_limits<T>::infinity();
	double lo = 0, hi = std::max(1.0, (double) g.shape());
	while (gamma_p(g.shape(), hi) < p) hi *= 2;
	for (int i = 0; i < 200 && hi - lo > 1e-16 * hi; ++i) {
		const double mid = 0.5 * (lo + hi);
		if (gamma_p(g.shape(), mid) < p) lo = mid; else hi = mid;
	}
	return (T) (0.5 * (lo + hi) * g.scale());
}
} }
#endif
