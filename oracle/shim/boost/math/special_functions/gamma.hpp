// regularised incomplete gamma functions, lgamma, tgamma (series / continued fraction, Numerical-Recipes style
// restatement of the published algorithms; the reference calls boost::math for DB *build* only)
#ifndef HU_SHIM_BOOST_MATH_GAMMA
#define HU_SHIM_BOOST_MATH_GAMMA
#include <cmath>
#include <limits>
namespace boost { namespace math {
template <class T> inline T lgamma(T x) { return (T) ::lgamma((double) x); }
template <class T> inline T tgamma(T x) { return (T) ::tgamma((double) x); }
inline double gamma_p(double a, double x) {
	if (x <= 0) return 0;
	const double eps = std::numeric_limits<double>::epsilon();
	const double gln = ::lgamma(a);
	if (x < a + 1) { // series
		double ap = a, del = 1.0 / a, sum = del;
		for (int n = 0; n < 100000; ++n) {
			ap += 1; del *= x / ap; sum += del;
			if (std::fabs(del) < std::fabs(sum) * eps) break;
		}
		return sum * std::exp(-x + a * std::log(x) - gln);
	}
	// continued fraction (modified Lentz) for Q
	const double fpmin = std::numeric_limits<double>::min() / eps;
	double b = x + 1 - a, c = 1 / fpmin, d = 1 / b, h = d;
	for (int i = 1; i < 100000; ++i) {
		const double an = -i * (i - a);
		b += 2;
		d = an * d + b; if (std::fabs(d) < fpmin) d = fpmin;
		c = b + an / c; if (std::fabs(c) < fpmin) c = fpmin;
		d = 1 / d;
		const double del = d * c;
		h *= del;
		if (std::fabs(del - 1) <= eps) break;
	}
	return 1 - std::exp(-x + a * std::log(x) - gln) * h;
}
inline double gamma_q(double a, double x) { return 1 - gamma_p(a, x); }
} }
#endif
