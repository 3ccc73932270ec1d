// boost::random engines/distributions on top of <random>.  mt11213b has Boost's parameters, so the raw stream is the
// same; the DISTRIBUTIONS are libstdc++'s and do not reproduce Boost's variates (only the simulator and the OTU
// tools draw random numbers; nothing on the per-read path does).
#ifndef HU_SHIM_BOOST_RANDOM
#define HU_SHIM_BOOST_RANDOM
#include <random>
#include <algorithm>
#include <cstdint>
#include <iterator>
#include <vector>
namespace boost { namespace random {
typedef std::mersenne_twister_engine<uint32_t, 32, 351, 175, 19, 0xccab8ee7, 11, 0xffffffff, 7, 0x31b6ab00, 15, 0xffe50000, 17, 1812433253> mt11213b;
typedef std::mt19937 mt19937;
template <class T = double> class uniform_01 {
public:
	typedef T result_type;
	template <class E> T operator()(E& e) { return std::generate_canonical<T, 53>(e); }
};
template <class T = int> class uniform_smallint {
public:
	typedef T result_type;
	uniform_smallint(T lo = 0, T hi = 9) : d_(lo, hi) {}
	template <class E> T operator()(E& e) { return d_(e); }
private:
	std::uniform_int_distribution<T> d_;
};
template <class T = int> class uniform_int_distribution : public std::uniform_int_distribution<T> {
public:
	uniform_int_distribution(T lo = 0, T hi = 9) : std::uniform_int_distribution<T>(lo, hi) {}
};
template <class T = double> class normal_distribution : public std::normal_distribution<T> {
public:
	normal_distribution(T m = 0, T s = 1) : std::normal_distribution<T>(m, s) {}
};
template <class I = int, class W = double> class discrete_distribution {
public:
	typedef I result_type;
	class param_type {
	public:
		param_type() {}
		template <class It> param_type(It b, It e) : w(b, e) {}
		template <class Rng> explicit param_type(const Rng& r) : w(std::begin(r), std::end(r)) {}
		std::vector<W> w;
	};
	discrete_distribution() {}
	template <class It> discrete_distribution(It b, It e) : d_(b, e) {}
	template <class Rng> explicit discrete_distribution(const Rng& r) : d_(std::begin(r), std::end(r)) {}
	explicit discrete_distribution(const param_type& p) : d_(p.w.begin(), p.w.end()) {}
	template <class E> I operator()(E& e) { return (I) d_(e); }
	param_type param() const { const std::vector<double> pr = d_.probabilities(); return param_type(pr.begin(), pr.end()); }
	void param(const param_type& p) { d_ = std::discrete_distribution<long>(p.w.begin(), p.w.end()); }
private:
	std::discrete_distribution<long> d_;
};
template <class E, class I = long> class random_number_generator {
public:
	typedef I argument_type;
	typedef I result_type;
	explicit random_number_generator(E& e) : e_(e) {}
	I operator()(I n) { return std::uniform_int_distribution<I>(0, n - 1)(e_); }
private:
	E& e_;
};
}
using random::mt11213b;
using random::mt19937;
using random::uniform_01;
using random::uniform_smallint;
using random::normal_distribution;
using random::random_number_generator;
namespace range {
template <class C> C& random_shuffle(C& c) { std::mt19937 g(5489u); std::shuffle(c.begin(), c.end(), g); return c; }
template <class C, class G> C& random_shuffle(C& c, G& gen) { for (long i = (long) c.size() - 1; i > 0; --i) { typename C::value_type t = c[i]; const long j = gen(i + 1); c[i] = c[j]; c[j] = t; } return c; }
}
using range::random_shuffle;
}
#endif
