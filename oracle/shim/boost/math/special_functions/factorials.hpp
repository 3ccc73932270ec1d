#ifndef HU_SHIM_BOOST_MATH_FACTORIALS
#define HU_SHIM_BOOST_MATH_FACTORIALS
#include <cmath>
namespace boost { namespace math {
template <class T> inline T factorial(unsigned n) { return (T) ::tgamma((double) n + 1.0); }
} }
#endif
