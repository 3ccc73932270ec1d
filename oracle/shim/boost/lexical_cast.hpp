#ifndef HU_SHIM_BOOST_LEXICAL_CAST
#define HU_SHIM_BOOST_LEXICAL_CAST
#include <sstream>
#include <string>
#include <typeinfo>
#include <limits>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public:
	const char* what() const noexcept override { return "bad lexical cast: source type value could not be interpreted as target"; }
};
template <class T> struct lc_to {
	template <class S> static T run(const S& s) {
		std::stringstream ss;
		ss.precision(std::numeric_limits<double>::digits10 + 2);
		T out;
		if (!(ss << s) || !(ss >> out) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
		return out;
	}
};
template <> struct lc_to<std::string> {
	template <class S> static std::string run(const S& s) {
		std::ostringstream ss;
		// boost prints floating point with enough digits to round-trip (17 for double)
		ss.precision(std::numeric_limits<double>::digits10 + 2);
		if (!(ss << s)) throw bad_lexical_cast();
		return ss.str();
	}
};
template <class T, class S> T lexical_cast(const S& s) { return lc_to<T>::run(s); }
}
#endif
