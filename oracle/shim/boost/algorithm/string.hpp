#ifndef HU_SHIM_BOOST_ALGORITHM_STRING
#define HU_SHIM_BOOST_ALGORITHM_STRING
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>
namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of_pred {
	std::string set;
	bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_pred is_any_of(const std::string& s) { is_any_of_pred p; p.set = s; return p; }
inline is_any_of_pred is_any_of(const char* s) { is_any_of_pred p; p.set = s; return p; }
inline is_any_of_pred is_any_of(char c) { is_any_of_pred p; p.set = std::string(1, c); return p; }
// boost::split semantics: every separator character ends a token (empty tokens are kept) unless token_compress_on,
// which merges runs of adjacent separators into one
template <class Seq, class Pred> Seq& split(Seq& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off) {
	out.clear();
	std::string cur;
	size_t i = 0;
	const size_t n = in.size();
	for (;;) {
		cur.clear();
		while (i < n && !pred(in[i])) cur += in[i++];
		out.push_back(typename Seq::value_type(cur));
		if (i >= n) break;
		++i; // the separator
		if (mode == token_compress_on) while (i < n && pred(in[i])) ++i;
	}
	return out;
}
template <class Seq> std::string join(const Seq& parts, const std::string& sep) {
	std::string out;
	bool first = true;
	for (typename Seq::const_iterator it = parts.begin(); it != parts.end(); ++it) {
		if (!first) out += sep;
		out += *it;
		first = false;
	}
	return out;
}
inline void trim(std::string& s) {
	size_t b = 0, e = s.size();
	while (b < e && std::isspace((unsigned char) s[b])) ++b;
	while (e > b && std::isspace((unsigned char) s[e - 1])) --e;
	s = s.substr(b, e - b);
}
inline std::string trim_copy(std::string s) { trim(s); return s; }
inline void to_upper(std::string& s) { for (size_t i = 0; i < s.size(); ++i) s[i] = (char) std::toupper((unsigned char) s[i]); }
inline void to_lower(std::string& s) { for (size_t i = 0; i < s.size(); ++i) s[i] = (char) std::tolower((unsigned char) s[i]); }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline bool starts_with(const std::string& s, const std::string& p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(const std::string& s, const std::string& p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline void replace_all(std::string& s, const std::string& from, const std::string& to) {
	if (from.empty()) return;
	size_t pos = 0;
	while ((pos = s.find(from, pos)) != std::string::npos) { s.replace(pos, from.size(), to); pos += to.size(); }
}
}
#endif
