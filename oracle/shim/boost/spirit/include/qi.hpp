// Stand-in so that NewickTree.h PARSES: the qi grammar in that header is a class template that this build never
// instantiates (NewickTree.cpp, the only user, is not compiled; oracle/_ref gets its trees from .ptu files).
#ifndef HU_SHIM_BOOST_SPIRIT_QI
#define HU_SHIM_BOOST_SPIRIT_QI
#include <cctype>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>
namespace boost {
namespace shim_any {
struct any {
	any() {}
	template <class T> any(const T&) {}
	template <class T> any operator[](const T&) const { return any(); }
	template <class T> any operator()(const T&) const { return any(); }
	template <class T, class U> any operator()(const T&, const U&) const { return any(); }
	template <class T> any& operator=(const T&) { return *this; }
	template <class T> any& operator%=(const T&) { return *this; }
	any operator+() const { return any(); }
	any operator-() const { return any(); }
	any operator*() const { return any(); }
	any operator!() const { return any(); }
	void name(const std::string&) {}
};
#define HU_ANY_BINOP(OP) \
	template <class T> any operator OP(const any&, const T&) { return any(); } \
	inline any operator OP(char, const any&) { return any(); } \
	inline any operator OP(const char*, const any&) { return any(); } \
	inline any operator OP(std::ostream&, const any&) { return any(); }
HU_ANY_BINOP(>>)
HU_ANY_BINOP(>)
HU_ANY_BINOP(|)
HU_ANY_BINOP(-)
HU_ANY_BINOP(%)
HU_ANY_BINOP(<<)
inline any operator<<(const any&, std::ostream& (*)(std::ostream&)) { return any(); }
}
namespace spirit {
namespace qi {
using shim_any::any;
template <class It, class Sig = void, class S = void> struct rule : any { template <class T> rule& operator=(const T&) { return *this; } template <class T> rule& operator%=(const T&) { return *this; } };
template <class It, class Sig = void, class S = void> struct grammar {
	typedef grammar base_type;
	template <class R> grammar(R&, const std::string& = "") {}
};
static const any lexeme, double_, int_, _val, _1, _2, _3, _4, eps;
enum error_handler_result { fail, retry, accept, rethrow };
template <error_handler_result A, class R, class F> void on_error(R&, const F&) {}
static const any space;
// The one grammar the reference parses with Spirit is its Newick grammar (NewickTree.h:172-236):
//   tree = subtree ';'      subtree = -('(' subtree % ',' ')') >> -label >> -(':' double)
//   label = lexeme[+(print - space - "()[]':;,")] | '\'' lexeme[+(print - '\'')] '\''      skipper: space
// phrase_parse() below is a recursive-descent restatement of exactly that grammar for any attribute type with
// members name / length / children, so that NewickTree.cpp compiles and works unmodified.
namespace newick_detail {
template <class It> void skip(It& it, It end) { while (it != end && std::isspace((unsigned char) *it)) ++it; }
inline bool unquotedChar(char c) {
	return std::isprint((unsigned char) c) && !std::isspace((unsigned char) c) && c != '(' && c != ')' && c != '[' && c != ']' && c != '\'' && c != ':' && c != ';' && c != ',';
}
template <class It, class Attr> bool subtree(It& it, It end, Attr& node) {
	skip(it, end);
	if (it != end && *it == '(') { // optional descendant list; on failure the optional matches nothing
		It save = it;
		++it;
		std::vector<Attr> kids;
		bool ok = true;
		for (;;) {
			Attr kid = Attr();
			if (!subtree(it, end, kid)) { ok = false; break; }
			kids.push_back(kid);
			skip(it, end);
			if (it != end && *it == ',') { ++it; continue; }
			break;
		}
		skip(it, end);
		if (ok && it != end && *it == ')') { ++it; node.children = kids; }
		else it = save;
	}
	skip(it, end);
	if (it != end && unquotedChar(*it)) {
		std::string s;
		while (it != end && unquotedChar(*it)) s += *it++;
		node.name = s;
	} else if (it != end && *it == '\'') {
		It q = it;
		++q;
		std::string s;
		while (q != end && std::isprint((unsigned char) *q) && *q != '\'') s += *q++;
		if (s.empty() || q == end || *q != '\'') return false; // '>' is an expectation point
		it = ++q;
		node.name = s;
	}
	skip(it, end);
	if (it != end && *it == ':') {
		It q = it;
		++q;
		skip(q, end);
		std::string num;
		It r = q;
		while (r != end && (std::isdigit((unsigned char) *r) || *r == '+' || *r == '-' || *r == '.' || *r == 'e' || *r == 'E')) num += *r++;
		char* stop = 0;
		const double v = std::strtod(num.c_str(), &stop);
		const size_t used = (size_t) (stop - num.c_str());
		if (used == 0) return false; // expectation failure
		std::advance(q, (long) used);
		it = q;
		node.length = v;
	}
	return true;
}
}
template <class It, class G, class Sk, class Attr> bool phrase_parse(It& it, It end, const G&, const Sk&, Attr& tree) {
	It cur = it;
	if (!newick_detail::subtree(cur, end, tree)) return false;
	newick_detail::skip(cur, end);
	if (cur == end || *cur != ';') return false;
	++cur;
	newick_detail::skip(cur, end); // post-skip
	it = cur;
	return true;
}
}
namespace ascii { using shim_any::any; static const any char_, string, print, space, alnum, alpha, digit; }
}
namespace phoenix {
using shim_any::any;
template <int N, class T> any at_c(const T&) { return any(); }
template <class T, class U> any push_back(const T&, const U&) { return any(); }
template <class T, class A, class B> any construct(const A&, const B&) { return any(); }
template <class T> any val(const T&) { return any(); }
}
namespace fusion {}
}
#endif
