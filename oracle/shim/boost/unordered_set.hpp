#ifndef HU_SHIM_BOOST_UNORDERED_SET
#define HU_SHIM_BOOST_UNORDERED_SET
#include <unordered_set>
#include <memory>
namespace boost {
template <class K, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<K> >
using unordered_set = std::unordered_set<K, H, E, A>;
}
#endif
