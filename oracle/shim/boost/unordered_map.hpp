// boost::unordered_map -> std::unordered_map.  NOTE: iteration order differs from Boost's; nothing on the per-read
// path of the reference depends on it (maps are only looked up there).
#ifndef HU_SHIM_BOOST_UNORDERED_MAP
#define HU_SHIM_BOOST_UNORDERED_MAP
#include <unordered_map>
#include <memory>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<std::pair<const K, V> > >
using unordered_map = std::unordered_map<K, V, H, E, A>;
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<std::pair<const K, V> > >
using unordered_multimap = std::unordered_multimap<K, V, H, E, A>;
}
#endif
