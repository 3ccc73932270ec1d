// boost::shared_ptr -> std::shared_ptr (test infrastructure: lets the reference's sources compile without Boost)
#ifndef HU_SHIM_BOOST_SHARED_PTR
#define HU_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
}
#endif
