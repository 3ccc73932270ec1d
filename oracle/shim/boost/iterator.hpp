// intentionally empty: the reference includes <boost/iterator.hpp> but uses nothing from it
