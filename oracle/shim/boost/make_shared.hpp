#include "shared_ptr.hpp"
