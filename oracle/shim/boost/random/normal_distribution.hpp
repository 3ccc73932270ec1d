#include "mersenne_twister.hpp"
