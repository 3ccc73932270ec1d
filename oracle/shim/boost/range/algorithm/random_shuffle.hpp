#include "../../random/mersenne_twister.hpp"
