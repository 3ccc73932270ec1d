#include "qi.hpp"
