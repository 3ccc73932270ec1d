#include "../filtering_stream.hpp"
