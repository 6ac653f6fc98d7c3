#include <opencv2/core.hpp>   // oracle/_ref build shim, see opencv2/core.hpp
