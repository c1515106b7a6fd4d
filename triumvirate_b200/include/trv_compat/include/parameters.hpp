// Forwarding header: the reference's Cython sources name "include/parameters.hpp"
// (e.g. T/_threept.pyx:27); put triumvirate_b200/include/trv_compat on the include path.
#include "../../trv/parameters.hpp"
