// Forwarding header: the reference's Cython sources name "include/twopt.hpp"
// (T/_twopt.pyx:27); put triumvirate_b200/include/trv_compat on the include path.
#include "../../trv/twopt.hpp"
