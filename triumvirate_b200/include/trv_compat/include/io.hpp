// Forwarding header: callers of the reference name "io.hpp" / "include/io.hpp".
#include "../../trv/io.hpp"
