// Forwarding header: everything lives in the single stand-in Cabana_Core.hpp.
#include <Cabana_Core.hpp>
