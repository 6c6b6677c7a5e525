// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle_common.hpp).
// placeholder: EV oracle is added in oracle_ev.cpp (ev/Patch.cpp, ev/assembly.hpp,
// ev/solve_patch.hpp)
#include "oracle_common.hpp"
