// TEST INFRASTRUCTURE - NOT PRODUCT CODE. DOLFINx stand-in (see dolfinx/shim.hpp).
#pragma once
#include <dolfinx/shim.hpp>
