// TEST INFRASTRUCTURE - NOT PRODUCT CODE. Basix stand-in (see finite-element.h).
#pragma once
#include "finite-element.h"
