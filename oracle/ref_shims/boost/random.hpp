// TEST INFRASTRUCTURE: stand-in for <boost/random.hpp>.  The reference's uq sources only need the standard headers the
// real one drags in (no boost random generator is on the oracle's path).
#pragma once
#include <math.h>
#include <cmath>
#include <cstring>
#include <iostream>
