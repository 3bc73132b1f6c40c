// Stand-in: everything lives in <boost/geometry.hpp> of this shim.
#pragma once
#include <boost/geometry.hpp>
