// Stand-in for <boost/array.hpp>: TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Boost is not installed in this image; this header provides just the surface
// /root/reference/data/pillars.cpp:286-288 uses (operator[] on a fixed-size array key).
#pragma once
#include <array>
#include <cstddef>
namespace boost {
template <class T, std::size_t N>
struct array : public std::array<T, N> {};
}  // namespace boost
