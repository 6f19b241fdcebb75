// bitpit_common.hpp -- part of the minimal bitpit stand-in (see README.md in this directory).
// Written from scratch against the API surface minimmerflow uses; NOT bitpit code.
#ifndef MMF_COMPAT_BITPIT_COMMON_HPP
#define MMF_COMPAT_BITPIT_COMMON_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define BITPIT_UNUSED(variable) (void) (variable)
#define BITPIT_UNREACHABLE(message)                                                                \
    do {                                                                                           \
        assert(false && (message));                                                                \
        __builtin_unreachable();                                                                   \
    } while (0)

// array arithmetic / printing used by the solver sources (bitpit keeps these in the global namespace)
template <typename T, std::size_t d>
std::array<T, d> operator*(const T &a, const std::array<T, d> &x)
{
    std::array<T, d> y;
    for (std::size_t i = 0; i < d; ++i) y[i] = a * x[i];
    return y;
}

template <typename T, std::size_t d>
std::ostream &operator<<(std::ostream &out, const std::array<T, d> &x)
{
    for (std::size_t i = 0; i < d; ++i) out << (i ? " " : "") << x[i];
    return out;
}

#endif
