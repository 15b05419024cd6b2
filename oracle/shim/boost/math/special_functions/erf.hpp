#pragma once
#include <cmath>
namespace boost { namespace math { inline double erf(double x) { return std::erf(x); } } }
