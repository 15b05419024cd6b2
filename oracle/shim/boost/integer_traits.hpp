#pragma once
#include <limits>
namespace boost { template <class T> struct integer_traits { static const T const_max = std::numeric_limits<T>::max(); static const T const_min = std::numeric_limits<T>::min(); }; }
