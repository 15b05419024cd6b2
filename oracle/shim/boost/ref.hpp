#pragma once
// boost::reference_wrapper on top of the standard one (ref / cref live in bind.hpp)
#include <functional>
namespace boost {
template <class T> using reference_wrapper = std::reference_wrapper<T>;
}
