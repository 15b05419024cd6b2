#pragma once
namespace boost { namespace mpl { template <int N> struct int_ { static const int value = N; }; } }
