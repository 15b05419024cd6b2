#pragma once
namespace boost { namespace mpl { template <class A, class B> struct equal_to { static const bool value = (A::value == B::value); }; } }
