#pragma once
namespace boost { namespace math { namespace constants { template <class T> inline T root_two() { return static_cast<T>(1.4142135623730950488016887242096980785696718753769L); } } } }
