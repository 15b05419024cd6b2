#pragma once
namespace boost { template <class T, class S> inline T numeric_cast(S s) { return static_cast<T>(s); } }
