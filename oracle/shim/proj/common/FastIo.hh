// Shadows common/FastIo.hh of the reference: only the helper Cigar::toString needs.
#ifndef iSAAC_COMMON_FAST_IO_HH
#define iSAAC_COMMON_FAST_IO_HH
#include <string>
#include <algorithm>
#include "common/MathCompatibility.hh"
namespace isaac { namespace common {
template <typename ContainerT> inline void appendUnsignedInteger(ContainerT &s, unsigned value)
{
    char buf[16]; int n = 0;
    do { buf[n++] = '0' + (value % 10); value /= 10; } while (value);
    while (n) s.push_back(buf[--n]);
}
} }
#endif
