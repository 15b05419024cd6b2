// FragmentBuilder::consolidateDuplicateFragments (FragmentBuilder.cpp:279-324) for host AND device code: the reference's loop
// with the library's std::sort replaced by its replay (sort_replay.cuh), so
// that the entry that survives a group of duplicates -- and with it firstSeedIndex and the seed bookkeeping -- is the one the
// reference keeps.  tests/cpp/test_consolidate_replay.cu checks on the CPU that both give the same lists, byte for
// byte.  Used by the B1-B4 kernels of kernels_tile.cuh.
#pragma once
#include "sort_replay.cuh"
#include "../../include/isaac_ext.h"

namespace isaac_b200
{

/// FragmentMetadata::operator< (FragmentMetadata.hh:419-429) on records that carry an isaac_ext_fragment_t as member f
template <class Record> ISAAC_HD inline bool fragmentLessReplay(const Record &a, const Record &b)
{
    return a.f.contigId < b.f.contigId ||
           (a.f.contigId == b.f.contigId &&
            (a.f.position < b.f.position ||
             (a.f.position == b.f.position &&
              (a.f.reverse < b.f.reverse || (a.f.reverse == b.f.reverse && a.f.observedLength < b.f.observedLength)))));
}

/// \return the new size of list[0..n)
template <class Record> ISAAC_HD inline unsigned consolidateDuplicateFragmentsReplay(Record *list, unsigned n, const bool removeUnaligned)
{
    sort_replay::sort(list, n, [](const Record &a, const Record &b) { return fragmentLessReplay(a, b); });
    unsigned first = 0;
    while (first != n && removeUnaligned && !list[first].f.cigarLength) ++first;             // :288-296
    if (first) { for (unsigned k = first; k < n; ++k) list[k - first] = list[k]; n -= first; }
    if (n < 2) return n;
    unsigned last = 0;
    for (unsigned cur = 1; cur != n; ++cur)                                                  // :298-322
    {
        if (removeUnaligned && !list[cur].f.cigarLength) continue;
        isaac_ext_fragment_t &l = list[last].f;
        const isaac_ext_fragment_t &c = list[cur].f;
        if (l.position == c.position && l.contigId == c.contigId && l.reverse == c.reverse && l.observedLength == c.observedLength)
        {
            l.uniqueSeedCount = uint16_t(l.uniqueSeedCount + c.uniqueSeedCount);             // FragmentMetadata::consolidate (:470-475)
            if (c.nonUniqueSeedOffsetFirst < l.nonUniqueSeedOffsetFirst) l.nonUniqueSeedOffsetFirst = c.nonUniqueSeedOffsetFirst;
            if (c.nonUniqueSeedOffsetSecond > l.nonUniqueSeedOffsetSecond) l.nonUniqueSeedOffsetSecond = c.nonUniqueSeedOffsetSecond;
        }
        else
        {
            ++last;
            if (last != cur) list[last] = list[cur];
        }
    }
    return last + 1;
}

} // namespace isaac_b200
