// The library's own std::sort + the consolidate loop of FragmentBuilder::consolidateDuplicateFragments (FragmentBuilder.cpp:279-324) on
// WorkFragment lists: what csrc/consolidate_device.cuh (the libstdc++ replay the kernels run) is compared with.  Round 1 ran this
// on host threads inside the product; it is TEST CODE now.
#pragma once
#include <algorithm>
#include "../../isaac_aligner_b200/csrc/host_pipeline.cuh"

namespace isaac_b200
{

/// FragmentMetadata::operator< (FragmentMetadata.hh:419-429)
inline bool fragmentLess(const WorkFragment &a, const WorkFragment &b)
{
    return a.f.contigId < b.f.contigId ||
           (a.f.contigId == b.f.contigId &&
            (a.f.position < b.f.position ||
             (a.f.position == b.f.position &&
              (a.f.reverse < b.f.reverse || (a.f.reverse == b.f.reverse && a.f.observedLength < b.f.observedLength)))));
}

/// FragmentBuilder::consolidateDuplicateFragments (FragmentBuilder.cpp:279-324) on list[0..n); returns the new size.
/// std::sort is the same libstdc++ introsort the reference runs, on the same comparator and input order, so the entry
/// (and its firstSeedIndex) that survives a group of duplicates is the same one.
inline unsigned consolidateDuplicateFragments(WorkFragment *list, unsigned n, bool removeUnaligned)
{
    std::sort(list, list + n, fragmentLess);
    unsigned first = 0;
    while (first != n && removeUnaligned && !list[first].f.cigarLength) ++first;
    if (first) { std::copy(list + first, list + n, list); n -= first; }
    if (n < 2) return n;
    unsigned last = 0;
    for (unsigned cur = 1; cur != n; ++cur)
    {
        if (removeUnaligned && !list[cur].f.cigarLength) continue;
        isaac_ext_fragment_t &l = list[last].f;
        const isaac_ext_fragment_t &c = list[cur].f;
        if (l.position == c.position && l.contigId == c.contigId && l.reverse == c.reverse && l.observedLength == c.observedLength)
        {
            l.uniqueSeedCount = uint16_t(l.uniqueSeedCount + c.uniqueSeedCount);        // FragmentMetadata::consolidate (:470-475)
            l.nonUniqueSeedOffsetFirst = std::min(l.nonUniqueSeedOffsetFirst, c.nonUniqueSeedOffsetFirst);
            l.nonUniqueSeedOffsetSecond = std::max(l.nonUniqueSeedOffsetSecond, c.nonUniqueSeedOffsetSecond);
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
