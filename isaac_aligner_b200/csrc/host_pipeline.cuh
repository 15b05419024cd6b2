// Host side of the two TemplateBuilder-facing calls: the per-cluster bookkeeping of FragmentBuilder::build and
// ShadowAligner::rescueShadow (candidate lists, std::sort + consolidate, pairing, acceptance rules) between the kernel
// passes.  Everything here is integer bookkeeping on fragment records; every base comparison, score, k-mer scan and
// Smith-Waterman cell is computed by the kernels (kernels*.cuh).  The reference keeps this logic per cluster and per
// thread (MatchSelector.cpp:258-368); here a tile is processed phase by phase:
//
//   build:   P1 candidates (addMatch, repeat filter, consolidate)      -> K1 ungappedKernel
//            P2 consolidate, pair adjacent candidates                   -> simpleIndelKernel
//            P3 apply patches, consolidate, pick mismatchCount > 5      -> swForwardKernel + swTraceScoreKernel
//            P4 acceptance rule, consolidate, flatten
//   rescue:  R1 rescue windows from the template length statistics      -> shadowCandidates*Kernel, K1 ungappedKernel
//            (everything behind R1 is on the device: the list bookkeeping R2 / R3 and the flat result are kernels_rescue.cuh)
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "kernels2.cuh"
#include "kernels_indel.cuh"
#include "kernels_shadow.cuh"

namespace isaac_b200
{

template <class T> struct PinnedBuffer
{
    T *p = nullptr; size_t capacity = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= capacity) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; capacity = 0;
        const size_t want = std::max<size_t>(n + n / 4, 1024);
        const cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&p), want * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) capacity = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; capacity = 0; }
};

/// Phase timing to stderr when the environment variable ISAAC_EXT_TRACE is set (development aid).
struct PhaseTimer
{
    bool on; const char *what; std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(const char *w) : on(std::getenv("ISAAC_EXT_TRACE") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *phase)
    {
        if (!on) return;
        const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[isaac_ext] %s %-28s %8.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

/// runs f(threadIndex, begin, end) over a fixed partition of [0, n) (the same partition on every call with the same n)
template <class F> void parallelRanges(unsigned threads, size_t n, F f)
{
    threads = unsigned(std::max<size_t>(1, std::min<size_t>(threads, n / 64 + 1)));
    if (threads == 1) { f(0u, size_t(0), n); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back([=]() { f(t, n * t / threads, n * (t + 1) / threads); });
    for (std::thread &th : pool) th.join();
}
inline unsigned partitionCount(unsigned threads, size_t n) { return unsigned(std::max<size_t>(1, std::min<size_t>(threads, n / 64 + 1))); }

/// A fragment record being worked on plus the pool its CIGAR currently lives in.
struct WorkFragment
{
    isaac_ext_fragment_t f;
    uint32_t pool;            // index into HostPools::pools
    uint32_t slot;            // dense index of the record in the kernel pass that last scored it
};

struct HostPools
{
    const uint32_t *pools[4] = {nullptr, nullptr, nullptr, nullptr};      // 0 ungapped, 1 simple indel, 2 gapped
    const uint32_t *cigar(const WorkFragment &w) const { return pools[w.pool] + w.f.cigarOffset; }
    long beginClipped(const WorkFragment &w) const                        // FragmentMetadata::getBeginClippedLength (:148-159)
    {
        if (!w.f.cigarLength) return 0;
        const uint32_t word = cigar(w)[0];
        return (word & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP ? long(word >> 4) : 0;
    }
    long endClipped(const WorkFragment &w) const                          // getEndClippedLength (:161-172)
    {
        if (!w.f.cigarLength) return 0;
        const uint32_t word = cigar(w)[w.f.cigarLength - 1];
        return (word & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP ? long(word >> 4) : 0;
    }
    long unclippedPosition(const WorkFragment &w) const { return long(w.f.position) - beginClipped(w); }   // :185-188
};

__host__ __device__ inline bool lpEquals(double a, double b) { const double d = a - b; return 0.0000001 >= (d < 0 ? -d : d); }      // ISAAC_LP_EQUALS, Quality.hh:104-107
__host__ __device__ inline bool lpLess(double a, double b) { return !lpEquals(a, b) && a < b; }             // ISAAC_LP_LESS,   Quality.hh:109-112

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

/* the reference's packed Match fields (SeedId.hh:37-127, ReferencePosition.hh:51-188) */
__host__ __device__ inline unsigned matchSeed(const isaac_ext_match_t &m) { return unsigned((m.seedId >> 1) & 0xFF); }
__host__ __device__ inline bool matchReverse(const isaac_ext_match_t &m) { return m.seedId & 1; }
__host__ __device__ inline bool matchIsNoMatch(const isaac_ext_match_t &m) { return m.location == (((~uint64_t(0)) >> 41) << 41); }
__host__ __device__ inline bool matchIsTooMany(const isaac_ext_match_t &m) { return (m.location >> 1) == 0; }
__host__ __device__ inline unsigned matchContig(const isaac_ext_match_t &m) { return unsigned(m.location >> 41) - 1; }
__host__ __device__ inline long matchPosition(const isaac_ext_match_t &m) { return long((m.location >> 1) & ((uint64_t(1) << 40) - 1)); }
__host__ __device__ inline bool matchHasNeighbors(const isaac_ext_match_t &m) { return m.location & 1; }

/// the kernel-computed fields of 'scored' replace those of 'w'; the seed bookkeeping of 'w' stays
__host__ __device__ inline void adoptAlignment(WorkFragment &w, const isaac_ext_fragment_t &scored, uint32_t pool, uint32_t slot)
{
    isaac_ext_fragment_t &f = w.f;
    f.position = scored.position; f.logProbability = scored.logProbability; f.cigarOffset = scored.cigarOffset;
    f.smithWatermanScore = scored.smithWatermanScore; f.observedLength = scored.observedLength;
    f.mismatchCount = scored.mismatchCount; f.matchesInARow = scored.matchesInARow; f.gapCount = scored.gapCount;
    f.editDistance = scored.editDistance; f.lowClipped = scored.lowClipped; f.highClipped = scored.highClipped;
    f.cigarLength = scored.cigarLength; f.matchCount = scored.matchCount;
    w.pool = pool; w.slot = slot;
}

/// the 5-clause acceptance rule of the gapped alignment (FragmentBuilder.cpp:202-205, ShadowAligner.cpp:259-262)
__host__ __device__ inline bool acceptGapped(const isaac_ext_fragment_t &ungapped, const isaac_ext_fragment_t &gapped, unsigned gappedMismatchesMax)
{
    const unsigned observed = ungapped.cigarLength ? ungapped.observedLength : 0;       // getObservedLength()
    return gapped.matchCount && gapped.matchCount + ISAAC_EXT_BAND_WIDTH > observed &&
           gapped.mismatchCount <= gappedMismatchesMax && ungapped.mismatchCount > gapped.mismatchCount &&
           lpLess(ungapped.logProbability, gapped.logProbability);
}

} // namespace isaac_b200
