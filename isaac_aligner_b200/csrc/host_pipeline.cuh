// Small host-side helpers of the tile calls (page-locked buffers, phase timing, a fixed-partition parallel loop) and the record
// helpers the kernels of kernels_tile.cuh share with them: the work record of a candidate, the reference's packed Match fields,
// the 1e-7-tolerant comparisons, the adoption of a kernel's result into a work record, the 5-clause acceptance rule.
// (Round 1 ran the bookkeeping of FragmentBuilder::build between the kernel passes on host threads from here; it is all kernels
// now, see kernels_tile.cuh.)
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
    uint32_t pool;            // 0 ungapped, 1 simple indel, 2 gapped: the pass whose CIGAR pool holds the record's words
    uint32_t slot;            // dense index of the record in the kernel pass that last scored it
};

__host__ __device__ inline bool lpEquals(double a, double b) { const double d = a - b; return 0.0000001 >= (d < 0 ? -d : d); }      // ISAAC_LP_EQUALS, Quality.hh:104-107
__host__ __device__ inline bool lpLess(double a, double b) { return !lpEquals(a, b) && a < b; }             // ISAAC_LP_LESS,   Quality.hh:109-112

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
