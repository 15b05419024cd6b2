// Kernels of the candidate-extension path (first, scalar generation).
//   packReferenceKernel   K0  ASCII contig -> 2-bit + N-mask            (ContigLoader product, ContigLoader.cpp:29-65)
//   decodeBclKernel       K0b BCL bytes -> 2-bit + n-mask + qualities   (Read::decodeBcl, Read.cpp:32-73)
//   ungappedKernel        K1  UngappedAligner::alignUngapped            (UngappedAligner.cpp:39-92)
//                                                                        (GappedAligner.cpp:167-249)
//   bandedSwAsciiKernel   K2  BandedSmithWaterman::align on explicit (query, database) strings
#pragma once
#include "device_types.cuh"
#include "score.cuh"
#include "kernels_adapter.cuh"
#include "sw.cuh"

namespace isaac_b200
{

__device__ __forceinline__ unsigned asciiRefCode(unsigned char c)
{
    switch (c)
    {
    case 'A': case 'a': return 0u;
    case 'C': case 'c': return 1u;
    case 'G': case 'g': return 2u;
    case 'T': case 't': return 3u;
    default: return CODE_REF_N;
    }
}

/// One thread packs 32 bases: two 2-bit words and one mask word.  'ascii' holds 'count' bases that land at global
/// base index 'dstBase' (a multiple of 32).
__global__ void packReferenceKernel(const unsigned char *__restrict__ ascii, uint64_t count, uint64_t dstBase,
                                    uint32_t *__restrict__ bases2, uint32_t *__restrict__ nmask)
{
    const uint64_t groups = (count + 31) / 32;
    for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < groups; t += uint64_t(gridDim.x) * blockDim.x)
    {
        uint32_t lo = 0, hi = 0, m = 0;
        const uint64_t first = t * 32;
#pragma unroll 8
        for (unsigned k = 0; k < 32; ++k)
        {
            const uint64_t i = first + k;
            const unsigned c = i < count ? asciiRefCode(ascii[i]) : 0u;
            const unsigned two = c > 3u ? 0u : c;
            if (k < 16) lo |= two << (2 * k); else hi |= two << (2 * (k - 16));
            m |= (c > 3u ? 1u : 0u) << k;
        }
        const uint64_t w = (dstBase + first) >> 5;
        bases2[2 * w] = lo; bases2[2 * w + 1] = hi; nmask[w] = m;
    }
}

/// One thread decodes 32 consecutive cycles of one read.
__global__ void decodeBclKernel(const uint8_t *__restrict__ bcl, uint32_t clusterCount, uint32_t readCount,
                                uint32_t len0, uint32_t len1, uint32_t words2, uint32_t wordsN, uint32_t qualityStride,
                                uint32_t *__restrict__ bases2, uint32_t *__restrict__ nmask, uint8_t *__restrict__ quality)
{
    const uint32_t groupsPerRead = wordsN;
    const uint64_t total = uint64_t(clusterCount) * readCount * groupsPerRead;
    const uint32_t clusterBytes = len0 + (readCount > 1 ? len1 : 0);
    for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < total; t += uint64_t(gridDim.x) * blockDim.x)
    {
        const uint32_t grp = uint32_t(t % groupsPerRead);
        const uint64_t readId = t / groupsPerRead;
        const uint32_t readIndex = uint32_t(readId % readCount);
        const uint64_t cluster = readId / readCount;
        const uint32_t L = readIndex ? len1 : len0;
        const uint8_t *src = bcl + cluster * clusterBytes + (readIndex ? len0 : 0);
        uint32_t lo = 0, hi = 0, m = 0;
#pragma unroll 8
        for (unsigned k = 0; k < 32; ++k)
        {
            const uint32_t i = grp * 32 + k;
            if (i < L)
            {
                const unsigned b = src[i];
                const bool isN = !(b & 0xfcu);                              // oligo::isBclN (Nucleotides.hh:91-94)
                const unsigned two = isN ? 0u : (b & 3u);
                if (k < 16) lo |= two << (2 * k); else hi |= two << (2 * (k - 16));
                m |= (isN ? 1u : 0u) << k;
                quality[readId * qualityStride + i] = isN ? 2 : uint8_t(b >> 2);   // Read.cpp:60,66
            }
        }
        bases2[readId * words2 + 2 * grp] = lo;
        if (2 * grp + 1 < words2) bases2[readId * words2 + 2 * grp + 1] = hi;
        nmask[readId * wordsN + grp] = m;
    }
}

/// Both strands of every read as 4-bit codes in strand order (ReadSetView::codes4); one thread per 64-bit word.
__global__ void encodeStrandCodesKernel(const uint8_t *__restrict__ bcl, uint32_t clusterCount, uint32_t readCount,
                                        uint32_t len0, uint32_t len1, uint32_t wordsC, uint64_t *__restrict__ codes4,
                                        uint64_t *__restrict__ strand2, uint32_t qualityStride, uint8_t *__restrict__ qualityStrand)
{
    const uint64_t total = uint64_t(clusterCount) * readCount * 2 * wordsC;
    const uint32_t clusterBytes = len0 + (readCount > 1 ? len1 : 0);
    for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < total; t += uint64_t(gridDim.x) * blockDim.x)
    {
        const uint32_t w = uint32_t(t % wordsC);
        const uint64_t strandId = t / wordsC;
        const bool reverse = strandId & 1u;
        const uint64_t readId = strandId >> 1;
        const uint32_t readIndex = uint32_t(readId % readCount);
        const uint32_t L = readIndex ? len1 : len0;
        const uint8_t *src = bcl + (readId / readCount) * clusterBytes + (readIndex ? len0 : 0);
        uint64_t word = 0, two = 0;
        for (unsigned k = 0; k < 16; ++k)
        {
            const uint32_t p = w * 16 + k;
            if (p < L)
            {
                const unsigned b = src[reverse ? L - 1 - p : p];
                const unsigned code = !(b & 0xfcu) ? unsigned(CODE_READ_N) : (reverse ? 3u - (b & 3u) : (b & 3u));   // Read.cpp:56-69
                word |= uint64_t(code) << (4 * k);
                two |= code > 3u ? 1ull << (32 + k) : uint64_t(code) << (2 * k);
                qualityStrand[strandId * qualityStride + p] = !(b & 0xfcu) ? 2 : uint8_t(b >> 2);                 // Read.cpp:60,66
            }
            else if (p < qualityStride) qualityStrand[strandId * qualityStride + p] = 0;
        }
        codes4[t] = word;
        strand2[t] = two;
    }
}

__device__ __forceinline__ void initFragment(isaac_ext_fragment_t &o, const isaac_ext_candidate_t &c, uint32_t readCount)
{
    o.position = c.position; o.logProbability = 0.0; o.contigId = c.contigStrand >> 1; o.readId = c.readId;
    o.cigarOffset = 0; o.smithWatermanScore = 0; o.observedLength = 0; o.mismatchCount = 0; o.matchesInARow = 0;
    o.gapCount = 0; o.editDistance = 0; o.uniqueSeedCount = 0; o.repeatSeedsCount = 0;
    o.nonUniqueSeedOffsetFirst = 0xFFFF; o.nonUniqueSeedOffsetSecond = 0; o.firstSeedIndex = -1;
    o.lowClipped = 0; o.highClipped = 0; o.cigarLength = 0; o.reverse = uint8_t(c.contigStrand & 1u);
    o.readIndex = uint8_t(c.readId % readCount); o.matchCount = 0;
}

/// K1: one candidate per thread.
__global__ void ungappedKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, uint32_t n,
                               const isaac_ext_candidate_t *__restrict__ candidates,
                               isaac_ext_fragment_t *__restrict__ fragments, uint32_t *__restrict__ cigars,
                               uint64_t *__restrict__ masks, const uint32_t *__restrict__ adapterClip = nullptr)
{
    __shared__ double tables[201];
    const ScoreParams sp = stageScoreTables(spGlobal, tables);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = candidates[i];
        if (c.readId == 0xFFFFFFFFu) continue;          // a match slot of the tile pipeline that holds no candidate (kernels_tile.cuh)
        isaac_ext_fragment_t o;
        initFragment(o, c, reads.readCount);
        const unsigned contigId = c.contigStrand >> 1;
        const unsigned L = reads.length(c.readId);
        uint64_t *mask = masks ? masks + size_t(i) * ISAAC_EXT_MASK_WORDS : nullptr;
        if (mask) for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) mask[k] = 0;
        uint32_t *cigar = cigars + size_t(i) * 3;
        o.cigarOffset = i * 3;
        // resetAlignment + resetClipping: position is the unclipped candidate position, clips are 0 (UngappedAligner.cpp:48-49)
        FragmentState f = {c.position, 0u, 0u, bool(c.contigStrand & 1u)};
        long begin = 0, end = L;
        if (adapterClip) applyAdapterClip(adapterClip[i], L, f, begin, end);            // :59
        clipReadMasking(L, reads.endCyclesMasked[c.readId], f, begin, end);             // :60
        clipReference(long(ref.contigLength[contigId]), f, begin, end);                 // :62
        o.lowClipped = uint16_t(f.lowClipped); o.highClipped = uint16_t(f.highClipped); o.position = f.position;
        {
            uint32_t ops[3]; unsigned nOps = 0;
            if (begin) ops[nOps++] = cigarWord(uint32_t(begin), ISAAC_EXT_CIGAR_SOFT_CLIP);            // :64-68
            if (end - begin) ops[nOps++] = cigarWord(uint32_t(end - begin), ISAAC_EXT_CIGAR_ALIGN);    // :70-75
            if (long(L) - end) ops[nOps++] = cigarWord(uint32_t(L - end), ISAAC_EXT_CIGAR_SOFT_CLIP);  // :77-81
            const unsigned matchCount = scoreUngapped(ref, reads, sp, c.readId, L, f.reverse, ref.contigOffset[contigId],
                                                      f.position, unsigned(begin), unsigned(end), o, mask);
            for (unsigned k = 0; k < 3; ++k) cigar[k] = k < nOps ? ops[k] : 0u;
            o.cigarLength = matchCount ? uint16_t(nOps) : 0;                                           // setUnaligned (:86-89)
        }
        fragments[i] = o;
    }
}

constexpr unsigned SW_OPS_CAP = 64;

struct AsciiBaseSrc
{
    const unsigned char *query; const unsigned char *database;
    __device__ __forceinline__ static unsigned qcode(unsigned char c)
    {
        switch (c) { case 'A': return 0u; case 'C': return 1u; case 'G': return 2u; case 'T': return 3u; default: return CODE_READ_N; }
    }
    __device__ __forceinline__ unsigned q(unsigned i) const { return qcode(query[i]); }
    __device__ __forceinline__ unsigned d(unsigned k) const { return asciiRefCode(database[k]); }
};

} // namespace isaac_b200

namespace isaac_b200
{
/// Integer-pipe throughput probe for the Smith-Waterman roofline denominator (MEASURED_PEAKS.json has no INT32 figure).
/// 8 independent chains per thread, 'iters' x 16 x 8 operations per thread.  kind 0: add.s32 (IADD3 / IMAD.IADD),
/// 1: max.s32 (VIMNMX), 2: packed 16x2 max (VIMNMX.S16x2, counted as 2 operations by the caller).
template <int KIND> __global__ void intPeakKernel(int iters, unsigned seed, unsigned *out)
{
    unsigned a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 8u + k + seed;
    const unsigned b = (blockIdx.x + seed) | 1u;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 16; ++u)
        {
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                if (KIND == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                else if (KIND == 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b + u));
                else a[k] = __vmaxs2(a[k], b + u);
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 0x12345678u) out[0] = s;
}
} // namespace isaac_b200
