// Device-side views of the data that stays resident in HBM: the 2-bit packed reference with its N-mask and the
// decoded read set of the current tile.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/isaac_ext.h"

// Accessors the CPU harnesses of tests/cpp run too (the end clippers, tests/cpp/test_clip_host.cu): device code is unchanged,
// __ldg is the read-only load only where there is a device
#define ISAAC_VIEW_FN __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define ISAAC_VIEW_LOAD(p) __ldg(p)
#else
#define ISAAC_VIEW_LOAD(p) (*(p))
#endif

namespace isaac_b200
{

// Base codes used inside the kernels.  A,C,G,T = 0..3; the read's 'n' and the reference's 'N' get two different
// codes above 3 so that a plain equality test reproduces the reference's raw byte compare inside Smith-Waterman
// (BandedSmithWaterman.cpp:205: 'n' != 'N' != anything), while isMatch (Alignment.hh:44-47) is
// read == CODE_READ_N || (read == ref) -- ref 'N' can never equal a read code.
enum : unsigned { CODE_READ_N = 4, CODE_REF_N = 5 };

/// All contigs back to back, each starting on a 128-base boundary, padded with 128 zero bases at the end.
struct ReferenceView
{
    const uint32_t *bases2;          // 16 bases per word, base g at bits 2*(g%16)
    const uint32_t *nmask;           // 32 bases per word, bit set = 'N'
    const uint64_t *contigOffset;    // global base index of the first base of each contig
    const uint64_t *contigLength;
    uint32_t contigCount;
    uint64_t totalBases;             // padded size of the packed arrays in bases

    ISAAC_VIEW_FN unsigned code(uint64_t g) const
    {
        const unsigned c = (ISAAC_VIEW_LOAD(bases2 + (g >> 4)) >> ((unsigned(g) & 15u) * 2u)) & 3u;
        const unsigned n = (ISAAC_VIEW_LOAD(nmask + (g >> 5)) >> (unsigned(g) & 31u)) & 1u;
        return n ? unsigned(CODE_REF_N) : c;
    }
};

/// One tile's reads, forward strand only; the reverse strand is read back to front and complemented on the fly
/// (Read::decodeBcl, Read.cpp:32-73, builds exactly that).  readId = cluster * readCount + readIndex.
struct ReadSetView
{
    const uint32_t *bases2;          // readId * words2 + w, 16 bases per word
    const uint32_t *nmask;           // readId * wordsN + w, 32 bases per word, bit set = 'n'
    const uint8_t *quality;          // readId * qualityStride + i (BCL N gets quality 2)
    const uint16_t *endCyclesMasked; // per readId
    // Both strands again as 4-bit codes in strand order (0..3, CODE_READ_N), 16 bases per 64-bit word, for the
    // Smith-Waterman query stream: (readId * 2 + reverse) * wordsC + w.  The last TWO words of every strand are spare
    // zeros so that a 16-base fetch may start anywhere in the strand.
    const uint64_t *codes4;
    // The same strands once more for the scorer, 16 bases per 64-bit word at the same index: bits 0..31 the 2-bit codes in
    // strand order (complemented on the reverse strand), bits 32..47 one flag per base = 'n'.
    const uint64_t *strand2;
    uint32_t wordsC;
    const uint8_t *qualityStrand;    // qualities again per strand in strand order: (readId * 2 + reverse) * qualityStride + p
    uint32_t words2, wordsN, qualityStride;
    uint32_t readCount;
    uint32_t readLength[2];
    uint32_t firstCycle[2];
    uint32_t readTotal;              // clusterCount * readCount

    ISAAC_VIEW_FN unsigned length(unsigned readId) const { return readLength[readId % readCount]; }
    __device__ __forceinline__ const uint64_t *strandCodes(unsigned readId, bool reverse) const
    {
        return codes4 + (size_t(readId) * 2 + (reverse ? 1u : 0u)) * wordsC;
    }
    __device__ __forceinline__ const uint64_t *strandWords2(unsigned readId, bool reverse) const
    {
        return strand2 + (size_t(readId) * 2 + (reverse ? 1u : 0u)) * wordsC;
    }
    __device__ __forceinline__ const uint8_t *strandQuality(unsigned readId, bool reverse) const
    {
        return qualityStrand + (size_t(readId) * 2 + (reverse ? 1u : 0u)) * qualityStride;
    }
    /// highest strand position a 16-base fetch may start at
    __device__ __forceinline__ unsigned codesClamp() const { return (wordsC - 2u) * 16u; }

    /// base code and quality of strand-order position i
    ISAAC_VIEW_FN unsigned code(unsigned readId, unsigned L, bool reverse, unsigned i, unsigned &q) const
    {
        const unsigned f = reverse ? L - 1 - i : i;
        q = quality[size_t(readId) * qualityStride + f];
        const unsigned c = (bases2[size_t(readId) * words2 + (f >> 4)] >> ((f & 15u) * 2u)) & 3u;
        const unsigned n = (nmask[size_t(readId) * wordsN + (f >> 5)] >> (f & 31u)) & 1u;
        return n ? unsigned(CODE_READ_N) : (reverse ? 3u - c : c);
    }
};

/// Normalised penalties of AlignerBase (AlignerBase.cpp:32-43) and the Smith-Waterman scores (GappedAligner.cpp:41).
struct ScoreParams
{
    uint32_t mismatch, gapOpen, gapExtend, maxGapExtend;   // unsigned like the reference (AlignerBase.hh:51-54)
    int swMatch, swMismatch, swOpen, swExtend;             // open/extend positive
    const double *logMatch;                                // 100 entries, host libm (Quality.cpp:34-66)
    const double *logMismatch;
};

} // namespace isaac_b200
