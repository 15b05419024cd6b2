// Scoring of a CIGAR against the resident reference (AlignerBase::updateFragmentCigar) and the clipping rules of
// AlignerBase, as device functions shared by the ungapped, gapped and simple-indel kernels.
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

/// The mutable part of FragmentMetadata a kernel thread works on (FragmentMetadata.hh:330-414).
struct FragmentState
{
    int64_t position;
    uint32_t lowClipped, highClipped;
    bool reverse;

    // FragmentMetadata::incrementClipLeft / incrementClipRight (FragmentMetadata.hh:284-285)
    __device__ __forceinline__ void incrementClipLeft(unsigned bases)
    {
        position += bases;
        if (reverse) highClipped += bases; else lowClipped += bases;
    }
    __device__ __forceinline__ void incrementClipRight(unsigned bases)
    {
        if (reverse) lowClipped += bases; else highClipped += bases;
    }
};

/// AlignerBase::clipReadMasking (AlignerBase.cpp:89-119); begin/end index the strand-order sequence.
/// Read::getBeginCyclesMasked() is always 0 (Read.hh:79).
__device__ __forceinline__ void clipReadMasking(unsigned L, unsigned endCyclesMasked, FragmentState &f, long &begin, long &end)
{
    const long maskedBegin = f.reverse ? long(endCyclesMasked) : 0L;
    const long maskedEnd = f.reverse ? long(L) : long(L) - long(endCyclesMasked);
    if (maskedBegin > begin) { f.incrementClipLeft(unsigned(maskedBegin - begin)); begin = maskedBegin; }
    if (maskedEnd < end) { f.incrementClipRight(unsigned(end - maskedEnd)); end = maskedEnd; }
}

/// AlignerBase::clipReference (AlignerBase.cpp:50-82).  The second branch (:74-81) is reached when quality trimming
/// moved the first unmasked base of a reverse-strand read past the contig end: the position is pulled back to the
/// last contig base and the mapped range becomes empty.  (begin cannot go below 0 for candidates the host accepts,
/// position <= contigLength - 2 or no left masking; the reference would walk off its sequence there.)
__device__ __forceinline__ void clipReference(long referenceSize, FragmentState &f, long &begin, long &end)
{
    const long referenceLeft = referenceSize - f.position;
    if (referenceLeft >= 0)
    {
        if (referenceLeft < end - begin) end = begin + referenceLeft;
        if (0 > f.position) { begin -= f.position; f.position = 0; }
        end = max(end, begin);
    }
    else
    {
        f.position += referenceLeft - 1;
        begin = max(0L, begin + referenceLeft - 2);
        end = begin;
    }
}

/// AlignerBase::updateFragmentCigar (AlignerBase.cpp:121-227): walks the CIGAR against the reference and fills
/// the scores of 'out'.  logProbability is the reference's left-to-right FP64 sum starting from 0.0 (soft-clipped
/// bases add logMatch, inserted bases add nothing).  \return matchCount
__device__ __forceinline__ unsigned scoreCigar(const ReferenceView &ref, const ReadSetView &reads, const ScoreParams &sp,
                                               const unsigned readId, const unsigned L, const bool reverse,
                                               const uint64_t contigOffset, const long strandPosition,
                                               const uint32_t *cigar, const unsigned nOps,
                                               isaac_ext_fragment_t &out, uint64_t *mask)
{
    uint64_t g = contigOffset + uint64_t(strandPosition);
    unsigned currentBase = 0, matchCount = 0;
    unsigned mismatchCount = 0, matchesInARow = 0, gapCount = 0, editDistance = 0, sws = 0;
    double lp = 0.0;
    for (unsigned k = 0; k < nOps; ++k)
    {
        const uint32_t word = cigar[k];
        const unsigned length = word >> 4, op = word & 0xFu;
        if (op == ISAAC_EXT_CIGAR_ALIGN)
        {
            unsigned run = 0;
            for (unsigned j = 0; j < length; ++j, ++g, ++currentBase)
            {
                unsigned q;
                const unsigned rc = reads.code(readId, L, reverse, currentBase, q);
                const unsigned gc = ref.code(g);
                if (rc == CODE_READ_N || rc == gc)               // isMatch (Alignment.hh:44-47)
                {
                    ++matchCount; ++run;
                    lp += sp.logMatch[q];
                }
                else
                {
                    matchesInARow = max(matchesInARow, run); run = 0;
                    if (mask) mask[currentBase >> 6] |= 1ull << (currentBase & 63u);
                    ++mismatchCount;
                    lp += sp.logMismatch[q];
                    sws += sp.mismatch;
                }
                editDistance += rc != gc;                         // byte inequality, so Ns count (:175-179)
            }
            matchesInARow = max(matchesInARow, run);
        }
        else if (op == ISAAC_EXT_CIGAR_INSERT)
        {
            currentBase += length; editDistance += length; ++gapCount;
            sws += sp.gapOpen + min(sp.maxGapExtend, (length - 1) * sp.gapExtend);
        }
        else if (op == ISAAC_EXT_CIGAR_DELETE)
        {
            g += length; editDistance += length; ++gapCount;
            sws += sp.gapOpen + min(sp.maxGapExtend, (length - 1) * sp.gapExtend);
        }
        else   // SOFT_CLIP (:199-213)
        {
            for (unsigned j = 0; j < length; ++j)
            {
                const unsigned f = reverse ? L - 1 - (currentBase + j) : currentBase + j;
                lp += sp.logMatch[reads.quality[size_t(readId) * reads.qualityStride + f]];
            }
            currentBase += length;
        }
    }
    out.observedLength = uint32_t(g - contigOffset - uint64_t(strandPosition));
    out.position = strandPosition;
    out.logProbability = lp;
    out.mismatchCount = uint16_t(mismatchCount);
    out.matchesInARow = uint16_t(matchesInARow);
    out.gapCount = uint16_t(gapCount);
    out.editDistance = uint16_t(editDistance);
    out.smithWatermanScore = sws;
    out.matchCount = uint16_t(matchCount);
    return matchCount;
}

} // namespace isaac_b200
