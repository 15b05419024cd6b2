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

/// 16 base codes (4 bits each, strand order) starting at strand position 'pos' of a read strand, from the
/// nibble-packed copy of the read set (ReadSetView::codes4; two spare words per strand make w[1] readable).
__device__ __forceinline__ uint64_t readCodes16(const uint64_t *__restrict__ strandWords, unsigned pos)
{
    const uint64_t *w = strandWords + (pos >> 4);
    const unsigned s = (pos & 15u) * 4u;
    const uint64_t w0 = w[0], w1 = w[1];
    return s ? (w0 >> s) | (w1 << (64u - s)) : w0;
}

__device__ __forceinline__ uint64_t spread2to4(uint32_t x)
{
    uint64_t r = x;
    r = (r | (r << 16)) & 0x0000FFFF0000FFFFull;
    r = (r | (r << 8)) & 0x00FF00FF00FF00FFull;
    r = (r | (r << 4)) & 0x0F0F0F0F0F0F0F0Full;
    r = (r | (r << 2)) & 0x3333333333333333ull;
    return r;
}

__device__ __forceinline__ uint64_t spread1to4(uint32_t m16)
{
    uint64_t r = m16;
    r = (r | (r << 24)) & 0x000000FF000000FFull;
    r = (r | (r << 12)) & 0x000F000F000F000Full;
    r = (r | (r << 6)) & 0x0303030303030303ull;
    r = (r | (r << 3)) & 0x1111111111111111ull;
    return r;
}

/// 16 reference base codes (4 bits each: 0..3, CODE_REF_N for 'N') starting at global base index g.
__device__ __forceinline__ uint64_t referenceCodes16(const ReferenceView &ref, uint64_t g)
{
    const uint32_t *b = ref.bases2 + (g >> 4);
    const uint32_t two = __funnelshift_r(__ldg(b), __ldg(b + 1), (unsigned(g) & 15u) * 2u);
    const uint32_t *m = ref.nmask + (g >> 5);
    const uint32_t n16 = __funnelshift_r(__ldg(m), __ldg(m + 1), unsigned(g) & 31u) & 0xFFFFu;
    return spread2to4(two) | (spread1to4(n16) * uint64_t(CODE_REF_N));
}

/// AlignerBase::updateFragmentCigar (AlignerBase.cpp:121-227): walks the CIGAR against the reference and fills
/// the scores of 'out'.  logProbability is the reference's left-to-right FP64 sum starting from 0.0 (soft-clipped
/// bases add logMatch, inserted bases add nothing, AlignerBase.cpp:165-213).
///
/// The walk is organised as ONE loop over the L read bases with the CIGAR operation as loop-carried state, so that
/// all threads of a warp run the same trip count whatever their CIGARs look like; deletions are consumed when the
/// operation changes.  \return matchCount
__device__ __forceinline__ unsigned scoreCigar(const ReferenceView &ref, const ReadSetView &reads, const ScoreParams &sp,
                                               const unsigned readId, const unsigned L, const bool reverse,
                                               const uint64_t contigOffset, const long strandPosition,
                                               const uint32_t *cigar, const unsigned nOps,
                                               isaac_ext_fragment_t &out, uint64_t *mask)
{
    uint64_t g = contigOffset + uint64_t(strandPosition);
    const uint64_t *strandWords = reads.strandCodes(readId, reverse);
    const uint8_t *quality = reads.quality + size_t(readId) * reads.qualityStride;
    unsigned matchCount = 0, mismatchCount = 0, matchesInARow = 0, gapCount = 0, editDistance = 0, sws = 0, run = 0;
    double lp = 0.0;
    unsigned k = 0, remaining = 0, op = ISAAC_EXT_CIGAR_SOFT_CLIP;
    uint64_t qBuf = 0, rBuf = 0;
    unsigned rLeft = 0;
    for (unsigned p = 0; p < L; ++p)
    {
        if ((p & 15u) == 0) qBuf = readCodes16(strandWords, p);
        const unsigned rc = unsigned(qBuf) & 15u;
        qBuf >>= 4;
        const unsigned q = quality[reverse ? L - 1 - p : p];
        while (remaining == 0 && k < nOps)
        {
            if (op == ISAAC_EXT_CIGAR_ALIGN) { matchesInARow = max(matchesInARow, run); }   // :183
            const uint32_t word = cigar[k++];
            const unsigned length = word >> 4;
            op = word & 0xFu;
            if (op == ISAAC_EXT_CIGAR_DELETE)                                              // :192-198
            {
                g += length; editDistance += length; ++gapCount; rLeft = 0;
                sws += sp.gapOpen + min(sp.maxGapExtend, (length - 1) * sp.gapExtend);
            }
            else
            {
                remaining = length;
                run = 0;                                                                   // :158
                if (op == ISAAC_EXT_CIGAR_INSERT)                                          // :185-191
                {
                    editDistance += length; ++gapCount;
                    sws += sp.gapOpen + min(sp.maxGapExtend, (length - 1) * sp.gapExtend);
                }
            }
        }
        if (op == ISAAC_EXT_CIGAR_ALIGN)
        {
            if (rLeft == 0) { rBuf = referenceCodes16(ref, g); rLeft = 16; }
            const unsigned gc = unsigned(rBuf) & 15u;
            rBuf >>= 4; --rLeft; ++g;
            if (rc == CODE_READ_N || rc == gc)                   // isMatch (Alignment.hh:44-47)
            {
                ++matchCount; ++run;
                lp += sp.logMatch[q];
            }
            else
            {
                matchesInARow = max(matchesInARow, run); run = 0;
                if (mask) mask[p >> 6] |= 1ull << (p & 63u);     // addMismatchCycle (:171), as a bit over base index
                ++mismatchCount;
                lp += sp.logMismatch[q];
                sws += sp.mismatch;
            }
            editDistance += rc != gc;                             // byte inequality, so Ns count (:175-179)
        }
        else if (op == ISAAC_EXT_CIGAR_SOFT_CLIP)                 // :199-213
        {
            lp += sp.logMatch[q];
        }
        --remaining;
    }
    if (op == ISAAC_EXT_CIGAR_ALIGN) matchesInARow = max(matchesInARow, run);
    // a well-formed CIGAR has no operation left here (trailing deletions are stripped, BandedSmithWaterman.cpp:447-452)
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
