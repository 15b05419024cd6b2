// Sequencing-adapter clipping: matchSelector::FragmentSequencingAdapterClipper and
// matchSelector::SequencingAdapter as two small kernels that run in front of the scoring kernels.
//
//   adapterInitKernel   checkInitStrand (FragmentSequencingAdapterClipper.cpp:102-147): one thread per clipper slot = the
//                       (read, strand) of one FragmentBuilder read list or one rescueShadow call.  The FIRST candidate the
//                       reference's clipper sees on a strand decides where the adapter lies in that strand's sequence
//                       (SURVEY D9); the host names that candidate, the kernel walks it against the reference and looks the
//                       5-mer behind every mismatch up in the adapters' k-mer tables (SequencingAdapter.cpp:58-139).
//   adapterClipKernel   clip + decideWhichSideToClip (:149-277): one thread per candidate, turns the slot's adapter range
//                       into this candidate's clipped [begin, end) of the strand sequence (it depends on the candidate's own
//                       reference window), one packed word per candidate that the scoring kernels apply before the quality
//                       and reference clipping (UngappedAligner.cpp:59, GappedAligner.cpp:186).
//
// Without adapters (every BASELINE config) neither kernel runs and the scoring kernels get a null pointer.
#pragma once
#include "device_types.cuh"
#include "score.cuh"

namespace isaac_b200
{

constexpr unsigned ADAPTER_KMER = 5;                  // SequencingAdapter::adapterMatchBasesMin_ (SequencingAdapter.hh:40)
constexpr unsigned ADAPTER_KMERS = 1u << (2 * ADAPTER_KMER);
constexpr unsigned ADAPTER_STRIDE = 128;              // bytes per adapter sequence (length < 127, SequencingAdapter.cpp:35)
constexpr uint32_t ADAPTER_NO_CANDIDATE = 0xFFFFFFFFu;

/// flowcell::SequencingAdapterMetadata + SequencingAdapter::kmerPositions_ of every adapter, in list order
struct AdapterView
{
    uint32_t count;
    const uint8_t *codes;          // count * ADAPTER_STRIDE base codes 0..3
    const int8_t *kmerPositions;   // count * ADAPTER_KMERS: position of the 5-mer in the adapter, -1 unknown, -2 not unique
    const uint32_t *length;        // count
    const uint32_t *clipLength;    // count, 0 = unbounded (SequencingAdapterMetadata.hh:58)
    const uint8_t *reverse;        // count
};

/// strandAdapters_.strandRange_[reverse] of one clipper: [begin, end) in the strand sequence; empty = nothing to clip
struct AdapterRange
{
    uint16_t begin, end; uint32_t empty;
};

struct StrandBases
{
    const ReadSetView &reads; unsigned readId, L; bool reverse;
    __device__ __forceinline__ unsigned operator()(long i) const { unsigned q; return reads.code(readId, L, reverse, unsigned(i), q); }
};

/// isMatch (Alignment.hh:44-47)
__device__ __forceinline__ bool isMatchCode(unsigned read, unsigned reference) { return read == CODE_READ_N || read == reference; }

/// SequencingAdapter::getMatchRange (SequencingAdapter.cpp:58-139); all positions are offsets into the strand sequence.
/// \return true and [first, second) if the adapter is recognised around mismatchBase
__device__ inline bool adapterMatchRange(const AdapterView &ad, const unsigned a, const StrandBases &seq, const long sequenceBegin,
                                         const long sequenceEnd, const long mismatchBase, long &first, long &second)
{
    if (sequenceEnd - mismatchBase < long(ADAPTER_KMER)) return false;                           // oligo::generateKmer (KmerGenerator.hpp:150-169)
    unsigned kmer = 0;
    for (unsigned j = 0; j < ADAPTER_KMER; ++j) kmer = ((kmer << 2) | seq(mismatchBase + j)) & 0xFFFFu;   // 'n' translates to 4 and spills into the base before
    kmer &= ADAPTER_KMERS - 1u;
    const int pos = ad.kmerPositions[a * ADAPTER_KMERS + kmer];
    if (pos < 0) return false;                                                                   // isGoodPosition
    const bool unbounded = ad.clipLength[a] == 0, adapterReverse = ad.reverse[a] != 0;
    const unsigned mismatchBaseOffset = unsigned(mismatchBase - sequenceBegin);
    const unsigned adapterBasesBeforeSequence = mismatchBaseOffset < unsigned(pos) ? unsigned(pos) - mismatchBaseOffset : 0u;
    if (adapterBasesBeforeSequence && unbounded) return false;                                   // :74, :127-132
    const long testBase = mismatchBase - (long(pos) - long(adapterBasesBeforeSequence));
    const unsigned testSequenceLength = unsigned(sequenceEnd - testBase);
    const unsigned adapterSequenceSize = ad.length[a];
    const unsigned leftClippedAdapterLength = adapterSequenceSize - adapterBasesBeforeSequence;
    const unsigned overlapLength = min(testSequenceLength, leftClippedAdapterLength);
    if (overlapLength < leftClippedAdapterLength && unbounded && adapterReverse) return false;   // :81-88
    if (overlapLength < ADAPTER_KMER) return false;                                              // :91
    for (unsigned j = 0; j < overlapLength; ++j)
        if (ad.codes[a * ADAPTER_STRIDE + adapterBasesBeforeSequence + j] != seq(testBase + j)) return false;
    if (adapterReverse)                                                                          // :99-105
    {
        first = unbounded ? sequenceBegin
                          : testBase - long(min(unsigned(testBase - sequenceBegin), ad.clipLength[a] - adapterSequenceSize));
        second = testBase + overlapLength;
    }
    else                                                                                         // :107-111
    {
        first = testBase;
        second = unbounded ? sequenceEnd : testBase + long(min(overlapLength, ad.clipLength[a]));
    }
    return first != second;
}

/// checkInitStrand (FragmentSequencingAdapterClipper.cpp:102-147) for the first candidate of a strand
__device__ inline AdapterRange adapterInitStrand(const AdapterView &ad, const ReferenceView &ref, const ReadSetView &reads,
                                                 const isaac_ext_candidate_t c)
{
    const bool reverse = c.contigStrand & 1u;
    const unsigned contigId = c.contigStrand >> 1;
    const unsigned L = reads.length(c.readId);
    const StrandBases seq = {reads, c.readId, L, reverse};
    const long referenceSize = long(ref.contigLength[contigId]);
    const uint64_t contigOffset = ref.contigOffset[contigId];
    // the clipper's own clipReference (:40-59)
    long sequenceBegin = 0, sequenceEnd = L;
    const long referenceLeft = referenceSize - c.position;
    if (referenceLeft < sequenceEnd - sequenceBegin) sequenceEnd = sequenceBegin + referenceLeft;
    long newFragmentPos = c.position;
    if (0 > c.position) { sequenceBegin -= c.position; newFragmentPos = 0; }
    long adapterRangeBegin = sequenceEnd, adapterRangeEnd = sequenceBegin;                       // :124-125
    for (unsigned a = 0; a < ad.count; ++a)
    {
        // SequencingAdapter::isStrandCompatible (SequencingAdapter.hh:58-61)
        if (ad.clipLength[a] == 0 && reverse != (ad.reverse[a] != 0)) continue;
        // findSequencingAdapter (:79-100) from the end of what has been found so far
        const long searchBegin = adapterRangeEnd;
        long reference = newFragmentPos + (searchBegin - sequenceBegin);
        for (long current = searchBegin; current < sequenceEnd; ++current, ++reference)
        {
            if (!isMatchCode(seq(current), ref.code(contigOffset + uint64_t(reference))))
            {
                long first, second;
                if (adapterMatchRange(ad, a, seq, searchBegin, sequenceEnd, current, first, second))
                {
                    adapterRangeBegin = min(first, adapterRangeBegin);                           // :138-139
                    adapterRangeEnd = max(second, adapterRangeEnd);
                    break;
                }
            }
        }
    }
    AdapterRange r;
    r.begin = uint16_t(max(adapterRangeBegin, 0L)); r.end = uint16_t(max(adapterRangeEnd, 0L));
    r.empty = sequenceBegin == adapterRangeEnd ? 1u : 0u;                                        // :145
    return r;
}

/// countMatches / countMismatches over [sequenceBegin, sequenceEnd) against the contig at referenceBegin (Alignment.hh:55-87)
__device__ inline unsigned adapterCountMatches(const ReferenceView &ref, const uint64_t contigOffset, const StrandBases &seq,
                                               const long sequenceBegin, const long sequenceEnd, const long referenceBegin)
{
    unsigned matches = 0;
    for (long i = sequenceBegin; i < sequenceEnd; ++i)
        matches += isMatchCode(seq(i), ref.code(contigOffset + uint64_t(referenceBegin + (i - sequenceBegin)))) ? 1u : 0u;
    return matches;
}

/// clip + decideWhichSideToClip (:149-277) for one candidate.  \return the packed clipped range: begin | end << 16
__device__ inline uint32_t adapterClipCandidate(const AdapterRange range, const ReferenceView &ref, const ReadSetView &reads,
                                                const isaac_ext_candidate_t c)
{
    const unsigned L = reads.length(c.readId);
    if (range.empty) return L << 16;
    const bool reverse = c.contigStrand & 1u;
    const unsigned contigId = c.contigStrand >> 1;
    const StrandBases seq = {reads, c.readId, L, reverse};
    const long referenceSize = long(ref.contigLength[contigId]);
    const uint64_t contigOffset = ref.contigOffset[contigId];
    const long sequenceBegin = 0, sequenceEnd = L, contigPosition = c.position;
    const long rangeBegin = min(long(range.begin), sequenceEnd), rangeEnd = min(long(range.end), sequenceEnd);
    const unsigned backwardsClipped = unsigned(rangeBegin - sequenceBegin);                      // :158-159
    const unsigned forwardsClipped = unsigned(sequenceEnd - rangeEnd);
    bool clipBackwards = backwardsClipped < forwardsClipped;
    bool clip = true;
    const unsigned sequenceLength = L;
    const int difference = int(backwardsClipped - forwardsClipped);
    if (backwardsClipped && forwardsClipped && (difference < 0 ? -difference : difference) < 9)  // :166 (abs of the int conversion)
    {
        if (contigPosition >= 0 && referenceSize >= contigPosition + long(sequenceLength))       // :169
        {
            const long referenceEnd = contigPosition + sequenceLength;
            const unsigned backwardsMatches = adapterCountMatches(ref, contigOffset, seq, sequenceBegin, rangeBegin, contigPosition);
            const unsigned forwardsMatches = adapterCountMatches(ref, contigOffset, seq, rangeEnd, sequenceEnd, referenceEnd - forwardsClipped);
            clipBackwards = backwardsMatches < forwardsMatches || (backwardsMatches == forwardsMatches && backwardsClipped < forwardsClipped);
        }
    }
    else if (!backwardsClipped || !forwardsClipped)                                              // :190-216
    {
        // clipping all the way to one end: the clipped side must carry a decent amount of mismatches.  The reference reads
        // its contig without a bounds check here; positions outside the contig count as mismatches in this build.
        auto percentMismatches = [&](const long b, const long e, const long referenceBegin) {
            unsigned mismatches = 0;
            for (long i = b; i < e; ++i)
            {
                const long r = referenceBegin + (i - b);
                const bool inside = r >= 0 && r < referenceSize;
                mismatches += inside && isMatchCode(seq(i), ref.code(contigOffset + uint64_t(r))) ? 0u : 1u;
            }
            return mismatches * 100u / unsigned(e - b);
        };
        if (clipBackwards && !backwardsClipped)
            clip = percentMismatches(sequenceBegin, rangeEnd, contigPosition) > 40u;             // TOO_GOOD_READ_MISMATCH_PERCENT
        else if (!clipBackwards && !forwardsClipped)
        {
            const unsigned basesClipped = unsigned(sequenceEnd - rangeBegin);
            clip = percentMismatches(rangeBegin, sequenceEnd, contigPosition + long(sequenceLength) - long(basesClipped)) > 40u;
        }
    }
    if (!clip) return L << 16;
    return clipBackwards ? uint32_t(rangeEnd) | (L << 16) : uint32_t(rangeBegin) << 16;           // :240-262
}

/// what the scoring kernels do with the packed word in place of FragmentSequencingAdapterClipper::clip
__device__ __forceinline__ void applyAdapterClip(const uint32_t packed, const unsigned L, FragmentState &f, long &begin, long &end)
{
    const unsigned b = packed & 0xFFFFu, e = (packed >> 16) & 0x7FFFu;       // bit 31: kernels_avoid.cuh
    if (b) { f.incrementClipLeft(b); begin = b; }
    if (e < L) { f.incrementClipRight(L - e); end = e; }
}

__global__ void adapterInitKernel(const AdapterView ad, const ReferenceView ref, const ReadSetView reads, uint32_t slots,
                                  const isaac_ext_candidate_t *__restrict__ firstCandidates, AdapterRange *__restrict__ ranges)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = firstCandidates[s];
        AdapterRange r = {0, 0, 1u};
        if (c.readId != ADAPTER_NO_CANDIDATE) r = adapterInitStrand(ad, ref, reads, c);
        ranges[s] = r;
    }
}

/// slotOf == nullptr: slot = readId * 2 + reverse (the clipper of a FragmentBuilder read list, FragmentBuilder.cpp:164)
__global__ void adapterClipKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n,
                                  const isaac_ext_candidate_t *__restrict__ candidates, const uint32_t *__restrict__ slotOf,
                                  const AdapterRange *__restrict__ ranges, uint32_t *__restrict__ clipOut)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = candidates[i];
        if (c.readId == ADAPTER_NO_CANDIDATE) continue;      // a match slot of the tile pipeline that holds no candidate
        const uint32_t slot = slotOf ? slotOf[i] : c.readId * 2u + (c.contigStrand & 1u);
        clipOut[i] = adapterClipCandidate(ranges[slot], ref, reads, c);
    }
}

/// rescueShadow keeps one clipper per call (ShadowAligner.cpp:207): request r owns candidates [taskBegin[r], +taskCount[r])
/// of the pool; its first candidate position initialises the strand (:222).  One warp per request.
__global__ void shadowAdapterSlotsKernel(uint32_t requests, const uint32_t *__restrict__ taskBegin, const uint32_t *__restrict__ taskCount,
                                         const isaac_ext_candidate_t *__restrict__ pool, isaac_ext_candidate_t *__restrict__ first,
                                         uint32_t *__restrict__ slotOf)
{
    const uint32_t lane = threadIdx.x & 31u, warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < requests; r += warpsPerGrid)
    {
        const uint32_t begin = taskBegin[r], count = taskCount[r];
        if (lane == 0)
        {
            isaac_ext_candidate_t c;
            c.position = 0; c.readId = ADAPTER_NO_CANDIDATE; c.contigStrand = 0;
            first[r] = count ? pool[begin] : c;
        }
        for (uint32_t k = lane; k < count; k += 32u) slotOf[begin + k] = r;
    }
}

/// the micro entry points: every candidate is its own clipper (checkInitStrand + clip on the same candidate, the way the
/// reference's testSequencingAdapter.cpp:159-182 drives one alignment)
__global__ void adapterSelfClipKernel(const AdapterView ad, const ReferenceView ref, const ReadSetView reads, uint32_t n,
                                      const isaac_ext_candidate_t *__restrict__ candidates, uint32_t *__restrict__ clipOut)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = candidates[i];
        clipOut[i] = adapterClipCandidate(adapterInitStrand(ad, ref, reads, c), ref, reads, c);
    }
}

} // namespace isaac_b200
