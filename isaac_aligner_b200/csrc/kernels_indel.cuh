// SimpleIndelAligner on the GPU: one thread per (head, tail) pair of adjacent candidates of a read.
//
// Restates SimpleIndelAligner::alignSimpleDeletion / alignSimpleInsertion (reference
// lib/alignment/fragmentBuilder/SimpleIndelAligner.cpp:50-229, 241-438).  The host enumerates the pairs exactly like
// alignSimpleIndels does (:460-518); every pair only reads the pre-pass state of its two fragments and only patches
// the earlier one, so all pairs of a read -- and of a tile -- are independent and run in one launch.
#pragma once
#include "device_types.cuh"
#include "score.cuh"
#include "sw.cuh"

namespace isaac_b200
{

/// The fields of one fragment the indel search reads (FragmentMetadata getters used in SimpleIndelAligner.cpp).
struct IndelSide
{
    int64_t position;                 // FragmentMetadata::position
    uint32_t beginClipped, endClipped;// getBeginClippedLength / getEndClippedLength
    uint32_t observedLength;          // getObservedLength() (0 when unaligned)
    uint32_t smithWatermanScore, mismatchCount;
    uint16_t lowClipped, highClipped;
    uint32_t seedOffset, seedLength;  // strand-order offset of the anchoring seed (:493-494)
    __host__ __device__ long unclippedPosition() const { return long(position) - long(beginClipped); }
};

struct IndelTask
{
    IndelSide head, tail;             // named as in the function that handles the pair
    uint32_t readId, contigId;
    uint8_t reverse, insertion, pad[6];
};

struct IndelResult
{
    isaac_ext_fragment_t fragment;    // the patched fragment (seed bookkeeping fields are filled by the host)
    uint32_t cigar[5];
    uint32_t accepted;
};

constexpr unsigned GAP_FLANK_BASES = 32, GAP_FLANK_MISMATCHES_MAX = 8;     // SimpleIndelAligner.hh:36-37

struct IndelView
{
    const ReferenceView &ref; const uint64_t *strandWords; uint64_t contigOffset; long contigLength;
    __device__ __forceinline__ unsigned readCode(long p) const { return unsigned(strandWords[p >> 4] >> ((unsigned(p) & 15u) * 4u)) & 15u; }
    /// !isMatch(read[p], reference[r]) (Alignment.hh:44-47).  Reference positions outside the contig (the reference
    /// would read past its vector there) count as mismatches.
    __device__ __forceinline__ unsigned mismatch(long p, long r) const
    {
        if (r < 0 || r >= contigLength) return 1u;
        const unsigned rc = readCode(p), gc = ref.code(contigOffset + uint64_t(r));
        return !(rc == CODE_READ_N || rc == gc);
    }
    /// countMismatches (Alignment.hh:119-159): stops at the end of the contig
    __device__ __forceinline__ unsigned count(long p, long r, unsigned length) const
    {
        unsigned n = 0;
        for (unsigned i = 0; i < length && r + long(i) < contigLength; ++i) n += mismatch(p + i, r + i);
        return n;
    }
};

__global__ void simpleIndelKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, uint32_t n,
                                  const IndelTask *__restrict__ tasks, IndelResult *__restrict__ results,
                                  const uint32_t *__restrict__ slots = nullptr, const uint32_t *__restrict__ slotCount = nullptr,
                                  uint32_t *__restrict__ cigarsOut = nullptr)
{
    __shared__ double tables[201];
    const ScoreParams sp = stageScoreTables(spGlobal, tables);
    // tile pipeline (kernels_tile.cuh): the pairs live in the match slots slots[0 .. *slotCount) (a dense list made by
    // cub::DeviceSelect), results go to the same slots; otherwise tasks[0 .. n)
    const uint32_t count = slotCount ? *slotCount : n;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x)
    {
        const uint32_t i = slots ? slots[j] : j;
        const IndelTask t = tasks[i];
        IndelResult &out = results[i];
        out.accepted = 0;
        const IndelSide &head = t.head, &tail = t.tail;
        const unsigned L = reads.length(t.readId);
        const IndelView v = {ref, reads.strandCodes(t.readId, t.reverse), ref.contigOffset[t.contigId], long(ref.contigLength[t.contigId])};
        if (head.seedOffset < head.beginClipped) continue;                                               // :60-64, :252-256
        if (tail.beginClipped + tail.observedLength < tail.seedOffset + tail.seedLength) continue;        // :66-72, :258-262
        uint32_t ops[5]; unsigned nOps = 0;
        long strandPosition; uint16_t low, high;
        if (!t.insertion)
        {
            // ---- alignSimpleDeletion: the deletion may sit inside the head seed (:77)
            const unsigned tailOffset = head.seedOffset;
            long tailIt = tailOffset;
            unsigned tailLength = tail.beginClipped + tail.observedLength - tailOffset;                  // :85
            const unsigned tailMismatches = v.count(tailIt, head.unclippedPosition() + tailOffset, tailLength);
            if (!tailMismatches) continue;                                                               // :90-94
            const unsigned deletionLength = unsigned(tail.unclippedPosition() - head.unclippedPosition());
            unsigned rightRealigned = v.count(tailIt, tail.unclippedPosition() + tailOffset, tailLength);
            unsigned leftRealigned = 0;
            const unsigned lf = min(GAP_FLANK_BASES, tailOffset);
            unsigned leftFlank = v.count(tailIt - lf, head.unclippedPosition() + tailOffset - lf, lf);   // :106-108
            unsigned rightFlank = v.count(tailIt, tail.unclippedPosition() + tailOffset, min(GAP_FLANK_BASES, tailLength));
            long refIt = head.unclippedPosition() + tailOffset;
            unsigned best = tailMismatches, bestLeftFlank = leftFlank, bestRightFlank = rightFlank, bestOffset = ~0u;
            for (unsigned deletionOffset = tailOffset; best && deletionOffset <= tail.seedOffset;
                 ++deletionOffset, ++tailIt, ++refIt, --tailLength)                                      // :126-161
            {
                const unsigned thisOffset = leftRealigned + rightRealigned;
                if (best > thisOffset) { bestOffset = deletionOffset; best = thisOffset; bestLeftFlank = leftFlank; bestRightFlank = rightFlank; }
                const unsigned newLeft = v.mismatch(tailIt, refIt);
                leftRealigned += newLeft; leftFlank += newLeft;
                if (deletionOffset >= GAP_FLANK_BASES) leftFlank -= v.mismatch(tailIt - GAP_FLANK_BASES, refIt - GAP_FLANK_BASES);
                const unsigned disappearing = v.mismatch(tailIt, refIt + deletionLength);
                rightRealigned -= disappearing; rightFlank -= disappearing;
                if (tailLength > GAP_FLANK_BASES)
                    rightFlank += v.mismatch(tailIt + GAP_FLANK_BASES, refIt + deletionLength + GAP_FLANK_BASES);
            }
            if (!(bestLeftFlank <= GAP_FLANK_MISMATCHES_MAX && bestRightFlank <= GAP_FLANK_MISMATCHES_MAX && bestOffset != ~0u)) continue;
            const unsigned clip = head.beginClipped;
            const unsigned leftMapped = bestOffset - clip;
            const unsigned headMismatches = v.count(clip, head.position, leftMapped);
            const unsigned newMismatches = headMismatches + best;
            const unsigned sws = sp.mismatch * newMismatches + sp.gapOpen + min(sp.maxGapExtend, (deletionLength - 1) * sp.gapExtend);
            if (!(head.smithWatermanScore > sws || (head.smithWatermanScore == sws && head.mismatchCount > newMismatches))) continue;
            long position = head.position;
            if (clip) ops[nOps++] = cigarWord(clip, ISAAC_EXT_CIGAR_SOFT_CLIP);
            if (leftMapped)
            {
                ops[nOps++] = cigarWord(leftMapped, ISAAC_EXT_CIGAR_ALIGN);
                ops[nOps++] = cigarWord(deletionLength, ISAAC_EXT_CIGAR_DELETE);
            }
            else position += deletionLength;                                                             // :191-195
            const unsigned rightMapped = head.observedLength + head.endClipped - leftMapped - tail.endClipped;
            if (rightMapped) ops[nOps++] = cigarWord(rightMapped, ISAAC_EXT_CIGAR_ALIGN);
            if (tail.endClipped) ops[nOps++] = cigarWord(tail.endClipped, ISAAC_EXT_CIGAR_SOFT_CLIP);
            strandPosition = position;            // resetAlignment unclips, updateFragmentCigar gets position + clip (:209-213)
            low = head.lowClipped; high = head.highClipped;
            if (t.reverse) low = tail.lowClipped; else high = tail.highClipped;                          // rightClipped() (:211)
        }
        else
        {
            // ---- alignSimpleInsertion: the insertion must fit between the two seeds (:268-277)
            const unsigned tailOffset = head.seedOffset + head.seedLength;
            const unsigned observedEnd = tail.beginClipped + tail.observedLength;
            const unsigned insertionLength = unsigned(head.unclippedPosition() - tail.unclippedPosition());
            if (tail.seedOffset - head.seedOffset < insertionLength + head.seedLength) continue;
            long tailIt = long(tailOffset) + insertionLength;
            unsigned tailLength = observedEnd - tailOffset - insertionLength;
            const unsigned tailMismatches = v.count(tailIt, head.unclippedPosition() + tailOffset, tailLength);
            unsigned leftFlank = v.count(tailIt - insertionLength - GAP_FLANK_BASES, head.unclippedPosition() + tailOffset - GAP_FLANK_BASES, GAP_FLANK_BASES);
            unsigned rightFlank = v.count(tailIt, head.unclippedPosition() + tailOffset, min(GAP_FLANK_BASES, tailLength));
            unsigned rightRealigned = tailMismatches, leftRealigned = 0;
            long refIt = head.unclippedPosition() + tailOffset;
            unsigned best = tailMismatches, bestOffset = tailOffset, bestLeftFlank = leftFlank, bestRightFlank = rightFlank;
            for (unsigned insertionOffset = tailOffset; best && insertionOffset <= tail.seedOffset - insertionLength;
                 ++insertionOffset, ++tailIt, ++refIt, --tailLength)                                     // :339-377
            {
                const unsigned thisOffset = leftRealigned + rightRealigned;
                if (best > thisOffset) { bestOffset = insertionOffset; best = thisOffset; bestLeftFlank = leftFlank; bestRightFlank = rightFlank; }
                const unsigned newLeft = v.mismatch(tailIt - insertionLength, refIt);
                leftRealigned += newLeft; leftFlank += newLeft;
                if (insertionOffset >= GAP_FLANK_BASES)
                    leftFlank -= v.mismatch(tailIt - insertionLength - GAP_FLANK_BASES, refIt - GAP_FLANK_BASES);
                const unsigned disappearing = v.mismatch(tailIt, refIt);
                rightRealigned -= disappearing; rightFlank -= disappearing;
                if (tailLength > GAP_FLANK_BASES) rightFlank += v.mismatch(tailIt + GAP_FLANK_BASES, refIt + GAP_FLANK_BASES);
            }
            const unsigned clip = head.beginClipped;
            const unsigned leftMapped = bestOffset - clip;
            const unsigned headMismatches = v.count(clip, head.position, leftMapped);
            const unsigned newMismatches = headMismatches + best;
            const unsigned sws = sp.mismatch * newMismatches + sp.gapOpen + min(sp.maxGapExtend, (insertionLength - 1) * sp.gapExtend);
            if (!(bestLeftFlank <= GAP_FLANK_MISMATCHES_MAX && bestRightFlank <= GAP_FLANK_MISMATCHES_MAX)) continue;
            if (!(tail.smithWatermanScore > sws || (tail.smithWatermanScore == sws && tail.mismatchCount > newMismatches))) continue;
            if (clip) ops[nOps++] = cigarWord(clip, ISAAC_EXT_CIGAR_SOFT_CLIP);
            ops[nOps++] = cigarWord(leftMapped, ISAAC_EXT_CIGAR_ALIGN);
            ops[nOps++] = cigarWord(insertionLength, ISAAC_EXT_CIGAR_INSERT);
            const unsigned rightMapped = head.observedLength + head.endClipped - leftMapped - tail.endClipped - insertionLength;
            ops[nOps++] = cigarWord(rightMapped, ISAAC_EXT_CIGAR_ALIGN);
            if (tail.endClipped) ops[nOps++] = cigarWord(tail.endClipped, ISAAC_EXT_CIGAR_SOFT_CLIP);
            strandPosition = head.position;                                                              // :421-422
            low = tail.lowClipped; high = tail.highClipped;
            if (t.reverse) high = head.highClipped; else low = head.lowClipped;                          // leftClipped() (:420)
        }
        // ---- re-score the patched CIGAR (updateFragmentCigar, :212-213 / :421-422)
        isaac_ext_fragment_t o;
        isaac_ext_candidate_t c = {strandPosition, t.readId, (t.contigId << 1) | t.reverse};
        initFragment(o, c, reads.readCount);
        scoreCigar(ref, reads, sp, t.readId, L, t.reverse != 0, v.contigOffset, strandPosition, ops, nOps, o, nullptr);
        o.lowClipped = low; o.highClipped = high;
        o.cigarLength = uint16_t(nOps);
        o.cigarOffset = i * 5;
        out.fragment = o;
        for (unsigned k = 0; k < 5; ++k) out.cigar[k] = k < nOps ? ops[k] : 0u;
        if (cigarsOut) for (unsigned k = 0; k < 5; ++k) cigarsOut[size_t(i) * 5 + k] = k < nOps ? ops[k] : 0u;      // the words as a pool of their own
        out.accepted = 1;
    }
}

} // namespace isaac_b200
