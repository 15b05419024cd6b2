// What MatchSelector does to a cluster's reads before and to its template after the hot path:
//   trimLowQualityEndsKernel   alignment::trimLowQualityEnds (Quality.cpp:71-120): quality trimming of the read ends
//   clipTemplateEndsKernel     matchSelector::SemialignedEndsClipper::clip (SemialignedEndsClipper.cpp:32-205, clipMismatches
//                              Alignment.hh:55-87) then matchSelector::OverlappingEndsClipper::clip
//                              (OverlappingEndsClipper.cpp:45-180) on every kept template
// Both are a handful of base / quality comparisons per read on data that is already resident.
#pragma once
#include "device_types.cuh"
#include "sw.cuh"

namespace isaac_b200
{

/// one read per thread
__global__ void trimLowQualityEndsKernel(const ReadSetView reads, const uint32_t baseQualityCutoff, uint16_t *__restrict__ endCyclesMasked)
{
    const unsigned MASK_READ_LENGTH_MIN = 35;                                                    // Quality.cpp:71
    for (uint32_t readId = blockIdx.x * blockDim.x + threadIdx.x; readId < reads.readTotal; readId += gridDim.x * blockDim.x)
    {
        const unsigned L = reads.length(readId);
        unsigned masked = 0;
        if (baseQualityCutoff && L >= MASK_READ_LENGTH_MIN)
        {
            const uint8_t *q = reads.quality + size_t(readId) * reads.qualityStride;
            int qscoreSum = 0, peakSum = 0;
            for (unsigned k = 0; k + MASK_READ_LENGTH_MIN != L; ++k)                             // reverse qualities: last cycle first
            {
                qscoreSum += int(baseQualityCutoff) - int(q[L - 1 - k]);
                if (qscoreSum < 0) break;
                if (qscoreSum > peakSum) { peakSum = qscoreSum; masked = k + 1; }                 // trims the base at trimPos as well
            }
        }
        endCyclesMasked[readId] = uint16_t(masked);
    }
}

constexpr unsigned CLIP_CIGAR_CAP = 64;

/// a template fragment being clipped: record + CIGAR in registers / local memory
struct ClipFragment
{
    isaac_ext_fragment_t f;
    uint32_t cigar[CLIP_CIGAR_CAP];
    unsigned n;
    unsigned L;
};

ISAAC_VIEW_FN bool clipIsAligned(const ClipFragment &x) { return x.n != 0; }

/// clipMismatches<5> (Alignment.hh:55-87) over strand positions seq0, seq0 + step, ... (count of them) against reference
/// bases g0, g0 + step, ... (refCount of them)
ISAAC_VIEW_FN void clipMismatches(const ReferenceView &ref, const ReadSetView &reads, const ClipFragment &x,
                                               long seq0, long g0, int step, unsigned count, uint64_t refCount,
                                               unsigned &clippedBases, unsigned &clippedEdits)
{
    const unsigned CONSECUTIVE_MATCHES_MIN = 5;                                                  // SemialignedEndsClipper.hh:37
    unsigned matchesInARow = 0, editDistanceMismatches = 0, editDistanceMismatchesUnclipped = 0, ret = 0;
    while (ret != count && uint64_t(ret) != refCount && CONSECUTIVE_MATCHES_MIN > matchesInARow)
    {
        unsigned q;
        const unsigned s = reads.code(x.f.readId, x.L, x.f.reverse != 0, unsigned(seq0 + long(ret) * step), q);
        const unsigned r = ref.code(uint64_t(g0 + long(ret) * step));
        const bool differ = s != r;                                   // 'n' differs from everything, 'N' from every read base
        if (s == CODE_READ_N || !differ) { ++matchesInARow; editDistanceMismatchesUnclipped += differ; }   // isMatch (:44-47)
        else { matchesInARow = 0; editDistanceMismatchesUnclipped = 0; }
        editDistanceMismatches += differ;
        ++ret;
    }
    const bool found = CONSECUTIVE_MATCHES_MIN == matchesInARow;
    clippedBases = found ? ret - matchesInARow : 0u;
    clippedEdits = found ? editDistanceMismatches - editDistanceMismatchesUnclipped : 0u;
}

/// SemialignedEndsClipper::clipLeftSide (:32-93)
ISAAC_VIEW_FN bool clipLeftSide(const ReferenceView &ref, const ReadSetView &reads, ClipFragment &x)
{
    unsigned first = 0;
    uint32_t op = x.cigar[0];
    unsigned softClippedBeginBases = 0;
    if ((op & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP)
    {
        if (2 > x.n) return false;
        first = 1; softClippedBeginBases = op >> 4;
        op = x.cigar[1];
    }
    if ((op & 0xFu) != ISAAC_EXT_CIGAR_ALIGN) return false;
    unsigned mappedBeginBases = op >> 4;
    const uint64_t contigOffset = ref.contigOffset[x.f.contigId], contigLength = ref.contigLength[x.f.contigId];
    unsigned clippedBases, clippedEdits;
    clipMismatches(ref, reads, x, long(softClippedBeginBases), long(contigOffset) + x.f.position, 1, mappedBeginBases,
                   contigLength - uint64_t(x.f.position), clippedBases, clippedEdits);
    if (!clippedBases) return false;
    x.f.observedLength -= clippedBases;
    softClippedBeginBases += clippedBases;
    mappedBeginBases -= clippedBases;
    x.f.position += clippedBases;
    x.f.editDistance -= uint16_t(clippedEdits);
    // [S][M] + the operations after the first ALIGN
    if (first == 0)
    {
        for (unsigned k = x.n; k > 1; --k) x.cigar[k] = x.cigar[k - 1];
        ++x.n;
    }
    x.cigar[0] = cigarWord(softClippedBeginBases, ISAAC_EXT_CIGAR_SOFT_CLIP);
    x.cigar[1] = cigarWord(mappedBeginBases, ISAAC_EXT_CIGAR_ALIGN);
    return true;
}

/// SemialignedEndsClipper::clipRightSide (:95-157)
ISAAC_VIEW_FN bool clipRightSide(const ReferenceView &ref, const ReadSetView &reads, ClipFragment &x)
{
    unsigned last = x.n - 1;
    uint32_t op = x.cigar[last];
    unsigned softClippedEndBases = 0;
    if ((op & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP)
    {
        if (2 > x.n) return false;
        --last; softClippedEndBases = op >> 4;
        op = x.cigar[last];
    }
    if ((op & 0xFu) != ISAAC_EXT_CIGAR_ALIGN) return false;
    unsigned mappedEndBases = op >> 4;
    const uint64_t contigOffset = ref.contigOffset[x.f.contigId];
    const long endPosition = x.f.position + long(x.f.observedLength);                              // one past the last aligned base
    unsigned clippedBases, clippedEdits;
    clipMismatches(ref, reads, x, long(x.L) - 1 - long(softClippedEndBases), long(contigOffset) + endPosition - 1, -1, mappedEndBases,
                   uint64_t(endPosition), clippedBases, clippedEdits);
    if (!clippedBases) return false;
    x.f.observedLength -= clippedBases;
    softClippedEndBases += clippedBases;
    x.f.editDistance -= uint16_t(clippedEdits);
    mappedEndBases -= clippedBases;
    x.cigar[last] = cigarWord(mappedEndBases, ISAAC_EXT_CIGAR_ALIGN);
    x.cigar[last + 1] = cigarWord(softClippedEndBases, ISAAC_EXT_CIGAR_SOFT_CLIP);
    x.n = last + 2;
    return true;
}

/// OverlappingEndsClipper::clip (:45-180)
ISAAC_VIEW_FN void clipOverlappingEnds(const ReferenceView &ref, const ReadSetView &reads, ClipFragment &r1, ClipFragment &r2)
{
    if (!clipIsAligned(r1) || !clipIsAligned(r2) || r1.f.gapCount || r2.f.gapCount) return;
    // (:62-66 compares r1.contigId with itself: chimeric pairs are NOT skipped)
    if ((r1.f.reverse != 0) == (r2.f.reverse != 0)) return;
    ClipFragment &left = r1.f.position < r2.f.position ? r1 : r2;
    ClipFragment &right = r1.f.position <= r2.f.position ? r2 : r1;
    if (left.f.reverse) return;
    const long overlapLength = left.f.position + long(left.f.observedLength) - right.f.position;
    if (0 >= overlapLength) return;
    // the overlapping end of the left read
    unsigned leftEndSoftClip = 0, leftEndOffset = left.L, leftLast = left.n - 1;
    uint32_t leftOp = left.cigar[leftLast];
    if ((leftOp & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP)
    {
        leftEndOffset -= leftOp >> 4; leftEndSoftClip = leftOp >> 4;
        --leftLast; leftOp = left.cigar[leftLast];
    }
    if (overlapLength >= long(leftOp >> 4)) return;
    // the overlapping start of the right read
    unsigned rightStartOffset = 0, rightFirst = 0;
    uint32_t rightOp = right.cigar[0];
    if ((rightOp & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP)
    {
        rightStartOffset += rightOp >> 4;
        rightFirst = 1; rightOp = right.cigar[1];
    }
    if (overlapLength >= long(rightOp >> 4)) return;
    const unsigned overlap = unsigned(overlapLength);
    // diffBaseQualities (:33-43): left forward qualities minus right reverse(strand)-order qualities over the overlap
    int diff = 0;
    for (unsigned i = 0; i < overlap; ++i)
    {
        unsigned ql, qr;
        reads.code(left.f.readId, left.L, false, leftEndOffset - overlap + i, ql);
        reads.code(right.f.readId, right.L, true, rightStartOffset + i, qr);
        diff += int(ql) - int(qr);
    }
    if (0 < diff)
    {
        // left one is better, clip right (:127-149)
        const uint64_t g = ref.contigOffset[right.f.contigId] + uint64_t(right.f.position);
        unsigned edits = 0;
        for (unsigned i = 0; i < overlap; ++i)
        {
            unsigned q;
            edits += reads.code(right.f.readId, right.L, true, rightStartOffset + i, q) != ref.code(g + i);
        }
        if (rightFirst == 0)
        {
            for (unsigned k = right.n; k > 1; --k) right.cigar[k] = right.cigar[k - 1];
            ++right.n;
        }
        right.cigar[0] = cigarWord(rightStartOffset + overlap, ISAAC_EXT_CIGAR_SOFT_CLIP);
        right.cigar[1] = cigarWord((rightOp >> 4) - overlap, ISAAC_EXT_CIGAR_ALIGN);
        right.f.position += overlap;                                                              // incrementClipLeft (FragmentMetadata.hh:284)
        if (right.f.reverse) right.f.highClipped += uint16_t(overlap); else right.f.lowClipped += uint16_t(overlap);
        right.f.observedLength -= overlap;
        right.f.editDistance -= uint16_t(edits);
    }
    else
    {
        // right one is better, clip left (:151-176)
        const uint64_t g = ref.contigOffset[left.f.contigId] + uint64_t(left.f.position + long(left.f.observedLength) - overlapLength);
        unsigned edits = 0;
        for (unsigned i = 0; i < overlap; ++i)
        {
            unsigned q;
            edits += reads.code(left.f.readId, left.L, false, leftEndOffset - overlap + i, q) != ref.code(g + i);
        }
        left.cigar[leftLast] = cigarWord((leftOp >> 4) - overlap, ISAAC_EXT_CIGAR_ALIGN);
        left.cigar[leftLast + 1] = cigarWord(leftEndSoftClip + overlap, ISAAC_EXT_CIGAR_SOFT_CLIP);
        left.n = leftLast + 2;
        if (left.f.reverse) left.f.lowClipped += uint16_t(overlap); else left.f.highClipped += uint16_t(overlap);   // incrementClipRight (:285)
        left.f.observedLength -= overlap;
        left.f.editDistance -= uint16_t(edits);
    }
}

/// The end clippers on the template of cluster c (MatchSelector.cpp:336-346).  fragments[c * readCount + r] / cigarsIn + cigarOffset
/// in; the record and its (possibly longer) CIGAR out at cigarsOut + cigarOffset + 4 * (c * readCount + r), where the caller left room
/// for cigarLength + 4 words per fragment.  \return false, nothing written, when a CIGAR does not fit the clippers' scratch.
ISAAC_VIEW_FN bool clipTemplateEndsOfCluster(const ReferenceView &ref, const ReadSetView &reads, const uint32_t c, const uint32_t clipFlags,
                                             const isaac_ext_template_t *__restrict__ templates, isaac_ext_fragment_t *__restrict__ fragments,
                                             const uint32_t *__restrict__ cigarsIn, uint32_t *__restrict__ cigarsOut)
{
    const unsigned readCount = reads.readCount;
    ClipFragment x[2];
    bool tooLong = false;
    for (unsigned r = 0; r < readCount; ++r)
    {
        const size_t i = size_t(c) * readCount + r;
        x[r].f = fragments[i];
        x[r].n = x[r].f.cigarLength;
        x[r].L = reads.length(x[r].f.readId);
        if (x[r].n + 4 > CLIP_CIGAR_CAP) { tooLong = true; x[r].n = 0; continue; }
        for (unsigned k = 0; k < x[r].n; ++k) x[r].cigar[k] = cigarsIn[x[r].f.cigarOffset + k];
    }
    if (tooLong) return false;
    if (templates[c].built)
    {
        if (clipFlags & ISAAC_EXT_CLIP_SEMIALIGNED)                                           // SemialignedEndsClipper::clip (:183-205)
        {
            for (unsigned k = 0; k < readCount; ++k)
            {
                if (!clipIsAligned(x[k])) continue;
                bool changed = clipLeftSide(ref, reads, x[k]);
                if (clipRightSide(ref, reads, x[k])) changed = true;
                if (changed && 2 == readCount)
                {
                    ClipFragment &mate = x[1 - k];
                    if (!clipIsAligned(mate)) { mate.f.position = x[k].f.position; break; }
                }
            }
        }
        if ((clipFlags & ISAAC_EXT_CLIP_OVERLAPPING) && 2 == readCount) clipOverlappingEnds(ref, reads, x[0], x[1]);
    }
    for (unsigned r = 0; r < readCount; ++r)
    {
        const size_t i = size_t(c) * readCount + r;
        const uint32_t out = x[r].f.cigarOffset + 4u * uint32_t(i);
        for (unsigned k = 0; k < x[r].n; ++k) cigarsOut[out + k] = x[r].cigar[k];
        x[r].f.cigarOffset = x[r].n ? out : x[r].f.cigarOffset;
        x[r].f.cigarLength = uint16_t(x[r].n);
        fragments[i] = x[r].f;
    }
    return true;
}

/// One cluster per thread.
__global__ void clipTemplateEndsKernel(const ReferenceView ref, const ReadSetView reads, const uint32_t clusterCount,
                                       const uint32_t clipFlags, const isaac_ext_template_t *__restrict__ templates,
                                       isaac_ext_fragment_t *__restrict__ fragments, const uint32_t *__restrict__ cigarsIn,
                                       uint32_t *__restrict__ cigarsOut, uint32_t *__restrict__ errorFlag)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < clusterCount; c += gridDim.x * blockDim.x)
        if (!clipTemplateEndsOfCluster(ref, reads, c, clipFlags, templates, fragments, cigarsIn, cigarsOut)) atomicOr(errorFlag, 4u);
}

} // namespace isaac_b200
