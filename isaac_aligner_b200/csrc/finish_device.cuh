// The finish pass of alignment::TemplateBuilder for host AND device code: the BamTemplate of one cluster from its candidate lists
// (the build pass) and the answers to the ShadowAligner::rescueShadow calls plan_device.cuh recorded for it (the rescue pass).
// One cluster per thread; everything a thread needs beyond registers is a slice of a global scratch buffer whose size is known
// from the sizes of the rescue answers, so there is no container, no allocation and no recursion.
//
// What makes this pass different from the plan pass is that libm decides integers here: mapping scores are
// floor(-10 * log10(other / total)) of sums of exp(logProbability) (TemplateBuilder.cpp:233-285, 398-465, 495-676, 868-1008).
// exp / log10 are glibc's own, replayed instruction for instruction (glibc_math.cuh); double -> unsigned conversions follow the
// x86-64 instruction the reference's build uses (cvttsd2si, below); the order-dependent parts are kept in the reference's order:
// sums of exp() run in list order, std::sort + std::unique_copy of the probability lists (:694-714) run as the libstdc++
// replay of sort_replay.cuh under the reference's tolerance-based comparators.
//
// The lists of equally good pairs / orphans the reference keeps are only ever read at the index --scatter-repeats picks; the pair
// list of locateBestPair (up to n0 * n1 entries) is therefore walked twice instead of stored (like plan_device.cuh), the short
// orphan lists live in the scratch slice.
//
// tests/cpp/test_template_worker.cu runs this function on the CPU between the checker's build and rescue results and
// tests/test_template_worker.py requires the reference's own TemplateBuilder's templates, bit for bit; on the GPU it is the body of
// finishTemplatesKernel (kernels_templates.cuh).
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include "../../include/isaac_ext.h"
#include "glibc_math.cuh"
#include "plan_device.cuh"
#include "sort_replay.cuh"

namespace isaac_b200
{

constexpr uint32_t FINISH_NO_MATCH_CONTIG = 0x7FFFFFu;          // ReferencePosition::MAX_CONTIG_ID (ReferencePosition.hh:177)
constexpr unsigned FINISH_DODGY_BUT_CLEAN = 10;                 // DODGY_BUT_CLEAN_ALIGNMENT_SCORE (TemplateBuilder.hh:149)
constexpr uint32_t FINISH_POOL_SHIFT = 30;                      // cigarOffset of a record inside the pipeline: pool << 30 | word index
constexpr uint32_t FINISH_POOL_MASK = (1u << FINISH_POOL_SHIFT) - 1;
constexpr uint32_t FINISH_POOL_RESCUE = 3;                      // pools 0..2 are the build pass's (ungapped, simple indel, gapped)

/// unsigned(x) / implicit double -> unsigned of the reference's x86-64 build: cvttsd2si into a 64-bit register, low half kept.
/// Out of range and NaN give the "integer indefinite" 0x8000000000000000, i.e. 0 (CUDA's conversions saturate instead).
ISAAC_HD inline uint32_t toUnsignedX86(const double x)
{
    if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return 0u;
    return uint32_t(uint64_t((long long)x));
}

ISAAC_HD inline double finishFloor(const double x)
{
#ifdef __CUDA_ARCH__
    return ::floor(x);
#else
    return __builtin_floor(x);
#endif
}

/// floor(-10.0 * log10(other / total)) as the reference evaluates it
ISAAC_HD inline double mappingScoreOf(const double other, const double total)
{
    return finishFloor(glibc_math::mul(-10.0, glibc_math::log10(other / total)));
}

struct FinishView
{
    // build pass: the candidate list of (cluster, readIndex) is fragments[listBegin[l] .. + listCount[l]), l = cluster * readCount + readIndex
    const isaac_ext_fragment_t *fragments;
    const uint32_t *listBegin, *listCount;
    const uint8_t *built;
    const uint32_t *cigarPools[4];          // [pool] of a record's cigarOffset; [FINISH_POOL_RESCUE] = CIGAR words of the rescue pass
    // rescue pass: call i of the tile answered by rescueFragments[requestFragmentBegin[i] .. requestFragmentBegin[i + 1])
    const isaac_ext_fragment_t *rescueFragments;
    const uint64_t *requestFragmentBegin;
    const uint8_t *rescued;
    const uint32_t *clusterRequestBegin;    // first call of every cluster
    // run
    uint32_t readCount;
    uint32_t tlsMin, tlsMax, bestModel[2];
    uint32_t scatterRepeats, mapqThreshold;
    int32_t dodgyAlignmentScore;
    double rogRead[2], rogAll, logMismatchQ40;
};

/// Where the alignment of a template fragment came from: the candidate record as it was when the template took it.  The flags the
/// template logic applies afterwards (setNoMatch, filterLowQualityFragments) and the end clippers leave mismatchCount and
/// mismatchCycles of the reference's FragmentMetadata alone, and TileStats::recordFragment reads them (TileStats.hh:125-142): the
/// per-cycle statistics re-derive the cycles from this alignment.
struct FinishSource
{
    int64_t position;
    uint32_t cigarOffset;           // pool << 30 | word index
    uint32_t contigId;
    uint16_t cigarLength;
    uint8_t reverse, valid;
    uint32_t pad;
};

/// FragmentMetadata as TemplateBuilder sees it; f.cigarOffset carries pool << 30 | word index
struct FinishFragment
{
    isaac_ext_fragment_t f;
    uint32_t alignmentScore;
    FinishSource src;
};

/// TemplateBuilder::ShadowProbability (TemplateBuilder.hh:165-207)
struct FinishShadowProbability
{
    uint64_t pos; double logProbability; long observedLength;
};
/// TemplateBuilder::PairProbability (:213-246)
struct FinishPairProbability
{
    FinishShadowProbability r1, r2;
};

ISAAC_HD inline bool finishIsAligned(const isaac_ext_fragment_t &f) { return f.cigarLength != 0; }
ISAAC_HD inline uint64_t finishReferencePosition(const uint64_t contigId, const uint64_t position) { return (((contigId + 1) << 40) | position) << 1; }
ISAAC_HD inline uint64_t finishFStrandPosition(const isaac_ext_fragment_t &f)                       // FragmentMetadata.hh:90-95
{
    return f.contigId != FINISH_NO_MATCH_CONTIG ? finishReferencePosition(f.contigId, uint64_t(f.position)) : (uint64_t(FINISH_NO_MATCH_CONTIG) << 40) << 1;
}
ISAAC_HD inline uint64_t finishRStrandPosition(const isaac_ext_fragment_t &f)                       // :97-103
{
    const long end = f.position + long(f.observedLength);
    return f.contigId != FINISH_NO_MATCH_CONTIG ? finishReferencePosition(f.contigId, uint64_t((end > 1 ? end : 1) - 1)) : (uint64_t(FINISH_NO_MATCH_CONTIG) << 40) << 1;
}
ISAAC_HD inline FinishShadowProbability finishShadowProbability(const isaac_ext_fragment_t &s)
{
    FinishShadowProbability p;
    p.pos = (finishFStrandPosition(s) & ~uint64_t(1)) | uint64_t(s.reverse != 0);
    p.logProbability = s.logProbability; p.observedLength = long(planObservedLength(s));
    return p;
}
ISAAC_HD inline bool finishShadowLess(const FinishShadowProbability &a, const FinishShadowProbability &b)
{
    return a.pos < b.pos ||
        (a.pos == b.pos && (planLpLess(a.logProbability, b.logProbability) ||
                            (planLpEquals(a.logProbability, b.logProbability) && a.observedLength < b.observedLength)));
}
ISAAC_HD inline bool finishShadowSame(const FinishShadowProbability &a, const FinishShadowProbability &b)
{
    return a.pos == b.pos && planLpEquals(a.logProbability, b.logProbability) && a.observedLength == b.observedLength;
}
ISAAC_HD inline double finishPairLp(const FinishPairProbability &p) { return p.r1.logProbability + p.r2.logProbability; }
ISAAC_HD inline bool finishPairLess(const FinishPairProbability &a, const FinishPairProbability &b)
{
    return a.r1.pos < b.r1.pos || (a.r1.pos == b.r1.pos &&
        (a.r2.pos < b.r2.pos || (a.r2.pos == b.r2.pos &&
            (planLpLess(finishPairLp(b), finishPairLp(a)) || (planLpEquals(finishPairLp(a), finishPairLp(b)) &&
                (a.r1.observedLength < b.r1.observedLength || (a.r1.observedLength == b.r1.observedLength &&
                    a.r2.observedLength < b.r2.observedLength)))))));
}
ISAAC_HD inline bool finishPairSame(const FinishPairProbability &a, const FinishPairProbability &b)
{
    return a.r1.pos == b.r1.pos && a.r2.pos == b.r2.pos && planLpEquals(finishPairLp(a), finishPairLp(b)) &&
        a.r1.observedLength == b.r1.observedLength && a.r2.observedLength == b.r2.observedLength;
}

/// sumUniqueShadowProbabilities (TemplateBuilder.cpp:694-703): std::sort, then std::unique_copy into a summing output iterator.
/// libstdc++'s unique_copy for forward iterators compares every element with the last one it KEPT; exp() is summed in kept order.
ISAAC_HD inline double finishSumUniqueShadows(FinishShadowProbability *v, const unsigned n)
{
    double ret = 0.0;
    sort_replay::sort(v, n, [](const FinishShadowProbability &a, const FinishShadowProbability &b) { return finishShadowLess(a, b); });
    unsigned kept = 0;
    for (unsigned i = 0; i < n; ++i)
    {
        if (i && finishShadowSame(v[kept], v[i])) continue;
        kept = i;
        ret += glibc_math::exp(v[i].logProbability);
    }
    return ret;
}
/// sumUniquePairProbabilities (:705-714)
ISAAC_HD inline double finishSumUniquePairs(FinishPairProbability *v, const unsigned n)
{
    double ret = 0.0;
    sort_replay::sort(v, n, [](const FinishPairProbability &a, const FinishPairProbability &b) { return finishPairLess(a, b); });
    unsigned kept = 0;
    for (unsigned i = 0; i < n; ++i)
    {
        if (i && finishPairSame(v[kept], v[i])) continue;
        kept = i;
        ret += glibc_math::exp(finishPairLp(v[i]));
    }
    return ret;
}

/// bytes of scratch a cluster needs: 'shadows' = fragments in the answers to its rescueShadow calls, 'candidates' = an upper
/// bound of the entries of its two candidate lists together (its matches)
ISAAC_HD inline uint64_t finishScratchBytes(const uint64_t shadows, const uint64_t candidates)
{
    return 2 * sizeof(FinishShadowProbability) * (shadows + candidates) + sizeof(FinishPairProbability) * shadows +
           2 * sizeof(uint32_t) * (2 * candidates + 2);
}

/// The per-thread TemplateBuilder.  run() = buildTemplate(..., mapqThreshold) of one cluster (TemplateBuilder.cpp:97-175).
struct FinishWorker
{
    const FinishView &v;
    uint32_t clusterId;
    const isaac_ext_fragment_t *frags[2];
    int n[2];
    uint64_t nextRequest;
    // BamTemplate
    FinishFragment bam[2]; uint32_t bamAlignmentScore; bool bamProperPair;
    // scratch slice of this cluster
    FinishShadowProbability *allShadow[2]; unsigned allShadowCount[2];
    FinishPairProbability *allPairs; unsigned allPairCount;
    uint32_t *best[2]; unsigned bestCount[2];                      // BestPairInfo::bestPairFragments: indices into frags[]
    uint32_t *bestShadow[2]; unsigned bestShadowCount[2];          // bestOrphanShadows_: indices into v.rescueFragments
    // BestPairInfo scalars of the rescued pair
    double bestLp, totalProbability; unsigned long bestScore; unsigned resolved;

    ISAAC_HD FinishWorker(const FinishView &view, const uint32_t cluster, unsigned char *scratch, const uint64_t shadows, const uint64_t candidates)
        : v(view), clusterId(cluster), nextRequest(view.clusterRequestBegin ? view.clusterRequestBegin[cluster] : 0)
    {
        for (unsigned r = 0; r < 2; ++r)
        {
            const bool have = r < v.readCount && v.built[cluster];
            const size_t l = size_t(cluster) * v.readCount + r;
            frags[r] = have ? v.fragments + v.listBegin[l] : nullptr;
            n[r] = have ? int(v.listCount[l]) : 0;
        }
        allShadow[0] = reinterpret_cast<FinishShadowProbability *>(scratch);
        allShadow[1] = allShadow[0] + (shadows + candidates);
        allPairs = reinterpret_cast<FinishPairProbability *>(allShadow[1] + (shadows + candidates));
        best[0] = reinterpret_cast<uint32_t *>(allPairs + shadows);
        best[1] = best[0] + (candidates + 1);
        bestShadow[0] = best[1] + (candidates + 1);
        bestShadow[1] = bestShadow[0] + candidates;
        allShadowCount[0] = allShadowCount[1] = allPairCount = 0;
        bestCount[0] = bestCount[1] = bestShadowCount[0] = bestShadowCount[1] = 0;
        bestLp = -DBL_MAX; totalProbability = 0.0; bestScore = ~0ul; resolved = 0;
        bamAlignmentScore = 0; bamProperPair = false;
    }

    ISAAC_HD const uint32_t *cigarOf(const isaac_ext_fragment_t &f) const { return v.cigarPools[f.cigarOffset >> FINISH_POOL_SHIFT] + (f.cigarOffset & FINISH_POOL_MASK); }

    ISAAC_HD static void setUnaligned(FinishFragment &t) { t.f.cigarLength = 0; t.f.cigarOffset = 0; t.alignmentScore = ~0u; }                 // FragmentMetadata.hh:252
    ISAAC_HD static void setNoMatch(FinishFragment &t) { setUnaligned(t); t.f.contigId = FINISH_NO_MATCH_CONTIG; t.f.position = 0; }            // :258-259
    /// FragmentMetadata(cluster, cigarBuffer, readIndex) (:63-75)
    ISAAC_HD static FinishFragment unaligned(const uint32_t readId, const unsigned readIndex)
    {
        FinishFragment t;
        isaac_ext_fragment_t &f = t.f;
        f.position = 0; f.logProbability = 0.0; f.contigId = FINISH_NO_MATCH_CONTIG; f.readId = readId; f.cigarOffset = 0; f.smithWatermanScore = 0;
        f.observedLength = 0; f.mismatchCount = 0; f.matchesInARow = 0; f.gapCount = 0; f.editDistance = 0; f.uniqueSeedCount = 0;
        f.repeatSeedsCount = 0; f.nonUniqueSeedOffsetFirst = 0xFFFF; f.nonUniqueSeedOffsetSecond = 0; f.firstSeedIndex = -1; f.lowClipped = 0;
        f.highClipped = 0; f.cigarLength = 0; f.reverse = 0; f.readIndex = uint8_t(readIndex); f.matchCount = 0;
        t.alignmentScore = ~0u;
        t.src.position = 0; t.src.cigarOffset = 0; t.src.contigId = 0; t.src.cigarLength = 0; t.src.reverse = 0; t.src.valid = 0; t.src.pad = 0;
        return t;
    }
    ISAAC_HD static void rememberSource(FinishFragment &t)
    {
        t.src.position = t.f.position; t.src.cigarOffset = t.f.cigarOffset; t.src.contigId = t.f.contigId; t.src.cigarLength = t.f.cigarLength;
        t.src.reverse = t.f.reverse; t.src.valid = 1; t.src.pad = 0;
    }
    ISAAC_HD static FinishFragment fromRecord(const isaac_ext_fragment_t &f) { FinishFragment t; t.f = f; t.alignmentScore = ~0u; rememberSource(t); return t; }
    ISAAC_HD FinishFragment fromRescue(const uint32_t index) const
    {
        FinishFragment t = fromRecord(v.rescueFragments[index]);
        t.f.cigarOffset = (FINISH_POOL_RESCUE << FINISH_POOL_SHIFT) | t.f.cigarOffset;
        rememberSource(t);
        return t;
    }

    /// isVeryBadAlignment (TemplateBuilder.cpp:52-58); Cigar::getMappedLength (Cigar.hh:137-153)
    ISAAC_HD bool veryBad(const isaac_ext_fragment_t &f, const uint32_t *cigar) const
    {
        unsigned mapped = 0;
        for (unsigned k = 0; k < f.cigarLength; ++k) if ((cigar[k] & 0xFu) == ISAAC_EXT_CIGAR_ALIGN) mapped += cigar[k] >> 4;
        return f.matchesInARow < 32 && (f.mismatchCount > mapped / 8 || f.logProbability < glibc_math::mul(v.logMismatchQ40 / 4, double(mapped)));
    }

    /// TemplateLengthStatistics::checkModel == Nominal (TemplateLengthStatistics.hh:104-176, .cpp:67-77)
    ISAAC_HD bool nominal(const isaac_ext_fragment_t &a, const isaac_ext_fragment_t &b) const
    {
        if (a.contigId != b.contigId) return false;
        const unsigned model = (a.position <= b.position ? 0u : 4u) | (a.reverse ? 2u : 0u) | (b.reverse ? 1u : 0u);
        if (model != v.bestModel[0] && model != v.bestModel[1]) return false;
        const long oa = long(planObservedLength(a)), ob = long(planObservedLength(b));
        long length;
        if (a.position < b.position) { length = b.position + ob - a.position; if (length < oa) length = oa; }
        else { length = a.position + oa - b.position; if (length < ob) length = ob; }
        return !((unsigned long)length > (unsigned long)v.tlsMax) && !((unsigned long)length < (unsigned long)v.tlsMin);
    }

    /// updateMappingScore (:233-285)
    ISAAC_HD bool updateMappingScore(FinishFragment &fragment, const int listFragment, const isaac_ext_fragment_t *list, const int count, const bool forceWellAnchored) const
    {
        if (forceWellAnchored || planWellAnchored(fragment.f))
        {
            double neighborProbability = v.rogRead[list[listFragment].readIndex];
            for (int i = 0; i < count; ++i)
                if (listFragment != i) neighborProbability += glibc_math::exp(list[i].logProbability);
            fragment.alignmentScore = toUnsignedX86(mappingScoreOf(neighborProbability, neighborProbability + glibc_math::exp(list[listFragment].logProbability)));
            return true;
        }
        fragment.alignmentScore = 0;
        return false;
    }

    ISAAC_HD bool flagDodgyTemplate(FinishFragment &orphan, FinishFragment &shadow)                  // :467-493
    {
        if (-1 == v.dodgyAlignmentScore) { setNoMatch(orphan); setNoMatch(shadow); bamAlignmentScore = ~0u; return false; }
        orphan.alignmentScore = ~0u; shadow.alignmentScore = ~0u; bamAlignmentScore = ~0u;
        return true;
    }
    ISAAC_HD bool flagDodgyTemplate(FinishFragment &orphan)                                          // :1010-1033
    {
        if (-1 == v.dodgyAlignmentScore) { setNoMatch(orphan); bamAlignmentScore = ~0u; return false; }
        orphan.alignmentScore = ~0u; bamAlignmentScore = ~0u;
        return true;
    }

    /// the answer to the next rescueShadow call of this cluster: [begin, end) in v.rescueFragments
    ISAAC_HD bool nextAnswer(uint32_t &begin, uint32_t &end)
    {
        const uint64_t i = nextRequest++;
        begin = uint32_t(v.requestFragmentBegin[i]); end = uint32_t(v.requestFragmentBegin[i + 1]);
        return v.rescued[i] != 0;
    }

    /// the bookkeeping both rescue loops share (:541-585, :778-815): is (orphan, best rescued shadow) the best pair so far?
    ISAAC_HD void considerRescuedPair(const unsigned orphanIndex, const int oi, const isaac_ext_fragment_t &orphan, const uint32_t rescuedIndex,
                                      unsigned *bestOrphanIndex)
    {
        const isaac_ext_fragment_t &bestRescued = v.rescueFragments[rescuedIndex];
        const double currentLp = orphan.logProbability + bestRescued.logProbability;
        const unsigned long templateScore = (unsigned long)(orphan.smithWatermanScore + bestRescued.smithWatermanScore);
        if (0 == resolved || templateScore < bestScore || (templateScore == bestScore && planLpLess(bestLp, currentLp)))
        {
            bestLp = currentLp; bestScore = templateScore;
            bestCount[orphanIndex] = 0; best[orphanIndex][bestCount[orphanIndex]++] = uint32_t(oi);
            bestShadowCount[orphanIndex] = 0; bestShadow[orphanIndex][bestShadowCount[orphanIndex]++] = rescuedIndex;
            if (bestOrphanIndex) *bestOrphanIndex = orphanIndex;
        }
        else if (templateScore == bestScore && planLpEquals(currentLp, bestLp))
        {
            best[orphanIndex][bestCount[orphanIndex]++] = uint32_t(oi);
            bestShadow[orphanIndex][bestShadowCount[orphanIndex]++] = rescuedIndex;
        }
        ++resolved;
    }

    /// TemplateBuilder::rescueShadow (:495-676): one of the two candidate lists is empty
    ISAAC_HD bool rescueShadowTemplate()
    {
        const unsigned orphanIndex = n[0] ? 0u : 1u, shadowIndex = (orphanIndex + 1) % 2;
        const isaac_ext_fragment_t *orphans = frags[orphanIndex];
        const int count = n[orphanIndex];
        const int bestOrphan = planBestFragment(PlanView{nullptr, nullptr, nullptr, nullptr, 2u, v.tlsMax, {v.bestModel[0], v.bestModel[1]}, v.scatterRepeats}, orphans, count, clusterId);
        bestCount[orphanIndex] = 0; best[orphanIndex][bestCount[orphanIndex]++] = uint32_t(bestOrphan);
        allShadowCount[orphanIndex] = 0;
        for (int oi = 0; oi < count; ++oi)
        {
            const isaac_ext_fragment_t &orphan = orphans[oi];
            uint32_t begin = 0, end = 0;
            if (!planLpLess(orphan.logProbability + 100.0, orphans[bestOrphan].logProbability))      // ORPHAN_LOG_PROBABILITY_SLACK_
            {
                if (nextAnswer(begin, end) && !veryBad(v.rescueFragments[begin], v.cigarPools[FINISH_POOL_RESCUE] + v.rescueFragments[begin].cigarOffset))
                    considerRescuedPair(orphanIndex, oi, orphan, begin, nullptr);
            }
            for (uint32_t s = begin; s < end; ++s)
            {
                allShadow[orphanIndex][allShadowCount[orphanIndex]++] = finishShadowProbability(v.rescueFragments[s]);
                totalProbability += glibc_math::exp(orphan.logProbability + v.rescueFragments[s].logProbability);
            }
        }
        const double totalShadowProbability = 0 < resolved ? finishSumUniqueShadows(allShadow[orphanIndex], allShadowCount[orphanIndex]) : 0.0;

        bool ret = true;
        FinishFragment &orphan = bam[orphanIndex];
        if (0 < resolved)
        {
            const unsigned repeatIndex = v.scatterRepeats ? clusterId % bestCount[orphanIndex] : 0u;
            orphan = fromRecord(orphans[best[orphanIndex][repeatIndex]]);
            FinishFragment shadow = fromRescue(bestShadow[orphanIndex][repeatIndex]);
            const bool assumeWellAnchored = updateMappingScore(orphan, int(best[orphanIndex][repeatIndex]), orphans, count,
                                                               0 == unsigned(orphan.f.editDistance) + shadow.f.editDistance);
            if (assumeWellAnchored)
            {
                const double shadowRog = v.rogRead[shadow.f.readIndex];
                const double otherShadows = (totalShadowProbability - glibc_math::exp(shadow.f.logProbability)) + shadowRog;
                shadow.alignmentScore = toUnsignedX86(mappingScoreOf(otherShadows, totalShadowProbability + shadowRog));
                const double otherPairs = (totalProbability - glibc_math::exp(bestLp)) + v.rogAll;
                bamAlignmentScore = toUnsignedX86(mappingScoreOf(otherPairs, totalProbability + v.rogAll));
                if (!orphan.alignmentScore || !planWellAnchored(orphan.f))
                {
                    if (bamAlignmentScore > FINISH_DODGY_BUT_CLEAN) bamAlignmentScore = FINISH_DODGY_BUT_CLEAN;
                    if (shadow.alignmentScore > FINISH_DODGY_BUT_CLEAN) shadow.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                    if (orphan.alignmentScore > FINISH_DODGY_BUT_CLEAN) orphan.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                }
            }
            else ret = flagDodgyTemplate(orphan, shadow);
            bam[shadowIndex] = shadow;
            bamProperPair = nominal(orphan.f, shadow.f);
        }
        else
        {
            orphan = fromRecord(orphans[bestOrphan]);
            FinishFragment &shadow = bam[shadowIndex];
            if (veryBad(orphan.f, cigarOf(orphan.f))) { setNoMatch(orphan); setNoMatch(shadow); ret = false; }
            else
            {
                shadow.f.contigId = orphan.f.contigId; shadow.f.position = orphan.f.position; shadow.f.readIndex = uint8_t(shadowIndex);
                shadow.alignmentScore = 0; shadow.f.cigarLength = 0;
                if (!updateMappingScore(orphan, bestOrphan, orphans, count, 0 == orphan.f.editDistance)) ret = flagDodgyTemplate(orphan, shadow);
                else
                {
                    if (!planWellAnchored(orphan.f) && orphan.alignmentScore > FINISH_DODGY_BUT_CLEAN) orphan.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                    bamAlignmentScore = 0;
                }
            }
        }
        return ret;
    }

    /// buildDisjoinedTemplate + scoreDisjoinedTemplate (:716-1008); known = the best pair among the seed candidates, knownTotal /
    /// knownLp its probability sums (unused here: the rescued pair gets its own)
    ISAAC_HD bool buildDisjoinedTemplate(const PlanBestPair &known)
    {
        const PlanView pv = {nullptr, nullptr, nullptr, nullptr, 2u, v.tlsMax, {v.bestModel[0], v.bestModel[1]}, v.scatterRepeats};
        const int bestDisjoined[2] = {planBestFragment(pv, frags[0], n[0], clusterId), planBestFragment(pv, frags[1], n[1], clusterId)};
        unsigned bestOrphanIndex = 0;
        // bestOrphans.init(bestDisjoinedFragments[0], [1]) (:730-731)
        bestLp = -DBL_MAX; bestScore = ~0ul; resolved = 0; totalProbability = 0.0;
        bestCount[0] = bestCount[1] = 0;
        best[0][bestCount[0]++] = uint32_t(bestDisjoined[0]); best[1][bestCount[1]++] = uint32_t(bestDisjoined[1]);
        allPairCount = 0;
        for (unsigned orphanIndex = 0; 2 > orphanIndex; ++orphanIndex)
        {
            allShadowCount[orphanIndex] = 0; bestShadowCount[orphanIndex] = 0;
            const isaac_ext_fragment_t *orphans = frags[orphanIndex];
            for (int oi = 0; oi < n[orphanIndex]; ++oi)
            {
                const isaac_ext_fragment_t &orphan = orphans[oi];
                const bool skip = known.resolved ? unsigned(orphan.editDistance) > known.editDistance + 3u
                                                 : planLpLess(orphan.logProbability + 100.0, orphans[bestDisjoined[orphanIndex]].logProbability);
                uint32_t begin = 0, end = 0;
                if (!skip && nextAnswer(begin, end))
                {
                    const isaac_ext_fragment_t &bestRescued = v.rescueFragments[begin];
                    const unsigned rescuedEditDistance = unsigned(orphan.editDistance) + bestRescued.editDistance;
                    if (veryBad(bestRescued, v.cigarPools[FINISH_POOL_RESCUE] + bestRescued.cigarOffset)) {}
                    else if (!known.resolved || known.editDistance + 3u >= rescuedEditDistance)
                        considerRescuedPair(orphanIndex, oi, orphan, begin, &bestOrphanIndex);
                }
                for (uint32_t s = begin; s < end; ++s)
                {
                    FinishPairProbability p;
                    if (0 == orphanIndex) { p.r1 = finishShadowProbability(orphan); p.r2 = finishShadowProbability(v.rescueFragments[s]); }
                    else { p.r1 = finishShadowProbability(v.rescueFragments[s]); p.r2 = finishShadowProbability(orphan); }
                    allPairs[allPairCount++] = p;
                    allShadow[orphanIndex][allShadowCount[orphanIndex]++] = finishShadowProbability(v.rescueFragments[s]);
                }
            }
        }
        const unsigned bestShadowIndex = (bestOrphanIndex + 1) % 2;
        double totalShadowProbability = 0.0, totalOrphanProbability = 0.0;
        if (0 < resolved)
        {
            for (int i = 0; i < n[bestShadowIndex]; ++i) allShadow[bestOrphanIndex][allShadowCount[bestOrphanIndex]++] = finishShadowProbability(frags[bestShadowIndex][i]);
            totalShadowProbability = finishSumUniqueShadows(allShadow[bestOrphanIndex], allShadowCount[bestOrphanIndex]);
            for (int i = 0; i < n[bestOrphanIndex]; ++i) allShadow[bestShadowIndex][allShadowCount[bestShadowIndex]++] = finishShadowProbability(frags[bestOrphanIndex][i]);
            totalOrphanProbability = finishSumUniqueShadows(allShadow[bestShadowIndex], allShadowCount[bestShadowIndex]);
            totalProbability += finishSumUniquePairs(allPairs, allPairCount);
        }

        // scoreDisjoinedTemplate (:868-1008)
        bool ret = true;
        if (0 < resolved)
        {
            const unsigned repeatIndex = v.scatterRepeats ? clusterId % bestCount[bestOrphanIndex] : 0u;
            const isaac_ext_fragment_t &bestOrphan = frags[bestOrphanIndex][best[bestOrphanIndex][repeatIndex]];
            FinishFragment shadow = fromRescue(bestShadow[bestOrphanIndex][repeatIndex]);
            const unsigned orphanRead = bestOrphan.readIndex, shadowRead = shadow.f.readIndex;
            bool rediscovered = !repeatIndex && known.resolved;
            if (rediscovered)
            {
                const isaac_ext_fragment_t &ko = frags[orphanRead][known.picked[orphanRead]], &ks = frags[shadowRead][known.picked[shadowRead]];
                rediscovered = ko.position == bestOrphan.position && ko.contigId == bestOrphan.contigId && ko.reverse == bestOrphan.reverse &&
                               ko.observedLength == bestOrphan.observedLength &&
                               ks.position == shadow.f.position && ks.contigId == shadow.f.contigId && ks.reverse == shadow.f.reverse &&
                               ks.observedLength == shadow.f.observedLength;
            }
            FinishFragment &orphan = bam[orphanRead];
            orphan = fromRecord(bestOrphan);
            const bool shadowWellAnchored = rediscovered && planWellAnchored(frags[shadowRead][known.picked[shadowRead]]);
            const bool assumeWellAnchored = updateMappingScore(orphan, int(best[orphanRead][repeatIndex]), frags[orphanRead], n[orphanRead],
                                                               0 == unsigned(orphan.f.editDistance) + shadow.f.editDistance || shadowWellAnchored);
            bamProperPair = nominal(orphan.f, shadow.f);
            if (assumeWellAnchored)
            {
                const double shadowRog = v.rogRead[shadowRead];
                const double otherShadows = (totalShadowProbability - glibc_math::exp(shadow.f.logProbability)) + shadowRog;
                shadow.alignmentScore = toUnsignedX86(mappingScoreOf(otherShadows, totalShadowProbability + shadowRog));
                const double orphanRog = v.rogRead[orphanRead];
                const double otherOrphans = (totalOrphanProbability - glibc_math::exp(bestOrphan.logProbability)) + orphanRog;
                orphan.alignmentScore = toUnsignedX86(mappingScoreOf(otherOrphans, totalOrphanProbability + orphanRog));
                const double otherPairs = (totalProbability - glibc_math::exp(bestLp)) + v.rogAll;
                bamAlignmentScore = toUnsignedX86(mappingScoreOf(otherPairs, totalProbability + v.rogAll));
                if ((!orphan.alignmentScore || !planWellAnchored(orphan.f)) && (!shadow.alignmentScore || !shadowWellAnchored))
                {
                    if (bamAlignmentScore > FINISH_DODGY_BUT_CLEAN) bamAlignmentScore = FINISH_DODGY_BUT_CLEAN;
                    if (shadow.alignmentScore > FINISH_DODGY_BUT_CLEAN) shadow.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                    if (orphan.alignmentScore > FINISH_DODGY_BUT_CLEAN) orphan.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                }
                bam[shadowRead] = shadow;
            }
            else
            {
                ret = flagDodgyTemplate(orphan, shadow);
                bam[shadowRead] = shadow;
            }
        }
        else if (known.resolved)
        {
            ret = flagDodgyTemplate(bam[0], bam[1]);
        }
        else
        {
            FinishFragment &read1 = bam[0], &read2 = bam[1];
            read1 = fromRecord(frags[0][bestDisjoined[0]]);
            read2 = fromRecord(frags[1][bestDisjoined[1]]);
            bamAlignmentScore = 0; bamProperPair = false;
            const bool r1 = updateMappingScore(read1, bestDisjoined[0], frags[0], n[0], 0 == read1.f.editDistance);
            const bool r2 = updateMappingScore(read2, bestDisjoined[1], frags[1], n[1], 0 == read2.f.editDistance);
            if (!r1 && !r2) ret = flagDodgyTemplate(read1, read2);
            else
            {
                if (!planWellAnchored(read1.f) && read1.alignmentScore > FINISH_DODGY_BUT_CLEAN) read1.alignmentScore = FINISH_DODGY_BUT_CLEAN;
                if (!planWellAnchored(read2.f) && read2.alignmentScore > FINISH_DODGY_BUT_CLEAN) read2.alignmentScore = FINISH_DODGY_BUT_CLEAN;
            }
        }
        return ret;
    }

    /// pickBestPair (:1060-1086) with locateBestPair (:287-391) and buildPairedEndTemplate (:398-465)
    ISAAC_HD bool pickBestPair()
    {
        const PlanView pv = {nullptr, nullptr, nullptr, nullptr, 2u, v.tlsMax, {v.bestModel[0], v.bestModel[1]}, v.scatterRepeats};
        const PlanBestPair known = planLocateBestPair(pv, frags[0], n[0], frags[1], n[1], clusterId);
        if (known.resolved)
        {
            // the sums locateBestPair keeps next to the best pair: total probability in walk order, the best pair's own sum
            double total = 0.0;
            planForEachPair(pv, frags[0], n[0], frags[1], n[1], [&](const int a, const int b) {
                total += glibc_math::exp(frags[0][a].logProbability + frags[1][b].logProbability);
            });
            const double knownLp = frags[0][known.first[0]].logProbability + frags[1][known.first[1]].logProbability;
            // buildPairedEndTemplate
            FinishFragment &read1 = bam[0], &read2 = bam[1];
            read1 = fromRecord(frags[0][known.picked[0]]);
            read2 = fromRecord(frags[1][known.picked[1]]);
            const bool r1 = updateMappingScore(read1, known.picked[0], frags[0], n[0], planWellAnchored(read2.f));
            const bool r2 = updateMappingScore(read2, known.picked[1], frags[1], n[1], planWellAnchored(read1.f));
            bamProperPair = nominal(read1.f, read2.f);
            bool paired = false;
            if (r1 || r2)
            {
                const double otherPairs = (total - glibc_math::exp(knownLp)) + v.rogAll;
                bamAlignmentScore = toUnsignedX86(mappingScoreOf(otherPairs, total + v.rogAll));
                paired = r1 && r2 && !read1.f.repeatSeedsCount && !read2.f.repeatSeedsCount;
            }
            else bamAlignmentScore = ~0u;
            if (paired && !known.editDistance) return true;
        }
        return buildDisjoinedTemplate(known);
    }

    /// pickBestFragment (:1035-1058): single-ended
    ISAAC_HD bool pickBestFragment()
    {
        if (!n[0]) return false;
        const PlanView pv = {nullptr, nullptr, nullptr, nullptr, 1u, v.tlsMax, {v.bestModel[0], v.bestModel[1]}, v.scatterRepeats};
        const int bestFragment = planBestFragment(pv, frags[0], n[0], clusterId);
        bam[0] = fromRecord(frags[0][bestFragment]);
        if (!updateMappingScore(bam[0], bestFragment, frags[0], n[0], false)) return flagDodgyTemplate(bam[0]);
        return true;
    }

    /// BamTemplate::filterLowQualityFragments (BamTemplate.cpp:46-72)
    ISAAC_HD bool filterLowQualityFragments(const unsigned threshold)
    {
        bool ret = false;
        unsigned alignmentScore = 0;
        for (unsigned i = 0; i < v.readCount; ++i)
        {
            FinishFragment &fragment = bam[i];
            if (threshold > fragment.alignmentScore)
            {
                fragment.f.cigarLength = 0; fragment.f.cigarOffset = 0; fragment.alignmentScore = 0;
                const FinishFragment &mate = bam[(i + 1) % v.readCount];
                fragment.f.position = mate.f.position; fragment.f.contigId = mate.f.contigId;
            }
            else if (finishIsAligned(fragment.f)) ret = true;
            alignmentScore += fragment.alignmentScore;
        }
        bamAlignmentScore = alignmentScore;
        return ret;
    }

    ISAAC_HD bool run()
    {
        for (unsigned r = 0; r < 2; ++r) bam[r] = unaligned(clusterId * v.readCount + (r < v.readCount - 1 ? r : v.readCount - 1), r);   // BamTemplate::initialize
        bamAlignmentScore = 0; bamProperPair = false;
        if (!v.built[clusterId]) return false;
        bool ret;
        if (2 == v.readCount)
        {
            if (n[0] && n[1]) ret = pickBestPair();
            else if (n[0] || n[1]) ret = rescueShadowTemplate();
            else ret = false;
        }
        else ret = pickBestFragment();
        if (ret && ~0u != bamAlignmentScore)                                                       // :112-124
        {
            if (!bamProperPair) ret = filterLowQualityFragments(v.mapqThreshold);
            else if (v.mapqThreshold > bamAlignmentScore) { filterLowQualityFragments(~0u); ret = false; }
        }
        return ret;
    }
};

/// the BamTemplate of one cluster as the flat records of the result; fragment.cigarOffset stays pool << 30 | word index, the
/// caller gathers the words (gatherTemplateCigars below / gatherTemplateCigarsKernel)
ISAAC_HD inline void finishCluster(const FinishView &v, const uint32_t cluster, unsigned char *scratch, const uint64_t shadows, const uint64_t candidates,
                                   isaac_ext_template_t &o, isaac_ext_fragment_t *fragments, FinishSource *sources = nullptr)
{
    FinishWorker w(v, cluster, scratch, shadows, candidates);
    const bool ok = w.run();
    o.alignmentScore = w.bamAlignmentScore; o.properPair = w.bamProperPair; o.built = ok; o.hadFragments = v.built[cluster]; o.pad = 0;
    o.fragmentAlignmentScore[0] = o.fragmentAlignmentScore[1] = 0;
    for (unsigned r = 0; r < v.readCount; ++r)
    {
        isaac_ext_fragment_t f = w.bam[r].f;
        f.readId = cluster * v.readCount + r;
        o.fragmentAlignmentScore[r] = w.bam[r].alignmentScore;
        fragments[r] = f;
        if (sources) sources[r] = w.bam[r].src;
    }
}

/// the rest-of-genome correction of a run (RestOfGenomeCorrection.hh:45-86, Quality.hh:87-91: the genome length passes through
/// 'unsigned'); host only (the host's libm like the reference's), once per call
inline void finishRestOfGenome(FinishView &v, const uint64_t *contigLength, const uint32_t contigCount, const uint32_t *readLength)
{
    uint64_t genomeLength = 0;
    for (uint32_t c = 0; c < contigCount; ++c) genomeLength += contigLength[c];
    auto correction = [&](unsigned length) {
        const double c = __builtin_exp(__builtin_log(2.0) + __builtin_log(double(unsigned(genomeLength))) - (__builtin_log(4.0) * double(length)));
        return c > DBL_MIN ? c : DBL_MIN;
    };
    unsigned total = 0;
    v.rogRead[0] = v.rogRead[1] = 0.0;
    for (unsigned r = 0; r < v.readCount; ++r) { v.rogRead[r] = correction(readLength[r]); total += readLength[r]; }
    v.rogAll = correction(total);
}

} // namespace isaac_b200
