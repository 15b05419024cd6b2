// alignment::TemplateBuilder per cluster on the host: pair selection, the decision which orphans get a
// ShadowAligner::rescueShadow call, mapping scores.  Restated line by line from lib/alignment/TemplateBuilder.cpp (cited at every
// function) over the flat results of the two batch calls; see isaac_ext_templates.cuh for how the plan / finish passes use it.
// Plain host C++ (no CUDA call): tests/cpp/test_template_worker.cpp drives it on the CPU between the checker's build and rescue
// results and compares the templates with the reference's own TemplateBuilder (tests/test_template_worker.py).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../isaac_aligner_b200/csrc/host_pipeline.cuh"

namespace
{

const uint32_t NO_MATCH_CONTIG = 0x7FFFFFu;                    // ReferencePosition::MAX_CONTIG_ID (ReferencePosition.hh:177)
const unsigned TRACKED_REPEATS_MAX_ONE_READ = 1000;            // TemplateBuilder.hh:145
const unsigned SKIP_ORPHAN_EDIT_DISTANCE = 3;                  // :146
const unsigned DODGY_BUT_CLEAN_ALIGNMENT_SCORE = 10;           // :149
const double ORPHAN_LOG_PROBABILITY_SLACK = 100.0;             // :141
const unsigned TEMPLATE_LENGTH_THRESHOLD = 50000;              // TemplateLengthStatistics.hh:220
const unsigned WEAK_SEED_LENGTH_T = 32;                        // Alignment.hh:36

/// FragmentMetadata as TemplateBuilder sees it: the flat record + alignmentScore + where its CIGAR words live.
struct TFrag
{
    isaac_ext_fragment_t f;
    uint32_t alignmentScore;
    const uint32_t *cigar;

    bool isAligned() const { return f.cigarLength != 0; }                                        // FragmentMetadata.hh:248
    bool isNoMatch() const { return f.contigId == NO_MATCH_CONTIG; }                             // :260
    unsigned observedLength() const { return isAligned() ? f.observedLength : 0; }               // :85
    bool isWellAnchored() const                                                                  // :477-483
    {
        return f.uniqueSeedCount ||
            (f.nonUniqueSeedOffsetFirst != 0xFFFF && f.nonUniqueSeedOffsetSecond > f.nonUniqueSeedOffsetFirst &&
             unsigned(f.nonUniqueSeedOffsetSecond - f.nonUniqueSeedOffsetFirst) >= WEAK_SEED_LENGTH_T);
    }
    static uint64_t referencePosition(uint64_t contigId, uint64_t position, bool neighbors = false)   // ReferencePosition.hh:69-73
    {
        return ((((contigId + 1) << 40) | position) << 1) | uint64_t(neighbors);
    }
    static uint64_t noMatchPosition() { return (uint64_t(NO_MATCH_CONTIG) << 40) << 1; }          // :60-61
    uint64_t fStrandPosition() const { return !isNoMatch() ? referencePosition(f.contigId, uint64_t(f.position)) : noMatchPosition(); }   // FragmentMetadata.hh:90-95
    uint64_t rStrandPosition() const                                                             // :97-103
    {
        return !isNoMatch() ? referencePosition(f.contigId, uint64_t(std::max(f.position + long(f.observedLength), 1L) - 1)) : noMatchPosition();
    }
    void setUnaligned() { cigar = nullptr; f.cigarLength = 0; alignmentScore = -1U; }            // :252
    void setNoMatch() { setUnaligned(); f.contigId = NO_MATCH_CONTIG; f.position = 0; }          // :258-259
    unsigned mappedLength() const                                                                // Cigar.hh:137-153
    {
        unsigned ret = 0;
        for (unsigned k = 0; k < f.cigarLength; ++k) if ((cigar[k] & 0xFu) == ISAAC_EXT_CIGAR_ALIGN) ret += cigar[k] >> 4;
        return ret;
    }
    bool samePlace(const TFrag &that) const                                                      // operator== (:431-438)
    {
        return f.position == that.f.position && f.contigId == that.f.contigId && f.reverse == that.f.reverse &&
            f.observedLength == that.f.observedLength;
    }
    /// FragmentMetadata(cluster, cigarBuffer, readIndex) (:63-75)
    static TFrag unaligned(uint32_t readId, unsigned readIndex)
    {
        TFrag t;
        std::memset(&t.f, 0, sizeof(t.f));
        t.f.contigId = NO_MATCH_CONTIG; t.f.readId = readId; t.f.readIndex = uint8_t(readIndex);
        t.f.firstSeedIndex = -1; t.f.nonUniqueSeedOffsetFirst = 0xFFFF;
        t.alignmentScore = -1U; t.cigar = nullptr;
        return t;
    }
};

/// TemplateBuilder.cpp:52-58
inline bool isVeryBadAlignment(const TFrag &t, double logMismatchQ40)
{
    const unsigned mapped = t.mappedLength();
    return t.f.matchesInARow < 32 && (t.f.mismatchCount > mapped / 8 || t.f.logProbability < logMismatchQ40 / 4 * mapped);
}

/// TemplateBuilder::ShadowProbability (TemplateBuilder.hh:165-207)
struct ShadowProbability
{
    uint64_t pos; double logProbability; long observedLength;
    explicit ShadowProbability(const TFrag &s)
        : pos((s.fStrandPosition() & ~uint64_t(1)) | uint64_t(s.f.reverse != 0)), logProbability(s.f.logProbability), observedLength(s.observedLength()) {}
    bool operator<(const ShadowProbability &that) const
    {
        return pos < that.pos ||
            (pos == that.pos && (lpLess(logProbability, that.logProbability) ||
                                 (lpEquals(logProbability, that.logProbability) && observedLength < that.observedLength)));
    }
    bool operator==(const ShadowProbability &that) const
    {
        return pos == that.pos && lpEquals(logProbability, that.logProbability) && observedLength == that.observedLength;
    }
};

/// TemplateBuilder::PairProbability (TemplateBuilder.hh:213-246)
struct PairProbability
{
    ShadowProbability r1, r2;
    PairProbability(const TFrag &a, const TFrag &b) : r1(a), r2(b) {}
    double logProbability() const { return r1.logProbability + r2.logProbability; }
    bool operator<(const PairProbability &that) const
    {
        return r1.pos < that.r1.pos || (r1.pos == that.r1.pos &&
            (r2.pos < that.r2.pos || (r2.pos == that.r2.pos &&
                (lpLess(that.logProbability(), logProbability()) || (lpEquals(logProbability(), that.logProbability()) &&
                    (r1.observedLength < that.r1.observedLength || (r1.observedLength == that.r1.observedLength &&
                        r2.observedLength < that.r2.observedLength)))))));
    }
    bool operator==(const PairProbability &that) const
    {
        return r1.pos == that.r1.pos && r2.pos == that.r2.pos && lpEquals(logProbability(), that.logProbability()) &&
            r1.observedLength == that.r1.observedLength && r2.observedLength == that.r2.observedLength;
    }
};

inline double logProbabilityOf(const ShadowProbability &p) { return p.logProbability; }
inline double logProbabilityOf(const PairProbability &p) { return p.logProbability(); }

/// sumUniqueShadowProbabilities / sumUniquePairProbabilities (TemplateBuilder.cpp:694-714): std::sort, then std::unique_copy
/// into a summing output iterator.  libstdc++'s unique_copy for forward iterators compares every element with the last one
/// it KEPT (not with its predecessor), which matters for the tolerance-based operator==; exp() is summed in kept order.
template <class P> double sumUniqueProbabilities(std::vector<P> &v)
{
    double ret = 0.0;
    std::sort(v.begin(), v.end());
    size_t kept = 0;
    for (size_t i = 0; i < v.size(); ++i)
    {
        if (i && v[kept] == v[i]) continue;
        kept = i;
        ret += exp(logProbabilityOf(v[i]));
    }
    return ret;
}

/// TemplateLengthStatistics::alignmentModel / getLength / matchModel / checkModel (TemplateLengthStatistics.hh:104-176, .cpp:67-77)
struct TemplateModel
{
    unsigned min, max, best[2];
    explicit TemplateModel(const isaac_ext_tls_t &t) : min(t.min), max(t.max) { best[0] = t.bestModel[0]; best[1] = t.bestModel[1]; }
    static unsigned alignmentModel(const TFrag &a, const TFrag &b)
    {
        if (a.f.contigId != b.f.contigId) return 8;
        return (a.f.position <= b.f.position ? 0u : 4u) | (a.f.reverse ? 2u : 0u) | (b.f.reverse ? 1u : 0u);
    }
    static unsigned long length(const TFrag &a, const TFrag &b)
    {
        if (a.f.position < b.f.position) return std::max<long>(b.f.position + long(b.observedLength()) - a.f.position, a.observedLength());
        return std::max<long>(a.f.position + long(a.observedLength()) - b.f.position, b.observedLength());
    }
    bool matchModel(const TFrag &a, const TFrag &b) const
    {
        const unsigned long len = length(a, b);
        const unsigned model = alignmentModel(a, b);
        return len <= max + TEMPLATE_LENGTH_THRESHOLD && (model == best[0] || model == best[1]);
    }
    bool nominal(const TFrag &a, const TFrag &b) const                                            // Nominal == checkModel(a, b)
    {
        if (a.f.contigId != b.f.contigId) return false;
        const unsigned model = alignmentModel(a, b);
        if (model != best[0] && model != best[1]) return false;
        const unsigned long len = length(a, b);
        return !(len > max) && !(len < min);
    }
};

/// TemplateBuilder::BestPairInfo (TemplateBuilder.hh:256-312); fragments are indices into the read's candidate list
struct BestPairInfo
{
    std::vector<int> best[2];
    double bestTemplateLogProbability; unsigned long bestTemplateScore; unsigned resolvedTemplateCount, bestPairEditDistance;
    double totalTemplateProbability;
    void clear()
    {
        bestTemplateLogProbability = -DBL_MAX; bestTemplateScore = -1UL; resolvedTemplateCount = 0; bestPairEditDistance = 0;
        totalTemplateProbability = 0.0; best[0].clear(); best[1].clear();
    }
    void init(int a, int b) { clear(); best[0].push_back(a); best[1].push_back(b); }
};

/// one rescueShadow call as the per-cluster code sees it
struct RescueAnswer
{
    bool rescued = false;
    const isaac_ext_fragment_t *begin = nullptr, *end = nullptr;
    const uint32_t *cigars = nullptr;
};

struct TemplateContext
{
    TemplateModel model;
    bool scatterRepeats; int dodgyAlignmentScore; unsigned mapqThreshold;
    double rogRead[2], rogAll, logMismatchQ40;
    unsigned readCount;
};

/// The per-thread TemplateBuilder.  run() = buildTemplate(…, mapqThreshold) of one cluster (TemplateBuilder.cpp:97-175).
struct TemplateWorker
{
    const TemplateContext &cx;
    // inputs of the current cluster
    std::vector<TFrag> frags[2];
    uint32_t clusterId = 0;
    // rescue plumbing
    bool planning = true;
    std::vector<isaac_ext_rescue_request_t> *requests = nullptr;       // plan: appended to
    const isaac_ext_rescue_result_t *rescueResult = nullptr;          // finish: answers, consumed from nextRequest on
    uint64_t nextRequest = 0;
    // BamTemplate
    TFrag bam[2]; uint32_t bamAlignmentScore = 0; bool bamProperPair = false;
    // scratch
    std::vector<TFrag> shadowList, bestOrphanShadows[2];
    std::vector<ShadowProbability> allShadowProbabilities[2];
    std::vector<PairProbability> allPairProbabilities;
    BestPairInfo bestCombinationPairInfo, bestRescuedPair;
    std::vector<int> bestFragmentsScratch;
    // CIGARs of rescued shadows kept by the template (cloneWithCigar, :678-687)
    std::vector<uint32_t> ownCigars;

    explicit TemplateWorker(const TemplateContext &c) : cx(c) {}

    RescueAnswer rescueShadow(const TFrag &orphan, long bestTemplateLength)
    {
        RescueAnswer a;
        if (planning)
        {
            isaac_ext_rescue_request_t q;
            q.orphanPosition = orphan.f.position; q.bestTemplateLength = bestTemplateLength; q.orphanReadId = orphan.f.readId;
            q.orphanContigStrand = (orphan.f.contigId << 1) | (orphan.f.reverse ? 1u : 0u);
            q.orphanObservedLength = orphan.f.observedLength; q.pad = 0;
            requests->push_back(q);
            return a;
        }
        const uint64_t i = nextRequest++;
        a.rescued = rescueResult->rescued[i] != 0;
        a.begin = rescueResult->fragments + rescueResult->requestFragmentBegin[i];
        a.end = rescueResult->fragments + rescueResult->requestFragmentBegin[i + 1];
        a.cigars = rescueResult->cigars;
        return a;
    }
    void loadShadowList(const RescueAnswer &a)
    {
        shadowList.clear();
        for (const isaac_ext_fragment_t *p = a.begin; p != a.end; ++p)
        {
            TFrag t; t.f = *p; t.alignmentScore = -1U; t.cigar = a.cigars + p->cigarOffset;
            shadowList.push_back(t);
        }
    }
    TFrag cloneWithCigar(const TFrag &right)                                                     // :678-687
    {
        // the clone must stay valid while this cluster is processed: keep the words in the worker (pointers are fixed up
        // against ownCigars when the template is written out, a vector may reallocate)
        TFrag ret = right;
        ret.f.cigarOffset = uint32_t(ownCigars.size());
        ownCigars.insert(ownCigars.end(), right.cigar, right.cigar + right.f.cigarLength);
        ret.cigar = nullptr;                       // marks "in ownCigars at f.cigarOffset"
        return ret;
    }
    const uint32_t *cigarOf(const TFrag &t) const { return t.cigar ? t.cigar : (t.f.cigarLength ? ownCigars.data() + t.f.cigarOffset : nullptr); }
    /// isVeryBadAlignment needs the CIGAR words whatever pool they are in
    bool veryBad(const TFrag &t) const { TFrag x = t; x.cigar = cigarOf(t); return isVeryBadAlignment(x, cx.logMismatchQ40); }

    int getBestFragment(const std::vector<TFrag> &list)                                          // :177-226
    {
        std::vector<int> &bestFragments = bestFragmentsScratch;
        bestFragments.clear();
        unsigned bestFragmentScore = -1U;
        double bestFragmentLogProbability = -DBL_MAX;
        for (int i = 0; i < int(list.size()); ++i)
        {
            const TFrag &t = list[i];
            if (bestFragmentScore > t.f.smithWatermanScore ||
                (bestFragmentScore == t.f.smithWatermanScore && lpLess(bestFragmentLogProbability, t.f.logProbability)))
            {
                bestFragmentScore = t.f.smithWatermanScore; bestFragmentLogProbability = t.f.logProbability;
                bestFragments.clear(); bestFragments.push_back(i);
            }
            else if (bestFragmentScore == t.f.smithWatermanScore && lpEquals(bestFragmentLogProbability, t.f.logProbability))
            {
                bestFragments.push_back(i);
            }
        }
        const unsigned repeatIndex = cx.scatterRepeats ? (clusterId % bestFragments.size()) : 0;
        return bestFragments[repeatIndex];
    }

    bool updateMappingScore(TFrag &fragment, int listFragment, const std::vector<TFrag> &list, bool forceWellAnchored) const   // :233-285
    {
        if (forceWellAnchored || fragment.isWellAnchored())
        {
            double neighborProbability = cx.rogRead[list[listFragment].f.readIndex];
            for (int i = 0; i < int(list.size()); ++i)
                if (listFragment != i) neighborProbability += exp(list[i].f.logProbability);
            fragment.alignmentScore = unsigned(floor(-10.0 * log10(neighborProbability / (neighborProbability + exp(list[listFragment].f.logProbability)))));
            return true;
        }
        fragment.alignmentScore = 0;
        return false;
    }

    void locateBestPair(BestPairInfo &ret)                                                       // :287-391
    {
        const std::vector<TFrag> &f0 = frags[0], &f1 = frags[1];
        ret.init(0, 0);
        int contigBegin[2] = {0, 0}, contigEnd[2] = {0, 0};
        const int size[2] = {int(f0.size()), int(f1.size())};
        const std::vector<TFrag> *lists[2] = {&f0, &f1};
        while (size[0] != contigBegin[0] && size[1] != contigBegin[1])
        {
            for (int i = 0; i < 2; ++i)
            {
                contigEnd[i] = contigBegin[i] + 1;
                while (size[i] != contigEnd[i] && (*lists[i])[contigEnd[i]].f.contigId == (*lists[i])[contigBegin[i]].f.contigId) ++contigEnd[i];
            }
            if (f0[contigBegin[0]].f.contigId == f1[contigBegin[1]].f.contigId)
            {
                for (int a = contigBegin[0]; a != contigEnd[0]; ++a)
                {
                    for (int b = contigBegin[1]; b != contigEnd[1]; ++b)
                    {
                        if (!cx.model.matchModel(f0[a], f1[b])) continue;
                        const double currentLogProbability = f0[a].f.logProbability + f1[b].f.logProbability;
                        const double currentProbability = exp(currentLogProbability);
                        const unsigned long templateScore = (unsigned long)(f0[a].f.smithWatermanScore + f1[b].f.smithWatermanScore);
                        ret.totalTemplateProbability += currentProbability;
                        if (0 == ret.resolvedTemplateCount || ret.bestTemplateScore > templateScore ||
                            (templateScore == ret.bestTemplateScore && lpLess(ret.bestTemplateLogProbability, currentLogProbability)))
                        {
                            ret.best[0].clear(); ret.best[1].clear();
                            ret.best[0].push_back(a); ret.best[1].push_back(b);
                            ret.bestTemplateScore = templateScore; ret.bestTemplateLogProbability = currentLogProbability;
                        }
                        else if (templateScore == ret.bestTemplateScore && lpEquals(currentLogProbability, ret.bestTemplateLogProbability))
                        {
                            ret.best[0].push_back(a); ret.best[1].push_back(b);
                        }
                        ++ret.resolvedTemplateCount;
                    }
                }
                contigBegin[0] = contigEnd[0]; contigBegin[1] = contigEnd[1];
            }
            else
            {
                const int i = f0[contigBegin[0]].f.contigId < f1[contigBegin[1]].f.contigId ? 0 : 1;
                contigBegin[i] = contigEnd[i];
            }
        }
        if (ret.resolvedTemplateCount)
            ret.bestPairEditDistance = unsigned(f0[ret.best[0][0]].f.editDistance) + f1[ret.best[1][0]].f.editDistance;
    }

    long bestTemplateLength(const BestPairInfo &info) const                                      // TemplateBuilder.hh:289-300
    {
        if (!info.resolvedTemplateCount) return 0;
        const TFrag &a = frags[0][info.best[0][0]], &b = frags[1][info.best[1][0]];
        const uint64_t start = std::min(a.fStrandPosition(), b.fStrandPosition());
        const uint64_t end = std::max(a.rStrandPosition(), b.rStrandPosition());
        const uint64_t mask = (uint64_t(1) << 40) - 1;
        return long((end >> 1) & mask) - long((start >> 1) & mask);
    }

    bool buildPairedEndTemplate(BestPairInfo &info)                                              // :398-465
    {
        TFrag &read1 = bam[0], &read2 = bam[1];
        if (cx.scatterRepeats)
        {
            const unsigned repeatIndex = clusterId % info.best[0].size();
            std::swap(info.best[0][0], info.best[0][repeatIndex]);
            std::swap(info.best[1][0], info.best[1][repeatIndex]);
        }
        read1 = frags[0][info.best[0][0]];
        read2 = frags[1][info.best[1][0]];
        const bool r1WellAnchored = updateMappingScore(read1, info.best[0][0], frags[0], read2.isWellAnchored());
        const bool r2WellAnchored = updateMappingScore(read2, info.best[1][0], frags[1], read1.isWellAnchored());
        bamProperPair = cx.model.nominal(read1, read2);
        if (r1WellAnchored || r2WellAnchored)
        {
            const double otherPairsProbability = (info.totalTemplateProbability - exp(info.bestTemplateLogProbability)) + cx.rogAll;
            bamAlignmentScore = unsigned(floor(-10.0 * log10(otherPairsProbability / (info.totalTemplateProbability + cx.rogAll))));
            return r1WellAnchored && r2WellAnchored && !read1.f.repeatSeedsCount && !read2.f.repeatSeedsCount;
        }
        bamAlignmentScore = -1U;
        return false;
    }

    bool flagDodgyTemplate(TFrag &orphan, TFrag &shadow)                                         // :467-493
    {
        if (-1 == cx.dodgyAlignmentScore)
        {
            orphan.setNoMatch(); shadow.setNoMatch(); bamAlignmentScore = -1U;
            return false;
        }
        orphan.alignmentScore = -1U; shadow.alignmentScore = -1U; bamAlignmentScore = -1U;
        return true;
    }
    bool flagDodgyTemplate(TFrag &orphan)                                                        // :1010-1033
    {
        if (-1 == cx.dodgyAlignmentScore) { orphan.setNoMatch(); bamAlignmentScore = -1U; return false; }
        orphan.alignmentScore = -1U; bamAlignmentScore = -1U;
        return true;
    }

    bool rescueShadowTemplate()                                                                  // TemplateBuilder::rescueShadow, :495-676
    {
        const unsigned orphanIndex = frags[0].empty() ? 1 : 0;
        const unsigned shadowIndex = (orphanIndex + 1) % 2;
        const std::vector<TFrag> &orphans = frags[orphanIndex];
        const int bestOrphanIterator = getBestFragment(orphans);
        BestPairInfo &bestPair = bestRescuedPair;
        bestPair.clear();
        bestPair.best[orphanIndex].push_back(bestOrphanIterator);
        allShadowProbabilities[orphanIndex].clear();
        for (int oi = 0; oi < int(orphans.size()); ++oi)
        {
            const TFrag &orphan = orphans[oi];
            shadowList.clear();
            if (lpLess(orphan.f.logProbability + ORPHAN_LOG_PROBABILITY_SLACK, orphans[bestOrphanIterator].f.logProbability))
            {
                // orphan too bad to try rescuing shadows
            }
            else
            {
                const RescueAnswer answer = rescueShadow(orphan, 0);
                loadShadowList(answer);
                if (answer.rescued)
                {
                    const TFrag &bestRescued = shadowList.front();
                    const double currentTemplateLogProbability = orphan.f.logProbability + bestRescued.f.logProbability;
                    const unsigned long templateScore = (unsigned long)(orphan.f.smithWatermanScore + bestRescued.f.smithWatermanScore);
                    if (!veryBad(bestRescued))
                    {
                        if (0 == bestPair.resolvedTemplateCount || templateScore < bestPair.bestTemplateScore ||
                            (templateScore == bestPair.bestTemplateScore && lpLess(bestPair.bestTemplateLogProbability, currentTemplateLogProbability)))
                        {
                            bestPair.bestTemplateLogProbability = currentTemplateLogProbability;
                            bestPair.bestTemplateScore = templateScore;
                            bestPair.best[orphanIndex].clear();
                            bestPair.best[orphanIndex].push_back(oi);
                            bestOrphanShadows[orphanIndex].clear();
                            bestOrphanShadows[orphanIndex].push_back(cloneWithCigar(bestRescued));
                        }
                        else if (templateScore == bestPair.bestTemplateScore && lpEquals(currentTemplateLogProbability, bestPair.bestTemplateLogProbability))
                        {
                            bestPair.best[orphanIndex].push_back(oi);
                            bestOrphanShadows[orphanIndex].push_back(cloneWithCigar(bestRescued));
                        }
                        ++bestPair.resolvedTemplateCount;
                    }
                }
            }
            for (const TFrag &shadow : shadowList)
            {
                allShadowProbabilities[orphanIndex].push_back(ShadowProbability(shadow));
                bestPair.totalTemplateProbability += exp(orphan.f.logProbability + shadow.f.logProbability);
            }
        }
        const double totalShadowProbability = 0 < bestPair.resolvedTemplateCount ? sumUniqueProbabilities(allShadowProbabilities[orphanIndex]) : 0.0;

        bool ret = true;
        TFrag &orphan = bam[orphanIndex];
        if (0 < bestPair.resolvedTemplateCount)
        {
            const unsigned repeatIndex = cx.scatterRepeats ? clusterId % bestPair.best[orphanIndex].size() : 0;
            orphan = orphans[bestPair.best[orphanIndex][repeatIndex]];
            TFrag &bestShadow = bestOrphanShadows[orphanIndex][repeatIndex];
            const bool assumeWellAnchored = updateMappingScore(orphan, bestPair.best[orphanIndex][repeatIndex], orphans,
                                                               0 == unsigned(orphan.f.editDistance) + bestShadow.f.editDistance);
            if (assumeWellAnchored)
            {
                const double shadowRog = cx.rogRead[bestShadow.f.readIndex];
                const double otherShadowsProbability = (totalShadowProbability - exp(bestShadow.f.logProbability)) + shadowRog;
                bestShadow.alignmentScore = floor(-10.0 * log10(otherShadowsProbability / (totalShadowProbability + shadowRog)));
                const double otherPairsProbability = (bestPair.totalTemplateProbability - exp(bestPair.bestTemplateLogProbability)) + cx.rogAll;
                bamAlignmentScore = unsigned(floor(-10.0 * log10(otherPairsProbability / (bestPair.totalTemplateProbability + cx.rogAll))));
                if (!orphan.alignmentScore || !orphan.isWellAnchored())
                {
                    bamAlignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, bamAlignmentScore);
                    bestShadow.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, bestShadow.alignmentScore);
                    orphan.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, orphan.alignmentScore);
                }
            }
            else
            {
                ret = flagDodgyTemplate(orphan, bestShadow);
            }
            bam[shadowIndex] = bestShadow;
            bamProperPair = cx.model.nominal(orphan, bestShadow);
        }
        else
        {
            orphan = orphans[bestOrphanIterator];
            TFrag &shadow = bam[shadowIndex];
            if (veryBad(orphan))
            {
                orphan.setNoMatch(); shadow.setNoMatch();
                ret = false;
            }
            else
            {
                shadow.f.contigId = orphan.f.contigId; shadow.f.position = orphan.f.position; shadow.f.readIndex = uint8_t(shadowIndex);
                shadow.alignmentScore = 0; shadow.f.cigarLength = 0;
                if (!updateMappingScore(orphan, bestOrphanIterator, orphans, 0 == orphan.f.editDistance))
                {
                    ret = flagDodgyTemplate(orphan, shadow);
                }
                else
                {
                    if (!orphan.isWellAnchored()) orphan.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, orphan.alignmentScore);
                    bamAlignmentScore = 0;
                }
            }
        }
        return ret;
    }

    bool buildDisjoinedTemplate(const BestPairInfo &knownBestPair)                               // :716-866
    {
        const int bestDisjoinedFragments[2] = {getBestFragment(frags[0]), getBestFragment(frags[1])};
        const long knownBestTemplateLength = bestTemplateLength(knownBestPair);
        unsigned bestOrphanIndex = 0;
        BestPairInfo &bestOrphans = bestRescuedPair;
        bestOrphans.init(bestDisjoinedFragments[0], bestDisjoinedFragments[1]);
        allPairProbabilities.clear();
        for (unsigned orphanIndex = 0; 2 > orphanIndex; ++orphanIndex)
        {
            allShadowProbabilities[orphanIndex].clear();
            bestOrphanShadows[orphanIndex].clear();
            const std::vector<TFrag> &orphans = frags[orphanIndex];
            for (int oi = 0; oi < int(orphans.size()); ++oi)
            {
                const TFrag &orphan = orphans[oi];
                const bool skipThisOrphan = knownBestPair.resolvedTemplateCount ?
                    unsigned(orphan.f.editDistance) > (knownBestPair.bestPairEditDistance + SKIP_ORPHAN_EDIT_DISTANCE) :
                    lpLess(orphan.f.logProbability + ORPHAN_LOG_PROBABILITY_SLACK, orphans[bestDisjoinedFragments[orphanIndex]].f.logProbability);
                shadowList.clear();
                if (!skipThisOrphan)
                {
                    const RescueAnswer answer = rescueShadow(orphan, knownBestTemplateLength);
                    loadShadowList(answer);
                    if (answer.rescued)
                    {
                        const TFrag &bestRescued = shadowList.front();
                        const double currentTemplateLogProbability = orphan.f.logProbability + bestRescued.f.logProbability;
                        const unsigned rescuedEditDistance = unsigned(orphan.f.editDistance) + bestRescued.f.editDistance;
                        if (veryBad(bestRescued))
                        {
                            // rescued shadow too bad
                        }
                        else if (!knownBestPair.resolvedTemplateCount || (knownBestPair.bestPairEditDistance + SKIP_ORPHAN_EDIT_DISTANCE) >= rescuedEditDistance)
                        {
                            const unsigned long templateScore = (unsigned long)(orphan.f.smithWatermanScore + bestRescued.f.smithWatermanScore);
                            if (0 == bestOrphans.resolvedTemplateCount || templateScore < bestOrphans.bestTemplateScore ||
                                (templateScore == bestOrphans.bestTemplateScore && lpLess(bestOrphans.bestTemplateLogProbability, currentTemplateLogProbability)))
                            {
                                bestOrphans.bestTemplateLogProbability = currentTemplateLogProbability;
                                bestOrphans.bestTemplateScore = templateScore;
                                bestOrphans.best[orphanIndex].clear();
                                bestOrphans.best[orphanIndex].push_back(oi);
                                bestOrphanShadows[orphanIndex].clear();
                                bestOrphanShadows[orphanIndex].push_back(cloneWithCigar(bestRescued));
                                bestOrphanIndex = orphanIndex;
                            }
                            else if (templateScore == bestOrphans.bestTemplateScore && lpEquals(currentTemplateLogProbability, bestOrphans.bestTemplateLogProbability))
                            {
                                bestOrphans.best[orphanIndex].push_back(oi);
                                bestOrphanShadows[orphanIndex].push_back(cloneWithCigar(bestRescued));
                            }
                            ++bestOrphans.resolvedTemplateCount;
                        }
                    }
                }
                for (const TFrag &shadow : shadowList)
                {
                    allPairProbabilities.push_back(0 == orphanIndex ? PairProbability(orphan, shadow) : PairProbability(shadow, orphan));
                    allShadowProbabilities[orphanIndex].push_back(ShadowProbability(shadow));
                }
            }
        }
        const unsigned bestShadowIndex = (bestOrphanIndex + 1) % 2;
        double totalShadowProbability = 0.0, totalOrphanProbability = 0.0;
        if (0 < bestOrphans.resolvedTemplateCount)
        {
            for (const TFrag &shadow : frags[bestShadowIndex]) allShadowProbabilities[bestOrphanIndex].push_back(ShadowProbability(shadow));
            totalShadowProbability = sumUniqueProbabilities(allShadowProbabilities[bestOrphanIndex]);
            for (const TFrag &orphan : frags[bestOrphanIndex]) allShadowProbabilities[bestShadowIndex].push_back(ShadowProbability(orphan));
            totalOrphanProbability = sumUniqueProbabilities(allShadowProbabilities[bestShadowIndex]);
            bestOrphans.totalTemplateProbability += sumUniqueProbabilities(allPairProbabilities);
        }
        return scoreDisjoinedTemplate(bestOrphans, knownBestPair, bestOrphanIndex, totalShadowProbability, totalOrphanProbability, bestDisjoinedFragments);
    }

    bool scoreDisjoinedTemplate(const BestPairInfo &bestOrphans, const BestPairInfo &knownBestPair, const unsigned bestOrphanIndex,
                                const double totalShadowProbability, const double totalOrphanProbability,
                                const int (&bestDisjoinedFragments)[2])                          // :868-1008
    {
        bool ret = true;
        if (0 < bestOrphans.resolvedTemplateCount)
        {
            const unsigned repeatIndex = cx.scatterRepeats ? clusterId % bestOrphans.best[bestOrphanIndex].size() : 0;
            const TFrag &bestOrphan = frags[bestOrphanIndex][bestOrphans.best[bestOrphanIndex][repeatIndex]];
            TFrag &bestShadow = bestOrphanShadows[bestOrphanIndex][repeatIndex];
            const unsigned orphanRead = bestOrphan.f.readIndex, shadowRead = bestShadow.f.readIndex;
            const bool rediscovered = !repeatIndex && knownBestPair.resolvedTemplateCount &&
                frags[orphanRead][knownBestPair.best[orphanRead][0]].samePlace(bestOrphan) &&
                frags[shadowRead][knownBestPair.best[shadowRead][0]].samePlace(bestShadow);
            TFrag &orphan = bam[orphanRead];
            orphan = bestOrphan;
            const bool shadowWellAnchored = rediscovered && frags[shadowRead][knownBestPair.best[shadowRead][0]].isWellAnchored();
            const bool assumeWellAnchored = updateMappingScore(
                orphan, bestOrphans.best[orphanRead][repeatIndex], frags[orphanRead],
                0 == unsigned(orphan.f.editDistance) + bestShadow.f.editDistance || shadowWellAnchored);
            bamProperPair = cx.model.nominal(orphan, bestShadow);
            if (assumeWellAnchored)
            {
                const double shadowRog = cx.rogRead[shadowRead];
                const double otherShadowsProbability = (totalShadowProbability - exp(bestShadow.f.logProbability)) + shadowRog;
                bestShadow.alignmentScore = floor(-10.0 * log10(otherShadowsProbability / (totalShadowProbability + shadowRog)));
                const double orphanRog = cx.rogRead[orphanRead];
                const double otherOrphansProbability = (totalOrphanProbability - exp(bestOrphan.f.logProbability)) + orphanRog;
                orphan.alignmentScore = floor(-10.0 * log10(otherOrphansProbability / (totalOrphanProbability + orphanRog)));
                const double otherPairsProbability = (bestOrphans.totalTemplateProbability - exp(bestOrphans.bestTemplateLogProbability)) + cx.rogAll;
                bamAlignmentScore = unsigned(floor(-10.0 * log10(otherPairsProbability / (bestOrphans.totalTemplateProbability + cx.rogAll))));
                if ((!orphan.alignmentScore || !orphan.isWellAnchored()) && (!bestShadow.alignmentScore || !shadowWellAnchored))
                {
                    bamAlignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, bamAlignmentScore);
                    bestShadow.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, bestShadow.alignmentScore);
                    orphan.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, orphan.alignmentScore);
                }
                bam[shadowRead] = bestShadow;
            }
            else
            {
                ret = flagDodgyTemplate(orphan, bestShadow);
                bam[shadowRead] = bestShadow;
            }
        }
        else if (knownBestPair.resolvedTemplateCount)
        {
            ret = flagDodgyTemplate(bam[0], bam[1]);
        }
        else
        {
            TFrag &read1 = bam[0], &read2 = bam[1];
            read1 = frags[0][bestDisjoinedFragments[0]];
            read2 = frags[1][bestDisjoinedFragments[1]];
            bamAlignmentScore = 0;
            bamProperPair = false;
            const bool assumeR1WellAnchored = updateMappingScore(read1, bestDisjoinedFragments[0], frags[0], 0 == read1.f.editDistance);
            const bool assumeR2WellAnchored = updateMappingScore(read2, bestDisjoinedFragments[1], frags[1], 0 == read2.f.editDistance);
            if (!assumeR1WellAnchored && !assumeR2WellAnchored)
            {
                ret = flagDodgyTemplate(read1, read2);
            }
            else
            {
                if (!read1.isWellAnchored()) read1.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, read1.alignmentScore);
                if (!read2.isWellAnchored()) read2.alignmentScore = std::min(DODGY_BUT_CLEAN_ALIGNMENT_SCORE, read2.alignmentScore);
            }
        }
        return ret;
    }

    bool pickBestFragment()                                                                      // :1035-1058
    {
        if (frags[0].empty()) return false;
        const int best = getBestFragment(frags[0]);
        bam[0] = frags[0][best];
        if (!updateMappingScore(bam[0], best, frags[0], false)) return flagDodgyTemplate(bam[0]);
        return true;
    }

    bool pickBestPair()                                                                          // :1060-1086
    {
        locateBestPair(bestCombinationPairInfo);
        if (!bestCombinationPairInfo.resolvedTemplateCount || !buildPairedEndTemplate(bestCombinationPairInfo) ||
            bestCombinationPairInfo.bestPairEditDistance)
        {
            return buildDisjoinedTemplate(bestCombinationPairInfo);
        }
        return true;
    }

    /// BamTemplate::filterLowQualityFragments (BamTemplate.cpp:46-72)
    bool filterLowQualityFragments(unsigned threshold)
    {
        bool ret = false;
        unsigned alignmentScore = 0;
        for (unsigned i = 0; i < cx.readCount; ++i)
        {
            TFrag &fragment = bam[i];
            if (threshold > fragment.alignmentScore)
            {
                fragment.f.cigarLength = 0; fragment.f.cigarOffset = 0; fragment.alignmentScore = 0;
                const TFrag &mate = bam[(i + 1) % cx.readCount];
                fragment.f.position = mate.f.position; fragment.f.contigId = mate.f.contigId;
            }
            else if (fragment.isAligned())
            {
                ret = true;
            }
            alignmentScore += fragment.alignmentScore;
        }
        bamAlignmentScore = alignmentScore;
        return ret;
    }

    /// buildTemplate(…, mapqThreshold) over the loaded fragment lists (:97-175); the BamTemplate is in bam[] afterwards
    bool run()
    {
        ownCigars.clear();
        for (unsigned r = 0; r < 2; ++r) bam[r] = TFrag::unaligned(clusterId * cx.readCount + std::min(r, cx.readCount - 1), r);     // BamTemplate::initialize
        bamAlignmentScore = 0; bamProperPair = false;
        bool ret;
        if (2 == cx.readCount)
        {
            if (!frags[0].empty() && !frags[1].empty()) ret = pickBestPair();
            else if (!frags[0].empty() || !frags[1].empty()) ret = rescueShadowTemplate();
            else ret = false;
        }
        else
        {
            ret = pickBestFragment();
        }
        if (ret && -1U != bamAlignmentScore)                                                      // :112-124
        {
            if (!bamProperPair) ret = filterLowQualityFragments(cx.mapqThreshold);
            else if (cx.mapqThreshold > bamAlignmentScore) { filterLowQualityFragments(-1U); ret = false; }
        }
        return ret;
    }
};

/// TemplateContext of a run: the template length model, the options and the rest-of-genome correction
/// (RestOfGenomeCorrection.hh:45-86, Quality.hh:87-91: the genome length passes through 'unsigned')
inline TemplateContext makeTemplateContext(const isaac_ext_tls_t &tls, const isaac_ext_template_options_t &options,
                                           const std::vector<uint64_t> &contigLength, const uint32_t readCount,
                                           const uint32_t (&readLength)[2], const double logMismatchQ40)
{
    TemplateContext cx = {TemplateModel(tls), options.scatterRepeats != 0, options.dodgyAlignmentScore, options.mapqThreshold,
                          {0.0, 0.0}, 0.0, logMismatchQ40, readCount};
    uint64_t genomeLength = 0;
    for (uint64_t l : contigLength) genomeLength += l;
    auto correction = [&](unsigned length) {
        const double c = exp(log(2.0) + log(double(unsigned(genomeLength))) - (log(4.0) * double(length)));
        return std::max(c, DBL_MIN);
    };
    unsigned total = 0;
    for (unsigned r = 0; r < readCount; ++r) { cx.rogRead[r] = correction(readLength[r]); total += readLength[r]; }
    cx.rogAll = correction(total);
    return cx;
}

/// the candidate lists of cluster c from the flat result of isaac_ext_build_fragments (getFragments() of the reference's builder)
inline void loadClusterFragments(TemplateWorker &w, const isaac_ext_build_result_t &built, const bool clusterBuilt,
                                 const unsigned readCount, const uint32_t c)
{
    w.clusterId = c;
    for (unsigned r = 0; r < 2; ++r)
    {
        w.frags[r].clear();
        if (r >= readCount || !clusterBuilt) continue;
        for (uint64_t i = built.readFragmentBegin[size_t(c) * readCount + r]; i < built.readFragmentBegin[size_t(c) * readCount + r + 1]; ++i)
        {
            TFrag t; t.f = built.fragments[i]; t.alignmentScore = -1U; t.cigar = built.cigars + t.f.cigarOffset;
            w.frags[r].push_back(t);
        }
    }
}

/// the BamTemplate a finished worker holds, as the flat records of the result: template o, fragments[readCount], their CIGAR
/// words appended to 'pool' (cigarOffset relative to the pool)
inline void storeTemplate(const TemplateWorker &w, const bool ok, const bool hadFragments, const uint32_t c, const unsigned readCount,
                          isaac_ext_template_t &o, isaac_ext_fragment_t *fragments, std::vector<uint32_t> &pool)
{
    std::memset(&o, 0, sizeof(o));
    o.hadFragments = hadFragments;
    o.built = ok; o.alignmentScore = w.bamAlignmentScore; o.properPair = w.bamProperPair;
    for (unsigned r = 0; r < readCount; ++r)
    {
        const TFrag &src = w.bam[r];
        isaac_ext_fragment_t f = src.f;
        f.readId = c * readCount + r;
        o.fragmentAlignmentScore[r] = src.alignmentScore;
        const uint32_t *words = w.cigarOf(src);
        f.cigarOffset = uint32_t(pool.size());
        if (f.cigarLength && words) pool.insert(pool.end(), words, words + f.cigarLength);
        fragments[r] = f;
    }
}

/// a cluster without candidate fragments: BamTemplate::initialize (MatchSelector.cpp:300-314, 351-358)
inline void resetToUnaligned(TemplateWorker &w, const uint32_t c, const unsigned readCount)
{
    w.clusterId = c;
    w.frags[0].clear(); w.frags[1].clear(); w.ownCigars.clear();
    for (unsigned r = 0; r < 2; ++r) w.bam[r] = TFrag::unaligned(c * readCount + std::min(r, readCount - 1), r);
    w.bamAlignmentScore = 0; w.bamProperPair = false;
}

} // namespace
