// The plan pass of isaac_ext_build_templates for host AND device code: which ShadowAligner::rescueShadow calls the reference's
// TemplateBuilder makes for a cluster, in the order it makes them (TemplateBuilder.cpp:97-175 buildTemplate, :287-391
// locateBestPair, :398-465 buildPairedEndTemplate, :495-676 rescueShadow, :716-866 buildDisjoinedTemplate, :1060-1086 pickBestPair).
// Every decision on that path is free of libm -- ISAAC_LP_LESS / ISAAC_LP_EQUALS on the fragments' ordered FP64 sums, edit
// distances, seed anchoring, positions -- so a one-thread-per-cluster kernel behind device-resident build lists gives the
// requests template_worker.cuh records in its planning mode, bit for bit.  No std::vector: the lists of equally good fragments
// / pairs the reference keeps (of which only the entry picked by --scatter-repeats is ever read) are walked twice instead.
// tests/cpp/test_template_worker.cu checks on the CPU that both give the same request lists (tests/test_template_worker.py).
// On the GPU it is the body of planRequestsKernel (kernels_templates.cuh).
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include "../../include/isaac_ext.h"

#ifndef ISAAC_HD
#ifdef __CUDACC__
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif
#endif

namespace isaac_b200
{

/// the flat result of isaac_ext_build_fragments plus what the decisions read of the run
struct PlanView
{
    const isaac_ext_fragment_t *fragments;
    const uint32_t *listBegin, *listCount;  // the candidate list of (cluster, readIndex): fragments[listBegin[l] .. + listCount[l]), l = cluster * readCount + readIndex
    const uint8_t *built;                   // per cluster
    uint32_t readCount;
    uint32_t tlsMax, bestModel[2];          // TemplateLengthStatistics: getMax, getBestModel
    uint32_t scatterRepeats;
};

ISAAC_HD inline bool planLpEquals(const double a, const double b) { const double d = a - b; return 0.0000001 >= (d < 0 ? -d : d); }   // Quality.hh:104-107
ISAAC_HD inline bool planLpLess(const double a, const double b) { return !planLpEquals(a, b) && a < b; }                             // :109-112
ISAAC_HD inline unsigned planObservedLength(const isaac_ext_fragment_t &f) { return f.cigarLength ? f.observedLength : 0u; }          // FragmentMetadata.hh:85
ISAAC_HD inline bool planWellAnchored(const isaac_ext_fragment_t &f)                                                                  // :477-483
{
    return f.uniqueSeedCount || (f.nonUniqueSeedOffsetFirst != 0xFFFF && f.nonUniqueSeedOffsetSecond > f.nonUniqueSeedOffsetFirst &&
                                 unsigned(f.nonUniqueSeedOffsetSecond - f.nonUniqueSeedOffsetFirst) >= 32u);
}
/// TemplateLengthStatistics::matchModel (TemplateLengthStatistics.hh:104-176, .cpp:67-77)
ISAAC_HD inline bool planMatchModel(const PlanView &v, const isaac_ext_fragment_t &a, const isaac_ext_fragment_t &b)
{
    const long oa = long(planObservedLength(a)), ob = long(planObservedLength(b));
    long length;
    if (a.position < b.position) { length = b.position + ob - a.position; if (length < oa) length = oa; }
    else { length = a.position + oa - b.position; if (length < ob) length = ob; }
    const unsigned model = a.contigId != b.contigId ? 8u : ((a.position <= b.position ? 0u : 4u) | (a.reverse ? 2u : 0u) | (b.reverse ? 1u : 0u));
    // max_ + TEMPLATE_LENGTH_THRESHOLD is 32-bit arithmetic in the reference (it wraps for the cleared statistics, max_ = -1U)
    return (unsigned long)length <= (unsigned long)uint32_t(v.tlsMax + 50000u) && (model == v.bestModel[0] || model == v.bestModel[1]);
}

/// TemplateBuilder::getBestFragment (:177-226) on list[0..n): the entry --scatter-repeats picks among the equally good ones
ISAAC_HD inline int planBestFragment(const PlanView &v, const isaac_ext_fragment_t *list, const int n, const uint32_t clusterId)
{
    unsigned bestScore = 0xFFFFFFFFu;
    double bestLp = -DBL_MAX;
    int first = 0; unsigned ties = 0;
    for (int i = 0; i < n; ++i)
    {
        const isaac_ext_fragment_t &t = list[i];
        if (bestScore > t.smithWatermanScore || (bestScore == t.smithWatermanScore && planLpLess(bestLp, t.logProbability)))
        {
            bestScore = t.smithWatermanScore; bestLp = t.logProbability; first = i; ties = 1;
        }
        else if (bestScore == t.smithWatermanScore && planLpEquals(bestLp, t.logProbability)) ++ties;
    }
    unsigned pick = v.scatterRepeats ? clusterId % ties : 0u;
    for (int i = first; ; ++i)          // the pick-th of the entries that joined the list after its last restart
    {
        const isaac_ext_fragment_t &t = list[i];
        if (i == first || (bestScore == t.smithWatermanScore && planLpEquals(bestLp, t.logProbability)))
        {
            if (!pick) return i;
            --pick;
        }
    }
}

struct PlanBestPair { unsigned resolved, editDistance, ties; int first[2], picked[2]; };

/// one walk of locateBestPair (:287-391) over the pairs that match the model, in the reference's order; fn(a, b) per pair
template <class F>
ISAAC_HD inline void planForEachPair(const PlanView &v, const isaac_ext_fragment_t *f0, const int n0, const isaac_ext_fragment_t *f1, const int n1, F fn)
{
    int begin0 = 0, begin1 = 0;
    while (n0 != begin0 && n1 != begin1)
    {
        int end0 = begin0 + 1, end1 = begin1 + 1;
        while (n0 != end0 && f0[end0].contigId == f0[begin0].contigId) ++end0;
        while (n1 != end1 && f1[end1].contigId == f1[begin1].contigId) ++end1;
        if (f0[begin0].contigId == f1[begin1].contigId)
        {
            for (int a = begin0; a != end0; ++a)
                for (int b = begin1; b != end1; ++b)
                    if (planMatchModel(v, f0[a], f1[b])) fn(a, b);
            begin0 = end0; begin1 = end1;
        }
        else if (f0[begin0].contigId < f1[begin1].contigId) begin0 = end0;
        else begin1 = end1;
    }
}

/// locateBestPair + the --scatter-repeats swap of buildPairedEndTemplate (:403-408)
ISAAC_HD inline PlanBestPair planLocateBestPair(const PlanView &v, const isaac_ext_fragment_t *f0, const int n0, const isaac_ext_fragment_t *f1,
                                                const int n1, const uint32_t clusterId)
{
    PlanBestPair r = {0u, 0u, 0u, {0, 0}, {0, 0}};
    unsigned long bestScore = ~0ul;
    double bestLp = -DBL_MAX;
    planForEachPair(v, f0, n0, f1, n1, [&](const int a, const int b) {
        const double lp = f0[a].logProbability + f1[b].logProbability;
        const unsigned long score = (unsigned long)(f0[a].smithWatermanScore + f1[b].smithWatermanScore);
        if (0 == r.resolved || bestScore > score || (score == bestScore && planLpLess(bestLp, lp)))
        {
            r.first[0] = a; r.first[1] = b; r.ties = 1; bestScore = score; bestLp = lp;
        }
        else if (score == bestScore && planLpEquals(lp, bestLp)) ++r.ties;
        ++r.resolved;
    });
    if (!r.resolved) return r;
    r.editDistance = unsigned(f0[r.first[0]].editDistance) + f1[r.first[1]].editDistance;        // :389-390, before any swap
    r.picked[0] = r.first[0]; r.picked[1] = r.first[1];
    unsigned pick = v.scatterRepeats ? clusterId % r.ties : 0u;
    if (pick)
    {
        bool seenFirst = false;
        planForEachPair(v, f0, n0, f1, n1, [&](const int a, const int b) {
            if (!seenFirst) { seenFirst = a == r.first[0] && b == r.first[1]; if (!seenFirst) return; }
            const double lp = f0[a].logProbability + f1[b].logProbability;
            const unsigned long score = (unsigned long)(f0[a].smithWatermanScore + f1[b].smithWatermanScore);
            const bool member = (a == r.first[0] && b == r.first[1]) || (score == bestScore && planLpEquals(lp, bestLp));
            if (member && pick != ~0u)
            {
                if (!pick) { r.picked[0] = a; r.picked[1] = b; pick = ~0u; }
                else --pick;
            }
        });
    }
    return r;
}

ISAAC_HD inline void planRequest(const isaac_ext_fragment_t &orphan, const long bestTemplateLength, isaac_ext_rescue_request_t *out,
                                 const unsigned capacity, unsigned &count)
{
    if (count < capacity)
    {
        isaac_ext_rescue_request_t q;
        q.orphanPosition = orphan.position; q.bestTemplateLength = bestTemplateLength; q.orphanReadId = orphan.readId;
        q.orphanContigStrand = (orphan.contigId << 1) | (orphan.reverse ? 1u : 0u);
        q.orphanObservedLength = orphan.observedLength; q.pad = 0;
        out[count] = q;
    }
    ++count;
}

/// \return the number of rescueShadow calls buildTemplate makes for the cluster; the first min(count, capacity) are written to out
ISAAC_HD inline unsigned planClusterRequests(const PlanView &v, const uint32_t cluster, isaac_ext_rescue_request_t *out, const unsigned capacity)
{
    unsigned count = 0;
    if (v.readCount != 2 || !v.built[cluster]) return 0;                                           // single-ended: pickBestFragment, no rescue
    const size_t l = size_t(cluster) * 2;
    const isaac_ext_fragment_t *f[2] = {v.fragments + v.listBegin[l], v.fragments + v.listBegin[l + 1]};
    const int n[2] = {int(v.listCount[l]), int(v.listCount[l + 1])};
    if (n[0] && n[1])                                                                              // pickBestPair
    {
        const PlanBestPair best = planLocateBestPair(v, f[0], n[0], f[1], n[1], cluster);
        if (best.resolved)
        {
            const isaac_ext_fragment_t &read1 = f[0][best.picked[0]], &read2 = f[1][best.picked[1]];
            const bool paired = (planWellAnchored(read1) || planWellAnchored(read2)) && !read1.repeatSeedsCount && !read2.repeatSeedsCount;
            if (paired && !best.editDistance) return 0;                                            // :1064-1071
        }
        // buildDisjoinedTemplate
        const int bestDisjoined[2] = {planBestFragment(v, f[0], n[0], cluster), planBestFragment(v, f[1], n[1], cluster)};
        long knownBestTemplateLength = 0;                                                          // BestPairInfo::getBestTemplateLength (TemplateBuilder.hh:289-300)
        if (best.resolved)
        {
            const isaac_ext_fragment_t &a = f[0][best.picked[0]], &b = f[1][best.picked[1]];
            const uint64_t fa = (((uint64_t(a.contigId) + 1) << 40) | uint64_t(a.position)) << 1, fb = (((uint64_t(b.contigId) + 1) << 40) | uint64_t(b.position)) << 1;
            const long ea = a.position + long(a.observedLength), eb = b.position + long(b.observedLength);
            const uint64_t ra = (((uint64_t(a.contigId) + 1) << 40) | uint64_t((ea > 1 ? ea : 1) - 1)) << 1;
            const uint64_t rb = (((uint64_t(b.contigId) + 1) << 40) | uint64_t((eb > 1 ? eb : 1) - 1)) << 1;
            const uint64_t start = fa < fb ? fa : fb, end = ra < rb ? rb : ra, mask = (uint64_t(1) << 40) - 1;
            knownBestTemplateLength = long((end >> 1) & mask) - long((start >> 1) & mask);
        }
        for (unsigned orphanIndex = 0; orphanIndex < 2; ++orphanIndex)
            for (int oi = 0; oi < n[orphanIndex]; ++oi)
            {
                const isaac_ext_fragment_t &orphan = f[orphanIndex][oi];
                const bool skip = best.resolved ? unsigned(orphan.editDistance) > best.editDistance + 3u
                                                : planLpLess(orphan.logProbability + 100.0, f[orphanIndex][bestDisjoined[orphanIndex]].logProbability);
                if (!skip) planRequest(orphan, knownBestTemplateLength, out, capacity, count);
            }
    }
    else if (n[0] || n[1])                                                                         // TemplateBuilder::rescueShadow
    {
        const unsigned orphanIndex = n[0] ? 0u : 1u;
        const int bestOrphan = planBestFragment(v, f[orphanIndex], n[orphanIndex], cluster);
        for (int oi = 0; oi < n[orphanIndex]; ++oi)
        {
            const isaac_ext_fragment_t &orphan = f[orphanIndex][oi];
            if (!planLpLess(orphan.logProbability + 100.0, f[orphanIndex][bestOrphan].logProbability)) planRequest(orphan, 0, out, capacity, count);
        }
    }
    return count;
}

} // namespace isaac_b200
