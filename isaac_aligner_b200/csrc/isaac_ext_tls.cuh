// isaac_ext_determine_template_length: MatchSelector::determineTemplateLength (MatchSelector.cpp:188-249) for the resident tile.
// The candidate extension of every cluster (TemplateBuilder::buildFragments without gaps, :243) is one isaac_ext_build_fragments
// pass on the GPU; what is left for the host is the reference's sequential bookkeeping: one TemplateLengthDistribution::addTemplate
// per cluster in match-list order until the statistics are stable (TemplateLengthStatistics.cpp:95-160,266-357).  Clusters past
// the point of stability were extended for nothing; their results are not looked at, exactly like the reference never builds them.
// Included at the end of isaac_ext.cu.
#pragma once
#include <cmath>

namespace
{

/// alignment::TemplateLengthDistribution (TemplateLengthStatistics.hh:262-340)
struct TemplateLengthDistributionHost
{
    static const unsigned UPDATE_FREQUENCY = 10000;                 // :317
    static const unsigned TEMPLATE_LENGTH_THRESHOLD = 50000;        // :220
    static const unsigned INVALID_MODEL = 8;                        // InvalidAlignmentModel (:59)
    int mateDriftRange;
    unsigned min = -1U, max = -1U, median = -1U, lowStdDev = -1U, highStdDev = -1U, bestModels[2] = {INVALID_MODEL, INVALID_MODEL};
    bool stable = false;
    unsigned templateCount = 0, uniqueCount = 0, count = 0;
    std::vector<unsigned> histograms[INVALID_MODEL], lengthList;
    double lowerPercent, upperPercent, lowerPercent1z, upperPercent1z;

    explicit TemplateLengthDistributionHost(int drift) : mateDriftRange(drift)
    {
        // TemplateLengthStatistics.cpp:31-38 (boost::math::erf; STANDARD_DEVIATIONS_MAX = 3.0)
        const double interval = std::erf(3.0 / std::sqrt(2.0)), interval1z = std::erf(1.0 / std::sqrt(2.0));
        lowerPercent = (1.0 - interval) / 2.0; upperPercent = (1.0 + interval) / 2.0;
        lowerPercent1z = (1.0 - interval1z) / 2.0; upperPercent1z = (1.0 + interval1z) / 2.0;
    }

    struct Snapshot { unsigned min, median, max, low, high, m0, m1; };
    Snapshot snapshot() const { return Snapshot{min, median, max, lowStdDev, highStdDev, bestModels[0], bestModels[1]}; }
    bool sameNumbers(const Snapshot &o) const { return o.min == min && o.median == median && o.max == max && o.low == lowStdDev && o.high == highStdDev; }

    static unsigned alignmentClass(unsigned model) { return model < 4 ? model : ((~model) & 3); }

    /// updateStatistics (TemplateLengthStatistics.cpp:105-160)
    void updateStatistics()
    {
        const Snapshot old = snapshot();
        bestModels[0] = histograms[1].size() <= histograms[0].size() ? 0u : 1u;
        bestModels[1] = (bestModels[0] + 1) % 2;
        for (unsigned i = 2; i < INVALID_MODEL; ++i)
        {
            if (histograms[i].size() > histograms[bestModels[0]].size()) { bestModels[1] = bestModels[0]; bestModels[0] = i; }
            else if (histograms[i].size() > histograms[bestModels[1]].size()) bestModels[1] = i;
        }
        lengthList.clear();
        lengthList.insert(lengthList.end(), histograms[bestModels[0]].begin(), histograms[bestModels[0]].end());
        lengthList.insert(lengthList.end(), histograms[bestModels[1]].begin(), histograms[bestModels[1]].end());
        std::sort(lengthList.begin(), lengthList.end());
        const size_t n = lengthList.size();
        min = n ? lengthList[unsigned(n * lowerPercent)] : 0;
        median = n ? lengthList[unsigned(n * 0.5)] : TEMPLATE_LENGTH_THRESHOLD / 2;
        max = n ? lengthList[unsigned(n * upperPercent)] : TEMPLATE_LENGTH_THRESHOLD;
        lowStdDev = n ? median - lengthList[unsigned(n * lowerPercent1z)] : median;
        highStdDev = n ? lengthList[unsigned(n * upperPercent1z)] - median : median;
        if (sameNumbers(old) && old.m0 == bestModels[0] && old.m1 == bestModels[1]) stable = true;
    }

    /// addTemplate (:266-340) on the final fragment lists of the two reads of one cluster
    bool addTemplate(const isaac_ext_fragment_t *f0, size_t n0, const isaac_ext_fragment_t *f1, size_t n1, const uint32_t *cigars)
    {
        if (!n0 || !n1) return stable;
        ++templateCount;
        if (n0 > 1 || n1 > 1) return stable;
        ++uniqueCount;
        if (f0->contigId != f1->contigId) return stable;
        const isaac_ext_fragment_t *f[2] = {f0, f1};
        for (unsigned i = 0; i < 2; ++i)
        {
            const uint32_t firstOp = cigars[f[i]->cigarOffset], lastOp = cigars[f[i]->cigarOffset + f[i]->cigarLength - 1];
            if ((firstOp & 0xFu) == ISAAC_EXT_CIGAR_INSERT || (lastOp & 0xFu) == ISAAC_EXT_CIGAR_INSERT) return stable;
        }
        // TemplateLengthStatistics::getLength (TemplateLengthStatistics.hh:165-176)
        const unsigned long length = f0->position < f1->position
            ? (unsigned long)std::max<long>(f1->position + long(f1->observedLength) - f0->position, long(f0->observedLength))
            : (unsigned long)std::max<long>(f0->position + long(f0->observedLength) - f1->position, long(f1->observedLength));
        if (length > TEMPLATE_LENGTH_THRESHOLD) return stable;
        // alignmentModel (:153-163); the contigs are equal here
        const unsigned model = (f0->position <= f1->position ? 0u : 4u) | (f0->reverse ? 2u : 0u) | (f1->reverse ? 1u : 0u);
        histograms[model].push_back(unsigned(length));
        ++count;
        if (0 == count % UPDATE_FREQUENCY)
        {
            const Snapshot old = snapshot();
            updateStatistics();
            if (sameNumbers(old)) stable = true;
        }
        return stable;
    }

    /// finalize (:342-357)
    bool finalize()
    {
        const Snapshot old = snapshot();
        updateStatistics();
        if (sameNumbers(old)) stable = true;
        return stable;
    }
};

} // namespace

extern "C" int isaac_ext_determine_template_length(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const uint8_t *pf,
                                                   int32_t mateDriftRange, isaac_ext_tls_t *tlsOut, uint32_t *stableOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch || !tlsOut || !stableOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    TemplateLengthDistributionHost distribution(mateDriftRange);
    if (ctx->reads.readCount == 2)                                   // single-ended data: the cleared statistics (:204-208)
    {
        isaac_ext_build_batch_t noGaps = *batch;
        noGaps.withGaps = 0;                                        // buildFragments(..., false) (:243)
        isaac_ext_build_result_t built;
        const int rc = isaac_ext_build_fragments(ctx, &noGaps, &built);
        if (rc) return rc;
        for (uint32_t c = 0; c < ctx->clusterCount && !distribution.stable; ++c)
        {
            // good pf clusters with a real match list only (:226-230)
            if (batch->clusterMatchBegin[c] == batch->clusterMatchBegin[c + 1] || (pf && !pf[c])) continue;
            const uint64_t *begin = built.readFragmentBegin + size_t(c) * 2;
            distribution.addTemplate(built.fragments + begin[0], size_t(begin[1] - begin[0]), built.fragments + begin[1],
                                     size_t(begin[2] - begin[1]), built.cigars);
        }
        if (!distribution.stable) distribution.finalize();
    }
    tlsOut->min = distribution.min; tlsOut->max = distribution.max; tlsOut->median = distribution.median;
    tlsOut->lowStdDev = distribution.lowStdDev; tlsOut->highStdDev = distribution.highStdDev;
    tlsOut->bestModel[0] = distribution.bestModels[0]; tlsOut->bestModel[1] = distribution.bestModels[1];
    tlsOut->mateDriftRange = mateDriftRange;
    *stableOut = distribution.stable ? 1u : 0u;
    return ISAAC_EXT_OK;
}

/// matchSelector::TileBarcodeStats of the tile's templates (see include/isaac_ext.h)
extern "C" int isaac_ext_template_stats(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                        const isaac_ext_template_result_t *templates, const uint8_t *pf, uint64_t *statsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch || !tls || !templates || !statsOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reads first");
    const uint32_t n = ctx->clusterCount, rc = ctx->reads.readCount;
    const size_t counters = 4 * size_t(ISAAC_EXT_TEMPLATE_STATS_COUNTERS);
    std::memset(statsOut, 0, counters * sizeof(uint64_t));
    if (!n) return ISAAC_EXT_OK;
    if (!templates->templates || !templates->fragments || (templates->cigarWords && !templates->cigars))
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null template result");
    CK(cudaSetDevice(ctx->device));
    // the template type of every cluster (MatchSelector.cpp:300-365)
    std::vector<uint8_t> types(n);
    parallelRanges(ctx->hostThreads, n, [&](unsigned, size_t b, size_t e) {
        for (size_t c = b; c < e; ++c)
        {
            const uint64_t mb = batch->clusterMatchBegin[c], me = batch->clusterMatchBegin[c + 1];
            if (mb == me) types[c] = TEMPLATE_NMNM;
            else if (matchIsNoMatch(batch->matches[mb])) types[c] = ((batch->matches[mb].seedId >> 1) & 0xFFu) == 0xFFu ? TEMPLATE_QC : TEMPLATE_NMNM;   // SeedId::isNSeedId (SeedId.hh:117)
            else types[c] = templates->templates[c].hadFragments ? TEMPLATE_NORMAL : TEMPLATE_RM;
        }
    });
    if (!ctx->templates) ctx->templates = new TemplateState();
    TemplateState &st = *ctx->templates;
    const size_t count = size_t(n) * rc;
    CK(st.dStatTemplates.reserve(n)); CK(st.dStatFragments.reserve(count)); CK(st.dStatCigars.reserve(templates->cigarWords + 1));
    CK(st.dStatBytes.reserve(2 * size_t(n))); CK(st.dStats.reserve(counters));
    CK(cudaMemcpyAsync(st.dStatTemplates.p, templates->templates, size_t(n) * sizeof(isaac_ext_template_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(st.dStatFragments.p, templates->fragments, count * sizeof(isaac_ext_fragment_t), cudaMemcpyHostToDevice, ctx->stream));
    if (templates->cigarWords)
        CK(cudaMemcpyAsync(st.dStatCigars.p, templates->cigars, templates->cigarWords * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(st.dStatBytes.p, types.data(), n, cudaMemcpyHostToDevice, ctx->stream));
    if (pf) CK(cudaMemcpyAsync(st.dStatBytes.p + n, pf, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(st.dStats.p, 0, counters * sizeof(unsigned long long), ctx->stream));
    const TlsDevice t = {tls->min, tls->max, {tls->bestModel[0], tls->bestModel[1]}};
    templateStatsKernel<<<gridFor(ctx, n, 128, 16), 128, 0, ctx->stream>>>(ctx->reads, t, n, st.dStatTemplates.p, st.dStatFragments.p, st.dStatCigars.p,
                                                                          st.dStatBytes.p, pf ? st.dStatBytes.p + n : nullptr, st.dStats.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(statsOut, st.dStats.p, counters * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    return ctx->cuda(cudaStreamSynchronize(ctx->stream), "templateStatsKernel");
}
