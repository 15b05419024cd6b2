// isaac_ext_determine_template_length: MatchSelector::determineTemplateLength (MatchSelector.cpp:188-249) for the resident tile.
// The candidate extension of every cluster (TemplateBuilder::buildFragments without gaps, :243) is one isaac_ext_build_fragments
// pass on the GPU; what is left for the host is the reference's sequential bookkeeping: one TemplateLengthDistribution::addTemplate
// per cluster in match-list order until the statistics are stable (TemplateLengthStatistics.cpp:95-160,266-357).  Clusters past
// the point of stability were extended for nothing; their results are not looked at, exactly like the reference never builds them.
// Included at the end of isaac_ext.cu.
#pragma once
#include <cmath>

#include "template_length_host.cuh"

namespace
{
/// buffers of isaac_ext_template_stats
struct TemplateState
{
    DeviceBuffer<isaac_ext_template_t> dStatTemplates;  DeviceBuffer<isaac_ext_fragment_t> dStatFragments;
    DeviceBuffer<uint32_t> dStatCigars;  DeviceBuffer<uint8_t> dStatBytes;  DeviceBuffer<unsigned long long> dStats;
    ~TemplateState() { dStatTemplates.release(); dStatFragments.release(); dStatCigars.release(); dStatBytes.release(); dStats.release(); }
};
} // namespace
void releaseTemplates(TemplateState *state) { delete state; }

extern "C" int isaac_ext_determine_template_length(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const uint8_t *pf,
                                                   int32_t mateDriftRange, isaac_ext_tls_t *tlsOut, uint32_t *stableOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
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
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
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
