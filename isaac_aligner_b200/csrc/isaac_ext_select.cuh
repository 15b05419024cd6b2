// isaac_ext_select_tile: MatchSelector::parallelSelect for one tile (MatchSelector.cpp:370-443) as one C entry point over the
// passes of this library.  Included at the end of isaac_ext.cu.
#pragma once

struct SelectState
{
    std::vector<uint64_t> clusterMatchBegin;
    std::vector<uint16_t> endCyclesMasked;
    std::vector<uint64_t> cycleStats;
    isaac_ext_pack_result_t packed;
};
void releaseSelect(SelectState *state) { delete state; }

extern "C" int isaac_ext_select_tile(isaac_ext_ctx *ctx, const isaac_ext_tile_t *tile, isaac_ext_tile_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!tile || !result) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (tile->matchCount && !tile->matches) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null matches");
    if (!ctx->select) ctx->select = new SelectState();
    SelectState &st = *ctx->select;
    std::memset(result, 0, sizeof(*result));
    int rc = isaac_ext_set_reads(ctx, &tile->reads);
    if (rc) return rc;
    const uint32_t n = tile->reads.clusterCount;
    if (tile->baseQualityCutoff)                                                  // trimLowQualityEnds (MatchSelector.cpp:300)
    {
        st.endCyclesMasked.resize(size_t(n) * tile->reads.readCount);
        rc = isaac_ext_trim_low_quality_ends(ctx, tile->baseQualityCutoff, st.endCyclesMasked.data());
        if (rc) return rc;
        result->endCyclesMasked = st.endCyclesMasked.data();
    }
    // ---- the match list of every cluster: findNextCluster (MatchSelector.cpp:262-277) over the sorted records
    const uint64_t M = tile->matchCount;
    st.clusterMatchBegin.assign(size_t(n) + 1, M);
    std::atomic<int> bad(0);
    auto clusterOf = [&](uint64_t i) { return uint64_t((tile->matches[i].seedId >> 9) & 0x7FFFFFFFull); };      // SeedId.hh:37-127
    parallelRanges(ctx->hostThreads, M, [&](unsigned, size_t b, size_t e) {
        for (size_t i = b; i < e; ++i)
        {
            const uint64_t c = clusterOf(i);
            if (c >= n) { bad = 1; continue; }
            if (i == 0) { for (uint64_t k = 0; k <= c; ++k) st.clusterMatchBegin[k] = 0; continue; }
            const uint64_t p = clusterOf(i - 1);
            if (p > c) { bad = 1; continue; }
            for (uint64_t k = p + 1; k <= c; ++k) st.clusterMatchBegin[k] = i;     // the first record of clusters p + 1 .. c
        }
    });
    if (bad) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "matches must be sorted by cluster and belong to the tile");
    if (!M) std::fill(st.clusterMatchBegin.begin(), st.clusterMatchBegin.end(), uint64_t(0));
    isaac_ext_build_batch_t batch;
    batch.matches = tile->matches; batch.clusterMatchBegin = st.clusterMatchBegin.data(); batch.seeds = tile->seeds;
    batch.seedCount = tile->seedCount; batch.withGaps = tile->withGaps;
    // ---- template length statistics: the caller's, or this tile's own (MatchSelector.cpp:401-417)
    if (tile->tls) { result->tls = *tile->tls; result->tlsStable = 1; }
    else
    {
        rc = isaac_ext_determine_template_length(ctx, &batch, tile->pf, tile->mateDriftRange, &result->tls, &result->tlsStable);
        if (rc) return rc;
    }
    rc = isaac_ext_build_templates(ctx, &batch, &result->tls, &tile->options, &result->templates);
    if (rc) return rc;
    if (tile->cycleStats)                     // first: it reads the templates the build left on the device
    {
        st.cycleStats.resize(4 * size_t(ISAAC_EXT_TILE_CYCLE_STATS_WORDS));
        rc = isaac_ext_tile_cycle_stats(ctx, tile->pf, st.cycleStats.data());
        if (rc) return rc;
        result->cycleStats = st.cycleStats.data();
    }
    rc = isaac_ext_template_stats(ctx, &batch, &result->tls, &result->templates, tile->pf, result->stats);
    if (rc) return rc;
    if (tile->pack)
    {
        rc = isaac_ext_pack_fragments(ctx, &result->templates, tile->pack, &st.packed);
        if (rc) return rc;
        result->packedValid = 1;
    }
    return ISAAC_EXT_OK;
}

/// the records isaac_ext_select_tile left when tile.pack was given
extern "C" int isaac_ext_tile_packed(isaac_ext_ctx *ctx, isaac_ext_pack_result_t *packedOut)
{
    if (!ctx || !packedOut) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->select) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "no tile has been selected on this context");
    *packedOut = ctx->select->packed;
    return ISAAC_EXT_OK;
}
