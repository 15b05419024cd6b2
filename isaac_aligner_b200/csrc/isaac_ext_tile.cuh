// isaac_ext_build_fragments, isaac_ext_rescue_shadows and isaac_ext_build_templates (include/isaac_ext.h) as device-resident
// passes: the host uploads the tile's seed matches, launches kernels (kernels_tile.cuh for the bookkeeping between the scoring
// kernels), reads four totals on the way (they size the next pass's launches and buffers) and downloads the result.  No host
// thread touches a fragment record.  Included at the end of isaac_ext.cu.
#pragma once
#include "kernels_clip.cuh"
#include "kernels_tile.cuh"

namespace
{

/// the seed matches of a tile on the device: two slots, the tile in work and the one isaac_ext_prefetch_batch fills meanwhile
struct TileInput
{
    DeviceBuffer<isaac_ext_match_t> dMatches;  DeviceBuffer<uint64_t> dMatchBegin;  DeviceBuffer<isaac_ext_seed_t> dSeeds;
    isaac_ext_build_batch_t key{};  uint32_t clusters = 0;  bool staged = false;  cudaEvent_t ready = nullptr;
    void release()
    {
        dMatches.release(); dMatchBegin.release(); dSeeds.release();
        if (ready) cudaEventDestroy(ready);
        ready = nullptr;
    }
};

struct TileState
{
    TileInput input[2];  unsigned activeInput = 0;
    // candidate lists in match slots and the pools of the build pass
    DeviceBuffer<WorkFragment> dWork;  DeviceBuffer<uint32_t> dListBegin, dListCount;  DeviceBuffer<uint8_t> dBuilt;
    DeviceBuffer<isaac_ext_candidate_t> dCand1, dCand3, dAdapterFirst;
    DeviceBuffer<isaac_ext_fragment_t> dFrag1, dFrag3, dFinal;
    DeviceBuffer<uint32_t> dCig1, dCig3, dCigIndel, dGapCounts, dGapBegin;
    DeviceBuffer<IndelTask> dTasks;  DeviceBuffer<IndelResult> dIndel;  DeviceBuffer<uint8_t> dTaskValid;
    DeviceBuffer<uint32_t> dTaskSlots;      // the match slots that hold a simple-indel pair, dense, [slots] = their number
    // the flat result of isaac_ext_build_fragments
    DeviceBuffer<uint32_t> dWords, dFragmentBegin, dWordBegin, dOutCigars;
    DeviceBuffer<isaac_ext_fragment_t> dOutFragments;  DeviceBuffer<uint64_t> dOutBegin;
    PinnedBuffer<isaac_ext_fragment_t> hOutFragments;  PinnedBuffer<uint32_t> hOutCigars;  PinnedBuffer<uint64_t> hOutBegin;  PinnedBuffer<uint8_t> hBuilt;
    // templates
    DeviceBuffer<uint32_t> dRequestCounts, dRequestBegin;  DeviceBuffer<isaac_ext_rescue_request_t> dRequests;
    DeviceBuffer<unsigned char> dScratch;
    DeviceBuffer<isaac_ext_template_t> dTemplates;  DeviceBuffer<isaac_ext_fragment_t> dTemplateFragments;
    DeviceBuffer<uint32_t> dTemplateWords, dTemplateWordBegin, dTemplateCigars, dClippedCigars;
    PinnedBuffer<isaac_ext_template_t> hTemplates;  PinnedBuffer<isaac_ext_fragment_t> hTemplateFragments;  PinnedBuffer<uint32_t> hTemplateCigars;
    PinnedBuffer<uint32_t> hTotals;
    // isaac_ext_build_templates_deferred: two host result sets filled by a copy stream of their own, so that the download of one tile
    // runs next to the kernels of the next; copyDone[set] also holds back the kernels that rewrite the device result (finish pass)
    struct DeferredSet
    {
        PinnedBuffer<isaac_ext_template_t> templates;  PinnedBuffer<isaac_ext_fragment_t> fragments;  PinnedBuffer<uint32_t> cigars, flag;
        cudaEvent_t copyDone = nullptr;
        bool pending = false;
        const char *what = nullptr;
    };
    DeferredSet deferred[2];
    DeviceBuffer<uint32_t> dDeferredFlag;        // snapshot of the error flag at the end of the deferred tile's kernels
    cudaStream_t copyStream = nullptr;
    cudaEvent_t kernelsDone = nullptr;
    unsigned deferredSet = 0;
    bool copyInFlight = false;                   // a deferred download may still be reading the device result
    uint64_t matchTotal = 0;
    // what isaac_ext_tile_cycle_stats needs of the last isaac_ext_build_templates
    DeviceBuffer<FinishSource> dTemplateSources;  DeviceBuffer<unsigned long long> dCycleStats;  DeviceBuffer<uint8_t> dPf;
    bool templatesResident = false;
    ~TileState()
    {
        input[0].release(); input[1].release(); dWork.release(); dListBegin.release(); dListCount.release(); dBuilt.release();
        dCand1.release(); dCand3.release(); dAdapterFirst.release(); dFrag1.release(); dFrag3.release(); dFinal.release(); dCig1.release();
        dCig3.release(); dCigIndel.release(); dGapCounts.release(); dGapBegin.release(); dTasks.release(); dIndel.release(); dTaskValid.release(); dTaskSlots.release();
        dWords.release(); dFragmentBegin.release(); dWordBegin.release(); dOutCigars.release(); dOutFragments.release(); dOutBegin.release();
        hOutFragments.release(); hOutCigars.release(); hOutBegin.release(); hBuilt.release();
        dRequestCounts.release(); dRequestBegin.release(); dRequests.release(); dScratch.release(); dTemplates.release();
        dTemplateFragments.release(); dTemplateWords.release(); dTemplateWordBegin.release(); dTemplateCigars.release(); dClippedCigars.release();
        hTemplates.release(); hTemplateFragments.release(); hTemplateCigars.release(); hTotals.release();
        dTemplateSources.release(); dCycleStats.release(); dPf.release(); dDeferredFlag.release();
        for (DeferredSet &d : deferred)
        {
            d.templates.release(); d.fragments.release(); d.cigars.release(); d.flag.release();
            if (d.copyDone) cudaEventDestroy(d.copyDone);
        }
        if (copyStream) cudaStreamDestroy(copyStream);
        if (kernelsDone) cudaEventDestroy(kernelsDone);
    }
};

/// the clippers of a tile call: ranges[slot] from the first candidate of every slot (checkInitStrand, FragmentBuilder.cpp:173)
int initAdapterSlots(isaac_ext_ctx *ctx, uint32_t slots, const isaac_ext_candidate_t *dFirst)
{
    if (!ctx->adapters.count || !slots) return ISAAC_EXT_OK;
    CK(ctx->dAdapterRanges.reserve(slots));
    adapterInitKernel<<<gridFor(ctx, slots, 128, 16), 128, 0, ctx->stream>>>(ctx->adapters, ctx->ref, ctx->reads, slots, dFirst, ctx->dAdapterRanges.p);
    ++ctx->launches;
    return ctx->cuda(cudaGetLastError(), "adapterInitKernel");
}

/// exclusive prefix sum of n 32-bit counts; out[n] = total (cub::DeviceScan over n + 1 items, the last input is ignored)
cudaError_t exclusiveSum(isaac_ext_ctx *ctx, uint32_t *counts, uint32_t *out, size_t n)
{
    PipelineState &ps = ctx->pipeline;
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, counts, out, int(n) + 1, ctx->stream);
    if (e != cudaSuccess) return e;
    e = ps.dScanTemp.reserve(bytes + 16);
    if (e != cudaSuccess) return e;
    ++ctx->launches;
    return cub::DeviceScan::ExclusiveSum(ps.dScanTemp.p, bytes, counts, out, int(n) + 1, ctx->stream);
}

/// the error flag after a pass: bit TILE_ERROR_MATCHES = malformed input, anything else = a gapped CIGAR beyond its stride
int checkTileFlag(isaac_ext_ctx *ctx, const uint32_t flag, const char *malformed)
{
    if (!flag) return ISAAC_EXT_OK;
    cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream);
    if (flag & TILE_ERROR_MATCHES) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, malformed);
    if (flag & 4u) return ctx->fail(ISAAC_EXT_E_CAPACITY, "a template CIGAR exceeds 60 operations");
    return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR did not fit the cigar stride");
}

/// the tile's matches on 'stream' into 'in' (enqueue only)
int uploadBatch(isaac_ext_ctx *ctx, TileInput &in, const isaac_ext_build_batch_t *batch, const uint32_t n, cudaStream_t stream)
{
    const uint64_t M = batch->clusterMatchBegin[n];
    CK(in.dMatches.reserve(size_t(M) + 1)); CK(in.dMatchBegin.reserve(size_t(n) + 1)); CK(in.dSeeds.reserve(batch->seedCount));
    if (M) CK(cudaMemcpyAsync(in.dMatches.p, batch->matches, size_t(M) * sizeof(isaac_ext_match_t), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(in.dMatchBegin.p, batch->clusterMatchBegin, (size_t(n) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(in.dSeeds.p, batch->seeds, batch->seedCount * sizeof(isaac_ext_seed_t), cudaMemcpyHostToDevice, stream));
    in.key = *batch; in.clusters = n;
    return ISAAC_EXT_OK;
}

TileView tileViewOf(isaac_ext_ctx *ctx, TileState &ts, const isaac_ext_build_batch_t *batch)
{
    TileView v;
    const TileInput &in = ts.input[ts.activeInput];
    v.matches = in.dMatches.p; v.clusterMatchBegin = in.dMatchBegin.p; v.seeds = in.dSeeds.p; v.seedCount = batch->seedCount;
    v.clusters = ctx->clusterCount; v.readCount = ctx->reads.readCount; v.repeatThreshold = ctx->cfg.repeatThreshold;
    v.gapLimit = ctx->cfg.semialignedGapLimit; v.withGaps = batch->withGaps ? 1u : 0u; v.gappedMismatchesMax = ctx->cfg.gappedMismatchesMax;
    v.readLength[0] = ctx->reads.readLength[0]; v.readLength[1] = ctx->reads.readLength[1];
    v.contigLength = ctx->ref.contigLength; v.contigCount = ctx->ref.contigCount;
    v.work = ts.dWork.p; v.listBegin = ts.dListBegin.p; v.listCount = ts.dListCount.p; v.built = ts.dBuilt.p;
    v.frag1 = ts.dFrag1.p; v.cig1 = ts.dCig1.p; v.frag3 = ts.dFrag3.p; v.cig3 = ts.dCig3.p; v.cigIndel = ts.dCigIndel.p;
    return v;
}

/// FragmentBuilder::build of every cluster of the resident read set, results left on the device: the final candidate list of
/// (cluster, readIndex) is ts.dFinal[ts.dListBegin[l] .. + ts.dListCount[l]), CIGARs in the three pools
int tileBuildDevice(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch)
{
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!batch || !batch->clusterMatchBegin || !batch->seeds || !batch->seedCount) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null batch");
    if (batch->seedCount > TILE_MAX_SEEDS) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "more than 64 seeds per cluster");
    const uint32_t n = ctx->clusterCount, rc = ctx->reads.readCount;
    const size_t lists = size_t(n) * rc;
    const uint64_t M = batch->clusterMatchBegin[n];
    if (M && !batch->matches) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null matches");
    for (uint32_t s = 0; s < batch->seedCount; ++s)
        if (batch->seeds[s].readIndex >= rc) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "seed refers to an unknown read");
    if (M > (FINISH_POOL_MASK + 1ull) / 8) return ctx->fail(ISAAC_EXT_E_CAPACITY, "too many matches in one batch: split the tile");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->tile) ctx->tile = new TileState();
    TileState &ts = *ctx->tile;
    ts.matchTotal = M;
    ts.templatesResident = false;
    PhaseTimer timer("build");
    const size_t slots = size_t(M) + 1;
    CK(ts.dWork.reserve(slots)); CK(ts.dListBegin.reserve(lists + 1)); CK(ts.dListCount.reserve(lists + 1)); CK(ts.dBuilt.reserve(size_t(n) + 1));
    CK(ts.dCand1.reserve(slots)); CK(ts.dFrag1.reserve(slots)); CK(ts.dCig1.reserve(slots * 3)); CK(ts.dFinal.reserve(slots));
    CK(ts.dTasks.reserve(slots)); CK(ts.dIndel.reserve(slots)); CK(ts.dTaskValid.reserve(slots)); CK(ts.dCigIndel.reserve(slots * 5));
    CK(ts.dTaskSlots.reserve(slots + 1));
    CK(ts.dGapCounts.reserve(lists + 1)); CK(ts.dGapBegin.reserve(lists + 1)); CK(ts.hTotals.reserve(8));
    {
        // the matches: on the device already when isaac_ext_prefetch_batch was given this batch and the isaac_ext_set_reads of this
        // tile took it over (trusted once: the caller's buffers may change afterwards), else uploaded now
        TileInput &in = ts.input[ts.activeInput];
        if (in.staged && in.clusters == n && in.key.matches == batch->matches && in.key.clusterMatchBegin == batch->clusterMatchBegin &&
            in.key.seeds == batch->seeds && in.key.seedCount == batch->seedCount)
        {
            CK(cudaEventSynchronize(in.ready));
            in.staged = false;
        }
        else
        {
            in.staged = false;
            const int rcu = uploadBatch(ctx, in, batch, n, ctx->stream);
            if (rcu) return rcu;
        }
    }
    CK(cudaMemsetAsync(ts.dCand1.p, 0xFF, slots * sizeof(isaac_ext_candidate_t), ctx->stream));      // readId = TILE_NO_CANDIDATE everywhere
    CK(cudaMemsetAsync(ts.dTaskValid.p, 0, slots, ctx->stream));
    const bool withAdapters = ctx->adapters.count != 0;
    if (withAdapters) CK(ts.dAdapterFirst.reserve(lists * 2));
    TileView v = tileViewOf(ctx, ts, batch);
    const unsigned clusterGrid = gridFor(ctx, n, 128, 16), listGrid = gridFor(ctx, lists, 128, 16);

    // ---- B1, K1
    buildCandidatesKernel<<<clusterGrid, 128, 0, ctx->stream>>>(v, ts.dCand1.p, withAdapters ? ts.dAdapterFirst.p : nullptr, ctx->errorFlag.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    const uint32_t *clip = nullptr;
    if (withAdapters && M)
    {
        int rca = initAdapterSlots(ctx, uint32_t(lists * 2), ts.dAdapterFirst.p);
        if (!rca) rca = adapterSlotClip(ctx, uint32_t(M), ts.dCand1.p, nullptr, ctx->stream, &clip);
        if (rca) return rca;
    }
    int rcode = ungappedDevice(ctx, uint32_t(M), ts.dCand1.p, ts.dFrag1.p, ts.dCig1.p, nullptr, ctx->stream, clip);
    if (rcode) return rcode;
    // ---- B2, simple indels, B3
    pairIndelKernel<<<listGrid, 128, 0, ctx->stream>>>(v, ts.dTasks.p, ts.dTaskValid.p);
    ++ctx->launches;
    if (ctx->cfg.semialignedGapLimit && M)
    {
        // the slots that hold a pair as a dense list (a few percent of the slots): full warps for the indel search
        size_t bytes = 0;
        cub::CountingInputIterator<uint32_t> slotIds(0);
        CK(cub::DeviceSelect::Flagged(nullptr, bytes, slotIds, ts.dTaskValid.p, ts.dTaskSlots.p, ts.dTaskSlots.p + M, int(M), ctx->stream));
        CK(ctx->pipeline.dScanTemp.reserve(bytes + 16));
        CK(cub::DeviceSelect::Flagged(ctx->pipeline.dScanTemp.p, bytes, slotIds, ts.dTaskValid.p, ts.dTaskSlots.p, ts.dTaskSlots.p + M, int(M), ctx->stream));
        simpleIndelKernel<<<gridFor(ctx, (M + 7) / 8, 128, 8), 128, 0, ctx->stream>>>(ctx->ref, ctx->reads, ctx->sp, uint32_t(M), ts.dTasks.p, ts.dIndel.p,
                                                                                   ts.dTaskSlots.p, ts.dTaskSlots.p + M, ts.dCigIndel.p);
        ctx->launches += 2;
    }
    applyIndelKernel<<<listGrid, 128, 0, ctx->stream>>>(v, ts.dIndel.p, ts.dTaskValid.p, ts.dGapCounts.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    uint32_t n3 = 0;
    if (batch->withGaps)
    {
        CK(exclusiveSum(ctx, ts.dGapCounts.p, ts.dGapBegin.p, lists));
        CK(cudaMemcpyAsync(ts.hTotals.p, ts.dGapBegin.p + lists, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ts.hTotals.p + 1, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        rcode = checkTileFlag(ctx, ts.hTotals.p[1], "malformed match batch (offsets, seed index or contig out of range)");
        if (rcode) return rcode;
        n3 = ts.hTotals.p[0];
        timer.mark("B1 K1 B2 indel B3");
        if (uint64_t(n3) * TILE_GAPPED_STRIDE > FINISH_POOL_MASK) return ctx->fail(ISAAC_EXT_E_CAPACITY, "too many candidates for the gapped aligner in one batch: split the tile");
    }
    // ---- K2, B4
    if (n3)
    {
        CK(ts.dCand3.reserve(n3)); CK(ts.dFrag3.reserve(n3)); CK(ts.dCig3.reserve(size_t(n3) * TILE_GAPPED_STRIDE));
        v = tileViewOf(ctx, ts, batch);
        gapCandidatesKernel<<<listGrid, 128, 0, ctx->stream>>>(v, ts.dGapBegin.p, ts.dCand3.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        const uint32_t *clip3 = nullptr;
        rcode = adapterSlotClip(ctx, n3, ts.dCand3.p, nullptr, ctx->stream, &clip3);
        if (!rcode) rcode = gappedDevice(ctx, n3, ts.dCand3.p, TILE_GAPPED_STRIDE, ts.dFrag3.p, ts.dCig3.p, nullptr, ctx->stream, clip3);
        if (rcode) return rcode;
    }
    acceptGappedKernel<<<listGrid, 128, 0, ctx->stream>>>(v, n3 ? ts.dGapBegin.p : nullptr, ts.dFinal.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    timer.mark("K2 B4 (launched)");
    return ISAAC_EXT_OK;
}

} // namespace

void releaseTile(TileState *state) { delete state; }

/// isaac_ext_set_reads of a prefetched tile: the batch prefetched with it becomes the current one
void tileTakeOverPrefetchedBatch(isaac_ext_ctx *ctx)
{
    if (!ctx->tile) return;
    TileState &ts = *ctx->tile;
    if (ts.input[ts.activeInput ^ 1u].staged) ts.activeInput ^= 1u;
}

extern "C" int isaac_ext_prefetch_batch(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, uint32_t clusterCount)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch || !batch->clusterMatchBegin || !batch->seeds || !batch->seedCount || !clusterCount) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null batch");
    if (batch->clusterMatchBegin[clusterCount] && !batch->matches) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null matches");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->stageStream) CK(cudaStreamCreateWithFlags(&ctx->stageStream, cudaStreamNonBlocking));
    if (!ctx->tile) ctx->tile = new TileState();
    TileInput &standby = ctx->tile->input[ctx->tile->activeInput ^ 1u];
    if (!standby.ready) CK(cudaEventCreateWithFlags(&standby.ready, cudaEventDisableTiming));
    standby.staged = false;
    const int rc = uploadBatch(ctx, standby, batch, clusterCount, ctx->stageStream);
    if (rc) return rc;
    CK(cudaEventRecord(standby.ready, ctx->stageStream));
    standby.staged = true;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_build_fragments(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, isaac_ext_build_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!result) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null result");
    int rc = tileBuildDevice(ctx, batch);
    if (rc) return rc;
    TileState &ts = *ctx->tile;
    const uint32_t n = ctx->clusterCount;
    const size_t lists = size_t(n) * ctx->reads.readCount;
    const unsigned listGrid = gridFor(ctx, lists + 1, 128, 16);
    PhaseTimer timer("build result");
    // ---- the flat result: prefix sums over list sizes and CIGAR words, one dense copy, one download
    CK(ts.dWords.reserve(lists + 1)); CK(ts.dFragmentBegin.reserve(lists + 1)); CK(ts.dWordBegin.reserve(lists + 1)); CK(ts.dOutBegin.reserve(lists + 1));
    countListWordsKernel<<<listGrid, 128, 0, ctx->stream>>>(lists, ts.dListBegin.p, ts.dListCount.p, ts.dFinal.p, ts.dWords.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(exclusiveSum(ctx, ts.dListCount.p, ts.dFragmentBegin.p, lists));
    CK(exclusiveSum(ctx, ts.dWords.p, ts.dWordBegin.p, lists));
    CK(cudaMemcpyAsync(ts.hTotals.p, ts.dFragmentBegin.p + lists, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hTotals.p + 1, ts.dWordBegin.p + lists, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hTotals.p + 2, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = checkTileFlag(ctx, ts.hTotals.p[2], "malformed match batch (offsets, seed index or contig out of range)");
    if (rc) return rc;
    const uint32_t fragmentTotal = ts.hTotals.p[0], wordTotal = ts.hTotals.p[1];
    CK(ts.dOutFragments.reserve(size_t(fragmentTotal) + 1)); CK(ts.dOutCigars.reserve(size_t(wordTotal) + 1));
    CK(ts.hOutFragments.reserve(size_t(fragmentTotal) + 1)); CK(ts.hOutCigars.reserve(size_t(wordTotal) + 1)); CK(ts.hOutBegin.reserve(lists + 1));
    CK(ts.hBuilt.reserve(size_t(n) + 1));
    flattenListsKernel<<<listGrid, 128, 0, ctx->stream>>>(lists, ts.dListBegin.p, ts.dListCount.p, ts.dFinal.p, ts.dFragmentBegin.p, ts.dWordBegin.p,
                                                           ts.dCig1.p, ts.dCigIndel.p, ts.dCig3.p, ts.dOutFragments.p, ts.dOutCigars.p, ts.dOutBegin.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (fragmentTotal) CK(cudaMemcpyAsync(ts.hOutFragments.p, ts.dOutFragments.p, size_t(fragmentTotal) * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (wordTotal) CK(cudaMemcpyAsync(ts.hOutCigars.p, ts.dOutCigars.p, size_t(wordTotal) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hOutBegin.p, ts.dOutBegin.p, (lists + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hBuilt.p, ts.dBuilt.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    timer.mark("flatten + copies");
    result->fragments = ts.hOutFragments.p; result->readFragmentBegin = ts.hOutBegin.p; result->cigars = ts.hOutCigars.p;
    result->built = ts.hBuilt.p; result->fragmentCount = fragmentTotal; result->cigarWords = wordTotal;
    return ISAAC_EXT_OK;
}

namespace
{

/// What the rescue pass leaves on the device for request i of n: the shadow list ps.dOutFragments[ps.dOutBegin[i] .. ps.dOutBegin[i + 1])
/// with CIGAR words in ps.dOutCigars, ps.dRescued[i] = return value of rescueShadow.
struct RescueTotals { uint32_t fragments = 0, words = 0; };

/// ShadowAligner::rescueShadow behind R1 for n requests whose scan windows are in ps.dShadowTasks (ShadowAligner.cpp:195-290)
int rescueDeviceCore(isaac_ext_ctx *ctx, const uint32_t n, RescueTotals &totals)
{
    PipelineState &ps = ctx->pipeline;
    PhaseTimer timer("rescue");
    CK(ps.dOutBegin.reserve(size_t(n) + 1)); CK(ps.dRescued.reserve(size_t(n) + 1)); CK(ps.hTotals.reserve(4));
    totals = RescueTotals();
    if (!n)
    {
        CK(cudaMemsetAsync(ps.dOutBegin.p, 0, sizeof(uint64_t), ctx->stream));
        return ISAAC_EXT_OK;
    }
    // ---- K5: candidate positions of every request (:195-236).  Requests with a small window (nearly all) take one warp
    // each, the others one CTA each with the full 4^7 table and the 10000 cap.  Both kernels always get the whole request
    // list and decide per request with the one predicate shadowTaskIsSmall (kernels_shadow.cuh), so every request is
    // handled by exactly one of them whatever the mix of window sizes and read lengths.
    const unsigned grid = std::max(1u, std::min<unsigned>(n, unsigned(ctx->smCount) * 6));
    CK(ps.dTaskBegin.reserve(n)); CK(ps.dTaskCount.reserve(n)); CK(ps.dPoolSize.reserve(1)); CK(ps.dLargeTasks.reserve(size_t(n) + 1));
    CK(ps.dShadowScratch.reserve(size_t(grid) * SHADOW_SCRATCH));
    uint64_t capacity = std::max<uint64_t>(ps.dCand.capacity, uint64_t(n) * 24 + 4096);
    unsigned long long poolSize64 = 0;
    for (int attempt = 0; attempt < 2; ++attempt)
    {
        CK(ps.dCand.reserve(capacity));
        CK(cudaMemsetAsync(ps.dPoolSize.p, 0, sizeof(unsigned long long), ctx->stream));
        // a request neither kernel takes cannot exist, but an empty list is the safe reading of a skipped one
        CK(cudaMemsetAsync(ps.dTaskBegin.p, 0, size_t(n) * sizeof(uint32_t), ctx->stream));
        CK(cudaMemsetAsync(ps.dTaskCount.p, 0, size_t(n) * sizeof(uint32_t), ctx->stream));
        const uint32_t poolCapacity = uint32_t(std::min<uint64_t>(ps.dCand.capacity, 0xFFFFFFFFull));
        CK(cudaMemsetAsync(ps.dLargeTasks.p + n, 0, sizeof(uint32_t), ctx->stream));
        shadowCandidatesWarpKernel<<<gridFor(ctx, uint64_t(n) * 32, SHADOW_WARPS * 32, 6), SHADOW_WARPS * 32, 0, ctx->stream>>>(
            ctx->ref, ctx->reads, n, ps.dShadowTasks.p, ps.dCand.p, poolCapacity, ps.dPoolSize.p, ps.dTaskBegin.p, ps.dTaskCount.p, ctx->errorFlag.p,
            ps.dLargeTasks.p, ps.dLargeTasks.p + n);
        shadowCandidatesKernel<<<grid, SHADOW_BLOCK, 0, ctx->stream>>>(ctx->ref, ctx->reads, n, ps.dShadowTasks.p, ps.dShadowScratch.p, ps.dCand.p,
                                                                        poolCapacity, ps.dPoolSize.p, ps.dTaskBegin.p, ps.dTaskCount.p,
                                                                        ctx->errorFlag.p, ps.dLargeTasks.p, ps.dLargeTasks.p + n);
        ctx->launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&poolSize64, ps.dPoolSize.p, sizeof(poolSize64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (poolSize64 <= poolCapacity) break;
        CK(cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream));   // bit 2: the pool was too small, the counter holds the need
        if (attempt || poolSize64 > 0xFFFFFFF0ull)
            return ctx->fail(ISAAC_EXT_E_CAPACITY, "too many shadow candidate positions in one rescue batch: split the batch");
        capacity = poolSize64 + 1024;
    }
    const uint32_t poolSize = uint32_t(poolSize64);
    timer.mark("K5 shadow candidates");
    if (poolSize)
    {
        // ---- K1: UngappedAligner::alignUngapped of every candidate position (:205-236)
        CK(ps.dFrag.reserve(poolSize)); CK(ps.dCig.reserve(size_t(poolSize) * 3));
        const uint32_t *clip = nullptr;
        if (ctx->adapters.count)
        {
            CK(ps.dAdapterFirst.reserve(n)); CK(ps.dSlot.reserve(poolSize));
            shadowAdapterSlotsKernel<<<gridFor(ctx, uint64_t(n) * 32, 128, 16), 128, 0, ctx->stream>>>(
                n, ps.dTaskBegin.p, ps.dTaskCount.p, ps.dCand.p, ps.dAdapterFirst.p, ps.dSlot.p);
            ++ctx->launches;
            CK(cudaGetLastError());
            int rca = initAdapterSlots(ctx, n, ps.dAdapterFirst.p);
            if (!rca) rca = adapterSlotClip(ctx, poolSize, ps.dCand.p, ps.dSlot.p, ctx->stream, &clip);
            if (rca) return rca;
        }
        const int rc = ungappedDevice(ctx, poolSize, ps.dCand.p, ps.dFrag.p, ps.dCig.p, nullptr, ctx->stream, clip);
        if (rc) return rc;
    }
    // ---- R2: shadow lists, best shadow, neighbours to gap-align (:205-256)
    const unsigned rgrid = gridFor(ctx, n, 128, 16);
    CK(ps.dKept.reserve(size_t(poolSize) + 1)); CK(ps.dAdoptedBy.reserve(size_t(poolSize) + 1)); CK(ps.dListState.reserve(n));
    CK(ps.dCounts.reserve(size_t(n) * 3 + 1)); CK(ps.dBegins.reserve(size_t(n) * 3 + 3));
    uint32_t *gapCounts = ps.dCounts.p, *listCounts = ps.dCounts.p + n, *wordCounts = ps.dCounts.p + 2 * size_t(n);
    uint32_t *gapBegin = ps.dBegins.p, *fragmentBegin = ps.dBegins.p + (n + 1), *wordBegin = ps.dBegins.p + 2 * (size_t(n) + 1);
    shadowSelectKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dTaskCount.p, ps.dFrag.p, ps.dKept.p, ps.dListState.p, gapCounts);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(exclusiveSum(ctx, gapCounts, gapBegin, n));
    CK(cudaMemcpyAsync(ps.hTotals.p, gapBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t n3 = ps.hTotals.p[0];
    timer.mark("K1 ungapped + R2 lists");
    if (n3)
    {
        // ---- K2: the gapped aligner on the neighbours (:249-262), candidates written in list order by the device
        CK(ps.dCand3.reserve(n3)); CK(ps.dFrag3.reserve(n3)); CK(ps.dCig3.reserve(size_t(n3) * TILE_GAPPED_STRIDE));
        CK(ps.dSlot3.reserve(n3)); CK(ps.dSources.reserve(n3));
        shadowGapKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dListState.p, gapBegin, ps.dFrag.p, ps.dCig.p, ps.dKept.p,
                                                        ps.dCand3.p, ps.dSlot3.p, ps.dSources.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        const uint32_t *clip = nullptr;
        int rc = adapterSlotClip(ctx, n3, ps.dCand3.p, ps.dSlot3.p, ctx->stream, &clip);
        if (!rc) rc = gappedDevice(ctx, n3, ps.dCand3.p, TILE_GAPPED_STRIDE, ps.dFrag3.p, ps.dCig3.p, nullptr, ctx->stream, clip);
        if (rc) return rc;
    }
    // ---- R3: acceptance in list order, best shadow first (:255-290), then the flat result
    shadowAcceptKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dListState.p, gapBegin, ps.dSources.p, ps.dFrag.p, ps.dFrag3.p,
                                                       ctx->cfg.gappedMismatchesMax, ps.dKept.p, ps.dAdoptedBy.p, listCounts, wordCounts, ps.dRescued.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(exclusiveSum(ctx, listCounts, fragmentBegin, n));
    CK(exclusiveSum(ctx, wordCounts, wordBegin, n));
    CK(cudaMemcpyAsync(ps.hTotals.p + 1, fragmentBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ps.hTotals.p + 2, wordBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ps.hTotals.p + 3, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int rcFlag = checkTileFlag(ctx, ps.hTotals.p[3], "rescue request refers to an unknown read or contig");
    if (rcFlag) return rcFlag;
    totals.fragments = ps.hTotals.p[1]; totals.words = ps.hTotals.p[2];
    timer.mark("K2 gapped + R3 accept");
    CK(ps.dOutFragments.reserve(size_t(totals.fragments) + 1)); CK(ps.dOutCigars.reserve(size_t(totals.words) + 1));
    shadowFlattenKernel<<<gridFor(ctx, n, 128, 16), 128, 0, ctx->stream>>>(
        n, ps.dTaskBegin.p, listCounts, fragmentBegin, wordBegin, ps.dKept.p, ps.dAdoptedBy.p, ps.dFrag.p, ps.dCig.p, ps.dFrag3.p, ps.dCig3.p,
        TILE_GAPPED_STRIDE, ps.dOutFragments.p, ps.dOutCigars.p, ps.dOutBegin.p, totals.fragments);
    ++ctx->launches;
    CK(cudaGetLastError());
    timer.mark("flatten (launched)");
    return ISAAC_EXT_OK;
}

} // namespace

/// isaac_ext_rescue_shadows with its flat result in result set 'slot' (0 or 1) of the context
static int rescueShadowsInto(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t n, const isaac_ext_rescue_request_t *requests,
                             isaac_ext_rescue_result_t *result, const unsigned slot)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!tls || !result || (n && !requests)) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (ctx->reads.readCount != 2) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "shadow rescue needs paired reads (ShadowAligner.cpp:170)");
    CK(cudaSetDevice(ctx->device));
    PipelineState &ps = ctx->pipeline;
    if (!ctx->tile) ctx->tile = new TileState();
    TileState &ts = *ctx->tile;
    const ShadowWindowModel model = makeShadowWindowModel(*tls);
    CK(ps.hOutBegin[slot].reserve(size_t(n) + 1)); CK(ps.hRescued[slot].reserve(size_t(n) + 1));
    if (n && shadowModelCoherent(model))                                        // :164-168
    {
        // ---- R1: rescue windows (calculateShadowRescueRange :119-149, rescueShadow :170-198), one request per thread
        CK(ts.dRequests.reserve(n)); CK(ps.dShadowTasks.reserve(n));
        CK(cudaMemcpyAsync(ts.dRequests.p, requests, size_t(n) * sizeof(isaac_ext_rescue_request_t), cudaMemcpyHostToDevice, ctx->stream));
        shadowWindowsKernel<<<gridFor(ctx, n, 128, 16), 128, 0, ctx->stream>>>(n, ts.dRequests.p, model, ctx->reads.readLength[0], ctx->reads.readLength[1],
                                                                              ctx->reads.readTotal, ctx->ref.contigCount, ctx->ref.contigLength,
                                                                              ps.dShadowTasks.p, ctx->errorFlag.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        RescueTotals totals;
        const int rc = rescueDeviceCore(ctx, n, totals);
        if (rc) return rc;
        CK(ps.hOutFragments[slot].reserve(size_t(totals.fragments) + 1)); CK(ps.hOutCigars[slot].reserve(size_t(totals.words) + 1));
        if (totals.fragments) CK(cudaMemcpyAsync(ps.hOutFragments[slot].p, ps.dOutFragments.p, size_t(totals.fragments) * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (totals.words) CK(cudaMemcpyAsync(ps.hOutCigars[slot].p, ps.dOutCigars.p, size_t(totals.words) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ps.hOutBegin[slot].p, ps.dOutBegin.p, (size_t(n) + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ps.hRescued[slot].p, ps.dRescued.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        result->fragments = ps.hOutFragments[slot].p; result->requestFragmentBegin = ps.hOutBegin[slot].p; result->cigars = ps.hOutCigars[slot].p;
        result->rescued = ps.hRescued[slot].p; result->fragmentCount = totals.fragments; result->cigarWords = totals.words;
        return ISAAC_EXT_OK;
    }
    // nothing to rescue (no requests, or template length statistics without a coherent pair of models, :164-168)
    std::fill(ps.hOutBegin[slot].p, ps.hOutBegin[slot].p + n + 1, uint64_t(0));
    std::fill(ps.hRescued[slot].p, ps.hRescued[slot].p + n, uint8_t(0));
    result->fragments = ps.hOutFragments[slot].p; result->requestFragmentBegin = ps.hOutBegin[slot].p; result->cigars = ps.hOutCigars[slot].p;
    result->rescued = ps.hRescued[slot].p; result->fragmentCount = 0; result->cigarWords = 0;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_rescue_shadows(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t n,
                                        const isaac_ext_rescue_request_t *requests, isaac_ext_rescue_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    return rescueShadowsInto(ctx, tls, n, requests, result, 0);
}

/// alignment::TemplateBuilder over the resident tile: build, plan, rescue, finish, end clippers -- all on the
/// device.  The reference decides per cluster, while it walks the candidate lists, which orphans deserve a rescueShadow call;
/// whether a call is made never depends on the outcome of another call (TemplateBuilder.cpp:519-525, 746-753), so the calls of the
/// whole tile are planned first (plan_device.cuh), answered in one rescue pass and consumed in plan order by the finish pass
/// (finish_device.cuh).
static int buildTemplatesCore(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                              const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result, const bool deferred)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!batch || !tls || !options || !result) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    PhaseTimer timer("templates");
    int rc = tileBuildDevice(ctx, batch);
    if (rc) return rc;
    TileState &ts = *ctx->tile;
    PipelineState &ps = ctx->pipeline;
    const uint32_t n = ctx->clusterCount, readCount = ctx->reads.readCount;
    const size_t count = size_t(n) * readCount;
    const unsigned clusterGrid = gridFor(ctx, n, 128, 16);

    // ---- plan: the rescueShadow calls of every cluster and their scan windows
    PlanView pv;
    pv.fragments = ts.dFinal.p; pv.listBegin = ts.dListBegin.p; pv.listCount = ts.dListCount.p; pv.built = ts.dBuilt.p; pv.readCount = readCount;
    pv.tlsMax = tls->max; pv.bestModel[0] = tls->bestModel[0]; pv.bestModel[1] = tls->bestModel[1]; pv.scatterRepeats = options->scatterRepeats;
    const ShadowWindowModel model = makeShadowWindowModel(*tls);
    uint32_t requestTotal = 0;
    RescueTotals totals;
    CK(ts.dRequestBegin.reserve(size_t(n) + 1));
    if (2 == readCount && shadowModelCoherent(model))                           // ShadowAligner.cpp:164-168: no rescue without a coherent model
    {
        CK(ts.dRequestCounts.reserve(size_t(n) + 1));
        planCountKernel<<<clusterGrid, 128, 0, ctx->stream>>>(pv, n, ts.dRequestCounts.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(exclusiveSum(ctx, ts.dRequestCounts.p, ts.dRequestBegin.p, n));
        CK(cudaMemcpyAsync(ts.hTotals.p, ts.dRequestBegin.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ts.hTotals.p + 1, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        rc = checkTileFlag(ctx, ts.hTotals.p[1], "malformed match batch (offsets, seed index or contig out of range)");
        if (rc) return rc;
        requestTotal = ts.hTotals.p[0];
        timer.mark("build + plan count");
        if (requestTotal)
        {
            CK(ts.dRequests.reserve(requestTotal)); CK(ps.dShadowTasks.reserve(requestTotal));
            planWriteKernel<<<clusterGrid, 128, 0, ctx->stream>>>(pv, n, ts.dRequestBegin.p, model, ctx->reads.readLength[0], ctx->reads.readLength[1],
                                                                  ctx->ref.contigLength, ts.dRequests.p, ps.dShadowTasks.p);
            ++ctx->launches;
            CK(cudaGetLastError());
        }
        rc = rescueDeviceCore(ctx, requestTotal, totals);
        if (rc) return rc;
        timer.mark("rescue");
    }
    else
    {
        // an incoherent model answers every call "nothing rescued" with an empty list; the calls are still made and counted
        // (TemplateBuilder never looks at the model before it calls): plan them, leave every answer empty
        if (2 == readCount)
        {
            CK(ts.dRequestCounts.reserve(size_t(n) + 1));
            planCountKernel<<<clusterGrid, 128, 0, ctx->stream>>>(pv, n, ts.dRequestCounts.p);
            ++ctx->launches;
            CK(exclusiveSum(ctx, ts.dRequestCounts.p, ts.dRequestBegin.p, n));
            CK(cudaMemcpyAsync(ts.hTotals.p, ts.dRequestBegin.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            requestTotal = ts.hTotals.p[0];
        }
        else CK(cudaMemsetAsync(ts.dRequestBegin.p, 0, (size_t(n) + 1) * sizeof(uint32_t), ctx->stream));
        CK(ps.dOutBegin.reserve(size_t(requestTotal) + 1)); CK(ps.dRescued.reserve(size_t(requestTotal) + 1));
        CK(cudaMemsetAsync(ps.dOutBegin.p, 0, (size_t(requestTotal) + 1) * sizeof(uint64_t), ctx->stream));
        CK(cudaMemsetAsync(ps.dRescued.p, 0, size_t(requestTotal) + 1, ctx->stream));
    }

    // ---- finish: the BamTemplate of every cluster
    FinishView fv;
    fv.fragments = ts.dFinal.p; fv.listBegin = ts.dListBegin.p; fv.listCount = ts.dListCount.p; fv.built = ts.dBuilt.p;
    fv.cigarPools[0] = ts.dCig1.p; fv.cigarPools[1] = ts.dCigIndel.p; fv.cigarPools[2] = ts.dCig3.p; fv.cigarPools[FINISH_POOL_RESCUE] = ps.dOutCigars.p;
    fv.rescueFragments = ps.dOutFragments.p; fv.requestFragmentBegin = ps.dOutBegin.p; fv.rescued = ps.dRescued.p;
    fv.clusterRequestBegin = ts.dRequestBegin.p;
    fv.readCount = readCount; fv.tlsMin = tls->min; fv.tlsMax = tls->max; fv.bestModel[0] = tls->bestModel[0]; fv.bestModel[1] = tls->bestModel[1];
    fv.scatterRepeats = options->scatterRepeats; fv.mapqThreshold = options->mapqThreshold; fv.dodgyAlignmentScore = options->dodgyAlignmentScore;
    fv.logMismatchQ40 = ctx->logMismatchQ40;
    finishRestOfGenome(fv, ctx->contigLength.data(), uint32_t(ctx->contigLength.size()), ctx->reads.readLength);
    const uint64_t scratchBytes = finishScratchBytes(totals.fragments, ts.matchTotal) + uint64_t(n) * finishScratchBytes(0, 0);
    CK(ts.dScratch.reserve(scratchBytes));
    CK(ts.dTemplates.reserve(size_t(n) + 1)); CK(ts.dTemplateFragments.reserve(count + 1));
    CK(ts.dTemplateWords.reserve(count + 1)); CK(ts.dTemplateWordBegin.reserve(count + 1)); CK(ts.dTemplateSources.reserve(count + 1));
    if (ts.copyInFlight)
    {
        // the download of the previous (deferred) tile reads the buffers the finish pass is about to rewrite
        for (TileState::DeferredSet &d : ts.deferred) if (d.pending) CK(cudaStreamWaitEvent(ctx->stream, d.copyDone, 0));
        ts.copyInFlight = false;
    }
    finishTemplatesKernel<<<clusterGrid, 128, 0, ctx->stream>>>(fv, n, ts.input[ts.activeInput].dMatchBegin.p, ts.dScratch.p, ts.dTemplates.p, ts.dTemplateFragments.p,
                                                                ts.dTemplateWords.p, ts.dTemplateSources.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(exclusiveSum(ctx, ts.dTemplateWords.p, ts.dTemplateWordBegin.p, count));
    CK(cudaMemcpyAsync(ts.hTotals.p, ts.dTemplateWordBegin.p + count, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hTotals.p + 1, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = checkTileFlag(ctx, ts.hTotals.p[1], "malformed match batch (offsets, seed index or contig out of range)");
    if (rc) return rc;
    uint64_t words = ts.hTotals.p[0];
    timer.mark("finish");
    CK(ts.dTemplateCigars.reserve(words + 1));
    gatherTemplateCigarsKernel<<<gridFor(ctx, count, 128, 16), 128, 0, ctx->stream>>>(count, ts.dTemplateFragments.p, ts.dTemplateWordBegin.p, ts.dCig1.p,
                                                                                    ts.dCigIndel.p, ts.dCig3.p, ps.dOutCigars.p, ts.dTemplateCigars.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    const uint32_t *dCigars = ts.dTemplateCigars.p;

    // ---- end clippers on the kept templates (MatchSelector.cpp:336-346): one kernel pass over the tile
    if (options->clipFlags & (ISAAC_EXT_CLIP_SEMIALIGNED | ISAAC_EXT_CLIP_OVERLAPPING))
    {
        const uint64_t outWords = words + 4 * count;                        // a clip adds at most two operations per side
        if (outWords > 0xFFFFFFFFull) return ctx->fail(ISAAC_EXT_E_CAPACITY, "CIGAR pool of the tile exceeds 2^32 words");
        CK(ts.dClippedCigars.reserve(outWords));
        CK(cudaMemsetAsync(ts.dClippedCigars.p, 0, outWords * sizeof(uint32_t), ctx->stream));
        clipTemplateEndsKernel<<<clusterGrid, 128, 0, ctx->stream>>>(ctx->ref, ctx->reads, n, options->clipFlags, ts.dTemplates.p, ts.dTemplateFragments.p,
                                                                     ts.dTemplateCigars.p, ts.dClippedCigars.p, ctx->errorFlag.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        dCigars = ts.dClippedCigars.p; words = outWords;
    }
    if (deferred)
    {
        // ---- one download, on the copy stream: the call returns as soon as it is queued (isaac_ext_fetch_templates waits for it)
        if (!ts.copyStream)
        {
            CK(cudaStreamCreateWithFlags(&ts.copyStream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ts.kernelsDone, cudaEventDisableTiming));
            for (TileState::DeferredSet &d : ts.deferred) CK(cudaEventCreateWithFlags(&d.copyDone, cudaEventDisableTiming));
        }
        ts.deferredSet ^= 1u;
        TileState::DeferredSet &d = ts.deferred[ts.deferredSet];
        if (d.pending) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "two deferred tiles are waiting to be fetched: isaac_ext_fetch_templates first");
        CK(d.templates.reserve(size_t(n) + 1)); CK(d.fragments.reserve(count + 1)); CK(d.cigars.reserve(words + 1)); CK(d.flag.reserve(1));
        CK(ts.dDeferredFlag.reserve(2));
        CK(cudaMemcpyAsync(ts.dDeferredFlag.p + ts.deferredSet, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaEventRecord(ts.kernelsDone, ctx->stream));
        CK(cudaStreamWaitEvent(ts.copyStream, ts.kernelsDone, 0));
        CK(cudaMemcpyAsync(d.templates.p, ts.dTemplates.p, size_t(n) * sizeof(isaac_ext_template_t), cudaMemcpyDeviceToHost, ts.copyStream));
        CK(cudaMemcpyAsync(d.fragments.p, ts.dTemplateFragments.p, count * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ts.copyStream));
        if (words) CK(cudaMemcpyAsync(d.cigars.p, dCigars, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, ts.copyStream));
        CK(cudaMemcpyAsync(d.flag.p, ts.dDeferredFlag.p + ts.deferredSet, sizeof(uint32_t), cudaMemcpyDeviceToHost, ts.copyStream));
        CK(cudaEventRecord(d.copyDone, ts.copyStream));
        d.pending = true;
        ts.copyInFlight = true;
        timer.mark("gather + clip (download queued)");
        result->templates = d.templates.p; result->fragments = d.fragments.p; result->cigars = d.cigars.p;
        result->cigarWords = words; result->rescueRequests = requestTotal;
        ts.templatesResident = true;
        return ISAAC_EXT_OK;
    }
    // ---- one download
    CK(ts.hTemplates.reserve(size_t(n) + 1)); CK(ts.hTemplateFragments.reserve(count + 1)); CK(ts.hTemplateCigars.reserve(words + 1));
    CK(cudaMemcpyAsync(ts.hTemplates.p, ts.dTemplates.p, size_t(n) * sizeof(isaac_ext_template_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hTemplateFragments.p, ts.dTemplateFragments.p, count * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (words) CK(cudaMemcpyAsync(ts.hTemplateCigars.p, dCigars, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ts.hTotals.p + 1, ctx->errorFlag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = checkTileFlag(ctx, ts.hTotals.p[1], "malformed match batch (offsets, seed index or contig out of range)");
    if (rc) return rc;
    timer.mark("gather + clip + copies");
    result->templates = ts.hTemplates.p; result->fragments = ts.hTemplateFragments.p; result->cigars = ts.hTemplateCigars.p;
    result->cigarWords = words; result->rescueRequests = requestTotal;
    ts.templatesResident = true;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_build_templates(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                         const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result)
{
    return buildTemplatesCore(ctx, batch, tls, options, result, false);
}

extern "C" int isaac_ext_build_templates_deferred(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                                  const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result)
{
    return buildTemplatesCore(ctx, batch, tls, options, result, true);
}

extern "C" int isaac_ext_fetch_templates(isaac_ext_ctx *ctx, const isaac_ext_template_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!result || !ctx->tile) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "no deferred result to fetch");
    TileState &ts = *ctx->tile;
    for (TileState::DeferredSet &d : ts.deferred)
    {
        if (!d.pending || result->templates != d.templates.p) continue;
        CK(cudaEventSynchronize(d.copyDone));
        d.pending = false;
        return checkTileFlag(ctx, d.flag.p[0], "malformed match batch (offsets, seed index or contig out of range)");
    }
    return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "this result is not waiting to be fetched");
}

extern "C" int isaac_ext_tile_cycle_stats(isaac_ext_ctx *ctx, const uint8_t *pf, uint64_t *statsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!statsOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->tile || !ctx->tile->templatesResident)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "no templates on the device: isaac_ext_build_templates / isaac_ext_select_tile of the tile first");
    CK(cudaSetDevice(ctx->device));
    TileState &ts = *ctx->tile;
    PipelineState &ps = ctx->pipeline;
    const uint32_t n = ctx->clusterCount;
    const size_t words = 4 * size_t(ISAAC_EXT_TILE_CYCLE_STATS_WORDS);
    CK(ts.dCycleStats.reserve(words));
    CK(cudaMemsetAsync(ts.dCycleStats.p, 0, words * sizeof(unsigned long long), ctx->stream));
    if (pf) { CK(ts.dPf.reserve(n)); CK(cudaMemcpyAsync(ts.dPf.p, pf, n, cudaMemcpyHostToDevice, ctx->stream)); }
    tileCycleStatsKernel<<<gridFor(ctx, n, 128, 16), 128, 0, ctx->stream>>>(ctx->ref, ctx->reads, ctx->sp, n, ts.dTemplates.p, ts.dTemplateFragments.p,
                                                                          ts.dTemplateSources.p, ts.dCig1.p, ts.dCigIndel.p, ts.dCig3.p, ps.dOutCigars.p,
                                                                          pf ? ts.dPf.p : nullptr, ts.dCycleStats.p, ctx->errorFlag.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(statsOut, ts.dCycleStats.p, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag)
    {
        cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream);
        return ctx->fail(ISAAC_EXT_E_CAPACITY, "an alignment score exceeds TileStats::maxAlignmentScore_ (0x1FFF; the reference asserts)");
    }
    return ISAAC_EXT_OK;
}

/// TileStats::finalize (TileStats.hh:239-340) on the raw counters of one block: the five "fragments with X mismatches so far"
/// arrays (plain and uniquely aligned) become per-cycle counts of fragments that have exactly 1, <= 2, <= 3, <= 4, <= 5 mismatches
extern "C" void isaac_ext_tile_cycle_stats_finalize(uint64_t *block)
{
    for (unsigned group = 0; group < 2; ++group)
    {
        long *x[5];
        for (unsigned k = 0; k < 5; ++k) x[k] = reinterpret_cast<long *>(block) + (group ? TCS_X : TCS_UNIQUE_X) + k * 1024;
        // a mismatch seen at one cycle stays for all later cycles (:242-262, 295-313)
        for (unsigned k = 0; k < 5; ++k) for (unsigned c = 1; c < 1024; ++c) x[k][c] += x[k][c - 1];
        // a fragment with a second mismatch stops being a one-mismatch fragment (:264-276, 315-327)
        for (unsigned k = 0; k + 1 < 5; ++k) for (unsigned c = 0; c < 1024; ++c) x[k][c] -= x[k + 1][c];
        // two-mismatch fragments include the one-mismatch ones and so on (:278-290, 329-340)
        for (unsigned k = 1; k < 5; ++k) for (unsigned c = 0; c < 1024; ++c) x[k][c] += x[k - 1][c];
    }
}

extern "C" int isaac_ext_trim_low_quality_ends(isaac_ext_ctx *ctx, uint32_t baseQualityCutoff, uint16_t *endCyclesMaskedOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reads first");
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->reads.readTotal;
    trimLowQualityEndsKernel<<<gridFor(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->reads, baseQualityCutoff, ctx->slot().masked.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (endCyclesMaskedOut) CK(cudaMemcpyAsync(endCyclesMaskedOut, ctx->slot().masked.p, size_t(n) * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ISAAC_EXT_OK;
}
