// isaac_ext_pack_fragments: matchSelector::FragmentCollector::add for every stored template of the resident tile
// (pack_fragments.cuh).  Included by isaac_ext.cu.
#pragma once
#include "pack_fragments.cuh"

struct PackState
{
    DeviceBuffer<isaac_ext_template_t> dTemplates;  DeviceBuffer<isaac_ext_fragment_t> dFragments;  DeviceBuffer<uint32_t> dCigars;
    DeviceBuffer<uint8_t> dPf, dRecords, dInitialized;
    DeviceBuffer<int32_t> dXy;
    DeviceBuffer<uint64_t> dBarcodeSequence, dContigBinBegin, dFStrandPos, dRecordOffset;
    HostBuffer<uint64_t> recordOffset;
    DeviceBuffer<uint32_t> dBinIndex;
    DeviceBuffer<unsigned long long> dStored;
    PinnedBuffer<uint8_t> hRecords, hInitialized;
    PinnedBuffer<uint64_t> hFStrandPos;
    PinnedBuffer<unsigned long long> hStored;
    cudaEvent_t kernelBegin = nullptr, kernelEnd = nullptr;       // timing of the kernel alone (isaac_ext_pack_result_t::kernelMs)
    void release()
    {
        if (kernelBegin) cudaEventDestroy(kernelBegin);
        if (kernelEnd) cudaEventDestroy(kernelEnd);
        kernelBegin = kernelEnd = nullptr;
        dTemplates.release(); dFragments.release(); dCigars.release(); dPf.release(); dRecords.release(); dInitialized.release();
        dXy.release(); dBarcodeSequence.release(); dContigBinBegin.release(); dFStrandPos.release(); dBinIndex.release(); dStored.release(); dRecordOffset.release();
        hRecords.release(); hInitialized.release(); hFStrandPos.release(); hStored.release();
    }
};

void releasePack(PackState *state) { if (state) { state->release(); delete state; } }

extern "C" int isaac_ext_pack_fragments(isaac_ext_ctx *ctx, const isaac_ext_template_result_t *templates,
                                        const isaac_ext_pack_options_t *options, isaac_ext_pack_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!templates || !options || !result) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reads first");
    if (!templates->templates || !templates->fragments || (templates->cigarWords && !templates->cigars))
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null template result");
    const bool binMap = options->distributionBinSize != 0;
    if (binMap && (!options->contigBinBegin || !options->contigCount || (options->contigBinBegin[options->contigCount] && !options->binIndex)))
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "bin map without its vectors");
    const uint32_t n = ctx->clusterCount, rc = ctx->reads.readCount;
    const size_t count = size_t(n) * rc;
    // every aligned fragment's CIGAR must lie inside the pool it indexes
    {
        std::atomic<int> bad(0);
        parallelRanges(ctx->hostThreads, count, [&](unsigned, size_t b, size_t e) {
            for (size_t i = b; i < e; ++i)
            {
                const isaac_ext_fragment_t &f = templates->fragments[i];
                if (f.cigarLength && uint64_t(f.cigarOffset) + f.cigarLength > templates->cigarWords) bad = 1;
            }
        });
        if (bad) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "a fragment's CIGAR lies outside the CIGAR pool");
    }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->pack) ctx->pack = new PackState();
    PackState &st = *ctx->pack;

    PackView v{};
    v.clusterCount = n; v.readCount = rc;
    v.readLength[0] = ctx->reads.readLength[0]; v.readLength[1] = rc > 1 ? ctx->reads.readLength[1] : 0;
    packLayout(v);
    v.tile = options->tile; v.barcodeIdx = options->barcodeIdx; v.keepUnaligned = options->keepUnaligned;
    v.bcl = ctx->slot().bclStage.p; v.bclBytes = uint64_t(n) * (v.readLength[0] + v.readLength[1]);
    // record offsets: FragmentBuffer slots, or (compact) the prefix sum of the records' total lengths, computed here from the
    // host copy of the templates the kernel is about to get
    st.recordOffset.reserve(count + 1);
    uint64_t *offsets = st.recordOffset.p;
    {
        PackView h = v;
        h.templates = templates->templates; h.fragments = templates->fragments;
        parallelRanges(ctx->hostThreads, n, [&](unsigned, size_t b, size_t e) {
            for (size_t c = b; c < e; ++c)
                for (unsigned r = 0; r < rc; ++r)
                    offsets[c * rc + r + 1] = options->compact ? packRecordBytes(h, uint32_t(c), r)
                                                               : (r + 1 < rc ? h.readOffset[r + 1] : h.recordLength) - h.readOffset[r];
        });
        offsets[0] = 0;
        for (size_t i = 1; i <= count; ++i) offsets[i] += offsets[i - 1];
    }
    const size_t recordBytes = offsets[count];

    CK(st.dTemplates.reserve(n)); CK(st.dFragments.reserve(count)); CK(st.dCigars.reserve(templates->cigarWords + 1));
    CK(st.dRecords.reserve(recordBytes)); CK(st.dFStrandPos.reserve(count)); CK(st.dInitialized.reserve(count)); CK(st.dStored.reserve(1));
    CK(st.hRecords.reserve(recordBytes)); CK(st.hFStrandPos.reserve(count)); CK(st.hInitialized.reserve(count)); CK(st.hStored.reserve(1));
    CK(cudaMemcpyAsync(st.dTemplates.p, templates->templates, size_t(n) * sizeof(isaac_ext_template_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(st.dFragments.p, templates->fragments, count * sizeof(isaac_ext_fragment_t), cudaMemcpyHostToDevice, ctx->stream));
    if (templates->cigarWords)
        CK(cudaMemcpyAsync(st.dCigars.p, templates->cigars, templates->cigarWords * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    v.templates = st.dTemplates.p; v.fragments = st.dFragments.p; v.cigars = st.dCigars.p;
    if (options->pf)
    {
        CK(st.dPf.reserve(n));
        CK(cudaMemcpyAsync(st.dPf.p, options->pf, n, cudaMemcpyHostToDevice, ctx->stream));
        v.pf = st.dPf.p;
    }
    if (options->xy)
    {
        CK(st.dXy.reserve(2 * size_t(n)));
        CK(cudaMemcpyAsync(st.dXy.p, options->xy, 2 * size_t(n) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        v.xy = st.dXy.p;
    }
    if (options->barcodeSequence)
    {
        CK(st.dBarcodeSequence.reserve(n));
        CK(cudaMemcpyAsync(st.dBarcodeSequence.p, options->barcodeSequence, size_t(n) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        v.barcodeSequence = st.dBarcodeSequence.p;
    }
    if (binMap)
    {
        const uint64_t bins = options->contigBinBegin[options->contigCount];
        CK(st.dContigBinBegin.reserve(size_t(options->contigCount) + 1)); CK(st.dBinIndex.reserve(bins + 1));
        CK(cudaMemcpyAsync(st.dContigBinBegin.p, options->contigBinBegin, (size_t(options->contigCount) + 1) * sizeof(uint64_t),
                           cudaMemcpyHostToDevice, ctx->stream));
        if (bins) CK(cudaMemcpyAsync(st.dBinIndex.p, options->binIndex, bins * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        v.contigBinBegin = st.dContigBinBegin.p; v.binIndex = st.dBinIndex.p;
        v.contigCount = options->contigCount; v.distributionBinSize = options->distributionBinSize;
    }
    if (options->compact)
    {
        CK(st.dRecordOffset.reserve(count + 1));
        CK(cudaMemcpyAsync(st.dRecordOffset.p, offsets, (count + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        v.recordOffset = st.dRecordOffset.p;
    }
    v.records = st.dRecords.p; v.fStrandPos = st.dFStrandPos.p; v.initialized = st.dInitialized.p;
    CK(cudaMemsetAsync(st.dStored.p, 0, sizeof(unsigned long long), ctx->stream));

    // one warp per cluster; the staging area of the block's warps is dynamic shared memory (2 x 150: 8 x 832 B)
    const unsigned stagingBytes = packStagingBytes(v);
    unsigned warps = 8;
    while (warps > 1 && size_t(warps) * stagingBytes > 96 * 1024) warps /= 2;
    const size_t shared = size_t(warps) * stagingBytes;
    if (shared > 48 * 1024)
        CK(cudaFuncSetAttribute(packFragmentsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shared)));
    const unsigned block = warps * 32;
    if (!st.kernelBegin) { CK(cudaEventCreate(&st.kernelBegin)); CK(cudaEventCreate(&st.kernelEnd)); }
    CK(cudaEventRecord(st.kernelBegin, ctx->stream));
    packFragmentsKernel<<<gridFor(ctx, uint64_t(n) * 32, block, 8), block, shared, ctx->stream>>>(v, st.dStored.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaEventRecord(st.kernelEnd, ctx->stream));
    if (recordBytes) CK(cudaMemcpyAsync(st.hRecords.p, st.dRecords.p, recordBytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(st.hFStrandPos.p, st.dFStrandPos.p, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(st.hInitialized.p, st.dInitialized.p, count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(st.hStored.p, st.dStored.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    const int rcSync = ctx->cuda(cudaStreamSynchronize(ctx->stream), "packFragmentsKernel");
    if (rcSync) return rcSync;
    result->records = st.hRecords.p; result->fStrandPos = st.hFStrandPos.p; result->initialized = st.hInitialized.p;
    result->recordLength = v.recordLength; result->readOffset[0] = v.readOffset[0]; result->readOffset[1] = v.readOffset[1];
    result->headerLength = PACK_HEADER_BYTES;
    result->storedFragments = *st.hStored.p;
    result->recordOffset = offsets; result->recordBytes = recordBytes;
    result->kernelMs = 0.0f;
    CK(cudaEventElapsedTime(&result->kernelMs, st.kernelBegin, st.kernelEnd));
    return ISAAC_EXT_OK;
}
