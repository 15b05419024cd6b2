// End-to-end variants of the two micro entry points: host buffers in, 64-byte records + a DENSE CIGAR pool out, the
// batch cut into chunks whose H2D copy, kernels and D2H copies overlap on three streams (copy engines in both directions
// run next to the SMs).  Included at the end of isaac_ext.cu.
#pragma once
#include "kernels_compact.cuh"

namespace
{

// candidates per chunk: four chunks of the split Smith-Waterman path (2 waves x 148 SMs x 4 blocks x 128 pairs each);
// ISAAC_EXT_E2E_CHUNK overrides it for experiments
static uint32_t e2eChunk()
{
    static const uint32_t value = [] { const char *e = std::getenv("ISAAC_EXT_E2E_CHUNK"); const long v = e ? std::atol(e) : 0; return uint32_t(v > 0 ? v : 4 * 303104); }();
    return value;
}

struct E2eState
{
    static const int NB = 4;          // result buffer sets in flight: being copied out, being computed, queued behind
    cudaStream_t sH = nullptr, sC = nullptr, sD = nullptr;
    cudaEvent_t evC[NB] = {}, evD[NB] = {};
    std::vector<cudaEvent_t> evH;     // one per chunk: its candidates are on the device
    DeviceBuffer<isaac_ext_candidate_t> dCand;   // the whole call's candidates
    DeviceBuffer<isaac_ext_fragment_t> dFrag[NB];
    DeviceBuffer<uint32_t> dCig[NB], dPool[NB], dBlock[NB], dTotal[NB];
    // isaac_ext_align_batch_packed: the ungapped records of a chunk (never leave the device), what is kept, the packed records
    DeviceBuffer<isaac_ext_fragment_t> dFragU[NB];  DeviceBuffer<uint32_t> dCigU[NB];  DeviceBuffer<uint8_t> dStatus[NB];
    DeviceBuffer<isaac_ext_alignment_t> dPacked[NB];
    DeviceBuffer<unsigned long long> dRunning;   // [0], [1] running pool offset of the two passes, [2 + b] pool offset of the item in set b
    PinnedBuffer<uint32_t> hTotal;               // mapped: written by cigarScanBlockSumsKernel
    bool ready = false;
    void release()
    {
        for (int i = 0; i < NB; ++i)
        {
            dFrag[i].release(); dCig[i].release(); dPool[i].release(); dBlock[i].release(); dTotal[i].release();
            dFragU[i].release(); dCigU[i].release(); dStatus[i].release(); dPacked[i].release();
            if (evC[i]) cudaEventDestroy(evC[i]);
            if (evD[i]) cudaEventDestroy(evD[i]);
        }
        for (cudaEvent_t e : evH) cudaEventDestroy(e);
        evH.clear();
        dCand.release(); dRunning.release();
        if (sH) cudaStreamDestroy(sH);
        if (sC) cudaStreamDestroy(sC);
        if (sD) cudaStreamDestroy(sD);
        hTotal.release();
        ready = false;
    }
};

/// where one pass (ungapped or gapped) of a *_batch_compact call puts its results
struct E2ePass
{
    bool gapped;
    isaac_ext_fragment_t *fragmentsOut; uint32_t *poolOut; uint64_t poolCapacity; uint64_t *wordsOut;
    uint64_t base = 0;
    bool overflow = false;
};

} // namespace

/// Three streams: sH uploads the candidates chunk after chunk (nothing waits on the way back); sC validates every chunk and
/// runs its passes (ungapped, gapped, or both one after the other) into result sets that rotate as they leave the device; sD
/// copies records + dense pool out.  With both passes in one call the copy engine stays busy with the ungapped records while
/// the SMs run the Smith-Waterman of the same chunk.  The host only waits for an item's CIGAR word count (a mapped word) to
/// place its pool in the caller's buffer.
static int extendCompact(isaac_ext_ctx *ctx, E2eState &st, uint32_t n, const isaac_ext_candidate_t *candidates, E2ePass *passes,
                         const unsigned passCount)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    for (unsigned p = 0; p < passCount; ++p) if (passes[p].wordsOut) *passes[p].wordsOut = 0;
    if (!n) return ISAAC_EXT_OK;
    if (!candidates) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    for (unsigned p = 0; p < passCount; ++p)
        if (!passes[p].fragmentsOut || !passes[p].poolOut || !passes[p].wordsOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    const int NB = E2eState::NB;
    CK(cudaSetDevice(ctx->device));
    if (!st.ready)
    {
        CK(cudaStreamCreateWithFlags(&st.sH, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sC, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sD, cudaStreamNonBlocking));
        for (int i = 0; i < NB; ++i)
        {
            CK(cudaEventCreateWithFlags(&st.evC[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&st.evD[i], cudaEventDisableTiming));
        }
        CK(st.hTotal.reserve(NB));
        CK(st.dRunning.reserve(2 + NB));
        st.ready = true;
    }
    const uint32_t E2E_CHUNK = e2eChunk();
    bool anyGapped = false;
    for (unsigned p = 0; p < passCount; ++p) anyGapped |= passes[p].gapped;
    const uint32_t strideMax = anyGapped ? 32u : 3u;
    const uint32_t chunkMax = std::min(n, E2E_CHUNK);
    const uint32_t blocksMax = (chunkMax + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
    CK(st.dCand.reserve(n));
    for (int i = 0; i < NB; ++i)
    {
        CK(st.dFrag[i].reserve(chunkMax)); CK(st.dCig[i].reserve(size_t(chunkMax) * strideMax));
        CK(st.dPool[i].reserve(size_t(chunkMax) * strideMax)); CK(st.dBlock[i].reserve(blocksMax)); CK(st.dTotal[i].reserve(1));
    }
    const uint32_t chunks = (n + E2E_CHUNK - 1) / E2E_CHUNK;
    const uint32_t items = chunks * passCount;           // item w = pass w % passCount of chunk w / passCount
    while (st.evH.size() < chunks)
    {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st.evH.push_back(e);
    }
    auto chunkSize = [&](uint32_t k) { return std::min(E2E_CHUNK, n - k * E2E_CHUNK); };
    // ISAAC_EXT_TRACE: device-side time stamps of every item (upload end on sH, compute begin/end on sC, copy-out begin/end on sD)
    const bool trace = std::getenv("ISAAC_EXT_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) { tev.resize(size_t(items) * 5 + 1); for (cudaEvent_t &e : tev) cudaEventCreate(&e); cudaEventRecord(tev.back(), st.sH); }
    CK(cudaMemsetAsync(st.dRunning.p, 0, 2 * sizeof(unsigned long long), st.sC));
    for (uint32_t k = 0; k < chunks; ++k)
    {
        CK(cudaMemcpyAsync(st.dCand.p + size_t(k) * E2E_CHUNK, candidates + size_t(k) * E2E_CHUNK,
                           size_t(chunkSize(k)) * sizeof(isaac_ext_candidate_t), cudaMemcpyHostToDevice, st.sH));
        if (trace) cudaEventRecord(tev[size_t(k) * passCount * 5 + 4], st.sH);
        CK(cudaEventRecord(st.evH[k], st.sH));
    }
    auto enqueue = [&](uint32_t w) -> int {
        const int b = int(w % NB);
        const uint32_t k = w / passCount, p = w % passCount;
        const uint32_t m = chunkSize(k);
        const bool gapped = passes[p].gapped;
        const uint32_t stride = gapped ? 32u : 3u;
        isaac_ext_candidate_t *dCand = st.dCand.p + size_t(k) * E2E_CHUNK;
        if (p == 0) CK(cudaStreamWaitEvent(st.sC, st.evH[k], 0));
        if (w >= uint32_t(NB)) CK(cudaStreamWaitEvent(st.sC, st.evD[b], 0));      // the result set is free once item w - NB left the device
        if (trace) cudaEventRecord(tev[size_t(w) * 5], st.sC);
        if (p == 0)
        {
            validateCandidatesKernel<<<gridFor(ctx, m, 256, 8), 256, 0, st.sC>>>(ctx->ref, ctx->reads, m, dCand, ctx->errorFlag.p);
            ++ctx->launches;
        }
        const int r = gapped ? isaac_ext_gapped_batch_device(ctx, m, dCand, stride, st.dFrag[b].p, st.dCig[b].p, nullptr, st.sC)
                             : isaac_ext_ungapped_batch_device(ctx, m, dCand, st.dFrag[b].p, st.dCig[b].p, nullptr, st.sC);
        if (r) return r;
        const uint32_t blocks = (m + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
        cigarBlockSumsKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dBlock[b].p);
        cigarScanBlockSumsKernel<<<1, 1024, 0, st.sC>>>(blocks, st.dBlock[b].p, st.dTotal[b].p, st.dRunning.p + p, st.dRunning.p + 2 + b,
                                                        st.hTotal.p + b);
        cigarCompactKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dCig[b].p, stride, st.dBlock[b].p, st.dPool[b].p,
                                                                   uint32_t(st.dPool[b].capacity), st.dRunning.p + 2 + b);
        ctx->launches += 3;
        CK(cudaGetLastError());
        if (trace) cudaEventRecord(tev[size_t(w) * 5 + 1], st.sC);
        CK(cudaEventRecord(st.evC[b], st.sC));
        return ISAAC_EXT_OK;
    };
    int rc = ISAAC_EXT_OK;
    for (uint32_t w = 0; w + 1 < uint32_t(NB) && w < items && !rc; ++w) rc = enqueue(w);
    for (uint32_t w = 0; w < items && !rc; ++w)
    {
        if (w + NB - 1 < items) { rc = enqueue(w + NB - 1); if (rc) break; }
        const int b = int(w % NB);
        const uint32_t k = w / passCount;
        E2ePass &pass = passes[w % passCount];
        const uint32_t m = chunkSize(k);
        CK(cudaEventSynchronize(st.evC[b]));
        const uint32_t words = st.hTotal.p[b];
        CK(cudaStreamWaitEvent(st.sD, st.evC[b], 0));
        if (trace) cudaEventRecord(tev[size_t(w) * 5 + 2], st.sD);
        if (pass.base + words > pass.poolCapacity || pass.base + words > 0xFFFFFFFFull) pass.overflow = true;
        else
        {
            CK(cudaMemcpyAsync(pass.fragmentsOut + size_t(k) * E2E_CHUNK, st.dFrag[b].p, size_t(m) * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, st.sD));
            if (words) CK(cudaMemcpyAsync(pass.poolOut + pass.base, st.dPool[b].p, size_t(words) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st.sD));
        }
        if (trace) cudaEventRecord(tev[size_t(w) * 5 + 3], st.sD);
        CK(cudaEventRecord(st.evD[b], st.sD));
        pass.base += words;
    }
    if (rc) { cudaDeviceSynchronize(); return rc; }
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, st.sD));
    CK(cudaStreamSynchronize(st.sD));
    if (trace)
    {
        for (uint32_t w = 0; w < items; ++w)
        {
            float t[5] = {0, 0, 0, 0, 0};
            for (int j = 0; j < (w % passCount == 0 ? 5 : 4); ++j) cudaEventElapsedTime(&t[j], tev.back(), tev[size_t(w) * 5 + j]);
            std::fprintf(stderr, "[isaac_ext] e2e %s chunk %2u: uploaded %7.3f, compute %7.3f .. %7.3f ms, copy out %7.3f .. %7.3f ms\n",
                         passes[w % passCount].gapped ? "gapped  " : "ungapped", w / passCount, t[4], t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t &e : tev) cudaEventDestroy(e);
    }
    bool overflow = false;
    for (unsigned p = 0; p < passCount; ++p) { *passes[p].wordsOut = passes[p].base; overflow |= passes[p].overflow; }
    if (flag)
    {
        cudaMemset(ctx->errorFlag.p, 0, sizeof(uint32_t));
        if (flag & 2u) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate refers to an unknown read or contig");
        if (flag & 4u) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate position outside [-readLength, contigLength]");
        return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR exceeded 32 operations");
    }
    if (overflow) return ctx->fail(ISAAC_EXT_E_CAPACITY, "CIGAR pool too small: *cigarWordsOut holds the required number of words");
    return ISAAC_EXT_OK;
}


/// isaac_ext_align_batch_packed: like extendCompact with both passes of a chunk in ONE item: ungapped, gapped, the acceptance rule,
/// the pool of the accepted gapped CIGARs, the 32-byte records; only those and the pool cross PCIe.
static int alignPacked(isaac_ext_ctx *ctx, E2eState &st, uint32_t n, const isaac_ext_candidate_t *candidates, isaac_ext_alignment_t *alignmentsOut,
                       uint32_t *poolOut, const uint64_t poolCapacity, uint64_t *wordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (wordsOut) *wordsOut = 0;
    if (!n) return ISAAC_EXT_OK;
    if (!candidates || !alignmentsOut || !poolOut || !wordsOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    const int NB = E2eState::NB;
    CK(cudaSetDevice(ctx->device));
    if (!st.ready)
    {
        CK(cudaStreamCreateWithFlags(&st.sH, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sC, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sD, cudaStreamNonBlocking));
        for (int i = 0; i < NB; ++i)
        {
            CK(cudaEventCreateWithFlags(&st.evC[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&st.evD[i], cudaEventDisableTiming));
        }
        CK(st.hTotal.reserve(NB));
        CK(st.dRunning.reserve(2 + NB));
        st.ready = true;
    }
    const uint32_t E2E_CHUNK = e2eChunk(), stride = 32u;
    const uint32_t chunkMax = std::min(n, E2E_CHUNK);
    const uint32_t blocksMax = (chunkMax + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
    CK(st.dCand.reserve(n));
    for (int i = 0; i < NB; ++i)
    {
        CK(st.dFrag[i].reserve(chunkMax)); CK(st.dCig[i].reserve(size_t(chunkMax) * stride));
        CK(st.dPool[i].reserve(size_t(chunkMax) * stride)); CK(st.dBlock[i].reserve(blocksMax)); CK(st.dTotal[i].reserve(1));
        CK(st.dFragU[i].reserve(chunkMax)); CK(st.dCigU[i].reserve(size_t(chunkMax) * 3)); CK(st.dStatus[i].reserve(chunkMax));
        CK(st.dPacked[i].reserve(chunkMax));
    }
    const uint32_t chunks = (n + E2E_CHUNK - 1) / E2E_CHUNK;
    while (st.evH.size() < chunks)
    {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st.evH.push_back(e);
    }
    auto chunkSize = [&](uint32_t k) { return std::min(E2E_CHUNK, n - k * E2E_CHUNK); };
    CK(cudaMemsetAsync(st.dRunning.p, 0, 2 * sizeof(unsigned long long), st.sC));
    for (uint32_t k = 0; k < chunks; ++k)
    {
        CK(cudaMemcpyAsync(st.dCand.p + size_t(k) * E2E_CHUNK, candidates + size_t(k) * E2E_CHUNK,
                           size_t(chunkSize(k)) * sizeof(isaac_ext_candidate_t), cudaMemcpyHostToDevice, st.sH));
        CK(cudaEventRecord(st.evH[k], st.sH));
    }
    auto enqueue = [&](uint32_t k) -> int {
        const int b = int(k % NB);
        const uint32_t m = chunkSize(k);
        isaac_ext_candidate_t *dCand = st.dCand.p + size_t(k) * E2E_CHUNK;
        CK(cudaStreamWaitEvent(st.sC, st.evH[k], 0));
        if (k >= uint32_t(NB)) CK(cudaStreamWaitEvent(st.sC, st.evD[b], 0));      // the result set is free once chunk k - NB left the device
        validateCandidatesKernel<<<gridFor(ctx, m, 256, 8), 256, 0, st.sC>>>(ctx->ref, ctx->reads, m, dCand, ctx->errorFlag.p);
        ++ctx->launches;
        int r = isaac_ext_ungapped_batch_device(ctx, m, dCand, st.dFragU[b].p, st.dCigU[b].p, nullptr, st.sC);
        if (!r) r = isaac_ext_gapped_batch_device(ctx, m, dCand, stride, st.dFrag[b].p, st.dCig[b].p, nullptr, st.sC);
        if (r) return r;
        const uint32_t blocks = (m + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
        selectAlignmentKernel<<<gridFor(ctx, m, 256, 8), 256, 0, st.sC>>>(ctx->reads, m, st.dFragU[b].p, st.dCigU[b].p, st.dFrag[b].p, st.dCig[b].p, stride,
                                                                          ctx->cfg.gappedMismatchesMax, st.dStatus[b].p);
        cigarBlockSumsKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dBlock[b].p);
        cigarScanBlockSumsKernel<<<1, 1024, 0, st.sC>>>(blocks, st.dBlock[b].p, st.dTotal[b].p, st.dRunning.p, st.dRunning.p + 2 + b, st.hTotal.p + b);
        cigarCompactKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dCig[b].p, stride, st.dBlock[b].p, st.dPool[b].p,
                                                                   uint32_t(st.dPool[b].capacity), st.dRunning.p + 2 + b);
        packAlignmentsKernel<<<gridFor(ctx, m, 256, 8), 256, 0, st.sC>>>(m, st.dFrag[b].p, st.dStatus[b].p, st.dPacked[b].p, ctx->errorFlag.p);
        ctx->launches += 5;
        CK(cudaGetLastError());
        CK(cudaEventRecord(st.evC[b], st.sC));
        return ISAAC_EXT_OK;
    };
    int rc = ISAAC_EXT_OK;
    uint64_t base = 0;
    bool overflow = false;
    for (uint32_t k = 0; k + 1 < uint32_t(NB) && k < chunks && !rc; ++k) rc = enqueue(k);
    for (uint32_t k = 0; k < chunks && !rc; ++k)
    {
        if (k + NB - 1 < chunks) { rc = enqueue(k + NB - 1); if (rc) break; }
        const int b = int(k % NB);
        const uint32_t m = chunkSize(k);
        CK(cudaEventSynchronize(st.evC[b]));
        const uint32_t words = st.hTotal.p[b];
        CK(cudaStreamWaitEvent(st.sD, st.evC[b], 0));
        if (base + words > poolCapacity || base + words > 0xFFFFFFFFull) overflow = true;
        else
        {
            CK(cudaMemcpyAsync(alignmentsOut + size_t(k) * E2E_CHUNK, st.dPacked[b].p, size_t(m) * sizeof(isaac_ext_alignment_t), cudaMemcpyDeviceToHost, st.sD));
            if (words) CK(cudaMemcpyAsync(poolOut + base, st.dPool[b].p, size_t(words) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st.sD));
        }
        CK(cudaEventRecord(st.evD[b], st.sD));
        base += words;
    }
    if (rc) { cudaDeviceSynchronize(); return rc; }
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, st.sD));
    CK(cudaStreamSynchronize(st.sD));
    *wordsOut = base;
    if (flag)
    {
        cudaMemset(ctx->errorFlag.p, 0, sizeof(uint32_t));
        if (flag & 2u) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate refers to an unknown read or contig");
        if (flag & 4u) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate position outside [-readLength, contigLength]");
        if (flag & 16u) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "a score does not fit the 32-byte record: use isaac_ext_extend_batch_compact");
        return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR exceeded 32 operations");
    }
    if (overflow) return ctx->fail(ISAAC_EXT_E_CAPACITY, "CIGAR pool too small: *cigarWordsOut holds the required number of words");
    return ISAAC_EXT_OK;
}
