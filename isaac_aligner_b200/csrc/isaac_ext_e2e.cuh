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
    cudaStream_t sH = nullptr, sC = nullptr, sD = nullptr;
    cudaEvent_t evH[2] = {nullptr, nullptr}, evC[2] = {nullptr, nullptr}, evD[2] = {nullptr, nullptr};
    DeviceBuffer<isaac_ext_candidate_t> dCand[2];
    DeviceBuffer<isaac_ext_fragment_t> dFrag[2];
    DeviceBuffer<uint32_t> dCig[2], dPool[2], dBlock[2], dTotal[2];
    PinnedBuffer<uint32_t> hTotal;
    bool ready = false;
    void release()
    {
        for (int i = 0; i < 2; ++i)
        {
            dCand[i].release(); dFrag[i].release(); dCig[i].release(); dPool[i].release(); dBlock[i].release(); dTotal[i].release();
            if (evH[i]) cudaEventDestroy(evH[i]);
            if (evC[i]) cudaEventDestroy(evC[i]);
            if (evD[i]) cudaEventDestroy(evD[i]);
        }
        if (sH) cudaStreamDestroy(sH);
        if (sC) cudaStreamDestroy(sC);
        if (sD) cudaStreamDestroy(sD);
        hTotal.release();
        ready = false;
    }
};

} // namespace

static int extendCompact(isaac_ext_ctx *ctx, E2eState &st, bool gapped, uint32_t n, const isaac_ext_candidate_t *candidates,
                         isaac_ext_fragment_t *fragmentsOut, uint32_t *poolOut, uint64_t poolCapacity, uint64_t *wordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (wordsOut) *wordsOut = 0;
    if (!n) return ISAAC_EXT_OK;
    if (!candidates || !fragmentsOut || !poolOut || !wordsOut) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    int rc = ISAAC_EXT_OK;
    CK(cudaSetDevice(ctx->device));
    if (!st.ready)
    {
        CK(cudaStreamCreateWithFlags(&st.sH, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sC, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&st.sD, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i)
        {
            CK(cudaEventCreateWithFlags(&st.evH[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&st.evC[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&st.evD[i], cudaEventDisableTiming));
        }
        CK(st.hTotal.reserve(2));
        st.ready = true;
    }
    const uint32_t E2E_CHUNK = e2eChunk();
    const uint32_t stride = gapped ? 32u : 3u;
    const uint32_t chunkMax = std::min(n, E2E_CHUNK);
    const uint32_t blocksMax = (chunkMax + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
    for (int i = 0; i < 2; ++i)
    {
        CK(st.dCand[i].reserve(chunkMax)); CK(st.dFrag[i].reserve(chunkMax)); CK(st.dCig[i].reserve(size_t(chunkMax) * stride));
        CK(st.dPool[i].reserve(size_t(chunkMax) * stride)); CK(st.dBlock[i].reserve(blocksMax)); CK(st.dTotal[i].reserve(1));
    }
    const uint32_t chunks = (n + E2E_CHUNK - 1) / E2E_CHUNK;
    auto chunkSize = [&](uint32_t k) { return std::min(E2E_CHUNK, n - k * E2E_CHUNK); };
    // ISAAC_EXT_TRACE: device-side time stamps of every chunk (compute begin/end on sC, copy-out begin/end on sD)
    const bool trace = std::getenv("ISAAC_EXT_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) { tev.resize(size_t(chunks) * 4 + 1); for (cudaEvent_t &e : tev) cudaEventCreate(&e); cudaEventRecord(tev.back(), st.sH); }
    auto enqueue = [&](uint32_t k) -> int {
        const int b = k & 1;
        const uint32_t m = chunkSize(k);
        // the host checks a chunk while the device works on the previous ones; a bad candidate fails the whole call
        const int bad = validateCandidates(ctx, m, candidates + size_t(k) * E2E_CHUNK);
        if (bad) { cudaDeviceSynchronize(); return bad; }
        if (k >= 2) CK(cudaStreamWaitEvent(st.sH, st.evD[b], 0));          // the buffer set is free once chunk k-2 left the device
        CK(cudaMemcpyAsync(st.dCand[b].p, candidates + size_t(k) * E2E_CHUNK, size_t(m) * sizeof(isaac_ext_candidate_t), cudaMemcpyHostToDevice, st.sH));
        CK(cudaEventRecord(st.evH[b], st.sH));
        CK(cudaStreamWaitEvent(st.sC, st.evH[b], 0));
        if (trace) cudaEventRecord(tev[size_t(k) * 4], st.sC);
        const int r = gapped ? isaac_ext_gapped_batch_device(ctx, m, st.dCand[b].p, stride, st.dFrag[b].p, st.dCig[b].p, nullptr, st.sC)
                             : isaac_ext_ungapped_batch_device(ctx, m, st.dCand[b].p, st.dFrag[b].p, st.dCig[b].p, nullptr, st.sC);
        if (r) return r;
        const uint32_t blocks = (m + COMPACT_BLOCK * COMPACT_ITEMS - 1) / (COMPACT_BLOCK * COMPACT_ITEMS);
        cigarBlockSumsKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dBlock[b].p);
        cigarScanBlockSumsKernel<<<1, 1024, 0, st.sC>>>(blocks, st.dBlock[b].p, st.dTotal[b].p);
        cigarCompactKernel<<<blocks, COMPACT_BLOCK, 0, st.sC>>>(m, st.dFrag[b].p, st.dCig[b].p, stride, st.dBlock[b].p, st.dPool[b].p,
                                                                   uint32_t(st.dPool[b].capacity));
        ctx->launches += 3;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(st.hTotal.p + b, st.dTotal[b].p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st.sC));
        if (trace) cudaEventRecord(tev[size_t(k) * 4 + 1], st.sC);
        CK(cudaEventRecord(st.evC[b], st.sC));
        return ISAAC_EXT_OK;
    };
    uint64_t base = 0;
    bool overflow = false;
    rc = enqueue(0);
    if (rc) return rc;
    for (uint32_t k = 0; k < chunks; ++k)
    {
        if (k + 1 < chunks) { rc = enqueue(k + 1); if (rc) return rc; }
        const int b = k & 1;
        const uint32_t m = chunkSize(k);
        CK(cudaEventSynchronize(st.evC[b]));
        const uint32_t words = st.hTotal.p[b];
        CK(cudaStreamWaitEvent(st.sD, st.evC[b], 0));
        if (trace) cudaEventRecord(tev[size_t(k) * 4 + 2], st.sD);
        if (base + words > poolCapacity || base + words > 0xFFFFFFFFull) overflow = true;
        else
        {
            if (base)
            {
                addCigarBaseKernel<<<gridFor(ctx, m, 256, 8), 256, 0, st.sD>>>(m, st.dFrag[b].p, uint32_t(base));
                ++ctx->launches;
            }
            CK(cudaMemcpyAsync(fragmentsOut + size_t(k) * E2E_CHUNK, st.dFrag[b].p, size_t(m) * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, st.sD));
            if (words) CK(cudaMemcpyAsync(poolOut + base, st.dPool[b].p, size_t(words) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st.sD));
        }
        if (trace) cudaEventRecord(tev[size_t(k) * 4 + 3], st.sD);
        CK(cudaEventRecord(st.evD[b], st.sD));
        base += words;
    }
    CK(cudaStreamSynchronize(st.sD));
    if (trace)
    {
        for (uint32_t k = 0; k < chunks; ++k)
        {
            float t[4];
            for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev.back(), tev[size_t(k) * 4 + j]);
            std::fprintf(stderr, "[isaac_ext] e2e %s chunk %2u: compute %7.3f .. %7.3f ms, copy out %7.3f .. %7.3f ms\n",
                         gapped ? "gapped" : "ungapped", k, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t &e : tev) cudaEventDestroy(e);
    }
    *wordsOut = base;
    if (overflow) return ctx->fail(ISAAC_EXT_E_CAPACITY, "CIGAR pool too small: *cigarWordsOut holds the required number of words");
    if (gapped)
    {
        uint32_t flag = 0;
        CK(cudaMemcpy(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost));
        if (flag) { cudaMemset(ctx->errorFlag.p, 0, sizeof(uint32_t)); return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR exceeded 32 operations"); }
    }
    return ISAAC_EXT_OK;
}
