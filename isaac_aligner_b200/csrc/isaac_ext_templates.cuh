// isaac_ext_build_templates: alignment::TemplateBuilder on top of the two batch calls of the hot path.
//
// The reference decides per cluster, while it walks the candidate lists, which orphans deserve a ShadowAligner::rescueShadow
// call.  Whether a call is made never depends on the outcome of another call (TemplateBuilder.cpp:519-525, 746-753: only on the
// orphan and on the best pair found among the seed candidates), so the same per-cluster code runs twice:
//   plan    every rescueShadow call is recorded as a request and answered "nothing rescued";
//   finish  after ONE isaac_ext_rescue_shadows batch over all requests of the tile, the calls are answered from its results in
//           the order they were recorded.
// Everything in between -- pair enumeration, the 1e-7-tolerant comparisons, exp / log10 / floor of the mapping scores, the
// std::sort + unique probability sums -- is restated line by line from lib/alignment/TemplateBuilder.cpp (cited below) on the
// host threads; it is list bookkeeping over a handful of fragments per cluster.  Included at the end of isaac_ext.cu.
#pragma once
#include <cfloat>
#include <cmath>

#include "kernels_clip.cuh"

#include "template_worker.cuh"

namespace
{

struct TemplateState
{
    HostBuffer<isaac_ext_fragment_t> buildFragments;  HostBuffer<uint64_t> buildBegin;  HostBuffer<uint32_t> buildCigars;
    std::vector<uint8_t> buildFlags;
    std::vector<uint64_t> clusterRequestBegin;
    HostBuffer<isaac_ext_template_t> templates;  HostBuffer<isaac_ext_fragment_t> fragments;  HostBuffer<uint32_t> cigars;
    // end clippers
    DeviceBuffer<isaac_ext_template_t> dTemplates;  DeviceBuffer<isaac_ext_fragment_t> dFragments;
    DeviceBuffer<uint32_t> dCigarsIn, dCigarsOut;
    HostBuffer<uint32_t> clippedCigars;
    // isaac_ext_template_stats
    DeviceBuffer<isaac_ext_template_t> dStatTemplates;  DeviceBuffer<isaac_ext_fragment_t> dStatFragments;
    DeviceBuffer<uint32_t> dStatCigars;  DeviceBuffer<uint8_t> dStatBytes;  DeviceBuffer<unsigned long long> dStats;
    ~TemplateState()
    {
        dTemplates.release(); dFragments.release(); dCigarsIn.release(); dCigarsOut.release();
        dStatTemplates.release(); dStatFragments.release(); dStatCigars.release(); dStatBytes.release(); dStats.release();
    }
};

template <class T> void swapBuffers(HostBuffer<T> &a, HostBuffer<T> &b) { std::swap(a.p, b.p); std::swap(a.capacity, b.capacity); }

} // namespace

void releaseTemplates(TemplateState *state) { delete state; }

extern "C" int isaac_ext_build_templates(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                         const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!batch || !tls || !options || !result) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->templates) ctx->templates = new TemplateState();
    TemplateState &st = *ctx->templates;
    PhaseTimer timer("templates");

    // ---- the seed candidates of every cluster (FragmentBuilder::build)
    isaac_ext_build_result_t built;
    int rc = isaac_ext_build_fragments(ctx, batch, &built);
    if (rc) return rc;
    // keep this result alive across the rescue call: the pipeline gets the buffers of the previous tile in exchange
    PipelineState &ps = ctx->pipeline;
    swapBuffers(st.buildFragments, ps.outFragments); swapBuffers(st.buildBegin, ps.outBegin); swapBuffers(st.buildCigars, ps.outCigars);
    st.buildFlags.swap(ps.outFlags);
    timer.mark("build_fragments");

    const uint32_t n = ctx->clusterCount, readCount = ctx->reads.readCount;
    const unsigned T = ctx->hostThreads;
    const TemplateContext cx = makeTemplateContext(*tls, *options, ctx->contigLength, readCount, ctx->reads.readLength, ctx->logMismatchQ40);
    auto loadCluster = [&](TemplateWorker &w, uint32_t c) { loadClusterFragments(w, built, st.buildFlags[c] != 0, readCount, c); };

    // ---- plan / rescue / finish, software-pipelined over slices of the tile: while the GPU answers the rescueShadow calls of
    // slice s (one isaac_ext_rescue_shadows batch, on a helper thread), the host threads finish slice s - 1 and plan slice s + 1.
    // Clusters are independent, so slicing changes nothing in the results.
    // (measured on 500 k pairs and 16 host threads: 109 ms in one piece, 103 ms in two slices, 107 ms in four)
    unsigned slices = n >= 100000 ? 2u : 1u;
    if (const char *e = std::getenv("ISAAC_EXT_TEMPLATE_SLICES")) slices = unsigned(std::max(1, std::min(16, std::atoi(e))));
    slices = std::max(1u, std::min(slices, std::max(1u, n)));
    auto sliceBegin = [&](unsigned s) { return uint32_t(uint64_t(n) * s / slices); };
    st.templates.reserve(n); st.fragments.reserve(size_t(n) * readCount);
    st.clusterRequestBegin.assign(size_t(n) + 1, 0);
    struct Slice
    {
        std::vector<isaac_ext_rescue_request_t> requests;
        isaac_ext_rescue_result_t rescued;
        int rc = ISAAC_EXT_OK;
        std::thread worker;
    };
    std::vector<Slice> slice(slices);
    struct CigarPart { std::vector<uint32_t> words; size_t firstCluster, endCluster; };
    std::vector<CigarPart> cigarParts;
    uint64_t requestTotal = 0;

    auto planSlice = [&](unsigned s) {
        const uint32_t cb = sliceBegin(s), ce = sliceBegin(s + 1);
        const unsigned parts = partitionCount(T, ce - cb);
        std::vector<std::vector<isaac_ext_rescue_request_t>> partRequests(parts);
        parallelRanges(T, ce - cb, [&](unsigned t, size_t b, size_t e) {
            TemplateWorker w(cx);
            w.planning = true; w.requests = &partRequests[t];
            for (size_t c = cb + b; c < cb + e; ++c)
            {
                const size_t before = partRequests[t].size();
                if (st.buildFlags[c]) { loadCluster(w, uint32_t(c)); w.run(); }
                st.clusterRequestBegin[c + 1] = partRequests[t].size() - before;     // per cluster for now, offsets below
            }
        });
        uint64_t at = 0;                                                             // offsets within the slice's own batch
        for (size_t c = cb; c < ce; ++c) { const uint64_t k = st.clusterRequestBegin[c + 1]; st.clusterRequestBegin[c + 1] = at; at += k; }
        slice[s].requests.clear();
        for (const std::vector<isaac_ext_rescue_request_t> &p : partRequests) slice[s].requests.insert(slice[s].requests.end(), p.begin(), p.end());
        requestTotal += slice[s].requests.size();
    };
    auto startRescue = [&](unsigned s) {
        Slice &sl = slice[s];
        std::memset(&sl.rescued, 0, sizeof(sl.rescued));
        if (sl.requests.empty()) return;
        sl.worker = std::thread([&sl, ctx, tls, s] {
            sl.rc = rescueShadowsInto(ctx, tls, uint32_t(sl.requests.size()), sl.requests.data(), &sl.rescued, s & 1u);
        });
    };
    auto finishSlice = [&](unsigned s) {
        Slice &sl = slice[s];
        const uint32_t cb = sliceBegin(s), ce = sliceBegin(s + 1);
        const unsigned parts = partitionCount(T, ce - cb);
        const size_t firstPart = cigarParts.size();
        cigarParts.resize(firstPart + parts);
        for (unsigned t = 0; t < parts; ++t) cigarParts[firstPart + t].firstCluster = cigarParts[firstPart + t].endCluster = ce;
        parallelRanges(T, ce - cb, [&](unsigned t, size_t b, size_t e) {
            TemplateWorker w(cx);
            w.planning = false; w.rescueResult = &sl.rescued;
            CigarPart &part = cigarParts[firstPart + t];
            std::vector<uint32_t> &pool = part.words;
            part.firstCluster = cb + b; part.endCluster = cb + e;
            for (size_t c = cb + b; c < cb + e; ++c)
            {
                bool ok = false;
                const bool hadFragments = st.buildFlags[c] != 0;
                if (hadFragments)
                {
                    loadCluster(w, uint32_t(c));
                    w.nextRequest = st.clusterRequestBegin[c + 1];                   // offset of the cluster's first call in the slice's batch
                    ok = w.run();
                }
                else
                {
                    resetToUnaligned(w, uint32_t(c), readCount);
                }
                storeTemplate(w, ok, hadFragments, uint32_t(c), readCount, st.templates.p[c], st.fragments.p + c * readCount, pool);   // cigarOffset relative to this part, rebased below
            }
        });
    };
    auto joinRescue = [&](unsigned s) -> int {
        if (slice[s].worker.joinable()) slice[s].worker.join();
        return slice[s].rc;
    };

    planSlice(0);
    startRescue(0);
    for (unsigned s = 0; s < slices; ++s)
    {
        if (s + 1 < slices) planSlice(s + 1);                       // host, while the GPU works on slice s
        rc = joinRescue(s);
        if (rc) { for (unsigned k = s + 1; k < slices; ++k) joinRescue(k); return rc; }
        if (s + 1 < slices) startRescue(s + 1);                     // its result set (s + 1) & 1 was last read by finishSlice(s - 1)
        finishSlice(s);                                             // host, while the GPU works on slice s + 1
    }
    timer.mark("plan / rescue_shadows / finish");

    uint64_t words = 0;
    for (CigarPart &part : cigarParts) { const uint64_t k = part.words.size(); part.endCluster = std::max(part.endCluster, part.firstCluster); words += k; }
    st.cigars.reserve(words);
    {
        std::vector<uint64_t> partBase(cigarParts.size(), 0);
        uint64_t at = 0;
        for (size_t k = 0; k < cigarParts.size(); ++k) { partBase[k] = at; at += cigarParts[k].words.size(); }
        parallelRanges(T, cigarParts.size(), [&](unsigned, size_t b, size_t e) {
            for (size_t k = b; k < e; ++k)
            {
                const CigarPart &part = cigarParts[k];
                if (!part.words.empty()) std::memcpy(st.cigars.p + partBase[k], part.words.data(), part.words.size() * sizeof(uint32_t));
                for (size_t i = part.firstCluster * readCount; i < part.endCluster * readCount; ++i) st.fragments.p[i].cigarOffset += uint32_t(partBase[k]);
            }
        });
    }
    timer.mark("cigar pool");

    result->templates = st.templates.p; result->fragments = st.fragments.p; result->cigars = st.cigars.p;
    result->cigarWords = words; result->rescueRequests = requestTotal;

    // ---- end clippers on the kept templates (MatchSelector.cpp:336-346): one kernel pass over the tile
    if (options->clipFlags & (ISAAC_EXT_CLIP_SEMIALIGNED | ISAAC_EXT_CLIP_OVERLAPPING))
    {
        const size_t count = size_t(n) * readCount;
        const size_t outWords = words + 4 * count;                       // a clip adds at most two operations per side
        if (outWords > 0xFFFFFFFFull) return ctx->fail(ISAAC_EXT_E_CAPACITY, "CIGAR pool of the tile exceeds 2^32 words");
        CK(cudaSetDevice(ctx->device));
        CK(st.dTemplates.reserve(n)); CK(st.dFragments.reserve(count)); CK(st.dCigarsIn.reserve(words + 1)); CK(st.dCigarsOut.reserve(outWords));
        st.clippedCigars.reserve(outWords);
        CK(cudaMemcpyAsync(st.dTemplates.p, st.templates.p, size_t(n) * sizeof(isaac_ext_template_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(st.dFragments.p, st.fragments.p, count * sizeof(isaac_ext_fragment_t), cudaMemcpyHostToDevice, ctx->stream));
        if (words) CK(cudaMemcpyAsync(st.dCigarsIn.p, st.cigars.p, words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(st.dCigarsOut.p, 0, outWords * sizeof(uint32_t), ctx->stream));
        clipTemplateEndsKernel<<<gridFor(ctx, n, 128, 16), 128, 0, ctx->stream>>>(
            ctx->ref, ctx->reads, n, options->clipFlags, st.dTemplates.p, st.dFragments.p, st.dCigarsIn.p, st.dCigarsOut.p, ctx->errorFlag.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(st.fragments.p, st.dFragments.p, count * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(st.clippedCigars.p, st.dCigarsOut.p, outWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        uint32_t flag = 0;
        CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (flag) { cudaMemset(ctx->errorFlag.p, 0, sizeof(uint32_t)); return ctx->fail(ISAAC_EXT_E_CAPACITY, "a template CIGAR exceeds 60 operations"); }
        result->cigars = st.clippedCigars.p; result->cigarWords = outWords;
        timer.mark("end clippers");
    }
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_trim_low_quality_ends(isaac_ext_ctx *ctx, uint32_t baseQualityCutoff, uint16_t *endCyclesMaskedOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reads first");
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->reads.readTotal;
    trimLowQualityEndsKernel<<<gridFor(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->reads, baseQualityCutoff, ctx->readMasked.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (endCyclesMaskedOut) CK(cudaMemcpyAsync(endCyclesMaskedOut, ctx->readMasked.p, size_t(n) * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ISAAC_EXT_OK;
}
