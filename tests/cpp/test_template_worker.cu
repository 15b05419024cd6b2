// template_worker.cuh on the CPU: the plan and finish passes of isaac_ext_build_templates (isaac_ext_templates.cuh) around build
// and rescue results that the CHECKER supplies (tests/test_template_worker.py: oracle_build_fragments / oracle_rescue_shadows of
// the reference build), so that the host half of the product's TemplateBuilder is compared with the reference's own
// TemplateBuilder without a GPU.  Host code of an nvcc-compiled shared library, no CUDA call.  TEST CODE, not a product path.
#include <cstring>
#include <vector>

#include "../../include/isaac_ext.h"
#include "../../isaac_aligner_b200/csrc/host_pipeline.cuh"
using namespace isaac_b200;
#include "template_worker.cuh"
#include "../../isaac_aligner_b200/csrc/plan_device.cuh"
#include "../../isaac_aligner_b200/csrc/shadow_window_device.cuh"
#include "../../isaac_aligner_b200/csrc/finish_device.cuh"

namespace
{
double logMismatchQ40() { return std::log(std::pow(10.0, 40.0 / -10.0) / 3.0); }          // Quality.cpp:34-66, as isaac_ext_create computes it
}

/// plan: the rescueShadow calls every cluster makes, in cluster order.  clusterRequestBegin: clusterCount + 1 offsets.
extern "C" int template_worker_plan(uint32_t clusterCount, uint32_t readCount, const uint32_t *readLength, uint32_t contigCount,
                                    const uint64_t *contigLength, const isaac_ext_tls_t *tls, const isaac_ext_template_options_t *options,
                                    const isaac_ext_build_result_t *built, uint64_t requestCapacity, isaac_ext_rescue_request_t *requestsOut,
                                    uint64_t *clusterRequestBegin, unsigned threads)
{
    const std::vector<uint64_t> lengths(contigLength, contigLength + contigCount);
    const uint32_t rl[2] = {readLength[0], readCount > 1 ? readLength[1] : 0};
    const TemplateContext cx = makeTemplateContext(*tls, *options, lengths, readCount, rl, logMismatchQ40());
    const unsigned parts = partitionCount(threads, clusterCount);
    std::vector<std::vector<isaac_ext_rescue_request_t>> partRequests(parts);
    parallelRanges(threads, clusterCount, [&](unsigned t, size_t b, size_t e) {
        TemplateWorker w(cx);
        w.planning = true; w.requests = &partRequests[t];
        for (size_t c = b; c < e; ++c)
        {
            const size_t before = partRequests[t].size();
            if (built->built[c]) { loadClusterFragments(w, *built, true, readCount, uint32_t(c)); w.run(); }
            clusterRequestBegin[c + 1] = partRequests[t].size() - before;
        }
    });
    uint64_t at = 0;
    clusterRequestBegin[0] = 0;
    for (size_t c = 0; c < clusterCount; ++c) { const uint64_t k = clusterRequestBegin[c + 1]; at += k; clusterRequestBegin[c + 1] = at; }
    if (at > requestCapacity) return ISAAC_EXT_E_CAPACITY;
    uint64_t o = 0;
    for (const std::vector<isaac_ext_rescue_request_t> &p : partRequests) for (const isaac_ext_rescue_request_t &q : p) requestsOut[o++] = q;
    return ISAAC_EXT_OK;
}

/// finish: the templates, the rescueShadow calls answered from 'rescued' in the order plan recorded them
extern "C" int template_worker_finish(uint32_t clusterCount, uint32_t readCount, const uint32_t *readLength, uint32_t contigCount,
                                      const uint64_t *contigLength, const isaac_ext_tls_t *tls, const isaac_ext_template_options_t *options,
                                      const isaac_ext_build_result_t *built, const isaac_ext_rescue_result_t *rescued,
                                      const uint64_t *clusterRequestBegin, isaac_ext_template_t *templatesOut,
                                      isaac_ext_fragment_t *fragmentsOut, uint64_t cigarCapacity, uint32_t *cigarsOut, uint64_t *cigarWordsOut,
                                      unsigned threads)
{
    const std::vector<uint64_t> lengths(contigLength, contigLength + contigCount);
    const uint32_t rl[2] = {readLength[0], readCount > 1 ? readLength[1] : 0};
    const TemplateContext cx = makeTemplateContext(*tls, *options, lengths, readCount, rl, logMismatchQ40());
    const unsigned parts = partitionCount(threads, clusterCount);
    std::vector<std::vector<uint32_t>> pools(parts);
    std::vector<size_t> partBegin(parts, clusterCount), partEnd(parts, clusterCount);
    parallelRanges(threads, clusterCount, [&](unsigned t, size_t b, size_t e) {
        TemplateWorker w(cx);
        w.planning = false; w.rescueResult = rescued;
        partBegin[t] = b; partEnd[t] = e;
        for (size_t c = b; c < e; ++c)
        {
            bool ok = false;
            const bool hadFragments = built->built[c] != 0;
            if (hadFragments)
            {
                loadClusterFragments(w, *built, true, readCount, uint32_t(c));
                w.nextRequest = clusterRequestBegin[c];
                ok = w.run();
            }
            else resetToUnaligned(w, uint32_t(c), readCount);
            storeTemplate(w, ok, hadFragments, uint32_t(c), readCount, templatesOut[c], fragmentsOut + c * readCount, pools[t]);
        }
    });
    uint64_t words = 0;
    for (unsigned t = 0; t < parts; ++t)
    {
        if (words + pools[t].size() > cigarCapacity) return ISAAC_EXT_E_CAPACITY;
        if (!pools[t].empty()) std::memcpy(cigarsOut + words, pools[t].data(), pools[t].size() * sizeof(uint32_t));
        for (size_t c = partBegin[t]; c < partEnd[t]; ++c)
            for (unsigned r = 0; r < readCount; ++r) fragmentsOut[c * readCount + r].cigarOffset += uint32_t(words);
        words += pools[t].size();
    }
    *cigarWordsOut = words;
    return ISAAC_EXT_OK;
}

/// the same request list from plan_device.cuh (the libm-free plan pass written for device code), cluster after cluster
extern "C" int plan_device_requests(uint32_t clusterCount, uint32_t readCount, const isaac_ext_tls_t *tls, const isaac_ext_template_options_t *options,
                                    const isaac_ext_build_result_t *built, uint64_t requestCapacity, isaac_ext_rescue_request_t *requestsOut,
                                    uint64_t *clusterRequestBegin)
{
    const size_t lists = size_t(clusterCount) * readCount;
    std::vector<uint32_t> listBegin(lists), listCount(lists);
    for (size_t l = 0; l < lists; ++l) { listBegin[l] = uint32_t(built->readFragmentBegin[l]); listCount[l] = uint32_t(built->readFragmentBegin[l + 1] - built->readFragmentBegin[l]); }
    PlanView v;
    v.fragments = built->fragments; v.listBegin = listBegin.data(); v.listCount = listCount.data(); v.built = built->built; v.readCount = readCount;
    v.tlsMax = tls->max; v.bestModel[0] = tls->bestModel[0]; v.bestModel[1] = tls->bestModel[1]; v.scatterRepeats = options->scatterRepeats;
    uint64_t at = 0;
    clusterRequestBegin[0] = 0;
    for (uint32_t c = 0; c < clusterCount; ++c)
    {
        const unsigned room = at < requestCapacity ? unsigned(std::min<uint64_t>(requestCapacity - at, 1u << 30)) : 0u;
        at += planClusterRequests(v, c, requestsOut + at, room);
        clusterRequestBegin[c + 1] = at;
    }
    return at > requestCapacity ? ISAAC_EXT_E_CAPACITY : ISAAC_EXT_OK;
}

/// shadow_window_device.cuh: the scan window of every request; tasksOut: 4 x int64 per request = windowBegin, windowEnd, shadowReadId,
/// contigStrand; rangeOut: first / second of calculateShadowRescueRange
extern "C" int shadow_windows_device(const uint32_t *readLength, const isaac_ext_tls_t *tls, uint32_t requestCount,
                                     const isaac_ext_rescue_request_t *requests, const uint64_t *contigLength, int64_t *tasksOut, int64_t *rangeOut)
{
    struct Task { int64_t windowBegin, windowEnd; uint32_t shadowReadId, contigStrand; };
    const ShadowWindowModel m = makeShadowWindowModel(*tls);
    for (uint32_t i = 0; i < requestCount; ++i)
    {
        Task t;
        shadowWindowOf(m, requests[i], readLength, long(contigLength[requests[i].orphanContigStrand >> 1]), t);
        tasksOut[4 * size_t(i)] = t.windowBegin; tasksOut[4 * size_t(i) + 1] = t.windowEnd;
        tasksOut[4 * size_t(i) + 2] = t.shadowReadId; tasksOut[4 * size_t(i) + 3] = t.contigStrand;
        long first, second;
        shadowRescueRange(m, requests[i], readLength, first, second);
        rangeOut[2 * size_t(i)] = first; rangeOut[2 * size_t(i) + 1] = second;
    }
    return shadowModelCoherent(m) ? 0 : 1;
}

/// finish_device.cuh (the finish pass written for a one-thread-per-cluster kernel: glibc's exp / log10 replayed, no containers)
/// cluster after cluster, on the same inputs as template_worker_finish; the CIGAR words are gathered the way
/// gatherTemplateCigarsKernel gathers them
extern "C" int finish_device_templates(uint32_t clusterCount, uint32_t readCount, const uint32_t *readLength, uint32_t contigCount,
                                       const uint64_t *contigLength, const isaac_ext_tls_t *tls, const isaac_ext_template_options_t *options,
                                       const isaac_ext_build_result_t *built, const isaac_ext_rescue_result_t *rescued,
                                       const uint64_t *clusterRequestBegin, isaac_ext_template_t *templatesOut,
                                       isaac_ext_fragment_t *fragmentsOut, uint64_t cigarCapacity, uint32_t *cigarsOut, uint64_t *cigarWordsOut)
{
    const size_t lists = size_t(clusterCount) * readCount;
    std::vector<uint32_t> listBegin(lists), listCount(lists), requestBegin(size_t(clusterCount) + 1);
    for (size_t l = 0; l < lists; ++l) { listBegin[l] = uint32_t(built->readFragmentBegin[l]); listCount[l] = uint32_t(built->readFragmentBegin[l + 1] - built->readFragmentBegin[l]); }
    for (size_t c = 0; c <= clusterCount; ++c) requestBegin[c] = uint32_t(clusterRequestBegin[c]);
    FinishView v;
    v.fragments = built->fragments; v.listBegin = listBegin.data(); v.listCount = listCount.data(); v.built = built->built;
    v.cigarPools[0] = built->cigars; v.cigarPools[1] = v.cigarPools[2] = nullptr; v.cigarPools[FINISH_POOL_RESCUE] = rescued->cigars;
    v.rescueFragments = rescued->fragments; v.requestFragmentBegin = rescued->requestFragmentBegin; v.rescued = rescued->rescued;
    v.clusterRequestBegin = requestBegin.data();
    v.readCount = readCount; v.tlsMin = tls->min; v.tlsMax = tls->max; v.bestModel[0] = tls->bestModel[0]; v.bestModel[1] = tls->bestModel[1];
    v.scatterRepeats = options->scatterRepeats; v.mapqThreshold = options->mapqThreshold; v.dodgyAlignmentScore = options->dodgyAlignmentScore;
    v.logMismatchQ40 = logMismatchQ40();
    const uint32_t rl[2] = {readLength[0], readCount > 1 ? readLength[1] : 0};
    finishRestOfGenome(v, contigLength, contigCount, rl);
    std::vector<unsigned char> scratch;
    uint64_t words = 0;
    for (uint32_t c = 0; c < clusterCount; ++c)
    {
        const uint64_t shadows = rescued->requestFragmentBegin[requestBegin[c + 1]] - rescued->requestFragmentBegin[requestBegin[c]];
        uint64_t candidates = 0;
        for (unsigned r = 0; r < readCount; ++r) candidates += listCount[size_t(c) * readCount + r];
        scratch.assign(finishScratchBytes(shadows, candidates), 0xA5);
        finishCluster(v, c, scratch.data(), shadows, candidates, templatesOut[c], fragmentsOut + size_t(c) * readCount);
        for (unsigned r = 0; r < readCount; ++r)
        {
            isaac_ext_fragment_t &f = fragmentsOut[size_t(c) * readCount + r];
            const uint32_t *src = f.cigarLength ? v.cigarPools[f.cigarOffset >> FINISH_POOL_SHIFT] + (f.cigarOffset & FINISH_POOL_MASK) : nullptr;
            if (words + f.cigarLength > cigarCapacity) return ISAAC_EXT_E_CAPACITY;
            for (unsigned k = 0; k < f.cigarLength; ++k) cigarsOut[words + k] = src[k];
            f.cigarOffset = uint32_t(words);
            words += f.cigarLength;
        }
    }
    *cigarWordsOut = words;
    return ISAAC_EXT_OK;
}
