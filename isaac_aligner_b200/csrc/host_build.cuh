// Buffers of the rescue pass and the grow-only host array of the context.
// Included by isaac_ext.cu after the context definition.
#pragma once
#include "host_pipeline.cuh"

namespace isaac_b200
{

/// grow-only host array that is never zero-filled (std::vector::resize would touch hundreds of MB per call)
template <class T> struct HostBuffer
{
    T *p = nullptr; size_t capacity = 0;
    void reserve(size_t n)
    {
        if (n <= capacity) return;
        std::free(p);
        capacity = n + n / 4 + 1024;
        p = static_cast<T *>(std::malloc(capacity * sizeof(T)));
    }
    ~HostBuffer() { std::free(p); }
};

/// Buffers of the rescue pass (isaac_ext_rescue_shadows and the rescue step of isaac_ext_build_templates), owned by the context and
/// reused across calls; the build pass has its own in TileState (isaac_ext_tile.cuh).
struct PipelineState
{
    DeviceBuffer<isaac_ext_candidate_t> dCand, dCand3, dAdapterFirst;
    DeviceBuffer<isaac_ext_fragment_t> dFrag, dFrag3, dOutFragments;
    DeviceBuffer<uint32_t> dCig, dCig3, dOutCigars;
    DeviceBuffer<ShadowTask> dShadowTasks;
    DeviceBuffer<int> dShadowScratch;
    DeviceBuffer<uint32_t> dTaskBegin, dTaskCount, dLargeTasks, dSlot;
    DeviceBuffer<unsigned long long> dPoolSize;
    // the list bookkeeping on the device (kernels_rescue.cuh)
    DeviceBuffer<uint32_t> dKept, dAdoptedBy, dCounts, dBegins, dSlot3, dSources;
    DeviceBuffer<ShadowListState> dListState;
    DeviceBuffer<uint8_t> dRescued, dScanTemp;
    DeviceBuffer<uint64_t> dOutBegin;
    // the flat result of isaac_ext_rescue_shadows in page-locked memory
    PinnedBuffer<uint32_t> hTotals, hOutCigars[2];
    PinnedBuffer<isaac_ext_fragment_t> hOutFragments[2];
    PinnedBuffer<uint64_t> hOutBegin[2];
    PinnedBuffer<uint8_t> hRescued[2];

    void release()
    {
        dCand.release(); dCand3.release(); dAdapterFirst.release(); dFrag.release(); dFrag3.release(); dOutFragments.release();
        dCig.release(); dCig3.release(); dOutCigars.release(); dShadowTasks.release(); dShadowScratch.release(); dTaskBegin.release();
        dTaskCount.release(); dLargeTasks.release(); dSlot.release(); dPoolSize.release(); dKept.release(); dAdoptedBy.release();
        dCounts.release(); dBegins.release(); dSlot3.release(); dSources.release(); dListState.release(); dRescued.release();
        dScanTemp.release(); dOutBegin.release(); hTotals.release();
        for (int k = 0; k < 2; ++k) { hOutCigars[k].release(); hOutFragments[k].release(); hOutBegin[k].release(); hRescued[k].release(); }
    }
};

} // namespace isaac_b200
