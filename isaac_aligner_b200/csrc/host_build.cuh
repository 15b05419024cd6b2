// isaac_ext_build_fragments / isaac_ext_rescue_shadows: the phase drivers (see host_pipeline.cuh for the plan).
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

/// Buffers of the two pipelines, owned by the context and reused across calls.
struct PipelineState
{
    // host (pinned): kernel inputs and outputs of each pass
    PinnedBuffer<isaac_ext_candidate_t> hCand1, hCand3;
    PinnedBuffer<isaac_ext_fragment_t> hFrag1, hFrag3;
    PinnedBuffer<uint32_t> hCig1, hCig3;
    PinnedBuffer<IndelTask> hTasks;
    PinnedBuffer<IndelResult> hIndel;
    PinnedBuffer<ShadowTask> hShadowTasks;
    // device
    DeviceBuffer<isaac_ext_candidate_t> dCand;
    DeviceBuffer<isaac_ext_fragment_t> dFrag;
    DeviceBuffer<uint32_t> dCig;
    DeviceBuffer<IndelTask> dTasks;
    DeviceBuffer<IndelResult> dIndel;
    DeviceBuffer<ShadowTask> dShadowTasks;
    DeviceBuffer<int> dShadowScratch;
    DeviceBuffer<uint32_t> dTaskBegin, dTaskCount, dLargeTasks;
    DeviceBuffer<unsigned long long> dPoolSize;
    // sequencing adapters: first candidate of every clipper slot, slot of every candidate (rescue)
    PinnedBuffer<isaac_ext_candidate_t> hAdapterFirst;
    DeviceBuffer<isaac_ext_candidate_t> dAdapterFirst;
    PinnedBuffer<uint32_t> hSlot;
    DeviceBuffer<uint32_t> dSlot;
    // flattened results handed back to the caller
    HostBuffer<WorkFragment> work;          // grow-only, never value-initialised (every live element is written before it is read)
    std::vector<uint32_t> indelCigars;
    HostBuffer<isaac_ext_fragment_t> outFragments;
    HostBuffer<uint64_t> outBegin;
    HostBuffer<uint32_t> outCigars;
    uint64_t outFragmentCount = 0, outCigarWords = 0;
    std::vector<uint8_t> outFlags;
    // isaac_ext_rescue_shadows: list bookkeeping on the device (kernels_rescue.cuh) and its flat result in pinned memory
    DeviceBuffer<uint32_t> dKept, dAdoptedBy, dCounts, dBegins, dSlot3, dSources, dCig3, dOutCigars;
    DeviceBuffer<ShadowListState> dListState;
    DeviceBuffer<uint8_t> dRescued, dScanTemp;
    DeviceBuffer<isaac_ext_candidate_t> dCand3;
    DeviceBuffer<isaac_ext_fragment_t> dFrag3, dOutFragments;
    DeviceBuffer<uint64_t> dOutBegin;
    // two result sets: isaac_ext_build_templates reads one while the next slice's rescue pass fills the other
    PinnedBuffer<uint32_t> hTotals, hOutCigars[2];
    PinnedBuffer<isaac_ext_fragment_t> hOutFragments[2];
    PinnedBuffer<uint64_t> hOutBegin[2];
    PinnedBuffer<uint8_t> hRescued[2];

    void release()
    {
        hCand1.release(); hCand3.release(); hFrag1.release(); hFrag3.release(); hCig1.release(); hCig3.release();
        hTasks.release(); hIndel.release(); hShadowTasks.release();
        hAdapterFirst.release(); dAdapterFirst.release(); hSlot.release(); dSlot.release();
        dKept.release(); dAdoptedBy.release(); dCounts.release(); dBegins.release(); dSlot3.release(); dSources.release(); dCig3.release();
        dOutCigars.release(); dListState.release(); dRescued.release(); dScanTemp.release(); dCand3.release(); dFrag3.release();
        dOutFragments.release(); dOutBegin.release(); hTotals.release();
        for (int k = 0; k < 2; ++k) { hOutCigars[k].release(); hOutFragments[k].release(); hOutBegin[k].release(); hRescued[k].release(); }
        dCand.release(); dFrag.release(); dCig.release(); dTasks.release(); dIndel.release(); dShadowTasks.release();
        dShadowScratch.release(); dTaskBegin.release(); dTaskCount.release(); dPoolSize.release(); dLargeTasks.release();
    }
};

} // namespace isaac_b200
