// K6: per-tile statistics of a batch of fragment records, reduced on the device into a small vector of u64 counters
// that the ranks of a multi-GPU run sum with one NCCL all-reduce (the reference sums its per-thread
// MatchSelectorStats the same way at the end of a tile, MatchSelector.cpp:439-442; counters are plain u64 sums,
// include/alignment/matchSelector/TileStats.hh:68-93).
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

// layout of the counter vector (ISAAC_EXT_STATS_COUNTERS entries)
enum : unsigned
{
    STAT_FRAGMENTS = 0,        // records seen
    STAT_ALIGNED = 1,          // cigarLength != 0
    STAT_GAPPED = 2,           // gapCount != 0
    STAT_PERFECT = 3,          // aligned with editDistance == 0
    STAT_MISMATCHES = 4,       // sum of mismatchCount over aligned records
    STAT_EDIT_DISTANCE = 5,    // sum of editDistance
    STAT_GAPS = 6,             // sum of gapCount
    STAT_BASES = 7,            // sum of observedLength
    STAT_MISMATCH_HISTOGRAM = 8,   // 33 bins: mismatchCount clipped at 32
    STAT_COUNT = 64
};

__global__ void tileStatsKernel(uint32_t n, const isaac_ext_fragment_t *__restrict__ fragments, unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long block[STAT_COUNT];
    for (unsigned i = threadIdx.x; i < STAT_COUNT; i += blockDim.x) block[i] = 0;
    __syncthreads();
    unsigned long long local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_fragment_t f = fragments[i];
        ++local[STAT_FRAGMENTS];
        if (f.cigarLength)
        {
            ++local[STAT_ALIGNED];
            local[STAT_GAPPED] += f.gapCount != 0;
            local[STAT_PERFECT] += f.editDistance == 0;
            local[STAT_MISMATCHES] += f.mismatchCount;
            local[STAT_EDIT_DISTANCE] += f.editDistance;
            local[STAT_GAPS] += f.gapCount;
            local[STAT_BASES] += f.observedLength;
            atomicAdd(&block[STAT_MISMATCH_HISTOGRAM + min(unsigned(f.mismatchCount), 32u)], 1ull);
        }
    }
#pragma unroll
    for (unsigned k = 0; k < 8; ++k)
    {
        unsigned long long v = local[k];
        for (unsigned d = 16; d; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&block[k], v);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < STAT_COUNT; i += blockDim.x)
        if (block[i]) atomicAdd(&stats[i], block[i]);
}

} // namespace isaac_b200
