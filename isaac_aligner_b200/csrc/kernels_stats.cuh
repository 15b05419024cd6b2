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

// ---- matchSelector::TileBarcodeStats of a tile's templates (TileBarcodeStats.hh:40-160), recorded the way
// MatchSelectorStats::recordTemplate does (MatchSelectorStats.hh:77-103): one thread per cluster, the block's counters in shared
// memory, one global atomic per non-zero counter and block.

enum : unsigned
{
    TS_YIELD = 0, TS_YIELD_Q30, TS_QUALITY_SUM, TS_CLUSTERS, TS_UNANCHORED, TS_NMNM, TS_RM, TS_QC, TS_ALIGNED, TS_UNIQUE,
    TS_UNIQUE_PERFECT, TS_SCORE_SUM, TS_BASES, TS_UNIQUE_BASES, TS_MISMATCHES, TS_UNIQUE_MISMATCHES, TS_MODEL = 16,
    TS_NOMINAL = 25, TS_FRAGMENTS = 29, TS_COUNT = ISAAC_EXT_TEMPLATE_STATS_COUNTERS
};
enum : unsigned { TEMPLATE_NORMAL = 0, TEMPLATE_NMNM = 1, TEMPLATE_QC = 2, TEMPLATE_RM = 3 };       // TemplateAlignmentType (TileBarcodeStats.hh:30-37)

struct TlsDevice { uint32_t min, max, bestModel[2]; };

__global__ void templateStatsKernel(const ReadSetView reads, const TlsDevice tls, uint32_t clusters,
                                    const isaac_ext_template_t *__restrict__ templates, const isaac_ext_fragment_t *__restrict__ fragments,
                                    const uint32_t *__restrict__ cigars, const uint8_t *__restrict__ types, const uint8_t *__restrict__ pf,
                                    unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long block[4 * TS_COUNT];
    for (unsigned i = threadIdx.x; i < 4 * TS_COUNT; i += blockDim.x) block[i] = 0;
    __syncthreads();
    const unsigned rc = reads.readCount;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < clusters; c += gridDim.x * blockDim.x)
    {
        const isaac_ext_template_t t = templates[c];
        const bool passes = !pf || pf[c];
        bool unique[2] = {false, false}, anchored = false;
        isaac_ext_fragment_t f[2];
        for (unsigned r = 0; r < rc; ++r)
        {
            f[r] = fragments[size_t(c) * rc + r];
            const uint32_t score = t.fragmentAlignmentScore[r];
            const bool aligned = f[r].cigarLength != 0, hasScore = score != 0xFFFFFFFFu;
            unique[r] = aligned && hasScore && score > 3u;                                       // FragmentMetadata.hh:268
            anchored |= score != 0u;                                                             // BamTemplate::isUnanchored (BamTemplate.hh:90-95)
            // FragmentMetadataTileStatsAdapter (FragmentMetadataTileStatsAdapter.hh:43-110) -> recordFragment (TileBarcodeStats.hh:127-156)
            const unsigned L = reads.readLength[r];
            const uint8_t *q = reads.quality + size_t(c * rc + r) * reads.qualityStride;
            unsigned q30 = 0, sum = 0;
            for (unsigned i = 0; i < L; ++i) { const unsigned v = q[i]; q30 += v >= 30u; sum += v; }
            unsigned alignedBases = 0;
            for (unsigned k = 0; k < f[r].cigarLength; ++k)
            {
                const uint32_t op = cigars[f[r].cigarOffset + k];
                if ((op & 0xFu) == ISAAC_EXT_CIGAR_ALIGN) alignedBases += op >> 4;
            }
            for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
            {
                unsigned long long *s = block + (r * 2 + p) * TS_COUNT;
                atomicAdd(s + TS_YIELD, (unsigned long long)L);
                atomicAdd(s + TS_YIELD_Q30, (unsigned long long)q30);
                atomicAdd(s + TS_QUALITY_SUM, (unsigned long long)sum);
                atomicAdd(s + TS_FRAGMENTS, 1ull);
                if (aligned)
                {
                    if (hasScore && score) atomicAdd(s + TS_SCORE_SUM, (unsigned long long)score);
                    if (f[r].mismatchCount) atomicAdd(s + TS_MISMATCHES, (unsigned long long)f[r].mismatchCount);
                    atomicAdd(s + TS_BASES, (unsigned long long)alignedBases);
                    atomicAdd(s + TS_ALIGNED, 1ull);
                }
                if (unique[r])
                {
                    if (f[r].mismatchCount) atomicAdd(s + TS_UNIQUE_MISMATCHES, (unsigned long long)f[r].mismatchCount);
                    atomicAdd(s + TS_UNIQUE, 1ull);
                    atomicAdd(s + TS_UNIQUE_BASES, (unsigned long long)alignedBases);
                    if (!f[r].editDistance) atomicAdd(s + TS_UNIQUE_PERFECT, 1ull);
                }
            }
        }
        // BamTemplateTileStatsAdapter (BamTemplateTileStatsAdapter.hh:46-118) -> recordTemplate (TileBarcodeStats.hh:115-126);
        // pair-level counters live under the read index of fragment 0
        unsigned model = 8u, check = 3u;                                                         // InvalidAlignmentModel, NoMatch
        if (rc == 2 && unique[0] && unique[1] && f[0].contigId == f[1].contigId)
        {
            model = (f[0].position <= f[1].position ? 0u : 4u) | (f[0].reverse ? 2u : 0u) | (f[1].reverse ? 1u : 0u);   // TemplateLengthStatistics.hh:153-163
            if (model == tls.bestModel[0] || model == tls.bestModel[1])                           // checkModel (:104-118)
            {
                const long length = f[0].position < f[1].position
                    ? max(f[1].position + long(f[1].observedLength) - f[0].position, long(f[0].observedLength))
                    : max(f[0].position + long(f[0].observedLength) - f[1].position, long(f[1].observedLength));
                check = (unsigned long)length > tls.max ? 0u : (unsigned long)length < tls.min ? 1u : 2u;
            }
        }
        const unsigned type = types[c];
        for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
        {
            unsigned long long *s = block + (f[0].readIndex * 2 + p) * TS_COUNT;
            atomicAdd(s + TS_MODEL + model, 1ull);
            atomicAdd(s + TS_NOMINAL + check, 1ull);
            atomicAdd(s + TS_CLUSTERS, 1ull);
            if (!anchored) atomicAdd(s + TS_UNANCHORED, 1ull);
            if (type == TEMPLATE_NMNM) atomicAdd(s + TS_NMNM, 1ull);
            if (type == TEMPLATE_RM) atomicAdd(s + TS_RM, 1ull);
            if (type == TEMPLATE_QC) atomicAdd(s + TS_QC, 1ull);
        }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < 4 * TS_COUNT; i += blockDim.x)
        if (block[i]) atomicAdd(&stats[i], block[i]);
}

} // namespace isaac_b200
