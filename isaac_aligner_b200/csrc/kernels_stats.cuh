// K6: per-tile statistics of a batch of fragment records, reduced on the device into a small vector of u64 counters
// that the ranks of a multi-GPU run sum with one NCCL all-reduce (the reference sums its per-thread
// MatchSelectorStats the same way at the end of a tile, MatchSelector.cpp:439-442; counters are plain u64 sums,
// include/alignment/matchSelector/TileStats.hh:68-93).
#pragma once
#include "device_types.cuh"
#include "finish_device.cuh"
#include "score.cuh"

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

// ---- matchSelector::TileStats of a tile's templates (TileStats.hh:68-142): the alignment score histograms and the per-cycle
// arrays, recorded the way MatchSelectorStats::recordTemplate does (MatchSelectorStats.hh:77-103).  One cluster per thread; the
// counters are global u64 words (one block of TILE_CYCLE_STATS_WORDS per read index and pass filter), the score histograms
// warp-aggregated (most fragments of a tile share a few scores).
// mismatchCycles of a FragmentMetadata are filled by the updateFragmentCigar call that scored its alignment and are NOT touched by
// what happens to the fragment afterwards (setNoMatch, filterLowQualityFragments, the end clippers only change mismatchCount): the
// reference reads the first mismatchCount entries of that list.  Here the list is re-derived by scoring the alignment the template
// took the fragment from (FinishSource) once more, and cut to the final mismatchCount.
enum : unsigned
{
    TCS_SCORE_FRAGMENTS = 0, TCS_SCORE_MISMATCHES = 8192, TCS_SCORE_TEMPLATES = 16384, TCS_SCORE_TEMPLATE_MISMATCHES = 24576,
    TCS_BLANKS = 32768, TCS_UNIQUE_BLANKS = 33792, TCS_MISMATCHES = 34816, TCS_UNIQUE_MISMATCHES = 35840,
    TCS_UNIQUE_X = 36864,           // five arrays: 1, 2, 3, 4, "more" (= exactly 5) mismatches so far
    TCS_X = 41984,                  // five arrays
    TCS_UNIQUE_FRAGMENTS = 47104, TCS_WORDS = ISAAC_EXT_TILE_CYCLE_STATS_WORDS
};
static_assert(TCS_WORDS == 47105, "TileStats layout");

__global__ void tileCycleStatsKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, const uint32_t clusters,
                                     const isaac_ext_template_t *__restrict__ templates, const isaac_ext_fragment_t *__restrict__ fragments,
                                     const FinishSource *__restrict__ sources, const uint32_t *pool0, const uint32_t *pool1, const uint32_t *pool2,
                                     const uint32_t *pool3, const uint8_t *__restrict__ pf, unsigned long long *__restrict__ stats,
                                     uint32_t *__restrict__ errorFlag)
{
    __shared__ double tables[201];
    const ScoreParams sp = stageScoreTables(spGlobal, tables);
    const unsigned rc = reads.readCount, lane = threadIdx.x & 31u;
    const uint32_t perGrid = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < clusters; base += perGrid)
    {
        const uint32_t c = base + lane;
        const bool live = c < clusters;
        const bool passes = live && (!pf || pf[c]);
        isaac_ext_template_t t;
        if (live) t = templates[c];
        unsigned templateMismatches = 0, firstReadIndex = 0;
        for (unsigned r = 0; r < rc; ++r)
        {
            isaac_ext_fragment_t f;
            uint32_t score = 0xFFFFFFFFu;
            if (live) { f = fragments[size_t(c) * rc + r]; score = t.fragmentAlignmentScore[r]; }
            const bool hasScore = live && score != 0xFFFFFFFFu;
            if (hasScore && score > 0x1FFFu) { atomicOr(errorFlag, 64u); score = 0x1FFFu; }      // the reference asserts (TileStats.hh:128)
            const bool unique = live && f.cigarLength != 0 && hasScore && score > 3u;            // FragmentMetadata.hh:268
            if (live) { templateMismatches += f.mismatchCount; if (!r) firstReadIndex = f.readIndex; }
            // ---- alignment score histogram of the fragments: one add per distinct (score, pass) of the warp
            {
                const unsigned key = hasScore ? score * 2u + (passes ? 1u : 0u) : 0xFFFFFFFFu - lane;
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
                if (hasScore && lane == __ffs(peers) - 1u)
                {
                    atomicAdd(stats + size_t(r * 2) * TCS_WORDS + TCS_SCORE_FRAGMENTS + score, (unsigned long long)__popc(peers));
                    if (passes) atomicAdd(stats + size_t(r * 2 + 1) * TCS_WORDS + TCS_SCORE_FRAGMENTS + score, (unsigned long long)__popc(peers));
                }
                if (hasScore && f.mismatchCount)
                    for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
                        atomicAdd(stats + size_t(r * 2 + p) * TCS_WORDS + TCS_SCORE_MISMATCHES + score, (unsigned long long)f.mismatchCount);
            }
            if (!live) continue;
            const unsigned L = reads.readLength[r], firstCycle = reads.firstCycle[r];
            const uint32_t readId = c * rc + r;
            // ---- blanks: the 'n' of the forward sequence per cycle
            for (unsigned w = 0; w * 32u < L; ++w)
            {
                uint32_t n = reads.nmask[size_t(readId) * reads.wordsN + w];
                while (n)
                {
                    const unsigned i = w * 32u + (__ffs(n) - 1u);
                    n &= n - 1u;
                    if (i >= L) break;
                    for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
                    {
                        unsigned long long *s = stats + size_t(r * 2 + p) * TCS_WORDS;
                        atomicAdd(s + TCS_BLANKS + firstCycle + i, 1ull);
                        if (unique) atomicAdd(s + TCS_UNIQUE_BLANKS + firstCycle + i, 1ull);
                    }
                }
            }
            if (unique) for (unsigned p = 0; p < (passes ? 2u : 1u); ++p) atomicAdd(stats + size_t(r * 2 + p) * TCS_WORDS + TCS_UNIQUE_FRAGMENTS, 1ull);
            // ---- mismatch cycles
            const FinishSource src = sources[size_t(c) * rc + r];
            unsigned wanted = f.mismatchCount;
            if (!wanted || !src.valid || !src.cigarLength) continue;
            uint64_t mask[ISAAC_EXT_MASK_WORDS];
            for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) mask[k] = 0;
            {
                const uint32_t pool = src.cigarOffset >> FINISH_POOL_SHIFT;
                const uint32_t *cigar = (pool == 0 ? pool0 : pool == 1 ? pool1 : pool == 2 ? pool2 : pool3) + (src.cigarOffset & FINISH_POOL_MASK);
                isaac_ext_fragment_t scratch;
                scoreCigar(ref, reads, sp, readId, L, src.reverse != 0, ref.contigOffset[src.contigId], long(src.position), cigar, src.cigarLength, scratch, mask);
            }
            const unsigned total = wanted;
            unsigned k = 0;
            for (unsigned w = 0; w < ISAAC_EXT_MASK_WORDS && k < total; ++w)
            {
                uint64_t m = mask[w];
                while (m && k < total)
                {
                    const unsigned i = w * 64u + (__ffsll((long long)m) - 1u);
                    m &= m - 1ull;
                    const unsigned cycle = src.reverse ? firstCycle + L - 1u - i : firstCycle + i;           // AlignerBase.cpp:171
                    const unsigned number = src.reverse ? total - k : k + 1u;                                // cycleMismatchNumber
                    ++k;
                    for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
                    {
                        unsigned long long *s = stats + size_t(r * 2 + p) * TCS_WORDS;
                        atomicAdd(s + TCS_MISMATCHES + cycle, 1ull);
                        if (number <= 5u) atomicAdd(s + TCS_X + (number - 1u) * 1024u + cycle, 1ull);
                        if (unique)
                        {
                            atomicAdd(s + TCS_UNIQUE_MISMATCHES + cycle, 1ull);
                            if (number <= 5u) atomicAdd(s + TCS_UNIQUE_X + (number - 1u) * 1024u + cycle, 1ull);
                        }
                    }
                }
            }
        }
        // ---- the template's own score (recordTemplate, TileStats.hh:113-123): under the read index of fragment 0
        {
            uint32_t score = live ? t.alignmentScore : 0xFFFFFFFFu;
            const bool hasScore = live && score != 0xFFFFFFFFu;
            if (hasScore && score > 0x1FFFu) { atomicOr(errorFlag, 64u); score = 0x1FFFu; }
            const unsigned key = hasScore ? (score * 2u + (passes ? 1u : 0u)) * 2u + firstReadIndex : 0xFFFFFFFFu - lane;
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
            if (hasScore && lane == __ffs(peers) - 1u)
                for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
                    atomicAdd(stats + size_t(firstReadIndex * 2 + p) * TCS_WORDS + TCS_SCORE_TEMPLATES + score, (unsigned long long)__popc(peers));
            if (hasScore && templateMismatches)
                for (unsigned p = 0; p < (passes ? 2u : 1u); ++p)
                    atomicAdd(stats + size_t(firstReadIndex * 2 + p) * TCS_WORDS + TCS_SCORE_TEMPLATE_MISMATCHES + score, (unsigned long long)templateMismatches);
        }
    }
}

} // namespace isaac_b200
