// The list bookkeeping of ShadowAligner::rescueShadow (ShadowAligner.cpp:205-290) on the device, between and after the scoring
// kernels of isaac_ext_rescue_shadows.  One thread walks one request's candidates in position order, exactly like the reference's
// loops do (the 1e-7-tolerant "better than the best so far" rule is order dependent):
//
//   shadowSelectKernel   after the ungapped pass: which candidates stay in shadowList (aligned, capacity 1000), the best one,
//                        and how many of them go to the gapped aligner (:238-256)
//   shadowGapKernel      writes those gapped candidates, in list order, at the request's place of the gapped batch
//   shadowAcceptKernel   after the gapped pass: the acceptance rule, the new best, best first (:255-288); final size of the list
//                        and of its CIGARs
//   shadowFlattenKernel  one thread per request: the final records and their CIGAR words to the flat result arrays
//
// Prefix sums between them (cub::DeviceScan) give every request its place in the gapped batch and in the result.
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

constexpr unsigned SHADOW_LIST_CAPACITY_D = 1000;     // TemplateBuilder::TRACKED_REPEATS_MAX_ONE_READ (TemplateBuilder.hh:145, .cpp:82)
constexpr uint32_t SHADOW_ADOPTED = 0x80000000u;

__device__ __forceinline__ bool lpEqualsD(double a, double b) { return 0.0000001 >= fabs(a - b); }     // ISAAC_LP_EQUALS (Quality.hh:104-107)
__device__ __forceinline__ bool lpLessD(double a, double b) { return !lpEqualsD(a, b) && a < b; }      // ISAAC_LP_LESS (:109-112)

/// per-request state carried between the kernels
struct ShadowListState
{
    uint32_t size;        // entries of shadowList
    int32_t best;         // index of the best entry, -1 = none
    uint32_t full;        // the list ran into its capacity: rescueShadow returns false (:212-215)
    uint32_t gapTargets;  // entries that go to the gapped aligner
};

/// kept[begin + j] = pool index of the j-th list entry
__global__ void shadowSelectKernel(uint32_t requests, const uint32_t *__restrict__ taskBegin, const uint32_t *__restrict__ taskCount,
                                   const isaac_ext_fragment_t *__restrict__ ungapped, uint32_t *__restrict__ kept,
                                   ShadowListState *__restrict__ state, uint32_t *__restrict__ gapCounts)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < requests; i += gridDim.x * blockDim.x)
    {
        const uint32_t begin = taskBegin[i], count = taskCount[i];
        ShadowListState s = {0u, -1, 0u, 0u};
        double bestLp = 0.0;
        for (uint32_t k = 0; k < count; ++k)
        {
            if (s.size == SHADOW_LIST_CAPACITY_D) { s.full = 1; break; }                          // :212-215
            const isaac_ext_fragment_t *f = ungapped + begin + k;
            if (!f->cigarLength) continue;                                                       // alignUngapped returned 0 (:223)
            const double lp = f->logProbability;
            if (s.best < 0 || lpLessD(bestLp, lp)) { s.best = int32_t(s.size); bestLp = lp; }     // :227-230
            kept[begin + s.size++] = begin + k;
        }
        if (!s.full && s.best >= 0 && ISAAC_EXT_SW_MISMATCH_CUTOFF < ungapped[kept[begin + s.best]].mismatchCount)   // :243
            for (uint32_t j = 0; j + 1 < s.size; ++j)
            {
                const isaac_ext_fragment_t *a = ungapped + kept[begin + j], *b = ungapped + kept[begin + j + 1];
                if (b->position - a->position < long(ISAAC_EXT_SW_DISTANCE_CUTOFF) && ISAAC_EXT_SW_MISMATCH_CUTOFF < a->mismatchCount)   // :249-253
                    ++s.gapTargets;
            }
        state[i] = s;
        gapCounts[i] = s.gapTargets;
    }
}

/// the gapped candidates of request i at [gapBegin[i], +gapTargets): unclipped position (FragmentMetadata::getUnclippedPosition,
/// FragmentMetadata.hh:185-188), the request as adapter-clipper slot, the list index they came from
__global__ void shadowGapKernel(uint32_t requests, const uint32_t *__restrict__ taskBegin, const ShadowListState *__restrict__ state,
                                const uint32_t *__restrict__ gapBegin, const isaac_ext_fragment_t *__restrict__ ungapped,
                                const uint32_t *__restrict__ ungappedCigars, const uint32_t *__restrict__ kept,
                                isaac_ext_candidate_t *__restrict__ candidates, uint32_t *__restrict__ slots, uint32_t *__restrict__ sources)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < requests; i += gridDim.x * blockDim.x)
    {
        const ShadowListState s = state[i];
        if (!s.gapTargets) continue;
        const uint32_t begin = taskBegin[i];
        uint32_t at = gapBegin[i];
        for (uint32_t j = 0; j + 1 < s.size; ++j)
        {
            const uint32_t ia = kept[begin + j];
            const isaac_ext_fragment_t *a = ungapped + ia, *b = ungapped + kept[begin + j + 1];
            if (b->position - a->position < long(ISAAC_EXT_SW_DISTANCE_CUTOFF) && ISAAC_EXT_SW_MISMATCH_CUTOFF < a->mismatchCount)
            {
                const uint32_t word = ungappedCigars[size_t(ia) * 3];
                const long clipped = (word & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP ? long(word >> 4) : 0;
                isaac_ext_candidate_t c;
                c.position = a->position - clipped; c.readId = a->readId; c.contigStrand = (a->contigId << 1) | (a->reverse ? 1u : 0u);
                candidates[at] = c; slots[at] = i; sources[at] = j;
                ++at;
            }
        }
    }
}

/// After the gapped pass.  finalOrder[begin + j] = pool index of the entry at final position j, | SHADOW_ADOPTED + the index of
/// its gapped record in adoptedBy[begin + j]; counts[i] = final size, words[i] = CIGAR words of the list, rescued[i].
__global__ void shadowAcceptKernel(uint32_t requests, const uint32_t *__restrict__ taskBegin, const ShadowListState *__restrict__ state,
                                   const uint32_t *__restrict__ gapBegin, const uint32_t *__restrict__ sources,
                                   const isaac_ext_fragment_t *__restrict__ ungapped, const isaac_ext_fragment_t *__restrict__ gapped,
                                   uint32_t gappedMismatchesMax, uint32_t *__restrict__ kept, uint32_t *__restrict__ adoptedBy,
                                   uint32_t *__restrict__ counts, uint32_t *__restrict__ words, uint8_t *__restrict__ rescued)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < requests; i += gridDim.x * blockDim.x)
    {
        const ShadowListState s = state[i];
        const uint32_t begin = taskBegin[i];
        int32_t best = s.best;
        double bestLp = best >= 0 ? ungapped[kept[begin + best]].logProbability : 0.0;
        for (uint32_t j = 0; j < s.size; ++j) adoptedBy[begin + j] = 0;
        for (uint32_t t = 0; t < s.gapTargets; ++t)
        {
            const uint32_t g = gapBegin[i] + t, j = sources[g];
            const isaac_ext_fragment_t *u = ungapped + kept[begin + j], *a = gapped + g;
            // the 5-clause acceptance rule (ShadowAligner.cpp:259-262); the ungapped entry is aligned, so getObservedLength() is its field
            if (a->matchCount && a->matchCount + ISAAC_EXT_BAND_WIDTH > u->observedLength && a->mismatchCount <= gappedMismatchesMax &&
                u->mismatchCount > a->mismatchCount && lpLessD(u->logProbability, a->logProbability))
            {
                adoptedBy[begin + j] = g + 1;
                if (lpLessD(bestLp, a->logProbability)) { best = int32_t(j); bestLp = a->logProbability; }    // :265-268
            }
        }
        const bool ok = !s.full && best >= 0;
        if (ok && best != 0)                                                                      // :285-288
        {
            const uint32_t k0 = kept[begin], a0 = adoptedBy[begin];
            kept[begin] = kept[begin + best]; adoptedBy[begin] = adoptedBy[begin + best];
            kept[begin + best] = k0; adoptedBy[begin + best] = a0;
        }
        uint32_t w = 0;
        for (uint32_t j = 0; j < s.size; ++j)
            w += adoptedBy[begin + j] ? gapped[adoptedBy[begin + j] - 1].cigarLength : ungapped[kept[begin + j]].cigarLength;
        counts[i] = s.size; words[i] = w; rescued[i] = ok ? 1 : 0;
    }
}

/// One thread per request (a list has one to three entries, rarely more): its final list to fragmentsOut[fragmentBegin[i] ..] with
/// the CIGAR words behind each other from wordBegin[i] on (fragment.cigarOffset indexes cigarsOut).  An adopted entry takes the
/// alignment of its gapped record; the seed bookkeeping of a rescued shadow is the same in both records.
__global__ void shadowFlattenKernel(uint32_t requests, const uint32_t *__restrict__ taskBegin, const uint32_t *__restrict__ counts,
                                    const uint32_t *__restrict__ fragmentBegin, const uint32_t *__restrict__ wordBegin,
                                    const uint32_t *__restrict__ kept, const uint32_t *__restrict__ adoptedBy,
                                    const isaac_ext_fragment_t *__restrict__ ungapped, const uint32_t *__restrict__ ungappedCigars,
                                    const isaac_ext_fragment_t *__restrict__ gapped, const uint32_t *__restrict__ gappedCigars,
                                    uint32_t gappedStride, isaac_ext_fragment_t *__restrict__ fragmentsOut, uint32_t *__restrict__ cigarsOut,
                                    uint64_t *__restrict__ fragmentBeginOut, const uint32_t fragmentTotal)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) fragmentBeginOut[requests] = fragmentTotal;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < requests; i += gridDim.x * blockDim.x)
    {
        const uint32_t begin = taskBegin[i], size = counts[i], fb = fragmentBegin[i];
        fragmentBeginOut[i] = fb;
        uint32_t offset = wordBegin[i];
        for (uint32_t j = 0; j < size; ++j)
        {
            const uint32_t adopted = adoptedBy[begin + j], source = kept[begin + j];
            isaac_ext_fragment_t f = adopted ? gapped[adopted - 1] : ungapped[source];
            const uint32_t *src = adopted ? gappedCigars + size_t(adopted - 1) * gappedStride : ungappedCigars + size_t(source) * 3;
            for (uint32_t k = 0; k < f.cigarLength; ++k) cigarsOut[offset + k] = src[k];
            f.cigarOffset = offset;
            offset += f.cigarLength;
            fragmentsOut[fb + j] = f;
        }
    }
}

} // namespace isaac_b200
