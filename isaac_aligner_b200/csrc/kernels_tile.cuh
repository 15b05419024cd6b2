// The per-cluster bookkeeping of a tile ON THE DEVICE: what FragmentBuilder::build does around its aligners
// (FragmentBuilder.cpp:82-324) and what TemplateBuilder::buildTemplate does around ShadowAligner::rescueShadow
// (TemplateBuilder.cpp:97-1089), as kernels between the scoring kernels, so that a tile is one upload (seed matches; the BCL bytes
// are resident already) and one download (templates).  Nothing here visits the host.
//
//   build   B1 buildCandidatesKernel    matches -> candidate lists (addMatch, repeat-seed filter, consolidate)        | 1 cluster / thread
//           K1 ungappedKernel            over the match slots (a slot that holds no candidate is skipped)
//           B2 pairIndelKernel           adopt K1, consolidate, orderByUnclippedPosition, pair neighbours               | 1 read list / thread
//              simpleIndelKernel         over the match slots that hold a pair
//           B3 applyIndelKernel          patch, consolidate, count the lists' Smith-Waterman candidates                 | 1 read list / thread
//              cub::DeviceScan, gapCandidatesKernel: the dense batch of the gapped pass
//           K2 swForwardKernel + swTraceScoreKernel
//           B4 acceptGappedKernel        5-clause acceptance, final consolidate, final records                          | 1 read list / thread
//   plan    planCountKernel / planWriteKernel: the rescueShadow calls of every cluster (plan_device.cuh) and their scan windows
//           (shadow_window_device.cuh), cub::DeviceScan between them                                                    | 1 cluster / thread
//   rescue  K5, K1, shadow list kernels, K2 (kernels_shadow.cuh, kernels_rescue.cuh)
//   finish  finishTemplatesKernel (finish_device.cuh), gatherTemplateCigarsKernel                                       | 1 cluster / thread
//
// Memory: the candidate list of a read lives in the slots of its cluster's matches (a cluster never has more candidates than
// matches), so no pass needs a prefix sum to find its place: slot = match index.  Records are WorkFragment (72 B); the CIGAR of a
// record stays in the pool of the kernel pass that produced it (ungapped 3 words per slot, simple indel 5 per slot, gapped 32 per
// dense index) and is named by pool << 30 | word index once a list is final.
//
// Order-dependent steps keep the reference's order: std::sort is libstdc++'s own sequence of comparisons and moves
// (sort_replay.cuh), so the entry that survives consolidateDuplicateFragments is the reference's.
#pragma once
#include "consolidate_device.cuh"
#include "finish_device.cuh"
#include "host_pipeline.cuh"
#include "plan_device.cuh"
#include "shadow_window_device.cuh"

namespace isaac_b200
{

constexpr uint32_t TILE_NO_CANDIDATE = 0xFFFFFFFFu;       // readId of a match slot that holds no candidate (== ADAPTER_NO_CANDIDATE)
constexpr unsigned TILE_MAX_SEEDS = 64;                   // seeds of a cluster (both reads); the reference's default is 4 per read
constexpr uint32_t TILE_ERROR_MATCHES = 8u;               // errorFlag bit: malformed match batch
constexpr uint32_t TILE_GAPPED_STRIDE = 32;
#ifndef FINISH_MIN_BLOCKS
#define FINISH_MIN_BLOCKS 8
#endif

struct TileView
{
    const isaac_ext_match_t *matches;
    const uint64_t *clusterMatchBegin;
    const isaac_ext_seed_t *seeds;
    uint32_t seedCount, clusters, readCount, repeatThreshold, gapLimit, withGaps, gappedMismatchesMax;
    uint32_t readLength[2];
    const uint64_t *contigLength;
    uint32_t contigCount;
    // the lists
    WorkFragment *work;
    uint32_t *listBegin, *listCount;      // per (cluster, readIndex)
    uint8_t *built;                       // per cluster: return value of build()
    // pools of the three scoring passes
    const isaac_ext_fragment_t *frag1; const uint32_t *cig1;
    const isaac_ext_fragment_t *frag3; const uint32_t *cig3;
    const uint32_t *cigIndel;
};

__device__ __forceinline__ const uint32_t *tileCigar(const TileView &v, const WorkFragment &w)
{
    return (w.pool == 0 ? v.cig1 : w.pool == 1 ? v.cigIndel : v.cig3) + w.f.cigarOffset;
}
__device__ __forceinline__ long tileBeginClipped(const TileView &v, const WorkFragment &w)       // FragmentMetadata::getBeginClippedLength (:148-159)
{
    if (!w.f.cigarLength) return 0;
    const uint32_t word = tileCigar(v, w)[0];
    return (word & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP ? long(word >> 4) : 0;
}
__device__ __forceinline__ long tileEndClipped(const TileView &v, const WorkFragment &w)         // getEndClippedLength (:161-172)
{
    if (!w.f.cigarLength) return 0;
    const uint32_t word = tileCigar(v, w)[w.f.cigarLength - 1];
    return (word & 0xFu) == ISAAC_EXT_CIGAR_SOFT_CLIP ? long(word >> 4) : 0;
}
__device__ __forceinline__ long tileUnclippedPosition(const TileView &v, const WorkFragment &w) { return long(w.f.position) - tileBeginClipped(v, w); }   // :185-188
__device__ __forceinline__ isaac_ext_candidate_t tileCandidateOf(const isaac_ext_fragment_t &f, const long position)
{
    isaac_ext_candidate_t c;
    c.position = position; c.readId = f.readId; c.contigStrand = (f.contigId << 1) | (f.reverse ? 1u : 0u);
    return c;
}

/// B1: FragmentBuilder::build up to alignFragments (FragmentBuilder.cpp:92-134) + the first consolidate (:159).
/// A seed's matches all become candidates unless the seed is a repeat: it has a TooManyMatch record or repeatThreshold matches
/// (:101-121: the running count stops the additions at the threshold, removeRepeatSeedAlignments :128-134 then drops the earlier
/// ones as well), so two walks over the cluster's matches give the lists in the reference's order.
__global__ void buildCandidatesKernel(const TileView v, isaac_ext_candidate_t *__restrict__ candidates, isaac_ext_candidate_t *__restrict__ adapterFirst,
                                      uint32_t *__restrict__ errorFlag)
{
    const uint64_t M = v.clusterMatchBegin[v.clusters];
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < v.clusters; c += gridDim.x * blockDim.x)
    {
        const uint64_t mb = v.clusterMatchBegin[c], me = v.clusterMatchBegin[c + 1];
        const size_t l0 = size_t(c) * v.readCount;
        for (unsigned r = 0; r < v.readCount; ++r) { v.listBegin[l0 + r] = uint32_t(mb); v.listCount[l0 + r] = 0; }
        if (adapterFirst) for (unsigned k = 0; k < 2 * v.readCount; ++k) adapterFirst[l0 * 2 + k].readId = TILE_NO_CANDIDATE;
        v.built[c] = 0;
        if (me < mb || me > M) { atomicOr(errorFlag, TILE_ERROR_MATCHES); continue; }
        if (mb == me) continue;
        uint32_t seedMatches[TILE_MAX_SEEDS];
        uint64_t tooMany = 0;
        for (unsigned s = 0; s < v.seedCount; ++s) seedMatches[s] = 0;
        uint64_t end = mb;
        bool bad = false;
        for (; end < me && !matchIsNoMatch(v.matches[end]); ++end)
        {
            const isaac_ext_match_t match = v.matches[end];
            const unsigned seedIndex = matchSeed(match);
            if (seedIndex >= v.seedCount) { bad = true; break; }                       // seedMatchCounts_.at() would throw
            if (matchIsTooMany(match)) tooMany |= uint64_t(1) << seedIndex;
            else ++seedMatches[seedIndex];
        }
        if (bad) { atomicOr(errorFlag, TILE_ERROR_MATCHES); continue; }
        unsigned repeatSeedsCount = 0, count[2] = {0u, 0u};
        uint64_t repeat = 0;
        for (unsigned s = 0; s < v.seedCount; ++s)
        {
            const bool hasTooMany = (tooMany >> s) & 1u;
            if (hasTooMany || seedMatches[s] >= v.repeatThreshold)
            {
                repeat |= uint64_t(1) << s;
                if (v.repeatThreshold && (hasTooMany || seedMatches[s])) ++repeatSeedsCount;
            }
            else count[v.seeds[s].readIndex] += seedMatches[s];
        }
        WorkFragment *list[2] = {v.work + mb, v.work + mb + count[0]};
        unsigned filled[2] = {0u, 0u};
        for (uint64_t m = mb; m < end; ++m)
        {
            const isaac_ext_match_t match = v.matches[m];
            const unsigned seedIndex = matchSeed(match);
            if (((repeat >> seedIndex) & 1u) || matchIsTooMany(match)) continue;
            // addMatch (:219-249) with getReadPosition (:326-343)
            const isaac_ext_seed_t seed = v.seeds[seedIndex];
            const bool reverse = matchReverse(match);
            const long seedPosition = matchPosition(match);
            const long readLength = v.readLength[seed.readIndex];
            WorkFragment w;
            isaac_ext_fragment_t &f = w.f;
            f.position = reverse ? seedPosition + long(seed.length) + long(seed.offset) - readLength : seedPosition - long(seed.offset);
            f.logProbability = 0.0; f.contigId = matchContig(match); f.readId = c * v.readCount + seed.readIndex; f.cigarOffset = 0;
            f.smithWatermanScore = 0; f.observedLength = 0; f.mismatchCount = 0; f.matchesInARow = 0; f.gapCount = 0; f.editDistance = 0;
            f.uniqueSeedCount = 0; f.repeatSeedsCount = uint16_t(repeatSeedsCount);                                     // :167
            f.nonUniqueSeedOffsetFirst = 0xFFFF; f.nonUniqueSeedOffsetSecond = 0; f.firstSeedIndex = int16_t(seedIndex);
            f.lowClipped = 0; f.highClipped = 0; f.cigarLength = 0; f.reverse = reverse; f.readIndex = uint8_t(seed.readIndex); f.matchCount = 0;
            if (seed.length != 64 && matchHasNeighbors(match))                          // STRONG_SEED_LENGTH (Alignment.hh:38)
            {
                f.nonUniqueSeedOffsetFirst = seed.offset; f.nonUniqueSeedOffsetSecond = seed.offset;
            }
            else f.uniqueSeedCount = 1;
            w.pool = 0; w.slot = 0;
            // a read position behind the end of the contig cannot come from a seed that lies on the contig; the reference does not
            // survive such a match either (it reads qualities and bases off the end of its vectors and trips the assertion of
            // Quality.hh's lookup, checked with the reference build), so the batch is refused rather than aligned
            if (f.contigId >= v.contigCount || f.position > long(v.contigLength[f.contigId])) { bad = true; continue; }
            list[seed.readIndex][filled[seed.readIndex]++] = w;
        }
        if (bad) { atomicOr(errorFlag, TILE_ERROR_MATCHES); continue; }
        v.built[c] = (filled[0] || filled[1]) ? 1 : 0;                                  // return value of build() (:136-144)
        for (unsigned r = 0; r < v.readCount; ++r)
        {
            const unsigned n = consolidateDuplicateFragmentsReplay(list[r], filled[r], false);    // :159
            const uint32_t begin = uint32_t(list[r] - v.work);
            v.listBegin[l0 + r] = begin; v.listCount[l0 + r] = n;
            for (unsigned k = 0; k < n; ++k)
            {
                WorkFragment &w = list[r][k];
                w.slot = begin + k;
                const isaac_ext_candidate_t cand = tileCandidateOf(w.f, w.f.position);
                candidates[begin + k] = cand;
                // one FragmentSequencingAdapterClipper per read list (:164): its two strands are initialised by the first fragment
                // of that strand in list order (:173)
                if (adapterFirst && adapterFirst[(l0 + r) * 2 + (w.f.reverse ? 1 : 0)].readId == TILE_NO_CANDIDATE)
                    adapterFirst[(l0 + r) * 2 + (w.f.reverse ? 1 : 0)] = cand;
            }
        }
    }
}

/// B2: adopt the ungapped alignments, consolidate (:179), SimpleIndelAligner::alignSimpleIndels pairing (SimpleIndelAligner.cpp:460-518):
/// the pair (h, h + 1) of a list becomes the task of slot listBegin + h
__global__ void pairIndelKernel(const TileView v, IndelTask *__restrict__ tasks, uint8_t *__restrict__ taskValid)
{
    const size_t lists = size_t(v.clusters) * v.readCount;
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l < lists; l += size_t(gridDim.x) * blockDim.x)
    {
        WorkFragment *list = v.work + v.listBegin[l];
        unsigned n = v.listCount[l];
        if (!n) continue;
        for (unsigned k = 0; k < n; ++k) adoptAlignment(list[k], v.frag1[list[k].slot], 0, list[k].slot);
        n = consolidateDuplicateFragmentsReplay(list, n, true);
        v.listCount[l] = n;
        if (!v.gapLimit || n < 2) continue;
        sort_replay::sort(list, n, [&v](const WorkFragment &x, const WorkFragment &y) {            // orderByUnclippedPosition (:443-449)
            return x.f.contigId < y.f.contigId || (x.f.contigId == y.f.contigId && tileUnclippedPosition(v, x) < tileUnclippedPosition(v, y));
        });
        for (unsigned h = 0; h + 1 < n; ++h)
        {
            const WorkFragment &head = list[h], &tail = list[h + 1];
            if (head.f.contigId != tail.f.contigId || head.f.reverse != tail.f.reverse) continue;
            const isaac_ext_seed_t headSeed = v.seeds[head.f.firstSeedIndex], tailSeed = v.seeds[tail.f.firstSeedIndex];
            const long distance = tileUnclippedPosition(v, tail) - tileUnclippedPosition(v, head);
            if (!((distance < 0 ? -distance : distance) < long(v.gapLimit))) continue;                // :490
            const long readLength = v.readLength[head.f.readIndex];
            const long headSeedOffset = head.f.reverse ? readLength - headSeed.offset - headSeed.length : headSeed.offset;   // :493-494
            const long tailSeedOffset = head.f.reverse ? readLength - tailSeed.offset - tailSeed.length : tailSeed.offset;
            auto side = [&v](const WorkFragment &w, const long seedOffset, const unsigned seedLength) {
                IndelSide s;
                s.position = w.f.position; s.beginClipped = uint32_t(tileBeginClipped(v, w)); s.endClipped = uint32_t(tileEndClipped(v, w));
                s.observedLength = w.f.cigarLength ? w.f.observedLength : 0;
                s.smithWatermanScore = w.f.smithWatermanScore; s.mismatchCount = w.f.mismatchCount;
                s.lowClipped = w.f.lowClipped; s.highClipped = w.f.highClipped;
                s.seedOffset = uint32_t(seedOffset); s.seedLength = seedLength;
                return s;
            };
            IndelTask task;
            task.readId = head.f.readId; task.contigId = head.f.contigId; task.reverse = head.f.reverse;
            for (unsigned k = 0; k < 6; ++k) task.pad[k] = 0;
            if (0 < tailSeedOffset - headSeedOffset)
            {
                // seeds ordered like the alignments: a deletion, patch the head (:497-503)
                task.insertion = 0;
                task.head = side(head, headSeedOffset, headSeed.length);
                task.tail = side(tail, tailSeedOffset, tailSeed.length);
            }
            else
            {
                // alignSimpleInsertion(*tail as head, ..., *head as tail) (:504-509); still patches list[h]
                task.insertion = 1;
                task.head = side(tail, tailSeedOffset, tailSeed.length);
                task.tail = side(head, headSeedOffset, headSeed.length);
            }
            const uint32_t slot = v.listBegin[l] + h;
            tasks[slot] = task; taskValid[slot] = 1;
        }
    }
}

__device__ __forceinline__ bool tileGoesToGappedAligner(const WorkFragment &w) { return ISAAC_EXT_SW_MISMATCH_CUTOFF < w.f.mismatchCount; }   // FragmentBuilder.cpp:190-200

/// B3: apply the simple-indel patches, consolidate (:184), count the fragments for the gapped aligner
__global__ void applyIndelKernel(const TileView v, const IndelResult *__restrict__ indel, const uint8_t *__restrict__ taskValid, uint32_t *__restrict__ gapCounts)
{
    const size_t lists = size_t(v.clusters) * v.readCount;
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l < lists; l += size_t(gridDim.x) * blockDim.x)
    {
        WorkFragment *list = v.work + v.listBegin[l];
        unsigned n = v.listCount[l];
        uint32_t targets = 0;
        if (n)
        {
            if (v.gapLimit)
            {
                for (unsigned h = 0; h + 1 < n; ++h)
                {
                    const uint32_t slot = v.listBegin[l] + h;
                    if (taskValid[slot] && indel[slot].accepted) adoptAlignment(list[h], indel[slot].fragment, 1, slot);
                }
                n = consolidateDuplicateFragmentsReplay(list, n, true);
                v.listCount[l] = n;
            }
            if (v.withGaps) for (unsigned k = 0; k < n; ++k) targets += tileGoesToGappedAligner(list[k]) ? 1u : 0u;
        }
        gapCounts[l] = targets;
    }
}

/// the dense batch of the gapped pass: alignGapped starts from resetAlignment(), the unclipped position of the current alignment
/// (GappedAligner.cpp:175)
__global__ void gapCandidatesKernel(const TileView v, const uint32_t *__restrict__ gapBegin, isaac_ext_candidate_t *__restrict__ candidates)
{
    const size_t lists = size_t(v.clusters) * v.readCount;
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l < lists; l += size_t(gridDim.x) * blockDim.x)
    {
        const WorkFragment *list = v.work + v.listBegin[l];
        uint32_t at = gapBegin[l];
        if (at == gapBegin[l + 1]) continue;
        for (unsigned k = 0; k < v.listCount[l]; ++k)
            if (tileGoesToGappedAligner(list[k])) candidates[at++] = tileCandidateOf(list[k].f, tileUnclippedPosition(v, list[k]));
    }
}

/// B4: the acceptance rule (:202-209), the final consolidate (:213), the final records: finalFragments[slot], cigarOffset = pool << 30 | word
__global__ void acceptGappedKernel(const TileView v, const uint32_t *__restrict__ gapBegin, isaac_ext_fragment_t *__restrict__ finalFragments)
{
    const size_t lists = size_t(v.clusters) * v.readCount;
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l < lists; l += size_t(gridDim.x) * blockDim.x)
    {
        WorkFragment *list = v.work + v.listBegin[l];
        unsigned n = v.listCount[l];
        if (!n) continue;
        if (gapBegin && gapBegin[l] != gapBegin[l + 1])
        {
            uint32_t at = gapBegin[l];
            for (unsigned k = 0; k < n; ++k)
                if (tileGoesToGappedAligner(list[k]))
                {
                    const isaac_ext_fragment_t g = v.frag3[at];
                    if (acceptGapped(list[k].f, g, v.gappedMismatchesMax)) adoptAlignment(list[k], g, 2, at);
                    ++at;
                }
        }
        n = consolidateDuplicateFragmentsReplay(list, n, true);
        v.listCount[l] = n;
        for (unsigned k = 0; k < n; ++k)
        {
            isaac_ext_fragment_t f = list[k].f;
            f.cigarOffset = (list[k].pool << FINISH_POOL_SHIFT) | f.cigarOffset;
            finalFragments[v.listBegin[l] + k] = f;
        }
    }
}

/// the flat result of isaac_ext_build_fragments: CIGAR words per list (for the prefix sums) ...
__global__ void countListWordsKernel(const size_t lists, const uint32_t *__restrict__ listBegin, const uint32_t *__restrict__ listCount,
                                     const isaac_ext_fragment_t *__restrict__ finalFragments, uint32_t *__restrict__ words)
{
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l < lists; l += size_t(gridDim.x) * blockDim.x)
    {
        uint32_t w = 0;
        for (unsigned k = 0; k < listCount[l]; ++k) w += finalFragments[listBegin[l] + k].cigarLength;
        words[l] = w;
    }
}
/// ... and the dense copy: fragments of list l at fragmentBegin[l], their words from wordBegin[l] on
__global__ void flattenListsKernel(const size_t lists, const uint32_t *__restrict__ listBegin, const uint32_t *__restrict__ listCount,
                                   const isaac_ext_fragment_t *__restrict__ finalFragments, const uint32_t *__restrict__ fragmentBegin,
                                   const uint32_t *__restrict__ wordBegin, const uint32_t *pool0, const uint32_t *pool1, const uint32_t *pool2,
                                   isaac_ext_fragment_t *__restrict__ fragmentsOut, uint32_t *__restrict__ cigarsOut, uint64_t *__restrict__ beginOut)
{
    for (size_t l = blockIdx.x * size_t(blockDim.x) + threadIdx.x; l <= lists; l += size_t(gridDim.x) * blockDim.x)
    {
        beginOut[l] = fragmentBegin[l];
        if (l == lists) continue;
        uint32_t at = wordBegin[l];
        for (unsigned k = 0; k < listCount[l]; ++k)
        {
            isaac_ext_fragment_t f = finalFragments[listBegin[l] + k];
            const uint32_t pool = f.cigarOffset >> FINISH_POOL_SHIFT;
            const uint32_t *src = (pool == 0 ? pool0 : pool == 1 ? pool1 : pool2) + (f.cigarOffset & FINISH_POOL_MASK);
            for (unsigned i = 0; i < f.cigarLength; ++i) cigarsOut[at + i] = src[i];
            f.cigarOffset = at;
            at += f.cigarLength;
            fragmentsOut[fragmentBegin[l] + k] = f;
        }
    }
}

/// plan: the number of rescueShadow calls of every cluster ...
__global__ void planCountKernel(const PlanView v, const uint32_t clusters, uint32_t *__restrict__ counts)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < clusters; c += gridDim.x * blockDim.x)
        counts[c] = planClusterRequests(v, c, nullptr, 0u);
}
/// ... and the calls themselves at requestBegin[c], each with its scan window (R1 of the rescue pass)
__global__ void planWriteKernel(const PlanView v, const uint32_t clusters, const uint32_t *__restrict__ requestBegin, const ShadowWindowModel model,
                                const uint32_t readLength0, const uint32_t readLength1, const uint64_t *__restrict__ contigLength,
                                isaac_ext_rescue_request_t *__restrict__ requests, ShadowTask *__restrict__ tasks)
{
    const uint32_t len[2] = {readLength0, readLength1};
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < clusters; c += gridDim.x * blockDim.x)
    {
        const uint32_t begin = requestBegin[c], count = requestBegin[c + 1] - begin;
        if (!count) continue;
        planClusterRequests(v, c, requests + begin, count);
        for (uint32_t i = begin; i < begin + count; ++i)
            shadowWindowOf(model, requests[i], len, long(contigLength[requests[i].orphanContigStrand >> 1]), tasks[i]);
    }
}
/// R1 for requests that come from the caller (isaac_ext_rescue_shadows); bad requests get an empty window and raise the flag
__global__ void shadowWindowsKernel(const uint32_t n, const isaac_ext_rescue_request_t *__restrict__ requests, const ShadowWindowModel model,
                                    const uint32_t readLength0, const uint32_t readLength1, const uint32_t readTotal, const uint32_t contigCount,
                                    const uint64_t *__restrict__ contigLength, ShadowTask *__restrict__ tasks, uint32_t *__restrict__ errorFlag)
{
    const uint32_t len[2] = {readLength0, readLength1};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_rescue_request_t q = requests[i];
        if (q.orphanReadId >= readTotal || (q.orphanContigStrand >> 1) >= contigCount)
        {
            atomicOr(errorFlag, TILE_ERROR_MATCHES);
            tasks[i] = ShadowTask{0, 0, 0u, 0u};
            continue;
        }
        shadowWindowOf(model, q, len, long(contigLength[q.orphanContigStrand >> 1]), tasks[i]);
    }
}

/// finish: the BamTemplate of every cluster (finish_device.cuh).  The scratch slice of cluster c starts where the slices of the
/// clusters before it end; finishScratchBytes is linear in (shadows, candidates), so that place follows from the rescue pass's and
/// the match batch's own offsets without another prefix sum.
__global__ void __launch_bounds__(128, FINISH_MIN_BLOCKS) finishTemplatesKernel(const FinishView v, const uint32_t clusters, const uint64_t *__restrict__ clusterMatchBegin,
                                      unsigned char *__restrict__ scratch, isaac_ext_template_t *__restrict__ templates,
                                      isaac_ext_fragment_t *__restrict__ fragments, uint32_t *__restrict__ cigarLengths,
                                      FinishSource *__restrict__ sources)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < clusters; c += gridDim.x * blockDim.x)
    {
        const uint64_t shadowsBefore = v.clusterRequestBegin ? v.requestFragmentBegin[v.clusterRequestBegin[c]] : 0;
        const uint64_t shadows = v.clusterRequestBegin ? v.requestFragmentBegin[v.clusterRequestBegin[c + 1]] - shadowsBefore : 0;
        const uint64_t candidatesBefore = clusterMatchBegin[c], candidates = clusterMatchBegin[c + 1] - candidatesBefore;
        unsigned char *slice = scratch + (finishScratchBytes(shadowsBefore, candidatesBefore) + uint64_t(c) * finishScratchBytes(0, 0) - finishScratchBytes(0, 0));
        isaac_ext_template_t o;
        isaac_ext_fragment_t f[2];
        FinishSource src[2];
        finishCluster(v, c, slice, shadows, candidates, o, f, src);
        templates[c] = o;
        for (unsigned r = 0; r < v.readCount; ++r)
        {
            fragments[size_t(c) * v.readCount + r] = f[r];
            cigarLengths[size_t(c) * v.readCount + r] = f[r].cigarLength;
            sources[size_t(c) * v.readCount + r] = src[r];
        }
    }
}

/// the CIGAR words of the templates gathered into one pool in (cluster, read) order; cigarOffset = place in that pool (an unaligned
/// fragment points at the place the next words go)
__global__ void gatherTemplateCigarsKernel(const size_t count, isaac_ext_fragment_t *__restrict__ fragments, const uint32_t *__restrict__ wordBegin,
                                           const uint32_t *pool0, const uint32_t *pool1, const uint32_t *pool2, const uint32_t *pool3,
                                           uint32_t *__restrict__ cigarsOut)
{
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
    {
        isaac_ext_fragment_t &f = fragments[i];
        const uint32_t tagged = f.cigarOffset, at = wordBegin[i];
        if (f.cigarLength)
        {
            const uint32_t pool = tagged >> FINISH_POOL_SHIFT;
            const uint32_t *src = (pool == 0 ? pool0 : pool == 1 ? pool1 : pool == 2 ? pool2 : pool3) + (tagged & FINISH_POOL_MASK);
            for (unsigned k = 0; k < f.cigarLength; ++k) cigarsOut[at + k] = src[k];
        }
        f.cigarOffset = at;
    }
}

} // namespace isaac_b200
