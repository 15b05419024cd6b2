// --avoid-smith-waterman: GappedAligner::makesSenseToGapAlign (GappedAligner.cpp:88-165) as a pre-pass of the
// gapped kernels.  The reference asks, before every Smith-Waterman, whether the 7-mers the read shares with its database
// window vote for two DIFFERENT read offsets with at least 8 hits each; only then can a gap pay off.
//
// The reference keeps the read's 7-mer table in the GappedAligner, keyed by (tile, cluster, read) per strand, and does not
// rebuild it while the key stays (:95-121) -- even when the clipped range of the read differs between candidates.  The table
// a candidate is judged with therefore belongs to the first candidate of its key that reached this point ("owner"):
//
//   swHashOwnerKernel   replays that cache over the candidates in call order.  A key never matches across clusters, so the
//                       walk is cut where the cluster changes (the candidates of one cluster are contiguous in every caller,
//                       like in the reference, where one thread finishes a cluster before the next); one thread per run.
//   avoidSwKernel       one warp per candidate: the owner's unique 7-mers (28-bit code words, shared memory), the votes of
//                       the candidate's database 7-mers per offset (shared-memory counters), result = bit 31 of the
//                       candidate's prep word, which makes the gapped kernels return "no alignment" like the reference does.
#pragma once
#include "kernels2.cuh"

namespace isaac_b200
{

constexpr unsigned AVOID_KMER = 7;                    // GappedAligner::HASH_KMER_LENGTH (GappedAligner.hh:59)
constexpr unsigned AVOID_SUFFICIENT_HITS = 8;         // SUFFICIENT_NUMBER_OF_HITS (:75)
constexpr uint32_t AVOID_NO_OWNER = 0xFFFFFFFFu;
constexpr uint32_t PREP_SKIP_SW = 0x80000000u;
constexpr unsigned AVOID_WARPS = 4;

__global__ void swHashOwnerKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n,
                                  const isaac_ext_candidate_t *__restrict__ candidates, const uint32_t *__restrict__ adapterClip,
                                  uint32_t *__restrict__ owner)
{
    for (uint32_t first = blockIdx.x * blockDim.x + threadIdx.x; first < n; first += gridDim.x * blockDim.x)
    {
        const uint32_t cluster = candidates[first].readId / reads.readCount;
        if (first && candidates[first - 1].readId / reads.readCount == cluster) continue;      // not the start of a run
        uint32_t key[2] = {AVOID_NO_OWNER, AVOID_NO_OWNER}, slotOwner[2] = {AVOID_NO_OWNER, AVOID_NO_OWNER};   // hashedQuery*_ (:95-97)
        for (uint32_t j = first; j < n; ++j)
        {
            const isaac_ext_candidate_t c = candidates[j];
            if (c.readId / reads.readCount != cluster) break;
            const GappedPrep p = prepareGapped(ref, reads, c, adapterClip, j);
            if (!p.run) { owner[j] = AVOID_NO_OWNER; continue; }                                // returned before the heuristic (:204-208)
            const unsigned s = c.contigStrand & 1u;
            if (key[s] != c.readId) { key[s] = c.readId; slotOwner[s] = j; }
            owner[j] = slotOwner[s];
        }
    }
}

/// prepOut[j] = clip word of candidate j (adapterClip[j], or "nothing clipped") | PREP_SKIP_SW when gap-aligning it makes no sense.
/// Dynamic shared memory: AVOID_WARPS * (maxLength + 3 * maxLength + 32) words.
__global__ void __launch_bounds__(AVOID_WARPS * 32)
avoidSwKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n, const isaac_ext_candidate_t *__restrict__ candidates,
              const uint32_t *__restrict__ adapterClip, const uint32_t *__restrict__ owner, uint32_t maxLength,
              uint32_t *__restrict__ prepOut)
{
    extern __shared__ uint32_t avoidShared[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned votesSize = 3u * maxLength + 32u;
    uint32_t *queryKmers = avoidShared + size_t(warp) * (maxLength + votesSize);
    uint32_t *votes = queryKmers + maxLength;
    const uint32_t INVALID = 0xFFFFFFFFu, KMER_MASK = 0x0FFFFFFFu, NOT_ACGT = 0x0CCCCCCCu;
    for (uint32_t j = blockIdx.x * AVOID_WARPS + warp; j < n; j += gridDim.x * AVOID_WARPS)
    {
        const isaac_ext_candidate_t c = candidates[j];
        const unsigned L = reads.length(c.readId);
        const uint32_t word = adapterClip ? adapterClip[j] : (L << 16);
        const uint32_t o = owner[j];
        if (o == AVOID_NO_OWNER) { if (lane == 0) prepOut[j] = word; continue; }
        const GappedPrep p = prepareGapped(ref, reads, c, adapterClip, j);
        const GappedPrep po = prepareGapped(ref, reads, candidates[o], adapterClip, o);
        const uint64_t *strand = reads.strandCodes(c.readId, p.f.reverse);
        // queryKmerOffsets_ of the owner's query (:99-117): 7-mers without 'n'; a 7-mer seen twice is REPEAT_OFFSET_MAGIC
        const unsigned ownerKmers = po.sequenceLength >= AVOID_KMER ? po.sequenceLength - AVOID_KMER + 1 : 0u;
        for (unsigned t = lane; t < ownerKmers; t += 32u)
        {
            const uint32_t v = uint32_t(readCodes16(strand, unsigned(po.begin) + t)) & KMER_MASK;
            queryKmers[t] = (v & NOT_ACGT) ? INVALID : v;
        }
        for (unsigned k = lane; k < votesSize; k += 32u) votes[k] = 0;
        __syncwarp();
        uint32_t mine = 0;          // bit k: the 7-mer at offset lane + 32 k occurs again (reads have at most 1024 bases)
        for (unsigned t = lane, k = 0; t < ownerKmers; t += 32u, ++k)
        {
            const uint32_t v = queryKmers[t];
            bool rep = false;
            if (v != INVALID)
                for (unsigned u = 0; u < ownerKmers; ++u) rep |= (u != t && queryKmers[u] == v);
            if (rep) mine |= 1u << k;
        }
        __syncwarp();
        for (unsigned t = lane, k = 0; t < ownerKmers; t += 32u, ++k) if ((mine >> k) & 1u) queryKmers[t] = INVALID;
        __syncwarp();
        // the database window (:210-215) and its 7-mers without 'N' (:131-158): every one found (once) in the query votes for
        // firstBaseOffset = databaseOffset - queryOffset + queryLength
        const unsigned databaseLength = p.sequenceLength + 15u;
        const uint64_t databaseBegin = ref.contigOffset[p.contigId] + uint64_t(p.strandPosition - long(p.left));
        const unsigned databaseKmers = databaseLength - AVOID_KMER + 1;
        for (unsigned d = lane; d < databaseKmers; d += 32u)
        {
            const uint32_t v = uint32_t(referenceCodes16(ref, databaseBegin + d)) & KMER_MASK;
            if (v & NOT_ACGT) continue;
            for (unsigned t = 0; t < ownerKmers; ++t)
                if (queryKmers[t] == v)
                {
                    const int offset = int(d) - int(t) + int(p.sequenceLength);
                    if (offset >= 0 && unsigned(offset) < votesSize) atomicAdd(&votes[offset], 1u);
                    break;
                }
        }
        __syncwarp();
        // two different offsets confirmed by SUFFICIENT_NUMBER_OF_HITS each (:148-158)
        unsigned confirmed = 0;
        for (unsigned k = lane; k < votesSize; k += 32u) confirmed += votes[k] >= AVOID_SUFFICIENT_HITS ? 1u : 0u;
        for (unsigned s = 16; s; s >>= 1) confirmed += __shfl_xor_sync(0xFFFFFFFFu, confirmed, s);
        if (lane == 0) prepOut[j] = word | (confirmed >= 2u ? 0u : PREP_SKIP_SW);
        __syncwarp();
    }
}

} // namespace isaac_b200
