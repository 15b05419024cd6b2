// Compaction of the fixed-stride CIGAR slots a kernel pass wrote into a dense pool (deterministic 3-pass exclusive scan
// over cigarLength): the end-to-end entry points return 64-byte records + a dense CIGAR pool instead of 'stride' words
// per candidate, which is what bounds them on PCIe.
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

constexpr unsigned COMPACT_BLOCK = 256, COMPACT_ITEMS = 4;      // 1024 records per block

__global__ void __launch_bounds__(COMPACT_BLOCK)
cigarBlockSumsKernel(uint32_t n, const isaac_ext_fragment_t *__restrict__ fragments, uint32_t *__restrict__ blockSums)
{
    __shared__ uint32_t warpSums[COMPACT_BLOCK / 32];
    const uint32_t first = (blockIdx.x * COMPACT_BLOCK + threadIdx.x) * COMPACT_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (unsigned k = 0; k < COMPACT_ITEMS; ++k) if (first + k < n) s += fragments[first + k].cigarLength;
    for (unsigned d = 16; d; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31u) == 0) warpSums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (unsigned w = 0; w < COMPACT_BLOCK / 32; ++w) t += warpSums[w];
        blockSums[blockIdx.x] = t;
    }
}

/// exclusive scan of the block sums by one block; total[0] = grand total of this chunk.  'running' (optional) carries the
/// pool offset across the chunks of one call in stream order: chunkBase[0] = words of all earlier chunks, running[0] += total;
/// 'hostTotal' (optional) is a word of mapped pinned memory the host reads after the chunk's event (no copy engine involved).
__global__ void __launch_bounds__(1024) cigarScanBlockSumsKernel(uint32_t blocks, uint32_t *__restrict__ blockSums, uint32_t *__restrict__ total,
                                                                 unsigned long long *__restrict__ running = nullptr,
                                                                 unsigned long long *__restrict__ chunkBase = nullptr,
                                                                 volatile uint32_t *hostTotal = nullptr)
{
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < blocks ? blockSums[i] : 0u;
        uint32_t incl = v;
        for (unsigned d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31u) >= d) incl += o; }
        if ((threadIdx.x & 31u) == 31u) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            uint32_t w = warpSums[threadIdx.x];
            for (unsigned d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, w, d); if (threadIdx.x >= d) w += o; }
            warpSums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warpOffset = (threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u;
        if (i < blocks) blockSums[i] = carry + warpOffset + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += warpOffset + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        total[0] = carry;
        if (running) { chunkBase[0] = running[0]; running[0] += carry; }
        if (hostTotal) { hostTotal[0] = carry; __threadfence_system(); }
    }
}

/// writes the dense pool and points every record at its words (cigarOffset = poolBase + dense offset)
__global__ void __launch_bounds__(COMPACT_BLOCK)
cigarCompactKernel(uint32_t n, isaac_ext_fragment_t *__restrict__ fragments, const uint32_t *__restrict__ strided, uint32_t stride,
                   const uint32_t *__restrict__ blockOffsets, uint32_t *__restrict__ pool, uint32_t poolCapacity,
                   const unsigned long long *__restrict__ chunkBase = nullptr)
{
    __shared__ uint32_t warpSums[COMPACT_BLOCK / 32];
    const uint32_t first = (blockIdx.x * COMPACT_BLOCK + threadIdx.x) * COMPACT_ITEMS;
    uint32_t len[COMPACT_ITEMS], s = 0;
#pragma unroll
    for (unsigned k = 0; k < COMPACT_ITEMS; ++k) { len[k] = first + k < n ? fragments[first + k].cigarLength : 0u; s += len[k]; }
    uint32_t incl = s;
    for (unsigned d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31u) >= d) incl += o; }
    if ((threadIdx.x & 31u) == 31u) warpSums[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t offset = blockOffsets[blockIdx.x] + incl - s;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) offset += warpSums[w];
    const uint32_t poolBase = chunkBase ? uint32_t(chunkBase[0]) : 0u;      // the caller's pool is addressed with 32 bits
#pragma unroll
    for (unsigned k = 0; k < COMPACT_ITEMS; ++k)
    {
        if (first + k < n)
        {
            if (offset + len[k] <= poolCapacity)
                for (unsigned j = 0; j < len[k]; ++j) pool[offset + j] = strided[size_t(first + k) * stride + j];
            fragments[first + k].cigarOffset = offset + poolBase;
            offset += len[k];
        }
    }
}

/// FragmentBuilder::alignFragments' decision per candidate (FragmentBuilder.cpp:190-209): gapped[i] stays where the reference would
/// have run the gapped aligner on the ungapped alignment and accepts the result, else it is replaced by ungapped[i].  A kept ungapped
/// alignment whose CIGAR is the one its clip counts imply (isaac_ext_alignment_t) gets cigarLength 0 = no words for the pool; the
/// others (soft clips at a contig end are not counted in lowClipped / highClipped, AlignerBase.cpp:50-82) take their words along in
/// the gapped pass's CIGAR row.  status[i] = ISAAC_EXT_ALIGNMENT_ALIGNED / _GAPPED of the kept alignment.
__global__ void selectAlignmentKernel(const ReadSetView reads, uint32_t n, const isaac_ext_fragment_t *__restrict__ ungapped,
                                      const uint32_t *__restrict__ ungappedCigars, isaac_ext_fragment_t *__restrict__ gapped,
                                      uint32_t *__restrict__ gappedCigars, uint32_t gappedStride, uint32_t gappedMismatchesMax,
                                      uint8_t *__restrict__ status)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_fragment_t u = ungapped[i];
        const isaac_ext_fragment_t g = gapped[i];
        const double d = u.logProbability - g.logProbability;
        const bool lpLess = !(0.0000001 >= (d < 0 ? -d : d)) && u.logProbability < g.logProbability;          // ISAAC_LP_LESS (Quality.hh:104-112)
        const unsigned observed = u.cigarLength ? u.observedLength : 0u;
        const bool accept = u.cigarLength && ISAAC_EXT_SW_MISMATCH_CUTOFF < u.mismatchCount &&                 // :179 (unaligned are gone), :190-200
                            g.matchCount && g.matchCount + ISAAC_EXT_BAND_WIDTH > observed && g.mismatchCount <= gappedMismatchesMax &&
                            u.mismatchCount > g.mismatchCount && lpLess;                                       // :202-205
        if (accept) { status[i] = uint8_t(ISAAC_EXT_ALIGNMENT_ALIGNED | ISAAC_EXT_ALIGNMENT_GAPPED); continue; }
        isaac_ext_fragment_t k = u;
        status[i] = u.cigarLength ? uint8_t(ISAAC_EXT_ALIGNMENT_ALIGNED) : uint8_t(0);
        if (u.cigarLength)
        {
            const unsigned L = reads.length(u.readId);
            const unsigned left = u.reverse ? u.highClipped : u.lowClipped, right = u.reverse ? u.lowClipped : u.highClipped;
            uint32_t implied[3] = {0u, 0u, 0u}; unsigned words = 0;
            if (left) implied[words++] = (left << 4) | ISAAC_EXT_CIGAR_SOFT_CLIP;
            if (L > left + right) implied[words++] = ((L - left - right) << 4) | ISAAC_EXT_CIGAR_ALIGN;
            if (right) implied[words++] = (right << 4) | ISAAC_EXT_CIGAR_SOFT_CLIP;
            const uint32_t *actual = ungappedCigars + size_t(i) * 3;
            bool same = words == u.cigarLength;
            for (unsigned w = 0; same && w < words; ++w) same = implied[w] == actual[w];
            if (same) k.cigarLength = 0;
            else for (unsigned w = 0; w < u.cigarLength; ++w) gappedCigars[size_t(i) * gappedStride + w] = actual[w];
        }
        gapped[i] = k;
    }
}

/// the kept 64-byte records as 32-byte isaac_ext_alignment_t; bit 16 of 'flag' if a score does not fit
__global__ void packAlignmentsKernel(uint32_t n, const isaac_ext_fragment_t *__restrict__ kept, const uint8_t *__restrict__ status,
                                     isaac_ext_alignment_t *__restrict__ out, uint32_t *__restrict__ flag)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_fragment_t f = kept[i];
        isaac_ext_alignment_t a;
        a.position = f.position; a.logProbability = f.logProbability; a.observedLength = uint16_t(f.observedLength);
        a.mismatchCount = f.mismatchCount; a.matchesInARow = f.matchesInARow; a.editDistance = f.editDistance;
        a.smithWatermanScore = uint16_t(f.smithWatermanScore); a.lowClipped = f.lowClipped; a.highClipped = f.highClipped;
        a.gapsAndFlags = uint8_t((f.gapCount & ISAAC_EXT_ALIGNMENT_GAPS) | status[i]);
        a.cigarLength = uint8_t(f.cigarLength);
        if (f.smithWatermanScore > 0xFFFFu || f.observedLength > 0xFFFFu || f.gapCount > ISAAC_EXT_ALIGNMENT_GAPS || f.cigarLength > 0xFFu) atomicOr(flag, 16u);
        out[i] = a;
    }
}

/// What validateCandidates does on the host for the one-shot entry points, on the device copy of a chunk: a candidate that
/// names an unknown read / contig or lies outside [-ISAAC_EXT_MAX_CYCLES, contigLength] raises bit 1 / bit 2 of 'flag' (the
/// call then fails as a whole) and is replaced by a harmless one so that the kernels behind never read out of bounds.
__global__ void validateCandidatesKernel(ReferenceView ref, ReadSetView reads, uint32_t n, isaac_ext_candidate_t *__restrict__ candidates,
                                         uint32_t *__restrict__ flag)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = candidates[i];
        const uint32_t contig = c.contigStrand >> 1;
        uint32_t bad = 0;
        if (c.readId >= reads.readTotal || contig >= ref.contigCount) bad = 2u;
        else if (c.position > int64_t(ref.contigLength[contig]) || c.position < -int64_t(ISAAC_EXT_MAX_CYCLES)) bad = 4u;
        if (bad)
        {
            atomicOr(flag, bad);
            isaac_ext_candidate_t z = c;
            z.readId = 0; z.contigStrand = 0; z.position = 0;
            candidates[i] = z;
        }
    }
}

} // namespace isaac_b200
