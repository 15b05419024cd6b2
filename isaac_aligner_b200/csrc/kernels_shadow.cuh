// Shadow-rescue candidate finder (K5): ShadowAligner::hashShadowKmers + findShadowCandidatePositions
// (reference lib/alignment/ShadowAligner.cpp:53-112) for a batch of rescue requests.
//
// One CTA per request (persistent loop).  The 4^7-entry table of first 7-mer positions of the shadow read lives in
// shared memory (uint16, 32 KB); the rescue window of the resident 2-bit reference is scanned 8 positions per thread
// from one 16-base fetch.  The reference appends candidates in scan order, drops a candidate equal to the previously
// appended one, stops at 10000 and then sorts + uniques; the kernel reproduces exactly that with an ordered block
// compaction (so the 10000 cap cuts at the same place) followed by a bitonic sort and an ordered unique.
#pragma once
#include <climits>
#include "device_types.cuh"
#include "score.cuh"

namespace isaac_b200
{

struct ShadowTask
{
    int64_t windowBegin;      // candidatePositionOffset = max(0, rescue range begin)     (ShadowAligner.cpp:194)
    int64_t windowEnd;        // min(contig length, rescue range end + 1)                 (:197)
    uint32_t shadowReadId;
    uint32_t contigStrand;    // contig << 1 | shadow strand (TemplateLengthStatistics::mateOrientation)
};

constexpr unsigned SHADOW_BLOCK = 128;
constexpr unsigned SHADOW_PER_THREAD = 8;
constexpr unsigned SHADOW_SCRATCH = 16384;        // per-CTA candidate scratch (>= 10000, power of two for the sort)
constexpr unsigned SHADOW_TABLE = 1u << (2 * ISAAC_EXT_SHADOW_KMER);
constexpr uint16_t SHADOW_EMPTY = 0xFFFF;

/// 7-mer starting at nibble 0 of x (first base most significant, like oligo::KmerGenerator); false if it contains N
__device__ __forceinline__ bool kmerOf(uint64_t x, unsigned &kmer)
{
    if (x & 0x4444444ull) return false;            // codes >= 4 ('n' / 'N') have bit 2 set
    unsigned k = 0;
#pragma unroll
    for (unsigned i = 0; i < ISAAC_EXT_SHADOW_KMER; ++i) k = (k << 2) | (unsigned(x >> (4 * i)) & 3u);
    kmer = k;
    return true;
}

__device__ __forceinline__ void bitonicSort(int *a, unsigned n2)
{
    for (unsigned k = 2; k <= n2; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1)
        {
            for (unsigned i = threadIdx.x; i < n2; i += blockDim.x)
            {
                const unsigned l = i ^ j;
                if (l > i)
                {
                    const int x = a[i], y = a[l];
                    if (((i & k) == 0) == (x > y)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(SHADOW_BLOCK)
shadowCandidatesKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n, const ShadowTask *__restrict__ tasks,
                       int *__restrict__ scratch /* gridDim.x * SHADOW_SCRATCH */,
                       isaac_ext_candidate_t *__restrict__ pool, uint32_t poolCapacity, uint32_t *__restrict__ poolSize,
                       uint32_t *__restrict__ taskBegin, uint32_t *__restrict__ taskCount, uint32_t *__restrict__ errorFlag)
{
    __shared__ uint16_t table[SHADOW_TABLE];
    __shared__ int warpLast[SHADOW_BLOCK / 32];
    __shared__ unsigned warpCount[SHADOW_BLOCK / 32];
    __shared__ int carryLast;
    __shared__ unsigned pushed, base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int NONE = INT_MIN;
    int *cand = scratch + size_t(blockIdx.x) * SHADOW_SCRATCH;
    for (unsigned i = threadIdx.x; i < SHADOW_TABLE; i += blockDim.x) table[i] = SHADOW_EMPTY;
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < n; t += gridDim.x)
    {
        const ShadowTask task = tasks[t];
        const unsigned L = reads.length(task.shadowReadId);
        const uint64_t *strandWords = reads.strandCodes(task.shadowReadId, task.contigStrand & 1u);
        const long window = task.windowEnd - task.windowBegin;
        // ---- hashShadowKmers (:53-72): first position of every 7-mer of the shadow; warp 0 walks the read in order
        if (warp == 0 && L >= ISAAC_EXT_SHADOW_KMER)
        {
            for (unsigned p0 = 0; p0 + ISAAC_EXT_SHADOW_KMER <= L; p0 += 32)
            {
                const unsigned p = p0 + lane;
                unsigned kmer = 0;
                const bool valid = p + ISAAC_EXT_SHADOW_KMER <= L && kmerOf(readCodes16(strandWords, p), kmer);
                const unsigned same = __match_any_sync(0xFFFFFFFFu, valid ? kmer : (0x80000000u | lane));
                const bool first = (same & ((1u << lane) - 1u)) == 0;
                if (valid && first && table[kmer] == SHADOW_EMPTY) table[kmer] = uint16_t(p);
                __syncwarp();
            }
        }
        if (threadIdx.x == 0) { carryLast = NONE; pushed = 0; }
        __syncthreads();
        // ---- findShadowCandidatePositions (:74-102): scan the window in order
        const uint64_t g0 = ref.contigOffset[task.contigStrand >> 1] + uint64_t(task.windowBegin);
        const long tile = long(SHADOW_BLOCK) * SHADOW_PER_THREAD;
        for (long tileBegin = 0; tileBegin + long(ISAAC_EXT_SHADOW_KMER) <= window && pushed < ISAAC_EXT_SHADOW_POSITIONS; tileBegin += tile)
        {
            const long first = tileBegin + long(threadIdx.x) * SHADOW_PER_THREAD;
            int hits[SHADOW_PER_THREAD];
            unsigned nHits = 0;
            if (first + long(ISAAC_EXT_SHADOW_KMER) <= window)
            {
                uint64_t x = referenceCodes16(ref, g0 + uint64_t(first));
#pragma unroll
                for (unsigned k = 0; k < SHADOW_PER_THREAD; ++k)
                {
                    unsigned kmer;
                    if (first + long(k) + long(ISAAC_EXT_SHADOW_KMER) <= window && kmerOf(x, kmer))
                    {
                        const uint16_t pos = table[kmer];
                        if (pos != SHADOW_EMPTY) hits[nHits++] = int(first + long(k)) - int(pos);       // :89
                    }
                    x >>= 4;
                }
            }
            // candidate of the last hit before this thread (in scan order): warp scan, then across warps and tiles
            int myLast = nHits ? hits[nHits - 1] : NONE;
            int incl = myLast;
#pragma unroll
            for (unsigned d = 1; d < 32; d <<= 1)
            {
                const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d && incl == NONE) incl = o;
            }
            int prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0) prev = NONE;
            if (lane == 31) warpLast[warp] = incl;
            __syncthreads();
            if (prev == NONE)
            {
                for (int w = int(warp) - 1; w >= 0 && prev == NONE; --w) prev = warpLast[w];
                if (prev == NONE) prev = carryLast;
            }
            // "avoid spurious repetitions of start positions" (:90-91): keep a hit iff it differs from the previous hit
            unsigned keep = 0, nKeep = 0;
            for (unsigned k = 0; k < nHits; ++k)
            {
                if (hits[k] != prev) { keep |= 1u << k; ++nKeep; }
                prev = hits[k];
            }
            unsigned scan = nKeep;
#pragma unroll
            for (unsigned d = 1; d < 32; d <<= 1)
            {
                const unsigned o = __shfl_up_sync(0xFFFFFFFFu, scan, d);
                if (lane >= d) scan += o;
            }
            if (lane == 31) warpCount[warp] = scan;
            __syncthreads();
            unsigned offset = pushed + scan - nKeep;
            for (unsigned w = 0; w < warp; ++w) offset += warpCount[w];
            for (unsigned k = 0; k < nHits; ++k)
                if (keep & (1u << k))
                {
                    if (offset < ISAAC_EXT_SHADOW_POSITIONS) cand[offset] = hits[k];                    // capacity 10000 (:93-97)
                    ++offset;
                }
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1)
            {
                unsigned total = 0;
                for (unsigned w = 0; w < SHADOW_BLOCK / 32; ++w) total += warpCount[w];
                pushed = min(pushed + total, unsigned(ISAAC_EXT_SHADOW_POSITIONS));
                int last = carryLast;
                for (unsigned w = 0; w < SHADOW_BLOCK / 32; ++w) if (warpLast[w] != NONE) last = warpLast[w];
                carryLast = last;
            }
            __syncthreads();
        }
        // ---- sort + unique (:105-111)
        const unsigned count = pushed;
        unsigned n2 = 1;
        while (n2 < count) n2 <<= 1;
        for (unsigned i = count + threadIdx.x; i < n2; i += blockDim.x) cand[i] = INT_MAX;
        __syncthreads();
        if (count > 1) bitonicSort(cand, n2);
        // ordered unique into the global candidate pool
        if (threadIdx.x == 0) pushed = 0;
        __syncthreads();
        unsigned uniqueCount = 0;
        if (count)
        {
            // count first, then reserve, then write
            unsigned mine = 0;
            for (unsigned i = threadIdx.x; i < count; i += blockDim.x) mine += (i == 0 || cand[i] != cand[i - 1]);
            atomicAdd(&pushed, mine);
            __syncthreads();
            uniqueCount = pushed;
            if (threadIdx.x == 0)
            {
                base = atomicAdd(poolSize, uniqueCount);
                if (base + uniqueCount > poolCapacity) atomicOr(errorFlag, 2u);
            }
            __syncthreads();
            if (base + uniqueCount <= poolCapacity)
            {
                // rank of element i among the unique ones = number of unique heads in [0, i]; serial per chunk owner
                // (counts are tiny in practice: a handful of candidates per request)
                for (unsigned i = threadIdx.x; i < count; i += blockDim.x)
                {
                    if (i == 0 || cand[i] != cand[i - 1])
                    {
                        unsigned rank = 0;
                        for (unsigned k = 1; k <= i; ++k) rank += cand[k] != cand[k - 1];
                        isaac_ext_candidate_t c;
                        c.position = task.windowBegin + long(cand[i]);                                  // :216
                        c.readId = task.shadowReadId;
                        c.contigStrand = task.contigStrand;
                        pool[base + rank] = c;
                    }
                }
            }
        }
        if (threadIdx.x == 0) { taskBegin[t] = count ? base : 0u; taskCount[t] = uniqueCount; }
        // ---- clear the table entries this request set
        if (warp == 0 && L >= ISAAC_EXT_SHADOW_KMER)
        {
            for (unsigned p = lane; p + ISAAC_EXT_SHADOW_KMER <= L; p += 32)
            {
                unsigned kmer;
                if (kmerOf(readCodes16(strandWords, p), kmer)) table[kmer] = SHADOW_EMPTY;
            }
        }
        __syncthreads();
    }
}

} // namespace isaac_b200
