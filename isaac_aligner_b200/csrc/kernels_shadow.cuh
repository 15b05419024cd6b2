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

// Requests whose window fits one warp's scratch go to shadowCandidatesWarpKernel (below), the others to the CTA kernel; both
// kernels get the whole task list and skip what is not theirs.
constexpr unsigned SHADOW_WARP_WINDOW = 1024;       // window positions a warp handles (hits <= positions)
constexpr unsigned SHADOW_WARP_MAX_READ = 256;      // at most 250 distinct 7-mers in the 512-slot hash of a warp: it can never fill up
__device__ __forceinline__ bool shadowTaskIsSmall(const ShadowTask &task, const unsigned readLength)
{
    return readLength <= SHADOW_WARP_MAX_READ && task.windowEnd - task.windowBegin - long(ISAAC_EXT_SHADOW_KMER) + 1 <= long(SHADOW_WARP_WINDOW);
}

__global__ void __launch_bounds__(SHADOW_BLOCK)
shadowCandidatesKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n, const ShadowTask *__restrict__ tasks,
                       int *__restrict__ scratch /* gridDim.x * SHADOW_SCRATCH */,
                       isaac_ext_candidate_t *__restrict__ pool, uint32_t poolCapacity, unsigned long long *__restrict__ poolSize,
                       uint32_t *__restrict__ taskBegin, uint32_t *__restrict__ taskCount, uint32_t *__restrict__ errorFlag,
                       const uint32_t *__restrict__ taskList = nullptr, const uint32_t *__restrict__ taskListCount = nullptr)
{
    __shared__ uint16_t table[SHADOW_TABLE];
    __shared__ int warpLast[SHADOW_BLOCK / 32];
    __shared__ unsigned warpCount[SHADOW_BLOCK / 32];
    __shared__ int carryLast;
    __shared__ unsigned pushed;
    __shared__ unsigned long long base;         // 64-bit: n requests x up to 10000 candidates can pass 2^32
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int NONE = INT_MIN;
    int *cand = scratch + size_t(blockIdx.x) * SHADOW_SCRATCH;
    for (unsigned i = threadIdx.x; i < SHADOW_TABLE; i += blockDim.x) table[i] = SHADOW_EMPTY;
    __syncthreads();
    // with a task list (the requests the warp kernel left aside: windows beyond a warp's scratch, reads beyond its hash) the CTA
    // walks that list only
    const uint32_t total = taskListCount ? min(*taskListCount, n) : n;
    for (uint32_t k = blockIdx.x; k < total; k += gridDim.x)
    {
        const uint32_t t = taskList ? taskList[k] : k;
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
                base = atomicAdd(poolSize, (unsigned long long)uniqueCount);
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
        if (threadIdx.x == 0) { taskBegin[t] = count && base + uniqueCount <= poolCapacity ? uint32_t(base) : 0u; taskCount[t] = uniqueCount; }
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

// ---- the same for requests with a small rescue window (nearly all of them: the window is the template length range plus
// a read length, a few hundred bases): ONE WARP per request, eight requests per CTA, no block barrier anywhere.
//   * the shadow's 7-mers go into a 512-slot open-addressing hash per warp (key << 16 | first position; atomicMin keeps the
//     first position) instead of the 4^7-entry table, so that 8 warps x 6 KB fit a CTA and the SM holds 32 warps of them;
//   * lane l scans the l-th contiguous piece of the window, which keeps the hits in scan order across the warp; the
//     "drop a hit equal to the previous one" rule (:90-91) is applied inside the lane and then against the last hit of the
//     nearest lane before it that had one;
//   * the kept hits (at most one per window position, far below the 10000 cap) are compacted in order, sorted by a bitonic
//     network in shared memory and made unique in order.
constexpr unsigned SHADOW_WARP_SLOTS = 512;
constexpr unsigned SHADOW_WARPS = 8;

/// first slot of a 7-mer: multiplicative hashing (neighbouring 7-mers of a read share six bases, an xor-fold of their codes
/// sends them to neighbouring slots and linear probing then walks long runs: 3.8 probes per window position measured, 1.3 now)
__device__ __forceinline__ unsigned shadowWarpSlot(const unsigned kmer) { return (kmer * 0x9E3779B1u) >> 23; }
static_assert(SHADOW_WARP_SLOTS == 512, "shadowWarpSlot keeps the top 9 bits");

/// The k-mers here are 14-bit fields cut straight out of the 2-bit packed words (base i at bits 2i; any one-to-one code of a
/// 7-mer serves the hash as long as read and window use the same): no unpacking to one code per nibble on either side.
__global__ void __launch_bounds__(SHADOW_WARPS * 32)
shadowCandidatesWarpKernel(const ReferenceView ref, const ReadSetView reads, uint32_t n, const ShadowTask *__restrict__ tasks, isaac_ext_candidate_t *__restrict__ pool, uint32_t poolCapacity,
                           unsigned long long *__restrict__ poolSize, uint32_t *__restrict__ taskBegin, uint32_t *__restrict__ taskCount,
                           uint32_t *__restrict__ errorFlag, uint32_t *__restrict__ largeTasks, uint32_t *__restrict__ largeTaskCount)
{
    __shared__ uint32_t hashAll[SHADOW_WARPS][SHADOW_WARP_SLOTS];
    __shared__ short candAll[SHADOW_WARPS][SHADOW_WARP_WINDOW];     // a hit = window position - read position: within +-1024
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *hash = hashAll[warp];
    short *cand = candAll[warp];
    const int NONE = INT_MIN;
    const uint32_t EMPTY = 0xFFFFFFFFu;
    for (uint32_t t = blockIdx.x * SHADOW_WARPS + warp; t < n; t += gridDim.x * SHADOW_WARPS)
    {
        const ShadowTask task = tasks[t];
        const unsigned L = reads.length(task.shadowReadId);
        if (!shadowTaskIsSmall(task, L))
        {
            if (lane == 0) largeTasks[atomicAdd(largeTaskCount, 1u)] = t;      // the CTA kernel's share
            continue;
        }
        const uint64_t *strandWords = reads.strandWords2(task.shadowReadId, task.contigStrand & 1u);
        const long window = task.windowEnd - task.windowBegin;
        {
            uint4 *h4 = reinterpret_cast<uint4 *>(hash);
            for (unsigned i = lane; i < SHADOW_WARP_SLOTS / 4; i += 32) h4[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        }
        __syncwarp();
        // ---- hashShadowKmers (:53-72): the first position of every 7-mer of the shadow
        for (unsigned p = lane; p + ISAAC_EXT_SHADOW_KMER <= L; p += 32)
        {
            const uint64_t w0 = strandWords[p >> 4], w1 = strandWords[(p >> 4) + 1];
            const unsigned off = p & 15u;
            const uint32_t nFlags = (uint32_t(w0 >> 32) & 0xFFFFu) | (uint32_t(w1 >> 32) << 16);
            if ((nFlags >> off) & 0x7Fu) continue;                              // a 7-mer with 'n' is never looked up
            const unsigned kmer = __funnelshift_r(uint32_t(w0), uint32_t(w1), off * 2u) & 0x3FFFu;
            const uint32_t entry = (kmer << 16) | p;
            unsigned h = shadowWarpSlot(kmer);
            while (true)
            {
                const uint32_t seen = atomicCAS(&hash[h], EMPTY, entry);
                if (seen == EMPTY) break;
                if ((seen >> 16) == kmer) { atomicMin(&hash[h], entry); break; }
                h = (h + 1u) & (SHADOW_WARP_SLOTS - 1u);
            }
        }
        __syncwarp();
        // ---- findShadowCandidatePositions (:74-102): lane l scans positions [l * piece, (l + 1) * piece) of the window
        const unsigned positions = window >= long(ISAAC_EXT_SHADOW_KMER) ? unsigned(window - long(ISAAC_EXT_SHADOW_KMER) + 1) : 0u;
        const unsigned piece = (positions + 31u) / 32u;                         // <= 32
        const unsigned first = min(lane * piece, positions), last = min(first + piece, positions);
        const uint64_t gs = ref.contigOffset[task.contigStrand >> 1] + uint64_t(task.windowBegin) + first;
        short *mine = cand + first;                     // at most 'piece' hits, kept in place until the compaction
        unsigned nKeep = 0;
        int firstHit = NONE, lastHit = NONE;
        if (first < last)
        {
            // the piece's bases (at most 32 + 6) as three words aligned to its first base, its 'N' flags as two
            const uint32_t *b = ref.bases2 + (gs >> 4);
            const uint32_t w0 = __ldg(b), w1 = __ldg(b + 1), w2 = __ldg(b + 2), w3 = __ldg(b + 3);
            const unsigned s2 = (unsigned(gs) & 15u) * 2u;
            const uint32_t a0 = __funnelshift_r(w0, w1, s2), a1 = __funnelshift_r(w1, w2, s2), a2 = __funnelshift_r(w2, w3, s2);
            const uint32_t *m = ref.nmask + (gs >> 5);
            const uint32_t m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
            const unsigned s1 = unsigned(gs) & 31u;
            const uint32_t n0 = __funnelshift_r(m0, m1, s1), n1 = __funnelshift_r(m1, m2, s1);
            const bool anyN = (n0 | n1) != 0u;
            const unsigned count = last - first;
            auto look = [&](const unsigned q, const unsigned kmer) {
                if (anyN && (__funnelshift_r(n0, n1, q) & 0x7Fu)) return;
                unsigned h = shadowWarpSlot(kmer);
                uint32_t seen;
                while ((seen = hash[h]) != EMPTY && (seen >> 16) != kmer) h = (h + 1u) & (SHADOW_WARP_SLOTS - 1u);
                if (seen == EMPTY) return;
                const int hit = int(first + q) - int(seen & 0xFFFFu);                                   // :89
                if (firstHit == NONE) firstHit = hit;
                if (hit != lastHit) mine[nKeep++] = short(hit);                                         // :90-91 inside the lane
                lastHit = hit;
            };
            const unsigned firstHalf = min(count, 16u);
            for (unsigned q = 0; q < firstHalf; ++q) look(q, __funnelshift_r(a0, a1, q * 2u) & 0x3FFFu);
            for (unsigned q = 16; q < count; ++q) look(q, __funnelshift_r(a1, a2, (q - 16u) * 2u) & 0x3FFFu);
        }
        // the last hit of the nearest lane before this one that had a hit: its equal drops this lane's first kept hit
        int incl = lastHit;
#pragma unroll
        for (unsigned d = 1; d < 32; d <<= 1)
        {
            const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d && incl == NONE) incl = o;
        }
        int prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if (lane == 0) prev = NONE;
        const bool dropFirst = firstHit != NONE && firstHit == prev;
        const unsigned kept = nKeep - (dropFirst ? 1u : 0u);
        unsigned scan = kept;
#pragma unroll
        for (unsigned d = 1; d < 32; d <<= 1)
        {
            const unsigned o = __shfl_up_sync(0xFFFFFFFFu, scan, d);
            if (lane >= d) scan += o;
        }
        const unsigned count = __shfl_sync(0xFFFFFFFFu, scan, 31);
        __syncwarp();
        unsigned uniqueCount = 0;
        unsigned long long base = 0;
        if (count <= 32u)
        {
            // ---- the common case: the kept hits fit one register per lane.  Lane k takes the k-th hit in scan order: the lane whose
            // range [scan - kept, scan) holds k, then sort (bitonic network over the lanes) + unique (:105-111)
            const unsigned at = scan - kept;
            // source lane of slot 'lane': the number of lanes whose inclusive count is <= lane
            unsigned src = 0;
#pragma unroll
            for (unsigned d = 16; d; d >>= 1)
            {
                const unsigned probe = src + d - 1u;
                const unsigned upTo = __shfl_sync(0xFFFFFFFFu, scan, probe);
                if (upTo <= lane) src += d;
            }
            src = min(src, 31u);
            const unsigned srcAt = __shfl_sync(0xFFFFFFFFu, at, src), srcFirst = __shfl_sync(0xFFFFFFFFu, first + (dropFirst ? 1u : 0u), src);
            int value = lane < count ? int(cand[srcFirst + (lane - srcAt)]) : int(SHRT_MAX);
#pragma unroll
            for (unsigned k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (unsigned jj = k >> 1; jj > 0; jj >>= 1)
                {
                    const int other = __shfl_xor_sync(0xFFFFFFFFu, value, jj);
                    const bool up = (lane & k) == 0, lower = (lane & jj) == 0;
                    value = (up == lower) ? min(value, other) : max(value, other);
                }
            const int before = __shfl_up_sync(0xFFFFFFFFu, value, 1);
            const bool head = lane < count && (lane == 0 || value != before);
            const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
            uniqueCount = __popc(heads);
            if (lane == 0 && uniqueCount)
            {
                base = atomicAdd(poolSize, (unsigned long long)uniqueCount);
                if (base + uniqueCount > poolCapacity) atomicOr(errorFlag, 2u);
            }
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (head && base + uniqueCount <= poolCapacity)
            {
                isaac_ext_candidate_t c;
                c.position = task.windowBegin + long(value);                                            // :216
                c.readId = task.shadowReadId;
                c.contigStrand = task.contigStrand;
                pool[base + __popc(heads & ((1u << lane) - 1u))] = c;
            }
        }
        else
        {
            // ---- many hits (a low-complexity shadow): compaction in scan order lane after lane (the destination of a lane ends
            // in front of the pieces of the lanes behind it), bitonic sort in shared memory, unique in order
            const unsigned at = scan - kept;
            for (unsigned sl = 0; sl < 32; ++sl)
            {
                const unsigned k = __shfl_sync(0xFFFFFFFFu, kept, sl);
                if (!k) continue;
                const unsigned from = __shfl_sync(0xFFFFFFFFu, first + (dropFirst ? 1u : 0u), sl), to = __shfl_sync(0xFFFFFFFFu, at, sl);
                const short v = lane < k ? cand[from + lane] : short(0);
                __syncwarp();
                if (lane < k) cand[to + lane] = v;
                __syncwarp();
            }
            unsigned n2 = 1;
            while (n2 < count) n2 <<= 1;
            for (unsigned i = count + lane; i < n2; i += 32) cand[i] = SHRT_MAX;
            __syncwarp();
            for (unsigned k = 2; k <= n2; k <<= 1)
                for (unsigned jj = k >> 1; jj > 0; jj >>= 1)
                {
                    for (unsigned i = lane; i < n2; i += 32)
                    {
                        const unsigned l = i ^ jj;
                        if (l > i)
                        {
                            const short x = cand[i], y = cand[l];
                            if (((i & k) == 0) == (x > y)) { cand[i] = y; cand[l] = x; }
                        }
                    }
                    __syncwarp();
                }
            for (unsigned i0 = 0; i0 < count; i0 += 32)
            {
                const unsigned i = i0 + lane;
                const bool head = i < count && (i == 0 || cand[i] != cand[i - 1]);
                uniqueCount += __popc(__ballot_sync(0xFFFFFFFFu, head));
            }
            if (lane == 0 && uniqueCount)
            {
                base = atomicAdd(poolSize, (unsigned long long)uniqueCount);
                if (base + uniqueCount > poolCapacity) atomicOr(errorFlag, 2u);
            }
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (uniqueCount && base + uniqueCount <= poolCapacity)
            {
                unsigned before = 0;
                for (unsigned i0 = 0; i0 < count; i0 += 32)
                {
                    const unsigned i = i0 + lane;
                    const bool head = i < count && (i == 0 || cand[i] != cand[i - 1]);
                    const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
                    if (head)
                    {
                        isaac_ext_candidate_t c;
                        c.position = task.windowBegin + long(cand[i]);                                  // :216
                        c.readId = task.shadowReadId;
                        c.contigStrand = task.contigStrand;
                        pool[base + before + __popc(heads & ((1u << lane) - 1u))] = c;
                    }
                    before += __popc(heads);
                }
            }
        }
        if (lane == 0) { taskBegin[t] = count && base + uniqueCount <= poolCapacity ? uint32_t(base) : 0u; taskCount[t] = uniqueCount; }
        __syncwarp();
    }
}

} // namespace isaac_b200
