// Scoring of a CIGAR against the resident reference (AlignerBase::updateFragmentCigar) and the clipping rules of
// AlignerBase, as device functions shared by the ungapped, gapped and simple-indel kernels.
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

/// The mutable part of FragmentMetadata a kernel thread works on (FragmentMetadata.hh:330-414).
struct FragmentState
{
    int64_t position;
    uint32_t lowClipped, highClipped;
    bool reverse;

    // FragmentMetadata::incrementClipLeft / incrementClipRight (FragmentMetadata.hh:284-285)
    __device__ __forceinline__ void incrementClipLeft(unsigned bases)
    {
        position += bases;
        if (reverse) highClipped += bases; else lowClipped += bases;
    }
    __device__ __forceinline__ void incrementClipRight(unsigned bases)
    {
        if (reverse) lowClipped += bases; else highClipped += bases;
    }
};

/// AlignerBase::clipReadMasking (AlignerBase.cpp:89-119); begin/end index the strand-order sequence.
/// Read::getBeginCyclesMasked() is always 0 (Read.hh:79).
__device__ __forceinline__ void clipReadMasking(unsigned L, unsigned endCyclesMasked, FragmentState &f, long &begin, long &end)
{
    const long maskedBegin = f.reverse ? long(endCyclesMasked) : 0L;
    const long maskedEnd = f.reverse ? long(L) : long(L) - long(endCyclesMasked);
    if (maskedBegin > begin) { f.incrementClipLeft(unsigned(maskedBegin - begin)); begin = maskedBegin; }
    if (maskedEnd < end) { f.incrementClipRight(unsigned(end - maskedEnd)); end = maskedEnd; }
}

/// AlignerBase::clipReference (AlignerBase.cpp:50-82).  The second branch (:74-81) is reached when quality trimming
/// moved the first unmasked base of a reverse-strand read past the contig end: the position is pulled back to the
/// last contig base and the mapped range becomes empty.  (begin cannot go below 0 for candidates the host accepts,
/// position <= contigLength - 2 or no left masking; the reference would walk off its sequence there.)
__device__ __forceinline__ void clipReference(long referenceSize, FragmentState &f, long &begin, long &end)
{
    const long referenceLeft = referenceSize - f.position;
    if (referenceLeft >= 0)
    {
        if (referenceLeft < end - begin) end = begin + referenceLeft;
        if (0 > f.position) { begin -= f.position; f.position = 0; }
        end = max(end, begin);
    }
    else
    {
        f.position += referenceLeft - 1;
        begin = max(0L, begin + referenceLeft - 2);
        end = begin;
    }
}

/// 16 base codes (4 bits each, strand order) starting at strand position 'pos' of a read strand, from the
/// nibble-packed copy of the read set (ReadSetView::codes4; two spare words per strand make w[1] readable).
__device__ __forceinline__ uint64_t readCodes16(const uint64_t *__restrict__ strandWords, unsigned pos)
{
    const uint64_t *w = strandWords + (pos >> 4);
    const unsigned s = (pos & 15u) * 4u;
    const uint64_t w0 = w[0], w1 = w[1];
    return s ? (w0 >> s) | (w1 << (64u - s)) : w0;
}

__device__ __forceinline__ uint64_t spread2to4(uint32_t x)
{
    uint64_t r = x;
    r = (r | (r << 16)) & 0x0000FFFF0000FFFFull;
    r = (r | (r << 8)) & 0x00FF00FF00FF00FFull;
    r = (r | (r << 4)) & 0x0F0F0F0F0F0F0F0Full;
    r = (r | (r << 2)) & 0x3333333333333333ull;
    return r;
}

__device__ __forceinline__ uint64_t spread1to4(uint32_t m16)
{
    uint64_t r = m16;
    r = (r | (r << 24)) & 0x000000FF000000FFull;
    r = (r | (r << 12)) & 0x000F000F000F000Full;
    r = (r | (r << 6)) & 0x0303030303030303ull;
    r = (r | (r << 3)) & 0x1111111111111111ull;
    return r;
}

/// 16 reference base codes (4 bits each: 0..3, CODE_REF_N for 'N') starting at global base index g.
__device__ __forceinline__ uint64_t referenceCodes16(const ReferenceView &ref, uint64_t g)
{
    const uint32_t *b = ref.bases2 + (g >> 4);
    const uint32_t two = __funnelshift_r(__ldg(b), __ldg(b + 1), (unsigned(g) & 15u) * 2u);
    const uint32_t *m = ref.nmask + (g >> 5);
    const uint32_t n16 = __funnelshift_r(__ldg(m), __ldg(m + 1), unsigned(g) & 31u) & 0xFFFFu;
    return spread2to4(two) | (spread1to4(n16) * uint64_t(CODE_REF_N));
}

/// The score tables in shared memory: CigarScorer reads them with ld.shared.  Every thread of the block must call this
/// before anything else (barrier inside); returns the parameters with the table pointers redirected.
__device__ __forceinline__ ScoreParams stageScoreTables(const ScoreParams &global, double (&tables)[201])
{
    for (unsigned i = threadIdx.x; i < 201; i += blockDim.x) tables[i] = global.logMatch[i];
    __syncthreads();
    ScoreParams sp = global;
    sp.logMatch = tables; sp.logMismatch = tables + 100;
    return sp;
}

/// AlignerBase::updateFragmentCigar (AlignerBase.cpp:121-227) as a walker over the read, 16 bases (one 64-bit word of
/// 4-bit codes) per step:
///   * word level: the CIGAR operations overlapping the word are turned into one-bit-per-nibble masks (mismatch,
///     inserted, not-a-match, run boundary) by XOR-ing the read word with the 16 reference codes fetched for each
///     ALIGN piece; mismatchCount / matchCount / editDistance are popcounts of those masks;
///   * base level: only what must stay sequential is left in the per-base loop -- the reference's left-to-right FP64
///     sum (one shared-memory table lookup + one DADD per base; soft-clipped bases add logMatch, inserted bases add the
///     table's 0.0 entry, AlignerBase.cpp:165-213) and the longest-run-of-matches counter.
/// All threads of a warp run the same trip count whatever their CIGARs look like, and two walkers can be stepped side
/// by side (two independent FP64 chains).
struct CigarScorer
{
    const ReferenceView *ref; const ScoreParams *sp;
    const uint64_t *strandWords; const uint8_t *quality; const uint32_t *cigar; uint64_t *mask;
    uint64_t g, g0;
    double lp;
    unsigned L, nOps, k, remaining, op;
    unsigned matchCount, mismatchCount, matchesInARow, gapCount, editDistance, sws, run;
    uint32_t tableShared;       // shared-window address of the [0,100) match, [100,200) mismatch, [200] = 0.0 table
    bool fresh;

    __device__ __forceinline__ void start(const ReferenceView &r, const ReadSetView &reads, const ScoreParams &s, unsigned readId,
                                          unsigned length, bool reverse, uint64_t contigOffset, long strandPosition,
                                          const uint32_t *ops, unsigned n, uint64_t *maskOut)
    {
        ref = &r; sp = &s; strandWords = reads.strandWords2(readId, reverse); quality = reads.strandQuality(readId, reverse);
        cigar = ops; mask = maskOut; g0 = contigOffset + uint64_t(strandPosition); g = g0; lp = 0.0;
        L = length; nOps = n; k = 0; remaining = 0; op = ISAAC_EXT_CIGAR_SOFT_CLIP; fresh = false;
        matchCount = mismatchCount = matchesInARow = gapCount = editDistance = sws = run = 0;
        tableShared = uint32_t(__cvta_generic_to_shared(s.logMatch));
    }

    __device__ __forceinline__ void stepWord(const unsigned w)
    {
        const unsigned P0 = w * 16u;
        if (P0 >= L) return;
        const unsigned cnt = min(16u, L - P0);
        const uint64_t sword = strandWords[w];
        const uint32_t read2 = uint32_t(sword), readN = uint32_t(sword >> 32);     // 2-bit codes, 'n' flags of the 16 bases
        // per base of the word: codes differ (even bits of neq2, ALIGN pieces only), reference 'N', under an ALIGN operation,
        // inserted, first base of an ALIGN operation
        uint32_t neq2 = 0, refN = 0, aligned = 0, skip = 0, bnd = 0;
        unsigned off = 0;
        while (off < cnt)
        {
            while (remaining == 0 && k < nOps)
            {
                const uint32_t word = cigar[k++];
                const unsigned length = word >> 4;
                op = word & 0xFu;
                if (op == ISAAC_EXT_CIGAR_DELETE)                                                // :192-198
                {
                    g += length; editDistance += length; ++gapCount;
                    sws += sp->gapOpen + min(sp->maxGapExtend, (length - 1) * sp->gapExtend);
                }
                else
                {
                    remaining = length;
                    fresh = true;                                                                // a new run of matches (:158)
                    if (op == ISAAC_EXT_CIGAR_INSERT)                                            // :185-191
                    {
                        editDistance += length; ++gapCount;
                        sws += sp->gapOpen + min(sp->maxGapExtend, (length - 1) * sp->gapExtend);
                    }
                }
            }
            if (remaining == 0) { remaining = cnt - off; op = ISAAC_EXT_CIGAR_SOFT_CLIP; }        // malformed CIGAR: never for our callers
            const unsigned take = min(remaining, cnt - off);
            const uint32_t seg = ((1u << take) - 1u) << off;                                     // take <= 16
            if (op == ISAAC_EXT_CIGAR_ALIGN)
            {
                // 16 reference bases from g on, moved to the piece's place in the word
                const uint32_t *b2 = ref->bases2 + (g >> 4);
                const uint32_t d2 = __funnelshift_r(__ldg(b2), __ldg(b2 + 1), (unsigned(g) & 15u) * 2u);
                const uint32_t *bn = ref->nmask + (g >> 5);
                const uint32_t dN = __funnelshift_r(__ldg(bn), __ldg(bn + 1), unsigned(g) & 31u);
                const uint32_t x = read2 ^ (d2 << (2u * off));
                const uint32_t seg2 = ((take >= 16u ? 0u : (1u << (2u * take))) - 1u) << (2u * off);
                neq2 |= (x | (x >> 1)) & 0x55555555u & seg2;
                refN |= (dN << off) & seg;
                aligned |= seg;
                if (fresh) bnd |= 1u << off;
                g += take;
            }
            else if (op == ISAAC_EXT_CIGAR_INSERT) skip |= seg;
            fresh = false;
            off += take; remaining -= take;
        }
        uint32_t neq = neq2;                                         // even bits -> 16 dense bits
        neq = (neq | (neq >> 1)) & 0x33333333u;
        neq = (neq | (neq >> 2)) & 0x0F0F0F0Fu;
        neq = (neq | (neq >> 4)) & 0x00FF00FFu;
        neq = (neq | (neq >> 8)) & 0x0000FFFFu;
        const uint32_t valid = (1u << cnt) - 1u;
        const uint32_t mism16 = aligned & ~readN & (neq | refN);     // !isMatch (Alignment.hh:44-47): 'n' in the read matches anything
        const uint32_t match16 = aligned & ~mism16;
        editDistance += __popc(aligned & (neq | refN | readN));     // the bytes differ (:176-179): 'n' and 'N' differ from everything
        matchCount += __popc(match16);
        mismatchCount += __popc(mism16);
        if (mask && mism16) mask[P0 >> 6] |= uint64_t(mism16) << (P0 & 63u);   // addMismatchCycle (:171), as a bit over base index
        // ---- the sequential part: the FP64 sum, one table lookup + one DADD per base in read order.  The table index of
        // all 16 bases is prepared as bytes first: quality, + 100 for a mismatch, 200 (the +0.0 entry; the sum never is
        // -0.0, so it is unchanged bit for bit) for inserted bases and the positions past the end of the read.
        const uint4 qv = *reinterpret_cast<const uint4 *>(quality + P0);
        unsigned qw[4] = {qv.x, qv.y, qv.z, qv.w};
        const uint32_t zero16 = skip | (0xFFFFu & ~valid);
        if (zero16)
        {
#pragma unroll
            for (unsigned k = 0; k < 4; ++k)
            {
                const uint32_t bytes = ((((zero16 >> (4u * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
                qw[k] = (qw[k] & ~bytes) | (bytes & 0xC8C8C8C8u);
            }
        }
#pragma unroll
        for (unsigned k = 0; k < 4; ++k)
            qw[k] += ((((mism16 >> (4u * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 100u;
#pragma unroll
        for (unsigned b = 0; b < 16; ++b)
        {
            const unsigned idx = __byte_perm(qw[b >> 2], 0u, 0x4440u + (b & 3u));
            double v;
            asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tableShared + idx * 8u));
            lp += v;
        }
        // ---- longest run of matches (:158-170), on the 16 match bits of the word: 'run' enters from the previous word;
        // a run also ends in front of the first base of every ALIGN operation ('fresh', bits of bnd)
        uint32_t bnd16 = bnd;
        unsigned start = 0;
        while (bnd16)
        {
            const unsigned j = __ffs(bnd16) - 1u;
            bnd16 &= bnd16 - 1u;
            runSegment((match16 >> start) & ((1u << (j - start)) - 1u), j - start);
            run = 0; start = j;
        }
        runSegment(match16 >> start, 16u - start);
    }

    /// the word's match bits [0, len) (bits above are zero) appended to the current run
    __device__ __forceinline__ void runSegment(const uint32_t m, const unsigned len)
    {
        const unsigned lead = __ffs(~m) - 1u;                        // ones at the bottom; <= len
        matchesInARow = max(matchesInARow, run + lead);
        if (lead >= len) { run += len; return; }
        // longest run inside: y_n = starts of runs of at least n ones, y_(n+k) = y_n & (y_k >> n); greedy over k = 8,4,2,1
        const uint32_t p2 = m & (m >> 1), p4 = p2 & (p2 >> 2), p8 = p4 & (p4 >> 4);
        uint32_t cur = ~0u, t; unsigned n = 0;
        t = cur & p8;        if (t) { cur = t; n = 8; }
        t = cur & (p4 >> n); if (t) { cur = t; n += 4; }
        t = cur & (p2 >> n); if (t) { cur = t; n += 2; }
        t = cur & (m >> n);  if (t) { n += 1; }
        matchesInARow = max(matchesInARow, n);
        run = __clz(~(m << (32u - len)));                           // ones at the top of the segment
    }

    /// bit 4b of a nibble-flag word -> bit b
    static __device__ __forceinline__ uint32_t compactNibbleFlags(uint64_t c)
    {
        c = (c | (c >> 3)) & 0x0303030303030303ull;
        c = (c | (c >> 6)) & 0x000F000F000F000Full;
        c = (c | (c >> 12)) & 0x000000FF000000FFull;
        c = (c | (c >> 24)) & 0xFFFFull;
        return uint32_t(c);
    }

    __device__ __forceinline__ unsigned finish(isaac_ext_fragment_t &out)
    {
        // a well-formed CIGAR has no operation left here (trailing deletions are stripped, BandedSmithWaterman.cpp:447-452)
        out.observedLength = uint32_t(g - g0);
        out.logProbability = lp;
        out.mismatchCount = uint16_t(mismatchCount);
        out.matchesInARow = uint16_t(matchesInARow);
        out.gapCount = uint16_t(gapCount);
        out.editDistance = uint16_t(editDistance);
        out.smithWatermanScore = sws + mismatchCount * sp->mismatch;                             // :173
        out.matchCount = uint16_t(matchCount);
        return matchCount;
    }
};

/// updateFragmentCigar for the CIGAR of an UNGAPPED alignment, [begin S] [end - begin M] [L - end S] (UngappedAligner.cpp:64-81):
/// the same result as scoreCigar on those three operations, without walking operations.  Per 16-base word: one XOR of the read's
/// 2-bit codes against the reference window (its words are fetched once, the window runs on contiguously), the masks of the word,
/// popcounts; a word whose aligned bases all match takes a short cut through the counters.  What stays per base is the
/// reference's left-to-right FP64 sum.  \return matchCount
__device__ __forceinline__ unsigned scoreUngapped(const ReferenceView &ref, const ReadSetView &reads, const ScoreParams &sp,
                                                  const unsigned readId, const unsigned L, const bool reverse, const uint64_t contigOffset,
                                                  const long strandPosition, const unsigned begin, const unsigned end,
                                                  isaac_ext_fragment_t &out, uint64_t *mask)
{
    const uint64_t *strandWords = reads.strandWords2(readId, reverse);
    const uint8_t *quality = reads.strandQuality(readId, reverse);
    const uint32_t tableShared = uint32_t(__cvta_generic_to_shared(sp.logMatch));
    // reference index of strand position 0 if the alignment ran on to the left of 'begin' (may lie in front of the packed array
    // for a read clipped at the start of the first contig: such positions are never looked at)
    const long r0 = long(contigOffset) + strandPosition - long(begin);
    double lp = 0.0;
    unsigned matchCount = 0, mismatchCount = 0, editDistance = 0, matchesInARow = 0, run = 0;
    for (unsigned w = 0; w * 16u < L; ++w)
    {
        const unsigned P0 = w * 16u;
        const unsigned cnt = min(16u, L - P0);
        const uint64_t sword = strandWords[w];
        const uint32_t valid = (1u << cnt) - 1u;
        // the aligned bases of the word: [max(begin, P0), min(end, P0 + 16)) - P0
        const unsigned lo = begin > P0 ? min(begin - P0, 16u) : 0u, hi = end > P0 ? min(end - P0, 16u) : 0u;
        const uint32_t aligned = hi > lo ? (((1u << (hi - lo)) - 1u) << lo) : 0u;
        uint32_t mism16 = 0;
        if (aligned)
        {
            const uint32_t read2 = uint32_t(sword), readN = uint32_t(sword >> 32) & 0xFFFFu;
            const long rw = r0 + long(P0);                                   // reference index under bit 0 of the word
            uint32_t d2, dN;
            if (rw >= 0)
            {
                const uint32_t *b2 = ref.bases2 + (rw >> 4);
                d2 = __funnelshift_r(__ldg(b2), __ldg(b2 + 1), (unsigned(rw) & 15u) * 2u);
                const uint32_t *bn = ref.nmask + (rw >> 5);
                dN = __funnelshift_r(__ldg(bn), __ldg(bn + 1), unsigned(rw) & 31u);
            }
            else
            {
                const unsigned back = unsigned(-rw);                         // < 16: the word holds 'begin'
                d2 = __ldg(ref.bases2) << (2u * back);
                dN = __ldg(ref.nmask) << back;
            }
            const uint32_t x = read2 ^ d2;
            uint32_t neq = (x | (x >> 1)) & 0x55555555u;                     // even bits -> 16 dense bits
            neq = (neq | (neq >> 1)) & 0x33333333u;
            neq = (neq | (neq >> 2)) & 0x0F0F0F0Fu;
            neq = (neq | (neq >> 4)) & 0x00FF00FFu;
            neq = (neq | (neq >> 8)) & 0x0000FFFFu;
            const uint32_t differ = aligned & (neq | (dN & 0xFFFFu) | readN); // the bytes differ (:176-179): 'n' and 'N' differ from everything
            mism16 = differ & ~readN;                                        // !isMatch (Alignment.hh:44-47): 'n' in the read matches anything
            const uint32_t match16 = aligned & ~mism16;
            editDistance += __popc(differ);
            mismatchCount += __popc(mism16);
            matchCount += __popc(match16);
            if (mask && mism16) mask[P0 >> 6] |= uint64_t(mism16) << (P0 & 63u);   // addMismatchCycle (:171), as a bit over base index
            // longest run of matches (:158-170): the ALIGN operation starts a new run at 'begin'
            if (begin >= P0 && begin < P0 + 16u) run = 0;
            if (!mism16)
            {
                run += hi - lo;
                matchesInARow = max(matchesInARow, run);
            }
            else
            {
                const uint32_t m = match16 >> lo;                            // the aligned piece, bits [0, len)
                const unsigned len = hi - lo;
                const unsigned lead = __ffs(~m) - 1u;                        // ones at the bottom; < len here
                matchesInARow = max(matchesInARow, run + lead);
                const uint32_t p2 = m & (m >> 1), p4 = p2 & (p2 >> 2), p8 = p4 & (p4 >> 4);
                uint32_t cur = ~0u, t; unsigned n = 0;
                t = cur & p8;        if (t) { cur = t; n = 8; }
                t = cur & (p4 >> n); if (t) { cur = t; n += 4; }
                t = cur & (p2 >> n); if (t) { cur = t; n += 2; }
                t = cur & (m >> n);  if (t) { n += 1; }
                matchesInARow = max(matchesInARow, n);
                run = __clz(~(m << (32u - len)));                           // ones at the top of the piece
            }
        }
        // ---- the sequential part: the FP64 sum, one table lookup + one DADD per base in read order (soft-clipped bases add
        // logMatch, :199-213).  Table index bytes: quality, + 100 for a mismatch, 200 (the +0.0 entry) past the end of the read.
        const uint4 qv = *reinterpret_cast<const uint4 *>(quality + P0);
        unsigned qw[4] = {qv.x, qv.y, qv.z, qv.w};
        if (cnt < 16u)
        {
            const uint32_t zero16 = 0xFFFFu & ~valid;
#pragma unroll
            for (unsigned k = 0; k < 4; ++k)
            {
                const uint32_t bytes = ((((zero16 >> (4u * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
                qw[k] = (qw[k] & ~bytes) | (bytes & 0xC8C8C8C8u);
            }
        }
        if (mism16)
        {
#pragma unroll
            for (unsigned k = 0; k < 4; ++k)
                qw[k] += ((((mism16 >> (4u * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 100u;
        }
#pragma unroll
        for (unsigned b = 0; b < 16; ++b)
        {
            const unsigned idx = __byte_perm(qw[b >> 2], 0u, 0x4440u + (b & 3u));
            double v;
            asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tableShared + idx * 8u));
            lp += v;
        }
    }
    out.position = strandPosition;
    out.observedLength = end - begin;
    out.logProbability = lp;
    out.mismatchCount = uint16_t(mismatchCount);
    out.matchesInARow = uint16_t(matchesInARow);
    out.gapCount = 0;
    out.editDistance = uint16_t(editDistance);
    out.smithWatermanScore = mismatchCount * sp.mismatch;                                        // :173
    out.matchCount = uint16_t(matchCount);
    return matchCount;
}

/// updateFragmentCigar of one fragment.  \return matchCount
__device__ __forceinline__ unsigned scoreCigar(const ReferenceView &ref, const ReadSetView &reads, const ScoreParams &sp,
                                               const unsigned readId, const unsigned L, const bool reverse,
                                               const uint64_t contigOffset, const long strandPosition,
                                               const uint32_t *cigar, const unsigned nOps,
                                               isaac_ext_fragment_t &out, uint64_t *mask)
{
    CigarScorer s;
    s.start(ref, reads, sp, readId, L, reverse, contigOffset, strandPosition, cigar, nOps, mask);
    for (unsigned w = 0; w * 16u < L; ++w) s.stepWord(w);
    out.position = strandPosition;
    return s.finish(out);
}

} // namespace isaac_b200
