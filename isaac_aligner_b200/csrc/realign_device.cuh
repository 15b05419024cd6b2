// SURVEY 8(f) #4, last part: build::GapRealigner (GapRealigner.cpp:86-1290, gapRealigner/OverlappingGapsFilter.cpp:35-162,
// build/SemialignedEndsClipper.cpp:34-165) for one template of a bin -- what one thread of realignBinKernel runs.
//
// The reference walks a bin's index and realigns fragment after fragment against the gaps every fragment of the bin brought
// along.  Nothing a realign call writes is read by another call except through the mate link (updatePairDetails), and the gap lists
// are final before the first call, so the unit of parallel work is the template: its (up to) two index entries in index order.
// Per fragment: the gaps that touch its reference span, the combinations of them that do not overlap (OverlappingGapsFilter), for
// every combination and every gap of it as the pivot that stays put (before / after the gap) the start position that keeps the
// pivot base in place, the cost of that placement (mismatches of the pieces between the gaps + gap costs), the cheapest one applied:
// new CIGAR, gaps at the ends folded into soft clips, semialigned ends clipped, the pair's TLEN / proper-pair flag / mate position
// refreshed.
//
// Layout decisions that differ from the reference's pointer structures:
//   * positions are int64 P = ReferencePosition::getValue() >> 1 = (contig + 1) << 40 | position: order, +, - of the reference's class
//     are the integer ones (every position a call touches lies on the fragment's contig);
//   * a record may start at any byte of the bin (records lie back to back), so header fields and CIGAR words are read and written
//     byte by byte; the fragment's bases are packed once per call into 2-bit words + an N mask and every mismatch count is 16 bases
//     per XOR against the resident packed reference (the reference compares chars, GapRealigner.cpp:191-218);
//   * the found gaps are never copied out of the bin's two sorted lists unless there are few enough to be used (30): the list the
//     reference sorts and de-duplicates (:129-146) is "the deletions that begin before the span and end inside, by (start, length),
//     then every gap that begins inside";
//   * a CIGAR being built lives in thread-local words; the final one of an entry goes to a pool slot taken with one atomic add.
//
// Host + device: tests/cpp/test_realign_host.cu runs the same functions on the CPU against the reference's own classes.
#pragma once
#include <cstdint>
#include "device_types.cuh"

#ifndef ISAAC_HD
#ifdef __CUDACC__
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif
#endif

namespace isaac_b200
{

constexpr unsigned REALIGN_MAX_GAPS = 30;            // OverlappingGapsFilter::MAX_TRACKED_DELETIONS (OverlappingGapsFilter.hh:36)
constexpr unsigned REALIGN_MAX_OVERLAPS = 30;        // ::MAX_TRACKED_OVERLAPS (:35)
constexpr unsigned REALIGN_MAX_GAPS_AT_A_TIME = 10;  // GapRealigner::MAX_GAPS_AT_A_TIME (GapRealigner.hh:130)
constexpr unsigned REALIGN_FOUND_CAPACITY = 100;     // currentAttemptGaps_.reserve(MAX_GAPS_AT_A_TIME * 10) (:181)
constexpr unsigned REALIGN_CIGAR_CAP = 80;           // words of a CIGAR under construction: 2 per gap + clips + the clippers' splits
constexpr unsigned REALIGN_MAX_ORIGINAL_CIGAR = 64;
constexpr unsigned REALIGN_MAX_READ = 512;           // bases of a record the packed read buffer holds
constexpr uint16_t REALIGN_DODGY_ALIGNMENT_SCORE = 0xFFFFu;       // io::FragmentHeader::DODGY_ALIGNMENT_SCORE (Fragment.hh:304)

enum : uint32_t
{
    REALIGN_ERROR_UNSUPPORTED_RECORD = 1,   // read longer than REALIGN_MAX_READ or original CIGAR longer than REALIGN_MAX_ORIGINAL_CIGAR
    REALIGN_ERROR_POOL = 2,                 // the realigned CIGAR pool is full
    REALIGN_ERROR_OVERLAPS = 4,             // more than REALIGN_MAX_OVERLAPS overlap groups (the reference's FiniteCapacityVector asserts)
    REALIGN_ERROR_CIGAR = 8,                // a CIGAR under construction outgrew REALIGN_CIGAR_CAP / an unexpected operation / no mapped base
    REALIGN_ERROR_BARCODE = 16,             // FragmentHeader::barcode_ outside the barcode tables
    REALIGN_ERROR_BOUNDS = 32               // a record or an index entry lies outside the bin's data
};

/// byte offsets of io::FragmentHeader (Fragment.hh:260-404, checked against the reference's struct by tests/test_tile_write_bin_records.py)
enum : unsigned
{
    BIN_BAM_TLEN = 0, BIN_OBSERVED_LENGTH = 4, BIN_F_STRAND_POSITION = 8, BIN_LOW_CLIPPED = 16, BIN_HIGH_CLIPPED = 18,
    BIN_ALIGNMENT_SCORE = 20, BIN_TEMPLATE_ALIGNMENT_SCORE = 22, BIN_MATE_F_STRAND_POSITION = 24, BIN_READ_LENGTH = 32,
    BIN_CIGAR_LENGTH = 34, BIN_GAP_COUNT = 36, BIN_EDIT_DISTANCE = 38, BIN_FLAGS = 40, BIN_BARCODE = 56, BIN_CLUSTER_ID = 72,
    BIN_HEADER_BYTES = 112
};
enum : uint16_t
{
    BIN_FLAG_PAIRED = 1u << 0, BIN_FLAG_UNMAPPED = 1u << 1, BIN_FLAG_MATE_UNMAPPED = 1u << 2, BIN_FLAG_REVERSE = 1u << 3,
    BIN_FLAG_FIRST_READ = 1u << 5, BIN_FLAG_PROPER_PAIR = 1u << 8
};

ISAAC_HD inline uint16_t binGet16(const uint8_t *p) { return uint16_t(p[0] | (p[1] << 8)); }
ISAAC_HD inline uint32_t binGet32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
ISAAC_HD inline uint64_t binGet64(const uint8_t *p) { return uint64_t(binGet32(p)) | (uint64_t(binGet32(p + 4)) << 32); }
ISAAC_HD inline void binPut16(uint8_t *p, uint16_t v) { p[0] = uint8_t(v); p[1] = uint8_t(v >> 8); }
ISAAC_HD inline void binPut32(uint8_t *p, uint32_t v) { p[0] = uint8_t(v); p[1] = uint8_t(v >> 8); p[2] = uint8_t(v >> 16); p[3] = uint8_t(v >> 24); }
ISAAC_HD inline void binPut64(uint8_t *p, uint64_t v) { binPut32(p, uint32_t(v)); binPut32(p + 4, uint32_t(v >> 32)); }

/// FragmentHeader::getTotalLength (Fragment.hh:188-200)
ISAAC_HD inline uint64_t binRecordLength(const uint8_t *record)
{
    return BIN_HEADER_BYTES + uint64_t(binGet16(record + BIN_READ_LENGTH)) + 4ull * binGet16(record + BIN_CIGAR_LENGTH);
}

ISAAC_HD inline int64_t realignP(const uint64_t referencePositionValue) { return int64_t(referencePositionValue >> 1); }
ISAAC_HD inline uint64_t realignValue(const int64_t P) { return uint64_t(P) << 1; }
ISAAC_HD inline uint32_t realignContig(const int64_t P) { return uint32_t(uint64_t(P) >> 40) - 1u; }
ISAAC_HD inline int64_t realignPosition(const int64_t P) { return P & ((int64_t(1) << 40) - 1); }

/// gapRealigner::Gap in P units
struct RealignGap
{
    int64_t pos; int32_t length;
    ISAAC_HD bool isInsertion() const { return length < 0; }
    ISAAC_HD bool isDeletion() const { return length > 0; }
    ISAAC_HD uint32_t size() const { return uint32_t(length < 0 ? -length : length); }
    ISAAC_HD int64_t endPos(const bool fatInsertions) const { return (isDeletion() || fatInsertions) ? pos + int64_t(size()) : pos; }   // Gap.hh:60-63
};

/// what a realign call reads of the bin and where it leaves its results; every pointer is device memory in the kernel, host memory in
/// the CPU harness
struct RealignBinView
{
    uint8_t *data; uint64_t dataBytes;           // the bin's records
    const isaac_ext_bin_index_t *index; uint64_t indexCount;
    const uint32_t *recordIndex;                 // index entry of the record at byte offset o: recordIndex[o >> 6], 0xFFFFFFFF = not indexed
    const isaac_ext_gap_t *gaps;                 // gapGroups_ of every group back to back: by group, start, signed length; unique
    const uint32_t *gapGroupBegin;               // groups + 1
    const isaac_ext_gap_t *deletions;            // deletionEndGroups_: the deletions of every group by end position
    const uint32_t *deletionGroupBegin;
    const uint32_t *barcodeGapGroup;             // or null: one group
    const isaac_ext_tls_t *barcodeTls; uint32_t barcodeCount;
    ReferenceView ref;
    int64_t binStart, binEnd;                    // P
    bool vigorous, dodgy, clipSemialigned;
    unsigned mismatchCost, gapOpenCost, gapExtendCost;
    // results
    uint64_t *position; uint32_t *cigarOffset; uint32_t *cigarLength;
    uint32_t *cigarPool; unsigned long long *cigarPoolUsed; uint64_t cigarPoolCapacity;
    unsigned long long *realignedFragments;
    uint32_t *errorFlags;
    // the records a call updated: byte offset + the REALIGN_CHANGED_BYTES leading header bytes that hold every field store() writes,
    // so that only those travel back to the host (null in the CPU harness, which works on the caller's bytes directly)
    uint64_t *changedOffset; uint8_t *changedHeader; unsigned long long *changedCount;
};
constexpr unsigned REALIGN_CHANGED_BYTES = 42, REALIGN_CHANGED_STRIDE = 48;

ISAAC_HD inline void realignFlag(const RealignBinView &v, const uint32_t bit)
{
#ifdef __CUDA_ARCH__
    atomicOr(v.errorFlags, bit);
#else
    *v.errorFlags |= bit;
#endif
}
ISAAC_HD inline unsigned long long realignTake(unsigned long long *counter, const unsigned long long n)
{
#ifdef __CUDA_ARCH__
    return atomicAdd(counter, n);
#else
    const unsigned long long was = *counter; *counter += n; return was;
#endif
}

ISAAC_HD inline uint32_t realignFunnel(const uint32_t lo, const uint32_t hi, const unsigned shift)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, shift);
#else
    return uint32_t(((uint64_t(hi) << 32) | lo) >> (shift & 31u));
#endif
}
ISAAC_HD inline unsigned realignPopc(const uint32_t x)
{
#ifdef __CUDA_ARCH__
    return unsigned(__popc(x));
#else
    return unsigned(__builtin_popcount(x));
#endif
}

/// the forward bases of a record as the realigner compares them (oligo::getUppercaseBaseFromBcl, Nucleotides.hh:99-102): 2 bits per
/// base + one 'N' bit for the BCL bytes without quality; one spare word behind each array so that a 16-base window may start anywhere
struct RealignRead
{
    uint32_t codes[REALIGN_MAX_READ / 16 + 1];
    uint32_t n[REALIGN_MAX_READ / 32 + 1];
    const uint8_t *bcl;
    unsigned length;
    ISAAC_HD void load(const uint8_t *bases, const unsigned readLength)
    {
        bcl = bases; length = readLength;
        for (unsigned w = 0; w < readLength / 16 + 2 && w < REALIGN_MAX_READ / 16 + 1; ++w) codes[w] = 0;      // the words in use + the spare one
        for (unsigned w = 0; w < readLength / 32 + 2 && w < REALIGN_MAX_READ / 32 + 1; ++w) n[w] = 0;
        for (unsigned i = 0; i < readLength; ++i)
        {
            const uint32_t b = bases[i], called = (b & 0xFCu) != 0u;
            codes[i >> 4] |= (called ? (b & 3u) : 0u) << ((i & 15u) * 2u);
            n[i >> 5] |= (called ^ 1u) << (i & 31u);
        }
    }
    /// 'A' 'C' 'G' 'T' = 0..3, 'N' = 4
    ISAAC_HD unsigned base(const unsigned i) const { const uint8_t b = bcl[i]; return (b & 0xFCu) ? unsigned(b & 3u) : 4u; }
};

/// reference base as the same code, 'N' = 4; g = global base index
ISAAC_HD inline unsigned realignReferenceBase(const ReferenceView &ref, const uint64_t g)
{
    const unsigned n = (ISAAC_VIEW_LOAD(ref.nmask + (g >> 5)) >> (unsigned(g) & 31u)) & 1u;
    return n ? 4u : (ISAAC_VIEW_LOAD(ref.bases2 + (g >> 4)) >> ((unsigned(g) & 15u) * 2u)) & 3u;
}

/// 16 flags at the even bits -> 16 dense bits
ISAAC_HD inline uint32_t realignCompressEven(uint32_t x)
{
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

/// countMismatches (GapRealigner.cpp:191-218): read bases [readOffset, readOffset + length) against the contig of P from P on, cut
/// at the end of the contig; chars differ <=> codes differ ('N' equals 'N')
ISAAC_HD inline unsigned realignCountMismatches(const ReferenceView &ref, const RealignRead &read, const unsigned readOffset,
                                                const int64_t P, unsigned length)
{
    const uint32_t contig = realignContig(P);
    const int64_t position = realignPosition(P), contigLength = int64_t(ref.contigLength[contig]);
    if (position < contigLength) { if (int64_t(length) > contigLength - position) length = unsigned(contigLength - position); }
    else length = 0;          // the reference would read past its vector here; nothing to compare
    if (readOffset >= read.length) return 0;
    if (length > read.length - readOffset) length = read.length - readOffset;
    const uint64_t g0 = ref.contigOffset[contig] + uint64_t(position);
    unsigned mismatches = 0;
    for (unsigned k = 0; k < length; k += 16u)
    {
        const unsigned todo = length - k < 16u ? length - k : 16u;
        const unsigned r = readOffset + k;
        const uint64_t g = g0 + k;
        const uint32_t rw = realignFunnel(read.codes[r >> 4], read.codes[(r >> 4) + 1], (r & 15u) * 2u);
        const uint32_t rn = realignFunnel(read.n[r >> 5], read.n[(r >> 5) + 1], r & 31u);
        const uint32_t dw = realignFunnel(ISAAC_VIEW_LOAD(ref.bases2 + (g >> 4)), ISAAC_VIEW_LOAD(ref.bases2 + (g >> 4) + 1), (unsigned(g) & 15u) * 2u);
        const uint32_t dn = realignFunnel(ISAAC_VIEW_LOAD(ref.nmask + (g >> 5)), ISAAC_VIEW_LOAD(ref.nmask + (g >> 5) + 1), unsigned(g) & 31u);
        const uint32_t x = rw ^ dw;
        const uint32_t codesDiffer = realignCompressEven(x | (x >> 1));
        const uint32_t differ = ((codesDiffer & ~(rn | dn)) | (rn ^ dn)) & ((1u << todo) - 1u);
        mismatches += realignPopc(differ);
    }
    return mismatches;
}

/// a CIGAR being read or built: words in thread-local memory
struct RealignCigar
{
    uint32_t w[REALIGN_CIGAR_CAP];
    unsigned n;
    bool overflow;
    RealignCigar() = default;
    /// only the words in use travel (a plain struct copy moves all REALIGN_CIGAR_CAP of them through thread-local memory)
    ISAAC_HD RealignCigar(const RealignCigar &o) : n(o.n), overflow(o.overflow) { for (unsigned k = 0; k < o.n; ++k) w[k] = o.w[k]; }
    ISAAC_HD RealignCigar &operator=(const RealignCigar &o)
    {
        n = o.n; overflow = o.overflow;
        for (unsigned k = 0; k < o.n; ++k) w[k] = o.w[k];
        return *this;
    }
    ISAAC_HD void clear() { n = 0; overflow = false; }
    ISAAC_HD void push(const uint32_t length, const uint32_t op) { if (n < REALIGN_CIGAR_CAP) w[n++] = (length << 4) | op; else overflow = true; }
    ISAAC_HD uint32_t length(const unsigned i) const { return w[i] >> 4; }
    ISAAC_HD uint32_t op(const unsigned i) const { return w[i] & 0xFu; }
};

/// PackedFragmentBuffer::Index of the entry being realigned (PackedFragmentBuffer.hh:36-91)
struct RealignIndex
{
    int64_t pos;
    RealignCigar cigar;
    bool ownCigar;               // still the record's
    ISAAC_HD unsigned beginClippedLength() const { return (cigar.n && cigar.op(0) == ISAAC_EXT_CIGAR_SOFT_CLIP) ? cigar.length(0) : 0u; }
};

/// the header fields of the record a call works with; written back by store()
struct RealignFragment
{
    uint8_t *record;
    int32_t bamTlen; uint32_t observedLength; int64_t fStrandPosition, mateFStrandPosition;
    uint16_t lowClipped, highClipped, alignmentScore, templateAlignmentScore, readLength, cigarLength, editDistance, flags;
    uint64_t barcode;
    ISAAC_HD void load(uint8_t *r)
    {
        record = r;
        bamTlen = int32_t(binGet32(r + BIN_BAM_TLEN)); observedLength = binGet32(r + BIN_OBSERVED_LENGTH);
        fStrandPosition = realignP(binGet64(r + BIN_F_STRAND_POSITION)); mateFStrandPosition = realignP(binGet64(r + BIN_MATE_F_STRAND_POSITION));
        lowClipped = binGet16(r + BIN_LOW_CLIPPED); highClipped = binGet16(r + BIN_HIGH_CLIPPED);
        alignmentScore = binGet16(r + BIN_ALIGNMENT_SCORE); templateAlignmentScore = binGet16(r + BIN_TEMPLATE_ALIGNMENT_SCORE);
        readLength = binGet16(r + BIN_READ_LENGTH); cigarLength = binGet16(r + BIN_CIGAR_LENGTH);
        editDistance = binGet16(r + BIN_EDIT_DISTANCE); flags = binGet16(r + BIN_FLAGS); barcode = binGet64(r + BIN_BARCODE);
    }
    /// the fields a realign call may change
    ISAAC_HD void store() const
    {
        binPut32(record + BIN_BAM_TLEN, uint32_t(bamTlen)); binPut32(record + BIN_OBSERVED_LENGTH, observedLength);
        binPut64(record + BIN_F_STRAND_POSITION, realignValue(fStrandPosition));
        binPut64(record + BIN_MATE_F_STRAND_POSITION, realignValue(mateFStrandPosition));
        binPut16(record + BIN_EDIT_DISTANCE, editDistance); binPut16(record + BIN_FLAGS, flags);
    }
    ISAAC_HD bool reverse() const { return flags & BIN_FLAG_REVERSE; }
    ISAAC_HD unsigned leftClipped() const { return reverse() ? highClipped : lowClipped; }       // Fragment.hh:262-265
    ISAAC_HD unsigned rightClipped() const { return reverse() ? lowClipped : highClipped; }
    ISAAC_HD const uint8_t *bases() const { return record + BIN_HEADER_BYTES; }
    ISAAC_HD const uint8_t *cigarBytes() const { return record + BIN_HEADER_BYTES + readLength; }
};

/// the gaps of one lookup: RealignerGaps::findGaps (GapRealigner.cpp:99-149)
struct RealignFoundGaps
{
    RealignGap g[REALIGN_MAX_GAPS];
    unsigned count;             // what the reference's range holds; g[] is filled only while count <= REALIGN_MAX_GAPS
};

ISAAC_HD inline bool realignGapLess(const int64_t posA, const int32_t lengthA, const int64_t posB, const int32_t lengthB)
{
    return posA < posB || (posA == posB && lengthA < lengthB);                                   // orderByGapStartAndTypeLength (:46-52)
}

ISAAC_HD inline void realignFindGaps(const RealignBinView &v, const unsigned group, const int64_t rangeBegin, const int64_t rangeEnd,
                                     RealignFoundGaps &found)
{
    // gaps that begin inside: [first not below (rangeBegin, -1000000), first not below (rangeEnd, 0))
    uint32_t lo = v.gapGroupBegin[group], hi = v.gapGroupBegin[group + 1];
    const uint32_t groupEnd = hi;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (realignGapLess(realignP(v.gaps[mid].position), v.gaps[mid].length, rangeBegin, -1000000)) lo = mid + 1; else hi = mid;
    }
    const uint32_t startsBegin = lo;
    hi = groupEnd;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (realignGapLess(realignP(v.gaps[mid].position), v.gaps[mid].length, rangeEnd, 0)) lo = mid + 1; else hi = mid;
    }
    const uint32_t startsEnd = lo;
    // deletions that end inside (rangeBegin, rangeEnd]: the reference probes its end-ordered list with Gap(rangeBegin, 1) and
    // Gap(rangeEnd, 1), whose own ends lie one base further (:115-121)
    lo = v.deletionGroupBegin[group]; hi = v.deletionGroupBegin[group + 1];
    const uint32_t deletionsGroupEnd = hi;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (realignP(v.deletions[mid].position) + v.deletions[mid].length < rangeBegin + 1) lo = mid + 1; else hi = mid;
    }
    const uint32_t endsBegin = lo;
    hi = deletionsGroupEnd;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (realignP(v.deletions[mid].position) + v.deletions[mid].length < rangeEnd + 1) lo = mid + 1; else hi = mid;
    }
    const uint32_t endsEnd = lo;
    const uint32_t starts = startsEnd - startsBegin, ends = endsEnd - endsBegin;
    found.count = 0;
    if (REALIGN_FOUND_CAPACITY < starts + ends) return;                     // "Too many gaps": the range stays empty (:124-127)
    if (!ends || !starts)
    {
        // one list alone is handed on as it lies (the deletions in end order)
        found.count = starts + ends;
        if (found.count > REALIGN_MAX_GAPS) return;
        for (uint32_t k = 0; k < starts; ++k) found.g[k] = RealignGap{realignP(v.gaps[startsBegin + k].position), v.gaps[startsBegin + k].length};
        for (uint32_t k = 0; k < ends; ++k) found.g[k] = RealignGap{realignP(v.deletions[endsBegin + k].position), v.deletions[endsBegin + k].length};
        return;
    }
    // both: sorted and unique (:131-137).  A deletion that ends inside is also in the other list exactly if it begins inside, and
    // the ones that begin before the range sort in front of everything that begins inside
    unsigned early = 0;
    for (uint32_t k = endsBegin; k < endsEnd; ++k) early += realignP(v.deletions[k].position) < rangeBegin;
    found.count = early + starts;
    if (found.count > REALIGN_MAX_GAPS) return;
    unsigned n = 0;
    for (uint32_t k = endsBegin; k < endsEnd; ++k)
    {
        const RealignGap gap{realignP(v.deletions[k].position), v.deletions[k].length};
        if (gap.pos >= rangeBegin) continue;
        unsigned at = n++;
        while (at && realignGapLess(gap.pos, gap.length, found.g[at - 1].pos, found.g[at - 1].length)) { found.g[at] = found.g[at - 1]; --at; }
        found.g[at] = gap;
    }
    for (uint32_t k = 0; k < starts; ++k) found.g[n++] = RealignGap{realignP(v.gaps[startsBegin + k].position), v.gaps[startsBegin + k].length};
}

/// gapRealigner::OverlappingGapsFilter (OverlappingGapsFilter.hh:33-88, .cpp:35-162): masks of gaps of which a combination may hold
/// at most one
struct RealignOverlaps
{
    uint32_t mask[REALIGN_MAX_OVERLAPS];
    unsigned count;
    uint32_t maxChoice;

    ISAAC_HD bool build(const RealignFoundGaps &gaps)
    {
        count = 0;
        maxChoice = gaps.count > REALIGN_MAX_GAPS ? 0u : (1u << gaps.count) - 1u;
        if (!maxChoice) return true;
        // the ends of the gaps by (position, kind and index): deletion ends first (index), deletion starts (1024 + index), insertions
        // (2048 + index); keys relative to the lowest start so that position and tag share a word
        int64_t base = gaps.g[0].pos;
        for (unsigned k = 1; k < gaps.count; ++k) if (gaps.g[k].pos < base) base = gaps.g[k].pos;
        uint64_t events[2 * REALIGN_MAX_GAPS];
        unsigned n = 0;
        auto add = [&](const int64_t pos, const unsigned tag) {
            const uint64_t key = (uint64_t(pos - base) << 12) | tag;
            unsigned at = n++;
            while (at && key < events[at - 1]) { events[at] = events[at - 1]; --at; }
            events[at] = key;
        };
        for (unsigned k = 0; k < gaps.count; ++k)
        {
            if (gaps.g[k].isDeletion()) { add(gaps.g[k].pos, 1024u + k); add(gaps.g[k].endPos(false), k); }
            else add(gaps.g[k].endPos(false), 2048u + k);
        }
        uint32_t lastInsertionMask = 0;
        uint64_t lastInsertionPos = ~0ull;          // the reference starts from a position no gap has
        unsigned openDeletions = 0, openInsertions = 0;
        bool lastWasDeletionClose = true;
        bool fits = true;
        auto push = [&](const uint32_t m) { if (count < REALIGN_MAX_OVERLAPS) mask[count++] = m; else fits = false; };
        push(0);
        for (unsigned e = 0; e < n; ++e)
        {
            const unsigned tag = unsigned(events[e] & 0xFFFu);
            const uint64_t pos = events[e] >> 12;
            uint32_t &back = mask[count - 1];
            if (tag < 1024u)                                                                       // a deletion closes
            {
                const uint32_t gapMask = 1u << tag;
                if (lastWasDeletionClose) back &= ~gapMask;
                else if (openDeletions + openInsertions > 1)
                {
                    push(back & ~lastInsertionMask & ~gapMask);
                    lastInsertionMask = 0; openInsertions = 0;
                }
                else back = 0;
                lastWasDeletionClose = true;
                --openDeletions;
            }
            else if (tag < 2048u)                                                                  // a deletion opens
            {
                const uint32_t gapMask = 1u << (tag - 1024u);
                if (lastInsertionMask && lastInsertionPos != pos)
                {
                    if (openDeletions + openInsertions > 1) push((back & ~lastInsertionMask) | gapMask);
                    else back = gapMask;
                    lastInsertionMask = 0; openInsertions = 0;
                }
                else back |= gapMask;
                ++openDeletions;
                lastWasDeletionClose = false;
            }
            else                                                                                   // an insertion
            {
                const uint32_t gapMask = 1u << (tag - 2048u);
                if (lastInsertionMask && lastInsertionPos != pos)
                {
                    if (openDeletions + openInsertions > 1) push((back & ~lastInsertionMask) | gapMask);
                    else back = gapMask;
                    lastInsertionMask = gapMask; openInsertions = 1;
                }
                else { back |= gapMask; lastInsertionMask |= gapMask; ++openInsertions; }
                lastInsertionPos = pos;
                lastWasDeletionClose = false;
            }
            if (!fits) return false;
        }
        if (openDeletions + openInsertions <= 1) --count;
        return true;
    }
    /// the gaps of the combination that exclude each other, 0 = none (OverlappingGapsFilter.hh:56-67)
    ISAAC_HD uint32_t conflict(const uint32_t combination) const
    {
        for (unsigned k = 0; k < count; ++k)
        {
            const uint32_t both = combination & mask[k];
            if (both && (both & (both - 1u))) return both;
        }
        return 0;
    }
    /// the next combination without a conflict, 0 = no more (:69-84)
    ISAAC_HD uint32_t next(uint32_t combination) const
    {
        uint32_t increment = 1;
        while (combination < maxChoice)
        {
            combination += increment;
            const uint32_t both = conflict(combination);
            if (!both) return combination;
            increment = both & (0u - both);                                                        // 1 << lsbSet
        }
        return 0;
    }
};

struct RealignChoice { unsigned editDistance, mismatches, cost, mappedLength; };

/// one realign call and everything it needs of the thread: the reference's GapRealigner for one fragment
struct RealignWorker
{
    const RealignBinView &v;
    RealignRead read;
    RealignFoundGaps gaps;
    RealignOverlaps overlaps;
    ISAAC_HD explicit RealignWorker(const RealignBinView &view) : v(view) {}

    /// GapRealigner::findStartPos (GapRealigner.cpp:840-969)
    ISAAC_HD bool findStartPos(const uint16_t choice, const int64_t binStart, const int64_t binEnd, const RealignIndex &index,
                               const unsigned pivotGapIndex, const int64_t pivotPos, int64_t &result) const
    {
        int64_t lastGapEndPos = index.pos - int64_t(index.beginClippedLength());
        long offset = long(pivotPos - index.pos);
        for (unsigned k = 0; k < index.cigar.n; ++k)
        {
            if (lastGapEndPos > pivotPos) break;
            const uint32_t length = index.cigar.length(k), op = index.cigar.op(k);
            if (op == ISAAC_EXT_CIGAR_ALIGN) lastGapEndPos += length;
            else if (op == ISAAC_EXT_CIGAR_INSERT) offset += length;
            else if (op == ISAAC_EXT_CIGAR_DELETE)
            {
                lastGapEndPos += length;
                if (lastGapEndPos > pivotPos) return false;                 // an existing deletion spans the pivot
                offset -= length;
            }
            else if (op == ISAAC_EXT_CIGAR_SOFT_CLIP)
            {
                if (!k) offset += length;
                lastGapEndPos += length;
            }
        }
        if (0 > offset) return false;
        int64_t overlapPos = pivotPos;
        unsigned basesLeft = unsigned(offset);
        for (unsigned k = pivotGapIndex; k-- > 0;)
        {
            const RealignGap &gap = gaps.g[k];
            if (choice & (1 << k))
            {
                if (gap.endPos(false) > overlapPos) return false;
                if (gap.isInsertion())
                {
                    const unsigned insertionBases = basesLeft < gap.size() ? basesLeft : gap.size();
                    offset -= insertionBases;
                    basesLeft -= insertionBases;
                    if (!basesLeft) break;
                }
                else
                {
                    offset += gap.size();
                    overlapPos = gap.pos;
                }
            }
        }
        if (binStart + offset > pivotPos) return false;
        if (pivotPos - offset >= binEnd) return false;
        result = pivotPos - offset;
        return true;
    }

    /// GapRealigner::verifyGapsChoice (:492-642)
    ISAAC_HD RealignChoice verifyGapsChoice(const uint16_t choice, const int64_t newBeginPos, const RealignFragment &fragment) const
    {
        RealignChoice ret{0, 0, 0, 0};
        int basesLeft = fragment.readLength;
        int leftClippedLeft = int(fragment.leftClipped());
        const int rightClipped = int(fragment.rightClipped());
        int64_t lastGapEndPos = newBeginPos;
        int64_t lastGapBeginPos = 0;                                        // ReferencePosition(): no gap begins there
        for (unsigned k = 0; k < gaps.count; ++k)
        {
            if (!(choice & (1 << k))) continue;
            const RealignGap &gap = gaps.g[k];
            if (gap.endPos(true) <= lastGapEndPos || gap.pos < lastGapEndPos || gap.pos == lastGapBeginPos) { ret.cost = ~0u; return ret; }
            const int toGap = int(gap.pos - lastGapEndPos);
            const int mappedBases = basesLeft - rightClipped < toGap ? basesLeft - rightClipped : toGap;
            const unsigned length = unsigned(mappedBases - (mappedBases < leftClippedLeft ? mappedBases : leftClippedLeft));
            const unsigned mm = realignCountMismatches(v.ref, read, unsigned(int(fragment.readLength) - basesLeft + leftClippedLeft),
                                                       lastGapEndPos + leftClippedLeft, length);
            ret.mappedLength += length; ret.editDistance += mm; ret.mismatches += mm; ret.cost += mm * v.mismatchCost;
            basesLeft -= mappedBases;
            leftClippedLeft -= leftClippedLeft < mappedBases ? leftClippedLeft : mappedBases;
            unsigned clippedGapLength = 0;
            if (gap.isInsertion())
            {
                const int room = basesLeft - rightClipped, size = int(gap.size());
                clippedGapLength = unsigned(room < size ? room : size);
                basesLeft -= int(clippedGapLength);
                leftClippedLeft -= leftClippedLeft < size ? leftClippedLeft : size;
            }
            else clippedGapLength = leftClippedLeft ? 0u : gap.size();
            ret.editDistance += clippedGapLength;
            ret.cost += clippedGapLength ? (v.gapOpenCost + (clippedGapLength - 1u) * v.gapExtendCost) : 0u;
            lastGapEndPos = gap.endPos(false);
            lastGapBeginPos = gap.pos;
            if (basesLeft == leftClippedLeft + rightClipped) break;
        }
        if (basesLeft > leftClippedLeft + rightClipped)
        {
            const unsigned length = unsigned(basesLeft) - (unsigned(basesLeft) < unsigned(leftClippedLeft) ? unsigned(basesLeft) : unsigned(leftClippedLeft)) - unsigned(rightClipped);
            const int64_t firstUnclippedPos = lastGapEndPos + leftClippedLeft;
            if (realignPosition(firstUnclippedPos) > int64_t(v.ref.contigLength[realignContig(firstUnclippedPos)])) { ret.cost = ~0u; return ret; }
            const unsigned mm = realignCountMismatches(v.ref, read, unsigned(int(fragment.readLength) - basesLeft + leftClippedLeft), firstUnclippedPos, length);
            ret.mappedLength += length; ret.editDistance += mm; ret.mismatches += mm; ret.cost += mm * v.mismatchCost;
        }
        return ret;
    }

    /// GapRealigner::applyChoice (:651-832); on false 'index' is untouched
    ISAAC_HD bool applyChoice(const uint16_t choice, const int64_t binEnd, const int64_t contigEnd, RealignIndex &index,
                              const RealignFragment &fragment) const
    {
        int64_t newBeginPos = index.pos;
        RealignCigar out; out.clear();
        int basesLeft = fragment.readLength;
        int leftClippedLeft = int(fragment.leftClipped());
        const int rightClipped = int(fragment.rightClipped());
        int leftClippedInsertionBases = 0;
        if (fragment.leftClipped()) out.push(fragment.leftClipped(), ISAAC_EXT_CIGAR_SOFT_CLIP);
        int64_t lastGapEndPos = newBeginPos;
        uint32_t lastOperation = 0xFu;                                      // Cigar::UNKNOWN: neither INSERT nor DELETE
        for (unsigned k = 0; k < gaps.count; ++k)
        {
            if (!(choice & (1 << k))) continue;
            const RealignGap &gap = gaps.g[k];
            const int64_t gapClippedBeginPos = gap.pos > newBeginPos ? gap.pos : newBeginPos;
            if (gapClippedBeginPos < lastGapEndPos) return false;          // "Overlapping gaps are not allowed": the reference asserts
            const int toGap = int(gapClippedBeginPos - lastGapEndPos);
            const int mappedBases = basesLeft - rightClipped < toGap ? basesLeft - rightClipped : toGap;
            const unsigned softClippedMappedLength = unsigned(mappedBases - (mappedBases < leftClippedLeft ? mappedBases : leftClippedLeft));
            if (softClippedMappedLength) out.push(softClippedMappedLength, ISAAC_EXT_CIGAR_ALIGN);
            basesLeft -= mappedBases;
            leftClippedLeft -= mappedBases < leftClippedLeft ? mappedBases : leftClippedLeft;
            if (gap.isInsertion())
            {
                const int room = basesLeft - rightClipped, span = int(gap.endPos(true) - gapClippedBeginPos);
                const int clippedGapLength = room < span ? room : span;
                const int softClippedGapLength = clippedGapLength - (clippedGapLength < leftClippedLeft ? clippedGapLength : leftClippedLeft);
                if (softClippedGapLength)
                {
                    if (ISAAC_EXT_CIGAR_INSERT == lastOperation && !mappedBases && out.n)
                        out.w[out.n - 1] = ((out.length(out.n - 1) + uint32_t(softClippedGapLength)) << 4) | ISAAC_EXT_CIGAR_INSERT;
                    else { out.push(uint32_t(softClippedGapLength), ISAAC_EXT_CIGAR_INSERT); lastOperation = ISAAC_EXT_CIGAR_INSERT; }
                }
                basesLeft -= clippedGapLength;
                lastGapEndPos = gapClippedBeginPos;
                leftClippedLeft -= clippedGapLength < leftClippedLeft ? clippedGapLength : leftClippedLeft;
                leftClippedInsertionBases += clippedGapLength - softClippedGapLength;
            }
            else
            {
                const int clippedGapLength = int(gap.endPos(true) - gapClippedBeginPos);
                if (!leftClippedLeft)
                {
                    if (ISAAC_EXT_CIGAR_DELETE == lastOperation && !mappedBases && out.n)
                        out.w[out.n - 1] = ((out.length(out.n - 1) + uint32_t(clippedGapLength)) << 4) | ISAAC_EXT_CIGAR_DELETE;
                    else { out.push(uint32_t(clippedGapLength), ISAAC_EXT_CIGAR_DELETE); lastOperation = ISAAC_EXT_CIGAR_DELETE; }
                }
                else newBeginPos += clippedGapLength;
                lastGapEndPos = gap.endPos(false);
            }
            if (basesLeft == leftClippedLeft + rightClipped) break;
        }
        if (basesLeft > leftClippedLeft + rightClipped)
        {
            const int basesToTheEndOfContig = int(contigEnd - lastGapEndPos - leftClippedLeft);
            const int want = basesLeft - leftClippedLeft - rightClipped;
            const int mappedBases = basesToTheEndOfContig < want ? basesToTheEndOfContig : want;
            if (mappedBases) out.push(uint32_t(mappedBases), ISAAC_EXT_CIGAR_ALIGN);
            basesLeft -= leftClippedLeft + mappedBases;
            leftClippedLeft = 0;
        }
        if (basesLeft) out.push(uint32_t(basesLeft), ISAAC_EXT_CIGAR_SOFT_CLIP);
        newBeginPos += int(fragment.leftClipped()) - leftClippedInsertionBases;
        if (newBeginPos >= binEnd) return false;
        if (out.overflow) { realignFlag(v, REALIGN_ERROR_CIGAR); return false; }
        index.pos = newBeginPos;
        index.cigar = out;
        index.ownCigar = false;
        return true;
    }

    /// GapRealigner::compactCigar (:283-485); on false nothing is touched
    ISAAC_HD bool compactCigar(const int64_t binEnd, RealignIndex &index, RealignFragment &fragment) const
    {
        const RealignCigar &c = index.cigar;
        unsigned first = 0, softClipStart = 0;
        bool needCompacting = false;
        int64_t newPos = index.pos;
        for (; first < c.n; ++first)
        {
            const uint32_t length = c.length(first), op = c.op(first);
            if (op == ISAAC_EXT_CIGAR_ALIGN) break;
            else if (op == ISAAC_EXT_CIGAR_SOFT_CLIP) softClipStart += length;
            else if (op == ISAAC_EXT_CIGAR_INSERT) { needCompacting = true; softClipStart += length; }
            else if (op == ISAAC_EXT_CIGAR_DELETE)
            {
                needCompacting = true;
                if (binEnd <= newPos + int64_t(length)) return false;
                newPos += length;
            }
            else { realignFlag(v, REALIGN_ERROR_CIGAR); return false; }
        }
        if (first == c.n) return false;                                     // soft-clipped to nothing
        unsigned last = c.n - 1u, softClipEnd = 0;
        for (; last != first; --last)
        {
            const uint32_t length = c.length(last), op = c.op(last);
            if (op == ISAAC_EXT_CIGAR_ALIGN) break;
            else if (op == ISAAC_EXT_CIGAR_SOFT_CLIP) softClipEnd += length;
            else if (op == ISAAC_EXT_CIGAR_INSERT) { needCompacting = true; softClipEnd += length; }
            else if (op == ISAAC_EXT_CIGAR_DELETE) needCompacting = true;
            else { realignFlag(v, REALIGN_ERROR_CIGAR); return false; }
        }
        // the edit distance and the observed length of the middle [first, last]
        unsigned newEditDistance = 0, readOffset = softClipStart;
        const int64_t begin = needCompacting ? newPos : index.pos;
        int64_t newEndPos = begin;
        for (unsigned k = first; k <= last; ++k)
        {
            const uint32_t length = c.length(k), op = c.op(k);
            if (op == ISAAC_EXT_CIGAR_ALIGN)
            {
                newEditDistance += realignCountMismatches(v.ref, read, readOffset, newEndPos, length);
                newEndPos += length; readOffset += length;
            }
            else if (op == ISAAC_EXT_CIGAR_INSERT) { newEditDistance += length; readOffset += length; }
            else if (op == ISAAC_EXT_CIGAR_DELETE) { newEditDistance += length; newEndPos += length; }
            else { realignFlag(v, REALIGN_ERROR_CIGAR); return false; }
        }
        if (needCompacting)
        {
            RealignCigar out; out.clear();
            if (softClipStart) out.push(softClipStart, ISAAC_EXT_CIGAR_SOFT_CLIP);
            for (unsigned k = first; k <= last; ++k) out.push(c.length(k), c.op(k));
            if (softClipEnd) out.push(softClipEnd, ISAAC_EXT_CIGAR_SOFT_CLIP);
            if (out.overflow) { realignFlag(v, REALIGN_ERROR_CIGAR); return false; }
            index.cigar = out;
            index.pos = newPos;
        }
        fragment.editDistance = uint16_t(newEditDistance);
        fragment.fStrandPosition = index.pos;
        fragment.observedLength = uint32_t(newEndPos - fragment.fStrandPosition);
        return true;
    }

    /// alignment::clipMismatches<5> (Alignment.hh:55-87) with the bin's base extractor: read bases from 'readFrom' in steps of 'step'
    /// (count of them) against reference bases from global index g in the same steps (referenceCount of them).  'N' in the read
    /// matches nothing here (isMatch sees the upper-case 'N'), but does not differ from an 'N' of the reference
    ISAAC_HD void clipMismatches(const long readFrom, const int64_t g, const int step, const unsigned count, const int64_t referenceCount,
                                 unsigned &clippedBases, unsigned &clippedEdits) const
    {
        const unsigned CONSECUTIVE_MATCHES_MIN = 5;
        unsigned matchesInARow = 0, editDistanceMismatches = 0, editDistanceMismatchesUnclipped = 0, ret = 0;
        while (ret != count && int64_t(ret) != referenceCount && CONSECUTIVE_MATCHES_MIN > matchesInARow)
        {
            const unsigned s = read.base(unsigned(readFrom + long(ret) * step));
            const unsigned r = realignReferenceBase(v.ref, uint64_t(g + int64_t(ret) * step));
            const bool differ = s != r;
            if (!differ && r != 4u) { ++matchesInARow; }                    // equal chars that are not 'N': nothing to add to the unclipped count
            else { matchesInARow = 0; editDistanceMismatchesUnclipped = 0; }
            editDistanceMismatches += differ;
            ++ret;
        }
        const bool found = CONSECUTIVE_MATCHES_MIN == matchesInARow;
        clippedBases = found ? ret - matchesInARow : 0u;
        clippedEdits = found ? editDistanceMismatches - editDistanceMismatchesUnclipped : 0u;
    }

    /// build::SemialignedEndsClipper::clip (build/SemialignedEndsClipper.cpp:34-165)
    ISAAC_HD void clipSemialigned(const int64_t binEnd, RealignIndex &index, RealignFragment &fragment) const
    {
        const uint32_t contig = realignContig(index.pos);
        const int64_t contigOffset = int64_t(v.ref.contigOffset[contig]), contigLength = int64_t(v.ref.contigLength[contig]);
        {   // left side
            RealignCigar &c = index.cigar;
            unsigned first = 0, softClippedBeginBases = 0;
            if (c.n && c.op(0) == ISAAC_EXT_CIGAR_SOFT_CLIP) { first = 1; softClippedBeginBases = c.length(0); }
            if (first < c.n && c.op(first) == ISAAC_EXT_CIGAR_ALIGN)
            {
                unsigned mappedBeginBases = c.length(first);
                const int64_t position = realignPosition(index.pos);
                unsigned clippedBases, clippedEdits;
                clipMismatches(long(softClippedBeginBases), contigOffset + position, 1, mappedBeginBases, contigLength - position, clippedBases, clippedEdits);
                if (clippedBases && index.pos + int64_t(clippedBases) < binEnd)
                {
                    softClippedBeginBases += clippedBases; mappedBeginBases -= clippedBases;
                    index.pos += clippedBases;
                    fragment.fStrandPosition += clippedBases; fragment.observedLength -= clippedBases;
                    fragment.editDistance = uint16_t(fragment.editDistance - clippedEdits);
                    RealignCigar out; out.clear();
                    out.push(softClippedBeginBases, ISAAC_EXT_CIGAR_SOFT_CLIP); out.push(mappedBeginBases, ISAAC_EXT_CIGAR_ALIGN);
                    for (unsigned k = first + 1; k < c.n; ++k) out.push(c.length(k), c.op(k));
                    if (out.overflow) realignFlag(v, REALIGN_ERROR_CIGAR);
                    c = out;
                }
            }
        }
        {   // right side
            RealignCigar &c = index.cigar;
            unsigned end = c.n, softClippedEndBases = 0;
            if (end && c.op(end - 1) == ISAAC_EXT_CIGAR_SOFT_CLIP) { --end; softClippedEndBases = c.length(end); }
            if (end && c.op(end - 1) == ISAAC_EXT_CIGAR_ALIGN)
            {
                unsigned mappedEndBases = c.length(end - 1);
                const int64_t endPosition = realignPosition(index.pos) + int64_t(fragment.observedLength);   // one past the last reference base
                unsigned clippedBases, clippedEdits;
                clipMismatches(long(fragment.readLength) - 1 - long(softClippedEndBases), contigOffset + endPosition - 1, -1, mappedEndBases,
                               endPosition, clippedBases, clippedEdits);
                if (clippedBases)
                {
                    softClippedEndBases += clippedBases; mappedEndBases -= clippedBases;
                    fragment.observedLength -= clippedBases;
                    fragment.editDistance = uint16_t(fragment.editDistance - clippedEdits);
                    RealignCigar out; out.clear();
                    for (unsigned k = 0; k + 1 < end; ++k) out.push(c.length(k), c.op(k));
                    out.push(mappedEndBases, ISAAC_EXT_CIGAR_ALIGN); out.push(softClippedEndBases, ISAAC_EXT_CIGAR_SOFT_CLIP);
                    if (out.overflow) realignFlag(v, REALIGN_ERROR_CIGAR);
                    c = out;
                }
            }
        }
    }

    /// TemplateLengthStatistics::checkModel == Nominal for two bin records (TemplateLengthStatistics.hh:104-118,153-176)
    ISAAC_HD bool nominalModel(const RealignFragment &f1, const RealignFragment &f2) const
    {
        if (realignContig(f1.fStrandPosition) != realignContig(f2.fStrandPosition)) return false;
        const isaac_ext_tls_t &tls = v.barcodeTls[f1.barcode];
        const long p1 = long(realignPosition(f1.fStrandPosition)), p2 = long(realignPosition(f2.fStrandPosition));
        const unsigned model = ((p1 <= p2) ? 0u : 4u) | (f1.reverse() ? 2u : 0u) | (f2.reverse() ? 1u : 0u);
        if (model != tls.bestModel[0] && model != tls.bestModel[1]) return false;
        const long l1 = long(f1.observedLength), l2 = long(f2.observedLength);
        const unsigned long length = (p1 < p2) ? (unsigned long)((p2 + l2 - p1) > l1 ? (p2 + l2 - p1) : l1)
                                               : (unsigned long)((p1 + l1 - p2) > l2 ? (p1 + l1 - p2) : l2);
        return !(length > tls.max) && !(length < tls.min);
    }

    /// GapRealigner::updatePairDetails (:222-270); mate = the record at Index::mateDataOffset_ or null when the entry has none
    ISAAC_HD void updatePairDetails(RealignFragment &fragment, RealignFragment *mate) const
    {
        if (!mate || (fragment.flags & BIN_FLAG_MATE_UNMAPPED))
        {
            fragment.bamTlen = fragment.bamTlen < 0 ? int32_t(0u - fragment.observedLength + 1u) : int32_t(fragment.observedLength - 1u);
            if (mate)
            {
                mate->bamTlen = -fragment.bamTlen;
                fragment.mateFStrandPosition = fragment.fStrandPosition;
                mate->mateFStrandPosition = fragment.fStrandPosition;
                mate->fStrandPosition = fragment.fStrandPosition;
            }
            return;
        }
        const int64_t fragmentBeginPos = fragment.fStrandPosition, fragmentEndPos = fragmentBeginPos + int64_t(fragment.observedLength);
        const int64_t mateBeginPos = fragment.mateFStrandPosition, mateEndPos = mateBeginPos + int64_t(mate->observedLength);
        // io::FragmentHeader::getTlen (Fragment.hh:209-226); getLocation is P less a constant
        const uint64_t distance = uint64_t((fragmentEndPos > mateEndPos ? fragmentEndPos : mateEndPos) -
                                           (fragmentBeginPos < mateBeginPos ? fragmentBeginPos : mateBeginPos));
        const bool firstRead = fragment.flags & BIN_FLAG_FIRST_READ;
        const long tlen = fragmentBeginPos < mateBeginPos ? long(distance) : (fragmentBeginPos > mateBeginPos || !firstRead) ? long(0ull - distance) : long(distance);
        fragment.bamTlen = int32_t(tlen);
        mate->bamTlen = -fragment.bamTlen;
        mate->mateFStrandPosition = fragment.fStrandPosition;
        const bool proper = nominalModel(fragment, *mate);
        fragment.flags = uint16_t(proper ? (fragment.flags | BIN_FLAG_PROPER_PAIR) : (fragment.flags & ~BIN_FLAG_PROPER_PAIR));
        mate->flags = uint16_t(proper ? (mate->flags | BIN_FLAG_PROPER_PAIR) : (mate->flags & ~BIN_FLAG_PROPER_PAIR));
    }

    /// GapRealigner::getAlignmentCost (:1031-1046); false = no mapped base to take a percentage of
    ISAAC_HD bool alignmentCost(const RealignFragment &fragment, const RealignIndex &index, unsigned &cost, unsigned &editDistance,
                                int &mismatchesPercent) const
    {
        unsigned gapsCount = 0, mappedLength = 0;
        uint16_t totalGapsLength = 0;
        for (unsigned k = 0; k < index.cigar.n; ++k)
        {
            const uint32_t length = index.cigar.length(k), op = index.cigar.op(k);
            if (op == ISAAC_EXT_CIGAR_ALIGN) mappedLength += length;
            else if (op == ISAAC_EXT_CIGAR_INSERT || op == ISAAC_EXT_CIGAR_DELETE) { totalGapsLength = uint16_t(totalGapsLength + length); ++gapsCount; }
        }
        editDistance = fragment.editDistance;
        const unsigned mismatches = unsigned(fragment.editDistance) - unsigned(totalGapsLength);
        if (!mappedLength) return false;
        mismatchesPercent = int(mismatches * 100u / mappedLength);
        cost = mismatches * v.mismatchCost + gapsCount * v.gapOpenCost + v.gapExtendCost * (unsigned(totalGapsLength) - gapsCount);
        return true;
    }

    /// GapRealigner::realign (:1061-1267) for one index entry.  indexPosition = Index::pos_ as the bin's index holds it; mate = the
    /// record of the entry's mate, loaded, or null.  \return true when the entry leaves with a new CIGAR in 'index'
    ISAAC_HD bool realign(RealignIndex &index, RealignFragment &fragment, RealignFragment *mate)
    {
        bool realigned = false;
        if (fragment.flags & BIN_FLAG_UNMAPPED) return false;
        if (fragment.barcode >= v.barcodeCount) { realignFlag(v, REALIGN_ERROR_BARCODE); return false; }
        const unsigned group = v.barcodeGapGroup ? v.barcodeGapGroup[fragment.barcode] : 0u;
        bool readLoaded = false;
        bool makesSenseToTryAgain;
        do
        {
            makesSenseToTryAgain = false;
            int64_t binEnd = v.binEnd;
            {
                const int64_t contigLength = int64_t(v.ref.contigLength[realignContig(binEnd)]);
                if (realignPosition(binEnd) > contigLength) binEnd = binEnd - realignPosition(binEnd) + contigLength;
            }
            const bool paired = fragment.flags & BIN_FLAG_PAIRED;
            if (!(fragment.editDistance &&
                  (!paired || (!(fragment.flags & BIN_FLAG_MATE_UNMAPPED) && v.binStart <= fragment.mateFStrandPosition && binEnd > fragment.mateFStrandPosition)) &&
                  (v.dodgy || REALIGN_DODGY_ALIGNMENT_SCORE != fragment.alignmentScore || REALIGN_DODGY_ALIGNMENT_SCORE != fragment.templateAlignmentScore) &&
                  realignPosition(index.pos) >= int64_t(index.beginClippedLength())))
                break;
            index.pos = fragment.fStrandPosition;
            // extractRealignmentBounds (:154-189): the span of the read on the reference with its soft clips laid out
            int64_t beginPos = index.pos, endPos = index.pos;
            for (unsigned k = 0; k < index.cigar.n; ++k)
            {
                const uint32_t length = index.cigar.length(k), op = index.cigar.op(k);
                if (op == ISAAC_EXT_CIGAR_ALIGN || op == ISAAC_EXT_CIGAR_DELETE) endPos += length;
                else if (op == ISAAC_EXT_CIGAR_SOFT_CLIP) { if (!k) beginPos -= length; else endPos += length; }
            }
            realignFindGaps(v, group, beginPos, endPos, gaps);
            if (!v.vigorous && REALIGN_MAX_GAPS_AT_A_TIME < gaps.count) break;
            if (!gaps.count || gaps.count > REALIGN_MAX_GAPS) break;                // no combination to try
            if (!overlaps.build(gaps)) { realignFlag(v, REALIGN_ERROR_OVERLAPS); break; }
            unsigned bestEditDistance = 0, bestCost = 0;
            int originalMismatchesPercent = 0;
            if (!alignmentCost(fragment, index, bestCost, bestEditDistance, originalMismatchesPercent)) { realignFlag(v, REALIGN_ERROR_CIGAR); break; }
            if (!readLoaded)
            {
                if (fragment.readLength > REALIGN_MAX_READ) { realignFlag(v, REALIGN_ERROR_UNSUPPORTED_RECORD); break; }
                read.load(fragment.bases(), fragment.readLength);
                readLoaded = true;
            }
            int64_t bestStartPos = index.pos;
            uint32_t bestChoice = 0;
            unsigned evaluatedSoFar = 0;
            for (uint32_t choice = 0; (choice = overlaps.next(choice));)
            {
                if (((1u << REALIGN_MAX_GAPS_AT_A_TIME) - 1u) < evaluatedSoFar++) break;
                for (unsigned pivot = 0; pivot < gaps.count; ++pivot)
                {
                    if (!(choice & (1u << pivot))) continue;
                    const RealignGap &pivotGap = gaps.g[pivot];
                    for (unsigned after = 0; after < 2; ++after)
                    {
                        if (!after && !(pivotGap.pos >= v.binStart)) continue;
                        int64_t newStartPos;
                        if (!findStartPos(uint16_t(choice), v.binStart, binEnd, index, pivot + after, after ? pivotGap.endPos(false) : pivotGap.pos, newStartPos)) continue;
                        const RealignChoice c = verifyGapsChoice(uint16_t(choice), newStartPos, fragment);
                        if (c.mappedLength && (c.cost < bestCost || (c.cost == bestCost && c.editDistance < bestEditDistance)) &&
                            int(c.mismatches * 100u / c.mappedLength) <= originalMismatchesPercent)
                        {
                            bestEditDistance = c.editDistance; bestChoice = choice; bestStartPos = newStartPos; bestCost = c.cost;
                        }
                    }
                }
            }
            if (bestChoice && binEnd > bestStartPos)
            {
                RealignIndex tmp = index;
                tmp.pos = bestStartPos;
                const int64_t contigEnd = binEnd - realignPosition(binEnd) + int64_t(v.ref.contigLength[realignContig(binEnd)]);
                if (applyChoice(uint16_t(bestChoice), binEnd, contigEnd, tmp, fragment) && compactCigar(binEnd, tmp, fragment))
                {
                    if (v.clipSemialigned) clipSemialigned(binEnd, tmp, fragment);
                    index = tmp;
                    updatePairDetails(fragment, mate);
                    realigned = true;
                    makesSenseToTryAgain = v.vigorous;
                }
            }
        } while (makesSenseToTryAgain);
        return realigned;
    }
};

/// the index entry i and, when its mate is a later entry of the index, that one: BinSorter::realignGaps for one template
ISAAC_HD inline void realignTemplate(const RealignBinView &v, const uint64_t i)
{
    const uint64_t mateOffset = v.index[i].mateDataOffset, ownOffset = v.index[i].dataOffset;
    const bool hasMate = mateOffset != ownOffset;
    uint32_t mateEntry = hasMate ? v.recordIndex[mateOffset >> 6] : 0xFFFFFFFFu;
    if (mateEntry != 0xFFFFFFFFu && mateEntry < i) return;                  // the mate's thread does both
    RealignWorker worker(v);
    RealignFragment first, second;
    first.load(v.data + ownOffset);
    second.load(v.data + (hasMate ? mateOffset : ownOffset));
    auto run = [&](const uint64_t entry, RealignFragment &fragment, RealignFragment *mate) {
        RealignIndex index;
        index.pos = fragment.fStrandPosition;
        index.ownCigar = true;
        index.cigar.clear();
        bool supported = fragment.cigarLength <= REALIGN_MAX_ORIGINAL_CIGAR;
        if (supported) { for (unsigned k = 0; k < fragment.cigarLength; ++k) index.cigar.w[k] = binGet32(fragment.cigarBytes() + 4u * k); index.cigar.n = fragment.cigarLength; }
        else if (!(fragment.flags & BIN_FLAG_UNMAPPED)) realignFlag(v, REALIGN_ERROR_UNSUPPORTED_RECORD);
        const bool realigned = supported && index.cigar.n && worker.realign(index, fragment, mate);
        v.position[entry] = realignValue(index.pos);
        v.cigarLength[entry] = realigned ? index.cigar.n : fragment.cigarLength;
        v.cigarOffset[entry] = 0xFFFFFFFFu;
        if (realigned)
        {
            const unsigned long long at = realignTake(v.cigarPoolUsed, index.cigar.n);
            if (at + index.cigar.n <= v.cigarPoolCapacity)
            {
                for (unsigned k = 0; k < index.cigar.n; ++k) v.cigarPool[at + k] = index.cigar.w[k];
                v.cigarOffset[entry] = uint32_t(at);
            }
            else realignFlag(v, REALIGN_ERROR_POOL);
            realignTake(v.realignedFragments, 1);
        }
        return realigned;
    };
    // the positions the index holds are the records' as they were loaded: the second entry's must be taken before the first call
    // may move a shadow (updatePairDetails)
    const int64_t secondLoadedPos = hasMate ? second.fStrandPosition : 0;
    bool changed = run(i, first, hasMate ? &second : nullptr);
    if (mateEntry != 0xFFFFFFFFu)
    {
        // the mate's own call sees the record as the first call left it; its Index::pos_ is the loaded one until realign() refreshes it
        const bool secondChanged = run(mateEntry, second, &first);
        changed |= secondChanged;
        if (!secondChanged && (second.flags & BIN_FLAG_UNMAPPED)) v.position[mateEntry] = realignValue(secondLoadedPos);
    }
    if (!changed) return;
    first.store();
    if (hasMate) second.store();
    if (v.changedCount)
    {
        const unsigned n = hasMate ? 2u : 1u;
        const unsigned long long at = realignTake(v.changedCount, n);
        for (unsigned k = 0; k < n; ++k)
        {
            const RealignFragment &f = k ? second : first;
            v.changedOffset[at + k] = uint64_t(f.record - v.data);
            for (unsigned b = 0; b < REALIGN_CHANGED_BYTES; ++b) v.changedHeader[(at + k) * REALIGN_CHANGED_STRIDE + b] = f.record[b];
        }
    }
}

} // namespace isaac_b200
