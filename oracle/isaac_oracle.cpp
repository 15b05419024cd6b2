/*
 * isaac_oracle.cpp -- scalar CPU restatement of the Isaac candidate-extension path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): it is the checker the CUDA path is compared with and the
 * "port" CPU baseline of bench.py; nothing under isaac_aligner_b200/ may use it.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py diffs every function here against
 * oracle/_ref/libisaac_ref.so (the reference's own sources compiled unmodified, oracle/Makefile) on seeded
 * random batches, and tests/test_reference_goldens.py replays the literal vectors of the reference's CppUnit
 * suites (src/c++/lib/alignment/cppunit/test{BandedSmithWaterman,FragmentBuilder2,SimpleIndelAligner,
 * ShadowAligner}.cpp) through it.
 *
 * Every function cites the reference code it restates as path:line under /root/reference/src/c++/.
 * Nothing here is copied from the reference: the SSE2 kernel is restated lane by lane in scalar form, the
 * iterator-based clipping as index arithmetic.
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "oracle_api.h"

namespace
{

const unsigned BAND = ISAAC_EXT_BAND_WIDTH;

inline uint32_t cigarWord(uint32_t length, uint32_t op) { return (length << 4) | op; }   // Cigar.hh:156-161

/* -------------------------------------------------------------------------------------------------
 * BandedSmithWaterman::align   lib/alignment/BandedSmithWaterman.cpp:84-462
 * ------------------------------------------------------------------------------------------------- */
/// WIDTH = lanes of the band.  WIDTH = 16 is the reference (BandedSmithWaterman.hh:88-89 hard-wires it); any other even width is
/// the same recurrence, end-cell scan and traceback over more lanes: lane j of row i = database position i + WIDTH - 1 - j, the
/// database window has L + WIDTH - 1 bases, the byte-pair merge of the G direction codes (:197) applies to every pair of lanes.
/// The widened band of BASELINE configs[4] (2x250 bp) has no reference behaviour to match; this model is what the CUDA
/// warp-wavefront kernel is compared with, and tests/test_wide_band_oracle.py proves that its WIDTH = 16 instance (the one every
/// other function of this file uses) equals the reference's own code.
template <unsigned WIDTH> struct BandedSwT
{
    int match, mismatch, open, ext;     // open/ext positive as the constructor takes them (:36-43)
    int16_t init;                       // :44  numeric_limits<short>::min() + gapOpenScore
    bool valid;
    std::vector<uint8_t> T;             // 3 direction bytes x 16 lanes per row (:45, :306-308)

    BandedSwT(int m, int mm, int o, int e, unsigned maxReadLength)
        : match(m), mismatch(mm), open(o), ext(e), init(int16_t(-32768 + o)), T(size_t(maxReadLength) * 3 * WIDTH)
    {
        // overflow guard of the constructor (:47-53)
        const int maxScore = std::max(std::max(std::max(std::abs(m), std::abs(mm)), std::abs(o)), std::abs(e));
        valid = !(int(maxReadLength) * maxScore >= std::abs(int(init)));
    }

    static int16_t w16(int v) { return int16_t(v); }        // _mm_add/sub_epi16 and 'short' arithmetic wrap

    /// \return length of the stripped leading deletion (:437-453); appends to cigar
    unsigned align(const char *q, unsigned L, const char *db, std::vector<uint32_t> &cigar)
    {
        const size_t originalSize = cigar.size();
        int16_t G[WIDTH], E[WIDTH], F[WIDTH], nG[WIDTH], nE[WIDTH], nF[WIDTH];
        for (unsigned j = 0; j < WIDTH; ++j) { G[j] = init; E[j] = init; F[j] = 0; }     // :108-114 (F starts at 0)
        G[0] = 0;                                                                         // :115
        // the score byte pair the SSE code builds with unpack(W, B) (:230-244): low byte = score, high byte = 0xFF
        // when the bases differ, 0x00 when equal
        const int16_t wMatch = int16_t(uint16_t(uint8_t(match)));
        const int16_t wMismatch = int16_t(uint16_t(0xFF00u | uint8_t(mismatch)));
        for (unsigned i = 0; i < L; ++i)
        {
            uint8_t *TG = &T[size_t(i) * 3 * WIDTH], *TE = TG + WIDTH, *TF = TG + 2 * WIDTH;
            // F: insertion, from lane j-1 of the previous row, zeros shifted into lane 0 (:132-173)
            for (unsigned j = 0; j < WIDTH; ++j)
            {
                const int16_t gp = j ? G[j - 1] : 0, ep = j ? E[j - 1] : 0, fp = j ? F[j - 1] : 0;
                uint8_t tf = gp < ep ? 1 : 0;                                             // :142-145
                const int16_t a = w16(std::max(gp, ep) - open);                           // :150-151
                const int16_t b = w16(fp - ext);                                          // :154-158
                if (a < b) tf = 2;                                                        // :162-166 (max_epu8: 2 wins)
                nF[j] = std::max(a, b);                                                   // :171
                TF[j] = tf;
            }
            TF[0] = 0;                                                                    // :167
            nF[0] = init;                                                                 // :173
            // G: diagonal, same lane of the previous row (:176-190, :243-244)
            uint8_t tgE[WIDTH], tgF[WIDTH];
            for (unsigned j = 0; j < WIDTH; ++j)
            {
                tgE[j] = G[j] < E[j] ? 1 : 0;
                int16_t g = std::max(G[j], E[j]);
                tgF[j] = g < F[j] ? 2 : 0;
                g = std::max(g, F[j]);
                const bool differ = q[i] != db[i + (WIDTH - 1) - j];                               // raw byte compare (:200-205)
                nG[j] = w16(g + (differ ? wMismatch : wMatch));
            }
            // the direction bytes of G are merged with a 16-bit signed max over BYTE PAIRS (:197), not per byte
            for (unsigned p = 0; p < WIDTH / 2; ++p)
            {
                const int16_t x = int16_t(uint16_t(tgF[2 * p] | (tgF[2 * p + 1] << 8)));
                const int16_t y = int16_t(uint16_t(tgE[2 * p] | (tgE[2 * p + 1] << 8)));
                const uint16_t m = uint16_t(std::max(x, y));
                TG[2 * p] = m & 0xFF;
                TG[2 * p + 1] = m >> 8;
            }
            // E: deletion, serial from lane 15 down to lane 0 (:246-297)
            int16_t g = init, e = init, f = init;
            for (int j = int(WIDTH) - 1; j >= 0; --j)
            {
                int16_t mx = g; uint8_t t = 0;
                if (e > g && e > f) { mx = e; t = 1; }
                else if (f > g) { mx = f; t = 2; }
                nE[j] = mx; TE[j] = t;
                g = w16(nG[j] - open); e = w16(mx - ext); f = w16(nF[j] - open);
            }
            std::memcpy(G, nG, sizeof(G)); std::memcpy(E, nE, sizeof(E)); std::memcpy(F, nF, sizeof(F));
        }
        // end cell: lanes 15..0, matrices G,E,F in that order, strict '>' (:349-379)
        int16_t best = w16(G[WIDTH - 1] - 1);
        int ii = int(L) - 1, jj = ii; unsigned type = 0;
        const int16_t *M[3] = {G, E, F};
        for (int j = int(WIDTH) - 1; j >= 0; --j)
            for (unsigned t = 0; t < 3; ++t)
                if (M[t][j] > best) { best = M[t][j]; jj = j; type = t; }
        // traceback, ops emitted tail first (:381-435)
        static const uint32_t opOf[3] = {ISAAC_EXT_CIGAR_ALIGN, ISAAC_EXT_CIGAR_DELETE, ISAAC_EXT_CIGAR_INSERT};
        unsigned opLength = 0;
        if (jj > 0) cigar.push_back(cigarWord(jj, ISAAC_EXT_CIGAR_DELETE));
        while (ii >= 0 && jj >= 0 && jj <= int(WIDTH) - 1)
        {
            ++opLength;
            const unsigned next = T[(size_t(ii) * 3 + type) * WIDTH + jj];
            if (next != type) { cigar.push_back(cigarWord(opLength, opOf[type])); opLength = 0; }
            if (type == 0) { --ii; } else if (type == 1) { ++jj; } else { --ii; --jj; }
            type = next;
        }
        if (type != 1 && opLength) { cigar.push_back(cigarWord(opLength, opOf[type])); opLength = 0; }
        if (jj < int(WIDTH) - 1) { cigar.push_back(cigarWord(opLength + (WIDTH - 1) - jj, ISAAC_EXT_CIGAR_DELETE)); opLength = 0; }
        // strip the deletion at the alignment start, reverse, strip the one at the end (:437-453)
        unsigned ret = 0;
        if ((cigar.back() & 0xF) == ISAAC_EXT_CIGAR_DELETE) { ret = cigar.back() >> 4; cigar.pop_back(); }
        std::reverse(cigar.begin() + originalSize, cigar.end());
        if ((cigar.back() & 0xF) == ISAAC_EXT_CIGAR_DELETE) cigar.pop_back();
        return ret;
    }
};
typedef BandedSwT<BAND> BandedSw;       // the reference's band

/* -------------------------------------------------------------------------------------------------
 * Quality tables  lib/alignment/Quality.cpp:34-66, include/alignment/Quality.hh:52-86
 * ------------------------------------------------------------------------------------------------- */
struct QualityTables
{
    double logMatch[100], logMismatch[100];
    QualityTables()
    {
        logMatch[0] = std::log(1.0 - std::pow(10.0, 1.0 / -10.0));           // Q0 treated as Q1 (:41-42)
        for (int q = 1; q < 100; ++q) logMatch[q] = std::log(1.0 - std::pow(10.0, double(q) / -10.0));
        logMismatch[0] = logMatch[0];                                          // :60
        for (int q = 1; q < 100; ++q) logMismatch[q] = std::log(std::pow(10.0, double(q) / -10.0) / 3.0);
    }
};
const QualityTables tables;

inline bool lpEquals(double a, double b) { return 0.0000001 >= std::fabs(a - b); }   // Quality.hh:104-107
inline bool lpLess(double a, double b) { return !lpEquals(a, b) && a < b; }          // Quality.hh:109-112

/// isMatch  include/alignment/Alignment.hh:44-47
inline bool isMatch(char readBase, char refBase) { return readBase == 'n' || (readBase == refBase && refBase != 'N'); }

/* -------------------------------------------------------------------------------------------------
 * Read / Cluster   lib/alignment/Read.cpp:32-73, lib/alignment/Cluster.cpp:42-69
 * ------------------------------------------------------------------------------------------------- */
struct Read
{
    std::vector<char> seq[2];       // [0] forward, [1] reverse complement
    std::vector<char> qual[2];      // [1] reversed
    unsigned endCyclesMasked;
    unsigned length() const { return seq[0].size(); }
    void decode(const uint8_t *bcl, unsigned n)
    {
        static const char bases[] = "ACGT";
        for (int s = 0; s < 2; ++s) { seq[s].resize(n); qual[s].resize(n); }
        for (unsigned i = 0; i < n; ++i)
        {
            const uint8_t b = bcl[i];
            const bool isN = !(b & 0xfc);                                      // Nucleotides.hh:91-94
            seq[0][i] = isN ? 'n' : bases[b & 3];
            seq[1][n - 1 - i] = isN ? 'n' : bases[(~b) & 3];
            qual[0][i] = isN ? 2 : (b >> 2);
            qual[1][n - 1 - i] = qual[0][i];
        }
        endCyclesMasked = 0;
    }
};

struct Cluster
{
    Read reads[2];
    uint32_t id;
    void load(const isaac_ext_reads_t *r, uint32_t clusterId)
    {
        id = clusterId;
        const uint32_t total = r->readLength[0] + (r->readCount > 1 ? r->readLength[1] : 0);
        const uint8_t *p = r->bcl + size_t(clusterId) * total;
        for (uint32_t i = 0; i < r->readCount; ++i)
        {
            reads[i].decode(p, r->readLength[i]);
            p += r->readLength[i];
            reads[i].endCyclesMasked = r->endCyclesMasked ? r->endCyclesMasked[size_t(clusterId) * r->readCount + i] : 0;
        }
    }
};

/* -------------------------------------------------------------------------------------------------
 * FragmentMetadata (the fields the path touches)  include/alignment/FragmentMetadata.hh:47-444
 * ------------------------------------------------------------------------------------------------- */
struct Fragment
{
    uint32_t contigId; int64_t position; uint16_t lowClipped, highClipped; uint32_t observedLength;
    uint32_t readIndex; bool reverse; uint32_t cigarOffset, cigarLength; const std::vector<uint32_t> *cigarBuffer;
    uint32_t mismatchCount, matchesInARow, gapCount, editDistance; double logProbability;
    int firstSeedIndex; uint32_t repeatSeedsCount, uniqueSeedCount, nonUniqueFirst, nonUniqueSecond;
    uint32_t smithWatermanScore;
    std::vector<uint16_t> mismatchCycles;
    const Cluster *cluster;

    Fragment(const Cluster *c, const std::vector<uint32_t> *buffer, unsigned ri)
        : contigId(-1U), position(0), lowClipped(0), highClipped(0), observedLength(0), readIndex(ri), reverse(false),
          cigarOffset(0), cigarLength(0), cigarBuffer(buffer), mismatchCount(0), matchesInARow(0), gapCount(0),
          editDistance(0), logProbability(0.0), firstSeedIndex(-1), repeatSeedsCount(0), uniqueSeedCount(0),
          nonUniqueFirst(-1U), nonUniqueSecond(0), smithWatermanScore(0), cluster(c) {}

    const Read &read() const { return cluster->reads[readIndex]; }
    bool isAligned() const { return 0 != cigarLength; }                                     // :248
    void setUnaligned() { cigarBuffer = 0; cigarLength = 0; }                               // :252
    unsigned getObservedLength() const { return isAligned() ? observedLength : 0; }         // :85
    long beginClippedLength() const                                                         // :148-159
    {
        if (cigarBuffer && cigarLength) { const uint32_t w = (*cigarBuffer)[cigarOffset]; if ((w & 0xF) == ISAAC_EXT_CIGAR_SOFT_CLIP) return w >> 4; }
        return 0;
    }
    long endClippedLength() const                                                           // :161-172
    {
        if (cigarBuffer && cigarLength) { const uint32_t w = (*cigarBuffer)[cigarOffset + cigarLength - 1]; if ((w & 0xF) == ISAAC_EXT_CIGAR_SOFT_CLIP) return w >> 4; }
        return 0;
    }
    long unclippedPosition() const { return position - beginClippedLength(); }              // :185-188
    void incrementClipLeft(unsigned short bases) { position += bases; if (reverse) highClipped += bases; else lowClipped += bases; }   // :284
    void incrementClipRight(unsigned short bases) { if (reverse) lowClipped += bases; else highClipped += bases; }                      // :285
    uint16_t &leftClipped() { return reverse ? highClipped : lowClipped; }                  // :293
    uint16_t &rightClipped() { return reverse ? lowClipped : highClipped; }                 // :295
    void resetAlignment(const std::vector<uint32_t> &buffer)                                // :297-313
    {
        position = unclippedPosition();
        cigarOffset = buffer.size(); cigarLength = 0; cigarBuffer = &buffer; observedLength = 0;
        mismatchCycles.clear(); mismatchCount = 0; matchesInARow = 0; gapCount = 0; editDistance = 0;
        logProbability = 0.0; smithWatermanScore = 0;
    }
    void resetClipping() { lowClipped = 0; highClipped = 0; }                               // :314-319
};

struct Scores   // AlignerBase normalised penalties  lib/alignment/fragmentBuilder/AlignerBase.cpp:32-43
{
    unsigned mismatch, gapOpen, gapExtend, maxGapExtend;
    explicit Scores(const isaac_ext_config_t &c)
        : mismatch(c.gapMatchScore - c.gapMismatchScore), gapOpen(c.gapMatchScore - c.gapOpenScore),
          gapExtend(c.gapMatchScore - c.gapExtendScore), maxGapExtend(-c.minGapExtendScore) {}
};

struct Contig { const char *bases; uint64_t length; };

/// AlignerBase::clipReference  AlignerBase.cpp:50-82 ; begin/end are indices into the strand sequence
void clipReference(long referenceSize, Fragment &f, long &begin, long &end)
{
    const long referenceLeft = referenceSize - f.position;
    if (referenceLeft >= 0)
    {
        if (referenceLeft < end - begin) end = begin + referenceLeft;
        if (0 > f.position) { begin -= f.position; f.position = 0; }
        end = std::max(end, begin);
    }
    else
    {
        f.position += referenceLeft - 1;
        begin += referenceLeft - 1;
        --begin;
        end = begin;
    }
}

/// AlignerBase::clipReadMasking  AlignerBase.cpp:89-119 (beginCyclesMasked is always 0, Read.hh:79)
void clipReadMasking(const Read &read, Fragment &f, long &begin, long &end)
{
    const long n = read.length();
    const long maskedBegin = f.reverse ? long(read.endCyclesMasked) : 0;
    const long maskedEnd = f.reverse ? n : n - long(read.endCyclesMasked);
    if (maskedBegin > begin) { f.incrementClipLeft(maskedBegin - begin); begin = maskedBegin; }
    if (maskedEnd < end) { f.incrementClipRight(end - maskedEnd); end = maskedEnd; }
}

/// AlignerBase::updateFragmentCigar  AlignerBase.cpp:121-227
unsigned updateFragmentCigar(const Scores &s, const isaac_ext_reads_t &rm, const Contig &contig, Fragment &f,
                             long strandPosition, const std::vector<uint32_t> &cigarBuffer, unsigned cigarOffset)
{
    const Read &read = f.read();
    const std::vector<char> &sequence = read.seq[f.reverse];
    const std::vector<char> &quality = read.qual[f.reverse];
    const char *ref = contig.bases + strandPosition;
    const unsigned firstCycle = rm.firstCycle[f.readIndex];
    const unsigned lastCycle = firstCycle + rm.readLength[f.readIndex] - 1;
    f.cigarBuffer = &cigarBuffer;
    f.cigarOffset = cigarOffset;
    f.cigarLength = cigarBuffer.size() - cigarOffset;
    unsigned currentBase = 0, matchCount = 0;
    for (unsigned i = 0; i < f.cigarLength; ++i)
    {
        const uint32_t word = cigarBuffer[cigarOffset + i];
        const unsigned length = word >> 4, op = word & 0xF;
        if (op == ISAAC_EXT_CIGAR_ALIGN)
        {
            unsigned run = 0;
            for (unsigned j = 0; j < length; ++j)
            {
                if (isMatch(sequence[currentBase], *ref))
                {
                    ++matchCount; ++run;
                    f.logProbability += tables.logMatch[(unsigned char)quality[currentBase]];
                }
                else
                {
                    f.matchesInARow = std::max(f.matchesInARow, run); run = 0;
                    f.mismatchCycles.push_back(f.reverse ? lastCycle - currentBase : firstCycle + currentBase);
                    ++f.mismatchCount;
                    f.logProbability += tables.logMismatch[(unsigned char)quality[currentBase]];
                    f.smithWatermanScore += s.mismatch;
                }
                if (sequence[currentBase] != *ref) ++f.editDistance;       // Ns count (:175-179)
                ++ref; ++currentBase;
            }
            f.matchesInARow = std::max(f.matchesInARow, run);
        }
        else if (op == ISAAC_EXT_CIGAR_INSERT)
        {
            currentBase += length; f.editDistance += length; ++f.gapCount;
            f.smithWatermanScore += s.gapOpen + std::min(s.maxGapExtend, (length - 1) * s.gapExtend);
        }
        else if (op == ISAAC_EXT_CIGAR_DELETE)
        {
            ref += length; f.editDistance += length; ++f.gapCount;
            f.smithWatermanScore += s.gapOpen + std::min(s.maxGapExtend, (length - 1) * s.gapExtend);
        }
        else   // SOFT_CLIP: clipped bases count as matches for the probability, reference not advanced (:199-213)
        {
            for (unsigned j = 0; j < length; ++j) f.logProbability += tables.logMatch[(unsigned char)quality[currentBase + j]];
            currentBase += length;
        }
    }
    f.observedLength = (ref - contig.bases) - strandPosition;
    f.position = strandPosition;
    return matchCount;
}

/* -------------------------------------------------------------------------------------------------
 * matchSelector::SequencingAdapter  lib/alignment/matchSelector/SequencingAdapter.cpp:30-139
 * ------------------------------------------------------------------------------------------------- */
struct AdapterPort
{
    std::string sequence; bool reverse; unsigned clipLength;          // flowcell::SequencingAdapterMetadata (0 = unbounded)
    std::vector<signed char> kmerPositions;                           // -1 uninitialised, -2 not unique (SequencingAdapter.hh:41-42)
    static const unsigned MATCH_BASES_MIN = 5;                        // adapterMatchBasesMin_ (:40)
    static unsigned translate(char c)                                 // oligo::getTranslator(): everything but ACGT is 4 (Nucleotides.hh:41-59)
    {
        switch (c) { case 'a': case 'A': return 0; case 'c': case 'C': return 1; case 'g': case 'G': return 2; case 't': case 'T': return 3; default: return 4; }
    }
    AdapterPort(const std::string &s, bool r, unsigned clip) : sequence(s), reverse(r), clipLength(clip), kmerPositions(1u << (2 * MATCH_BASES_MIN), -1)
    {
        // oligo::KmerGenerator over the adapter (KmerGenerator.hpp:38-124): every window of 5 bases without a non-ACGT base
        for (size_t i = 0; i + MATCH_BASES_MIN <= sequence.size(); ++i)                               // :40-57
        {
            unsigned kmer = 0; bool valid = true;
            for (unsigned j = 0; j < MATCH_BASES_MIN; ++j) { const unsigned v = translate(sequence[i + j]); valid &= v < 4; kmer = (kmer << 2) | (v & 3); }
            if (!valid) continue;
            signed char &pos = kmerPositions[kmer];
            if (pos == -1) pos = (signed char)i; else pos = -2;
        }
    }
    bool isUnbounded() const { return !clipLength; }
    bool isStrandCompatible(bool strandReverse) const { return !isUnbounded() || strandReverse == reverse; }   // SequencingAdapter.hh:58-61

    /// getMatchRange (:58-139); positions index the strand sequence; \return false = (mismatchBase, mismatchBase)
    bool getMatchRange(const std::vector<char> &seq, long sequenceBegin, long sequenceEnd, long mismatchBase, long &first, long &second) const
    {
        if (sequenceEnd - mismatchBase < long(MATCH_BASES_MIN)) return false;                       // oligo::generateKmer (KmerGenerator.hpp:150-169)
        unsigned short kmer = 0;
        for (unsigned j = 0; j < MATCH_BASES_MIN; ++j) { kmer <<= 2; kmer |= translate(seq[mismatchBase + j]); }   // 'n' = 4 spills into the base before
        kmer &= (1u << (2 * MATCH_BASES_MIN)) - 1;
        const int pos = kmerPositions[kmer];
        if (pos < 0) return false;
        const unsigned mismatchBaseOffset = unsigned(mismatchBase - sequenceBegin);
        const unsigned adapterBasesBeforeSequence = mismatchBaseOffset < unsigned(pos) ? unsigned(pos) - mismatchBaseOffset : 0;
        if (adapterBasesBeforeSequence && isUnbounded()) return false;                              // :74, :127-132
        const long testBase = mismatchBase - (long(pos) - long(adapterBasesBeforeSequence));
        const unsigned testSequenceLength = unsigned(sequenceEnd - testBase);
        const unsigned adapterSequenceSize = unsigned(sequence.size());
        const unsigned leftClippedAdapterLength = adapterSequenceSize - adapterBasesBeforeSequence;
        const unsigned overlapLength = std::min(testSequenceLength, leftClippedAdapterLength);
        if (overlapLength < leftClippedAdapterLength && isUnbounded() && reverse) return false;     // :81-88
        if (!(overlapLength >= MATCH_BASES_MIN &&
              !sequence.compare(adapterBasesBeforeSequence, overlapLength, &seq[testBase], overlapLength))) return false;   // :91-92
        if (reverse)                                                                                // :99-105
        {
            first = isUnbounded() ? sequenceBegin : testBase - long(std::min<unsigned>(unsigned(testBase - sequenceBegin), clipLength - adapterSequenceSize));
            second = testBase + overlapLength;
        }
        else                                                                                        // :107-111
        {
            first = testBase;
            second = isUnbounded() ? sequenceEnd : testBase + long(std::min(overlapLength, clipLength));
        }
        return first != second;
    }
};

/// the adapter list of oracle_set_adapters (process-wide, like the reference build of the checker keeps it)
std::vector<AdapterPort> &portAdapters()
{
    static std::vector<AdapterPort> adapters;
    return adapters;
}

/* -------------------------------------------------------------------------------------------------
 * matchSelector::FragmentSequencingAdapterClipper  lib/alignment/matchSelector/FragmentSequencingAdapterClipper.cpp:40-277
 * ------------------------------------------------------------------------------------------------- */
struct AdapterClipperPort
{
    struct Range { bool initialized, empty; long begin, end; Range() : initialized(false), empty(true), begin(0), end(0) {} } strandRange[2];
    const std::vector<AdapterPort> &adapters;
    AdapterClipperPort() : adapters(portAdapters()) {}

    static unsigned countMatches(const std::vector<char> &seq, long b, long e, const Contig &contig, long referenceBegin)   // Alignment.hh:89-104
    {
        unsigned ret = 0;
        for (long i = b; i < e; ++i) ret += isMatch(seq[i], contig.bases[referenceBegin + (i - b)]);
        return ret;
    }
    /// percentMismatches (:62-71); like the CUDA path, reference positions outside the contig count as mismatches where the
    /// reference would read past its vector (:190-216)
    static unsigned percentMismatches(const std::vector<char> &seq, long b, long e, const Contig &contig, long referenceBegin)
    {
        unsigned mismatches = 0;
        for (long i = b; i < e; ++i)
        {
            const long r = referenceBegin + (i - b);
            mismatches += !(r >= 0 && r < long(contig.length) && isMatch(seq[i], contig.bases[r]));
        }
        return mismatches * 100 / unsigned(e - b);
    }

    void checkInitStrand(const Fragment &f, const Contig &contig)                                   // :102-147
    {
        Range &range = strandRange[f.reverse];
        if (range.initialized) return;
        const std::vector<char> &sequence = f.read().seq[f.reverse];
        long sequenceBegin = 0, sequenceEnd = long(sequence.size());
        // the clipper's own clipReference (:40-59)
        const long referenceLeft = long(contig.length) - f.position;
        if (referenceLeft < sequenceEnd - sequenceBegin) sequenceEnd = sequenceBegin + referenceLeft;
        long newFragmentPos = f.position;
        if (0 > f.position) { sequenceBegin -= f.position; newFragmentPos = 0; }
        range.begin = sequenceEnd; range.end = sequenceBegin;                                       // :124-125
        for (const AdapterPort &adapter : adapters)
        {
            if (!adapter.isStrandCompatible(f.reverse)) continue;
            // findSequencingAdapter (:79-100) from the end of what has been found so far
            const long searchBegin = range.end;
            long reference = newFragmentPos + (searchBegin - sequenceBegin);
            for (long current = searchBegin; current < sequenceEnd; ++current, ++reference)
            {
                if (!isMatch(sequence[current], contig.bases[reference]))
                {
                    long first, second;
                    if (adapter.getMatchRange(sequence, searchBegin, sequenceEnd, current, first, second))
                    {
                        range.begin = std::min(first, range.begin);                                 // :138-139
                        range.end = std::max(second, range.end);
                        break;
                    }
                }
            }
        }
        range.initialized = true;
        range.empty = sequenceBegin == range.end;                                                   // :145
    }

    /// clip + decideWhichSideToClip (:149-277); begin/end index the strand sequence
    void clip(const Contig &contig, Fragment &f, long &begin, long &end) const
    {
        const Range &range = strandRange[f.reverse];
        if (range.empty) return;
        const std::vector<char> &sequence = f.read().seq[f.reverse];
        const long contigPosition = f.position;
        const unsigned backwardsClipped = unsigned(range.begin - begin), forwardsClipped = unsigned(end - range.end);   // :158-159
        bool clipBackwards = backwardsClipped < forwardsClipped;
        const unsigned sequenceLength = unsigned(end - begin);
        bool doClip = true;
        if (backwardsClipped && forwardsClipped && std::abs(int(backwardsClipped - forwardsClipped)) < 9)             // :166
        {
            if (contigPosition >= 0 && contig.length >= uint64_t(contigPosition + sequenceLength))                     // :169
            {
                const long referenceEnd = contigPosition + sequenceLength;
                const unsigned backwardsMatches = countMatches(sequence, begin, range.begin, contig, contigPosition);
                const unsigned forwardsMatches = countMatches(sequence, range.end, end, contig, referenceEnd - forwardsClipped);
                clipBackwards = backwardsMatches < forwardsMatches || (backwardsMatches == forwardsMatches && backwardsClipped < forwardsClipped);
            }
        }
        else if (!backwardsClipped || !forwardsClipped)                                                                // :190-216
        {
            if (clipBackwards && !backwardsClipped)
                doClip = percentMismatches(sequence, begin, range.end, contig, contigPosition) > 40;                   // TOO_GOOD_READ_MISMATCH_PERCENT
            else if (!clipBackwards && !forwardsClipped)
            {
                const long basesClipped = end - range.begin;
                doClip = percentMismatches(sequence, range.begin, end, contig, contigPosition + sequenceLength - basesClipped) > 40;
            }
        }
        if (!doClip) return;
        if (clipBackwards) { f.incrementClipLeft(range.end - begin); begin = range.end; }                             // :240-249
        else { f.incrementClipRight(end - range.begin); end = range.begin; }                                           // :252-261
    }
};

/* -------------------------------------------------------------------------------------------------
 * GappedAligner::makesSenseToGapAlign (--avoid-smith-waterman)  lib/alignment/fragmentBuilder/GappedAligner.cpp:88-165
 * The 7-mer table of the query is cached per strand under (cluster, read) and NOT rebuilt while that key stays, whatever
 * the query range of the later calls is (:95-121).
 * ------------------------------------------------------------------------------------------------- */
struct GapHeuristicPort
{
    static const unsigned KMER = 7, SUFFICIENT_HITS = 8;                                            // GappedAligner.hh:59,75
    static const unsigned short UNINITIALIZED = 0xFFFF, REPEAT = 0xFFFE;                            // :70-71
    unsigned hashedCluster[2], hashedRead[2];
    std::vector<unsigned short> queryKmerOffsets;
    std::vector<unsigned char> trackedOffsets;
    GapHeuristicPort() : queryKmerOffsets(1u << (2 * KMER), UNINITIALIZED) { hashedCluster[0] = hashedCluster[1] = hashedRead[0] = hashedRead[1] = -1U; }

    /// oligo::KmerGenerator (KmerGenerator.hpp:38-124): the k-mers without a non-ACGT base, in order, with their positions
    template <class F> static void forEachKmer(const char *begin, const char *end, F f)
    {
        unsigned kmer = 0, valid = 0;
        for (const char *p = begin; p != end; ++p)
        {
            const unsigned v = AdapterPort::translate(*p);
            if (v > 3) { valid = 0; kmer = 0; continue; }
            kmer = ((kmer << 2) | v) & ((1u << (2 * KMER)) - 1);
            if (++valid >= KMER) f(kmer, long(p - begin) - long(KMER) + 1);
        }
    }

    bool makesSense(unsigned cluster, unsigned read, bool reverse, const char *queryBegin, const char *queryEnd,
                    const char *databaseBegin, const char *databaseEnd)
    {
        if (hashedCluster[reverse] != cluster || hashedRead[reverse] != read)                       // :95-121 (the tile is constant here)
        {
            std::fill(queryKmerOffsets.begin(), queryKmerOffsets.end(), UNINITIALIZED);
            forEachKmer(queryBegin, queryEnd, [&](unsigned kmer, long position) {
                queryKmerOffsets[kmer] = queryKmerOffsets[kmer] == UNINITIALIZED ? (unsigned short)position : REPEAT;
            });
            hashedCluster[reverse] = cluster; hashedRead[reverse] = read;
        }
        const int queryLength = int(queryEnd - queryBegin);
        trackedOffsets.assign(size_t(queryLength) * 2 + size_t(databaseEnd - databaseBegin) + 65536, 0);   // QUERY_LENGTH_MAX counters (:124)
        int lastConfirmedOffset = INT_MAX;
        bool ret = false, done = false;
        forEachKmer(databaseBegin, databaseEnd, [&](unsigned kmer, long databaseOffset) {
            if (done) return;
            const int queryOffset = queryKmerOffsets[kmer];
            if (queryOffset == REPEAT || queryOffset == UNINITIALIZED) return;                      // :136-141
            const int firstBaseOffset = int(databaseOffset) - queryOffset + queryLength;            // :144
            if (firstBaseOffset < 0) return;                                                        // (the reference would index before its array)
            if (++trackedOffsets[firstBaseOffset] == SUFFICIENT_HITS)                               // :147-157
            {
                if (lastConfirmedOffset == INT_MAX) lastConfirmedOffset = firstBaseOffset;
                else if (lastConfirmedOffset != firstBaseOffset) { ret = true; done = true; }
            }
        });
        return ret;
    }
};

/// UngappedAligner::alignUngapped  lib/alignment/fragmentBuilder/UngappedAligner.cpp:39-92
unsigned alignUngapped(const Scores &s, const isaac_ext_reads_t &rm, const Contig &contig, Fragment &f,
                       std::vector<uint32_t> &cigarBuffer, const AdapterClipperPort &clipper)
{
    const unsigned cigarOffset = cigarBuffer.size();
    f.resetAlignment(cigarBuffer);
    f.resetClipping();
    const Read &read = f.read();
    long begin = 0, end = read.length();
    clipper.clip(contig, f, begin, end);                                                            // :59
    clipReadMasking(read, f, begin, end);
    clipReference(contig.length, f, begin, end);
    if (begin) cigarBuffer.push_back(cigarWord(begin, ISAAC_EXT_CIGAR_SOFT_CLIP));
    if (end - begin) cigarBuffer.push_back(cigarWord(end - begin, ISAAC_EXT_CIGAR_ALIGN));
    if (long(read.length()) - end) cigarBuffer.push_back(cigarWord(read.length() - end, ISAAC_EXT_CIGAR_SOFT_CLIP));
    const unsigned ret = updateFragmentCigar(s, rm, contig, f, f.position, cigarBuffer, cigarOffset);
    if (!ret) f.setUnaligned();
    return ret;
}

/// getFlanks  lib/alignment/fragmentBuilder/GappedAligner.cpp:51-82
void getFlanks(long strandPosition, unsigned readLength, uint64_t referenceSize, unsigned &left, unsigned &right)
{
    const unsigned w = BAND;
    if (strandPosition >= w / 2)
    {
        if (strandPosition + readLength + (w - w / 2) < long(referenceSize)) { left = w / 2; right = w - left - 1; }
        else { right = referenceSize - readLength - strandPosition; left = w - right - 1; }
    }
    else { left = strandPosition; right = w - left - 1; }
}

/// GappedAligner::alignGapped  lib/alignment/fragmentBuilder/GappedAligner.cpp:167-249; heuristic = 0: --avoid-smith-waterman off
unsigned alignGapped(const Scores &s, BandedSw &sw, const isaac_ext_reads_t &rm, const Contig &contig, Fragment &f,
                     std::vector<uint32_t> &cigarBuffer, const AdapterClipperPort &clipper, GapHeuristicPort *heuristic)
{
    const unsigned cigarOffset = cigarBuffer.size();
    f.resetAlignment(cigarBuffer);
    f.resetClipping();
    const Read &read = f.read();
    const std::vector<char> &sequence = read.seq[f.reverse];
    long begin = 0, end = read.length();
    clipper.clip(contig, f, begin, end);                                                            // :186
    clipReadMasking(read, f, begin, end);
    clipReference(contig.length, f, begin, end);
    if (begin) cigarBuffer.push_back(cigarWord(begin, ISAAC_EXT_CIGAR_SOFT_CLIP));
    const unsigned sequenceLength = end - begin;
    long strandPosition = f.position;
    if (long(contig.length) < long(sequenceLength) + strandPosition + long(BAND)) return 0;   // :204-208
    unsigned left, right;
    getFlanks(strandPosition, sequenceLength, contig.length, left, right);
    if (heuristic && !heuristic->makesSense(f.cluster->id, f.readIndex, f.reverse, &sequence[0] + begin, &sequence[0] + end,
                                            contig.bases + strandPosition - left,
                                            contig.bases + strandPosition - left + (left + sequenceLength + right)))   // :218-226
        return 0;
    strandPosition += sw.align(&sequence[begin], sequenceLength, contig.bases + strandPosition - left, cigarBuffer);
    if (long(read.length()) - end) cigarBuffer.push_back(cigarWord(read.length() - end, ISAAC_EXT_CIGAR_SOFT_CLIP));
    strandPosition -= left;
    return updateFragmentCigar(s, rm, contig, f, strandPosition, cigarBuffer, cigarOffset);
}

void flatten(const Fragment &f, uint32_t readId, uint32_t cigarOffset, unsigned matchCount, const isaac_ext_reads_t &rm,
             isaac_ext_fragment_t &o, uint32_t *cigarOut, uint64_t *maskOut)
{
    std::memset(&o, 0, sizeof(o));
    o.position = f.position; o.logProbability = f.logProbability; o.contigId = f.contigId; o.readId = readId;
    o.cigarOffset = cigarOffset; o.smithWatermanScore = f.smithWatermanScore; o.observedLength = f.observedLength;
    o.mismatchCount = f.mismatchCount; o.matchesInARow = f.matchesInARow; o.gapCount = f.gapCount;
    o.editDistance = f.editDistance; o.uniqueSeedCount = f.uniqueSeedCount; o.repeatSeedsCount = f.repeatSeedsCount;
    o.nonUniqueSeedOffsetFirst = uint16_t(std::min<uint32_t>(f.nonUniqueFirst, 0xFFFF));
    o.nonUniqueSeedOffsetSecond = uint16_t(f.nonUniqueSecond);
    o.firstSeedIndex = int16_t(f.firstSeedIndex); o.lowClipped = f.lowClipped; o.highClipped = f.highClipped;
    o.cigarLength = f.cigarLength; o.reverse = f.reverse; o.readIndex = f.readIndex; o.matchCount = matchCount;
    if (f.cigarLength && f.cigarBuffer && cigarOut)
        std::copy(f.cigarBuffer->begin() + f.cigarOffset, f.cigarBuffer->begin() + f.cigarOffset + f.cigarLength, cigarOut);
    if (maskOut)
    {
        std::fill(maskOut, maskOut + ISAAC_EXT_MASK_WORDS, uint64_t(0));
        const unsigned firstCycle = rm.firstCycle[f.readIndex], lastCycle = firstCycle + rm.readLength[f.readIndex] - 1;
        for (uint16_t c : f.mismatchCycles)
        {
            const unsigned i = f.reverse ? lastCycle - c : c - firstCycle;
            maskOut[i / 64] |= uint64_t(1) << (i % 64);
        }
    }
}

template <class F> void parallelFor(uint32_t n, uint32_t threads, F f)
{
    if (threads <= 1 || n < 2 * threads) { f(0u, n); return; }
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t)
    {
        const uint32_t b = uint64_t(n) * t / threads, e = uint64_t(n) * (t + 1) / threads;
        pool.emplace_back([=]() { f(b, e); });
    }
    for (std::thread &th : pool) th.join();
}

int extendBatch(bool gapped, const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                uint32_t n, const isaac_ext_candidate_t *candidates, uint32_t cigarStride,
                isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *maskOut, uint32_t threads)
{
    const Scores scores(*cfg);
    const unsigned totalReadLength = reads->readLength[0] + (reads->readCount > 1 ? reads->readLength[1] : 0);
    if (gapped && !BandedSw(cfg->gapMatchScore, cfg->gapMismatchScore, -cfg->gapOpenScore, -cfg->gapExtendScore, totalReadLength).valid)
        return ISAAC_EXT_E_INVALID_ARG;
    bool overflow = false;
    parallelFor(n, threads, [&](uint32_t b, uint32_t e) {
        BandedSw sw(cfg->gapMatchScore, cfg->gapMismatchScore, -cfg->gapOpenScore, -cfg->gapExtendScore, totalReadLength);   // GappedAligner.cpp:41-42
        GapHeuristicPort heuristic;
        Cluster cluster; uint32_t loaded = -1U;
        std::vector<uint32_t> cigar;
        for (uint32_t i = b; i < e; ++i)
        {
            const isaac_ext_candidate_t &c = candidates[i];
            const uint32_t clusterId = c.readId / reads->readCount, readIndex = c.readId % reads->readCount;
            if (loaded != clusterId) { cluster.load(reads, clusterId); loaded = clusterId; }
            const Contig contig = {genome->contigBases[(c.contigStrand >> 1)], genome->contigLengths[(c.contigStrand >> 1)]};
            cigar.clear();
            Fragment f(&cluster, &cigar, readIndex);
            f.reverse = (c.contigStrand & 1); f.contigId = (c.contigStrand >> 1); f.position = c.position;
            AdapterClipperPort clipper;                                            // one clipper per candidate (testSequencingAdapter.cpp:159-182)
            clipper.checkInitStrand(f, contig);
            unsigned matchCount = alignUngapped(scores, *reads, contig, f, cigar, clipper);
            // the reference only gap-aligns fragments whose ungapped alignment kept at least one match
            // (FragmentBuilder.cpp:179 drops the others first, ShadowAligner.cpp:223-226 never lists them)
            if (gapped && matchCount)
            {
                Fragment tmp = f;                                                  // FragmentBuilder.cpp:199-200
                matchCount = alignGapped(scores, sw, *reads, contig, tmp, cigar, clipper, cfg->avoidSmithWaterman ? &heuristic : 0);
                f = tmp;
            }
            if (f.cigarLength > cigarStride) { overflow = true; f.cigarLength = 0; }
            flatten(f, c.readId, i * cigarStride, matchCount, *reads, fragmentsOut[i],
                    cigarOut ? cigarOut + size_t(i) * cigarStride : 0, maskOut ? maskOut + size_t(i) * ISAAC_EXT_MASK_WORDS : 0);
        }
    });
    return overflow ? ISAAC_EXT_E_CAPACITY : ISAAC_EXT_OK;
}

} // namespace

extern "C" const char *oracle_kind(void) { return "port"; }

extern "C" int oracle_set_adapters(uint32_t count, const isaac_ext_adapter_t *adapters)
{
    std::vector<AdapterPort> &list = portAdapters();
    list.clear();
    for (uint32_t a = 0; a < count; ++a) list.push_back(AdapterPort(adapters[a].sequence, adapters[a].reverse != 0, adapters[a].clipLength));
    return ISAAC_EXT_OK;
}

extern "C" int oracle_banded_sw_batch(uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                      const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                      int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                      uint32_t maxReadLength, uint32_t cigarStride, uint32_t *cigarOut,
                                      uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads)
{
    if (!BandedSw(matchScore, mismatchScore, gapOpenScore, gapExtendScore, maxReadLength).valid) return ISAAC_EXT_E_INVALID_ARG;
    for (uint32_t i = 0; i < n; ++i) if (queryLengths[i] > maxReadLength || !queryLengths[i]) return ISAAC_EXT_E_INVALID_ARG;
    parallelFor(n, threads, [&](uint32_t b, uint32_t e) {
        BandedSw sw(matchScore, mismatchScore, gapOpenScore, gapExtendScore, maxReadLength);
        std::vector<uint32_t> cigar;
        for (uint32_t i = b; i < e; ++i)
        {
            cigar.clear();
            offsetOut[i] = sw.align(queries + queryOffsets[i], queryLengths[i], databases + databaseOffsets[i], cigar);
            cigarLengthOut[i] = cigar.size();
            std::copy(cigar.begin(), cigar.begin() + std::min<size_t>(cigar.size(), cigarStride), cigarOut + size_t(i) * cigarStride);
        }
    });
    return ISAAC_EXT_OK;
}

/// BandedSwT<bandWidth>::align over a batch (bandWidth 16, 32 or 64): the model of the widened band, this library only
template <unsigned WIDTH> static int wideBatch(uint32_t n, const char *queries, const uint64_t *queryOffsets, const uint32_t *queryLengths,
                                               const char *databases, const uint64_t *databaseOffsets, int matchScore, int mismatchScore,
                                               int gapOpenScore, int gapExtendScore, uint32_t maxReadLength, uint32_t cigarStride,
                                               uint32_t *cigarOut, uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads)
{
    if (!BandedSwT<WIDTH>(matchScore, mismatchScore, gapOpenScore, gapExtendScore, maxReadLength).valid) return ISAAC_EXT_E_INVALID_ARG;
    for (uint32_t i = 0; i < n; ++i) if (queryLengths[i] > maxReadLength || !queryLengths[i]) return ISAAC_EXT_E_INVALID_ARG;
    parallelFor(n, threads, [&](uint32_t b, uint32_t e) {
        BandedSwT<WIDTH> sw(matchScore, mismatchScore, gapOpenScore, gapExtendScore, maxReadLength);
        std::vector<uint32_t> cigar;
        for (uint32_t i = b; i < e; ++i)
        {
            cigar.clear();
            offsetOut[i] = sw.align(queries + queryOffsets[i], queryLengths[i], databases + databaseOffsets[i], cigar);
            cigarLengthOut[i] = cigar.size();
            std::copy(cigar.begin(), cigar.begin() + std::min<size_t>(cigar.size(), cigarStride), cigarOut + size_t(i) * cigarStride);
        }
    });
    return ISAAC_EXT_OK;
}

extern "C" int oracle_banded_sw_wide_batch(uint32_t bandWidth, uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                           const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                           int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                           uint32_t maxReadLength, uint32_t cigarStride, uint32_t *cigarOut,
                                           uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads)
{
#define ISAAC_WIDE(W) wideBatch<W>(n, queries, queryOffsets, queryLengths, databases, databaseOffsets, matchScore, mismatchScore, gapOpenScore, \
                                   gapExtendScore, maxReadLength, cigarStride, cigarOut, cigarLengthOut, offsetOut, threads)
    switch (bandWidth)
    {
    case 16: return ISAAC_WIDE(16);
    case 32: return ISAAC_WIDE(32);
    case 64: return ISAAC_WIDE(64);
    default: return ISAAC_EXT_E_INVALID_ARG;
    }
#undef ISAAC_WIDE
}

extern "C" int oracle_ungapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                     const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                                     isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut,
                                     uint32_t threads)
{
    return extendBatch(false, genome, reads, config, n, candidates, 3, fragmentsOut, cigarOut, mismatchMaskOut, threads);
}

extern "C" int oracle_gapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                   const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                                   uint32_t cigarStride, isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut,
                                   uint64_t *mismatchMaskOut, uint32_t threads)
{
    return extendBatch(true, genome, reads, config, n, candidates, cigarStride, fragmentsOut, cigarOut, mismatchMaskOut, threads);
}

// SimpleIndelAligner, FragmentBuilder::build, ShadowAligner::rescueShadow
#include "isaac_oracle_build.inc"
