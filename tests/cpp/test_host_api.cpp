// The reference's own unit tests, restated against the classes of isaac_aligner_b200/host/isaac_b200.hh (GPU behind them).
// Mirrors lib/alignment/cppunit/testBandedSmithWaterman.cpp (testUngapped :79-103, testSingleDeletion :105-131,
// testSingleInsertion :133-154, testMultipleIndels :156-212, testOverflow :214-225) and the two-seed deletion case of
// testSimpleIndelAligner.cpp:264-277 through FragmentBuilder::build.  Exits 0 when every check passes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../isaac_aligner_b200/host/isaac_b200.hh"

using namespace isaac_b200;
using isaac_b200::alignment::Cigar;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)
#define CHECK_EQ(a, b) do { if (!((a) == (b))) { std::printf("FAILED %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); ++failures; } } while (0)

static std::vector<char> v(const std::string &s) { return std::vector<char>(s.begin(), s.end()); }

static std::string getGenome(unsigned size = 1000)
{
    static const std::string bases = "ACGT";
    std::string genome;
    unsigned state = 12345;
    while (size--) { state = state * 1103515245u + 12345u; genome.push_back(bases[(state >> 16) % 4]); }
    return genome;
}

static std::string align(const alignment::BandedSmithWaterman &bsw, const std::string &query, const std::string &database, unsigned *ret = 0)
{
    Cigar cigar;
    const std::vector<char> q = v(query), d = v(database);
    const unsigned r = bsw.align(q, d.begin(), d.end(), cigar);
    if (ret) *ret = r;
    return cigar.toString();
}

static void testBandedSmithWaterman()
{
    const alignment::BandedSmithWaterman bsw(2, -1, 15, 3, 300);
    const std::string genome = getGenome();
    {   // testUngapped
        const std::string database = genome.substr(100, 115);
        for (unsigned i = 0; i <= 15; ++i) CHECK_EQ(align(bsw, database.substr(i, 100), database), "100M");
    }
    {   // testSingleDeletion
        const unsigned left = 40, right = 40;
        const std::string deletion = "AGAGCAGCGAGCGACAGCAGCAGCAAA";
        for (unsigned deletionLength = 1; 13 >= deletionLength; ++deletionLength)
        {
            const unsigned dl = 7 - (deletionLength / 2);
            const std::string leftS = genome.substr(100 + dl, left - 1) + "T", rightS = genome.substr(100 + dl + left, right);
            const std::string databaseD = genome.substr(100, dl) + leftS + deletion.substr(0, deletionLength) + rightS +
                                          genome.substr(100 + dl + left + right, 15 - dl - deletionLength);
            CHECK_EQ(align(bsw, leftS + rightS, databaseD), "40M" + std::to_string(deletionLength) + "D40M");
        }
    }
    {   // testSingleInsertion
        const std::string database = genome.substr(100, 220);
        const unsigned queryLength = unsigned(database.size()) - 15;
        for (unsigned insertLength = 1; 9 >= insertLength; ++insertLength)
        {
            const unsigned left = 100, right = queryLength - left - insertLength, dl = 9;
            const std::string query = database.substr(dl, left) + std::string(insertLength, 'T') + database.substr(left + dl, right);
            CHECK_EQ(align(bsw, query, database), "100M" + std::to_string(insertLength) + "I" + std::to_string(right) + "M");
        }
    }
    {   // testMultipleIndels
        const unsigned dl = 6;
        const std::string dlS = genome.substr(100, dl), leftS = genome.substr(100 + dl, 19) + "T", centerS = genome.substr(100 + dl + 20, 19) + "T";
        const std::string rightS = genome.substr(100 + dl + 40, 20);
        auto tail = [&](unsigned n) { return genome.substr(100 + dl + 60, n); };
        CHECK_EQ(align(bsw, leftS + "A" + centerS + rightS, dlS + leftS + centerS + "ACAG" + rightS + tail(15 - dl + 1 - 4)), "20M1I20M4D20M");
        CHECK_EQ(align(bsw, leftS + "A" + centerS + "CG" + rightS, dlS + leftS + centerS + rightS + tail(15 - dl + 1 + 2)), "20M1I20M2I20M");
        CHECK_EQ(align(bsw, leftS + centerS + rightS, dlS + leftS + "AAG" + centerS + "ACAG" + rightS + tail(15 - dl - 3 - 4)), "20M3D20M4D20M");
    }
}

template <class F> static bool throwsInvalidParameter(F f)
{
    try { f(); } catch (const common::InvalidParameterException &) { return true; }
    return false;
}

static void testOverflow()
{
    CHECK(!throwsInvalidParameter([] { alignment::BandedSmithWaterman(2, -1, 6, 3, 5460); }));
    CHECK(throwsInvalidParameter([] { alignment::BandedSmithWaterman(2, -1, 7, 3, 4681); }));
    CHECK(throwsInvalidParameter([] { alignment::BandedSmithWaterman(2, -1, 17, 3, 3681); }));
    CHECK(throwsInvalidParameter([] { alignment::BandedSmithWaterman(2, -1, 11, 3, 13681); }));
}

static void testSimpleDeletionThroughFragmentBuilder()
{
    // testSimpleIndelAligner.cpp:264-277: a 14-base deletion between two 32-mer seeds -> 71M14D68M, 0 mismatches, edit distance 14
    const std::string read = "ATTTGGTTAAGGTAGCGGTAAAAGCGTGTTACCGCAATGTTCTGTCTCTTATACAACATCTAGATGTGTAT"
                             "AAGAGACAGGTGCACCGCCTATACACATCTAGAATAAGAGACAGGTGCACCGCCTATACACATCTAGA";
    const std::string ref = "ATTTGGTTAAGGTAGCGGTAAAAGCGTGTTACCGCAATGTTCTGTCTCTTATACAACATCTAGATGTGTATAAAAAAAAAAAAAAAAGAGACAGGTGCACCGCCTATACACATCTAGAATAAGAGACAGGTGCACCGCCTATACACATCTAGA";
    isaac_ext_config_t cfg = makeConfig(0, -1, -2, -1, -5, unsigned(read.size()), 10, 8, 5, 20000);
    Context context(cfg);
    std::vector<reference::Contig> contigs(1, reference::Contig(0, "vasja"));
    contigs[0].forward_ = v(ref);
    context.setReference(contigs);
    alignment::Cluster cluster;
    cluster.readCount = 1; cluster.readLength[0] = unsigned(read.size());
    for (char c : read) cluster.bcl.push_back(uint8_t((35 << 2) | std::string("ACGT").find(c)));
    const unsigned L = unsigned(read.size());
    alignment::SeedMetadataList seeds(2);
    seeds[0].offset = 0; seeds[0].length = 32; seeds[0].readIndex = 0;
    seeds[1].offset = uint16_t(L - 32 - 1); seeds[1].length = 32; seeds[1].readIndex = 0;
    const uint64_t headLocation = 0, tailLocation = (ref.size() - L) + seeds[1].offset;
    std::vector<alignment::Match> matches(2);
    matches[0].seedId = 0u << 1; matches[0].location = (((uint64_t(1) << 40) | headLocation) << 1);
    matches[1].seedId = 1u << 1; matches[1].location = (((uint64_t(1) << 40) | tailLocation) << 1);
    alignment::FragmentBuilder builder(context);
    CHECK(builder.build(seeds, matches.begin(), matches.end(), cluster, false));
    const std::vector<alignment::FragmentMetadata> &list = builder.getFragments()[0];
    CHECK(!list.empty());
    if (!list.empty())
    {
        CHECK_EQ(list[0].getCigarString(), "71M14D68M");
        CHECK_EQ(list[0].getMismatchCount(), 0u);
        CHECK_EQ(list[0].getEditDistance(), 14u);
    }
}

// testTemplateBuilder.cpp: testEmptyMatchList (:149-177, an empty match list leaves two unaligned fragments and template
// score 0) and the shape of testUnique (:230-281): a forward read 1 and a reverse read 2 a nominal template length apart,
// seeded without neighbours, come out as a proper pair with positive mapping scores.
static void testTemplateBuilder()
{
    const std::string genome = getGenome(4000);
    isaac_ext_config_t cfg = makeConfig(0, -3, -11, -4, -20, 200);          // bwa scores
    Context context(cfg);
    std::vector<reference::Contig> contigs(1, reference::Contig(0, "chr"));
    contigs[0].forward_ = v(genome);
    context.setReference(contigs);
    const unsigned L = 100, r1 = 1000, r2 = 1000 + 300 - L;          // template length 300
    const std::string read1 = genome.substr(r1, L);
    std::string read2;                                                // reverse complement of the mate
    for (unsigned i = 0; i < L; ++i)
    {
        const char c = genome[r2 + L - 1 - i];
        read2.push_back(c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A');
    }
    alignment::Cluster cluster;
    cluster.readCount = 2; cluster.readLength[0] = cluster.readLength[1] = L; cluster.firstCycle[0] = 1; cluster.firstCycle[1] = 1 + L;
    for (char c : read1 + read2) cluster.bcl.push_back(uint8_t((35 << 2) | std::string("ACGT").find(c)));
    alignment::SeedMetadataList seeds(2);
    seeds[0].offset = 0; seeds[0].length = 32; seeds[0].readIndex = 0;
    seeds[1].offset = 0; seeds[1].length = 32; seeds[1].readIndex = 1;
    const alignment::TemplateLengthStatistics tls = {200, 400, 300, 30, 30, {1, 6}, -1};      // FRp / RFm
    alignment::TemplateBuilder builder(context, false, alignment::TemplateBuilder::DODGY_ALIGNMENT_SCORE_UNALIGNED);
    {
        std::vector<alignment::Match> none;
        CHECK(!builder.buildFragments(seeds, none.begin(), none.end(), cluster, true));
        CHECK(!builder.buildTemplate(tls, 0));
        const alignment::BamTemplate &t = builder.getBamTemplate();
        CHECK_EQ(t.getFragmentCount(), 2u);
        CHECK_EQ(t.getAlignmentScore(), 0u);
        CHECK(!t.getFragmentMetadata(0).isAligned() && !t.getFragmentMetadata(1).isAligned());
        CHECK(!t.isProperPair());
    }
    {
        std::vector<alignment::Match> matches(2);
        // read 1 forward: the seed sits at the read start; read 2 reverse: its first 32 bases are the LAST 32 of the window
        matches[0].seedId = (0u << 1) | 0u; matches[0].location = (((uint64_t(1) << 40) | uint64_t(r1)) << 1);
        matches[1].seedId = (1u << 1) | 1u; matches[1].location = (((uint64_t(1) << 40) | uint64_t(r2 + L - 32)) << 1);
        CHECK(builder.buildFragments(seeds, matches.begin(), matches.end(), cluster, true));
        CHECK(builder.buildTemplate(tls, 0));
        const alignment::BamTemplate &t = builder.getBamTemplate();
        CHECK(t.isProperPair());
        CHECK(t.hasAlignmentScore() && t.getAlignmentScore() > 100);
        CHECK_EQ(t.getFragmentMetadata(0).getPosition(), long(r1));
        CHECK_EQ(t.getFragmentMetadata(1).getPosition(), long(r2));
        CHECK(!t.getFragmentMetadata(0).isReverse() && t.getFragmentMetadata(1).isReverse());
        CHECK_EQ(t.getFragmentMetadata(0).getCigarString(), "100M");
        CHECK_EQ(t.getFragmentMetadata(1).getCigarString(), "100M");
        CHECK(t.getFragmentAlignmentScore(0) > 100 && t.getFragmentAlignmentScore(1) > 100);
    }
    {   // only read 1 has a seed match: the mate must be rescued into a proper pair (testOrphan :179-228)
        std::vector<alignment::Match> matches(1);
        matches[0].seedId = (0u << 1) | 0u; matches[0].location = (((uint64_t(1) << 40) | uint64_t(r1)) << 1);
        CHECK(builder.buildFragments(seeds, matches.begin(), matches.end(), cluster, true));
        CHECK(builder.buildTemplate(tls, 0));
        const alignment::BamTemplate &t = builder.getBamTemplate();
        CHECK(t.isProperPair());
        CHECK_EQ(t.getFragmentMetadata(1).getPosition(), long(r2));
        CHECK_EQ(t.getFragmentMetadata(1).getCigarString(), "100M");
    }
}

// testSequencingAdapter.cpp: testMp51M49S (:184-204, a mate-pair junction adapter in the middle of the read: the longer, matching
// side is kept) and testStd38M62S (:376-401, an unbounded standard adapter: everything from the adapter on is clipped), aligned
// through FragmentBuilder::build with one seed match at position 0 like the reference's harness aligns at position 0.
static void testSequencingAdapter()
{
    struct Case { const char *read, *reference; bool matePair; const char *cigar; unsigned observedLength; };
    const Case cases[2] = {
        {"CGATTGTCTTTGCTGCCAATTTTAGCGTTGGCGTTAACGTCATGCTTAAGCCTGTCTCTTATACACATCTAGATGTGTATAAGAGACAGCTGCTACGCCA",
         "CGATTGTCTTTGCTGCCAATTTTAGCGTTGGCGTTAACGTCATGCTTAAGCCAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", true, "51M49S", 51},
        {"TGGTTAAGGTAGCGGTAAAAGCGTGTTACCGCAATGTTCTGTCTCTTATACACATCTAGATGTGTATAAGAGACAGGTGCACCGCCTATACACATCTAGA",
         "TGGTTAAGGTAGCGGTAAAAGCGTGTTACCGCAATGTTCTCTCTTCTCTGGAATATGATAAAAAAAAAAAAAAAAAGTGCACCGCCAAAAAAAAAAAAAA", false, "38M62S", 38}};
    for (const Case &c : cases)
    {
        const std::string read = c.read, ref = c.reference;
        isaac_ext_config_t cfg = makeConfig(2, -1, -15, -3, -25, unsigned(read.size()), 10, 8, 5, 0);
        Context context(cfg);
        std::vector<reference::Contig> contigs(1, reference::Contig(0, "vasja"));
        contigs[0].forward_ = v(ref);
        context.setReference(contigs);
        flowcell::SequencingAdapterMetadataList adapters;
        adapters.push_back(flowcell::SequencingAdapterMetadata("CTGTCTCTTATACACATCT", false, c.matePair ? 19u : 0u));
        adapters.push_back(flowcell::SequencingAdapterMetadata("AGATGTGTATAAGAGACAG", true, c.matePair ? 19u : 0u));
        context.setAdapters(adapters);
        alignment::Cluster cluster;
        cluster.readCount = 1; cluster.readLength[0] = unsigned(read.size());
        for (char b : read) cluster.bcl.push_back(uint8_t((35 << 2) | std::string("ACGT").find(b)));
        alignment::SeedMetadataList seeds(1);
        seeds[0].offset = 0; seeds[0].length = 32; seeds[0].readIndex = 0;
        std::vector<alignment::Match> matches(1);
        matches[0].seedId = 0u << 1; matches[0].location = ((uint64_t(1) << 40) << 1);
        alignment::FragmentBuilder builder(context);
        CHECK(builder.build(seeds, matches.begin(), matches.end(), cluster, false));
        const std::vector<alignment::FragmentMetadata> &list = builder.getFragments()[0];
        CHECK_EQ(list.size(), size_t(1));
        if (!list.empty())
        {
            CHECK_EQ(list[0].getCigarString(), c.cigar);
            CHECK_EQ(list[0].getMismatchCount(), 0u);
            CHECK_EQ(list[0].getObservedLength(), c.observedLength);
            CHECK_EQ(list[0].getPosition(), 0L);
        }
        // a 2-base sequence is refused like the SequencingAdapter constructor's assertions refuse what it cannot hash
        bool thrown = false;
        try { context.setAdapters(flowcell::SequencingAdapterMetadataList(1, flowcell::SequencingAdapterMetadata("AC", false))); }
        catch (const common::InvalidParameterException &) { thrown = true; }
        CHECK(thrown);
    }
}

/// One record of a bin: io::FragmentHeader + BCL bytes (quality 30) + CIGAR words
static void appendRecord(std::vector<char> &data, const std::string &bases, uint64_t position, const std::vector<uint32_t> &cigar,
                         unsigned editDistance, unsigned gapCount, unsigned observedLength)
{
    io::FragmentHeader h;
    std::memset(&h, 0, sizeof(h));
    h.observedLength_ = observedLength; h.fStrandPosition_ = io::referencePosition(0, position);
    h.alignmentScore_ = 100; h.templateAlignmentScore_ = 100; h.mateFStrandPosition_ = io::referencePosition(0, position);
    h.readLength_ = uint16_t(bases.size()); h.cigarLength_ = uint16_t(cigar.size()); h.gapCount_ = uint16_t(gapCount);
    h.editDistance_ = uint16_t(editDistance);
    h.flags_ = io::FragmentHeader::MATE_UNMAPPED | io::FragmentHeader::FIRST_READ | io::FragmentHeader::SECOND_READ;     // single-ended
    h.tile_ = 1; h.clusterX_ = h.clusterY_ = 0x7FFFFFFF;
    const char *raw = reinterpret_cast<const char *>(&h);
    data.insert(data.end(), raw, raw + sizeof(h));
    for (size_t i = 0; i < bases.size(); ++i) data.push_back(char((30 << 2) | std::string("ACGT").find(bases[i])));
    const char *words = reinterpret_cast<const char *>(cigar.data());
    data.insert(data.end(), words, words + 4 * cigar.size());
}

/// build::GapRealigner: a read that crosses a deletion 30 bases from its start and was laid down without a gap is repaired with the
/// deletion another read of the bin brought along (the situation the class exists for, lib/build/GapRealigner.cpp:1061-1267)
static void testGapRealigner()
{
    const std::string genome = getGenome(600);
    const std::string haplotype = genome.substr(0, 150) + genome.substr(155);             // five bases of the reference deleted
    Context context(makeConfig(2, -1, -15, -3, -25, 300));
    std::vector<reference::Contig> contigs(1, reference::Contig(0, "c0"));
    contigs[0].forward_ = v(genome);
    context.setReference(contigs);
    const uint32_t M = ISAAC_EXT_CIGAR_ALIGN, D = ISAAC_EXT_CIGAR_DELETE;
    std::vector<char> data;
    std::vector<build::Index> index;
    // read A: haplotype [100, 200) with its true alignment 50M5D50M at 100
    const std::vector<uint32_t> cigarA = {(50u << 4) | M, (5u << 4) | D, (50u << 4) | M};
    appendRecord(data, haplotype.substr(100, 100), 100, cigarA, 5, 1, 105);
    // read B: haplotype [120, 220) laid down as 100M at 120: everything behind the deletion is shifted
    const std::string readB = haplotype.substr(120, 100);
    unsigned mismatches = 0;
    for (unsigned i = 0; i < 100; ++i) mismatches += readB[i] != genome[120 + i];
    const std::vector<uint32_t> cigarB = {(100u << 4) | M};
    const size_t offsetB = data.size();
    appendRecord(data, readB, 120, cigarB, mismatches, 0, 100);
    for (size_t offset = 0, k = 0; offset < data.size(); ++k)
    {
        const io::FragmentHeader *h = reinterpret_cast<const io::FragmentHeader *>(&data[offset]);
        const uint32_t *cigar = reinterpret_cast<const uint32_t *>(&data[offset] + sizeof(io::FragmentHeader) + h->readLength_);
        const build::Index e = {h->fStrandPosition_, offset, offset, cigar, cigar + h->cigarLength_};
        index.push_back(e);
        offset += h->getTotalLength();
    }
    CHECK(mismatches > 20);
    const std::vector<isaac_ext_tls_t> tls(1, isaac_ext_tls_t{245, 455, 350, 35, 35, {1, 6}, -1});
    build::GapRealigner realigner(context, false, false, 1, 3, 4, 0, false, tls);
    realigner.realignBin(io::referencePosition(0, 0), io::referencePosition(0, genome.size()), data, index);
    CHECK_EQ(index[0].cigarEnd_ - index[0].cigarBegin_, 3);                               // read A is left alone
    CHECK_EQ(index[1].pos_, io::referencePosition(0, 120));
    CHECK_EQ(index[1].cigarEnd_ - index[1].cigarBegin_, 3);
    if (index[1].cigarEnd_ - index[1].cigarBegin_ == 3)
    {
        CHECK_EQ(index[1].cigarBegin_[0], (30u << 4) | M);
        CHECK_EQ(index[1].cigarBegin_[1], (5u << 4) | D);
        CHECK_EQ(index[1].cigarBegin_[2], (70u << 4) | M);
    }
    const io::FragmentHeader *b = reinterpret_cast<const io::FragmentHeader *>(&data[offsetB]);
    CHECK_EQ(b->editDistance_, 5);
    CHECK_EQ(b->observedLength_, 105u);
}

int main()
{
    testSequencingAdapter();
    testOverflow();
    testBandedSmithWaterman();
    testSimpleDeletionThroughFragmentBuilder();
    testTemplateBuilder();
    testGapRealigner();
    std::printf(failures ? "%d checks FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
