// isaac_b200.hh -- C++ classes carrying the reference's names on top of the C ABI (include/isaac_ext.h).
//
// These are the batch-of-one adapters that let code written against the reference's alignment classes (and the
// reference's own unit tests) run against the GPU path unchanged in spirit: same class names, same argument meaning,
// same error behaviour (the constructors throw isaac_b200::common::InvalidParameterException where the reference throws
// common::InvalidParameterException, BandedSmithWaterman.cpp:49-53).  Production use does NOT go through them: a tile
// is handed over in one isaac_ext_build_fragments / isaac_ext_rescue_shadows call.
//
//   reference class                                   here
//   alignment::BandedSmithWaterman  (.hh:37-105)      isaac_b200::alignment::BandedSmithWaterman
//   alignment::FragmentBuilder      (.hh:46-72)       isaac_b200::alignment::FragmentBuilder
//   alignment::ShadowAligner        (.hh:45-89)       isaac_b200::alignment::ShadowAligner
//   fragmentBuilder::UngappedAligner / GappedAligner  isaac_b200::alignment::fragmentBuilder::{UngappedAligner,GappedAligner}
//   alignment::FragmentMetadata     (.hh:47-444)      isaac_b200::alignment::FragmentMetadata (the fields of isaac_ext_fragment_t)
//   alignment::Cigar                (.hh:41-216)      isaac_b200::alignment::Cigar
//
// Header only; link with libisaac_ext.so.
#ifndef ISAAC_B200_HH
#define ISAAC_B200_HH

#include <cstdint>
#include <stdexcept>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/isaac_ext.h"

namespace isaac_b200
{
namespace common
{
struct InvalidParameterException : public std::logic_error { explicit InvalidParameterException(const std::string &m) : std::logic_error(m) {} };
struct DeviceException : public std::runtime_error { explicit DeviceException(const std::string &m) : std::runtime_error(m) {} };
} // namespace common

namespace reference
{
/// reference::Contig (include/reference/Contig.hh:31-39): 1 byte per base, upper-case ACGTN
struct Contig
{
    unsigned index_; std::string name_; std::vector<char> forward_;
    Contig(unsigned index, const std::string &name) : index_(index), name_(name) {}
    size_t getLength() const { return forward_.size(); }
};
} // namespace reference

namespace flowcell
{
/// flowcell::SequencingAdapterMetadata (SequencingAdapterMetadata.hh:36-72): clipLength 0 = unbounded
struct SequencingAdapterMetadata
{
    std::string sequence_; bool reverse_; unsigned clipLength_;
    SequencingAdapterMetadata(const std::string &sequence, bool reverse) : sequence_(sequence), reverse_(reverse), clipLength_(unsigned(sequence.size())) {}
    SequencingAdapterMetadata(const std::string &sequence, bool reverse, unsigned clipLength) : sequence_(sequence), reverse_(reverse), clipLength_(clipLength) {}
    const std::string &getSequence() const { return sequence_; }
    bool isReverse() const { return reverse_; }
    unsigned getClipLength() const { return clipLength_; }
    bool isUnbounded() const { return !clipLength_; }
};
typedef std::vector<SequencingAdapterMetadata> SequencingAdapterMetadataList;
} // namespace flowcell

/// RAII owner of one isaac_ext_ctx
class Context
{
public:
    explicit Context(const isaac_ext_config_t &config) : ctx_(0), config_(config)
    {
        const int rc = isaac_ext_create(&config, &ctx_);
        if (rc == ISAAC_EXT_E_INVALID_ARG) throw common::InvalidParameterException(isaac_ext_last_error(0));
        if (rc) throw common::DeviceException(isaac_ext_last_error(0));
    }
    ~Context() { isaac_ext_destroy(ctx_); }
    isaac_ext_ctx *get() const { return ctx_; }
    const isaac_ext_config_t &config() const { return config_; }
    void check(int rc) const
    {
        if (rc == ISAAC_EXT_E_INVALID_ARG) throw common::InvalidParameterException(isaac_ext_last_error(ctx_));
        if (rc) throw common::DeviceException(isaac_ext_last_error(ctx_));
    }
    void setReference(const std::vector<reference::Contig> &contigs)
    {
        std::vector<const char *> bases; std::vector<uint64_t> lengths;
        for (size_t i = 0; i < contigs.size(); ++i) { bases.push_back(contigs[i].forward_.data()); lengths.push_back(contigs[i].forward_.size()); }
        check(isaac_ext_set_reference(ctx_, uint32_t(contigs.size()), bases.data(), lengths.data()));
    }
    /// the matchSelector::SequencingAdapterList every later build() / rescueShadow() / alignUngapped() clips with
    void setAdapters(const flowcell::SequencingAdapterMetadataList &adapters)
    {
        std::vector<isaac_ext_adapter_t> flat;
        for (size_t i = 0; i < adapters.size(); ++i)
        {
            const isaac_ext_adapter_t a = {adapters[i].getSequence().c_str(), adapters[i].isReverse() ? 1u : 0u, adapters[i].getClipLength()};
            flat.push_back(a);
        }
        check(isaac_ext_set_adapters(ctx_, uint32_t(flat.size()), flat.empty() ? 0 : flat.data()));
    }
private:
    Context(const Context &); Context &operator=(const Context &);
    isaac_ext_ctx *ctx_;
    isaac_ext_config_t config_;
};

inline isaac_ext_config_t makeConfig(int gapMatchScore, int gapMismatchScore, int gapOpenScore, int gapExtendScore, int minGapExtendScore,
                                     unsigned maxReadLength, unsigned repeatThreshold = 10, unsigned maxSeedsPerRead = 8,
                                     unsigned gappedMismatchesMax = 5, unsigned semialignedGapLimit = 100, bool avoidSmithWaterman = false)
{
    isaac_ext_config_t c = {gapMatchScore, gapMismatchScore, gapOpenScore, gapExtendScore, minGapExtendScore, repeatThreshold,
                            maxSeedsPerRead, gappedMismatchesMax, semialignedGapLimit, avoidSmithWaterman ? 1u : 0u, maxReadLength, 0, 0};
    return c;
}

namespace alignment
{

/// alignment::Cigar (Cigar.hh:41-216): word = length << 4 | op
class Cigar : public std::vector<uint32_t>
{
public:
    enum OpCode { ALIGN = 0, INSERT = 1, DELETE = 2, SKIP = 3, SOFT_CLIP = 4, HARD_CLIP = 5, PAD = 6, MATCH = 7, MISMATCH = 8, UNKNOWN = 9 };
    static uint32_t encode(unsigned length, OpCode op) { return (length << 4) | op; }
    static std::pair<unsigned, OpCode> decode(uint32_t v) { return std::make_pair(v >> 4, OpCode(std::min<unsigned>(v & 0xF, UNKNOWN))); }
    void addOperation(unsigned length, OpCode op) { push_back(encode(length, op)); }
    static std::string toString(const uint32_t *begin, const uint32_t *end)
    {
        static const char ops[] = "MIDNSHP=X?";
        std::string s;
        for (; begin != end; ++begin) { s += std::to_string(*begin >> 4); s += ops[std::min<unsigned>(*begin & 0xF, 9)]; }
        return s;
    }
    std::string toString() const { return toString(data(), data() + size()); }
};

/// the FragmentMetadata fields the template layer reads (FragmentMetadata.hh:330-414) + its CIGAR
struct FragmentMetadata : public isaac_ext_fragment_t
{
    const std::vector<uint32_t> *cigarBuffer;
    FragmentMetadata() : cigarBuffer(0) { isaac_ext_fragment_t z = isaac_ext_fragment_t(); static_cast<isaac_ext_fragment_t &>(*this) = z; }
    FragmentMetadata(const isaac_ext_fragment_t &f, const std::vector<uint32_t> *buffer) : isaac_ext_fragment_t(f), cigarBuffer(buffer) {}
    bool isAligned() const { return 0 != cigarLength; }
    bool isReverse() const { return reverse; }
    unsigned getObservedLength() const { return isAligned() ? observedLength : 0; }
    unsigned getMismatchCount() const { return mismatchCount; }
    unsigned getEditDistance() const { return editDistance; }
    unsigned getGapCount() const { return gapCount; }
    long getPosition() const { return position; }
    unsigned getContigId() const { return contigId; }
    std::string getCigarString() const
    {
        return cigarBuffer && cigarLength ? Cigar::toString(cigarBuffer->data() + cigarOffset, cigarBuffer->data() + cigarOffset + cigarLength) : std::string();
    }
    bool isWellAnchored() const         // FragmentMetadata.hh:477-483, WEAK_SEED_LENGTH = 32
    {
        return uniqueSeedCount || (nonUniqueSeedOffsetSecond > nonUniqueSeedOffsetFirst && unsigned(nonUniqueSeedOffsetSecond - nonUniqueSeedOffsetFirst) >= 32);
    }
};

/// alignment::BandedSmithWaterman (BandedSmithWaterman.hh:37-105): same constructor arguments and overflow check, align()
/// appends to the CIGAR and returns the stripped leading deletion.
class BandedSmithWaterman
{
public:
    static const unsigned WIDEST_GAP_SIZE = ISAAC_EXT_BAND_WIDTH, distanceCutoff = ISAAC_EXT_SW_DISTANCE_CUTOFF, mismatchesCutoff = ISAAC_EXT_SW_MISMATCH_CUTOFF;
    BandedSmithWaterman(int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore, int maxReadLength)
        : match_(matchScore), mismatch_(mismatchScore), open_(gapOpenScore), extend_(gapExtendScore),
          context_(makeConfig(matchScore, mismatchScore, -gapOpenScore, -gapExtendScore, -gapExtendScore, unsigned(maxReadLength))) {}
    unsigned align(const std::vector<char> &query, std::vector<char>::const_iterator databaseBegin, std::vector<char>::const_iterator databaseEnd,
                   Cigar &cigar) const
    {
        const uint64_t zero = 0; const uint32_t length = uint32_t(query.size());
        if (size_t(databaseEnd - databaseBegin) != query.size() + WIDEST_GAP_SIZE - 1)
            throw common::InvalidParameterException("database must be query + 15 bases (BandedSmithWaterman.cpp:93)");
        uint32_t words[64], count = 0, offset = 0;
        context_.check(isaac_ext_banded_sw_batch(context_.get(), 1, query.data(), &zero, &length, &*databaseBegin, &zero,
                                                 match_, mismatch_, open_, extend_, 64, words, &count, &offset));
        cigar.insert(cigar.end(), words, words + count);
        return offset;
    }
private:
    int match_, mismatch_, open_, extend_;
    Context context_;
};

/// One cluster's reads in BCL form plus the read geometry (alignment::Cluster / flowcell::ReadMetadataList)
struct Cluster
{
    std::vector<uint8_t> bcl;            // read 0 then read 1, quality << 2 | base, 0..3 = N (BclClusters.hh:33-124)
    unsigned readLength[2]; unsigned firstCycle[2]; unsigned readCount;
    uint16_t endCyclesMasked[2];
    Cluster() : readCount(0) { readLength[0] = readLength[1] = 0; firstCycle[0] = 1; firstCycle[1] = 1; endCyclesMasked[0] = endCyclesMasked[1] = 0; }
    isaac_ext_reads_t view() const
    {
        isaac_ext_reads_t r = {1, readCount, {readLength[0], readLength[1]}, {firstCycle[0], firstCycle[1]}, bcl.data(), endCyclesMasked};
        return r;
    }
};

typedef isaac_ext_match_t Match;             // alignment::Match (Match.hh:38-73), bit-compatible
typedef isaac_ext_seed_t SeedMetadata;       // alignment::SeedMetadata (SeedMetadata.hh:46-98)
typedef std::vector<SeedMetadata> SeedMetadataList;
typedef isaac_ext_tls_t TemplateLengthStatistics;

/// alignment::FragmentBuilder (FragmentBuilder.hh:46-72): build() for one cluster, results through getFragments() /
/// getCigarBuffer() exactly like the reference (valid until the next build()).
class FragmentBuilder
{
public:
    FragmentBuilder(Context &context) : context_(context), fragments_(2) {}
    bool build(const SeedMetadataList &seedMetadataList, std::vector<Match>::const_iterator matchBegin,
               std::vector<Match>::const_iterator matchEnd, const Cluster &cluster, bool withGaps)
    {
        const isaac_ext_reads_t reads = cluster.view();
        context_.check(isaac_ext_set_reads(context_.get(), &reads));
        const uint64_t begin[2] = {0, uint64_t(matchEnd - matchBegin)};
        const isaac_ext_build_batch_t batch = {begin[1] ? &*matchBegin : 0, begin, seedMetadataList.data(), uint32_t(seedMetadataList.size()), withGaps ? 1u : 0u};
        isaac_ext_build_result_t r;
        context_.check(isaac_ext_build_fragments(context_.get(), &batch, &r));
        cigarBuffer_.assign(r.cigars, r.cigars + r.cigarWords);
        for (unsigned read = 0; read < 2; ++read)
        {
            fragments_[read].clear();
            if (read < cluster.readCount)
                for (uint64_t i = r.readFragmentBegin[read]; i < r.readFragmentBegin[read + 1]; ++i)
                    fragments_[read].push_back(FragmentMetadata(r.fragments[i], &cigarBuffer_));
        }
        return r.built[0] != 0;
    }
    const std::vector<std::vector<FragmentMetadata> > &getFragments() const { return fragments_; }
    const std::vector<uint32_t> &getCigarBuffer() const { return cigarBuffer_; }
private:
    Context &context_;
    std::vector<std::vector<FragmentMetadata> > fragments_;
    std::vector<uint32_t> cigarBuffer_;
};

/// alignment::ShadowAligner (ShadowAligner.hh:45-89): rescueShadow() for one orphan of the cluster last given to
/// FragmentBuilder::build (or to setCluster).
class ShadowAligner
{
public:
    ShadowAligner(Context &context) : context_(context) {}
    void setCluster(const Cluster &cluster) { const isaac_ext_reads_t reads = cluster.view(); context_.check(isaac_ext_set_reads(context_.get(), &reads)); }
    bool rescueShadow(const FragmentMetadata &orphan, std::vector<FragmentMetadata> &shadowList,
                      const TemplateLengthStatistics &templateLengthStatistics, long bestTemplateLength)
    {
        const isaac_ext_rescue_request_t q = {orphan.position, bestTemplateLength, orphan.readIndex, (orphan.contigId << 1) | (orphan.reverse ? 1u : 0u),
                                              orphan.observedLength, 0};
        isaac_ext_rescue_result_t r;
        context_.check(isaac_ext_rescue_shadows(context_.get(), &templateLengthStatistics, 1, &q, &r));
        shadowCigarBuffer_.assign(r.cigars, r.cigars + r.cigarWords);
        shadowList.clear();
        for (uint64_t i = 0; i < r.fragmentCount; ++i) shadowList.push_back(FragmentMetadata(r.fragments[i], &shadowCigarBuffer_));
        return r.rescued[0] != 0;
    }
    const std::vector<uint32_t> &getCigarBuffer() const { return shadowCigarBuffer_; }
private:
    Context &context_;
    std::vector<uint32_t> shadowCigarBuffer_;
};

/// alignment::BamTemplate (BamTemplate.hh:40-137): the two fragments chosen for a cluster, the template mapping score
/// and the proper-pair flag.  FragmentMetadata::alignmentScore travels next to the flat record.
class BamTemplate
{
public:
    BamTemplate() : fragmentCount_(0), alignmentScore_(0), properPair_(false) { fragmentAlignmentScore_[0] = fragmentAlignmentScore_[1] = -1U; }
    unsigned getFragmentCount() const { return fragmentCount_; }
    const FragmentMetadata &getFragmentMetadata(unsigned i) const { return fragments_[i]; }
    unsigned getFragmentAlignmentScore(unsigned i) const { return fragmentAlignmentScore_[i]; }
    unsigned getAlignmentScore() const { return alignmentScore_; }
    bool hasAlignmentScore() const { return -1U != alignmentScore_; }
    bool isProperPair() const { return properPair_; }
    bool isUnanchored() const { return 0 == fragmentAlignmentScore_[0] && (fragmentCount_ < 2 || 0 == fragmentAlignmentScore_[1]); }
private:
    friend class TemplateBuilder;
    FragmentMetadata fragments_[2];
    unsigned fragmentAlignmentScore_[2];
    unsigned fragmentCount_, alignmentScore_;
    bool properPair_;
};

/// alignment::TemplateBuilder (TemplateBuilder.hh:56-137) for one cluster: buildFragments() then buildTemplate(), the result
/// through getBamTemplate() like the reference.  Pair selection, shadow rescue and mapping scores happen behind
/// isaac_ext_build_templates.
class TemplateBuilder
{
public:
    typedef short DodgyAlignmentScore;
    static const DodgyAlignmentScore DODGY_ALIGNMENT_SCORE_UNKNOWN = 255;
    static const DodgyAlignmentScore DODGY_ALIGNMENT_SCORE_UNALIGNED = -1;
    TemplateBuilder(Context &context, bool scatterRepeats, DodgyAlignmentScore dodgyAlignmentScore)
        : context_(context), built_(false)
    {
        options_.scatterRepeats = scatterRepeats ? 1u : 0u; options_.dodgyAlignmentScore = dodgyAlignmentScore;
        options_.mapqThreshold = 0; options_.clipFlags = 0;
    }
    /// keeps the cluster and its matches for buildTemplate (the reference builds the candidate fragments here)
    bool buildFragments(const SeedMetadataList &seedMetadataList, std::vector<Match>::const_iterator matchBegin,
                        std::vector<Match>::const_iterator matchEnd, const Cluster &cluster, bool withGaps)
    {
        seeds_ = seedMetadataList; matches_.assign(matchBegin, matchEnd); withGaps_ = withGaps;
        const isaac_ext_reads_t reads = cluster.view();
        context_.check(isaac_ext_set_reads(context_.get(), &reads));
        readCount_ = cluster.readCount;
        built_ = false;
        return !matches_.empty();
    }
    /// \return false when the template ended up without a single aligned read
    bool buildTemplate(const TemplateLengthStatistics &templateLengthStatistics, unsigned mapqThreshold)
    {
        options_.mapqThreshold = mapqThreshold;
        const uint64_t begin[2] = {0, uint64_t(matches_.size())};
        const isaac_ext_build_batch_t batch = {matches_.empty() ? 0 : matches_.data(), begin, seeds_.data(), uint32_t(seeds_.size()), withGaps_ ? 1u : 0u};
        isaac_ext_template_result_t r;
        context_.check(isaac_ext_build_templates(context_.get(), &batch, &templateLengthStatistics, &options_, &r));
        cigarBuffer_.assign(r.cigars, r.cigars + r.cigarWords);
        bamTemplate_.fragmentCount_ = readCount_;
        for (unsigned i = 0; i < readCount_; ++i)
        {
            bamTemplate_.fragments_[i] = FragmentMetadata(r.fragments[i], &cigarBuffer_);
            bamTemplate_.fragmentAlignmentScore_[i] = r.templates[0].fragmentAlignmentScore[i];
        }
        bamTemplate_.alignmentScore_ = r.templates[0].alignmentScore;
        bamTemplate_.properPair_ = r.templates[0].properPair != 0;
        built_ = r.templates[0].built != 0;
        return built_;
    }
    const BamTemplate &getBamTemplate() const { return bamTemplate_; }
private:
    Context &context_;
    isaac_ext_template_options_t options_;
    SeedMetadataList seeds_;
    std::vector<Match> matches_;
    bool withGaps_, built_;
    unsigned readCount_;
    BamTemplate bamTemplate_;
    std::vector<uint32_t> cigarBuffer_;
};

namespace fragmentBuilder
{
/// fragmentBuilder::UngappedAligner / GappedAligner (UngappedAligner.hh:55-60, GappedAligner.hh:49-54): re-align one
/// fragment of the cluster resident in the context; the fragment is updated in place, the CIGAR appended to cigarBuffer,
/// the match count returned.
class UngappedAligner
{
public:
    UngappedAligner(Context &context) : context_(context) {}
    unsigned alignUngapped(FragmentMetadata &fragment, Cigar &cigarBuffer) const { return align(fragment, cigarBuffer, false); }
protected:
    unsigned align(FragmentMetadata &fragment, Cigar &cigarBuffer, bool gapped) const
    {
        long unclipped = fragment.position;                                   // resetAlignment() (FragmentMetadata.hh:297-313)
        if (fragment.cigarBuffer && fragment.cigarLength && ((*fragment.cigarBuffer)[fragment.cigarOffset] & 0xF) == Cigar::SOFT_CLIP)
            unclipped -= (*fragment.cigarBuffer)[fragment.cigarOffset] >> 4;
        const isaac_ext_candidate_t c = {unclipped, fragment.readId, (fragment.contigId << 1) | (fragment.reverse ? 1u : 0u)};
        isaac_ext_fragment_t out; uint32_t words[64];
        context_.check(gapped ? isaac_ext_gapped_batch(context_.get(), 1, &c, 64, &out, words, 0)
                              : isaac_ext_ungapped_batch(context_.get(), 1, &c, &out, words, 0));
        const unsigned matchCount = out.matchCount;
        if (gapped && !out.cigarLength) return 0;                             // alignGapped returned 0: fragment untouched
        out.uniqueSeedCount = fragment.uniqueSeedCount; out.repeatSeedsCount = fragment.repeatSeedsCount;
        out.nonUniqueSeedOffsetFirst = fragment.nonUniqueSeedOffsetFirst; out.nonUniqueSeedOffsetSecond = fragment.nonUniqueSeedOffsetSecond;
        out.firstSeedIndex = fragment.firstSeedIndex;
        out.cigarOffset = uint32_t(cigarBuffer.size());
        cigarBuffer.insert(cigarBuffer.end(), words, words + out.cigarLength);
        static_cast<isaac_ext_fragment_t &>(fragment) = out;
        fragment.cigarBuffer = &cigarBuffer;
        return matchCount;
    }
    Context &context_;
};

class GappedAligner : public UngappedAligner
{
public:
    GappedAligner(Context &context) : UngappedAligner(context) {}
    unsigned alignGapped(FragmentMetadata &fragment, Cigar &cigarBuffer) const { return align(fragment, cigarBuffer, true); }
};
} // namespace fragmentBuilder

} // namespace alignment

namespace io
{
/// io::FragmentHeader as the reference lays it out on x86-64 (include/io/Fragment.hh:73-404): the record of a bin, followed by
/// readLength_ BCL bytes and cigarLength_ CIGAR words
struct FragmentHeader
{
    int32_t bamTlen_; uint32_t observedLength_; uint64_t fStrandPosition_; uint16_t lowClipped_, highClipped_, alignmentScore_,
        templateAlignmentScore_; uint64_t mateFStrandPosition_; uint16_t readLength_, cigarLength_, gapCount_, editDistance_, flags_, pad0_[3];
    uint64_t tile_, barcode_, barcodeSequence_, clusterId_; int32_t clusterX_, clusterY_; uint64_t duplicateClusterRank_, mateAnchor_;
    uint32_t mateStorageBin_, pad1_;
    enum { PAIRED = 1, UNMAPPED = 2, MATE_UNMAPPED = 4, REVERSE = 8, MATE_REVERSE = 16, FIRST_READ = 32, SECOND_READ = 64, FAIL_FILTER = 128, PROPER_PAIR = 256 };
    unsigned getTotalLength() const { return unsigned(sizeof(FragmentHeader)) + readLength_ + 4u * cigarLength_; }
};
/// reference::ReferencePosition(contigId, position).getValue() (include/reference/ReferencePosition.hh:68-78)
inline uint64_t referencePosition(uint64_t contigId, uint64_t position) { return (((contigId + 1) << 40) | position) << 1; }
} // namespace io

namespace build
{
/// build::PackedFragmentBuffer::Index (include/build/PackedFragmentBuffer.hh:36-91)
struct Index
{
    uint64_t pos_;                       // ReferencePosition::getValue
    unsigned long dataOffset_, mateDataOffset_;
    const uint32_t *cigarBegin_, *cigarEnd_;
};

/// build::GapRealigner driven the way BinSorter drives it (lib/build/BinSorter.cpp:389-418): collectGaps + realignGaps of one bin
/// in one call on the context's GPU
class GapRealigner
{
public:
    GapRealigner(Context &context, bool realignGapsVigorously, bool realignDodgyFragments, unsigned /*realignedGapsPerFragment*/,
                 unsigned mismatchCost, unsigned gapOpenCost, unsigned gapExtendCost, bool clipSemialigned,
                 const std::vector<isaac_ext_tls_t> &barcodeTemplateLengthStatistics)
        : context_(context), tls_(barcodeTemplateLengthStatistics)
    {
        std::memset(&options_, 0, sizeof(options_));
        options_.realignGapsVigorously = realignGapsVigorously; options_.realignDodgyFragments = realignDodgyFragments;
        options_.mismatchCost = mismatchCost; options_.gapOpenCost = gapOpenCost; options_.gapExtendCost = gapExtendCost;
        options_.clipSemialigned = clipSemialigned;
    }
    /// data: the bin's records back to back (updated in place); index: its entries in processing order, their pos_ and CIGAR
    /// pointers are refreshed (a realigned entry points into this object's buffer, valid until the next call)
    void realignBin(uint64_t binStart, uint64_t binEnd, std::vector<char> &data, std::vector<Index> &index)
    {
        options_.binStart = binStart; options_.binEnd = binEnd;
        options_.barcodeCount = uint32_t(tls_.size()); options_.barcodeTls = tls_.data(); options_.barcodeGapGroup = 0;
        std::vector<isaac_ext_bin_index_t> flat;
        for (size_t i = 0; i < index.size(); ++i) { const isaac_ext_bin_index_t e = {index[i].dataOffset_, index[i].mateDataOffset_}; flat.push_back(e); }
        isaac_ext_realign_result_t r;
        context_.check(isaac_ext_realign_bin(context_.get(), &options_, reinterpret_cast<uint8_t *>(data.data()), data.size(), 0, 0,
                                             flat.data(), flat.size(), &r));
        realignedCigars_.assign(r.realignedCigars, r.realignedCigars + r.realignedCigarWords);
        for (size_t i = 0; i < index.size(); ++i)
        {
            index[i].pos_ = r.position[i];
            if (ISAAC_EXT_REALIGN_OWN_CIGAR != r.cigarOffset[i])
            {
                index[i].cigarBegin_ = realignedCigars_.data() + r.cigarOffset[i];
                index[i].cigarEnd_ = index[i].cigarBegin_ + r.cigarLength[i];
            }
        }
    }
private:
    Context &context_;
    std::vector<isaac_ext_tls_t> tls_;
    isaac_ext_realign_options_t options_;
    std::vector<uint32_t> realignedCigars_;
};
} // namespace build
} // namespace isaac_b200

#endif // ISAAC_B200_HH
