/*
 * ref_capi_realign.cpp -- oracle_realign_bin (oracle_api.h) over the reference's OWN build::RealignerGaps / build::GapRealigner.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  No realignment logic here: the bin's bytes are handed to the unmodified reference
 * classes compiled from /root/reference/src/c++/lib/build (oracle/Makefile) the way BinSorter::collectGaps / realignGaps drive them
 * (BinSorter.cpp:389-418), and the Index entries are flattened afterwards.  The reference keeps the two gap lists and the CIGAR
 * buffer private; this translation unit alone includes GapRealigner.hh with its two classes opened, to read them out and to size the buffer.
 */
#include <mutex>
#include <vector>
#include <cstring>

// everything GapRealigner.hh includes first, untouched; then the header itself with its two classes opened
#include "alignment/Cigar.hh"
#include "alignment/TemplateLengthStatistics.hh"
#include "build/BarcodeBamMapping.hh"
#include "build/PackedFragmentBuffer.hh"
#include "flowcell/BarcodeMetadata.hh"
#include "build/gapRealigner/Gap.hh"
#include "reference/Contig.hh"
#include "reference/ReferencePosition.hh"
#define class struct
#define private public
#include "build/GapRealigner.hh"
#undef private
#undef class
#include "reference/Contig.hh"
#include "io/Fragment.hh"

#include "oracle_api.h"

using namespace isaac;

const std::vector<reference::Contig> &oracleContigs(const oracle_genome_t *g);      // ref_capi.cpp

namespace
{
/// the contig list of lists GapRealigner wants, cached like the flat one
const std::vector<std::vector<reference::Contig> > &contigLists(const oracle_genome_t *g)
{
    static std::mutex mutex;
    static std::vector<std::vector<reference::Contig> > lists;
    static const reference::Contig *key = 0; static size_t keySize = 0;
    const std::vector<reference::Contig> &flat = oracleContigs(g);
    std::lock_guard<std::mutex> lock(mutex);
    if (key != flat.data() || keySize != flat.size() || lists.empty() || lists[0].size() != flat.size() ||
        (flat.size() && lists[0][0].forward_ != flat[0].forward_))
    {
        lists.assign(1, flat);
        key = flat.data(); keySize = flat.size();
    }
    return lists;
}
}

extern "C" int oracle_realign_bin(const oracle_genome_t *genome, const isaac_ext_realign_options_t *o, uint8_t *data, uint64_t dataBytes,
                                  const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index,
                                  uint64_t indexCount, uint64_t *positionOut, uint32_t *cigarOffsetOut, uint32_t *cigarLengthOut,
                                  uint32_t *cigarsOut, uint64_t cigarCapacity, isaac_ext_gap_t *gapsOut, isaac_ext_gap_t *deletionsOut,
                                  uint64_t gapCapacity, uint64_t *countsOut)
{
    try
    {
        const std::vector<std::vector<reference::Contig> > &contigList = contigLists(genome);
        flowcell::BarcodeMetadataList barcodes;
        std::vector<alignment::TemplateLengthStatistics> tls;
        unsigned groups = 1;
        for (uint32_t b = 0; b < o->barcodeCount; ++b)
        {
            flowcell::BarcodeMetadata barcode("fc", 0, 1, 0, false, flowcell::SequencingAdapterMetadataList());
            barcode.setIndex(b);
            barcodes.push_back(barcode);
            const isaac_ext_tls_t &t = o->barcodeTls[b];
            tls.push_back(alignment::TemplateLengthStatistics(
                t.min, t.max, t.median, t.lowStdDev, t.highStdDev, alignment::TemplateLengthStatistics::AlignmentModel(t.bestModel[0]),
                alignment::TemplateLengthStatistics::AlignmentModel(t.bestModel[1]), t.mateDriftRange));
            if (o->barcodeGapGroup) groups = std::max(groups, o->barcodeGapGroup[b] + 1);
        }
        build::PackedFragmentBuffer buffer;
        std::vector<char> &bytes = (std::vector<char> &)buffer;                 // PackedFragmentBuffer IS its bytes (private base)
        bytes.assign(reinterpret_cast<const char *>(data), reinterpret_cast<const char *>(data) + dataBytes);

        // BinSorter::collectGaps
        std::vector<build::RealignerGaps> realignerGaps(groups);
        std::vector<uint64_t> walked;
        if (!recordOffset)
        {
            for (uint64_t p = 0; p < dataBytes; p += buffer.getFragment(p).getTotalLength()) walked.push_back(p);
            recordOffset = walked.data(); recordCount = walked.size();
        }
        for (uint64_t r = 0; r < recordCount; ++r)
        {
            const io::FragmentAccessor &fragment = buffer.getFragment(recordOffset[r]);
            if (fragment.gapCount_)
            {
                const unsigned group = o->barcodeGapGroup ? o->barcodeGapGroup[fragment.barcode_] : 0;
                realignerGaps.at(group).addGaps(fragment.fStrandPosition_, fragment.cigarBegin(), fragment.cigarEnd());
            }
        }
        for (build::RealignerGaps &g : realignerGaps) g.finalizeGaps();
        uint64_t nGaps = 0, nDeletions = 0;
        for (unsigned group = 0; group < groups; ++group)
        {
            for (const build::gapRealigner::Gap &gap : realignerGaps[group].gapGroups_)
            {
                if (nGaps < gapCapacity) gapsOut[nGaps] = isaac_ext_gap_t{gap.pos_.getValue(), gap.length_, group};
                ++nGaps;
            }
            for (const build::gapRealigner::Gap &gap : realignerGaps[group].deletionEndGroups_)
            {
                if (nDeletions < gapCapacity) deletionsOut[nDeletions] = isaac_ext_gap_t{gap.pos_.getValue(), gap.length_, group};
                ++nDeletions;
            }
        }

        // BinSorter::realignGaps
        build::GapRealigner realigner(o->realignGapsVigorously, o->realignDodgyFragments, 1, o->mismatchCost, o->gapOpenCost,
                                      o->gapExtendCost, o->clipSemialigned, barcodes, tls, contigList);
        realigner.realignedCigars_.reserve(indexCount * 160 + 1024);            // never reallocates: Index points into it
        std::vector<build::PackedFragmentBuffer::Index> indexes;
        for (uint64_t i = 0; i < indexCount; ++i)
        {
            const io::FragmentAccessor &fragment = buffer.getFragment(index[i].dataOffset);
            build::PackedFragmentBuffer::Index idx(fragment.fStrandPosition_, index[i].dataOffset, index[i].mateDataOffset,
                                                  fragment.cigarBegin(), fragment.cigarEnd());
            idx.mateDataOffset_ = index[i].mateDataOffset;                      // that constructor stores dataOffset twice
            indexes.push_back(idx);
        }
        const reference::ReferencePosition binStart(o->binStart), binEnd(o->binEnd);
        for (build::PackedFragmentBuffer::Index &idx : indexes)
        {
            io::FragmentAccessor &fragment = buffer.getFragment(idx);
            const unsigned group = o->barcodeGapGroup ? o->barcodeGapGroup[fragment.barcode_] : 0;
            realigner.realign(realignerGaps.at(group), binStart, binEnd, idx, fragment, buffer);
        }
        uint64_t words = 0;
        for (uint64_t i = 0; i < indexCount; ++i)
        {
            const build::PackedFragmentBuffer::Index &idx = indexes[i];
            const io::FragmentAccessor &fragment = buffer.getFragment(idx);
            positionOut[i] = idx.pos_.getValue();
            cigarLengthOut[i] = uint32_t(idx.cigarEnd_ - idx.cigarBegin_);
            if (idx.cigarBegin_ == fragment.cigarBegin()) cigarOffsetOut[i] = 0xFFFFFFFFu;
            else
            {
                cigarOffsetOut[i] = uint32_t(words);
                for (const uint32_t *c = idx.cigarBegin_; c != idx.cigarEnd_; ++c, ++words) if (words < cigarCapacity) cigarsOut[words] = *c;
            }
        }
        std::memcpy(data, bytes.data(), dataBytes);
        countsOut[0] = nGaps; countsOut[1] = nDeletions; countsOut[2] = words;
        return (nGaps > gapCapacity || words > cigarCapacity) ? ISAAC_EXT_E_INVALID_ARG : ISAAC_EXT_OK;
    }
    catch (const std::exception &e)
    {
        std::cerr << "oracle_realign_bin: " << e.what() << std::endl;
        return ISAAC_EXT_E_INVALID_ARG;
    }
}
