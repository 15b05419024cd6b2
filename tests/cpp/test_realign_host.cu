// realign_device.cuh on the CPU: realignTemplate -- what every thread of realignBinKernel runs -- over a bin in host memory, with the
// gap lists built here the plain way (collect, std::sort, std::unique like RealignerGaps::finalizeGaps).  tests/test_realign_host.py
// compares it with the reference's own build::GapRealigner (oracle_realign_bin).  Host code of an nvcc-compiled shared library, no
// CUDA call.  TEST CODE, not a product path.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../isaac_aligner_b200/csrc/realign_device.cuh"

using namespace isaac_b200;

namespace
{
struct PackedReference
{
    std::vector<uint64_t> offset, length;
    std::vector<uint32_t> bases2, nmask;
    ReferenceView view{};
    PackedReference(uint32_t contigCount, const char *bases, const uint64_t *contigBegin) : offset(contigCount), length(contigCount)
    {
        uint64_t total = 0;
        for (uint32_t c = 0; c < contigCount; ++c) { offset[c] = total; length[c] = contigBegin[c + 1] - contigBegin[c]; total += (length[c] + 127) / 128 * 128; }
        total += 128;
        bases2.assign(total / 16 + 2, 0); nmask.assign(total / 32 + 2, 0);
        for (uint32_t c = 0; c < contigCount; ++c)
            for (uint64_t i = 0; i < length[c]; ++i)
            {
                const char b = bases[contigBegin[c] + i];
                const uint64_t g = offset[c] + i;
                const unsigned code = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 4u;
                if (code < 4) bases2[g >> 4] |= code << ((g & 15) * 2); else nmask[g >> 5] |= 1u << (g & 31);
            }
        view.bases2 = bases2.data(); view.nmask = nmask.data(); view.contigOffset = offset.data(); view.contigLength = length.data();
        view.contigCount = contigCount; view.totalBases = total;
    }
};
}

/// contigs: ASCII ACGTN back to back at contigBegin.  Outputs as isaac_ext_realign_result_t's arrays; gapsOut / deletionsOut with room
/// for gapCapacity entries each, countsOut = {gaps, deletions, cigar words, realigned fragments, error flags}
extern "C" int realign_bin_host(uint32_t contigCount, const char *bases, const uint64_t *contigBegin, const isaac_ext_realign_options_t *o,
                                uint8_t *data, uint64_t dataBytes, const uint64_t *recordOffset, uint64_t recordCount,
                                const isaac_ext_bin_index_t *index, uint64_t indexCount, uint64_t *positionOut, uint32_t *cigarOffsetOut,
                                uint32_t *cigarLengthOut, uint32_t *cigarsOut, uint64_t cigarCapacity, isaac_ext_gap_t *gapsOut,
                                isaac_ext_gap_t *deletionsOut, uint64_t gapCapacity, uint64_t *countsOut)
{
    PackedReference reference(contigCount, bases, contigBegin);
    // ---- BinSorter::collectGaps
    std::vector<uint64_t> walked;
    if (!recordOffset)
    {
        for (uint64_t p = 0; p < dataBytes; p += binRecordLength(data + p)) walked.push_back(p);
        recordOffset = walked.data(); recordCount = walked.size();
    }
    uint32_t groups = 1;
    if (o->barcodeGapGroup) for (uint32_t b = 0; b < o->barcodeCount; ++b) groups = std::max(groups, o->barcodeGapGroup[b] + 1);
    std::vector<isaac_ext_gap_t> gaps;
    for (uint64_t r = 0; r < recordCount; ++r)
    {
        const uint8_t *record = data + recordOffset[r];
        if (!binGet16(record + BIN_GAP_COUNT)) continue;
        const uint32_t group = o->barcodeGapGroup ? o->barcodeGapGroup[binGet64(record + BIN_BARCODE)] : 0;
        const unsigned readLength = binGet16(record + BIN_READ_LENGTH), cigarLength = binGet16(record + BIN_CIGAR_LENGTH);
        const uint64_t start = binGet64(record + BIN_F_STRAND_POSITION);
        uint64_t pos = start;
        for (unsigned k = 0; k < cigarLength; ++k)
        {
            const uint32_t w = binGet32(record + BIN_HEADER_BYTES + readLength + 4 * k), length = w >> 4, op = w & 0xF;
            if (op == ISAAC_EXT_CIGAR_ALIGN) pos += 2ull * length;
            else if (op == ISAAC_EXT_CIGAR_INSERT) gaps.push_back(isaac_ext_gap_t{pos, -int32_t(length), group});
            else if (op == ISAAC_EXT_CIGAR_DELETE) { gaps.push_back(isaac_ext_gap_t{pos, int32_t(length), group}); pos += 2ull * length; }
        }
    }
    auto byStart = [](const isaac_ext_gap_t &a, const isaac_ext_gap_t &b) {
        return a.group != b.group ? a.group < b.group : a.position != b.position ? a.position < b.position : a.length < b.length; };
    std::sort(gaps.begin(), gaps.end(), byStart);
    gaps.erase(std::unique(gaps.begin(), gaps.end(), [](const isaac_ext_gap_t &a, const isaac_ext_gap_t &b) {
        return a.group == b.group && a.position == b.position && a.length == b.length; }), gaps.end());
    // the deletions by end: the reference's std::sort on each group's list in gapGroups_ order (GapRealigner.cpp:91-93)
    std::vector<isaac_ext_gap_t> deletions;
    std::vector<uint32_t> gapGroupBegin(groups + 1, 0), deletionGroupBegin(groups + 1, 0);
    for (uint32_t g = 0; g < groups; ++g)
    {
        const size_t before = deletions.size();
        for (const isaac_ext_gap_t &gap : gaps) if (gap.group == g && gap.length > 0) deletions.push_back(gap);
        std::sort(deletions.begin() + before, deletions.end(), [](const isaac_ext_gap_t &a, const isaac_ext_gap_t &b) {
            return a.position + 2ull * uint64_t(a.length) < b.position + 2ull * uint64_t(b.length); });
        deletionGroupBegin[g + 1] = uint32_t(deletions.size());
        gapGroupBegin[g + 1] = uint32_t(std::count_if(gaps.begin(), gaps.end(), [g](const isaac_ext_gap_t &x) { return x.group <= g; }));
    }
    for (size_t k = 0; k < gaps.size() && k < gapCapacity; ++k) gapsOut[k] = gaps[k];
    for (size_t k = 0; k < deletions.size() && k < gapCapacity; ++k) deletionsOut[k] = deletions[k];
    // ---- BinSorter::realignGaps
    std::vector<uint32_t> recordIndex((dataBytes >> 6) + 1, 0xFFFFFFFFu);
    for (uint64_t i = 0; i < indexCount; ++i) recordIndex[index[i].dataOffset >> 6] = uint32_t(i);
    unsigned long long poolUsed = 0, realigned = 0;
    uint32_t errors = 0;
    RealignBinView v{};
    v.data = data; v.dataBytes = dataBytes; v.index = index; v.indexCount = indexCount; v.recordIndex = recordIndex.data();
    v.gaps = gaps.data(); v.gapGroupBegin = gapGroupBegin.data(); v.deletions = deletions.data(); v.deletionGroupBegin = deletionGroupBegin.data();
    v.barcodeGapGroup = o->barcodeGapGroup; v.barcodeTls = o->barcodeTls; v.barcodeCount = o->barcodeCount;
    v.ref = reference.view;
    v.binStart = realignP(o->binStart); v.binEnd = realignP(o->binEnd);
    v.vigorous = o->realignGapsVigorously; v.dodgy = o->realignDodgyFragments; v.clipSemialigned = o->clipSemialigned;
    v.mismatchCost = o->mismatchCost; v.gapOpenCost = o->gapOpenCost; v.gapExtendCost = o->gapExtendCost;
    v.position = positionOut; v.cigarOffset = cigarOffsetOut; v.cigarLength = cigarLengthOut;
    v.cigarPool = cigarsOut; v.cigarPoolUsed = &poolUsed; v.cigarPoolCapacity = cigarCapacity;
    v.realignedFragments = &realigned; v.errorFlags = &errors;
    for (uint64_t i = 0; i < indexCount; ++i) realignTemplate(v, i);
    countsOut[0] = gaps.size(); countsOut[1] = deletions.size(); countsOut[2] = poolUsed; countsOut[3] = realigned; countsOut[4] = errors;
    return 0;
}
