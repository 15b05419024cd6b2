// pack_fragments.cuh on the CPU: the functions the warps of packFragmentsKernel run, called lane after lane with the two
// phases of a record separated the way __syncwarp separates them.  Built as a shared library by tests/test_tile_write_bin_records.py,
// which compares the records with the reference's own io::FragmentHeader (oracle/_ref).  TEST CODE, not a product path.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../isaac_aligner_b200/csrc/pack_fragments.cuh"

using namespace isaac_b200;

extern "C" int pack_fragments_lanes(const isaac_ext_reads_t *reads, const isaac_ext_template_t *templates,
                                    const isaac_ext_fragment_t *fragments, const uint32_t *cigars,
                                    const isaac_ext_pack_options_t *options, unsigned lanes, unsigned misalign,
                                    uint8_t *recordsOut, uint64_t *fStrandPosOut, uint8_t *initializedOut, uint32_t *layoutOut,
                                    uint64_t *storedOut, uint64_t *recordOffsetOut)
{
    PackView v{};
    v.clusterCount = reads->clusterCount; v.readCount = reads->readCount;
    v.readLength[0] = reads->readLength[0]; v.readLength[1] = reads->readCount > 1 ? reads->readLength[1] : 0;
    packLayout(v);
    v.templates = templates; v.fragments = fragments; v.cigars = cigars;
    v.bcl = reads->bcl; v.bclBytes = uint64_t(v.clusterCount) * (v.readLength[0] + v.readLength[1]);
    v.pf = options->pf; v.xy = options->xy; v.barcodeSequence = options->barcodeSequence;
    if (options->distributionBinSize)
    {
        v.contigBinBegin = options->contigBinBegin; v.binIndex = options->binIndex;
        v.contigCount = options->contigCount; v.distributionBinSize = options->distributionBinSize;
    }
    v.tile = options->tile; v.barcodeIdx = options->barcodeIdx; v.keepUnaligned = options->keepUnaligned;
    // compact: the offsets the ABI computes on the host (isaac_ext_pack.cuh)
    std::vector<uint64_t> offsets(size_t(v.clusterCount) * v.readCount + 1, 0);
    size_t bytes = size_t(v.clusterCount) * v.recordLength;
    if (options->compact)
    {
        for (uint32_t c = 0; c < v.clusterCount; ++c)
            for (unsigned r = 0; r < v.readCount; ++r)
                offsets[size_t(c) * v.readCount + r + 1] = offsets[size_t(c) * v.readCount + r] + packRecordBytes(v, c, r);
        v.recordOffset = offsets.data();
        bytes = offsets.back();
        std::memcpy(recordOffsetOut, offsets.data(), offsets.size() * sizeof(uint64_t));
    }
    // the record buffer at any alignment the caller asks for (cudaMalloc gives 256 bytes, the slots inside are at odd offsets)
    std::vector<uint8_t> buffer(bytes + 32, 0xAB);
    uint8_t *base = buffer.data();
    while (reinterpret_cast<uintptr_t>(base) % 8 != misalign % 8) ++base;
    v.records = base; v.fStrandPos = fStrandPosOut; v.initialized = initializedOut;
    layoutOut[0] = v.recordLength; layoutOut[1] = v.readOffset[0]; layoutOut[2] = v.readOffset[1]; layoutOut[3] = PACK_HEADER_BYTES;
    std::vector<uint64_t> stagingWords(packStagingBytes(v) / 8 + 1);
    uint8_t *staging = reinterpret_cast<uint8_t *>(stagingWords.data());
    uint64_t stored = 0;
    for (uint32_t cluster = 0; cluster < v.clusterCount; ++cluster)
    {
        unsigned quality = 0;
        if (v.readCount == 2 && packStores(v, cluster))
            for (unsigned lane = 0; lane < lanes; ++lane) quality += packQualityShare(v, cluster, lane, lanes);
        for (unsigned r = 0; r < v.readCount; ++r)
        {
            std::memset(staging, 0xCD, packStagingBytes(v));          // what an earlier record left behind must not show
            unsigned used = 0;
            for (unsigned lane = 0; lane < lanes; ++lane) used = packStageRecord(v, cluster, r, quality, staging, lane, lanes);
            if (used) ++stored;
            for (unsigned lane = 0; lane < lanes; ++lane) packStoreRecord(v, cluster, r, used, staging, lane, lanes);
        }
    }
    std::memcpy(recordsOut, base, bytes);
    *storedOut = stored;
    // nothing may be written in front of or behind the records
    for (const uint8_t *p = buffer.data(); p < base; ++p) if (*p != 0xAB) return 2;
    for (const uint8_t *p = base + bytes; p < buffer.data() + buffer.size(); ++p) if (*p != 0xAB) return 3;
    return 0;
}
