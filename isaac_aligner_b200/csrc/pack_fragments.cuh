// the bin records matchSelector::FragmentCollector::add leaves in its FragmentBuffer for every fragment of a
// stored template (FragmentCollector.cpp:42-103): io::FragmentHeader (Fragment.hh:73-404) + the read's BCL bytes + its CIGAR.
//
// One warp per cluster.  Per record the warp assembles header, bases and CIGAR in a shared-memory staging area (lane 0 the 112
// header bytes, all lanes the bases and the CIGAR words), then streams the whole fixed-size slot -- the used bytes and the zeros
// behind them -- to the record buffer with aligned 32-bit stores, so that every byte of the buffer is written exactly once and no
// memset pass is needed.  It is byte work bound by HBM: per cluster about 300 B of BCL, 2 x 64 B fragment records and 16 B of
// template in, recordLength (1644 B for 2 x 150) out.
//
// Everything but the warp plumbing is ISAAC_HD and takes (lane, lanes): tests/cpp/test_pack_fragments.cpp runs the same functions
// lane after lane on the CPU against the reference's own io::FragmentHeader (tests/test_tile_write_bin_records.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include "../../include/isaac_ext.h"

#ifndef ISAAC_HD
#ifdef __CUDACC__
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif
#endif

namespace isaac_b200
{

/// io::FragmentHeader as g++ lays it out on x86-64 (Fragment.hh:260-404; offsets checked against the reference's struct by the test)
struct PackedFragmentHeader
{
    int32_t  bamTlen;                 //   0
    uint32_t observedLength;          //   4
    uint64_t fStrandPosition;         //   8  reference::ReferencePosition::value_
    uint16_t lowClipped;              //  16
    uint16_t highClipped;             //  18
    uint16_t alignmentScore;          //  20
    uint16_t templateAlignmentScore;  //  22
    uint64_t mateFStrandPosition;     //  24
    uint16_t readLength;              //  32
    uint16_t cigarLength;             //  34
    uint16_t gapCount;                //  36
    uint16_t editDistance;            //  38
    uint16_t flags;                   //  40  Flags bit fields, first member = bit 0 (Fragment.hh:348-366)
    uint16_t pad0[3];                 //  42
    uint64_t tile;                    //  48
    uint64_t barcode;                 //  56
    uint64_t barcodeSequence;         //  64
    uint64_t clusterId;               //  72
    int32_t  clusterX;                //  80
    int32_t  clusterY;                //  84
    uint64_t duplicateClusterRank;    //  88
    uint64_t mateAnchor;              //  96  FragmentIndexAnchor::value_
    uint32_t mateStorageBin;          // 104
    uint32_t pad1;                    // 108
};
static_assert(sizeof(PackedFragmentHeader) == 112, "io::FragmentHeader is 112 bytes");
static_assert(offsetof(PackedFragmentHeader, tile) == 48 && offsetof(PackedFragmentHeader, mateStorageBin) == 104, "io::FragmentHeader layout");

enum : uint16_t
{
    PACK_FLAG_PAIRED = 1u << 0, PACK_FLAG_UNMAPPED = 1u << 1, PACK_FLAG_MATE_UNMAPPED = 1u << 2, PACK_FLAG_REVERSE = 1u << 3,
    PACK_FLAG_MATE_REVERSE = 1u << 4, PACK_FLAG_FIRST_READ = 1u << 5, PACK_FLAG_SECOND_READ = 1u << 6, PACK_FLAG_FAIL_FILTER = 1u << 7,
    PACK_FLAG_PROPER_PAIR = 1u << 8
};
constexpr int32_t PACK_POSITION_NOT_SET = 0x7FFFFFFF;               // ClusterXy::POSITION_NOT_SET (Cluster.hh:47)
constexpr uint32_t PACK_NO_MATCH_CONTIG = 0x7FFFFFu;                // ReferencePosition::MAX_CONTIG_ID (ReferencePosition.hh:177)
constexpr unsigned PACK_HEADER_BYTES = 112;

/// Cigar::getMaxOpeations (Cigar.hh:181-193)
ISAAC_HD inline unsigned packMaxCigarOperations(const unsigned readLength) { return 2u + 2u + 1u + (readLength / 10u) * 2u; }
/// io::FragmentHeader::getMaxTotalLength (Fragment.hh:188-207): getTotalLength multiplies the cigar length by the word size and is
/// handed Cigar::getMaxLength, which is in bytes already, so a slot has room for four times the maximum number of operations
ISAAC_HD inline unsigned packMaxTotalLength(const unsigned readLength)
{
    return PACK_HEADER_BYTES + readLength + packMaxCigarOperations(readLength) * 4u * 4u;
}

/// reference::ReferencePosition(contigId, position).getValue() (ReferencePosition.hh:68-78, 174-176)
ISAAC_HD inline uint64_t packReferencePosition(const uint64_t contigId, const uint64_t position)
{
    return (((contigId + 1) << 40) | position) << 1;
}
ISAAC_HD inline uint64_t packNoMatchPosition() { return uint64_t(PACK_NO_MATCH_CONTIG) << 41; }            // :61-62
ISAAC_HD inline bool packIsNoMatch(const isaac_ext_fragment_t &f) { return f.contigId == PACK_NO_MATCH_CONTIG; }   // FragmentMetadata.hh:260
/// FragmentMetadata::getFStrandReferencePosition (FragmentMetadata.hh:90-95)
ISAAC_HD inline uint64_t packFStrandPosition(const isaac_ext_fragment_t &f)
{
    return packIsNoMatch(f) ? packNoMatchPosition() : packReferencePosition(f.contigId, uint64_t(f.position));
}
/// FragmentMetadata::getEndReferencePosition (:116-121)
ISAAC_HD inline uint64_t packEndPosition(const isaac_ext_fragment_t &f)
{
    return packIsNoMatch(f) ? packNoMatchPosition() : packReferencePosition(f.contigId, uint64_t(f.position + long(f.observedLength)));
}
/// FragmentMetadata::getStrandReferencePosition (:97-107)
ISAAC_HD inline uint64_t packStrandPosition(const isaac_ext_fragment_t &f)
{
    if (packIsNoMatch(f)) return packNoMatchPosition();
    if (!f.reverse) return packReferencePosition(f.contigId, uint64_t(f.position));
    const long end = f.position + long(f.observedLength);
    return packReferencePosition(f.contigId, uint64_t((end > 1L ? end : 1L) - 1));
}
/// ReferencePosition::getLocation (:99-104)
ISAAC_HD inline uint64_t packLocation(const uint64_t value) { return (value >> 1) - (1ull << 40); }

/// io::FragmentHeader::getTlen (Fragment.hh:209-237)
ISAAC_HD inline int32_t packTlen(const isaac_ext_fragment_t &fragment, const isaac_ext_fragment_t &mate, const bool firstRead)
{
    if (!fragment.cigarLength || !mate.cigarLength) return 0;
    const uint64_t fb = packFStrandPosition(fragment), fe = packEndPosition(fragment);
    const uint64_t mb = packFStrandPosition(mate), me = packEndPosition(mate);
    const uint64_t distance = packLocation(fe < me ? me : fe) - packLocation(mb < fb ? mb : fb);
    const long ret = fb < mb ? long(distance) : (fb > mb || !firstRead) ? long(0ull - distance) : long(distance);
    return int32_t(ret);
}

/// oligo::getReverseBcl (Nucleotides.hh:153-156)
ISAAC_HD inline uint8_t packReverseBcl(const uint8_t bcl) { return (bcl & 0xFCu) ? uint8_t((bcl & 0xFCu) | (3u - (bcl & 3u))) : uint8_t(0); }
/// the quality Read::decodeBcl keeps for a BCL byte (Read.cpp:54-68)
ISAAC_HD inline unsigned packBclQuality(const uint8_t bcl) { return (bcl & 0xFCu) ? unsigned(bcl >> 2) : 2u; }

/// what the pack pass reads of one tile; every pointer is device memory in the kernel, host memory in the CPU test
struct PackView
{
    const isaac_ext_template_t *templates;     // clusterCount
    const isaac_ext_fragment_t *fragments;     // clusterCount * readCount
    const uint32_t *cigars;
    const uint8_t *bcl;                        // clusterCount * (readLength[0] + readLength[1])
    uint64_t bclBytes;
    const uint8_t *pf;                         // or null
    const int32_t *xy;                         // or null
    const uint64_t *barcodeSequence;           // or null
    const uint64_t *contigBinBegin;            // or null (no bin map)
    const uint32_t *binIndex;
    uint32_t contigCount, distributionBinSize;
    uint64_t tile;
    uint32_t barcodeIdx, keepUnaligned;
    uint32_t clusterCount, readCount;
    uint32_t readLength[2];
    uint32_t recordLength, readOffset[2];
    const uint64_t *recordOffset;              // null: FragmentBuffer slots; else clusterCount * readCount + 1 byte offsets (compact)
    uint8_t *records;                          // clusterCount * recordLength, or recordOffset[clusterCount * readCount] bytes
    uint64_t *fStrandPos;                      // clusterCount * readCount
    uint8_t *initialized;                      // clusterCount * readCount
};

/// FragmentBuffer::getRecordLength / getReadOffsets (FragmentCollector.hh:283-308)
ISAAC_HD inline void packLayout(PackView &v)
{
    v.readOffset[0] = 0;
    v.readOffset[1] = v.readCount > 1 ? packMaxTotalLength(v.readLength[0]) : 0;
    v.recordLength = packMaxTotalLength(v.readLength[0]) + (v.readCount > 1 && v.readLength[1] ? packMaxTotalLength(v.readLength[1]) : 0);
}

/// BinIndexMap::getBinIndex (BinIndexMap.hh:96-107) of an FStrand position value; no-match positions and positions outside the
/// map (the reference asserts) give bin 0
ISAAC_HD inline uint32_t packBinIndex(const PackView &v, const isaac_ext_fragment_t &mate)
{
    if (!v.contigBinBegin || !v.distributionBinSize || packIsNoMatch(mate) || mate.contigId >= v.contigCount || mate.position < 0) return 0;
    const uint64_t index = uint64_t(mate.position) / v.distributionBinSize;
    const uint64_t begin = v.contigBinBegin[mate.contigId], end = v.contigBinBegin[mate.contigId + 1];
    return begin + index < end ? v.binIndex[begin + index] : 0u;
}

/// oligo::pack32BclBases (Nucleotides.hh:241-278) at 'offset' of the tile's BCL bytes; bytes past the end of the tile count as 0
ISAAC_HD inline uint64_t packShadowBases(const PackView &v, const uint64_t offset)
{
    uint64_t ret = 0;
    for (unsigned i = 0; i < 32; ++i)
        if (offset + i < v.bclBytes) ret |= uint64_t(v.bcl[offset + i] & 3u) << (2 * i);
    return ret;
}

/// a lane's share of BamTemplate::getQuality (BamTemplate.hh:69-74, FragmentMetadata.hh:270-275): the sum over the lanes is the
/// sum of the forward qualities of all reads of the cluster
ISAAC_HD inline unsigned packQualityShare(const PackView &v, const uint32_t cluster, const unsigned lane, const unsigned lanes)
{
    const unsigned total = v.readLength[0] + (v.readCount > 1 ? v.readLength[1] : 0);
    const uint8_t *b = v.bcl + size_t(cluster) * total;
    unsigned sum = 0;
    for (unsigned i = lane; i < total; i += lanes) sum += packBclQuality(b[i]);
    return sum;
}

/// MatchSelector::processMatchList stores the template (MatchSelector.cpp:311-314, 331-347, 355-358)
ISAAC_HD inline bool packStores(const PackView &v, const uint32_t cluster) { return v.templates[cluster].built || v.keepUnaligned; }

/// io::FragmentHeader's constructors (Fragment.hh:100-186) for read r of the cluster; 'quality' = BamTemplate::getQuality
ISAAC_HD inline PackedFragmentHeader packHeader(const PackView &v, const uint32_t cluster, const unsigned r, const unsigned quality)
{
    const isaac_ext_template_t t = v.templates[cluster];
    const unsigned rc = v.readCount;
    const isaac_ext_fragment_t &fragment = v.fragments[size_t(cluster) * rc + r];
    const bool aligned = fragment.cigarLength != 0;
    PackedFragmentHeader h;
    h.pad0[0] = h.pad0[1] = h.pad0[2] = 0; h.pad1 = 0;
    h.observedLength = aligned ? fragment.observedLength : 0u;                                   // FragmentMetadata.hh:85
    h.lowClipped = fragment.lowClipped; h.highClipped = fragment.highClipped;
    h.alignmentScore = uint16_t(t.fragmentAlignmentScore[r]);
    h.readLength = uint16_t(v.readLength[r]);
    h.cigarLength = fragment.cigarLength;
    h.gapCount = fragment.gapCount;
    h.editDistance = fragment.editDistance;
    h.tile = uint32_t(v.tile);                                                                  // Cluster::tile_ is unsigned (Cluster.hh:79)
    h.barcode = v.barcodeIdx;
    h.barcodeSequence = v.barcodeSequence ? v.barcodeSequence[cluster] : 0;
    h.clusterId = cluster;
    const bool xySet = v.xy && v.xy[2 * size_t(cluster)] != PACK_POSITION_NOT_SET;               // ClusterXy::isSet (Cluster.hh:52)
    h.clusterX = xySet ? v.xy[2 * size_t(cluster)] : PACK_POSITION_NOT_SET;
    h.clusterY = xySet ? v.xy[2 * size_t(cluster) + 1] : PACK_POSITION_NOT_SET;
    const bool failFilter = v.pf && !v.pf[cluster];
    if (rc == 2)
    {
        const isaac_ext_fragment_t &mate = v.fragments[size_t(cluster) * rc + (1 - r)];          // BamTemplate::getMateFragmentMetadata (BamTemplate.hh:99)
        const bool mateAligned = mate.cigarLength != 0;
        h.bamTlen = packTlen(fragment, mate, r == 0);
        h.fStrandPosition = aligned ? packFStrandPosition(fragment) : packFStrandPosition(mate);
        h.templateAlignmentScore = uint16_t(t.properPair ? t.alignmentScore : t.fragmentAlignmentScore[r]);
        h.mateFStrandPosition = mateAligned ? packFStrandPosition(mate) : packFStrandPosition(fragment);
        h.flags = uint16_t(PACK_FLAG_PAIRED | (aligned ? 0 : PACK_FLAG_UNMAPPED) | (mateAligned ? 0 : PACK_FLAG_MATE_UNMAPPED) |
                           (fragment.reverse ? PACK_FLAG_REVERSE : 0) | (mate.reverse ? PACK_FLAG_MATE_REVERSE : 0) |
                           (r == 0 ? PACK_FLAG_FIRST_READ : 0) | (r == 1 ? PACK_FLAG_SECOND_READ : 0) |
                           (failFilter ? PACK_FLAG_FAIL_FILTER : 0) | (t.properPair ? PACK_FLAG_PROPER_PAIR : 0));
        // getTemplateDuplicateRank (Fragment.hh:66-71): the middle term is 32-bit arithmetic
        const unsigned editDistance = unsigned(fragment.editDistance) + unsigned(mate.editDistance);
        const unsigned totalLength = v.readLength[0] + v.readLength[1];
        h.duplicateClusterRank = (uint64_t(quality) << 32) | uint64_t(uint32_t((totalLength - editDistance) << 16)) | uint64_t(t.alignmentScore);
        // FragmentIndexAnchor(mate) (Fragment.hh:489-497)
        const uint64_t mateBcl = size_t(cluster) * totalLength + (r == 0 ? v.readLength[0] : 0);
        h.mateAnchor = mateAligned ? packStrandPosition(mate) : packShadowBases(v, mateBcl);
        // FragmentCollector.cpp:57-71
        h.mateStorageBin = packIsNoMatch(fragment) ? 0u : packBinIndex(v, mate);
    }
    else
    {
        h.bamTlen = 0;
        h.fStrandPosition = packFStrandPosition(fragment);
        h.templateAlignmentScore = uint16_t(t.fragmentAlignmentScore[r]);
        h.mateFStrandPosition = packNoMatchPosition();
        h.flags = uint16_t((aligned ? 0 : PACK_FLAG_UNMAPPED) | PACK_FLAG_MATE_UNMAPPED | (fragment.reverse ? PACK_FLAG_REVERSE : 0) |
                           PACK_FLAG_FIRST_READ | PACK_FLAG_SECOND_READ | (failFilter ? PACK_FLAG_FAIL_FILTER : 0));
        h.duplicateClusterRank = 0; h.mateAnchor = 0; h.mateStorageBin = 0;
    }
    return h;
}

/// bytes of the record that carry data: header + bases + the CIGAR of an aligned fragment, never more than the slot
ISAAC_HD inline unsigned packUsedBytes(const PackView &v, const isaac_ext_fragment_t &fragment, const unsigned r)
{
    const unsigned used = PACK_HEADER_BYTES + v.readLength[r] + 4u * fragment.cigarLength;
    const unsigned slot = packMaxTotalLength(v.readLength[r]);
    return used < slot ? used : slot;
}

/// storeBclAndCigar (FragmentCollector.cpp:79-103) into staging[PACK_HEADER_BYTES ..), this lane's share
ISAAC_HD inline void packStageData(const PackView &v, const uint32_t cluster, const unsigned r, uint8_t *staging, const unsigned used,
                                   const unsigned lane, const unsigned lanes)
{
    const isaac_ext_fragment_t &fragment = v.fragments[size_t(cluster) * v.readCount + r];
    const unsigned L = v.readLength[r], total = v.readLength[0] + (v.readCount > 1 ? v.readLength[1] : 0);
    const uint8_t *b = v.bcl + size_t(cluster) * total + (r ? v.readLength[0] : 0);
    uint8_t *out = staging + PACK_HEADER_BYTES;
    if (fragment.reverse) for (unsigned i = lane; i < L; i += lanes) out[i] = packReverseBcl(b[L - 1 - i]);
    else for (unsigned i = lane; i < L; i += lanes) out[i] = b[i];
    const unsigned cigarBytes = used - PACK_HEADER_BYTES - L;
    const uint32_t *cigar = v.cigars + fragment.cigarOffset;
    for (unsigned k = lane; k < cigarBytes; k += lanes) out[L + k] = uint8_t(cigar[k >> 2] >> (8u * (k & 3u)));
}

/// four staging bytes starting at byte 'offset' of a 4-byte aligned staging area (which has 8 spare bytes behind its end)
ISAAC_HD inline uint32_t packLoadUnaligned(const uint32_t *staging, const unsigned offset)
{
    const uint32_t lo = staging[offset >> 2], hi = staging[(offset >> 2) + 1];
    const unsigned shift = (offset & 3u) * 8u;
    return shift ? (lo >> shift) | (hi << (32u - shift)) : lo;
}

/// the whole slot of one record: staging[0 .. used) then zeros up to 'slot', bytes up to the first 4-byte boundary of the
/// destination, aligned words, bytes behind the last boundary; this lane's share
ISAAC_HD inline void packStoreSlot(uint8_t *dst, const uint32_t *staging, const unsigned used, const unsigned slot,
                                   const unsigned lane, const unsigned lanes)
{
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(staging);
    unsigned head = unsigned((4u - unsigned(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
    if (head > slot) head = slot;
    for (unsigned i = lane; i < head; i += lanes) dst[i] = i < used ? bytes[i] : uint8_t(0);
    const unsigned words = (slot - head) / 4u;
    uint32_t *out = reinterpret_cast<uint32_t *>(dst + head);
    for (unsigned w = lane; w < words; w += lanes)
    {
        const unsigned i = head + 4u * w;
        uint32_t value = 0;
        if (i < used)
        {
            value = packLoadUnaligned(staging, i);
            if (i + 4u > used) value &= (1u << (8u * (used - i))) - 1u;
        }
        out[w] = value;
    }
    for (unsigned i = head + 4u * words + lane; i < slot; i += lanes) dst[i] = i < used ? bytes[i] : uint8_t(0);
}

/// staging bytes one warp needs: the largest used record + the spare words of packLoadUnaligned
ISAAC_HD inline unsigned packStagingBytes(const PackView &v)
{
    const unsigned a = packMaxTotalLength(v.readLength[0]), b = v.readCount > 1 ? packMaxTotalLength(v.readLength[1]) : 0;
    return (((a > b ? a : b) + 7u) & ~7u) + 8u;
}

/// First half of FragmentCollector::add for read r of a cluster, this lane's share: the index entry and the header (lane 0), bases
/// and CIGAR into the staging area.  \return the bytes of the record that carry data, 0 for a template that is not stored.
/// All lanes must be done with it before packStoreRecord starts.
ISAAC_HD inline unsigned packStageRecord(const PackView &v, const uint32_t cluster, const unsigned r, const unsigned quality,
                                         uint8_t *staging, const unsigned lane, const unsigned lanes)
{
    const size_t slotIndex = size_t(cluster) * v.readCount + r;
    if (!packStores(v, cluster))
    {
        if (lane == 0)
        {
            v.fStrandPos[slotIndex] = 0;                                                      // IndexRecord() (FragmentCollector.hh:66-68)
            v.initialized[slotIndex] = 0;
        }
        return 0;
    }
    const isaac_ext_fragment_t &fragment = v.fragments[slotIndex];
    const unsigned used = packUsedBytes(v, fragment, r);
    if (lane == 0)
    {
        *reinterpret_cast<PackedFragmentHeader *>(staging) = packHeader(v, cluster, r, quality);
        v.fStrandPos[slotIndex] = packFStrandPosition(fragment);                              // FragmentCollector.cpp:53
        v.initialized[slotIndex] = 1;
    }
    packStageData(v, cluster, r, staging, used, lane, lanes);
    return used;
}

/// Second half: the slot of read r of the cluster from the staging area, this lane's share
ISAAC_HD inline void packStoreRecord(const PackView &v, const uint32_t cluster, const unsigned r, const unsigned used,
                                     const uint8_t *staging, const unsigned lane, const unsigned lanes)
{
    if (v.recordOffset)         // compact: the used bytes only, at the offset the caller computed from the record lengths
        packStoreSlot(v.records + v.recordOffset[size_t(cluster) * v.readCount + r], reinterpret_cast<const uint32_t *>(staging), used, used,
                      lane, lanes);
    else
        packStoreSlot(v.records + size_t(cluster) * v.recordLength + v.readOffset[r], reinterpret_cast<const uint32_t *>(staging), used,
                      packMaxTotalLength(v.readLength[r]), lane, lanes);
}

/// FragmentHeader::getTotalLength of the record of (cluster, r) (Fragment.hh:192-202), 0 for a template that is not stored: the
/// caller's prefix sum over these is PackView::recordOffset
ISAAC_HD inline unsigned packRecordBytes(const PackView &v, const uint32_t cluster, const unsigned r)
{
    return packStores(v, cluster) ? packUsedBytes(v, v.fragments[size_t(cluster) * v.readCount + r], r) : 0u;
}

#ifdef __CUDACC__
/// FragmentCollector::add for every fragment of every stored template of the tile; dynamic shared memory =
/// warps per block * packStagingBytes
__global__ void packFragmentsKernel(const PackView v, unsigned long long *__restrict__ stored)
{
    extern __shared__ __align__(16) uint8_t packShared[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    uint8_t *staging = packShared + size_t(warp) * packStagingBytes(v);
    unsigned long long mine = 0;
    for (uint32_t cluster = blockIdx.x * warps + warp; cluster < v.clusterCount; cluster += gridDim.x * warps)
    {
        unsigned quality = 0;
        if (v.readCount == 2 && packStores(v, cluster))
        {
            quality = packQualityShare(v, cluster, lane, 32u);
            for (unsigned d = 16; d; d >>= 1) quality += __shfl_xor_sync(0xFFFFFFFFu, quality, d);
        }
        for (unsigned r = 0; r < v.readCount; ++r)
        {
            const unsigned used = packStageRecord(v, cluster, r, quality, staging, lane, 32u);
            if (lane == 0 && used) ++mine;
            __syncwarp();
            packStoreRecord(v, cluster, r, used, staging, lane, 32u);
            __syncwarp();
        }
    }
    if (mine) atomicAdd(stored, mine);
}
#endif

} // namespace isaac_b200
