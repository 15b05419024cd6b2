// Kernels of isaac_ext_realign_bin (last part): BinSorter::collectGaps and BinSorter::realignGaps of one bin.
//   countRecordGapsKernel / writeRecordGapsKernel   RealignerGaps::addGaps of every record with gaps (GapRealigner.hh:52-97): one
//                                                   thread per record, count pass + exclusive scan + write pass, no atomics
//   (cub merge sort, unique, select)                RealignerGaps::finalizeGaps (GapRealigner.cpp:86-94)
//   gapGroupBeginKernel                             first gap of every gap group in the two sorted lists
//   deletionEndTiesKernel                           do two deletions of a group end at the same base?  (then the reference's
//                                                   unstable std::sort decides their order, see isaac_ext_realign.cuh)
//   recordIndexKernel                               byte offset of a record -> its index entry, for the mate links
//   realignBinKernel                                GapRealigner::realign, one thread per template (realign_device.cuh)
// All of it is short integer work on data that crossed PCIe once; the realign kernel is bound by its divergent per-fragment search.
#pragma once
#include "realign_device.cuh"

namespace isaac_b200
{

/// layout-identical to isaac_ext_gap_t, with the comparisons the CUB passes need
struct GapRecord
{
    uint64_t position; int32_t length; uint32_t group;
    __host__ __device__ bool operator==(const GapRecord &o) const { return position == o.position && length == o.length && group == o.group; }
};
static_assert(sizeof(GapRecord) == sizeof(isaac_ext_gap_t), "GapRecord mirrors isaac_ext_gap_t");

/// by group, then orderByGapStartAndTypeLength (GapRealigner.cpp:46-52)
struct GapByStart
{
    __host__ __device__ bool operator()(const GapRecord &a, const GapRecord &b) const
    {
        return a.group != b.group ? a.group < b.group : a.position != b.position ? a.position < b.position : a.length < b.length;
    }
};
/// by group, then orderByDeletionGapEnd (:54-57)
struct GapByDeletionEnd
{
    __host__ __device__ static uint64_t end(const GapRecord &g) { return g.position + 2ull * uint64_t(g.length); }
    __host__ __device__ bool operator()(const GapRecord &a, const GapRecord &b) const
    {
        return a.group != b.group ? a.group < b.group : end(a) < end(b);
    }
};
struct GapIsDeletion { __host__ __device__ bool operator()(const GapRecord &g) const { return g.length > 0; } };

/// does the record at 'offset' lie inside the bin's data, header and all?
__device__ __forceinline__ bool recordInside(const uint8_t *data, const uint64_t dataBytes, const uint64_t offset)
{
    return offset + BIN_HEADER_BYTES <= dataBytes && offset + binRecordLength(data + offset) <= dataBytes;
}

/// insertions and deletions in the CIGAR of every record that says it has gaps (BinSorter.cpp:391-400)
__global__ void countRecordGapsKernel(const uint8_t *__restrict__ data, const uint64_t dataBytes, const uint64_t *__restrict__ recordOffset,
                                      const uint64_t recordCount, uint32_t *__restrict__ gapsOfRecord, const uint32_t barcodeCount,
                                      uint32_t *__restrict__ errorFlags)
{
    for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < recordCount; r += uint64_t(gridDim.x) * blockDim.x)
    {
        if (!recordInside(data, dataBytes, recordOffset[r])) { atomicOr(errorFlags, REALIGN_ERROR_BOUNDS); gapsOfRecord[r] = 0; continue; }
        const uint8_t *record = data + recordOffset[r];
        uint32_t n = 0;
        if (binGet16(record + BIN_GAP_COUNT))
        {
            if (binGet64(record + BIN_BARCODE) >= barcodeCount) atomicOr(errorFlags, REALIGN_ERROR_BARCODE);
            const unsigned readLength = binGet16(record + BIN_READ_LENGTH), cigarLength = binGet16(record + BIN_CIGAR_LENGTH);
            for (unsigned k = 0; k < cigarLength; ++k)
            {
                const uint32_t op = binGet32(record + BIN_HEADER_BYTES + readLength + 4u * k) & 0xFu;
                n += op == ISAAC_EXT_CIGAR_INSERT || op == ISAAC_EXT_CIGAR_DELETE;
            }
        }
        gapsOfRecord[r] = n;
    }
}

__global__ void writeRecordGapsKernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ recordOffset, const uint64_t recordCount,
                                      const uint32_t *__restrict__ gapsOfRecord, const uint32_t *__restrict__ gapBegin,
                                      const uint32_t *__restrict__ barcodeGapGroup, const uint32_t barcodeCount, GapRecord *__restrict__ gaps)
{
    for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < recordCount; r += uint64_t(gridDim.x) * blockDim.x)
    {
        if (!gapsOfRecord[r]) continue;                                        // also every record that failed the bounds check
        const uint8_t *record = data + recordOffset[r];
        const uint64_t barcode = binGet64(record + BIN_BARCODE);
        const uint32_t group = (barcodeGapGroup && barcode < barcodeCount) ? barcodeGapGroup[barcode] : 0u;
        const unsigned readLength = binGet16(record + BIN_READ_LENGTH), cigarLength = binGet16(record + BIN_CIGAR_LENGTH);
        uint64_t pos = binGet64(record + BIN_F_STRAND_POSITION);                 // ReferencePosition value: one base = 2
        uint32_t at = gapBegin[r];
        for (unsigned k = 0; k < cigarLength; ++k)
        {
            const uint32_t w = binGet32(record + BIN_HEADER_BYTES + readLength + 4u * k), length = w >> 4, op = w & 0xFu;
            if (op == ISAAC_EXT_CIGAR_ALIGN) pos += 2ull * length;
            else if (op == ISAAC_EXT_CIGAR_INSERT) gaps[at++] = GapRecord{pos, -int32_t(length), group};
            else if (op == ISAAC_EXT_CIGAR_DELETE) { gaps[at++] = GapRecord{pos, int32_t(length), group}; pos += 2ull * length; }
        }
    }
}

/// groupBegin[g] = first entry of 'sorted' whose group is not below g (g = 0..groups); count read from device memory
__global__ void gapGroupBeginKernel(const GapRecord *__restrict__ sorted, const uint32_t *__restrict__ count, const uint32_t groups,
                                    uint32_t *__restrict__ groupBegin)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > groups) return;
    uint32_t lo = 0, hi = *count;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sorted[mid].group < g) lo = mid + 1; else hi = mid; }
    groupBegin[g] = lo;
}

__global__ void deletionEndTiesKernel(const GapRecord *__restrict__ byEnd, const uint32_t *__restrict__ count, uint32_t *__restrict__ ties)
{
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x)
        if (byEnd[i].group == byEnd[i + 1].group && GapByDeletionEnd::end(byEnd[i]) == GapByDeletionEnd::end(byEnd[i + 1])) atomicAdd(ties, 1u);
}

/// also the bounds check of the index entries: the realign kernel runs only when no entry points outside the data
__global__ void recordIndexKernel(const uint8_t *__restrict__ data, const uint64_t dataBytes, const isaac_ext_bin_index_t *__restrict__ index,
                                  const uint64_t indexCount, uint32_t *__restrict__ recordIndex, uint32_t *__restrict__ errorFlags)
{
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < indexCount; i += uint64_t(gridDim.x) * blockDim.x)
    {
        if (!recordInside(data, dataBytes, index[i].dataOffset) || !recordInside(data, dataBytes, index[i].mateDataOffset))
        {
            atomicOr(errorFlags, REALIGN_ERROR_BOUNDS);
            continue;
        }
        recordIndex[index[i].dataOffset >> 6] = uint32_t(i);
    }
}

__global__ void __launch_bounds__(128)
realignBinKernel(const RealignBinView v)
{
    if (*v.errorFlags & REALIGN_ERROR_BOUNDS) return;                         // set by the passes before this one
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i < v.indexCount) realignTemplate(v, i);
}

} // namespace isaac_b200
