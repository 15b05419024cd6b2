// isaac_ext_realign_bin: build::GapRealigner over one bin (last part; kernels_realign.cuh, realign_device.cuh).
// Included by isaac_ext.cu.
//
// One upload of the bin (records, record offsets, index), then on the device: the gaps of every record (count / scan / write),
// sorted by (group, start, signed length) and made unique = RealignerGaps::gapGroups_; the deletions among them sorted by
// (group, end) = deletionEndGroups_; one thread per template realigns its one or two index entries; the records (updated in place),
// Index::pos_ and the CIGAR of every entry come back in one download.
//
// The one place where the reference's result depends on its C++ library: deletionEndGroups_ is ordered by an UNSTABLE std::sort on
// the end position alone (GapRealigner.cpp:91-93), and when a fragment's lookup finds only deletions that end inside its span the
// list is handed on in that order (:129-137).  The device sort is stable, i.e. ties keep the (start, length) order; when the bin has
// ties at all (deletionEndTiesKernel), the unique deletions come to the host, are put in gapGroups_ order like remove_copy_if leaves
// them, sorted there with the same std::sort call as the reference and go back: same library, same input order, same result.
#pragma once
#include <cub/cub.cuh>
#include "kernels_realign.cuh"

struct RealignState
{
    DeviceBuffer<uint8_t> dData, dTemp, dChangedHeader;
    DeviceBuffer<uint64_t> dRecordOffset, dPosition, dChangedOffset;
    DeviceBuffer<isaac_ext_bin_index_t> dIndex;
    DeviceBuffer<uint32_t> dGapsOfRecord, dGapBegin, dRecordIndex, dCigarOffset, dCigarLength, dCigarPool, dGroupBegin, dBarcodeGapGroup, dCounters;
    DeviceBuffer<GapRecord> dGapsRaw, dGaps, dDeletions;
    DeviceBuffer<isaac_ext_tls_t> dTls;
    DeviceBuffer<unsigned long long> dLongCounters;
    HostBuffer<uint64_t> walked;
    PinnedBuffer<uint64_t> hPosition, hChangedOffset;
    PinnedBuffer<uint8_t> hChangedHeader;
    PinnedBuffer<uint32_t> hCigarOffset, hCigarLength, hCigarPool, hCounters;
    PinnedBuffer<GapRecord> hGaps, hDeletions;
    PinnedBuffer<unsigned long long> hLongCounters;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;              // the slot's own stream
    void release()
    {
        for (cudaEvent_t &e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        dData.release(); dTemp.release(); dRecordOffset.release(); dPosition.release(); dIndex.release(); dGapsOfRecord.release();
        dGapBegin.release(); dRecordIndex.release(); dCigarOffset.release(); dCigarLength.release(); dCigarPool.release();
        dGroupBegin.release(); dBarcodeGapGroup.release(); dCounters.release(); dGapsRaw.release(); dGaps.release(); dDeletions.release();
        dTls.release(); dLongCounters.release(); dChangedHeader.release(); dChangedOffset.release(); hChangedOffset.release(); hChangedHeader.release();
        hPosition.release(); hCigarOffset.release(); hCigarLength.release(); hCigarPool.release(); hCounters.release(); hGaps.release();
        hDeletions.release(); hLongCounters.release();
    }
};

constexpr unsigned REALIGN_SLOTS = 3;      // measured on B200: 2 slots 14.0 ms, 3 slots 11.5 ms, 4 slots 14.8 ms per 16 bins
struct RealignSlots { RealignState slot[REALIGN_SLOTS]; };

void releaseRealign(RealignSlots *state) { if (state) { for (RealignState &s : state->slot) s.release(); delete state; } }

namespace
{
// dCounters: 0 unique gaps, 1 deletions, 2 ties among deletion ends, 3 error flags
enum { RC_GAPS = 0, RC_DELETIONS = 1, RC_TIES = 2, RC_ERRORS = 3 };
}

namespace
{
/// where a call leaves the per-entry results: the slot's page-locked buffers (null members) or memory of the caller
struct RealignOutputs
{
    uint64_t *position = nullptr; uint32_t *cigarOffset = nullptr, *cigarLength = nullptr, *cigars = nullptr;
    uint64_t cigarCapacity = 0;
};

#define CKR(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) { error = std::string(#call) + ": " + cudaGetErrorString(e_); return ISAAC_EXT_E_CUDA; } } while (0)

/// one bin on one slot (buffers + stream); touches nothing of the context but its resident reference, so that two slots can work
/// on two bins at a time (isaac_ext_realign_bins)
int realignBinOn(const isaac_ext_ctx *ctx, RealignState &st, std::string &error, uint64_t &launches, const isaac_ext_realign_options_t *options,
                 uint8_t *data, uint64_t dataBytes, const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index,
                 uint64_t indexCount, const RealignOutputs &out, isaac_ext_realign_result_t *result)
{
    auto fail = [&](int code, const char *what) { error = what; return code; };
    if (!options || !result || (dataBytes && !data) || (indexCount && !index)) return fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->haveReference) return fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference first");
    if (!options->barcodeCount || !options->barcodeTls) return fail(ISAAC_EXT_E_INVALID_ARG, "the barcodes' template length statistics are missing");
    if (indexCount >= 0xFFFFFFFFull || dataBytes >= (1ull << 38)) return fail(ISAAC_EXT_E_UNSUPPORTED, "bin too large");
    const int64_t binStart = realignP(options->binStart), binEnd = realignP(options->binEnd);
    if (realignContig(binEnd) >= ctx->ref.contigCount || realignContig(binStart) >= ctx->ref.contigCount)
        return fail(ISAAC_EXT_E_INVALID_ARG, "the bin lies on a contig the resident reference does not have");
    CKR(cudaSetDevice(ctx->device));
    for (cudaEvent_t &e : st.ev) if (!e) CKR(cudaEventCreate(&e));
    if (!st.stream) CKR(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
    cudaStream_t s = st.stream;
    PhaseTimer timer("realign_bin");

    // ---- the records of the bin: the caller's offsets, or the chain of FragmentHeader::getTotalLength walked here
    if (!recordOffset)
    {
        size_t n = 0;
        for (uint64_t p = 0; p < dataBytes; ++n)
        {
            if (p + BIN_HEADER_BYTES > dataBytes) return fail(ISAAC_EXT_E_INVALID_ARG, "the bin's data end inside a record");
            p += binRecordLength(data + p);
        }
        st.walked.reserve(n + 1);
        n = 0;
        for (uint64_t p = 0; p < dataBytes; p += binRecordLength(data + p)) st.walked.p[n++] = p;
        recordOffset = st.walked.p; recordCount = n;
    }
    // whether every record and every index entry lies inside the data is checked by the first kernel that touches them
    timer.mark("validate");
    uint32_t groups = 1;
    if (options->barcodeGapGroup)
        for (uint32_t b = 0; b < options->barcodeCount; ++b) groups = std::max(groups, options->barcodeGapGroup[b] + 1);

    // ---- upload
    CKR(st.dData.reserve(dataBytes + 8)); CKR(st.dRecordOffset.reserve(recordCount + 1)); CKR(st.dIndex.reserve(indexCount + 1));
    CKR(st.dGapsOfRecord.reserve(recordCount + 1)); CKR(st.dGapBegin.reserve(recordCount + 1));
    CKR(st.dTls.reserve(options->barcodeCount)); CKR(st.dBarcodeGapGroup.reserve(options->barcodeCount));
    CKR(st.dCounters.reserve(8)); CKR(st.dLongCounters.reserve(4)); CKR(st.dGroupBegin.reserve(2 * (size_t(groups) + 1)));
    CKR(st.hCounters.reserve(8)); CKR(st.hLongCounters.reserve(4));
    if (dataBytes) CKR(cudaMemcpyAsync(st.dData.p, data, dataBytes, cudaMemcpyHostToDevice, s));
    if (recordCount) CKR(cudaMemcpyAsync(st.dRecordOffset.p, recordOffset, recordCount * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (indexCount) CKR(cudaMemcpyAsync(st.dIndex.p, index, indexCount * sizeof(isaac_ext_bin_index_t), cudaMemcpyHostToDevice, s));
    CKR(cudaMemcpyAsync(st.dTls.p, options->barcodeTls, options->barcodeCount * sizeof(isaac_ext_tls_t), cudaMemcpyHostToDevice, s));
    if (options->barcodeGapGroup)
        CKR(cudaMemcpyAsync(st.dBarcodeGapGroup.p, options->barcodeGapGroup, options->barcodeCount * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CKR(cudaMemsetAsync(st.dCounters.p, 0, 8 * sizeof(uint32_t), s));
    CKR(cudaMemsetAsync(st.dLongCounters.p, 0, 4 * sizeof(unsigned long long), s));

    // ---- BinSorter::collectGaps
    CKR(cudaEventRecord(st.ev[0], s));
    auto temp = [&](size_t bytes) { return st.dTemp.reserve(bytes + 16); };
    uint32_t rawGaps = 0;
    if (recordCount)
    {
        countRecordGapsKernel<<<gridFor(ctx, recordCount, 256, 16), 256, 0, s>>>(st.dData.p, dataBytes, st.dRecordOffset.p, recordCount, st.dGapsOfRecord.p,
                                                                                options->barcodeCount, st.dCounters.p + RC_ERRORS);
        ++launches;
        CKR(cudaGetLastError());
        CKR(cudaMemsetAsync(st.dGapsOfRecord.p + recordCount, 0, sizeof(uint32_t), s));
        size_t bytes = 0;
        CKR(cub::DeviceScan::ExclusiveSum(nullptr, bytes, st.dGapsOfRecord.p, st.dGapBegin.p, int(recordCount + 1), s));
        CKR(temp(bytes));
        CKR(cub::DeviceScan::ExclusiveSum(st.dTemp.p, bytes, st.dGapsOfRecord.p, st.dGapBegin.p, int(recordCount + 1), s));
        ++launches;
        CKR(cudaMemcpyAsync(st.hCounters.p, st.dGapBegin.p + recordCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaStreamSynchronize(s));
        rawGaps = st.hCounters.p[0];
    }
    timer.mark("upload + count gaps");
    CKR(st.dGapsRaw.reserve(size_t(rawGaps) + 1)); CKR(st.dGaps.reserve(size_t(rawGaps) + 1)); CKR(st.dDeletions.reserve(size_t(rawGaps) + 1));
    if (rawGaps)
    {
        writeRecordGapsKernel<<<gridFor(ctx, recordCount, 256, 16), 256, 0, s>>>(st.dData.p, st.dRecordOffset.p, recordCount, st.dGapsOfRecord.p, st.dGapBegin.p,
                                                                                options->barcodeGapGroup ? st.dBarcodeGapGroup.p : nullptr,
                                                                                options->barcodeCount, st.dGapsRaw.p);
        ++launches;
        CKR(cudaGetLastError());
        size_t bytes = 0;
        CKR(cub::DeviceMergeSort::SortKeys(nullptr, bytes, st.dGapsRaw.p, int(rawGaps), GapByStart(), s));
        CKR(temp(bytes));
        CKR(cub::DeviceMergeSort::SortKeys(st.dTemp.p, bytes, st.dGapsRaw.p, int(rawGaps), GapByStart(), s));
        CKR(cub::DeviceSelect::Unique(nullptr, bytes, st.dGapsRaw.p, st.dGaps.p, st.dCounters.p + RC_GAPS, int(rawGaps), s));
        CKR(temp(bytes));
        CKR(cub::DeviceSelect::Unique(st.dTemp.p, bytes, st.dGapsRaw.p, st.dGaps.p, st.dCounters.p + RC_GAPS, int(rawGaps), s));
        // the deletions among the unique gaps in gapGroups_ order (remove_copy_if), then by end
        CKR(cudaMemcpyAsync(st.hCounters.p, st.dCounters.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaStreamSynchronize(s));
        const uint32_t uniqueGaps = st.hCounters.p[0];
        CKR(cub::DeviceSelect::If(nullptr, bytes, st.dGaps.p, st.dDeletions.p, st.dCounters.p + RC_DELETIONS, int(uniqueGaps), GapIsDeletion(), s));
        CKR(temp(bytes));
        CKR(cub::DeviceSelect::If(st.dTemp.p, bytes, st.dGaps.p, st.dDeletions.p, st.dCounters.p + RC_DELETIONS, int(uniqueGaps), GapIsDeletion(), s));
        CKR(cudaMemcpyAsync(st.hCounters.p + 1, st.dCounters.p + RC_DELETIONS, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaStreamSynchronize(s));
        const uint32_t deletions = st.hCounters.p[1];
        launches += 6;
        if (deletions)
        {
            CKR(cub::DeviceMergeSort::StableSortKeys(nullptr, bytes, st.dDeletions.p, int(deletions), GapByDeletionEnd(), s));
            CKR(temp(bytes));
            CKR(cub::DeviceMergeSort::StableSortKeys(st.dTemp.p, bytes, st.dDeletions.p, int(deletions), GapByDeletionEnd(), s));
            deletionEndTiesKernel<<<gridFor(ctx, deletions, 256, 8), 256, 0, s>>>(st.dDeletions.p, st.dCounters.p + RC_DELETIONS, st.dCounters.p + RC_TIES);
            launches += 3;
            CKR(cudaGetLastError());
            CKR(cudaMemcpyAsync(st.hCounters.p + 2, st.dCounters.p + RC_TIES, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CKR(cudaStreamSynchronize(s));
            if (st.hCounters.p[2])
            {
                // ties: the reference's own sort call on the reference's input order, group by group
                CKR(st.hGaps.reserve(uniqueGaps)); CKR(st.hDeletions.reserve(deletions));
                CKR(cudaMemcpyAsync(st.hGaps.p, st.dGaps.p, size_t(uniqueGaps) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
                CKR(cudaStreamSynchronize(s));
                size_t n = 0;
                for (uint32_t k = 0; k < uniqueGaps; ++k) if (st.hGaps.p[k].length > 0) st.hDeletions.p[n++] = st.hGaps.p[k];
                for (size_t b = 0; b < n;)
                {
                    size_t e = b;
                    while (e < n && st.hDeletions.p[e].group == st.hDeletions.p[b].group) ++e;
                    std::sort(st.hDeletions.p + b, st.hDeletions.p + e, [](const GapRecord &l, const GapRecord &r) {
                        return GapByDeletionEnd::end(l) < GapByDeletionEnd::end(r); });
                    b = e;
                }
                CKR(cudaMemcpyAsync(st.dDeletions.p, st.hDeletions.p, n * sizeof(GapRecord), cudaMemcpyHostToDevice, s));
            }
        }
    }
    gapGroupBeginKernel<<<1, 256, 0, s>>>(st.dGaps.p, st.dCounters.p + RC_GAPS, std::min(groups, 255u), st.dGroupBegin.p);
    gapGroupBeginKernel<<<1, 256, 0, s>>>(st.dDeletions.p, st.dCounters.p + RC_DELETIONS, std::min(groups, 255u), st.dGroupBegin.p + groups + 1);
    launches += 2;
    CKR(cudaGetLastError());
    if (groups > 255) return fail(ISAAC_EXT_E_UNSUPPORTED, "more than 255 gap groups");
    CKR(cudaEventRecord(st.ev[1], s));
    timer.mark("sort gaps");

    // ---- BinSorter::realignGaps
    CKR(st.dRecordIndex.reserve((dataBytes >> 6) + 2)); CKR(st.dPosition.reserve(indexCount + 1));
    CKR(st.dCigarOffset.reserve(indexCount + 1)); CKR(st.dCigarLength.reserve(indexCount + 1));
    CKR(st.dChangedOffset.reserve(2 * indexCount + 2)); CKR(st.dChangedHeader.reserve((2 * indexCount + 2) * REALIGN_CHANGED_STRIDE));
    if (!out.position) { CKR(st.hPosition.reserve(indexCount + 1)); CKR(st.hCigarOffset.reserve(indexCount + 1)); CKR(st.hCigarLength.reserve(indexCount + 1)); }
    uint64_t *const hPosition = out.position ? out.position : st.hPosition.p;
    uint32_t *const hCigarOffset = out.position ? out.cigarOffset : st.hCigarOffset.p, *const hCigarLength = out.position ? out.cigarLength : st.hCigarLength.p;
    uint64_t poolCapacity = std::max<uint64_t>(st.dCigarPool.capacity, indexCount * 8 + 4096);
    float realignMs = 0.0f;
    for (unsigned attempt = 0;; ++attempt)
    {
        CKR(st.dCigarPool.reserve(poolCapacity));
        CKR(cudaMemsetAsync(st.dRecordIndex.p, 0xFF, ((dataBytes >> 6) + 2) * sizeof(uint32_t), s));
        CKR(cudaMemsetAsync(st.dLongCounters.p, 0, 4 * sizeof(unsigned long long), s));
        RealignBinView v{};
        v.data = st.dData.p; v.dataBytes = dataBytes; v.index = st.dIndex.p; v.indexCount = indexCount; v.recordIndex = st.dRecordIndex.p;
        v.gaps = reinterpret_cast<const isaac_ext_gap_t *>(st.dGaps.p); v.gapGroupBegin = st.dGroupBegin.p;
        v.deletions = reinterpret_cast<const isaac_ext_gap_t *>(st.dDeletions.p); v.deletionGroupBegin = st.dGroupBegin.p + groups + 1;
        v.barcodeGapGroup = options->barcodeGapGroup ? st.dBarcodeGapGroup.p : nullptr; v.barcodeTls = st.dTls.p; v.barcodeCount = options->barcodeCount;
        v.ref = ctx->ref;
        v.binStart = binStart; v.binEnd = binEnd;
        v.vigorous = options->realignGapsVigorously != 0; v.dodgy = options->realignDodgyFragments != 0; v.clipSemialigned = options->clipSemialigned != 0;
        v.mismatchCost = options->mismatchCost; v.gapOpenCost = options->gapOpenCost; v.gapExtendCost = options->gapExtendCost;
        v.position = st.dPosition.p; v.cigarOffset = st.dCigarOffset.p; v.cigarLength = st.dCigarLength.p;
        v.cigarPool = st.dCigarPool.p; v.cigarPoolUsed = st.dLongCounters.p; v.cigarPoolCapacity = poolCapacity;
        v.realignedFragments = st.dLongCounters.p + 1; v.errorFlags = st.dCounters.p + RC_ERRORS;
        v.changedOffset = st.dChangedOffset.p; v.changedHeader = st.dChangedHeader.p; v.changedCount = st.dLongCounters.p + 2;
        CKR(cudaEventRecord(st.ev[2], s));
        if (indexCount)
        {
            recordIndexKernel<<<gridFor(ctx, indexCount, 256, 16), 256, 0, s>>>(st.dData.p, dataBytes, st.dIndex.p, indexCount, st.dRecordIndex.p, st.dCounters.p + RC_ERRORS);
            realignBinKernel<<<unsigned((indexCount + 127) / 128), 128, 0, s>>>(v);
            launches += 2;
            CKR(cudaGetLastError());
        }
        CKR(cudaEventRecord(st.ev[3], s));
        CKR(cudaMemcpyAsync(st.hCounters.p, st.dCounters.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaMemcpyAsync(st.hLongCounters.p, st.dLongCounters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CKR(cudaStreamSynchronize(s));
        CKR(cudaEventElapsedTime(&realignMs, st.ev[2], st.ev[3]));
        timer.mark("realign");
        const uint32_t errors = st.hCounters.p[RC_ERRORS];
        if ((errors & REALIGN_ERROR_POOL) && attempt == 0)
        {
            // the pool was sized for short CIGARs; the counter kept counting, so the need is known: the records go up again (the
            // kernel updates them in place) and the pass is repeated once
            poolCapacity = st.hLongCounters.p[0] + 64;
            CKR(cudaMemcpyAsync(st.dData.p, data, dataBytes, cudaMemcpyHostToDevice, s));
            CKR(cudaMemsetAsync(st.dCounters.p + RC_ERRORS, 0, sizeof(uint32_t), s));
            continue;
        }
        if (errors & REALIGN_ERROR_BOUNDS) return fail(ISAAC_EXT_E_INVALID_ARG, "a record or an index entry lies outside the bin's data");
        if (errors & REALIGN_ERROR_BARCODE) return fail(ISAAC_EXT_E_INVALID_ARG, "a record names a barcode outside the barcode tables");
        if (errors & REALIGN_ERROR_UNSUPPORTED_RECORD) return fail(ISAAC_EXT_E_UNSUPPORTED, "a record has more than 512 bases or a CIGAR of more than 64 operations");
        if (errors & REALIGN_ERROR_OVERLAPS) return fail(ISAAC_EXT_E_UNSUPPORTED, "more than 30 groups of overlapping gaps around one fragment (the reference asserts)");
        if (errors) return fail(ISAAC_EXT_E_INVALID_ARG, "a CIGAR of the bin is malformed (no mapped base, unknown operation or too many operations)");
        break;
    }
    // ---- download
    const uint64_t words = st.hLongCounters.p[0];
    const uint32_t uniqueGaps = st.hCounters.p[RC_GAPS], deletions = st.hCounters.p[RC_DELETIONS];
    if (out.position && words > out.cigarCapacity) return fail(ISAAC_EXT_E_CAPACITY, "the realigned CIGARs do not fit the caller's buffer");
    if (!out.position) CKR(st.hCigarPool.reserve(words + 1));
    uint32_t *const hCigarPool = out.position ? out.cigars : st.hCigarPool.p;
    CKR(st.hGaps.reserve(size_t(uniqueGaps) + 1)); CKR(st.hDeletions.reserve(size_t(deletions) + 1));
    // of the records only the headers a call rewrote travel back: offset + leading bytes, scattered into the caller's data below
    const uint64_t changed = st.hLongCounters.p[2];
    CKR(st.hChangedOffset.reserve(changed + 1)); CKR(st.hChangedHeader.reserve((changed + 1) * REALIGN_CHANGED_STRIDE));
    if (changed)
    {
        CKR(cudaMemcpyAsync(st.hChangedOffset.p, st.dChangedOffset.p, changed * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaMemcpyAsync(st.hChangedHeader.p, st.dChangedHeader.p, changed * REALIGN_CHANGED_STRIDE, cudaMemcpyDeviceToHost, s));
    }
    if (indexCount)
    {
        CKR(cudaMemcpyAsync(hPosition, st.dPosition.p, indexCount * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaMemcpyAsync(hCigarOffset, st.dCigarOffset.p, indexCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CKR(cudaMemcpyAsync(hCigarLength, st.dCigarLength.p, indexCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    if (words) CKR(cudaMemcpyAsync(hCigarPool, st.dCigarPool.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (uniqueGaps) CKR(cudaMemcpyAsync(st.hGaps.p, st.dGaps.p, size_t(uniqueGaps) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
    if (deletions) CKR(cudaMemcpyAsync(st.hDeletions.p, st.dDeletions.p, size_t(deletions) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
    CKR(cudaStreamSynchronize(s));
    for (uint64_t c = 0; c < changed; ++c)
    {
        // the targets are scattered over the whole bin: ask for the lines a few records ahead
        if (c + 16 < changed) __builtin_prefetch(data + st.hChangedOffset.p[c + 16], 1);
        std::memcpy(data + st.hChangedOffset.p[c], st.hChangedHeader.p + c * REALIGN_CHANGED_STRIDE, REALIGN_CHANGED_BYTES);
    }
    timer.mark("download");
    result->position = hPosition; result->cigarOffset = hCigarOffset; result->cigarLength = hCigarLength;
    result->realignedCigars = hCigarPool; result->realignedCigarWords = words; result->realignedFragments = st.hLongCounters.p[1];
    result->gaps = reinterpret_cast<const isaac_ext_gap_t *>(st.hGaps.p); result->deletionsByEnd = reinterpret_cast<const isaac_ext_gap_t *>(st.hDeletions.p);
    result->gapCount = uniqueGaps; result->deletionCount = deletions;
    result->collectMs = 0.0f; result->realignMs = realignMs;
    CKR(cudaEventElapsedTime(&result->collectMs, st.ev[0], st.ev[1]));
    return ISAAC_EXT_OK;
}
} // namespace

extern "C" int isaac_ext_realign_bin(isaac_ext_ctx *ctx, const isaac_ext_realign_options_t *options, uint8_t *data, uint64_t dataBytes,
                                     const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index,
                                     uint64_t indexCount, isaac_ext_realign_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->realign) ctx->realign = new RealignSlots();
    return realignBinOn(ctx, ctx->realign->slot[0], ctx->error, ctx->launches, options, data, dataBytes, recordOffset, recordCount, index, indexCount,
                        RealignOutputs(), result);
}

/// BinSorter::process runs on a pool of threads, one bin each (Build.cpp); here REALIGN_SLOTS slots take the jobs in turn, each on its
/// own stream, so that the upload of one bin runs next to the kernels and the downloads of the others
extern "C" int isaac_ext_realign_bins(isaac_ext_ctx *ctx, isaac_ext_realign_job_t *jobs, uint32_t jobCount)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (jobCount && !jobs) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null jobs");
    for (uint32_t j = 0; j < jobCount; ++j)
        if (!jobs[j].options || (jobs[j].indexCount && (!jobs[j].position || !jobs[j].cigarOffset || !jobs[j].cigarLength)) ||
            (jobs[j].realignedCigarCapacity && !jobs[j].realignedCigars))
            return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "a job without options or without its output arrays");
    if (!ctx->realign) ctx->realign = new RealignSlots();
    std::string errors[REALIGN_SLOTS];
    uint64_t launches[REALIGN_SLOTS] = {};
    std::atomic<uint32_t> next(0);
    auto work = [&](const unsigned slot) {
        for (uint32_t j; (j = next++) < jobCount;)
        {
            isaac_ext_realign_job_t &job = jobs[j];
            RealignOutputs out;
            // a job without index entries still gets non-null output pointers so that the caller's memory is the destination
            static uint64_t nothing64; static uint32_t nothing32;
            out.position = job.position ? job.position : &nothing64; out.cigarOffset = job.cigarOffset ? job.cigarOffset : &nothing32;
            out.cigarLength = job.cigarLength ? job.cigarLength : &nothing32; out.cigars = job.realignedCigars; out.cigarCapacity = job.realignedCigarCapacity;
            isaac_ext_realign_result_t r;
            job.status = realignBinOn(ctx, ctx->realign->slot[slot], errors[slot], launches[slot], job.options, job.data, job.dataBytes, job.recordOffset,
                                      job.recordCount, job.index, job.indexCount, out, &r);
            if (job.status) { job.realignedCigarWords = job.realignedFragments = 0; continue; }
            job.realignedCigarWords = r.realignedCigarWords; job.realignedFragments = r.realignedFragments;
        }
    };
    std::thread others[REALIGN_SLOTS - 1];
    const unsigned helpers = std::min<unsigned>(REALIGN_SLOTS - 1, jobCount > 1 ? jobCount - 1 : 0);
    for (unsigned k = 0; k < helpers; ++k) others[k] = std::thread(work, k + 1);
    work(0u);
    for (unsigned k = 0; k < helpers; ++k) others[k].join();
    int rc = ISAAC_EXT_OK;
    for (unsigned k = 0; k < REALIGN_SLOTS; ++k) ctx->launches += launches[k];
    for (uint32_t j = 0; j < jobCount && !rc; ++j) rc = jobs[j].status;
    if (rc) for (unsigned k = 0; k < REALIGN_SLOTS; ++k) if (!errors[k].empty()) { ctx->error = errors[k]; break; }
    return rc;
}
