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
    void release()
    {
        for (cudaEvent_t &e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
        dData.release(); dTemp.release(); dRecordOffset.release(); dPosition.release(); dIndex.release(); dGapsOfRecord.release();
        dGapBegin.release(); dRecordIndex.release(); dCigarOffset.release(); dCigarLength.release(); dCigarPool.release();
        dGroupBegin.release(); dBarcodeGapGroup.release(); dCounters.release(); dGapsRaw.release(); dGaps.release(); dDeletions.release();
        dTls.release(); dLongCounters.release(); dChangedHeader.release(); dChangedOffset.release(); hChangedOffset.release(); hChangedHeader.release();
        hPosition.release(); hCigarOffset.release(); hCigarLength.release(); hCigarPool.release(); hCounters.release(); hGaps.release();
        hDeletions.release(); hLongCounters.release();
    }
};

void releaseRealign(RealignState *state) { if (state) { state->release(); delete state; } }

namespace
{
// dCounters: 0 unique gaps, 1 deletions, 2 ties among deletion ends, 3 error flags
enum { RC_GAPS = 0, RC_DELETIONS = 1, RC_TIES = 2, RC_ERRORS = 3, RC_WORDS = 4 };
}

extern "C" int isaac_ext_realign_bin(isaac_ext_ctx *ctx, const isaac_ext_realign_options_t *options, uint8_t *data, uint64_t dataBytes,
                                     const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index,
                                     uint64_t indexCount, isaac_ext_realign_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!options || !result || (dataBytes && !data) || (indexCount && !index)) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (!ctx->haveReference) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference first");
    if (!options->barcodeCount || !options->barcodeTls) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "the barcodes' template length statistics are missing");
    if (indexCount >= 0xFFFFFFFFull || dataBytes >= (1ull << 38)) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "bin too large");
    const int64_t binStart = realignP(options->binStart), binEnd = realignP(options->binEnd);
    if (realignContig(binEnd) >= ctx->ref.contigCount || realignContig(binStart) >= ctx->ref.contigCount)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "the bin lies on a contig the resident reference does not have");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->realign) ctx->realign = new RealignState();
    RealignState &st = *ctx->realign;
    for (cudaEvent_t &e : st.ev) if (!e) CK(cudaEventCreate(&e));
    cudaStream_t s = ctx->stream;
    PhaseTimer timer("realign_bin");

    // ---- the records of the bin: the caller's offsets, or the chain of FragmentHeader::getTotalLength walked here
    if (!recordOffset)
    {
        size_t n = 0;
        for (uint64_t p = 0; p < dataBytes; ++n)
        {
            if (p + BIN_HEADER_BYTES > dataBytes) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "the bin's data end inside a record");
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
    CK(st.dData.reserve(dataBytes + 8)); CK(st.dRecordOffset.reserve(recordCount + 1)); CK(st.dIndex.reserve(indexCount + 1));
    CK(st.dGapsOfRecord.reserve(recordCount + 1)); CK(st.dGapBegin.reserve(recordCount + 1));
    CK(st.dTls.reserve(options->barcodeCount)); CK(st.dBarcodeGapGroup.reserve(options->barcodeCount));
    CK(st.dCounters.reserve(8)); CK(st.dLongCounters.reserve(4)); CK(st.dGroupBegin.reserve(2 * (size_t(groups) + 1)));
    CK(st.hCounters.reserve(8)); CK(st.hLongCounters.reserve(4));
    if (dataBytes) CK(cudaMemcpyAsync(st.dData.p, data, dataBytes, cudaMemcpyHostToDevice, s));
    if (recordCount) CK(cudaMemcpyAsync(st.dRecordOffset.p, recordOffset, recordCount * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (indexCount) CK(cudaMemcpyAsync(st.dIndex.p, index, indexCount * sizeof(isaac_ext_bin_index_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(st.dTls.p, options->barcodeTls, options->barcodeCount * sizeof(isaac_ext_tls_t), cudaMemcpyHostToDevice, s));
    if (options->barcodeGapGroup)
        CK(cudaMemcpyAsync(st.dBarcodeGapGroup.p, options->barcodeGapGroup, options->barcodeCount * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(st.dCounters.p, 0, 8 * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(st.dLongCounters.p, 0, 4 * sizeof(unsigned long long), s));

    // ---- BinSorter::collectGaps
    CK(cudaEventRecord(st.ev[0], s));
    auto temp = [&](size_t bytes) { return st.dTemp.reserve(bytes + 16); };
    uint32_t rawGaps = 0;
    if (recordCount)
    {
        countRecordGapsKernel<<<gridFor(ctx, recordCount, 256, 16), 256, 0, s>>>(st.dData.p, dataBytes, st.dRecordOffset.p, recordCount, st.dGapsOfRecord.p,
                                                                                options->barcodeCount, st.dCounters.p + RC_ERRORS);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(st.dGapsOfRecord.p + recordCount, 0, sizeof(uint32_t), s));
        size_t bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, st.dGapsOfRecord.p, st.dGapBegin.p, int(recordCount + 1), s));
        CK(temp(bytes));
        CK(cub::DeviceScan::ExclusiveSum(st.dTemp.p, bytes, st.dGapsOfRecord.p, st.dGapBegin.p, int(recordCount + 1), s));
        ++ctx->launches;
        CK(cudaMemcpyAsync(st.hCounters.p, st.dGapBegin.p + recordCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        rawGaps = st.hCounters.p[0];
    }
    timer.mark("upload + count gaps");
    CK(st.dGapsRaw.reserve(size_t(rawGaps) + 1)); CK(st.dGaps.reserve(size_t(rawGaps) + 1)); CK(st.dDeletions.reserve(size_t(rawGaps) + 1));
    if (rawGaps)
    {
        writeRecordGapsKernel<<<gridFor(ctx, recordCount, 256, 16), 256, 0, s>>>(st.dData.p, st.dRecordOffset.p, recordCount, st.dGapsOfRecord.p, st.dGapBegin.p,
                                                                                options->barcodeGapGroup ? st.dBarcodeGapGroup.p : nullptr,
                                                                                options->barcodeCount, st.dGapsRaw.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        size_t bytes = 0;
        CK(cub::DeviceMergeSort::SortKeys(nullptr, bytes, st.dGapsRaw.p, int(rawGaps), GapByStart(), s));
        CK(temp(bytes));
        CK(cub::DeviceMergeSort::SortKeys(st.dTemp.p, bytes, st.dGapsRaw.p, int(rawGaps), GapByStart(), s));
        CK(cub::DeviceSelect::Unique(nullptr, bytes, st.dGapsRaw.p, st.dGaps.p, st.dCounters.p + RC_GAPS, int(rawGaps), s));
        CK(temp(bytes));
        CK(cub::DeviceSelect::Unique(st.dTemp.p, bytes, st.dGapsRaw.p, st.dGaps.p, st.dCounters.p + RC_GAPS, int(rawGaps), s));
        // the deletions among the unique gaps in gapGroups_ order (remove_copy_if), then by end
        CK(cudaMemcpyAsync(st.hCounters.p, st.dCounters.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const uint32_t uniqueGaps = st.hCounters.p[0];
        CK(cub::DeviceSelect::If(nullptr, bytes, st.dGaps.p, st.dDeletions.p, st.dCounters.p + RC_DELETIONS, int(uniqueGaps), GapIsDeletion(), s));
        CK(temp(bytes));
        CK(cub::DeviceSelect::If(st.dTemp.p, bytes, st.dGaps.p, st.dDeletions.p, st.dCounters.p + RC_DELETIONS, int(uniqueGaps), GapIsDeletion(), s));
        CK(cudaMemcpyAsync(st.hCounters.p + 1, st.dCounters.p + RC_DELETIONS, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const uint32_t deletions = st.hCounters.p[1];
        ctx->launches += 6;
        if (deletions)
        {
            CK(cub::DeviceMergeSort::StableSortKeys(nullptr, bytes, st.dDeletions.p, int(deletions), GapByDeletionEnd(), s));
            CK(temp(bytes));
            CK(cub::DeviceMergeSort::StableSortKeys(st.dTemp.p, bytes, st.dDeletions.p, int(deletions), GapByDeletionEnd(), s));
            deletionEndTiesKernel<<<gridFor(ctx, deletions, 256, 8), 256, 0, s>>>(st.dDeletions.p, st.dCounters.p + RC_DELETIONS, st.dCounters.p + RC_TIES);
            ctx->launches += 3;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(st.hCounters.p + 2, st.dCounters.p + RC_TIES, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (st.hCounters.p[2])
            {
                // ties: the reference's own sort call on the reference's input order, group by group
                CK(st.hGaps.reserve(uniqueGaps)); CK(st.hDeletions.reserve(deletions));
                CK(cudaMemcpyAsync(st.hGaps.p, st.dGaps.p, size_t(uniqueGaps) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
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
                CK(cudaMemcpyAsync(st.dDeletions.p, st.hDeletions.p, n * sizeof(GapRecord), cudaMemcpyHostToDevice, s));
            }
        }
    }
    gapGroupBeginKernel<<<1, 256, 0, s>>>(st.dGaps.p, st.dCounters.p + RC_GAPS, std::min(groups, 255u), st.dGroupBegin.p);
    gapGroupBeginKernel<<<1, 256, 0, s>>>(st.dDeletions.p, st.dCounters.p + RC_DELETIONS, std::min(groups, 255u), st.dGroupBegin.p + groups + 1);
    ctx->launches += 2;
    CK(cudaGetLastError());
    if (groups > 255) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "more than 255 gap groups");
    CK(cudaEventRecord(st.ev[1], s));
    timer.mark("sort gaps");

    // ---- BinSorter::realignGaps
    CK(st.dRecordIndex.reserve((dataBytes >> 6) + 2)); CK(st.dPosition.reserve(indexCount + 1));
    CK(st.dCigarOffset.reserve(indexCount + 1)); CK(st.dCigarLength.reserve(indexCount + 1));
    CK(st.dChangedOffset.reserve(2 * indexCount + 2)); CK(st.dChangedHeader.reserve((2 * indexCount + 2) * REALIGN_CHANGED_STRIDE));
    CK(st.hPosition.reserve(indexCount + 1)); CK(st.hCigarOffset.reserve(indexCount + 1)); CK(st.hCigarLength.reserve(indexCount + 1));
    uint64_t poolCapacity = std::max<uint64_t>(st.dCigarPool.capacity, indexCount * 8 + 4096);
    float realignMs = 0.0f;
    for (unsigned attempt = 0;; ++attempt)
    {
        CK(st.dCigarPool.reserve(poolCapacity));
        CK(cudaMemsetAsync(st.dRecordIndex.p, 0xFF, ((dataBytes >> 6) + 2) * sizeof(uint32_t), s));
        CK(cudaMemsetAsync(st.dLongCounters.p, 0, 4 * sizeof(unsigned long long), s));
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
        CK(cudaEventRecord(st.ev[2], s));
        if (indexCount)
        {
            recordIndexKernel<<<gridFor(ctx, indexCount, 256, 16), 256, 0, s>>>(st.dData.p, dataBytes, st.dIndex.p, indexCount, st.dRecordIndex.p, st.dCounters.p + RC_ERRORS);
            realignBinKernel<<<unsigned((indexCount + 127) / 128), 128, 0, s>>>(v);
            ctx->launches += 2;
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(st.ev[3], s));
        CK(cudaMemcpyAsync(st.hCounters.p, st.dCounters.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(st.hLongCounters.p, st.dLongCounters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        const int rcSync = ctx->cuda(cudaStreamSynchronize(s), "realignBinKernel");
        if (rcSync) return rcSync;
        CK(cudaEventElapsedTime(&realignMs, st.ev[2], st.ev[3]));
        timer.mark("realign");
        const uint32_t errors = st.hCounters.p[RC_ERRORS];
        if ((errors & REALIGN_ERROR_POOL) && attempt == 0)
        {
            // the pool was sized for short CIGARs; the counter kept counting, so the need is known: the records go up again (the
            // kernel updates them in place) and the pass is repeated once
            poolCapacity = st.hLongCounters.p[0] + 64;
            CK(cudaMemcpyAsync(st.dData.p, data, dataBytes, cudaMemcpyHostToDevice, s));
            CK(cudaMemsetAsync(st.dCounters.p + RC_ERRORS, 0, sizeof(uint32_t), s));
            continue;
        }
        if (errors & REALIGN_ERROR_BOUNDS) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "a record or an index entry lies outside the bin's data");
        if (errors & REALIGN_ERROR_BARCODE) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "a record names a barcode outside the barcode tables");
        if (errors & REALIGN_ERROR_UNSUPPORTED_RECORD) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "a record has more than 512 bases or a CIGAR of more than 64 operations");
        if (errors & REALIGN_ERROR_OVERLAPS) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "more than 30 groups of overlapping gaps around one fragment (the reference asserts)");
        if (errors) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "a CIGAR of the bin is malformed (no mapped base, unknown operation or too many operations)");
        break;
    }
    // ---- download
    const uint64_t words = st.hLongCounters.p[0];
    const uint32_t uniqueGaps = st.hCounters.p[RC_GAPS], deletions = st.hCounters.p[RC_DELETIONS];
    CK(st.hCigarPool.reserve(words + 1)); CK(st.hGaps.reserve(size_t(uniqueGaps) + 1)); CK(st.hDeletions.reserve(size_t(deletions) + 1));
    // of the records only the headers a call rewrote travel back: offset + leading bytes, scattered into the caller's data below
    const uint64_t changed = st.hLongCounters.p[2];
    CK(st.hChangedOffset.reserve(changed + 1)); CK(st.hChangedHeader.reserve((changed + 1) * REALIGN_CHANGED_STRIDE));
    if (changed)
    {
        CK(cudaMemcpyAsync(st.hChangedOffset.p, st.dChangedOffset.p, changed * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(st.hChangedHeader.p, st.dChangedHeader.p, changed * REALIGN_CHANGED_STRIDE, cudaMemcpyDeviceToHost, s));
    }
    if (indexCount)
    {
        CK(cudaMemcpyAsync(st.hPosition.p, st.dPosition.p, indexCount * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(st.hCigarOffset.p, st.dCigarOffset.p, indexCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(st.hCigarLength.p, st.dCigarLength.p, indexCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    if (words) CK(cudaMemcpyAsync(st.hCigarPool.p, st.dCigarPool.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (uniqueGaps) CK(cudaMemcpyAsync(st.hGaps.p, st.dGaps.p, size_t(uniqueGaps) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
    if (deletions) CK(cudaMemcpyAsync(st.hDeletions.p, st.dDeletions.p, size_t(deletions) * sizeof(GapRecord), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint64_t c = 0; c < changed; ++c)
    {
        // the targets are scattered over the whole bin: ask for the lines a few records ahead
        if (c + 16 < changed) __builtin_prefetch(data + st.hChangedOffset.p[c + 16], 1);
        std::memcpy(data + st.hChangedOffset.p[c], st.hChangedHeader.p + c * REALIGN_CHANGED_STRIDE, REALIGN_CHANGED_BYTES);
    }
    timer.mark("download");
    result->position = st.hPosition.p; result->cigarOffset = st.hCigarOffset.p; result->cigarLength = st.hCigarLength.p;
    result->realignedCigars = st.hCigarPool.p; result->realignedCigarWords = words; result->realignedFragments = st.hLongCounters.p[1];
    result->gaps = reinterpret_cast<const isaac_ext_gap_t *>(st.hGaps.p); result->deletionsByEnd = reinterpret_cast<const isaac_ext_gap_t *>(st.hDeletions.p);
    result->gapCount = uniqueGaps; result->deletionCount = deletions;
    result->collectMs = 0.0f; result->realignMs = realignMs;
    CK(cudaEventElapsedTime(&result->collectMs, st.ev[0], st.ev[1]));
    return ISAAC_EXT_OK;
}
