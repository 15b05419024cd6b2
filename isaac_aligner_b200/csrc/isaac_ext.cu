// C ABI of the B200 candidate-extension path (include/isaac_ext.h): context, resident data, launches.
// There is no CPU implementation behind any of these entry points.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cub/cub.cuh>
#include "device_buffer.cuh"
#include "kernels.cuh"
#include "kernels2.cuh"
#include "kernels3.cuh"
#include "kernels_avoid.cuh"
#include "kernels_rescue.cuh"
#include "host_build.cuh"
#include "kernels_stats.cuh"
#include "sw_wide.cuh"

using namespace isaac_b200;

namespace
{
thread_local std::string g_createError;
struct E2eState;                       // isaac_ext_e2e.cuh
void releaseE2e(E2eState *state);
struct TemplateState;                  // isaac_ext_tls.cuh
struct TileState;                      // isaac_ext_tile.cuh
} // namespace
void releaseTemplates(TemplateState *state);
void releaseTile(TileState *state);
void tileTakeOverPrefetchedBatch(isaac_ext_ctx *ctx);
struct PackState;                      // isaac_ext_pack.cuh
void releasePack(PackState *state);
struct AsyncState;                     // isaac_ext_async.cuh
void releaseAsync(AsyncState *state);
struct SelectState;                    // isaac_ext_select.cuh
void releaseSelect(SelectState *state);
struct RealignSlots;                   // isaac_ext_realign.cuh
void releaseRealign(RealignSlots *state);

struct isaac_ext_ctx
{
    isaac_ext_config_t cfg;
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;
    // One Smith-Waterman path: the packed 16x2 forward kernel + the trace/score kernel (kernels3.cuh).  Tuning knobs
    // (environment, read once at isaac_ext_create): ISAAC_EXT_SW_CHUNK_WAVES = waves of forward blocks per chunk,
    // ISAAC_EXT_SW_BLOCKS_PER_SM bounds the persistent grid of the explicit-string micro kernel.
    unsigned swBlocksPerSm = 8;
    unsigned swChunkWaves = 2;
    cudaStream_t swStream[2] = {nullptr, nullptr};
    cudaEvent_t swDone[2] = {nullptr, nullptr}, swStart = nullptr;
    DeviceBuffer<uint32_t> swPlanes[2], swEndCells[2];
    unsigned hostThreads = 1;     // config.hostThreads (0 = hardware concurrency)
    uint32_t clusterCount = 0;    // of the resident read set
    PipelineState pipeline;       // buffers of isaac_ext_build_fragments / isaac_ext_rescue_shadows
    E2eState *e2e = nullptr;      // streams and chunk buffers of the *_batch_compact entry points
    TemplateState *templates = nullptr;   // buffers of isaac_ext_template_stats
    TileState *tile = nullptr;            // device-resident tile pipeline: isaac_ext_build_fragments / _rescue_shadows / _build_templates
    PackState *pack = nullptr;            // buffers of isaac_ext_pack_fragments
    AsyncState *async = nullptr;          // the call in flight between isaac_ext_submit_* and isaac_ext_wait
    SelectState *select = nullptr;        // isaac_ext_select_tile
    RealignSlots *realign = nullptr;      // buffers and streams of isaac_ext_realign_bin / _bins
    double logMismatchQ40 = 0.0;  // LOG_MISMATCH_Q40 (Quality.hh:100)

    // score tables (host libm, Quality.cpp:34-66) and parameters
    DeviceBuffer<double> tables;
    ScoreParams sp;

    // resident reference
    DeviceBuffer<uint32_t> refBases2, refNmask;
    DeviceBuffer<uint64_t> refContigOffset, refContigLength;
    std::vector<uint64_t> contigLength;
    ReferenceView ref{};
    bool haveReference = false;

    // resident read set: two slots, the tile the calls work on and the one isaac_ext_prefetch_reads fills meanwhile
    struct ReadSlot
    {
        DeviceBuffer<uint32_t> bases2, nmask;
        DeviceBuffer<uint8_t> quality, bclStage, qualityStrand;
        DeviceBuffer<uint64_t> codes4, strand2;
        DeviceBuffer<uint16_t> masked;
        ReadSetView view{};
        uint32_t clusterCount = 0;
        isaac_ext_reads_t key{};          // what was uploaded (prefetch: what set_reads must be called with to take the slot over)
        bool staged = false;              // filled by isaac_ext_prefetch_reads, not yet taken over
        cudaEvent_t ready = nullptr;
        void release()
        {
            bases2.release(); nmask.release(); quality.release(); bclStage.release(); qualityStrand.release(); codes4.release();
            strand2.release(); masked.release();
            if (ready) cudaEventDestroy(ready);
            ready = nullptr;
        }
    };
    ReadSlot readSlot[2];
    unsigned activeSlot = 0;
    cudaStream_t stageStream = nullptr;   // uploads + decodes the prefetched tile next to the kernels of the current one
    ReadSetView reads{};                  // = readSlot[activeSlot].view
    bool haveReads = false;
    ReadSlot &slot() { return readSlot[activeSlot]; }

    // sequencing adapters (isaac_ext_set_adapters; kernels_adapter.cuh); count == 0: no adapter kernel ever runs
    DeviceBuffer<uint8_t> adapterCodes, adapterReverse;
    DeviceBuffer<int8_t> adapterKmers;
    DeviceBuffer<uint32_t> adapterLength, adapterClipLength;
    AdapterView adapters{};
    DeviceBuffer<uint32_t> dAdapterClip;          // one packed clip word per candidate of the pass in flight
    DeviceBuffer<uint32_t> dSwOwner, dPrepWords;  // --avoid-smith-waterman: owner of the 7-mer table, clip word | skip flag
    DeviceBuffer<AdapterRange> dAdapterRanges;    // one per clipper slot of the tile call in flight

    // staging for the host-pointer entry points
    DeviceBuffer<isaac_ext_candidate_t> dCandidates;
    DeviceBuffer<isaac_ext_fragment_t> dFragments;
    DeviceBuffer<uint32_t> dCigars;
    DeviceBuffer<uint64_t> dMasks;
    DeviceBuffer<uint32_t> tbScratch;
    DeviceBuffer<uint32_t> errorFlag;
    DeviceBuffer<unsigned char> dAscii;
    DeviceBuffer<uint64_t> dOffsets;
    DeviceBuffer<uint32_t> dLengths;

    int fail(int code, const std::string &what) { error = what; return code; }
    int cuda(cudaError_t e, const char *what)
    {
        if (e == cudaSuccess) return ISAAC_EXT_OK;
        error = std::string(what) + ": " + cudaGetErrorString(e);
        return ISAAC_EXT_E_CUDA;
    }
};

#define CK(call) do { const int rc_ = ctx->cuda((call), #call); if (rc_) return rc_; } while (0)

/// a call submitted with isaac_ext_submit_* owns the context's stream, buffers and error text until isaac_ext_wait: every other
/// entry point but the two prefetches (they work on the standby slots and their own stream) refuses to run next to it
bool asyncCallInFlight(const isaac_ext_ctx *ctx);       // isaac_ext_async.cuh
#define REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx) \
    do { if (asyncCallInFlight(ctx)) return ISAAC_EXT_E_UNSUPPORTED; } while (0)

namespace
{

const unsigned SW_BLOCK = 128;

/// grid of a persistent-style launch: a multiple of the SM count, at most 'perSm' blocks per SM
unsigned gridFor(const isaac_ext_ctx *ctx, uint64_t items, unsigned block, unsigned perSm)
{
    const uint64_t needed = (items + block - 1) / block;
    const uint64_t cap = uint64_t(ctx->smCount) * perSm;
    return unsigned(std::max<uint64_t>(1, std::min(needed, cap)));
}

/// The overflow guard of the BandedSmithWaterman constructor (BandedSmithWaterman.cpp:47-53) plus the condition under
/// which the reference's wrapping int16 arithmetic never wraps (sw.cuh).
bool swScoresSupported(int match, int mismatch, int open, int ext, unsigned maxReadLength, std::string &why)
{
    const int init = -32768 + open;
    const int maxScore = std::max(std::max(std::max(std::abs(match), std::abs(mismatch)), std::abs(open)), std::abs(ext));
    if (long(maxReadLength) * maxScore >= std::abs(init))
    {
        why = "BandedSmithWaterman: unsupported read length for these scores: use smaller scores or shorter reads";
        return false;
    }
    if (match < 0 || mismatch > 0 || open < 0 || ext < 0 || ext > open)
    {
        why = "BandedSmithWaterman: scores must satisfy match >= 0 >= mismatch and 0 <= gapExtend <= gapOpen";
        return false;
    }
    return true;
}

int ensureTraceback(isaac_ext_ctx *ctx, unsigned grid, unsigned block, unsigned maxQueryLength)
{
    const size_t words = size_t(grid) * block * SW2_FLAG_WORDS * maxQueryLength;
    return ctx->cuda(ctx->tbScratch.reserve(words), "cudaMalloc(traceback scratch)");
}

int checkErrorFlag(isaac_ext_ctx *ctx)
{
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag)
    {
        cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream);
        return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR did not fit the cigar stride");
    }
    return ISAAC_EXT_OK;
}

} // namespace

extern "C" const char *isaac_ext_version(void) { return "isaac-ext-b200 0.1 (sm_100a)"; }

extern "C" const char *isaac_ext_last_error(const isaac_ext_ctx *ctx)
{
    return ctx ? ctx->error.c_str() : g_createError.c_str();
}

extern "C" uint64_t isaac_ext_launch_count(const isaac_ext_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int isaac_ext_create(const isaac_ext_config_t *config, isaac_ext_ctx **out)
{
    if (!config || !out) { g_createError = "null argument"; return ISAAC_EXT_E_INVALID_ARG; }
    *out = nullptr;
    std::string why;
    if (!swScoresSupported(config->gapMatchScore, config->gapMismatchScore, -config->gapOpenScore, -config->gapExtendScore,
                           config->maxReadLength, why))
    {
        g_createError = why;
        return ISAAC_EXT_E_INVALID_ARG;     // reference: common::InvalidParameterException
    }
    if (!config->maxReadLength)     // (reads longer than FragmentMetadata::maxCycles_ = 1024 are refused by isaac_ext_set_reads)
    {
        g_createError = "maxReadLength must not be 0";
        return ISAAC_EXT_E_INVALID_ARG;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || config->device >= count)
    {
        g_createError = "no usable CUDA device (this library has no CPU fallback)";
        return ISAAC_EXT_E_NO_DEVICE;
    }
    isaac_ext_ctx *ctx = new isaac_ext_ctx();
    ctx->cfg = *config;
    ctx->hostThreads = config->hostThreads ? config->hostThreads : std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = std::getenv("ISAAC_EXT_SW_CHUNK_WAVES")) ctx->swChunkWaves = std::max(1, std::min(64, std::atoi(e)));
    if (const char *e = std::getenv("ISAAC_EXT_SW_BLOCKS_PER_SM")) ctx->swBlocksPerSm = std::max(1, std::min(16, std::atoi(e)));
    ctx->device = config->device;
    int rc = ctx->cuda(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (!rc) rc = ctx->cuda(cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, ctx->device), "cudaDeviceGetAttribute");
    if (!rc) rc = ctx->cuda(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int k = 0; k < 2 && !rc; ++k)
    {
        rc = ctx->cuda(cudaStreamCreateWithFlags(&ctx->swStream[k], cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = ctx->cuda(cudaEventCreateWithFlags(&ctx->swDone[k], cudaEventDisableTiming), "cudaEventCreate");
    }
    if (!rc) rc = ctx->cuda(cudaEventCreateWithFlags(&ctx->swStart, cudaEventDisableTiming), "cudaEventCreate");
    if (!rc) rc = ctx->cuda(ctx->tables.reserve(201), "cudaMalloc(tables)");
    if (!rc) rc = ctx->cuda(ctx->errorFlag.reserve(1), "cudaMalloc(flag)");
    if (!rc) rc = ctx->cuda(cudaMemset(ctx->errorFlag.p, 0, sizeof(uint32_t)), "cudaMemset(flag)");
    if (!rc)
    {
        // Quality::logMatchLookup / logMismatchLookup, computed with the host libm like the reference (Quality.cpp:34-66)
        // entry 200 = 0.0 is what an inserted base adds to logProbability (x + 0.0 == x bit for bit)
        double t[201];
        t[200] = 0.0;
        t[0] = std::log(1.0 - std::pow(10.0, 1.0 / -10.0));
        for (int q = 1; q < 100; ++q) t[q] = std::log(1.0 - std::pow(10.0, double(q) / -10.0));
        t[100] = t[0];
        for (int q = 1; q < 100; ++q) t[100 + q] = std::log(std::pow(10.0, double(q) / -10.0) / 3.0);
        ctx->logMismatchQ40 = t[100 + 40];
        rc = ctx->cuda(cudaMemcpy(ctx->tables.p, t, sizeof(t), cudaMemcpyHostToDevice), "cudaMemcpy(tables)");
    }
    if (rc) { g_createError = ctx->error; delete ctx; return rc; }
    ctx->sp.mismatch = uint32_t(config->gapMatchScore - config->gapMismatchScore);     // AlignerBase.cpp:38-41
    ctx->sp.gapOpen = uint32_t(config->gapMatchScore - config->gapOpenScore);
    ctx->sp.gapExtend = uint32_t(config->gapMatchScore - config->gapExtendScore);
    ctx->sp.maxGapExtend = uint32_t(-config->minGapExtendScore);
    ctx->sp.swMatch = config->gapMatchScore; ctx->sp.swMismatch = config->gapMismatchScore;
    ctx->sp.swOpen = -config->gapOpenScore; ctx->sp.swExtend = -config->gapExtendScore;   // GappedAligner.cpp:41
    ctx->sp.logMatch = ctx->tables.p; ctx->sp.logMismatch = ctx->tables.p + 100;
    *out = ctx;
    return ISAAC_EXT_OK;
}

extern "C" void isaac_ext_destroy(isaac_ext_ctx *ctx)
{
    if (!ctx) return;
    releaseAsync(ctx->async);             // joins a submitted call that was never waited for
    ctx->async = nullptr;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->tables.release(); ctx->refBases2.release(); ctx->refNmask.release(); ctx->refContigOffset.release();
    ctx->refContigLength.release();
    if (ctx->stageStream) { cudaStreamSynchronize(ctx->stageStream); cudaStreamDestroy(ctx->stageStream); }
    ctx->readSlot[0].release(); ctx->readSlot[1].release(); ctx->dCandidates.release(); ctx->dFragments.release();
    ctx->dCigars.release(); ctx->dMasks.release(); ctx->tbScratch.release(); ctx->errorFlag.release();
    ctx->dAscii.release(); ctx->dOffsets.release(); ctx->dLengths.release();
    ctx->adapterCodes.release(); ctx->adapterReverse.release(); ctx->adapterKmers.release(); ctx->adapterLength.release();
    ctx->adapterClipLength.release(); ctx->dAdapterClip.release(); ctx->dAdapterRanges.release();
    ctx->dSwOwner.release(); ctx->dPrepWords.release();
    for (int k = 0; k < 2; ++k)
    {
        if (ctx->swStream[k]) { cudaStreamSynchronize(ctx->swStream[k]); cudaStreamDestroy(ctx->swStream[k]); }
        if (ctx->swDone[k]) cudaEventDestroy(ctx->swDone[k]);
        ctx->swPlanes[k].release(); ctx->swEndCells[k].release();
    }
    if (ctx->swStart) cudaEventDestroy(ctx->swStart);
    ctx->pipeline.release();
    releaseE2e(ctx->e2e);
    releaseTemplates(ctx->templates);
    releaseTile(ctx->tile);
    releasePack(ctx->pack);
    releaseSelect(ctx->select);
    releaseRealign(ctx->realign);
    delete ctx;
}

extern "C" int isaac_ext_set_reference(isaac_ext_ctx *ctx, uint32_t contigCount, const char *const *contigBases,
                                       const uint64_t *contigLengths)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!contigCount || !contigBases || !contigLengths) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "empty reference");
    CK(cudaSetDevice(ctx->device));
    std::vector<uint64_t> offset(contigCount);
    uint64_t total = 0;
    for (uint32_t c = 0; c < contigCount; ++c)
    {
        offset[c] = total;
        total += (contigLengths[c] + 127) / 128 * 128;     // every contig starts on a 128-base boundary
    }
    total += 128;                                          // zero padding so that window reads may run past the end
    CK(ctx->refBases2.reserve(total / 16));
    CK(ctx->refNmask.reserve(total / 32));
    CK(cudaMemsetAsync(ctx->refBases2.p, 0, total / 16 * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->refNmask.p, 0, total / 32 * sizeof(uint32_t), ctx->stream));
    CK(ctx->refContigOffset.reserve(contigCount));
    CK(ctx->refContigLength.reserve(contigCount));
    CK(cudaMemcpyAsync(ctx->refContigOffset.p, offset.data(), contigCount * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->refContigLength.p, contigLengths, contigCount * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    // pack in chunks so that a human-size contig does not need its ASCII form resident at once
    const uint64_t chunk = uint64_t(256) << 20;
    CK(ctx->dAscii.reserve(std::min<uint64_t>(chunk, *std::max_element(contigLengths, contigLengths + contigCount))));
    for (uint32_t c = 0; c < contigCount; ++c)
    {
        for (uint64_t done = 0; done < contigLengths[c]; done += chunk)
        {
            const uint64_t n = std::min(chunk, contigLengths[c] - done);
            CK(cudaMemcpyAsync(ctx->dAscii.p, contigBases[c] + done, n, cudaMemcpyHostToDevice, ctx->stream));
            packReferenceKernel<<<gridFor(ctx, (n + 31) / 32, 256, 16), 256, 0, ctx->stream>>>(
                ctx->dAscii.p, n, offset[c] + done, ctx->refBases2.p, ctx->refNmask.p);
            ++ctx->launches;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(ctx->stream));       // dAscii is reused by the next chunk
        }
    }
    ctx->contigLength.assign(contigLengths, contigLengths + contigCount);
    ctx->ref.bases2 = ctx->refBases2.p; ctx->ref.nmask = ctx->refNmask.p;
    ctx->ref.contigOffset = ctx->refContigOffset.p; ctx->ref.contigLength = ctx->refContigLength.p;
    ctx->ref.contigCount = contigCount;
    ctx->ref.totalBases = total;
    ctx->haveReference = true;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_set_adapters(isaac_ext_ctx *ctx, uint32_t count, const isaac_ext_adapter_t *adapters)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (count && !adapters) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null adapter list");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->adapters = AdapterView{};
    if (!count) return ISAAC_EXT_OK;
    std::vector<uint8_t> codes(size_t(count) * ADAPTER_STRIDE, 0), reverse(count);
    std::vector<int8_t> kmers(size_t(count) * ADAPTER_KMERS, int8_t(-1));            // UNINITIALIZED_POSITION (SequencingAdapter.hh:41)
    std::vector<uint32_t> length(count), clipLength(count);
    for (uint32_t a = 0; a < count; ++a)
    {
        const char *s = adapters[a].sequence;
        const size_t n = s ? std::strlen(s) : 0;
        // SequencingAdapter.cpp:35-38; shorter than one k-mer could never be found
        if (n < ADAPTER_KMER || n >= 127) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "adapter sequence must have 5..126 bases");
        if (adapters[a].clipLength && n > adapters[a].clipLength)
            return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "Clip length cannot be shorter than the adapter sequence");
        for (size_t i = 0; i < n; ++i)
        {
            const char *at = std::strchr("ACGT", s[i]);
            if (!at || !s[i]) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "adapter sequences must be upper-case ACGT");
            codes[size_t(a) * ADAPTER_STRIDE + i] = uint8_t(at - "ACGT");
        }
        // kmerPositions_ (SequencingAdapter.cpp:40-57): first position of every 5-mer, -2 once it repeats
        for (size_t i = 0; i + ADAPTER_KMER <= n; ++i)
        {
            unsigned kmer = 0;
            for (unsigned j = 0; j < ADAPTER_KMER; ++j) kmer = (kmer << 2) | codes[size_t(a) * ADAPTER_STRIDE + i + j];
            int8_t &pos = kmers[size_t(a) * ADAPTER_KMERS + kmer];
            if (pos == -1) pos = int8_t(i); else pos = int8_t(-2);                  // NON_UNIQUE_KMER_POSITION
        }
        length[a] = uint32_t(n); clipLength[a] = adapters[a].clipLength; reverse[a] = adapters[a].reverse ? 1 : 0;
    }
    CK(ctx->adapterCodes.reserve(codes.size())); CK(ctx->adapterReverse.reserve(count)); CK(ctx->adapterKmers.reserve(kmers.size()));
    CK(ctx->adapterLength.reserve(count)); CK(ctx->adapterClipLength.reserve(count));
    CK(cudaMemcpy(ctx->adapterCodes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->adapterReverse.p, reverse.data(), count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->adapterKmers.p, kmers.data(), kmers.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->adapterLength.p, length.data(), count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->adapterClipLength.p, clipLength.data(), count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->adapters = AdapterView{count, ctx->adapterCodes.p, ctx->adapterKmers.p, ctx->adapterLength.p, ctx->adapterClipLength.p,
                                ctx->adapterReverse.p};
    return ISAAC_EXT_OK;
}

/// Read::decodeBcl of a tile into 'slot', enqueued on 'stream' (upload, two decode kernels); the caller synchronises
static int loadReads(isaac_ext_ctx *ctx, isaac_ext_ctx::ReadSlot &slot, const isaac_ext_reads_t *r, cudaStream_t stream)
{
    if (!r || !r->bcl || !r->clusterCount || r->readCount < 1 || r->readCount > 2) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "bad read set");
    const uint32_t len0 = r->readLength[0], len1 = r->readCount > 1 ? r->readLength[1] : 0;
    const uint32_t maxLen = std::max(len0, len1);
    if (!len0 || (r->readCount > 1 && !len1) || maxLen > ISAAC_EXT_MAX_CYCLES || len0 + len1 > ctx->cfg.maxReadLength)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "read lengths exceed config.maxReadLength (flowcell::getMaxTotalReadLength)");
    CK(cudaSetDevice(ctx->device));
    const uint64_t readTotal = uint64_t(r->clusterCount) * r->readCount;
    const uint32_t wordsN = (maxLen + 31) / 32, words2 = wordsN * 2, qualityStride = wordsN * 32;
    const uint64_t bclBytes = uint64_t(r->clusterCount) * (len0 + len1);
    const uint32_t wordsC = (maxLen + 15) / 16 + 2;
    CK(slot.bases2.reserve(readTotal * words2));
    CK(slot.nmask.reserve(readTotal * wordsN));
    CK(slot.quality.reserve(readTotal * qualityStride));
    CK(slot.masked.reserve(readTotal));
    CK(slot.bclStage.reserve(bclBytes));
    CK(slot.codes4.reserve(readTotal * 2 * wordsC));
    CK(slot.strand2.reserve(readTotal * 2 * wordsC));
    CK(slot.qualityStrand.reserve(readTotal * 2 * qualityStride));
    CK(cudaMemcpyAsync(slot.bclStage.p, r->bcl, bclBytes, cudaMemcpyHostToDevice, stream));
    if (r->endCyclesMasked)
        CK(cudaMemcpyAsync(slot.masked.p, r->endCyclesMasked, readTotal * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
    else
        CK(cudaMemsetAsync(slot.masked.p, 0, readTotal * sizeof(uint16_t), stream));
    decodeBclKernel<<<gridFor(ctx, readTotal * wordsN, 256, 16), 256, 0, stream>>>(
        slot.bclStage.p, r->clusterCount, r->readCount, len0, len1, words2, wordsN, qualityStride, slot.bases2.p, slot.nmask.p, slot.quality.p);
    encodeStrandCodesKernel<<<gridFor(ctx, readTotal * 2 * wordsC, 256, 16), 256, 0, stream>>>(
        slot.bclStage.p, r->clusterCount, r->readCount, len0, len1, wordsC, slot.codes4.p, slot.strand2.p, qualityStride, slot.qualityStrand.p);
    ctx->launches += 2;
    CK(cudaGetLastError());
    ReadSetView &v = slot.view;
    v.codes4 = slot.codes4.p; v.strand2 = slot.strand2.p; v.wordsC = wordsC; v.qualityStrand = slot.qualityStrand.p;
    v.bases2 = slot.bases2.p; v.nmask = slot.nmask.p; v.quality = slot.quality.p; v.endCyclesMasked = slot.masked.p;
    v.words2 = words2; v.wordsN = wordsN; v.qualityStride = qualityStride; v.readCount = r->readCount;
    v.readLength[0] = len0; v.readLength[1] = len1; v.firstCycle[0] = r->firstCycle[0]; v.firstCycle[1] = r->firstCycle[1];
    v.readTotal = uint32_t(readTotal);
    slot.clusterCount = r->clusterCount;
    slot.key = *r;
    return ISAAC_EXT_OK;
}

static bool sameReads(const isaac_ext_reads_t &a, const isaac_ext_reads_t &b)
{
    return a.bcl == b.bcl && a.endCyclesMasked == b.endCyclesMasked && a.clusterCount == b.clusterCount && a.readCount == b.readCount &&
           a.readLength[0] == b.readLength[0] && a.readLength[1] == b.readLength[1] && a.firstCycle[0] == b.firstCycle[0] &&
           a.firstCycle[1] == b.firstCycle[1];
}

extern "C" int isaac_ext_prefetch_reads(isaac_ext_ctx *ctx, const isaac_ext_reads_t *r)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->stageStream) CK(cudaStreamCreateWithFlags(&ctx->stageStream, cudaStreamNonBlocking));
    isaac_ext_ctx::ReadSlot &standby = ctx->readSlot[ctx->activeSlot ^ 1u];
    if (!standby.ready) CK(cudaEventCreateWithFlags(&standby.ready, cudaEventDisableTiming));
    standby.staged = false;
    const int rc = loadReads(ctx, standby, r, ctx->stageStream);
    if (rc) return rc;
    CK(cudaEventRecord(standby.ready, ctx->stageStream));
    standby.staged = true;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_set_reads(isaac_ext_ctx *ctx, const isaac_ext_reads_t *r)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!r) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "bad read set");
    isaac_ext_ctx::ReadSlot &standby = ctx->readSlot[ctx->activeSlot ^ 1u];
    if (standby.staged && sameReads(standby.key, *r))
    {
        // the tile was prefetched: wait for its decode (long done when the previous tile took longer than the upload), take it over
        CK(cudaSetDevice(ctx->device));
        CK(cudaEventSynchronize(standby.ready));
        standby.staged = false;
        ctx->activeSlot ^= 1u;
        tileTakeOverPrefetchedBatch(ctx);       // the tile's matches, if they were prefetched too
    }
    else
    {
        const int rc = loadReads(ctx, ctx->slot(), r, ctx->stream);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->reads = ctx->slot().view;
    ctx->clusterCount = ctx->slot().clusterCount;
    ctx->haveReads = true;
    return ISAAC_EXT_OK;
}

static int validateCandidates(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *c)
{
    std::atomic<int> bad(0);
    parallelRanges(ctx->hostThreads, n, [&](unsigned, size_t b, size_t e) {
        int mine = 0;
        for (size_t i = b; i < e; ++i)
        {
            const uint32_t contig = c[i].contigStrand >> 1;
            if (c[i].readId >= ctx->reads.readTotal || contig >= ctx->ref.contigCount) { mine = 1; continue; }
            // FragmentBuilder::addMatch / ShadowAligner never place a read beyond the contig end
            if (c[i].position > int64_t(ctx->contigLength[contig]) || c[i].position < -int64_t(ISAAC_EXT_MAX_CYCLES)) mine = 2;
        }
        if (mine) bad = mine;
    });
    if (bad == 1) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate refers to an unknown read or contig");
    if (bad == 2) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "candidate position outside [-readLength, contigLength]");
    return ISAAC_EXT_OK;
}

/// The micro entry points treat every candidate as its own adapter clipper (checkInitStrand + clip on the same candidate,
/// like testSequencingAdapter.cpp:159-182); *clipOut = nullptr when no adapters are set.
static int adapterSelfClip(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *dCandidates, cudaStream_t stream,
                           const uint32_t **clipOut)
{
    *clipOut = nullptr;
    if (!ctx->adapters.count || !n) return ISAAC_EXT_OK;
    if (n > ctx->dAdapterClip.capacity) CK(cudaStreamSynchronize(stream));      // an earlier pass may still read the old buffer
    CK(ctx->dAdapterClip.reserve(n));
    adapterSelfClipKernel<<<gridFor(ctx, n, 128, 16), 128, 0, stream>>>(ctx->adapters, ctx->ref, ctx->reads, n, dCandidates, ctx->dAdapterClip.p);
    ++ctx->launches;
    *clipOut = ctx->dAdapterClip.p;
    return ctx->cuda(cudaGetLastError(), "adapterSelfClipKernel");
}

/// Clip words of candidates that share clippers: ranges[slot] was filled by adapterInitKernel, slot = dSlotOf[i] or, with
/// dSlotOf == nullptr, readId * 2 + reverse.
static int adapterSlotClip(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *dCandidates, const uint32_t *dSlotOf,
                           cudaStream_t stream, const uint32_t **clipOut)
{
    *clipOut = nullptr;
    if (!ctx->adapters.count || !n) return ISAAC_EXT_OK;
    if (n > ctx->dAdapterClip.capacity) CK(cudaStreamSynchronize(stream));
    CK(ctx->dAdapterClip.reserve(n));
    adapterClipKernel<<<gridFor(ctx, n, 128, 16), 128, 0, stream>>>(ctx->ref, ctx->reads, n, dCandidates, dSlotOf, ctx->dAdapterRanges.p,
                                                                    ctx->dAdapterClip.p);
    ++ctx->launches;
    *clipOut = ctx->dAdapterClip.p;
    return ctx->cuda(cudaGetLastError(), "adapterClipKernel");
}

static int ungappedDevice(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *dCandidates, isaac_ext_fragment_t *dFragmentsOut,
                          uint32_t *dCigarOut, uint64_t *dMismatchMaskOut, cudaStream_t stream, const uint32_t *adapterClip)
{
    if (!n) return ISAAC_EXT_OK;
    ungappedKernel<<<gridFor(ctx, n, 128, 16), 128, 0, stream>>>(ctx->ref, ctx->reads, ctx->sp, n, dCandidates, dFragmentsOut, dCigarOut,
                                                                 dMismatchMaskOut, adapterClip);
    ++ctx->launches;
    return ctx->cuda(cudaGetLastError(), "ungappedKernel");
}

extern "C" int isaac_ext_ungapped_batch_device(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates, void *dFragmentsOut,
                                               void *dCigarOut, void *dMismatchMaskOut, void *cudaStream)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!n) return ISAAC_EXT_OK;
    const uint32_t *clip = nullptr;
    const int rc = adapterSelfClip(ctx, n, static_cast<const isaac_ext_candidate_t *>(dCandidates), cudaStream_t(cudaStream), &clip);
    if (rc) return rc;
    return ungappedDevice(ctx, n, static_cast<const isaac_ext_candidate_t *>(dCandidates), static_cast<isaac_ext_fragment_t *>(dFragmentsOut),
                          static_cast<uint32_t *>(dCigarOut), static_cast<uint64_t *>(dMismatchMaskOut), cudaStream_t(cudaStream), clip);
}

/// Split path (kernels3.cuh): chunks of pairs alternate between two streams so that the forward kernel of one chunk
/// overlaps the trace+score kernel of the previous one; each stream owns one set of direction planes.
static int gappedSplit(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *dCandidates, uint32_t cigarStride,
                       isaac_ext_fragment_t *dFragments, uint32_t *dCigars, uint64_t *dMasks, cudaStream_t user,
                       const uint32_t *adapterClip)
{
    const unsigned maxLength = std::max(ctx->reads.readLength[0], ctx->reads.readLength[1]);
    const size_t rowWords = size_t(SW2_FLAG_WORDS) * maxLength;
    // whole waves of 4 resident forward blocks per SM; planes of one chunk bounded to 2 GiB
    size_t chunkPairs = size_t(ctx->smCount) * 4 * SW_BLOCK * ctx->swChunkWaves;
    const size_t cap = std::max<size_t>(SW_BLOCK, ((size_t(2) << 30) / 4 / rowWords) / SW_BLOCK * SW_BLOCK);
    chunkPairs = std::min(chunkPairs, cap);
    const size_t pairs = (size_t(n) + 1) / 2;
    const size_t stride = std::min(chunkPairs, (pairs + SW_BLOCK - 1) / SW_BLOCK * SW_BLOCK);
    const unsigned buffers = pairs > chunkPairs ? 2 : 1;
    for (unsigned k = 0; k < buffers; ++k)
    {
        CK(ctx->swPlanes[k].reserve(stride * rowWords));
        CK(ctx->swEndCells[k].reserve(stride));
    }
    CK(cudaEventRecord(ctx->swStart, user));
    for (unsigned k = 0; k < buffers; ++k) CK(cudaStreamWaitEvent(ctx->swStream[k], ctx->swStart, 0));
    unsigned k = 0;
    for (size_t first = 0; first < pairs; first += chunkPairs, k ^= 1u)
    {
        const size_t firstCandidate = first * 2;
        const uint32_t count = uint32_t(std::min<size_t>(chunkPairs * 2, n - firstCandidate));
        const uint32_t chunk = (count + 1) / 2;
        // row-relative cell values where the scores leave room for them (sw2.cuh), the plain ones otherwise
        (sw2RowRelativeFits(ctx->sp.swMatch, maxLength) ? swForwardKernel<true> : swForwardKernel<false>)
            <<<(chunk + SW_BLOCK - 1) / SW_BLOCK, SW_BLOCK, 0, ctx->swStream[k]>>>(
            ctx->ref, ctx->reads, ctx->sp, count, dCandidates + firstCandidate, ctx->swPlanes[k].p, uint32_t(stride),
            ctx->swEndCells[k].p, adapterClip ? adapterClip + firstCandidate : nullptr);
        swTraceScoreKernel<<<(count + SW_BLOCK - 1) / SW_BLOCK, SW_BLOCK, 0, ctx->swStream[k]>>>(
            ctx->ref, ctx->reads, ctx->sp, count, uint32_t(firstCandidate), dCandidates + firstCandidate, ctx->swPlanes[k].p,
            uint32_t(stride), ctx->swEndCells[k].p, cigarStride, dFragments + firstCandidate, dCigars + firstCandidate * cigarStride,
            dMasks ? dMasks + firstCandidate * ISAAC_EXT_MASK_WORDS : nullptr, ctx->errorFlag.p,
            adapterClip ? adapterClip + firstCandidate : nullptr);
        ctx->launches += 2;
    }
    for (unsigned b = 0; b < buffers; ++b)
    {
        CK(cudaEventRecord(ctx->swDone[b], ctx->swStream[b]));
        CK(cudaStreamWaitEvent(user, ctx->swDone[b], 0));
    }
    return ctx->cuda(cudaGetLastError(), "swForwardKernel / swTraceScoreKernel");
}

static int gappedDevice(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates, uint32_t cigarStride,
                        void *dFragmentsOut, void *dCigarOut, void *dMismatchMaskOut, void *cudaStream, const uint32_t *adapterClip);

extern "C" int isaac_ext_gapped_batch_device(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates, uint32_t cigarStride,
                                             void *dFragmentsOut, void *dCigarOut, void *dMismatchMaskOut, void *cudaStream)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!n) return ISAAC_EXT_OK;
    const uint32_t *clip = nullptr;
    const int rc = adapterSelfClip(ctx, n, static_cast<const isaac_ext_candidate_t *>(dCandidates), cudaStream_t(cudaStream), &clip);
    if (rc) return rc;
    return gappedDevice(ctx, n, dCandidates, cigarStride, dFragmentsOut, dCigarOut, dMismatchMaskOut, cudaStream, clip);
}

static int gappedDevice(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates, uint32_t cigarStride,
                        void *dFragmentsOut, void *dCigarOut, void *dMismatchMaskOut, void *cudaStream, const uint32_t *adapterClip)
{
    if (!n) return ISAAC_EXT_OK;
    if (ctx->cfg.avoidSmithWaterman)
    {
        // makesSenseToGapAlign of every candidate first (kernels_avoid.cuh); its verdict rides on the clip words
        const isaac_ext_candidate_t *cand = static_cast<const isaac_ext_candidate_t *>(dCandidates);
        const cudaStream_t stream = cudaStream_t(cudaStream);
        if (n > ctx->dSwOwner.capacity) CK(cudaStreamSynchronize(stream));
        CK(ctx->dSwOwner.reserve(n)); CK(ctx->dPrepWords.reserve(n));
        const uint32_t maxLength = std::max(ctx->reads.readLength[0], ctx->reads.readLength[1]);
        swHashOwnerKernel<<<gridFor(ctx, n, 128, 16), 128, 0, stream>>>(ctx->ref, ctx->reads, n, cand, adapterClip, ctx->dSwOwner.p);
        const size_t shared = size_t(AVOID_WARPS) * (maxLength + 3u * maxLength + 32u) * sizeof(uint32_t);
        if (shared > 48 * 1024)
            CK(cudaFuncSetAttribute(avoidSwKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shared)));
        avoidSwKernel<<<gridFor(ctx, (uint64_t(n) + AVOID_WARPS - 1) / AVOID_WARPS * (AVOID_WARPS * 32), AVOID_WARPS * 32, 8), AVOID_WARPS * 32, shared, stream>>>(
            ctx->ref, ctx->reads, n, cand, adapterClip, ctx->dSwOwner.p, maxLength, ctx->dPrepWords.p);
        ctx->launches += 2;
        CK(cudaGetLastError());
        adapterClip = ctx->dPrepWords.p;
    }
    return gappedSplit(ctx, n, static_cast<const isaac_ext_candidate_t *>(dCandidates), cigarStride,
                       static_cast<isaac_ext_fragment_t *>(dFragmentsOut), static_cast<uint32_t *>(dCigarOut),
                       static_cast<uint64_t *>(dMismatchMaskOut), cudaStream_t(cudaStream), adapterClip);
}

static int extendHost(isaac_ext_ctx *ctx, bool gapped, uint32_t n, const isaac_ext_candidate_t *candidates, uint32_t cigarStride,
                      isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *maskOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!n) return ISAAC_EXT_OK;
    if (!candidates || !fragmentsOut || !cigarOut || cigarStride < 3) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    int rc = validateCandidates(ctx, n, candidates);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dCandidates.reserve(n));
    CK(ctx->dFragments.reserve(n));
    CK(ctx->dCigars.reserve(size_t(n) * cigarStride));
    if (maskOut) CK(ctx->dMasks.reserve(size_t(n) * ISAAC_EXT_MASK_WORDS));
    CK(cudaMemcpyAsync(ctx->dCandidates.p, candidates, size_t(n) * sizeof(*candidates), cudaMemcpyHostToDevice, ctx->stream));
    rc = gapped ? isaac_ext_gapped_batch_device(ctx, n, ctx->dCandidates.p, cigarStride, ctx->dFragments.p, ctx->dCigars.p,
                                                maskOut ? ctx->dMasks.p : nullptr, ctx->stream)
                : isaac_ext_ungapped_batch_device(ctx, n, ctx->dCandidates.p, ctx->dFragments.p, ctx->dCigars.p,
                                                  maskOut ? ctx->dMasks.p : nullptr, ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(fragmentsOut, ctx->dFragments.p, size_t(n) * sizeof(*fragmentsOut), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cigarOut, ctx->dCigars.p, size_t(n) * cigarStride * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (maskOut)
        CK(cudaMemcpyAsync(maskOut, ctx->dMasks.p, size_t(n) * ISAAC_EXT_MASK_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    return gapped ? checkErrorFlag(ctx) : ctx->cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
}

extern "C" int isaac_ext_ungapped_batch(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                        isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut)
{
    return extendHost(ctx, false, n, candidates, 3, fragmentsOut, cigarOut, mismatchMaskOut);
}

extern "C" int isaac_ext_gapped_batch(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates, uint32_t cigarStride,
                                      isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut)
{
    return extendHost(ctx, true, n, candidates, cigarStride, fragmentsOut, cigarOut, mismatchMaskOut);
}

extern "C" int isaac_ext_banded_sw_batch(isaac_ext_ctx *ctx, uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                         const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                         int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                         uint32_t cigarStride, uint32_t *cigarOut, uint32_t *cigarLengthOut, uint32_t *offsetOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!n) return ISAAC_EXT_OK;
    if (!queries || !queryOffsets || !queryLengths || !databases || !databaseOffsets || !cigarOut || !cigarLengthOut || !offsetOut || !cigarStride)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    uint32_t maxLen = 0; uint64_t qBytes = 0, dBytes = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (!queryLengths[i]) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "empty query");
        maxLen = std::max(maxLen, queryLengths[i]);
        qBytes = std::max(qBytes, queryOffsets[i] + queryLengths[i]);
        dBytes = std::max(dBytes, databaseOffsets[i] + queryLengths[i] + 15);
    }
    std::string why;
    // assert(querySize <= maxReadLength_) (BandedSmithWaterman.cpp:94) and the constructor's guard
    if (maxLen > ctx->cfg.maxReadLength) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "query longer than config.maxReadLength");
    if (!swScoresSupported(matchScore, mismatchScore, gapOpenScore, gapExtendScore, ctx->cfg.maxReadLength, why))
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, why);
    for (uint64_t i = 0; i < qBytes; ++i)
    {
        const char c = queries[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'n') return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "query alphabet is ACGTn");
    }
    for (uint64_t i = 0; i < dBytes; ++i)
    {
        const char c = databases[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "database alphabet is ACGTN");
    }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dAscii.reserve(qBytes + dBytes));
    CK(ctx->dOffsets.reserve(size_t(n) * 2));
    CK(ctx->dLengths.reserve(size_t(n) * 3));
    CK(ctx->dCigars.reserve(size_t(n) * cigarStride));
    CK(cudaMemcpyAsync(ctx->dAscii.p, queries, qBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dAscii.p + qBytes, databases, dBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dOffsets.p, queryOffsets, size_t(n) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dOffsets.p + n, databaseOffsets, size_t(n) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dLengths.p, queryLengths, size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = gridFor(ctx, (n + 1) / 2, SW_BLOCK, ctx->swBlocksPerSm);      // two alignments per thread
    int rc = ensureTraceback(ctx, grid, SW_BLOCK, maxLen);
    if (rc) return rc;
    const SwScores sw = {matchScore, mismatchScore, gapOpenScore, gapExtendScore, -32768 + gapOpenScore};
    (sw2RowRelativeFits(matchScore, maxLen) ? bandedSwAsciiKernel2<true> : bandedSwAsciiKernel2<false>)<<<grid, SW_BLOCK, 0, ctx->stream>>>(n, ctx->dAscii.p, ctx->dOffsets.p, ctx->dLengths.p, ctx->dAscii.p + qBytes,
                                                                 ctx->dOffsets.p + n, sw, cigarStride, ctx->dCigars.p, ctx->dLengths.p + n,
                                                                 ctx->dLengths.p + 2 * size_t(n), ctx->tbScratch.p, ctx->errorFlag.p);
    ++ctx->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cigarOut, ctx->dCigars.p, size_t(n) * cigarStride * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cigarLengthOut, ctx->dLengths.p + n, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(offsetOut, ctx->dLengths.p + 2 * size_t(n), size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return checkErrorFlag(ctx);
}

/// launch of the warp-wavefront kernel over device-resident strings
static int bandedSwWideLaunch(isaac_ext_ctx *ctx, uint32_t bandWidth, uint32_t n, const unsigned char *dQueries, const uint64_t *dQueryOffsets,
                              const uint32_t *dQueryLengths, const unsigned char *dDatabases, const uint64_t *dDatabaseOffsets,
                              uint32_t maxQueryLength, int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                              uint32_t cigarStride, uint32_t *dCigarOut, uint32_t *dCigarLengthOut, uint32_t *dOffsetOut, cudaStream_t stream)
{
    if (bandWidth != 16 && bandWidth != 32) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "band width must be 16 or 32 (the band lies across the lanes of one warp)");
    std::string why;
    if (maxQueryLength > ctx->cfg.maxReadLength) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "query longer than config.maxReadLength");
    if (!swScoresSupported(matchScore, mismatchScore, gapOpenScore, gapExtendScore, ctx->cfg.maxReadLength, why))
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, why);
    const unsigned perWarp = 32u / bandWidth;
    const size_t shared = size_t(swWideSharedBytes(maxQueryLength, bandWidth)) * SW_WIDE_WARPS * perWarp;
    if (shared > 200 * 1024) return ctx->fail(ISAAC_EXT_E_CAPACITY, "query too long for the direction planes of the wavefront kernel in shared memory");
    const SwScores sw = {matchScore, mismatchScore, gapOpenScore, gapExtendScore, -32768 + gapOpenScore};
    const uint64_t perBlock = uint64_t(SW_WIDE_WARPS) * perWarp;
    // resident CTAs per SM by shared memory (227 KB), at most 8: a grid of whole waves
    const unsigned perSm = unsigned(std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / std::max<size_t>(shared, 1))));
    const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((n + perBlock - 1) / perBlock, uint64_t(ctx->smCount) * perSm)));
    if (bandWidth == 32)
    {
        if (shared > 48 * 1024) CK(cudaFuncSetAttribute(bandedSwWideKernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shared)));
        bandedSwWideKernel<32><<<grid, SW_WIDE_WARPS * 32, shared, stream>>>(n, dQueries, dQueryOffsets, dQueryLengths, dDatabases, dDatabaseOffsets, sw,
                                                                              maxQueryLength, cigarStride, dCigarOut, dCigarLengthOut, dOffsetOut, ctx->errorFlag.p);
    }
    else
    {
        if (shared > 48 * 1024) CK(cudaFuncSetAttribute(bandedSwWideKernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shared)));
        bandedSwWideKernel<16><<<grid, SW_WIDE_WARPS * 32, shared, stream>>>(n, dQueries, dQueryOffsets, dQueryLengths, dDatabases, dDatabaseOffsets, sw,
                                                                              maxQueryLength, cigarStride, dCigarOut, dCigarLengthOut, dOffsetOut, ctx->errorFlag.p);
    }
    ++ctx->launches;
    return ctx->cuda(cudaGetLastError(), "bandedSwWideKernel");
}

extern "C" int isaac_ext_banded_sw_wide_batch_device(isaac_ext_ctx *ctx, uint32_t bandWidth, uint32_t n, const void *dQueries,
                                                     const void *dQueryOffsets, const void *dQueryLengths, const void *dDatabases,
                                                     const void *dDatabaseOffsets, uint32_t maxQueryLength, int matchScore, int mismatchScore,
                                                     int gapOpenScore, int gapExtendScore, uint32_t cigarStride, void *dCigarOut,
                                                     void *dCigarLengthOut, void *dOffsetOut, void *cudaStream)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!n) return ISAAC_EXT_OK;
    if (!dQueries || !dQueryOffsets || !dQueryLengths || !dDatabases || !dDatabaseOffsets || !dCigarOut || !dCigarLengthOut || !dOffsetOut ||
        !cigarStride || !maxQueryLength)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    return bandedSwWideLaunch(ctx, bandWidth, n, static_cast<const unsigned char *>(dQueries), static_cast<const uint64_t *>(dQueryOffsets),
                              static_cast<const uint32_t *>(dQueryLengths), static_cast<const unsigned char *>(dDatabases),
                              static_cast<const uint64_t *>(dDatabaseOffsets), maxQueryLength, matchScore, mismatchScore, gapOpenScore,
                              gapExtendScore, cigarStride, static_cast<uint32_t *>(dCigarOut), static_cast<uint32_t *>(dCigarLengthOut),
                              static_cast<uint32_t *>(dOffsetOut), cudaStream_t(cudaStream));
}

extern "C" int isaac_ext_banded_sw_wide_batch(isaac_ext_ctx *ctx, uint32_t bandWidth, uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                              const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                              int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                              uint32_t cigarStride, uint32_t *cigarOut, uint32_t *cigarLengthOut, uint32_t *offsetOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!n) return ISAAC_EXT_OK;
    if (!queries || !queryOffsets || !queryLengths || !databases || !databaseOffsets || !cigarOut || !cigarLengthOut || !offsetOut || !cigarStride)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null buffer");
    if (bandWidth != 16 && bandWidth != 32) return ctx->fail(ISAAC_EXT_E_UNSUPPORTED, "band width must be 16 or 32 (the band lies across the lanes of one warp)");
    uint32_t maxLen = 0; uint64_t qBytes = 0, dBytes = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (!queryLengths[i]) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "empty query");
        maxLen = std::max(maxLen, queryLengths[i]);
        qBytes = std::max(qBytes, queryOffsets[i] + queryLengths[i]);
        dBytes = std::max(dBytes, databaseOffsets[i] + queryLengths[i] + bandWidth - 1);
    }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dAscii.reserve(qBytes + dBytes));
    CK(ctx->dOffsets.reserve(size_t(n) * 2));
    CK(ctx->dLengths.reserve(size_t(n) * 3));
    CK(ctx->dCigars.reserve(size_t(n) * cigarStride));
    CK(cudaMemcpyAsync(ctx->dAscii.p, queries, qBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dAscii.p + qBytes, databases, dBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dOffsets.p, queryOffsets, size_t(n) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dOffsets.p + n, databaseOffsets, size_t(n) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dLengths.p, queryLengths, size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    const int rc = bandedSwWideLaunch(ctx, bandWidth, n, ctx->dAscii.p, ctx->dOffsets.p, ctx->dLengths.p, ctx->dAscii.p + qBytes, ctx->dOffsets.p + n, maxLen,
                                      matchScore, mismatchScore, gapOpenScore, gapExtendScore, cigarStride, ctx->dCigars.p, ctx->dLengths.p + n,
                                      ctx->dLengths.p + 2 * size_t(n), ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(cigarOut, ctx->dCigars.p, size_t(n) * cigarStride * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cigarLengthOut, ctx->dLengths.p + n, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(offsetOut, ctx->dLengths.p + 2 * size_t(n), size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag)
    {
        cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream);
        if (flag & 32u) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "query alphabet is ACGTn, database alphabet is ACGTN");
        return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR did not fit the cigar stride");
    }
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_measure_int32_peak(isaac_ext_ctx *ctx, int kind, double *opsPerSecond)
{
    if (!ctx || !opsPerSecond || kind < 0 || kind > 2) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 2048; const unsigned grid = ctx->smCount * 8, block = 256;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep)       // first repetition warms up
    {
        CK(cudaEventRecord(e0, ctx->stream));
        if (kind == 0) intPeakKernel<0><<<grid, block, 0, ctx->stream>>>(iters, rep, ctx->errorFlag.p);
        else if (kind == 1) intPeakKernel<1><<<grid, block, 0, ctx->stream>>>(iters, rep, ctx->errorFlag.p);
        else intPeakKernel<2><<<grid, block, 0, ctx->stream>>>(iters, rep, ctx->errorFlag.p);
        ++ctx->launches;
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = double(iters) * 16 * 8 * double(grid) * block * (kind == 2 ? 2 : 1);
        if (rep) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream));
    *opsPerSecond = best;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_tile_stats_device(isaac_ext_ctx *ctx, uint32_t n, const void *dFragments, void *dStats, void *cudaStream)
{
    if (!ctx || !dStats || (n && !dFragments)) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    static_assert(STAT_COUNT == ISAAC_EXT_STATS_COUNTERS, "counter layout");
    if (!n) return ISAAC_EXT_OK;
    tileStatsKernel<<<gridFor(ctx, n, 256, 8), 256, 0, cudaStream_t(cudaStream)>>>(
        n, static_cast<const isaac_ext_fragment_t *>(dFragments), static_cast<unsigned long long *>(dStats));
    ++ctx->launches;
    return ctx->cuda(cudaGetLastError(), "tileStatsKernel");
}

// isaac_ext_build_fragments, isaac_ext_rescue_shadows, isaac_ext_build_templates
#include "isaac_ext_tile.cuh"

// isaac_ext_ungapped_batch_compact, isaac_ext_gapped_batch_compact
#include "isaac_ext_e2e.cuh"
#include "isaac_ext_tls.cuh"
#include "isaac_ext_pack.cuh"
#include "isaac_ext_async.cuh"
#include "isaac_ext_select.cuh"
#include "isaac_ext_realign.cuh"

namespace
{
void releaseE2e(E2eState *state) { if (state) { state->release(); delete state; } }
} // namespace

extern "C" int isaac_ext_ungapped_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                                isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                                uint64_t *cigarWordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->e2e) ctx->e2e = new E2eState();
    E2ePass pass = {false, fragmentsOut, cigarPoolOut, cigarPoolCapacity, cigarWordsOut};
    return extendCompact(ctx, *ctx->e2e, n, candidates, &pass, 1);
}

extern "C" int isaac_ext_gapped_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                              isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                              uint64_t *cigarWordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->e2e) ctx->e2e = new E2eState();
    E2ePass pass = {true, fragmentsOut, cigarPoolOut, cigarPoolCapacity, cigarWordsOut};
    return extendCompact(ctx, *ctx->e2e, n, candidates, &pass, 1);
}

extern "C" int isaac_ext_extend_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                              isaac_ext_fragment_t *ungappedOut, uint32_t *ungappedPoolOut, uint64_t ungappedPoolCapacity,
                                              uint64_t *ungappedWordsOut, isaac_ext_fragment_t *gappedOut, uint32_t *gappedPoolOut,
                                              uint64_t gappedPoolCapacity, uint64_t *gappedWordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->e2e) ctx->e2e = new E2eState();
    E2ePass passes[2] = {{false, ungappedOut, ungappedPoolOut, ungappedPoolCapacity, ungappedWordsOut},
                         {true, gappedOut, gappedPoolOut, gappedPoolCapacity, gappedWordsOut}};
    return extendCompact(ctx, *ctx->e2e, n, candidates, passes, 2);
}

extern "C" int isaac_ext_align_batch_packed(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                            isaac_ext_alignment_t *alignmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                            uint64_t *cigarWordsOut)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    REFUSE_NEXT_TO_A_SUBMITTED_CALL(ctx);
    if (!ctx->e2e) ctx->e2e = new E2eState();
    return alignPacked(ctx, *ctx->e2e, n, candidates, alignmentsOut, cigarPoolOut, cigarPoolCapacity, cigarWordsOut);
}
