/*
 * ref_capi.cpp -- flat C interface (oracle_api.h) over the reference's OWN classes.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This file contains no alignment logic: it builds the
 * reference's value types (reference::Contig, alignment::Cluster, flowcell::ReadMetadataList, ...) from the
 * flat batch structs, calls the unmodified reference code compiled from /root/reference/src/c++/lib/alignment
 * (see oracle/Makefile) and flattens the FragmentMetadata results.  It is linked into oracle/_ref/libisaac_ref.so,
 * which is git-ignored and rebuilt wherever /root/reference is mounted.
 */
#include <mutex>
#include <thread>
#include <vector>
#include <memory>
#include <cstring>

#include "alignment/BandedSmithWaterman.hh"
#include "alignment/matchSelector/TileStats.hh"
#include "alignment/FragmentBuilder.hh"
#include "alignment/ShadowAligner.hh"
#include "alignment/TemplateBuilder.hh"
#include "alignment/RestOfGenomeCorrection.hh"
#include "alignment/Quality.hh"
#include "alignment/matchSelector/SemialignedEndsClipper.hh"
#include "alignment/matchSelector/OverlappingEndsClipper.hh"
#include "alignment/Cluster.hh"
#include "alignment/fragmentBuilder/UngappedAligner.hh"
#include "alignment/fragmentBuilder/GappedAligner.hh"
#include "alignment/matchSelector/TileBarcodeStats.hh"
#include "alignment/matchSelector/FragmentMetadataTileStatsAdapter.hh"
#include "alignment/matchSelector/BamTemplateTileStatsAdapter.hh"
#include "alignment/matchSelector/FragmentSequencingAdapterClipper.hh"
#include "reference/Contig.hh"
#include "oligo/Nucleotides.hh"
#include "io/Fragment.hh"

#include "oracle_api.h"

using namespace isaac;

namespace
{

template <class F> void parallelFor(uint32_t n, uint32_t threads, F f)
{
    if (threads <= 1 || n < 2 * threads) { f(0, 0u, n); return; }
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t)
    {
        const uint32_t b = uint64_t(n) * t / threads, e = uint64_t(n) * (t + 1) / threads;
        pool.emplace_back([=]() { f(t, b, e); });
    }
    for (std::thread &th : pool) th.join();
}

/// reference::Contig copies of the caller's genome, cached across calls (keyed on the caller's pointers) so that timing
/// a batch does not include copying a human-size genome every time
const std::vector<reference::Contig> &makeContigs(const oracle_genome_t *g)
{
    static std::mutex mutex;
    static std::vector<reference::Contig> cached;
    static std::vector<std::pair<const char *, uint64_t> > key;
    std::lock_guard<std::mutex> lock(mutex);
    std::vector<std::pair<const char *, uint64_t> > want;
    for (uint32_t c = 0; c < g->contigCount; ++c)
    {
        // the caller may reuse an address for another genome: fold a content fingerprint into the key (every byte up
        // to 64 MB, a strided sample beyond)
        const char *b = g->contigBases[c];
        const uint64_t n = g->contigLengths[c], step = n > (64u << 20) ? n / (1u << 20) : 1;
        uint64_t h = 1469598103934665603ull;
        for (uint64_t i = 0; i < n; i += step) h = (h ^ uint8_t(b[i])) * 1099511628211ull;
        want.push_back(std::make_pair(b, n ^ (h << 20)));
    }
    if (want != key)
    {
        cached.clear();
        for (uint32_t c = 0; c < g->contigCount; ++c)
        {
            cached.push_back(reference::Contig(c, "c" + std::to_string(c)));
            cached.back().forward_.assign(g->contigBases[c], g->contigBases[c] + g->contigLengths[c]);
        }
        key = want;
    }
    return cached;
}

} // namespace
const std::vector<reference::Contig> &oracleContigs(const oracle_genome_t *g) { return makeContigs(g); }   // ref_capi_realign.cpp
namespace
{

flowcell::ReadMetadataList makeReadMetadata(const isaac_ext_reads_t *r)
{
    std::vector<flowcell::ReadMetadata> v;
    unsigned offset = 0;
    for (uint32_t i = 0; i < r->readCount; ++i)
    {
        v.push_back(flowcell::ReadMetadata(r->firstCycle[i], r->firstCycle[i] + r->readLength[i] - 1, i, offset));
        offset += r->readLength[i];
    }
    return flowcell::ReadMetadataList(v);
}

/// holds the BCL bytes of one cluster as the std::vector<char> Cluster::init wants
struct ClusterHolder
{
    std::vector<char> bcl;
    alignment::Cluster cluster;
    ClusterHolder(unsigned maxReadLength) : cluster(maxReadLength) {}
    void load(const isaac_ext_reads_t *r, const flowcell::ReadMetadataList &rml, uint32_t clusterId, bool pf = true)
    {
        const uint32_t total = r->readLength[0] + (r->readCount > 1 ? r->readLength[1] : 0);
        bcl.assign(r->bcl + size_t(clusterId) * total, r->bcl + size_t(clusterId + 1) * total);
        cluster.init(rml, bcl.begin(), 0, clusterId, alignment::ClusterXy(0, 0), pf, 0);
        for (uint32_t i = 0; i < r->readCount; ++i)
        {
            cluster[i].maskCyclesFromEnd(r->endCyclesMasked ? r->endCyclesMasked[size_t(clusterId) * r->readCount + i] : 0);
        }
    }
};

void flatten(const alignment::FragmentMetadata &f, uint32_t readId, uint32_t cigarOffset, unsigned matchCount,
             const flowcell::ReadMetadataList &rml, isaac_ext_fragment_t &o, uint32_t *cigarOut, uint32_t cigarStride,
             uint64_t *maskOut)
{
    std::memset(&o, 0, sizeof(o));
    o.position = f.position;
    o.logProbability = f.logProbability;
    o.contigId = f.contigId;
    o.readId = readId;
    o.cigarOffset = cigarOffset;
    o.smithWatermanScore = f.smithWatermanScore;
    o.observedLength = f.observedLength;
    o.mismatchCount = f.mismatchCount;
    o.matchesInARow = f.matchesInARow;
    o.gapCount = f.gapCount;
    o.editDistance = f.editDistance;
    o.uniqueSeedCount = f.uniqueSeedCount;
    o.repeatSeedsCount = f.repeatSeedsCount;
    o.nonUniqueSeedOffsetFirst = uint16_t(std::min<unsigned>(f.nonUniqueSeedOffsets.first, 0xFFFF));
    o.nonUniqueSeedOffsetSecond = uint16_t(f.nonUniqueSeedOffsets.second);
    o.firstSeedIndex = int16_t(f.firstSeedIndex);
    o.lowClipped = f.lowClipped;
    o.highClipped = f.highClipped;
    o.cigarLength = f.cigarLength;
    o.reverse = f.reverse;
    o.readIndex = f.readIndex;
    o.matchCount = matchCount;
    if (f.cigarLength && f.cigarBuffer && cigarOut)
    {
        ISAAC_ASSERT_MSG(f.cigarLength <= cigarStride, "cigar stride too small");
        std::copy(f.cigarBuffer->begin() + f.cigarOffset, f.cigarBuffer->begin() + f.cigarOffset + f.cigarLength, cigarOut);
    }
    if (maskOut)
    {
        std::fill(maskOut, maskOut + ISAAC_EXT_MASK_WORDS, 0UL);
        const unsigned firstCycle = rml[f.readIndex].getFirstCycle(), lastCycle = rml[f.readIndex].getLastCycle();
        for (const unsigned short *c = f.getMismatchCyclesBegin(); c != f.getMismatchCyclesEnd(); ++c)
        {
            const unsigned i = f.reverse ? lastCycle - *c : *c - firstCycle;
            maskOut[i / 64] |= 1UL << (i % 64);
        }
    }
}

} // namespace

extern "C" const char *oracle_kind(void) { return "reference"; }

extern "C" int oracle_banded_sw_batch(uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                      const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                      int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                      uint32_t maxReadLength, uint32_t cigarStride, uint32_t *cigarOut,
                                      uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads)
{
    try
    {
        parallelFor(n, threads, [&](uint32_t, uint32_t b, uint32_t e) {
            const alignment::BandedSmithWaterman sw(matchScore, mismatchScore, gapOpenScore, gapExtendScore, maxReadLength);
            std::vector<char> query, database;
            alignment::Cigar cigar;
            for (uint32_t i = b; i < e; ++i)
            {
                query.assign(queries + queryOffsets[i], queries + queryOffsets[i] + queryLengths[i]);
                database.assign(databases + databaseOffsets[i], databases + databaseOffsets[i] + queryLengths[i] + 15);
                cigar.clear();
                offsetOut[i] = sw.align(query, database.begin(), database.end(), cigar);
                cigarLengthOut[i] = cigar.size();
                std::copy(cigar.begin(), cigar.begin() + std::min<size_t>(cigar.size(), cigarStride), cigarOut + size_t(i) * cigarStride);
            }
        });
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
    return ISAAC_EXT_OK;
}

/// the adapter list of oracle_set_adapters: what isaac-align builds from --default-adapters (process-wide, test harness only)
static alignment::matchSelector::SequencingAdapterList &currentAdapters()
{
    static alignment::matchSelector::SequencingAdapterList adapters;
    return adapters;
}

extern "C" int oracle_set_adapters(uint32_t count, const isaac_ext_adapter_t *adapters)
{
    alignment::matchSelector::SequencingAdapterList &list = currentAdapters();
    list.clear();
    for (uint32_t a = 0; a < count; ++a)
        list.push_back(alignment::matchSelector::SequencingAdapter(
            flowcell::SequencingAdapterMetadata(adapters[a].sequence, adapters[a].reverse != 0, adapters[a].clipLength)));
    return ISAAC_EXT_OK;
}

static int extendBatch(const bool gapped, const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                       const isaac_ext_config_t *cfg, uint32_t n, const isaac_ext_candidate_t *candidates,
                       uint32_t cigarStride, isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut,
                       uint64_t *mismatchMaskOut, uint32_t threads)
{
    try
    {
        const std::vector<reference::Contig> &contigs = makeContigs(genome);
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
        const alignment::matchSelector::SequencingAdapterList &adapterList = currentAdapters();
        const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
        parallelFor(n, threads, [&](uint32_t, uint32_t b, uint32_t e) {
            const alignment::fragmentBuilder::UngappedAligner ungapped(
                cfg->gapMatchScore, cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore, cfg->minGapExtendScore);
            alignment::fragmentBuilder::GappedAligner gappedAligner(
                layouts, cfg->avoidSmithWaterman != 0, cfg->gapMatchScore, cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore, cfg->minGapExtendScore);
            ClusterHolder holder(maxReadLength);
            uint32_t loaded = -1U;
            alignment::Cigar cigar;
            for (uint32_t i = b; i < e; ++i)
            {
                const isaac_ext_candidate_t &c = candidates[i];
                const uint32_t clusterId = c.readId / reads->readCount, readIndex = c.readId % reads->readCount;
                if (loaded != clusterId) { holder.load(reads, rml, clusterId); loaded = clusterId; }
                cigar.clear();
                alignment::FragmentMetadata fragment(&holder.cluster, &cigar, readIndex);
                fragment.reverse = (c.contigStrand & 1);
                fragment.contigId = (c.contigStrand >> 1);
                fragment.position = c.position;
                alignment::matchSelector::FragmentSequencingAdapterClipper clipper(adapterList);
                clipper.checkInitStrand(fragment, contigs[(c.contigStrand >> 1)]);
                unsigned matchCount = ungapped.alignUngapped(fragment, cigar, rml, clipper, contigs[(c.contigStrand >> 1)]);
                // the reference only gap-aligns fragments whose ungapped alignment kept at least one match
                // (FragmentBuilder.cpp:179 drops the others first, ShadowAligner.cpp:223-226 never lists them)
                if (gapped && matchCount)
                {
                    alignment::FragmentMetadata tmp = fragment;
                    matchCount = gappedAligner.alignGapped(tmp, cigar, rml, clipper, contigs[(c.contigStrand >> 1)]);
                    fragment = tmp;
                }
                flatten(fragment, c.readId, i * cigarStride, matchCount, rml, fragmentsOut[i],
                        cigarOut ? cigarOut + size_t(i) * cigarStride : 0, cigarStride,
                        mismatchMaskOut ? mismatchMaskOut + size_t(i) * ISAAC_EXT_MASK_WORDS : 0);
            }
        });
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
    return ISAAC_EXT_OK;
}

extern "C" int oracle_ungapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                     const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                                     isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut,
                                     uint32_t threads)
{
    return extendBatch(false, genome, reads, config, n, candidates, 3, fragmentsOut, cigarOut, mismatchMaskOut, threads);
}

extern "C" int oracle_gapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                   const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                                   uint32_t cigarStride, isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut,
                                   uint64_t *mismatchMaskOut, uint32_t threads)
{
    return extendBatch(true, genome, reads, config, n, candidates, cigarStride, fragmentsOut, cigarOut, mismatchMaskOut, threads);
}

/* ------------------------------------------------------------------------------------------------------------------
 * FragmentBuilder::build and ShadowAligner::rescueShadow over flat batches (glue only, the logic is the reference's)
 * ------------------------------------------------------------------------------------------------------------------ */
namespace
{

struct FlatOut
{
    std::vector<isaac_ext_fragment_t> fragments;
    std::vector<uint32_t> cigars;
    void add(const alignment::FragmentMetadata &f, uint32_t readId, const flowcell::ReadMetadataList &rml)
    {
        isaac_ext_fragment_t o;
        const uint32_t offset = cigars.size();
        if (f.cigarLength && f.cigarBuffer)
            cigars.insert(cigars.end(), f.cigarBuffer->begin() + f.cigarOffset, f.cigarBuffer->begin() + f.cigarOffset + f.cigarLength);
        flatten(f, readId, offset, 0, rml, o, 0, 0, 0);
        fragments.push_back(o);
    }
};

int concatenate(const std::vector<FlatOut> &parts, uint64_t fragmentCapacity, isaac_ext_fragment_t *fragmentsOut,
                uint64_t cigarCapacity, uint32_t *cigarsOut, uint64_t *fragmentCount, uint64_t *cigarWords)
{
    uint64_t nf = 0, nc = 0;
    for (const FlatOut &p : parts) { nf += p.fragments.size(); nc += p.cigars.size(); }
    *fragmentCount = nf; *cigarWords = nc;
    if (nf > fragmentCapacity || nc > cigarCapacity) return ISAAC_EXT_E_CAPACITY;
    nf = 0; nc = 0;
    for (const FlatOut &p : parts)
    {
        for (isaac_ext_fragment_t f : p.fragments) { f.cigarOffset += nc; fragmentsOut[nf++] = f; }
        std::copy(p.cigars.begin(), p.cigars.end(), cigarsOut + nc);
        nc += p.cigars.size();
    }
    return ISAAC_EXT_OK;
}

} // namespace

extern "C" int oracle_build_fragments(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                      const isaac_ext_build_batch_t *batch, uint64_t fragmentCapacity, isaac_ext_fragment_t *fragmentsOut,
                                      uint64_t *readFragmentBegin, uint64_t cigarCapacity, uint32_t *cigarsOut, uint8_t *builtOut,
                                      uint64_t *fragmentCount, uint64_t *cigarWords, uint32_t threads)
{
    try
    {
        const std::vector<reference::Contig> &contigs = makeContigs(genome);
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
        const alignment::matchSelector::SequencingAdapterList &adapterList = currentAdapters();
        alignment::SeedMetadataList seeds;
        for (uint32_t s = 0; s < batch->seedCount; ++s)
            seeds.push_back(alignment::SeedMetadata(batch->seeds[s].offset, batch->seeds[s].length, batch->seeds[s].readIndex, s));
        const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
        const uint32_t n = reads->clusterCount, rc = reads->readCount;
        if (threads < 1) threads = 1;
        if (n < 2 * threads) threads = 1;
        std::vector<FlatOut> parts(threads);
        std::vector<uint64_t> counts(size_t(n) * rc, 0);
        parallelFor(n, threads, [&](uint32_t t, uint32_t b, uint32_t e) {
            alignment::FragmentBuilder builder(layouts, cfg->repeatThreshold, cfg->maxSeedsPerRead, cfg->gappedMismatchesMax,
                                               cfg->avoidSmithWaterman, cfg->gapMatchScore, cfg->gapMismatchScore,
                                               cfg->gapOpenScore, cfg->gapExtendScore, cfg->minGapExtendScore, cfg->semialignedGapLimit);
            ClusterHolder holder(maxReadLength);
            std::vector<alignment::Match> matches;
            for (uint32_t c = b; c < e; ++c)
            {
                holder.load(reads, rml, c);
                matches.clear();
                for (uint64_t m = batch->clusterMatchBegin[c]; m < batch->clusterMatchBegin[c + 1]; ++m)
                    matches.push_back(alignment::Match(alignment::SeedId(batch->matches[m].seedId),
                                                       reference::ReferencePosition(batch->matches[m].location)));
                builtOut[c] = builder.build(contigs, rml, seeds, adapterList, matches.begin(), matches.end(), holder.cluster,
                                            batch->withGaps != 0);
                if (builtOut[c] || !matches.empty())
                {
                    for (uint32_t r = 0; r < rc; ++r)
                    {
                        for (const alignment::FragmentMetadata &f : builder.getFragments()[r]) parts[t].add(f, c * rc + r, rml);
                        counts[size_t(c) * rc + r] = builder.getFragments()[r].size();
                    }
                }
            }
        });
        readFragmentBegin[0] = 0;
        for (size_t i = 0; i < counts.size(); ++i) readFragmentBegin[i + 1] = readFragmentBegin[i] + counts[i];
        return concatenate(parts, fragmentCapacity, fragmentsOut, cigarCapacity, cigarsOut, fragmentCount, cigarWords);
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}

extern "C" int oracle_rescue_shadows(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                     const isaac_ext_tls_t *tls, uint32_t requestCount, const isaac_ext_rescue_request_t *requests,
                                     uint64_t fragmentCapacity, isaac_ext_fragment_t *fragmentsOut, uint64_t *requestFragmentBegin,
                                     uint64_t cigarCapacity, uint32_t *cigarsOut, uint8_t *rescuedOut,
                                     uint64_t *fragmentCount, uint64_t *cigarWords, uint32_t threads)
{
    try
    {
        const std::vector<reference::Contig> &contigs = makeContigs(genome);
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
        const alignment::matchSelector::SequencingAdapterList &adapterList = currentAdapters();
        const alignment::TemplateLengthStatistics stats(
            tls->min, tls->max, tls->median, tls->lowStdDev, tls->highStdDev,
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[0]),
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[1]), tls->mateDriftRange);
        const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
        const uint32_t rc = reads->readCount;
        if (threads < 1) threads = 1;
        if (requestCount < 2 * threads) threads = 1;
        std::vector<FlatOut> parts(threads);
        std::vector<uint64_t> counts(requestCount, 0);
        parallelFor(requestCount, threads, [&](uint32_t t, uint32_t b, uint32_t e) {
            alignment::ShadowAligner aligner(layouts, cfg->gappedMismatchesMax, cfg->avoidSmithWaterman, cfg->gapMatchScore,
                                             cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore, cfg->minGapExtendScore);
            ClusterHolder holder(maxReadLength);
            uint32_t loaded = -1U;
            std::vector<alignment::FragmentMetadata> shadowList;
            shadowList.reserve(1000);            // TemplateBuilder::TRACKED_REPEATS_MAX_ONE_READ (TemplateBuilder.hh:145, TemplateBuilder.cpp:82)
            for (uint32_t i = b; i < e; ++i)
            {
                const isaac_ext_rescue_request_t &q = requests[i];
                const uint32_t clusterId = q.orphanReadId / rc, readIndex = q.orphanReadId % rc;
                if (loaded != clusterId) { holder.load(reads, rml, clusterId); loaded = clusterId; }
                alignment::FragmentMetadata orphan(&holder.cluster, 0, readIndex);
                orphan.reverse = q.orphanContigStrand & 1;
                orphan.contigId = q.orphanContigStrand >> 1;
                orphan.position = q.orphanPosition;
                orphan.observedLength = q.orphanObservedLength;
                shadowList.clear();
                rescuedOut[i] = aligner.rescueShadow(contigs, orphan, shadowList, rml, adapterList, stats, q.bestTemplateLength);
                const uint32_t shadowReadId = clusterId * rc + (readIndex + 1) % 2;
                for (const alignment::FragmentMetadata &f : shadowList) parts[t].add(f, shadowReadId, rml);
                counts[i] = shadowList.size();
            }
        });
        requestFragmentBegin[0] = 0;
        for (uint32_t i = 0; i < requestCount; ++i) requestFragmentBegin[i + 1] = requestFragmentBegin[i] + counts[i];
        return concatenate(parts, fragmentCapacity, fragmentsOut, cigarCapacity, cigarsOut, fragmentCount, cigarWords);
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * TemplateBuilder over a flat batch (glue only; the pair selection and mapping scores are the reference's)
 * ------------------------------------------------------------------------------------------------------------------ */
extern "C" int oracle_build_templates(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                      const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                      const isaac_ext_template_options_t *options, isaac_ext_template_t *templatesOut,
                                      isaac_ext_fragment_t *fragmentsOut, uint64_t cigarCapacity, uint32_t *cigarsOut,
                                      uint64_t *cigarWords, uint32_t threads)
{
    try
    {
        const std::vector<reference::Contig> &contigs = makeContigs(genome);
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
        const alignment::matchSelector::SequencingAdapterList &adapterList = currentAdapters();
        alignment::SeedMetadataList seeds;
        for (uint32_t s = 0; s < batch->seedCount; ++s)
            seeds.push_back(alignment::SeedMetadata(batch->seeds[s].offset, batch->seeds[s].length, batch->seeds[s].readIndex, s));
        const alignment::TemplateLengthStatistics stats(
            tls->min, tls->max, tls->median, tls->lowStdDev, tls->highStdDev,
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[0]),
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[1]), tls->mateDriftRange);
        const alignment::RestOfGenomeCorrection rog(contigs, rml);
        const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
        const uint32_t n = reads->clusterCount, rc = reads->readCount;
        if (threads < 1) threads = 1;
        if (n < 2 * threads) threads = 1;
        std::vector<FlatOut> parts(threads);
        parallelFor(n, threads, [&](uint32_t t, uint32_t b, uint32_t e) {
            // on the heap like MatchSelector.cpp:150: the builder embeds ~32 MB of fixed-capacity vectors
            std::unique_ptr<alignment::TemplateBuilder> builder(new alignment::TemplateBuilder(
                layouts, cfg->repeatThreshold, cfg->maxSeedsPerRead, options->scatterRepeats != 0, cfg->gappedMismatchesMax,
                cfg->avoidSmithWaterman, cfg->gapMatchScore, cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore,
                cfg->minGapExtendScore, cfg->semialignedGapLimit,
                alignment::TemplateBuilder::DodgyAlignmentScore(options->dodgyAlignmentScore)));
            ClusterHolder holder(maxReadLength);
            std::vector<alignment::Match> matches;
            alignment::matchSelector::SemialignedEndsClipper semialignedClipper;
            alignment::matchSelector::OverlappingEndsClipper overlappingClipper;
            for (uint32_t c = b; c < e; ++c)
            {
                holder.load(reads, rml, c);
                matches.clear();
                for (uint64_t m = batch->clusterMatchBegin[c]; m < batch->clusterMatchBegin[c + 1]; ++m)
                    matches.push_back(alignment::Match(alignment::SeedId(batch->matches[m].seedId),
                                                       reference::ReferencePosition(batch->matches[m].location)));
                alignment::BamTemplate &bam = builder->getBamTemplate();
                isaac_ext_template_t &o = templatesOut[c];
                std::memset(&o, 0, sizeof(o));
                // MatchSelector.cpp:300-349: clusters without matches and clusters whose fragments did not build get an
                // initialised (unaligned) template
                if (!matches.empty() && !matches.front().location.isNoMatch() &&
                    builder->buildFragments(contigs, rml, seeds, adapterList, matches.begin(), matches.end(), holder.cluster,
                                            batch->withGaps != 0))
                {
                    o.hadFragments = 1;
                    o.built = builder->buildTemplate(contigs, rog, rml, adapterList, holder.cluster, stats, options->mapqThreshold);
                    if (o.built)                                                       // MatchSelector.cpp:336-346
                    {
                        if (options->clipFlags & ISAAC_EXT_CLIP_SEMIALIGNED) { semialignedClipper.reset(); semialignedClipper.clip(contigs, bam); }
                        if (options->clipFlags & ISAAC_EXT_CLIP_OVERLAPPING) { overlappingClipper.reset(); overlappingClipper.clip(contigs, bam); }
                    }
                }
                else
                {
                    bam.initialize(rml, holder.cluster);
                }
                o.alignmentScore = bam.getAlignmentScore();
                o.properPair = bam.isProperPair();
                for (uint32_t r = 0; r < rc; ++r)
                {
                    const alignment::FragmentMetadata &f = bam.getFragmentMetadata(r);
                    o.fragmentAlignmentScore[r] = f.alignmentScore;
                    parts[t].add(f, c * rc + r, rml);
                }
            }
        });
        uint64_t fragmentCount = 0;
        return concatenate(parts, uint64_t(n) * rc, fragmentsOut, cigarCapacity, cigarsOut, &fragmentCount, cigarWords);
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}

/* The TileBarcodeStats of MatchSelectorStats (MatchSelectorStats.hh:77-103) for one tile and one barcode: every cluster goes
 * through the template pipeline like in oracle_build_templates and is recorded the way MatchSelector::processMatchList records it
 * (MatchSelector.cpp:300-365; pfOnly off).  MatchSelectorStats.hh itself needs Boost.Filesystem and the barcode metadata; its
 * recordTemplate dispatch (:77-103) is restated here on the reference's own TileBarcodeStats and its two adapters. */
static void flattenStats(const alignment::matchSelector::TileBarcodeStats &s, uint64_t *out)
{
    out[0] += s.yield_; out[1] += s.yieldQ30_; out[2] += s.qualityScoreSum_; out[3] += s.clusterCount_;
    out[4] += s.unanchoredClusterCount_; out[5] += s.nmnmClusterCount_; out[6] += s.rmClusterCount_; out[7] += s.qcClusterCount_;
    out[8] += s.alignedFragmentCount_; out[9] += s.uniquelyAlignedFragmentCount_; out[10] += s.uniquelyAlignedPerfectFragmentCount_;
    out[11] += s.alignmentScoreSum_; out[12] += s.basesOutsideIndels_; out[13] += s.uniquelyAlignedBasesOutsideIndels_;
    out[14] += s.mismatches_; out[15] += s.uniquelyAlignedMismatches_;
    for (unsigned m = 0; m < 9; ++m) out[16 + m] += s.alignmentModelCounts_[m];
    for (unsigned m = 0; m < 4; ++m) out[25 + m] += s.nominalModelCounts_[m];
    out[29] += s.fragmentCount_;
}

/// every cluster of the tile through the template pipeline, recorded into parts[thread][readIndex * 2 + passesFilter] the way
/// MatchSelectorStats::recordTemplate dispatches (MatchSelectorStats.hh:77-103); Stats = TileBarcodeStats or TileStats
template <class Stats>
static void recordTileTemplates(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls, const isaac_ext_template_options_t *options,
                                const uint8_t *pf, std::vector<std::vector<Stats> > &parts, uint32_t threads)
{
    using namespace alignment::matchSelector;
    const std::vector<reference::Contig> &contigs = makeContigs(genome);
    const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
    const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
    const SequencingAdapterList &adapterList = currentAdapters();
    alignment::SeedMetadataList seeds;
    for (uint32_t s = 0; s < batch->seedCount; ++s)
        seeds.push_back(alignment::SeedMetadata(batch->seeds[s].offset, batch->seeds[s].length, batch->seeds[s].readIndex, s));
    const alignment::TemplateLengthStatistics stats(
        tls->min, tls->max, tls->median, tls->lowStdDev, tls->highStdDev,
        alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[0]),
        alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[1]), tls->mateDriftRange);
    const alignment::RestOfGenomeCorrection rog(contigs, rml);
    const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
    const uint32_t n = reads->clusterCount;
    parallelFor(n, threads, [&](uint32_t t, uint32_t b, uint32_t e) {
        std::unique_ptr<alignment::TemplateBuilder> builder(new alignment::TemplateBuilder(
            layouts, cfg->repeatThreshold, cfg->maxSeedsPerRead, options->scatterRepeats != 0, cfg->gappedMismatchesMax,
            cfg->avoidSmithWaterman, cfg->gapMatchScore, cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore,
            cfg->minGapExtendScore, cfg->semialignedGapLimit,
            alignment::TemplateBuilder::DodgyAlignmentScore(options->dodgyAlignmentScore)));
        ClusterHolder holder(maxReadLength);
        std::vector<alignment::Match> matches;
        SemialignedEndsClipper semialignedClipper;
        OverlappingEndsClipper overlappingClipper;
        std::vector<Stats> &mine = parts[t];
        for (uint32_t c = b; c < e; ++c)
        {
            holder.load(reads, rml, c, !pf || pf[c]);
            matches.clear();
            for (uint64_t m = batch->clusterMatchBegin[c]; m < batch->clusterMatchBegin[c + 1]; ++m)
                matches.push_back(alignment::Match(alignment::SeedId(batch->matches[m].seedId),
                                                   reference::ReferencePosition(batch->matches[m].location)));
            alignment::BamTemplate &bam = builder->getBamTemplate();
            TemplateAlignmentType type = Normal;
            if (matches.empty() || matches.front().location.isNoMatch())                      // MatchSelector.cpp:302-313
            {
                bam.initialize(rml, holder.cluster);
                type = !matches.empty() && matches.front().getSeedId().isNSeedId() ? Qc : NmNm;
            }
            else if (builder->buildFragments(contigs, rml, seeds, adapterList, matches.begin(), matches.end(), holder.cluster,
                                             batch->withGaps != 0))
            {
                if (builder->buildTemplate(contigs, rog, rml, adapterList, holder.cluster, stats, options->mapqThreshold)) // :329-347
                {
                    if (options->clipFlags & ISAAC_EXT_CLIP_SEMIALIGNED) { semialignedClipper.reset(); semialignedClipper.clip(contigs, bam); }
                    if (options->clipFlags & ISAAC_EXT_CLIP_OVERLAPPING) { overlappingClipper.reset(); overlappingClipper.clip(contigs, bam); }
                }
            }
            else
            {
                bam.initialize(rml, holder.cluster);
                type = Rm;
            }
            // MatchSelectorStats::recordTemplate (MatchSelectorStats.hh:77-103)
            BamTemplateTileStatsAdapter templateAdapter(stats, bam, type);
            const unsigned r0 = bam.getFragmentMetadata(0).getReadIndex();
            if (bam.getPassesFilter()) mine[r0 * 2 + 1].recordTemplate(templateAdapter);
            mine[r0 * 2].recordTemplate(templateAdapter);
            for (unsigned i = 0; bam.getFragmentCount() > i; ++i)
            {
                const alignment::FragmentMetadata &fragment = bam.getFragmentMetadata(i);
                FragmentMetadataTileStatsAdapter fragmentAdapter(fragment);
                if (bam.getPassesFilter()) mine[fragment.getReadIndex() * 2 + 1].recordFragment(fragmentAdapter, rml.at(i));
                mine[fragment.getReadIndex() * 2].recordFragment(fragmentAdapter, rml.at(i));
            }
        }
    });
}

extern "C" int oracle_template_stats(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                     const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                     const isaac_ext_template_options_t *options, const uint8_t *pf, uint64_t *statsOut, uint32_t threads)
{
    using namespace alignment::matchSelector;
    try
    {
        if (threads < 1) threads = 1;
        if (reads->clusterCount < 2 * threads) threads = 1;
        std::vector<std::vector<TileBarcodeStats> > parts(threads, std::vector<TileBarcodeStats>(4));   // [readIndex * 2 + passesFilter]
        // TileBarcodeStats::reset() zeroes the eight real models but not alignmentModelCounts_[InvalidAlignmentModel]
        // (TileBarcodeStats.hh:62-69): the reference counts the pairs without a model on top of uninitialised memory
        for (std::vector<TileBarcodeStats> &p : parts)
            for (TileBarcodeStats &s : p) s.alignmentModelCounts_[alignment::TemplateLengthStatistics::InvalidAlignmentModel] = 0;
        recordTileTemplates(genome, reads, cfg, batch, tls, options, pf, parts, threads);
        std::memset(statsOut, 0, 4 * ISAAC_EXT_TEMPLATE_STATS_COUNTERS * sizeof(uint64_t));
        for (const std::vector<TileBarcodeStats> &p : parts)
            for (unsigned k = 0; k < 4; ++k) flattenStats(p[k], statsOut + k * ISAAC_EXT_TEMPLATE_STATS_COUNTERS);
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
    return ISAAC_EXT_OK;
}

/* The TileStats of MatchSelectorStats (TileStats.hh:68-142) for one tile: the same walk, recorded into the reference's own TileStats.
 * statsOut: 4 blocks of ISAAC_EXT_TILE_CYCLE_STATS_WORDS u64 in the member order of the struct (= its memory layout), summed over
 * the threads with TileStats::operator+=; finalize != 0 applies TileStats::finalize to every block. */
extern "C" int oracle_tile_cycle_stats(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *cfg,
                                       const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                       const isaac_ext_template_options_t *options, const uint8_t *pf, uint64_t *statsOut, uint32_t finalize,
                                       uint32_t threads)
{
    using namespace alignment::matchSelector;
    static_assert(sizeof(TileStats) == ISAAC_EXT_TILE_CYCLE_STATS_WORDS * sizeof(uint64_t), "TileStats is a plain block of 64-bit counters");
    try
    {
        if (threads < 1) threads = 1;
        if (reads->clusterCount < 2 * threads) threads = 1;
        std::vector<std::vector<TileStats> > parts(threads, std::vector<TileStats>(4));
        recordTileTemplates(genome, reads, cfg, batch, tls, options, pf, parts, threads);
        for (unsigned k = 0; k < 4; ++k)
        {
            TileStats sum;
            for (const std::vector<TileStats> &p : parts) sum += p[k];
            if (finalize) sum.finalize();
            std::memcpy(statsOut + size_t(k) * ISAAC_EXT_TILE_CYCLE_STATS_WORDS, &sum, sizeof(sum));
        }
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
    return ISAAC_EXT_OK;
}

/* MatchSelector::determineTemplateLength (MatchSelector.cpp:188-249) for one tile: the loop is restated here (MatchSelector.cpp
 * itself drags in the whole workflow), everything it calls -- TemplateBuilder::buildFragments without gaps and
 * TemplateLengthDistribution::addTemplate / finalize -- is the reference's own code. */
extern "C" int oracle_determine_template_length(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                                const isaac_ext_config_t *cfg, const isaac_ext_build_batch_t *batch,
                                                const uint8_t *pf, int32_t mateDriftRange, isaac_ext_tls_t *tlsOut, uint32_t *stableOut)
{
    try
    {
        const std::vector<reference::Contig> &contigs = makeContigs(genome);
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const flowcell::FlowcellLayoutList layouts(1, flowcell::Layout(rml));
        const alignment::matchSelector::SequencingAdapterList &adapterList = currentAdapters();
        alignment::SeedMetadataList seeds;
        for (uint32_t s = 0; s < batch->seedCount; ++s)
            seeds.push_back(alignment::SeedMetadata(batch->seeds[s].offset, batch->seeds[s].length, batch->seeds[s].readIndex, s));
        alignment::TemplateLengthDistribution distribution(mateDriftRange);
        distribution.reset(contigs, rml);
        if (reads->readCount == 2)
        {
            std::unique_ptr<alignment::TemplateBuilder> builder(new alignment::TemplateBuilder(
                layouts, cfg->repeatThreshold, cfg->maxSeedsPerRead, false, cfg->gappedMismatchesMax,
                cfg->avoidSmithWaterman, cfg->gapMatchScore, cfg->gapMismatchScore, cfg->gapOpenScore, cfg->gapExtendScore,
                cfg->minGapExtendScore, cfg->semialignedGapLimit, alignment::TemplateBuilder::DODGY_ALIGNMENT_SCORE_UNALIGNED));
            ClusterHolder holder(std::max(reads->readLength[0], reads->readLength[1]));
            std::vector<alignment::Match> matches;
            for (uint32_t c = 0; c < reads->clusterCount && !distribution.getStatistics().isStable(); ++c)
            {
                matches.clear();
                for (uint64_t m = batch->clusterMatchBegin[c]; m < batch->clusterMatchBegin[c + 1]; ++m)
                    matches.push_back(alignment::Match(alignment::SeedId(batch->matches[m].seedId),
                                                       reference::ReferencePosition(batch->matches[m].location)));
                if (matches.empty() || (pf && !pf[c]) || matches.front().location.isNoMatch()) continue;      // :226-230
                holder.load(reads, rml, c);
                builder->buildFragments(contigs, rml, seeds, adapterList, matches.begin(), matches.end(), holder.cluster, false);
                distribution.addTemplate(builder->getFragments());
            }
            if (!distribution.isStable()) distribution.finalize();
        }
        const alignment::TemplateLengthStatistics &s = distribution.getStatistics();
        tlsOut->min = s.getMin(); tlsOut->max = s.getMax(); tlsOut->median = s.getMedian();
        tlsOut->lowStdDev = s.getLowStdDev(); tlsOut->highStdDev = s.getHighStdDev();
        tlsOut->bestModel[0] = s.getBestModel(0); tlsOut->bestModel[1] = s.getBestModel(1);
        tlsOut->mateDriftRange = mateDriftRange;
        *stableOut = s.isStable();
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
    return ISAAC_EXT_OK;
}

extern "C" int oracle_trim_low_quality_ends(const isaac_ext_reads_t *reads, uint32_t baseQualityCutoff, uint16_t *endCyclesMaskedOut)
{
    try
    {
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const unsigned maxReadLength = std::max(reads->readLength[0], reads->readLength[1]);
        ClusterHolder holder(maxReadLength);
        isaac_ext_reads_t unmasked = *reads;
        unmasked.endCyclesMasked = 0;                                                 // Cluster::init starts from unmasked reads
        for (uint32_t c = 0; c < reads->clusterCount; ++c)
        {
            holder.load(&unmasked, rml, c);
            alignment::trimLowQualityEnds(holder.cluster, baseQualityCutoff);
            for (uint32_t r = 0; r < reads->readCount; ++r)
                endCyclesMaskedOut[size_t(c) * reads->readCount + r] = uint16_t(holder.cluster[r].getEndCyclesMasked());
        }
        return ISAAC_EXT_OK;
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}

/* FragmentCollector::add for every stored template of a tile (FragmentCollector.cpp:42-77, storeBclAndCigar :79-103; the
 * FragmentBuffer layout FragmentCollector.hh:283-308).  FragmentCollector.hh itself needs Boost.Filesystem (BinMetadata.hh), so
 * its two short functions are followed here line by line; every byte of the header comes from the reference's own
 * io::FragmentHeader constructors, getMaxTotalLength, BamTemplate and FragmentMetadata. */
extern "C" int oracle_pack_fragments(const isaac_ext_reads_t *reads, const isaac_ext_template_t *templates,
                                     const isaac_ext_fragment_t *fragments, const uint32_t *cigars, uint64_t cigarWords,
                                     const isaac_ext_pack_options_t *options, const uint8_t *barcodeBytes, uint32_t barcodeLength,
                                     uint8_t *recordsOut, uint64_t *fStrandPosOut, uint8_t *initializedOut, uint8_t *headerMaskOut,
                                     uint32_t *layoutOut)
{
    try
    {
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const uint32_t n = reads->clusterCount, rc = reads->readCount;
        const unsigned len0 = reads->readLength[0], len1 = rc > 1 ? reads->readLength[1] : 0;
        // FragmentBuffer::getRecordLength / getReadOffsets
        const unsigned recordLength = io::FragmentHeader::getMaxTotalLength(len0) + (len1 ? io::FragmentHeader::getMaxTotalLength(len1) : 0);
        const unsigned readOffsets[2] = {0, rc > 1 ? io::FragmentHeader::getMaxTotalLength(len0) : 0};
        layoutOut[0] = recordLength; layoutOut[1] = readOffsets[0]; layoutOut[2] = readOffsets[1]; layoutOut[3] = sizeof(io::FragmentHeader);
        if (headerMaskOut)
        {
            std::memset(headerMaskOut, 0, sizeof(io::FragmentHeader));
#define ORACLE_MEMBER(m) std::memset(headerMaskOut + offsetof(io::FragmentHeader, m), 0xFF, sizeof(io::FragmentHeader::m))
            ORACLE_MEMBER(bamTlen_); ORACLE_MEMBER(observedLength_); ORACLE_MEMBER(fStrandPosition_); ORACLE_MEMBER(lowClipped_);
            ORACLE_MEMBER(highClipped_); ORACLE_MEMBER(alignmentScore_); ORACLE_MEMBER(templateAlignmentScore_);
            ORACLE_MEMBER(mateFStrandPosition_); ORACLE_MEMBER(readLength_); ORACLE_MEMBER(cigarLength_); ORACLE_MEMBER(gapCount_);
            ORACLE_MEMBER(editDistance_); ORACLE_MEMBER(tile_); ORACLE_MEMBER(barcode_); ORACLE_MEMBER(barcodeSequence_);
            ORACLE_MEMBER(clusterId_); ORACLE_MEMBER(clusterX_); ORACLE_MEMBER(clusterY_); ORACLE_MEMBER(duplicateClusterRank_);
            ORACLE_MEMBER(mateAnchor_); ORACLE_MEMBER(mateStorageBin_);
#undef ORACLE_MEMBER
            // the ten bit fields of Flags: found by setting each of them in an all-zero header
            io::FragmentHeader probe;
            std::memset(&probe, 0, sizeof(probe));
            probe.flags_.paired_ = probe.flags_.unmapped_ = probe.flags_.mateUnmapped_ = probe.flags_.reverse_ = probe.flags_.mateReverse_ =
                probe.flags_.firstRead_ = probe.flags_.secondRead_ = probe.flags_.failFilter_ = probe.flags_.properPair_ = probe.flags_.duplicate_ = true;
            for (size_t i = 0; i < sizeof(probe); ++i) headerMaskOut[i] |= reinterpret_cast<const unsigned char *>(&probe)[i];
        }
        if (!recordsOut) return ISAAC_EXT_OK;
        std::memset(recordsOut, 0, size_t(n) * recordLength);                       // FragmentBuffer::resize value-initialises data_
        const unsigned maxReadLength = std::max(len0, len1);
        std::vector<unsigned> cigarBuffer(cigars, cigars + cigarWords);
        alignment::BamTemplate bam(cigarBuffer);
        alignment::Cluster cluster(maxReadLength);
        std::vector<char> bcl;
        for (uint32_t c = 0; c < n; ++c)
        {
            for (uint32_t r = 0; r < rc; ++r) { fStrandPosOut[size_t(c) * rc + r] = 0; initializedOut[size_t(c) * rc + r] = 0; }
            if (!(templates[c].built || options->keepUnaligned)) continue;
            bcl.clear();
            if (barcodeBytes) bcl.insert(bcl.end(), barcodeBytes + size_t(c) * barcodeLength, barcodeBytes + size_t(c + 1) * barcodeLength);
            bcl.insert(bcl.end(), reads->bcl + size_t(c) * (len0 + len1), reads->bcl + size_t(c + 1) * (len0 + len1));
            bcl.resize(bcl.size() + 64, 0);      // pack32BclBases reads 32 bytes whatever the read length
            const alignment::ClusterXy xy = options->xy ? alignment::ClusterXy(options->xy[2 * size_t(c)], options->xy[2 * size_t(c) + 1])
                                                        : alignment::ClusterXy();
            cluster.init(rml, bcl.begin(), options->tile, c, xy, options->pf ? options->pf[c] != 0 : true, barcodeBytes ? barcodeLength : 0);
            bam.initialize(rml, cluster);
            bam.setAlignmentScore(templates[c].alignmentScore);
            bam.setProperPair(templates[c].properPair);
            for (uint32_t r = 0; r < rc; ++r)
            {
                const isaac_ext_fragment_t &f = fragments[size_t(c) * rc + r];
                alignment::FragmentMetadata &m = bam.getFragmentMetadata(r);
                m.contigId = f.contigId; m.position = f.position; m.observedLength = f.observedLength; m.reverse = f.reverse;
                m.lowClipped = f.lowClipped; m.highClipped = f.highClipped; m.editDistance = f.editDistance; m.gapCount = f.gapCount;
                m.mismatchCount = f.mismatchCount; m.cigarOffset = f.cigarOffset; m.cigarLength = f.cigarLength;
                m.alignmentScore = templates[c].fragmentAlignmentScore[r];
            }
            for (unsigned k = 0; k < bam.getFragmentCount(); ++k)                     // BufferingFragmentStorage::add
            {
                // FragmentCollector::add
                const alignment::FragmentMetadata &fragment = bam.getFragmentMetadata(k);
                char *dataBytes = reinterpret_cast<char *>(recordsOut) + size_t(c) * recordLength + readOffsets[fragment.getReadIndex()];
                const size_t slot = size_t(c) * rc + fragment.getReadIndex();
                initializedOut[slot] = 1;
                fStrandPosOut[slot] = fragment.getFStrandReferencePosition().getValue();
                // storeBclAndCigar
                char *variableData = dataBytes + sizeof(io::FragmentHeader);
                std::vector<char>::const_iterator bclData = fragment.getBclData();
                if (fragment.isReverse())
                    variableData = std::transform(std::reverse_iterator<std::vector<char>::const_iterator>(bclData + fragment.getReadLength()),
                                                  std::reverse_iterator<std::vector<char>::const_iterator>(bclData), variableData, oligo::getReverseBcl);
                else
                    variableData = std::copy(bclData, bclData + fragment.getReadLength(), variableData);
                if (fragment.isAligned())
                {
                    const alignment::Cigar::const_iterator cigarBegin = fragment.cigarBuffer->begin() + fragment.cigarOffset;
                    std::memcpy(variableData, &*cigarBegin, size_t(fragment.cigarLength) * sizeof(unsigned));
                }
                io::FragmentHeader header;
                if (2 == bam.getFragmentCount())
                {
                    const alignment::FragmentMetadata &mate = bam.getMateFragmentMetadata(fragment);
                    unsigned mateStorageBin = 0;
                    if (!fragment.isNoMatch() && options->distributionBinSize)             // BinIndexMap::getBinIndex (BinIndexMap.hh:96-107)
                    {
                        const reference::ReferencePosition pos = mate.getFStrandReferencePosition();
                        const uint64_t begin = options->contigBinBegin[pos.getContigId()];
                        mateStorageBin = options->binIndex[begin + pos.getPosition() / options->distributionBinSize];
                    }
                    header = io::FragmentHeader(bam, fragment, mate, options->barcodeIdx, mateStorageBin);
                }
                else
                {
                    header = io::FragmentHeader(bam, fragment, options->barcodeIdx);
                }
                std::memcpy(dataBytes, &header, sizeof(header));
            }
        }
        return ISAAC_EXT_OK;
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}

/* calculateShadowRescueRange is a free function of ShadowAligner.cpp with external linkage (:119-149), not declared in a header */
namespace isaac { namespace alignment {
std::pair<long, long> calculateShadowRescueRange(const FragmentMetadata &orphan, const TemplateLengthStatistics &templateLengthStatistics,
                                                 const long bestTemplateLength);
} }

extern "C" int oracle_shadow_rescue_range(const isaac_ext_reads_t *reads, const isaac_ext_tls_t *tls, uint32_t requestCount,
                                          const isaac_ext_rescue_request_t *requests, int64_t *rangeOut, uint8_t *orientationOut)
{
    try
    {
        const flowcell::ReadMetadataList rml = makeReadMetadata(reads);
        const alignment::TemplateLengthStatistics stats(
            tls->min, tls->max, tls->median, tls->lowStdDev, tls->highStdDev,
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[0]),
            alignment::TemplateLengthStatistics::AlignmentModel(tls->bestModel[1]), tls->mateDriftRange);
        ClusterHolder holder(std::max(reads->readLength[0], reads->readLength[1]));
        for (uint32_t i = 0; i < requestCount; ++i)
        {
            const isaac_ext_rescue_request_t &q = requests[i];
            holder.load(reads, rml, q.orphanReadId / reads->readCount);
            alignment::FragmentMetadata orphan;
            orphan.cluster = &holder.cluster;
            orphan.readIndex = q.orphanReadId % reads->readCount;
            orphan.contigId = q.orphanContigStrand >> 1;
            orphan.reverse = q.orphanContigStrand & 1;
            orphan.position = q.orphanPosition;
            orphan.observedLength = q.orphanObservedLength;
            const std::pair<long, long> range = alignment::calculateShadowRescueRange(orphan, stats, q.bestTemplateLength);
            rangeOut[2 * size_t(i)] = range.first; rangeOut[2 * size_t(i) + 1] = range.second;
            orientationOut[i] = stats.mateOrientation(orphan.readIndex, orphan.reverse);
        }
        return ISAAC_EXT_OK;
    }
    catch (const std::exception &e)
    {
        return ISAAC_EXT_E_INVALID_ARG;
    }
}
