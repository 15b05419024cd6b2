// isaac_ext_build_fragments and isaac_ext_rescue_shadows (include/isaac_ext.h).  Included at the end of isaac_ext.cu.
#pragma once

namespace
{

/// the clippers of a tile call: ranges[slot] from the first candidate of every slot (checkInitStrand, FragmentBuilder.cpp:173)
int initAdapterSlots(isaac_ext_ctx *ctx, uint32_t slots, const isaac_ext_candidate_t *dFirst)
{
    if (!ctx->adapters.count || !slots) return ISAAC_EXT_OK;
    CK(ctx->dAdapterRanges.reserve(slots));
    adapterInitKernel<<<gridFor(ctx, slots, 128, 16), 128, 0, ctx->stream>>>(ctx->adapters, ctx->ref, ctx->reads, slots, dFirst, ctx->dAdapterRanges.p);
    ++ctx->launches;
    return ctx->cuda(cudaGetLastError(), "adapterInitKernel");
}

/// the candidates of the tile calls share clippers: slot = hSlotOf[i], or readId * 2 + reverse with hSlotOf == nullptr
int tileAdapterClip(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *dCand, const uint32_t *hSlotOf, const uint32_t **clip)
{
    *clip = nullptr;
    if (!ctx->adapters.count || !n) return ISAAC_EXT_OK;
    PipelineState &ps = ctx->pipeline;
    if (hSlotOf)
    {
        CK(ps.dSlot.reserve(n));
        CK(cudaMemcpyAsync(ps.dSlot.p, hSlotOf, size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    return adapterSlotClip(ctx, n, dCand, hSlotOf ? ps.dSlot.p : nullptr, ctx->stream, clip);
}

int runUngapped(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *hCand, isaac_ext_fragment_t *hFrag, uint32_t *hCig)
{
    PipelineState &ps = ctx->pipeline;
    if (!n) return ISAAC_EXT_OK;
    CK(ps.dCand.reserve(n)); CK(ps.dFrag.reserve(n)); CK(ps.dCig.reserve(size_t(n) * 3));
    CK(cudaMemcpyAsync(ps.dCand.p, hCand, size_t(n) * sizeof(*hCand), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *clip = nullptr;
    int rc = tileAdapterClip(ctx, n, ps.dCand.p, nullptr, &clip);
    if (rc) return rc;
    rc = ungappedDevice(ctx, n, ps.dCand.p, ps.dFrag.p, ps.dCig.p, nullptr, ctx->stream, clip);
    if (rc) return rc;
    CK(cudaMemcpyAsync(hFrag, ps.dFrag.p, size_t(n) * sizeof(*hFrag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hCig, ps.dCig.p, size_t(n) * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return ctx->cuda(cudaStreamSynchronize(ctx->stream), "ungapped pass");
}

int runGapped(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *hCand, uint32_t stride, isaac_ext_fragment_t *hFrag, uint32_t *hCig,
              const uint32_t *hSlotOf = nullptr)
{
    PipelineState &ps = ctx->pipeline;
    if (!n) return ISAAC_EXT_OK;
    CK(ps.dCand.reserve(n)); CK(ps.dFrag.reserve(n)); CK(ps.dCig.reserve(size_t(n) * stride));
    CK(cudaMemcpyAsync(ps.dCand.p, hCand, size_t(n) * sizeof(*hCand), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *clip = nullptr;
    int rc = tileAdapterClip(ctx, n, ps.dCand.p, hSlotOf, &clip);
    if (rc) return rc;
    rc = gappedDevice(ctx, n, ps.dCand.p, stride, ps.dFrag.p, ps.dCig.p, nullptr, ctx->stream, clip);
    if (rc) return rc;
    CK(cudaMemcpyAsync(hFrag, ps.dFrag.p, size_t(n) * sizeof(*hFrag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hCig, ps.dCig.p, size_t(n) * stride * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return checkErrorFlag(ctx);
}

const uint32_t GAPPED_STRIDE = 32;

isaac_ext_candidate_t candidateOf(const isaac_ext_fragment_t &f, long position)
{
    isaac_ext_candidate_t c;
    c.position = position; c.readId = f.readId; c.contigStrand = (f.contigId << 1) | (f.reverse ? 1u : 0u);
    return c;
}

/// copies the surviving fragments of every group, in order, into the flat result arrays (two parallel passes:
/// count per partition, then fill; the output buffers are grow-only and never zero-filled)
template <class ListOf>
void flatten(isaac_ext_ctx *ctx, const HostPools &pools, size_t groups, ListOf listOf)
{
    PipelineState &ps = ctx->pipeline;
    const unsigned T = ctx->hostThreads;
    const unsigned parts = partitionCount(T, groups);
    std::vector<uint64_t> partFragments(parts + 1, 0), partWords(parts + 1, 0);
    ps.outBegin.reserve(groups + 1);
    parallelRanges(T, groups, [&](unsigned t, size_t b, size_t e) {
        uint64_t nf = 0, nw = 0;
        for (size_t g = b; g < e; ++g)
        {
            const std::pair<const WorkFragment *, unsigned> l = listOf(g);
            nf += l.second;
            for (unsigned k = 0; k < l.second; ++k) nw += l.first[k].f.cigarLength;
        }
        partFragments[t + 1] = nf; partWords[t + 1] = nw;
    });
    for (unsigned p = 0; p < parts; ++p) { partFragments[p + 1] += partFragments[p]; partWords[p + 1] += partWords[p]; }
    ps.outFragments.reserve(partFragments[parts]);
    ps.outCigars.reserve(partWords[parts]);
    ps.outFragmentCount = partFragments[parts]; ps.outCigarWords = partWords[parts];
    parallelRanges(T, groups, [&](unsigned t, size_t b, size_t e) {
        uint64_t nf = partFragments[t], at = partWords[t];
        for (size_t g = b; g < e; ++g)
        {
            const std::pair<const WorkFragment *, unsigned> l = listOf(g);
            ps.outBegin.p[g] = nf;
            for (unsigned k = 0; k < l.second; ++k)
            {
                isaac_ext_fragment_t f = l.first[k].f;
                std::copy(pools.cigar(l.first[k]), pools.cigar(l.first[k]) + f.cigarLength, ps.outCigars.p + at);
                f.cigarOffset = uint32_t(at);
                at += f.cigarLength;
                ps.outFragments.p[nf++] = f;
            }
        }
    });
    ps.outBegin.p[groups] = partFragments[parts];
}

} // namespace

extern "C" int isaac_ext_build_fragments(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, isaac_ext_build_result_t *result)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!batch || !result || !batch->clusterMatchBegin || !batch->seeds || !batch->seedCount)
        return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null batch");
    CK(cudaSetDevice(ctx->device));
    PipelineState &ps = ctx->pipeline;
    const uint32_t nClusters = ctx->clusterCount, rc = ctx->reads.readCount;
    const size_t lists = size_t(nClusters) * rc;
    const uint64_t M = batch->clusterMatchBegin[nClusters];
    if (M && !batch->matches) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null matches");
    for (uint32_t s = 0; s < batch->seedCount; ++s)
        if (batch->seeds[s].readIndex >= rc) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "seed refers to an unknown read");
    const unsigned T = ctx->hostThreads;
    const unsigned parts = partitionCount(T, nClusters);
    const unsigned repeatThreshold = ctx->cfg.repeatThreshold;
    ps.work.reserve(M);
    ps.outFlags.assign(nClusters, 0);
    std::vector<uint64_t> listBegin(lists, 0);
    std::vector<uint32_t> listCount(lists, 0);
    std::vector<uint64_t> partCount(parts + 1, 0);
    std::atomic<int> bad(0);
    HostPools pools;

    PhaseTimer timer("build");
    // ---- P1: FragmentBuilder::build up to alignFragments (FragmentBuilder.cpp:92-134) + consolidate (:159)
    parallelRanges(T, nClusters, [&](unsigned t, size_t b, size_t e) {
        std::vector<unsigned> seedMatchCounts(batch->seedCount);
        std::vector<WorkFragment> tmp[2];
        uint64_t total = 0;
        for (size_t c = b; c < e; ++c)
        {
            const uint64_t mb = batch->clusterMatchBegin[c], me = batch->clusterMatchBegin[c + 1];
            if (me < mb || me > M) { bad = 1; continue; }
            if (mb == me) continue;
            std::fill(seedMatchCounts.begin(), seedMatchCounts.end(), 0u);
            unsigned repeatSeedsCount = 0;
            tmp[0].clear(); tmp[1].clear();
            for (uint64_t m = mb; m < me && !matchIsNoMatch(batch->matches[m]); ++m)
            {
                const isaac_ext_match_t &match = batch->matches[m];
                const unsigned seedIndex = matchSeed(match);
                if (seedIndex >= batch->seedCount) { bad = 1; break; }          // seedMatchCounts_.at() would throw
                if (repeatThreshold > seedMatchCounts[seedIndex])
                {
                    if (matchIsTooMany(match)) { seedMatchCounts[seedIndex] = repeatThreshold; ++repeatSeedsCount; }
                    else if (repeatThreshold == ++seedMatchCounts[seedIndex]) ++repeatSeedsCount;
                    else
                    {
                        // addMatch (:219-249) with getReadPosition (:326-343)
                        const isaac_ext_seed_t &seed = batch->seeds[seedIndex];
                        const bool reverse = matchReverse(match);
                        const long seedPosition = matchPosition(match);
                        const long readLength = ctx->reads.readLength[seed.readIndex];
                        WorkFragment w;
                        std::memset(&w, 0, sizeof(w));
                        w.f.position = reverse ? seedPosition + long(seed.length) + long(seed.offset) - readLength
                                               : seedPosition - long(seed.offset);
                        w.f.contigId = matchContig(match);
                        w.f.readId = uint32_t(c) * rc + seed.readIndex;
                        w.f.readIndex = uint8_t(seed.readIndex);
                        w.f.reverse = reverse;
                        w.f.firstSeedIndex = int16_t(seedIndex);
                        w.f.nonUniqueSeedOffsetFirst = 0xFFFF;
                        if (seed.length != 64 && matchHasNeighbors(match))      // STRONG_SEED_LENGTH (Alignment.hh:38)
                        {
                            w.f.nonUniqueSeedOffsetFirst = std::min<uint16_t>(w.f.nonUniqueSeedOffsetFirst, seed.offset);
                            w.f.nonUniqueSeedOffsetSecond = std::max<uint16_t>(w.f.nonUniqueSeedOffsetSecond, seed.offset);
                        }
                        else w.f.uniqueSeedCount = 1;
                        if (w.f.contigId >= ctx->ref.contigCount || w.f.position > long(ctx->contigLength[w.f.contigId])) { bad = 1; continue; }
                        tmp[seed.readIndex].push_back(w);
                    }
                }
            }
            if (repeatSeedsCount)                                               // removeRepeatSeedAlignments (:128-134, 261-266)
                for (std::vector<WorkFragment> &list : tmp)
                    list.erase(std::remove_if(list.begin(), list.end(), [&](const WorkFragment &w) {
                                   return seedMatchCounts[w.f.firstSeedIndex] >= repeatThreshold; }), list.end());
            ps.outFlags[c] = !(tmp[0].empty() && tmp[1].empty());               // return value of build() (:136-144)
            uint64_t at = mb;
            for (unsigned r = 0; r < rc; ++r)
            {
                const unsigned n = consolidateDuplicateFragments(tmp[r].data(), unsigned(tmp[r].size()), false);    // :159
                for (unsigned k = 0; k < n; ++k) { tmp[r][k].f.repeatSeedsCount = uint16_t(repeatSeedsCount); ps.work.p[at + k] = tmp[r][k]; }   // :167
                listBegin[c * rc + r] = at; listCount[c * rc + r] = n;
                at += n; total += n;
            }
        }
        partCount[t + 1] = total;
    });
    timer.mark("P1 candidates");
    if (bad) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "malformed match batch (offsets, seed index or contig out of range)");
    for (unsigned p = 0; p < parts; ++p) partCount[p + 1] += partCount[p];
    const uint64_t n1 = partCount[parts];
    if (n1 > 0xFFFFFFF0ull / GAPPED_STRIDE) return ctx->fail(ISAAC_EXT_E_CAPACITY, "too many candidates in one batch");
    CK(ps.hCand1.reserve(n1)); CK(ps.hFrag1.reserve(n1)); CK(ps.hCig1.reserve(n1 * 3));
    const bool withAdapters = ctx->adapters.count != 0;
    if (withAdapters) CK(ps.hAdapterFirst.reserve(lists * 2));
    parallelRanges(T, nClusters, [&](unsigned t, size_t b, size_t e) {
        uint64_t at = partCount[t];
        for (size_t l = b * rc; l < e * rc; ++l)
        {
            // one FragmentSequencingAdapterClipper per read list (:164): its two strands are slots l * 2 + reverse, each
            // initialised by the first fragment of that strand in list order (:173)
            if (withAdapters) ps.hAdapterFirst.p[l * 2].readId = ps.hAdapterFirst.p[l * 2 + 1].readId = ADAPTER_NO_CANDIDATE;
            for (unsigned k = 0; k < listCount[l]; ++k)
            {
                WorkFragment &w = ps.work.p[listBegin[l] + k];
                w.slot = uint32_t(at);
                ps.hCand1.p[at] = candidateOf(w.f, w.f.position);
                if (withAdapters && ps.hAdapterFirst.p[l * 2 + (w.f.reverse ? 1 : 0)].readId == ADAPTER_NO_CANDIDATE)
                    ps.hAdapterFirst.p[l * 2 + (w.f.reverse ? 1 : 0)] = ps.hCand1.p[at];
                ++at;
            }
        }
    });
    timer.mark("P1 dense write");
    if (withAdapters)
    {
        CK(ps.dAdapterFirst.reserve(lists * 2));
        CK(cudaMemcpyAsync(ps.dAdapterFirst.p, ps.hAdapterFirst.p, lists * 2 * sizeof(isaac_ext_candidate_t), cudaMemcpyHostToDevice, ctx->stream));
        const int rca = initAdapterSlots(ctx, uint32_t(lists * 2), ps.dAdapterFirst.p);
        if (rca) return rca;
    }
    // ---- K1: UngappedAligner::alignUngapped of every candidate (:174)
    int rcode = runUngapped(ctx, uint32_t(n1), ps.hCand1.p, ps.hFrag1.p, ps.hCig1.p);
    if (rcode) return rcode;
    timer.mark("K1 ungapped + copies");
    pools.pools[0] = ps.hCig1.p;

    // ---- P2: consolidate (:179), SimpleIndelAligner::alignSimpleIndels pairing (SimpleIndelAligner.cpp:460-518)
    std::vector<std::vector<IndelTask>> partTasks(parts);
    std::vector<std::vector<uint64_t>> partTargets(parts);
    const unsigned gapLimit = ctx->cfg.semialignedGapLimit;
    parallelRanges(T, nClusters, [&](unsigned t, size_t b, size_t e) {
        for (size_t l = b * rc; l < e * rc; ++l)
        {
            WorkFragment *list = ps.work.p + listBegin[l];
            unsigned n = listCount[l];
            for (unsigned k = 0; k < n; ++k) adoptAlignment(list[k], ps.hFrag1.p[list[k].slot], 0, list[k].slot);
            n = consolidateDuplicateFragments(list, n, true);
            listCount[l] = n;
            if (!gapLimit || n < 2) continue;
            std::sort(list, list + n, [&](const WorkFragment &x, const WorkFragment &y) {           // orderByUnclippedPosition (:443-449)
                return x.f.contigId < y.f.contigId || (x.f.contigId == y.f.contigId && pools.unclippedPosition(x) < pools.unclippedPosition(y));
            });
            for (unsigned h = 0; h + 1 < n; ++h)
            {
                const WorkFragment &head = list[h], &tail = list[h + 1];
                if (head.f.contigId != tail.f.contigId || head.f.reverse != tail.f.reverse) continue;
                const isaac_ext_seed_t &headSeed = batch->seeds[head.f.firstSeedIndex], &tailSeed = batch->seeds[tail.f.firstSeedIndex];
                const long distance = pools.unclippedPosition(tail) - pools.unclippedPosition(head);
                if (!(std::labs(distance) < long(gapLimit))) continue;                              // :490
                const long readLength = ctx->reads.readLength[head.f.readIndex];
                const long headSeedOffset = head.f.reverse ? readLength - headSeed.offset - headSeed.length : headSeed.offset;   // :493-494
                const long tailSeedOffset = head.f.reverse ? readLength - tailSeed.offset - tailSeed.length : tailSeed.offset;
                auto side = [&](const WorkFragment &w, long seedOffset, unsigned seedLength) {
                    IndelSide s;
                    s.position = w.f.position; s.beginClipped = uint32_t(pools.beginClipped(w)); s.endClipped = uint32_t(pools.endClipped(w));
                    s.observedLength = w.f.cigarLength ? w.f.observedLength : 0;
                    s.smithWatermanScore = w.f.smithWatermanScore; s.mismatchCount = w.f.mismatchCount;
                    s.lowClipped = w.f.lowClipped; s.highClipped = w.f.highClipped;
                    s.seedOffset = uint32_t(seedOffset); s.seedLength = seedLength;
                    return s;
                };
                IndelTask task;
                std::memset(&task, 0, sizeof(task));
                task.readId = head.f.readId; task.contigId = head.f.contigId; task.reverse = head.f.reverse;
                if (0 < tailSeedOffset - headSeedOffset)
                {
                    // seeds ordered like the alignments: a deletion, patch the head (:497-503)
                    task.insertion = 0;
                    task.head = side(head, headSeedOffset, headSeed.length);
                    task.tail = side(tail, tailSeedOffset, tailSeed.length);
                }
                else
                {
                    // alignSimpleInsertion(*tail as head, ..., *head as tail) (:504-509); still patches list[h]
                    task.insertion = 1;
                    task.head = side(tail, tailSeedOffset, tailSeed.length);
                    task.tail = side(head, headSeedOffset, headSeed.length);
                }
                partTasks[t].push_back(task);
                partTargets[t].push_back(listBegin[l] + h);
            }
        }
    });
    timer.mark("P2 consolidate + pairing");
    std::vector<uint64_t> taskBegin(parts + 1, 0);
    for (unsigned p = 0; p < parts; ++p) taskBegin[p + 1] = taskBegin[p] + partTasks[p].size();
    const uint64_t nTasks = taskBegin[parts];
    if (nTasks)
    {
        CK(ps.hTasks.reserve(nTasks)); CK(ps.hIndel.reserve(nTasks));
        CK(ps.dTasks.reserve(nTasks)); CK(ps.dIndel.reserve(nTasks));
        for (unsigned p = 0; p < parts; ++p) std::copy(partTasks[p].begin(), partTasks[p].end(), ps.hTasks.p + taskBegin[p]);
        CK(cudaMemcpyAsync(ps.dTasks.p, ps.hTasks.p, nTasks * sizeof(IndelTask), cudaMemcpyHostToDevice, ctx->stream));
        simpleIndelKernel<<<gridFor(ctx, nTasks, 128, 8), 128, 0, ctx->stream>>>(ctx->ref, ctx->reads, ctx->sp, uint32_t(nTasks), ps.dTasks.p, ps.dIndel.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ps.hIndel.p, ps.dIndel.p, nTasks * sizeof(IndelResult), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ps.indelCigars.resize(nTasks * 5);
        for (uint64_t i = 0; i < nTasks; ++i) std::copy(ps.hIndel.p[i].cigar, ps.hIndel.p[i].cigar + 5, ps.indelCigars.begin() + i * 5);
        pools.pools[1] = ps.indelCigars.data();
    }

    timer.mark("simple indel kernel + copies");
    // ---- P3: apply the patches, consolidate (:184), pick the fragments for the gapped aligner (:190-200)
    const bool withGaps = batch->withGaps != 0;
    std::vector<std::vector<uint64_t>> gapTargets(parts);
    std::vector<uint64_t> gapBegin(parts + 1, 0);
    parallelRanges(T, nClusters, [&](unsigned t, size_t b, size_t e) {
        for (size_t i = 0; i < partTasks[t].size(); ++i)
        {
            const IndelResult &r = ps.hIndel.p[taskBegin[t] + i];
            if (r.accepted) adoptAlignment(ps.work.p[partTargets[t][i]], r.fragment, 1, uint32_t(taskBegin[t] + i));
        }
        for (size_t l = b * rc; l < e * rc; ++l)
        {
            WorkFragment *list = ps.work.p + listBegin[l];
            if (gapLimit) listCount[l] = consolidateDuplicateFragments(list, listCount[l], true);
            if (!withGaps) continue;
            for (unsigned k = 0; k < listCount[l]; ++k)
                if (ISAAC_EXT_SW_MISMATCH_CUTOFF < list[k].f.mismatchCount) gapTargets[t].push_back(listBegin[l] + k);
        }
        gapBegin[t + 1] = gapTargets[t].size();
    });
    timer.mark("P3 apply + consolidate");
    for (unsigned p = 0; p < parts; ++p) gapBegin[p + 1] += gapBegin[p];
    const uint64_t n3 = gapBegin[parts];
    if (n3)
    {
        CK(ps.hCand3.reserve(n3)); CK(ps.hFrag3.reserve(n3)); CK(ps.hCig3.reserve(n3 * GAPPED_STRIDE));
        parallelRanges(T, nClusters, [&](unsigned t, size_t, size_t) {
            for (size_t i = 0; i < gapTargets[t].size(); ++i)
            {
                const WorkFragment &w = ps.work.p[gapTargets[t][i]];
                // alignGapped starts from resetAlignment(): the unclipped position of the current alignment (GappedAligner.cpp:175)
                ps.hCand3.p[gapBegin[t] + i] = candidateOf(w.f, pools.unclippedPosition(w));
            }
        });
        rcode = runGapped(ctx, uint32_t(n3), ps.hCand3.p, GAPPED_STRIDE, ps.hFrag3.p, ps.hCig3.p);
        if (rcode) return rcode;
        pools.pools[2] = ps.hCig3.p;
    }

    timer.mark("K2 gapped + copies");
    // ---- P4: acceptance rule (:202-209), final consolidate (:213)
    parallelRanges(T, nClusters, [&](unsigned t, size_t b, size_t e) {
        for (size_t i = 0; i < gapTargets[t].size(); ++i)
        {
            WorkFragment &w = ps.work.p[gapTargets[t][i]];
            const isaac_ext_fragment_t &g = ps.hFrag3.p[gapBegin[t] + i];
            if (acceptGapped(w.f, g, ctx->cfg.gappedMismatchesMax)) adoptAlignment(w, g, 2, uint32_t(gapBegin[t] + i));
        }
        for (size_t l = b * rc; l < e * rc; ++l)
            if (listCount[l]) listCount[l] = consolidateDuplicateFragments(ps.work.p + listBegin[l], listCount[l], true);
    });
    timer.mark("P4 accept + consolidate");
    flatten(ctx, pools, lists, [&](size_t l) { return std::pair<const WorkFragment *, unsigned>(ps.work.p + listBegin[l], listCount[l]); });
    timer.mark("flatten");
    result->fragments = ps.outFragments.p; result->readFragmentBegin = ps.outBegin.p; result->cigars = ps.outCigars.p;
    result->built = ps.outFlags.data(); result->fragmentCount = ps.outFragmentCount; result->cigarWords = ps.outCigarWords;
    return ISAAC_EXT_OK;
}

namespace
{
/// TemplateLengthStatistics helpers (TemplateLengthStatistics.cpp:80-84,186-238; .hh:189-214): pure integer logic on
/// the two best alignment models.
struct TlsHost
{
    unsigned mateMin, mateMax, models[2];
    explicit TlsHost(const isaac_ext_tls_t &t)
    {
        mateMin = -1 == t.mateDriftRange ? t.min : t.median - t.mateDriftRange;
        mateMax = -1 == t.mateDriftRange ? t.max : t.median + t.mateDriftRange;
        models[0] = t.bestModel[0]; models[1] = t.bestModel[1];
    }
    static unsigned alignmentClass(unsigned m) { return m < 4 ? m : ((~m) & 3); }
    bool coherent() const { return models[0] < 8 && models[1] < 8 && models[0] != models[1] && alignmentClass(models[0]) == alignmentClass(models[1]); }
    bool validModel(bool reverse, unsigned readIndex) const
    {
        const unsigned shift = (readIndex + 1) % 2;
        return reverse == ((models[0] >> shift) & 1) || reverse == ((models[1] >> shift) & 1);
    }
    bool firstFragment(bool reverse, unsigned readIndex) const
    {
        const unsigned shift = (readIndex + 1) % 2;
        for (unsigned i = 0; i < 2; ++i) if (reverse == ((models[i] >> shift) & 1)) return ((models[i] >> 2) & 1) == readIndex;
        return false;
    }
    bool mateOrientation(unsigned readIndex, bool reverse) const
    {
        const unsigned shift = (readIndex + 1) % 2;
        for (unsigned i = 0; i < 2; ++i) if (reverse == ((models[i] >> shift) & 1)) return (models[i] >> readIndex) & 1;
        return (models[0] >> readIndex) & 1;
    }
    long mateMinPosition(unsigned readIndex, bool reverse, long position, const uint32_t *len) const
    {
        if (!validModel(reverse, readIndex)) return position;
        return firstFragment(reverse, readIndex) ? position + long(mateMin) - long(len[(readIndex + 1) % 2]) : position - long(mateMax) + long(len[readIndex]);
    }
    long mateMaxPosition(unsigned readIndex, bool reverse, long position, const uint32_t *len) const
    {
        if (!validModel(reverse, readIndex)) return position;
        return firstFragment(reverse, readIndex) ? position + long(mateMax) - long(len[(readIndex + 1) % 2]) : position - long(mateMin) + long(len[readIndex]);
    }
};
const unsigned SHADOW_LIST_CAPACITY = 1000;      // TemplateBuilder::TRACKED_REPEATS_MAX_ONE_READ (TemplateBuilder.hh:145, .cpp:82)
} // namespace

/// exclusive prefix sum of n 32-bit counts; out[n] = total (cub::DeviceScan over n + 1 items, the last input is ignored)
static cudaError_t exclusiveSum(isaac_ext_ctx *ctx, uint32_t *counts, uint32_t *out, uint32_t n)
{
    PipelineState &ps = ctx->pipeline;
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, counts, out, int(n) + 1, ctx->stream);
    if (e != cudaSuccess) return e;
    e = ps.dScanTemp.reserve(bytes + 16);
    if (e != cudaSuccess) return e;
    ++ctx->launches;
    return cub::DeviceScan::ExclusiveSum(ps.dScanTemp.p, bytes, counts, out, int(n) + 1, ctx->stream);
}

/// isaac_ext_rescue_shadows with its flat result in result set 'slot' (0 or 1) of the context
static int rescueShadowsInto(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t n, const isaac_ext_rescue_request_t *requests,
                             isaac_ext_rescue_result_t *result, const unsigned slot)
{
    if (!ctx) return ISAAC_EXT_E_INVALID_ARG;
    if (!ctx->haveReference || !ctx->haveReads) return ctx->fail(ISAAC_EXT_E_NO_REFERENCE, "set_reference / set_reads first");
    if (!tls || !result || (n && !requests)) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "null argument");
    if (ctx->reads.readCount != 2) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "shadow rescue needs paired reads (ShadowAligner.cpp:170)");
    CK(cudaSetDevice(ctx->device));
    PipelineState &ps = ctx->pipeline;
    const unsigned T = ctx->hostThreads;
    const TlsHost stats(*tls);
    uint32_t fragmentTotal = 0, wordTotal = 0;
    PhaseTimer timer("rescue");
    if (n && stats.coherent())                                                   // :164-168
    {
        // ---- R1: rescue windows (calculateShadowRescueRange :119-149, rescueShadow :170-198)
        CK(ps.hShadowTasks.reserve(n));
        std::atomic<int> bad(0);
        parallelRanges(T, n, [&](unsigned, size_t b, size_t e) {
            for (size_t i = b; i < e; ++i)
            {
                const isaac_ext_rescue_request_t &q = requests[i];
                ShadowTask &task = ps.hShadowTasks.p[i];
                const unsigned contigId = q.orphanContigStrand >> 1;
                if (q.orphanReadId >= ctx->reads.readTotal || contigId >= ctx->ref.contigCount) { bad = 1; task = ShadowTask{0, 0, 0, 0}; continue; }
                const unsigned orphanReadIndex = q.orphanReadId % 2;
                const bool orphanReverse = q.orphanContigStrand & 1;
                const unsigned shadowReadIndex = (orphanReadIndex + 1) % 2;
                const uint32_t *len = ctx->reads.readLength;
                long shadowMin = stats.mateMinPosition(orphanReadIndex, orphanReverse, q.orphanPosition, len);
                long shadowMax = stats.mateMaxPosition(orphanReadIndex, orphanReverse, q.orphanPosition, len) + long(len[shadowReadIndex]) - 1;
                if (q.bestTemplateLength)
                {
                    const long fStrand = q.orphanPosition;                                            // FragmentMetadata.hh:90-95
                    const long rStrand = std::max(q.orphanPosition + long(q.orphanObservedLength), 1L) - 1;   // :97-103
                    if (shadowMin < fStrand) shadowMin = std::min(rStrand - q.bestTemplateLength, shadowMin);
                    if (shadowMax > fStrand) shadowMax = std::max(fStrand + q.bestTemplateLength, shadowMax);
                }
                const long first = shadowMin - 10, second = shadowMax + 10;                            // :147
                task.shadowReadId = q.orphanReadId - orphanReadIndex + shadowReadIndex;
                task.contigStrand = (contigId << 1) | (stats.mateOrientation(orphanReadIndex, orphanReverse) ? 1u : 0u);
                task.windowBegin = std::max(0L, first);                                                // :194
                task.windowEnd = std::min(long(ctx->contigLength[contigId]), second + 1);              // :197
                if (second < first || second + 1 + long(len[shadowReadIndex]) < 0) task.windowEnd = task.windowBegin;   // :179-190
            }
        });
        timer.mark("R1 windows");
        if (bad) return ctx->fail(ISAAC_EXT_E_INVALID_ARG, "rescue request refers to an unknown read or contig");
        // ---- K5: candidate positions of every request (:195-236).  Requests with a small window (nearly all) take one warp
        // each, the others one CTA each with the full 4^7 table and the 10000 cap.  Both kernels always get the whole request
        // list and decide per request with the one predicate shadowTaskIsSmall (kernels_shadow.cuh), so every request is
        // handled by exactly one of them whatever the mix of window sizes and read lengths.
        const unsigned grid = std::max(1u, std::min<unsigned>(n, unsigned(ctx->smCount) * 6));
        CK(ps.dShadowTasks.reserve(n)); CK(ps.dTaskBegin.reserve(n)); CK(ps.dTaskCount.reserve(n)); CK(ps.dPoolSize.reserve(1));
        CK(ps.dShadowScratch.reserve(size_t(grid) * SHADOW_SCRATCH));
        CK(cudaMemcpyAsync(ps.dShadowTasks.p, ps.hShadowTasks.p, size_t(n) * sizeof(ShadowTask), cudaMemcpyHostToDevice, ctx->stream));
        uint64_t capacity = std::max<uint64_t>(ps.dCand.capacity, uint64_t(n) * 24 + 4096);
        unsigned long long poolSize64 = 0;
        for (int attempt = 0; attempt < 2; ++attempt)
        {
            CK(ps.dCand.reserve(capacity));
            CK(cudaMemsetAsync(ps.dPoolSize.p, 0, sizeof(unsigned long long), ctx->stream));
            // a request neither kernel takes cannot exist, but an empty list is the safe reading of a skipped one
            CK(cudaMemsetAsync(ps.dTaskBegin.p, 0, size_t(n) * sizeof(uint32_t), ctx->stream));
            CK(cudaMemsetAsync(ps.dTaskCount.p, 0, size_t(n) * sizeof(uint32_t), ctx->stream));
            const uint32_t poolCapacity = uint32_t(std::min<uint64_t>(ps.dCand.capacity, 0xFFFFFFFFull));
            {
                shadowCandidatesWarpKernel<<<gridFor(ctx, uint64_t(n) * 32, SHADOW_WARPS * 32, 6), SHADOW_WARPS * 32, 0, ctx->stream>>>(
                    ctx->ref, ctx->reads, n, ps.dShadowTasks.p, ps.dCand.p, poolCapacity, ps.dPoolSize.p, ps.dTaskBegin.p, ps.dTaskCount.p,
                    ctx->errorFlag.p);
                ++ctx->launches;
            }
            {
                shadowCandidatesKernel<<<grid, SHADOW_BLOCK, 0, ctx->stream>>>(ctx->ref, ctx->reads, n, ps.dShadowTasks.p, ps.dShadowScratch.p,
                                                                                ps.dCand.p, poolCapacity, ps.dPoolSize.p, ps.dTaskBegin.p,
                                                                                ps.dTaskCount.p, ctx->errorFlag.p, true);
                ++ctx->launches;
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&poolSize64, ps.dPoolSize.p, sizeof(poolSize64), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (poolSize64 <= poolCapacity) break;
            CK(cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream));   // bit 2: the pool was too small, the counter holds the need
            if (attempt || poolSize64 > 0xFFFFFFF0ull)
                return ctx->fail(ISAAC_EXT_E_CAPACITY, "too many shadow candidate positions in one rescue batch: split the batch");
            capacity = poolSize64 + 1024;
        }
        const uint32_t poolSize = uint32_t(poolSize64);
        timer.mark("K5 shadow candidates");
        if (poolSize)
        {
            // ---- K1: UngappedAligner::alignUngapped of every candidate position (:205-236)
            CK(ps.dFrag.reserve(poolSize)); CK(ps.dCig.reserve(size_t(poolSize) * 3));
            const uint32_t *clip = nullptr;
            if (ctx->adapters.count)
            {
                CK(ps.dAdapterFirst.reserve(n)); CK(ps.dSlot.reserve(poolSize));
                shadowAdapterSlotsKernel<<<gridFor(ctx, uint64_t(n) * 32, 128, 16), 128, 0, ctx->stream>>>(
                    n, ps.dTaskBegin.p, ps.dTaskCount.p, ps.dCand.p, ps.dAdapterFirst.p, ps.dSlot.p);
                ++ctx->launches;
                CK(cudaGetLastError());
                int rca = initAdapterSlots(ctx, n, ps.dAdapterFirst.p);
                if (!rca) rca = adapterSlotClip(ctx, poolSize, ps.dCand.p, ps.dSlot.p, ctx->stream, &clip);
                if (rca) return rca;
            }
            const int rc = ungappedDevice(ctx, poolSize, ps.dCand.p, ps.dFrag.p, ps.dCig.p, nullptr, ctx->stream, clip);
            if (rc) return rc;
        }
        // ---- R2 on the device: shadow lists, best shadow, neighbours to gap-align (:205-256)
        const unsigned rgrid = gridFor(ctx, n, 128, 16);
        CK(ps.dKept.reserve(size_t(poolSize) + 1)); CK(ps.dAdoptedBy.reserve(size_t(poolSize) + 1)); CK(ps.dListState.reserve(n));
        CK(ps.dCounts.reserve(size_t(n) * 3 + 1)); CK(ps.dBegins.reserve(size_t(n) * 3 + 3)); CK(ps.dRescued.reserve(n)); CK(ps.hTotals.reserve(4));
        uint32_t *gapCounts = ps.dCounts.p, *listCounts = ps.dCounts.p + n, *wordCounts = ps.dCounts.p + 2 * size_t(n);
        uint32_t *gapBegin = ps.dBegins.p, *fragmentBegin = ps.dBegins.p + (n + 1), *wordBegin = ps.dBegins.p + 2 * (size_t(n) + 1);
        shadowSelectKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dTaskCount.p, ps.dFrag.p, ps.dKept.p, ps.dListState.p, gapCounts);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(exclusiveSum(ctx, gapCounts, gapBegin, n));
        CK(cudaMemcpyAsync(ps.hTotals.p, gapBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const uint32_t n3 = ps.hTotals.p[0];
        timer.mark("K1 ungapped + R2 lists");
        if (n3)
        {
            // ---- K2: the gapped aligner on the neighbours (:249-262), candidates written in list order by the device
            CK(ps.dCand3.reserve(n3)); CK(ps.dFrag3.reserve(n3)); CK(ps.dCig3.reserve(size_t(n3) * GAPPED_STRIDE));
            CK(ps.dSlot3.reserve(n3)); CK(ps.dSources.reserve(n3));
            shadowGapKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dListState.p, gapBegin, ps.dFrag.p, ps.dCig.p, ps.dKept.p,
                                                            ps.dCand3.p, ps.dSlot3.p, ps.dSources.p);
            ++ctx->launches;
            CK(cudaGetLastError());
            const uint32_t *clip = nullptr;
            int rc = adapterSlotClip(ctx, n3, ps.dCand3.p, ps.dSlot3.p, ctx->stream, &clip);
            if (!rc) rc = gappedDevice(ctx, n3, ps.dCand3.p, GAPPED_STRIDE, ps.dFrag3.p, ps.dCig3.p, nullptr, ctx->stream, clip);
            if (rc) return rc;
        }
        // ---- R3 on the device: acceptance in list order, best shadow first (:255-290), then the flat result
        shadowAcceptKernel<<<rgrid, 128, 0, ctx->stream>>>(n, ps.dTaskBegin.p, ps.dListState.p, gapBegin, ps.dSources.p, ps.dFrag.p, ps.dFrag3.p,
                                                           ctx->cfg.gappedMismatchesMax, ps.dKept.p, ps.dAdoptedBy.p, listCounts, wordCounts,
                                                           ps.dRescued.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        CK(exclusiveSum(ctx, listCounts, fragmentBegin, n));
        CK(exclusiveSum(ctx, wordCounts, wordBegin, n));
        CK(cudaMemcpyAsync(ps.hTotals.p + 1, fragmentBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ps.hTotals.p + 2, wordBegin + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        uint32_t flag = 0;
        CK(cudaMemcpyAsync(&flag, ctx->errorFlag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (flag)
        {
            CK(cudaMemsetAsync(ctx->errorFlag.p, 0, sizeof(uint32_t), ctx->stream));
            return ctx->fail(ISAAC_EXT_E_CAPACITY, "a gapped CIGAR did not fit the cigar stride");
        }
        fragmentTotal = ps.hTotals.p[1]; wordTotal = ps.hTotals.p[2];
        timer.mark("K2 gapped + R3 accept");
        CK(ps.dOutFragments.reserve(fragmentTotal + 1)); CK(ps.dOutCigars.reserve(wordTotal + 1)); CK(ps.dOutBegin.reserve(size_t(n) + 1));
        CK(ps.hOutFragments[slot].reserve(fragmentTotal + 1)); CK(ps.hOutCigars[slot].reserve(wordTotal + 1)); CK(ps.hOutBegin[slot].reserve(size_t(n) + 1));
        CK(ps.hRescued[slot].reserve(n));
        shadowFlattenKernel<<<gridFor(ctx, uint64_t(n) * 32, 128, 16), 128, 0, ctx->stream>>>(
            n, ps.dTaskBegin.p, listCounts, fragmentBegin, wordBegin, ps.dKept.p, ps.dAdoptedBy.p, ps.dFrag.p, ps.dCig.p, ps.dFrag3.p, ps.dCig3.p,
            GAPPED_STRIDE, ps.dOutFragments.p, ps.dOutCigars.p, ps.dOutBegin.p);
        ++ctx->launches;
        CK(cudaGetLastError());
        if (fragmentTotal) CK(cudaMemcpyAsync(ps.hOutFragments[slot].p, ps.dOutFragments.p, size_t(fragmentTotal) * sizeof(isaac_ext_fragment_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (wordTotal) CK(cudaMemcpyAsync(ps.hOutCigars[slot].p, ps.dOutCigars.p, size_t(wordTotal) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ps.hOutBegin[slot].p, ps.dOutBegin.p, size_t(n) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ps.hRescued[slot].p, ps.dRescued.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ps.hOutBegin[slot].p[n] = fragmentTotal;
        timer.mark("flatten + copies");
        result->fragments = ps.hOutFragments[slot].p; result->requestFragmentBegin = ps.hOutBegin[slot].p; result->cigars = ps.hOutCigars[slot].p;
        result->rescued = ps.hRescued[slot].p; result->fragmentCount = fragmentTotal; result->cigarWords = wordTotal;
        return ISAAC_EXT_OK;
    }
    // nothing to rescue (no requests, or template length statistics without a coherent pair of models, :164-168)
    CK(ps.hOutBegin[slot].reserve(size_t(n) + 1)); CK(ps.hRescued[slot].reserve(size_t(n) + 1));
    std::fill(ps.hOutBegin[slot].p, ps.hOutBegin[slot].p + n + 1, uint64_t(0));
    std::fill(ps.hRescued[slot].p, ps.hRescued[slot].p + n, uint8_t(0));
    result->fragments = ps.hOutFragments[slot].p; result->requestFragmentBegin = ps.hOutBegin[slot].p; result->cigars = ps.hOutCigars[slot].p;
    result->rescued = ps.hRescued[slot].p; result->fragmentCount = 0; result->cigarWords = 0;
    return ISAAC_EXT_OK;
}

extern "C" int isaac_ext_rescue_shadows(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t n,
                                        const isaac_ext_rescue_request_t *requests, isaac_ext_rescue_result_t *result)
{
    return rescueShadowsInto(ctx, tls, n, requests, result, 0);
}
