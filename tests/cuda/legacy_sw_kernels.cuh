// Earlier generations of the Smith-Waterman kernels, kept for reference only: NOT compiled into libisaac_ext.so and not
// reachable from the product (round 1 could still select them with an environment variable; that dispatch is gone).
//   gappedKernel        scalar, one alignment per thread, forward + traceback + re-score in one kernel  (454 GCUPS)
//   bandedSwAsciiKernel scalar BandedSmithWaterman::align on explicit strings
//   gappedKernel2       packed 16x2, two alignments per thread, fused                                   (740 GCUPS)
// The product runs swForwardKernel + swTraceScoreKernel (csrc/kernels3.cuh) and bandedSwAsciiKernel2 (csrc/kernels2.cuh).
// To build them again: include this file behind csrc/kernels2.cuh.
#pragma once
namespace isaac_b200
{

/// \param src       src.q(i) = code of query base i, src.d(k) = code of database base k (k < L + 15)
/// \param ops       thread-local buffer receiving the CIGAR in final (head first) order
/// \return stripped leading deletion length (return value of BandedSmithWaterman::align)
template <class BaseSrc>
__device__ __forceinline__ unsigned bandedSwAlign(const BaseSrc &src, const unsigned L, const SwScores s,
                                                   uint32_t *__restrict__ tb, const size_t tbStride,
                                                   uint32_t *ops, const unsigned cap, unsigned &nOps, bool &overflow)
{
    int G[16], E[16], F[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { G[j] = s.init; E[j] = s.init; F[j] = 0; }   // :108-114, F really starts at 0
    G[0] = 0;                                                                  // :115

    // W: nibble j = database code seen by lane j, i.e. db[i + 15 - j] (:117-122, :202-203)
    unsigned long long W = 0;
#pragma unroll 1
    for (unsigned k = 0; k < 15; ++k) W = (W << 4) | src.d(k);

#pragma unroll 1
    for (unsigned i = 0; i < L; ++i)
    {
        W = (W << 4) | src.d(i + 15);
        const unsigned qc = src.q(i);
        unsigned TG = 0, TE = 0, TF = 0;
        int cg = s.init, ce = s.init, cf = s.init;     // E carries from lane j+1 (:248-250)
        unsigned tgEhi = 0, tgFhi = 0;
#pragma unroll
        for (int j = 15; j >= 0; --j)
        {
            // ---- F: insertion, from lane j-1 of the previous row; zeros are shifted into lane 0 (:132-173)
            const int gp = j ? G[j - 1] : 0, ep = j ? E[j - 1] : 0, fp = j ? F[j - 1] : 0;
            unsigned tf = gp < ep ? 1u : 0u;
            const int a = max(gp, ep) - s.open;
            const int b = fp - s.ext;
            if (a < b) tf = 2u;                        // _mm_max_epu8: 2 overrides 1 (:166)
            int nF = max(a, b);
            if (j == 0) { tf = 0u; nF = s.init; }      // :167, :173
            // ---- G: diagonal, from the same lane of the previous row (:176-190)
            const unsigned tgE = G[j] < E[j] ? 1u : 0u;
            int g = max(G[j], E[j]);
            const unsigned tgF = g < F[j] ? 2u : 0u;
            g = max(g, F[j]);
            const unsigned dc = unsigned(W >> (4 * j)) & 0xFu;
            const int nG = g + (qc != dc ? s.mismatch : s.match);                  // raw compare (:200-205, :230-244)
            // ---- direction of G: _mm_max_epi16 applied to BYTE pairs (:197) -> lanes (2p, 2p+1) are coupled
            unsigned tg;
            if (j & 1) { tgEhi = tgE; tgFhi = tgF; tg = tgF ? 2u : tgE; }
            else { tg = tgFhi ? tgF : (tgEhi ? tgE : max(tgF, tgE)); }
            // ---- E: deletion, serial from lane 15 down (:261-297)
            int nE; unsigned te;
            if (ce > cg && ce > cf) { nE = ce; te = 1u; }
            else if (cf > cg) { nE = cf; te = 2u; }
            else { nE = cg; te = 0u; }
            cg = nG - s.open; ce = nE - s.ext; cf = nF - s.open;
            G[j] = nG; E[j] = nE; F[j] = nF;
            TG |= tg << (2 * j); TE |= te << (2 * j); TF |= tf << (2 * j);
        }
        uint32_t *row = tb + size_t(i) * 3 * tbStride;                             // :306-308
        row[0] = TG; row[tbStride] = TE; row[2 * tbStride] = TF;
    }

    // ---- end cell: lanes 15..0, matrices G,E,F in that order, strict '>' (:349-379)
    int best = G[15] - 1;
    int ii = int(L) - 1, jj = ii;
    unsigned type = 0;
#pragma unroll
    for (int j = 15; j >= 0; --j)
    {
        if (G[j] > best) { best = G[j]; jj = j; type = 0; }
        if (E[j] > best) { best = E[j]; jj = j; type = 1; }
        if (F[j] > best) { best = F[j]; jj = j; type = 2; }
    }

    // ---- traceback, operations come out tail first (:381-435); written from the back of ops[]
    unsigned w = cap;   // next free slot is ops[w-1]
    overflow = false;
    auto push = [&](unsigned length, unsigned type3) {
        // type3: 0 ALIGN, 1 DELETE, 2 INSERT (opCodes[] :383)
        const uint32_t op = type3 == 0 ? ISAAC_EXT_CIGAR_ALIGN : (type3 == 1 ? ISAAC_EXT_CIGAR_DELETE : ISAAC_EXT_CIGAR_INSERT);
        if (w == 0) { overflow = true; return; }
        ops[--w] = cigarWord(length, op);
    };
    unsigned opLength = 0;
    if (jj > 0) push(jj, 1);
    while (ii >= 0 && jj >= 0 && jj <= 15)
    {
        ++opLength;
        const unsigned next = (tb[(size_t(ii) * 3 + type) * tbStride] >> (2 * jj)) & 3u;
        if (next != type) { push(opLength, type); opLength = 0; }
        if (type == 0) { --ii; } else if (type == 1) { ++jj; } else { --ii; --jj; }
        type = next;
    }
    if (type != 1 && opLength) { push(opLength, type); opLength = 0; }
    if (jj < 15) { push(opLength + 15 - jj, 1); opLength = 0; }

    // ---- ops[w..cap) is now head first.  Strip a deletion at the start (its length is returned) and one at the
    //      end (:437-453).
    unsigned ret = 0;
    unsigned e = cap;
    if (w < e && (ops[w] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { ret = ops[w] >> 4; ++w; }
    if (w < e && (ops[e - 1] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { --e; }
    nOps = e - w;
    // compact to the front
    for (unsigned k = 0; k < nOps; ++k) ops[k] = ops[w + k];
    return ret;
}


struct ResidentBaseSrc
{
    const ReferenceView &ref; const ReadSetView &reads;
    unsigned readId, L; bool reverse; unsigned qBegin; uint64_t dBegin;
    __device__ __forceinline__ unsigned q(unsigned i) const { unsigned qq; return reads.code(readId, L, reverse, qBegin + i, qq); }
    __device__ __forceinline__ unsigned d(unsigned k) const { return ref.code(dBegin + k); }
};


/// K2+K4: one candidate per thread: clip, banded Smith-Waterman, traceback, re-score the gapped CIGAR.
__global__ void gappedKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, uint32_t n,
                             const isaac_ext_candidate_t *__restrict__ candidates, uint32_t cigarStride,
                             isaac_ext_fragment_t *__restrict__ fragments, uint32_t *__restrict__ cigars,
                             uint64_t *__restrict__ masks, uint32_t *__restrict__ tbScratch, uint32_t *__restrict__ errorFlag,
                             const uint32_t *__restrict__ adapterClip = nullptr)
{
    __shared__ double tables[201];
    const ScoreParams sp = stageScoreTables(spGlobal, tables);
    const size_t tbStride = size_t(gridDim.x) * blockDim.x;
    uint32_t *tb = tbScratch + (blockIdx.x * blockDim.x + threadIdx.x);
    const SwScores sw = {sp.swMatch, sp.swMismatch, sp.swOpen, sp.swExtend, -32768 + sp.swOpen};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const isaac_ext_candidate_t c = candidates[i];
        isaac_ext_fragment_t o;
        initFragment(o, c, reads.readCount);
        const unsigned contigId = c.contigStrand >> 1;
        const unsigned L = reads.length(c.readId);
        const long contigLength = long(ref.contigLength[contigId]);
        uint64_t *mask = masks ? masks + size_t(i) * ISAAC_EXT_MASK_WORDS : nullptr;
        if (mask) for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) mask[k] = 0;
        uint32_t *cigar = cigars + size_t(i) * cigarStride;
        o.cigarOffset = i * cigarStride;
        FragmentState f = {c.position, 0u, 0u, bool(c.contigStrand & 1u)};              // GappedAligner.cpp:175-176
        long begin = 0, end = L;
        if (adapterClip) applyAdapterClip(adapterClip[i], L, f, begin, end);            // :186
        clipReadMasking(L, reads.endCyclesMasked[c.readId], f, begin, end);             // :187
        clipReference(contigLength, f, begin, end);                                     // :189
        o.lowClipped = uint16_t(f.lowClipped); o.highClipped = uint16_t(f.highClipped); o.position = f.position;
        const unsigned sequenceLength = unsigned(end - begin);
        long strandPosition = f.position;
        // no gapped alignment if the reference is too short (:204-208)
        if (sequenceLength && !(contigLength < long(sequenceLength) + strandPosition + 16) &&
            !(adapterClip && (adapterClip[i] >> 31)))                                    // --avoid-smith-waterman (:218-226)
        {
            // getFlanks (:51-82)
            unsigned left, right;
            if (strandPosition >= 8)
            {
                if (strandPosition + sequenceLength + 8 < contigLength) { left = 8; right = 7; }
                else { right = unsigned(contigLength - sequenceLength - strandPosition); left = 16 - right - 1; }
            }
            else { left = unsigned(strandPosition); right = 16 - left - 1; }
            (void)right;
            const ResidentBaseSrc src = {ref, reads, c.readId, L, f.reverse, unsigned(begin),
                                         ref.contigOffset[contigId] + uint64_t(strandPosition - left)};
            uint32_t ops[SW_OPS_CAP + 2];
            unsigned nSw = 0; bool overflow = false;
            unsigned nOps = 0;
            uint32_t *swOps = ops + 1;
            const unsigned ret = bandedSwAlign(src, sequenceLength, sw, tb, tbStride, swOps, SW_OPS_CAP, nSw, overflow);   // :231
            uint32_t *all = swOps;
            nOps = nSw;
            if (begin) { ops[0] = cigarWord(uint32_t(begin), ISAAC_EXT_CIGAR_SOFT_CLIP); all = ops; ++nOps; }           // :191-195
            if (long(L) - end) all[nOps++] = cigarWord(uint32_t(L - end), ISAAC_EXT_CIGAR_SOFT_CLIP);                   // :233-237
            strandPosition += long(ret) - long(left);                                                                    // :231,240
            if (overflow || nOps > cigarStride) { atomicOr(errorFlag, 1u); }
            else
            {
                const unsigned matchCount = scoreCigar(ref, reads, sp, c.readId, L, f.reverse, ref.contigOffset[contigId],
                                                       strandPosition, all, nOps, o, mask);                              // :245
                for (unsigned k = 0; k < nOps; ++k) cigar[k] = all[k];
                o.cigarLength = uint16_t(nOps);
                (void)matchCount;
            }
        }
        fragments[i] = o;
    }
}


/// BandedSmithWaterman::align on explicit strings (unit parity with testBandedSmithWaterman.cpp and kernel timing).
__global__ void bandedSwAsciiKernel(uint32_t n, const unsigned char *__restrict__ queries, const uint64_t *__restrict__ queryOffsets,
                                    const uint32_t *__restrict__ queryLengths, const unsigned char *__restrict__ databases,
                                    const uint64_t *__restrict__ databaseOffsets, const SwScores sw, uint32_t cigarStride,
                                    uint32_t *__restrict__ cigars, uint32_t *__restrict__ cigarLengths, uint32_t *__restrict__ offsets,
                                    uint32_t *__restrict__ tbScratch, uint32_t *__restrict__ errorFlag)
{
    const size_t tbStride = size_t(gridDim.x) * blockDim.x;
    uint32_t *tb = tbScratch + (blockIdx.x * blockDim.x + threadIdx.x);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const AsciiBaseSrc src = {queries + queryOffsets[i], databases + databaseOffsets[i]};
        uint32_t ops[SW_OPS_CAP];
        unsigned nOps = 0; bool overflow = false;
        const unsigned ret = bandedSwAlign(src, queryLengths[i], sw, tb, tbStride, ops, SW_OPS_CAP, nOps, overflow);
        if (overflow) atomicOr(errorFlag, 1u);
        offsets[i] = ret;
        cigarLengths[i] = nOps;
        for (unsigned k = 0; k < nOps && k < cigarStride; ++k) cigars[size_t(i) * cigarStride + k] = ops[k];
    }
}


#ifndef ISAAC_SW2_MIN_BLOCKS
#define ISAAC_SW2_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, ISAAC_SW2_MIN_BLOCKS)
gappedKernel2(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, uint32_t n,
              const isaac_ext_candidate_t *__restrict__ candidates, uint32_t cigarStride,
              isaac_ext_fragment_t *__restrict__ fragments, uint32_t *__restrict__ cigars,
              uint64_t *__restrict__ masks, uint32_t *__restrict__ tbScratch, uint32_t *__restrict__ errorFlag,
              const uint32_t *__restrict__ adapterClip = nullptr)
{
    // the two 100-entry log-probability tables are looked up once per base: keep them in shared memory
    __shared__ double tables[201];      // [0,100) logMatch, [100,200) logMismatch, [200] = 0.0 (contiguous in global too)
    for (unsigned i = threadIdx.x; i < 201; i += blockDim.x) tables[i] = spGlobal.logMatch[i];
    __syncthreads();
    ScoreParams sp = spGlobal;
    sp.logMatch = tables; sp.logMismatch = tables + 100;

    const size_t tbStride = size_t(gridDim.x) * blockDim.x;
    uint32_t *tb = tbScratch + (blockIdx.x * blockDim.x + threadIdx.x);
    const SwScores sw = {sp.swMatch, sp.swMismatch, sp.swOpen, sp.swExtend, -32768 + sp.swOpen};
    const uint32_t pairs = (n + 1) / 2;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += gridDim.x * blockDim.x)
    {
        const uint32_t iA = 2 * t, iB = 2 * t + 1;
        const bool haveB = iB < n;
        const GappedPrep pa = prepareGapped(ref, reads, candidates[iA], adapterClip, iA);
        GappedPrep pb = prepareGapped(ref, reads, candidates[haveB ? iB : iA], adapterClip, haveB ? iB : iA);
        if (!haveB) pb.run = false;
        const unsigned LA = pa.run ? pa.sequenceLength : 0u, LB = pb.run ? pb.sequenceLength : 0u;
        int jj[2] = {0, 0}; unsigned type[2] = {0, 0};
        if (LA | LB)
        {
            // a half that is not aligned (run == false) streams its own first bases: harmless, never stored
            ResidentPairSrc src = {ref,
                                   {reads.strandCodes(pa.c.readId, pa.f.reverse), reads.strandCodes(pb.c.readId, pb.f.reverse)},
                                   {LA ? unsigned(pa.begin) : 0u, LB ? unsigned(pb.begin) : 0u},
                                   {LA ? ref.contigOffset[pa.contigId] + uint64_t(pa.strandPosition - long(pa.left)) : 0ull,
                                    LB ? ref.contigOffset[pb.contigId] + uint64_t(pb.strandPosition - long(pb.left)) : 0ull},
                                   {0, 0}, {0, 0}, reads.codesClamp()};
            sw2Forward(src, LA, LB, sw, tb, tbStride, jj, type);                                                 // :231
        }
        // ---- traceback of both halves in one pass over the rows
        uint32_t opsA[SW_OPS_CAP + 2], opsB[SW_OPS_CAP + 2];
        Sw2Walker wa, wb;
        wa.start(LA, jj[0], type[0], opsA + 1, SW_OPS_CAP);
        wb.start(LB, jj[1], type[1], opsB + 1, SW_OPS_CAP);
        sw2TracebackPair(tb, tbStride, wa, wb);
        unsigned nSwA = 0, nSwB = 0;
        const unsigned retA = LA ? wa.finish(nSwA) : 0u, retB = LB ? wb.finish(nSwB) : 0u;
        // ---- soft clips, position (:233-240) and updateFragmentCigar of both halves side by side (:245)
        unsigned nOpsA = 0, nOpsB = 0;
        uint32_t *allA = assembleGappedCigar(pa, opsA, nSwA, nOpsA), *allB = assembleGappedCigar(pb, opsB, nSwB, nOpsB);
        const long posA = pa.strandPosition + long(retA) - long(pa.left), posB = pb.strandPosition + long(retB) - long(pb.left);
        const bool okA = pa.run && !(wa.overflow || nOpsA > cigarStride), okB = haveB && pb.run && !(wb.overflow || nOpsB > cigarStride);
        if ((pa.run && !okA) || (haveB && pb.run && !okB)) atomicOr(errorFlag, 1u);
        uint64_t *maskA = masks ? masks + size_t(iA) * ISAAC_EXT_MASK_WORDS : nullptr;
        uint64_t *maskB = masks && haveB ? masks + size_t(iB) * ISAAC_EXT_MASK_WORDS : nullptr;
        if (maskA) for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) maskA[k] = 0;
        if (maskB) for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) maskB[k] = 0;
        CigarScorer sa, sb;
        sa.start(ref, reads, sp, pa.c.readId, okA ? pa.L : 0u, pa.f.reverse, ref.contigOffset[pa.contigId], posA, allA, nOpsA, maskA);
        sb.start(ref, reads, sp, pb.c.readId, okB ? pb.L : 0u, pb.f.reverse, ref.contigOffset[pb.contigId], posB, allB, nOpsB, maskB);
        const unsigned Lmax = max(sa.L, sb.L);
        for (unsigned w = 0; w * 16u < Lmax; ++w) { sa.stepWord(w); sb.stepWord(w); }
        {
            isaac_ext_fragment_t o;
            initFragment(o, pa.c, reads.readCount);
            o.cigarOffset = iA * cigarStride;
            o.lowClipped = uint16_t(pa.f.lowClipped); o.highClipped = uint16_t(pa.f.highClipped); o.position = pa.f.position;
            if (okA)
            {
                sa.finish(o);
                o.position = posA;
                for (unsigned k = 0; k < nOpsA; ++k) cigars[size_t(iA) * cigarStride + k] = allA[k];
                o.cigarLength = uint16_t(nOpsA);
            }
            fragments[iA] = o;
        }
        if (haveB)
        {
            isaac_ext_fragment_t o;
            initFragment(o, pb.c, reads.readCount);
            o.cigarOffset = iB * cigarStride;
            o.lowClipped = uint16_t(pb.f.lowClipped); o.highClipped = uint16_t(pb.f.highClipped); o.position = pb.f.position;
            if (okB)
            {
                sb.finish(o);
                o.position = posB;
                for (unsigned k = 0; k < nOpsB; ++k) cigars[size_t(iB) * cigarStride + k] = allB[k];
                o.cigarLength = uint16_t(nOpsB);
            }
            fragments[iB] = o;
        }
    }
}

} // namespace isaac_b200
