// Second-generation Smith-Waterman kernels: two alignments per thread in packed 16x2 form (sw2.cuh).
//   bandedSwAsciiKernel2  BandedSmithWaterman::align on explicit strings, pairs (2t, 2t+1)
#pragma once
#include "kernels.cuh"
#include "sw2.cuh"

namespace isaac_b200
{

/// What GappedAligner::alignGapped derives for one candidate before it calls the Smith-Waterman (GappedAligner.cpp:174-215).
struct GappedPrep
{
    isaac_ext_candidate_t c;
    FragmentState f;
    long begin, end, strandPosition;
    unsigned L, sequenceLength, left, contigId;
    bool run;
};

__device__ __forceinline__ GappedPrep prepareGapped(const ReferenceView &ref, const ReadSetView &reads,
                                                    const isaac_ext_candidate_t c,
                                                    const uint32_t *__restrict__ adapterClip = nullptr, const uint32_t index = 0)
{
    GappedPrep p;
    p.c = c;
    p.contigId = c.contigStrand >> 1;
    p.L = reads.length(c.readId);
    const long contigLength = long(ref.contigLength[p.contigId]);
    p.f = FragmentState{c.position, 0u, 0u, bool(c.contigStrand & 1u)};          // :175-176
    p.begin = 0; p.end = p.L;
    if (adapterClip) applyAdapterClip(adapterClip[index], p.L, p.f, p.begin, p.end);   // :186
    clipReadMasking(p.L, reads.endCyclesMasked[c.readId], p.f, p.begin, p.end);  // :187
    clipReference(contigLength, p.f, p.begin, p.end);                            // :189
    p.sequenceLength = unsigned(p.end - p.begin);
    p.strandPosition = p.f.position;
    // no gapped alignment if the reference is too short (:204-208)
    p.run = p.sequenceLength && !(contigLength < long(p.sequenceLength) + p.strandPosition + 16);
    // --avoid-smith-waterman: makesSenseToGapAlign said no (:218-226, kernels_avoid.cuh)
    if (adapterClip && (adapterClip[index] >> 31)) p.run = false;
    // getFlanks (:51-82): once the "too short" test has passed the right flank is never squeezed
    p.left = p.run ? (p.strandPosition >= 8 ? 8u : unsigned(p.strandPosition)) : 0u;
    return p;
}

/// The CIGAR of one gapped candidate assembled around the Smith-Waterman operations (GappedAligner.cpp:191-240):
/// ops[1..1+nSw) hold the SW CIGAR on entry; returns the first word of the complete CIGAR and its length.
__device__ __forceinline__ uint32_t *assembleGappedCigar(const GappedPrep &p, uint32_t *ops, unsigned nSw, unsigned &nOps)
{
    uint32_t *all = ops + 1;
    nOps = nSw;
    if (p.begin) { ops[0] = cigarWord(uint32_t(p.begin), ISAAC_EXT_CIGAR_SOFT_CLIP); all = ops; ++nOps; }      // :191-195
    if (long(p.L) - p.end) all[nOps++] = cigarWord(uint32_t(p.L - p.end), ISAAC_EXT_CIGAR_SOFT_CLIP);        // :233-237
    return all;
}

/// Sequential code streams of a pair of (read strand window, reference window).  q2()/d2() must be called in
/// increasing order of their argument (sw2Forward does); every 16th call refills from memory, in lock-step for the
/// whole warp.
struct ResidentPairSrc
{
    const ReferenceView &ref;
    const uint64_t *qWords[2]; unsigned qPos[2]; uint64_t dPos[2];
    uint64_t qBuf[2], dBuf[2];
    unsigned qClamp;        // the shorter alignment of a pair keeps streaming past its own end: keep it inside its buffers
    __device__ __forceinline__ uint32_t q2(unsigned i)
    {
        if ((i & 15u) == 0)
        {
            qBuf[0] = readCodes16(qWords[0], min(qPos[0] + i, qClamp));
            qBuf[1] = readCodes16(qWords[1], min(qPos[1] + i, qClamp));
        }
        const uint32_t r = (uint32_t(qBuf[0]) & 15u) | ((uint32_t(qBuf[1]) & 15u) << 16);
        qBuf[0] >>= 4; qBuf[1] >>= 4;
        return r;
    }
    __device__ __forceinline__ uint32_t d2(unsigned k)
    {
        if ((k & 15u) == 0)
        {
            dBuf[0] = referenceCodes16(ref, min(dPos[0] + k, ref.totalBases - 64));
            dBuf[1] = referenceCodes16(ref, min(dPos[1] + k, ref.totalBases - 64));
        }
        const uint32_t r = (uint32_t(dBuf[0]) & 15u) | ((uint32_t(dBuf[1]) & 15u) << 16);
        dBuf[0] >>= 4; dBuf[1] >>= 4;
        return r;
    }
};


struct AsciiPairSrc
{
    const unsigned char *query[2]; const unsigned char *database[2]; unsigned n[2];
    __device__ __forceinline__ uint32_t q2(unsigned i) const
    {
        return (i < n[0] ? AsciiBaseSrc::qcode(query[0][i]) : 0u) | ((i < n[1] ? AsciiBaseSrc::qcode(query[1][i]) : 0u) << 16);
    }
    __device__ __forceinline__ uint32_t d2(unsigned k) const
    {
        return (k < n[0] + 15 ? asciiRefCode(database[0][k]) : 0u) | ((k < n[1] + 15 ? asciiRefCode(database[1][k]) : 0u) << 16);
    }
};

template <bool ROW_RELATIVE>
__global__ void __launch_bounds__(128)
bandedSwAsciiKernel2(uint32_t n, const unsigned char *__restrict__ queries, const uint64_t *__restrict__ queryOffsets,
                     const uint32_t *__restrict__ queryLengths, const unsigned char *__restrict__ databases,
                     const uint64_t *__restrict__ databaseOffsets, const SwScores sw, uint32_t cigarStride,
                     uint32_t *__restrict__ cigars, uint32_t *__restrict__ cigarLengths, uint32_t *__restrict__ offsets,
                     uint32_t *__restrict__ tbScratch, uint32_t *__restrict__ errorFlag)
{
    const size_t tbStride = size_t(gridDim.x) * blockDim.x;
    uint32_t *tb = tbScratch + (blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t pairs = (n + 1) / 2;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += gridDim.x * blockDim.x)
    {
        const uint32_t iA = 2 * t, iB = 2 * t + 1 < n ? 2 * t + 1 : 2 * t;
        const bool haveB = 2 * t + 1 < n;
        const unsigned LA = queryLengths[iA], LB = haveB ? queryLengths[iB] : 0u;
        AsciiPairSrc src = {{queries + queryOffsets[iA], queries + queryOffsets[iB]},
                            {databases + databaseOffsets[iA], databases + databaseOffsets[iB]}, {LA, LB}};
        int jj[2]; unsigned type[2];
        sw2Forward<ROW_RELATIVE>(src, LA, LB, sw, tb, tbStride, jj, type);
        uint32_t opsA[SW_OPS_CAP], opsB[SW_OPS_CAP];
        Sw2Walker wa, wb;
        wa.start(LA, jj[0], type[0], opsA, SW_OPS_CAP);
        wb.start(LB, jj[1], type[1], opsB, SW_OPS_CAP);
        sw2TracebackPair(tb, tbStride, wa, wb);
        for (unsigned h = 0; h < (haveB ? 2u : 1u); ++h)
        {
            const uint32_t i = h ? iB : iA;
            Sw2Walker &w = h ? wb : wa;
            unsigned nOps = 0;
            const unsigned ret = w.finish(nOps);
            if (w.overflow) atomicOr(errorFlag, 1u);
            offsets[i] = ret;
            cigarLengths[i] = nOps;
            for (unsigned k = 0; k < nOps && k < cigarStride; ++k) cigars[size_t(i) * cigarStride + k] = w.ops[k];
        }
    }
}

} // namespace isaac_b200
