// Third-generation gapped path: the packed 16x2 Smith-Waterman of sw2.cuh split into two kernels.
//
//   swForwardKernel      forward DP of candidate pairs (2t, 2t+1): writes the flag masks of every row and the
//                        end cell of both halves.  128 registers, nothing but the DP loop in its instruction footprint.
//   swTraceScoreKernel   one thread per candidate: traceback from the planes, CIGAR assembly, updateFragmentCigar.  These
//                        phases are chains of dependent loads and FP64 adds; with a third of the registers they run at
//                        three times the occupancy of the fused kernel and no longer evict the DP loop from the
//                        instruction cache (the fused kernel's top stall was "no instruction").
//
// The batch is cut into chunks of pairs; the plane buffers of two chunks are alive at a time so that the forward kernel
// of chunk k+1 overlaps the trace kernel of chunk k on two streams (isaac_ext.cu).
#pragma once
#include "kernels2.cuh"

// rows of direction planes a walk keeps in flight, and resident blocks per SM the trace kernel is compiled for
// (measured on B200, 10M x 150 bp: 4 rows beat 8, 12 and 16 rows, which cost occupancy; 12 blocks = 40 registers with
// a few spilled words beat 10 and 8 blocks: the kernel waits on loads, not on the issue slots)
#ifndef ISAAC_TRACE_ROWS_AHEAD
#define ISAAC_TRACE_ROWS_AHEAD 4
#endif
#ifndef ISAAC_TRACE_MIN_BLOCKS
#define ISAAC_TRACE_MIN_BLOCKS 12
#endif

namespace isaac_b200
{

/// plane layout of a chunk: planes[(row * SW2_FLAG_WORDS + k) * pairStride + pair]
template <bool ROW_RELATIVE>
__global__ void __launch_bounds__(128, 4)
swForwardKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams sp, uint32_t n,
                const isaac_ext_candidate_t *__restrict__ candidates, uint32_t *__restrict__ planes, uint32_t pairStride,
                uint32_t *__restrict__ endCells, const uint32_t *__restrict__ adapterClip = nullptr)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pairs = (n + 1) / 2;
    if (t >= pairs) return;
    const uint32_t iA = 2 * t, iB = 2 * t + 1;
    const bool haveB = iB < n;
    const GappedPrep pa = prepareGapped(ref, reads, candidates[iA], adapterClip, iA);
    GappedPrep pb = prepareGapped(ref, reads, candidates[haveB ? iB : iA], adapterClip, haveB ? iB : iA);
    if (!haveB) pb.run = false;
    const unsigned LA = pa.run ? pa.sequenceLength : 0u, LB = pb.run ? pb.sequenceLength : 0u;
    int jj[2] = {0, 0}; unsigned type[2] = {0, 0};
    if (LA | LB)
    {
        const SwScores sw = {sp.swMatch, sp.swMismatch, sp.swOpen, sp.swExtend, -32768 + sp.swOpen};
        ResidentPairSrc src = {ref,
                               {reads.strandCodes(pa.c.readId, pa.f.reverse), reads.strandCodes(pb.c.readId, pb.f.reverse)},
                               {LA ? unsigned(pa.begin) : 0u, LB ? unsigned(pb.begin) : 0u},
                               {LA ? ref.contigOffset[pa.contigId] + uint64_t(pa.strandPosition - long(pa.left)) : 0ull,
                                LB ? ref.contigOffset[pb.contigId] + uint64_t(pb.strandPosition - long(pb.left)) : 0ull},
                               {0, 0}, {0, 0}, reads.codesClamp()};
        sw2Forward<ROW_RELATIVE>(src, LA, LB, sw, planes + t, pairStride, jj, type);
    }
    endCells[t] = uint32_t(jj[0] & 0xFF) | (type[0] << 8) | (uint32_t(jj[1] & 0xFF) << 16) | (type[1] << 24);
}

/// One thread per candidate of the chunk; 'base' = index of the chunk's first candidate in the batch (the output
/// pointers are already advanced to it).
__global__ void __launch_bounds__(128, ISAAC_TRACE_MIN_BLOCKS)
swTraceScoreKernel(const ReferenceView ref, const ReadSetView reads, const ScoreParams spGlobal, uint32_t n, uint32_t base,
                   const isaac_ext_candidate_t *__restrict__ candidates, const uint32_t *__restrict__ planes, uint32_t pairStride,
                   const uint32_t *__restrict__ endCells, uint32_t cigarStride, isaac_ext_fragment_t *__restrict__ fragments,
                   uint32_t *__restrict__ cigars, uint64_t *__restrict__ masks, uint32_t *__restrict__ errorFlag,
                   const uint32_t *__restrict__ adapterClip = nullptr)
{
    __shared__ double tables[201];
    const ScoreParams sp = stageScoreTables(spGlobal, tables);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pair = i >> 1;
    const GappedPrep p = prepareGapped(ref, reads, candidates[i], adapterClip, i);
    isaac_ext_fragment_t o;
    initFragment(o, p.c, reads.readCount);
    o.cigarOffset = (base + i) * cigarStride;      // records and CIGAR rows are addressed from the batch start
    o.lowClipped = uint16_t(p.f.lowClipped); o.highClipped = uint16_t(p.f.highClipped); o.position = p.f.position;
    uint64_t *mask = masks ? masks + size_t(i) * ISAAC_EXT_MASK_WORDS : nullptr;
    if (mask) for (unsigned k = 0; k < ISAAC_EXT_MASK_WORDS; ++k) mask[k] = 0;
    if (p.run)
    {
        const uint32_t cell = endCells[pair] >> ((i & 1u) * 16u);
        uint32_t ops[SW_OPS_CAP + 2];
        Sw2Walker w;
        w.start(p.sequenceLength, int(cell & 0xFFu), (cell >> 8) & 0xFFu, ops + 1, SW_OPS_CAP);
        // every row from L-1 down to 0 is visited once: the one word a walk on the diagonal needs of AHEAD rows is kept in flight;
        // the five masks are fetched only for a row in which some walk of the warp leaves the diagonal
        const unsigned half = i & 1u;
        const uint32_t *tb = planes + pair;
        const size_t rowStride = size_t(SW2_FLAG_WORDS) * pairStride;
        auto load = [&](int r, uint32_t &q) { q = tb[size_t(max(r, 0)) * rowStride]; };
        constexpr int AHEAD = ISAAC_TRACE_ROWS_AHEAD;
        uint32_t q[AHEAD];
        const int top = w.ii;
#pragma unroll
        for (int k = 0; k < AHEAD; ++k) load(top - k, q[k]);
        for (int r = top; r >= 0 && w.active; r -= AHEAD)
        {
#pragma unroll
            for (int k = 0; k < AHEAD; ++k)
            {
                const int row = r - k;
                if (row >= 0 && w.active && w.ii == row)
                {
                    // when all lanes stay on the diagonal the row costs a handful of instructions; otherwise all lanes run the
                    // same general step once (no per-lane slow path for the others to wait for)
                    const bool diagonal = w.staysOnDiagonal(q[k], half);
                    if (__all_sync(__activemask(), diagonal)) w.diagonalStep();
                    else
                    {
                        const uint32_t *f = tb + size_t(row) * rowStride + pairStride;
                        const size_t ps = pairStride;
                        uint32_t p[SW2_PLANE_WORDS];
                        sw2Planes(f[0], f[ps], f[2 * ps], f[3 * ps], f[4 * ps], p);
                        w.stepsInRow(row, p[half * 3u], p[half * 3u + 1u], p[half * 3u + 2u]);
                    }
                }
                load(r - k - AHEAD, q[k]);
            }
        }
        unsigned nSw = 0, nOps = 0;
        const unsigned ret = w.finish(nSw);
        uint32_t *all = assembleGappedCigar(p, ops, nSw, nOps);
        const long position = p.strandPosition + long(ret) - long(p.left);                       // GappedAligner.cpp:231,240
        if (w.overflow || nOps > cigarStride) atomicOr(errorFlag, 1u);
        else
        {
            scoreCigar(ref, reads, sp, p.c.readId, p.L, p.f.reverse, ref.contigOffset[p.contigId], position, all, nOps, o, mask);
            for (unsigned k = 0; k < nOps; ++k) cigars[size_t(i) * cigarStride + k] = all[k];
            o.cigarLength = uint16_t(nOps);
        }
    }
    fragments[i] = o;
}

} // namespace isaac_b200
