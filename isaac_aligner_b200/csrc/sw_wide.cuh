// K3: banded Smith-Waterman as a warp-level wavefront with a band of W lanes (BASELINE configs[4]: 2x250 bp reads, widened band).
//
// The thread-per-alignment kernels (sw2.cuh) keep all 16 lanes of the reference's band in the registers of one thread; a band of
// 32 lanes and reads of 250+ bases do not fit that way.  Here the band lies ACROSS the lanes of a warp: lane j of a W-lane
// segment owns band lane j (W = 32: one alignment per warp, W = 16: two), rows run one after the other, and everything the
// recurrence (BandedSmithWaterman.cpp:127-347) needs from a neighbouring band lane comes through __shfl_sync:
//   F (insertion)  lane j - 1 of the previous row            two shuffles up (G and E packed as 16x2, F)
//   G (diagonal)   the same lane of the previous row          no exchange; the byte-pair merge of its direction codes (:197) takes
//                                                             the partner lane's flags: one shuffle xor 1
//   E (deletion)   serial from lane W-1 down to 0 in the reference (:246-297).  E[j] = max over k > j of
//                  (max(G[k], F[k]) - open) - (k - j - 1) * ext, a max-plus suffix scan, done in int32 on values shifted by
//                  k * ext: log2(W) shuffle steps; the direction code of a lane is then the reference's local three-way
//                  compare (priority G, then F, then E) of what lane j + 1 hands down: two shuffles down
// The database character of lane j in row i is db[i + W - 1 - j] (lane <-> database skew, :202-203): query and window are staged
// in shared memory, read once per row.  Direction codes never leave the SM: every lane stores one byte per row (TG | TE << 2 |
// TF << 4) in shared memory, W bytes per row against the reference's 3 x W; the segment's first lane walks them back (:381-435)
// and assembles the CIGAR (:437-453).  No tensor cores: integer max-plus, not a contraction.
//
// Checker: the band-width-parametrised scalar model BandedSwT<W> of the CPU restatement (its W = 16 instance is the reference's own result on 120 000 cases,
// tests/test_wide_band_oracle.py); tests/test_gpu_wide_band.py compares this kernel with it at W = 16 and W = 32.
#pragma once
#include "device_types.cuh"
#include "sw.cuh"

namespace isaac_b200
{

constexpr unsigned SW_WIDE_WARPS = 4;          // warps per CTA
constexpr unsigned SW_WIDE_OPS_CAP = 96;       // CIGAR operations a traceback may emit before the strips

/// shared memory per alignment in bytes: W direction bytes per row + the staged query and database window (word aligned)
__host__ __device__ inline unsigned swWideSharedBytes(const unsigned maxQueryLength, const unsigned W)
{
    return maxQueryLength * W + ((maxQueryLength + 3u) & ~3u) + ((maxQueryLength + W - 1u + 3u) & ~3u);
}

template <unsigned W>
__global__ void __launch_bounds__(SW_WIDE_WARPS * 32)
bandedSwWideKernel(const uint32_t n, const unsigned char *__restrict__ queries, const uint64_t *__restrict__ queryOffsets,
                   const uint32_t *__restrict__ queryLengths, const unsigned char *__restrict__ databases,
                   const uint64_t *__restrict__ databaseOffsets, const SwScores sw, const uint32_t maxQueryLength,
                   const uint32_t cigarStride, uint32_t *__restrict__ cigarOut, uint32_t *__restrict__ cigarLengthOut,
                   uint32_t *__restrict__ offsetOut, uint32_t *__restrict__ errorFlag)
{
    static_assert(W == 16 || W == 32, "the band lies across the lanes of one warp");
    constexpr unsigned PER_WARP = 32 / W;
    extern __shared__ uint32_t sharedWords[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, j = lane % W, segment = lane / W;
    const unsigned perAlignment = swWideSharedBytes(maxQueryLength, W) / 4u;
    unsigned char *directions = reinterpret_cast<unsigned char *>(sharedWords + (warp * PER_WARP + segment) * perAlignment);
    unsigned char *qStage = directions + maxQueryLength * W;
    unsigned char *dStage = qStage + ((maxQueryLength + 3u) & ~3u);
    const int init = sw.init, open = sw.open, ext = sw.ext;
    const uint32_t alignmentsPerGrid = gridDim.x * SW_WIDE_WARPS * PER_WARP;
    // every segment of a warp runs the same number of rounds (the shuffles and ballots are warp-wide): a segment without an
    // alignment in the last round idles with L = 0
    for (uint32_t base = (blockIdx.x * SW_WIDE_WARPS + warp) * PER_WARP; base < n; base += alignmentsPerGrid)
    {
        const uint32_t a = base + segment;
        const bool live = a < n;
        const unsigned L = live ? queryLengths[a] : 0u;
        const unsigned Lmax = W == 32 ? L : max(L, __shfl_xor_sync(0xFFFFFFFFu, L, 16));
        if (live)
        {
            const unsigned char *q = queries + queryOffsets[a], *d = databases + databaseOffsets[a];
            // staged once; the alphabets the recurrence is defined on (ACGTn / ACGTN) are checked on the way: bit 5 of the flag
            bool bad = false;
            for (unsigned k = j; k < L; k += W)
            {
                const unsigned char c = q[k];
                qStage[k] = c;
                bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'n');
            }
            for (unsigned k = j; k < L + W - 1u; k += W)
            {
                const unsigned char c = d[k];
                dStage[k] = c;
                bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N');
            }
            if (bad) atomicOr(errorFlag, 32u);
        }
        __syncwarp();
        int G = j ? init : 0, E = init, F = 0;                               // :108-115 (F really starts at 0)
        for (unsigned i = 0; i < Lmax; ++i)
        {
            const bool row = i < L;
            // ---- F: insertion, from lane j - 1 of the previous row, zeros shifted into lane 0 (:132-173)
            const uint32_t gePacked = (uint32_t(G) << 16) | (uint32_t(E) & 0xFFFFu);
            const uint32_t geUp = __shfl_up_sync(0xFFFFFFFFu, gePacked, 1, W);
            const int fUp = __shfl_up_sync(0xFFFFFFFFu, F, 1, W);
            const int gp = j ? int(geUp) >> 16 : 0, ep = j ? int(int16_t(geUp & 0xFFFFu)) : 0, fp = j ? fUp : 0;
            unsigned tf = gp < ep ? 1u : 0u;
            const int fa = max(gp, ep) - open, fb = fp - ext;
            if (fa < fb) tf = 2u;
            int nF = max(fa, fb);
            if (!j) { tf = 0u; nF = init; }                                  // :167,173
            // ---- G: diagonal, the same lane of the previous row; raw byte compare (:176-205, 230-244)
            const unsigned tgE = G < E ? 1u : 0u;
            int g = max(G, E);
            const unsigned tgF = g < F ? 2u : 0u;
            g = max(g, F);
            const bool differ = row && qStage[i] != dStage[i + (W - 1u) - j];
            const int nG = g + (differ ? sw.mismatch : sw.match);
            // the direction bytes of G are merged with a signed 16-bit max over BYTE PAIRS of lanes (2p, 2p + 1) (:197)
            const unsigned mineFlags = tgE | (tgF << 2);
            const unsigned partnerFlags = __shfl_xor_sync(0xFFFFFFFFu, mineFlags, 1);
            const unsigned lowFlags = (j & 1u) ? partnerFlags : mineFlags, highFlags = (j & 1u) ? mineFlags : partnerFlags;
            const unsigned x = (lowFlags >> 2) | ((highFlags >> 2) << 8), y = (lowFlags & 3u) | ((highFlags & 3u) << 8);
            const unsigned merged = max(x, y);                               // both are small and non-negative
            const unsigned tg = (j & 1u) ? merged >> 8 : merged & 0xFFu;
            // ---- E: deletion, the reference's serial pass over the lanes as a max-plus suffix scan (:246-297)
            const int opening = max(nG, nF) - open;                          // what lane j offers the lanes below it
            int s = opening - int(j) * ext;
#pragma unroll
            for (unsigned dlt = 1; dlt < W; dlt <<= 1)
            {
                const int t = __shfl_down_sync(0xFFFFFFFFu, s, dlt, W);
                if (j + dlt < W) s = max(s, t);
            }
            const int above = __shfl_down_sync(0xFFFFFFFFu, s, 1, W);        // suffix maximum over the lanes k > j
            const int nE = j == W - 1u ? init : max(above + int(j + 1u) * ext, init - int(W - 1u - j) * ext);
            // direction of E: the three values lane j + 1 hands down, priority G, then F, then E (:250-262)
            const uint32_t gfPacked = (uint32_t(nG) << 16) | (uint32_t(nF) & 0xFFFFu);
            const uint32_t gfDown = __shfl_down_sync(0xFFFFFFFFu, gfPacked, 1, W);
            const int eDown = __shfl_down_sync(0xFFFFFFFFu, nE, 1, W);
            const int gIn = j == W - 1u ? init : (int(gfDown) >> 16) - open;
            const int fIn = j == W - 1u ? init : int(int16_t(gfDown & 0xFFFFu)) - open;
            const int eIn = j == W - 1u ? init : eDown - ext;
            const unsigned te = (eIn > gIn && eIn > fIn) ? 1u : (fIn > gIn ? 2u : 0u);
            // ---- the direction codes of the row
            if (row) directions[i * W + j] = (unsigned char)(tg | (te << 2) | (tf << 4));
            if (row) { G = nG; E = nE; F = nF; }
        }
        __syncwarp();
        // ---- end cell: lanes W-1 .. 0, matrices G, E, F in that order, strict '>' (:349-379): the largest value, ties to
        // the cell met first
        int bestValue = G; unsigned bestType = 0;
        if (E > bestValue) { bestValue = E; bestType = 1; }
        if (F > bestValue) { bestValue = F; bestType = 2; }
        // key: value (biased to be positive) above the order of the scan turned upside down
        uint32_t key = (uint32_t(bestValue + 40000) << 8) | (255u - ((W - 1u - j) * 3u + bestType));
#pragma unroll
        for (unsigned dlt = W / 2; dlt; dlt >>= 1) key = max(key, __shfl_xor_sync(0xFFFFFFFFu, key, dlt, W));
        // ---- traceback and CIGAR by the segment's first lane, operations emitted tail first (:381-453)
        if (live && j == 0)
        {
            const unsigned order = 255u - (key & 0xFFu);
            int jj = int(W - 1u - order / 3u);
            unsigned type = order % 3u;
            int ii = int(L) - 1;
            uint32_t ops[SW_WIDE_OPS_CAP];
            unsigned nOps = 0, opLength = 0;
            bool overflow = false;
            auto emit = [&](const unsigned length, const unsigned t) {
                const uint32_t op = t == 0 ? ISAAC_EXT_CIGAR_ALIGN : t == 1 ? ISAAC_EXT_CIGAR_DELETE : ISAAC_EXT_CIGAR_INSERT;
                if (nOps < SW_WIDE_OPS_CAP) ops[nOps++] = (length << 4) | op; else overflow = true;
            };
            if (jj > 0) emit(unsigned(jj), 1u);
            while (ii >= 0 && jj >= 0 && jj <= int(W) - 1)
            {
                ++opLength;
                const unsigned next = (unsigned(directions[unsigned(ii) * W + unsigned(jj)]) >> (type * 2u)) & 3u;
                if (next != type) { emit(opLength, type); opLength = 0; }
                if (type == 0) --ii; else if (type == 1) ++jj; else { --ii; --jj; }
                type = next;
            }
            if (type != 1 && opLength) { emit(opLength, type); opLength = 0; }
            if (jj < int(W) - 1) { emit(opLength + (W - 1u) - unsigned(jj), 1u); opLength = 0; }
            // strip the deletion at the alignment start (the last one emitted), reverse, strip the one at the end
            unsigned offset = 0;
            if (nOps && (ops[nOps - 1] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { offset = ops[nOps - 1] >> 4; --nOps; }
            unsigned first = 0;
            if (nOps && (ops[0] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) first = 1;       // the end of the alignment was emitted first
            const unsigned count = nOps - first;
            if (overflow || count > cigarStride) atomicOr(errorFlag, 1u);
            for (unsigned k = 0; k < count && k < cigarStride; ++k) cigarOut[size_t(a) * cigarStride + k] = ops[nOps - 1u - k];
            cigarLengthOut[a] = count;
            offsetOut[a] = offset;
        }
        __syncwarp();
    }
}

} // namespace isaac_b200
