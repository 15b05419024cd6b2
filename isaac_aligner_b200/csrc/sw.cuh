// Banded affine-gap Smith-Waterman, one alignment per thread (scalar lanes).
//
// Restates the semantics of BandedSmithWaterman::align (reference lib/alignment/BandedSmithWaterman.cpp:84-462,
// SURVEY.md Appendix A) for a GPU thread: the 16 band lanes live in registers, one query row per iteration; lanes are
// walked from 15 down to 0 so that F (needs lane j-1 of the previous row), G (lane j of the previous row) and E (lane
// j+1 of the current row) are all updated in place.  The three direction codes of a cell are packed 2 bits per lane
// into one 32-bit word per matrix and row and stored to a thread-interleaved scratch (tb[(row*3+matrix)*stride + slot])
// so that a warp's stores coalesce; the traceback reads single words back.
//
// Arithmetic: the reference works in wrapping int16.  With gapExtend <= gapOpen and the constructor's overflow guard
// (BandedSmithWaterman.cpp:47-53) no intermediate leaves the int16 range, so plain int arithmetic is bit-identical;
// the host rejects any other score set (isaac_ext_create).
#pragma once
#include "device_types.cuh"

namespace isaac_b200
{

struct SwScores
{
    int match, mismatch, open, ext;   // open/ext positive (BandedSmithWaterman.hh:44-46)
    int init;                         // -32768 + open (BandedSmithWaterman.cpp:44)
};

__host__ __device__ __forceinline__ uint32_t cigarWord(uint32_t length, uint32_t op) { return (length << 4) | op; }

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

} // namespace isaac_b200
