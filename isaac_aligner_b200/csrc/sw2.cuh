// Banded affine-gap Smith-Waterman, TWO alignments per thread packed as 16x2 SIMD-in-register.
//
// Same semantics as sw.cuh (BandedSmithWaterman::align, reference lib/alignment/BandedSmithWaterman.cpp:84-462), laid
// out for the sm_100a integer pipes:
//   * every 32-bit register holds the same band lane of two independent alignments (A in the low half, B in the high
//     half), so the native packed instructions VIMNMX.U16x2 / VIADDMNMX.U16x2 do two cells at a time and no lane ever
//     has to be shifted inside a register: "lane j-1" is simply another register;
//   * cell values are kept biased (value + 32768) as unsigned 16-bit.  Under the score conditions the host enforces
//     no value leaves the int16 range, hence adding a constant to both halves is ONE 32-bit add of c * 0x10001 (no
//     borrow can cross the halves), and the match/mismatch score of both alignments is one IMAD;
//   * the "which operand won" flags the traceback needs are recovered without predicates: max(a,b) != a  <=>  a < b,
//     so flag = min((max ^ a), 1) per half (XOR + VIMNMX.U16x2), shifted into a per-row accumulator by an IMAD on the
//     FMA pipe.  5 flag words per row and thread pair are stored (10 bytes per alignment row):
//        fE  bit j: G[j] <  E[j]                 (previous row; TG and the TF of lane j+1)
//        fF  bit j: max(G[j],E[j]) < F[j]        (previous row; TG)
//        fAB bit j: max(G,E)[j-1]-open < F[j-1]-ext   (TF of lane j)
//        fGF bit j: newG[j] < newF[j]            (TE of lane j-1)
//        fHE bit j: max(newG,newF)[j+1]-open < newE[j+1]-ext   (TE of lane j)
//     The direction codes (including the reference's _mm_max_epi16-on-bytes coupling of lanes 2p/2p+1, :197) are
//     decoded from these bits only along the traceback path.
#pragma once
#include "device_types.cuh"
#include "sw.cuh"

namespace isaac_b200
{

struct Sw2Consts
{
    uint32_t negOpen32;    // (-open) * 0x10001 as a 32-bit addend
    uint32_t negExt16x2;   // (65536 - ext) in both halves, for the wrapping per-half add of VIADDMNMX
    uint32_t match32;      // match * 0x10001
    int delta;             // mismatch - match
    uint32_t init2;        // biased init (= open) in both halves
};

__device__ __forceinline__ Sw2Consts makeSw2Consts(const SwScores s)
{
    Sw2Consts c;
    c.negOpen32 = uint32_t(-s.open) * 0x10001u;
    c.negExt16x2 = (uint32_t(65536 - s.ext) & 0xFFFFu) * 0x10001u;
    c.match32 = uint32_t(s.match) * 0x10001u;
    c.delta = s.mismatch - s.match;
    c.init2 = uint32_t(s.init + 32768) * 0x10001u;
    return c;
}

constexpr unsigned SW2_FLAG_WORDS = 5;

/// End-cell scan of one half (:349-379): lanes 15..0, matrices G,E,F in that order, strict '>'.
__device__ __forceinline__ void sw2ScanEnd(const uint32_t (&G)[16], const uint32_t (&E)[16], const uint32_t (&F)[16],
                                           const unsigned half, int &jj, unsigned &type)
{
    const unsigned sh = half * 16;
    int best = int((G[15] >> sh) & 0xFFFFu) - 1;
    type = 0;
#pragma unroll
    for (int j = 15; j >= 0; --j)
    {
        const int g = int((G[j] >> sh) & 0xFFFFu), e = int((E[j] >> sh) & 0xFFFFu), f = int((F[j] >> sh) & 0xFFFFu);
        if (g > best) { best = g; jj = j; type = 0; }
        if (e > best) { best = e; jj = j; type = 1; }
        if (f > best) { best = f; jj = j; type = 2; }
    }
}

/// Traceback of one half from the stored flag words (:381-453).  ops receives the CIGAR head first.
__device__ __forceinline__ unsigned sw2Traceback(const uint32_t *__restrict__ tb, const size_t tbStride, const unsigned half,
                                                 const unsigned L, int jj, unsigned type, uint32_t *ops, const unsigned cap,
                                                 unsigned &nOps, bool &overflow)
{
    const unsigned sh = half * 16;
    int ii = int(L) - 1;
    unsigned w = cap;
    overflow = false;
    auto push = [&](unsigned length, unsigned type3) {
        const uint32_t op = type3 == 0 ? ISAAC_EXT_CIGAR_ALIGN : (type3 == 1 ? ISAAC_EXT_CIGAR_DELETE : ISAAC_EXT_CIGAR_INSERT);
        if (w == 0) { overflow = true; return; }
        ops[--w] = cigarWord(length, op);
    };
    unsigned opLength = 0;
    if (jj > 0) push(unsigned(jj), 1);
    // Every row from L-1 down to 0 is visited exactly once, so the flag words of the rows ahead can be pulled into
    // L1 while the current row is decoded: the walk is a chain of dependent loads otherwise.
    constexpr int PREFETCH_ROWS = 8;
    auto prefetchRow = [&](int r) {
        if (r >= 0)
        {
            const uint32_t *p = tb + size_t(r) * SW2_FLAG_WORDS * tbStride;
#pragma unroll
            for (unsigned k = 0; k < SW2_FLAG_WORDS; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + k * tbStride));
        }
    };
    for (int r = 0; r < PREFETCH_ROWS; ++r) prefetchRow(ii - r);
    int cachedRow = -1;
    uint32_t row[SW2_FLAG_WORDS] = {0, 0, 0, 0, 0};
    while (ii >= 0 && jj >= 0 && jj <= 15)
    {
        ++opLength;
        if (cachedRow != ii)
        {
            const uint32_t *p = tb + size_t(ii) * SW2_FLAG_WORDS * tbStride;
#pragma unroll
            for (unsigned k = 0; k < SW2_FLAG_WORDS; ++k) row[k] = p[k * tbStride];
            cachedRow = ii;
            prefetchRow(ii - PREFETCH_ROWS);
        }
        unsigned next;
        if (type == 0)
        {
            const unsigned fE = (row[0] >> sh) & 0xFFFFu, fF = (row[1] >> sh) & 0xFFFFu;
            const unsigned lo = unsigned(jj) & ~1u, hi = lo + 1;
            const unsigned tgElo = (fE >> lo) & 1u, tgEhi = (fE >> hi) & 1u;
            const unsigned tgFlo = ((fF >> lo) & 1u) * 2u, tgFhi = ((fF >> hi) & 1u) * 2u;
            // _mm_max_epi16 on byte pairs (:197): the odd lane decides which whole pair wins
            if (jj & 1) next = tgFhi ? 2u : tgEhi;
            else next = tgFhi ? tgFlo : (tgEhi ? tgElo : max(tgFlo, tgElo));
        }
        else if (type == 1)
        {
            const unsigned fHE = (row[4] >> sh) & 0xFFFFu, fGF = (row[3] >> sh) & 0xFFFFu;
            next = ((fHE >> jj) & 1u) ? 1u : (((fGF >> (jj + 1)) & 1u) ? 2u : 0u);
        }
        else
        {
            const unsigned fAB = (row[2] >> sh) & 0xFFFFu, fE = (row[0] >> sh) & 0xFFFFu;
            next = jj == 0 ? 0u : (((fAB >> jj) & 1u) ? 2u : ((fE >> (jj - 1)) & 1u));
        }
        if (next != type) { push(opLength, type); opLength = 0; }
        if (type == 0) { --ii; } else if (type == 1) { ++jj; } else { --ii; --jj; }
        type = next;
    }
    if (type != 1 && opLength) { push(opLength, type); opLength = 0; }
    if (jj < 15) { push(opLength + 15 - jj, 1); opLength = 0; }
    unsigned ret = 0;
    unsigned e = cap;
    if (w < e && (ops[w] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { ret = ops[w] >> 4; ++w; }
    if (w < e && (ops[e - 1] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { --e; }
    nOps = e - w;
    for (unsigned k = 0; k < nOps; ++k) ops[k] = ops[w + k];
    return ret;
}

/// One traceback in progress (one half of a pair).
struct Sw2Walker
{
    int ii, jj;                 // current row / lane
    unsigned type;              // 0 G (ALIGN), 1 E (DELETE), 2 F (INSERT)
    unsigned opLength;
    unsigned w;                 // next free slot is ops[w-1]; operations are written tail first from the back
    bool active, overflow;
    uint32_t *ops; unsigned cap;

    __device__ __forceinline__ void push(unsigned length, unsigned type3)
    {
        const uint32_t op = type3 == 0 ? ISAAC_EXT_CIGAR_ALIGN : (type3 == 1 ? ISAAC_EXT_CIGAR_DELETE : ISAAC_EXT_CIGAR_INSERT);
        if (w == 0) { overflow = true; return; }
        ops[--w] = cigarWord(length, op);
    }
    __device__ __forceinline__ void start(unsigned L, int endLane, unsigned endType, uint32_t *buffer, unsigned capacity)
    {
        ii = int(L) - 1; jj = endLane; type = endType; opLength = 0; ops = buffer; cap = capacity; w = capacity;
        overflow = false; active = L != 0;
        if (active && jj > 0) push(unsigned(jj), 1);                           // :388-391
    }
    /// all steps of this walk that happen in row 'row' (:392-423); f = the 16 flag bits of this half of the 5 row words
    __device__ __forceinline__ void stepRow(int row, unsigned fE, unsigned fF, unsigned fAB, unsigned fGF, unsigned fHE)
    {
        if (!active || ii != row) return;
        // fast path, the overwhelmingly common step: on the diagonal, and neither lane of the byte pair (2p, 2p+1)
        // has a direction flag set, so the direction of G is G again whatever the _mm_max_epi16 coupling does
        if (type == 0 && (((fE | fF) >> (unsigned(jj) & ~1u)) & 3u) == 0)
        {
            ++opLength; --ii;
            active = ii >= 0;
            return;
        }
        while (active && ii == row)
        {
            ++opLength;
            unsigned next;
            if (type == 0)
            {
                // _mm_max_epi16 on byte pairs (:197): the odd lane decides which whole pair wins
                const unsigned lo = unsigned(jj) & ~1u;
                const unsigned loE = (fE >> lo) & 1u, hiE = (fE >> lo) & 2u, loF = (fF >> lo) & 1u, hiF = (fF >> lo) & 2u;
                if (jj & 1) next = hiF ? 2u : (hiE >> 1);
                else next = hiF ? loF * 2u : (hiE ? loE : max(loF * 2u, loE));
            }
            else if (type == 1) next = ((fHE >> jj) & 1u) ? 1u : (((fGF >> (jj + 1)) & 1u) ? 2u : 0u);
            else next = jj == 0 ? 0u : (((fAB >> jj) & 1u) ? 2u : ((fE >> (jj - 1)) & 1u));
            if (next != type) { push(opLength, type); opLength = 0; }
            if (type == 0) { --ii; } else if (type == 1) { ++jj; } else { --ii; --jj; }
            type = next;
            active = ii >= 0 && jj >= 0 && jj <= 15;
        }
    }
    /// :425-453: flush, strip a deletion at either end; \return the stripped leading deletion, ops[0..nOps) head first
    __device__ __forceinline__ unsigned finish(unsigned &nOps)
    {
        if (type != 1 && opLength) { push(opLength, type); opLength = 0; }
        if (jj < 15) { push(opLength + 15 - jj, 1); opLength = 0; }
        unsigned ret = 0, e = cap;
        if (w < e && (ops[w] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { ret = ops[w] >> 4; ++w; }
        if (w < e && (ops[e - 1] & 0xFu) == ISAAC_EXT_CIGAR_DELETE) { --e; }
        nOps = e - w;
        for (unsigned k = 0; k < nOps; ++k) ops[k] = ops[w + k];
        return ret;
    }
};

/// Traceback of both halves of a pair in one pass over the rows (every row from L-1 down to 0 is visited exactly once
/// by each walk, so the row order is known in advance): the five flag words of four rows ahead are kept in flight in
/// registers, turning the walk's chain of dependent loads into a software pipeline.
__device__ __forceinline__ void sw2TracebackPair(const uint32_t *__restrict__ tb, const size_t tbStride,
                                                 Sw2Walker &a, Sw2Walker &b)
{
    const int top = max(a.active ? a.ii : -1, b.active ? b.ii : -1);
    if (top < 0) return;
    auto load = [&](int r, uint32_t (&w)[SW2_FLAG_WORDS]) {
        const uint32_t *p = tb + size_t(max(r, 0)) * SW2_FLAG_WORDS * tbStride;
#pragma unroll
        for (unsigned k = 0; k < SW2_FLAG_WORDS; ++k) w[k] = p[k * tbStride];
    };
    auto process = [&](int r, const uint32_t (&w)[SW2_FLAG_WORDS]) {
        if (r < 0) return;
        a.stepRow(r, w[0] & 0xFFFFu, w[1] & 0xFFFFu, w[2] & 0xFFFFu, w[3] & 0xFFFFu, w[4] & 0xFFFFu);
        b.stepRow(r, w[0] >> 16, w[1] >> 16, w[2] >> 16, w[3] >> 16, w[4] >> 16);
    };
    uint32_t w0[SW2_FLAG_WORDS], w1[SW2_FLAG_WORDS], w2[SW2_FLAG_WORDS], w3[SW2_FLAG_WORDS];
    load(top, w0); load(top - 1, w1); load(top - 2, w2); load(top - 3, w3);
    for (int r = top; r >= 0 && (a.active || b.active); r -= 4)
    {
        process(r, w0); load(r - 4, w0);
        process(r - 1, w1); load(r - 5, w1);
        process(r - 2, w2); load(r - 6, w2);
        process(r - 3, w3); load(r - 7, w3);
    }
}

/// Forward pass over max(LA, LB) rows for the pair.  src.q(half, i) / src.d(half, k) return base codes (they must
/// tolerate indices past the own length of a half: any code will do there).  On return jj/type hold the end cell of
/// each half.
template <class PairSrc>
__device__ __forceinline__ void sw2Forward(PairSrc &src, const unsigned LA, const unsigned LB, const SwScores s,
                                           uint32_t *__restrict__ tb, const size_t tbStride,
                                           int (&jj)[2], unsigned (&type)[2])
{
    const Sw2Consts c = makeSw2Consts(s);
    uint32_t G[16], E[16], F[16], D[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { G[j] = c.init2; E[j] = c.init2; F[j] = 0x80008000u; D[j] = 0; }   // :108-114
    G[0] = 0x80008000u;                                                                               // :115
    // D[j] = database codes seen by lane j = db[i + 15 - j]; preload db[0..14] (:117-122)
#pragma unroll
    for (int k = 0; k < 15; ++k) D[14 - k] = src.d2(k);
    jj[0] = int(LA) - 1; jj[1] = int(LB) - 1; type[0] = 0; type[1] = 0;
    const unsigned Lmax = max(LA, LB);
#pragma unroll 1
    for (unsigned i = 0; i < Lmax; ++i)
    {
#pragma unroll
        for (int j = 15; j > 0; --j) D[j] = D[j - 1];
        D[0] = src.d2(i + 15);
        const uint32_t Q = src.q2(i);
        uint32_t fE = 0, fF = 0, fAB = 0, fGF = 0, fHE = 0;
        uint32_t mCur = __vmaxu2(G[15], E[15]);
        uint32_t xE = mCur ^ G[15];
        uint32_t hOnext = 0, nEnext = 0;
#pragma unroll
        for (int j = 15; j >= 0; --j)
        {
            // ---- F of lane j from lane j-1 of the previous row (:132-173)
            uint32_t nF, xAB, mPrev = 0, xEprev = 0;
            if (j > 0)
            {
                mPrev = __vmaxu2(G[j - 1], E[j - 1]);
                xEprev = mPrev ^ G[j - 1];
                const uint32_t a = mPrev + c.negOpen32;
                nF = __viaddmax_u16x2(F[j - 1], c.negExt16x2, a);
                xAB = nF ^ a;
            }
            else { nF = c.init2; xAB = 0; }                                       // :167, :173
            // ---- G of lane j from the same lane (:176-190, :230-244)
            const uint32_t g = __vmaxu2(mCur, F[j]);
            const uint32_t xF = g ^ mCur;
            const uint32_t t = __vminu2(D[j] ^ Q, 0x00010001u);
            const uint32_t nG = g + c.match32 + t * uint32_t(c.delta);
            // ---- E of lane j from lane j+1 of THIS row (:261-297)
            uint32_t nE, xHE;
            if (j < 15)
            {
                nE = __viaddmax_u16x2(nEnext, c.negExt16x2, hOnext);
                xHE = nE ^ hOnext;
            }
            else { nE = c.init2; xHE = 0; }
            const uint32_t h = __vmaxu2(nG, nF);
            const uint32_t xGF = h ^ nG;
            hOnext = h + c.negOpen32;
            nEnext = nE;
            // ---- flags, bit j of each half
            fE = fE * 2u + __vminu2(xE, 0x00010001u);
            fF = fF * 2u + __vminu2(xF, 0x00010001u);
            fAB = fAB * 2u + __vminu2(xAB, 0x00010001u);
            fGF = fGF * 2u + __vminu2(xGF, 0x00010001u);
            fHE = fHE * 2u + __vminu2(xHE, 0x00010001u);
            G[j] = nG; E[j] = nE; F[j] = nF;
            mCur = mPrev; xE = xEprev;
        }
        uint32_t *row = tb + size_t(i) * SW2_FLAG_WORDS * tbStride;
        row[0] = fE; row[tbStride] = fF; row[2 * tbStride] = fAB; row[3 * tbStride] = fGF; row[4 * tbStride] = fHE;
        if (LA != LB)
        {
            if (i + 1 == LA) sw2ScanEnd(G, E, F, 0, jj[0], type[0]);
            if (i + 1 == LB) sw2ScanEnd(G, E, F, 1, jj[1], type[1]);
        }
    }
    if (LA == LB && LA)
    {
        sw2ScanEnd(G, E, F, 0, jj[0], type[0]);
        sw2ScanEnd(G, E, F, 1, jj[1], type[1]);
    }
}

} // namespace isaac_b200
