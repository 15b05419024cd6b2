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
//   * the "which operand won" flags the traceback needs: max(a,b) != a  <=>  a < b.  For the three plain maxima the
//     VIMNMX.U16x2 itself delivers "max == a" of both halves as predicates and a predicated add sets the lane's bit
//     (sw2MaxFlag); for the two fused add-max (VIADDMNMX.U16x2 has no predicate output) flag = min(max - a, 1) per
//     half (the halves of max - a cannot borrow from each other because max >= a in both: ONE 32-bit subtract + one
//     VIMNMX.U16x2), shifted into a per-row accumulator.  Five 16-lane flag masks per half and row:
//        fE  bit j: G[j] <  E[j]                 (previous row; TG and the TF of lane j+1)
//        fF  bit j: max(G[j],E[j]) < F[j]        (previous row; TG)
//        fAB bit j: max(G,E)[j-1]-open < F[j-1]-ext   (TF of lane j)
//        fGF bit j: newG[j] < newF[j]            (TE of lane j-1)
//        fHE bit j: max(newG,newF)[j+1]-open < newE[j+1]-ext   (TE of lane j)
//   * what a row leaves for the traceback is SIX words per thread pair (12 bytes per alignment row, what the reference's
//     3 x 16 direction bytes compress to), half A in the low and half B in the high 16 bits of each: first the lanes whose
//     G direction is not "stay on the diagonal" (sw2OffDiagonal: a handful of operations on fE and fF, including the
//     reference's _mm_max_epi16-on-byte-pairs coupling of lanes 2p/2p+1 in TG, :197) -- the one word nearly every step of
//     nearly every walk reads --, then the five masks as they are.  The walk turns them into the six bit planes
//     (tg1,tg2) (te1,te2) (tf1,tf2), plane 1 = "direction is E", plane 2 = "direction is F" of the three direction matrices
//     TG/TE/TF (sw2Planes, a dozen bitwise operations on whole masks), only in the rows where it leaves the diagonal: the
//     forward loop, which is what the pass waits for, does not pay for directions nobody looks at.
#pragma once
#include "device_types.cuh"
#include "sw.cuh"

namespace isaac_b200
{

/// may the forward pass keep row-relative values?  The largest stored value is score + 32768 + 2 * readLength * match
__host__ __device__ inline bool sw2RowRelativeFits(const int match, const unsigned maxReadLength)
{
    return 2ll * (long long)(maxReadLength) * match < 32768ll;
}

struct Sw2Consts
{
    uint32_t negOpen32;    // (-open) * 0x10001 as a 32-bit addend
    uint32_t negExt16x2;   // (65536 - ext) in both halves, for the wrapping per-half add of VIADDMNMX
    uint32_t negOpenRow32; // the same two for the step from one row to the next (F): (-(open + match)), 65536 - (ext + match)
    uint32_t negExtRow16x2;
    uint32_t match32;      // match * 0x10001
    int delta;             // mismatch - match
    uint32_t init2;        // biased init (= open) in both halves
};

// Row-relative values.  Every cell of row i is kept as value - (i + 1) * match (+ a constant that keeps it non-negative, see
// sw2Forward): the diagonal step nG = max(...) + match + t * (mismatch - match) then is one multiply-add instead of an add and a
// multiply-add, the step from row to row (F) takes its constants less 'match', the step inside a row (E) is unchanged, and the two
// cells a row seeds with 'init' take a register that drops by 'match' per row.  Every comparison the recurrence makes is between
// cells of one row, so all the flags -- the only thing that leaves the kernel besides the end cell, which is also picked inside
// one row -- are the ones the plain values give.
__device__ __forceinline__ Sw2Consts makeSw2Consts(const SwScores s)
{
    Sw2Consts c;
    c.negOpen32 = uint32_t(-s.open) * 0x10001u;
    c.negExt16x2 = (uint32_t(65536 - s.ext) & 0xFFFFu) * 0x10001u;
    c.negOpenRow32 = uint32_t(-(s.open + s.match)) * 0x10001u;
    c.negExtRow16x2 = (uint32_t(65536 - (s.ext + s.match)) & 0xFFFFu) * 0x10001u;
    c.match32 = uint32_t(s.match) * 0x10001u;
    c.delta = s.mismatch - s.match;
    c.init2 = uint32_t(s.init + 32768) * 0x10001u;
    return c;
}

constexpr unsigned SW2_FLAG_WORDS = 6;           // sw2OffDiagonal, then fE, fF, fAB, fGF, fHE of a row: half A in the low, half B in the high 16 bits
constexpr unsigned SW2_PLANE_WORDS = 6;

/// max(a, b) of both halves, and bit J (half A) / bit 16+J (half B) of 'flags' set where the maximum is not a, i.e. a < b.
/// ptxas folds the max and the two equality tests into ONE VIMNMX.U16x2 with two predicate outputs; each flag then
/// costs one predicated add instead of subtract + min + shift-add.
template <int J>
__device__ __forceinline__ uint32_t sw2MaxFlag(const uint32_t a, const uint32_t b, uint32_t &flags)
{
    uint32_t m;
    asm("{.reg .pred pu, pv;\n\t.reg .u16 rs0, rs1, rs2, rs3;\n\t"
        "max.u16x2 %0, %2, %3;\n\tmov.b32 {rs0, rs1}, %0;\n\tmov.b32 {rs2, rs3}, %2;\n\t"
        "setp.eq.u16 pv, rs0, rs2;\n\tsetp.eq.u16 pu, rs1, rs3;\n\t"
        "@!pv or.b32 %1, %1, %4;\n\t@!pu or.b32 %1, %1, %5;}"
        : "=r"(m), "+r"(flags) : "r"(a), "r"(b), "n"(1u << J), "n"(0x10000u << J));
    return m;
}

template <int J> struct Sw2Lane { static constexpr int value = J; };

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

/// The six direction planes of a row from its five flag masks; every word holds the 16 lanes of half A in its low and
/// of half B in its high 16 bits, shifts by one lane are masked so that nothing crosses the halves.
__device__ __forceinline__ void sw2Planes(uint32_t fE, uint32_t fF, uint32_t fAB, uint32_t fGF, uint32_t fHE, uint32_t (&p)[SW2_PLANE_WORDS])
{
    const uint32_t EVEN = 0x55555555u, ODD = 0xAAAAAAAAu;
    // TG (:176-197).  Odd lane (high byte of the pair): F wins over E.  Even lane: the pair with the larger high byte
    // wins as a whole, so the odd lane's flags decide which of the even lane's own flags survive.
    const uint32_t hF = (fF >> 1) & EVEN, hE = (fE >> 1) & EVEN;                 // the odd lane's flags at the even lane's bit
    const uint32_t tg2 = (fF & ODD) | (fF & EVEN & (hF | ~hE));
    const uint32_t tg1 = (fE & ~fF & ODD) | (fE & EVEN & ~hF & (hE | ~fF));
    // TE (:261-297): 1 if e beats both, else 2 if f > g (flags of lane j+1)
    const uint32_t te1 = fHE;
    const uint32_t te2 = ~fHE & ((fGF >> 1) & 0x7FFF7FFFu);
    // TF (:142-167): 2 if a < b, else 1 if G[j-1] < E[j-1]; lane 0 is forced to 0
    const uint32_t tf2 = fAB & 0xFFFEFFFEu;
    const uint32_t tf1 = ~fAB & ((fE << 1) & 0xFFFEFFFEu);
    // stored per half: words 0..2 = (tg1 | tg2 << 16), (te1 | te2 << 16), (tf1 | tf2 << 16) of half A, words 3..5 of half B,
    // so that a walk reads three words per row and finds both planes of its state in one of them
    p[0] = __byte_perm(tg1, tg2, 0x5410); p[1] = __byte_perm(te1, te2, 0x5410); p[2] = __byte_perm(tf1, tf2, 0x5410);
    p[3] = __byte_perm(tg1, tg2, 0x7632); p[4] = __byte_perm(te1, te2, 0x7632); p[5] = __byte_perm(tf1, tf2, 0x7632);
}

/// The lanes of a row whose G direction is not "stay on the diagonal" (tg1 | tg2 of sw2Planes), from the two masks it depends on:
/// what a walk that is in G -- nearly every step of nearly every walk -- needs of a row.
__device__ __forceinline__ uint32_t sw2OffDiagonal(const uint32_t fE, const uint32_t fF)
{
    const uint32_t EVEN = 0x55555555u, ODD = 0xAAAAAAAAu;
    const uint32_t hF = (fF >> 1) & EVEN, hE = (fE >> 1) & EVEN;
    const uint32_t tg2 = (fF & ODD) | (fF & EVEN & (hF | ~hE));
    const uint32_t tg1 = (fE & ~fF & ODD) | (fE & EVEN & ~hF & (hE | ~fF));
    return tg1 | tg2;
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
    /// the general step(s) of this walk in row 'row' (:392-423) on the three plane words of its half (plane 1 in the low, plane 2 in
    /// the high 16 bits): one step, or several while the walk stays in the row (a deletion)
    __device__ __forceinline__ void stepsInRow(int row, const uint32_t tg, const uint32_t te, const uint32_t tf)
    {
        do
        {
            ++opLength;
            const uint32_t p = (type == 0 ? tg : (type == 1 ? te : tf)) >> jj;
            const unsigned next = (p & 1u) | ((p >> 15) & 2u);
            if (next != type) { push(opLength, type); opLength = 0; }
            ii -= int(type != 1);
            jj += int(type == 1) - int(type == 2);
            type = next;
            active = ii >= 0 && jj >= 0 && jj <= 15;
        } while (active && ii == row);
    }
    /// is the step of this walk in its current row "on the diagonal and staying there"?  offDiagonal = sw2OffDiagonal of the row
    __device__ __forceinline__ bool staysOnDiagonal(const uint32_t offDiagonal, const unsigned half) const
    {
        return type == 0 && ((offDiagonal >> (jj + int(half) * 16)) & 1u) == 0;
    }
    __device__ __forceinline__ void diagonalStep() { ++opLength; --ii; active = ii >= 0; }
    /// all steps of this walk in row 'row' from the six words of the row
    __device__ __forceinline__ void stepRow(int row, const uint32_t (&f)[SW2_FLAG_WORDS], const unsigned half)
    {
        if (!active || ii != row) return;
        if (staysOnDiagonal(f[0], half)) { diagonalStep(); return; }
        uint32_t p[SW2_PLANE_WORDS];
        sw2Planes(f[1], f[2], f[3], f[4], f[5], p);
        stepsInRow(row, p[half * 3u], p[half * 3u + 1u], p[half * 3u + 2u]);
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
/// by each walk, so the row order is known in advance): the six words of four rows ahead are kept in flight in
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
        a.stepRow(r, w, 0u);
        b.stepRow(r, w, 1u);
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

/// Forward pass over max(LA, LB) rows for the pair.  src.q2(i) / src.d2(k) return the base codes of both halves (they
/// must tolerate indices past the own length of a half: any code will do there).  On return jj/type hold the end cell
/// of each half.
/// ROW_RELATIVE: the cells are kept less (row + 1) * match (makeSw2Consts), which needs room for Lmax * match above the largest
/// score; where the scores and the read length do not leave it (sw2RowRelativeFits) the plain values are used, one add per cell more
template <bool ROW_RELATIVE, class PairSrc>
__device__ __forceinline__ void sw2Forward(PairSrc &src, const unsigned LA, const unsigned LB, const SwScores s,
                                           uint32_t *__restrict__ tb, const size_t tbStride,
                                           int (&jj)[2], unsigned (&type)[2])
{
    Sw2Consts c = makeSw2Consts(s);
    if (!ROW_RELATIVE) { c.negOpenRow32 = c.negOpen32; c.negExtRow16x2 = c.negExt16x2; }
    const unsigned Lmax = max(LA, LB);
    // the constant that keeps the row-relative values non-negative: the last row has dropped by Lmax * match (the host admits only
    // scores and read lengths for which value + 32768 + Lmax * match stays below 65536, swScoresSupported)
    const uint32_t lift = ROW_RELATIVE ? Lmax * c.match32 : 0u;
    uint32_t initRow = c.init2 + lift;                              // 'init' as the row about to be computed holds it
    uint32_t G[16], E[16], F[16], D[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { G[j] = initRow; E[j] = initRow; F[j] = 0x80008000u + lift; D[j] = 0; }   // :108-114
    G[0] = 0x80008000u + lift;                                                                        // :115
    // D[j] = database codes seen by lane j = db[i + 15 - j]; preload db[0..14] (:117-122)
#pragma unroll
    for (int k = 0; k < 15; ++k) D[14 - k] = src.d2(k);
    jj[0] = int(LA) - 1; jj[1] = int(LB) - 1; type[0] = 0; type[1] = 0;
#pragma unroll 1
    for (unsigned i = 0; i < Lmax; ++i)
    {
        if (ROW_RELATIVE) initRow -= c.match32;
#pragma unroll
        for (int j = 15; j > 0; --j) D[j] = D[j - 1];
        D[0] = src.d2(i + 15);
        const uint32_t Q = src.q2(i);
        uint32_t fE = 0, fF = 0, fAB = 0, fGF = 0, fHE = 0;
        uint32_t mCur = sw2MaxFlag<15>(G[15], E[15], fE);
        uint32_t hOnext = 0, nEnext = 0;
        auto lane = [&](auto laneTag) {
            constexpr int j = decltype(laneTag)::value;
            // ---- F of lane j from lane j-1 of the previous row (:132-173)
            uint32_t nF, mPrev = 0;
            if (j > 0)
            {
                constexpr int jm = j > 0 ? j - 1 : 0;
                mPrev = sw2MaxFlag<jm>(G[jm], E[jm], fE);
                const uint32_t a = mPrev + c.negOpenRow32;
                nF = __viaddmax_u16x2(F[jm], c.negExtRow16x2, a);
                fAB = fAB * 2u + __vminu2(nF - a, 0x00010001u);
            }
            else { nF = initRow; fAB = fAB * 2u; }                                // :167, :173
            // ---- G of lane j from the same lane (:176-190, :230-244)
            const uint32_t g = sw2MaxFlag<j>(mCur, F[j], fF);
            const uint32_t t = __vminu2(D[j] ^ Q, 0x00010001u);
            const uint32_t nG = ROW_RELATIVE ? g + t * uint32_t(c.delta) : g + c.match32 + t * uint32_t(c.delta);
            // ---- E of lane j from lane j+1 of THIS row (:261-297)
            uint32_t nE;
            if (j < 15)
            {
                nE = __viaddmax_u16x2(nEnext, c.negExt16x2, hOnext);
                fHE = fHE * 2u + __vminu2(nE - hOnext, 0x00010001u);
            }
            else { nE = initRow; fHE = fHE * 2u; }
            const uint32_t h = sw2MaxFlag<j>(nG, nF, fGF);
            hOnext = h + c.negOpen32;
            nEnext = nE;
            G[j] = nG; E[j] = nE; F[j] = nF;
            mCur = mPrev;
        };
        lane(Sw2Lane<15>()); lane(Sw2Lane<14>()); lane(Sw2Lane<13>()); lane(Sw2Lane<12>());
        lane(Sw2Lane<11>()); lane(Sw2Lane<10>()); lane(Sw2Lane<9>()); lane(Sw2Lane<8>());
        lane(Sw2Lane<7>()); lane(Sw2Lane<6>()); lane(Sw2Lane<5>()); lane(Sw2Lane<4>());
        lane(Sw2Lane<3>()); lane(Sw2Lane<2>()); lane(Sw2Lane<1>()); lane(Sw2Lane<0>());
        // the five masks leave as they are (:306-308 stores direction bytes), in front of them the one word nearly every step of
        // nearly every walk needs: the lanes whose G direction leaves the diagonal.  Turning the masks into directions is left
        // to the walk, for the few rows in which it needs them
        const uint32_t flags[SW2_FLAG_WORDS] = {sw2OffDiagonal(fE, fF), fE, fF, fAB, fGF, fHE};
        uint32_t *row = tb + size_t(i) * SW2_FLAG_WORDS * tbStride;
#pragma unroll
        for (unsigned k = 0; k < SW2_FLAG_WORDS; ++k) row[k * tbStride] = flags[k];
        if (i + 1 == LA || i + 1 == LB)                        // last row of a half: its end cell (one copy of the scan)
        {
#pragma unroll 1
            for (unsigned half = 0; half < 2; ++half)
                if (i + 1 == (half ? LB : LA)) sw2ScanEnd(G, E, F, half, jj[half], type[half]);
        }
    }
}

} // namespace isaac_b200
