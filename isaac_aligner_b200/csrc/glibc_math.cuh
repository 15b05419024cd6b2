// exp() and log10() exactly as glibc 2.39 computes them on an x86-64 host with FMA, for host AND device code.
//
// Why: alignment::TemplateBuilder turns probabilities into mapping scores with floor(-10 * log10(other / total)) where the
// operands are sums of exp(logProbability) (TemplateBuilder.cpp:233-285,398-465,495-676,868-1008).  (total - exp(best)) cancels,
// so the LAST bit of every exp() decides integers: a device-side exp() that is merely accurate (CUDA's is, to 1 ulp) would give
// different alignment scores than the reference for a few clusters per million.  Bit-exact mapping scores on the device need
// the host library's own results.  glibc's exp / log have been Szabolcs Nagy's table-driven "optimized-routines" code since
// 2.28 (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c, tables e_exp_data.c / e_log_data.c); log10 is the older
// __ieee754_log10 (e_log10.c) on top of that log.  On x86-64 the ifunc resolvers pick the variants compiled with -mfma -mavx2
// (__exp_fma, __log_fma) on every CPU that has FMA + AVX2; this header replays the instruction sequence of exactly those
// variants (which products are fused is the compiler's choice and was read off the disassembly), with the tables taken from
// the same library (glibc_math_tables.inc, tools/extract_glibc_tables.py).  IEEE double add / multiply / fma are correctly
// rounded on both sides, so equal sequences give equal bits.  tests/cpp/test_glibc_math.cpp compares the host build of these
// functions with the libm of the box for hundreds of millions of arguments; tests/test_gpu_glibc_math.py does the same for the
// device build through isaac_ext_selftest_glibc_math.
//
// Domain: every double.  NaN payloads and the errno side effects of the library wrappers are not reproduced.
#pragma once
#include <cstdint>
#include <cstring>

#ifndef ISAAC_HD
#ifdef __CUDACC__
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif
#endif

namespace isaac_b200
{
namespace glibc_math
{

// the tables once for host code and once more, under nvcc, in device memory (plain global memory: the lookups of a warp
// diverge, constant memory would serialise them)
#define ISAAC_GLIBC_TABLE(name, n) static const uint64_t name[n]
#include "glibc_math_tables.inc"
#undef ISAAC_GLIBC_TABLE
#ifdef __CUDACC__
#define ISAAC_GLIBC_TABLE(name, n) static __device__ const uint64_t name##_DEVICE[n]
#include "glibc_math_tables.inc"
#undef ISAAC_GLIBC_TABLE
#endif
#ifdef __CUDA_ARCH__
#define ISAAC_GLIBC_WORD(name, i) (name##_DEVICE[i])
#else
#define ISAAC_GLIBC_WORD(name, i) (name[i])
#endif

ISAAC_HD inline uint64_t toBits(const double x)
{
#ifdef __CUDA_ARCH__
    return uint64_t(__double_as_longlong(x));
#else
    uint64_t u; std::memcpy(&u, &x, sizeof(u)); return u;
#endif
}
ISAAC_HD inline double fromBits(const uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; std::memcpy(&x, &u, sizeof(x)); return x;
#endif
}
#define ISAAC_GLIBC_DOUBLE(name, i) fromBits(ISAAC_GLIBC_WORD(name, i))

// one IEEE operation each, never fused with a neighbour (the host compiler contracts a * b + c on its own with -mfma)
ISAAC_HD inline double mul(const double a, const double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    double r = a * b;
#if defined(__x86_64__)
    asm volatile("" : "+x"(r));
#else
    volatile double v = r; r = v;
#endif
    return r;
#endif
}
ISAAC_HD inline double add(const double a, const double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    double r = a + b;
#if defined(__x86_64__)
    asm volatile("" : "+x"(r));
#else
    volatile double v = r; r = v;
#endif
    return r;
#endif
}
ISAAC_HD inline double sub(const double a, const double b) { return add(a, -b); }
/// a * b + c with one rounding
ISAAC_HD inline double fma(const double a, const double b, const double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

/// __exp_fma of glibc 2.39 (e_exp.c: exp = 2^(k/128) * exp(r), 128-entry table, degree-5 polynomial)
ISAAC_HD inline double exp(const double x)
{
    const uint64_t ix = toBits(x);
    uint32_t abstop = uint32_t(ix >> 52) & 0x7ffu;
    if (abstop - 0x3c9u > 0x3eu)                                    // |x| < 2^-54 or |x| >= 512
    {
        if (int32_t(abstop - 0x3c9u) < 0) return add(x, 1.0);       // tiny: 1 + x
        if (abstop > 0x408u)                                        // |x| >= 1024, inf, nan
        {
            if (ix == 0xfff0000000000000ull) return 0.0;            // exp(-inf)
            if (abstop == 0x7ffu) return add(x, 1.0);               // +inf, nan
            return int64_t(ix) < 0 ? 0.0 : fromBits(0x7ff0000000000000ull);   // __math_uflow(0) / __math_oflow(0)
        }
        abstop = 0;                                                 // 512 <= |x| < 1024: the result may over/underflow
    }
    const double invLn2N = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 0), shift = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 1);
    const double negLn2hiN = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 2), negLn2loN = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 3);
    const double C2 = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 4), C3 = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 5);
    const double C4 = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 6), C5 = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_CONST, 7);
    double kd = fma(x, invLn2N, shift);                             // z + Shift in one rounding
    const uint64_t ki = toBits(kd);
    kd = sub(kd, shift);
    double r = fma(kd, negLn2hiN, x);
    r = fma(kd, negLn2loN, r);
    const unsigned idx = 2u * unsigned(ki & 127u);
    const uint64_t top = ki << 45;
    const double tail = ISAAC_GLIBC_DOUBLE(GLIBC_EXP_TAB, idx);
    uint64_t sbits = ISAAC_GLIBC_WORD(GLIBC_EXP_TAB, idx + 1) + top;
    const double p23 = fma(r, C3, C2);
    const double tailr = add(r, tail);
    const double r2 = mul(r, r);
    const double p45 = fma(r, C5, C4);
    const double low = fma(p23, r2, tailr);
    const double r4 = mul(r2, r2);
    const double tmp = fma(r4, p45, low);
    if (abstop == 0)                                                // specialcase()
    {
        if ((ki & 0x80000000ull) == 0)                              // k > 0: scale by 2^-1009 first
        {
            sbits -= uint64_t(1009) << 52;
            const double scale = fromBits(sbits);
            return mul(fma(scale, tmp, scale), fromBits(uint64_t(0x3ff + 1009) << 52));
        }
        sbits += uint64_t(1022) << 52;                              // k < 0: the result may be subnormal
        const double scale = fromBits(sbits);
        const double p = mul(tmp, scale);
        double y = add(scale, p);
        if (y < 1.0)
        {
            const double hi = add(y, 1.0);
            double lo = sub(scale, y);
            lo = add(lo, p);
            double t = sub(1.0, hi);
            t = add(t, y);
            t = add(t, lo);
            t = add(t, hi);
            y = sub(t, 1.0);
            if (y == 0.0) y = 0.0;                                  // no -0
        }
        return mul(y, fromBits(uint64_t(0x3ff - 1022) << 52));
    }
    const double scale = fromBits(sbits);
    return fma(scale, tmp, scale);
}

/// __log_fma of glibc 2.39 (e_log.c) for the arguments __ieee754_log10 hands it: normal, 0.5 <= x < 2
ISAAC_HD inline double logNormalized(const double x)
{
    const uint64_t ix = toBits(x);
    if (ix - 0x3fee000000000000ull < 0x0003090000000000ull)         // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    {
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = sub(x, 1.0);
#define ISAAC_B(i) ISAAC_GLIBC_DOUBLE(GLIBC_LOG_CONST, 7 + (i))
        double p12 = fma(r, ISAAC_B(2), ISAAC_B(1));
        double p45 = fma(r, ISAAC_B(5), ISAAC_B(4));
        const double r2 = mul(r, r);
        double p78 = fma(r, ISAAC_B(8), ISAAC_B(7));
        p12 = fma(r2, ISAAC_B(3), p12);
        p45 = fma(r2, ISAAC_B(6), p45);
        const double r3 = mul(r, r2);
        p78 = fma(r2, ISAAC_B(9), p78);
        p78 = fma(r3, ISAAC_B(10), p78);
        double p = fma(p78, r3, p45);
        p = fma(p, r3, p12);
        const double two27 = fromBits(uint64_t(0x3ff + 27) << 52);
        const double w = fma(r, two27, r);
        const double rhi = fma(-two27, r, w);
        const double rhi2 = mul(rhi, rhi);
        const double rlo = sub(r, rhi);
        const double hi = fma(rhi2, ISAAC_B(0), r);
        const double rMinusHi = sub(r, hi);
        const double rPlusRhi = add(r, rhi);
        double lo = fma(rhi2, ISAAC_B(0), rMinusHi);
        lo = fma(mul(ISAAC_B(0), rlo), rPlusRhi, lo);
        const double y = fma(p, r3, lo);
#undef ISAAC_B
        return add(hi, y);
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const unsigned i = unsigned(tmp >> 45) & 127u;
    const int k = int(int64_t(tmp) >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = ISAAC_GLIBC_DOUBLE(GLIBC_LOG_TAB, 2 * i), logc = ISAAC_GLIBC_DOUBLE(GLIBC_LOG_TAB, 2 * i + 1);
    const double z = fromBits(iz);
    const double kd = double(k);
    const double ln2hi = ISAAC_GLIBC_DOUBLE(GLIBC_LOG_CONST, 0), ln2lo = ISAAC_GLIBC_DOUBLE(GLIBC_LOG_CONST, 1);
#define ISAAC_A(i) ISAAC_GLIBC_DOUBLE(GLIBC_LOG_CONST, 2 + (i))
    const double w = fma(kd, ln2hi, logc);
    const double r = fma(z, invc, -1.0);
    const double p12 = fma(r, ISAAC_A(2), ISAAC_A(1));
    const double hi = add(r, w);
    const double r2 = mul(r, r);
    double lo = sub(w, hi);
    lo = add(lo, r);
    lo = fma(kd, ln2lo, lo);
    const double r3 = mul(r, r2);
    const double p34 = fma(r, ISAAC_A(4), ISAAC_A(3));
    lo = fma(r2, ISAAC_A(0), lo);
    const double p = fma(p34, r2, p12);
    const double y = fma(r3, p, lo);
#undef ISAAC_A
    return add(hi, y);
}

/// log10() of glibc 2.39: the wrapper (w_log10_compat.c) around __ieee754_log10 (e_log10.c)
ISAAC_HD inline double log10(double x)
{
    int64_t hx = int64_t(toBits(x));
    int k = -1023;
    if (hx < int64_t(0x0010000000000000ll))
    {
        if ((uint64_t(hx) & 0x7fffffffffffffffull) == 0) return fromBits(0xfff0000000000000ull);   // log10(+-0) = -inf
        if (hx < 0) return fromBits(0x7ff8000000000000ull);                                        // log10(negative) = nan
        x = mul(x, ISAAC_GLIBC_DOUBLE(GLIBC_LOG10_CONST, 3));                                      // subnormal: scale up by 2^54
        hx = int64_t(toBits(x));
        k = -1077;
    }
    if (uint64_t(hx) > 0x7fefffffffffffffull) return add(x, x);                                    // inf, nan
    k += int(hx >> 52);
    const int64_t i = k < 0 ? 1 : 0;
    const uint64_t mantissa = uint64_t(hx) & 0x000fffffffffffffull;
    const double y = double(int64_t(k) + i);
    const double normalized = fromBits(mantissa | (uint64_t(0x3ff - i) << 52));
    const double t1 = mul(y, ISAAC_GLIBC_DOUBLE(GLIBC_LOG10_CONST, 0));                            // y * log10_2lo
    const double l = logNormalized(normalized);
    const double t2 = mul(y, ISAAC_GLIBC_DOUBLE(GLIBC_LOG10_CONST, 2));                            // y * log10_2hi
    return add(add(mul(l, ISAAC_GLIBC_DOUBLE(GLIBC_LOG10_CONST, 1)), t1), t2);
}

} // namespace glibc_math
} // namespace isaac_b200
