// isaac_aligner_b200/csrc/glibc_math.cuh (host build) against the libm of this box: exp() and log10() must agree bit for bit on
// every argument the template scores can produce and on random bit patterns.  Usage: test_glibc_math [millions of samples]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../isaac_aligner_b200/csrc/glibc_math.cuh"

using namespace isaac_b200;

static uint64_t bitsOf(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static double ofBits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static bool same(double a, double b) { return bitsOf(a) == bitsOf(b) || (a != a && b != b); }

int main(int argc, char **argv)
{
    const long millions = argc > 1 ? std::atol(argv[1]) : 20;
    std::mt19937_64 rng(20261017);
    long badExp = 0, badLog = 0, n = 0;
    auto checkExp = [&](double x) {
        volatile double vx = x;                       // keep the library call a library call
        const double want = std::exp(vx), got = glibc_math::exp(x);
        if (!same(want, got) && badExp++ < 10) std::printf("exp(%a): libm %a replay %a\n", x, want, got);
    };
    auto checkLog = [&](double x) {
        volatile double vx = x;
        const double want = std::log10(vx), got = glibc_math::log10(x);
        if (!same(want, got) && badLog++ < 10) std::printf("log10(%a): libm %a replay %a\n", x, want, got);
    };
    // hand-picked: zeros, ones, thresholds of every branch
    const double special[] = {0.0, -0.0, 1.0, -1.0, 0x1p-54, -0x1p-54, 0x1p-55, 511.9999, 512.0, -512.0, 709.78, 709.79, -708.3, -708.4, -745.13,
                              -745.14, -1023.9, -1024.0, 1024.0, 1e308, -1e308, INFINITY, -INFINITY, 0x1p-1022, 0x1p-1074, 0.5, 2.0, 10.0, 0.1,
                              1.0 - 0x1p-4, 1.0 + 0x1.09p-4, 0x1.fffffffffffffp-1, 0x1.0000000000001p0, 1e-300, 1e-310, 0.9375, 1.0646972656};
    for (double x : special) { checkExp(x); checkLog(x); checkLog(-x); }
    std::uniform_real_distribution<double> lp(-1100.0, 0.0), wide(-1100.0, 720.0), unit(0.0, 1.0), nearOne(0.9, 1.1);
    for (long i = 0; i < millions * 1000000; ++i, ++n)
    {
        // exp: log probabilities (sums of a few hundred table entries), the whole finite range, random bit patterns
        checkExp(lp(rng));
        if ((i & 3) == 0) checkExp(wide(rng));
        if ((i & 7) == 0) checkExp(ofBits(rng()));
        // log10: ratios in (0, 1], exponents across the whole range incl. subnormals, around 1, random bit patterns
        const double u = unit(rng);
        checkLog(u);
        checkLog(std::ldexp(u, -int(rng() % 1080)));
        if ((i & 3) == 0) checkLog(nearOne(rng));
        if ((i & 7) == 0) checkLog(std::fabs(ofBits(rng())));
    }
    std::printf("%ld rounds, exp mismatches %ld, log10 mismatches %ld\n", n, badExp, badLog);
    if (badExp || badLog) return 1;
    std::printf("all checks passed\n");
    return 0;
}
