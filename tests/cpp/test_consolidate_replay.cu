// consolidate_device.cuh against consolidateDuplicateFragments of host_consolidate_reference.hh (std::sort of this box's libstdc++), on the CPU:
// random candidate lists like FragmentBuilder sees them (a few loci hit by several seeds, both strands, aligned and unaligned
// entries), both values of removeUnaligned: the same surviving records, byte for byte.
#include <cstdio>
#include <cstring>
#include <vector>

#include "host_consolidate_reference.hh"
#include "../../isaac_aligner_b200/csrc/consolidate_device.cuh"

using namespace isaac_b200;

static uint64_t state = 0xC0501DA7Eull;
static unsigned rnd(unsigned n) { state = state * 6364136223846793005ull + 1442695040888963407ull; return unsigned((state >> 33) % n); }

int main()
{
    long failures = 0, lists = 0, merged = 0;
    for (int rep = 0; rep < 300000; ++rep)
    {
        const unsigned n = rnd(rep % 50 == 0 ? 200 : 40);
        std::vector<WorkFragment> v(n);
        for (unsigned i = 0; i < n; ++i)
        {
            WorkFragment &w = v[i];
            std::memset(&w, 0, sizeof(w));
            w.f.contigId = rnd(2); w.f.position = 5000 + long(rnd(5)) * (rnd(3) ? 1 : 211); w.f.reverse = uint8_t(rnd(2));
            w.f.cigarLength = uint16_t(rnd(5) ? 1 + rnd(3) : 0);
            w.f.observedLength = w.f.cigarLength ? (rnd(4) ? 100u : 100u - rnd(3)) : 0u;
            w.f.uniqueSeedCount = uint16_t(rnd(3)); w.f.firstSeedIndex = int16_t(rnd(8));
            w.f.nonUniqueSeedOffsetFirst = rnd(3) ? uint16_t(0xFFFF) : uint16_t(rnd(100)); w.f.nonUniqueSeedOffsetSecond = uint16_t(rnd(100));
            w.f.readId = i; w.pool = rnd(3); w.slot = i;
        }
        const bool removeUnaligned = rnd(2) != 0;
        std::vector<WorkFragment> a = v, b = v;
        const unsigned na = consolidateDuplicateFragments(a.data(), n, removeUnaligned);
        const unsigned nb = consolidateDuplicateFragmentsReplay(b.data(), n, removeUnaligned);
        ++lists; merged += n - na;
        if (na != nb || (na && std::memcmp(a.data(), b.data(), size_t(na) * sizeof(WorkFragment)) != 0))
            if (++failures < 5) std::printf("FAILED: list %d of %u entries: %u vs %u survivors\n", rep, n, na, nb);
    }
    std::printf("%ld lists, %ld entries merged or removed, %ld failures\n", lists, merged, failures);
    std::printf(failures ? "FAILED\n" : "all checks passed\n");
    return failures ? 1 : 0;
}
