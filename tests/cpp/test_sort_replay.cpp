// isaac_aligner_b200/csrc/sort_replay.cuh against the std::sort of this box's libstdc++ (the one the reference checker is built
// with): the SAME permutation, element for element, on lists full of equivalent keys -- random lists of every length up to a
// few hundred, sorted / reversed / organ-pipe / few-distinct-keys patterns, the FragmentMetadata ordering of the reference
// (FragmentMetadata.hh:419-429), and median-of-three killer sequences that drive the introsort into its heap-sort fallback.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static long heapSortCalls = 0;
#define ISAAC_SORT_REPLAY_COUNT_HEAPSORT heapSortCalls
#include "../../isaac_aligner_b200/csrc/sort_replay.cuh"

struct Item { int key; int id; };
struct Fragment { uint32_t contigId; int64_t position; uint8_t reverse; uint32_t observedLength; int id; };

static bool keyLess(const Item &a, const Item &b) { return a.key < b.key; }
static bool fragmentLess(const Fragment &a, const Fragment &b)                 // FragmentMetadata::operator<
{
    return a.contigId < b.contigId ||
           (a.contigId == b.contigId &&
            (a.position < b.position ||
             (a.position == b.position && (a.reverse < b.reverse || (a.reverse == b.reverse && a.observedLength < b.observedLength)))));
}

static uint64_t state = 0x15AAC0DEull;
static unsigned rnd(unsigned n) { state = state * 6364136223846793005ull + 1442695040888963407ull; return unsigned((state >> 33) % n); }

static long failures = 0, lists = 0;

static void checkItems(std::vector<Item> v)
{
    for (size_t i = 0; i < v.size(); ++i) v[i].id = int(i);
    std::vector<Item> a = v, b = v;
    std::sort(a.begin(), a.end(), keyLess);
    isaac_b200::sort_replay::sort(b.data(), unsigned(b.size()), keyLess);
    ++lists;
    for (size_t i = 0; i < v.size(); ++i)
        if (a[i].key != b[i].key || a[i].id != b[i].id) { if (++failures < 5) std::printf("FAILED: list of %zu differs at %zu\n", v.size(), i); return; }
}

/// Musser's median-of-three killer: quadratic partitions for a median-of-3 quicksort, so that introsort runs out of depth
static std::vector<Item> killer(unsigned n)
{
    n &= ~1u;
    std::vector<Item> v(n);
    const unsigned k = n / 2;
    for (unsigned i = 1; i <= k; ++i)
    {
        if (i % 2) { v[i - 1].key = int(i); v[i].key = int(k + i); }
        v[k + i - 1].key = int(2 * i);
    }
    return v;
}

int main()
{
    // random lists, many equivalent keys
    for (unsigned n = 0; n <= 300; ++n)
        for (unsigned distinct : {1u, 2u, 3u, 5u, 17u, 1000u})
            for (int rep = 0; rep < 40; ++rep)
            {
                std::vector<Item> v(n);
                for (Item &x : v) x.key = int(rnd(distinct));
                checkItems(v);
            }
    // patterns
    for (unsigned n : {17u, 33u, 64u, 100u, 257u, 1000u, 5000u})
    {
        std::vector<Item> v(n);
        for (unsigned i = 0; i < n; ++i) v[i].key = int(i);
        checkItems(v);
        for (unsigned i = 0; i < n; ++i) v[i].key = int(n - i);
        checkItems(v);
        for (unsigned i = 0; i < n; ++i) v[i].key = int(i < n / 2 ? i : n - i);
        checkItems(v);
        for (unsigned i = 0; i < n; ++i) v[i].key = int(i % 4);
        checkItems(v);
        checkItems(killer(n));
        std::vector<Item> k2 = killer(n);
        for (Item &x : k2) x.key /= 3;                                      // the killer with groups of equivalent keys
        checkItems(k2);
    }
    const long heapSorts = heapSortCalls;
    // the reference's candidate ordering: lists like FragmentBuilder sees them (a few loci, both strands, repeated by several seeds)
    for (int rep = 0; rep < 200000; ++rep)
    {
        const unsigned n = rnd(90);
        std::vector<Fragment> v(n);
        for (unsigned i = 0; i < n; ++i)
        {
            v[i].contigId = rnd(2); v[i].position = 1000 + int64_t(rnd(6)) * (rnd(2) ? 1 : 37); v[i].reverse = uint8_t(rnd(2));
            v[i].observedLength = rnd(3) ? 100 : 100 - rnd(3); v[i].id = int(i);
        }
        std::vector<Fragment> a = v, b = v;
        std::sort(a.begin(), a.end(), fragmentLess);
        isaac_b200::sort_replay::sort(b.data(), n, fragmentLess);
        ++lists;
        for (unsigned i = 0; i < n; ++i)
            if (a[i].id != b[i].id) { if (++failures < 5) std::printf("FAILED: fragment list of %u differs at %u\n", n, i); break; }
    }
    if (!heapSorts) { std::printf("FAILED: the heap-sort fallback was never reached\n"); ++failures; }
    std::printf("%ld lists, heap-sort fallback reached %ld times, %ld failures\n", lists, heapSorts, failures);
    std::printf(failures ? "FAILED\n" : "all checks passed\n");
    return failures ? 1 : 0;
}
