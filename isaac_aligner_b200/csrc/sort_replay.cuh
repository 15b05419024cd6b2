// libstdc++'s std::sort replayed step by step, callable from host and device code.
//
// Why: FragmentBuilder::consolidateDuplicateFragments (FragmentBuilder.cpp:279-324) std::sorts a read's candidate list with a
// comparator under which many entries are equivalent (same contig, position, strand, observed length) and then keeps the FIRST
// entry of every group of equivalents -- its firstSeedIndex and seed bookkeeping survive.  std::sort is not
// stable, so which entry comes first depends on the exact sequence of comparisons and swaps of the library's introsort.  The
// host phases of isaac_ext_build_fragments call std::sort itself (the same libstdc++ the reference is built with); moving
// those phases to the device needs the same permutation there.  This header restates
// bits/stl_algo.h / bits/stl_heap.h of GCC's libstdc++ (__introsort_loop with its median-of-three pivot and unguarded
// partition, the heap sort it falls back to when the depth limit 2 * floor(log2 n) is used up, __final_insertion_sort with
// its threshold of 16) on a plain pointer range.  tests/cpp/test_sort_replay.cpp checks on the CPU that it produces the same
// permutation as std::sort, element for element, on millions of lists full of equivalent keys.
//
// Used by consolidate_device.cuh (the B1-B4 kernels of kernels_tile.cuh) and by finish_device.cuh (the probability lists of TemplateBuilder).
#pragma once

#if defined(__CUDACC__)
#define ISAAC_HD __host__ __device__
#else
#define ISAAC_HD
#endif

namespace isaac_b200
{
namespace sort_replay
{

constexpr int THRESHOLD = 16;            // std::_S_threshold

template <class T> ISAAC_HD inline void swapValues(T &a, T &b) { const T t = a; a = b; b = t; }

/// std::__move_median_to_first
template <class T, class Less> ISAAC_HD inline void moveMedianToFirst(T *result, T *a, T *b, T *c, Less less)
{
    if (less(*a, *b))
    {
        if (less(*b, *c)) swapValues(*result, *b);
        else if (less(*a, *c)) swapValues(*result, *c);
        else swapValues(*result, *a);
    }
    else if (less(*a, *c)) swapValues(*result, *a);
    else if (less(*b, *c)) swapValues(*result, *c);
    else swapValues(*result, *b);
}

/// std::__unguarded_partition
template <class T, class Less> ISAAC_HD inline T *unguardedPartition(T *first, T *last, T *pivot, Less less)
{
    while (true)
    {
        while (less(*first, *pivot)) ++first;
        --last;
        while (less(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swapValues(*first, *last);
        ++first;
    }
}

/// std::__push_heap
template <class T, class Less> ISAAC_HD inline void pushHeap(T *first, long holeIndex, long topIndex, T value, Less less)
{
    long parent = (holeIndex - 1) / 2;
    while (holeIndex > topIndex && less(first[parent], value))
    {
        first[holeIndex] = first[parent];
        holeIndex = parent;
        parent = (holeIndex - 1) / 2;
    }
    first[holeIndex] = value;
}

/// std::__adjust_heap
template <class T, class Less> ISAAC_HD inline void adjustHeap(T *first, long holeIndex, long len, T value, Less less)
{
    const long topIndex = holeIndex;
    long secondChild = holeIndex;
    while (secondChild < (len - 1) / 2)
    {
        secondChild = 2 * (secondChild + 1);
        if (less(first[secondChild], first[secondChild - 1])) --secondChild;
        first[holeIndex] = first[secondChild];
        holeIndex = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2)
    {
        secondChild = 2 * (secondChild + 1);
        first[holeIndex] = first[secondChild - 1];
        holeIndex = secondChild - 1;
    }
    pushHeap(first, holeIndex, topIndex, value, less);
}

/// std::__partial_sort(first, last, last): __make_heap + __sort_heap (the __heap_select loop over [middle, last) is empty)
template <class T, class Less> ISAAC_HD inline void heapSort(T *first, T *last, Less less)
{
#ifdef ISAAC_SORT_REPLAY_COUNT_HEAPSORT
    ++ISAAC_SORT_REPLAY_COUNT_HEAPSORT;              // test hook: how often the depth limit was reached
#endif
    const long len = last - first;
    if (len >= 2)
    {
        long parent = (len - 2) / 2;
        while (true)
        {
            const T value = first[parent];
            adjustHeap(first, parent, len, value, less);
            if (parent == 0) break;
            --parent;
        }
    }
    while (last - first > 1)
    {
        --last;
        const T value = *last;                       // std::__pop_heap(first, last, last)
        *last = *first;
        adjustHeap(first, 0L, long(last - first), value, less);
    }
}

/// std::__unguarded_linear_insert
template <class T, class Less> ISAAC_HD inline void unguardedLinearInsert(T *last, Less less)
{
    const T value = *last;
    T *next = last - 1;
    while (less(value, *next))
    {
        *last = *next;
        last = next;
        --next;
    }
    *last = value;
}

/// std::__insertion_sort
template <class T, class Less> ISAAC_HD inline void insertionSort(T *first, T *last, Less less)
{
    if (first == last) return;
    for (T *i = first + 1; i != last; ++i)
    {
        if (less(*i, *first))
        {
            const T value = *i;
            for (T *p = i; p != first; --p) *p = *(p - 1);         // std::move_backward(first, i, i + 1)
            *first = value;
        }
        else unguardedLinearInsert(i, less);
    }
}

/// std::sort(first, first + n, less) of libstdc++: same comparisons, same moves, same result
template <class T, class Less> ISAAC_HD inline void sort(T *first, const unsigned n, Less less)
{
    if (!n) return;
    T *last = first + n;
    // std::__introsort_loop with its recursion on the right part turned into a stack (at most 2 * lg(n) frames)
    struct Frame { T *first, *last; int depth; };
    Frame stack[66];
    int top = 0;
    int lg = 0;
    for (unsigned k = n; k > 1; k >>= 1) ++lg;                     // std::__lg
    stack[top++] = Frame{first, last, 2 * lg};
    while (top)
    {
        Frame f = stack[--top];
        while (f.last - f.first > THRESHOLD)
        {
            if (f.depth == 0) { heapSort(f.first, f.last, less); break; }
            --f.depth;
            T *mid = f.first + (f.last - f.first) / 2;             // std::__unguarded_partition_pivot
            moveMedianToFirst(f.first, f.first + 1, mid, f.last - 1, less);
            T *cut = unguardedPartition(f.first + 1, f.last, f.first, less);
            // the library recurses into [cut, last) first and then loops on [first, cut): the two ranges are disjoint, so the
            // order in which they are finished does not change what happens inside either of them
            stack[top++] = Frame{cut, f.last, f.depth};
            f.last = cut;
        }
    }
    // std::__final_insertion_sort
    if (last - first > THRESHOLD)
    {
        insertionSort(first, first + THRESHOLD, less);
        for (T *i = first + THRESHOLD; i != last; ++i) unguardedLinearInsert(i, less);
    }
    else insertionSort(first, last, less);
}

} // namespace sort_replay
} // namespace isaac_b200
