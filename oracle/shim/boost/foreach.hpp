// Stand-in for BOOST_FOREACH on top of range-for; also accepts std::pair<It,It> as a range like Boost does.
#pragma once
#include <utility>
namespace boost_shim {
template <class It> struct PairRange { It b, e; It begin() const { return b; } It end() const { return e; } };
template <class C> C &rng(C &c) { return c; }
template <class C> const C &rng(const C &c) { return c; }
template <class It> PairRange<It> rng(const std::pair<It, It> &p) { return PairRange<It>{p.first, p.second}; }
}
#define BOOST_FOREACH(decl, col) for (decl : boost_shim::rng(col))
