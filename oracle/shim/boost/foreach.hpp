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
#include <iterator>
namespace boost_shim {
template <class It> struct ReversedRange { std::reverse_iterator<It> b, e; std::reverse_iterator<It> begin() const { return b; } std::reverse_iterator<It> end() const { return e; } };
template <class It> ReversedRange<It> rrng(const std::pair<It, It> &p) { return ReversedRange<It>{std::reverse_iterator<It>(p.second), std::reverse_iterator<It>(p.first)}; }
template <class C> auto rrng(C &c) -> ReversedRange<decltype(c.begin())> { return ReversedRange<decltype(c.begin())>{std::reverse_iterator<decltype(c.begin())>(c.end()), std::reverse_iterator<decltype(c.begin())>(c.begin())}; }
}
#define BOOST_REVERSE_FOREACH(decl, col) for (decl : boost_shim::rrng(col))
