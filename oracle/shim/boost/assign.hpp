// Stand-in for boost::assign::list_of(x)(y)... convertible to vector / std::array.
#pragma once
#include <vector>
#include <array>
#include <string>
namespace boost { namespace assign {
template <class T> struct ListOf {
    std::vector<T> v;
    ListOf &operator()(const T &t) { v.push_back(t); return *this; }
    template <class U> operator std::vector<U>() const { return std::vector<U>(v.begin(), v.end()); }
    template <class U, std::size_t N> operator std::array<U, N>() const { std::array<U, N> a{}; for (std::size_t i = 0; i < N && i < v.size(); ++i) a[i] = v[i]; return a; }
};
template <class T> ListOf<T> list_of(const T &t) { ListOf<T> l; l.v.push_back(t); return l; }
inline ListOf<std::string> list_of(const char *t) { ListOf<std::string> l; l.v.push_back(t); return l; }
} }
