// Stand-in for boost::bind / ref / cref on top of <functional>.
#pragma once
#include <functional>
#include <type_traits>
#include <utility>
namespace boost {
using std::bind;
using std::ref;
// the reference takes the address of boost::cref<char>, so it must be a real function template
template <class T> inline const std::reference_wrapper<const T> cref(const T &t) { return std::cref(t); }
}
using namespace std::placeholders;
namespace boost_shim {
template <class B> struct NotBind { B b; template <class... A> bool operator()(A &&...a) { return !b(std::forward<A>(a)...); } };
template <class L, class R> struct LessBind { L l; R r; template <class... A> bool operator()(A &&...a) { return l(a...) < r(a...); } };
template <class L, class V> struct EqualsValue { L l; V v; bool eq; template <class... A> bool operator()(A &&...a) { return (l(a...) == v) == eq; } };
}
// boost::bind expressions support operator! and relational operators; std::bind does not.
template <class B, class = typename std::enable_if<std::is_bind_expression<B>::value>::type>
boost_shim::NotBind<B> operator!(B b) { return boost_shim::NotBind<B>{b}; }
template <class L, class R, class = typename std::enable_if<std::is_bind_expression<L>::value && std::is_bind_expression<R>::value>::type>
boost_shim::LessBind<L, R> operator<(L l, R r) { return boost_shim::LessBind<L, R>{l, r}; }
// bind expression compared with a plain value (BamTemplate.hh:90-95)
template <class L, class V, class = typename std::enable_if<std::is_bind_expression<L>::value && std::is_arithmetic<V>::value>::type>
boost_shim::EqualsValue<L, V> operator!=(L l, V v) { return boost_shim::EqualsValue<L, V>{l, v, false}; }
template <class L, class V, class = typename std::enable_if<std::is_bind_expression<L>::value && std::is_arithmetic<V>::value>::type>
boost_shim::EqualsValue<L, V> operator==(L l, V v) { return boost_shim::EqualsValue<L, V>{l, v, true}; }
