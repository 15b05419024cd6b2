// Stand-in for boost::bind / ref / cref on top of <functional>.
#pragma once
#include <functional>
#include <type_traits>
#include <utility>
namespace boost {
using std::bind;
using std::ref;
// The reference takes the address of boost::cref<char> / boost::cref<unsigned char>, so it must be a real function template.
// It calls cref<unsigned char> on 'char' elements (FragmentMetadataTileStatsAdapter.hh:49-50): the argument is a converted
// temporary and a reference to it dangles once cref returns (with the real Boost too; there the stack slot happens to
// survive until the comparison).  This stand-in carries the value instead, which is what the code means.
template <class T> struct ValueRef { T value; operator const T &() const { return value; } };
template <class T> inline const ValueRef<T> cref(const T &t) { return ValueRef<T>{t}; }
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
// bind expression >= plain value (FragmentMetadataTileStatsAdapter.hh:49-50: quality >= 30u)
namespace boost_shim {
template <class L, class V> struct AtLeastValue { L l; V v; template <class... A> bool operator()(A &&...a) { return l(a...) >= v; } };
}
template <class L, class V, class = typename std::enable_if<std::is_bind_expression<L>::value && std::is_arithmetic<V>::value>::type>
boost_shim::AtLeastValue<L, V> operator>=(L l, V v) { return boost_shim::AtLeastValue<L, V>{l, v}; }
// boost::bind evaluates a nested "bind == value" with the outer call's arguments (TileStats.hh:118-121: plus(_2, bind(cref, _1) == 'n'));
// std::bind does that for the types it knows as bind expressions
namespace std {
template <class L, class V> struct is_bind_expression<boost_shim::EqualsValue<L, V> > : true_type {};
}
