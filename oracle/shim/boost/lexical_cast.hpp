#pragma once
// Stand-in for boost::lexical_cast: numbers to strings through std::to_string (no stream, no locale), anything else through a
// string stream.
#include <sstream>
#include <string>
#include <type_traits>
namespace boost {
namespace lexical_cast_shim {
template <class T, class S> struct Cast { static T run(const S &s) { std::stringstream ss; ss << s; T t; ss >> t; return t; } };
template <class S> struct Cast<std::string, S>
{
    template <class U = S> static typename std::enable_if<std::is_arithmetic<U>::value, std::string>::type run(const S &s) { return std::to_string(s); }
    template <class U = S> static typename std::enable_if<!std::is_arithmetic<U>::value, std::string>::type run(const S &s) { std::stringstream ss; ss << s; return ss.str(); }
};
}
template <class T, class S> T lexical_cast(const S &s) { return lexical_cast_shim::Cast<T, S>::run(s); }
}
