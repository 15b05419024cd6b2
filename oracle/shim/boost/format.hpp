// Stand-in for boost::format: only builds message strings for asserts/exceptions; formatting is approximated
// by appending the arguments (messages are diagnostics, never results).
#pragma once
#include <string>
#include <sstream>
#include <ostream>
namespace boost {
class format {
    std::string fmt_; std::ostringstream args_;
public:
    format(const char *f) : fmt_(f) {}
    format(const std::string &f) : fmt_(f) {}
    format(const format &o) : fmt_(o.fmt_) { args_ << o.args_.str(); }
    template <class T> format &operator%(const T &v) { args_ << " [" << v << "]"; return *this; }
    std::string str() const { return fmt_ + args_.str(); }
};
inline std::ostream &operator<<(std::ostream &os, const format &f) { return os << f.str(); }
}
