#pragma once
// Stand-in for the little of boost::filesystem the headers on the gap realigner's include path mention: a path that holds a string.
#include <string>
#include <ostream>
namespace boost { namespace filesystem {
class path
{
    std::string s_;
public:
    path() {}
    path(const std::string &s) : s_(s) {}
    path(const char *s) : s_(s) {}
    const std::string &string() const { return s_; }
    const char *c_str() const { return s_.c_str(); }
    bool empty() const { return s_.empty(); }
    path operator/(const path &r) const { return path(s_ + "/" + r.s_); }
    path &operator/=(const path &r) { s_ += "/" + r.s_; return *this; }
    bool operator==(const path &r) const { return s_ == r.s_; }
    bool operator!=(const path &r) const { return s_ != r.s_; }
    bool operator<(const path &r) const { return s_ < r.s_; }
    path filename() const { const size_t p = s_.rfind('/'); return p == std::string::npos ? *this : path(s_.substr(p + 1)); }
    path parent_path() const { const size_t p = s_.rfind('/'); return p == std::string::npos ? path() : path(s_.substr(0, p)); }
};
inline std::ostream &operator<<(std::ostream &os, const path &p) { return os << p.string(); }
}}
