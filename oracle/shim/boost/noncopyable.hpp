// Stand-in for boost::noncopyable (test infrastructure: lets the unmodified reference sources compile without Boost).
#pragma once
namespace boost { class noncopyable { protected: noncopyable() {} ~noncopyable() {} private: noncopyable(const noncopyable&); noncopyable& operator=(const noncopyable&); }; }
