// Shadows common/FiniteCapacityVector.hh of the reference (a boost::ublas::bounded_array wrapper): same interface
// on a plain fixed array + size.
#ifndef iSAAC_COMMON_FINITE_CAPACITY_VECTOR_HH
#define iSAAC_COMMON_FINITE_CAPACITY_VECTOR_HH
#include <cstddef>
#include <algorithm>
#include <cassert>
namespace isaac { namespace common {
template <class T, std::size_t N> class FiniteCapacityVector {
    T data_[N]; std::size_t size_;
public:
    typedef T *iterator; typedef const T *const_iterator; typedef const T &const_reference; typedef T value_type;
    FiniteCapacityVector() : size_(0) {}
    FiniteCapacityVector(std::size_t s, const T &v) : size_(0) { resize(s, v); }
    iterator begin() { return data_; } iterator end() { return data_ + size_; }
    const_iterator begin() const { return data_; } const_iterator end() const { return data_ + size_; }
    bool empty() const { return 0 == size_; }
    std::size_t size() const { return size_; }
    static std::size_t capacity() { return N; }
    T &operator[](std::size_t i) { assert(i < size_); return data_[i]; }
    const T &operator[](std::size_t i) const { assert(i < size_); return data_[i]; }
    void resize(std::size_t s) { assert(s <= N); for (std::size_t i = size_; i < s; ++i) data_[i] = T(); size_ = s; }
    void resize(std::size_t s, const T &v) { assert(s <= N); for (std::size_t i = size_; i < s; ++i) data_[i] = v; size_ = s; }
    void clear() { size_ = 0; }
    void push_back(const T &x) { assert(size_ < N); data_[size_++] = x; }
    T pop_back() { assert(size_); return data_[--size_]; }
    T &front() { assert(size_); return data_[0]; } const T &front() const { assert(size_); return data_[0]; }
    T &back() { assert(size_); return data_[size_ - 1]; } const T &back() const { assert(size_); return data_[size_ - 1]; }
    void erase(iterator b, iterator e) { size_ = std::copy(e, end(), b) - begin(); }
};
} }
#endif
