// Stand-in for boost::function_output_iterator: an output iterator whose assignment calls a function
// (TemplateBuilder.cpp:699,710 sums probabilities through std::unique_copy with it).
#pragma once
#include <iterator>
namespace boost {
template <class F> class function_output_iterator
{
    F f_;
    struct Proxy { F *f; template <class T> Proxy &operator=(const T &v) { (*f)(v); return *this; } };
public:
    typedef std::output_iterator_tag iterator_category;
    typedef void value_type; typedef void difference_type; typedef void pointer; typedef void reference;
    explicit function_output_iterator(const F &f = F()) : f_(f) {}
    Proxy operator*() { return Proxy{&f_}; }
    function_output_iterator &operator++() { return *this; }
    function_output_iterator &operator++(int) { return *this; }
};
template <class F> function_output_iterator<F> make_function_output_iterator(const F &f) { return function_output_iterator<F>(f); }
}
