#pragma once
#include <type_traits>
namespace boost { using std::is_unsigned; using std::is_signed; using std::is_integral; using std::is_same; using std::remove_const; }
