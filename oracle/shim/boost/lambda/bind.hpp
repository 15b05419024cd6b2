// Stand-in for boost::lambda placeholders (std::bind nested binds evaluate eagerly, which is what the callers need).
#pragma once
#include <functional>
namespace boost { namespace lambda { using std::bind; using namespace std::placeholders; } }
