// Force-included (-include) ahead of every reference translation unit: standard headers the Boost ones used to
// drag in, plus the stand-ins.  Test infrastructure only.
#pragma once
#include <cassert>
#include <algorithm>
#include <limits>
#include <numeric>
#include <cmath>
#include <string>
#include <cstring>
#include <vector>
#include <iostream>
#include <iterator>
#include <iomanip>
#include <functional>
#include <stdint.h>
#include <boost/noncopyable.hpp>
#include <boost/format.hpp>
#include <boost/foreach.hpp>
#include <boost/bind.hpp>
#include <boost/integer_traits.hpp>
#include <boost/numeric/conversion/cast.hpp>
#include <boost/assign.hpp>
#include <boost/array.hpp>
#include <boost/lexical_cast.hpp>
#include <boost/ref.hpp>
#include <boost/filesystem.hpp>
#define BOOST_STATIC_ASSERT(x) static_assert(x, "")
#define BOOST_CURRENT_FUNCTION __func__
#include "common/Debug.hh"
#include "common/Exceptions.hh"
#include "alignment/SeedMetadata.hh"   // the reference's flowcell/Layout.hh brings this in
#include "common/FiniteCapacityVector.hh"
