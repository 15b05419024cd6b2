#pragma once
#define BOOST_MPL_ASSERT(x) static_assert(true, "")
