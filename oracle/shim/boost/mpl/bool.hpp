#pragma once
namespace boost { namespace mpl { template <bool B> struct bool_ { static const bool value = B; }; } }
