// Stand-in for the cmake-generated config.h: the C library here provides all of these.
#pragma once
#define HAVE_FLOORF 1
#define HAVE_ROUND 1
#define HAVE_ROUNDF 1
#define HAVE_POWF 1
#define HAVE_ERF 1
#define HAVE_ERFF 1
#define HAVE_ERFC 1
#define HAVE_ERFCF 1
